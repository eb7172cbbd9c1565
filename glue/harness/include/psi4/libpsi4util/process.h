// HARNESS STAND-IN for psi4/libpsi4util/process.h: Process::environment.options / get_n_threads().
#pragma once
#include "psi4/liboptions/liboptions.h"
namespace psi {
class Process {
   public:
    class Environment {
       public:
        Options options;
        int get_n_threads() const { return 1; }
    };
    static Environment environment;
};
}  // namespace psi
