// HARNESS STAND-IN for psi4/libmints/basisset.h (the real one pulls in <libint2/shell.h>): a basis set is its size.
#pragma once
namespace psi {
class BasisSet {
    int nbf_;

   public:
    explicit BasisSet(int nbf) : nbf_(nbf) {}
    int nbf() const { return nbf_; }
};
}  // namespace psi
