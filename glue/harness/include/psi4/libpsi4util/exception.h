// HARNESS STAND-IN for psi4/libpsi4util/exception.h:48.
#pragma once
#include <stdexcept>
#include <string>
namespace psi {
class PsiException : public std::runtime_error {
   public:
    PsiException(const std::string& msg, const char* /*file*/, int /*line*/) : std::runtime_error(msg) {}
};
#define PSIEXCEPTION(message) psi::PsiException(message, __FILE__, __LINE__)
}  // namespace psi
