// Declaration-only stand-in used ONLY by tests/test_glue_compiles.py: psi4's libmints/basisset.h holds a
// std::vector<libint2::Shell> member; the glue never touches it.
#pragma once
namespace libint2 {
struct Shell {};
}  // namespace libint2
