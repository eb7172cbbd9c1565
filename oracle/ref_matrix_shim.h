/*
 * ref_matrix_shim.h -- stand-in for the parts of psi4's Matrix / Dimension / libqt that Matrix::power
 * (libmints/matrix.cc:2370-2424) touches, so that function can be compiled unmodified from /root/reference into
 * oracle/_ref/libref_matrix.so (recipe: oracle/ref_build.py).  TEST INFRASTRUCTURE ONLY.  The fitting metric
 * J^-1/2 = power(-0.5, condition) decides which near-null fitting directions are dropped (|lambda| < cond * |lambda|max),
 * which is why the host driver's matrix_power is checked against this and not only against numpy.
 */
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace psi {

#define PSIEXCEPTION(msg) std::runtime_error(msg)

class Dimension {  // libmints/dimension.h:93-
    std::vector<int> v_;

   public:
    Dimension(int n, const std::string& = "") : v_(n, 0) {}
    int& operator[](int i) { return v_[i]; }
    const int& operator[](int i) const { return v_[i]; }
};

namespace linalg {
namespace detail {
inline double** matrix(int nrow, int ncol) {  // libmints/matrix.cc: contiguous block + row pointers
    double** m = (double**)std::malloc(sizeof(double*) * nrow);
    m[0] = (double*)std::calloc((size_t)nrow * ncol, sizeof(double));
    for (int i = 1; i < nrow; i++) m[i] = m[0] + (size_t)i * ncol;
    return m;
}
inline void free(double** m) {
    std::free(m[0]);
    std::free(m);
}
}  // namespace detail
}  // namespace linalg

// libqt's wrappers are the reference's own (lapack_intfc.cc C_DSYEV, blas_intfc.cc C_DSCAL, blas_intfc23.cc C_DGEMM,
// compiled as they are and linked in); prototypes from libqt/qt.h.
int C_DSYEV(char jobz, char uplo, int n, double* a, int lda, double* w, double* work, int lwork);
void C_DSCAL(size_t len, double alpha, double* vec, int inc);
void C_DGEMM(char transa, char transb, int m, int n, int k, double alpha, double* a, int lda, double* b, int ldb,
             double beta, double* c, int ldc);

class Matrix {
   public:
    int symmetry_ = 0, nirrep_ = 1;
    int rowspi_[1] = {0};
    double** matrix_[1] = {nullptr};
    Dimension power(double alpha, double cutoff);
};

}  // namespace psi
