/*
 * ref_shim.h -- the little of psi4's runtime that the reference's own DF-JK loops touch, so that the function bodies
 * of lib3index/dfhelper.cc can be compiled UNMODIFIED, straight from /root/reference, into oracle/_ref/libref_dfjk.so
 * (recipe: oracle/ref_build.py, which slices the function definitions out of the reference file at build time; no
 * reference source is stored in this repository).
 *
 * TEST INFRASTRUCTURE ONLY: the library built from this is the checker of the checker -- tests compare the C
 * restatement (dfjk_oracle.c) with it bit for bit; bench.py may time it as the "reference" CPU baseline.
 *
 * Stand-ins, each with the reference declaration it replaces:
 *   Matrix / SharedMatrix   libmints/matrix.h:69- (only pointer()[0] and colspi()[0] are used by the sliced code)
 *   outfile->Printf         libpsi4util/PsiOutStream.h
 *   timer_on / timer_off    libqt/qt.h:79-80
 *   PSIEXCEPTION            libpsi4util/exception.h:48
 *   C_DGEMM/C_DGEMV/C_DCOPY libqt/blas_intfc23.cc:324-328, :424-433, libqt/blas_intfc.cc (row-major over Fortran BLAS)
 */
#pragma once
#include <algorithm>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace psi {

struct Dimension1 {
    int n;
    int operator[](int) const { return n; }
};

class Matrix {
    double* data_;
    double* row0_[1];
    Dimension1 rows_, cols_;

   public:
    Matrix(double* data, int rows, int cols) : data_(data), rows_{rows}, cols_{cols} { row0_[0] = data_; }
    double** pointer(int = 0) { return row0_; }  // pointer()[0] = contiguous row-major block (matrix.cc:3586-3593)
    const Dimension1& rowspi() const { return rows_; }
    const Dimension1& colspi() const { return cols_; }
};
using SharedMatrix = std::shared_ptr<Matrix>;

struct PsiOutStream {
    void Printf(const char*, ...) {}
};
static std::shared_ptr<PsiOutStream> outfile = std::make_shared<PsiOutStream>();
inline void timer_on(const std::string&) {}
inline void timer_off(const std::string&) {}
#define PSIEXCEPTION(msg) std::runtime_error(msg)

typedef void (*dgemm_fn)(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                         const int*, const double*, const int*, const double*, double*, const int*);
typedef void (*dgemv_fn)(const char*, const int*, const int*, const double*, const double*, const int*, const double*,
                         const int*, const double*, double*, const int*);
typedef void (*dcopy_fn)(const int*, const double*, const int*, double*, const int*);
extern dgemm_fn REF_DGEMM;
extern dgemv_fn REF_DGEMV;
extern dcopy_fn REF_DCOPY;

inline void C_DGEMM(char transa, char transb, int m, int n, int k, double alpha, double* a, int lda, double* b, int ldb,
                    double beta, double* c, int ldc) {
    if (m == 0 || n == 0 || k == 0) return;
    REF_DGEMM(&transb, &transa, &n, &m, &k, &alpha, b, &ldb, a, &lda, &beta, c, &ldc);
}
inline void C_DGEMV(char trans, int m, int n, double alpha, double* a, int lda, double* x, int incx, double beta,
                    double* y, int incy) {
    if (m == 0 || n == 0) return;
    trans = (trans == 'N' || trans == 'n') ? 'T' : 'N';
    REF_DGEMV(&trans, &n, &m, &alpha, a, &lda, x, &incx, &beta, y, &incy);
}
inline void C_DCOPY(size_t length, double* x, int incx, double* y, int incy) {
    int n = (int)length;
    REF_DCOPY(&n, x, &incx, y, &incy);
}

}  // namespace psi
