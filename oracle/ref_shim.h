/*
 * ref_shim.h -- the little of psi4's runtime that the reference's own DF-JK loops touch, so that the function bodies
 * of lib3index/dfhelper.cc can be compiled UNMODIFIED, straight from /root/reference, into oracle/_ref/libref_dfjk.so
 * (recipe: oracle/ref_build.py, which slices the function definitions out of the reference file at build time and pipes
 * them to the compiler; no reference source is stored in this repository).
 *
 * TEST INFRASTRUCTURE ONLY: the library built from this is the checker of the checker -- tests compare the C
 * restatement (dfjk_oracle.c) with it bit for bit; bench.py may time it as the "reference" CPU baseline.
 *
 * Stand-ins, each with the reference declaration it replaces:
 *   Matrix / SharedMatrix   libmints/matrix.h:69- (only pointer()[0] and colspi()[0] are used by the sliced code)
 *   outfile->Printf         libpsi4util/PsiOutStream.h
 *   timer_on / timer_off    libqt/qt.h:79-80
 *   PSIEXCEPTION            libpsi4util/exception.h:48
 * (libqt's C_DGEMM / C_DGEMV / C_DCOPY are the reference's own, compiled from libqt/blas_intfc23.cc and blas_intfc.cc.)
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

// The BLAS wrappers are NOT stand-ins: libqt/blas_intfc23.cc (C_DGEMM :324-328, C_DGEMV :424-433) and
// libqt/blas_intfc.cc (C_DCOPY) compile as they are (g++ -DFC_SYMBOL=2 on the reference files, see ref_build.py) and are
// linked in; these are their prototypes (libqt/qt.h:100-, :160-).  The Fortran symbols they call (dgemm_, dgemv_, dcopy_)
// are forwarded to the LP64 OpenBLAS bundled with scipy by ref_entry.inl.
void C_DGEMM(char transa, char transb, int m, int n, int k, double alpha, double* a, int lda, double* b, int ldb,
             double beta, double* c, int ldc);
void C_DGEMV(char trans, int m, int n, double alpha, double* a, int lda, double* x, int incx, double beta, double* y,
             int incy);
void C_DCOPY(size_t length, double* x, int inc_x, double* y, int inc_y);

}  // namespace psi
