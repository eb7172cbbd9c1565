// ref_matrix_entry.inl -- C entry point over the sliced Matrix::power.
#include <dlfcn.h>
typedef void (*dsyev_fn)(char*, char*, int*, double*, int*, double*, double*, int*, int*);
typedef void (*dscal_fn)(int*, double*, double*, int*);
typedef void (*dgemm_fn)(char*, char*, int*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
static dsyev_fn REFM_DSYEV = nullptr;
static dscal_fn REFM_DSCAL = nullptr;
static dgemm_fn REFM_DGEMM = nullptr;
extern "C" {  // the Fortran symbols the reference's libqt wrappers call (FC_SYMBOL == 2), forwarded to OpenBLAS
void dsyev_(char* jz, char* ul, int* n, double* a, int* lda, double* w, double* work, int* lwork, int* info) {
    REFM_DSYEV(jz, ul, n, a, lda, w, work, lwork, info);
}
void dscal_(int* n, double* al, double* x, int* inc) { REFM_DSCAL(n, al, x, inc); }
void dgemm_(char* ta, char* tb, int* m, int* n, int* k, double* al, double* a, int* lda, double* b, int* ldb, double* be,
            double* c, int* ldc) {
    REFM_DGEMM(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}
}

extern "C" {
int refm_init(const char* blas_path) {
    void* h = dlopen(blas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return 1;
    const char* pre[] = {"scipy_", "", nullptr};
    for (int i = 0; pre[i]; i++) {
        std::string p = pre[i];
        REFM_DSYEV = (dsyev_fn)dlsym(h, (p + "dsyev_").c_str());
        REFM_DSCAL = (dscal_fn)dlsym(h, (p + "dscal_").c_str());
        REFM_DGEMM = (dgemm_fn)dlsym(h, (p + "dgemm_").c_str());
        if (REFM_DSYEV && REFM_DSCAL && REFM_DGEMM) return 0;
    }
    return 2;
}

// A (n x n, row-major, symmetric) <- A^alpha in place; returns the number of eigenvalues kept, < 0 on error.
int refm_power(double* A, int n, double alpha, double cutoff) {
    if (!REFM_DSYEV) return -1;
    try {
        std::vector<double*> rows(n);
        for (int i = 0; i < n; i++) rows[i] = A + (size_t)i * n;
        psi::Matrix m;
        m.rowspi_[0] = n;
        m.matrix_[0] = rows.data();
        psi::Dimension rem = m.power(alpha, cutoff);
        return rem[0];
    } catch (const std::exception&) {
        return -2;
    }
}
}
