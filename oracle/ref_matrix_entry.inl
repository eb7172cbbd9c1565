// ref_matrix_entry.inl -- C entry point over the sliced Matrix::power.
#include <dlfcn.h>
namespace psi {
dsyev_fn REFM_DSYEV = nullptr;
dscal_fn REFM_DSCAL = nullptr;
dgemm_fn REFM_DGEMM = nullptr;
}  // namespace psi

extern "C" {
int refm_init(const char* blas_path) {
    void* h = dlopen(blas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return 1;
    const char* pre[] = {"scipy_", "", nullptr};
    for (int i = 0; pre[i]; i++) {
        std::string p = pre[i];
        psi::REFM_DSYEV = (psi::dsyev_fn)dlsym(h, (p + "dsyev_").c_str());
        psi::REFM_DSCAL = (psi::dscal_fn)dlsym(h, (p + "dscal_").c_str());
        psi::REFM_DGEMM = (psi::dgemm_fn)dlsym(h, (p + "dgemm_").c_str());
        if (psi::REFM_DSYEV && psi::REFM_DSCAL && psi::REFM_DGEMM) return 0;
    }
    return 2;
}

// A (n x n, row-major, symmetric) <- A^alpha in place; returns the number of eigenvalues kept, < 0 on error.
int refm_power(double* A, int n, double alpha, double cutoff) {
    if (!psi::REFM_DSYEV) return -1;
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
