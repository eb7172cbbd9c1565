// ref_entry.inl -- C entry points over the sliced reference functions; same signatures as the restatement's
// oracle_build_JK / oracle_contract_metric_AO_core_symm (dfjk_oracle.c) so one Python wrapper drives both.
#include <dlfcn.h>

// Fortran BLAS as the reference's libqt wrappers call it (FC_SYMBOL == 2: lower case + underscore), forwarded to the
// OpenBLAS resolved by ref_init.  The other ~25 level-2/3 routines blas_intfc23.cc declares stay unresolved: nothing on
// this path calls them (the library is loaded with lazy binding).
typedef void (*dgemm_fn)(char*, char*, int*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
typedef void (*dgemv_fn)(char*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
typedef void (*dcopy_fn)(int*, double*, int*, double*, int*);
static dgemm_fn REF_DGEMM = nullptr;
static dgemv_fn REF_DGEMV = nullptr;
static dcopy_fn REF_DCOPY = nullptr;
extern "C" {
void dgemm_(char* ta, char* tb, int* m, int* n, int* k, double* al, double* a, int* lda, double* b, int* ldb, double* be,
            double* c, int* ldc) {
    REF_DGEMM(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}
void dgemv_(char* t, int* m, int* n, double* al, double* a, int* lda, double* x, int* incx, double* be, double* y,
            int* incy) {
    REF_DGEMV(t, m, n, al, a, lda, x, incx, be, y, incy);
}
void dcopy_(int* n, double* x, int* incx, double* y, int* incy) { REF_DCOPY(n, x, incx, y, incy); }
}

namespace {
// the packed tensors are owned by the caller: hand them to the unique_ptr members for the call, take them back after
struct Borrow {
    std::unique_ptr<double[]>& slot;
    Borrow(std::unique_ptr<double[]>& s, const double* p) : slot(s) { slot.reset(const_cast<double*>(p)); }
    ~Borrow() { slot.release(); }
};

void fill_tables(psi::DFHelper& d, size_t nbf, size_t naux, int nthreads, const size_t* fun_index, const size_t* small_skips,
                 const size_t* big_skips, const size_t* symm_small_skips, const size_t* symm_ignored_columns) {
    d.nbf_ = nbf;
    d.naux_ = naux;
    d.nthreads_ = (size_t)nthreads;
    d.schwarz_fun_index_.assign(fun_index, fun_index + nbf * nbf);
    d.small_skips_.assign(small_skips, small_skips + nbf + 1);
    d.big_skips_.assign(big_skips, big_skips + nbf + 1);
    d.symm_small_skips_.assign(symm_small_skips, symm_small_skips + nbf);
    d.symm_ignored_columns_.assign(symm_ignored_columns, symm_ignored_columns + nbf);
    d.symm_big_skips_.assign(nbf + 1, 0);
    for (size_t i = 1; i <= nbf; i++) d.symm_big_skips_[i] = d.symm_big_skips_[i - 1] + d.symm_small_skips_[i - 1] * naux;
}
}  // namespace

extern "C" {

int ref_init(const char* blas_path) {
    void* h = dlopen(blas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return 1;
    const char* pre[] = {"scipy_", "", nullptr};
    for (int i = 0; pre[i]; i++) {
        std::string p = pre[i];
        REF_DGEMM = (dgemm_fn)dlsym(h, (p + "dgemm_").c_str());
        REF_DGEMV = (dgemv_fn)dlsym(h, (p + "dgemv_").c_str());
        REF_DCOPY = (dcopy_fn)dlsym(h, (p + "dcopy_").c_str());
        if (REF_DGEMM && REF_DGEMV && REF_DCOPY) return 0;
    }
    return 2;
}

// DFHelper::build_JK (dfhelper.cc:3015) as MemDFJK::compute_JK calls it (MemDFJK.cc:97-111: outputs zeroed first,
// wK hermitivitized afterwards when lr_symmetric).  q_block > 0 makes Qshell_blocks_for_JK_build cut the auxiliary
// index into "shells" of q_block functions and sets memory_ so that exactly one of them fits per block.
int ref_build_JK(size_t nbf, size_t naux, int nthreads, const size_t* fun_index, const size_t* small_skips,
                 const size_t* big_skips, const size_t* symm_small_skips, const size_t* symm_ignored_columns,
                 const double* Ppq, const double* m1Ppq, const double* wPpq, int nmat, double* const* Cleft,
                 double* const* Cright, const int* nocc, double* const* D, double* const* J, double* const* K,
                 double* const* wK, int do_J, int do_K, int do_wK, int lr_symmetric, size_t q_block) {
    if (!REF_DGEMM) return 1;
    try {
        psi::DFHelper d;
        fill_tables(d, nbf, naux, nthreads, fun_index, small_skips, big_skips, symm_small_skips, symm_ignored_columns);
        d.do_wK_ = do_wK != 0;
        size_t max_nocc = 0;
        for (int i = 0; i < nmat; i++) max_nocc = std::max(max_nocc, (size_t)nocc[i]);
        if (q_block == 0 || q_block > naux) q_block = naux;
        d.Qshell_aggs_.clear();
        for (size_t q = 0; q < naux; q += q_block) d.Qshell_aggs_.push_back(q);
        d.Qshell_aggs_.push_back(naux);
        d.Qshells_ = d.Qshell_aggs_.size() - 1;
        {
            // exactly the constraint of Qshell_blocks_for_JK_build (:814-869) at tmpbs = q_block: one "shell" fits, two
            // do not.  (compute_wK evaluates the rule with lr_symmetric = false, :3381; with wK tasked memory_ is the
            // larger of the two needs, so the cheaper pass may merge a short tail shell into its last block.)
            size_t T1 = nbf * max_nocc, T2 = lr_symmetric ? nbf * nbf : nbf * max_nocc;
            size_t T3 = std::max(d.nthreads_ * nbf * nbf, d.nthreads_ * nbf * max_nocc);
            d.memory_ = d.big_skips_[nbf] + T1 * q_block + T3 + (lr_symmetric ? T2 : T2 * q_block);
            if (do_wK) d.memory_ = std::max(d.memory_, d.big_skips_[nbf] + T1 * q_block + T3 + T1 * q_block);
        }
        Borrow b0(d.Ppq_, Ppq), b1(d.m1Ppq_, m1Ppq), b2(d.wPpq_, wPpq);
        std::vector<psi::SharedMatrix> Cl, Cr, Dm, Jm, Km, wKm;
        for (int i = 0; i < nmat; i++) {
            Cl.push_back(std::make_shared<psi::Matrix>(Cleft[i], (int)nbf, nocc[i]));
            Cr.push_back(std::make_shared<psi::Matrix>(lr_symmetric ? Cleft[i] : Cright[i], (int)nbf, nocc[i]));
            Dm.push_back(std::make_shared<psi::Matrix>(D ? D[i] : nullptr, (int)nbf, (int)nbf));
            if (do_J) {
                std::memset(J[i], 0, sizeof(double) * nbf * nbf);  // JK::zero(), MemDFJK.cc:100
                Jm.push_back(std::make_shared<psi::Matrix>(J[i], (int)nbf, (int)nbf));
            }
            if (do_K) {
                std::memset(K[i], 0, sizeof(double) * nbf * nbf);
                Km.push_back(std::make_shared<psi::Matrix>(K[i], (int)nbf, (int)nbf));
            }
            if (do_wK) {
                std::memset(wK[i], 0, sizeof(double) * nbf * nbf);
                wKm.push_back(std::make_shared<psi::Matrix>(wK[i], (int)nbf, (int)nbf));
            }
        }
        d.build_JK(Cl, Cr, Dm, Jm, Km, wKm, max_nocc, do_J != 0, do_K != 0, do_wK != 0, lr_symmetric != 0);
        if (lr_symmetric && do_wK)  // MemDFJK.cc:104-110, Matrix::hermitivitize: (A + A^T)/2
            for (int i = 0; i < nmat; i++)
                for (size_t r = 0; r < nbf; r++)
                    for (size_t c = 0; c < r; c++) {
                        double v = 0.5 * (wK[i][r * nbf + c] + wK[i][c * nbf + r]);
                        wK[i][r * nbf + c] = wK[i][c * nbf + r] = v;
                    }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_build_JK: %s\n", e.what());
        return 3;
    }
    return 0;
}

// Table-building half of DFHelper::prepare_sparsity (dfhelper.cc:370-420) on given Schwarz maxima; outputs sized like
// the members: fun_index nbf*nbf, small_skips / big_skips / symm_big_skips nbf+1, the others nbf, shell_mask pshells^2.
int ref_prepare_sparsity_tables(size_t nbf, size_t naux, size_t pshells, double cutoff, const double* shell_max_vals,
                                const double* fun_max_vals, double max_val, size_t* fun_index, size_t* small_skips,
                                size_t* big_skips, size_t* symm_small_skips, size_t* symm_ignored_columns,
                                size_t* symm_big_skips, size_t* shell_mask) {
    try {
        psi::DFHelper d;
        d.nbf_ = nbf;
        d.naux_ = naux;
        d.pshells_ = pshells;
        d.cutoff_ = cutoff;
        // the resizes of the function's first half (:306-314)
        d.schwarz_shell_mask_.resize(pshells * pshells);
        d.schwarz_fun_index_.resize(nbf * nbf);
        d.symm_ignored_columns_.resize(nbf);
        d.symm_big_skips_.resize(nbf + 1);
        d.symm_small_skips_.resize(nbf);
        d.small_skips_.resize(nbf + 1);
        d.big_skips_.resize(nbf + 1);
        std::vector<double> sm(shell_max_vals, shell_max_vals + pshells * pshells), fm(fun_max_vals, fun_max_vals + nbf * nbf);
        d.prepare_sparsity_tables(sm, fm, max_val);
        std::copy(d.schwarz_fun_index_.begin(), d.schwarz_fun_index_.end(), fun_index);
        std::copy(d.small_skips_.begin(), d.small_skips_.end(), small_skips);
        std::copy(d.big_skips_.begin(), d.big_skips_.end(), big_skips);
        std::copy(d.symm_small_skips_.begin(), d.symm_small_skips_.end(), symm_small_skips);
        std::copy(d.symm_ignored_columns_.begin(), d.symm_ignored_columns_.end(), symm_ignored_columns);
        std::copy(d.symm_big_skips_.begin(), d.symm_big_skips_.end(), symm_big_skips);
        std::copy(d.schwarz_shell_mask_.begin(), d.schwarz_shell_mask_.end(), shell_mask);
        return d.sparsity_prepared_ ? 0 : 4;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_prepare_sparsity_tables: %s\n", e.what());
        return 3;
    }
}

// DFHelper::contract_metric_AO_core_symm (dfhelper.cc:1653-1678)
int ref_contract_metric_AO_core_symm(size_t nbf, size_t naux, int nthreads, const size_t* fun_index,
                                     const size_t* small_skips, const size_t* big_skips, const size_t* symm_small_skips,
                                     const size_t* symm_ignored_columns, const size_t* /*symm_big_skips*/,
                                     const double* Qpq, double* Ppq, const double* metp, size_t begin, size_t end) {
    if (!REF_DGEMM) return 1;
    try {
        psi::DFHelper d;
        fill_tables(d, nbf, naux, nthreads, fun_index, small_skips, big_skips, symm_small_skips, symm_ignored_columns);
        d.contract_metric_AO_core_symm(const_cast<double*>(Qpq), Ppq, const_cast<double*>(metp), begin, end);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_contract_metric_AO_core_symm: %s\n", e.what());
        return 3;
    }
    return 0;
}
}
