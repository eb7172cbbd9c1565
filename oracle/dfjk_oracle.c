/*
 * dfjk_oracle.c -- CPU restatement of psi4's in-core MemDFJK / DFHelper J/K build.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under psi4_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, as the checker and as the timed CPU baseline.
 *
 * PARITY PINNING: see the "Oracle" section of DESIGN.md for the current status.  The
 * reference (psi4) cannot be built or imported in this image (no Libint2 / LibXC /
 * gau2grid), and the reference's tests hold no element-wise J/K
 * golden arrays for this path -- only SCF energies, which need real integrals.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/psi4/src/psi4/).  Loop order, BLAS calls (through the same
 * row-major-over-Fortran convention as libqt) and OpenMP structure are kept so that
 * the timing of this code is a fair stand-in for the reference's "JK: JK" region.
 *
 * BLAS: resolved at run time by dlopen of the LP64 OpenBLAS bundled with scipy
 * (symbols scipy_dgemm_, scipy_dgemv_, scipy_dcopy_), passed in by oracle_init().
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*dgemm_fn)(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                         const int*, const double*, const int*, const double*, double*, const int*);
typedef void (*dgemv_fn)(const char*, const int*, const int*, const double*, const double*, const int*, const double*,
                         const int*, const double*, double*, const int*);
typedef void (*dcopy_fn)(const int*, const double*, const int*, double*, const int*);
typedef void (*setthr_fn)(int);
typedef int (*getthr_fn)(void);
typedef char* (*getcfg_fn)(void);

static dgemm_fn F_DGEMM = NULL;
static dgemv_fn F_DGEMV = NULL;
static dcopy_fn F_DCOPY = NULL;
static setthr_fn blas_set_threads = NULL;
static getthr_fn blas_get_threads = NULL;
static getcfg_fn blas_get_config = NULL;
static char blas_cfg[256] = "unresolved";

/* Resolve BLAS.  Returns 0 on success. */
int oracle_init(const char* blas_path) {
    void* h = dlopen(blas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        fprintf(stderr, "oracle_init: dlopen(%s) failed: %s\n", blas_path, dlerror());
        return 1;
    }
    const char* pre[] = {"scipy_", "", NULL};
    for (int i = 0; pre[i]; i++) {
        char nm[64];
        snprintf(nm, sizeof nm, "%sdgemm_", pre[i]);
        F_DGEMM = (dgemm_fn)dlsym(h, nm);
        snprintf(nm, sizeof nm, "%sdgemv_", pre[i]);
        F_DGEMV = (dgemv_fn)dlsym(h, nm);
        snprintf(nm, sizeof nm, "%sdcopy_", pre[i]);
        F_DCOPY = (dcopy_fn)dlsym(h, nm);
        snprintf(nm, sizeof nm, "%sopenblas_set_num_threads", pre[i]);
        blas_set_threads = (setthr_fn)dlsym(h, nm);
        snprintf(nm, sizeof nm, "%sopenblas_get_num_threads", pre[i]);
        blas_get_threads = (getthr_fn)dlsym(h, nm);
        snprintf(nm, sizeof nm, "%sopenblas_get_config", pre[i]);
        blas_get_config = (getcfg_fn)dlsym(h, nm);
        if (F_DGEMM && F_DGEMV && F_DCOPY) break;
    }
    if (!(F_DGEMM && F_DGEMV && F_DCOPY)) return 2;
    if (blas_get_config) snprintf(blas_cfg, sizeof blas_cfg, "%s", blas_get_config());
    return 0;
}
const char* oracle_blas_config(void) { return blas_cfg; }
void oracle_set_blas_threads(int n) {
    if (blas_set_threads) blas_set_threads(n);
}
int oracle_get_blas_threads(void) { return blas_get_threads ? blas_get_threads() : 1; }
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* libqt/blas_intfc23.cc:324-328 -- row-major DGEMM through column-major Fortran BLAS. */
static void C_DGEMM(char transa, char transb, int m, int n, int k, double alpha, const double* a, int lda,
                    const double* b, int ldb, double beta, double* c, int ldc) {
    if (m == 0 || n == 0 || k == 0) return;
    F_DGEMM(&transb, &transa, &n, &m, &k, &alpha, b, &ldb, a, &lda, &beta, c, &ldc);
}
/* libqt/blas_intfc23.cc:424-433 */
static void C_DGEMV(char trans, int m, int n, double alpha, const double* a, int lda, const double* x, int incx,
                    double beta, double* y, int incy) {
    if (m == 0 || n == 0) return;
    trans = (trans == 'N' || trans == 'n') ? 'T' : 'N';
    F_DGEMV(&trans, &n, &m, &alpha, a, &lda, x, &incx, &beta, y, &incy);
}
static void C_DCOPY(size_t n, const double* x, int incx, double* y, int incy) {
    int nn = (int)n;
    F_DCOPY(&nn, x, &incx, y, &incy);
}

/* ------------------------------------------------------------------------------------------
 * Sparsity tables.  lib3index/dfhelper.cc:299-420 (prepare_sparsity), second half: given the
 * per-function-pair Schwarz maxima fun_max_vals[nbf*nbf] and the cutoff, build the mask and
 * the packed-index tables.  tolerance = cutoff^2 / max_val (:371); pair kept iff
 * fun_max_vals >= tolerance (:380).
 * ------------------------------------------------------------------------------------------ */
void oracle_schwarz_mask(size_t nbf, const double* fun_max_vals, double cutoff, unsigned char* keep) {
    double max_val = 0.0;
    for (size_t i = 0; i < nbf * nbf; i++)
        if (fun_max_vals[i] > max_val) max_val = fun_max_vals[i];
    double tolerance = cutoff * cutoff / max_val;
    for (size_t i = 0; i < nbf * nbf; i++) keep[i] = (fun_max_vals[i] >= tolerance);
}

/* dfhelper.cc:377-416 given keep(m,n). All tables size_t as in dfhelper.h:440-452. */
void oracle_prepare_sparsity(size_t nbf, size_t naux, const unsigned char* keep, size_t* schwarz_fun_index,
                             size_t* small_skips /*nbf+1*/, size_t* big_skips /*nbf+1*/,
                             size_t* symm_small_skips /*nbf*/, size_t* symm_ignored_columns /*nbf*/,
                             size_t* symm_big_skips /*nbf+1*/) {
    for (size_t i = 0, count = 0; i < nbf; i++) {
        count = 0;
        for (size_t j = 0; j < nbf; j++) {
            if (keep[i * nbf + j]) {
                count++;
                schwarz_fun_index[i * nbf + j] = count;
            } else
                schwarz_fun_index[i * nbf + j] = 0;
        }
        small_skips[i] = count;
    }
    big_skips[0] = 0;
    size_t coltots = 0;
    for (size_t j = 0; j < nbf; j++) {
        size_t cols = small_skips[j];
        size_t size = cols * naux;
        coltots += cols;
        big_skips[j + 1] = size + big_skips[j];
    }
    small_skips[nbf] = coltots;
    for (size_t i = 0; i < nbf; i++) {
        size_t size = 0, skip = 0;
        for (size_t j = 0; j < nbf; j++) {
            if (schwarz_fun_index[i * nbf + j]) {
                if (j >= i)
                    size++;
                else
                    skip++;
            }
        }
        symm_small_skips[i] = size;
        symm_ignored_columns[i] = skip;
    }
    symm_big_skips[0] = 0;
    for (size_t i = 1; i < nbf + 1; i++) symm_big_skips[i] = symm_big_skips[i - 1] + symm_small_skips[i - 1] * naux;
}

/* Pack a dense (naux, nbf, nbf) tensor into pQq order:
 * B(Q,m,n) -> Ppq[ big_skips[m] + Q*sp(m) + fun_index[m*nbf+n] - 1 ]   (dfhelper.cc:1274-1276, :1671-1672) */
void oracle_pack_pQq(size_t nbf, size_t naux, const double* dense_Qmn, const size_t* fun_index,
                     const size_t* small_skips, const size_t* big_skips, double* Ppq) {
#pragma omp parallel for schedule(static)
    for (size_t m = 0; m < nbf; m++) {
        size_t sp = small_skips[m];
        for (size_t Q = 0; Q < naux; Q++) {
            for (size_t n = 0; n < nbf; n++) {
                size_t f = fun_index[m * nbf + n];
                if (f) Ppq[big_skips[m] + Q * sp + f - 1] = dense_Qmn[(Q * nbf + m) * nbf + n];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Context mirroring the DFHelper members the JK build reads.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    size_t nbf, naux;
    int nthreads;
    const size_t* schwarz_fun_index;
    const size_t* small_skips;
    const size_t* big_skips;
    const size_t* symm_small_skips;
    const size_t* symm_ignored_columns;
    /* Q extent actually present in the host tensors: rows [0, naux_store) of every m block.
     * naux_store == naux for the full tensor; a Q-slice (bench cpu_baseline sample) stores
     * fewer rows with big_skips built for naux_store. */
} oracle_ctx;

static void fill(double* b, size_t count, double value, int nthreads) {
    /* dfhelper.cc:3224-3229 */
#pragma omp parallel for simd num_threads(nthreads) schedule(static)
    for (size_t i = 0; i < count; i++) b[i] = value;
}

/* dfhelper.cc:3162-3223 */
static void compute_J_symm(const oracle_ctx* c, int nmat, double* const* D, double* const* J, const double* Mp,
                           double* T1p, double* T2p, double** D_buffers, size_t bcount, size_t block_size) {
    const size_t nbf_ = c->nbf, naux_ = c->naux;
    const int nthreads_ = c->nthreads;
    for (int i = 0; i < nmat; i++) {
        const double* Dp = D[i];
        double* Jp = J[i];
        fill(T1p, nthreads_ * naux_, 0.0, nthreads_);
#pragma omp parallel for schedule(guided) num_threads(nthreads_)
        for (size_t k = 0; k < nbf_; k++) {
            size_t si = c->small_skips[k];
            size_t mi = c->symm_small_skips[k];
            size_t skip = c->symm_ignored_columns[k];
            size_t jump = c->big_skips[k] + bcount * si;
            int rank = 0;
#ifdef _OPENMP
            rank = omp_get_thread_num();
#endif
            for (size_t m = k, sp_count = (size_t)-1; m < nbf_; m++) {
                if (c->schwarz_fun_index[k * nbf_ + m]) {
                    sp_count++;
                    D_buffers[rank][sp_count] = (m == k ? Dp[nbf_ * k + m] : 2 * Dp[nbf_ * k + m]);
                }
            }
            C_DGEMV('N', (int)block_size, (int)mi, 1.0, &Mp[jump + skip], (int)si, &D_buffers[rank][0], 1, 1.0,
                    &T1p[rank * naux_], 1);
        }
        for (size_t k = 1; k < (size_t)nthreads_; k++)
            for (size_t l = 0; l < naux_; l++) T1p[l] += T1p[k * naux_ + l];
#pragma omp parallel for schedule(guided) num_threads(nthreads_)
        for (size_t k = 0; k < nbf_; k++) {
            size_t si = c->small_skips[k];
            size_t mi = c->symm_small_skips[k];
            size_t skip = c->symm_ignored_columns[k];
            size_t jump = c->big_skips[k] + bcount * si;
            C_DGEMV('T', (int)block_size, (int)mi, 1.0, &Mp[jump + skip], (int)si, T1p, 1, 0.0, &T2p[k * nbf_], 1);
        }
        for (size_t k = 0; k < nbf_; k++) {
            for (size_t m = k + 1, count = 0; m < nbf_; m++) {
                if (c->schwarz_fun_index[k * nbf_ + m]) {
                    count++;
                    Jp[k * nbf_ + m] += T2p[k * nbf_ + count];
                    Jp[m * nbf_ + k] += T2p[k * nbf_ + count];
                }
            }
        }
        for (size_t k = 0; k < nbf_; k++) Jp[k * nbf_ + k] += T2p[k * nbf_];
    }
}

/* dfhelper.cc:3230-3285 */
static void compute_J(const oracle_ctx* c, int nmat, double* const* D, double* const* J, const double* Mp, double* T1p,
                      double* T2p, double** D_buffers, size_t bcount, size_t block_size) {
    const size_t nbf_ = c->nbf, naux_ = c->naux;
    const int nthreads_ = c->nthreads;
    for (int i = 0; i < nmat; i++) {
        const double* Dp = D[i];
        double* Jp = J[i];
        fill(T1p, nthreads_ * naux_, 0.0, nthreads_);
#pragma omp parallel for schedule(guided) num_threads(nthreads_)
        for (size_t k = 0; k < nbf_; k++) {
            size_t sp_size = c->small_skips[k];
            size_t jump = c->big_skips[k] + bcount * sp_size;
            int rank = 0;
#ifdef _OPENMP
            rank = omp_get_thread_num();
#endif
            for (size_t m = 0, sp_count = (size_t)-1; m < nbf_; m++) {
                if (c->schwarz_fun_index[k * nbf_ + m]) {
                    sp_count++;
                    D_buffers[rank][sp_count] = Dp[nbf_ * k + m];
                }
            }
            C_DGEMV('N', (int)block_size, (int)sp_size, 1.0, &Mp[jump], (int)sp_size, &D_buffers[rank][0], 1, 1.0,
                    &T1p[rank * naux_], 1);
        }
        for (size_t k = 1; k < (size_t)nthreads_; k++)
            for (size_t l = 0; l < naux_; l++) T1p[l] += T1p[k * naux_ + l];
#pragma omp parallel for schedule(guided) num_threads(nthreads_)
        for (size_t k = 0; k < nbf_; k++) {
            size_t sp_size = c->small_skips[k];
            size_t jump = c->big_skips[k] + bcount * sp_size;
            C_DGEMV('T', (int)block_size, (int)sp_size, 1.0, &Mp[jump], (int)sp_size, T1p, 1, 0.0, &T2p[k * nbf_], 1);
        }
        for (size_t k = 0; k < nbf_; k++) {
            for (size_t m = 0, count = (size_t)-1; m < nbf_; m++) {
                if (c->schwarz_fun_index[k * nbf_ + m]) {
                    count++;
                    Jp[k * nbf_ + m] += T2p[k * nbf_ + count];
                }
            }
        }
    }
}

/* dfhelper.cc:2162-2186 */
static void first_transform_pQq(const oracle_ctx* c, size_t bsize, size_t bcount, size_t block_size, const double* Mp,
                                double* Tp, const double* Bp, double** C_buffers) {
    const size_t nbf_ = c->nbf;
    const int nthreads_ = c->nthreads;
#pragma omp parallel for schedule(guided) num_threads(nthreads_)
    for (size_t k = 0; k < nbf_; k++) {
        size_t sp_size = c->small_skips[k];
        size_t jump = c->big_skips[k] + bcount * sp_size;
        int rank = 0;
#ifdef _OPENMP
        rank = omp_get_thread_num();
#endif
        for (size_t m = 0, sp_count = (size_t)-1; m < nbf_; m++) {
            if (c->schwarz_fun_index[k * nbf_ + m]) {
                sp_count++;
                C_DCOPY(bsize, &Bp[m * bsize], 1, &C_buffers[rank][sp_count * bsize], 1);
            }
        }
        C_DGEMM('N', 'N', (int)block_size, (int)bsize, (int)sp_size, 1.0, &Mp[jump], (int)sp_size, &C_buffers[rank][0],
                (int)bsize, 0.0, &Tp[k * block_size * bsize], (int)bsize);
    }
}

/* dfhelper.cc:3350-3377 */
static void compute_K(const oracle_ctx* c, int nmat, double* const* Cleft, double* const* Cright, const int* nocc_i,
                      double* const* K, double* T1p, double* T2p, const double* Mp, size_t bcount, size_t block_size,
                      double** C_buffers, int lr_symmetric) {
    const size_t nbf_ = c->nbf;
    for (int i = 0; i < nmat; i++) {
        size_t nocc = (size_t)nocc_i[i];
        if (!nocc) continue;
        const double* Clp = Cleft[i];
        const double* Crp = Cright[i];
        double* Kp = K[i];
        first_transform_pQq(c, nocc, bcount, block_size, Mp, T1p, Clp, C_buffers);
        double* T2use = T2p;
        if (lr_symmetric)
            T2use = T1p;
        else
            first_transform_pQq(c, nocc, bcount, block_size, Mp, T2p, Crp, C_buffers);
        C_DGEMM('N', 'T', (int)nbf_, (int)nbf_, (int)(nocc * block_size), 1.0, T1p, (int)(nocc * block_size), T2use,
                (int)(nocc * block_size), 1.0, Kp, (int)nbf_);
    }
}

/* Timing of the last oracle_build_JK call (seconds): J, K half-transforms+GEMM, wK. */
static double last_t[4];
static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}
void oracle_last_timings(double* out4) { memcpy(out4, last_t, sizeof last_t); }

/*
 * dfhelper.cc:3015-3043 (build_JK) + :3044-3161 (compute_JK) + :3378-3438 (compute_wK), in-core
 * (AO_core_ = true) case, plus the caller-side steps of MemDFJK::compute_JK (libfock/MemDFJK.cc:97-111):
 * zero() of the requested outputs and hermitivitize() of wK when lr_symmetric.
 *
 *  naux_store / q_block: the tensors hold Q rows [0, naux_store) (big_skips built with that
 *  extent).  The Q loop runs in blocks of q_block rows (0 => one block), standing in for
 *  Qshell_blocks_for_JK_build (:814-869), whose only effect in core is this blocking with
 *  beta=1 accumulation across blocks.
 *
 *  Buffers are allocated on every call exactly as the reference does (:3071-3108).
 */
int oracle_build_JK(size_t nbf, size_t naux_store, int nthreads, const size_t* schwarz_fun_index,
                    const size_t* small_skips, const size_t* big_skips, const size_t* symm_small_skips,
                    const size_t* symm_ignored_columns, const double* Ppq, const double* m1Ppq, const double* wPpq,
                    int nmat, double* const* Cleft, double* const* Cright, const int* nocc, double* const* D,
                    double* const* J, double* const* K, double* const* wK, int do_J, int do_K, int do_wK,
                    int lr_symmetric, size_t q_block) {
    if (!F_DGEMM) return 1;
    oracle_ctx c = {nbf, naux_store, nthreads, schwarz_fun_index, small_skips, big_skips, symm_small_skips,
                    symm_ignored_columns};
    const size_t nbf_ = nbf, naux_ = naux_store;
    const size_t nthreads_ = (size_t)nthreads;
    if (q_block == 0 || q_block > naux_) q_block = naux_;
    size_t max_nocc = 0;
    for (int i = 0; i < nmat; i++)
        if ((size_t)nocc[i] > max_nocc) max_nocc = (size_t)nocc[i];

    /* MemDFJK.cc:100 / jk.cc:690-703 */
    for (int i = 0; i < nmat; i++) {
        if (do_J) memset(J[i], 0, sizeof(double) * nbf * nbf);
        if (do_K) memset(K[i], 0, sizeof(double) * nbf * nbf);
        if (do_wK) memset(wK[i], 0, sizeof(double) * nbf * nbf);
    }
    memset(last_t, 0, sizeof last_t);

    size_t cbuf = nbf_ * (max_nocc > nbf_ ? max_nocc : nbf_);
    double** C_buffers = (double**)malloc(sizeof(double*) * nthreads_);
    if (do_J || do_K) {
        /* :3071-3081 */
#pragma omp parallel num_threads(nthreads)
        {
            int rank = 0;
#ifdef _OPENMP
            rank = omp_get_thread_num();
#endif
            C_buffers[rank] = (double*)calloc(cbuf, sizeof(double));
        }
        size_t totsb = q_block;
        size_t Ktmp_size = (!max_nocc ? totsb * 1 : totsb * max_nocc);
        Ktmp_size = Ktmp_size * nbf_ > nthreads_ * naux_ ? Ktmp_size * nbf_ : nthreads_ * naux_;
        if (Ktmp_size < nbf_ * nbf_) Ktmp_size = nbf_ * nbf_;
        double* T1p = (double*)malloc(sizeof(double) * Ktmp_size);
        size_t K2 = lr_symmetric ? nbf_ * nbf_ : (nbf_ * nbf_ > Ktmp_size ? nbf_ * nbf_ : Ktmp_size);
        if (K2 < nbf_ * nbf_) K2 = nbf_ * nbf_;
        if (K2 < nthreads_ * naux_) K2 = nthreads_ * naux_;
        double* T2p = (double*)malloc(sizeof(double) * K2);
        if (!T1p || !T2p) return 3;

        size_t bcount = 0;
        while (bcount < naux_) {
            size_t block_size = naux_ - bcount < q_block ? naux_ - bcount : q_block;
            if (do_J) {
                double t0 = now_s();
                if (lr_symmetric)
                    compute_J_symm(&c, nmat, D, J, Ppq, T1p, T2p, C_buffers, bcount, block_size);
                else
                    compute_J(&c, nmat, D, J, Ppq, T1p, T2p, C_buffers, bcount, block_size);
                last_t[0] += now_s() - t0;
            }
            if (do_K) {
                double t0 = now_s();
                compute_K(&c, nmat, Cleft, lr_symmetric ? Cleft : Cright, nocc, K, T1p, T2p, Ppq, bcount, block_size,
                          C_buffers, lr_symmetric);
                last_t[1] += now_s() - t0;
            }
            bcount += block_size;
        }
        free(T1p);
        free(T2p);
        for (size_t r = 0; r < nthreads_; r++) free(C_buffers[r]);
    }

    if (do_wK) {
        /* :3378-3438 */
        double t0 = now_s();
#pragma omp parallel num_threads(nthreads)
        {
            int rank = 0;
#ifdef _OPENMP
            rank = omp_get_thread_num();
#endif
            C_buffers[rank] = (double*)calloc(cbuf, sizeof(double));
        }
        size_t totsb = q_block;
        size_t Ktmp_size = (!max_nocc ? totsb * 1 : totsb * max_nocc);
        Ktmp_size = Ktmp_size * nbf_ > nthreads_ * naux_ ? Ktmp_size * nbf_ : nthreads_ * naux_;
        double* T1p = (double*)malloc(sizeof(double) * Ktmp_size);
        double* T2p = (double*)malloc(sizeof(double) * Ktmp_size);
        if (!T1p || !T2p) return 3;
        for (size_t bcount = 0; bcount < naux_;) {
            size_t block_size = naux_ - bcount < q_block ? naux_ - bcount : q_block;
            for (int i = 0; i < nmat; i++) {
                size_t no = (size_t)nocc[i];
                if (!no) continue;
                const double* Clp = Cleft[i];
                const double* Crp = lr_symmetric ? Cleft[i] : Cright[i];
                first_transform_pQq(&c, no, bcount, block_size, m1Ppq, T1p, Clp, C_buffers);
                first_transform_pQq(&c, no, bcount, block_size, wPpq, T2p, Crp, C_buffers);
                C_DGEMM('N', 'T', (int)nbf_, (int)nbf_, (int)(no * block_size), 1.0, T1p, (int)(no * block_size), T2p,
                        (int)(no * block_size), 1.0, wK[i], (int)nbf_);
            }
            bcount += block_size;
        }
        free(T1p);
        free(T2p);
        for (size_t r = 0; r < nthreads_; r++) free(C_buffers[r]);
        /* MemDFJK.cc:104-110; Matrix::hermitivitize = (A + A^T)/2 (libmints/matrix.cc) */
        if (lr_symmetric) {
            for (int i = 0; i < nmat; i++) {
                double* w = wK[i];
                for (size_t a = 0; a < nbf; a++)
                    for (size_t b = 0; b < a; b++) {
                        double v = 0.5 * (w[a * nbf + b] + w[b * nbf + a]);
                        w[a * nbf + b] = w[b * nbf + a] = v;
                    }
            }
        }
        last_t[2] += now_s() - t0;
    }
    free(C_buffers);
    return 0;
}

/* libfock/jk.cc:314-354 (compute_D), C1 case: D = Cl * Cr^T, row-major nbf x nocc inputs. */
void oracle_compute_D(size_t nbf, int nocc, const double* Cl, const double* Cr, double* D) {
    memset(D, 0, sizeof(double) * nbf * nbf);
    if (!nocc) return;
    C_DGEMM('N', 'T', (int)nbf, (int)nbf, nocc, 1.0, Cl, nocc, Cr, nocc, 0.0, D, (int)nbf);
}

/* ------------------------------------------------------------------------------------------
 * Synthetic DF tensor generator (bench / large-size tests).  NOT part of the reference: psi4
 * has no synthetic workload.  Bit-identical twin of the device generator in
 * psi4_b200/csrc/synth.cuh: value(Q,m,n) = amp[m*nbf+n] * u(seed, Q, min(m,n), max(m,n)),
 * u in [-1,1) from a splitmix64 counter hash; exact in IEEE double on both sides.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline double synth_u(uint64_t seed, uint64_t Q, uint64_t m, uint64_t n, uint64_t nbf) {
    uint64_t lo = m < n ? m : n, hi = m < n ? n : m;
    uint64_t ctr = (Q * nbf + lo) * nbf + hi;
    uint64_t h = splitmix64(splitmix64(seed) ^ ctr);
    return (double)(int64_t)(h >> 11) * (1.0 / 4503599627370496.0) - 1.0; /* 2^-52 */
}
/* Fill packed rows Q in [q0, q0+nq) of every m block: out has big_skips built for nq rows. */
void oracle_synth_fill(size_t nbf, size_t q0, size_t nq, uint64_t seed, const double* amp, const size_t* fun_index,
                       const size_t* small_skips, const size_t* big_skips_nq, double* out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t m = 0; m < nbf; m++) {
        size_t sp = small_skips[m];
        for (size_t n = 0; n < nbf; n++) {
            size_t f = fun_index[m * nbf + n];
            if (!f) continue;
            double a = amp[m * nbf + n];
            for (size_t q = 0; q < nq; q++)
                out[big_skips_nq[m] + q * sp + f - 1] = a * synth_u(seed, q0 + q, m, n, nbf);
        }
    }
}
