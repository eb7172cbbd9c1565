/*
 * b200jk.h -- C ABI of the B200-native density-fitted J/K engine (libb200jk.so).
 *
 * This is the drop-in boundary behind psi4's MEM_DF JK object.  Every entry point names the
 * reference interface it replaces (paths relative to psi4/src/psi4/ in the psi4 tree).
 * Plain C types only: no C++ classes, no exceptions, no torch types.  All matrices are
 * row-major contiguous doubles, exactly what psi4's Matrix::get_pointer() / pointer()[0]
 * hands out for a C1 matrix (libmints/matrix.h:553).
 *
 * Call order for one SCF:
 *     b200jk_create[_rank]            <- MemDFJK ctor / DFHelper ctor      (libfock/MemDFJK.cc:56-64)
 *     b200jk_set_layout               <- DFHelper::prepare_sparsity tables (lib3index/dfhelper.cc:299-420)
 *     b200jk_upload | _upload_rows    <- Ppq_/m1Ppq_/wPpq_ after prepare_AO_core (:514-588, :589-699)
 *     b200jk_compute  (every iteration) <- DFHelper::build_JK              (:3015-3043)
 *     b200jk_destroy
 *
 * Threading: calls on one handle must be serialised by the caller (as JK::compute is, from the
 * Python main thread).  Several handles may coexist (the SAD guess builds its own JK, sad.cc:706).
 *
 * There is no CPU fallback: if no CUDA device / insufficient HBM, calls fail with an error code
 * (mirrors SCF_SUBTYPE=INCORE throwing at dfhelper.cc:259-262 rather than degrading).
 */
#ifndef B200JK_H
#define B200JK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200jk b200jk_t;

/* Status codes (psi4 glue turns non-zero into PSIEXCEPTION, libpsi4util/exception.h:48). */
enum {
    B200JK_OK = 0,
    B200JK_ERR_INVALID = 1,  /* bad argument / call order                      */
    B200JK_ERR_CUDA = 2,     /* CUDA runtime error (message in last_error)     */
    B200JK_ERR_OOM = 3,      /* tensor + work buffers do not fit in HBM        */
    B200JK_ERR_NCCL = 4,     /* NCCL error                                     */
    B200JK_ERR_NODEVICE = 5  /* no CUDA device: engine refuses to run          */
};

/* Which three-index tensor (lib3index/dfhelper.h:380,385-386). */
enum {
    B200JK_TENSOR_PPQ = 0,   /* Ppq_   = J^-1/2 (A|mn)            J and K      */
    B200JK_TENSOR_M1PPQ = 1, /* m1Ppq_ = J^-1   (A|mn)            wK left      */
    B200JK_TENSOR_WPPQ = 2   /* wPpq_  = (A|erf(w r)/r|mn)        wK right     */
};

#define B200JK_NCCL_ID_BYTES 128

/* ---- construction --------------------------------------------------------------------------- */

/* One process driving `ngpu` devices (psi4 is single-process: libfock/jk.cc:143-150 constructs
 * exactly one MemDFJK).  The auxiliary index Q is split into ngpu contiguous shards; partial
 * J/K/wK are summed with one NCCL all-reduce per build.  dev_ids may be NULL (0..ngpu-1). */
int b200jk_create(b200jk_t** out, int ngpu, const int* dev_ids);

/* One process per GPU (torchrun-style launch): this rank owns Q shard `rank` of `world`.
 * nccl_id: B200JK_NCCL_ID_BYTES bytes from b200jk_nccl_unique_id() on rank 0, broadcast by the
 * caller (ignored when world == 1). */
int b200jk_create_rank(b200jk_t** out, int device, int rank, int world, const void* nccl_id);
int b200jk_nccl_unique_id(void* out_bytes /* B200JK_NCCL_ID_BYTES */);

void b200jk_destroy(b200jk_t* h);

/* ---- layout + tensor upload ------------------------------------------------------------------ */

/* Tables of DFHelper::prepare_sparsity (dfhelper.cc:377-398), passed verbatim:
 *   small_skips[nbf+1]  small_skips_    sp(m), last entry = total kept pairs
 *   big_skips[nbf+1]    big_skips_      offset of row-block m in the packed tensor (naux*sp(m) each)
 *   fun_index[nbf*nbf]  schwarz_fun_index_  1-based rank of n among kept partners of m, 0 = dropped
 * Element B(Q,m,n) lives at host_pQq[big_skips[m] + Q*sp(m) + fun_index[m*nbf+n] - 1]
 * (dfhelper.cc:1274-1276, :1671-1672).  The diagonal must be kept (assumed at :3213). */
int b200jk_set_layout(b200jk_t* h, size_t nbf, size_t naux, const size_t* small_skips, const size_t* big_skips,
                      const size_t* fun_index);

/* Copy a whole host tensor (big_skips[nbf] doubles) into HBM, Q-sharded.  The host buffer is not
 * retained; psi4 may release Ppq_ afterwards. */
int b200jk_upload(b200jk_t* h, int which, const double* host_pQq);

/* Streaming variant for the p-blocked construction loop of prepare_AO_core (:540-587): rows
 * m in [m0, m1) only; host_rows points at the first double of row-block m0, i.e. what psi4 holds
 * at Ppq_ + big_skips[m0].
 * Both producers (this call and b200jk_fit_rows) are pipelined: the call returns as soon as the caller's block has
 * been READ -- a pageable block is copied into the engine's page-locked staging ring, a block inside a range
 * registered with b200jk_register_host is DMA'd in place and the call waits for that copy only -- so the buffer may
 * be refilled at once (psi4 reuses one block buffer, :553-585) while the transfer and the kernels run behind the
 * call.  The block with m1 == nbf completes the tensor and drains the pipeline; compute / download calls are ordered
 * behind whatever is still in flight. */
int b200jk_upload_rows(b200jk_t* h, int which, size_t m0, size_t m1, const double* host_rows);

/* ---- on-device fitting --------------------------------------------------------
 * Moves DFHelper::contract_metric_AO_core_symm (dfhelper.cc:1653-1678) to the GPU: psi4 keeps computing the
 * UNFITTED integrals on the CPU (compute_sparse_pQq_blocking_p_symm, :1284-1347) and hands each block over instead of
 * running naux DGEMMs per basis function itself; the host never holds the fitted tensor and uploads half the bytes.
 *
 * b200jk_set_metric: naux x naux row-major metric power that prepare_AO_core contracts with (metp, :560-566):
 *   J^-1/2 for Ppq_, J^-1 for m1Ppq_ (:642-650).  NULL = no contraction (wPpq_, :688-692: plain scatter + mirror).
 *   Each Q shard keeps only its rows.
 * b200jk_fit_rows: rows m in [m0, m1) of the symmetric-packed unfitted buffer exactly as psi4's Mp holds them:
 *   host_sym[symm_big_skips[m] - symm_big_skips[m0] + Q*mi(m) + (f(m,n) - f(m,m))],  n >= m kept, mi(m) = symm_small_skips[m]
 *   (:1338-1340).  Computes Ppq[m][Q][n>=m] = sum_R metric[Q,R] (R|mn) and the mirror copy B(Q,n,m) = B(Q,m,n)
 *   (:1666-1677).  The block with m1 == nbf completes the tensor.  The pair mask must be symmetric. */
int b200jk_set_metric(b200jk_t* h, const double* metric);
int b200jk_fit_rows(b200jk_t* h, int which, size_t m0, size_t m1, const double* host_sym);
/* device time (ms) and flops of the metric contraction accumulated since the fit_rows call with m0 == 0 */
int b200jk_fit_stats(const b200jk_t* h, double* ms_gemm, double* flops);

/* ---- the hot path ----------------------------------------------------------------------------- */

/* DFHelper::build_JK (dfhelper.cc:3015-3043) + the zero()/hermitivitize() wrapper of
 * MemDFJK::compute_JK (libfock/MemDFJK.cc:97-111).
 *   nmat            number of (C_left, C_right, D) triples (C_left_ao_.size())
 *   Cl[i], Cr[i]    nbf x nocc[i], row-major, ld = nocc[i].  Cr == NULL  <=> lr_symmetric_
 *                   (jk.cc:597-602): K uses T2=T1 (:3367-3368), J uses the upper triangle of D (:3188).
 *   nocc[i]         Cleft[i]->colspi()[0]; 0 => K[i]/wK[i] left zero (:3354-3357)
 *   D[i]            nbf x nbf density built by the caller (jk.cc:314-354); only read if do_J
 *   J[i],K[i],wK[i] nbf x nbf outputs, OVERWRITTEN with the result (the reference zeroes them in
 *                   MemDFJK.cc:100 and accumulates with beta=1; the sum is identical).  Arrays for
 *                   untasked products may be NULL.
 *   wK              requires tensors M1PPQ and WPPQ; symmetrised (A+A^T)/2 when lr_symmetric
 *                   (MemDFJK.cc:104-110).
 *   repeated C_left When Cr != NULL and Cl[i] is the matrix Cl[i-1] again (same pointer, or same nocc and values) --
 *                   the reference's response builds push ONE occupied block as every C_left (twoel_Hx,
 *                   libscf_solver/rhf.cc:466-484) -- the first half transform is kept instead of recomputed
 *                   (dfhelper.cc:3364 recomputes it); results are bit-identical.  B200JK_NO_T1_REUSE=1 disables it.
 *   worker ranks    In rank mode a rank other than 0 may pass J = K = wK = NULL: it takes part in the build and the
 *                   all-reduce and brings nothing home.
 * Host pointers are not retained past the call. */
int b200jk_compute(b200jk_t* h, int nmat, const double* const* Cl, const double* const* Cr, const int* nocc,
                   const double* const* D, double* const* J, double* const* K, double* const* wK, int do_J,
                   int do_K, int do_wK);

/* Optional: page-lock caller memory that persists across builds (psi4 allocates D_ao_/J_ao_/K_ao_ once in
 * JK::allocate_JK / USO2AO, jk.cc:355-446) so b200jk_compute DMAs straight from/to it instead of staging through
 * the engine's pinned buffers.  Any D/J/K/wK pointer inside a registered range is used in place; everything else
 * is staged.  Unregister before freeing the memory (b200jk_destroy unregisters what is left). */
int b200jk_register_host(b200jk_t* h, void* ptr, size_t bytes);
int b200jk_unregister_host(b200jk_t* h, void* ptr);

/* Same build with every operand already resident in this rank's HBM (device pointers); used by
 * the kernel-only timing of bench.py and by callers that keep C/D on the GPU.  Rank mode only
 * (one shard per handle).  Outputs hold the all-reduced result on every rank. */
int b200jk_compute_device(b200jk_t* h, int nmat, const double* const* dCl, const double* const* dCr,
                          const int* nocc, const double* const* dD, double* const* dJ, double* const* dK,
                          double* const* dwK, int do_J, int do_K, int do_wK);

/* ---- density-fitted SCF gradient: the tensor contractions of DFJKGrad -------------------
 * Replaces the CPU work of scfgrad/jk_grad.cc build_Amn_terms (:294-475), build_AB_inv_terms (:634-721), build_UV_terms
 * (:722-833) and the AO back-transform at the top of build_Amn_x_terms (:1010-1023) -- everything of
 * DFJKGrad::compute_gradient except the Libint2 derivative integrals (A|B)^x, (A|mn)^x and their final dot products,
 * which stay in psi4.  Nothing is recomputed: the intermediates come from the fitted tensor that is already resident
 * (one first-J-sweep pass and one half transform per spin) and the metric power the caller passes.
 *
 * b200jk_grad_begin   nspin = 1: restricted (the reference's Ca_ == Cb_: one transform, factor 2), 2: unrestricted.
 *                     C[s]: nbf x nocc[s] occupied orbitals (Ca_occ, Cb_occ); Dt: nbf x nbf total density (Da + Db,
 *                     symmetric); Jm12: naux x naux, the same J^-1/2 the tensor was fitted with.
 * b200jk_grad_vectors d[naux]        = J^-1 (A|mn) Dt_mn                       (the "c" entry after :659)
 *                     V[naux x naux] = f sum_spin sum_ij (A|ij)(B|ij), fitted   (the "V" entry, :812)
 * b200jk_grad_rows    Kmn[(a1-a0) x nbf x nbf] = f sum_spin C (A|ij) C^T for aux rows [a0, a1): what the reference
 *                     holds in Kmnp for one block of auxiliary shells (:1010-1023); call it block by block and dot
 *                     each block with (A|mn)^x on the host.
 * b200jk_grad_end     frees the intermediates.
 * One-shard handles only (the gradient runs once per geometry). */
int b200jk_grad_begin(b200jk_t* h, int nspin, const double* const* C, const int* nocc, const double* Dt,
                      const double* Jm12);
int b200jk_grad_vectors(b200jk_t* h, double* d, double* V);
int b200jk_grad_rows(b200jk_t* h, size_t a0, size_t a1, double* Kmn);
int b200jk_grad_end(b200jk_t* h);

/* ---- setup: Matrix::power on the device --------------------------------------------------------------------------
 * Matrix::power(alpha, cutoff) (libmints/matrix.cc:2370-2424) as DFHelper::prepare_metric / compute_metric call it for
 * the fitting metric (lib3index/dfhelper.cc:1462-1517; alpha = -1/2 for Ppq_, -1 for m1Ppq_): eigendecomposition,
 * eigenvalues with |lambda| < cutoff * max|lambda| dropped when alpha < 0 (and any non-finite lambda^alpha), then
 * V f(Lambda) V^T.  A and out are n x n row-major host arrays (A symmetric; may alias).  remaining: eigenvalues kept (the
 * Dimension the reference returns); ms_device: device time of the decomposition + reconstruction.  The decomposition is
 * cuSOLVER's (dlopen'ed); the drop rule is evaluated on the host with the reference's own libm pow; the reconstruction
 * runs in the engine's kernels. */
int b200jk_matrix_power(b200jk_t* h, size_t n, const double* A, double alpha, double cutoff, double* out,
                        int* remaining, double* ms_device);

/* ---- introspection ---------------------------------------------------------------------------- */

typedef struct {
    /* device time (CUDA events on the engine's stream, max over local shards) of the last compute */
    double ms_total;      /* first kernel to last kernel incl. all-reduce          */
    double ms_j;          /* J sweeps (K1+K2)                                      */
    double ms_half;       /* half-transforms (K3)                                  */
    double ms_kgemm;      /* K GEMM (K4) incl. split-K reduction                   */
    double ms_allreduce;  /* NCCL all-reduce                                       */
    double ms_h2d, ms_d2h; /* host<->device copies of C/D and J/K/wK                */
    /* algorithmic work of the last compute on THIS handle's shards  */
    double j_bytes;       /* bytes of B the J sweeps must read                     */
    double half_flops;    /* 2*A*P*o per transform                                 */
    double half_bytes;    /* 8*A*P read + 8*N*A*o written per transform            */
    double kgemm_flops;   /* flops executed by the K GEMM (N(N+1)*A*o if symmetric)*/
    uint64_t launches;    /* kernels launched by the last compute                  */
    uint64_t hbm_tensor_bytes; /* bytes of HBM holding the packed tensors          */
    uint64_t hbm_work_bytes;   /* bytes of HBM in work buffers                     */
    int n_shards;         /* local shards (GPUs driven by this handle)             */
    int q_begin, q_end;   /* Q range of local shard 0                              */
    int reduce_kind;      /* cross-GPU sum: 0 none (one GPU), 1 fixed-rank-order peer-memory kernel over NVLink,
                             2 NCCL all-reduce (B200JK_REDUCE=nccl, or no peer access between the devices) */
    int kgemm_kind;       /* arm the K GEMM took: 0 FP64 tensor pipe (DMMA), 1 INT8 tensor cores by residues       */
    int kgemm_moduli;     /* moduli of the residue arm (13: 50 bits below the row norm)                            */
    int half_kind;        /* arm the half transform took: 0 FP64 tensor pipe (DMMA), 1 INT8 tensor cores by residues */
    /* sub-phases of the INT8 arms (CUDA events between their launches, max over local shards; zero on the DMMA arms) */
    double ms_half_i8[4];  /* conversion to residue planes (+ fused first J sweep), C^T gather, tcgen05 GEMM, CRT    */
    double ms_kgemm_i8[3]; /* row scales + residue planes of T, tcgen05 GEMM, CRT                                    */
    double half_i8_ops;    /* 2 x int8 multiply-adds the half transform issued to the tensor cores (all moduli,
                              whole padded tiles), this handle's shards                                              */
    double half_i8_plane_bytes;   /* residue-plane bytes its GEMM read                                               */
    double half_i8_convert_bytes; /* bytes the conversions that RAN moved: f64 rows read + planes written (0: cached) */
    double kgemm_i8_ops;   /* the same count for the K GEMM                                                          */
    int half_moduli;       /* moduli of the half transform's residue arm (12: 47 bits below |B row| |C column|)      */
    int half_i8_chunks;    /* row-block chunks the scratch arena forced (summed over transforms)                     */
    int half_i8_cached;    /* transforms that found their residue planes cached from an earlier build                */
    int half_i8_resident_rows; /* row-blocks [0, n) whose residue planes stay in HBM across builds (nbf: all of them)  */
} b200jk_stats;

int b200jk_get_stats(const b200jk_t* h, b200jk_stats* out);
const char* b200jk_last_error(const b200jk_t* h);

/* DFHelper::get_core_size analogue (dfhelper.cc:216-236) for HBM: bytes needed per GPU for the
 * tensors (+2 more with wK) and work buffers at the given max_nocc; usable before upload. */
int b200jk_hbm_estimate(const b200jk_t* h, size_t max_nocc, int do_wK, uint64_t* bytes_per_gpu);

/* Limit the HBM the half-transformed intermediate T may take (bytes per GPU; 0 = automatic).
 * Stands in for DFHelper::set_memory + Qshell_blocks_for_JK_build (:814-869): smaller budgets
 * make the K build loop over Q chunks. */
int b200jk_set_work_budget(b200jk_t* h, uint64_t bytes);

/* Arm of the K GEMM  K += T1 T2^T  (the C_DGEMM of DFHelper::compute_K, lib3index/dfhelper.cc:3374).
 *   arm 0  automatic: the INT8 residue arm from 128 basis functions and 2^24 elements of T on, when its planes fit
 *   arm 1  FP64 tensor pipe (DMMA m8n8k4, kgemm_ws_kernel)
 *   arm 2  INT8 tensor cores (tcgen05.mma.kind::i8): the rows of T are scaled to integers of `moduli`-dependent width,
 *          multiplied exactly modulo `moduli` coprime numbers <= 256 and rebuilt by the Chinese remainder theorem.
 *          The only rounding is the scaling of T: |dK[m,n]| <= 2^-bits * sqrt(kdim) * |T[m,:]| |T[n,:]| worst case,
 *          ~2^-bits |T[m,:]| |T[n,:]| observed, bits = 50.7 at 13 moduli (default, = what a DGEMM in double commits),
 *          46.9 at 12, 43.0 at 11; bit-identical run to run for any split of the work.
 * moduli: 0 = default (13), else 6..13.  Environment: B200JK_KGEMM=dmma|i8, B200JK_I8_MODULI=n. */
int b200jk_set_kgemm(b200jk_t* h, int arm, int moduli);

/* The same choice for the half transform  T[m,Q,i] = sum_n B[Q,m,n] C[n,i]  (DFHelper::first_transform_pQq,
 * lib3index/dfhelper.cc:2162-2186): arm 0 automatic, 1 FP64 tensor pipe (half_ws_kernel), 2 INT8 tensor cores.  The
 * residue arm scales every row (m,Q) of the resident tensor (once per tensor) and every column of C (once per build) to
 * integers, multiplies modulo `moduli` coprime numbers (default 12: 46.9 bits below |B[Q,m,:]| |C[:,i]|) and rebuilds T
 * by the Chinese remainder theorem.  automatic = from 256 basis functions, 16 occupied orbitals and 2^27 tensor elements per
 * shard on.  The residue planes of as many row-blocks as the scratch arena holds stay in HBM until the tensor changes (all of
 * them from two GPUs on at C60; stats.half_i8_resident_rows), the rest is converted in every build.  The first J sweep is one
 * more column of the residue GEMM (the density row of each row-block; needs nocc not a multiple of 16), with the arm's
 * norm-wise error 2^-46.9 |B[Q,m,:]| |D'[m,:]|.
 * Environment: B200JK_HALF=dmma|i8, B200JK_I8_HALF_MODULI=n, B200JK_I8_CLUSTER=1|2|4, B200JK_I8_RESIDENT=0 (planes resident only
 * when all of them fit), B200JK_I8_JCOL=0 (first J sweep as its own FP64 kernel), B200JK_I8_ARENA_MB=n (upper bound of the
 * scratch arena; tests). */
int b200jk_set_half(b200jk_t* h, int arm, int moduli);

/* ---- synthetic workload support (bench.py / large-size tests only; not in the reference) ------ */

/* Fill tensor `which` on the device with value(Q,m,n) = amp[m*nbf+n] * u(seed,Q,min(m,n),max(m,n)),
 * u in [-1,1) from a counter hash -- bit-identical to oracle_synth_fill() in oracle/dfjk_oracle.c.
 * Lets C60-size tensors (123 GB) exist without a host copy. */
int b200jk_fill_synthetic(b200jk_t* h, int which, uint64_t seed, const double* amp /* nbf*nbf host */);

/* Read back rows Q in [q0,q1) of row-block m (packed, sp(m) per row) from local shard holding them. */
int b200jk_download_rows(b200jk_t* h, int which, size_t m, size_t q0, size_t q1, double* host_out);

/* Device pointer helpers for b200jk_compute_device callers that have no CUDA runtime binding. */
int b200jk_dev_alloc(b200jk_t* h, size_t bytes, void** dptr);
int b200jk_dev_free(b200jk_t* h, void* dptr);
int b200jk_dev_copy(b200jk_t* h, void* dst, const void* src, size_t bytes, int kind /*1=H2D 2=D2H 3=D2D*/);

/* FP64 calibration kernels (register-resident loops, no memory traffic) for the roofline denominator that
 * MEASURED_PEAKS.json lacks.  kind: 0 = DMMA m8n8k4, 1 = DFMA, 2 = both interleaved in every warp.
 * out4 = { burst TFLOP/s (best single launch from idle), sustained TFLOP/s (back-to-back launches for
 * `seconds`), SM MHz during the burst launch, SM MHz during the last sustained launch }; the clocks come
 * from clock64()/globaltimer inside the kernel. */
int b200jk_fp64_peak(b200jk_t* h, int kind, double seconds, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* B200JK_H */
