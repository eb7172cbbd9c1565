// engine.cu -- libb200jk.so: the C ABI of include/b200jk.h over the sm_100a kernels.
//
// Host-side orchestration of the in-core DF-JK build (what DFHelper::compute_JK /
// compute_wK do on the CPU, lib3index/dfhelper.cc:3044-3161, :3378-3438) with the packed
// (Q|mn) tensor resident in HBM, sharded over the auxiliary index Q.
//
// No CPU fallback: every entry point that needs a device fails with an error code if none.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/b200jk.h"
#include "aux_kernels.cuh"
#include "dmma_gemm.cuh"
#include "dmma_ws.cuh"
#include "fit_kernels.cuh"
#include "j_kernels.cuh"
#include "gemm_strided.cuh"
#include "peer_reduce.cuh"
#include "i8_kgemm.cuh"
#include "i8_half.cuh"

using namespace b2k;

// ------------------------------------------------------------------------------------------------
// NCCL, resolved lazily (single-GPU handles never touch it).  If the process already has a
// libnccl.so.2 loaded (e.g. torch's bundled copy) the soname lookup returns that one.
// ------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
static_assert(sizeof(ncclUniqueId) == B200JK_NCCL_ID_BYTES, "nccl id size");
struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return false;
        }
#define SYM(f)                                                           \
    *(void**)(&f) = dlsym(lib, "nccl" #f);                               \
    if (!f) {                                                            \
        err = "libnccl lacks nccl" #f;                                   \
        return false;                                                    \
    }
        SYM(GetUniqueId) SYM(CommInitRank) SYM(CommInitAll) SYM(CommDestroy) SYM(AllReduce) SYM(AllGather) SYM(GroupStart)
        SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
        return true;
    }
};
Nccl g_nccl;
constexpr int kNcclDouble = 8;  // ncclFloat64
constexpr int kNcclSum = 0;
constexpr int kNcclMin = 3;
constexpr int kNcclInt8 = 0, kNcclInt32 = 2;

// Peer-memory window of one shard for the fixed-order cross-GPU sum (peer_reduce.cuh).  The window IS the shard's
// partial-result buffer `out`; the flag block is a separate small allocation that never moves.
struct PeerCtx {
    unsigned* flags = nullptr;                 // own flag block: PR_CHANNELS x PR_FLAG_WORDS words
    unsigned* status = nullptr;                // page-locked host word: a barrier of the peer sum timed out
    unsigned* peer_flags[PR_MAXP] = {nullptr}; // every rank's flag block, mapped here (own entry = flags)
    double* peer_win[PR_MAXP] = {nullptr};     // every rank's window, mapped here (own entry = out)
    void* ipc_flags_base[PR_MAXP] = {nullptr}; // what cudaIpcOpenMemHandle returned (rank mode), to close later
    void* ipc_win_base[PR_MAXP] = {nullptr};
    void* xchg = nullptr;                      // device scratch of the handle exchange: world x 128 bytes
    bool flags_open = false, win_open = false;
    const double* win_key = nullptr;           // the `out` pointer the peers currently map
    unsigned epoch[PR_CHANNELS] = {0, 0};
};

struct Phase {
    int tag;  // 0 J, 1 half, 2 kgemm, 3 allreduce, 4 h2d, 5 d2h, 6 total
    cudaEvent_t a, b;
};

struct Shard {
    int dev = 0;
    int q0 = 0, q1 = 0, nq = 0;
    cudaStream_t stream = nullptr;  // compute stream
    cudaStream_t copy = nullptr;    // copy stream: D upload and K download overlap the kernels
    double* tensor[3] = {nullptr, nullptr, nullptr};
    CUtensorMap* d_amaps[3] = {nullptr, nullptr, nullptr};  // per-row-block TMA descriptors of each tensor
    int nsm = 148;
    int* d_counter = nullptr;  // work-queue head of the persistent kernels
    int2 *d_tiles_sym = nullptr, *d_tiles_full = nullptr;  // K-GEMM tile lists, largest live area first
    int ntiles_sym = 0, ntiles_full = 0;
    // on-device fitting (f2): mirror ranks, this shard's metric rows, staging
    int* d_mpos = nullptr;
    double* d_metric = nullptr;  // [nq][apitch]
    bool have_metric = false;
    double* Dm = nullptr;        // [nmat][nbf][ldd] density operands of the fused first J sweep
    size_t dm_cap = 0;
    std::vector<char> fused;     // per density of the current build: first J sweep done inside the half transform
    // unfitted integrals of a group, double-buffered so the H2D of the next group (copy stream) runs under the
    // kernels of the current one; fit_raw_free[b]: the kernels that read buffer b have finished
    double *fit_raw[2] = {nullptr, nullptr}, *fit_t = nullptr;
    size_t fit_raw_cap[2] = {0, 0}, fit_t_cap = 0;
    cudaEvent_t fit_raw_free[2] = {nullptr, nullptr};
    unsigned fit_next = 0;
    // pre-gathered C^T of the half transform (gather_ct_kernel) and its per-row-block TMA maps, cached on
    // (buffer, rows, box rows): in an SCF nocc never changes, so the maps are encoded once
    double* Cg = nullptr;
    size_t cg_cap = 0;
    // a few map sets, each keyed on (buffer, rows, box rows): an open-shell build alternates between two occupied
    // counts (n_alpha != n_beta) and must not re-encode nbf maps -- and drain the stream -- twice per iteration
    struct CgMaps {
        CUtensorMap* d = nullptr;
        const double* ptr = nullptr;
        int rows = 0, box = 0;
        uint64_t used = 0;
    } cgmaps[4];
    uint64_t cg_clock = 0;
    bool screened = false;  // some row-block keeps fewer than nbf partners
    char* fit_meta[2] = {nullptr, nullptr};  // page-locked index tables of a group (see fit_group)
    size_t fit_meta_cap[2] = {0, 0};
    cudaEvent_t fit_meta_done[2] = {nullptr, nullptr};
    size_t* d_fit_dst_off = nullptr;
    size_t* d_fit_src_off = nullptr;
    int *d_fit_dst_ld = nullptr, *d_fit_mi = nullptr, *d_fit_j0 = nullptr, *d_fit_m = nullptr;
    size_t fit_cols_cap = 0, fit_nm_cap = 0;
    size_t tensor_doubles = 0;
    size_t* d_row_off = nullptr;
    size_t* d_row_off_unit = nullptr;  // the same per Q row (sum of ld up to m): layout of the pre-gathered C^T
    int *d_ldm = nullptr, *d_sp = nullptr, *d_ign = nullptr, *d_cols = nullptr;
    size_t* d_cols_off = nullptr;
    // work buffers (grown on demand, kept across calls: SURVEY.md Appendix B "allocate once per handle")
    double *in = nullptr, *out = nullptr, *Ctl = nullptr, *Ctr = nullptr, *dpart = nullptr, *dvec = nullptr;
    double *T1 = nullptr, *T2 = nullptr, *ws = nullptr;
    size_t in_cap = 0, out_cap = 0, ct_cap = 0, dpart_cap = 0, T_cap = 0, T2_cap = 0, ws_cap = 0;
    ncclComm_t comm = nullptr;
    PeerCtx peer;
    std::vector<Phase> phases;
    std::vector<cudaEvent_t> evpool;
    size_t evused = 0;
    uint64_t launches = 0;
    // half transforms skipped in the current build because C_left repeated (Task::same_left): sum of q rows, q rows x nocc
    double skipped_q = 0, skipped_qo = 0;
    I8Plan i8;        // residue planes / workspace of the INT8-tensor-core K GEMM (i8_kgemm.cuh)
    I8HalfPlan i8h;   // row scales, C planes and scratch arena of the INT8-tensor-core half transform (i8_half.cuh)
    int half_kind = 0;  // arm the last half transform of the current build took (0 DMMA, 1 INT8 residues)
    int kgemm_kind = 0, kgemm_moduli = 0;  // arm the last K GEMM of the current build took (0 DMMA, 1 INT8 residues)
    int j_reads = 0;  // passes over the tensor the J sweeps of the current build made (first sweeps + batched second sweeps)
    // sub-phases of the INT8 arms of the current build: (tag, event) in stream order (tags: i8_half.cuh / i8_kgemm.cuh `mark`)
    std::vector<std::pair<int, cudaEvent_t>> marks;
    double i8h_ops = 0, i8h_plane_bytes = 0, i8h_convert_bytes = 0, i8k_ops = 0;
    int i8h_nmod = 0, i8h_chunks = 0, i8h_cached = 0, i8h_resident_rows = 0;
    int i8k_ok_kdim = -1, i8k_ok_nop = 0, i8k_ok_nmod = 0;  // shape of the last K GEMM the residue arm completed (its buffers suffice)
};
}  // namespace

struct b200jk {
    std::vector<Shard> sh;
    int rank = 0, world = 1;  // rank mode (one shard per process) when world > 1 and sh.size()==1
    bool rank_mode = false;
    int reduce_kind = 0;      // cross-GPU sum: 0 none (one GPU), 1 fixed-order peer-memory kernel, 2 NCCL all-reduce
    size_t nbf = 0, naux = 0;
    bool have_layout = false;
    bool uploaded[3] = {false, false, false};
    std::vector<size_t> small_skips, big_skips, row_off_unit;  // row_off_unit: sum of ldm up to m (per q row)
    std::vector<int> sp, ign, ldm, cols;
    std::vector<int> mpos;                 // rank of m among the kept partners of n = cols[m][k] (mirror destination)
    std::vector<size_t> symm_big_skips;    // dfhelper.cc:413-416: offsets of the symmetric-packed (n >= m) blocks
    double ms_fit_gemm = 0, fit_flops = 0; // last b200jk_fit_rows: device time and flops of the metric contraction
    std::vector<size_t> cols_off;
    int max_sp = 0;
    uint64_t work_budget = 0;
    int kgemm_arm = 0, kgemm_nmod = 0;  // b200jk_set_kgemm: 0 automatic / 1 DMMA / 2 INT8 residues; moduli (0 = default)
    int half_arm = 0, half_nmod = 0;    // b200jk_set_half: the same for the half transform
    double* pin_in = nullptr;
    double* pin_out = nullptr;
    size_t pin_in_cap = 0, pin_out_cap = 0;
    std::string err;
    b200jk_stats stats;
    std::vector<std::pair<const char*, size_t>> pinned;  // caller ranges registered with b200jk_register_host
    // page-locked staging ring of the tensor producers (b200jk_upload_rows / b200jk_fit_rows): the caller's block is
    // copied here by a few threads and the call returns; the DMA and the kernels run behind it
    struct StageSlot {
        double* buf = nullptr;
        size_t cap = 0;                 // doubles
        std::vector<cudaEvent_t> done;  // per shard: the last DMA out of this slot
    } stage[2];
    unsigned stage_next = 0;
    double stage_wait_s = 0, stage_alloc_s = 0, stage_copy_s = 0;  // host time of the producers (B200JK_STAGE_TRACE)
    struct FitTiming {
        int shard;
        cudaEvent_t a, b;
        double flops;
    };
    std::vector<FitTiming> fit_pending;  // metric-GEMM event pairs not yet read back
    // DF-JK gradient intermediates (grad_host.inl), alive between b200jk_grad_begin and b200jk_grad_end
    struct Grad {
        struct Spin {
            double* C = nullptr;     // [nbf][o] occupied orbitals
            double* cfit = nullptr;  // [naux][o][op] fitted (A|ij)
            int o = 0, op = 0;
        } spin[2];
        double *Jm12 = nullptr, *d = nullptr, *V = nullptr, *rows = nullptr, *M1 = nullptr;
        size_t rows_cap = 0, M1_cap = 0;
        int nspin = 0;
        bool active = false;
    } grad;
};

namespace {
void grad_free(b200jk* h);  // grad_host.inl
// true if [p, p+bytes) lies inside a range the caller registered (page-locked): DMA can use it directly
bool is_pinned(const b200jk* h, const void* p, size_t bytes) {
    const char* c = (const char*)p;
    for (auto& r : h->pinned)
        if (c >= r.first && c + bytes <= r.first + r.second) return true;
    return false;
}
}  // namespace

namespace {

int fail(b200jk* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(h, e_ == cudaErrorMemoryAllocation ? B200JK_ERR_OOM : B200JK_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                       \
    } while (0)
#define NK(call)                                                                                             \
    do {                                                                                                     \
        int r_ = (call);                                                                                     \
        if (r_ != 0)                                                                                         \
            return fail(h, B200JK_ERR_NCCL, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// cudaMalloc that gives the engine's own elastic scratch back first: the residue arenas of the INT8 arms take what is
// free when they are sized, and a later request (T2 of a non-symmetric build, a second tensor, gradient buffers) must
// not fail because of them.  They are rebuilt (smaller) by the next build that wants them.
cudaError_t malloc_elastic(b200jk* h, void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaErrorMemoryAllocation || !h) return e;
    cudaGetLastError();
    int dev = -1;
    cudaGetDevice(&dev);
    bool freed = false;
    for (auto& s : h->sh) {
        if (s.dev != dev) continue;
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.i8h.arena) {
            s.i8h.release_arena();
            freed = true;
        }
        if (s.i8.planes[0] || s.i8.planes[1] || s.i8.ws) {
            s.i8.release();
            freed = true;
        }
    }
    return freed ? cudaMalloc(p, bytes) : e;
}

template <class T>
int upload_vec(b200jk* h, T** dptr, const std::vector<T>& v) {
    if (*dptr) cudaFree(*dptr);
    *dptr = nullptr;
    size_t n = std::max<size_t>(v.size(), 1);
    CK(malloc_elastic(h, (void**)dptr, n * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

int grow(b200jk* h, double** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CK(malloc_elastic(h, (void**)p, need * sizeof(double)));
    *cap = need;
    return 0;
}

cudaEvent_t get_event(Shard& s) {
    if (s.evused == s.evpool.size()) {
        cudaEvent_t e;
        cudaSetDevice(s.dev);  // events belong to the device that is current when they are created
        cudaEventCreate(&e);
        s.evpool.push_back(e);
    }
    return s.evpool[s.evused++];
}
// phase marker handed to the INT8 arms: one event per sub-phase boundary on the compute stream
void i8_mark(void* ctx, int tag) {
    Shard& s = *static_cast<Shard*>(ctx);
    cudaEvent_t e = get_event(s);
    cudaEventRecord(e, s.stream);
    s.marks.emplace_back(tag, e);
}
struct PhaseScope {
    Shard& s;
    Phase ph;
    PhaseScope(Shard& s_, int tag) : s(s_) {
        ph.tag = tag;
        ph.a = get_event(s);
        ph.b = get_event(s);
        cudaEventRecord(ph.a, s.stream);
    }
    ~PhaseScope() {
        cudaEventRecord(ph.b, s.stream);
        s.phases.push_back(ph);
    }
};

inline int round_up(int x, int a) { return (x + a - 1) / a * a; }

#include "peer_host.inl"

// Pick the split-K factor of the K GEMM: enough CTAs to fill 148 SMs in nearly whole waves while
// each split keeps >= 64 k-tiles (prologue/epilogue amortised) and the partial-tile workspace stays bounded.
void choose_split(int ntiles, int kdim, int nsm, int* nsplit, int* klen) {
    int nkt = (kdim + BK - 1) / BK;
    int max_s = std::max(1, nkt / 64);
    size_t ws_cap = (size_t)1 << 30;  // 1 GiB of partials
    max_s = (int)std::min<size_t>(max_s, std::max<size_t>(1, ws_cap / ((size_t)ntiles * BM * 128 * 8)));
    max_s = std::min(max_s, 512);
    int best = 1;
    double best_eff = 0;
    for (int s = 1; s <= max_s; s++) {
        long ctas = (long)ntiles * s;
        long waves = (ctas + nsm - 1) / nsm;
        double eff = (double)ctas / (double)(waves * nsm);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = s;
        }
        if (eff >= 0.97 && waves >= 4) {
            best = s;
            break;
        }
    }
    int kl = ((nkt + best - 1) / best) * BK;
    *klen = kl;
    *nsplit = (kdim + kl - 1) / kl;
}


// ---- TMA descriptors ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int load_encode(b200jk* h) {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess)
        return fail(h, B200JK_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = (EncodeTiledFn)fn;
    return 0;
}

// 2-D f64 tensor map: `inner` contiguous elements per row, `rows` rows of pitch_bytes; box = 16 x box_rows,
// 128-byte swizzle, out-of-bounds elements read as zero (ragged k tails / edge rows need no masking).
int make_map(b200jk* h, CUtensorMap* out, const double* base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
             uint32_t box_rows) {
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(h, B200JK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%llu rows=%llu pitch=%llu", (int)r,
                    (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)pitch_bytes);
    return 0;
}

bool use_legacy() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200JK_LEGACY");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

template <int NB>
int launch_half_ws(b200jk* h, Shard& s, const CUtensorMap& ctmap, const CUtensorMap& ctmap_last, const CUtensorMap& dmap,
                   const HalfWsParams& p) {
    static bool attr_set[64] = {false};
    constexpr size_t smem = ws_smem_bytes<NB>();
    if (!attr_set[s.dev]) {
        CK(cudaFuncSetAttribute(half_ws_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[s.dev] = true;
    }
    int grid = std::min(p.nitems, s.nsm);
    CK(cudaMemsetAsync(s.d_counter, 0, sizeof(int), s.stream));
    half_ws_kernel<NB><<<grid, WS_THREADS, smem, s.stream>>>(ctmap, ctmap_last, dmap, p);
    s.launches++;
    CK(cudaGetLastError());
    return 0;
}

// Density operand of the fused first J sweep (see HalfWsParams): nullptr = plain half transform.
struct FuseJ {
    const double* Dm;
    int ldd;
    double* dpart;
    int dstride;
    int gemm_col = 0;  // INT8 arm only: the sweep rides on the GEMM (one more column), not on the conversion (i8_half.cuh)
};

// orbital tiling of the half transform: nit tiles of iw columns (whole 8-column DMMA blocks), NB blocks per warp pair
void half_tiling(int o, int op, int* nit, int* iw, int* NB) {
    *nit = (o + 127) / 128;
    *iw = std::min(128, round_up((op + *nit - 1) / *nit, 8));
    *NB = (*iw + 15) / 16;
}
// the density row needs a free B-operand row in the last orbital tile
bool can_fuse_j(int o) {
    static int off = -1;
    if (off < 0) {
        const char* e = getenv("B200JK_NO_JFUSE");
        off = (e && e[0] == '1') ? 1 : 0;
    }
    if (off || use_legacy() || o <= 0) return false;
    int nit, iw, NB;
    half_tiling(o, round_up(o, 2), &nit, &iw, &NB);
    return o - (nit - 1) * iw < 16 * NB;
}

// Pre-gather C^T for the screened row-blocks (all-TMA K3)?  The packed copy costs one pass of 8*P*(o+1) bytes per
// transform, independent of how many Q rows the shard holds, and buys back the LSU gather (about 3.5 % of K3):
// worth it from ~1500 Q rows per GPU.  B200JK_CGATHER=1 / 0 forces it on / off.
bool want_cgather(const Shard& s, int qc) {
    static int mode = -2;
    if (mode == -2) {
        const char* e = getenv("B200JK_CGATHER");
        mode = !e || !*e ? -1 : (e[0] == '1' ? 1 : 0);
    }
    if (use_legacy() || !s.screened) return false;
    return mode == 1 || (mode == -1 && qc >= 1500);
}

int run_half_ws(b200jk* h, Shard& s, int which, const double* Ct, int ldc, int o, int op, int qbeg, int qc, double* T,
                const FuseJ* fj) {
    int nit, iw, NB;
    half_tiling(o, op, &nit, &iw, &NB);
    const CUtensorMap* cgmaps = nullptr;
    if (want_cgather(s, qc)) {
        const int R = o + 1;
        const size_t unit = h->row_off_unit[h->nbf];  // sum of ld(m)
        int rc = grow(h, &s.Cg, &s.cg_cap, unit * (size_t)R);
        if (rc) return rc;
        gather_ct_kernel<<<dim3((unsigned)h->nbf, (unsigned)((R + 7) / 8)), 256, 0, s.stream>>>(
            Ct, ldc, o, fj ? fj->Dm : nullptr, fj ? fj->ldd : 0, s.d_sp, s.d_ldm, s.d_row_off_unit, s.d_cols, s.d_cols_off, s.Cg);
        s.launches++;
        CK(cudaGetLastError());
        Shard::CgMaps* hit = nullptr;
        Shard::CgMaps* victim = &s.cgmaps[0];
        for (auto& c : s.cgmaps) {
            if (c.d && c.ptr == s.Cg && c.rows == R && c.box == 16 * NB) hit = &c;
            if (c.used < victim->used) victim = &c;
        }
        if (!hit) {
            hit = victim;
            std::vector<CUtensorMap> maps(h->nbf);
            for (size_t m = 0; m < h->nbf; m++)
                if ((rc = make_map(h, &maps[m], s.Cg + h->row_off_unit[m] * (size_t)R, (uint64_t)h->sp[m], (uint64_t)R,
                                   (uint64_t)h->ldm[m] * 8, (uint32_t)(16 * NB))))
                    return rc;
            if (!hit->d) CK(cudaMalloc((void**)&hit->d, h->nbf * sizeof(CUtensorMap)));
            // (pageable source: the copy is staged before the call returns, and it is stream-ordered before K3)
            CK(cudaMemcpyAsync(hit->d, maps.data(), h->nbf * sizeof(CUtensorMap), cudaMemcpyHostToDevice, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            hit->ptr = s.Cg;
            hit->rows = R;
            hit->box = 16 * NB;
        }
        hit->used = ++s.cg_clock;
        cgmaps = hit->d;
    }
    CUtensorMap ctmap, ctmap_last, dmap;
    int rc = make_map(h, &ctmap, Ct, h->nbf, (uint64_t)o, (uint64_t)ldc * 8, (uint32_t)(16 * NB));
    if (rc) return rc;
    ctmap_last = ctmap;
    dmap = ctmap;
    const int rd = o - (nit - 1) * iw;
    if (fj) {
        if ((rc = make_map(h, &ctmap_last, Ct, h->nbf, (uint64_t)o, (uint64_t)ldc * 8, (uint32_t)rd))) return rc;
        if ((rc = make_map(h, &dmap, fj->Dm, h->nbf, h->nbf, (uint64_t)fj->ldd * 8, 1))) return rc;
    }
    HalfWsParams p;
    p.fuse = fj ? 1 : 0;
    p.rd = rd;
    p.Dm = fj ? fj->Dm : nullptr;
    p.ldd = fj ? fj->ldd : 0;
    p.dpart = fj ? fj->dpart : nullptr;
    p.dstride = fj ? fj->dstride : 0;
    p.cgmaps = cgmaps;
    p.amaps = s.d_amaps[which];
    p.sp = s.d_sp;
    p.cols = s.d_cols;
    p.cols_off = s.d_cols_off;
    p.Ct = Ct;
    p.ldc = ldc;
    p.o = o;
    p.op = op;
    p.iw = iw;
    p.nit = nit;
    p.nqt = (qc + BM - 1) / BM;
    p.qbeg = qbeg;
    p.qc = qc;
    p.nbf = (int)h->nbf;
    p.nitems = nit * p.nqt * (int)h->nbf;
    p.counter = s.d_counter;
    p.T = T;
    switch (NB) {
        case 1: return launch_half_ws<1>(h, s, ctmap, ctmap_last, dmap, p);
        case 2: return launch_half_ws<2>(h, s, ctmap, ctmap_last, dmap, p);
        case 3: return launch_half_ws<3>(h, s, ctmap, ctmap_last, dmap, p);
        case 4: return launch_half_ws<4>(h, s, ctmap, ctmap_last, dmap, p);
        case 5: return launch_half_ws<5>(h, s, ctmap, ctmap_last, dmap, p);
        case 6: return launch_half_ws<6>(h, s, ctmap, ctmap_last, dmap, p);
        case 7: return launch_half_ws<7>(h, s, ctmap, ctmap_last, dmap, p);
        default: return launch_half_ws<8>(h, s, ctmap, ctmap_last, dmap, p);
    }
}

int run_kgemm_ws(b200jk* h, Shard& s, const double* T1, const double* T2, int kdim, bool symmetric, double* Kout) {
    static bool attr_set[64] = {false};
    constexpr size_t smem = ws_smem_bytes<8>();
    if (!attr_set[s.dev]) {
        CK(cudaFuncSetAttribute(kgemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[s.dev] = true;
    }
    const int ntiles = symmetric ? s.ntiles_sym : s.ntiles_full;
    const int2* tiles = symmetric ? s.d_tiles_sym : s.d_tiles_full;
    // Dynamic scheduling: aim at ~48 work items per SM so the tail is one small item (at 16 the last wave cost ~5 %),
    // while every item keeps >= 64 k-tiles (pipeline fill + partial-tile store amortised) and the partial-tile
    // workspace stays <= 1 GiB.
    const int nkt = (kdim + BK - 1) / BK;
    int nsplit = (48 * s.nsm + ntiles - 1) / ntiles;
    nsplit = std::min(nsplit, std::max(1, nkt / 64));
    nsplit = (int)std::min<size_t>((size_t)nsplit, std::max<size_t>(1, ((size_t)1 << 30) / ((size_t)ntiles * BM * 128 * 8)));
    nsplit = std::max(nsplit, 1);
    const int klen = ((nkt + nsplit - 1) / nsplit) * BK;
    nsplit = (kdim + klen - 1) / klen;
    size_t need = (size_t)nsplit * ntiles * BM * 128;
    int rc = grow(h, &s.ws, &s.ws_cap, need);
    if (rc) return rc;
    CUtensorMap m1, m2;
    if ((rc = make_map(h, &m1, T1, (uint64_t)kdim, h->nbf, (uint64_t)kdim * 8, BM))) return rc;
    if ((rc = make_map(h, &m2, T2, (uint64_t)kdim, h->nbf, (uint64_t)kdim * 8, 128))) return rc;
    KgemmWsParams p;
    p.nbf = (int)h->nbf;
    p.kdim = kdim;
    p.klen = klen;
    p.ntiles = ntiles;
    p.nitems = ntiles * nsplit;
    p.symmetric = symmetric ? 1 : 0;
    p.tiles = tiles;
    p.counter = s.d_counter;
    p.ws = s.ws;
    // the split-K reduction runs inside the kernel (per-tile arrival counters behind the queue head);
    // B200JK_KREDUCE=separate keeps the second pass for A/B measurement
    static int separate = -1;
    if (separate < 0) {
        const char* e = getenv("B200JK_KREDUCE");
        separate = (e && !strcmp(e, "separate")) ? 1 : 0;
    }
    static int tri = -1;
    if (tri < 0) {
        const char* e = getenv("B200JK_KTRI");
        tri = (e && e[0] == '0') ? 0 : 1;
    }
    p.tri = (tri && !separate) ? 1 : 0;  // (the separate reduce kernel reads partial tiles in the plain row order: same)
    p.arrive = separate ? nullptr : s.d_counter + 1;
    p.nsplit = nsplit;
    p.K = Kout;
    CK(cudaMemsetAsync(s.d_counter, 0, sizeof(int) * (1 + (size_t)s.ntiles_full * WS_CONSUMER_WARPS), s.stream));
    kgemm_ws_kernel<<<std::min(p.nitems, s.nsm), WS_THREADS, smem, s.stream>>>(m1, m2, p);
    s.launches++;
    CK(cudaGetLastError());
    if (separate) {
        kgemm_reduce_list_kernel<<<ntiles, 256, 0, s.stream>>>(s.ws, nsplit, ntiles, tiles, symmetric ? 1 : 0, p.nbf, Kout);
        s.launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

template <int NB>
int launch_half(b200jk* h, Shard& s, const HalfParams& p, dim3 grid) {
    static bool attr_set[64] = {false};
    constexpr size_t smem = gemm_smem_bytes<NB>();
    if (!attr_set[s.dev]) {
        CK(cudaFuncSetAttribute(half_transform_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[s.dev] = true;
    }
    half_transform_kernel<NB><<<grid, GEMM_THREADS, smem, s.stream>>>(p);
    s.launches++;
    CK(cudaGetLastError());
    return 0;
}

int kgemm_want_i8(const b200jk* h, int kdim);
int kgemm_moduli(const b200jk* h);

// Which arm the half transform takes.  B200JK_HALF=dmma|i8 (or b200jk_set_half) forces one.
int half_want_i8(const b200jk* h, const Shard& s, int o) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("B200JK_HALF");
        env = !e ? 0 : (!strcmp(e, "dmma") ? 1 : (!strcmp(e, "i8") ? 2 : 0));
    }
    const int arm = h->half_arm ? h->half_arm : env;
    if (arm == 1 || use_legacy() || o <= 0) return 0;
    if (arm == 2) return 1;
    // automatic: from 256 basis functions, 16 occupied orbitals and 2^27 tensor elements per shard on (smaller builds
    // are launch-bound and stay on the single DMMA kernel)
    return h->nbf >= 256 && o >= 16 && (double)s.nq * (double)h->small_skips[h->nbf] >= 134217728.0;
}
int half_moduli(const b200jk* h) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("B200JK_I8_HALF_MODULI");
        env = e ? atoi(e) : 0;
    }
    const int n = h->half_nmod ? h->half_nmod : (env ? env : 12);
    return std::max(I8_MINMOD, std::min(I8_MAXMOD, n));
}

// Size the scratch arena of the INT8 half transform for this build (largest nocc, Q chunk).  Returns -1 when not even one
// row-block fits beside everything else; *cacheable: the planes of the whole chunk fit at once, so they stay valid across
// builds (i8h_cached) -- and the first J sweep must then NOT ride on the conversion, or the first build (conversion runs)
// and the later ones (skipped) would sum d_Q in different orders.
int half_i8_cluster() {
    static int cl_env = -1;
    if (cl_env < 0) {
        const char* e = getenv("B200JK_I8_CLUSTER");
        cl_env = e ? atoi(e) : 1;
    }
    return cl_env;
}
int half_i8_arena(b200jk* h, Shard& s, int max_o, int qc, int op, bool two_operands, bool* cacheable) {
    const int nmod = half_moduli(h), cluster = half_i8_cluster();
    const int nbf = (int)h->nbf;
    size_t mn = 0, all = 0;
    i8h_arena_need(s.i8h, nmod, qc, max_o, cluster, &mn, &all);
    if (s.i8h.arena_cap < mn || (s.i8h.arena_cap < all && s.i8h.arena_cap < ((size_t)8 << 30))) {
        // (re)size the arena: everything in one chunk if it fits beside what the K GEMM's residue arm will ask for
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t avail = free_b + s.i8h.arena_cap;
        size_t k4 = 0;
        if (kgemm_want_i8(h, qc * op) && !s.i8.planes_cap[0])  // planes of one operand (two when C_right differs) + residues
            k4 = (size_t)kgemm_moduli(h) * nbf * ((size_t)qc * op + 128) * (two_operands ? 2 : 1) + ((size_t)4 << 30);
        const size_t reserve = (size_t)4 << 30;
        if (avail < k4 + reserve + mn) return -1;
        size_t want = std::min(all, avail - k4 - reserve);
        static long cap_mb = -1;  // B200JK_I8_ARENA_MB: upper bound of the arena (tests of the partly resident plan)
        if (cap_mb < 0) {
            const char* e = getenv("B200JK_I8_ARENA_MB");
            cap_mb = e ? atol(e) : 0;
        }
        if (cap_mb > 0) want = std::max(mn, std::min(want, (size_t)cap_mb << 20));
        if (want > s.i8h.arena_cap) {
            s.i8h.release_arena();
            if (cudaMalloc((void**)&s.i8h.arena, want) != cudaSuccess) {
                cudaGetLastError();
                s.i8h.arena = nullptr;
                return -1;
            }
            s.i8h.arena_cap = want;
        }
    }
    if (cacheable) *cacheable = s.i8h.arena_cap >= all;
    return 0;
}

// K3 on the INT8 tensor cores (i8_half.cuh).  Returns -1 when no scratch arena can be had (the caller then takes the
// DMMA arm).  max_o: the largest nocc of the build.
int run_half_i8(b200jk* h, Shard& s, int which, const double* Ct, int ldc, int o, int op, int max_o, int qbeg, int qc, double* T,
                const FuseJ* fj, bool two_operands) {
    const int nmod = half_moduli(h), cluster = half_i8_cluster();
    const int nbf = (int)h->nbf;
    if (half_i8_arena(h, s, max_o, qc, op, two_operands, nullptr) < 0) return -1;
    std::string err;
    I8HalfInfo info;
    I8HalfFuseJ f8 = {nullptr, 0, nullptr, 0, nullptr, nullptr};
    if (fj) {  // the first J sweep rides on the conversion (the DMMA arm carries it as an extra operand row instead)
        f8.Dm = fj->Dm;
        f8.ldd = fj->ldd;
        f8.dpart = fj->dpart;
        f8.dstride = fj->dstride;
        f8.gemm_col = fj->gemm_col;
    }
    const uint64_t l0 = s.i8h.launches;
    s.i8h.mark = i8_mark;
    s.i8h.mark_ctx = &s;
    int rc = i8_half_run(s.i8h, s.stream, s.nsm, s.tensor[which], which, s.d_row_off, s.d_ldm, s.d_sp, s.d_cols, s.d_cols_off, nbf, s.nq,
                         Ct, ldc, o, op, max_o, qbeg, qc, T, (size_t)qc * op, nmod, cluster, fj ? &f8 : nullptr, &info, &err);
    s.launches += s.i8h.launches - l0;
    if (rc == 3) return -1;
    if (rc) return fail(h, B200JK_ERR_CUDA, "INT8 half transform: %s", err.c_str());
    s.half_kind = 1;
    s.i8h_ops += info.mma_ops;
    s.i8h_plane_bytes += info.plane_bytes;
    s.i8h_convert_bytes += info.convert_bytes;
    s.i8h_nmod = info.nmod;
    s.i8h_chunks += info.nchunks;
    s.i8h_cached += info.cached;
    s.i8h_resident_rows = info.resident_rows;
    return 0;
}

// T[m][q][i] for q in the chunk [qbeg, qbeg+qc): one launch over (i-tiles, q-tiles, m).
int run_half(b200jk* h, Shard& s, int which, const double* Ct, int ldc, int o, int op, int qbeg, int qc, double* T,
             const FuseJ* fj = nullptr, int max_o = 0, bool two_operands = false) {
    if (half_want_i8(h, s, o)) {
        int rc = run_half_i8(h, s, which, Ct, ldc, o, op, std::max(max_o, o), qbeg, qc, T, fj, two_operands);
        if (rc >= 0) return rc;
    }
    s.half_kind = 0;
    if (!use_legacy()) return run_half_ws(h, s, which, Ct, ldc, o, op, qbeg, qc, T, fj);
    const double* tensor = s.tensor[which];
    int nit = (o + 127) / 128;
    int iw = round_up((op + nit - 1) / nit, 2);
    int NB = (iw + 15) / 16;
    HalfParams p;
    p.tensor = tensor;
    p.row_off = s.d_row_off;
    p.ldm = s.d_ldm;
    p.sp = s.d_sp;
    p.cols = s.d_cols;
    p.cols_off = s.d_cols_off;
    p.Ct = Ct;
    p.ldc = ldc;
    p.o = o;
    p.op = op;
    p.iw = iw;
    p.qbeg = qbeg;
    p.qc = qc;
    p.nbf = (int)h->nbf;
    p.T = T;
    dim3 grid(nit, (qc + BM - 1) / BM, (unsigned)h->nbf);
    switch (NB) {
        case 1: return launch_half<1>(h, s, p, grid);
        case 2: return launch_half<2>(h, s, p, grid);
        case 3: return launch_half<3>(h, s, p, grid);
        case 4: return launch_half<4>(h, s, p, grid);
        case 5: return launch_half<5>(h, s, p, grid);
        case 6: return launch_half<6>(h, s, p, grid);
        case 7: return launch_half<7>(h, s, p, grid);
        default: return launch_half<8>(h, s, p, grid);
    }
}

// Which arm the K GEMM takes.  B200JK_KGEMM=dmma|i8 (or b200jk_set_kgemm) forces one; automatic = the INT8 residue
// arm from 128 basis functions and 2^24 elements of T on (below that its five launches cost more than the DMMA kernel).
int kgemm_want_i8(const b200jk* h, int kdim) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("B200JK_KGEMM");
        env = !e ? 0 : (!strcmp(e, "dmma") ? 1 : (!strcmp(e, "i8") ? 2 : 0));
    }
    const int arm = h->kgemm_arm ? h->kgemm_arm : env;
    if (arm == 1 || use_legacy()) return 0;
    if (arm == 2) return 1;
    return h->nbf >= 128 && kdim >= 4096 && (size_t)h->nbf * (size_t)kdim >= ((size_t)1 << 24);
}
int kgemm_moduli(const b200jk* h) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("B200JK_I8_MODULI");
        env = e ? atoi(e) : 0;
    }
    const int n = h->kgemm_nmod ? h->kgemm_nmod : (env ? env : I8_MAXMOD);
    return std::max(I8_MINMOD, std::min(I8_MAXMOD, n));
}

// K4 on the INT8 tensor cores (i8_kgemm.cuh).  Returns -1 when the residue planes do not fit in what is free right
// now (the caller then takes the DMMA arm: slower, never wrong).
int run_kgemm_i8(b200jk* h, Shard& s, const double* T1, const double* T2, int kdim, bool symmetric, double* Kout) {
    static int klen_env = -1;
    if (klen_env < 0) {
        const char* e = getenv("B200JK_I8_KLEN");
        klen_env = e ? atoi(e) : 0;
    }
    const int nmod = kgemm_moduli(h);
    const int klen = klen_env > 0 ? klen_env : 8192;
    const int nbf = (int)h->nbf, nop = symmetric ? 1 : 2;
    const size_t ntile_est = (size_t)((nbf + I8_TM - 1) / I8_TM) * ((nbf + I8_TN - 1) / I8_TN + 1);
    const size_t ws_est = ((size_t)(kdim + klen - 1) / klen) * nmod * ntile_est * I8_TILE_BYTES;
    const size_t plane_need = (size_t)nmod * nbf * (((size_t)kdim + 127) / 128 * 128);  // per operand, one pass over k
    size_t budget = plane_need * nop;
    // (ws_est is an upper estimate -- the symmetric tile list is about half of it -- so a workspace grown to what a build of
    // this very shape needed counts as sufficient: otherwise every build would ask the driver for the free memory)
    const bool same_shape = s.i8k_ok_kdim == kdim && s.i8k_ok_nop == nop && s.i8k_ok_nmod == nmod && s.i8.ws_cap > 0;
    const bool have = s.i8.planes_cap[0] >= plane_need && (nop == 1 || s.i8.planes_cap[1] >= plane_need) &&
                      (s.i8.ws_cap >= ws_est || same_shape);
    if (!have) {
        // (cudaMemGetInfo only when something must grow: in the steady state of an SCF it would be a driver call per build
        // that now and then waits tens of milliseconds behind the copies of the other stream)
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const size_t held = s.i8.planes_cap[0] + s.i8.planes_cap[1] + s.i8.ws_cap;
        const size_t reserve = (size_t)1 << 30;
        const size_t avail = free_b + held;
        const size_t min_planes = (size_t)nmod * nbf * nop * (size_t)std::min(kdim, 4 * klen);
        if (avail < ws_est + reserve + min_planes) return -1;
        budget = std::min(budget, avail - ws_est - reserve);
    }
    std::string err;
    I8RunInfo info;
    const uint64_t l0 = s.i8.launches;
    s.i8.mark = i8_mark;
    s.i8.mark_ctx = &s;
    int rc = i8_kgemm_run(s.i8, (I8EncodeFn)g_encode, s.stream, s.nsm, T1, T2, (size_t)kdim, nbf, kdim, symmetric, Kout, nbf, nmod,
                          klen, budget, &info, &err);
    s.launches += s.i8.launches - l0;
    if (rc == 3) {  // lost a race for the memory: release and let the DMMA arm run
        cudaGetLastError();
        s.i8.release();
        return -1;
    }
    if (rc) return fail(h, B200JK_ERR_CUDA, "INT8 K GEMM: %s", err.c_str());
    s.kgemm_kind = 1;
    s.kgemm_moduli = info.nmod;
    s.i8k_ok_kdim = kdim;
    s.i8k_ok_nop = nop;
    s.i8k_ok_nmod = nmod;
    s.i8k_ops += info.mma_ops;
    return 0;
}

int run_kgemm(b200jk* h, Shard& s, const double* T1, const double* T2, int kdim, bool symmetric, double* Kout) {
    if (kgemm_want_i8(h, kdim)) {
        int rc = run_kgemm_i8(h, s, T1, T2, kdim, symmetric, Kout);
        if (rc >= 0) return rc;
    }
    s.kgemm_kind = 0;
    s.kgemm_moduli = 0;
    if (!use_legacy()) return run_kgemm_ws(h, s, T1, T2, kdim, symmetric, Kout);
    static bool attr_set[64] = {false};
    constexpr size_t smem = gemm_smem_bytes<8>();
    if (!attr_set[s.dev]) {
        CK(cudaFuncSetAttribute(kgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[s.dev] = true;
    }
    int n1d = ((int)h->nbf + BM - 1) / BM;
    int ntiles = symmetric ? n1d * (n1d + 1) / 2 : n1d * n1d;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s.dev);
    int nsplit, klen;
    choose_split(ntiles, kdim, nsm, &nsplit, &klen);
    size_t need = (size_t)nsplit * ntiles * BM * 128;
    int rc = grow(h, &s.ws, &s.ws_cap, need);
    if (rc) return rc;
    KgemmParams p;
    p.T1 = T1;
    p.T2 = T2;
    p.nbf = (int)h->nbf;
    p.kdim = kdim;
    p.klen = klen;
    p.ntile1d = n1d;
    p.symmetric = symmetric ? 1 : 0;
    p.ws = s.ws;
    kgemm_kernel<<<dim3(ntiles, nsplit), GEMM_THREADS, smem, s.stream>>>(p);
    s.launches++;
    CK(cudaGetLastError());
    kgemm_reduce_kernel<<<ntiles, 256, 0, s.stream>>>(s.ws, nsplit, ntiles, n1d, p.symmetric, p.nbf, Kout);
    s.launches++;
    CK(cudaGetLastError());
    return 0;
}

// J sweeps of one shard: per density the first sweep (unless the fused half transform already produced d_part) and
// the fixed-order reduction to d[q]; then the second sweep for up to J_MAX_ND densities per pass over the tensor.
template <int ND>
int launch_j_mn(b200jk* h, Shard& s, const JParams& p, const JBatch& bt) {
    static bool attr_set[64] = {false};
    if (!attr_set[s.dev]) {
        CK(cudaFuncSetAttribute(j_mn_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[s.dev] = true;
    }
    j_mn_kernel<ND><<<dim3((h->max_sp + 127) / 128, (unsigned)h->nbf), J_THREADS, (size_t)ND * s.nq * sizeof(double), s.stream>>>(p, bt);
    s.launches++;
    CK(cudaGetLastError());
    return 0;
}

int run_j_all(b200jk* h, Shard& s, int nmat, const double* const* D, bool symmetric, double* Jout0, size_t n2) {
    if (s.nq == 0) return 0;  // an empty Q shard (naux < number of GPUs) contributes the zeros of the memset
    JParams p;
    p.tensor = s.tensor[B200JK_TENSOR_PPQ];
    p.row_off = s.d_row_off;
    p.ldm = s.d_ldm;
    p.sp = s.d_sp;
    p.ign = s.d_ign;
    p.cols = s.d_cols;
    p.cols_off = s.d_cols_off;
    p.nbf = (int)h->nbf;
    p.nq = s.nq;
    p.symmetric = symmetric ? 1 : 0;
    p.J = nullptr;
    static bool attr_set[64] = {false};
    if (!attr_set[s.dev]) {
        CK(cudaFuncSetAttribute(j_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[s.dev] = true;
    }
    const size_t sm1 = (size_t)(h->max_sp + 2) * sizeof(double);
    const size_t smq = (size_t)s.nq * sizeof(double);
    if (sm1 > 200 * 1024 || smq > 200 * 1024)
        return fail(h, B200JK_ERR_INVALID, "J kernels: nbf %d / shard naux %d exceed the shared-memory staging limit",
                    h->max_sp, s.nq);
    s.j_reads = 0;
    for (int i = 0; i < nmat; i++) {
        p.D = D[i];
        p.dpart = s.dpart + (size_t)i * h->nbf * (size_t)s.nq;
        p.d = s.dvec + (size_t)i * s.nq;
        if (!s.fused[i]) {
            j_dq_kernel<<<dim3((s.nq + J1_ROWS - 1) / J1_ROWS, (unsigned)h->nbf), J_THREADS, sm1, s.stream>>>(p);
            s.j_reads++;
            s.launches++;
            CK(cudaGetLastError());
        }
        j_dq_reduce_kernel<<<(s.nq + JR_Q - 1) / JR_Q, JR_Q * JR_M, 0, s.stream>>>(p.dpart, p.nbf, s.nq, p.d);
        s.launches++;
        CK(cudaGetLastError());
    }
    const int nd_max = (int)std::max<size_t>(1, std::min<size_t>(J_MAX_ND, (200 * 1024) / smq));
    for (int i0 = 0; i0 < nmat; i0 += nd_max) {
        const int nd = std::min(nd_max, nmat - i0);
        JBatch bt;
        for (int k = 0; k < J_MAX_ND; k++) {
            bt.d[k] = s.dvec + (size_t)(i0 + std::min(k, nd - 1)) * s.nq;
            bt.J[k] = Jout0 + (size_t)(i0 + std::min(k, nd - 1)) * n2;
        }
        int rc;
        switch (nd) {
            case 1: rc = launch_j_mn<1>(h, s, p, bt); break;
            case 2: rc = launch_j_mn<2>(h, s, p, bt); break;
            case 3: rc = launch_j_mn<3>(h, s, p, bt); break;
            default: rc = launch_j_mn<4>(h, s, p, bt); break;
        }
        if (rc) return rc;
        s.j_reads++;
    }
    return 0;
}

int run_transpose(b200jk* h, Shard& s, const double* C, int o, double* Ct, int ldc, int orows) {
    dim3 grid((ldc + 31) / 32, (orows + 31) / 32);
    transpose_c_kernel<<<grid, dim3(32, 8), 0, s.stream>>>(C, (int)h->nbf, o, Ct, ldc, orows);
    s.launches++;
    CK(cudaGetLastError());
    return 0;
}

struct Task {
    int nmat;
    const int* nocc;
    bool lr, do_J, do_K, do_wK;
    int max_o;
    size_t n2;
    int nprod;  // tasked products
    // same_left[i]: C_left[i] is the matrix C_left[i-1] again (same nocc, same values).  The reference's response
    // callers pass ONE occupied block as every C_left (twoel_Hx: `Cl.push_back(Co)`, libscf_solver/rhf.cc:466-484,
    // uhf.cc, TDSCF) and DFHelper re-transforms it for every right-hand side (dfhelper.cc:3364); here the first half
    // transform is kept and only T2 = transform(C_right[i]) is recomputed.
    std::vector<char> same_left;
};

// Grow the per-shard work buffers for this task and pick the Q chunk of the K build.
int ensure_work(b200jk* h, Shard& s, const Task& t, int* qc_out) {
    CK(cudaSetDevice(s.dev));
    int rc;
    size_t N = h->nbf;
    {
        size_t need = (size_t)t.nmat * t.nprod * t.n2;
        const bool multi = h->sh.size() > 1 || (h->rank_mode && h->world > 1);
        if (multi) need = std::max(need, kPeerBlockBytes / sizeof(double));  // its own IPC-exportable block
        // rank mode: the peers map this buffer -- they let go of it (collectively) before it moves
        if (need > s.out_cap && (rc = peer_close_windows(h))) return rc;
        if ((rc = grow(h, &s.out, &s.out_cap, need))) return rc;
    }
    if (t.do_J) {
        // one d_part (and one d) per density: the fused half transforms of all densities run before the J sweeps
        if ((rc = grow(h, &s.dpart, &s.dpart_cap, (size_t)t.nmat * (N + 1) * (size_t)s.nq))) return rc;
        s.dvec = s.dpart + (size_t)t.nmat * N * (size_t)s.nq;
        if (t.do_K && (rc = grow(h, &s.Dm, &s.dm_cap, (size_t)t.nmat * N * (size_t)round_up((int)N, 2)))) return rc;
    }
    s.fused.assign(t.nmat, 0);
    *qc_out = 0;
    if ((t.do_K || t.do_wK) && t.max_o > 0) {
        int op = round_up(t.max_o, 2);
        int ldc = round_up((int)N, 4);
        size_t ct = (size_t)round_up(op, 32) * ldc;
        if (ct > s.ct_cap) {
            if (s.Ctl) CK(cudaFree(s.Ctl));
            if (s.Ctr) CK(cudaFree(s.Ctr));
            s.Ctl = s.Ctr = nullptr;
            s.ct_cap = 0;
            CK(malloc_elastic(h, (void**)&s.Ctl, ct * sizeof(double)));
            CK(malloc_elastic(h, (void**)&s.Ctr, ct * sizeof(double)));
            s.ct_cap = ct;
        }
        // the pre-gathered C^T (see want_cgather) is claimed before the T buffers size themselves from what is left
        if (want_cgather(s, s.nq) && (rc = grow(h, &s.Cg, &s.cg_cap, h->row_off_unit[N] * (size_t)(t.max_o + 1)))) return rc;
        bool two = !t.lr || t.do_wK;
        size_t per_q = N * (size_t)op;  // doubles of T per q row
        int qc = s.nq;
        size_t have = std::min(s.T_cap, two ? s.T2_cap : s.T_cap);
        if (per_q * qc > have) {
            // need to (re)allocate: budget = user limit or what is free now (+ what we would release)
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            // the INT8 scratch is elastic: it is given back (malloc_elastic) when T needs the room
            size_t reclaim = (s.T_cap + s.T2_cap) * sizeof(double) + s.i8h.arena_cap + s.i8.planes_cap[0] + s.i8.planes_cap[1] + s.i8.ws_cap;
            size_t budget = free_b + reclaim;
            size_t reserve = ((size_t)1 << 30) + ((size_t)1 << 29);  // split-K partials + slack
            budget = budget > reserve ? budget - reserve : 0;
            if (h->work_budget && h->work_budget < budget) budget = h->work_budget;
            size_t nT = two ? 2 : 1;
            size_t max_q = budget / (per_q * sizeof(double) * nT);
            if (max_q < (size_t)qc) {
                qc = (int)(max_q >= 128 ? (max_q / 128) * 128 : max_q);
                if (qc < 1)
                    return fail(h, B200JK_ERR_OOM, "not enough HBM for one Q row of the half-transformed tensor (%zu B)",
                                per_q * 8 * nT);
            }
            if (s.T1) CK(cudaFree(s.T1));
            if (s.T2) CK(cudaFree(s.T2));
            s.T1 = s.T2 = nullptr;
            s.T_cap = s.T2_cap = 0;
            CK(malloc_elastic(h, (void**)&s.T1, per_q * qc * sizeof(double)));
            s.T_cap = per_q * qc;
            if (two) {
                CK(malloc_elastic(h, (void**)&s.T2, per_q * qc * sizeof(double)));
                s.T2_cap = per_q * qc;
            }
        } else {
            qc = (int)std::min<size_t>(s.nq, have / per_q);
        }
        *qc_out = qc;
    }
    return 0;
}

// Device pointers of the three output groups inside s.out: [J x nmat | K x nmat | wK x nmat] (tasked ones only).
struct OutLayout {
    size_t offJ, offK, offW, countJ, countKW, total;
};
OutLayout out_layout(const Task& t) {
    OutLayout o;
    o.offJ = 0;
    o.countJ = (size_t)(t.do_J ? t.nmat : 0) * t.n2;
    o.offK = o.countJ;
    o.offW = o.offK + (size_t)(t.do_K ? t.nmat : 0) * t.n2;
    o.countKW = (size_t)((t.do_K ? t.nmat : 0) + (t.do_wK ? t.nmat : 0)) * t.n2;
    o.total = o.countJ + o.countKW;
    return o;
}

// The J sweeps of one build on one shard (K1, K2 per density).
int run_device_J(b200jk* h, Shard& s, const Task& t, const double* const* dD) {
    if (!t.do_J) return 0;
    CK(cudaSetDevice(s.dev));
    const OutLayout ol = out_layout(t);
    CK(cudaMemsetAsync(s.out + ol.offJ, 0, ol.countJ * sizeof(double), s.stream));
    PhaseScope ps(s, 0);
    return run_j_all(h, s, t.nmat, dD, t.lr, s.out + ol.offJ, t.n2);
}

// The K and wK builds of one shard (K3, K4 per density and Q chunk).  When J is tasked too (dD != nullptr) the first
// half transform of each density also produces d_part, the first J sweep (see HalfWsParams).
int run_device_K(b200jk* h, Shard& s, const Task& t, const double* const* dCl, const double* const* dCr,
                 const double* const* dD, int qc) {
    s.skipped_q = s.skipped_qo = 0;
    if (!t.do_K && !t.do_wK) return 0;
    CK(cudaSetDevice(s.dev));
    const size_t N = h->nbf, n2 = t.n2;
    const OutLayout ol = out_layout(t);
    double* outK = s.out + ol.offK;
    double* outW = s.out + ol.offW;
    CK(cudaMemsetAsync(s.out + ol.offK, 0, ol.countKW * sizeof(double), s.stream));
    int rc;
    const int ldc = round_up((int)N, 4);
    for (int pass = 0; pass < 2; pass++) {
        bool wk = pass == 1;
        if (wk ? !t.do_wK : !t.do_K) continue;
        const int tenL = wk ? B200JK_TENSOR_M1PPQ : B200JK_TENSOR_PPQ;
        const int tenR = wk ? B200JK_TENSOR_WPPQ : B200JK_TENSOR_PPQ;
        int t1_of = -1;  // matrix whose first half transform s.T1 holds in full (single Q chunk only)
        for (int i = 0; i < t.nmat; i++) {
            int o = t.nocc[i];
            if (!o) continue;  // dfhelper.cc:3354-3357
            int op = round_up(o, 2);
            bool one_T = t.lr && !wk;  // T2 = T1 (:3367-3368)
            // C_left repeated (response right-hand sides): T1 of the previous matrix is this matrix's T1
            const bool reuse_T1 = !one_T && t.same_left[i] && t1_of == i - 1 && qc >= s.nq;
            {
                PhaseScope ps(s, 1);
                if (!reuse_T1 && (rc = run_transpose(h, s, dCl[i], o, s.Ctl, ldc, op))) return rc;
                if (!one_T && (rc = run_transpose(h, s, t.lr ? dCl[i] : dCr[i], o, s.Ctr, ldc, op))) return rc;
            }
            double* Kout = (wk ? outW : outK) + i * n2;
            FuseJ fj_store, *fj = nullptr;
            // the density row can ride on either transform of the Ppq tensor: on T1 normally, on T2 when T1 is reused
            // (INT8 arm whose planes stay resident across builds: the conversion the sweep would ride on is skipped from the
            // second build on, so the sweep never rides on it -- J then takes the same kernels in every build)
            bool planes_stay = false;
            if (!wk && half_want_i8(h, s, o) && half_i8_arena(h, s, t.max_o, std::min(qc, s.nq), op, !one_T, &planes_stay) < 0)
                planes_stay = false;
            // ... unless it can ride on the GEMM itself as one more column (exact in the integers: the same bits whether the
            // planes were converted in this build or are cached); B200JK_I8_JCOL=0 keeps the separate first sweep
            static int jcol = -1;
            if (jcol < 0) {
                const char* e = getenv("B200JK_I8_JCOL");
                jcol = (e && e[0] == '0') ? 0 : 1;
            }
            const bool i8_arm = !wk && half_want_i8(h, s, o) && s.i8h.arena;
            const bool ride_gemm = i8_arm && jcol && i8h_can_fuse_col(s.i8h, o, half_i8_cluster());
            // (part of the planes stays resident even when not all of them fit, unless B200JK_I8_RESIDENT=0: riding on the
            // conversion is then the A/B arm only)
            static int resident = -1;
            if (resident < 0) {
                const char* e = getenv("B200JK_I8_RESIDENT");
                resident = (e && e[0] == '0') ? 0 : 1;
            }
            const bool ride_convert = !i8_arm || (!planes_stay && !resident);
            if (!wk && t.do_J && dD && dD[i] && can_fuse_j(o) && (ride_gemm || ride_convert)) {
                const int ldd = round_up((int)N, 2);
                fj_store.Dm = s.Dm + (size_t)i * N * ldd;
                fj_store.ldd = ldd;
                fj_store.dpart = s.dpart + (size_t)i * N * (size_t)s.nq;
                fj_store.dstride = s.nq;
                fj_store.gemm_col = ride_gemm ? 1 : 0;
                fj = &fj_store;
                PhaseScope ps(s, 0);
                j_prep_dm_kernel<<<dim3((ldd + 127) / 128, (unsigned)N), 128, 0, s.stream>>>(dD[i], (int)N, ldd, t.lr ? 1 : 0,
                                                                                          s.Dm + (size_t)i * N * ldd);
                s.launches++;
                CK(cudaGetLastError());
                s.fused[i] = 1;
            }
            for (int qb = 0; qb < s.nq; qb += qc) {
                int nqc = std::min(qc, s.nq - qb);
                {
                    PhaseScope ps(s, 1);
                    if (reuse_T1) {
                        s.skipped_q += nqc;
                        s.skipped_qo += (double)nqc * o;
                    } else if ((rc = run_half(h, s, tenL, s.Ctl, ldc, o, op, qb, nqc, s.T1, fj, t.max_o, !one_T))) {
                        return rc;
                    }
                    FuseJ* fj2 = reuse_T1 ? fj : nullptr;
                    if (!one_T && (rc = run_half(h, s, tenR, s.Ctr, ldc, o, op, qb, nqc, s.T2, fj2, t.max_o, true))) return rc;
                }
                {
                    PhaseScope ps(s, 2);
                    if ((rc = run_kgemm(h, s, s.T1, one_T ? s.T1 : s.T2, nqc * op, one_T, Kout))) return rc;
                }
            }
            t1_of = (qc >= s.nq) ? i : -1;
            if (wk && t.lr) {
                hermitivitize_kernel<<<dim3(((unsigned)N + 127) / 128, (unsigned)N), 128, 0, s.stream>>>(Kout, (int)N);
                s.launches++;
                CK(cudaGetLastError());
            }
        }
    }
    return 0;
}

void account_work(b200jk* h, const Task& t) {
    // algorithmic work of this handle's shards
    b200jk_stats& st = h->stats;
    size_t N = h->nbf;
    double P = (double)h->small_skips[N];
    double Ptri = 0;
    for (size_t m = 0; m < N; m++) Ptri += h->sp[m] - h->ign[m];
    double Aloc = 0;
    for (auto& s : h->sh) Aloc += s.nq;
    st.j_bytes = st.half_flops = st.half_bytes = st.kgemm_flops = 0;
    // J traffic = the passes over the tensor the sweeps actually make: a fused first sweep rides on the half
    // transform's read, and one second sweep serves up to J_MAX_ND densities (never credit bytes that were not read)
    if (t.do_J && !h->sh.empty()) st.j_bytes = (double)h->sh[0].j_reads * 8.0 * Aloc * (t.lr ? Ptri : P);
    for (int i = 0; i < t.nmat; i++) {
        double o = t.nocc[i];
        if (o == 0) continue;
        if (t.do_K) {
            int ntr = t.lr ? 1 : 2;
            st.half_flops += ntr * 2.0 * Aloc * P * o;
            st.half_bytes += ntr * (8.0 * Aloc * P + 8.0 * N * Aloc * o);
            st.kgemm_flops += t.lr ? (double)N * (N + 1) * Aloc * o : 2.0 * N * N * Aloc * o;
        }
        if (t.do_wK) {
            st.half_flops += 2 * 2.0 * Aloc * P * o;
            st.half_bytes += 2 * (8.0 * Aloc * P + 8.0 * N * Aloc * o);
            st.kgemm_flops += 2.0 * N * N * Aloc * o;
        }
    }
    // transforms not executed because C_left repeated are not credited (never count un-executed flops)
    for (auto& s : h->sh) {
        st.half_flops -= 2.0 * P * s.skipped_qo;
        st.half_bytes -= 8.0 * P * s.skipped_q + 8.0 * N * s.skipped_qo;
    }
}

int collect_stats(b200jk* h) {
    b200jk_stats& st = h->stats;
    st.ms_total = st.ms_j = st.ms_half = st.ms_kgemm = st.ms_allreduce = st.ms_h2d = st.ms_d2h = 0;
    st.launches = 0;
    for (int i = 0; i < 4; i++) st.ms_half_i8[i] = 0;
    for (int i = 0; i < 3; i++) st.ms_kgemm_i8[i] = 0;
    st.half_i8_ops = st.half_i8_plane_bytes = st.half_i8_convert_bytes = st.kgemm_i8_ops = 0;
    st.half_moduli = st.half_i8_chunks = st.half_i8_cached = st.half_i8_resident_rows = 0;
    for (auto& s : h->sh) {
        // sub-phases of the INT8 arms: the time between two consecutive marks belongs to the later one's tag
        double sub_h[4] = {0, 0, 0, 0}, sub_k[3] = {0, 0, 0};
        for (size_t i = 1; i < s.marks.size(); i++) {
            const int tag = s.marks[i].first;
            if (tag % 10 == 0) continue;  // a chunk / pass starts: what lies before it is not this arm's
            float ms = 0;
            cudaEventElapsedTime(&ms, s.marks[i - 1].second, s.marks[i].second);
            if (tag >= 11 && tag <= 14) sub_h[tag - 11] += ms;
            if (tag >= 21 && tag <= 23) sub_k[tag - 21] += ms;
        }
        for (int i = 0; i < 4; i++) st.ms_half_i8[i] = std::max(st.ms_half_i8[i], sub_h[i]);
        for (int i = 0; i < 3; i++) st.ms_kgemm_i8[i] = std::max(st.ms_kgemm_i8[i], sub_k[i]);
        st.half_i8_ops += s.i8h_ops;
        st.half_i8_plane_bytes += s.i8h_plane_bytes;
        st.half_i8_convert_bytes += s.i8h_convert_bytes;
        st.kgemm_i8_ops += s.i8k_ops;
        st.half_moduli = std::max(st.half_moduli, s.i8h_nmod);
        st.half_i8_chunks += s.i8h_chunks;
        st.half_i8_cached += s.i8h_cached;
        st.half_i8_resident_rows = std::max(st.half_i8_resident_rows, s.i8h_resident_rows);
        double acc[7] = {0, 0, 0, 0, 0, 0, 0};
        for (auto& ph : s.phases) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ph.a, ph.b);
            acc[ph.tag] += ms;
        }
        st.ms_j = std::max(st.ms_j, acc[0]);
        st.ms_half = std::max(st.ms_half, acc[1]);
        st.ms_kgemm = std::max(st.ms_kgemm, acc[2]);
        st.ms_allreduce = std::max(st.ms_allreduce, acc[3]);
        st.ms_h2d = std::max(st.ms_h2d, acc[4]);
        st.ms_d2h = std::max(st.ms_d2h, acc[5]);
        st.ms_total = std::max(st.ms_total, acc[6]);
        st.launches += s.launches;
    }
    return 0;
}

// memcpy split over a few threads: the staging copies of D / J / K (tens of MB) otherwise cost more than the PCIe hop.
void par_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t chunk = (size_t)4 << 20;
    if (bytes < 2 * chunk) {
        memcpy(dst, src, bytes);
        return;
    }
    int nt = (int)std::min<size_t>(bytes >= ((size_t)64 << 20) ? 8 : 4, bytes / chunk);
    std::vector<std::thread> th;
    size_t per = (bytes / nt + 63) & ~(size_t)63;
    for (int i = 1; i < nt; i++) {
        size_t a = per * i, b = std::min(bytes, per * (i + 1));
        if (a < b) th.emplace_back([=] { memcpy((char*)dst + a, (const char*)src + a, b - a); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto& x : th) x.join();
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// B200JK_STAGE_TRACE=1: one stderr line per producer call with where the host time went
bool stage_trace() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("B200JK_STAGE_TRACE");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

// Next slot of the staging ring, free (every DMA issued out of it has finished) and at least `doubles` large.
// Page-locking host memory is slow (measured 2.2 GB/s for cudaHostAlloc on the B200 box), so a slot is allocated
// once at the producers' group size (`reserve`) and only ever grows for a single row-block larger than that.
int stage_acquire(b200jk* h, size_t doubles, size_t reserve, b200jk::StageSlot** out) {
    b200jk::StageSlot& sl = h->stage[h->stage_next++ & 1];
    if (sl.done.size() != h->sh.size()) sl.done.assign(h->sh.size(), nullptr);
    double t0 = now_s();
    for (size_t si = 0; si < h->sh.size(); si++) {
        CK(cudaSetDevice(h->sh[si].dev));
        if (!sl.done[si])
            CK(cudaEventCreateWithFlags(&sl.done[si], cudaEventDisableTiming));
        else
            CK(cudaEventSynchronize(sl.done[si]));
    }
    h->stage_wait_s += now_s() - t0;
    if (doubles > sl.cap) {
        t0 = now_s();
        if (sl.buf) CK(cudaFreeHost(sl.buf));
        sl.buf = nullptr;
        sl.cap = 0;
        const size_t want = std::max(doubles, reserve);
        CK(cudaHostAlloc((void**)&sl.buf, want * sizeof(double), cudaHostAllocPortable));
        sl.cap = want;
        h->stage_alloc_s += now_s() - t0;
    }
    *out = &sl;
    return 0;
}

// Group size of the tensor producers; B200JK_STAGE_BYTES overrides it (tests drive many groups through the ring).
size_t stage_budget(size_t dflt) {
    const char* e = getenv("B200JK_STAGE_BYTES");
    if (e && *e) {
        unsigned long long v = strtoull(e, nullptr, 10);
        if (v) return (size_t)v;
    }
    return dflt;
}

// Everything the tensor producers queued has finished (streams of every shard drained).
int producers_sync(b200jk* h) {
    for (auto& s : h->sh) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.copy));
        CK(cudaStreamSynchronize(s.stream));
    }
    return 0;
}

int check_compute_args(b200jk* h, int nmat, const int* nocc, bool do_J, bool do_K, bool do_wK) {
    if (!h->have_layout) return fail(h, B200JK_ERR_INVALID, "b200jk_compute before b200jk_set_layout");
    if (nmat < 0 || (nmat > 0 && !nocc)) return fail(h, B200JK_ERR_INVALID, "bad nmat/nocc");
    if ((do_J || do_K) && !h->uploaded[0]) return fail(h, B200JK_ERR_INVALID, "tensor Ppq not uploaded");
    if (do_wK && !(h->uploaded[1] && h->uploaded[2]))
        return fail(h, B200JK_ERR_INVALID, "wK tasked but m1Ppq/wPpq not uploaded");
    for (int i = 0; i < nmat; i++)
        if (nocc[i] < 0) return fail(h, B200JK_ERR_INVALID, "negative nocc[%d]", i);
    return 0;
}

void begin_compute(b200jk* h) {
    for (auto& s : h->sh) {
        s.phases.clear();
        s.marks.clear();
        s.evused = 0;
        s.launches = 0;
        s.i8h_ops = s.i8h_plane_bytes = s.i8h_convert_bytes = s.i8k_ops = 0;
        s.i8h_nmod = s.i8h_chunks = s.i8h_cached = s.i8h_resident_rows = 0;
    }
}

int setup_shards(b200jk* h, int n, const int* devs) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(h, B200JK_ERR_NODEVICE, "no CUDA device visible: the B200 JK engine has no CPU fallback");
    }
    h->sh.resize(n);
    for (int i = 0; i < n; i++) {
        Shard& s = h->sh[i];
        s.dev = devs ? devs[i] : i;
        if (s.dev < 0 || s.dev >= ndev) return fail(h, B200JK_ERR_INVALID, "device %d out of range (%d visible)", s.dev, ndev);
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking));
        CK(cudaDeviceGetAttribute(&s.nsm, cudaDevAttrMultiProcessorCount, s.dev));
    }
    return load_encode(h);
}

void free_shard(Shard& s) {
    cudaSetDevice(s.dev);
    if (s.stream) cudaStreamSynchronize(s.stream);
    for (int w = 0; w < 3; w++) {
        if (s.tensor[w]) cudaFree(s.tensor[w]);
        if (s.d_amaps[w]) cudaFree(s.d_amaps[w]);
    }
    for (auto& c : s.cgmaps)
        if (c.d) cudaFree(c.d);
    void* ptrs[] = {s.Cg, s.d_row_off_unit, s.d_counter, s.d_tiles_sym, s.d_tiles_full, s.d_mpos, s.d_metric, s.fit_raw[0], s.fit_raw[1], s.fit_t, s.Dm,
                    s.d_fit_dst_off, s.d_fit_src_off, s.d_fit_dst_ld, s.d_fit_mi, s.d_fit_j0, s.d_fit_m,
                    s.d_row_off, s.d_ldm, s.d_sp, s.d_ign, s.d_cols, s.d_cols_off, s.in, s.out,
                    s.Ctl,       s.Ctr,   s.dpart, s.T1,  s.T2,     s.ws};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    s.i8.release();
    s.i8h.release();
    for (auto e : s.evpool) cudaEventDestroy(e);
    for (auto e : s.fit_raw_free)
        if (e) cudaEventDestroy(e);
    for (auto e : s.fit_meta_done)
        if (e) cudaEventDestroy(e);
    for (auto p : s.fit_meta)
        if (p) cudaFreeHost(p);
    if (s.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s.comm);
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.copy) cudaStreamDestroy(s.copy);
}

int alloc_tensor(b200jk* h, int which) {
    for (auto& s : h->sh) {
        if (s.tensor[which]) continue;
        CK(cudaSetDevice(s.dev));
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t need = s.tensor_doubles * sizeof(double);
        if (need + ((size_t)256 << 20) > free_b)
            return fail(h, B200JK_ERR_OOM,
                        "in-core tensor needs %.2f GiB on device %d but only %.2f GiB are free "
                        "(no out-of-core / CPU fallback; use more GPUs)",
                        need / 1073741824.0, s.dev, free_b / 1073741824.0);
        CK(malloc_elastic(h, (void**)&s.tensor[which], std::max<size_t>(need, 8)));
        CK(cudaMemsetAsync(s.tensor[which], 0, need, s.stream));
        // one TMA descriptor per row-block m: dims {sp(m), nq}, row pitch ld(m) doubles
        std::vector<CUtensorMap> maps(h->nbf);
        if (s.nq > 0) {
            for (size_t m = 0; m < h->nbf; m++) {
                int rc = make_map(h, &maps[m], s.tensor[which] + h->row_off_unit[m] * (size_t)s.nq, (uint64_t)h->sp[m],
                                  (uint64_t)s.nq, (uint64_t)h->ldm[m] * 8, BM);
                if (rc) return rc;
            }
        }
        CK(cudaMalloc((void**)&s.d_amaps[which], sizeof(CUtensorMap) * h->nbf));
        CK(cudaMemcpy(s.d_amaps[which], maps.data(), sizeof(CUtensorMap) * h->nbf, cudaMemcpyHostToDevice));
    }
    return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int b200jk_create(b200jk_t** out, int ngpu, const int* dev_ids) {
    if (!out || ngpu < 1) return B200JK_ERR_INVALID;
    b200jk* h = new b200jk();
    memset(&h->stats, 0, sizeof h->stats);
    *out = h;
    int rc = setup_shards(h, ngpu, dev_ids);
    if (rc) return rc;
    if (ngpu > 1) {
        if (!g_nccl.load(h->err)) return B200JK_ERR_NCCL;
        std::vector<int> devs(ngpu);
        std::vector<ncclComm_t> comms(ngpu);
        for (int i = 0; i < ngpu; i++) devs[i] = h->sh[i].dev;
        NK(g_nccl.CommInitAll(comms.data(), ngpu, devs.data()));
        for (int i = 0; i < ngpu; i++) h->sh[i].comm = comms[i];
        if ((rc = peer_setup_local(h))) return rc;
    }
    h->world = ngpu;
    h->stats.n_shards = ngpu;
    h->stats.reduce_kind = h->reduce_kind;
    return 0;
}

int b200jk_nccl_unique_id(void* out_bytes) {
    std::string err;
    if (!out_bytes || !g_nccl.load(err)) return B200JK_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id)) return B200JK_ERR_NCCL;
    memcpy(out_bytes, &id, sizeof id);
    return 0;
}

int b200jk_create_rank(b200jk_t** out, int device, int rank, int world, const void* nccl_id) {
    if (!out || world < 1 || rank < 0 || rank >= world) return B200JK_ERR_INVALID;
    b200jk* h = new b200jk();
    memset(&h->stats, 0, sizeof h->stats);
    *out = h;
    int rc = setup_shards(h, 1, &device);
    if (rc) return rc;
    h->rank = rank;
    h->world = world;
    h->rank_mode = true;
    h->stats.n_shards = 1;
    if (world > 1) {
        if (!nccl_id) return fail(h, B200JK_ERR_INVALID, "world > 1 needs an NCCL unique id");
        if (!g_nccl.load(h->err)) return B200JK_ERR_NCCL;
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof id);
        CK(cudaSetDevice(device));
        NK(g_nccl.CommInitRank(&h->sh[0].comm, world, id, rank));
        if ((rc = peer_setup_rank(h))) return rc;
    }
    h->stats.reduce_kind = h->reduce_kind;
    return 0;
}

int b200jk_register_host(b200jk_t* h, void* ptr, size_t bytes) {
    if (!h || !ptr || !bytes) return B200JK_ERR_INVALID;
    if (h->sh.empty()) return fail(h, B200JK_ERR_NODEVICE, "handle has no device");
    if (is_pinned(h, ptr, bytes)) return 0;
    CK(cudaSetDevice(h->sh[0].dev));
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    h->pinned.push_back({(const char*)ptr, bytes});
    return 0;
}

int b200jk_unregister_host(b200jk_t* h, void* ptr) {
    if (!h || !ptr) return B200JK_ERR_INVALID;
    for (size_t i = 0; i < h->pinned.size(); i++) {
        if (h->pinned[i].first == (const char*)ptr) {
            // the bookkeeping entry goes whatever CUDA says (psi4 may already have freed the matrix): a stale range
            // would make a later allocation at the same address look page-locked
            cudaError_t e = cudaHostUnregister(ptr);
            h->pinned.erase(h->pinned.begin() + i);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(h, B200JK_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
            }
            return 0;
        }
    }
    return fail(h, B200JK_ERR_INVALID, "pointer was not registered");
}

void b200jk_destroy(b200jk_t* h) {
    if (!h) return;
    for (auto& r : h->pinned) cudaHostUnregister((void*)r.first);
    for (auto& f : h->fit_pending) {
        cudaEventDestroy(f.a);
        cudaEventDestroy(f.b);
    }
    grad_free(h);
    for (auto& s : h->sh) peer_free(h, s);
    for (auto& s : h->sh) free_shard(s);
    for (auto& sl : h->stage) {
        for (auto e : sl.done)
            if (e) cudaEventDestroy(e);
        if (sl.buf) cudaFreeHost(sl.buf);
    }
    if (h->pin_in) cudaFreeHost(h->pin_in);
    if (h->pin_out) cudaFreeHost(h->pin_out);
    delete h;
}

const char* b200jk_last_error(const b200jk_t* h) { return h ? h->err.c_str() : "null handle"; }

int b200jk_set_layout(b200jk_t* h, size_t nbf, size_t naux, const size_t* small_skips, const size_t* big_skips,
                      const size_t* fun_index) {
    if (!h) return B200JK_ERR_INVALID;
    if (h->sh.empty()) return fail(h, B200JK_ERR_NODEVICE, "handle has no device");
    if (!nbf || !naux || !small_skips || !big_skips || !fun_index) return fail(h, B200JK_ERR_INVALID, "null/zero layout");
    if (nbf > 0x7fffffffu / 4 || naux > 0x7fffffffu / 4) return fail(h, B200JK_ERR_INVALID, "nbf/naux too large");
    // A second set_layout is a re-initialisation (jk.initialize() twice on one MemDFJK is legal in the reference and
    // simply recomputes, MemDFJK.cc:71-96): the resident tensors, their TMA maps and everything sized by the old
    // tables go; work buffers are kept (they are sized by capacity, not by shape).
    for (auto& s : h->sh) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaStreamSynchronize(s.copy));
        for (int w = 0; w < 3; w++) {
            if (s.tensor[w]) CK(cudaFree(s.tensor[w]));
            if (s.d_amaps[w]) CK(cudaFree(s.d_amaps[w]));
            s.tensor[w] = nullptr;
            s.d_amaps[w] = nullptr;
        }
        for (auto& c : s.cgmaps) {
            if (c.d) CK(cudaFree(c.d));
            c = Shard::CgMaps();
        }
        if (s.d_metric) CK(cudaFree(s.d_metric));
        s.d_metric = nullptr;
        s.have_metric = false;
    }
    for (int w = 0; w < 3; w++) h->uploaded[w] = false;
    h->nbf = nbf;
    h->naux = naux;
    h->small_skips.assign(small_skips, small_skips + nbf + 1);
    h->big_skips.assign(big_skips, big_skips + nbf + 1);
    h->sp.resize(nbf);
    h->ign.resize(nbf);
    h->ldm.resize(nbf);
    h->cols_off.resize(nbf);
    h->row_off_unit.resize(nbf + 1);
    h->cols.clear();
    h->cols.reserve(small_skips[nbf]);
    h->max_sp = 0;
    size_t tot = 0, run = 0;
    h->row_off_unit[0] = 0;
    for (size_t m = 0; m < nbf; m++) {
        size_t cnt = 0, skip = 0;
        h->cols_off[m] = h->cols.size();
        for (size_t n = 0; n < nbf; n++) {
            size_t f = fun_index[m * nbf + n];
            if (f) {
                if (f != cnt + 1) return fail(h, B200JK_ERR_INVALID, "fun_index row %zu is not a 1-based running rank", m);
                cnt++;
                if (n < m) skip++;
                h->cols.push_back((int)n);
            }
        }
        if (cnt != small_skips[m]) return fail(h, B200JK_ERR_INVALID, "small_skips[%zu] != kept partners in fun_index", m);
        if (big_skips[m] != run) return fail(h, B200JK_ERR_INVALID, "big_skips[%zu] inconsistent with small_skips*naux", m);
        if (!fun_index[m * nbf + m]) return fail(h, B200JK_ERR_INVALID, "diagonal pair (%zu,%zu) screened out", m, m);
        run += cnt * naux;
        tot += cnt;
        h->sp[m] = (int)cnt;
        h->ign[m] = (int)skip;
        h->ldm[m] = round_up((int)cnt, 4);
        h->row_off_unit[m + 1] = h->row_off_unit[m] + h->ldm[m];
        h->max_sp = std::max(h->max_sp, (int)cnt);
    }
    if (small_skips[nbf] != tot || big_skips[nbf] != run) return fail(h, B200JK_ERR_INVALID, "table totals inconsistent");
    // mirror ranks (the mask must be symmetric: dfhelper.cc:1670-1672 reads f(onu,omu) whenever f(omu,onu) != 0)
    h->mpos.assign(h->cols.size(), 0);
    h->symm_big_skips.assign(nbf + 1, 0);
    for (size_t m = 0; m < nbf; m++) {
        for (int k = 0; k < h->sp[m]; k++) {
            size_t n = (size_t)h->cols[h->cols_off[m] + k];
            size_t f = fun_index[n * nbf + m];
            if (!f) return fail(h, B200JK_ERR_INVALID, "pair mask is not symmetric at (%zu,%zu)", m, n);
            h->mpos[h->cols_off[m] + k] = (int)f - 1;
        }
        h->symm_big_skips[m + 1] = h->symm_big_skips[m] + (size_t)(h->sp[m] - h->ign[m]) * naux;
    }

    // Q shards: contiguous, near-equal.  rank mode: this process owns shard `rank` of `world`.
    int nshard_total = h->rank_mode ? h->world : (int)h->sh.size();
    for (size_t i = 0; i < h->sh.size(); i++) {
        Shard& s = h->sh[i];
        int idx = h->rank_mode ? h->rank : (int)i;
        s.q0 = (int)((naux * (size_t)idx) / nshard_total);
        s.q1 = (int)((naux * (size_t)(idx + 1)) / nshard_total);
        s.nq = s.q1 - s.q0;
        s.tensor_doubles = h->row_off_unit[nbf] * (size_t)s.nq;
        std::vector<size_t> row_off(nbf);
        for (size_t m = 0; m < nbf; m++) row_off[m] = h->row_off_unit[m] * (size_t)s.nq;
        CK(cudaSetDevice(s.dev));
        int rc;
        if ((rc = upload_vec(h, &s.d_row_off, row_off))) return rc;
        if ((rc = upload_vec(h, &s.d_row_off_unit, h->row_off_unit))) return rc;
        s.screened = h->small_skips[nbf] < nbf * nbf;
        if ((rc = upload_vec(h, &s.d_ldm, h->ldm))) return rc;
        if ((rc = upload_vec(h, &s.d_sp, h->sp))) return rc;
        if ((rc = upload_vec(h, &s.d_ign, h->ign))) return rc;
        if ((rc = upload_vec(h, &s.d_cols, h->cols))) return rc;
        if ((rc = upload_vec(h, &s.d_cols_off, h->cols_off))) return rc;
        if ((rc = upload_vec(h, &s.d_mpos, h->mpos))) return rc;
        {
            std::string e8;
            if (i8h_set_layout(s.i8h, h->sp, &e8)) return fail(h, B200JK_ERR_CUDA, "INT8 half transform layout: %s", e8.c_str());
        }
        {
            // K-GEMM tile lists sorted by live area (rows x cols inside nbf), largest first
            const int n1d = ((int)nbf + BM - 1) / BM;
            auto live = [&](int t) { return std::min<int>(BM, (int)nbf - t * BM); };
            for (int sym = 0; sym < 2; sym++) {
                std::vector<int2> tl;
                for (int a = 0; a < n1d; a++)
                    for (int b = sym ? a : 0; b < n1d; b++) tl.push_back(make_int2(a, b));
                std::stable_sort(tl.begin(), tl.end(), [&](const int2& x, const int2& y) {
                    return live(x.x) * live(x.y) > live(y.x) * live(y.y);
                });
                if ((rc = upload_vec(h, sym ? &s.d_tiles_sym : &s.d_tiles_full, tl))) return rc;
                (sym ? s.ntiles_sym : s.ntiles_full) = (int)tl.size();
            }
            // work-queue head + the K GEMM's per-(tile, consumer warp) arrival counters
            std::vector<int> zero(1 + (size_t)s.ntiles_full * WS_CONSUMER_WARPS, 0);
            if ((rc = upload_vec(h, &s.d_counter, zero))) return rc;
        }
    }
    h->stats.q_begin = h->sh[0].q0;
    h->stats.q_end = h->sh[0].q1;
    h->have_layout = true;
    return 0;
}

int b200jk_upload_rows(b200jk_t* h, int which, size_t m0, size_t m1, const double* host_rows) {
    if (!h) return B200JK_ERR_INVALID;
    if (!h->have_layout) return fail(h, B200JK_ERR_INVALID, "upload before set_layout");
    if (which < 0 || which > 2 || m0 > m1 || m1 > h->nbf || !host_rows) return fail(h, B200JK_ERR_INVALID, "bad upload args");
    const double t_call = now_s();
    h->stage_wait_s = h->stage_alloc_s = h->stage_copy_s = 0;
    int rc = alloc_tensor(h, which);
    if (rc) return rc;
    // Q range the local shards hold (everything for an in-process handle, one shard's rows in rank mode): only that
    // part of each row-block is touched
    size_t qlo = h->naux, qhi = 0;
    for (auto& s : h->sh)
        if (s.nq) {
            qlo = std::min<size_t>(qlo, s.q0);
            qhi = std::max<size_t>(qhi, s.q1);
        }
    const size_t total_bytes = (h->big_skips[m1] - h->big_skips[m0]) * sizeof(double);
    const bool direct = is_pinned(h, host_rows, total_bytes);  // page-locked by the caller: DMA straight out of it
    // Pageable source: groups of row-blocks (<= 256 MB of the local Q range) go through the page-locked ring -- a few
    // threads fill one slot while the DMA engines drain the other; the call returns when the last slot is filled.
    const size_t budget = stage_budget((size_t)256 << 20);
    size_t ma = m0;
    while (ma < m1 && qlo < qhi) {
        size_t mb = ma, doubles = 0;
        while (mb < m1) {
            size_t add = (qhi - qlo) * (size_t)h->sp[mb];
            if (mb > ma && (doubles + add) * 8 > budget) break;
            doubles += add;
            mb++;
        }
        b200jk::StageSlot* slot = nullptr;
        std::vector<size_t> soff(mb - ma, 0);
        if (!direct) {
            if ((rc = stage_acquire(h, doubles, std::min<size_t>(budget / 8, h->big_skips[h->nbf]), &slot))) return rc;
            const double tc = now_s();
            // the local Q range of each row-block is one contiguous slab of the caller's block
            size_t off = 0;
            std::vector<std::thread> th;
            const int nt = 8;
            for (size_t m = ma; m < mb; m++) {
                soff[m - ma] = off;
                off += (qhi - qlo) * (size_t)h->sp[m];
            }
            for (int t = 0; t < nt; t++)
                th.emplace_back([&, t] {
                    for (size_t m = ma + t; m < mb; m += nt) {
                        size_t sp = h->sp[m];
                        memcpy(slot->buf + soff[m - ma], host_rows + (h->big_skips[m] - h->big_skips[m0]) + qlo * sp,
                               (qhi - qlo) * sp * sizeof(double));
                    }
                });
            for (auto& x : th) x.join();
            h->stage_copy_s += now_s() - tc;
        }
        for (size_t si = 0; si < h->sh.size(); si++) {
            Shard& s = h->sh[si];
            if (!s.nq) continue;
            CK(cudaSetDevice(s.dev));
            for (size_t m = ma; m < mb; m++) {
                size_t sp = h->sp[m];
                const double* src = direct ? host_rows + (h->big_skips[m] - h->big_skips[m0]) + (size_t)s.q0 * sp
                                           : slot->buf + soff[m - ma] + ((size_t)s.q0 - qlo) * sp;
                double* dst = s.tensor[which] + h->row_off_unit[m] * (size_t)s.nq;
                CK(cudaMemcpy2DAsync(dst, (size_t)h->ldm[m] * 8, src, sp * 8, sp * 8, (size_t)s.nq, cudaMemcpyHostToDevice,
                                     s.stream));
            }
            if (slot) CK(cudaEventRecord(slot->done[si], s.stream));
        }
        ma = mb;
    }
    // the caller's own memory must stay untouched until the DMA has read it; staged data is already ours.  The last
    // block of the tensor drains everything so "uploaded" means resident.
    const double ts = now_s();
    if (direct || m1 == h->nbf)
        if ((rc = producers_sync(h))) return rc;
    if (stage_trace())
        fprintf(stderr, "[b200jk] upload_rows [%zu,%zu) %.1f MB %s: total %.4f s (ring wait %.4f, pin alloc %.4f, copy %.4f, drain %.4f)\n",
                m0, m1, total_bytes / 1e6, direct ? "direct" : "staged", now_s() - t_call, h->stage_wait_s, h->stage_alloc_s,
                h->stage_copy_s, now_s() - ts);
    if (m1 == h->nbf) h->uploaded[which] = true;  // streaming: the last block completes the tensor
    for (auto& s : h->sh) s.i8h.expo_valid[which] = false;  // the row scales of the INT8 half transform follow the tensor
    h->stats.hbm_tensor_bytes = 0;
    for (int w = 0; w < 3; w++)
        if (h->sh[0].tensor[w]) h->stats.hbm_tensor_bytes += h->sh[0].tensor_doubles * 8;
    return 0;
}

int b200jk_upload(b200jk_t* h, int which, const double* host_pQq) {
    if (!h) return B200JK_ERR_INVALID;
    return b200jk_upload_rows(h, which, 0, h->nbf, host_pQq);
}

int b200jk_fill_synthetic(b200jk_t* h, int which, uint64_t seed, const double* amp) {
    if (!h) return B200JK_ERR_INVALID;
    if (!h->have_layout || which < 0 || which > 2 || !amp) return fail(h, B200JK_ERR_INVALID, "bad fill_synthetic args");
    int rc = alloc_tensor(h, which);
    if (rc) return rc;
    size_t n2 = h->nbf * h->nbf;
    uint64_t seedh = splitmix64(seed);
    for (auto& s : h->sh) {
        CK(cudaSetDevice(s.dev));
        double* damp = nullptr;
        CK(cudaMalloc((void**)&damp, n2 * 8));
        CK(cudaMemcpyAsync(damp, amp, n2 * 8, cudaMemcpyHostToDevice, s.stream));
        if (s.nq) {
            synth_fill_kernel<<<dim3((s.nq + 7) / 8, (unsigned)h->nbf), 256, 0, s.stream>>>(
                s.tensor[which], s.d_row_off, s.d_ldm, s.d_sp, s.d_cols, s.d_cols_off, damp, (int)h->nbf, s.nq, s.q0, seedh);
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaFree(damp));
    }
    h->uploaded[which] = true;
    for (auto& s : h->sh) s.i8h.expo_valid[which] = false;
    h->stats.hbm_tensor_bytes = 0;
    for (int w = 0; w < 3; w++)
        if (h->sh[0].tensor[w]) h->stats.hbm_tensor_bytes += h->sh[0].tensor_doubles * 8;
    return 0;
}

int b200jk_download_rows(b200jk_t* h, int which, size_t m, size_t q0, size_t q1, double* host_out) {
    if (!h) return B200JK_ERR_INVALID;
    if (!h->have_layout || which < 0 || which > 2 || m >= h->nbf || q0 > q1 || q1 > h->naux || !host_out)
        return fail(h, B200JK_ERR_INVALID, "bad download args");
    size_t sp = h->sp[m];
    bool any = false;
    for (auto& s : h->sh) {
        size_t a = std::max<size_t>(q0, s.q0), b = std::min<size_t>(q1, s.q1);
        if (a >= b || !s.tensor[which]) continue;
        any = true;
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.stream));  // uploads / fitting kernels queued behind an earlier call
        const double* src = s.tensor[which] + h->row_off_unit[m] * (size_t)s.nq + (a - s.q0) * (size_t)h->ldm[m];
        CK(cudaMemcpy2D(host_out + (a - q0) * sp, sp * 8, src, (size_t)h->ldm[m] * 8, sp * 8, b - a, cudaMemcpyDeviceToHost));
    }
    if (!any && q1 > q0 && !h->rank_mode) return fail(h, B200JK_ERR_INVALID, "tensor not resident");
    return 0;
}

int b200jk_set_work_budget(b200jk_t* h, uint64_t bytes) {
    if (!h) return B200JK_ERR_INVALID;
    h->work_budget = bytes;
    for (auto& s : h->sh) {  // force re-planning
        cudaSetDevice(s.dev);
        if (s.T1) cudaFree(s.T1);
        if (s.T2) cudaFree(s.T2);
        s.T1 = s.T2 = nullptr;
        s.T_cap = s.T2_cap = 0;
    }
    return 0;
}

int b200jk_set_kgemm(b200jk_t* h, int arm, int moduli) {
    if (!h || arm < 0 || arm > 2 || (moduli && (moduli < I8_MINMOD || moduli > I8_MAXMOD))) return B200JK_ERR_INVALID;
    h->kgemm_arm = arm;
    h->kgemm_nmod = moduli;
    return 0;
}

int b200jk_set_half(b200jk_t* h, int arm, int moduli) {
    if (!h || arm < 0 || arm > 2 || (moduli && (moduli < I8_MINMOD || moduli > I8_MAXMOD))) return B200JK_ERR_INVALID;
    h->half_arm = arm;
    h->half_nmod = moduli;
    for (auto& s : h->sh)
        for (int w = 0; w < 3; w++) s.i8h.expo_valid[w] = s.i8h.expo_valid[w] && (moduli == 0 || s.i8h.expo_nmod[w] == moduli);
    return 0;
}

int b200jk_hbm_estimate(const b200jk_t* h, size_t max_nocc, int do_wK, uint64_t* bytes_per_gpu) {
    if (!h || !bytes_per_gpu) return B200JK_ERR_INVALID;
    if (!h->have_layout) return B200JK_ERR_INVALID;
    const Shard& s = h->sh[0];
    uint64_t b = s.tensor_doubles * 8 * (do_wK ? 3 : 1);
    uint64_t op = (max_nocc + 1) & ~(uint64_t)1;
    b += (uint64_t)h->nbf * s.nq * op * 8 * 2;          // T1, T2 (full shard)
    b += (uint64_t)h->nbf * s.nq * 8;                   // dpart
    b += (uint64_t)6 * h->nbf * h->nbf * 8 + ((uint64_t)1 << 30);  // in/out + split-K partials
    if (want_cgather(s, s.nq)) b += (uint64_t)h->row_off_unit[h->nbf] * (max_nocc + 1) * 8;  // pre-gathered C^T
    *bytes_per_gpu = b;
    return 0;
}

int b200jk_get_stats(const b200jk_t* h, b200jk_stats* out) {
    if (!h || !out) return B200JK_ERR_INVALID;
    *out = h->stats;
    return 0;
}

static int compute_impl(b200jk_t* h, bool host_ops, int nmat, const double* const* Cl, const double* const* Cr,
                        const int* nocc, const double* const* D, double* const* J, double* const* K,
                        double* const* wK, int do_J, int do_K, int do_wK) {
    if (!h) return B200JK_ERR_INVALID;
    int rc = check_compute_args(h, nmat, nocc, do_J, do_K, do_wK);
    if (rc) return rc;
    if (!host_ops && h->sh.size() != 1) return fail(h, B200JK_ERR_INVALID, "compute_device needs a one-shard handle");
    if (nmat == 0) return 0;
    if ((do_K || do_wK) && !Cl) return fail(h, B200JK_ERR_INVALID, "K tasked without C_left");
    if (do_J && !D) return fail(h, B200JK_ERR_INVALID, "J tasked without D");
    // rank mode: a rank other than 0 may pass no output arrays at all -- it takes part in the build and the all-reduce
    // but brings nothing home (one host process needs the result once; psi4 is a single process)
    const bool fetch = !(host_ops && h->rank_mode && h->rank != 0 && !J && !K && !wK);
    if (fetch && ((do_J && !J) || (do_K && !K) || (do_wK && !wK)))
        return fail(h, B200JK_ERR_INVALID, "tasked output array is NULL");

    Task t;
    t.nmat = nmat;
    t.nocc = nocc;
    t.lr = (Cr == nullptr);
    t.do_J = do_J;
    t.do_K = do_K;
    t.do_wK = do_wK;
    t.n2 = h->nbf * h->nbf;
    t.nprod = (do_J ? 1 : 0) + (do_K ? 1 : 0) + (do_wK ? 1 : 0);
    t.max_o = 0;
    for (int i = 0; i < nmat; i++) t.max_o = std::max(t.max_o, nocc[i]);
    if (t.nprod == 0) return 0;
    const size_t N = h->nbf, n2 = t.n2;
    t.same_left.assign(nmat, 0);
    if ((do_K || do_wK) && !t.lr && !getenv("B200JK_NO_T1_REUSE"))
        for (int i = 1; i < nmat; i++) {
            if (!nocc[i] || nocc[i] != nocc[i - 1]) continue;
            // device operands: identity only; host operands: identity or equal contents (USO2AO makes per-matrix
            // AO copies of one SO block, jk.cc:404-446)
            t.same_left[i] = Cl[i] == Cl[i - 1] ||
                             (host_ops && memcmp(Cl[i], Cl[i - 1], N * (size_t)nocc[i] * sizeof(double)) == 0);
        }
    begin_compute(h);

    std::vector<int> qc(h->sh.size());
    for (size_t i = 0; i < h->sh.size(); i++)
        if ((rc = ensure_work(h, h->sh[i], t, &qc[i]))) return rc;

    // Order of one build: C up -> K/wK kernels -> [D up on the copy stream meanwhile] -> J sweeps; the K results are
    // all-reduced and brought home on the copy stream while J runs.  Event pairs bracket every phase.
    const OutLayout ol = out_layout(t);
    const size_t nsh = h->sh.size();
    std::vector<Phase> total(nsh);
    for (size_t i = 0; i < nsh; i++) {
        Shard& s = h->sh[i];
        CK(cudaSetDevice(s.dev));
        total[i].tag = 6;
        total[i].a = get_event(s);
        total[i].b = get_event(s);
        CK(cudaEventRecord(total[i].a, s.stream));
    }

    std::vector<std::vector<const double*>> dCl(nsh), dCr(nsh), dD(nsh);
    const bool needC = do_K || do_wK;
    size_t c_doubles = 0, d_doubles = do_J ? (size_t)nmat * n2 : 0;
    if (needC)
        for (int i = 0; i < nmat; i++) c_doubles += N * (size_t)nocc[i] * (t.lr ? 1 : 2);
    std::vector<size_t> offCl(nmat, 0), offCr(nmat, 0);
    if (host_ops) {
        size_t in_doubles = std::max<size_t>(c_doubles + d_doubles, 1);
        if (in_doubles > h->pin_in_cap) {
            if (h->pin_in) CK(cudaFreeHost(h->pin_in));
            h->pin_in = nullptr;
            CK(cudaHostAlloc((void**)&h->pin_in, in_doubles * 8, cudaHostAllocPortable));
            h->pin_in_cap = in_doubles;
        }
        if (ol.total > h->pin_out_cap) {
            if (h->pin_out) CK(cudaFreeHost(h->pin_out));
            h->pin_out = nullptr;
            CK(cudaHostAlloc((void**)&h->pin_out, ol.total * 8, cudaHostAllocPortable));
            h->pin_out_cap = ol.total;
        }
        size_t off = 0;
        if (needC) {
            for (int i = 0; i < nmat; i++) {
                size_t c = N * (size_t)nocc[i];
                offCl[i] = off;
                if (c) memcpy(h->pin_in + off, Cl[i], c * 8);
                off += c;
                if (!t.lr) {
                    offCr[i] = off;
                    if (c) memcpy(h->pin_in + off, Cr[i], c * 8);
                    off += c;
                }
            }
        }
        for (size_t si = 0; si < nsh; si++) {
            Shard& s = h->sh[si];
            CK(cudaSetDevice(s.dev));
            if ((rc = grow(h, &s.in, &s.in_cap, in_doubles))) return rc;
            if (c_doubles) {
                PhaseScope ps(s, 4);
                CK(cudaMemcpyAsync(s.in, h->pin_in, c_doubles * 8, cudaMemcpyHostToDevice, s.stream));
            }
            dCl[si].assign(nmat, nullptr);
            dCr[si].assign(nmat, nullptr);
            dD[si].assign(nmat, nullptr);
            for (int i = 0; i < nmat; i++) {
                if (needC) dCl[si][i] = s.in + offCl[i];
                if (needC && !t.lr) dCr[si][i] = s.in + offCr[i];
                if (do_J) dD[si][i] = s.in + c_doubles + (size_t)i * n2;
            }
        }
    } else {
        dCl[0].assign(nmat, nullptr);
        dCr[0].assign(nmat, nullptr);
        dD[0].assign(nmat, nullptr);
        for (int i = 0; i < nmat; i++) {
            if (needC) dCl[0][i] = Cl[i];
            if (needC && !t.lr) dCr[0][i] = Cr[i];
            if (do_J) dD[0][i] = D[i];
        }
    }

    // D goes up on the copy stream: under the K kernels normally, ahead of them when the first J sweep is fused
    // into the half transform (the density row is then an operand of K3)
    bool fuse_any = false;
    if (do_J && do_K)
        for (int i = 0; i < nmat; i++) fuse_any = fuse_any || (nocc[i] > 0 && can_fuse_j(nocc[i]));
    std::vector<cudaEvent_t> evK(nsh), evD(nsh);
    for (size_t si = 0; si < nsh; si++) {
        evK[si] = get_event(h->sh[si]);
        evD[si] = get_event(h->sh[si]);
    }
    auto upload_D = [&]() -> int {
        // densities in caller memory registered with b200jk_register_host go up by DMA as they are; others are staged
        std::vector<const double*> srcD(nmat);
        for (int i = 0; i < nmat; i++) {
            if (is_pinned(h, D[i], n2 * 8)) {
                srcD[i] = D[i];
            } else {
                par_memcpy(h->pin_in + c_doubles + (size_t)i * n2, D[i], n2 * 8);
                srcD[i] = h->pin_in + c_doubles + (size_t)i * n2;
            }
        }
        for (size_t si = 0; si < nsh; si++) {
            Shard& s = h->sh[si];
            CK(cudaSetDevice(s.dev));
            Phase ph;
            ph.tag = 4;
            ph.a = get_event(s);
            ph.b = get_event(s);
            CK(cudaEventRecord(ph.a, s.copy));
            for (int i = 0; i < nmat; i++)
                CK(cudaMemcpyAsync(s.in + c_doubles + (size_t)i * n2, srcD[i], n2 * 8, cudaMemcpyHostToDevice, s.copy));
            CK(cudaEventRecord(ph.b, s.copy));
            s.phases.push_back(ph);
            CK(cudaEventRecord(evD[si], s.copy));
            CK(cudaStreamWaitEvent(s.stream, evD[si], 0));
        }
        return 0;
    };
    if (host_ops && do_J && fuse_any && (rc = upload_D())) return rc;

    // K / wK kernels
    for (size_t si = 0; si < nsh; si++) {
        Shard& s = h->sh[si];
        if ((rc = run_device_K(h, s, t, dCl[si].data(), dCr[si].data(), do_J ? dD[si].data() : nullptr, qc[si]))) return rc;
        CK(cudaSetDevice(s.dev));
        CK(cudaEventRecord(evK[si], s.stream));
    }
    if (host_ops && do_J && !fuse_any && (rc = upload_D())) return rc;

    // J sweeps
    for (size_t si = 0; si < nsh; si++)
        if ((rc = run_device_J(h, h->sh[si], t, dD[si].data()))) return rc;

    // K results: reduce + download on the copy stream (overlaps the J sweeps); J results follow on the compute stream
    Shard& s0 = h->sh[0];
    double* const* outs[3] = {do_J ? J : nullptr, do_K ? K : nullptr, do_wK ? wK : nullptr};
    cudaEvent_t evKhome = nullptr;
    if (host_ops) {
        for (size_t si = 0; si < nsh; si++) {
            CK(cudaSetDevice(h->sh[si].dev));
            CK(cudaStreamWaitEvent(h->sh[si].copy, evK[si], 0));
        }
        // (result into every shard's window: in rank mode any rank may ask for the outputs, and a rank cannot know
        // whether its peers do; the extra NVLink writes are count doubles per rank, ~0.05 ms at C60)
        if ((rc = reduce_shards(h, ol.offK, ol.countKW, true, 0, -1))) return rc;
        CK(cudaSetDevice(s0.dev));
        if (ol.countKW && fetch) {
            Phase ph;
            ph.tag = 5;
            ph.a = get_event(s0);
            ph.b = get_event(s0);
            CK(cudaEventRecord(ph.a, s0.copy));
            size_t off = ol.offK;
            for (int pr = 1; pr < 3; pr++) {
                if (!outs[pr]) continue;
                for (int i = 0; i < nmat; i++, off += n2) {
                    double* dst = is_pinned(h, outs[pr][i], n2 * 8) ? outs[pr][i] : h->pin_out + off;  // registered: DMA home
                    CK(cudaMemcpyAsync(dst, s0.out + off, n2 * 8, cudaMemcpyDeviceToHost, s0.copy));
                }
            }
            CK(cudaEventRecord(ph.b, s0.copy));
            s0.phases.push_back(ph);
        }
        evKhome = get_event(s0);
        CK(cudaEventRecord(evKhome, s0.copy));
        if ((rc = reduce_shards(h, ol.offJ, ol.countJ, false, 1, -1))) return rc;
        CK(cudaSetDevice(s0.dev));
        if (ol.countJ && fetch) {
            PhaseScope ps(s0, 5);
            for (int i = 0; i < nmat; i++) {
                size_t off = ol.offJ + (size_t)i * n2;
                double* dst = is_pinned(h, J[i], n2 * 8) ? J[i] : h->pin_out + off;
                CK(cudaMemcpyAsync(dst, s0.out + off, n2 * 8, cudaMemcpyDeviceToHost, s0.stream));
            }
        }
        // every shard's compute stream also waits for its copy stream so `total` covers both
        for (size_t si = 0; si < nsh; si++) {
            Shard& s = h->sh[si];
            CK(cudaSetDevice(s.dev));
            cudaEvent_t e = get_event(s);
            CK(cudaEventRecord(e, s.copy));
            CK(cudaStreamWaitEvent(s.stream, e, 0));
        }
    } else {
        // the same two sums as the host-operand arm (K group, then J group), result on every rank
        if ((rc = reduce_shards(h, ol.offK, ol.countKW, false, 0, -1))) return rc;
        if ((rc = reduce_shards(h, ol.offJ, ol.countJ, false, 1, -1))) return rc;
        CK(cudaSetDevice(s0.dev));
        size_t off = 0;
        for (int pr = 0; pr < 3; pr++) {
            if (!outs[pr]) continue;
            for (int i = 0; i < nmat; i++, off += n2)
                CK(cudaMemcpyAsync(outs[pr][i], s0.out + off, n2 * 8, cudaMemcpyDeviceToDevice, s0.stream));
        }
    }
    for (size_t i = 0; i < nsh; i++) {
        Shard& s = h->sh[i];
        CK(cudaSetDevice(s.dev));
        CK(cudaEventRecord(total[i].b, s.stream));
        s.phases.push_back(total[i]);
    }
    if (host_ops && fetch) {
        // K / wK leave the staging buffer while the GPU is still busy with J
        CK(cudaSetDevice(s0.dev));
        CK(cudaEventSynchronize(evKhome));
        size_t off = ol.offK;
        for (int pr = 1; pr < 3; pr++) {
            if (!outs[pr]) continue;
            for (int i = 0; i < nmat; i++, off += n2)
                if (!is_pinned(h, outs[pr][i], n2 * 8)) par_memcpy(outs[pr][i], h->pin_out + off, n2 * 8);
        }
    }
    for (auto& s : h->sh) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaStreamSynchronize(s.copy));
    }
    if (host_ops && do_J && fetch)
        for (int i = 0; i < nmat; i++)
            if (!is_pinned(h, J[i], n2 * 8)) par_memcpy(J[i], h->pin_out + ol.offJ + (size_t)i * n2, n2 * 8);
    if ((rc = peer_check_status(h))) return rc;
    account_work(h, t);
    collect_stats(h);
    h->stats.reduce_kind = h->reduce_kind;
    h->stats.hbm_work_bytes = 0;
    {
        Shard& s = h->sh[0];
        h->stats.hbm_work_bytes =
            8 * (s.in_cap + s.out_cap + 2 * s.ct_cap + s.dpart_cap + s.T_cap + s.T2_cap + s.ws_cap + s.cg_cap) +
            s.i8.planes_cap[0] + s.i8.planes_cap[1] + s.i8.ws_cap;
        h->stats.kgemm_kind = s.kgemm_kind;
        h->stats.half_kind = s.half_kind;
        h->stats.hbm_work_bytes += s.i8h.arena_cap;
        h->stats.kgemm_moduli = s.kgemm_moduli;
    }
    return 0;
}

int b200jk_compute(b200jk_t* h, int nmat, const double* const* Cl, const double* const* Cr, const int* nocc,
                   const double* const* D, double* const* J, double* const* K, double* const* wK, int do_J,
                   int do_K, int do_wK) {
    return compute_impl(h, true, nmat, Cl, Cr, nocc, D, J, K, wK, do_J, do_K, do_wK);
}

int b200jk_compute_device(b200jk_t* h, int nmat, const double* const* dCl, const double* const* dCr,
                          const int* nocc, const double* const* dD, double* const* dJ, double* const* dK,
                          double* const* dwK, int do_J, int do_K, int do_wK) {
    return compute_impl(h, false, nmat, dCl, dCr, nocc, dD, dJ, dK, dwK, do_J, do_K, do_wK);
}

int b200jk_dev_alloc(b200jk_t* h, size_t bytes, void** dptr) {
    if (!h || !dptr || h->sh.empty()) return B200JK_ERR_INVALID;
    CK(cudaSetDevice(h->sh[0].dev));
    CK(malloc_elastic(h, dptr, std::max<size_t>(bytes, 8)));
    return 0;
}
int b200jk_dev_free(b200jk_t* h, void* dptr) {
    if (!h || h->sh.empty()) return B200JK_ERR_INVALID;
    CK(cudaSetDevice(h->sh[0].dev));
    CK(cudaFree(dptr));
    return 0;
}
int b200jk_dev_copy(b200jk_t* h, void* dst, const void* src, size_t bytes, int kind) {
    if (!h || h->sh.empty()) return B200JK_ERR_INVALID;
    CK(cudaSetDevice(h->sh[0].dev));
    cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy(dst, src, bytes, k));
    return 0;
}

int b200jk_fp64_peak(b200jk_t* h, int kind, double seconds, double* out4) {
    if (!h || !out4 || h->sh.empty() || kind < 0 || kind > 4) return B200JK_ERR_INVALID;
    Shard& s = h->sh[0];
    CK(cudaSetDevice(s.dev));
    int nsm = 148;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s.dev));
    // kinds 3 / 4: DMMA at the occupancy of the GEMM kernels (8 / 4 warps per SM, one CTA per SM)
    const int iters = 4000, blocks = nsm * (kind >= 3 ? 1 : 4), threads = (kind == 4 ? 128 : 256);
    double* out = nullptr;
    unsigned long long* clk = nullptr;
    CK(cudaMalloc((void**)&out, 8));
    CK(cudaMalloc((void**)&clk, sizeof(unsigned long long) * 2 * blocks));
    std::vector<unsigned long long> hclk(2 * blocks);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    const double flops = (double)blocks * threads * iters * 16.0 *
                         ((kind == 0 || kind >= 2 ? 512.0 / 32.0 : 0.0) + (kind == 1 || kind == 2 ? 2.0 : 0.0));
    auto launch = [&]() {
        if (kind == 0 || kind >= 3)
            fp64_peak_kernel<0><<<blocks, threads, 0, s.stream>>>(out, clk, iters);
        else if (kind == 1)
            fp64_peak_kernel<1><<<blocks, threads, 0, s.stream>>>(out, clk, iters);
        else
            fp64_peak_kernel<2><<<blocks, threads, 0, s.stream>>>(out, clk, iters);
    };
    auto mhz = [&]() {
        cudaMemcpy(hclk.data(), clk, sizeof(unsigned long long) * 2 * blocks, cudaMemcpyDeviceToHost);
        double cyc = 0, ns = 0;
        for (int i = 0; i < blocks; i++) {
            cyc += (double)hclk[2 * i];
            ns += (double)hclk[2 * i + 1];
        }
        return ns > 0 ? cyc / ns * 1e3 : 0.0;
    };
    // burst: best single launch out of a few, from idle
    CK(cudaStreamSynchronize(s.stream));
    double burst = 0, burst_mhz = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a, s.stream));
        launch();
        CK(cudaEventRecord(b, s.stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > burst) {
            burst = tf;
            burst_mhz = mhz();
        }
    }
    // sustained: back-to-back launches for `seconds`
    int n = 0;
    float total_ms = 0;
    CK(cudaEventRecord(a, s.stream));
    do {
        for (int k = 0; k < 8; k++, n++) launch();
        CK(cudaEventRecord(b, s.stream));
        CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&total_ms, a, b));
    } while (total_ms < seconds * 1e3);
    out4[0] = burst;
    out4[1] = flops * n / (total_ms * 1e-3) / 1e12;
    out4[2] = burst_mhz;
    out4[3] = mhz();
    CK(cudaEventDestroy(a));
    CK(cudaEventDestroy(b));
    CK(cudaFree(out));
    CK(cudaFree(clk));
    return 0;
}

}  // extern "C"

#include "fit_host.inl"
#include "grad_host.inl"
#include "power_host.inl"
