// dmma_ws.cuh -- persistent, warp-specialised FP64 tensor-core pipeline for the two compute-bound
// kernels of the DF-K build (second generation of dmma_gemm.cuh, same math, same tile shapes):
//
//   K3  half_ws_kernel<NB>  : T[m,q,i] = sum_k B_m[q,k] * C[n_k,i]      (lib3index/dfhelper.cc:2162-2186)
//   K4  kgemm_ws_kernel     : K[m,n]  += sum_{q,i} T1[m,qi] * T2[n,qi]  (lib3index/dfhelper.cc:3374)
//
// One CTA per SM, 12 warps = 3 warpgroups (registers re-balanced with setmaxnreg: 40 / 232 / 232):
//   warps 0..3  producer warpgroup.  One elected lane issues TMA tile loads (cp.async.bulk.tensor.2d,
//               SWIZZLE_128B, completion on an mbarrier); when a row-block is screened (sp(m) < nbf) all
//               128 producer threads gather the C operand along the kept-partner list with 8-byte cp.async
//               into the same swizzled layout and signal the same mbarrier (cp.async.mbarrier.arrive.noinc).
//   warps 4..11 consumers.  4 (M) x 2 (N) warp grid, warp tile 32 x 8*NB, DMMA m8n8k4 with register
//               accumulators.  They wait on full[stage], read fragments, and release empty[stage].
// No __syncthreads in the main loop: warps drift apart, so while one warp waits for data or stores its
// tile the other warp on the same scheduler keeps the DMMA pipe busy.  The CTA is persistent: the producer
// pulls work items from a global atomic counter (dynamic scheduling: edge tiles and screened row-blocks
// cost less than interior ones, a static split leaves SMs idle) and tags the first stage of each item with
// its id, so the consumers need no other communication; the producer prefetches the next item's first
// stages while the consumers are still in the epilogue of the previous one.
//
// Shared tile = rows x 128 bytes (16 doubles of k), TMA 128-byte swizzle: 16-byte chunk c of row r lands
// at chunk c ^ (r & 7).  DMMA k-step s of a stage uses k = {2s, 2s+1, 2s+8, 2s+9} (a permutation of k is
// free in a dot product as long as both operands agree): lane (g = lane/4, t = lane%4) reads chunk
// (s + 4*(t/2)) ^ g, half t%2 -- the 16 lanes of a half-warp then hit 16 distinct 8-byte bank pairs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmma_gemm.cuh"

namespace b2k {

constexpr int WS_STAGES = 6;  // 7 stages fit as well but measured no faster: the ring is not latency-bound
constexpr int WS_CONSUMER_WARPS = 8;
constexpr int WS_PRODUCER_WARPS = 4;   // one warpgroup (setmaxnreg works per warpgroup)
constexpr int WS_PRODUCER_THREADS = 32 * WS_PRODUCER_WARPS;
constexpr int WS_THREADS = 32 * (WS_CONSUMER_WARPS + WS_PRODUCER_WARPS);
constexpr int WS_ROW_BYTES = 128;            // BK doubles
constexpr int WS_A_STAGE = BM * WS_ROW_BYTES;  // 16 KB
// The 16 row blocks (8 rows each) of a 128-row tile are dealt round-robin to the four M warps (= the four SM
// sub-partitions): block b belongs to warp b % 4, so a ragged tile (e.g. 80 live rows on the last q tile of a
// Q shard) still loads all four DMMA pipes evenly instead of leaving two of them idle.
constexpr int WS_A_MB_STRIDE = 32 * WS_ROW_BYTES;  // warp's next row block is 4 blocks = 32 rows further
__device__ __forceinline__ int live_row_blocks(int rows_live, int wm) {
    int nblk = (min(max(rows_live, 0), BM) + 7) >> 3;  // live 8-row blocks in the tile
    return max(0, min(4, (nblk - wm + 3) >> 2));      // those with index = wm (mod 4)
}

template <int NB>
constexpr size_t ws_smem_bytes() {
    return 1024 + (size_t)WS_STAGES * (WS_A_STAGE + 16 * NB * WS_ROW_BYTES) + 2 * WS_STAGES * sizeof(uint64_t) +
           (WS_STAGES + 2) * sizeof(int);
}

// ---- mbarrier / TMA primitives (PTX ISA 8.x, sm_90+) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(n)
                 : "memory");
}
__device__ __forceinline__ void reg_dealloc_producer() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n"); }
__device__ __forceinline__ void reg_alloc_consumer() { asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n"); }
__device__ __forceinline__ void producer_bar_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(WS_PRODUCER_THREADS) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// Non-blocking probe of a phase (consumers look one stage ahead so that the probe's latency overlaps their DMMAs).
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// Producer-side wait: the producer is normally several stages ahead and would otherwise spin on try_wait every few
// tens of cycles, stealing issue slots from the two consumer warps that share its SM sub-partition.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok;
    for (;;) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(160);
    }
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

struct WsCarve {
    uint8_t* As;
    uint8_t* Bs;
    uint64_t* full;
    uint64_t* empty;
    volatile int* meta;   // [WS_STAGES] item id carried by the first stage of an item (-1 = no more work)
    volatile int* sched;  // [2] producer-group broadcast slots for the next item id
};
template <int NB>
__device__ __forceinline__ WsCarve ws_carve(uint8_t* raw) {
    WsCarve c;
    // offset arithmetic (not an integer round trip) so the compiler keeps the shared address space -> LDS, not LD
    uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    c.As = base;
    c.Bs = base + WS_STAGES * WS_A_STAGE;
    c.full = reinterpret_cast<uint64_t*>(c.Bs + WS_STAGES * 16 * NB * WS_ROW_BYTES);
    c.empty = c.full + WS_STAGES;
    c.meta = reinterpret_cast<volatile int*>(c.empty + WS_STAGES);
    c.sched = c.meta + WS_STAGES;
    return c;
}

// ---------------------------------------------------------------------------------------------
// Consumer main loop of one work item, specialised at compile time on this warp's live blocks (ML rows x NL columns)
// and software-pipelined: fragments are double-buffered in registers, the loads of k-step s+1 -- and, on the last
// k-step of a stage, of the NEXT stage's first k-step -- are issued before the DMMAs of k-step s, and the stage is
// released as soon as its last fragment has been read.  The two consumer warps of an SM sub-partition run in
// lock-step (they share the pipe and the stage ring), so without this the barrier-probe / address / LDS latency at
// the top of every stage left the DMMA pipe idle (~6 % in K3).
// Entry: the caller has waited for full[g % WS_STAGES].  Exit: g advanced by nkt.
// ---------------------------------------------------------------------------------------------
template <int NB, int ML, int NL>
__device__ __forceinline__ void consume_item(const WsCarve& sm, int b_stage_bytes, int a_row0, int b_row0,
                                             const int (&off)[4], double (&acc)[4][NB][2], int nkt, uint32_t& g, int lane) {
    if constexpr (ML == 0 || NL == 0) {
        // nothing to compute in this warp's corner of the tile: keep the ring protocol going
        for (int kt = 0; kt < nkt; kt++, g++) {
            const int s = g % WS_STAGES;
            if (kt > 0) mbar_wait(&sm.full[s], (g / WS_STAGES) & 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }
    } else {
        double a[2][ML], b[2][NL];
        int s = g % WS_STAGES;
        const uint8_t* ap = sm.As + s * WS_A_STAGE + a_row0;
        const uint8_t* bp = sm.Bs + s * b_stage_bytes + b_row0;
#pragma unroll
        for (int mb = 0; mb < ML; mb++) a[0][mb] = *reinterpret_cast<const double*>(ap + mb * WS_A_MB_STRIDE + off[0]);
#pragma unroll
        for (int nb = 0; nb < NL; nb++) b[0][nb] = *reinterpret_cast<const double*>(bp + nb * 8 * WS_ROW_BYTES + off[0]);
        for (int kt = 0; kt < nkt; kt++, g++) {
            const bool has_next = kt + 1 < nkt;
            const int sn = (g + 1) % WS_STAGES;
            const uint32_t pn = ((g + 1) / WS_STAGES) & 1;
            const uint32_t ready = has_next ? mbar_test(&sm.full[sn], pn) : 0;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const int cur = ks & 1, nxt = cur ^ 1;
                if (ks < 3) {
#pragma unroll
                    for (int mb = 0; mb < ML; mb++)
                        a[nxt][mb] = *reinterpret_cast<const double*>(ap + mb * WS_A_MB_STRIDE + off[ks + 1]);
#pragma unroll
                    for (int nb = 0; nb < NL; nb++)
                        b[nxt][nb] = *reinterpret_cast<const double*>(bp + nb * 8 * WS_ROW_BYTES + off[ks + 1]);
                } else {
                    // every fragment of this stage has been read: hand it back, then reach into the next stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[g % WS_STAGES]);
                    if (has_next) {
                        if (!ready) mbar_wait(&sm.full[sn], pn);
                        ap = sm.As + sn * WS_A_STAGE + a_row0;
                        bp = sm.Bs + sn * b_stage_bytes + b_row0;
#pragma unroll
                        for (int mb = 0; mb < ML; mb++)
                            a[nxt][mb] = *reinterpret_cast<const double*>(ap + mb * WS_A_MB_STRIDE + off[0]);
#pragma unroll
                        for (int nb = 0; nb < NL; nb++)
                            b[nxt][nb] = *reinterpret_cast<const double*>(bp + nb * 8 * WS_ROW_BYTES + off[0]);
                    }
                }
#pragma unroll
                for (int mb = 0; mb < ML; mb++)
#pragma unroll
                    for (int nb = 0; nb < NL; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[cur][mb], b[cur][nb]);
            }
        }
    }
}
template <int NB, int ML, int NL>
struct ItemDispatchN {
    static __device__ __forceinline__ void run(const WsCarve& sm, int bsb, int a_row0, int b_row0, const int (&off)[4],
                                               double (&acc)[4][NB][2], int nkt, uint32_t& g, int lane, int nbv) {
        if (nbv == NL)
            consume_item<NB, ML, NL>(sm, bsb, a_row0, b_row0, off, acc, nkt, g, lane);
        else
            ItemDispatchN<NB, ML, NL - 1>::run(sm, bsb, a_row0, b_row0, off, acc, nkt, g, lane, nbv);
    }
};
template <int NB, int ML>
struct ItemDispatchN<NB, ML, 0> {
    static __device__ __forceinline__ void run(const WsCarve& sm, int bsb, int a_row0, int b_row0, const int (&off)[4],
                                               double (&acc)[4][NB][2], int nkt, uint32_t& g, int lane, int) {
        consume_item<NB, 0, 0>(sm, bsb, a_row0, b_row0, off, acc, nkt, g, lane);
    }
};
// mbv / nbv are warp-uniform and constant over the item: one dispatch per item, not per stage.
template <int NB>
__device__ __forceinline__ void consume_item_any(const WsCarve& sm, int bsb, int a_row0, int b_row0, const int (&off)[4],
                                                 double (&acc)[4][NB][2], int nkt, uint32_t& g, int lane, int mbv, int nbv) {
    if (mbv == 4)
        ItemDispatchN<NB, 4, NB>::run(sm, bsb, a_row0, b_row0, off, acc, nkt, g, lane, nbv);
    else if (mbv == 3)
        ItemDispatchN<NB, 3, NB>::run(sm, bsb, a_row0, b_row0, off, acc, nkt, g, lane, nbv);
    else if (mbv == 2)
        ItemDispatchN<NB, 2, NB>::run(sm, bsb, a_row0, b_row0, off, acc, nkt, g, lane, nbv);
    else if (mbv == 1)
        ItemDispatchN<NB, 1, NB>::run(sm, bsb, a_row0, b_row0, off, acc, nkt, g, lane, nbv);
    else
        consume_item<NB, 0, 0>(sm, bsb, a_row0, b_row0, off, acc, nkt, g, lane);
}

// ---------------------------------------------------------------------------------------------
// Diagonal tile of a symmetric product (K4 with T2 == T1), upper triangle only.  In units of 8 x 8 blocks the tile is
// 16 x 16 and block (b, cb) is needed iff cb >= b: 136 of 256 blocks.  With the usual row-block deal (block b -> warp
// b % 4) the four SM sub-partitions would get 40 / 36 / 32 / 28 blocks; dealing the odd groups in reverse -- row
// blocks {wm, 7 - wm, 8 + wm, 15 - wm} -- gives every sub-partition exactly 34 (9 in its left-half warp, 25 in its
// right-half warp).  Everything about the shape is a compile-time function of (WM, WN), so the DMMA stream stays
// unpredicated; same software pipeline as consume_item.
// ---------------------------------------------------------------------------------------------
template <int WM, int WN>
struct DiagShape {
    static __host__ __device__ constexpr int rowsel(int mb) { return (mb & 1) ? 3 - WM : WM; }
    static __host__ __device__ constexpr int nb0(int mb) {
        const int b = mb * 4 + rowsel(mb) - 8 * WN;
        return b < 0 ? 0 : (b > 8 ? 8 : b);
    }
    static __host__ __device__ constexpr bool row_live(int mb) { return nb0(mb) < 8; }
    static __host__ __device__ constexpr int a_off(int mb) { return mb * WS_A_MB_STRIDE + (rowsel(mb) - WM) * 8 * WS_ROW_BYTES; }
    static __host__ __device__ constexpr int nb_first() {
        int f = 8;
        for (int mb = 0; mb < 4; mb++) f = nb0(mb) < f ? nb0(mb) : f;
        return f;
    }
};
__device__ __forceinline__ int diag_row_of(int mb, int wm) { return mb * 4 + ((mb & 1) ? 3 - wm : wm); }  // 8-row block

template <int WM, int WN>
__device__ __forceinline__ void consume_item_diag(const WsCarve& sm, int b_stage_bytes, int a_row0, int b_row0,
                                                  const int (&off)[4], double (&acc)[4][8][2], int nkt, uint32_t& g, int lane) {
    using S = DiagShape<WM, WN>;
    constexpr int NB = 8;
    constexpr int NF = S::nb_first();
    double a[2][4], b[2][NB];
    int s = g % WS_STAGES;
    const uint8_t* ap = sm.As + s * WS_A_STAGE + a_row0;
    const uint8_t* bp = sm.Bs + s * b_stage_bytes + b_row0;
    auto load = [&](int buf, int o) {
#pragma unroll
        for (int mb = 0; mb < 4; mb++)
            if (S::row_live(mb)) a[buf][mb] = *reinterpret_cast<const double*>(ap + S::a_off(mb) + o);
#pragma unroll
        for (int nb = NF; nb < NB; nb++) b[buf][nb] = *reinterpret_cast<const double*>(bp + nb * 8 * WS_ROW_BYTES + o);
    };
    load(0, off[0]);
    for (int kt = 0; kt < nkt; kt++, g++) {
        const bool has_next = kt + 1 < nkt;
        const int sn = (g + 1) % WS_STAGES;
        const uint32_t pn = ((g + 1) / WS_STAGES) & 1;
        const uint32_t ready = has_next ? mbar_test(&sm.full[sn], pn) : 0;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const int cur = ks & 1, nxt = cur ^ 1;
            if (ks < 3) {
                load(nxt, off[ks + 1]);
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[g % WS_STAGES]);
                if (has_next) {
                    if (!ready) mbar_wait(&sm.full[sn], pn);
                    ap = sm.As + sn * WS_A_STAGE + a_row0;
                    bp = sm.Bs + sn * b_stage_bytes + b_row0;
                    load(nxt, off[0]);
                }
            }
#pragma unroll
            for (int mb = 0; mb < 4; mb++)
#pragma unroll
                for (int nb = 0; nb < NB; nb++)
                    if (nb >= S::nb0(mb)) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[cur][mb], b[cur][nb]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3, persistent.  Work item w -> (it, qt, m): it fastest so the CTAs sharing an A tile run together.
// ---------------------------------------------------------------------------------------------
struct HalfWsParams {
    const CUtensorMap* amaps;   // [nbf] per-row-block maps of the shard tensor: dims {sp(m), nq}, pitch ld(m)
    const int* sp;              // [nbf]
    const int* cols;            // kept-partner lists
    const size_t* cols_off;     // [nbf]
    const double* Ct;           // [o x ldc] (gather source)
    int ldc;
    int o, op, iw;
    int nit, nqt;               // i tiles, q tiles per row-block
    int qbeg, qc;               // chunk of the shard's Q rows
    int nbf;
    int nitems;
    int* counter;               // work queue head (zeroed before launch)
    double* T;                  // [nbf][qc][op]
    // Fused first J sweep (dfhelper.cc:3193 / :3258): the last orbital tile carries one extra B-operand row, the
    // density row D'[m, n_k], so its accumulator column is d_part[m][q] = sum_k B_m[q,k] D'[m,n_k] -- the tensor
    // tile is already in shared memory, the separate HBM sweep of j_dq_kernel disappears.
    int fuse;                   // 0 / 1
    int rd;                     // row of the density inside the last i-tile (= o - (nit-1)*iw < 16*NB)
    const double* Dm;           // [nbf][ldd]  D (general) or D symmetrised from its upper triangle (lr_symmetric)
    int ldd;
    double* dpart;              // [nbf][dstride]
    int dstride;
    // Pre-gathered C^T (gather_ct_kernel): per-row-block maps {sp(m), o+1} over the packed copy, box 16 x 16*NB.  When
    // set, screened row-blocks take the all-TMA path too (the density row is row o of the packed block, i.e. tile row rd).
    const CUtensorMap* cgmaps;
};

template <int NB>
__global__ void __launch_bounds__(WS_THREADS, 1)
    half_ws_kernel(const __grid_constant__ CUtensorMap ctmap, const __grid_constant__ CUtensorMap ctmap_last,
                   const __grid_constant__ CUtensorMap dmap, HalfWsParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int BN = 16 * NB;
    constexpr int B_STAGE = BN * WS_ROW_BYTES;
    WsCarve sm = ws_carve<NB>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < WS_STAGES; s++) {
            mbar_init(&sm.full[s], 1 + WS_PRODUCER_THREADS);  // expect_tx arrive + one arrive per producer thread
            mbar_init(&sm.empty[s], WS_CONSUMER_WARPS);       // one arrive per consumer warp
        }
        mbar_fence_init();
        sm.sched[0] = atomicAdd(p.counter, 1);
    }
    __syncthreads();

    if (warp < WS_PRODUCER_WARPS) {
        // ================= producer warpgroup =================
        reg_dealloc_producer();
        if (tid == 0) prefetch_tmap(&ctmap);
        uint32_t g = 0;  // global stage counter
        int slot = 0;
        int w = sm.sched[0];
        while (w < p.nitems) {
            int w_next = 0;
            if (tid == 0) w_next = atomicAdd(p.counter, 1);  // latency hidden behind this item's k loop
            const int it = w % p.nit;
            const int qt = (w / p.nit) % p.nqt;
            const int m = w / (p.nit * p.nqt);
            const int K = p.sp[m];
            const int nkt = (K + BK - 1) / BK;
            const bool dense = (K == p.nbf);
            const int i0 = it * p.iw;
            const CUtensorMap* amap = p.amaps + m;
            const bool fused = p.fuse && it == p.nit - 1;  // this tile also carries the density row (row p.rd)
            if (dense) {
                // both operands by TMA; one thread drives the whole stage
                if (tid == 0) {
                    for (int kt = 0; kt < nkt; kt++) {
                        const uint32_t gg = g + kt;
                        const int s = gg % WS_STAGES;
                        mbar_wait_backoff(&sm.empty[s], ((gg / WS_STAGES) & 1) ^ 1);
                        if (kt == 0) sm.meta[s] = w;
                        if (!fused) {
                            mbar_arrive_expect_tx(&sm.full[s], WS_A_STAGE + B_STAGE);
                            tma_load_2d(sm.As + s * WS_A_STAGE, amap, kt * BK, p.qbeg + qt * BM, &sm.full[s]);
                            tma_load_2d(sm.Bs + s * B_STAGE, &ctmap, kt * BK, i0, &sm.full[s]);
                        } else {
                            // C^T box stops at row rd (its own map) so that it cannot race with the density row;
                            // rows above rd keep stale finite-or-not data that only feeds discarded columns
                            mbar_arrive_expect_tx(&sm.full[s], WS_A_STAGE + (p.rd + 1) * WS_ROW_BYTES);
                            tma_load_2d(sm.As + s * WS_A_STAGE, amap, kt * BK, p.qbeg + qt * BM, &sm.full[s]);
                            tma_load_2d(sm.Bs + s * B_STAGE, &ctmap_last, kt * BK, i0, &sm.full[s]);
                            tma_load_2d(sm.Bs + s * B_STAGE + p.rd * WS_ROW_BYTES, &dmap, kt * BK, m, &sm.full[s]);
                        }
                        mbar_arrive_n(&sm.full[s], WS_PRODUCER_THREADS);
                    }
                }
                g += nkt;
            } else if (p.cgmaps) {
                // screened row-block, C^T pre-gathered into the tensor's column order: both operands by TMA
                if (tid == 0) {
                    const CUtensorMap* cmap = p.cgmaps + m;
                    for (int kt = 0; kt < nkt; kt++) {
                        const uint32_t gg = g + kt;
                        const int s = gg % WS_STAGES;
                        mbar_wait_backoff(&sm.empty[s], ((gg / WS_STAGES) & 1) ^ 1);
                        if (kt == 0) sm.meta[s] = w;
                        mbar_arrive_expect_tx(&sm.full[s], WS_A_STAGE + B_STAGE);
                        tma_load_2d(sm.As + s * WS_A_STAGE, amap, kt * BK, p.qbeg + qt * BM, &sm.full[s]);
                        tma_load_2d(sm.Bs + s * B_STAGE, cmap, kt * BK, i0, &sm.full[s]);
                        mbar_arrive_n(&sm.full[s], WS_PRODUCER_THREADS);
                    }
                }
                g += nkt;
            } else {
                const int irows = min(p.o, i0 + BN) - i0;  // valid C^T rows of this tile
                const int* cols = p.cols + p.cols_off[m];
                const double* drow = p.Dm + (size_t)m * p.ldd;
                // Each thread owns one PAIR of k positions (2*pj, 2*pj+1) of the stage = one 16-byte chunk of every
                // tile row.  Kept partners come in runs (the functions of an atom), so most pairs are two adjacent,
                // 16-byte-aligned columns of C^T and go with ONE 16-byte cp.async; the others with two 8-byte ones.
                const int pj = tid & 7;
                int c0n = (2 * pj < K) ? __ldg(cols + 2 * pj) : -1;
                int c1n = (2 * pj + 1 < K) ? __ldg(cols + 2 * pj + 1) : -1;
                for (int kt = 0; kt < nkt; kt++, g++) {
                    const int s = g % WS_STAGES;
                    const uint32_t ph = (g / WS_STAGES) & 1;
                    mbar_wait(&sm.empty[s], ph ^ 1);
                    if (tid == 0) {
                        if (kt == 0) sm.meta[s] = w;
                        mbar_arrive_expect_tx(&sm.full[s], WS_A_STAGE);
                        tma_load_2d(sm.As + s * WS_A_STAGE, amap, kt * BK, p.qbeg + qt * BM, &sm.full[s]);
                    }
                    const int c0 = c0n, c1 = c1n;
                    const int kn = (kt + 1) * BK + 2 * pj;
                    c0n = (kt + 1 < nkt && kn < K) ? __ldg(cols + kn) : -1;
                    c1n = (kt + 1 < nkt && kn + 1 < K) ? __ldg(cols + kn + 1) : -1;
                    uint8_t* bs = sm.Bs + s * B_STAGE;
                    if (c0 < 0 || (!(c0 & 1) && (c1 == c0 + 1 || c1 < 0))) {
                        const int nb = c0 < 0 ? 0 : (c1 < 0 ? 8 : 16);  // the rest of the chunk is zero-filled
#pragma unroll 3
                        for (int r = tid >> 3; r < BN; r += WS_PRODUCER_THREADS / 8) {
                            const bool drow_here = fused && r == p.rd;
                            const int bytes = (r < irows || drow_here) ? nb : 0;
                            const double* src =
                                !bytes ? p.Ct : (drow_here ? drow + c0 : p.Ct + (size_t)(i0 + r) * p.ldc + c0);
                            double* dst = reinterpret_cast<double*>(bs + r * WS_ROW_BYTES + ((pj ^ (r & 7)) << 4));
                            cp_async16_ca(dst, src, bytes);
                        }
                    } else {
#pragma unroll 3
                        for (int r = tid >> 3; r < BN; r += WS_PRODUCER_THREADS / 8) {
                            const bool drow_here = fused && r == p.rd;
                            const bool live = r < irows || drow_here;
                            const double* base = drow_here ? drow : p.Ct + (size_t)(i0 + r) * p.ldc;
                            double* dst = reinterpret_cast<double*>(bs + r * WS_ROW_BYTES + ((pj ^ (r & 7)) << 4));
                            cp_async8(dst, live ? base + c0 : p.Ct, live ? 8 : 0);  // c0 >= 0 here
                            const bool l1 = live && c1 >= 0;
                            cp_async8(dst + 1, l1 ? base + c1 : p.Ct, l1 ? 8 : 0);
                        }
                    }
                    cp_async_mbar_arrive_noinc(&sm.full[s]);
                }
            }
            if (tid == 0) sm.sched[slot ^ 1] = w_next;
            producer_bar_sync();
            slot ^= 1;
            w = sm.sched[slot];
        }
        if (tid == 0) {  // sentinel stage: tells every consumer warp to stop
            const int s = g % WS_STAGES;
            mbar_wait_backoff(&sm.empty[s], ((g / WS_STAGES) & 1) ^ 1);
            sm.meta[s] = -1;
            mbar_arrive_n(&sm.full[s], 1 + WS_PRODUCER_THREADS);
        }
        cp_async_wait<0>();
        return;
    }

    // ================= consumers =================
    reg_alloc_consumer();
    const int cw = warp - WS_PRODUCER_WARPS;
    const int wm = cw & 3, wn = cw >> 2, gq = lane >> 2, t = lane & 3;
    int off[4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) off[ks] = (((ks + 4 * (t >> 1)) ^ gq) << 4) + ((t & 1) << 3);
    const int a_row0 = (wm * 8 + gq) * WS_ROW_BYTES;  // row blocks interleaved over the four M warps
    const int b_row0 = (wn * 8 * NB + gq) * WS_ROW_BYTES;
    uint32_t g = 0;
    for (;;) {
        int s = g % WS_STAGES;
        mbar_wait(&sm.full[s], (g / WS_STAGES) & 1);
        const int w = sm.meta[s];
        if (w < 0) break;
        const int it = w % p.nit;
        const int qt = (w / p.nit) % p.nqt;
        const int m = w / (p.nit * p.nqt);
        const int K = p.sp[m];
        const int nkt = (K + BK - 1) / BK;
        const int q0 = qt * BM, i0 = it * p.iw;
        const bool fused = p.fuse && it == p.nit - 1;
        const int icols = min(p.iw, p.o - i0) + (fused ? 1 : 0);  // + the density column
        const int mbv = live_row_blocks(p.qc - q0, wm);
        const int nbv = max(0, min(NB, (icols - wn * 8 * NB + 7) / 8));
        double acc[4][NB][2];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < NB; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

        consume_item_any<NB>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane, mbv, nbv);

        double* Tm = p.T + (size_t)m * p.qc * p.op;
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
            int q = q0 + mb * 32 + wm * 8 + gq;
            if (q < p.qc) {
#pragma unroll
                for (int nb = 0; nb < NB; nb++) {
                    int ic = wn * 8 * NB + nb * 8 + t * 2;
                    int i = i0 + ic;
                    double v0 = acc[mb][nb][0], v1 = acc[mb][nb][1];
                    if (fused && (ic == p.rd || ic + 1 == p.rd)) {
                        // density column: goes to d_part, and is NOT part of T (it may sit on T's zero pad column)
                        p.dpart[(size_t)m * p.dstride + p.qbeg + q] = (ic == p.rd) ? v0 : v1;
                        if (ic == p.rd) v0 = 0.0; else v1 = 0.0;
                    }
                    if (ic < p.iw && i < p.op) *reinterpret_cast<double2*>(Tm + (size_t)q * p.op + i) = make_double2(v0, v1);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4, persistent.  Work item w -> (tile = w % ntiles, split = w / ntiles): the tiles of one split run
// together (L2 reuse of T); `tiles` lists (tm, tn) by decreasing live area so the cheap edge tiles of the
// last split fill the tail.
// ---------------------------------------------------------------------------------------------
struct KgemmWsParams {
    int nbf, kdim, klen, ntiles, nitems;
    int symmetric;      // T2 == T1: only tiles with tn >= tm are listed; diagonal tiles skip their lower-left quadrant
    const int2* tiles;  // [ntiles] (tm, tn)
    int* counter;
    double* ws;         // [nsplit][ntiles][128*128]
    // Split-K reduction folded into the kernel: arrive[tile*8 + cw] counts the splits whose consumer warp cw has
    // stored its part of the partial tile; the warp that makes it nsplit sums that part over all splits in split
    // order (fixed order: deterministic, whoever arrives last) and accumulates it into K (beta = 1 across Q chunks),
    // mirrored for T2 == T1.  nullptr: partials only, kgemm_reduce_list_kernel follows (B200JK_KREDUCE=separate).
    int* arrive;        // [ntiles * 8], zeroed before launch
    int nsplit;
    int tri;            // triangular consumer on whole diagonal tiles (B200JK_KTRI=0 turns it off for A/B runs)
    double* K;          // [nbf][nbf]
};

// Count consumer warp cw's part of item w's partial tile as stored; true if that makes the tile's part complete
// (every split has arrived), i.e. this warp must now sum it.
__device__ __forceinline__ bool kgemm_arrive(const KgemmWsParams& p, int w, int cw, int lane) {
    __threadfence();  // this lane's part of the partial tile is visible device-wide ...
    __syncwarp();
    int last = 0;
    if (lane == 0) last = (atomicAdd(p.arrive + (w % p.ntiles) * WS_CONSUMER_WARPS + cw, 1) == p.nsplit - 1) ? 1 : 0;
    return __shfl_sync(0xffffffffu, last, 0) != 0;  // ... before the warp is counted
}

// Consumer warp cw's 32 x 64 part of the tile of item w, summed over the splits in split order (fixed order: the
// result does not depend on who arrives last), accumulated into K (beta = 1 across Q chunks) and mirrored when
// T2 == T1.  One 8-row group at a time with eight splits' worth of loads (64 x 16 bytes per lane) in flight: this
// runs when the warp's accumulators are dead.  Dead edge columns of a partial tile hold zeros, so only the stores
// are predicated.
__device__ __noinline__ void kgemm_reduce_part(const KgemmWsParams& p, int w, int cw, int lane) {
    constexpr int NB = 8, BN = 128;
    const int wm = cw & 3, wn = cw >> 2, gq = lane >> 2, t = lane & 3;
    const int tile = w % p.ntiles;
    const int2 tl = p.tiles[tile];
    int mbv = live_row_blocks(p.nbf - tl.x * BM, wm);
    const int nbv = max(0, min(NB, (p.nbf - (tl.y * BN + wn * 8 * NB) + 7) / 8));
    const bool diag = p.symmetric && tl.x == tl.y;
    const bool tri = diag && p.tri && (tl.x + 1) * BM <= p.nbf;
    if (diag && !tri && wn == 0) mbv = min(mbv, 2);
    __threadfence();
    const double* src0 = p.ws + (size_t)tile * (BM * BN) + wn * 8 * NB + t * 2;
    const size_t sstride = (size_t)p.ntiles * (BM * BN);
    constexpr int U = 8;
#pragma unroll 1
    for (int mb = 0; mb < mbv; mb++) {
        const int r = (tri ? diag_row_of(mb, wm) : mb * 4 + wm) * 8 + gq;
        const double* base = src0 + r * BN;
        double2 sum[NB];
#pragma unroll
        for (int nb = 0; nb < NB; nb++) sum[nb] = make_double2(0.0, 0.0);
        int sp = 0;
#pragma unroll 1
        for (; sp + U <= p.nsplit; sp += U) {
            double2 v[U][NB];
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int nb = 0; nb < NB; nb++)
                    v[u][nb] = __ldcg(reinterpret_cast<const double2*>(base + (size_t)(sp + u) * sstride + nb * 8));
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int nb = 0; nb < NB; nb++) {
                    sum[nb].x += v[u][nb].x;
                    sum[nb].y += v[u][nb].y;
                }
        }
#pragma unroll 1
        for (; sp < p.nsplit; sp++) {
#pragma unroll
            for (int nb = 0; nb < NB; nb++) {
                const double2 v = __ldcg(reinterpret_cast<const double2*>(base + (size_t)sp * sstride + nb * 8));
                sum[nb].x += v.x;
                sum[nb].y += v.y;
            }
        }
        const int m = tl.x * BM + r;
        if (m < p.nbf) {
#pragma unroll
            for (int nb = 0; nb < NB; nb++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int c = wn * 8 * NB + nb * 8 + t * 2 + e;
                    const int n = tl.y * BN + c;
                    if (nb < nbv && n < p.nbf && (!diag || c >= r)) {
                        const double v = e ? sum[nb].y : sum[nb].x;
                        p.K[(size_t)m * p.nbf + n] += v;
                        if (p.symmetric && n != m) p.K[(size_t)n * p.nbf + m] += v;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(WS_THREADS, 1)
    kgemm_ws_kernel(const __grid_constant__ CUtensorMap t1map, const __grid_constant__ CUtensorMap t2map, KgemmWsParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int NB = 8, BN = 128;
    constexpr int B_STAGE = BN * WS_ROW_BYTES;
    WsCarve sm = ws_carve<NB>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < WS_STAGES; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], WS_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < WS_PRODUCER_WARPS) {
        reg_dealloc_producer();
        if (tid == 0) {
            prefetch_tmap(&t1map);
            prefetch_tmap(&t2map);
            uint32_t g = 0;
            int w = atomicAdd(p.counter, 1);
            while (w < p.nitems) {
                const int w_next = atomicAdd(p.counter, 1);
                const int2 tl = p.tiles[w % p.ntiles];
                const int kb = (w / p.ntiles) * p.klen;
                const int ke = min(p.kdim, kb + p.klen);
                const int nkt = (ke - kb + BK - 1) / BK;
                for (int kt = 0; kt < nkt; kt++, g++) {
                    const int s = g % WS_STAGES;
                    mbar_wait_backoff(&sm.empty[s], ((g / WS_STAGES) & 1) ^ 1);
                    if (kt == 0) sm.meta[s] = w;
                    mbar_arrive_expect_tx(&sm.full[s], WS_A_STAGE + B_STAGE);
                    tma_load_2d(sm.As + s * WS_A_STAGE, &t1map, kb + kt * BK, tl.x * BM, &sm.full[s]);
                    tma_load_2d(sm.Bs + s * B_STAGE, &t2map, kb + kt * BK, tl.y * BN, &sm.full[s]);
                }
                w = w_next;
            }
            const int s = g % WS_STAGES;
            mbar_wait_backoff(&sm.empty[s], ((g / WS_STAGES) & 1) ^ 1);
            sm.meta[s] = -1;
            mbar_arrive(&sm.full[s]);
        }
        return;
    }

    reg_alloc_consumer();
    const int cw = warp - WS_PRODUCER_WARPS;
    const int wm = cw & 3, wn = cw >> 2, gq = lane >> 2, t = lane & 3;
    int off[4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) off[ks] = (((ks + 4 * (t >> 1)) ^ gq) << 4) + ((t & 1) << 3);
    const int a_row0 = (wm * 8 + gq) * WS_ROW_BYTES;  // row blocks interleaved over the four M warps
    const int b_row0 = (wn * 8 * NB + gq) * WS_ROW_BYTES;
    uint32_t g = 0;
    int pending = -1;  // the item whose partial tile this warp has stored but not yet counted
    for (;;) {
        int s = g % WS_STAGES;
        mbar_wait(&sm.full[s], (g / WS_STAGES) & 1);
        const int w = sm.meta[s];
        if (w < 0) break;
        const int2 tl = p.tiles[w % p.ntiles];
        const int kb = (w / p.ntiles) * p.klen;
        const int ke = min(p.kdim, kb + p.klen);
        const int nkt = (ke - kb + BK - 1) / BK;
        int mbv = live_row_blocks(p.nbf - tl.x * BM, wm);
        const int nbv = max(0, min(NB, (p.nbf - (tl.y * BN + wn * 8 * NB) + 7) / 8));
        // Diagonal tile of a symmetric product: its lower-left 64 x 64 quadrant (rows 64.., columns ..63) is the mirror
        // of the upper-right one and is not computed -- the warps of the left column half stop after their first two
        // row groups.  (The diagonal quadrants are still computed whole; only their upper triangles are kept.)
        const bool diag = p.symmetric && tl.x == tl.y;
        // a diagonal tile that lies wholly inside the matrix takes the triangular consumer (34 of 64 block products per
        // sub-partition instead of 48); the ragged last one keeps the quadrant rule
        const bool tri = diag && p.tri && (tl.x + 1) * BM <= p.nbf;
        if (diag && !tri && wn == 0) mbv = min(mbv, 2);
        double acc[4][NB][2];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < NB; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
        if (tri) {
            switch (cw) {
                case 0: consume_item_diag<0, 0>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane); break;
                case 1: consume_item_diag<1, 0>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane); break;
                case 2: consume_item_diag<2, 0>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane); break;
                case 3: consume_item_diag<3, 0>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane); break;
                case 4: consume_item_diag<0, 1>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane); break;
                case 5: consume_item_diag<1, 1>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane); break;
                case 6: consume_item_diag<2, 1>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane); break;
                default: consume_item_diag<3, 1>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane); break;
            }
        } else {
            consume_item_any<NB>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane, mbv, nbv);
        }
        // The PREVIOUS item of this warp is counted now, one whole item after its partial tile was stored: the fence
        // finds those stores long drained.  (Fencing right behind the stores held the DMMA pipe of this sub-partition
        // idle for the drain time of 128 KB, once per item.)
        const bool reduce_prev = p.arrive && pending >= 0 && kgemm_arrive(p, pending, cw, lane);
        double* wsp = p.ws + (size_t)w * (BM * BN);
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
            int r = (tri ? diag_row_of(mb, wm) : mb * 4 + wm) * 8 + gq;
#pragma unroll
            for (int nb = 0; nb < NB; nb++) {
                int c = wn * 8 * NB + nb * 8 + t * 2;
                *reinterpret_cast<double2*>(wsp + r * BN + c) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            }
        }
        if (reduce_prev) kgemm_reduce_part(p, pending, cw, lane);
        pending = w;
    }
    // the last item of this CTA: nothing follows it, retire it now
    if (p.arrive && pending >= 0 && kgemm_arrive(p, pending, cw, lane)) kgemm_reduce_part(p, pending, cw, lane);
}

// ---------------------------------------------------------------------------------------------
// Fitting GEMM (DFHelper::contract_metric_AO_core_symm, dfhelper.cc:1653-1678):
//   F[q, j] = sum_R Met[q, R] * U[j, R]      q in this shard's aux rows, j = staged (m, n>=m) pair column
// Same persistent pipeline as the K GEMM (both operands by TMA, contraction index R contiguous), no split-K;
// the epilogue scatters column j straight into the packed pQq tensor: dst_off[j] + q * dst_ld[j].
// ---------------------------------------------------------------------------------------------
struct FitGemmParams {
    int nq, ncols, kdim;        // shard rows, staged pair columns, naux
    int ntm, ntn, nitems;       // tiles along q, along j
    const size_t* dst_off;      // [ncols] offset (doubles, from tensor base) of element (q=0) of column j
    const int* dst_ld;          // [ncols] row pitch of the row-block column j belongs to
    double* tensor;
    int* counter;
};

__global__ void __launch_bounds__(WS_THREADS, 1)
    fit_gemm_ws_kernel(const __grid_constant__ CUtensorMap metmap, const __grid_constant__ CUtensorMap umap, FitGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int NB = 8, BN = 128;
    constexpr int B_STAGE = BN * WS_ROW_BYTES;
    WsCarve sm = ws_carve<NB>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nkt = (p.kdim + BK - 1) / BK;
    if (tid == 0) {
        for (int s = 0; s < WS_STAGES; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], WS_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < WS_PRODUCER_WARPS) {
        reg_dealloc_producer();
        if (tid == 0) {
            prefetch_tmap(&metmap);
            prefetch_tmap(&umap);
            uint32_t g = 0;
            int w = atomicAdd(p.counter, 1);
            while (w < p.nitems) {
                const int w_next = atomicAdd(p.counter, 1);
                const int tm = w % p.ntm, tn = w / p.ntm;  // q tiles fastest: CTAs in flight share the U tile
                for (int kt = 0; kt < nkt; kt++, g++) {
                    const int s = g % WS_STAGES;
                    mbar_wait_backoff(&sm.empty[s], ((g / WS_STAGES) & 1) ^ 1);
                    if (kt == 0) sm.meta[s] = w;
                    mbar_arrive_expect_tx(&sm.full[s], WS_A_STAGE + B_STAGE);
                    tma_load_2d(sm.As + s * WS_A_STAGE, &metmap, kt * BK, tm * BM, &sm.full[s]);
                    tma_load_2d(sm.Bs + s * B_STAGE, &umap, kt * BK, tn * BN, &sm.full[s]);
                }
                w = w_next;
            }
            const int s = g % WS_STAGES;
            mbar_wait_backoff(&sm.empty[s], ((g / WS_STAGES) & 1) ^ 1);
            sm.meta[s] = -1;
            mbar_arrive(&sm.full[s]);
        }
        return;
    }

    reg_alloc_consumer();
    const int cw = warp - WS_PRODUCER_WARPS;
    const int wm = cw & 3, wn = cw >> 2, gq = lane >> 2, t = lane & 3;
    int off[4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) off[ks] = (((ks + 4 * (t >> 1)) ^ gq) << 4) + ((t & 1) << 3);
    const int a_row0 = (wm * 8 + gq) * WS_ROW_BYTES;
    const int b_row0 = (wn * 8 * NB + gq) * WS_ROW_BYTES;
    uint32_t g = 0;
    for (;;) {
        int s = g % WS_STAGES;
        mbar_wait(&sm.full[s], (g / WS_STAGES) & 1);
        const int w = sm.meta[s];
        if (w < 0) break;
        const int tm = w % p.ntm, tn = w / p.ntm;
        const int mbv = live_row_blocks(p.nq - tm * BM, wm);
        const int nbv = max(0, min(NB, (p.ncols - (tn * BN + wn * 8 * NB) + 7) / 8));
        double acc[4][NB][2];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < NB; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
        consume_item_any<NB>(sm, B_STAGE, a_row0, b_row0, off, acc, nkt, g, lane, mbv, nbv);
#pragma unroll
        for (int nb = 0; nb < NB; nb++) {
            const int j0 = tn * BN + wn * 8 * NB + nb * 8 + t * 2;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = j0 + e;
                if (j < p.ncols) {
                    double* dst = p.tensor + p.dst_off[j];
                    const size_t ld = (size_t)p.dst_ld[j];
#pragma unroll
                    for (int mb = 0; mb < 4; mb++) {
                        const int q = tm * BM + mb * 32 + wm * 8 + gq;
                        if (q < p.nq) dst[(size_t)q * ld] = acc[mb][nb][e];
                    }
                }
            }
        }
    }
}

// K[m][n] += sum_s ws[s][tile][r][c]  (fixed order); mirrored for off-diagonal tiles when symmetric.
__global__ void kgemm_reduce_list_kernel(const double* __restrict__ ws, int nsplit, int ntiles,
                                         const int2* __restrict__ tiles, int symmetric, int nbf, double* __restrict__ K) {
    const int tile = blockIdx.x;
    const int2 tl = tiles[tile];
    for (int e = threadIdx.x; e < BM * 128; e += blockDim.x) {
        int r = e >> 7, c = e & 127;
        int m = tl.x * BM + r, n = tl.y * 128 + c;
        if (m < nbf && n < nbf) {
            double s = 0.0;
            for (int sp = 0; sp < nsplit; sp++) s += ws[((size_t)sp * ntiles + tile) * (BM * 128) + e];
            if (symmetric && tl.y == tl.x) {
                // diagonal tile: the upper triangle is authoritative, the lower one is its mirror
                if (c >= r) {
                    K[(size_t)m * nbf + n] += s;
                    if (c > r) K[(size_t)n * nbf + m] += s;
                }
            } else {
                K[(size_t)m * nbf + n] += s;
                if (symmetric) K[(size_t)n * nbf + m] += s;
            }
        }
    }
}

}  // namespace b2k
