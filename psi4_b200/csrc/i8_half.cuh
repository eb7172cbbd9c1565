// i8_half.cuh -- the half transform on the INT8 tensor cores: an alternative arm of K3
//   T[m,q,i] = sum_k B_m[q,k] * C[n_k(m), i]      (DFHelper::first_transform_pQq, lib3index/dfhelper.cc:2162-2186)
// next to half_ws_kernel (DMMA), built from the residue machinery of i8_kgemm.cuh.
//
// Per row-block m the product is (nq x sp(m)) x (sp(m) x nocc): the big operand is the resident tensor itself, touched
// once with only nocc multiply-adds per element, so the conversion to residues is what has to be cheap:
//   once per tensor   every row (m, q) of the packed tensor gets a power-of-two scale 2^eB(m,q), |B'| <= R (row 2-norm);
//   once per build    the columns of C get scales 2^eC(i) (full column norm: a bound for every kept-partner subset) and
//                     the residue planes of C^T, rc[j][i][n];
//   per chunk of row-blocks (as many as the scratch arena holds):
//     convert   f64 rows -> residue planes, stored as the GEMM reads them: per row-block and q-tile a run of 16 KB
//               tiles (128 rows x 128 bytes of k), each already in the 128-byte-swizzled order of the shared-memory
//               operand, so one 1-D bulk copy of 16 contiguous kilobytes fetches a stage (a tensor-map box over
//               row-major planes reads 128 separate 128-byte pieces: 3.5 TB/s instead of HBM speed);
//     gather    rc along the kept-partner list of every row-block -> the B operand in the same tiled, swizzled form;
//     GEMM      persistent tcgen05 kernel over (row-block, modulus, q-tile group, orbital tile); the CTAs of a cluster take
//               consecutive q-tiles and share the B tile by multicast; one k-range (sp(m) <= nbf), the accumulator is
//               reduced mod p_j and stored as bytes;
//     CRT       the bytes of all moduli -> the integer sum_k B' C' (floating-point CRT) -> * 2^-(eB + eC) -> T.
// Exact except for the scaling of the two operands: |dT[m,q,i]| <~ 2^-bits |B_m[q,:]| |C[:,i]|.
#pragma once
#include <stdlib.h>
#include <string.h>

#include "i8_kgemm.cuh"

namespace b2k {

constexpr int I8H_TILE = I8_TM * I8_BK;  // 16 KB: 128 rows x 128 bytes of k

// ---- row scales of the packed tensor (once per tensor) ------------------------------------------------------
// one warp per row (m, q); expo[m * nq + q] = largest e with |B_m[q,:]| 2^e <= Rb
__global__ void __launch_bounds__(256) i8h_rowscale_kernel(const double* __restrict__ tensor, const size_t* __restrict__ row_off,
                                                           const int* __restrict__ ldm, int nq, int nbf, double Rb,
                                                           int* __restrict__ expo) {
    const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long)nbf * nq) return;
    const int m = (int)(gw / nq), q = (int)(gw - (long)m * nq);
    const int ld = ldm[m];  // multiple of 4, pad columns are zero
    const double* src = tensor + row_off[m] + (size_t)q * ld;
    double s0 = 0, s1 = 0;
    for (int k = 2 * lane; k < ld; k += 64) {
        const double2 x = *reinterpret_cast<const double2*>(src + k);
        s0 = fma(x.x, x.x, s0);
        s1 = fma(x.y, x.y, s1);
    }
    double s = s0 + s1;
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
    if (lane == 0) {
        const double nrm = sqrt(s) * (1.0 + 1e-12);
        int ex = 0;
        if (nrm > 0 && isfinite(nrm)) {
            int t;
            frexp(Rb / nrm, &t);
            ex = t - 1;
            if (ldexp(nrm, ex) > Rb) ex--;
            ex = max(-1000, min(1000, ex));
        }
        expo[gw] = ex;
    }
}

// First J sweep riding on the conversion (the converter reads every f64 row of the tensor anyway):
//   dpart[m * dstride + q] = sum_k B_m[q,k] * Dm[m][n_k(m)]     (DGEMV 'N', dfhelper.cc:3193 / :3258; Dm as j_prep_dm_kernel
// makes it), lane partials in k order, xor-shuffle tree: a fixed order.  Dm == nullptr: plain conversion.
//
// gemm_col = 1: the sweep rides on the GEMM instead -- the density row of row-block m, as residues, is gathered into the first
// free column (index nocc) of the row-block's C^T tile, the GEMM computes it like any orbital, and the CRT kernel rebuilds
//   dpart[m][q] = sum_k B'_m[q,k] D'[m,n_k] * 2^-(eB(m,q) + eD(m))
// from the residues: exact in the integers, so the same bits whether the planes were converted in this build or cached from
// an earlier one (a sweep riding on the conversion cannot be used when the conversion is skipped).
struct I8HalfFuseJ {
    const double* Dm;
    int ldd;
    double* dpart;
    int dstride;
    const int* cols;
    const size_t* cols_off;
    int gemm_col;
};

// ---- residue planes of a chunk of row-blocks ------------------------------------------------------------------
// grid (ceil(qc / 32), nmc): one CTA = 32 rows q of one row-block m, one warp = four of them in turn.  Tile (m, t, kb) of
// plane j lies at  planes + j * plane_stride + ((kboff[m] - kboff[m0]) * nqt + t * nkb(m) + kb) * 16 KB,
// row r = q % 128 at r * 128, 16-byte chunk c of the row at chunk c ^ (r & 7) (TMA SWIZZLE_128B order).
// Fused first J sweep: the density row of m, gathered along the kept-partner list, is staged in shared memory once per
// CTA (gathering it per row through the LSU cost 20 ms of the 62 the C60 conversion took; the sweep it replaces costs 7.5).
constexpr int I8H_ROWS = 32;
template <int NMOD>
__global__ void __launch_bounds__(256) i8h_convert_kernel(const double* __restrict__ tensor, const size_t* __restrict__ row_off,
                                                          const int* __restrict__ ldm, const int* __restrict__ sp,
                                                          const int* __restrict__ kboff, const int* __restrict__ expo, int nq, int m0,
                                                          int qbeg, int qc, int nqt, int8_t* __restrict__ planes, size_t plane_stride,
                                                          const I8HalfFuseJ fj) {
    extern __shared__ double i8h_dg[];  // [nkb(m) * 128] density row in the tensor's column order, zero beyond sp(m)
    const int m = m0 + blockIdx.y;
    const int K = sp[m], nk = kboff[m + 1] - kboff[m];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool fuse = fj.Dm != nullptr;
    if (fuse) {
        const double* drow = fj.Dm + (size_t)m * fj.ldd;
        const int* cm = fj.cols + fj.cols_off[m];
        for (int k = threadIdx.x; k < nk * I8_BK; k += 256) i8h_dg[k] = k < K ? drow[__ldg(cm + k)] : 0.0;
        __syncthreads();
    }
    const double CM = 6755399441055744.0;  // 1.5 * 2^52 (see i8_convert_kernel: one FP64 FMA per element and modulus)
    const size_t tile0 = (size_t)(kboff[m] - kboff[m0]) * nqt;
    for (int rr = warp; rr < I8H_ROWS; rr += 8) {
        const int q = blockIdx.x * I8H_ROWS + rr;
        if (q >= qc) break;
        const int t = q >> 7, r = q & 127;
        const double* src = tensor + row_off[m] + (size_t)(qbeg + q) * ldm[m];
        int8_t* dst = planes + (tile0 + (size_t)t * nk) * I8H_TILE + r * I8_BK + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2);
        const double scale = ldexp(1.0, expo[(size_t)m * nq + qbeg + q]);
        int k = 4 * lane;
        double2 a = make_double2(0.0, 0.0), b = a;
        if (k < K) {
            a = *reinterpret_cast<const double2*>(src + k);
            b = *reinterpret_cast<const double2*>(src + k + 2);
        }
        double dq = 0.0;
        for (int kb = 0; kb < nk; kb++, k += 128) {
            double y[4] = {a.x, a.y, b.x, b.y};
            a = make_double2(0.0, 0.0);
            b = a;
            if (k + 128 < K) {  // the next 32 bytes of the lane are requested before the current ones are converted
                a = *reinterpret_cast<const double2*>(src + k + 128);
                b = *reinterpret_cast<const double2*>(src + k + 130);
            }
            if (fuse) {
                const double2 d0 = *reinterpret_cast<const double2*>(i8h_dg + k), d1 = *reinterpret_cast<const double2*>(i8h_dg + k + 2);
                dq = fma(y[0], d0.x, dq);
                dq = fma(y[1], d0.y, dq);
                dq = fma(y[2], d1.x, dq);
                dq = fma(y[3], d1.y, dq);
            }
            int ylo[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                y[c] = __dadd_rn(__fma_rn(y[c], scale, CM), -CM);
                ylo[c] = __double2loint(__dadd_rn(y[c], CM));
            }
            int8_t* d = dst + (size_t)kb * I8H_TILE;
            *reinterpret_cast<uint32_t*>(d) = i8_pack_bytes(ylo[0], ylo[1], ylo[2], ylo[3]);
#pragma unroll
            for (int j = 1; j < NMOD; j++) {
                const int pj = c_i8.p[j];
                const double inv = c_i8.invpd[j];
                int rv[4];
#pragma unroll
                for (int c = 0; c < 4; c++) rv[c] = ylo[c] - pj * __double2loint(__fma_rn(y[c], inv, CM));
                *reinterpret_cast<uint32_t*>(d + (size_t)j * plane_stride) = i8_pack_bytes(rv[0], rv[1], rv[2], rv[3]);
            }
        }
        if (fuse) {
#pragma unroll
            for (int w = 16; w > 0; w >>= 1) dq += __shfl_xor_sync(0xffffffffu, dq, w);
            if (lane == 0) fj.dpart[(size_t)m * fj.dstride + qbeg + q] = dq;
        }
    }
}

// ---- gathered C^T of the chunk, tiled like the planes --------------------------------------------------------------------
// block (m, kb, il) of plane j: cg + j * cg_plane + ((kboff[m] - kboff[m0] + kb) * nit + il) * ntile_n * 128, row = orbital
// within the tile, byte kk = k % 128 at chunk (kk / 16) ^ (row & 7).  grid (nmc, ceil(opw / 4)); one thread = four
// consecutive k of four orbitals, all moduli; k beyond sp(m) and orbitals beyond nocc are written as zeros.
// rD / dcol: residue planes rD[j][m][n] of the density rows and the column (>= nocc) that carries row m of them in
// row-block m's tile (first J sweep as a column of the GEMM, I8HalfFuseJ::gemm_col); dcol < 0: none.
// The source planes (a few MB) live in L2 and every word costs two dependent accesses to it, so the loop over the moduli
// is unrolled at compile time: all 2 * NMOD loads of an orbital are in flight before the first store (with a run-time
// trip count ncu showed 26 % issue-active and 35 long-scoreboard stall cycles per instruction: 3.0 ms per build whatever
// the shard size, a fifth of the build on the 592-row shard of the 8-GPU run).  Two variants through shared memory were
// measured and dropped: a 16-byte unit (kept-partner runs at C60 are mostly shorter than 16: 4.2 ms on its byte path) and
// a word unit with one staged modulus at a time (issue-bound on re-computed addresses: 5.9 ms).
template <int NMOD>
__global__ void __launch_bounds__(128) i8h_gather_kernel(const int8_t* __restrict__ rc, size_t rc_ld, size_t rc_plane, int o,
                                                         const int* __restrict__ sp, const int* __restrict__ kboff,
                                                         const int* __restrict__ cols, const size_t* __restrict__ cols_off, int m0,
                                                         int nit, int ntile_n, int8_t* __restrict__ cg, size_t cg_plane,
                                                         const int8_t* __restrict__ rD, size_t rd_plane, int dcol) {
    const int m = m0 + blockIdx.x;
    const int K = sp[m], nk = kboff[m + 1] - kboff[m];
    const int* c = cols + cols_off[m];
    const size_t blk = (size_t)ntile_n * I8_BK;
    int8_t* base = cg + (size_t)(kboff[m] - kboff[m0]) * nit * blk;
    const int i0 = blockIdx.y * 4, opw = nit * ntile_n;
    for (int k = 4 * threadIdx.x; k < nk * I8_BK; k += 4 * 128) {
        int n[4];
#pragma unroll
        for (int e = 0; e < 4; e++) n[e] = (k + e < K) ? __ldg(c + k + e) : -1;
        // kept partners come in runs: four consecutive columns are one unaligned 32-bit window of the rc row (two aligned
        // loads and a funnel shift instead of four byte loads)
        const bool run = n[0] >= 0 && n[1] == n[0] + 1 && n[2] == n[0] + 2 && n[3] == n[0] + 3;
        const int nw = n[0] & ~3, sh = (n[0] & 3) * 8;
        const int kb = k >> 7, kk = k & 127;
#pragma unroll
        for (int di = 0; di < 4; di++) {
            const int i = i0 + di;
            if (i >= opw) break;
            const int il = i / ntile_n, row = i - il * ntile_n;
            int8_t* d = base + ((size_t)kb * nit + il) * blk + row * I8_BK + (((kk >> 4) ^ (row & 7)) << 4) + (kk & 15);
            const bool isd = i == dcol;
            const int8_t* src = isd ? rD + (size_t)m * rc_ld : rc + (size_t)i * rc_ld;
            const size_t pstride = isd ? rd_plane : rc_plane;
            uint32_t w[NMOD];
            if ((i < o || isd) && n[0] >= 0) {
                if (run) {
                    uint32_t w0[NMOD], w1[NMOD];
#pragma unroll
                    for (int j = 0; j < NMOD; j++) {
                        w0[j] = *reinterpret_cast<const uint32_t*>(src + (size_t)j * pstride + nw);
                        w1[j] = sh ? *reinterpret_cast<const uint32_t*>(src + (size_t)j * pstride + nw + 4) : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < NMOD; j++) w[j] = __funnelshift_r(w0[j], w1[j], sh);
                } else {
#pragma unroll
                    for (int j = 0; j < NMOD; j++) {
                        uint32_t v = 0;
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            if (n[e] >= 0) v |= ((uint32_t)(uint8_t)src[(size_t)j * pstride + n[e]]) << (8 * e);
                        w[j] = v;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < NMOD; j++) w[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < NMOD; j++) *reinterpret_cast<uint32_t*>(d + (size_t)j * cg_plane) = w[j];
        }
    }
}
template <int NMOD>
inline void i8h_launch_gather(const int8_t* rc, size_t rc_ld, size_t rc_plane, int o, const int* sp, const int* kboff, const int* cols,
                              const size_t* cols_off, int m0, int nmc, int nit, int ntile_n, int8_t* cg, size_t cg_plane, const int8_t* rD,
                              size_t rd_plane, int dcol, cudaStream_t st) {
    i8h_gather_kernel<NMOD><<<dim3((unsigned)nmc, (unsigned)((nit * ntile_n + 3) / 4)), 128, 0, st>>>(rc, rc_ld, rc_plane, o, sp, kboff, cols, cols_off, m0,
                                                                                                    nit, ntile_n, cg, cg_plane, rD, rd_plane, dcol);
}

// ---- GEMM ------------------------------------------------------------------------------------------------------------------
struct I8HalfParams {
    int nitems;  // super-items: (row-block, modulus, q-tile group, orbital tile)
    int nmod, nqt, nit, ntile_n, m0, qc, opw;
    size_t ws_mod_stride;  // bytes between the moduli in ws = nmc * qc * opw
    size_t plane_stride, cg_plane;
    const int* kboff;
    const int8_t *planes, *cg;
    uint8_t* ws;  // [nmod][nmc][qc][opw]
};

// The gathered C^T tile (ntile_n x 128 bytes per k-block) is larger than the A tile and would be re-streamed from L2 for
// every q-tile; the CL CTAs of a cluster therefore take CL consecutive q-tiles of the same (row-block, modulus, orbital
// tile): each loads 1/CL of the B tile and multicasts it into the shared memory of all of them (completion counted on
// every CTA's own full barrier); a stage is refilled only when the MMAs of ALL CTAs have read it (tcgen05.commit
// multicast onto every CTA's empty barrier, count CL).  Static round-robin over the super-items: the A planes are
// streamed once whatever the order, so no queue is needed.
__device__ __forceinline__ uint32_t i8_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t i8_cluster_id() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t i8_ncluster() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void i8_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_load_mc(void* smem_dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}

template <int CL>
__global__ void __launch_bounds__(I8_THREADS, 1) i8h_gemm_kernel(const I8HalfParams p) {
    extern __shared__ uint8_t i8_raw[];
    uint8_t* base = i8_raw + ((1024u - (smem_u32(i8_raw) & 1023u)) & 1023u);
    uint8_t* As = base;
    uint8_t* Bs = base + I8_STAGES * I8_A_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(Bs + I8_STAGES * I8_B_STAGE);
    uint64_t* empty = full + I8_STAGES;
    uint64_t* tfull = empty + I8_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = CL > 1 ? (int)i8_cluster_ctarank() : 0;
    const int ncl = CL > 1 ? (int)i8_ncluster() : (int)gridDim.x, cid = CL > 1 ? (int)i8_cluster_id() : (int)blockIdx.x;
    if (tid == 0) {
        for (int s = 0; s < I8_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);  // one commit from the MMA thread of every CTA of the cluster
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) i8_cluster_sync();  // every CTA's barriers exist before anybody multicasts into them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint16_t mask = (uint16_t)((1u << CL) - 1u);
    const int nqg = (p.nqt + CL - 1) / CL;
    const int per_j = nqg * p.nit, per_m = p.nmod * per_j;
    const int nsuper = p.nitems;
    const uint32_t bblk = (uint32_t)p.ntile_n * I8_BK, slice = bblk / CL;
    const uint32_t bytes = I8_A_STAGE + bblk;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int it = cid; it < nsuper; it += ncl) {
                const int mloc = it / per_m, r = it - mloc * per_m;
                const int j = r / per_j, r2 = r - j * per_j;
                const int qg = r2 / p.nit, il = r2 - qg * p.nit;
                const int m = p.m0 + mloc;
                const int kb0 = p.kboff[m] - p.kboff[p.m0], nkb = p.kboff[m + 1] - p.kboff[m];
                const int qt = min(qg * CL + rank, p.nqt - 1);  // a CTA past the last q-tile recomputes it and stores nothing
                const int8_t* asrc = p.planes + (size_t)j * p.plane_stride + ((size_t)kb0 * p.nqt + (size_t)qt * nkb) * I8H_TILE;
                const int8_t* bsrc = p.cg + (size_t)j * p.cg_plane + ((size_t)kb0 * p.nit + il) * bblk + (size_t)rank * slice;
                for (int kb = 0; kb < nkb; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], bytes);
                    bulk_load(As + stage * I8_A_STAGE, asrc + (size_t)kb * I8H_TILE, I8H_TILE, &full[stage]);
                    if (CL > 1)
                        bulk_load_mc(Bs + stage * I8_B_STAGE + rank * slice, bsrc + (size_t)kb * p.nit * bblk, slice, &full[stage], mask);
                    else
                        bulk_load(Bs + stage * I8_B_STAGE, bsrc + (size_t)kb * p.nit * bblk, bblk, &full[stage]);
                    if (++stage == I8_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            // instruction descriptor: D = s32, A = B = s8, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.ntile_n >> 3) << 17) | ((128u >> 4) << 24);
            for (int it = cid; it < nsuper; it += ncl) {
                const int m = p.m0 + it / per_m;
                const int nkb = p.kboff[m + 1] - p.kboff[m];
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem + as * I8_TN;
                for (int kb = 0; kb < nkb; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t ad = i8_smem_desc(smem_u32(As + stage * I8_A_STAGE));
                    const uint64_t bd = i8_smem_desc(smem_u32(Bs + stage * I8_B_STAGE));
#pragma unroll
                    for (int k = 0; k < 4; k++) mma_i8_ss(dcol, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    if (CL > 1)
                        tc_commit_mc(&empty[stage], mask);  // every CTA of the cluster learns that this one has read the stage
                    else
                        tc_commit(&empty[stage]);
                    if (++stage == I8_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(&tfull[as]);
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else {
        uint32_t as = 0, aphase = 0;
        const int lane_base = (warp & 3) * 32;
        for (int it = cid; it < nsuper; it += ncl) {
            const int mloc = it / per_m, r = it - mloc * per_m;
            const int j = r / per_j, r2 = r - j * per_j;
            const int qg = r2 / p.nit, il = r2 - qg * p.nit;
            const int qt = qg * CL + rank;
            const int rows = qt < p.nqt ? min(I8_TM, p.qc - qt * I8_TM) : 0;
            const int ncols = p.ntile_n;
            const uint32_t pm = (uint32_t)c_i8.p[j], magic = c_i8.magic[j], off = c_i8.off[j];
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)lane_base << 16) + as * I8_TN;
            const int row = lane_base + lane;
            const bool live = row < rows;
            uint8_t* dst = p.ws + (size_t)j * p.ws_mod_stride + ((size_t)mloc * p.qc + (size_t)qt * I8_TM + row) * p.opw + (size_t)il * p.ntile_n;
            uint32_t v[32];
            int c0 = 0;
            if (rows > 0) {
                for (; c0 + 32 <= ncols; c0 += 32) {
                    tmem_ld32(taddr + c0, v);
                    uint4 w0, w1;
                    w0.x = i8_pack4(v + 0, pm, magic, off);
                    w0.y = i8_pack4(v + 4, pm, magic, off);
                    w0.z = i8_pack4(v + 8, pm, magic, off);
                    w0.w = i8_pack4(v + 12, pm, magic, off);
                    w1.x = i8_pack4(v + 16, pm, magic, off);
                    w1.y = i8_pack4(v + 20, pm, magic, off);
                    w1.z = i8_pack4(v + 24, pm, magic, off);
                    w1.w = i8_pack4(v + 28, pm, magic, off);
                    if (live) {
                        *reinterpret_cast<uint4*>(dst + c0) = w0;
                        *reinterpret_cast<uint4*>(dst + c0 + 16) = w1;
                    }
                }
                if (c0 < ncols) {
                    tmem_ld16(taddr + c0, v);
                    uint4 w0;
                    w0.x = i8_pack4(v + 0, pm, magic, off);
                    w0.y = i8_pack4(v + 4, pm, magic, off);
                    w0.z = i8_pack4(v + 8, pm, magic, off);
                    w0.w = i8_pack4(v + 12, pm, magic, off);
                    if (live) *reinterpret_cast<uint4*>(dst + c0) = w0;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) i8_cluster_sync();  // nobody leaves while a peer may still multicast into its shared memory or signal its barriers
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

// ---- CRT: residues -> T ------------------------------------------------------------------------------------------------
struct I8HalfCrtParams {
    const uint8_t* ws;
    size_t ws_mod_stride, Tpitch;
    int opw, o, op, qc, qbeg, nq, m0;
    const int *eB, *eC;
    double* T;
    unsigned long long M_lo, M_hi, H_lo, H_hi;
    // first J sweep riding on the GEMM (I8HalfFuseJ::gemm_col): column dcol (>= o; -1: none) holds the residues of
    // sum_k B'_m[q,k] D'[m,n_k]; the value, * 2^-(eB(m,q) + eD(m)), goes to dpart[m * dstride + qbeg + q]
    int dcol, dstride;
    const int* eD;
    double* dpart;
};
// grid (ceil(qc * opw / 4 / 256), nmc): one thread = four consecutive orbitals of one (m, q)
template <int NMOD>
__global__ void __launch_bounds__(256) i8h_crt_kernel(const I8HalfCrtParams p) {
    const int per_q = p.opw >> 2;
    const long idx = (long)blockIdx.x * 256 + threadIdx.x;
    const int q = (int)(idx / per_q), i4 = (int)(idx - (long)q * per_q) * 4;
    if (q >= p.qc || (i4 >= p.op && !(p.dcol >= i4 && p.dcol < i4 + 4))) return;
    const int mloc = blockIdx.y, m = p.m0 + mloc;
    const uint8_t* src = p.ws + ((size_t)mloc * p.qc + q) * p.opw + i4;
    uint32_t w[NMOD];
#pragma unroll
    for (int j = 0; j < NMOD; j++) w[j] = *reinterpret_cast<const uint32_t*>(src + (size_t)j * p.ws_mod_stride);
    const int eb = p.eB[(size_t)m * p.nq + p.qbeg + q];
    double* dst = p.T + (size_t)m * p.Tpitch + (size_t)q * p.op + i4;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int i = i4 + c;
        if (i == p.dcol) {
            int rr[NMOD];
#pragma unroll
            for (int j = 0; j < NMOD; j++) rr[j] = (int)((w[j] >> (8 * c)) & 255u);
            const int e2 = -(eb + p.eD[m]);
            const double f = i8_crt_value_fast<NMOD>(rr);
            p.dpart[(size_t)m * p.dstride + p.qbeg + q] = (e2 > -1000 && e2 < 1000) ? f * __hiloint2double((1023 + e2) << 20, 0) : ldexp(f, e2);
        }
        if (i >= p.op) continue;
        double val = 0.0;
        if (i < p.o) {
            int rr[NMOD];
#pragma unroll
            for (int j = 0; j < NMOD; j++) rr[j] = (int)((w[j] >> (8 * c)) & 255u);
            // * 2^-(eB + eC): a power of two built from its exponent field is the same rounding as ldexp (one correctly
            // rounded scaling) at a twentieth of its instructions; exponents outside the normal range take ldexp
            const int e2 = -(eb + p.eC[i]);
            const double f = i8_crt_value_fast<NMOD>(rr);
            val = (e2 > -1000 && e2 < 1000) ? f * __hiloint2double((1023 + e2) << 20, 0) : ldexp(f, e2);
        }
        dst[c] = val;
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
struct I8HalfPlan {  // per shard
    // k-blocks per row-block (from the engine's tables, at set_layout): nkb(m) = ceil(sp(m) / 128), kboff = running sum
    std::vector<int> kboff;  // [nbf + 1]
    int max_nkb_row = 0;     // largest nkb(m)
    double sum_sp = 0;       // kept pairs (sum of sp(m))
    std::vector<double> sp_prefix;  // [nbf + 1] running sum of sp(m)
    int* d_kboff = nullptr;
    // scales of the tensor rows, per tensor; valid until the tensor changes
    int* expoB[3] = {nullptr, nullptr, nullptr};
    bool expo_valid[3] = {false, false, false};
    int expo_nmod[3] = {0, 0, 0};
    // C operand
    int8_t* rc = nullptr;
    size_t rc_cap = 0;
    int* expoC = nullptr;
    size_t expoC_cap = 0;
    double* normpart = nullptr;
    size_t normpart_cap = 0;
    // density rows of the first J sweep when it rides on the GEMM: residue planes rD[j][m][n], row scales
    int8_t* rD = nullptr;
    size_t rD_cap = 0;
    int* expoD = nullptr;
    size_t expoD_cap = 0;
    double* normpartD = nullptr;
    size_t normpartD_cap = 0;
    // scratch arena: planes | cg | ws
    uint8_t* arena = nullptr;
    size_t arena_cap = 0;
    uint64_t launches = 0;
    bool consts = false;
    int crt_nmod = 0;  // moduli the constants of the floating-point CRT were uploaded for
    // When the arena holds the planes of the whole Q range in one chunk they stay valid until the tensor changes: the next
    // build (and the second density of an open-shell build) skips the conversion -- which is most of the arm's time.
    int cache_which = -1, cache_qbeg = -1, cache_qc = -1, cache_nmod = -1, cache_mr = -1;
    uint64_t conversions_skipped = 0;
    cudaEvent_t prof[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // optional: start / convert / gather / GEMM / CRT of chunk 0
    // optional phase marker of the caller (the engine records an event on the stream): 10 = a chunk starts, 11 / 12 / 13 / 14 =
    // its conversion / gather / GEMM / CRT is launched
    void (*mark)(void* ctx, int tag) = nullptr;
    void* mark_ctx = nullptr;
    void release_arena() {
        if (arena) cudaFree(arena);
        arena = nullptr;
        arena_cap = 0;
        cache_which = -1;
    }
    void release() {
        release_arena();
        for (int i = 0; i < 3; i++) {
            if (expoB[i]) cudaFree(expoB[i]);
            expoB[i] = nullptr;
            expo_valid[i] = false;
        }
        void* ptrs[] = {d_kboff, rc, expoC, normpart, rD, expoD, normpartD};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        d_kboff = nullptr;
        rc = nullptr;
        expoC = nullptr;
        normpart = nullptr;
        rD = nullptr;
        expoD = nullptr;
        normpartD = nullptr;
        rc_cap = expoC_cap = normpart_cap = rD_cap = expoD_cap = normpartD_cap = 0;
        kboff.clear();
    }
};

inline int i8h_set_layout(I8HalfPlan& pl, const std::vector<int>& sp, std::string* err) {
    const size_t nbf = sp.size();
    pl.kboff.resize(nbf + 1);
    pl.kboff[0] = 0;
    pl.max_nkb_row = 0;
    pl.sum_sp = 0;
    pl.sp_prefix.assign(nbf + 1, 0.0);
    for (size_t m = 0; m < nbf; m++) {
        pl.sum_sp += (double)sp[m];
        pl.sp_prefix[m + 1] = pl.sp_prefix[m] + (double)sp[m];
        pl.kboff[m + 1] = pl.kboff[m] + (sp[m] + I8_BK - 1) / I8_BK;
        pl.max_nkb_row = std::max(pl.max_nkb_row, pl.kboff[m + 1] - pl.kboff[m]);
    }
    if (pl.d_kboff) cudaFree(pl.d_kboff);
    pl.d_kboff = nullptr;
    I8CK(cudaMalloc((void**)&pl.d_kboff, (nbf + 1) * sizeof(int)));
    I8CK(cudaMemcpy(pl.d_kboff, pl.kboff.data(), (nbf + 1) * sizeof(int), cudaMemcpyHostToDevice));
    for (int i = 0; i < 3; i++) {  // a new layout: the row scales (sized nbf x nq of the old one) and the cached planes go
        if (pl.expoB[i]) cudaFree(pl.expoB[i]);
        pl.expoB[i] = nullptr;
        pl.expo_valid[i] = false;
    }
    pl.cache_which = -1;
    return 0;
}

// true if the planes of (tensor which, Q range, moduli) are still in the arena from an earlier call
inline bool i8h_cached(const I8HalfPlan& pl, int which, int qbeg, int qc, int nmod) {
    return pl.arena && pl.cache_which == which && pl.cache_qbeg == qbeg && pl.cache_qc == qc && pl.cache_nmod == nmod &&
           pl.expo_valid[which] && pl.expo_nmod[which] == nmod;
}

template <int NMOD>
inline void i8h_launch_convert(const double* tensor, const size_t* row_off, const int* ldm, const int* sp, const int* kboff,
                               const int* expo, int nq, int m0, int nmc, int qbeg, int qc, int nqt, int8_t* planes, size_t plane_stride,
                               const I8HalfFuseJ& fj, int max_nkb, cudaStream_t st) {
    const size_t smem = fj.Dm ? (size_t)max_nkb * I8_BK * sizeof(double) : 0;
    i8h_convert_kernel<NMOD><<<dim3((unsigned)((qc + I8H_ROWS - 1) / I8H_ROWS), (unsigned)nmc), 256, smem, st>>>(
        tensor, row_off, ldm, sp, kboff, expo, nq, m0, qbeg, qc, nqt, planes, plane_stride, fj);
}
template <int NMOD>
inline void i8h_launch_crt(const I8HalfCrtParams& p, int nmc, cudaStream_t st) {
    const long per_m = (long)p.qc * (p.opw / 4);
    i8h_crt_kernel<NMOD><<<dim3((unsigned)((per_m + 255) / 256), (unsigned)nmc), 256, 0, st>>>(p);
}
#define I8H_DISPATCH(nmod, F, ...)                 \
    switch (nmod) {                                \
        case 6: F<6>(__VA_ARGS__); break;          \
        case 7: F<7>(__VA_ARGS__); break;          \
        case 8: F<8>(__VA_ARGS__); break;          \
        case 9: F<9>(__VA_ARGS__); break;          \
        case 10: F<10>(__VA_ARGS__); break;        \
        case 11: F<11>(__VA_ARGS__); break;        \
        case 12: F<12>(__VA_ARGS__); break;        \
        default: F<13>(__VA_ARGS__); break;        \
    }

template <int CL>
inline int i8h_launch_gemm(const I8HalfParams& gp, int nsm, cudaStream_t st, std::string* err) {
    static bool attr[64] = {false};  // function attributes are per device: one process may drive all eight
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 63]) {
        I8CK(cudaFuncSetAttribute(i8h_gemm_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)i8_smem_bytes()));
        attr[dev & 63] = true;
    }
    if (CL == 1) {
        i8h_gemm_kernel<CL><<<(unsigned)std::min(gp.nitems, nsm), I8_THREADS, i8_smem_bytes(), st>>>(gp);
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(I8_THREADS);
    cfg.dynamicSmemBytes = i8_smem_bytes();
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(nsm / CL * CL));
    static int max_clusters = 0;
    if (!max_clusters) {
        I8CK(cudaOccupancyMaxActiveClusters(&max_clusters, i8h_gemm_kernel<CL>, &cfg));
        if (max_clusters < 1) {
            max_clusters = 0;
            if (err) *err = "no cluster of this size can be resident";
            return 2;
        }
    }
    const int ncl = std::min(max_clusters, gp.nitems);
    cfg.gridDim = dim3((unsigned)(ncl * CL));
    I8CK(cudaLaunchKernelEx(&cfg, i8h_gemm_kernel<CL>, gp));
    return 0;
}

struct I8HalfInfo {
    int nmod = 0, nchunks = 0, ntile_n = 0, nit = 0, cluster = 0, cached = 0;
    double bits = 0;
    size_t arena = 0;
    // work of this call: 2 x int8 multiply-adds issued to the tensor cores (all moduli, whole padded tiles), bytes of residue
    // planes the GEMM read, bytes the conversions that RAN moved (f64 rows read + planes written)
    double mma_ops = 0, plane_bytes = 0, convert_bytes = 0;
    int resident_rows = 0;  // row-blocks [0, resident_rows) keep their planes across builds
    int resident_hit = 0;   // ... and this call found them valid (their conversion was skipped)
};

// orbital tiling: nit tiles of ntile_n columns (whole 8-row swizzle atoms per cluster slice), opw = nit * ntile_n >= o
inline void i8h_tiling(int o, int cluster, int* nit, int* ntile_n) {
    const int ngran = cluster > 1 ? 8 * cluster : 16;
    const int opw0 = (o + 15) / 16 * 16;
    *nit = (opw0 + I8_TN - 1) / I8_TN;
    *ntile_n = ((opw0 + *nit - 1) / *nit + ngran - 1) / ngran * ngran;
}

// The first J sweep can ride on the GEMM when the tiling has a free column behind the nocc orbitals
inline bool i8h_can_fuse_col(const I8HalfPlan& pl, int o, int cluster) {
    (void)pl;
    if (cluster != 2 && cluster != 4) cluster = 1;
    int nit, ntile_n;
    i8h_tiling(o, cluster, &nit, &ntile_n);
    return o > 0 && o < nit * ntile_n;
}

// bytes of arena one row-block with nkb k-blocks needs
inline size_t i8h_cost(int nmod, int nkb, int nqt, int qc, int nit, int ntile_n) {
    return (size_t)nmod * ((size_t)nkb * ((size_t)nqt * I8H_TILE + (size_t)nit * ntile_n * I8_BK) + (size_t)qc * nit * ntile_n);
}
// Smallest arena one call needs (one row-block per chunk) and the arena that holds everything in one chunk.
inline void i8h_arena_need(const I8HalfPlan& pl, int nmod, int qc, int max_o, int cluster, size_t* min_bytes, size_t* all_bytes) {
    const size_t nbf = pl.kboff.size() - 1;
    int nit, ntile_n;
    i8h_tiling(max_o, cluster, &nit, &ntile_n);
    const int nqt = (qc + I8_TM - 1) / I8_TM;
    size_t mx = 0, all = 0;
    for (size_t m = 0; m < nbf; m++) {
        const size_t c = i8h_cost(nmod, pl.kboff[m + 1] - pl.kboff[m], nqt, qc, nit, ntile_n);
        mx = std::max(mx, c);
        all += c;
    }
    *min_bytes = 2 * mx + 2 * 65536 + 2048;  // i8_half_run keeps a work region of two row-blocks' worth at least
    *all_bytes = all + 65536 + 2048;
}

// T[m][q][i] (pitch Tpitch per m, op per q) for q in [qbeg, qbeg + qc) of tensor `which`.
//   Ct: C^T, o rows of pitch ldc (the engine's transpose_c_kernel output).  max_o: the largest nocc of the build (the
//   chunk plan is made for it, so that it does not change between the densities of an open-shell build).
// Returns 0, 2 (CUDA) or 3 (arena too small).
inline int i8_half_run(I8HalfPlan& pl, cudaStream_t st, int nsm, const double* tensor, int which, const size_t* d_row_off,
                       const int* d_ldm, const int* d_sp, const int* d_cols, const size_t* d_cols_off, int nbf, int nq,
                       const double* Ct, int ldc, int o, int op, int max_o, int qbeg, int qc, double* T, size_t Tpitch, int nmod,
                       int cluster, const I8HalfFuseJ* fuse, I8HalfInfo* info, std::string* err) {
    nmod = std::max(I8_MINMOD, std::min(I8_MAXMOD, nmod));
    if (cluster != 2 && cluster != 4) cluster = 1;
    int rc;
    if (!pl.consts) {
        I8Consts c;
        i8_fill_consts(c);
        I8CK(cudaMemcpyToSymbolAsync(c_i8, &c, sizeof c, 0, cudaMemcpyHostToDevice, st));
        I8CK(cudaStreamSynchronize(st));
        pl.consts = true;
    }
    if (pl.crt_nmod != nmod) {
        I8CrtFast f;
        i8_fill_crt_fast(f, nmod);
        I8CK(cudaMemcpyToSymbolAsync(c_i8f, &f, sizeof f, 0, cudaMemcpyHostToDevice, st));
        I8CK(cudaStreamSynchronize(st));
        pl.crt_nmod = nmod;
    }
    unsigned __int128 M;
    const double Rb = i8_row_bound(nmod, nbf, &M);
    // scales of the tensor rows: once per tensor (and per number of moduli)
    if (!pl.expo_valid[which] || pl.expo_nmod[which] != nmod) {
        if (!pl.expoB[which]) I8CK(cudaMalloc((void**)&pl.expoB[which], (size_t)nbf * nq * sizeof(int)));
        const long warps = (long)nbf * nq;
        i8h_rowscale_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(tensor, d_row_off, d_ldm, nq, nbf, Rb, pl.expoB[which]);
        pl.launches++;
        pl.expo_valid[which] = true;
        pl.expo_nmod[which] = nmod;
        if (pl.cache_which == which) pl.cache_which = -1;  // the tensor changed: its resident planes are stale too
    }
    // C operand: column scales and residue planes rc[j][i][n] (the row kernels of i8_kgemm.cuh on the rows of C^T)
    const size_t rc_ld = ((size_t)nbf + 127) / 128 * 128, rc_plane = rc_ld * (size_t)o;
    const int nchunkn = 4;
    if ((rc = i8_grow(&pl.rc, &pl.rc_cap, rc_plane * nmod, err))) return rc;
    if ((rc = i8_grow(&pl.expoC, &pl.expoC_cap, (size_t)o, err))) return rc;
    if ((rc = i8_grow(&pl.normpart, &pl.normpart_cap, (size_t)o * nchunkn, err))) return rc;
    i8_rownorm_kernel<<<dim3(nchunkn, o), 256, 0, st>>>(Ct, (size_t)ldc, nbf, nchunkn, pl.normpart);
    i8_rowscale_kernel<<<(o + 127) / 128, 128, 0, st>>>(pl.normpart, nchunkn, o, Rb, pl.expoC);
    i8_convert(nmod, Ct, (size_t)ldc, nbf, o, pl.expoC, pl.rc, rc_ld, rc_plane, st);
    pl.launches += 3;
    // density rows of a first J sweep that rides on the GEMM: the same row kernels on the rows of D'
    const bool fuse_col = fuse && fuse->gemm_col;
    const size_t rd_plane = rc_ld * (size_t)nbf;
    if (fuse_col) {
        if (!i8h_can_fuse_col(pl, o, cluster)) {
            if (err) *err = "the first J sweep cannot ride on the GEMM: no free column behind nocc";
            return 2;
        }
        if ((rc = i8_grow(&pl.rD, &pl.rD_cap, rd_plane * nmod, err))) return rc;
        if ((rc = i8_grow(&pl.expoD, &pl.expoD_cap, (size_t)nbf, err))) return rc;
        if ((rc = i8_grow(&pl.normpartD, &pl.normpartD_cap, (size_t)nbf * nchunkn, err))) return rc;
        i8_rownorm_kernel<<<dim3(nchunkn, nbf), 256, 0, st>>>(fuse->Dm, (size_t)fuse->ldd, nbf, nchunkn, pl.normpartD);
        i8_rowscale_kernel<<<(nbf + 127) / 128, 128, 0, st>>>(pl.normpartD, nchunkn, nbf, Rb, pl.expoD);
        i8_convert(nmod, fuse->Dm, (size_t)fuse->ldd, nbf, nbf, pl.expoD, pl.rD, rc_ld, rd_plane, st);
        pl.launches += 3;
    }

    int nit, ntile_n, nit_p, ntile_p;
    i8h_tiling(o, cluster, &nit, &ntile_n);
    i8h_tiling(max_o, cluster, &nit_p, &ntile_p);  // the plan is made for the largest nocc
    const int opw = nit * ntile_n;
    const int nqt = (qc + I8_TM - 1) / I8_TM;
    // ---- plan: resident planes + work region ------------------------------------------------------------------------
    // The arena is split into a RESIDENT region -- the residue planes of row-blocks [0, m_r), which stay valid until the
    // tensor changes, so that later builds (and the second density of an open-shell build) skip their conversion -- and
    // a WORK region that holds, chunk by chunk, the gathered C^T and the residues of the GEMM (and, for the row-blocks
    // behind m_r, their planes, converted in every build).  Everything fits: m_r = nbf, one chunk.  Otherwise a fifth of
    // the arena is work region and the rest resident: C60 on one GPU keeps a third of its 132 GB of planes.
    // B200JK_I8_RESIDENT=0: nothing resident unless everything fits (the round-2 rule, for A/B).
    static int resident_env = -1;
    if (resident_env < 0) {
        const char* e = getenv("B200JK_I8_RESIDENT");
        resident_env = (e && e[0] == '0') ? 0 : 1;
    }
    const size_t align = 1024, slack = 65536;
    auto planes_of = [&](int m) { return (size_t)nmod * (size_t)(pl.kboff[m + 1] - pl.kboff[m]) * (size_t)nqt * I8H_TILE; };
    auto cgws_of = [&](int m) {
        const size_t nk = (size_t)(pl.kboff[m + 1] - pl.kboff[m]);
        const size_t c1 = (size_t)nmod * (nk * (size_t)nit * ntile_n * I8_BK + (size_t)qc * nit * ntile_n);
        const size_t c2 = (size_t)nmod * (nk * (size_t)nit_p * ntile_p * I8_BK + (size_t)qc * nit_p * ntile_p);
        return std::max(c1, c2);
    };
    size_t planes_all = 0, cgws_all = 0, max_block = 0;
    for (int m = 0; m < nbf; m++) {
        planes_all += planes_of(m);
        cgws_all += cgws_of(m);
        max_block = std::max(max_block, planes_of(m) + cgws_of(m));
    }
    int m_r = 0;
    size_t resident_bytes = 0;
    if (planes_all + cgws_all + slack + align <= pl.arena_cap) {
        m_r = nbf;
        resident_bytes = planes_all;
    } else if (resident_env) {
        const size_t work = std::max(pl.arena_cap / 5, 2 * max_block + slack);
        if (work + align < pl.arena_cap) {
            const size_t room = pl.arena_cap - work - align;
            while (m_r < nbf && resident_bytes + planes_of(m_r) <= room) resident_bytes += planes_of(m_r++);
        }
    }
    const size_t work_off = (resident_bytes + align - 1) / align * align;
    if (work_off + max_block + slack > pl.arena_cap) {
        if (err) *err = "scratch arena too small for one row-block of residue planes";
        return 3;
    }
    const size_t work_cap = pl.arena_cap - work_off - slack;
    struct Chunk {
        int m0, m1;
        bool resident;
    };
    std::vector<Chunk> chunks;
    for (int part = 0; part < 2; part++) {
        const int lo = part == 0 ? 0 : m_r, hi = part == 0 ? m_r : nbf;
        int m0 = lo;
        size_t cost = 0;
        for (int m = lo; m < hi; m++) {
            const size_t c = cgws_of(m) + (part == 0 ? 0 : planes_of(m));
            if (m > m0 && cost + c > work_cap) {
                chunks.push_back(Chunk{m0, m, part == 0});
                m0 = m;
                cost = 0;
            }
            cost += c;
        }
        if (hi > m0) chunks.push_back(Chunk{m0, hi, part == 0});
    }
    const size_t resident_stride = (size_t)pl.kboff[m_r] * (size_t)nqt * I8H_TILE;  // one resident plane
    const bool cached = m_r > 0 && pl.arena && pl.cache_which == which && pl.cache_qbeg == qbeg && pl.cache_qc == qc &&
                        pl.cache_nmod == nmod && pl.cache_mr == m_r && pl.expo_valid[which] && pl.expo_nmod[which] == nmod;
    if (fuse && !fuse_col && m_r > 0) {
        if (err) *err = "the first J sweep cannot ride on a conversion that is skipped (planes resident)";
        return 2;
    }
    if (!cached) pl.cache_which = -1;
    int nch = 0;
    double converted_sp = 0, converted_planes = 0;
    for (auto& c : chunks) {
        const int nmc = c.m1 - c.m0;
        const size_t nkb_c = (size_t)(pl.kboff[c.m1] - pl.kboff[c.m0]);
        const bool prof = pl.prof[0] && nch == 0;
        const bool convert = !(c.resident && cached);
        // carve: resident chunk = planes in the resident region, C^T | residues in the work region; otherwise all three there
        int8_t* planes;
        size_t plane_stride;
        uint8_t* wbase = pl.arena + work_off;
        if (c.resident) {
            planes = reinterpret_cast<int8_t*>(pl.arena) + (size_t)pl.kboff[c.m0] * (size_t)nqt * I8H_TILE;
            plane_stride = resident_stride;
        } else {
            planes = reinterpret_cast<int8_t*>(wbase);
            plane_stride = nkb_c * (size_t)nqt * I8H_TILE;
            wbase += plane_stride * nmod;
        }
        const size_t cg_plane = nkb_c * (size_t)nit * ntile_n * I8_BK;
        int8_t* cg = reinterpret_cast<int8_t*>(wbase);
        uint8_t* ws = wbase + cg_plane * nmod;
        if ((size_t)(ws - pl.arena) + (size_t)nmod * nmc * qc * opw > pl.arena_cap) {
            if (err) *err = "scratch arena accounting";
            return 3;
        }
        if (prof) cudaEventRecord(pl.prof[0], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 10);
        if (!convert) pl.conversions_skipped++;
        I8HalfFuseJ fjv = {nullptr, 0, nullptr, 0, nullptr, nullptr};
        if (fuse && !fuse_col) {
            fjv = *fuse;
            fjv.cols = d_cols;
            fjv.cols_off = d_cols_off;
        }
        if (convert) {
            I8H_DISPATCH(nmod, i8h_launch_convert, tensor, d_row_off, d_ldm, d_sp, pl.d_kboff, pl.expoB[which], nq, c.m0, nmc, qbeg, qc, nqt,
                         planes, plane_stride, fjv, pl.max_nkb_row, st);
            converted_sp += pl.sp_prefix[c.m1] - pl.sp_prefix[c.m0];
            converted_planes += (double)nmod * nqt * (double)nkb_c * I8H_TILE;
        }
        if (prof) cudaEventRecord(pl.prof[1], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 11);
        I8H_DISPATCH(nmod, i8h_launch_gather, pl.rc, rc_ld, rc_plane, o, d_sp, pl.d_kboff, d_cols, d_cols_off, c.m0, nmc, nit, ntile_n, cg, cg_plane,
                     fuse_col ? pl.rD : nullptr, rd_plane, fuse_col ? o : -1, st);
        if (prof) cudaEventRecord(pl.prof[2], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 12);
        I8HalfParams gp;
        gp.nmod = nmod;
        gp.nqt = nqt;
        gp.nit = nit;
        gp.ntile_n = ntile_n;
        gp.m0 = c.m0;
        gp.qc = qc;
        gp.opw = opw;
        gp.ws_mod_stride = (size_t)nmc * qc * opw;
        gp.plane_stride = plane_stride;
        gp.cg_plane = cg_plane;
        gp.kboff = pl.d_kboff;
        gp.planes = planes;
        gp.cg = cg;
        gp.ws = ws;
        gp.nitems = nmc * nmod * ((nqt + cluster - 1) / cluster) * nit;
        if ((rc = cluster == 1 ? i8h_launch_gemm<1>(gp, nsm, st, err)
                               : (cluster == 2 ? i8h_launch_gemm<2>(gp, nsm, st, err) : i8h_launch_gemm<4>(gp, nsm, st, err))))
            return rc;
        if (prof) cudaEventRecord(pl.prof[3], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 13);
        I8HalfCrtParams cp;
        cp.ws = ws;
        cp.ws_mod_stride = gp.ws_mod_stride;
        cp.Tpitch = Tpitch;
        cp.opw = opw;
        cp.o = o;
        cp.op = op;
        cp.qc = qc;
        cp.qbeg = qbeg;
        cp.nq = nq;
        cp.m0 = c.m0;
        cp.eB = pl.expoB[which];
        cp.eC = pl.expoC;
        cp.T = T;
        cp.M_lo = (unsigned long long)M;
        cp.M_hi = (unsigned long long)(M >> 64);
        cp.H_lo = (unsigned long long)(M / 2);
        cp.H_hi = (unsigned long long)((M / 2) >> 64);
        cp.dcol = fuse_col ? o : -1;
        cp.dstride = fuse_col ? fuse->dstride : 0;
        cp.eD = pl.expoD;
        cp.dpart = fuse_col ? fuse->dpart : nullptr;
        I8H_DISPATCH(nmod, i8h_launch_crt, cp, nmc, st);
        if (prof) cudaEventRecord(pl.prof[4], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 14);
        pl.launches += convert ? 4 : 3;
        I8CK(cudaGetLastError());
        nch++;
    }
    if (m_r > 0) {
        pl.cache_which = which;
        pl.cache_qbeg = qbeg;
        pl.cache_qc = qc;
        pl.cache_nmod = nmod;
        pl.cache_mr = m_r;
    }
    if (info) {
        info->cached = (cached && m_r == nbf) ? 1 : 0;
        info->resident_rows = m_r;
        info->resident_hit = cached ? 1 : 0;
        info->nmod = nmod;
        info->nchunks = nch;
        info->ntile_n = ntile_n;
        info->nit = nit;
        info->cluster = cluster;
        info->bits = log2(Rb);
        info->arena = pl.arena_cap;
        const double nkb_all = (double)pl.kboff[nbf];
        info->mma_ops = 2.0 * nmod * ((double)((nqt + cluster - 1) / cluster * cluster) * I8_TM) * (nkb_all * I8_BK) * ((double)nit * ntile_n);
        info->plane_bytes = (double)nmod * nqt * nkb_all * I8H_TILE;
        info->convert_bytes = 8.0 * qc * converted_sp + converted_planes;
    }
    return 0;
}

}  // namespace b2k
