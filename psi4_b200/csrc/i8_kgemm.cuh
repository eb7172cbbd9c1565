// i8_kgemm.cuh -- the K GEMM on the INT8 tensor cores (tcgen05.mma.kind::i8, accumulators in TMEM), exact in the
// integers: an alternative arm of K4  K[m,n] += sum_k T1[m,k] * T2[n,k]  (lib3index/dfhelper.cc:3374,
// C_DGEMM('N','T',nbf,nbf,nocc*block)) next to the DMMA kernel of dmma_ws.cuh.
//
// FP64 has no tcgen05 path; the FP64 pipe tops out at 37 TFLOP/s while the int8 tensor pipe runs 4.5 Pop/s.  The
// product is therefore computed by residues (the "Ozaki scheme II" construction; cuSOLVER/cuBLAS 13 ship the same idea
// as their FP64 emulation -- neither is in this CUDA 12.9 image):
//   1. every row of T gets a power-of-two scale 2^e(m) such that the integer row  T'[m,:] = rint(T[m,:] * 2^e(m))  has
//      2-norm <= R, with R^2 < M/2 and M = p_0 * ... * p_{n-1} the product of n pairwise coprime moduli <= 256
//      (Cauchy-Schwarz: every |sum_k T1'[m,k] T2'[n,k]| < M/2);
//   2. T' is stored as n planes of int8 residues (symmetric representatives, |r| <= 128);
//   3. per modulus one int8 GEMM with int32 accumulation in TMEM (exact: k-ranges of <= 65536, |sum| <= 2^30), the
//      accumulator reduced mod p_j in the epilogue and written as one byte per element and k-range;
//   4. the residues of all k-ranges are summed, the integer sum_k T1' T2' is rebuilt by the Chinese remainder theorem
//      (Garner's mixed-radix digits, 128-bit Horner), converted to double and scaled by 2^-(e(m)+e(n)).
// Steps 3 and 4 are exact; the only rounding is the quantisation of the inputs in step 1, |T' 2^-e - T| <= 2^-e-1 with
// 2^e >= R / (2 |T[m,:]|): with 13 moduli R = 2^50.7, i.e. every element carries 50 bits below the NORM of its row --
// the same norm-wise error a DGEMM in double commits -- and integer arithmetic makes the result independent of the
// split-K factor, the scheduling and the tile shape (bit-identical run to run by construction).
//
// Kernel: one CTA per SM, persistent over (k-range, modulus, tile) items in that order so that the ~2.5 residue
// plane slabs in flight stay in L2; warp 0 = TMA producer (3-D maps {k, row, plane}, SWIZZLE_128B, 4-stage ring),
// warp 1 = MMA issuer (one lane, 4 x tcgen05.mma M128 N<=256 K32 per stage, tcgen05.commit frees the stage), warps 2-5 =
// epilogue (tcgen05.ld 32 columns at a time, mod p, pack, 32-byte stores) on the other of two TMEM accumulators.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "dmma_ws.cuh"

namespace b2k {

constexpr int I8_MAXMOD = 13;  // more moduli would need |T'| >= 2^52: the residues are taken in double arithmetic
constexpr int I8_MINMOD = 6;
constexpr int I8_STAGES = 4;
constexpr int I8_QRING = 4;   // item ids in flight between the producer and the MMA / epilogue warps
constexpr int I8_BK = 128;  // bytes (= int8 elements) of k per stage
constexpr int I8_TM = 128, I8_TN = 256;
constexpr int I8_A_STAGE = I8_TM * I8_BK;  // 16 KB
constexpr int I8_B_STAGE = I8_TN * I8_BK;  // 32 KB
constexpr int I8_THREADS = 192;
constexpr int I8_TILE_BYTES = I8_TM * I8_TN;  // residue bytes one item writes
constexpr int I8_MAX_KLEN = 65536;
constexpr size_t i8_smem_bytes() { return 1024 + (size_t)I8_STAGES * (I8_A_STAGE + I8_B_STAGE) + 256; }

static const int kI8Moduli[16] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193};

struct I8Consts {
    int p[16];
    float invp[16];
    double invpd[16];
    unsigned magic[16];  // floor(2^32 / p)
    unsigned off[16];    // smallest multiple of p that is >= 2^30
    int ginv[16][16];    // ginv[i][j] = p_i^-1 mod p_j (i < j)
};
__constant__ I8Consts c_i8;

// Constants of the floating-point CRT for one number of moduli (i8_crt_value_fast): value / M = frac(sum_j r_j q_j / p_j),
// q_j = (M / p_j)^-1 mod p_j as a symmetric representative; 1 / p_j = ih_j + il_j with ih_j a multiple of 2^-32.
struct I8CrtFast {
    int q[16];
    double ih[16], il[16];
    double Md;
    int nmod;
};
__constant__ I8CrtFast c_i8f;

struct I8Tile {
    int row0, col0, ncols, bmap;  // rows [row0, row0+128) of T1 against rows [col0, col0+ncols) of T2; which B map
};

struct I8GemmParams {
    int nitems, ntile, nmod, kdim, klen;  // item = (split * nmod + mod) * ntile + tile
    const I8Tile* tiles;
    uint8_t* ws;   // [nitems][128][256] residues in [0, p)
    int* counter;  // work-queue head (zeroed before the launch)
};

// ---- tcgen05 / TMA primitives ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t i8_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_i8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
        "[%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// x (int32 accumulator, |x| <= 2^30) -> x mod p in [0, p)
__device__ __forceinline__ uint32_t i8_mod_acc(uint32_t v, uint32_t p, uint32_t magic, uint32_t off) {
    const uint32_t x = v + off;  // off = multiple of p >= 2^30: x in [0, 2^31 + p)
    const uint32_t q = __umulhi(x, magic);
    uint32_t r = x - q * p;
    if (r >= p) r -= p;
    return r;
}
__device__ __forceinline__ uint32_t i8_pack4(const uint32_t* v, uint32_t p, uint32_t magic, uint32_t off) {
    return i8_mod_acc(v[0], p, magic, off) | (i8_mod_acc(v[1], p, magic, off) << 8) | (i8_mod_acc(v[2], p, magic, off) << 16) |
           (i8_mod_acc(v[3], p, magic, off) << 24);
}

// ---- step 3: residue GEMM --------------------------------------------------------------------------------
// One work item of the persistent pipeline: D[128 x ncols] (int32, TMEM) = A[128 x K] B[ncols x K]^T for one modulus,
// K = nkb blocks of 128 bytes; the accumulator is reduced mod p and stored as bytes, `rows` rows of `ncols` bytes.
struct I8Item {
    const CUtensorMap *amap, *bmap;
    int ax, ay, az, bx, by, bz;  // TMA coordinates of k-block 0 (x advances by 128 per k-block)
    int nkb, ncols, mod, rows;
    uint8_t* dst;      // residues of row 0
    size_t dst_pitch;  // bytes between rows
};

// The K GEMM's items: (k-range, modulus, tile) in that order.
struct I8KgemmTraits {
    typedef I8GemmParams Params;
    static __device__ __forceinline__ int nitems(const Params& p) { return p.nitems; }
    static __device__ __forceinline__ int* counter(const Params& p) { return p.counter; }
    static __device__ __forceinline__ void decode(const Params& p, int it, const CUtensorMap* amap, const CUtensorMap* b0,
                                                  const CUtensorMap* b1, const CUtensorMap* b2, I8Item& x) {
        const int per_split = p.nmod * p.ntile;
        const int split = it / per_split, rem = it - split * per_split;
        const int mod = rem / p.ntile, tile = rem - mod * p.ntile;
        const I8Tile t = p.tiles[tile];
        const int k0 = split * p.klen;
        x.amap = amap;
        x.bmap = t.bmap == 0 ? b0 : (t.bmap == 1 ? b1 : b2);
        x.ax = x.bx = k0;
        x.ay = t.row0;
        x.by = t.col0;
        x.az = x.bz = mod;
        x.nkb = (min(p.klen, p.kdim - k0) + I8_BK - 1) / I8_BK;
        x.ncols = t.ncols;
        x.mod = mod;
        x.rows = I8_TM;
        x.dst = p.ws + (size_t)it * I8_TILE_BYTES;
        x.dst_pitch = I8_TN;
    }
};

template <class Traits>
__global__ void __launch_bounds__(I8_THREADS, 1)
i8_pipeline_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap0,
                   const __grid_constant__ CUtensorMap bmap1, const __grid_constant__ CUtensorMap bmap2,
                   const typename Traits::Params p) {
    extern __shared__ uint8_t i8_raw[];
    uint8_t* base = i8_raw + ((1024u - (smem_u32(i8_raw) & 1023u)) & 1023u);
    uint8_t* As = base;
    uint8_t* Bs = base + I8_STAGES * I8_A_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(Bs + I8_STAGES * I8_B_STAGE);
    uint64_t* empty = full + I8_STAGES;
    uint64_t* tfull = empty + I8_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* qfull = tempty + 2;
    uint64_t* qempty = qfull + I8_QRING;
    int* qitem = reinterpret_cast<int*>(qempty + I8_QRING);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qitem + I8_QRING);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < I8_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        for (int s = 0; s < I8_QRING; s++) {
            mbar_init(&qfull[s], 1);
            mbar_init(&qempty[s], 5);  // the MMA thread and the four epilogue warps
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int nitems = Traits::nitems(p);

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&amap);
            prefetch_tmap(&bmap0);
            uint32_t stage = 0, phase = 0, qs = 0, qph = 0;
            for (;;) {
                // dynamic queue: the CTAs take the items in their global order, so the ~148 items in flight always
                // belong to neighbouring slabs of the operands and those stay in L2 (a static round-robin lets the
                // CTAs drift apart over ~600 items each: 27.9 ms instead of 16.6 for the C60 K GEMM)
                const int it = atomicAdd(Traits::counter(p), 1);
                mbar_wait(&qempty[qs], qph ^ 1);
                qitem[qs] = it < nitems ? it : -1;
                mbar_arrive(&qfull[qs]);
                if (++qs == I8_QRING) {
                    qs = 0;
                    qph ^= 1;
                }
                if (it >= nitems) break;
                I8Item x;
                Traits::decode(p, it, &amap, &bmap0, &bmap1, &bmap2, x);
                const uint32_t bytes = I8_A_STAGE + (uint32_t)x.ncols * I8_BK;
                for (int kb = 0; kb < x.nkb; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], bytes);
                    tma_load_3d(As + stage * I8_A_STAGE, x.amap, x.ax + kb * I8_BK, x.ay, x.az, &full[stage]);
                    tma_load_3d(Bs + stage * I8_B_STAGE, x.bmap, x.bx + kb * I8_BK, x.by, x.bz, &full[stage]);
                    if (++stage == I8_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0, qs = 0, qph = 0;
            for (;;) {
                mbar_wait(&qfull[qs], qph);
                const int it = qitem[qs];
                mbar_arrive(&qempty[qs]);
                if (++qs == I8_QRING) {
                    qs = 0;
                    qph ^= 1;
                }
                if (it < 0) break;
                I8Item x;
                Traits::decode(p, it, &amap, &bmap0, &bmap1, &bmap2, x);
                // instruction descriptor: D = s32, A = B = s8, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
                const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(x.ncols >> 3) << 17) | ((128u >> 4) << 24);
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem + as * I8_TN;
                for (int kb = 0; kb < x.nkb; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t ad = i8_smem_desc(smem_u32(As + stage * I8_A_STAGE));
                    const uint64_t bd = i8_smem_desc(smem_u32(Bs + stage * I8_B_STAGE));
#pragma unroll
                    for (int k = 0; k < 4; k++) mma_i8_ss(dcol, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    tc_commit(&empty[stage]);  // arrives when these MMAs have read the stage
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
        uint32_t as = 0, aphase = 0, qs = 0, qph = 0;
        const int lane_base = (warp & 3) * 32;  // the TMEM lanes this warp may read
        for (;;) {
            mbar_wait(&qfull[qs], qph);
            const int it = qitem[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&qempty[qs]);
            if (++qs == I8_QRING) {
                qs = 0;
                qph ^= 1;
            }
            if (it < 0) break;
            I8Item x;
            Traits::decode(p, it, &amap, &bmap0, &bmap1, &bmap2, x);
            const int ncols = x.ncols;
            const uint32_t pm = (uint32_t)c_i8.p[x.mod], magic = c_i8.magic[x.mod], off = c_i8.off[x.mod];
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)lane_base << 16) + as * I8_TN;
            const int row = lane_base + lane;
            const bool live = row < x.rows;
            uint8_t* dst = x.dst + (size_t)row * x.dst_pitch;
            uint32_t v[32];
            int c0 = 0;
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
            if (c0 < ncols) {  // ncols is a multiple of 16
                tmem_ld16(taddr + c0, v);
                uint4 w0;
                w0.x = i8_pack4(v + 0, pm, magic, off);
                w0.y = i8_pack4(v + 4, pm, magic, off);
                w0.z = i8_pack4(v + 8, pm, magic, off);
                w0.w = i8_pack4(v + 12, pm, magic, off);
                if (live) *reinterpret_cast<uint4*>(dst + c0) = w0;
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

// low bytes of four words -> one word
__device__ __forceinline__ uint32_t i8_pack_bytes(int a, int b, int c, int d) {
    return __byte_perm(__byte_perm((uint32_t)a, (uint32_t)b, 0x0040), __byte_perm((uint32_t)c, (uint32_t)d, 0x0040), 0x5410);
}

// ---- step 1: row norms and scales ------------------------------------------------------------------------
// part[row][chunk] = sum of squares over a contiguous k-chunk, fixed order (thread-strided, then a tree)
__global__ void __launch_bounds__(256) i8_rownorm_kernel(const double* __restrict__ T, size_t pitch, int kdim, int nchunk,
                                                         double* __restrict__ part) {
    const int row = blockIdx.y, chunk = blockIdx.x;
    const int len = (((kdim + nchunk - 1) / nchunk) + 1) & ~1;
    const int k0 = chunk * len, k1 = min(kdim, k0 + len);
    const double* r = T + (size_t)row * pitch;
    double s0 = 0, s1 = 0;
    if (((pitch & 1) == 0) && ((reinterpret_cast<uintptr_t>(T) & 15) == 0)) {
        for (int k = k0 + 2 * (int)threadIdx.x; k + 1 < k1; k += 512) {
            const double2 x = *reinterpret_cast<const double2*>(r + k);
            s0 = fma(x.x, x.x, s0);
            s1 = fma(x.y, x.y, s1);
        }
        if (((k1 - k0) & 1) && threadIdx.x == 0) s0 = fma(r[k1 - 1], r[k1 - 1], s0);
    } else {
        for (int k = k0 + (int)threadIdx.x; k < k1; k += 256) s0 = fma(r[k], r[k], s0);
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = s0 + s1;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)row * nchunk + chunk] = sh[0];
}

// e[row]: the largest exponent with |T[row,:]| * 2^e <= Rb (Rb already leaves room for the rounding of every element)
__global__ void i8_rowscale_kernel(const double* __restrict__ part, int nchunk, int nrows, double Rb, int* __restrict__ e) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    double s = 0;
    for (int c = 0; c < nchunk; c++) s += part[(size_t)row * nchunk + c];
    double nrm = sqrt(s) * (1.0 + 1e-12);
    int ex = 0;
    if (nrm > 0 && isfinite(nrm)) {
        int t;
        frexp(Rb / nrm, &t);  // Rb / nrm = f * 2^t, f in [0.5, 1): 2^(t-1) <= Rb / nrm
        ex = t - 1;
        if (ldexp(nrm, ex) > Rb) ex--;
        ex = max(-1000, min(1000, ex));
    }
    e[row] = ex;
}

// ---- step 2: residue planes --------------------------------------------------------------------------------
// planes[j][row][k] = (rint(T[row][k] * 2^e[row]) mod p_j) as a signed byte; one thread = four consecutive k
template <int NMOD>
__global__ void __launch_bounds__(256) i8_convert_kernel(const double* __restrict__ T, size_t pitch, int kdim,
                                                         const int* __restrict__ e, int8_t* __restrict__ planes, size_t ldk,
                                                         size_t plane_stride) {
    const int row = blockIdx.y;
    const int k = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (k >= kdim) return;
    const double* r = T + (size_t)row * pitch + k;
    double y[4];
    if (k + 3 < kdim && ((pitch & 1) == 0) && ((reinterpret_cast<uintptr_t>(T) & 15) == 0)) {
        const double2 a = *reinterpret_cast<const double2*>(r), b = *reinterpret_cast<const double2*>(r + 2);
        y[0] = a.x;
        y[1] = a.y;
        y[2] = b.x;
        y[3] = b.y;
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++) y[c] = (k + c < kdim) ? r[c] : 0.0;
    }
    // Rounding runs on the FP64 FMA pipe alone: (v + 1.5*2^52) - 1.5*2^52 = rint(v) for |v| < 2^51, and the low word of
    // (v + 1.5*2^52) is v mod 2^32 in two's complement (FRND/F2I.F64 go through the slow conversion pipe).  Per modulus
    // only q = rint(y / p) needs FP64 (one FMA): r = y - p q lies in [-127, 127], so its byte is (y - p q) mod 256 and
    // comes from the low words of y and q with one integer multiply-add (46 -> 14 FP64 operations per element).
    const double CM = 6755399441055744.0;
    const double scale = ldexp(1.0, e[row]);
    int8_t* dst = planes + (size_t)row * ldk + k;
    int ylo[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        y[c] = __dadd_rn(__fma_rn(y[c], scale, CM), -CM);  // |y| <= 2^51: exact integers
        ylo[c] = __double2loint(__dadd_rn(y[c], CM));
    }
    *reinterpret_cast<uint32_t*>(dst) = i8_pack_bytes(ylo[0], ylo[1], ylo[2], ylo[3]);  // p = 256: the low byte
#pragma unroll
    for (int j = 1; j < NMOD; j++) {
        const int pj = c_i8.p[j];
        const double inv = c_i8.invpd[j];
        int r[4];
#pragma unroll
        for (int c = 0; c < 4; c++) r[c] = ylo[c] - pj * __double2loint(__fma_rn(y[c], inv, CM));
        *reinterpret_cast<uint32_t*>(dst + (size_t)j * plane_stride) = i8_pack_bytes(r[0], r[1], r[2], r[3]);
    }
}

// ---- step 4: sum over k-ranges, CRT, scale, accumulate ------------------------------------------------------
struct I8CrtParams {
    int ntile, nsplit, nbf, symmetric, ldK;
    const I8Tile* tiles;
    const uint8_t* ws;
    const int *eA, *eB;
    double* K;
    unsigned long long M_lo, M_hi, H_lo, H_hi;  // M = product of the moduli, H = M / 2 (values above H are negative)
};

__device__ __forceinline__ int i8_modp_small(int t, int p, float invp) {  // 0 <= t < 2^22
    const int q = (int)((float)t * invp);
    int r = t - q * p;
    if (r < 0) r += p;
    if (r >= p) r -= p;
    return r;
}

// The integer whose residues (any non-negative representatives < 2^22) are r[0..NMOD): Garner's mixed-radix digits
// v_j = (..((r_j - v_0) / p_0 - v_1) / p_1 ..) mod p_j, then value = v_0 + p_0 (v_1 + p_1 (v_2 + ...)) in 128 bits, taken
// in (-M/2, M/2], converted to double (M = product of the moduli, H = M / 2).
template <int NMOD>
__device__ __forceinline__ double i8_crt_value(const int (&r)[NMOD], unsigned long long M_lo, unsigned long long M_hi,
                                               unsigned long long H_lo, unsigned long long H_hi) {
    int v[NMOD];
#pragma unroll
    for (int j = 0; j < NMOD; j++) {
        const int pj = c_i8.p[j];
        const float ij = c_i8.invp[j];
        int x = i8_modp_small(r[j], pj, ij);
#pragma unroll
        for (int i = 0; i < j; i++) x = i8_modp_small((x + 2 * pj - v[i]) * c_i8.ginv[i][j], pj, ij);
        v[j] = x;
    }
    unsigned long long lo = (unsigned long long)v[NMOD - 1], hi = 0;
#pragma unroll
    for (int j = NMOD - 2; j >= 0; j--) {
        const unsigned long long pj = (unsigned long long)c_i8.p[j];
        const unsigned long long l2 = lo * pj;
        hi = hi * pj + __umul64hi(lo, pj);
        lo = l2 + (unsigned long long)v[j];
        if (lo < l2) hi++;
    }
    const bool neg = hi > H_hi || (hi == H_hi && lo > H_lo);
    if (neg) {  // M - value
        const unsigned long long l2 = M_lo - lo;
        hi = M_hi - hi - (M_lo < lo ? 1ull : 0ull);
        lo = l2;
    }
    const double d = (double)hi * 18446744073709551616.0 + (double)lo;
    return neg ? -d : d;
}

template <int NMOD>
__global__ void __launch_bounds__(256) i8_crt_kernel(const I8CrtParams p) {
    const int tile = blockIdx.y;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 6), col4 = (threadIdx.x & 63) * 4;
    const I8Tile t = p.tiles[tile];
    const int gm = t.row0 + row;
    if (col4 >= t.ncols || gm >= p.nbf) return;
    int r[NMOD][4];
#pragma unroll
    for (int j = 0; j < NMOD; j++) r[j][0] = r[j][1] = r[j][2] = r[j][3] = 0;
    const uint8_t* src = p.ws + (size_t)tile * I8_TILE_BYTES + (size_t)row * I8_TN + col4;
    for (int s = 0; s < p.nsplit; s++) {
#pragma unroll
        for (int j = 0; j < NMOD; j++) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(src + ((size_t)s * NMOD + j) * p.ntile * I8_TILE_BYTES);
            r[j][0] += w & 255u;
            r[j][1] += (w >> 8) & 255u;
            r[j][2] += (w >> 16) & 255u;
            r[j][3] += w >> 24;
        }
    }
    const int ea = p.eA[gm];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int gn = t.col0 + col4 + c;
        if (gn >= p.nbf || (p.symmetric && gn < gm)) continue;
        int rr[NMOD];
#pragma unroll
        for (int j = 0; j < NMOD; j++) rr[j] = r[j][c];
        const double d = i8_crt_value<NMOD>(rr, p.M_lo, p.M_hi, p.H_lo, p.H_hi);
        const double val = ldexp(d, -(ea + p.eB[gn]));
        double* kp = p.K + (size_t)gm * p.ldK + gn;
        const double nv = *kp + val;
        *kp = nv;
        if (p.symmetric && gn != gm) p.K[(size_t)gn * p.ldK + gm] = nv;
    }
}

// value = the integer in (-M/2, M/2] with residues r[0..NMOD) (representatives in [0, 256)), as a double good to
// ~2^-67 M absolute -- four orders of magnitude below the rounding of the operands that produced the residues.  The
// exact Garner / 128-bit route of i8_crt_value costs ~1000 instructions per element, fine for the nbf^2 elements of K but
// not for the nbf*naux*nocc elements of the half transform; this one costs ~70: s_j = r_j q_j (|s_j| < 2^15), the high parts
// s_j ih_j are multiples of 2^-32 below 2^7 and sum exactly in double, the low parts carry the rest.
template <int NMOD>
__device__ __forceinline__ double i8_crt_value_fast(const int (&r)[NMOD]) {
    double fh = 0.0, fl = 0.0;
#pragma unroll
    for (int j = 0; j < NMOD; j++) {
        const int s = r[j] * c_i8f.q[j] + 32768;  // in [0, 65536)
        const double sd = __hiloint2double(0x43300000, s) - 4503599627403264.0;  // 2^52 + 32768
        fh = __fma_rn(sd, c_i8f.ih[j], fh);
        fl = __fma_rn(sd, c_i8f.il[j], fl);
    }
    const double CM = 6755399441055744.0;
    fh -= __dadd_rn(__dadd_rn(fh, CM), -CM);
    double f = fh + fl;
    f -= __dadd_rn(__dadd_rn(f, CM), -CM);
    return f * c_i8f.Md;
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*I8EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct I8Plan {  // per device: buffers grown on demand, kept across builds
    int8_t* planes[2] = {nullptr, nullptr};
    size_t planes_cap[2] = {0, 0};
    double* normpart = nullptr;
    size_t normpart_cap = 0;
    int* expo = nullptr;  // [2][nbf]
    size_t expo_cap = 0;
    uint8_t* ws = nullptr;
    size_t ws_cap = 0;
    I8Tile* d_tiles = nullptr;
    int* d_counter = nullptr;
    int tiles_nbf = -1, tiles_sym = -1, ntile = 0;
    double tile_area = 0;  // sum over the tile list of 128 x (padded columns)
    int bbox[3] = {256, 0, 0};
    bool consts = false, attr = false;
    uint64_t launches = 0;
    cudaEvent_t prof[4] = {nullptr, nullptr, nullptr, nullptr};  // optional: start / planes done / GEMM done / CRT done of pass 0
    // optional phase marker of the caller (the engine records an event on the stream): 20 = a pass starts, 21 = its planes
    // are written, 22 = its GEMM is launched, 23 = its CRT is launched
    void (*mark)(void* ctx, int tag) = nullptr;
    void* mark_ctx = nullptr;
    void release() {
        for (int i = 0; i < 2; i++) {
            if (planes[i]) cudaFree(planes[i]);
            planes[i] = nullptr;
            planes_cap[i] = 0;
        }
        if (normpart) cudaFree(normpart);
        if (expo) cudaFree(expo);
        if (ws) cudaFree(ws);
        if (d_tiles) cudaFree(d_tiles);
        if (d_counter) cudaFree(d_counter);
        d_counter = nullptr;
        normpart = nullptr;
        expo = nullptr;
        ws = nullptr;
        d_tiles = nullptr;
        normpart_cap = expo_cap = ws_cap = 0;
        tiles_nbf = tiles_sym = -1;
    }
};

inline int i8_egcd_inv(int a, int m) {  // a^-1 mod m
    int g = m, x = 0, y = 1, aa = a % m;
    while (aa) {
        int q = g / aa, t = g - q * aa;
        g = aa;
        aa = t;
        t = x - q * y;
        x = y;
        y = t;
    }
    return ((x % m) + m) % m;
}

inline void i8_fill_consts(I8Consts& c) {
    for (int i = 0; i < 16; i++) {
        const int p = kI8Moduli[i];
        c.p[i] = p;
        c.invp[i] = 1.0f / (float)p;
        c.invpd[i] = 1.0 / (double)p;
        c.magic[i] = (unsigned)(((unsigned long long)1 << 32) / (unsigned)p);
        c.off[i] = (unsigned)((((1u << 30) + (unsigned)p - 1) / (unsigned)p) * (unsigned)p);
        for (int j = 0; j < 16; j++) c.ginv[i][j] = (i < j) ? i8_egcd_inv(p, kI8Moduli[j]) : 0;
    }
}

inline void i8_fill_crt_fast(I8CrtFast& f, int nmod) {
    unsigned __int128 M = 1;
    for (int i = 0; i < nmod; i++) M *= (unsigned)kI8Moduli[i];
    for (int j = 0; j < 16; j++) {
        f.q[j] = 0;
        f.ih[j] = f.il[j] = 0.0;
    }
    for (int j = 0; j < nmod; j++) {
        const int p = kI8Moduli[j];
        const int mj = (int)((M / (unsigned)p) % (unsigned)p);
        int q = i8_egcd_inv(mj, p);
        if (q > p / 2) q -= p;
        f.q[j] = q;
        const long double inv = 1.0L / (long double)p;
        const double ih = (double)(rintl(inv * 4294967296.0L) / 4294967296.0L);
        f.ih[j] = ih;
        f.il[j] = (double)(inv - (long double)ih);
    }
    f.Md = (double)(long double)M;
    f.nmod = nmod;
}

// Rb(nmod, kdim): bound on the 2-norm of a scaled row BEFORE rounding, so that after rounding (each element moves by at
// most 1/2, the norm by at most sqrt(kdim)/2) the norm is <= R, R^2 < M / 2 and R <= 2^51
inline double i8_row_bound(int nmod, int kdim, unsigned __int128* Mout) {
    unsigned __int128 M = 1;
    for (int i = 0; i < nmod; i++) M *= (unsigned)kI8Moduli[i];
    if (Mout) *Mout = M;
    long double R = sqrtl((long double)(M / 2)) * (1.0L - 1e-9L);
    const long double cap = 2251799813685248.0L;  // 2^51
    if (R > cap) R = cap;
    R -= 0.5L * sqrtl((long double)kdim) + 1.0L;
    return (double)R;
}

struct I8RunInfo {
    int nmod = 0, klen = 0, nsplit = 0, ntile = 0, passes = 0;
    double bits = 0;
    double mma_ops = 0;  // 2 x int8 multiply-adds issued to the tensor cores (all moduli, whole padded tiles and k-blocks)
};

#define I8CK(call)                                                                                          \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return e_ == cudaErrorMemoryAllocation ? 3 : 2;                                                 \
        }                                                                                                   \
    } while (0)

template <class T>
inline int i8_grow(T** p, size_t* cap, size_t need, std::string* err) {
    if (need <= *cap) return 0;
    if (*p) I8CK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    I8CK(cudaMalloc((void**)p, need * sizeof(T)));
    *cap = need;
    return 0;
}

inline int i8_make_map(I8EncodeFn enc, CUtensorMap* out, const int8_t* base, uint64_t kdim, uint64_t rows, uint64_t nplanes,
                       uint64_t ldk, uint64_t plane_stride, uint32_t box_rows, std::string* err) {
    cuuint64_t dims[3] = {kdim, rows, nplanes};
    cuuint64_t strides[2] = {ldk, plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)I8_BK, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled (int8 planes) failed: " + std::to_string((int)r);
        return 2;
    }
    return 0;
}

template <int NMOD>
inline void i8_launch_convert(const double* T, size_t pitch, int kdim, int nrows, const int* e, int8_t* planes, size_t ldk,
                              size_t plane_stride, cudaStream_t st) {
    dim3 grid((kdim + 1023) / 1024, nrows);
    i8_convert_kernel<NMOD><<<grid, 256, 0, st>>>(T, pitch, kdim, e, planes, ldk, plane_stride);
}
inline void i8_convert(int nmod, const double* T, size_t pitch, int kdim, int nrows, const int* e, int8_t* planes, size_t ldk,
                       size_t plane_stride, cudaStream_t st) {
    switch (nmod) {
        case 6: i8_launch_convert<6>(T, pitch, kdim, nrows, e, planes, ldk, plane_stride, st); break;
        case 7: i8_launch_convert<7>(T, pitch, kdim, nrows, e, planes, ldk, plane_stride, st); break;
        case 8: i8_launch_convert<8>(T, pitch, kdim, nrows, e, planes, ldk, plane_stride, st); break;
        case 9: i8_launch_convert<9>(T, pitch, kdim, nrows, e, planes, ldk, plane_stride, st); break;
        case 10: i8_launch_convert<10>(T, pitch, kdim, nrows, e, planes, ldk, plane_stride, st); break;
        case 11: i8_launch_convert<11>(T, pitch, kdim, nrows, e, planes, ldk, plane_stride, st); break;
        case 12: i8_launch_convert<12>(T, pitch, kdim, nrows, e, planes, ldk, plane_stride, st); break;
        default: i8_launch_convert<13>(T, pitch, kdim, nrows, e, planes, ldk, plane_stride, st); break;
    }
}
inline void i8_crt(int nmod, const I8CrtParams& p, cudaStream_t st) {
    dim3 grid(I8_TM / 4, p.ntile);
    switch (nmod) {
        case 6: i8_crt_kernel<6><<<grid, 256, 0, st>>>(p); break;
        case 7: i8_crt_kernel<7><<<grid, 256, 0, st>>>(p); break;
        case 8: i8_crt_kernel<8><<<grid, 256, 0, st>>>(p); break;
        case 9: i8_crt_kernel<9><<<grid, 256, 0, st>>>(p); break;
        case 10: i8_crt_kernel<10><<<grid, 256, 0, st>>>(p); break;
        case 11: i8_crt_kernel<11><<<grid, 256, 0, st>>>(p); break;
        case 12: i8_crt_kernel<12><<<grid, 256, 0, st>>>(p); break;
        default: i8_crt_kernel<13><<<grid, 256, 0, st>>>(p); break;
    }
}

// K[m][n] (+)= sum_k T1[m][k] T2[n][k], m, n < nbf, k < kdim; T row pitch `pitch` doubles; symmetric: T2 == T1, the
// upper triangle is computed and mirrored.  plane_budget: bytes the residue planes may take (the k index is processed
// in passes when they do not fit).  Returns 0, or 2 (CUDA) / 3 (out of memory) with a message in *err.
inline int i8_kgemm_run(I8Plan& pl, I8EncodeFn enc, cudaStream_t st, int nsm, const double* T1, const double* T2, size_t pitch,
                        int nbf, int kdim, bool symmetric, double* K, int ldK, int nmod, int klen, size_t plane_budget,
                        I8RunInfo* info, std::string* err) {
    nmod = std::max(I8_MINMOD, std::min(I8_MAXMOD, nmod));
    klen = std::max(I8_BK, std::min(I8_MAX_KLEN, (klen / I8_BK) * I8_BK));
    if (!pl.consts) {
        I8Consts c;
        i8_fill_consts(c);
        I8CK(cudaMemcpyToSymbolAsync(c_i8, &c, sizeof c, 0, cudaMemcpyHostToDevice, st));
        I8CK(cudaStreamSynchronize(st));  // c lives on this stack frame
        pl.consts = true;
    }
    if (!pl.d_counter) I8CK(cudaMalloc((void**)&pl.d_counter, sizeof(int)));
    if (!pl.attr) {
        I8CK(cudaFuncSetAttribute(i8_pipeline_kernel<I8KgemmTraits>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)i8_smem_bytes()));
        pl.attr = true;
    }
    // tile list: 128 rows of T1 against column segments of <= 256 rows of T2 (symmetric: from the row block's own
    // first row on); the remainder segment is rounded up to 16 columns and gets a map with its own box
    if (pl.tiles_nbf != nbf || pl.tiles_sym != (symmetric ? 1 : 0)) {
        std::vector<I8Tile> tiles;
        pl.bbox[0] = 256;
        pl.bbox[1] = pl.bbox[2] = 0;
        for (int row0 = 0; row0 < nbf; row0 += I8_TM) {
            const int cstart = symmetric ? row0 : 0;
            for (int col0 = cstart; col0 < nbf; col0 += I8_TN) {
                const int w = std::min(I8_TN, ((nbf - col0) + 15) / 16 * 16);
                int which = 0;
                if (w != 256) {
                    if (pl.bbox[1] == 0 || pl.bbox[1] == w) {
                        pl.bbox[1] = w;
                        which = 1;
                    } else {
                        pl.bbox[2] = w;  // at most two remainder widths exist (row0 mod 256 is 0 or 128)
                        which = 2;
                    }
                }
                tiles.push_back(I8Tile{row0, col0, w, which});
            }
        }
        pl.tile_area = 0;
        for (auto& t : tiles) pl.tile_area += (double)I8_TM * t.ncols;
        if (pl.d_tiles) I8CK(cudaFree(pl.d_tiles));
        pl.d_tiles = nullptr;
        I8CK(cudaMalloc((void**)&pl.d_tiles, tiles.size() * sizeof(I8Tile)));
        I8CK(cudaMemcpyAsync(pl.d_tiles, tiles.data(), tiles.size() * sizeof(I8Tile), cudaMemcpyHostToDevice, st));
        I8CK(cudaStreamSynchronize(st));
        pl.ntile = (int)tiles.size();
        pl.tiles_nbf = nbf;
        pl.tiles_sym = symmetric ? 1 : 0;
    }
    const int nop = symmetric ? 1 : 2;
    // k passes
    const size_t per_k = (size_t)nmod * nbf * nop;
    long kpass = (long)std::min<size_t>((size_t)kdim, std::max<size_t>(plane_budget / per_k, (size_t)klen));
    if (kpass < kdim) kpass = std::max<long>(klen, kpass / klen * klen);
    const int nchunk = 16;
    int rc;
    if ((rc = i8_grow(&pl.normpart, &pl.normpart_cap, (size_t)nbf * nchunk, err))) return rc;
    if ((rc = i8_grow(&pl.expo, &pl.expo_cap, (size_t)2 * nbf, err))) return rc;
    unsigned __int128 M;
    int passes = 0;
    for (long kb0 = 0; kb0 < kdim; kb0 += kpass, passes++) {
        const int kk = (int)std::min<long>(kpass, kdim - kb0);
        const size_t ldk = ((size_t)kk + 127) / 128 * 128;
        const size_t plane_stride = ldk * nbf;
        const double Rb = i8_row_bound(nmod, kk, &M);
        const int nsplit = (kk + klen - 1) / klen;
        const size_t nitems = (size_t)nsplit * nmod * pl.ntile;
        if ((rc = i8_grow(&pl.ws, &pl.ws_cap, nitems * I8_TILE_BYTES, err))) return rc;
        CUtensorMap amap, bmap[3];
        if (pl.prof[0] && passes == 0) cudaEventRecord(pl.prof[0], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 20);
        for (int op = 0; op < nop; op++) {
            const double* T = (op == 0 ? T1 : T2) + kb0;
            if ((rc = i8_grow(&pl.planes[op], &pl.planes_cap[op], plane_stride * nmod, err))) return rc;
            i8_rownorm_kernel<<<dim3(nchunk, nbf), 256, 0, st>>>(T, pitch, kk, nchunk, pl.normpart);
            i8_rowscale_kernel<<<(nbf + 127) / 128, 128, 0, st>>>(pl.normpart, nchunk, nbf, Rb, pl.expo + (size_t)op * nbf);
            i8_convert(nmod, T, pitch, kk, nbf, pl.expo + (size_t)op * nbf, pl.planes[op], ldk, plane_stride, st);
            pl.launches += 3;
        }
        if (pl.prof[1] && passes == 0) cudaEventRecord(pl.prof[1], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 21);
        const int8_t* pb = pl.planes[symmetric ? 0 : 1];
        if ((rc = i8_make_map(enc, &amap, pl.planes[0], (uint64_t)kk, (uint64_t)nbf, (uint64_t)nmod, ldk, plane_stride, I8_TM, err)))
            return rc;
        for (int b = 0; b < 3; b++) {
            const int box = pl.bbox[b] ? pl.bbox[b] : 16;
            if ((rc = i8_make_map(enc, &bmap[b], pb, (uint64_t)kk, (uint64_t)nbf, (uint64_t)nmod, ldk, plane_stride, (uint32_t)box, err)))
                return rc;
        }
        I8GemmParams gp;
        gp.nitems = (int)nitems;
        gp.ntile = pl.ntile;
        gp.nmod = nmod;
        gp.kdim = kk;
        gp.klen = klen;
        gp.tiles = pl.d_tiles;
        gp.ws = pl.ws;
        gp.counter = pl.d_counter;
        I8CK(cudaMemsetAsync(pl.d_counter, 0, sizeof(int), st));
        i8_pipeline_kernel<I8KgemmTraits><<<(unsigned)std::min<size_t>(nitems, (size_t)nsm), I8_THREADS, i8_smem_bytes(), st>>>(amap, bmap[0], bmap[1],
                                                                                                            bmap[2], gp);
        if (pl.prof[2] && passes == 0) cudaEventRecord(pl.prof[2], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 22);
        I8CrtParams cp;
        cp.ntile = pl.ntile;
        cp.nsplit = nsplit;
        cp.nbf = nbf;
        cp.symmetric = symmetric ? 1 : 0;
        cp.ldK = ldK;
        cp.tiles = pl.d_tiles;
        cp.ws = pl.ws;
        cp.eA = pl.expo;
        cp.eB = pl.expo + (symmetric ? 0 : (size_t)nbf);
        cp.K = K;
        cp.M_lo = (unsigned long long)M;
        cp.M_hi = (unsigned long long)(M >> 64);
        cp.H_lo = (unsigned long long)(M / 2);
        cp.H_hi = (unsigned long long)((M / 2) >> 64);
        i8_crt(nmod, cp, st);
        if (pl.prof[3] && passes == 0) cudaEventRecord(pl.prof[3], st);
        if (pl.mark) pl.mark(pl.mark_ctx, 23);
        pl.launches += 2;
        I8CK(cudaGetLastError());
        if (info) {
            info->nsplit = nsplit;
            info->bits = log2(Rb);
            info->mma_ops += 2.0 * nmod * (double)ldk * pl.tile_area;
        }
    }
    if (info) {
        info->nmod = nmod;
        info->klen = klen;
        info->ntile = pl.ntile;
        info->passes = passes;
    }
    return 0;
}

}  // namespace b2k
