// dmma_gemm.cuh -- FP64 tensor-core (DMMA m8n8k4) tile engine shared by the two compute-bound
// kernels of the DF-K build:
//
//   K3  half_transform : T[m,q,i] = sum_k B_m[q,k] * C[n_k,i]      (lib3index/dfhelper.cc:2162-2186)
//   K4  k_gemm         : K[m,n]  += sum_{q,i} T1[m,qi] * T2[n,qi]  (lib3index/dfhelper.cc:3374)
//
// Both are "NT" products of two K-contiguous operands, which is exactly the row.col operand
// order of mma.sync.m8n8k4.f64.  sm_100a has no tcgen05 kind for f64 (SASS: DMMA.8x8x4), so the
// pipeline is: cp.async (LDGSTS) 4-stage ring -> padded shared tiles (conflict-free 8-byte
// fragment loads) -> warp-level DMMA with register accumulators.
//
// CTA tile: 128 x (16*NB) x 16, 8 warps as 4 (M) x 2 (N); warp tile 32 x (8*NB).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2k {

constexpr int BM = 128;       // CTA rows (A operand)
constexpr int BK = 16;        // k per stage
constexpr int LDT = 20;       // padded smem row pitch in doubles: (g*20 + t) mod 16 distinct for g<4,t<4
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 256;

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
// 16-byte copy through L1 (the gathered C^T rows are re-read by every item of a row-block); src_bytes < 16 zero-fills
__device__ __forceinline__ void cp_async16_ca(double* smem, const double* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ROWS x BK tile of a row-major, k-contiguous matrix.  Rows >= nrows and k >= kend are zero-filled
// (cp.async src-size), so edge tiles and ragged k tails need no separate code path.
// Requires: g 16-byte aligned, ld even, k0 even.
template <int ROWS>
__device__ __forceinline__ void load_tile_plain(double* s, const double* __restrict__ g, size_t ld, int row0,
                                                int nrows, int k0, int kend, int tid) {
#pragma unroll
    for (int id = tid; id < ROWS * 8; id += GEMM_THREADS) {
        int r = id >> 3, c = id & 7;
        int gr = row0 + r, gk = k0 + c * 2;
        int rem = kend - gk;
        int bytes = (gr < nrows) ? (rem >= 2 ? 16 : (rem == 1 ? 8 : 0)) : 0;
        const double* src = bytes ? g + (size_t)gr * ld + gk : g;
        cp_async16(s + r * LDT + c * 2, src, bytes);
    }
}

// ROWS x BK tile gathered along k: s[r][kk] = g[(row0+r)*ld + col], col = the kept-partner index
// of position k0+kk (or -1 past the end).  Every thread owns one fixed kk (256 % 16 == 0), so the
// index is fetched once per stage by the caller and passed in.
template <int ROWS>
__device__ __forceinline__ void load_tile_gather(double* s, const double* __restrict__ g, size_t ld, int row0,
                                                 int nrows, int col, int tid) {
    int kk = tid & 15;
#pragma unroll
    for (int r = tid >> 4; r < ROWS; r += GEMM_THREADS / 16) {
        int gr = row0 + r;
        int bytes = (col >= 0 && gr < nrows) ? 8 : 0;
        const double* src = bytes ? g + (size_t)gr * ld + col : g;
        cp_async8(s + r * LDT + kk, src, bytes);
    }
}

// One BK=16 stage of DMMAs for this warp.  FULL: every 8x8 block of the warp tile is live -> no
// predicates at all (a predicated mma.sync costs WARPSYNC + branch per DMMA and serialises the pipe).
// Otherwise mbv / nbv = number of live 8-row / 8-col blocks (warp-uniform): dead edge blocks are skipped.
template <int NB, bool FULL>
__device__ __forceinline__ void mma_stage(const double* __restrict__ As, const double* __restrict__ Bs,
                                          double (&acc)[4][NB][2], int wm, int wn, int g, int t, int mbv, int nbv) {
    const double* Ap = As + (wm * 32 + g) * LDT + t;
    const double* Bp = Bs + (wn * 8 * NB + g) * LDT + t;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
        double a[4], b[NB];
#pragma unroll
        for (int mb = 0; mb < 4; mb++) a[mb] = Ap[mb * 8 * LDT + ks * 4];
#pragma unroll
        for (int nb = 0; nb < NB; nb++) b[nb] = Bp[nb * 8 * LDT + ks * 4];
        if (FULL) {
#pragma unroll
            for (int mb = 0; mb < 4; mb++)
#pragma unroll
                for (int nb = 0; nb < NB; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
        } else {
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
                if (mb < mbv) {
#pragma unroll
                    for (int nb = 0; nb < NB; nb++)
                        if (nb < nbv) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
                }
            }
        }
    }
}

template <int NB>
constexpr size_t gemm_smem_bytes() {
    return (size_t)STAGES * (BM + 16 * NB) * LDT * sizeof(double);
}

// ---------------------------------------------------------------------------------------------
// K3: half transform.  grid = (n_itiles, n_qtiles, nbf).  One CTA: row-block m, rows
// [q0,q0+128) of this shard's Q chunk, orbital columns [i0, i0+iw).
//   Bm   : packed row-block m of the tensor, [nq_shard x ldm] (ldm = padded sp(m))
//   Ct   : C transposed, [o x ldc]; gathered along n through cols (kept partners of m)
//   T    : [nbf][qc][op] chunk-local intermediate
// DENSE rows (sp(m) == nbf) skip the gather and use 16-byte copies.
// ---------------------------------------------------------------------------------------------
struct HalfParams {
    const double* tensor;       // shard tensor base
    const size_t* row_off;      // [nbf] offset (doubles) of row-block m in the shard tensor
    const int* ldm;             // [nbf] padded row pitch
    const int* sp;              // [nbf] kept partners
    const int* cols;            // concatenated kept-partner lists
    const size_t* cols_off;     // [nbf] offset into cols
    const double* Ct;           // [o x ldc]
    int ldc;
    int o;                      // occupied count (valid rows of Ct)
    int op;                     // padded o (even) = T inner pitch
    int iw;                     // orbital columns per i-tile (even)
    int qbeg;                   // first shard-local q row of this chunk
    int qc;                     // rows in this chunk
    int nbf;
    double* T;                  // [nbf][qc][op]
};

template <int NB>
__global__ void __launch_bounds__(GEMM_THREADS, 1) half_transform_kernel(HalfParams p) {
    extern __shared__ __align__(16) double smem[];
    const int m = blockIdx.z;
    const int q0 = blockIdx.y * BM;
    const int i0 = blockIdx.x * p.iw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2, g = lane >> 2, t = lane & 3;
    constexpr int BN = 16 * NB;

    const int K = p.sp[m];
    const int ldm = p.ldm[m];
    const double* A = p.tensor + p.row_off[m] + (size_t)p.qbeg * ldm;
    const int* cols = p.cols + p.cols_off[m];
    const bool dense = (K == p.nbf);
    const int icols = min(p.iw, p.o - i0);  // live orbital columns in this tile
    const int nkt = (K + BK - 1) / BK;

    double* As = smem;
    double* Bs = smem + STAGES * BM * LDT;

    const int mbv = max(0, min(4, (p.qc - (q0 + wm * 32) + 7) / 8));
    const int nbv = max(0, min(NB, (icols - wn * 8 * NB + 7) / 8));
    const bool full = (mbv == 4) && (nbv == NB);

    double acc[4][NB][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < NB; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int kk = tid & 15;
    auto fetch_col = [&](int kt) -> int {
        int k = kt * BK + kk;
        return (kt < nkt && k < K) ? __ldg(cols + k) : -1;
    };
    auto issue = [&](int kt, int col) {
        if (kt < nkt) {
            int st = kt % STAGES;
            load_tile_plain<BM>(As + st * BM * LDT, A, (size_t)ldm, q0, p.qc, kt * BK, K, tid);
            if (dense)
                load_tile_plain<BN>(Bs + st * BN * LDT, p.Ct, (size_t)p.ldc, i0, i0 + icols, kt * BK, K, tid);
            else
                load_tile_gather<BN>(Bs + st * BN * LDT, p.Ct, (size_t)p.ldc, i0, i0 + icols, col, tid);
        }
        cp_async_commit();
    };

    int col_next = dense ? -1 : fetch_col(0);
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        int col = col_next;
        if (!dense) col_next = fetch_col(s + 1);
        issue(s, col);
    }
    for (int kt = 0; kt < nkt; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        int col = col_next;
        if (!dense) col_next = fetch_col(kt + STAGES);
        issue(kt + STAGES - 1, col);
        int st = kt % STAGES;
        if (full)
            mma_stage<NB, true>(As + st * BM * LDT, Bs + st * BN * LDT, acc, wm, wn, g, t, mbv, nbv);
        else
            mma_stage<NB, false>(As + st * BM * LDT, Bs + st * BN * LDT, acc, wm, wn, g, t, mbv, nbv);
    }
    cp_async_wait<0>();

    // epilogue: T[m][q][i], 16-byte stores (column index even, op even)
    double* Tm = p.T + (size_t)m * p.qc * p.op;
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
        int q = q0 + wm * 32 + mb * 8 + g;
        if (q < p.qc) {
#pragma unroll
            for (int nb = 0; nb < NB; nb++) {
                int ic = wn * 8 * NB + nb * 8 + t * 2;  // column inside the tile
                int i = i0 + ic;
                if (ic < p.iw && i < p.op) {
                    double2 v = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
                    *reinterpret_cast<double2*>(Tm + (size_t)q * p.op + i) = v;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4: K GEMM, split-K.  grid = (ntiles, nsplit).  Tile list covers the full square, or only the
// upper triangle of tiles when T2 == T1 (lr_symmetric); the reduction kernel mirrors.
// Each CTA writes its 128x128 partial to ws[split][tile]; kgemm_reduce_kernel sums the splits in
// fixed order (deterministic) and accumulates into K.
// ---------------------------------------------------------------------------------------------
struct KgemmParams {
    const double* T1;   // [nbf][kdim]
    const double* T2;   // [nbf][kdim]
    int nbf;
    int kdim;           // qc*op (even)
    int klen;           // k per split (multiple of BK)
    int ntile1d;        // ceil(nbf/128)
    int symmetric;      // tiles: upper triangle only
    double* ws;         // [nsplit][ntiles][128*128]
};

__device__ __forceinline__ void tile_coords(int tile, int n1d, int symmetric, int& tm, int& tn) {
    if (!symmetric) {
        tm = tile / n1d;
        tn = tile % n1d;
    } else {
        // row-major enumeration of the upper triangle
        int r = 0, rem = tile;
        while (rem >= n1d - r) {
            rem -= n1d - r;
            r++;
        }
        tm = r;
        tn = r + rem;
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) kgemm_kernel(KgemmParams p) {
    extern __shared__ __align__(16) double smem[];
    constexpr int NB = 8, BN = 128;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2, g = lane >> 2, t = lane & 3;
    int tm, tn;
    tile_coords(blockIdx.x, p.ntile1d, p.symmetric, tm, tn);
    const int m0 = tm * BM, n0 = tn * BN;
    const int kb = blockIdx.y * p.klen;
    const int ke = min(p.kdim, kb + p.klen);
    const int nkt = (ke - kb + BK - 1) / BK;

    double* As = smem;
    double* Bs = smem + STAGES * BM * LDT;
    const int mbv = max(0, min(4, (p.nbf - (m0 + wm * 32) + 7) / 8));
    const int nbv = max(0, min(NB, (p.nbf - (n0 + wn * 8 * NB) + 7) / 8));
    const bool full = (mbv == 4) && (nbv == NB);

    double acc[4][NB][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < NB; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto issue = [&](int kt) {
        if (kt < nkt) {
            int st = kt % STAGES;
            load_tile_plain<BM>(As + st * BM * LDT, p.T1, (size_t)p.kdim, m0, p.nbf, kb + kt * BK, ke, tid);
            load_tile_plain<BN>(Bs + st * BN * LDT, p.T2, (size_t)p.kdim, n0, p.nbf, kb + kt * BK, ke, tid);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) issue(s);
    for (int kt = 0; kt < nkt; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(kt + STAGES - 1);
        int st = kt % STAGES;
        if (full)
            mma_stage<NB, true>(As + st * BM * LDT, Bs + st * BN * LDT, acc, wm, wn, g, t, mbv, nbv);
        else
            mma_stage<NB, false>(As + st * BM * LDT, Bs + st * BN * LDT, acc, wm, wn, g, t, mbv, nbv);
    }
    cp_async_wait<0>();

    double* w = p.ws + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (BM * BN);
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
        int r = wm * 32 + mb * 8 + g;
#pragma unroll
        for (int nb = 0; nb < NB; nb++) {
            int c = wn * 8 * NB + nb * 8 + t * 2;
            *reinterpret_cast<double2*>(w + r * BN + c) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
        }
    }
}

// K[m][n] += sum_s ws[s][tile][r][c]  (fixed order); mirrored for off-diagonal tiles when symmetric.
__global__ void kgemm_reduce_kernel(const double* __restrict__ ws, int nsplit, int ntiles, int ntile1d, int symmetric,
                                    int nbf, double* __restrict__ K) {
    const int tile = blockIdx.x;
    int tm, tn;
    tile_coords(tile, ntile1d, symmetric, tm, tn);
    for (int e = threadIdx.x; e < BM * 128; e += blockDim.x) {
        int r = e >> 7, c = e & 127;
        int m = tm * BM + r, n = tn * 128 + c;
        if (m < nbf && n < nbf) {
            double s = 0.0;
            for (int sp = 0; sp < nsplit; sp++) s += ws[((size_t)sp * ntiles + tile) * (BM * 128) + e];
            K[(size_t)m * nbf + n] += s;
            if (symmetric && tn > tm) K[(size_t)n * nbf + m] += s;
        }
    }
}

}  // namespace b2k
