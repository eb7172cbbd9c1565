// j_kernels.cuh -- the two HBM-bound sweeps of the DF-J build over the packed (Q|mn) tensor.
//
//   K1  j_dq   : d[q]    = sum_m sum_k B_m[q,k] * Dp_m[k]      (DGEMV 'N', dfhelper.cc:3193 / :3258)
//   K2  j_mn   : J[m,n_k] = sum_q B_m[q,k] * d[q]  + unpack    (DGEMV 'T' + loops, :3208-3221 / :3272-3283)
//
// Symmetric densities (lr_symmetric) read only the columns n >= m of each row-block (start column
// ign(m)), weight 2 off the diagonal, and mirror on unpack -- exactly compute_J_symm.  General
// densities read every kept column -- compute_J.  Both are pure streaming reads of B: every
// warp-level load is one contiguous 512-byte (K1) or 2 x 512-byte (K2) segment.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2k {

struct JParams {
    const double* tensor;     // shard tensor base
    const size_t* row_off;    // [nbf]
    const int* ldm;           // [nbf]
    const int* sp;            // [nbf]
    const int* ign;           // [nbf] kept partners with n < m
    const int* cols;
    const size_t* cols_off;
    int nbf;
    int nq;                   // shard rows
    int symmetric;
    const double* D;          // [nbf][nbf]
    double* dpart;            // [nbf][nq]   per-row-block partial d
    double* d;                // [nq]
    double* J;                // [nbf][nbf]
};

constexpr int J1_ROWS = 64;      // q rows per CTA in K1
constexpr int J_THREADS = 256;

// K1.  grid = (ceil(nq/64), nbf).  smem: gathered density row Dp (sp(m) doubles, zero below the
// symmetric start column).  Each warp owns 8 rows, two at a time; lanes stride the row in double2.
__global__ void __launch_bounds__(J_THREADS) j_dq_kernel(JParams p) {
    extern __shared__ __align__(16) double Dp[];
    const int m = blockIdx.y;
    const int sp = p.sp[m];
    const int ldm = p.ldm[m];
    const int kstart = p.symmetric ? p.ign[m] : 0;
    const int* cols = p.cols + p.cols_off[m];
    const int spe = (sp + 1) & ~1;  // even length (pad column reads hit zero weight)
    for (int k = threadIdx.x; k < spe; k += blockDim.x) {
        double v = 0.0;
        if (k >= kstart && k < sp) {
            int n = cols[k];
            v = p.D[(size_t)m * p.nbf + n];
            if (p.symmetric && n != m) v *= 2.0;
        }
        Dp[k] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* Bm = p.tensor + p.row_off[m];
    const int k0 = kstart & ~1;
    const int qbase = blockIdx.x * J1_ROWS + warp * 8;
#pragma unroll 1
    for (int r = 0; r < 8; r += 2) {
        int qa = qbase + r, qb = qa + 1;
        if (qa >= p.nq) break;
        bool hb = qb < p.nq;
        const double* ra = Bm + (size_t)qa * ldm;
        const double* rb = Bm + (size_t)(hb ? qb : qa) * ldm;
        double sa0 = 0, sa1 = 0, sb0 = 0, sb1 = 0;
#pragma unroll 4
        for (int k = k0 + lane * 2; k < spe; k += 64) {
            double2 w = *reinterpret_cast<const double2*>(Dp + k);
            double2 a = __ldcs(reinterpret_cast<const double2*>(ra + k));
            double2 b = __ldcs(reinterpret_cast<const double2*>(rb + k));
            sa0 = fma(a.x, w.x, sa0);
            sa1 = fma(a.y, w.y, sa1);
            sb0 = fma(b.x, w.x, sb0);
            sb1 = fma(b.y, w.y, sb1);
        }
        double sa = sa0 + sa1, sb = sb0 + sb1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, off);
            sb += __shfl_xor_sync(0xffffffffu, sb, off);
        }
        if (lane == 0) {
            p.dpart[(size_t)m * p.nq + qa] = sa;
            if (hb) p.dpart[(size_t)m * p.nq + qb] = sb;
        }
    }
}

// d[q] = sum_m dpart[m][q] in a fixed order (deterministic; replaces the serial thread reduce :3197-3199).
// block = 32 q x 32 m-slices: slice y sums m = y, y+32, ... into shared memory, then the 32 slice sums of a q are
// added in slice order.  (One thread per q walking all nbf rows left a handful of CTAs on a latency-bound loop: at
// 592 Q rows per GPU that was ~0.2 ms of a 1.2 ms J phase.)
constexpr int JR_Q = 32, JR_M = 32;
__global__ void __launch_bounds__(JR_Q * JR_M) j_dq_reduce_kernel(const double* __restrict__ dpart, int nbf, int nq,
                                                                   double* __restrict__ d) {
    __shared__ double part[JR_M][JR_Q + 1];
    const int qx = threadIdx.x & (JR_Q - 1), my = threadIdx.x / JR_Q;
    const int q = blockIdx.x * JR_Q + qx;
    double s0 = 0, s1 = 0;
    if (q < nq) {
        int m = my;
        for (; m + JR_M < nbf; m += 2 * JR_M) {
            s0 += dpart[(size_t)m * nq + q];
            s1 += dpart[(size_t)(m + JR_M) * nq + q];
        }
        if (m < nbf) s0 += dpart[(size_t)m * nq + q];
    }
    part[my][qx] = s0 + s1;
    __syncthreads();
    if (my == 0 && q < nq) {
        double s = 0.0;
#pragma unroll
        for (int y = 0; y < JR_M; y++) s += part[y][qx];
        d[q] = s;
    }
}

// Density operand of the fused first sweep (half_ws_kernel): Dm[m][n] = D[m][n], or, for lr_symmetric builds,
// D[min(m,n)][max(m,n)] -- the symmetric path reads only the upper triangle of D (dfhelper.cc:3185-3190: D_mm and
// 2 D_mn for n > m over the n >= m half of a tensor that is symmetric in mn  ==  the full sum with this Dm).
__global__ void j_prep_dm_kernel(const double* __restrict__ D, int nbf, int ldd, int symmetric, double* __restrict__ Dm) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int m = blockIdx.y;
    if (n >= ldd) return;
    double v = 0.0;
    if (n < nbf) v = (symmetric && n < m) ? D[(size_t)n * nbf + m] : D[(size_t)m * nbf + n];
    Dm[(size_t)m * ldd + n] = v;
}

// K2.  grid = (ceil(max_sp/128), nbf).  CTA: row-block m, packed columns [kt*128, kt*128+128).
// Warp w sums rows q = w, w+8, ...; lane owns columns 2*lane,+1 and 64+2*lane,+1.  Cross-warp
// reduction in shared memory in fixed order, then the sparse->dense unpack writes J directly.
// ND densities ride on ONE read of the tensor (UHF passes two, response builds several: the reference re-reads B
// per density, dfhelper.cc:3177 loops over i): d vectors side by side in shared memory, ND accumulator sets per lane;
// per density the arithmetic is exactly the ND = 1 kernel's.
constexpr int J_MAX_ND = 4;
struct JBatch {
    const double* d[J_MAX_ND];  // [nq] each
    double* J[J_MAX_ND];        // [nbf][nbf] each
};
template <int ND>
__global__ void __launch_bounds__(J_THREADS) j_mn_kernel(JParams p, JBatch bt) {
    __shared__ double red[8][128];
    extern __shared__ __align__(16) double dq[];  // d[dens][q] staged once per CTA
    const int m = blockIdx.y;
    const int sp = p.sp[m];
    const int kt0 = blockIdx.x * 128;
    if (kt0 >= sp) return;
    const int kstart = p.symmetric ? p.ign[m] : 0;
    if (kt0 + 128 <= kstart) return;
#pragma unroll
    for (int dn = 0; dn < ND; dn++)
        for (int q = threadIdx.x; q < p.nq; q += blockDim.x) dq[dn * p.nq + q] = bt.d[dn][q];
    __syncthreads();
    const int ldm = p.ldm[m];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* Bm = p.tensor + p.row_off[m];
    const int ka = kt0 + lane * 2, kb = ka + 64;
    const bool va = ka < ldm, vb = kb < ldm;  // ldm is a multiple of 4: ka+1 < ldm when ka < ldm
    double a0[ND], a1[ND], b0[ND], b1[ND];
#pragma unroll
    for (int dn = 0; dn < ND; dn++) a0[dn] = a1[dn] = b0[dn] = b1[dn] = 0.0;
    const double* base = Bm + (size_t)warp * ldm;
    const size_t step = (size_t)8 * ldm;
    int q = warp;
#pragma unroll 4
    for (; q < p.nq; q += 8, base += step) {
        double w[ND];
#pragma unroll
        for (int dn = 0; dn < ND; dn++) w[dn] = dq[dn * p.nq + q];
        if (va) {
            double2 x = __ldcs(reinterpret_cast<const double2*>(base + ka));
#pragma unroll
            for (int dn = 0; dn < ND; dn++) {
                a0[dn] = fma(x.x, w[dn], a0[dn]);
                a1[dn] = fma(x.y, w[dn], a1[dn]);
            }
        }
        if (vb) {
            double2 y = __ldcs(reinterpret_cast<const double2*>(base + kb));
#pragma unroll
            for (int dn = 0; dn < ND; dn++) {
                b0[dn] = fma(y.x, w[dn], b0[dn]);
                b1[dn] = fma(y.y, w[dn], b1[dn]);
            }
        }
    }
#pragma unroll
    for (int dn = 0; dn < ND; dn++) {
        if (dn) __syncthreads();
        red[warp][lane * 2] = a0[dn];
        red[warp][lane * 2 + 1] = a1[dn];
        red[warp][64 + lane * 2] = b0[dn];
        red[warp][64 + lane * 2 + 1] = b1[dn];
        __syncthreads();
        if (threadIdx.x < 128) {
            int k = kt0 + threadIdx.x;
            if (k < sp && k >= kstart) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < 8; w++) s += red[w][threadIdx.x];
                int n = p.cols[p.cols_off[m] + k];
                double* J = bt.J[dn];
                J[(size_t)m * p.nbf + n] += s;
                if (p.symmetric && n != m) J[(size_t)n * p.nbf + m] += s;
            }
        }
    }
}

}  // namespace b2k
