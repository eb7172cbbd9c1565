// gemm_strided.cuh -- a plain FP64 tensor-core GEMM with arbitrary operand strides, for the contractions AROUND the
// hot path that are neither of its two pipelined shapes (the DF-JK gradient intermediates of scfgrad/jk_grad.cc:
// (A|mi) -> (A|ij), the metric contractions, the AO back-transform).
//
//   C[b][m*sCm + n*sCn] = alpha * sum_k A[b][m*sAm + k*sAk] * B[b][n*sBn + k*sBk]  +  beta * C[...]
//
// CTA tile 64 x 64 x 16, four warps (2 x 2), warp tile 32 x 32 as 4 x 4 DMMA m8n8k4 blocks; operands are staged
// through registers into padded shared tiles (the next k slab is in flight while the current one is multiplied), so
// any layout works -- coalescing is whatever the strides allow.  This is NOT the speed-of-light kernel of the build
// (that is dmma_ws.cuh); it runs the once-per-geometry gradient contractions at several TFLOP/s.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmma_gemm.cuh"

namespace b2k {

struct GemmStrided {
    int M, N, K, batch;
    const double* A;
    long long sAm, sAk, bA;
    const double* B;
    long long sBn, sBk, bB;
    double* C;
    long long sCm, sCn, bC;
    double alpha, beta;
};

constexpr int GS_T = 64, GS_K = 16, GS_LD = GS_K + 1, GS_THREADS = 128;

__global__ void __launch_bounds__(GS_THREADS) gemm_strided_kernel(GemmStrided d) {
    __shared__ double As[GS_T][GS_LD], Bs[GS_T][GS_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1, gq = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * GS_T, n0 = blockIdx.x * GS_T;
    const double* A = d.A + (long long)blockIdx.z * d.bA;
    const double* B = d.B + (long long)blockIdx.z * d.bB;
    double* C = d.C + (long long)blockIdx.z * d.bC;
    // element e of a 64 x 16 slab -> (row, kk): the index with the smaller stride runs fastest across the threads
    const bool a_kfast = (d.sAk <= d.sAm), b_kfast = (d.sBk <= d.sBn);
    double ra[8], rb[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = tid + i * GS_THREADS;
            const int ar = a_kfast ? (e >> 4) : (e & 63), ak = a_kfast ? (e & 15) : (e >> 6);
            const int br = b_kfast ? (e >> 4) : (e & 63), bk = b_kfast ? (e & 15) : (e >> 6);
            ra[i] = (m0 + ar < d.M && k0 + ak < d.K) ? A[(long long)(m0 + ar) * d.sAm + (long long)(k0 + ak) * d.sAk] : 0.0;
            rb[i] = (n0 + br < d.N && k0 + bk < d.K) ? B[(long long)(n0 + br) * d.sBn + (long long)(k0 + bk) * d.sBk] : 0.0;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = tid + i * GS_THREADS;
            const int ar = a_kfast ? (e >> 4) : (e & 63), ak = a_kfast ? (e & 15) : (e >> 6);
            const int br = b_kfast ? (e >> 4) : (e & 63), bk = b_kfast ? (e & 15) : (e >> 6);
            As[ar][ak] = ra[i];
            Bs[br][bk] = rb[i];
        }
    };
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    fetch(0);
    for (int k0 = 0; k0 < d.K; k0 += GS_K) {
        __syncthreads();  // everyone has finished reading the previous slab
        stash();
        __syncthreads();
        if (k0 + GS_K < d.K) fetch(k0 + GS_K);
#pragma unroll
        for (int ks = 0; ks < GS_K / 4; ks++) {
            double a[4], b[4];
#pragma unroll
            for (int mb = 0; mb < 4; mb++) a[mb] = As[wm * 32 + mb * 8 + gq][ks * 4 + t];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) b[nb] = Bs[wn * 32 + nb * 8 + gq][ks * 4 + t];
#pragma unroll
            for (int mb = 0; mb < 4; mb++)
#pragma unroll
                for (int nb = 0; nb < 4; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
        }
    }
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
        const int m = m0 + wm * 32 + mb * 8 + gq;
        if (m >= d.M) continue;
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int n = n0 + wn * 32 + nb * 8 + t * 2 + e;
                if (n < d.N) {
                    double* c = C + (long long)m * d.sCm + (long long)n * d.sCn;
                    const double v = d.alpha * acc[mb][nb][e];
                    *c = d.beta == 0.0 ? v : v + d.beta * *c;
                }
            }
        }
    }
}

}  // namespace b2k
