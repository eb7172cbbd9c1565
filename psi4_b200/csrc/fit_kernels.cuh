// fit_kernels.cuh -- support kernels of the on-device fitting step:
//   B = J^-1/2 (A|mn)   DFHelper::contract_metric_AO_core_symm, lib3index/dfhelper.cc:1653-1678
// The host hands over the UNFITTED integrals in the symmetric-packed layout that
// compute_sparse_pQq_blocking_p_symm writes (dfhelper.cc:1284-1347): for each m the block [naux][mi(m)] over the kept
// partners n >= m.  On the device a staged group of row-blocks is transposed to [pair column][naux] (so that the
// contraction index is contiguous for both GEMM operands), multiplied by the metric with fit_gemm_ws_kernel
// (dmma_ws.cuh), scattered into the packed pQq tensor, and mirrored into the n < m columns (:1666-1677).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2k {

struct FitGroup {            // per row-block m of the staged group (device arrays, length nm)
    const size_t* src_off;   // offset (doubles) of block m inside the staged raw buffer
    const int* mi;           // kept partners n >= m
    const int* j0;           // first staged column of block m
    const int* mglob;        // global m
};

// Ut[(j0[b] + k)][R] = U_b[R][k]; rows padded to ldt doubles, pad zeroed.  grid = (ceil(mi_max/32), ceil(naux/32), nm)
__global__ void fit_transpose_kernel(const double* __restrict__ raw, FitGroup g, int naux, int ldt, double* __restrict__ Ut) {
    __shared__ double tile[32][33];
    const int b = blockIdx.z;
    const int mi = g.mi[b];
    const int k0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    if (k0 >= mi) return;
    const double* U = raw + g.src_off[b];
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        int R = r0 + rr, k = k0 + threadIdx.x;
        tile[rr][threadIdx.x] = (R < naux && k < mi) ? U[(size_t)R * mi + k] : 0.0;
    }
    __syncthreads();
    for (int kk = threadIdx.y; kk < 32; kk += blockDim.y) {
        int k = k0 + kk, R = r0 + threadIdx.x;
        if (k < mi && R < ldt) Ut[(size_t)(g.j0[b] + k) * ldt + R] = (R < naux) ? tile[threadIdx.x][kk] : 0.0;
    }
}

// No metric (wPpq_, dfhelper.cc:688-692): plain scatter of this shard's rows.  grid = (ceil(mi_max/128), nq, nm)
__global__ void fit_scatter_kernel(const double* __restrict__ raw, FitGroup g, int q0, const size_t* __restrict__ row_off,
                                   const int* __restrict__ ldm, const int* __restrict__ ign, double* __restrict__ tensor) {
    const int b = blockIdx.z, q = blockIdx.y;
    const int mi = g.mi[b], m = g.mglob[b];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= mi) return;
    tensor[row_off[m] + (size_t)q * ldm[m] + ign[m] + k] = raw[g.src_off[b] + (size_t)(q0 + q) * mi + k];
}

// Mirror: B(q, n, m) = B(q, m, n) for kept n > m (dfhelper.cc:1666-1677).  Reads are coalesced along the row of
// block m; writes land in row-block n at the rank of m among n's partners (mpos).  grid = (ceil(mi_max/128), nq, nm)
__global__ void fit_mirror_kernel(FitGroup g, const size_t* __restrict__ row_off, const int* __restrict__ ldm,
                                  const int* __restrict__ ign, const int* __restrict__ cols,
                                  const size_t* __restrict__ cols_off, const int* __restrict__ mpos,
                                  double* __restrict__ tensor) {
    const int b = blockIdx.z, q = blockIdx.y;
    const int mi = g.mi[b], m = g.mglob[b];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= mi || k == 0) return;  // k == 0 is the diagonal n == m
    const size_t e = cols_off[m] + ign[m] + k;
    const int n = cols[e];
    const double v = tensor[row_off[m] + (size_t)q * ldm[m] + ign[m] + k];
    tensor[row_off[n] + (size_t)q * ldm[n] + mpos[e]] = v;
}

}  // namespace b2k
