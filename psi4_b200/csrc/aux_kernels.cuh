// aux_kernels.cuh -- small support kernels: operand staging, synthetic tensor fill, FP64 peak probes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2k {

// Ct[i][n] = C[n][i] (C row-major nbf x o, as psi4 hands it over); rows i in [o, orows) zeroed.
__global__ void transpose_c_kernel(const double* __restrict__ C, int nbf, int o, double* __restrict__ Ct, int ldc,
                                   int orows) {
    __shared__ double tile[32][33];
    int n0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int n = n0 + r, i = i0 + threadIdx.x;
        tile[r][threadIdx.x] = (n < nbf && i < o) ? C[(size_t)n * o + i] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r, n = n0 + threadIdx.x;
        if (i < orows && n < ldc) Ct[(size_t)i * ldc + n] = (n < nbf) ? tile[threadIdx.x][r] : 0.0;
    }
}

// Packed copy of the C^T operand of the half transform, one block per row-block m in the tensor's own column order:
// Cg[m][r][k] = Ct[r][n_k(m)] for r < o, and row r == o holds the density row Dm[m][n_k(m)] of the fused first J sweep
// (zeros when Dm == nullptr).  Same pitch ld(m) as the tensor, so K3 can fetch both operands of a screened row-block
// by TMA (through a per-m map with dims {sp(m), o+1}) instead of gathering C^T through the LSU for every work item.
// grid = (nbf, ceil((o+1)/8)); block 256.
__global__ void __launch_bounds__(256) gather_ct_kernel(const double* __restrict__ Ct, int ldc, int o,
                                                         const double* __restrict__ Dm, int ldd,
                                                         const int* __restrict__ sp, const int* __restrict__ ldm,
                                                         const size_t* __restrict__ row_off_unit,
                                                         const int* __restrict__ cols, const size_t* __restrict__ cols_off,
                                                         double* __restrict__ Cg) {
    const int m = blockIdx.x;
    const int K = sp[m], ld = ldm[m];
    const int R = o + 1;
    const int* c = cols + cols_off[m];
    double* out = Cg + row_off_unit[m] * (size_t)R;
    const int r0 = blockIdx.y * 8;
    for (int k = threadIdx.x; k < ld; k += 256) {
        const int col = k < K ? __ldg(c + k) : -1;
#pragma unroll
        for (int dr = 0; dr < 8; dr++) {
            const int r = r0 + dr;
            if (r >= R) break;
            double v = 0.0;
            if (col >= 0) v = r < o ? Ct[(size_t)r * ldc + col] : (Dm ? Dm[(size_t)m * ldd + col] : 0.0);
            out[(size_t)r * ld + k] = v;
        }
    }
}

// wK post-step of MemDFJK::compute_JK (libfock/MemDFJK.cc:104-110): A <- (A + A^T)/2.
__global__ void hermitivitize_kernel(double* __restrict__ A, int n) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    if (c < n && c < r) {
        double v = 0.5 * (A[(size_t)r * n + c] + A[(size_t)c * n + r]);
        A[(size_t)r * n + c] = v;
        A[(size_t)c * n + r] = v;
    }
}

// ---- synthetic tensor: bit-identical twin of oracle_synth_fill (oracle/dfjk_oracle.c) ----------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ double synth_u(uint64_t seedh, uint64_t Q, uint64_t m, uint64_t n, uint64_t nbf) {
    uint64_t lo = m < n ? m : n, hi = m < n ? n : m;
    uint64_t ctr = (Q * nbf + lo) * nbf + hi;
    uint64_t h = splitmix64(seedh ^ ctr);
    return (double)(int64_t)(h >> 11) * (1.0 / 4503599627370496.0) - 1.0;
}
// grid = (ceil(nq/8), nbf); each warp fills one packed row.
__global__ void synth_fill_kernel(double* __restrict__ tensor, const size_t* __restrict__ row_off,
                                  const int* __restrict__ ldm, const int* __restrict__ sp,
                                  const int* __restrict__ cols, const size_t* __restrict__ cols_off,
                                  const double* __restrict__ amp, int nbf, int nq, int qglobal0, uint64_t seedh) {
    int m = blockIdx.y;
    int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    int lane = threadIdx.x & 31;
    int s = sp[m], ld = ldm[m];
    double* row = tensor + row_off[m] + (size_t)q * ld;
    const int* c = cols + cols_off[m];
    for (int k = lane; k < ld; k += 32) {
        double v = 0.0;
        if (k < s) {
            int n = c[k];
            v = amp[(size_t)m * nbf + n] * synth_u(seedh, (uint64_t)(qglobal0 + q), (uint64_t)m, (uint64_t)n, (uint64_t)nbf);
        }
        row[k] = v;
    }
}

// ---- FP64 ceilings: register-resident loops, no memory traffic ---------------------------------
// kind 0: DMMA m8n8k4 only; 1: DFMA only; 2: both interleaved in every warp (are they one pipe?).
// Each block also reports (clock64 delta, globaltimer delta) so the SM clock under this load is
// measured from inside the kernel rather than sampled by nvidia-smi.
template <int KIND>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, unsigned long long* clk, int iters) {
    double c[16][2];
    double f[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        c[i][0] = c[i][1] = 0.0;
        f[i] = threadIdx.x * 1e-3 + i;
    }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    unsigned long long t0, t1, g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (KIND == 0 || KIND == 2)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1])
                             : "d"(a), "d"(b));
            if (KIND == 1 || KIND == 2) f[i] = fma(f[i], a, b);
        }
    }
    t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1] + f[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0) {
        clk[2 * blockIdx.x] = t1 - t0;
        clk[2 * blockIdx.x + 1] = g1 - g0;
    }
}

}  // namespace b2k
