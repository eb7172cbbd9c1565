// Standalone check and timing of the INT8-tensor-core K GEMM (psi4_b200/csrc/i8_kgemm.cuh) against a double-double
// reference of the same product, and against the FP64 figure of the DMMA kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o k4_i8_test k4_i8_test.cu
//   ./k4_i8_test nbf kdim nmod klen symmetric [nsample] [reps]
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../psi4_b200/csrc/i8_kgemm.cuh"

using namespace b2k;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// rows with different magnitudes, elements with ~12 binades of dynamic range inside a row (like a half-transformed tensor:
// a few large entries per row, most small)
__global__ void fill_kernel(double* T, size_t pitch, int nbf, int kdim, uint64_t seed) {
    const int row = blockIdx.y;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < kdim; k += gridDim.x * blockDim.x) {
        const uint64_t h = mix64(seed + (uint64_t)row * 0x100000001B3ull + (uint64_t)k);
        const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        const int sh = (int)((h >> 3) & 15);
        const double rs = ldexp(1.0, (row % 9) - 4);
        T[(size_t)row * pitch + k] = u * ldexp(1.0, -sh) * rs * 0.05;
    }
}

// double-double dot product of rows m and n
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
__global__ void ref_kernel(const double* T1, const double* T2, size_t pitch, int kdim, const int2* pairs, int npairs, double* out,
                           double* norms) {
    const int pidx = blockIdx.x;
    if (pidx >= npairs) return;
    const double* a = T1 + (size_t)pairs[pidx].x * pitch;
    const double* b = T2 + (size_t)pairs[pidx].y * pitch;
    double hi = 0, lo = 0, na = 0, nb = 0;
    for (int k = threadIdx.x; k < kdim; k += blockDim.x) {
        const double p = a[k] * b[k];
        const double pe = fma(a[k], b[k], -p);
        double s, e;
        two_sum(hi, p, s, e);
        hi = s;
        lo += e + pe;
        na = fma(a[k], a[k], na);
        nb = fma(b[k], b[k], nb);
    }
    __shared__ double sh[256], sl[256], sa[256], sb[256];
    sh[threadIdx.x] = hi;
    sl[threadIdx.x] = lo;
    sa[threadIdx.x] = na;
    sb[threadIdx.x] = nb;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            double s, e;
            two_sum(sh[threadIdx.x], sh[threadIdx.x + w], s, e);
            sh[threadIdx.x] = s;
            sl[threadIdx.x] += sl[threadIdx.x + w] + e;
            sa[threadIdx.x] += sa[threadIdx.x + w];
            sb[threadIdx.x] += sb[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[pidx] = sh[0] + sl[0];
        norms[pidx] = sqrt(sa[0] * sb[0]);
    }
}

int main(int argc, char** argv) {
    const int nbf = argc > 1 ? atoi(argv[1]) : 300;
    const int kdim = argc > 2 ? atoi(argv[2]) : 5000;
    const int nmod = argc > 3 ? atoi(argv[3]) : 13;
    const int klen = argc > 4 ? atoi(argv[4]) : 8192;
    const int sym = argc > 5 ? atoi(argv[5]) : 1;
    int nsample = argc > 6 ? atoi(argv[6]) : 4096;
    const int reps = argc > 7 ? atoi(argv[7]) : 3;
    const size_t pitch = (size_t)kdim;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    I8EncodeFn enc = (I8EncodeFn)fn;

    double *T1, *T2, *K, *K2;
    CK(cudaMalloc(&T1, (size_t)nbf * pitch * 8));
    fill_kernel<<<dim3(64, nbf), 256>>>(T1, pitch, nbf, kdim, 12345);
    if (sym) {
        T2 = T1;
    } else {
        CK(cudaMalloc(&T2, (size_t)nbf * pitch * 8));
        fill_kernel<<<dim3(64, nbf), 256>>>(T2, pitch, nbf, kdim, 99991);
    }
    CK(cudaMalloc(&K, (size_t)nbf * nbf * 8));
    CK(cudaMalloc(&K2, (size_t)nbf * nbf * 8));
    CK(cudaDeviceSynchronize());

    I8Plan pl;
    for (int i = 0; i < 4; i++) CK(cudaEventCreate(&pl.prof[i]));
    std::string err;
    I8RunInfo info;
    const size_t budget = (size_t)48 << 30;
    float best[3] = {1e30f, 1e30f, 1e30f}, best_tot = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaMemsetAsync(K, 0, (size_t)nbf * nbf * 8, 0));
        int rc = i8_kgemm_run(pl, enc, 0, nsm, T1, T2, pitch, nbf, kdim, sym != 0, K, nbf, nmod, klen, budget, &info, &err);
        if (rc) {
            printf("i8_kgemm_run failed rc=%d: %s\n", rc, err.c_str());
            return 1;
        }
        CK(cudaDeviceSynchronize());
        float t[3], tot;
        for (int i = 0; i < 3; i++) {
            CK(cudaEventElapsedTime(&t[i], pl.prof[i], pl.prof[i + 1]));
            if (t[i] < best[i]) best[i] = t[i];
        }
        CK(cudaEventElapsedTime(&tot, pl.prof[0], pl.prof[3]));
        if (tot < best_tot) best_tot = tot;
    }
    const double credited = sym ? (double)nbf * (nbf + 1) * kdim : 2.0 * nbf * nbf * kdim;
    printf("i8 K GEMM nbf=%d kdim=%d nmod=%d klen=%d sym=%d: ntile=%d nsplit=%d passes=%d bits=%.2f\n", nbf, kdim, info.nmod, info.klen,
           sym, info.ntile, info.nsplit, info.passes, info.bits);
    printf("  ms: planes %.3f  gemm %.3f  crt %.3f  total %.3f (pass 0)  -> %.1f TFLOP/s FP64-equivalent (credited %.3e flop)\n", best[0],
           best[1], best[2], best_tot, credited / best_tot * 1e-9, credited);
    {
        double area = 0;  // executed int8 MACs
        std::vector<I8Tile> tiles(info.ntile);
        CK(cudaMemcpy(tiles.data(), pl.d_tiles, tiles.size() * sizeof(I8Tile), cudaMemcpyDeviceToHost));
        for (auto& t : tiles) area += 128.0 * t.ncols;
        printf("  int8 tensor rate %.2f Pop/s (executed %.3e MAC x %d moduli)\n", 2.0 * area * kdim * info.nmod / best[1] * 1e-12,
               area * kdim, info.nmod);
    }

    // run-to-run and split-independence: the result must not depend on klen (integer arithmetic)
    CK(cudaMemsetAsync(K2, 0, (size_t)nbf * nbf * 8, 0));
    int rc = i8_kgemm_run(pl, enc, 0, nsm, T1, T2, pitch, nbf, kdim, sym != 0, K2, nbf, nmod, klen == 8192 ? 4096 : 8192, budget, nullptr,
                          &err);
    if (rc) {
        printf("second run failed: %s\n", err.c_str());
        return 1;
    }
    CK(cudaDeviceSynchronize());
    std::vector<double> hK((size_t)nbf * nbf), hK2((size_t)nbf * nbf);
    CK(cudaMemcpy(hK.data(), K, hK.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hK2.data(), K2, hK2.size() * 8, cudaMemcpyDeviceToHost));
    size_t ndiff = 0;
    for (size_t i = 0; i < hK.size(); i++) ndiff += (memcmp(&hK[i], &hK2[i], 8) != 0);
    printf("  other klen: %zu of %zu elements differ bitwise\n", ndiff, hK.size());
    if (sym) {
        size_t nasym = 0;
        for (int m = 0; m < nbf; m++)
            for (int n = 0; n < m; n++) nasym += (hK[(size_t)m * nbf + n] != hK[(size_t)n * nbf + m]);
        printf("  asymmetric pairs: %zu\n", nasym);
    }

    // reference on sampled pairs (all pairs when the matrix is small)
    std::vector<int2> pairs;
    if ((size_t)nbf * nbf <= (size_t)nsample) {
        for (int m = 0; m < nbf; m++)
            for (int n = 0; n < nbf; n++) pairs.push_back(make_int2(m, n));
    } else {
        uint64_t s = 777;
        for (int i = 0; i < nsample; i++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            int m = (int)((s >> 33) % nbf);
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            int n = (int)((s >> 33) % nbf);
            if (i < 64) n = m;                                  // diagonal
            else if (i < 128) { m = nbf - 1 - (i & 7); }        // last rows
            else if (i < 192) { n = nbf - 1 - (i & 7); }
            pairs.push_back(make_int2(m, n));
        }
    }
    int2* dp;
    double *dref, *dnorm;
    CK(cudaMalloc(&dp, pairs.size() * sizeof(int2)));
    CK(cudaMalloc(&dref, pairs.size() * 8));
    CK(cudaMalloc(&dnorm, pairs.size() * 8));
    CK(cudaMemcpy(dp, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice));
    ref_kernel<<<(unsigned)pairs.size(), 256>>>(T1, T2, pitch, kdim, dp, (int)pairs.size(), dref, dnorm);
    CK(cudaDeviceSynchronize());
    std::vector<double> ref(pairs.size()), nrm(pairs.size());
    CK(cudaMemcpy(ref.data(), dref, ref.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(nrm.data(), dnorm, nrm.size() * 8, cudaMemcpyDeviceToHost));
    double max_abs = 0, max_rel = 0, max_k = 0;
    int worst = 0;
    for (size_t i = 0; i < pairs.size(); i++) {
        const double got = hK[(size_t)pairs[i].x * nbf + pairs[i].y];
        const double d = fabs(got - ref[i]);
        if (d > max_abs) {
            max_abs = d;
            worst = (int)i;
        }
        if (nrm[i] > 0 && d / nrm[i] > max_rel) max_rel = d / nrm[i];
        if (fabs(ref[i]) > max_k) max_k = fabs(ref[i]);
    }
    printf("  vs double-double on %zu pairs: max|dK| = %.3e (at (%d,%d): got %.17g want %.17g), max|dK|/(|a||b|) = %.3e = 2^%.1f, max|K| = %.3e\n",
           pairs.size(), max_abs, pairs[worst].x, pairs[worst].y, hK[(size_t)pairs[worst].x * nbf + pairs[worst].y], ref[worst], max_rel,
           max_rel > 0 ? log2(max_rel) : -999.0, max_k);
    const bool ok = ndiff == 0 && max_rel < ldexp(1.0, -(int)info.bits + 2) * sqrt((double)kdim);
    printf("%s\n", ok ? "OK" : "FAILED");
    return ok ? 0 : 2;
}
