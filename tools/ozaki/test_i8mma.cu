// Bring-up test of tcgen05.mma.kind::i8 on sm_100a (standalone: nvcc -arch=sm_100a -o test_i8mma test_i8mma.cu -lcuda).
// C[128 x N] (int32) = A[128 x K] (int8, K contiguous) * B[N x K]^T, one CTA, operands by TMA (SWIZZLE_128B),
// accumulator in TMEM, read back with tcgen05.ld.  Checked against the host.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 bytes, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                         // version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128) i8mma_test_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                                          int K, int32_t* C) {
    extern __shared__ uint8_t raw[];
    uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* As = base;                 // 128 rows x 128 B
    uint8_t* Bs = base + 128 * 128;     // N rows x 128 B
    uint64_t* full = reinterpret_cast<uint64_t*>(Bs + N * 128);
    uint64_t* done = full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(full, 1);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(N < 32 ? 32 : N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const int nkb = K / 128;
    for (int kb = 0; kb < nkb; kb++) {
        if (tid == 0) {
            mbar_expect_tx(full, 128 * 128 + N * 128);
            tma_load_2d(As, &amap, kb * 128, 0, full);
            tma_load_2d(Bs, &bmap, kb * 128, 0, full);
            mbar_wait(full, kb & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint64_t ad = make_desc(smem_u32(As)), bd = make_desc(smem_u32(Bs));
            for (int k = 0; k < 4; k++) mma_i8(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);  // +32 bytes of K per step
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(done)) : "memory");
            mbar_wait(done, kb & 1);  // serial bring-up: smem is reused for the next k-block
        }
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // warp w reads TMEM lanes 32w..32w+31 = rows of C; 32 columns at a time
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                       "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                       "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 32; j++) C[(size_t)tid * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(N < 32 ? 32 : N) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    constexpr int N = 64;
    const int K = 512, M = 128;
    std::vector<int8_t> A((size_t)M * K), B((size_t)N * K);
    srand(7);
    for (auto& x : A) x = (int8_t)(rand() % 129 - 64);
    for (auto& x : B) x = (int8_t)(rand() % 129 - 64);
    int8_t *dA, *dB;
    int32_t* dC;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dC, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    auto mk = [&](CUtensorMap* m, void* p, int rows) {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)K};
        cuuint32_t box[2] = {128, (cuuint32_t)rows};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    CUtensorMap am, bm;
    mk(&am, dA, M);
    mk(&bm, dB, N);
    const int smem = 1024 + 128 * 128 + N * 128 + 64;
    CK(cudaFuncSetAttribute(i8mma_test_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    i8mma_test_kernel<N><<<1, 128, smem>>>(am, bm, K, dC);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> C((size_t)M * N);
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            int32_t ref = 0;
            for (int k = 0; k < K; k++) ref += (int32_t)A[(size_t)m * K + k] * (int32_t)B[(size_t)n * K + k];
            if (ref != C[(size_t)m * N + n]) {
                if (bad < 8) printf("mismatch C[%d][%d] = %d, want %d\n", m, n, C[(size_t)m * N + n], ref);
                bad++;
            }
        }
    printf("i8mma test: %ld mismatches of %d\n", bad, M * N);
    return bad ? 2 : 0;
}
