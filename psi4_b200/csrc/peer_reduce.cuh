// peer_reduce.cuh -- the cross-GPU sum of the partial [J|K|wK] matrices over NVLink peer memory, in a FIXED rank
// order ("deterministic across runs for fixed G").
//
// The reference has no counterpart: its Q-block loop accumulates with beta = 1 in block order
// (lib3index/dfhelper.cc:3124-3159).  With the auxiliary index sharded over GPUs that loop becomes a sum over shards,
// and this kernel keeps its defining property -- one summation order, ((p0 + p1) + p2) + ... by shard index, for every
// element, whatever the message size, stream, or which entry point asked for it.  (An NCCL all-reduce picks its
// algorithm and chunking per message size: the same build reduced as one [J|K] message or as two gave different last
// bits at 4 ranks.  NCCL stays as the bootstrap channel and as the B200JK_REDUCE=nccl A/B arm.)
//
// One launch per rank, all ranks co-resident on their own GPUs:
//   barrier A   every rank tells every peer "my partials are complete" (flag store into the peer's window)
//   reduce      rank r owns elements [count*r/W, count*(r+1)/W): loads them from all W windows in rank order, sums,
//               and stores the result into every window (all-reduce), into the root's window only (reduce), and /
//               or straight into page-locked host memory of this process (the D2H leg rides on the reduction)
//   barrier B   "I have finished reading your window and writing mine"; nobody leaves before every peer has said so,
//               so the kernels that follow in the stream may read the result / overwrite the window.
// Flags are epoch numbers (monotonic per channel), never reset, so a fast rank that is already at the next build's
// barrier A cannot confuse a slow one.  Spins are bounded: after `timeout_ns` the kernel records a status word and
// proceeds, the host turns that into an error instead of a hung GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2k {

constexpr int PR_MAXP = 16;           // ranks per node
constexpr int PR_CHANNELS = 2;        // independent flag sets (K group on the copy stream, J group on the compute stream)
constexpr int PR_FLAG_WORDS = 64;     // per channel: [0,16) barrier A, [16,32) barrier B, 32 CTA counter
constexpr int PR_THREADS = 512;

struct PeerReduceParams {
    double* win[PR_MAXP];         // every rank's window (its partial [J|K|wK] buffer), mapped in this process
    unsigned* flags[PR_MAXP];     // every rank's flag block (this channel)
    int rank, world;
    size_t off, count;            // doubles
    unsigned epoch;
    int root;                     // >= 0: result into that rank's window only; -1: into every window
    double* host_dst;             // optional: page-locked host image of the window (element e -> host_dst[e]); the owner
    int host_all;                 //   of a slice writes it iff host_all (every rank shares the host buffer) or rank == root
    unsigned long long timeout_ns;
    unsigned* status;             // page-locked host word of this rank: non-zero = a barrier gave up waiting
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pr_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// peer data: read once, never cached in this SM's L1 (the peer rewrites it every build)
__device__ __forceinline__ double2 ld_peer2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer1(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// spin until *flag >= epoch (wrap-safe compare); false on timeout
__device__ __forceinline__ bool pr_wait(const unsigned* flag, unsigned epoch, unsigned long long timeout_ns) {
    const unsigned long long t0 = pr_now_ns();
    unsigned it = 0;
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        if ((++it & 1023u) == 0 && pr_now_ns() - t0 > timeout_ns) return false;
    }
    return true;
}

// element e (and e+1 when V == 2) of every window, summed in rank order; all W loads are issued before the first add
template <int W, int V>
__device__ __forceinline__ void pr_sum_store(const PeerReduceParams& p, int world, size_t e, bool to_host) {
    double acc[V];
    if constexpr (W > 0) {
        double v[W][V];
#pragma unroll
        for (int r = 0; r < W; r++) {
            if constexpr (V == 2) {
                const double2 x = ld_peer2(p.win[r] + e);
                v[r][0] = x.x;
                v[r][1] = x.y;
            } else {
                v[r][0] = ld_peer1(p.win[r] + e);
            }
        }
#pragma unroll
        for (int k = 0; k < V; k++) acc[k] = v[0][k];
#pragma unroll
        for (int r = 1; r < W; r++)
#pragma unroll
            for (int k = 0; k < V; k++) acc[k] += v[r][k];
    } else {  // any other world size: one load at a time
#pragma unroll
        for (int k = 0; k < V; k++) acc[k] = ld_peer1(p.win[0] + e + k);
        for (int r = 1; r < world; r++)
#pragma unroll
            for (int k = 0; k < V; k++) acc[k] += ld_peer1(p.win[r] + e + k);
    }
    if (p.root < 0) {
        for (int r = 0; r < world; r++) {
            if constexpr (V == 2)
                *reinterpret_cast<double2*>(p.win[r] + e) = make_double2(acc[0], acc[1]);
            else
                p.win[r][e] = acc[0];
        }
    } else {
        if constexpr (V == 2)
            *reinterpret_cast<double2*>(p.win[p.root] + e) = make_double2(acc[0], acc[1]);
        else
            p.win[p.root][e] = acc[0];
    }
    if (to_host) {
        if constexpr (V == 2)
            *reinterpret_cast<double2*>(p.host_dst + e) = make_double2(acc[0], acc[1]);
        else
            p.host_dst[e] = acc[0];
    }
}

// W = compile-time world size (2, 4, 8: fully unrolled loads) or 0 (run-time loop)
template <int W>
__global__ void __launch_bounds__(PR_THREADS) peer_reduce_kernel(PeerReduceParams p) {
    unsigned* myf = p.flags[p.rank];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    // ---- barrier A ----
    if (tid < p.world) {
        if (blockIdx.x == 0) st_release_sys(p.flags[tid] + p.rank, p.epoch);
        if (!pr_wait(myf + tid, p.epoch, p.timeout_ns)) *p.status = 1u + (unsigned)tid;
    }
    __syncthreads();

    // ---- my slice, summed in rank order ----
    const size_t per = ((p.count + (size_t)p.world * 2 - 1) / ((size_t)p.world * 2)) * 2;  // even slice length
    const size_t a = per * (size_t)p.rank < p.count ? per * (size_t)p.rank : p.count;
    const size_t b = a + per < p.count ? a + per : p.count;
    const bool to_host = p.host_dst && (p.host_all || p.rank == p.root);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (((p.off + a) & 1) == 0) {
        const size_t n2 = (b - a) >> 1;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n2; i += stride)
            pr_sum_store<W, 2>(p, p.world, p.off + a + 2 * i, to_host);
        if (((b - a) & 1) && blockIdx.x == 0 && tid == 0)  // odd tail (last slice of an odd count)
            pr_sum_store<W, 1>(p, p.world, p.off + b - 1, to_host);
    } else {
        for (size_t i = a + (size_t)blockIdx.x * blockDim.x + tid; i < b; i += stride)
            pr_sum_store<W, 1>(p, p.world, p.off + i, to_host);
    }

    // ---- barrier B: the last CTA of this rank speaks for all of them ----
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        s_last = (atomicAdd(myf + 32, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        if (tid == 0) {
            __threadfence_system();
            myf[32] = 0;  // ready for the next launch on this channel
        }
        if (tid < p.world) {
            st_release_sys(p.flags[tid] + PR_MAXP + p.rank, p.epoch);
            if (!pr_wait(myf + PR_MAXP + tid, p.epoch, p.timeout_ns)) *p.status = 0x101u + (unsigned)tid;
        }
    }
}

}  // namespace b2k
