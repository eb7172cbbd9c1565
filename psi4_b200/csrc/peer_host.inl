// peer_host.inl -- host side of the fixed-order cross-GPU sum (peer_reduce.cuh); included inside engine.cu's
// anonymous namespace.  Two ways to get every shard's window mapped everywhere:
//   one process, N GPUs (b200jk_create):     cudaDeviceEnablePeerAccess, the pointers are already valid on every device
//   one process per GPU (b200jk_create_rank): CUDA IPC handles of the flag block (once) and of the partial-result
//                                             buffer (whenever it is reallocated), exchanged with one ncclAllGather
// If some pair of devices has no peer access the handle keeps using NCCL for the sum (stats.reduce_kind says which).

typedef CUresult (*AddrRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
AddrRangeFn g_addr_range = nullptr;

bool reduce_forced_nccl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200JK_REDUCE");
        v = (e && !strcmp(e, "nccl")) ? 1 : 0;
    }
    return v == 1;
}

unsigned long long peer_timeout_ns() {
    static unsigned long long v = 0;
    if (!v) {
        const char* e = getenv("B200JK_PEER_TIMEOUT_MS");
        unsigned long long ms = e && *e ? strtoull(e, nullptr, 10) : 20000ull;
        v = std::max<unsigned long long>(ms, 1) * 1000000ull;
    }
    return v;
}

constexpr size_t kPeerBlockBytes = (size_t)2 << 20;  // allocations this large own their IPC-exportable block
constexpr size_t kPeerMsgBytes = 128;

struct PeerMsg {
    cudaIpcMemHandle_t handle;  // of the allocation that contains the pointer
    uint64_t offset;            // pointer - allocation base
    uint64_t ok;
};
static_assert(sizeof(PeerMsg) <= kPeerMsgBytes, "peer message size");

int peer_alloc_flags(b200jk* h, Shard& s) {
    CK(cudaSetDevice(s.dev));
    if (!s.peer.flags) {
        CK(cudaMalloc((void**)&s.peer.flags, kPeerBlockBytes));
        CK(cudaMemset(s.peer.flags, 0, kPeerBlockBytes));
    }
    if (!s.peer.status) {
        CK(cudaHostAlloc((void**)&s.peer.status, sizeof(unsigned), cudaHostAllocPortable | cudaHostAllocMapped));
        *s.peer.status = 0;
    }
    return 0;
}

// one process drives every GPU: peer access all-to-all; windows are looked up at launch time
int peer_setup_local(b200jk* h) {
    const int n = (int)h->sh.size();
    h->reduce_kind = 2;
    if (n < 2) {
        h->reduce_kind = 0;
        return 0;
    }
    if (reduce_forced_nccl() || n > PR_MAXP) return 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, h->sh[i].dev, h->sh[j].dev));
            if (!can) return 0;  // stay on NCCL
        }
    for (int i = 0; i < n; i++) {
        CK(cudaSetDevice(h->sh[i].dev));
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(h->sh[j].dev, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
            } else if (e != cudaSuccess) {
                cudaGetLastError();
                return 0;
            }
        }
        int rc = peer_alloc_flags(h, h->sh[i]);
        if (rc) return rc;
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) h->sh[i].peer.peer_flags[j] = h->sh[j].peer.flags;
    h->reduce_kind = 1;
    return 0;
}

// rank mode: all ranks agree (min over ranks) on a 0/1 flag
int peer_agree(b200jk* h, Shard& s, int mine, int* all) {
    int* d = (int*)s.peer.xchg;
    CK(cudaMemcpyAsync(d, &mine, sizeof(int), cudaMemcpyHostToDevice, s.stream));
    NK(g_nccl.AllReduce(d, d, 1, kNcclInt32, kNcclMin, s.comm, s.stream));
    CK(cudaMemcpyAsync(all, d, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return 0;
}

// rank mode: publish `ptr` (device memory of this rank) to every peer and map theirs.  mapped[r] / bases[r] are
// filled for r != rank; *ok = every rank mapped every peer.
int peer_exchange(b200jk* h, Shard& s, void* ptr, void** bases, void** mapped, bool* ok) {
    const int W = h->world, me = h->rank;
    if (!g_addr_range) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return fail(h, B200JK_ERR_CUDA, "cuMemGetAddressRange not available from the driver");
        g_addr_range = (AddrRangeFn)fn;
    }
    CK(cudaSetDevice(s.dev));
    PeerMsg msg;
    memset(&msg, 0, sizeof msg);
    CUdeviceptr base = 0;
    size_t size = 0;
    int good = 1;
    if (g_addr_range(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) good = 0;
    if (good && cudaIpcGetMemHandle(&msg.handle, (void*)base) != cudaSuccess) {
        cudaGetLastError();
        good = 0;
    }
    msg.offset = good ? (uint64_t)((CUdeviceptr)ptr - base) : 0;
    msg.ok = (uint64_t)good;
    std::vector<char> all((size_t)W * kPeerMsgBytes, 0);
    char* d = (char*)s.peer.xchg;
    CK(cudaMemcpyAsync(d + (size_t)me * kPeerMsgBytes, &msg, sizeof msg, cudaMemcpyHostToDevice, s.stream));
    NK(g_nccl.AllGather(d + (size_t)me * kPeerMsgBytes, d, kPeerMsgBytes, kNcclInt8, s.comm, s.stream));
    CK(cudaMemcpyAsync(all.data(), d, all.size(), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (int r = 0; r < W; r++) {
        bases[r] = mapped[r] = nullptr;
        if (r == me) continue;
        PeerMsg pm;
        memcpy(&pm, all.data() + (size_t)r * kPeerMsgBytes, sizeof pm);
        if (!pm.ok) {
            good = 0;
            continue;
        }
        void* b = nullptr;
        if (cudaIpcOpenMemHandle(&b, pm.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
            continue;
        }
        bases[r] = b;
        mapped[r] = (char*)b + pm.offset;
    }
    int all_good = 0;
    int rc = peer_agree(h, s, good, &all_good);
    if (rc) return rc;
    *ok = all_good != 0;
    if (!*ok)
        for (int r = 0; r < W; r++)
            if (bases[r]) {
                cudaIpcCloseMemHandle(bases[r]);
                bases[r] = mapped[r] = nullptr;
            }
    return 0;
}

// rank mode, once per handle (collective: every rank is inside b200jk_create_rank)
int peer_setup_rank(b200jk* h) {
    h->reduce_kind = h->world > 1 ? 2 : 0;
    if (h->world < 2) return 0;
    Shard& s = h->sh[0];
    CK(cudaSetDevice(s.dev));
    CK(cudaMalloc(&s.peer.xchg, PR_MAXP * kPeerMsgBytes));
    // every rank takes the same branch: the switch is an environment variable of the job, the world size is shared
    if (reduce_forced_nccl() || h->world > PR_MAXP) return 0;
    int rc = peer_alloc_flags(h, s);
    if (rc) return rc;
    void* mapped[PR_MAXP];
    bool ok = false;
    if ((rc = peer_exchange(h, s, s.peer.flags, s.peer.ipc_flags_base, mapped, &ok))) return rc;
    if (!ok) return 0;
    for (int r = 0; r < h->world; r++) s.peer.peer_flags[r] = r == h->rank ? s.peer.flags : (unsigned*)mapped[r];
    s.peer.flags_open = true;
    h->reduce_kind = 1;
    return 0;
}

// rank mode: nobody maps anybody's window any more (collective; called before a window is reallocated or freed)
int peer_close_windows(b200jk* h) {
    if (!h->rank_mode || h->world < 2 || h->reduce_kind != 1) return 0;
    Shard& s = h->sh[0];
    if (!s.peer.win_open) return 0;
    CK(cudaSetDevice(s.dev));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaStreamSynchronize(s.copy));
    for (int r = 0; r < h->world; r++)
        if (s.peer.ipc_win_base[r]) {
            cudaIpcCloseMemHandle(s.peer.ipc_win_base[r]);
            s.peer.ipc_win_base[r] = nullptr;
        }
    s.peer.win_open = false;
    s.peer.win_key = nullptr;
    int all = 0;
    return peer_agree(h, s, 1, &all);  // barrier: every rank has unmapped before anyone frees
}

// rank mode: (re)publish the window after `out` moved
int peer_open_windows(b200jk* h) {
    Shard& s = h->sh[0];
    if (s.peer.win_open && s.peer.win_key == s.out) return 0;
    int rc = peer_close_windows(h);
    if (rc) return rc;
    void* mapped[PR_MAXP];
    bool ok = false;
    if ((rc = peer_exchange(h, s, s.out, s.peer.ipc_win_base, mapped, &ok))) return rc;
    if (!ok) return fail(h, B200JK_ERR_CUDA, "peer mapping of the partial-result window failed on some rank");
    for (int r = 0; r < h->world; r++) s.peer.peer_win[r] = r == h->rank ? s.out : (double*)mapped[r];
    s.peer.win_open = true;
    s.peer.win_key = s.out;
    return 0;
}

void peer_free(b200jk* h, Shard& s) {
    cudaSetDevice(s.dev);
    for (int r = 0; r < PR_MAXP; r++) {
        if (s.peer.ipc_win_base[r]) cudaIpcCloseMemHandle(s.peer.ipc_win_base[r]);
        if (s.peer.ipc_flags_base[r]) cudaIpcCloseMemHandle(s.peer.ipc_flags_base[r]);
        s.peer.ipc_win_base[r] = s.peer.ipc_flags_base[r] = nullptr;
    }
    if (h->rank_mode && h->world > 1 && h->reduce_kind == 1 && s.comm && s.peer.xchg) {
        // every rank has unmapped before anyone frees what it exported (destroy is collective, like ncclCommDestroy)
        int all = 0;
        peer_agree(h, s, 1, &all);
    }
    if (s.peer.flags) cudaFree(s.peer.flags);
    if (s.peer.xchg) cudaFree(s.peer.xchg);
    if (s.peer.status) cudaFreeHost(s.peer.status);
    s.peer.status = nullptr;
    s.peer.flags = nullptr;
    s.peer.xchg = nullptr;
}

// Sum `count` doubles at s.out + off over all Q shards, result in every shard's buffer (root < 0) or in shard
// `root`'s only.  use_copy: issue on the copy stream (K results are reduced and sent home while the J sweeps still
// run on the compute stream); channel: which flag set (two sums may be in flight at once, one per stream).
int reduce_shards(b200jk* h, size_t off, size_t count, bool use_copy, int channel, int root) {
    bool multi = h->sh.size() > 1 || (h->rank_mode && h->world > 1);
    if (!multi || !count) return 0;
    if (h->reduce_kind == 1) {
        const int W = h->rank_mode ? h->world : (int)h->sh.size();
        if (h->rank_mode) {
            int rc = peer_open_windows(h);
            if (rc) return rc;
        }
        for (size_t si = 0; si < h->sh.size(); si++) {
            Shard& s = h->sh[si];
            CK(cudaSetDevice(s.dev));
            cudaStream_t st = use_copy ? s.copy : s.stream;
            PeerReduceParams p;
            memset(&p, 0, sizeof p);
            for (int r = 0; r < W; r++) {
                p.win[r] = h->rank_mode ? s.peer.peer_win[r] : h->sh[r].out;
                p.flags[r] = s.peer.peer_flags[r] + channel * PR_FLAG_WORDS;
            }
            p.rank = h->rank_mode ? h->rank : (int)si;
            p.world = W;
            p.off = off;
            p.count = count;
            p.epoch = ++s.peer.epoch[channel];
            p.root = root;
            p.host_dst = nullptr;
            p.host_all = 0;
            p.timeout_ns = peer_timeout_ns();
            p.status = s.peer.status;
            const size_t per = (count + W - 1) / W;
            int grid = (int)std::min<size_t>(std::max<size_t>(1, (per / 2 + PR_THREADS - 1) / PR_THREADS), (size_t)s.nsm);
            Phase ph;
            ph.tag = 3;
            ph.a = get_event(s);
            ph.b = get_event(s);
            CK(cudaEventRecord(ph.a, st));
            switch (W) {
                case 2: peer_reduce_kernel<2><<<grid, PR_THREADS, 0, st>>>(p); break;
                case 4: peer_reduce_kernel<4><<<grid, PR_THREADS, 0, st>>>(p); break;
                case 8: peer_reduce_kernel<8><<<grid, PR_THREADS, 0, st>>>(p); break;
                default: peer_reduce_kernel<0><<<grid, PR_THREADS, 0, st>>>(p); break;
            }
            s.launches++;
            CK(cudaGetLastError());
            CK(cudaEventRecord(ph.b, st));
            s.phases.push_back(ph);
        }
        return 0;
    }
    if (h->sh.size() > 1) NK(g_nccl.GroupStart());
    for (auto& s : h->sh) {
        CK(cudaSetDevice(s.dev));
        cudaStream_t st = use_copy ? s.copy : s.stream;
        Phase ph;
        ph.tag = 3;
        ph.a = get_event(s);
        ph.b = get_event(s);
        CK(cudaEventRecord(ph.a, st));
        NK(g_nccl.AllReduce(s.out + off, s.out + off, count, kNcclDouble, kNcclSum, s.comm, st));
        CK(cudaEventRecord(ph.b, st));
        s.phases.push_back(ph);
    }
    if (h->sh.size() > 1) NK(g_nccl.GroupEnd());
    return 0;
}

// after the build's final synchronisation: did any barrier of the peer sum give up waiting?  (the status word lives
// in page-locked host memory, so looking at it costs nothing)
int peer_check_status(b200jk* h) {
    if (h->reduce_kind != 1) return 0;
    for (auto& s : h->sh) {
        if (!s.peer.status || !*(volatile unsigned*)s.peer.status) continue;
        const unsigned st = *(volatile unsigned*)s.peer.status;
        *s.peer.status = 0;
        return fail(h, B200JK_ERR_CUDA,
                    "cross-GPU sum timed out at barrier %c waiting for rank %u: a peer never launched its half of the build",
                    st >= 0x100u ? 'B' : 'A', (st & 0xffu) - 1u);
    }
    return 0;
}
