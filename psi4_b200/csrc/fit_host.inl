// fit_host.inl -- host orchestration of the on-device fitting step (included at the end of engine.cu).
//
// Reference: DFHelper::prepare_AO_core (lib3index/dfhelper.cc:514-588) runs, for each block of basis functions,
//   compute_sparse_pQq_blocking_p_symm  -> unfitted (A|mn), n >= m, symmetric-packed      (:1284-1347, stays on the CPU)
//   contract_metric_AO_core_symm        -> Ppq[m] = metric . (A|mn), then the mirror copy  (:1653-1678, moved here)
// b200jk_set_metric uploads this shard's rows of the metric power; b200jk_fit_rows takes one block of the unfitted
// buffer exactly as psi4 holds it (Mp) and leaves the fitted, mirrored rows in the HBM-resident packed tensor.

namespace {

template <class T>
int fit_grow(b200jk* h, T** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CK(cudaMalloc((void**)p, need * sizeof(T)));
    *cap = need;
    return 0;
}

// One group of row-blocks [ma, mb) on one shard, queued and NOT waited for: the H2D of the raw integrals goes on the
// copy stream into one of two device buffers (so it runs under the kernels of the previous group), transpose / metric
// GEMM / mirror follow on the compute stream.  dma_src is page-locked (the caller's registered block or a slot of the
// staging ring); ev_h2d is recorded when the copy has left it.
int fit_group(b200jk* h, Shard& s, int si, int which, size_t ma, size_t mb, const double* dma_src, bool with_metric,
              cudaEvent_t ev_h2d) {
    const size_t A = h->naux;
    const int apitch = round_up((int)A, 2);
    const size_t nm = mb - ma;
    size_t ncols = 0, raw = 0;
    int mi_max = 0;
    for (size_t m = ma; m < mb; m++) {
        int mi = h->sp[m] - h->ign[m];
        raw += A * (size_t)mi;
        ncols += mi;
        mi_max = std::max(mi_max, mi);
    }
    CK(cudaSetDevice(s.dev));
    int rc;
    const int rb = (int)(s.fit_next++ & 1);
    if (!s.fit_raw_free[rb]) {
        CK(cudaEventCreateWithFlags(&s.fit_raw_free[rb], cudaEventDisableTiming));
        CK(cudaEventRecord(s.fit_raw_free[rb], s.stream));
    }
    // index tables of the group, built in a page-locked block (two per shard) so their upload is asynchronous too:
    // [src_off nm x8][dst_off ncols x8][mi nm x4][j0 nm x4][m nm x4][dst_ld ncols x4]
    const size_t meta_bytes = (nm + ncols) * 8 + (3 * nm + ncols) * 4;
    if (!s.fit_meta_done[rb]) {
        CK(cudaEventCreateWithFlags(&s.fit_meta_done[rb], cudaEventDisableTiming));
    } else {
        CK(cudaEventSynchronize(s.fit_meta_done[rb]));  // its previous contents have gone up
    }
    if (meta_bytes > s.fit_meta_cap[rb]) {
        if (s.fit_meta[rb]) CK(cudaFreeHost(s.fit_meta[rb]));
        s.fit_meta[rb] = nullptr;
        s.fit_meta_cap[rb] = 0;
        CK(cudaHostAlloc((void**)&s.fit_meta[rb], meta_bytes, cudaHostAllocDefault));
        s.fit_meta_cap[rb] = meta_bytes;
    }
    size_t* src_off = (size_t*)s.fit_meta[rb];
    size_t* dst_off = src_off + nm;
    int* mi = (int*)(dst_off + ncols);
    int* j0 = mi + nm;
    int* mg = j0 + nm;
    int* dst_ld = mg + nm;
    {
        size_t r = 0, c = 0;
        for (size_t b = 0; b < nm; b++) {
            size_t m = ma + b;
            mi[b] = h->sp[m] - h->ign[m];
            j0[b] = (int)c;
            mg[b] = (int)m;
            src_off[b] = r;
            r += A * (size_t)mi[b];
            for (int k = 0; k < mi[b]; k++) {
                dst_off[c + k] = h->row_off_unit[m] * (size_t)s.nq + h->ign[m] + k;
                dst_ld[c + k] = h->ldm[m];
            }
            c += mi[b];
        }
    }
    // (a growing buffer is freed with cudaFree, which drains the device first)
    if ((rc = fit_grow(h, &s.fit_raw[rb], &s.fit_raw_cap[rb], raw))) return rc;
    size_t cap_nm = s.fit_nm_cap, cap_cols = s.fit_cols_cap;
    if ((rc = fit_grow(h, &s.d_fit_src_off, &cap_nm, nm))) return rc;
    cap_nm = s.fit_nm_cap;
    if ((rc = fit_grow(h, &s.d_fit_mi, &cap_nm, nm))) return rc;
    cap_nm = s.fit_nm_cap;
    if ((rc = fit_grow(h, &s.d_fit_j0, &cap_nm, nm))) return rc;
    cap_nm = s.fit_nm_cap;
    if ((rc = fit_grow(h, &s.d_fit_m, &cap_nm, nm))) return rc;
    s.fit_nm_cap = cap_nm;
    if ((rc = fit_grow(h, &s.d_fit_dst_off, &cap_cols, ncols))) return rc;
    cap_cols = s.fit_cols_cap;
    if ((rc = fit_grow(h, &s.d_fit_dst_ld, &cap_cols, ncols))) return rc;
    s.fit_cols_cap = cap_cols;
    // copy stream: wait until the kernels that read this raw buffer two groups ago are done, then bring the block up
    CK(cudaStreamWaitEvent(s.copy, s.fit_raw_free[rb], 0));
    CK(cudaMemcpyAsync(s.fit_raw[rb], dma_src, raw * 8, cudaMemcpyHostToDevice, s.copy));
    CK(cudaEventRecord(ev_h2d, s.copy));
    CK(cudaStreamWaitEvent(s.stream, ev_h2d, 0));
    // index tables: stream-ordered behind the previous group's kernels, which read the same device arrays
    CK(cudaMemcpyAsync(s.d_fit_src_off, src_off, nm * sizeof(size_t), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_fit_mi, mi, nm * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_fit_j0, j0, nm * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_fit_m, mg, nm * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    if (with_metric && s.nq && ncols) {
        CK(cudaMemcpyAsync(s.d_fit_dst_off, dst_off, ncols * sizeof(size_t), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_fit_dst_ld, dst_ld, ncols * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    }
    CK(cudaEventRecord(s.fit_meta_done[rb], s.stream));
    FitGroup g{s.d_fit_src_off, s.d_fit_mi, s.d_fit_j0, s.d_fit_m};
    if (s.nq == 0 || ncols == 0) {
        CK(cudaEventRecord(s.fit_raw_free[rb], s.stream));
        return 0;
    }
    if (with_metric) {
        if ((rc = fit_grow(h, &s.fit_t, &s.fit_t_cap, ncols * (size_t)apitch))) return rc;
        dim3 tg((mi_max + 31) / 32, (apitch + 31) / 32, (unsigned)nm);
        fit_transpose_kernel<<<tg, dim3(32, 8), 0, s.stream>>>(s.fit_raw[rb], g, (int)A, apitch, s.fit_t);
        CK(cudaGetLastError());
        CK(cudaEventRecord(s.fit_raw_free[rb], s.stream));  // the GEMM reads the transposed copy
        static bool attr_set[64] = {false};
        constexpr size_t smem = ws_smem_bytes<8>();
        if (!attr_set[s.dev]) {
            CK(cudaFuncSetAttribute(fit_gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[s.dev] = true;
        }
        CUtensorMap metmap, umap;
        if ((rc = make_map(h, &metmap, s.d_metric, A, (uint64_t)s.nq, (uint64_t)apitch * 8, BM))) return rc;
        if ((rc = make_map(h, &umap, s.fit_t, A, ncols, (uint64_t)apitch * 8, 128))) return rc;
        FitGemmParams p;
        p.nq = s.nq;
        p.ncols = (int)ncols;
        p.kdim = (int)A;
        p.ntm = (s.nq + BM - 1) / BM;
        p.ntn = ((int)ncols + 127) / 128;
        p.nitems = p.ntm * p.ntn;
        p.dst_off = s.d_fit_dst_off;
        p.dst_ld = s.d_fit_dst_ld;
        p.tensor = s.tensor[which];
        p.counter = s.d_counter;
        b200jk::FitTiming ft;
        ft.shard = si;
        ft.flops = 2.0 * (double)s.nq * (double)ncols * (double)A;
        CK(cudaEventCreate(&ft.a));
        CK(cudaEventCreate(&ft.b));
        CK(cudaMemsetAsync(s.d_counter, 0, sizeof(int), s.stream));
        CK(cudaEventRecord(ft.a, s.stream));
        fit_gemm_ws_kernel<<<std::min(p.nitems, s.nsm), WS_THREADS, smem, s.stream>>>(metmap, umap, p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ft.b, s.stream));
        h->fit_pending.push_back(ft);
    } else {
        dim3 sg((mi_max + 127) / 128, (unsigned)s.nq, (unsigned)nm);
        fit_scatter_kernel<<<sg, 128, 0, s.stream>>>(s.fit_raw[rb], g, s.q0, s.d_row_off, s.d_ldm, s.d_ign, s.tensor[which]);
        CK(cudaGetLastError());
        CK(cudaEventRecord(s.fit_raw_free[rb], s.stream));
    }
    dim3 mgrid((mi_max + 127) / 128, (unsigned)s.nq, (unsigned)nm);
    fit_mirror_kernel<<<mgrid, 128, 0, s.stream>>>(g, s.d_row_off, s.d_ldm, s.d_ign, s.d_cols, s.d_cols_off, s.d_mpos,
                                                   s.tensor[which]);
    CK(cudaGetLastError());
    return 0;
}

// Read back the metric-GEMM timings queued so far (waits for those kernels).
int fit_resolve(b200jk* h) {
    for (auto& f : h->fit_pending) {
        CK(cudaSetDevice(h->sh[f.shard].dev));
        CK(cudaEventSynchronize(f.b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, f.a, f.b));
        h->ms_fit_gemm += ms;
        h->fit_flops += f.flops;
        CK(cudaEventDestroy(f.a));
        CK(cudaEventDestroy(f.b));
    }
    h->fit_pending.clear();
    return 0;
}

}  // namespace

extern "C" int b200jk_set_metric(b200jk_t* h, const double* metric) {
    if (!h) return B200JK_ERR_INVALID;
    if (!h->have_layout) return fail(h, B200JK_ERR_INVALID, "set_metric before set_layout");
    const size_t A = h->naux;
    const int apitch = round_up((int)A, 2);
    for (auto& s : h->sh) {
        CK(cudaSetDevice(s.dev));
        s.have_metric = false;
        if (!metric) continue;
        if (!s.d_metric) CK(cudaMalloc((void**)&s.d_metric, std::max<size_t>((size_t)s.nq * apitch * 8, 8)));
        if (s.nq) {
            CK(cudaMemsetAsync(s.d_metric, 0, (size_t)s.nq * apitch * 8, s.stream));
            CK(cudaMemcpy2DAsync(s.d_metric, (size_t)apitch * 8, metric + (size_t)s.q0 * A, A * 8, A * 8, (size_t)s.nq,
                                 cudaMemcpyHostToDevice, s.stream));
            CK(cudaStreamSynchronize(s.stream));
        }
        s.have_metric = true;
    }
    return 0;
}

extern "C" int b200jk_fit_rows(b200jk_t* h, int which, size_t m0, size_t m1, const double* host_sym) {
    if (h && which >= 0 && which <= 2)
        for (auto& sh : h->sh) sh.i8h.expo_valid[which] = false;  // the row scales of the INT8 half transform follow the tensor
    if (!h) return B200JK_ERR_INVALID;
    if (!h->have_layout) return fail(h, B200JK_ERR_INVALID, "fit_rows before set_layout");
    if (which < 0 || which > 2 || m0 > m1 || m1 > h->nbf || !host_sym) return fail(h, B200JK_ERR_INVALID, "bad fit_rows args");
    const bool with_metric = h->sh[0].have_metric;
    const double t_call = now_s();
    h->stage_wait_s = h->stage_alloc_s = h->stage_copy_s = 0;
    int rc = alloc_tensor(h, which);
    if (rc) return rc;
    if (m0 == 0) {
        if ((rc = fit_resolve(h))) return rc;
        h->ms_fit_gemm = h->fit_flops = 0;
    }
    const size_t A = h->naux;
    // Groups of row-blocks of <= 256 MB of raw integrals.  Every local shard needs the whole group (the contraction
    // runs over all of the auxiliary index), so a pageable block is staged ONCE in the page-locked ring and DMA'd to
    // all shards concurrently; the call returns as soon as the caller's block has been read -- psi4 computes the next
    // block of integrals (dfhelper.cc:566-585) while the GPUs transpose, contract and mirror this one.
    const size_t budget = stage_budget((size_t)256 << 20);
    const size_t total_bytes = (h->symm_big_skips[m1] - h->symm_big_skips[m0]) * sizeof(double);
    const bool direct = is_pinned(h, host_sym, total_bytes);
    std::vector<std::pair<int, cudaEvent_t>> h2d;  // direct mode: copies that still read the caller's block
    size_t ma = m0;
    while (ma < m1) {
        size_t mb = ma, bytes = 0;
        while (mb < m1) {
            size_t add = A * (size_t)(h->sp[mb] - h->ign[mb]) * 8;
            if (mb > ma && bytes + add > budget) break;
            bytes += add;
            mb++;
        }
        const double* src = host_sym + (h->symm_big_skips[ma] - h->symm_big_skips[m0]);
        b200jk::StageSlot* slot = nullptr;
        if (!direct) {
            if ((rc = stage_acquire(h, std::max<size_t>(bytes / 8, 1), std::min<size_t>(budget / 8, h->symm_big_skips[h->nbf]),
                                    &slot)))
                return rc;
            const double tc = now_s();
            par_memcpy(slot->buf, src, bytes);
            h->stage_copy_s += now_s() - tc;
            src = slot->buf;
        }
        for (size_t si = 0; si < h->sh.size(); si++) {
            Shard& s = h->sh[si];
            CK(cudaSetDevice(s.dev));
            cudaEvent_t ev;
            if (slot) {
                ev = slot->done[si];
            } else {
                CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                h2d.push_back({s.dev, ev});
            }
            if ((rc = fit_group(h, s, (int)si, which, ma, mb, src, with_metric, ev))) return rc;
        }
        ma = mb;
    }
    const double ts = now_s();
    for (auto& e : h2d) {
        CK(cudaSetDevice(e.first));
        CK(cudaEventSynchronize(e.second));
        CK(cudaEventDestroy(e.second));
    }
    if (stage_trace())
        fprintf(stderr, "[b200jk] fit_rows [%zu,%zu) %.1f MB %s: queued in %.4f s (ring wait %.4f, pin alloc %.4f, copy %.4f, h2d wait %.4f)\n",
                m0, m1, total_bytes / 1e6, direct ? "direct" : "staged", now_s() - t_call, h->stage_wait_s, h->stage_alloc_s,
                h->stage_copy_s, now_s() - ts);
    if (m1 == h->nbf) {
        // the last block completes the tensor: drain the pipeline so "uploaded" means resident and fitted
        if ((rc = producers_sync(h))) return rc;
        if ((rc = fit_resolve(h))) return rc;
        h->uploaded[which] = true;
        h->stats.hbm_tensor_bytes = 0;
        for (int w = 0; w < 3; w++)
            if (h->sh[0].tensor[w]) h->stats.hbm_tensor_bytes += h->sh[0].tensor_doubles * 8;
    }
    return 0;
}

extern "C" int b200jk_fit_stats(const b200jk_t* h, double* ms_gemm, double* flops) {
    if (!h || !ms_gemm || !flops) return B200JK_ERR_INVALID;
    if (fit_resolve(const_cast<b200jk_t*>(h))) return B200JK_ERR_CUDA;
    *ms_gemm = h->ms_fit_gemm;
    *flops = h->fit_flops;
    return 0;
}
