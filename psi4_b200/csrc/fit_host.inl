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

// One staged group of row-blocks [ma, mb) on one shard.
int fit_group(b200jk* h, Shard& s, int which, size_t ma, size_t mb, const double* host_group, bool with_metric,
              cudaEvent_t ev_a, cudaEvent_t ev_b) {
    const size_t A = h->naux;
    const int apitch = round_up((int)A, 2);
    const size_t nm = mb - ma;
    std::vector<size_t> src_off(nm), dst_off;
    std::vector<int> mi(nm), j0(nm), mg(nm), dst_ld;
    size_t ncols = 0, raw = 0;
    int mi_max = 0;
    for (size_t b = 0; b < nm; b++) {
        size_t m = ma + b;
        mi[b] = h->sp[m] - h->ign[m];
        j0[b] = (int)ncols;
        mg[b] = (int)m;
        src_off[b] = raw;
        raw += A * (size_t)mi[b];
        ncols += mi[b];
        mi_max = std::max(mi_max, mi[b]);
    }
    dst_off.resize(ncols);
    dst_ld.resize(ncols);
    for (size_t b = 0; b < nm; b++) {
        size_t m = ma + b;
        for (int k = 0; k < mi[b]; k++) {
            dst_off[j0[b] + k] = h->row_off_unit[m] * (size_t)s.nq + h->ign[m] + k;
            dst_ld[j0[b] + k] = h->ldm[m];
        }
    }
    CK(cudaSetDevice(s.dev));
    int rc;
    if ((rc = fit_grow(h, &s.fit_raw, &s.fit_raw_cap, raw))) return rc;
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
    CK(cudaMemcpyAsync(s.fit_raw, host_group, raw * 8, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_fit_src_off, src_off.data(), nm * sizeof(size_t), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_fit_mi, mi.data(), nm * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_fit_j0, j0.data(), nm * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_fit_m, mg.data(), nm * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    FitGroup g{s.d_fit_src_off, s.d_fit_mi, s.d_fit_j0, s.d_fit_m};
    if (s.nq == 0 || ncols == 0) {
        CK(cudaStreamSynchronize(s.stream));
        return 0;
    }
    if (with_metric) {
        if ((rc = fit_grow(h, &s.fit_t, &s.fit_t_cap, ncols * (size_t)apitch))) return rc;
        CK(cudaMemcpyAsync(s.d_fit_dst_off, dst_off.data(), ncols * sizeof(size_t), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_fit_dst_ld, dst_ld.data(), ncols * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        dim3 tg((mi_max + 31) / 32, (apitch + 31) / 32, (unsigned)nm);
        fit_transpose_kernel<<<tg, dim3(32, 8), 0, s.stream>>>(s.fit_raw, g, (int)A, apitch, s.fit_t);
        CK(cudaGetLastError());
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
        CK(cudaMemsetAsync(s.d_counter, 0, sizeof(int), s.stream));
        CK(cudaEventRecord(ev_a, s.stream));
        fit_gemm_ws_kernel<<<std::min(p.nitems, s.nsm), WS_THREADS, smem, s.stream>>>(metmap, umap, p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev_b, s.stream));
    } else {
        dim3 sg((mi_max + 127) / 128, (unsigned)s.nq, (unsigned)nm);
        fit_scatter_kernel<<<sg, 128, 0, s.stream>>>(s.fit_raw, g, s.q0, s.d_row_off, s.d_ldm, s.d_ign, s.tensor[which]);
        CK(cudaGetLastError());
    }
    dim3 mgrid((mi_max + 127) / 128, (unsigned)s.nq, (unsigned)nm);
    fit_mirror_kernel<<<mgrid, 128, 0, s.stream>>>(g, s.d_row_off, s.d_ldm, s.d_ign, s.d_cols, s.d_cols_off, s.d_mpos,
                                                   s.tensor[which]);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s.stream));  // host vectors and the staging buffers are reused by the next group
    if (with_metric) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ev_a, ev_b));
        h->ms_fit_gemm += ms;
        h->fit_flops += 2.0 * (double)s.nq * (double)ncols * (double)A;
    }
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
    if (!h) return B200JK_ERR_INVALID;
    if (!h->have_layout) return fail(h, B200JK_ERR_INVALID, "fit_rows before set_layout");
    if (which < 0 || which > 2 || m0 > m1 || m1 > h->nbf || !host_sym) return fail(h, B200JK_ERR_INVALID, "bad fit_rows args");
    const bool with_metric = h->sh[0].have_metric;
    int rc = alloc_tensor(h, which);
    if (rc) return rc;
    if (m0 == 0) h->ms_fit_gemm = h->fit_flops = 0;
    const size_t A = h->naux;
    // staged groups of row-blocks: <= ~1.5 GB of raw integrals each (and as much again transposed)
    const size_t budget = (size_t)3 << 29;
    for (auto& s : h->sh) {
        CK(cudaSetDevice(s.dev));
        cudaEvent_t ea, eb;
        CK(cudaEventCreate(&ea));
        CK(cudaEventCreate(&eb));
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
            if ((rc = fit_group(h, s, which, ma, mb, src, with_metric, ea, eb))) return rc;
            ma = mb;
        }
        CK(cudaEventDestroy(ea));
        CK(cudaEventDestroy(eb));
    }
    if (m1 == h->nbf) {
        h->uploaded[which] = true;
        h->stats.hbm_tensor_bytes = 0;
        for (int w = 0; w < 3; w++)
            if (h->sh[0].tensor[w]) h->stats.hbm_tensor_bytes += h->sh[0].tensor_doubles * 8;
    }
    return 0;
}

extern "C" int b200jk_fit_stats(const b200jk_t* h, double* ms_gemm, double* flops) {
    if (!h || !ms_gemm || !flops) return B200JK_ERR_INVALID;
    *ms_gemm = h->ms_fit_gemm;
    *flops = h->fit_flops;
    return 0;
}
