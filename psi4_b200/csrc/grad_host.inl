// grad_host.inl -- SURVEY.md 8f row f4: the tensor contractions of the density-fitted SCF gradient (DFJKGrad,
// scfgrad/jk_grad.cc:175-1245) on the tensor that is already resident in HBM; included at the end of engine.cu.
//
// The reference recomputes the UNFITTED integrals (A|mn) block by block on the CPU and forms (jk_grad.cc):
//   c_A      = (A|mn) Dt_mn                                  build_Amn_terms      :294-475
//   (A|ij)   = C^T (A|mn) C                                  build_Amn_terms
//   d        = J^-1 c ,  (A|ij) <- J^-1 (B|ij)               build_AB_inv_terms   :634-721
//   V_AB     = f * sum_spin sum_ij (A|ij)(B|ij)              build_UV_terms       :722-833   (f = 2 if restricted)
//   Kmn[A]   = f * sum_spin C (A|ij) C^T                     build_Amn_x_terms    :1010-1023
// and then contracts d d^T / V with (A|B)^x and d Dt / Kmn with (A|mn)^x (Libint2 derivative integrals, which stay on
// the host).  The engine holds B = J^-1/2 (A|mn), so with t_Q = B_Q . Dt and (Q|ij) = C^T B_Q C:
//   d = J^-1/2 t ,   (A|ij)_fitted = J^-1/2 (Q|ij) ,
// i.e. every intermediate above comes from ONE pass over the resident tensor (the first J sweep + the half transform of
// the K build) plus small GEMMs with the metric power -- no integral is recomputed.  The caller passes J^-1/2.
//
// Scope: handles with one Q shard (one GPU).  The gradient runs once per geometry, not once per SCF iteration.

namespace {

int launch_gemm(b200jk* h, Shard& s, const GemmStrided& g) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return 0;
    dim3 grid((g.N + GS_T - 1) / GS_T, (g.M + GS_T - 1) / GS_T, g.batch);
    gemm_strided_kernel<<<grid, GS_THREADS, 0, s.stream>>>(g);
    s.launches++;
    CK(cudaGetLastError());
    return 0;
}

GemmStrided gemm_desc(int M, int N, int K, const double* A, long long sAm, long long sAk, const double* B, long long sBn,
                      long long sBk, double* C, long long sCm, long long sCn, double alpha, double beta) {
    GemmStrided g;
    g.M = M;
    g.N = N;
    g.K = K;
    g.batch = 1;
    g.A = A;
    g.sAm = sAm;
    g.sAk = sAk;
    g.bA = 0;
    g.B = B;
    g.sBn = sBn;
    g.sBk = sBk;
    g.bB = 0;
    g.C = C;
    g.sCm = sCm;
    g.sCn = sCn;
    g.bC = 0;
    g.alpha = alpha;
    g.beta = beta;
    return g;
}

void grad_free(b200jk* h) {
    if (h->sh.empty()) return;
    Shard& s = h->sh[0];
    cudaSetDevice(s.dev);
    for (auto& sp : h->grad.spin) {
        if (sp.C) cudaFree(sp.C);
        if (sp.cfit) cudaFree(sp.cfit);
        sp = b200jk::Grad::Spin();
    }
    void* ptrs[] = {h->grad.Jm12, h->grad.d, h->grad.V, h->grad.rows, h->grad.M1};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    h->grad = b200jk::Grad();
}

}  // namespace

extern "C" {

int b200jk_grad_begin(b200jk_t* h, int nspin, const double* const* C, const int* nocc, const double* Dt,
                      const double* Jm12) {
    if (!h) return B200JK_ERR_INVALID;
    if (!h->have_layout || !h->uploaded[0]) return fail(h, B200JK_ERR_INVALID, "grad_begin: tensor Ppq not resident");
    if (h->sh.size() != 1 || (h->rank_mode && h->world > 1))
        return fail(h, B200JK_ERR_INVALID, "grad_begin: the gradient contractions run on a one-shard handle (one GPU)");
    if (nspin < 1 || nspin > 2 || !C || !nocc || !Dt || !Jm12) return fail(h, B200JK_ERR_INVALID, "grad_begin: bad arguments");
    for (int sp = 0; sp < nspin; sp++)
        if (nocc[sp] < 0 || (nocc[sp] > 0 && !C[sp])) return fail(h, B200JK_ERR_INVALID, "grad_begin: bad C / nocc");
    grad_free(h);
    Shard& s = h->sh[0];
    CK(cudaSetDevice(s.dev));
    CK(cudaStreamSynchronize(s.stream));
    const size_t N = h->nbf, A = h->naux, n2 = N * N;
    const int ldc = round_up((int)N, 4);
    int rc;
    b200jk::Grad& g = h->grad;
    g.nspin = nspin;

    // work buffers of a K-only build (half-transformed T, C^T, the packed C^T of screened row-blocks) + d_part
    Task t;
    t.nmat = 1;
    int max_o = std::max(nocc[0], nspin > 1 ? nocc[1] : 0);
    t.nocc = &max_o;
    t.lr = true;
    t.do_J = true;
    t.do_K = true;
    t.do_wK = false;
    t.max_o = max_o;
    t.n2 = n2;
    t.nprod = 2;
    t.same_left.assign(1, 0);
    int qc = 0;
    begin_compute(h);
    if ((rc = ensure_work(h, s, t, &qc))) return rc;

    // ---- t_Q = B_Q . Dt (first J sweep, symmetric density), d = J^-1/2 t ----
    CK(cudaMalloc((void**)&g.Jm12, A * A * sizeof(double)));
    CK(cudaMalloc((void**)&g.d, A * sizeof(double)));
    CK(cudaMalloc((void**)&g.V, A * A * sizeof(double)));
    CK(cudaMemcpyAsync(g.Jm12, Jm12, A * A * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    if ((rc = grow(h, &s.in, &s.in_cap, n2))) return rc;
    CK(cudaMemcpyAsync(s.in, Dt, n2 * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemsetAsync(g.V, 0, A * A * sizeof(double), s.stream));
    {
        JParams p;
        p.tensor = s.tensor[B200JK_TENSOR_PPQ];
        p.row_off = s.d_row_off;
        p.ldm = s.d_ldm;
        p.sp = s.d_sp;
        p.ign = s.d_ign;
        p.cols = s.d_cols;
        p.cols_off = s.d_cols_off;
        p.nbf = (int)N;
        p.nq = s.nq;
        p.symmetric = 1;
        p.D = s.in;
        p.dpart = s.dpart;
        p.d = s.dvec;
        p.J = nullptr;
        static bool attr_set[64] = {false};
        if (!attr_set[s.dev]) {
            CK(cudaFuncSetAttribute(j_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[s.dev] = true;
        }
        const size_t sm1 = (size_t)(h->max_sp + 2) * sizeof(double);
        if (sm1 > 200 * 1024) return fail(h, B200JK_ERR_INVALID, "grad_begin: nbf exceeds the J kernels' staging limit");
        j_dq_kernel<<<dim3((s.nq + J1_ROWS - 1) / J1_ROWS, (unsigned)N), J_THREADS, sm1, s.stream>>>(p);
        s.launches++;
        CK(cudaGetLastError());
        j_dq_reduce_kernel<<<(s.nq + JR_Q - 1) / JR_Q, JR_Q * JR_M, 0, s.stream>>>(s.dpart, p.nbf, s.nq, s.dvec);
        s.launches++;
        CK(cudaGetLastError());
        // d[a] = sum_Q Jm12[a][Q] t[Q]
        if ((rc = launch_gemm(h, s, gemm_desc((int)A, 1, (int)A, g.Jm12, (long long)A, 1, s.dvec, 0, 1, g.d, 1, 0, 1.0, 0.0))))
            return rc;
    }

    // ---- per spin: (Q|mi) by the half transform, (Q|ij) = C^T (Q|mi), fitted c = J^-1/2 (Q|ij), V += f c c^T ----
    const double f = nspin == 1 ? 2.0 : 1.0;  // "restricted": one transform counted twice (jk_grad.cc:808-810)
    for (int sp = 0; sp < nspin; sp++) {
        const int o = nocc[sp];
        b200jk::Grad::Spin& gs = g.spin[sp];
        gs.o = o;
        if (!o) continue;  // jk_grad.cc:455-456: "skip if there are no beta electrons"
        const int op = round_up(o, 2);
        gs.op = op;
        const size_t oop = (size_t)o * op;
        double* qij = nullptr;
        CK(cudaMalloc((void**)&gs.C, N * (size_t)o * sizeof(double)));
        CK(cudaMalloc((void**)&qij, A * oop * sizeof(double)));
        CK(cudaMalloc((void**)&gs.cfit, A * oop * sizeof(double)));
        CK(cudaMemcpyAsync(gs.C, C[sp], N * (size_t)o * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        if ((rc = run_transpose(h, s, gs.C, o, s.Ctl, ldc, op))) return rc;
        for (int qb = 0; qb < s.nq; qb += qc) {
            const int nqc = std::min(qc, s.nq - qb);
            // T[m][q][i] = sum_n B(q, m, n) C[n][i]   (first_transform_pQq; jk_grad.cc:437 does it as one DGEMM)
            if ((rc = run_half(h, s, B200JK_TENSOR_PPQ, s.Ctl, ldc, o, op, qb, nqc, s.T1, nullptr))) return rc;
            // (Q|ij): out[q][i][j] = sum_m C[m][i] T[m][q][j]   (jk_grad.cc:441-444, one DGEMM per aux row)
            GemmStrided gm = gemm_desc(o, op, (int)N, gs.C, 1, (long long)o, s.T1, 1, (long long)nqc * op,
                                       qij + (size_t)qb * oop, (long long)op, 1, 1.0, 0.0);
            gm.batch = nqc;
            gm.bA = 0;
            gm.bB = op;
            gm.bC = (long long)oop;
            if ((rc = launch_gemm(h, s, gm))) return rc;
        }
        // c[a][ij] = sum_Q Jm12[Q][a] (Q|ij)   (Jm12 symmetric; jk_grad.cc:705: J^-1 (B|ij))
        if ((rc = launch_gemm(h, s, gemm_desc((int)A, (int)oop, (int)A, g.Jm12, 1, (long long)A, qij, 1, (long long)oop,
                                              gs.cfit, (long long)oop, 1, 1.0, 0.0))))
            return rc;
        // V[a][b] += f sum_ij c[a][ij] c[b][ij]   (jk_grad.cc:776-777, :800-801, :808-810)
        if ((rc = launch_gemm(h, s, gemm_desc((int)A, (int)A, (int)oop, gs.cfit, (long long)oop, 1, gs.cfit, (long long)oop, 1,
                                              g.V, (long long)A, 1, f, 1.0))))
            return rc;
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaFree(qij));
    }
    CK(cudaStreamSynchronize(s.stream));
    g.active = true;
    return 0;
}

int b200jk_grad_vectors(b200jk_t* h, double* d, double* V) {
    if (!h) return B200JK_ERR_INVALID;
    if (!h->grad.active) return fail(h, B200JK_ERR_INVALID, "grad_vectors before grad_begin");
    Shard& s = h->sh[0];
    CK(cudaSetDevice(s.dev));
    const size_t A = h->naux;
    if (d) CK(cudaMemcpy(d, h->grad.d, A * sizeof(double), cudaMemcpyDeviceToHost));
    if (V) CK(cudaMemcpy(V, h->grad.V, A * A * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int b200jk_grad_rows(b200jk_t* h, size_t a0, size_t a1, double* Kmn) {
    if (!h) return B200JK_ERR_INVALID;
    if (!h->grad.active) return fail(h, B200JK_ERR_INVALID, "grad_rows before grad_begin");
    if (a0 > a1 || a1 > h->naux || !Kmn) return fail(h, B200JK_ERR_INVALID, "grad_rows: bad row range");
    Shard& s = h->sh[0];
    CK(cudaSetDevice(s.dev));
    b200jk::Grad& g = h->grad;
    const size_t N = h->nbf, n2 = N * N, na = a1 - a0;
    if (!na) return 0;
    int rc;
    int max_op = std::max(g.spin[0].op, g.spin[1].op);
    if ((rc = grow(h, &g.rows, &g.rows_cap, na * n2))) return rc;
    if ((rc = grow(h, &g.M1, &g.M1_cap, na * N * (size_t)std::max(max_op, 1)))) return rc;
    CK(cudaMemsetAsync(g.rows, 0, na * n2 * sizeof(double), s.stream));
    const double f = g.nspin == 1 ? 2.0 : 1.0;
    for (int sp = 0; sp < g.nspin; sp++) {
        const b200jk::Grad::Spin& gs = g.spin[sp];
        if (!gs.o) continue;
        const int o = gs.o, op = gs.op;
        const size_t oop = (size_t)o * op;
        // M1[a][m][j] = sum_i C[m][i] c[a][j][i]   ((A|ij) is symmetric in ij; jk_grad.cc:1016-1018)
        GemmStrided g1 = gemm_desc((int)N, o, o, gs.C, (long long)o, 1, gs.cfit + a0 * oop, (long long)op, 1, g.M1,
                                   (long long)op, 1, 1.0, 0.0);
        g1.batch = (int)na;
        g1.bB = (long long)oop;
        g1.bC = (long long)(N * op);
        if ((rc = launch_gemm(h, s, g1))) return rc;
        // Kmn[a][m][n] += f sum_j M1[a][m][j] C[n][j]   (jk_grad.cc:1021-1022)
        GemmStrided g2 = gemm_desc((int)N, (int)N, o, g.M1, (long long)op, 1, gs.C, (long long)o, 1, g.rows, (long long)N, 1, f, 1.0);
        g2.batch = (int)na;
        g2.bA = (long long)(N * op);
        g2.bC = (long long)n2;
        if ((rc = launch_gemm(h, s, g2))) return rc;
    }
    CK(cudaMemcpyAsync(Kmn, g.rows, na * n2 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return 0;
}

int b200jk_grad_end(b200jk_t* h) {
    if (!h) return B200JK_ERR_INVALID;
    grad_free(h);
    return 0;
}

}  // extern "C"
