/*
 * ints.c -- host-side Gaussian integral front end (libb200ints.so): the input producer of the DF-JK
 * path.  psi4 gets these numbers from Libint2 (not vendored); the engine only needs the
 * values, so this is an independent McMurchie-Davidson implementation over CARTESIAN shells:
 *
 *   one-electron   S, T, V                      (libmints OneBodyAOInt role)
 *   two-centre     (A|B)                        lib3index/fittingmetric.cc:72-155
 *   three-centre   (A|mn)                       lib3index/dfhelper.cc:1284-1347 (compute_sparse_pQq_blocking_p_symm)
 *   pair diagonal  (mn|mn) shell-pair blocks    lib3index/dfhelper.cc:330-369 (Schwarz screening input)
 *
 * Spherical-harmonic transformation, normalisation and contraction bookkeeping live in
 * psi4_b200/integrals.py.  Cartesian component order inside a shell: lx descending, then ly descending
 * (xx, xy, xz, yy, yz, zz ...).  Contraction coefficients passed in already include the primitive
 * normalisation.  Stays on the CPU, like its counterpart in the reference.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LMAX 5               /* per-shell angular momentum limit (h) */
#define LPAIR (2 * LMAX)     /* Hermite order of a product */
#define LTOT (2 * LPAIR)     /* Hermite order of a quartet */
#define NCART(l) (((l) + 1) * ((l) + 2) / 2)

typedef struct {
    int nshell;
    const double* xyz;   /* [nshell][3] bohr */
    const int* l;        /* [nshell] */
    const int* nprim;    /* [nshell] */
    const int* poff;     /* [nshell] offset into exps/coefs */
    const double* exps;
    const double* coefs; /* include primitive normalisation */
} basis_t;

/* ---- Boys function F_0..F_nmax(x) --------------------------------------------------------- */
static void boys(int nmax, double x, double* F) {
    if (x < 35.0) {
        /* series for the highest order, then downward recursion (stable) */
        double ex = exp(-x);
        double term = 1.0 / (2.0 * nmax + 1.0), sum = term;
        for (int k = 1; k < 400; k++) {
            term *= 2.0 * x / (2.0 * nmax + 2.0 * k + 1.0);
            sum += term;
            if (term < 1e-17 * sum) break;
        }
        F[nmax] = ex * sum;
        for (int n = nmax; n > 0; n--) F[n - 1] = (2.0 * x * F[n] + ex) / (2.0 * n - 1.0);
    } else {
        /* asymptotic F_0 (erf(sqrt(x)) == 1 to double precision) + upward recursion (stable for large x) */
        double ex = exp(-x);
        F[0] = 0.5 * sqrt(M_PI / x);
        for (int n = 0; n < nmax; n++) F[n + 1] = ((2.0 * n + 1.0) * F[n] - ex) / (2.0 * x);
    }
}

/* ---- Hermite expansion coefficients E[i][j][t], 1-D -------------------------------------- */
typedef double herm_t[LMAX + 3][LMAX + 3][2 * LMAX + 6];
static void hermite_E(int la, int lb, double a, double b, double A, double B, herm_t E) {
    double p = a + b, P = (a * A + b * B) / p;
    double XPA = P - A, XPB = P - B, mu = a * b / p, XAB = A - B;
    double i2p = 0.5 / p;
    for (int i = 0; i <= la; i++)
        for (int j = 0; j <= lb; j++)
            for (int t = 0; t <= la + lb + 1; t++) E[i][j][t] = 0.0;
    E[0][0][0] = exp(-mu * XAB * XAB);
    for (int i = 0; i < la; i++)
        for (int t = 0; t <= i + 1; t++) {
            double v = XPA * E[i][0][t] + (t + 1) * E[i][0][t + 1];
            if (t > 0) v += i2p * E[i][0][t - 1];
            E[i + 1][0][t] = v;
        }
    for (int i = 0; i <= la; i++)
        for (int j = 0; j < lb; j++)
            for (int t = 0; t <= i + j + 1; t++) {
                double v = XPB * E[i][j][t] + (t + 1) * E[i][j][t + 1];
                if (t > 0) v += i2p * E[i][j][t - 1];
                E[i][j + 1][t] = v;
            }
}

/* ---- Hermite Coulomb integrals R_{tuv}(alpha, PC), t+u+v <= L ------------------------------ */
#define RD (LTOT + 1)
static void hermite_R(int L, double alpha, double X, double Y, double Z, double* R0 /* [RD][RD][RD] */, double* work) {
    /* work: [(L+1)][RD][RD][RD] levels n */
    double F[LTOT + 2];
    boys(L, alpha * (X * X + Y * Y + Z * Z), F);
    size_t lvl = (size_t)RD * RD * RD;
#define RN(n, t, u, v) work[(size_t)(n)*lvl + ((size_t)(t)*RD + (u)) * RD + (v)]
    double m2a = 1.0;
    for (int n = 0; n <= L; n++) {
        RN(n, 0, 0, 0) = m2a * F[n];
        m2a *= -2.0 * alpha;
    }
    for (int tot = 1; tot <= L; tot++) {
        for (int n = 0; n <= L - tot; n++) {
            for (int t = 0; t <= tot; t++)
                for (int u = 0; u <= tot - t; u++) {
                    int v = tot - t - u;
                    double val;
                    if (t > 0) {
                        val = X * RN(n + 1, t - 1, u, v);
                        if (t > 1) val += (t - 1) * RN(n + 1, t - 2, u, v);
                    } else if (u > 0) {
                        val = Y * RN(n + 1, t, u - 1, v);
                        if (u > 1) val += (u - 1) * RN(n + 1, t, u - 2, v);
                    } else {
                        val = Z * RN(n + 1, t, u, v - 1);
                        if (v > 1) val += (v - 1) * RN(n + 1, t, u, v - 2);
                    }
                    RN(n, t, u, v) = val;
                }
        }
    }
    for (int t = 0; t <= L; t++)
        for (int u = 0; u <= L - t; u++)
            for (int v = 0; v <= L - t - u; v++) R0[((size_t)t * RD + u) * RD + v] = RN(0, t, u, v);
#undef RN
}

static int cart_list(int l, int (*c)[3]) {
    int n = 0;
    for (int lx = l; lx >= 0; lx--)
        for (int ly = l - lx; ly >= 0; ly--) {
            c[n][0] = lx;
            c[n][1] = ly;
            c[n][2] = l - lx - ly;
            n++;
        }
    return n;
}

typedef struct {
    double x, y, z;
    int l, nprim;
    const double *e, *c;
} sh_t;
static sh_t get_shell(const basis_t* b, int s) {
    sh_t r = {b->xyz[3 * s], b->xyz[3 * s + 1], b->xyz[3 * s + 2], b->l[s], b->nprim[s], b->exps + b->poff[s],
              b->coefs + b->poff[s]};
    return r;
}
static const double UNIT_E[1] = {0.0}, UNIT_C[1] = {1.0};
static sh_t unit_shell(sh_t at) { /* s "function" == 1 on the same centre: turns a pair into a single centre */
    sh_t r = {at.x, at.y, at.z, 0, 1, UNIT_E, UNIT_C};
    return r;
}

/* Range separation: omega > 0 switches the operator from 1/r12 to erf(omega r12)/r12 (IntegralFactory::erf_eri,
 * used for wPpq_ in DFHelper::prepare_AO_wK_core, lib3index/dfhelper.cc:589-699).  For two Gaussian charge
 * distributions with reduced exponent alpha this is the Coulomb formula with alpha_w = alpha w^2 / (alpha + w^2) in the
 * Hermite integrals and an extra factor sqrt(alpha_w / alpha).  Set by the *_erf entry points before their parallel
 * region and reset after it. */
static double g_omega = 0.0;

/* ---- contracted cartesian shell quartet (ab|cd) -> out[na][nb][nc][nd] -------------------------- */
#define MAXC NCART(LMAX)
static void eri_quartet(sh_t A, sh_t B, sh_t C, sh_t D, double* out, double* work /* >= (LTOT+2)*RD^3 + RD^3 */) {
    int ca[MAXC][3], cb[MAXC][3], cc[MAXC][3], cd[MAXC][3];
    int na = cart_list(A.l, ca), nb = cart_list(B.l, cb), nc = cart_list(C.l, cc), nd = cart_list(D.l, cd);
    int Lb = A.l + B.l, Lk = C.l + D.l, L = Lb + Lk;
    size_t nout = (size_t)na * nb * nc * nd;
    memset(out, 0, sizeof(double) * nout);
    double* R0 = work;
    double* Rw = work + (size_t)RD * RD * RD;
    static const int HB = LPAIR + 1;
    double* Hcd = (double*)malloc(sizeof(double) * HB * HB * HB);
    herm_t E1x, E1y, E1z, E2x, E2y, E2z;
    for (int ia = 0; ia < A.nprim; ia++)
        for (int ib = 0; ib < B.nprim; ib++) {
            double a = A.e[ia], b = B.e[ib], p = a + b;
            if (p <= 0.0) continue;
            double Px = (a * A.x + b * B.x) / p, Py = (a * A.y + b * B.y) / p, Pz = (a * A.z + b * B.z) / p;
            hermite_E(A.l, B.l, a, b, A.x, B.x, E1x);
            hermite_E(A.l, B.l, a, b, A.y, B.y, E1y);
            hermite_E(A.l, B.l, a, b, A.z, B.z, E1z);
            double cab = A.c[ia] * B.c[ib];
            for (int ic = 0; ic < C.nprim; ic++)
                for (int id = 0; id < D.nprim; id++) {
                    double c = C.e[ic], d = D.e[id], q = c + d;
                    double Qx = (c * C.x + d * D.x) / q, Qy = (c * C.y + d * D.y) / q, Qz = (c * C.z + d * D.z) / q;
                    hermite_E(C.l, D.l, c, d, C.x, D.x, E2x);
                    hermite_E(C.l, D.l, c, d, C.y, D.y, E2y);
                    hermite_E(C.l, D.l, c, d, C.z, D.z, E2z);
                    double alpha = p * q / (p + q);
                    double pref = 2.0 * pow(M_PI, 2.5) / (p * q * sqrt(p + q)) * cab * C.c[ic] * D.c[id];
                    if (g_omega > 0.0) {
                        double aw = alpha * g_omega * g_omega / (alpha + g_omega * g_omega);
                        pref *= sqrt(aw / alpha);
                        alpha = aw;
                    }
                    hermite_R(L, alpha, Px - Qx, Py - Qy, Pz - Qz, R0, Rw);
                    for (int kc = 0; kc < nc; kc++)
                        for (int kd = 0; kd < nd; kd++) {
                            int lx = cc[kc][0] + cd[kd][0], ly = cc[kc][1] + cd[kd][1], lz = cc[kc][2] + cd[kd][2];
                            /* H[t][u][v] = sum_{tau,nu,phi} (-1)^(tau+nu+phi) E2 R[t+tau][u+nu][v+phi] */
                            for (int t = 0; t <= Lb; t++)
                                for (int u = 0; u <= Lb - t; u++)
                                    for (int v = 0; v <= Lb - t - u; v++) {
                                        double s = 0.0;
                                        for (int tau = 0; tau <= lx; tau++) {
                                            double ex = E2x[cc[kc][0]][cd[kd][0]][tau];
                                            for (int nu = 0; nu <= ly; nu++) {
                                                double exy = ex * E2y[cc[kc][1]][cd[kd][1]][nu];
                                                for (int phi = 0; phi <= lz; phi++) {
                                                    double sg = ((tau + nu + phi) & 1) ? -1.0 : 1.0;
                                                    s += sg * exy * E2z[cc[kc][2]][cd[kd][2]][phi] *
                                                         R0[((size_t)(t + tau) * RD + (u + nu)) * RD + (v + phi)];
                                                }
                                            }
                                        }
                                        Hcd[(t * HB + u) * HB + v] = s;
                                    }
                            for (int ka = 0; ka < na; ka++)
                                for (int kb = 0; kb < nb; kb++) {
                                    int mx = ca[ka][0] + cb[kb][0], my = ca[ka][1] + cb[kb][1], mz = ca[ka][2] + cb[kb][2];
                                    double s = 0.0;
                                    for (int t = 0; t <= mx; t++) {
                                        double ex = E1x[ca[ka][0]][cb[kb][0]][t];
                                        for (int u = 0; u <= my; u++) {
                                            double exy = ex * E1y[ca[ka][1]][cb[kb][1]][u];
                                            for (int v = 0; v <= mz; v++)
                                                s += exy * E1z[ca[ka][2]][cb[kb][2]][v] * Hcd[(t * HB + u) * HB + v];
                                        }
                                    }
                                    out[(((size_t)ka * nb + kb) * nc + kc) * nd + kd] += pref * s;
                                }
                        }
                }
        }
    free(Hcd);
}
static size_t work_doubles(void) { return (size_t)(LTOT + 3) * RD * RD * RD; }

static void cart_offsets(const basis_t* b, int* off) {
    int o = 0;
    for (int s = 0; s < b->nshell; s++) {
        off[s] = o;
        o += NCART(b->l[s]);
    }
    off[b->nshell] = o;
}

/* ---- public entry points (ctypes) --------------------------------------------------------------- */
static basis_t mk(int ns, const double* xyz, const int* l, const int* np, const int* po, const double* e, const double* c) {
    basis_t b = {ns, xyz, l, np, po, e, c};
    return b;
}

int ints_ncart(int ns, const int* l) {
    int n = 0;
    for (int s = 0; s < ns; s++) n += NCART(l[s]);
    return n;
}

/* S, T, V over cartesian functions; V = sum_C -Z_C <a| 1/r_C |b> */
int ints_one_electron(int ns, const double* xyz, const int* l, const int* np, const int* po, const double* e,
                      const double* c, int natom, const double* Z, const double* axyz, double* S, double* T, double* V) {
    basis_t bs = mk(ns, xyz, l, np, po, e, c);
    for (int s = 0; s < ns; s++)
        if (l[s] > LMAX) return 1;
    int* off = (int*)malloc(sizeof(int) * (ns + 1));
    cart_offsets(&bs, off);
    int n = off[ns];
    memset(S, 0, sizeof(double) * n * n);
    memset(T, 0, sizeof(double) * n * n);
    memset(V, 0, sizeof(double) * n * n);
#pragma omp parallel
    {
        double* R0 = (double*)malloc(sizeof(double) * work_doubles());
        double* Rw = R0 + (size_t)RD * RD * RD;
        herm_t Ex, Ey, Ez;
#pragma omp for schedule(dynamic) collapse(2)
        for (int sa = 0; sa < ns; sa++)
            for (int sb = 0; sb < ns; sb++) {
                sh_t A = get_shell(&bs, sa), B = get_shell(&bs, sb);
                int ca[MAXC][3], cb[MAXC][3];
                int na = cart_list(A.l, ca), nb = cart_list(B.l, cb);
                for (int ia = 0; ia < A.nprim; ia++)
                    for (int ib = 0; ib < B.nprim; ib++) {
                        double a = A.e[ia], b = B.e[ib], p = a + b;
                        double cc = A.c[ia] * B.c[ib];
                        /* lb+2 needed for the kinetic energy */
                        hermite_E(A.l, B.l + 2, a, b, A.x, B.x, Ex);
                        hermite_E(A.l, B.l + 2, a, b, A.y, B.y, Ey);
                        hermite_E(A.l, B.l + 2, a, b, A.z, B.z, Ez);
                        double s3 = pow(M_PI / p, 1.5);
                        double Px = (a * A.x + b * B.x) / p, Py = (a * A.y + b * B.y) / p, Pz = (a * A.z + b * B.z) / p;
                        for (int ka = 0; ka < na; ka++)
                            for (int kb = 0; kb < nb; kb++) {
                                int i = ca[ka][0], k = ca[ka][1], m = ca[ka][2];
                                int j = cb[kb][0], ll = cb[kb][1], nn = cb[kb][2];
                                double sx = Ex[i][j][0], sy = Ey[k][ll][0], sz = Ez[m][nn][0];
                                size_t idx = (size_t)(off[sa] + ka) * n + off[sb] + kb;
                                S[idx] += cc * s3 * sx * sy * sz;
                                /* T_ij(1D) = -2b^2 S_{i,j+2} + b(2j+1) S_ij - j(j-1)/2 S_{i,j-2} */
                                double tx = -2.0 * b * b * Ex[i][j + 2][0] + b * (2 * j + 1) * sx -
                                            (j > 1 ? 0.5 * j * (j - 1) * Ex[i][j - 2][0] : 0.0);
                                double ty = -2.0 * b * b * Ey[k][ll + 2][0] + b * (2 * ll + 1) * sy -
                                            (ll > 1 ? 0.5 * ll * (ll - 1) * Ey[k][ll - 2][0] : 0.0);
                                double tz = -2.0 * b * b * Ez[m][nn + 2][0] + b * (2 * nn + 1) * sz -
                                            (nn > 1 ? 0.5 * nn * (nn - 1) * Ez[m][nn - 2][0] : 0.0);
                                T[idx] += cc * s3 * (tx * sy * sz + sx * ty * sz + sx * sy * tz);
                            }
                        int L = A.l + B.l;
                        for (int at = 0; at < natom; at++) {
                            hermite_R(L, p, Px - axyz[3 * at], Py - axyz[3 * at + 1], Pz - axyz[3 * at + 2], R0, Rw);
                            double pref = -Z[at] * 2.0 * M_PI / p * cc;
                            for (int ka = 0; ka < na; ka++)
                                for (int kb = 0; kb < nb; kb++) {
                                    int i = ca[ka][0], k = ca[ka][1], m = ca[ka][2];
                                    int j = cb[kb][0], ll = cb[kb][1], nn = cb[kb][2];
                                    double s = 0.0;
                                    for (int t = 0; t <= i + j; t++)
                                        for (int u = 0; u <= k + ll; u++)
                                            for (int v = 0; v <= m + nn; v++)
                                                s += Ex[i][j][t] * Ey[k][ll][u] * Ez[m][nn][v] *
                                                     R0[((size_t)t * RD + u) * RD + v];
                                    V[(size_t)(off[sa] + ka) * n + off[sb] + kb] += pref * s;
                                }
                        }
                    }
            }
        free(R0);
    }
    free(off);
    return 0;
}

/* (A|B) over the auxiliary basis, cartesian: out[na][na] */
int ints_two_center(int ns, const double* xyz, const int* l, const int* np, const int* po, const double* e,
                    const double* c, double* out) {
    basis_t bs = mk(ns, xyz, l, np, po, e, c);
    for (int s = 0; s < ns; s++)
        if (l[s] > LMAX) return 1;
    int* off = (int*)malloc(sizeof(int) * (ns + 1));
    cart_offsets(&bs, off);
    int n = off[ns];
#pragma omp parallel
    {
        double* work = (double*)malloc(sizeof(double) * work_doubles());
        double* blk = (double*)malloc(sizeof(double) * MAXC * MAXC);
#pragma omp for schedule(dynamic)
        for (int sa = 0; sa < ns; sa++)
            for (int sb = 0; sb <= sa; sb++) {
                sh_t A = get_shell(&bs, sa), B = get_shell(&bs, sb);
                eri_quartet(A, unit_shell(A), B, unit_shell(B), blk, work);
                int na = NCART(A.l), nb = NCART(B.l);
                for (int i = 0; i < na; i++)
                    for (int j = 0; j < nb; j++) {
                        out[(size_t)(off[sa] + i) * n + off[sb] + j] = blk[i * nb + j];
                        out[(size_t)(off[sb] + j) * n + off[sa] + i] = blk[i * nb + j];
                    }
            }
        free(work);
        free(blk);
    }
    free(off);
    return 0;
}

/* (A|mn) cartesian: out[naux_cart][nbf_cart][nbf_cart], m <= n computed and mirrored (dfhelper.cc:1284-1347) */
int ints_three_center(int nsa, const double* axyz, const int* al, const int* anp, const int* apo, const double* ae,
                      const double* ac, int nsp, const double* pxyz, const int* pl, const int* pnp, const int* ppo,
                      const double* pe, const double* pc, double* out) {
    basis_t ab = mk(nsa, axyz, al, anp, apo, ae, ac), pb = mk(nsp, pxyz, pl, pnp, ppo, pe, pc);
    for (int s = 0; s < nsa; s++)
        if (al[s] > LMAX) return 1;
    for (int s = 0; s < nsp; s++)
        if (pl[s] > LMAX) return 1;
    int* aoff = (int*)malloc(sizeof(int) * (nsa + 1));
    int* poff = (int*)malloc(sizeof(int) * (nsp + 1));
    cart_offsets(&ab, aoff);
    cart_offsets(&pb, poff);
    size_t N = poff[nsp];
#pragma omp parallel
    {
        double* work = (double*)malloc(sizeof(double) * work_doubles());
        double* blk = (double*)malloc(sizeof(double) * MAXC * MAXC * MAXC);
#pragma omp for schedule(dynamic) collapse(2)
        for (int sm = 0; sm < nsp; sm++)
            for (int sq = 0; sq < nsa; sq++) {
                sh_t M = get_shell(&pb, sm), Q = get_shell(&ab, sq);
                int nq = NCART(Q.l), nm = NCART(M.l);
                for (int sn = 0; sn <= sm; sn++) {
                    sh_t Nn = get_shell(&pb, sn);
                    int nn = NCART(Nn.l);
                    eri_quartet(Q, unit_shell(Q), M, Nn, blk, work);
                    for (int q = 0; q < nq; q++)
                        for (int i = 0; i < nm; i++)
                            for (int j = 0; j < nn; j++) {
                                double v = blk[((size_t)q * nm + i) * nn + j];
                                size_t Qg = aoff[sq] + q, mg = poff[sm] + i, ng = poff[sn] + j;
                                out[(Qg * N + mg) * N + ng] = v;
                                out[(Qg * N + ng) * N + mg] = v;
                            }
                }
            }
        free(work);
        free(blk);
    }
    free(aoff);
    free(poff);
    return 0;
}

/* (A|erf(omega r)/r|mn), same layout (the unfitted wPpq_ integrals, dfhelper.cc:683-692) */
int ints_three_center_erf(int nsa, const double* axyz, const int* al, const int* anp, const int* apo, const double* ae,
                          const double* ac, int nsp, const double* pxyz, const int* pl, const int* pnp, const int* ppo,
                          const double* pe, const double* pc, double omega, double* out) {
    if (!(omega > 0.0)) return 2;
    g_omega = omega;
    int rc = ints_three_center(nsa, axyz, al, anp, apo, ae, ac, nsp, pxyz, pl, pnp, ppo, pe, pc, out);
    g_omega = 0.0;
    return rc;
}

/* general erf-attenuated (ab|cd) block (tests) */
int ints_quartet_erf(const double* xyz4, const int* l4, const int* np4, const double* const* e4, const double* const* c4,
                     double omega, double* out);

/* (MU NU | MU NU) cartesian block of one shell pair: out[nm][nn][nm][nn] */
int ints_pair_diagonal(int ns, const double* xyz, const int* l, const int* np, const int* po, const double* e,
                       const double* c, int MU, int NU, double* out) {
    basis_t bs = mk(ns, xyz, l, np, po, e, c);
    if (l[MU] > LMAX || l[NU] > LMAX) return 1;
    double* work = (double*)malloc(sizeof(double) * work_doubles());
    sh_t A = get_shell(&bs, MU), B = get_shell(&bs, NU);
    eri_quartet(A, B, A, B, out, work);
    free(work);
    return 0;
}

/* general (ab|cd) cartesian block between shells of up to four different bases (tests) */
int ints_quartet(const double* xyz4, const int* l4, const int* np4, const double* const* e4, const double* const* c4,
                 double* out) {
    sh_t s[4];
    for (int i = 0; i < 4; i++) {
        if (l4[i] > LMAX) return 1;
        sh_t t = {xyz4[3 * i], xyz4[3 * i + 1], xyz4[3 * i + 2], l4[i], np4[i], e4[i], c4[i]};
        s[i] = t;
    }
    double* work = (double*)malloc(sizeof(double) * work_doubles());
    eri_quartet(s[0], s[1], s[2], s[3], out, work);
    free(work);
    return 0;
}

int ints_quartet_erf(const double* xyz4, const int* l4, const int* np4, const double* const* e4, const double* const* c4,
                     double omega, double* out) {
    if (!(omega > 0.0)) return 2;
    g_omega = omega;
    int rc = ints_quartet(xyz4, l4, np4, e4, c4, out);
    g_omega = 0.0;
    return rc;
}
