/*
 * C restatement of ODINN.jl's SIA2D hot path -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 * PARITY UNPINNED BY STORED NUMBERS (see oracle/__init__.py): pinned by tests/test_oracle_c.py
 * against the NumPy oracle, which is pinned by the reference's identities and FD protocol.
 *
 * Included twice by sia2d_c.c with  REAL / SUF  = double / _f64 and float / _f32.
 * Every function cites the reference file:line (ODINN.jl v1.1.0, commit 31dfbf2) it follows.
 * Arithmetic follows the reference's order of operations (S = B + H, slopes divided by Δ,
 * clamp bounds η₀·H/Δ); only the loops are fused and the temporaries of the Julia code kept
 * per dual node instead of per array.  Layout: Julia column-major, element (i,j) at i + j*nx.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

typedef struct {
    double dx, dy, eta0, n, p, q, rho, g, C, A;
    const REAL* Afield; /* (nx-1) x (ny-1), ld = nx-1, or NULL for the glacier-wide scalar A */
} FN(sia2d_par);

/* Γ (include_A = false), src/models/target/target_utils.jl:3-13 */
static inline REAL FN(gamma_noA)(const FN(sia2d_par) * P) { return (REAL)(2.0 * pow(P->rho * P->g, P->n) / (P->n + 2.0)); }
/* S, src/models/target/target_utils.jl:15-19 */
static inline REAL FN(slide)(const FN(sia2d_par) * P) { return (REAL)(P->C * pow(P->rho * P->g, P->p - P->q)); }

static inline REAL FN(rpow)(REAL x, REAL y) { return (REAL)pow((double)x, (double)y); }
static inline REAL FN(rmax)(REAL a, REAL b) { return a > b ? a : b; }
static inline REAL FN(rmin)(REAL a, REAL b) { return a < b ? a : b; }
/* H = map(x -> ifelse(x > 0.0, x, 0.0), H), src/inverse/SIA2D/adjoint.jl:52 */
static inline REAL FN(pos)(REAL x) { return x > (REAL)0 ? x : (REAL)0; }

/* One dual node (a,b): ∇Sx, ∇Sy, ∇S, H̄ (adjoint.jl:58-67) and D, α, β, ∂A_spatial
 * (src/models/target/target_A.jl:16-30, 32-46, 48-62, 71-72). */
typedef struct { REAL gSx, gSy, Hb, D, alpha, beta, gA; } FN(node);

static inline FN(node) FN(eval_node)(const FN(sia2d_par) * P, int nx, const REAL* H, const REAL* B, int a, int b,
                                     REAL Gam, REAL Sl, int partials) {
    FN(node) r;
    const REAL dx = (REAL)P->dx, dy = (REAL)P->dy;
    REAL h00 = FN(pos)(H[a + b * nx]), h10 = FN(pos)(H[a + 1 + b * nx]);
    REAL h01 = FN(pos)(H[a + (b + 1) * nx]), h11 = FN(pos)(H[a + 1 + (b + 1) * nx]);
    REAL s00 = B[a + b * nx] + h00, s10 = B[a + 1 + b * nx] + h10; /* S = B .+ H, :54 */
    REAL s01 = B[a + (b + 1) * nx] + h01, s11 = B[a + 1 + (b + 1) * nx] + h11;
    REAL dSdx0 = (s10 - s00) / dx, dSdx1 = (s11 - s01) / dx; /* diff_x(S)/Δx, :58 */
    REAL dSdy0 = (s01 - s00) / dy, dSdy1 = (s11 - s10) / dy; /* diff_y(S)/Δy, :59 */
    r.gSx = (REAL)0.5 * (dSdx0 + dSdx1);                     /* avg_y, :60 */
    r.gSy = (REAL)0.5 * (dSdy0 + dSdy1);                     /* avg_x, :61 */
    REAL gS = (REAL)sqrt((double)(r.gSx * r.gSx + r.gSy * r.gSy)); /* :64 */
    r.Hb = (REAL)0.25 * (h00 + h10 + h01 + h11);             /* avg(H), :67 */
    REAL A = P->Afield ? P->Afield[a + b * (nx - 1)] : (REAL)P->A;
    REAL n = (REAL)P->n, p = (REAL)P->p, q = (REAL)P->q;
    if (P->n == 3.0 && P->C == 0.0) {
        REAL H2 = r.Hb * r.Hb, H4 = H2 * H2, g2 = gS * gS;
        r.gA = Gam * H4 * r.Hb * g2;
        r.D = A * r.gA;
        r.alpha = A * Gam * (REAL)5 * H4 * g2;
        r.beta = A * Gam * (REAL)2 * H4 * r.Hb;
    } else {
        r.gA = Gam * FN(rpow)(r.Hb, n + 2) * FN(rpow)(gS, n - 1);
        r.D = Sl * FN(rpow)(r.Hb, p - q + 1) * FN(rpow)(gS, p - 1) + A * r.gA;
        if (partials) {
            r.alpha = (p - q + 1) * Sl * FN(rpow)(r.Hb, p - q) * FN(rpow)(gS, p - 1) +
                      A * Gam * (n + 2) * FN(rpow)(r.Hb, n + 1) * FN(rpow)(gS, n - 1);
            r.beta = Sl * (p - 1) * FN(rpow)(r.Hb, p - q + 1) * FN(rpow)(gS, p - 3) +
                     A * Gam * (n - 1) * FN(rpow)(r.Hb, n + 2) * FN(rpow)(gS, n - 3);
        } else {
            r.alpha = r.beta = 0;
        }
    }
    return r;
}

/* clamp_borders_dx / _dy value for the edge between a lower cell (thickness Hlo) and an upper cell (Hup),
 * src/inverse/SIA2D/inversion_utils.jl:17-20, 31-34. */
static inline REAL FN(clampv)(REAL dS, REAL eta0, REAL Hlo, REAL Hup, REAL d) {
    return FN(rmax)(FN(rmin)(dS, eta0 * Hup / d), -eta0 * Hlo / d);
}

/* F1: Huginn.SIA2D!(dH, H, simulation, t, θ) [NOT IN TREE], restated from adjoint.jl:47-104 and the
 * forward-form twin :522-533, 552-553.  work: (nx-1)*(ny-1) REALs. */
void FN(sia2d_rhs)(int nx, int ny, const REAL* H, const REAL* B, REAL* dH, REAL* work, const FN(sia2d_par) * P) {
    const REAL Gam = FN(gamma_noA)(P), Sl = FN(slide)(P);
    const REAL dx = (REAL)P->dx, dy = (REAL)P->dy, eta0 = (REAL)P->eta0;
    const int mx = nx - 1;
    REAL* D = work;
#pragma omp parallel
    {
#pragma omp for schedule(static)
        for (int b = 0; b < ny - 1; ++b)
            for (int a = 0; a < nx - 1; ++a) D[a + b * mx] = FN(eval_node)(P, nx, H, B, a, b, Gam, Sl, 0).D; /* :79-84 */
#pragma omp for schedule(static)
        for (int j = 0; j < ny; ++j) {
            for (int i = 0; i < nx; ++i) {
                if (i == 0 || j == 0 || i == nx - 1 || j == ny - 1) { dH[i + j * nx] = 0; continue; }
                REAL hc = FN(pos)(H[i + j * nx]), hw = FN(pos)(H[i - 1 + j * nx]), he = FN(pos)(H[i + 1 + j * nx]);
                REAL hs = FN(pos)(H[i + (j - 1) * nx]), hn = FN(pos)(H[i + (j + 1) * nx]);
                REAL sc = B[i + j * nx] + hc, sw = B[i - 1 + j * nx] + hw, se = B[i + 1 + j * nx] + he;
                REAL ss = B[i + (j - 1) * nx] + hs, sn = B[i + (j + 1) * nx] + hn;
                /* dSdx_edges, dSdy_edges (:87-88), clamped (:93-94) */
                REAL cw = FN(clampv)((sc - sw) / dx, eta0, hw, hc, dx), ce = FN(clampv)((se - sc) / dx, eta0, hc, he, dx);
                REAL cs = FN(clampv)((sc - ss) / dy, eta0, hs, hc, dy), cn = FN(clampv)((sn - sc) / dy, eta0, hc, hn, dy);
                REAL d00 = D[i - 1 + (j - 1) * mx], d10 = D[i + (j - 1) * mx], d01 = D[i - 1 + j * mx], d11 = D[i + j * mx];
                REAL Fw = -((REAL)0.5 * (d00 + d01)) * cw, Fe = -((REAL)0.5 * (d10 + d11)) * ce; /* Dx = avg_y(D), :96 */
                REAL Fs = -((REAL)0.5 * (d00 + d10)) * cs, Fn = -((REAL)0.5 * (d01 + d11)) * cn; /* Dy = avg_x(D), :97 */
                dH[i + j * nx] = -((Fe - Fw) / dx + (Fn - Fs) / dy);
            }
        }
    }
}

/* A1 + A2: VJP_λ_∂SIA∂H_discrete (adjoint.jl:31-151) and the glacier-wide contraction of
 * VJP_λ_∂SIA∂θ_discrete (adjoint.jl:178-255 with target_A.jl:64-92):
 *   out  = (∂SIA/∂H)^T λ            (skipped when out == NULL)
 *   *S   = Σ ∂A_spatial ∘ D†        (skipped when S == NULL); vjpA (dual grid, ld nx-1) optional.
 * work: 4*(nx-1)*(ny-1) REALs. */
void FN(sia2d_vjp)(int nx, int ny, const REAL* lam, const REAL* H, const REAL* B, REAL* out, double* S, REAL* vjpA,
                   REAL* work, const FN(sia2d_par) * P) {
    const REAL Gam = FN(gamma_noA)(P), Sl = FN(slide)(P);
    const REAL dx = (REAL)P->dx, dy = (REAL)P->dy, eta0 = (REAL)P->eta0;
    const int mx = nx - 1, my = ny - 1;
    const long nn = (long)mx * my;
    REAL *D = work, *aD = work + nn, *Pn = work + 2 * nn, *Qn = work + 3 * nn;
    double Sacc = 0.0;
#define LAMI(i, j) (((i) >= 1 && (i) <= nx - 2 && (j) >= 1 && (j) <= ny - 2) ? lam[(i) + (j) * nx] : (REAL)0) /* λ_inn, :99 */
#pragma omp parallel
    {
#pragma omp for schedule(static) reduction(+ : Sacc)
        for (int b = 0; b < my; ++b) {
            for (int a = 0; a < mx; ++a) {
                FN(node) nd = FN(eval_node)(P, nx, H, B, a, b, Gam, Sl, 1);
                REAL h00 = FN(pos)(H[a + b * nx]), h10 = FN(pos)(H[a + 1 + b * nx]);
                REAL h01 = FN(pos)(H[a + (b + 1) * nx]), h11 = FN(pos)(H[a + 1 + (b + 1) * nx]);
                REAL s00 = B[a + b * nx] + h00, s10 = B[a + 1 + b * nx] + h10;
                REAL s01 = B[a + (b + 1) * nx] + h01, s11 = B[a + 1 + (b + 1) * nx] + h11;
                REAL l00 = LAMI(a, b), l10 = LAMI(a + 1, b), l01 = LAMI(a, b + 1), l11 = LAMI(a + 1, b + 1);
                /* Fx† = diff_x_adjoint(-λ_inn, Δx), Fy† (:100-101); Dx† = avg_y_adjoint(-Fx† ∘ clamp) (:102-104).
                 * Edges on border rows/columns (not in dSdx_edges) carry λ_inn = 0 at both ends. */
                REAL gx0 = -((l10 - l00) / dx) * FN(clampv)((s10 - s00) / dx, eta0, h00, h10, dx);
                REAL gx1 = -((l11 - l01) / dx) * FN(clampv)((s11 - s01) / dx, eta0, h01, h11, dx);
                REAL gy0 = -((l01 - l00) / dy) * FN(clampv)((s01 - s00) / dy, eta0, h00, h01, dy);
                REAL gy1 = -((l11 - l10) / dy) * FN(clampv)((s11 - s10) / dy, eta0, h10, h11, dy);
                REAL Dadj = (REAL)0.5 * (gx0 + gx1) + (REAL)0.5 * (gy0 + gy1);
                D[a + b * mx] = nd.D;
                aD[a + b * mx] = nd.alpha * Dadj;           /* α .* D_adjoint, :124 */
                Pn[a + b * mx] = nd.beta * nd.gSx * Dadj;   /* βx .* D_adjoint, :125 */
                Qn[a + b * mx] = nd.beta * nd.gSy * Dadj;   /* βy .* D_adjoint, :126 */
                REAL v = nd.gA * Dadj;                      /* ∂A_spatial ∘ D†, adjoint.jl:250 */
                Sacc += (double)v;
                if (vjpA) vjpA[a + b * mx] = v;
            }
        }
        if (out) {
#define ND(arr, a, b) (((a) >= 0 && (a) < mx && (b) >= 0 && (b) < my) ? arr[(a) + (b) * mx] : (REAL)0)
#pragma omp for schedule(static)
            for (int j = 0; j < ny; ++j) {
                for (int i = 0; i < nx; ++i) {
                    REAL hc = FN(pos)(H[i + j * nx]);
                    if (!(hc > 0)) { out[i + j * nx] = 0; continue; } /* dλ .* (H .> 0), :148 */
                    /* first term, :123-127 via the transposes of inversion_utils.jl:3-15, 45-66 */
                    REAL t1 = (REAL)0.25 * (ND(aD, i - 1, j - 1) + ND(aD, i, j - 1) + ND(aD, i - 1, j) + ND(aD, i, j));
                    t1 += ((REAL)0.5 * (ND(Pn, i - 1, j - 1) + ND(Pn, i - 1, j)) - (REAL)0.5 * (ND(Pn, i, j - 1) + ND(Pn, i, j))) / dx;
                    t1 += ((REAL)0.5 * (ND(Qn, i - 1, j - 1) + ND(Qn, i, j - 1)) - (REAL)0.5 * (ND(Qn, i - 1, j) + ND(Qn, i, j))) / dy;
                    /* second term, :130-144 with clamp_borders_d{x,y}_adjoint! (inversion_utils.jl:22-29, 36-43) */
                    REAL t2 = 0;
                    REAL sc = B[i + j * nx] + hc, lc = LAMI(i, j);
                    if (j >= 1 && j <= ny - 2) {
                        if (i >= 1) { /* x-edge (i-1,j): this cell is the upper cell */
                            REAL hw = FN(pos)(H[i - 1 + j * nx]);
                            REAL dS = (sc - (B[i - 1 + j * nx] + hw)) / dx;
                            REAL dC = -((lc - LAMI(i - 1, j)) / dx) * ((REAL)0.5 * (ND(D, i - 1, j - 1) + ND(D, i - 1, j)));
                            if (dS < eta0 * hc / dx && dS > -eta0 * hw / dx) t2 += dC / dx;
                            if (dS > eta0 * hc / dx) t2 += eta0 * dC / dx;
                        }
                        if (i <= nx - 2) { /* x-edge (i,j): lower cell */
                            REAL he = FN(pos)(H[i + 1 + j * nx]);
                            REAL dS = ((B[i + 1 + j * nx] + he) - sc) / dx;
                            REAL dC = -((LAMI(i + 1, j) - lc) / dx) * ((REAL)0.5 * (ND(D, i, j - 1) + ND(D, i, j)));
                            if (dS < eta0 * he / dx && dS > -eta0 * hc / dx) t2 -= dC / dx;
                            if (dS < -eta0 * hc / dx) t2 -= eta0 * dC / dx;
                        }
                    }
                    if (i >= 1 && i <= nx - 2) {
                        if (j >= 1) { /* y-edge (i,j-1): upper cell */
                            REAL hs = FN(pos)(H[i + (j - 1) * nx]);
                            REAL dS = (sc - (B[i + (j - 1) * nx] + hs)) / dy;
                            REAL dC = -((lc - LAMI(i, j - 1)) / dy) * ((REAL)0.5 * (ND(D, i - 1, j - 1) + ND(D, i, j - 1)));
                            if (dS < eta0 * hc / dy && dS > -eta0 * hs / dy) t2 += dC / dy;
                            if (dS > eta0 * hc / dy) t2 += eta0 * dC / dy;
                        }
                        if (j <= ny - 2) { /* y-edge (i,j): lower cell */
                            REAL hn = FN(pos)(H[i + (j + 1) * nx]);
                            REAL dS = ((B[i + (j + 1) * nx] + hn) - sc) / dy;
                            REAL dC = -((LAMI(i, j + 1) - lc) / dy) * ((REAL)0.5 * (ND(D, i - 1, j) + ND(D, i, j)));
                            if (dS < eta0 * hn / dy && dS > -eta0 * hc / dy) t2 -= dC / dy;
                            if (dS < -eta0 * hc / dy) t2 -= eta0 * dC / dy;
                        }
                    }
                    out[i + j * nx] = t1 + t2;
                }
            }
#undef ND
        }
    }
#undef LAMI
    if (S) *S = Sacc;
}

/* Fixed-step forward solve with tstops, saveat = tstops: the loop of oracle/sia2d_numpy.py::solve_forward for
 * method 0 = explicit Euler, 1 = Shu-Osher SSPRK(3,3) (the reference's solve call: src/simulations/inversions/
 * inversion_utils.jl:551-572 with tstops :487-495; the integrator is a user parameter there).  Same operation order as the
 * NumPy loop:  u1 = H + h f(H);  u2 = 0.75 H + 0.25 (u1 + h f(u1));  H = H/3 + (2/3)(u2 + h f(u2)).
 * out: n_t planes of nx*ny (out[0] = H0).  work: (nx-1)*(ny-1) + 3*nx*ny REALs. */
void FN(sia2d_solve_fixed)(int nx, int ny, const REAL* H0, const REAL* B, const FN(sia2d_par) * P, int n_t, const double* t,
                           int nsub, int method, REAL* out, REAL* work) {
    const long N = (long)nx * ny;
    REAL *wk = work, *f = work + (long)(nx - 1) * (ny - 1), *u1 = f + N, *u2 = u1 + N;
    REAL* H = out;
    for (long k = 0; k < N; ++k) H[k] = H0[k];
    for (int j = 1; j < n_t; ++j) {
        REAL* Hn = out + (long)j * N;
        for (long k = 0; k < N; ++k) Hn[k] = H[k];
        H = Hn;
        const REAL h = (REAL)((t[j] - t[j - 1]) / nsub);
        for (int s = 0; s < nsub; ++s) {
            FN(sia2d_rhs)(nx, ny, H, B, f, wk, P);
            if (method == 0) {
#pragma omp parallel for schedule(static)
                for (long k = 0; k < N; ++k) H[k] = H[k] + h * f[k];
                continue;
            }
#pragma omp parallel for schedule(static)
            for (long k = 0; k < N; ++k) u1[k] = H[k] + h * f[k];
            FN(sia2d_rhs)(nx, ny, u1, B, f, wk, P);
#pragma omp parallel for schedule(static)
            for (long k = 0; k < N; ++k) u2[k] = (REAL)0.75 * H[k] + (REAL)0.25 * (u1[k] + h * f[k]);
            FN(sia2d_rhs)(nx, ny, u2, B, f, wk, P);
#pragma omp parallel for schedule(static)
            for (long k = 0; k < N; ++k) H[k] = H[k] / (REAL)3 + ((REAL)2 / (REAL)3) * (u2[k] + h * f[k]);
        }
    }
}

/* DiscreteAdjoint reverse loop of SIA2D_grad_batch! (src/inverse/SIA2D/gradient.jl:191-253) for LossH(L2Sum), glacier-wide A:
 * the loop of oracle/sia2d_numpy.py::loss_and_grad_discrete.  Hs, Href, W: n_t planes (W = is_in_glacier mask / (nx ny),
 * gradient.jl:161; the loss weight of snapshot j is t[j] - t[j-1], 0 for j = 0).  *loss = sum_j dt_j sum W (H_j - Href_j)^2,
 * *Ssum = sum_j dt_{j-1} S(lambda_{j-1}, H_j)  with  d(theta) = (dA/dtheta) * Ssum;  lam: nx*ny (lambda at t_0 on return).
 * work: 4*(nx-1)*(ny-1) + nx*ny REALs. */
void FN(sia2d_grad_discrete)(int nx, int ny, const REAL* B, const FN(sia2d_par) * P, int n_t, const double* t, const REAL* Hs,
                             const REAL* Href, const REAL* W, double* loss, double* Ssum, REAL* lam, REAL* work) {
    const long N = (long)nx * ny;
    REAL* v = work + 4 * (long)(nx - 1) * (ny - 1);
    double ell = 0.0, Sa = 0.0;
    for (long k = 0; k < N; ++k) lam[k] = 0;
    for (int j = n_t - 1; j >= 0; --j) {
        const REAL *Hj = Hs + (long)j * N, *Rj = Href + (long)j * N, *Wj = W + (long)j * N;
        const double dtH = j > 0 ? t[j] - t[j - 1] : 0.0;
        double lj = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : lj)
        for (long k = 0; k < N; ++k) { double d = (double)Hj[k] - (double)Rj[k]; lj += (double)Wj[k] * d * d; } /* Losses.jl:142-152 */
        ell += dtH * lj;                                                                  /* gradient.jl:218-232 */
        if (j == 0) break;
        FN(sia2d_vjp)(nx, ny, lam, Hj, B, v, NULL, NULL, work, P);                        /* :235-237 */
        const REAL dt = (REAL)(t[j] - t[j - 1]), c = (REAL)(2.0 * dtH);
#pragma omp parallel for schedule(static)
        for (long k = 0; k < N; ++k) lam[k] = lam[k] + dt * v[k] + c * Wj[k] * (Hj[k] - Rj[k]);  /* :242, Losses.jl:270-291 */
        double S = 0.0;
        FN(sia2d_vjp)(nx, ny, lam, Hj, B, NULL, &S, NULL, work, P);                       /* :245-246 (the UPDATED lambda) */
        Sa += (double)dt * S;                                                             /* :249 */
    }
    *loss = ell;
    *Ssum = Sa;
}

#undef FN
#undef CAT
#undef CAT_
