// Register-marching F1 / A1+A2 kernels.
//
// The first (shared-memory tiled) kernels were instruction-issue bound (ncu, profiles/r01_v1_*: 254 / 427
// thread-instructions per cell, 84 % / 70 % issue-slot utilisation, 15 % / 11 % DRAM).  Here every dual node
// and every edge is evaluated exactly once:
//   * one warp owns a strip of 32 consecutive columns (lane = column, the contiguous axis, so every row access
//     is one coalesced line) and marches over a chunk of rows;
//   * the 3x3 neighbourhood lives in registers: the previous row is carried, x-neighbours come from
//     warp shuffles (halo exchange), so there is no shared memory and no block barrier at all;
//   * lanes 0 and 31 are halo lanes: a strip produces 30 output columns;
//   * all 1/Δ, 1/2 and 1/4 factors are folded into per-warp constants ("raw" quantities below are sums of
//     differences that have not been scaled yet).
// Work items (glacier, first column, row range) come from a table built at ensemble creation, so ragged
// ensembles are one launch.  Reference semantics and citations: see sia2d_kernels.cuh.
#pragma once
#include "sia2d_kernels.cuh"

namespace odinn {

constexpr int STRIP = 30;        // output columns per warp
#ifndef ODINN_MARCH_WARPS
#define ODINN_MARCH_WARPS 4   // (sweep profiles/r01_v6_sweep.txt: fp64 F1 0.340 -> 0.298 ms against 8 warps)
#endif
constexpr int MARCH_WARPS = ODINN_MARCH_WARPS;   // warps per CTA
#ifndef ODINN_VJP1_MIN_CTAS
#define ODINN_VJP1_MIN_CTAS 4   // resident CTAs per SM the A1+A2 kernel is compiled for: 128 registers instead of 142,
                                // 16 resident warps instead of 8 (fp64 A1+A2 0.938 -> 0.694 ms)
#endif
constexpr unsigned FULL = 0xffffffffu;
// L2 prefetch distance in rows ahead of the register prefetch queue (0 = off); see sia2d_march2.cuh / profiles/r01_v4_sweep.txt
#ifndef ODINN_L2PF_ROWS1
#define ODINN_L2PF_ROWS1 8
#endif
__device__ __forceinline__ void prefetch_l2_row(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Rows of register prefetch (ncu: the 1-row version stalled on long_scoreboard).  Deeper queues cost
// registers, i.e. resident warps: F1 is light (4 rows), the VJP kernel is not.
#ifndef ODINN_PF_RHS
#define ODINN_PF_RHS 4
#endif
#ifndef ODINN_PF_VJP32
#define ODINN_PF_VJP32 2
#endif
#ifndef ODINN_PF_VJP64
#define ODINN_PF_VJP64 2
#endif
template <typename T> struct PfVjp { static constexpr int value = ODINN_PF_VJP32; };
template <> struct PfVjp<double> { static constexpr int value = ODINN_PF_VJP64; };

template <typename T>
__device__ __forceinline__ T shfl_dn(T v) { return __shfl_down_sync(FULL, v, 1); }
template <typename T>
__device__ __forceinline__ T shfl_up(T v) { return __shfl_up_sync(FULL, v, 1); }

__device__ __forceinline__ float fmx(float a, float b) { return fmaxf(a, b); }
// fp64: compare + select (DSETP, 2 x SEL) instead of fmax / fmin, whose NaN handling costs DSETP.MAX + 2 MOV + FSEL + SEL + LOP3 per call
// on sm_100 (nine calls per row of the A1 step); the operands are finite, where both forms return the same value.
__device__ __forceinline__ double fmx(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ float fmn(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double fmn(double a, double b) { return a < b ? a : b; }

// Node diffusivity from RAW sums: Hs = Σ of the 4 thicknesses (H̄ = Hs/4), g2 = |∇S|².
// Returns D and, when PARTIALS, α = ∂D/∂H̄, β = (1/∇S)∂D/∂∇S, gA = Γ H̄^{n+2} ∇S^{n-1} (target_A.jl:16-92).
template <typename T, bool CUBIC, bool PARTIALS>
__device__ __forceinline__ void node_raw(const PhysDev<T>& ph, T A, T Hs, T g2, T& D, T& alpha, T& beta, T& gA) {
    if (CUBIC) {
        const T K = ph.Gam * T(1.0 / 1024.0);  // Γ/4^5
        T H2 = Hs * Hs;
        T H4 = H2 * H2;
        T w = H4 * Hs;
        if (!PARTIALS) {
            D = ((A * K) * g2) * w;
            return;
        }
        T tg = K * g2;
        gA = tg * w;
        D = A * gA;
        if (PARTIALS) {
            alpha = (T(20) * A) * (tg * H4);  // 5 A Γ H̄^4 ∇S^2
            beta = (T(2) * A * K) * w;        // 2 A Γ H̄^5
        }
    } else {
        node_diffusivity<T, false, PARTIALS>(ph, A, T(0.25) * Hs, g2, D, alpha, beta, gA);
    }
}

// Sub-gradient of the flux clamp  c = max(min(e, up), lo)  pushed onto the two cells of an edge
// (clamp_borders_d{x,y}_adjoint!, inversion_utils.jl:22-29 / 36-43; strict inequalities, ties give zero):
//   strictly inside      : ∂dS = dC            -> lower cell gets -dC,      upper cell +dC
//   e < lo (= -η₀H_lo)   : ∂H_lo = -η₀ dC      -> lower cell gets -η₀ dC
//   e > up (=  η₀H_up)   : ∂H_up = +η₀ dC      -> upper cell gets +η₀ dC
// dC here already carries the 1/Δ of diff_adjoint.  With η₀ = 1 (ETA1) the cases merge:
// lower gets -dC unless (e >= up or e == lo), upper gets +dC unless (e <= lo or e == up).
template <typename T, bool ETA1>
__device__ __forceinline__ void subgrad_cmp(T dC, bool gt_lo, bool lt_lo, bool lt_up, bool gt_up, T eta0, T& to_lower, T& to_upper) {
    if (ETA1) {
        to_lower = (lt_up && (gt_lo || lt_lo)) ? -dC : T(0);
        to_upper = (gt_lo && (lt_up || gt_up)) ? dC : T(0);
    } else {
        T pass = (lt_up && gt_lo) ? dC : T(0);
        T edC = eta0 * dC;
        to_lower = -pass - (lt_lo ? edC : T(0));
        to_upper = pass + (gt_up ? edC : T(0));
    }
}
// fp32: plain comparisons of the raw differences.
template <typename T, bool ETA1>
__device__ __forceinline__ void subgrad(float dC, float e, float lo, float up, float delta, float eta0, float& to_lower,
                                        float& to_upper) {
    (void)delta;
    subgrad_cmp<float, ETA1>(dC, e > lo, lo > e, up > e, e > up, eta0, to_lower, to_upper);
}
// fp64: the reference compares the DIVIDED quantities dS/Δ and ±η₀H/Δ (inversion_utils.jl:24-28, 38-42); two raw
// values an ulp apart can round to the same quotient, which turns a strict inequality into a tie (see gt_div).
// The sign of a floating-point difference is exact, so d1 = e - lo and d2 = up - e classify the edge with two DADDs;
// the true divisions are needed only when a difference is below 1e-14·|e| -- a rare slow path taken warp-uniformly.
// (Ice-free flat regions, e = lo = up = 0, stay on the fast path: tol = 0.)
template <typename T, bool ETA1>
__device__ __forceinline__ void subgrad(double dC, double e, double lo, double up, double delta, double eta0,
                                        double& to_lower, double& to_upper) {
    const double d1 = e - lo, d2 = up - e;
    bool gt_lo = d1 > 0.0, lt_lo = d1 < 0.0, lt_up = d2 > 0.0, gt_up = d2 < 0.0;
    // near-tie with a bound implies |bound| ~ |e|, so the test is relative to |e| alone; e == 0 gives tol == 0 (never near)
    const double tol = 1e-14 * fabs(e);
    const bool near = (fabs(d1) < tol) || (fabs(d2) < tol);
    // warp-uniform branch: keeps the six divisions out of the common path (a per-lane `if` gets if-converted and the
    // divisions then run on every step -- profiles/r01_v5: 48 DFMA + 6 MUFU per step)
    if (__any_sync(FULL, near)) {
        if (near) {
            const double qe = e / delta, ql = lo / delta, qu = up / delta;
            gt_lo = qe > ql; lt_lo = ql > qe; lt_up = qu > qe; gt_up = qe > qu;
        }
    }
    subgrad_cmp<double, ETA1>(dC, gt_lo, lt_lo, lt_up, gt_up, eta0, to_lower, to_upper);
}

#ifndef ODINN_SUBGRAD_XY
#define ODINN_SUBGRAD_XY 1
#endif
// Both edges of a marching step at once (cubic-form steps).  fp64: ONE warp vote covers the near-tie slow paths of the y- and the
// x-edge, and the near-tie test itself -- |d| < 2^-46 |e| (1.4e-14 |e|; any threshold far above the 2^-52 at which a quotient can
// round onto its neighbour serves) -- compares the high words on the integer pipe instead of a DMUL, a DSETP.MIN chain and a DSETP on
// the FP64 pipe, which this kernel saturates.  d == 0 with e != 0 (an exact tie) is `near` as before; e == 0 never is.
template <typename T, bool ETA1>
__device__ __forceinline__ void subgrad_xy(float dCy, float ey, float loy, float upy, float dly, float dCx, float ex, float lox, float upx,
                                           float dlx, float eta0, float& yl, float& yu, float& xl, float& xu) {
    subgrad<T, ETA1>(dCy, ey, loy, upy, dly, eta0, yl, yu);
    subgrad<T, ETA1>(dCx, ex, lox, upx, dlx, eta0, xl, xu);
}
__device__ __forceinline__ int hi_abs(double v) { return __double2hiint(v) & 0x7fffffff; }
template <typename T, bool ETA1>
__device__ __forceinline__ void subgrad_xy(double dCy, double ey, double loy, double upy, double dly, double dCx, double ex, double lox,
                                           double upx, double dlx, double eta0, double& yl, double& yu, double& xl, double& xu) {
    const double y1 = ey - loy, y2 = upy - ey, x1 = ex - lox, x2 = upx - ex;
    bool y_gl = y1 > 0.0, y_ll = y1 < 0.0, y_lu = y2 > 0.0, y_gu = y2 < 0.0;
    bool x_gl = x1 > 0.0, x_ll = x1 < 0.0, x_lu = x2 > 0.0, x_gu = x2 < 0.0;
    const int ty = hi_abs(ey) - (46 << 20), tx = hi_abs(ex) - (46 << 20);
    const bool near_y = hi_abs(y1) < ty || hi_abs(y2) < ty;
    const bool near_x = hi_abs(x1) < tx || hi_abs(x2) < tx;
    if (__any_sync(FULL, near_y || near_x)) {   // warp-uniform: keeps the divisions out of the common path
        if (near_y) {
            const double qe = ey / dly, ql = loy / dly, qu = upy / dly;
            y_gl = qe > ql; y_ll = ql > qe; y_lu = qu > qe; y_gu = qe > qu;
        }
        if (near_x) {
            const double qe = ex / dlx, ql = lox / dlx, qu = upx / dlx;
            x_gl = qe > ql; x_ll = ql > qe; x_lu = qu > qe; x_gu = qe > qu;
        }
    }
    subgrad_cmp<double, ETA1>(dCy, y_gl, y_ll, y_lu, y_gu, eta0, yl, yu);
    subgrad_cmp<double, ETA1>(dCx, x_gl, x_ll, x_lu, x_gu, eta0, xl, xu);
}

// --------------------------------------------------------------------------------------------
// F1.  One marching step consumes cell row `row+1` and produces dual-node row `row`, the y-edges
// row→row+1, the x-edges of row `row` and (OUT) the output row `row`.  MASKED steps carry the
// row-boundary logic (clamped row pointer, zero border rows); the main loop runs unmasked steps only.
// --------------------------------------------------------------------------------------------
// STAGE fuses a low-storage Runge-Kutta stage into the epilogue:  out = sa·U0 + sb·(H + sdt·SIA2D(H))
// (removes the separate axpy passes of the time loop: 4 words/cell instead of 3 + 4).
// DFIELD (with AFIELD): the node plane holds the diffusivity D itself (per-cell laws, sia2d_law.cuh) instead of A.
// RK > 0: an RDPK3Sp35 stage as the epilogue (RkFuse, common.cuh; compile-time mode RKM_*) instead of the two-register stage of STAGE.
// Straight-line: every lane loads and computes (clamped column index), only the stores and the norm accumulation are predicated.
template <typename T, bool CUBIC, bool AFIELD, bool ETA1, bool STAGE, bool DFIELD = false, int RK = 0>
struct RhsMarch {
    static constexpr int PF = ODINN_PF_RHS;
    static constexpr bool RK_FIRST = RK == RKM_FIRST, RK_U = RK == RKM_MID_U || RK == RKM_LAST, RK_WR = RK == RKM_MID || RK == RKM_MID_U,
                          RK_NORM = RK == RKM_LAST;
    // per-warp / per-lane constants
    const T *hp, *bp, *ap, *up;
    T* op;
    // RK: the stage's planes (one element offset `oo` of the output row serves them all), the glacier's (b h, e h), the operands of
    // the output row and of the next one (loaded one step ahead), the error-norm accumulator
    RkFuse<T> rk;
    long long oo;
    T rbh, reh, r_s2, r_e, r_u, n_s2, n_e, n_u;
    double nacc;

    // operation order of rk_stage / rk_stage1_main in rdpk.cu; k = SIA2D(S1) at the cell
    __device__ __forceinline__ void rk_epilogue(T k) {
        const T s1 = hraw;
        T s1n, er;
        if (RK_FIRST) {
            s1n = s1 + rbh * k;
            er = reh * k;
        } else {
            const T s2 = r_s2 + rk.d * s1;
            T v = rk.g1 * s1 + rk.g2 * s2;
            if (RK_U) v = v + rk.g3 * r_u;
            s1n = v + rbh * k;
            er = r_e + reh * k;
            if (RK_WR) { if (store_lane) rk.S2out[oo] = s2; }
        }
        if (store_lane) *op = s1n;
        if (RK_FIRST || RK_WR) { if (store_lane) rk.est[oo] = er; }
        if (RK_NORM) {
            // er / den through a float-seeded reciprocal + two Newton steps (relative error ~1e-16; a true DDIV with its slow path made
            // this launch 1.03 ms against 0.67 ms for the other stages at the bench workload); den >= abstol > 0 is far inside the float range
            const double m = fmax(fabs((double)r_u), fabs((double)s1n));
            const double den = (double)rk.abstol + (double)rk.reltol * m;
            double x = (double)__frcp_rn((float)den);
            x = fma(x, fma(-den, x, 1.0), x);
            x = fma(x, fma(-den, x, 1.0), x);
            const double r = (double)er * x;
            nacc += store_lane ? r * r : 0.0;
        }
    }
    int ld, nym1, ny2;
    T eta0, hdx, hdy, kx, ky, A;  // kx, ky are zeroed on border columns
    T sa, sb, sdt, hraw;
    bool store_lane;
    PhysDev<T> ph;
    // carried row state
    T h, b, eh, ex, hx, ehE, Dp, Fy;
    T hq[PF], bq[PF];  // prefetched cell rows row+1 .. row+PF

    template <bool OUT, bool MASKED>
    __device__ __forceinline__ void step(int row) {
        T h1 = hq[0], b1 = bq[0];
#pragma unroll
        for (int k = 0; k + 1 < PF; ++k) { hq[k] = hq[k + 1]; bq[k] = bq[k + 1]; }
        if (MASKED) {
            int stp = (row + 1 + PF <= nym1) ? ld : 0;
            hp += stp;
            bp += stp;
        } else {
            hp += ld;
            bp += ld;
        }
        hq[PF - 1] = __ldg(hp);
        bq[PF - 1] = __ldg(bp);
        if (ODINN_L2PF_ROWS1 > 0 && !MASKED) {
            if (row + 1 + PF + ODINN_L2PF_ROWS1 <= nym1) {
                prefetch_l2_row(hp + (long long)ODINN_L2PF_ROWS1 * ld);
                prefetch_l2_row(bp + (long long)ODINN_L2PF_ROWS1 * ld);
                if (STAGE) prefetch_l2_row(up + (long long)ODINN_L2PF_ROWS1 * ld);  // U0 has no register queue (see sia2d_rhs_march2)
                if (RK) {   // (rows ahead of the OUTPUT row here: the epilogue planes are read PF + 1 rows behind the input rows)
                    const long long oa = oo + (long long)(ODINN_L2PF_ROWS1 + PF + 1) * ld;
                    if (!RK_FIRST) { prefetch_l2_row(rk.S2in + oa); prefetch_l2_row(rk.est + oa); }
                    if (RK_U && rk.u != rk.S2in) prefetch_l2_row(rk.u + oa);
                }
            }
        }
        T u0 = T(0);
        if (STAGE && OUT) { if (store_lane) u0 = __ldg(up); }
        if (RK) {   // operands of the NEXT output row (plain loads: S2 and est are rewritten in place by their owner)
            const long long on = oo + (MASKED ? ((row + 1 <= nym1) ? ld : 0) : ld);
            if (!RK_FIRST) { r_s2 = n_s2; r_e = n_e; n_s2 = rk.S2in[on]; n_e = rk.est[on]; }
            if (RK_U) { r_u = n_u; n_u = rk.u[on]; }
        }
        const T hraw1 = h1;
        h1 = fmx(h1, T(0));                // adjoint.jl:52
        b1 = surf_store<T>(b1, h1);
        T eh1 = ETA1 ? h1 : eta0 * h1;
        T hE1 = shfl_dn(h1), bE1 = shfl_dn(b1);
        T ex1 = sdiff<T>(bE1, b1, hE1, h1);  // raw x-edge difference S[i+1]-S[i], row+1
        T hx1 = h1 + hE1;
        T ehE1 = ETA1 ? hE1 : eta0 * hE1;
        T ey = sdiff<T>(b1, b, h1, h);       // raw y-edge difference S[j+1]-S[j]
        T eyE = shfl_dn(ey);
        // node (i, row): ∇Sx = ½(ex+ex1)/Δx, ∇Sy = ½(ey+eyE)/Δy, H̄ = ¼(hx+hx1)   (adjoint.jl:58-67)
        T u = (ex + ex1) * hdx, v = (ey + eyE) * hdy;
        T g2 = u * u + v * v;
        T Anode = A;
        if (AFIELD) {
            Anode = __ldg(ap);
            if (MASKED) { if (row >= 0 && row < ny2) ap += ld; } else ap += ld;
        }
        T D1, al, be, gA;
        if (DFIELD) D1 = Anode;
        else node_raw<T, CUBIC, false>(ph, Anode, hx + hx1, g2, D1, al, be, gA);
        T D1W = shfl_up(D1);
        // raw fluxes: y-edge (i, row→row+1) and x-edge (i→i+1, row)          (adjoint.jl:93-97)
        T Fy1 = (D1W + D1) * fmx(fmn(ey, eh1), -eh);
        T Fx = (Dp + D1) * fmx(fmn(ex, ehE), -eh);
        T FxW = shfl_up(Fx);
        if (OUT) {
            // dH = -(∂x Fx + ∂y Fy), F = -½(D+D)·clamp/Δ  ⇒  dH = ½/Δx² ΔFx_raw + ½/Δy² ΔFy_raw
            T outv = kx * (Fx - FxW) + ky * (Fy1 - Fy);
            if (MASKED) { if (row < 1 || row >= nym1) outv = T(0); }
            if (STAGE) outv = sa * u0 + sb * (hraw + sdt * outv);
            if (RK) rk_epilogue(outv);
            else if (store_lane) *op = outv;
        }
        op += ld;
        if (STAGE) { up += ld; hraw = hraw1; }
        if (RK) { oo += ld; hraw = hraw1; }
        h = h1; b = b1; eh = eh1; ex = ex1; hx = hx1; ehE = ehE1; Dp = D1; Fy = Fy1;
    }
};

template <typename T, bool CUBIC, bool AFIELD, bool ETA1, bool STAGE, bool DFIELD = false, int RK = 0>
__global__ void __launch_bounds__(MARCH_WARPS * 32)
sia2d_rhs_march(const GDesc<T>* __restrict__ descs, const int4* __restrict__ items, int n_items,
                const T* __restrict__ H, const T* __restrict__ B, const T* __restrict__ Af, T* dH,
                PhysDev<T> ph, const T* U0, T sa, T sb, T sdt, T A_ovr = T(0), int use_A_ovr = 0,
                const double* __restrict__ stage_tab = nullptr, const int* __restrict__ interval = nullptr,
                RkFuse<T> rkf = RkFuse<T>(), double* __restrict__ partial = nullptr) {
    // Replayed from a CUDA graph (odinn_solve_forward): the stage coefficients of interval *interval come from a device table
    // (9 doubles per interval, stage_tab already offset to this launch's stage), so one captured graph serves every interval.
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<T> d = descs[it.x];
    const int i = it.y + lane, r0 = it.z, r1 = it.w;
    const int ic = min(max(i, 0), d.nx - 1);
    const bool col_inner = (i >= 1 && i <= d.nx - 2);
    RhsMarch<T, CUBIC, AFIELD, ETA1, STAGE, DFIELD, RK> m;
    constexpr int PF = ODINN_PF_RHS;
    // Everything above reads tables that no F1 kernel writes (a kernel that does write them -- set_A_kernel -- never triggers early, so
    // it has completed before this prologue starts); from here on the launch depends on the previous kernel of the stream.
    pdl_wait();
    if (RK) { if (rkf.st[it.x].done) return; }   // the glacier has landed on the tstop: nothing to integrate (rk_integrate commits by copy then)
    if (STAGE && stage_tab != nullptr) {
        const double* sp = stage_tab + (long long)(*interval) * 9;
        sa = (T)sp[0]; sb = (T)sp[1]; sdt = (T)sp[2];
    }
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.eta0 = ph.eta0;
    m.hdx = T(0.5) * d.inv_dx;
    m.hdy = T(0.5) * d.inv_dy;
    m.kx = col_inner ? m.hdx * d.inv_dx : T(0);  // ½/Δx²
    m.ky = col_inner ? m.hdy * d.inv_dy : T(0);
    m.A = use_A_ovr ? A_ovr : d.A;  // (A ≡ 1: the ∂A_spatial flux of the continuous θ-VJP, sia2d_cont.cuh)
    m.store_lane = (lane >= 1 && lane <= STRIP && i < d.nx);
    const int rc = max(r0 - 1, 0);
    m.hp = H + d.off + ic + (long long)rc * d.ld;
    m.bp = B + d.off + ic + (long long)rc * d.ld;
    m.ap = AFIELD ? Af + d.off + min(ic, d.nx - 2) + (long long)min(rc, d.ny - 2) * d.ld : nullptr;
    m.op = dH + d.off + ic + (long long)(r0 - 1) * d.ld;  // dereferenced for rows >= r0 only
    m.up = STAGE ? U0 + d.off + ic + (long long)(r0 - 1) * d.ld : nullptr;
    m.sa = sa;
    m.sb = sb;
    m.sdt = sdt;
    m.hraw = T(0);
    if (RK) {
        m.rk = rkf;
        const double hh = rkf.st[it.x].h;
        m.rbh = (T)(rkf.b * hh);
        m.reh = (T)(rkf.e * hh);
        m.oo = d.off + ic + (long long)(r0 - 1) * d.ld;
        m.r_s2 = m.r_e = m.r_u = m.n_s2 = m.n_e = m.n_u = T(0);
        m.nacc = 0.0;
    }

    // ---- cell row r0-1 (rows outside the grid are clamped: they only feed masked quantities) ----
    m.h = fmx(__ldg(m.hp), T(0));
    m.b = surf_store<T>(__ldg(m.bp), m.h);
    m.eh = ETA1 ? m.h : m.eta0 * m.h;
    {
        T hE = shfl_dn(m.h), bE = shfl_dn(m.b);
        m.ex = sdiff<T>(bE, m.b, hE, m.h);
        m.hx = m.h + hE;
        m.ehE = ETA1 ? hE : m.eta0 * hE;
    }
    m.Dp = T(0);
    m.Fy = T(0);
#pragma unroll
    for (int k = 0; k < PF; ++k) {  // rows r0 .. r0+PF-1 (clamped)
        if (r0 + k >= 1 && r0 + k <= m.nym1) { m.hp += d.ld; m.bp += d.ld; }
        m.hq[k] = __ldg(m.hp);
        m.bq[k] = __ldg(m.bp);
    }

    int row = r0 - 1;
    m.template step<false, true>(row);  // warm-up: node row r0-1, no output
    ++row;
    const int main_end = min(r1, d.ny - 1 - PF);
    for (; row < min(r1, 1); ++row) m.template step<true, true>(row);
#pragma unroll 4
    for (; row < main_end; ++row) m.template step<true, false>(row);
    for (; row < r1; ++row) m.template step<true, true>(row);
    if (RK) {
        if (RK == RKM_LAST) {   // (only storing lanes accumulated)
            double a = m.nacc;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
            if (lane == 0) partial[item] = a;
        }
    }
}

// --------------------------------------------------------------------------------------------
// A1 + A2 (math: see sia2d_vjp_kernel; same arithmetic, marched, constants folded)
// --------------------------------------------------------------------------------------------
// DFIELD (with AFIELD): the node planes hold D, α = ∂D/∂H̄ and β (as the target's ∂Diffusivity∂∇H returns it) of a
// per-cell law; the θ-integrand plane then receives D† itself (gA ≡ 1) for the law's own pullback (sia2d_law.cuh).
// WRITE_F (cubic form only): the same pass also writes dH = SIA2D(H) -- see VjpMarch2 in sia2d_march2.cuh.
// RKA (cubic form, WRITE_H only): one RDPK3Sp35 stage of the continuous adjoint's reverse ODE in one pass -- H is the interpolant
// la0 Ha + la1 Hb of two forward snapshots formed when a row leaves the prefetch queues, the stage update of RkFuse (common.cuh) is the
// epilogue (see VjpMarch2 in sia2d_march2.cuh).
template <typename T, bool CUBIC, bool AFIELD, bool WRITE_H, bool WRITE_S, bool ETA1, bool DFIELD = false, bool WRITE_F = false, bool RKA = false>
struct VjpMarch {
    static constexpr int PF = PfVjp<T>::value;
    static constexpr bool CUBIC_FORM = CUBIC && !DFIELD;  // n = 3, C = 0: step_cubic()
    // RKA state: second snapshot row pointer + queue, interpolation weights, the stage, the glacier's (b h, e h), the offset of the
    // output row in the stage's planes, its operands (raw S1, S2, est, u), the error-norm accumulator, two more L2-prefetch pointers
    const T* h2p;
    T hq2[PfVjp<T>::value];
    T la0, la1;
    RkFuse<T> rk;
    const T* s1b;
    long long oo;
    T rbh, reh, r_s1, r_s2, r_e, r_u;
    double nacc;
    const T *pf2p, *pf3p;

    // res = k = (dSIA/dH)^T S1 at H_itp: the stage update of rk_stage / rk_stage1_main (rdpk.cu), straight-line
    __device__ __forceinline__ void rk_emit(T k) {
        const T s1 = r_s1;
        T s1n, er;
        if (rk.flags & RKF_FIRST) {
            s1n = s1 + rbh * k;
            er = reh * k;
        } else {
            const T s2 = r_s2 + rk.d * s1;
            T v = rk.g1 * s1 + rk.g2 * s2;
            if (rk.flags & RKF_U) v = v + rk.g3 * r_u;
            s1n = v + rbh * k;
            er = r_e + reh * k;
            if (rk.flags & RKF_WS2) { if (store_lane) rk.S2out[oo] = s2; }
        }
        if (store_lane) *op = s1n;
        if (rk.flags & RKF_WEST) { if (store_lane) rk.est[oo] = er; }
        if (rk.flags & RKF_NORM) {
            const double m = fmax(fabs((double)r_u), fabs((double)s1n));
            const double den = (double)rk.abstol + (double)rk.reltol * m;
            double x = (double)__frcp_rn((float)den);   // float-seeded reciprocal + two Newton steps (see RhsMarch::rk_epilogue)
            x = fma(x, fma(-den, x, 1.0), x);
            x = fma(x, fma(-den, x, 1.0), x);
            const double r = (double)er * x;
            nacc += store_lane ? r * r : 0.0;
        }
    }
    const T *hp, *bp, *lp, *ap, *alp, *bep;
    const T* pfp;  // CUBIC_FORM: this lane's L2-prefetch sector (10 lanes per input plane), ODINN_L2PF_ROWS1 rows ahead
    T *op, *vp;  // output row pointer; gridded-A integrand pointer (or null)
    T* fp;       // WRITE_F: dH row pointer
    // CUBIC_FORM constants and carried values (same scheme as VjpMarch2::compute_cubic)
    T hdxs, hdys, lmx, lmy, qx2, qy2, nodem;
    T ly, fx, Qp, cx, Fyp;
    int ld, nym1, ny2, r0;
    T eta0, hdx, hdy, nhx2, nhy2, qx, qy, A, dx, dy;
    T lmask;     // 1 on inner columns, 0 on border columns (λ_inn zero-extension)
    bool store_lane, node_col_ok, own_lane;
    PhysDev<T> ph;
    // carried row state
    T h, b, l, eh, ex, hx, ehE, fxr, px, Dp, aDp, Pp, Qrow_p, yu_p, acc;
    T hq[PF], bq[PF], lq[PF];

    // n = 3, C = 0 form of the step (scalar A or gridded A): shared node products, λ scaled once when loaded, the Q term in the
    // west-going message, one L2 prefetch per row for all planes, optional F1 output.  Same operator as step().
    template <bool OUT, bool MASKED>
    __device__ __forceinline__ void step_cubic(int row) {
        T h1 = hq[0], b1 = bq[0], l1 = lq[0];
        if (RKA) h1 = la0 * h1 + la1 * hq2[0];               // H_itp = (1 - a) Ha + a Hb   (rk_lerp of rdpk.cu)
#pragma unroll
        for (int k = 0; k + 1 < PF; ++k) { hq[k] = hq[k + 1]; bq[k] = bq[k + 1]; lq[k] = lq[k + 1]; if (RKA) hq2[k] = hq2[k + 1]; }
        T l1x = l1 * lmx, l1y = l1 * lmy;
        if (MASKED) {
            int stp = (row + 1 + PF <= nym1) ? ld : 0;
            hp += stp;
            bp += stp;
            lp += stp;
            pfp += stp;
            if (RKA) { h2p += stp; pf2p += stp; if (pf3p) pf3p += stp; }
            if (!(row >= 0 && row + 1 < nym1)) l1x = l1y = T(0);  // λ_inn zero-extended on border rows
        } else {
            hp += ld;
            bp += ld;
            lp += ld;
            pfp += ld;
            if (RKA) { h2p += ld; pf2p += ld; if (pf3p) pf3p += ld; }
        }
        hq[PF - 1] = __ldg(hp);
        if (RKA) hq2[PF - 1] = __ldg(h2p);
        bq[PF - 1] = __ldg(bp);
        lq[PF - 1] = __ldg(lp);
        if (ODINN_L2PF_ROWS1 > 0 && !MASKED) {
            if (row + 1 + PF + ODINN_L2PF_ROWS1 <= nym1) {
                prefetch_l2_row(pfp);
                if (RKA) { prefetch_l2_row(pf2p); if (pf3p) prefetch_l2_row(pf3p); }
            }
        }
        if (RKA && WRITE_H && OUT) {   // operands of the row this step emits (plain loads: S2 and est are rewritten in place by this thread)
            r_s1 = s1b[oo];
            if (!(rk.flags & RKF_FIRST)) { r_s2 = rk.S2in[oo]; r_e = rk.est[oo]; }
            if (rk.flags & (RKF_U | RKF_NORM)) r_u = rk.u[oo];
        }
        h1 = fmx(h1, T(0));
        b1 = surf_store<T>(b1, h1);
        T eh1 = ETA1 ? h1 : eta0 * h1;
        T hE1 = shfl_dn(h1), bE1 = shfl_dn(b1), lE1x = shfl_dn(l1x);
        // x-edge (i→i+1, row+1)
        T ex1 = sdiff<T>(bE1, b1, hE1, h1);
        T hx1 = h1 + hE1;
        T ehE1 = ETA1 ? hE1 : eta0 * hE1;
        T fx1 = lE1x - l1x;                                  // -½/Δx² · Fx†   (adjoint.jl:100)
        const T cx1 = fmx(fmn(ex1, ehE1), -eh1);
        T px1 = fx1 * cx1;                                   // (adjoint.jl:102)
        // y-edge (i, row→row+1)
        T ey = sdiff<T>(b1, b, h1, h);
        T fy = l1y - ly;
        const T cy = fmx(fmn(ey, eh1), -eh);
        T py = fy * cy;
        T eyE = shfl_dn(ey), pyE = shfl_dn(py);
        // node (i, row)
        T gxr = ex + ex1, gyr = ey + eyE;
        T u = gxr * hdxs, v = gyr * hdys;                    // scaled by √K: u² + v² = K |∇S|²
        T g2 = u * u + v * v;
        T Anode = A;
        if (AFIELD) {
            Anode = __ldg(ap);
            bool adv = true;
            if (MASKED) adv = (row >= 0 && row < ny2);
            if (adv) ap += ld;
        }
        T Hs = hx + hx1;
        T H2 = Hs * Hs;
        T H4 = H2 * H2;
        T mk = g2 * Hs;
        T D1 = H4 * (mk * Anode);
        T Dadj = ((py + pyE) + (px + px1)) * nodem;          // D† (adjoint.jl:102-104), zero outside the dual grid
        bool node_ok = node_col_ok;
        if (MASKED) { node_ok = node_ok && row >= 0 && row < nym1; if (!(row >= 0 && row < nym1)) Dadj = T(0); }
        T z = H4 * Dadj;
        T zA = z * Anode;
        T aD1 = g2 * zA;                                     // α D† / 5
        T bD = Hs * zA;                                      // β D† / (2K)
        T P1 = bD * gxr;
        T Q1 = bD * gyr;
        if (WRITE_S) {
            if (OUT) {
                T vS = mk * z;                               // ∂A_spatial ∘ D† (adjoint.jl:250)
                acc += vS;                                   // (halo lanes are dropped when the strip is reduced)
                if (AFIELD) { if (own_lane && node_ok) *vp = vS; }
            }
            if (AFIELD) vp += ld;
        }
        T D1W = T(0);
        if (WRITE_H || WRITE_F) D1W = shfl_up(D1);
        if (WRITE_F) {
            T Fy1 = (D1W + D1) * cy;
            T Fx = (Dp + D1) * cx;
            T FxW = shfl_up(Fx);
            if (OUT) {
                T outv = lmy * (Fyp - Fy1) + lmx * (FxW - Fx);
                if (MASKED) { if (row < 1 || row >= nym1) outv = T(0); }
                if (store_lane) *fp = outv;
            }
            fp += ld;
            Fyp = Fy1;
        }
        if (WRITE_H) {
            T yl, yu1, xl, xu;
#if ODINN_SUBGRAD_XY
            subgrad_xy<T, ETA1>(fy * (D1W + D1), ey, -eh, eh1, dy,     // ∂Cy/Δy = -Fy†·Dy/Δy
                                fx * (Dp + D1), ex, -eh, ehE, dx, eta0, yl, yu1, xl, xu);
#else
            subgrad<T, ETA1>(fy * (D1W + D1), ey, -eh, eh1, dy, eta0, yl, yu1);
            subgrad<T, ETA1>(fx * (Dp + D1), ex, -eh, ehE, dx, eta0, xl, xu);
#endif
            T SAW = (Qp - Q1) * qy2 + (aDp + aD1) * T(5);
            T SP = (Pp + P1) * qx2;
            T ZW = shfl_up(SAW + SP + xu);                   // everything column i-1 sends to cell (i, row)
            if (OUT) {
                T res = ZW + (SAW - SP + xl) + (yl + yu_p);
                if (!(h > T(0))) res = T(0);                 // adjoint.jl:148
                if (RKA) rk_emit(res);
                else if (store_lane) *op = res;
            }
            op += ld;
            if (RKA) oo += ld;
            yu_p = yu1;
        }
        h = h1; b = b1; eh = eh1; ex = ex1; hx = hx1; ehE = ehE1; cx = cx1;
        fx = fx1; ly = l1y; px = px1;
        Dp = D1; aDp = aD1; Pp = P1; Qp = Q1;
    }

    template <bool OUT, bool MASKED>
    __device__ __forceinline__ void step(int row) {
        if (CUBIC_FORM) { step_cubic<OUT, MASKED>(row); return; }
        T h1 = hq[0], b1 = bq[0], l1 = lq[0] * lmask;
#pragma unroll
        for (int k = 0; k + 1 < PF; ++k) { hq[k] = hq[k + 1]; bq[k] = bq[k + 1]; lq[k] = lq[k + 1]; }
        if (MASKED) {
            int stp = (row + 1 + PF <= nym1) ? ld : 0;
            hp += stp;
            bp += stp;
            lp += stp;
            if (!(row >= 0 && row + 1 < nym1)) l1 = T(0);  // λ_inn zero-extended on border rows
        } else {
            hp += ld;
            bp += ld;
            lp += ld;
        }
        hq[PF - 1] = __ldg(hp);
        bq[PF - 1] = __ldg(bp);
        lq[PF - 1] = __ldg(lp);
        if (ODINN_L2PF_ROWS1 > 0 && !MASKED) {
            if (row + 1 + PF + ODINN_L2PF_ROWS1 <= nym1) {
                prefetch_l2_row(hp + (long long)ODINN_L2PF_ROWS1 * ld);
                prefetch_l2_row(bp + (long long)ODINN_L2PF_ROWS1 * ld);
                prefetch_l2_row(lp + (long long)ODINN_L2PF_ROWS1 * ld);
            }
        }
        h1 = fmx(h1, T(0));
        b1 = surf_store<T>(b1, h1);
        T eh1 = ETA1 ? h1 : eta0 * h1;
        T hE1 = shfl_dn(h1), bE1 = shfl_dn(b1), lE1 = shfl_dn(l1);
        // x-edge (i→i+1, row+1)
        T ex1 = sdiff<T>(bE1, b1, hE1, h1);
        T hx1 = h1 + hE1;
        T ehE1 = ETA1 ? hE1 : eta0 * hE1;
        T fxr1 = lE1 - l1;                                   // raw Fx† = λ̃[i+1]-λ̃[i]   (adjoint.jl:100)
        T px1 = fxr1 * fmx(fmn(ex1, ehE1), -eh1);            // raw Fx†·clamp(dSdx)       (adjoint.jl:102)
        // y-edge (i, row→row+1)
        T ey = sdiff<T>(b1, b, h1, h);
        T fyr = l1 - l;
        T py = fyr * fmx(fmn(ey, eh1), -eh);
        T eyE = shfl_dn(ey), pyE = shfl_dn(py);
        // node (i, row)
        T gxr = ex + ex1, gyr = ey + eyE;  // raw: ∇Sx = hdx·gxr, ∇Sy = hdy·gyr
        T u = gxr * hdx, v = gyr * hdy;
        T Anode = A;
        T D1, al, be, gA;
        if (AFIELD) {
            Anode = __ldg(ap);
            if (DFIELD) { al = __ldg(alp); be = __ldg(bep); }
            bool adv = true;
            if (MASKED) adv = (row >= 0 && row < ny2);
            if (adv) { ap += ld; if (DFIELD) { alp += ld; bep += ld; } }
        }
        if (DFIELD) { D1 = Anode; gA = T(1); }
        else node_raw<T, CUBIC, true>(ph, Anode, hx + hx1, u * u + v * v, D1, al, be, gA);
        T Dadj = (px + px1) * nhx2 + (py + pyE) * nhy2;  // D† (adjoint.jl:102-104)
        bool node_ok = node_col_ok;
        if (MASKED) node_ok = node_ok && row >= 0 && row < nym1;
        if (!node_ok) Dadj = T(0);            // nodes outside the dual grid: zero-extension of the transposes
        T bD = be * Dadj;
        T aD1 = al * Dadj;                    // α D†
        T P1 = bD * gxr;                      // β D† ∇Sx / hdx
        T Q1 = bD * gyr;                      // β D† ∇Sy / hdy
        if (WRITE_S) {
            if (OUT) {                        // node rows r0..r1-1 belong to this item, columns i0..i0+29
                T vS = gA * Dadj;             // ∂A_spatial ∘ D† (adjoint.jl:250)
                if (own_lane) acc += vS;
                if (AFIELD) { if (own_lane && node_ok) *vp = vS; }
            }
            if (AFIELD) vp += ld;
        }
        if (WRITE_H) {
            T D1W = shfl_up(D1), Q1W = shfl_up(Q1);
            T Qrow1 = Q1W + Q1;
            // y-edge sub-gradient (inversion_utils.jl:36-43): lower cell = row, upper cell = row+1
            T yl, yu1;
            {
                T dC = (fyr * nhy2) * (D1W + D1);  // ∂Cy/Δy = -Fy†·Dy/Δy
                subgrad<T, ETA1>(dC, ey, -eh, eh1, dy, eta0, yl, yu1);
            }
            // x-edge sub-gradient (inversion_utils.jl:22-29): lower cell = i, upper cell = i+1
            T xl, xu;
            {
                T dC = (fxr * nhx2) * (Dp + D1);
                subgrad<T, ETA1>(dC, ex, -eh, ehE, dx, eta0, xl, xu);
            }
            T aDc = T(0.25) * (aDp + aD1);
            T Pc = qx * (Pp + P1);
            T ZW = shfl_up(aDc + Pc + xu);  // everything column i-1 sends to cell (i, row)
            if (OUT) {
                T res = ZW + (aDc - Pc + xl) + qy * (Qrow_p - Qrow1) + (yl + yu_p);
                if (!(h > T(0))) res = T(0);  // adjoint.jl:148
                if (store_lane) *op = res;
            }
            op += ld;
            Qrow_p = Qrow1;
            yu_p = yu1;
        }
        h = h1; b = b1; l = l1; eh = eh1; ex = ex1; hx = hx1; ehE = ehE1; fxr = fxr1; px = px1;
        Dp = D1; aDp = aD1; Pp = P1;
    }
};

// RKA: H2 is the upper snapshot; the glacier's interpolation weight is a = (sign (t_g + lc h_g) - lta) / (ltb - lta) from its controller state.
template <typename T, bool CUBIC, bool AFIELD, bool WRITE_H, bool WRITE_S, bool ETA1, bool DFIELD = false, bool WRITE_F = false, bool RKA = false>
__global__ void __launch_bounds__(MARCH_WARPS * 32, RKA ? 3 : ODINN_VJP1_MIN_CTAS)   // (RKA in fp64: 152 bytes of spills at 128 registers)
sia2d_vjp_march(const GDesc<T>* __restrict__ descs, const int4* __restrict__ items, int n_items,
                const T* __restrict__ lam, const T* __restrict__ H, const T* __restrict__ B, const T* __restrict__ Af,
                T* __restrict__ out, T* __restrict__ vjpA, double* __restrict__ partial, PhysDev<T> ph,
                const T* __restrict__ alF = nullptr, const T* __restrict__ beF = nullptr, T* __restrict__ dH = nullptr,
                RkFuse<T> rkf = RkFuse<T>(), const T* __restrict__ H2 = nullptr, double lc = 0.0, double lsign = 1.0, double lta = 0.0,
                double ltb = 1.0) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<T> d = descs[it.x];
    const int i = it.y + lane, r0 = it.z, r1 = it.w;
    const int ic = min(max(i, 0), d.nx - 1);
    const bool col_inner = (i >= 1 && i <= d.nx - 2);
    VjpMarch<T, CUBIC, AFIELD, WRITE_H, WRITE_S, ETA1, DFIELD, WRITE_F, RKA> m;
    constexpr int PF = PfVjp<T>::value;
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.r0 = r0;
    m.eta0 = ph.eta0;
    m.dx = d.dx;
    m.dy = d.dy;
    m.hdx = T(0.5) * d.inv_dx;
    m.hdy = T(0.5) * d.inv_dy;
    m.nhx2 = -m.hdx * d.inv_dx;  // -½/Δx²
    m.nhy2 = -m.hdy * d.inv_dy;
    m.qx = m.hdx * m.hdx;        // ¼/Δx²
    m.qy = m.hdy * m.hdy;
    m.A = d.A;
    m.lmask = col_inner ? T(1) : T(0);
    m.store_lane = (lane >= 1 && lane <= STRIP && i < d.nx);
    m.node_col_ok = (i >= 0 && i <= d.nx - 2);
    m.own_lane = (lane < STRIP);  // nodes i0 .. i0+29 are owned by this strip
    const int rc = max(r0 - 1, 0);
    m.hp = H + d.off + ic + (long long)rc * d.ld;
    m.bp = B + d.off + ic + (long long)rc * d.ld;
    m.lp = lam + d.off + ic + (long long)rc * d.ld;
    m.ap = AFIELD ? Af + d.off + min(ic, d.nx - 2) + (long long)min(rc, d.ny - 2) * d.ld : nullptr;
    m.alp = DFIELD ? alF + d.off + min(ic, d.nx - 2) + (long long)min(rc, d.ny - 2) * d.ld : nullptr;
    m.bep = DFIELD ? beF + d.off + min(ic, d.nx - 2) + (long long)min(rc, d.ny - 2) * d.ld : nullptr;
    m.op = WRITE_H ? out + d.off + ic + (long long)(r0 - 1) * d.ld : nullptr;
    m.vp = (WRITE_S && AFIELD) ? vjpA + d.off + ic + (long long)(r0 - 1) * d.ld : nullptr;
    m.fp = WRITE_F ? dH + d.off + ic + (long long)(r0 - 1) * d.ld : nullptr;
    {
        // one prefetch instruction per row covers the three input planes: 10 lanes per plane, one 32-byte sector each
        // (a 32-column row segment spans up to 9 sectors in fp64; lanes 30, 31 touch the next strip's first sectors)
        const int pl = min(lane / 10, 2), sec = lane - 10 * pl;
        const T* pb = pl == 0 ? H : (pl == 1 ? B : lam);
        constexpr int per = 32 / (int)sizeof(T);  // elements per sector
        m.pfp = pb + d.off + min(max(it.y + per * sec, 0), d.nx - 1) + (long long)rc * d.ld;
    }

    m.h2p = nullptr; m.pf2p = m.pf3p = nullptr;
    if (RKA) {
        const RkState st = rkf.st[it.x];
        // the glacier has landed on the stop: nothing to integrate (rk_integrate commits by copy then); RKF_LERP_ONLY: see VjpMarch2
        if (st.done && !(rkf.flags & RKF_LERP_ONLY)) return;
        const double tt = lsign * (st.t + lc * st.h);
        m.la1 = (T)((tt - lta) / (ltb - lta));
        m.la0 = T(1) - m.la1;
        m.rk = rkf;
        m.rbh = (T)(rkf.b * st.h);
        m.reh = (T)(rkf.e * st.h);
        m.s1b = lam;
        m.oo = d.off + ic + (long long)(r0 - 1) * d.ld;
        m.r_s1 = m.r_s2 = m.r_e = m.r_u = T(0);
        m.nacc = 0.0;
        m.h2p = H2 + d.off + ic + (long long)rc * d.ld;
        {   // H2, S2in, est (10 lanes each); u on a third instruction when it is read and is not the S2 input
            const int pl = min(lane / 10, 2), sec = lane - 10 * pl;
            const bool first = (rkf.flags & RKF_FIRST) != 0;
            const T* pb = pl == 0 ? H2 : (first ? H2 : (pl == 1 ? rkf.S2in : (const T*)rkf.est));
            constexpr int per = 32 / (int)sizeof(T);
            const long long po = d.off + min(max(it.y + per * sec, 0), d.nx - 1) + (long long)rc * d.ld;
            m.pf2p = pb + po;
            if ((rkf.flags & (RKF_U | RKF_NORM)) && rkf.u != rkf.S2in && lane < 10) m.pf3p = rkf.u + po;
        }
    }

    // ---- cell row r0-1 ----
    m.h = fmx(RKA ? m.la0 * __ldg(m.hp) + m.la1 * __ldg(m.h2p) : __ldg(m.hp), T(0));
    m.b = surf_store<T>(__ldg(m.bp), m.h);
    m.l = __ldg(m.lp) * m.lmask;
    if (!(r0 >= 2 && r0 <= m.nym1)) m.l = T(0);  // row r0-1 must be an inner row
    m.eh = ETA1 ? m.h : m.eta0 * m.h;
    {
        T hE = shfl_dn(m.h), bE = shfl_dn(m.b), lE = shfl_dn(m.l);
        m.ex = sdiff<T>(bE, m.b, hE, m.h);
        m.hx = m.h + hE;
        m.ehE = ETA1 ? hE : m.eta0 * hE;
        m.fxr = lE - m.l;
        m.cx = fmx(fmn(m.ex, m.ehE), -m.eh);
        m.px = m.fxr * m.cx;
    }
    if (CUBIC && !DFIELD) {
        const T Kc = ph.Gam * T(1.0 / 1024.0);  // Γ/4^5
        const T sK = tsqrt(Kc);
        m.hdxs = m.hdx * sK;
        m.hdys = m.hdy * sK;
        m.lmx = m.lmask * m.nhx2;
        m.lmy = m.lmask * m.nhy2;
        m.qx2 = T(2) * Kc * m.qx;
        m.qy2 = T(2) * Kc * m.qy;
        m.nodem = m.node_col_ok ? T(1) : T(0);
        const T lx = m.l * m.nhx2;              // m.l is the masked λ row r0-1
        m.ly = m.l * m.nhy2;
        m.fx = shfl_dn(lx) - lx;
        m.px = m.fx * m.cx;
        m.Qp = T(0);
    }
    m.Dp = m.aDp = m.Pp = m.Qrow_p = m.yu_p = m.acc = m.Fyp = T(0);
#pragma unroll
    for (int k = 0; k < PF; ++k) {  // rows r0 .. r0+PF-1 (clamped)
        if (r0 + k >= 1 && r0 + k <= m.nym1) {
            m.hp += d.ld; m.bp += d.ld; m.lp += d.ld; m.pfp += d.ld;
            if (RKA) { m.h2p += d.ld; m.pf2p += d.ld; if (m.pf3p) m.pf3p += d.ld; }
        }
        m.hq[k] = __ldg(m.hp);
        if (RKA) m.hq2[k] = __ldg(m.h2p);
        m.bq[k] = __ldg(m.bp);
        m.lq[k] = __ldg(m.lp);
    }
    m.pfp += (long long)ODINN_L2PF_ROWS1 * d.ld;
    if (RKA) { m.pf2p += (long long)ODINN_L2PF_ROWS1 * d.ld; if (m.pf3p) m.pf3p += (long long)ODINN_L2PF_ROWS1 * d.ld; }

    int row = r0 - 1;
    m.template step<false, true>(row);  // warm-up
    ++row;
    const int main_end = min(r1, d.ny - 1 - PF);
    for (; row < min(r1, 1); ++row) m.template step<true, true>(row);
#pragma unroll 2
    for (; row < main_end; ++row) m.template step<true, false>(row);
    for (; row < r1; ++row) m.template step<true, true>(row);

    if (RKA) {
        if (rkf.flags & RKF_NORM) {   // (only storing lanes accumulated)
            double a = m.nacc;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
            if (lane == 0) partial[item] = a;
        }
    }
    if (WRITE_S) {
        double a = m.own_lane ? (double)m.acc : 0.0;  // (the cubic form accumulates in the halo lanes too)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
        if (lane == 0) partial[item] = a;
    }
}

// Second stage of the A2 reduction: one CTA per glacier sums its work items' partials in a fixed order.
static __global__ void __launch_bounds__(NT)
reduce_items_kernel(const int* __restrict__ item_start, const double* __restrict__ partial, double* __restrict__ S) {
    __shared__ double sRed[NT / 32];
    int g = blockIdx.x;
    int t0 = item_start[g], t1 = item_start[g + 1];
    double acc = 0.0;
    for (int t = t0 + threadIdx.x; t < t1; t += NT) acc += partial[t];
    double s = block_sum(acc, sRed);
    if (threadIdx.x == 0) S[g] = s;
}

}  // namespace odinn
