// fp32 register-marching kernels, two columns per lane, packed f32x2 arithmetic (Blackwell FADD2/FMUL2/FFMA2).
//
// ncu of the one-column kernels (profiles/r01_v3_*): DRAM traffic == algorithmic bytes, but 67 (F1) / 121 (A1+A2)
// thread-instructions per cell keep the issue slots 70-74 % busy at 66 % / 52 % of the HBM roofline -- the kernels
// are instruction-issue bound.  Measured pipe rates on B200 (tools/microbench/pipes.cu): FFMA 3.6, FFMA2 1.8,
// FMNMX 2.0, SHFL 1.0 warp-instructions / clk / SM.  A packed FFMA2 costs ONE issue slot for two columns, and with
// two columns per lane only every second x-neighbour is in another lane, so shuffles, loads, stores and address
// arithmetic per cell halve as well.  Same arithmetic, same operation order per column as sia2d_march.cuh
// (results are bit-identical to the one-column kernels; the parity tests cover both).
//
// Geometry: a warp owns 64 consecutive columns  base .. base+63  (base even), lane l holds columns base+2l (.x)
// and base+2l+1 (.y); lanes 0 and 31 are halo lanes, so a strip produces the 60 columns base+2 .. base+61.
// Every row access of a warp is one 256-byte LDG.64 / STG.64.  Requires even ld and even plane offsets.
#pragma once
#include "sia2d_march.cuh"

namespace odinn {

constexpr int STRIP2 = 60;       // output columns per warp
#ifndef ODINN_MARCH2_WARPS
#define ODINN_MARCH2_WARPS 4
#endif
constexpr int MARCH2_WARPS = ODINN_MARCH2_WARPS;
#ifndef ODINN_PF2_RHS
#define ODINN_PF2_RHS 4
#endif
#ifndef ODINN_PF2_VJP
#define ODINN_PF2_VJP 2
#endif

typedef float2 f2;
__device__ __forceinline__ f2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __ffma2_rn(b, mk2(-1.0f, -1.0f), a); }  // a - b, one FFMA2
__device__ __forceinline__ f2 max2(f2 a, f2 b) { return mk2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
__device__ __forceinline__ f2 min2(f2 a, f2 b) { return mk2(fminf(a.x, b.x), fminf(a.y, b.y)); }
__device__ __forceinline__ f2 neg2(f2 a) { return mk2(-a.x, -a.y); }
// value of the column to the east / west of each of the lane's two columns
__device__ __forceinline__ f2 east2(f2 v) { return mk2(v.y, __shfl_down_sync(FULL, v.x, 1)); }
__device__ __forceinline__ f2 west2(f2 v) { return mk2(__shfl_up_sync(FULL, v.y, 1), v.x); }
__device__ __forceinline__ f2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
// raw surface difference, fp32 form (see sdiff<float>): (b1 - b0) + (h1 - h0)
__device__ __forceinline__ f2 sdiff2(f2 b1, f2 b0, f2 h1, f2 h0) { return add2(sub2(b1, b0), sub2(h1, h0)); }
// clamp(e, -lo, up) = max(min(e, up), -lo)
__device__ __forceinline__ f2 clamp2(f2 e, f2 up, f2 lo) {
    return mk2(fmaxf(fminf(e.x, up.x), -lo.x), fmaxf(fminf(e.y, up.y), -lo.y));
}

template <bool CUBIC, bool PARTIALS>
__device__ __forceinline__ void node_raw2(const PhysDev<float>& ph, f2 A, f2 Hs, f2 g2, f2& D, f2& alpha, f2& beta, f2& gA) {
    if (CUBIC) {
        const float K = ph.Gam * (1.0f / 1024.0f);  // Γ/4^5
        f2 H2 = mul2(Hs, Hs);
        f2 H4 = mul2(H2, H2);
        f2 w = mul2(H4, Hs);
        if (!PARTIALS) {
            D = mul2(mul2(mul2(A, bc2(K)), g2), w);
            return;
        }
        f2 tg = mul2(bc2(K), g2);
        gA = mul2(tg, w);
        D = mul2(A, gA);
        alpha = mul2(mul2(bc2(20.0f), A), mul2(tg, H4));
        beta = mul2(mul2(mul2(bc2(2.0f), A), bc2(K)), w);
    } else {
        float Dx, ax, bx, gx, Dy, ay, by, gy;
        node_diffusivity<float, false, PARTIALS>(ph, A.x, 0.25f * Hs.x, g2.x, Dx, ax, bx, gx);
        node_diffusivity<float, false, PARTIALS>(ph, A.y, 0.25f * Hs.y, g2.y, Dy, ay, by, gy);
        D = mk2(Dx, Dy);
        if (PARTIALS) { alpha = mk2(ax, ay); beta = mk2(bx, by); gA = mk2(gx, gy); }
    }
}

// --------------------------------------------------------------------------------------------
// F1 (see RhsMarch for the step structure; every quantity is a pair of adjacent columns)
// --------------------------------------------------------------------------------------------
template <bool CUBIC, bool AFIELD, bool ETA1, bool STAGE>
struct RhsMarch2 {
    static constexpr int PF = ODINN_PF2_RHS;
    const float *hp, *bp, *ap, *up;
    float* op;
    int ld, nym1, ny2;
    float eta0;
    f2 hdx, hdy, kx, ky, A;  // kx, ky are zeroed on border / out-of-grid columns
    f2 sa, sb, sdt, hraw;
    bool store_pair, store_x, y_oob;
    PhysDev<float> ph;
    f2 h, b, eh, ex, hx, ehE, Dp, Fy;
    f2 hq[PF], bq[PF];

    __device__ __forceinline__ void sanitize(f2& hv, f2& bv) const {
        if (y_oob) { hv.y = 0.0f; bv.y = bv.x; }  // the pair straddles the last column (odd nx): keep it finite
    }

    template <bool OUT, bool MASKED>
    __device__ __forceinline__ void step(int row) {
        f2 h1 = hq[0], b1 = bq[0];
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
        hq[PF - 1] = ldg2(hp);
        bq[PF - 1] = ldg2(bp);
        f2 u0 = bc2(0.0f);
        if (STAGE && OUT) { if (store_pair) u0 = ldg2(up); else if (store_x) u0.x = __ldg(up); }
        sanitize(h1, b1);
        const f2 hraw1 = h1;
        h1 = max2(h1, bc2(0.0f));                 // adjoint.jl:52
        f2 eh1 = ETA1 ? h1 : mul2(bc2(eta0), h1);
        f2 hE1 = east2(h1), bE1 = east2(b1);
        f2 ex1 = sdiff2(bE1, b1, hE1, h1);        // raw x-edge difference S[i+1]-S[i], row+1
        f2 hx1 = add2(h1, hE1);
        f2 ehE1 = ETA1 ? hE1 : mul2(bc2(eta0), hE1);
        f2 ey = sdiff2(b1, b, h1, h);             // raw y-edge difference S[j+1]-S[j]
        f2 eyE = east2(ey);
        f2 u = mul2(add2(ex, ex1), hdx), v = mul2(add2(ey, eyE), hdy);
        f2 g2 = fma2(v, v, mul2(u, u));
        f2 Anode = A;
        if (AFIELD) {
            Anode = ldg2(ap);
            if (MASKED) { if (row >= 0 && row < ny2) ap += ld; } else ap += ld;
        }
        f2 D1, al, be, gA;
        node_raw2<CUBIC, false>(ph, Anode, add2(hx, hx1), g2, D1, al, be, gA);
        f2 D1W = west2(D1);
        f2 Fy1 = mul2(add2(D1W, D1), clamp2(ey, eh1, eh));
        f2 Fx = mul2(add2(Dp, D1), clamp2(ex, ehE, eh));
        f2 FxW = west2(Fx);
        if (OUT) {
            f2 outv = fma2(ky, sub2(Fy1, Fy), mul2(kx, sub2(Fx, FxW)));
            if (MASKED) { if (row < 1 || row >= nym1) outv = bc2(0.0f); }
            if (STAGE) outv = fma2(sb, fma2(sdt, outv, hraw), mul2(sa, u0));
            if (store_pair) *reinterpret_cast<float2*>(op) = outv;
            else if (store_x) *op = outv.x;
        }
        op += ld;
        if (STAGE) { up += ld; hraw = hraw1; }
        h = h1; b = b1; eh = eh1; ex = ex1; hx = hx1; ehE = ehE1; Dp = D1; Fy = Fy1;
    }
};

template <bool CUBIC, bool AFIELD, bool ETA1, bool STAGE>
__global__ void __launch_bounds__(MARCH2_WARPS * 32)
sia2d_rhs_march2(const GDesc<float>* __restrict__ descs, const int4* __restrict__ items, int n_items,
                 const float* __restrict__ H, const float* __restrict__ B, const float* __restrict__ Af, float* dH,
                 PhysDev<float> ph, const float* U0, float sa, float sb, float sdt) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * MARCH2_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<float> d = descs[it.x];
    const int c0 = it.y + 2 * lane, r0 = it.z, r1 = it.w;  // it.y even
    // pair index clamped into the grid (pairs start at even columns; the last pair may straddle nx when nx is odd)
    const int cmax = (d.nx - 1) & ~1;
    const int ic = min(max(c0, 0), cmax);
    RhsMarch2<CUBIC, AFIELD, ETA1, STAGE> m;
    constexpr int PF = ODINN_PF2_RHS;
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.eta0 = ph.eta0;
    const float hdx = 0.5f * d.inv_dx, hdy = 0.5f * d.inv_dy;
    m.hdx = bc2(hdx);
    m.hdy = bc2(hdy);
    const bool inx = (c0 >= 1 && c0 <= d.nx - 2), iny = (c0 + 1 >= 1 && c0 + 1 <= d.nx - 2);
    m.kx = mk2(inx ? hdx * d.inv_dx : 0.0f, iny ? hdx * d.inv_dx : 0.0f);  // ½/Δx²
    m.ky = mk2(inx ? hdy * d.inv_dy : 0.0f, iny ? hdy * d.inv_dy : 0.0f);
    m.A = bc2(d.A);
    const bool out_lane = (lane >= 1 && lane <= 30 && c0 >= 0);
    m.store_pair = out_lane && (c0 + 1 < d.nx);
    m.store_x = out_lane && (c0 + 1 == d.nx);
    m.y_oob = (ic + 1 >= d.nx);
    const int rc = max(r0 - 1, 0);
    m.hp = H + d.off + ic + (long long)rc * d.ld;
    m.bp = B + d.off + ic + (long long)rc * d.ld;
    m.ap = AFIELD ? Af + d.off + ic + (long long)min(rc, d.ny - 2) * d.ld : nullptr;
    m.op = dH + d.off + ic + (long long)(r0 - 1) * d.ld;  // dereferenced for rows >= r0 only
    m.up = STAGE ? U0 + d.off + ic + (long long)(r0 - 1) * d.ld : nullptr;
    m.sa = bc2(sa);
    m.sb = bc2(sb);
    m.sdt = bc2(sdt);
    m.hraw = bc2(0.0f);

    // ---- cell row r0-1 ----
    {
        f2 hv = ldg2(m.hp), bv = ldg2(m.bp);
        m.sanitize(hv, bv);
        m.h = max2(hv, bc2(0.0f));
        m.b = bv;
    }
    m.eh = ETA1 ? m.h : mul2(bc2(m.eta0), m.h);
    {
        f2 hE = east2(m.h), bE = east2(m.b);
        m.ex = sdiff2(bE, m.b, hE, m.h);
        m.hx = add2(m.h, hE);
        m.ehE = ETA1 ? hE : mul2(bc2(m.eta0), hE);
    }
    m.Dp = bc2(0.0f);
    m.Fy = bc2(0.0f);
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        if (r0 + k >= 1 && r0 + k <= m.nym1) { m.hp += d.ld; m.bp += d.ld; }
        m.hq[k] = ldg2(m.hp);
        m.bq[k] = ldg2(m.bp);
    }

    int row = r0 - 1;
    m.template step<false, true>(row);
    ++row;
    const int main_end = min(r1, d.ny - 1 - PF);
    for (; row < min(r1, 1); ++row) m.template step<true, true>(row);
#pragma unroll 4
    for (; row < main_end; ++row) m.template step<true, false>(row);
    for (; row < r1; ++row) m.template step<true, true>(row);
}

}  // namespace odinn
