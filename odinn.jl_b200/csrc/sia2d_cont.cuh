// Continuous-adjoint VJPs (differentiate-then-discretise), register-marching form.
//
// Reference (ODINN.jl v1.1.0):
//   A1c  VJP_λ_∂SIA∂H_continuous   src/inverse/SIA2D/adjoint.jl:442-555
//        dλ_inn = ∇·(D∇λ) - avg(∂D/∂H̄)·avg(⟨∇S,∇λ⟩) + avg_y(∂x(β∇Sx⟨∇S,∇λ⟩)) + avg_x(∂y(β∇Sy⟨∇S,∇λ⟩)),  border 0,
//        no flux clamp, no H > 0 mask, λ used as is (not zero-extended).
//   A2c  VJP_λ_∂SIA∂θ_continuous   src/inverse/SIA2D/adjoint.jl:582-662
//        ∂θ_k = Σ_ij λ[i,j]·pad(∂x(avg_y(∂D/∂θ_k)·clamp(dSdx)) + ∂y(...))[i,j].  For a glacier-wide law
//        ∂D/∂θ_k = ∂A_spatial·(∂A/∂θ_k) (target_A.jl:85-87), so  ∂θ = (∂A/∂θ)·Σ λ ⊙ SIA2D_{A≡1, C≡0}(H):
//        one F1 evaluation with unit A and no sliding (sia2d_rhs_march with the A override) + dot_kernel below.
//
// Same marching scheme as sia2d_march.cuh: one warp owns 32 columns (30 outputs), the previous row is carried in
// registers, x-neighbours come from warp shuffles.  Per dual node (i,row):
//   ls = ½[(ex·lx + ex1·lx1)/Δx² + (ey·ly + eyE·lyE)/Δy²]          ⟨∇S,∇λ⟩ averaged to the node (adjoint.jl:535-539)
//   X  = ls·β·∇Sx ,  Y = ls·β·∇Sy                                   (adjoint.jl:544-545)
// and per inner cell (i,row), from the four surrounding nodes:
//   div = ½/Δx²·Δx(Fx_raw) + ½/Δy²·Δy(Fy_raw),  F_raw = (D+D)·Δλ      (adjoint.jl:522-532)
//   t2  = ¼Σα · ¼Σls                                                (adjoint.jl:541)
//   t3  = ½/Δx·[(X[i,·]-X[i-1,·]) summed over the two node rows] + ½/Δy·[(Y[·,row]-Y[·,row-1]) summed over the two node columns]
#pragma once
#include "sia2d_march.cuh"

namespace odinn {

template <typename T, bool CUBIC, bool AFIELD>
struct VjpcMarch {
    static constexpr int PF = 2;
    const T *hp, *bp, *lp, *ap;
    T* op;
    int ld, nym1, ny2;
    T hdx, hdy, kx, ky, lsx, lsy, tx, ty, A;  // kx, ky, tx, ty zeroed on border columns
    T cmask;                                  // 1 on inner columns
    bool store_lane;
    PhysDev<T> ph;
    // carried row state: cell row `row`
    T h, b, l, ex, lx, hx;
    // carried node row `row-1`
    T Dp, ap_, lsp, Xp, Yp, Fy;
    T hq[PF], bq[PF], lq[PF];

    template <bool OUT, bool MASKED>
    __device__ __forceinline__ void step(int row) {
        T h1 = hq[0], b1 = bq[0], l1 = lq[0];
#pragma unroll
        for (int k = 0; k + 1 < PF; ++k) { hq[k] = hq[k + 1]; bq[k] = bq[k + 1]; lq[k] = lq[k + 1]; }
        if (MASKED) {
            int stp = (row + 1 + PF <= nym1) ? ld : 0;
            hp += stp; bp += stp; lp += stp;
        } else {
            hp += ld; bp += ld; lp += ld;
        }
        hq[PF - 1] = __ldg(hp);
        bq[PF - 1] = __ldg(bp);
        lq[PF - 1] = __ldg(lp);
        h1 = fmx(h1, T(0));                       // adjoint.jl:463
        b1 = surf_store<T>(b1, h1);
        T hE1 = shfl_dn(h1), bE1 = shfl_dn(b1), lE1 = shfl_dn(l1);
        T ex1 = sdiff<T>(bE1, b1, hE1, h1);       // raw S[i+1]-S[i], row+1
        T lx1 = lE1 - l1;                         // raw λ[i+1]-λ[i], row+1
        T hx1 = h1 + hE1;
        T ey = sdiff<T>(b1, b, h1, h);            // raw S[j+1]-S[j]
        T ly = l1 - l;
        T eyE = shfl_dn(ey), lyE = shfl_dn(ly);
        // node (i, row)
        T gxr = ex + ex1, gyr = ey + eyE;
        T u = gxr * hdx, v = gyr * hdy;           // ∇Sx, ∇Sy
        T Anode = A;
        if (AFIELD) {
            Anode = __ldg(ap);
            if (MASKED) { if (row >= 0 && row < ny2) ap += ld; } else ap += ld;
        }
        T D1, al, be, gA;
        node_raw<T, CUBIC, true>(ph, Anode, hx + hx1, u * u + v * v, D1, al, be, gA);
        T ls1 = lsx * (ex * lx + ex1 * lx1) + lsy * (ey * ly + eyE * lyE);
        T lb = ls1 * be;
        T X1 = lb * u, Y1 = lb * v;
        // fluxes of ∇·(D∇λ): y-edge (i, row→row+1) and x-edge (i→i+1, row)
        T D1W = shfl_up(D1);
        T Fy1 = (D1W + D1) * ly;
        T Fx = (Dp + D1) * lx;
        T FxW = shfl_up(Fx);
        // node-column sums over the two node rows row-1, row
        T asum = ap_ + al, lssum = lsp + ls1, Xsum = Xp + X1, Ydif = Y1 - Yp;
        T asumW = shfl_up(asum), lssumW = shfl_up(lssum), XsumW = shfl_up(Xsum), YdifW = shfl_up(Ydif);
        if (OUT) {
            T div = kx * (Fx - FxW) + ky * (Fy1 - Fy);
            T t2 = (T(0.0625) * cmask) * ((asumW + asum) * (lssumW + lssum));
            T t3 = tx * (Xsum - XsumW) + ty * (YdifW + Ydif);
            T res = div - t2 + t3;
            if (MASKED) { if (row < 1 || row >= nym1) res = T(0); }
            if (store_lane) *op = res;
        }
        op += ld;
        h = h1; b = b1; l = l1; ex = ex1; lx = lx1; hx = hx1;
        Dp = D1; ap_ = al; lsp = ls1; Xp = X1; Yp = Y1; Fy = Fy1;
    }
};

template <typename T, bool CUBIC, bool AFIELD>
__global__ void __launch_bounds__(MARCH_WARPS * 32)
sia2d_vjpc_march(const GDesc<T>* __restrict__ descs, const int4* __restrict__ items, int n_items,
                 const T* __restrict__ lam, const T* __restrict__ H, const T* __restrict__ B, const T* __restrict__ Af,
                 T* __restrict__ out, PhysDev<T> ph) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<T> d = descs[it.x];
    const int i = it.y + lane, r0 = it.z, r1 = it.w;
    const int ic = min(max(i, 0), d.nx - 1);
    const bool col_inner = (i >= 1 && i <= d.nx - 2);
    VjpcMarch<T, CUBIC, AFIELD> m;
    constexpr int PF = VjpcMarch<T, CUBIC, AFIELD>::PF;
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.hdx = T(0.5) * d.inv_dx;
    m.hdy = T(0.5) * d.inv_dy;
    m.kx = col_inner ? m.hdx * d.inv_dx : T(0);  // ½/Δx²
    m.ky = col_inner ? m.hdy * d.inv_dy : T(0);
    m.lsx = m.hdx * d.inv_dx;                    // ½/Δx²
    m.lsy = m.hdy * d.inv_dy;
    m.tx = col_inner ? m.hdx : T(0);             // ½/Δx
    m.ty = col_inner ? m.hdy : T(0);
    m.cmask = col_inner ? T(1) : T(0);
    m.A = d.A;
    m.store_lane = (lane >= 1 && lane <= STRIP && i < d.nx);
    const int rc = max(r0 - 1, 0);
    m.hp = H + d.off + ic + (long long)rc * d.ld;
    m.bp = B + d.off + ic + (long long)rc * d.ld;
    m.lp = lam + d.off + ic + (long long)rc * d.ld;
    m.ap = AFIELD ? Af + d.off + min(ic, d.nx - 2) + (long long)min(rc, d.ny - 2) * d.ld : nullptr;
    m.op = out + d.off + ic + (long long)(r0 - 1) * d.ld;

    m.h = fmx(__ldg(m.hp), T(0));
    m.b = surf_store<T>(__ldg(m.bp), m.h);
    m.l = __ldg(m.lp);
    {
        T hE = shfl_dn(m.h), bE = shfl_dn(m.b), lE = shfl_dn(m.l);
        m.ex = sdiff<T>(bE, m.b, hE, m.h);
        m.lx = lE - m.l;
        m.hx = m.h + hE;
    }
    m.Dp = m.ap_ = m.lsp = m.Xp = m.Yp = m.Fy = T(0);
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        if (r0 + k >= 1 && r0 + k <= m.nym1) { m.hp += d.ld; m.bp += d.ld; m.lp += d.ld; }
        m.hq[k] = __ldg(m.hp);
        m.bq[k] = __ldg(m.bp);
        m.lq[k] = __ldg(m.lp);
    }
    int row = r0 - 1;
    m.template step<false, true>(row);  // warm-up: node row r0-1
    ++row;
    const int main_end = min(r1, d.ny - 1 - PF);
    for (; row < min(r1, 1); ++row) m.template step<true, true>(row);
#pragma unroll 2
    for (; row < main_end; ++row) m.template step<true, false>(row);
    for (; row < r1; ++row) m.template step<true, true>(row);
}

// partial[tile] = Σ_tile a·b over the inner cells of the glacier (border excluded: the `pad` of adjoint.jl:653-654)
template <typename T>
__global__ void __launch_bounds__(NT)
dot_inner_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const T* __restrict__ a,
                 const T* __restrict__ b, double* __restrict__ partial) {
    __shared__ double sRed[NT / 32];
    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    const int tx = threadIdx.x & 31, tr = threadIdx.x >> 5;
    const int i = x0 + tx;
    double acc = 0.0;
#pragma unroll
    for (int rr = 0; rr < TY / 8; ++rr) {
        int j = y0 + tr + rr * 8;
        if (i >= 1 && i <= d.nx - 2 && j >= 1 && j <= d.ny - 2) {
            long long p = d.off + (long long)j * d.ld + i;
            acc += (double)__ldg(a + p) * (double)__ldg(b + p);
        }
    }
    double s = block_sum(acc, sRed);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

}  // namespace odinn
