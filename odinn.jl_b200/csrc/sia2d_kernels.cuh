// F1 (forward RHS) and A1/A2 (discrete VJPs) of the SIA2D stencil, batched over an ensemble.
//
// Reference semantics (ODINN.jl v1.1.0):
//   forward intermediates      src/inverse/SIA2D/adjoint.jl:47-104   (= Huginn.SIA2D!, not in tree)
//   flux clamp + sub-gradients src/inverse/SIA2D/inversion_utils.jl:17-43
//   operator transposes        src/inverse/SIA2D/inversion_utils.jl:3-15, 45-66
//   dH-VJP                     src/inverse/SIA2D/adjoint.jl:99-148
//   dθ-VJP                     src/inverse/SIA2D/adjoint.jl:235-250, src/models/target/target_A.jl:64-92
//
// Index convention: cell (i,j), i the contiguous axis.  Dual node (a,b) sits between cells
// a..a+1 x b..b+1 (valid a in 0..nx-2, b in 0..ny-2).  x-edge (a,j) joins cells (a,j),(a+1,j);
// y-edge (i,b) joins cells (i,b),(i,b+1).
//
// sB holds S = B + H in fp64 and B in fp32 (see sdiff below).
#pragma once
#include "common.cuh"

namespace odinn {

constexpr int CW = TX + 2;  // cell tile width incl. halo
constexpr int CH = TY + 2;
constexpr int NW = TX + 1;  // node tile
constexpr int NH = TY + 1;

// Surface differences.  fp64 follows the reference literally: S = B + H is rounded first and slopes are
// differences of S (adjoint.jl:54-59) -- structured ties of the strict clamp inequalities (e.g. a margin edge
// over a locally flat bed) are then resolved by the same rounding noise as in the reference.  fp32 keeps B
// and forms (B1-B0) + (H1-H0), so that a 3000 m bed does not cost 3 decimal digits of slope.
template <typename T> __device__ __forceinline__ T surf_store(T b, T h);
template <> __device__ __forceinline__ double surf_store<double>(double b, double h) { return b + h; }
template <> __device__ __forceinline__ float surf_store<float>(float b, float h) { return b; }
template <typename T> __device__ __forceinline__ T sdiff(T b1, T b0, T h1, T h0);
template <> __device__ __forceinline__ double sdiff<double>(double s1, double s0, double, double) { return s1 - s0; }
template <> __device__ __forceinline__ float sdiff<float>(float b1, float b0, float h1, float h0) { return (b1 - b0) + (h1 - h0); }

// Clamped edge slope (raw, i.e. not yet divided by Δ):  max(min(dS, η₀ H_up), -η₀ H_lo)
// (inversion_utils.jl:17-20, 31-34).  H_lo is the thickness at the lower-index cell.
template <typename T>
__device__ __forceinline__ T clamp_raw(T dS, T eta0, T H_lo, T H_up) {
    return tmax(tmin(dS, eta0 * H_up), -(eta0 * H_lo));
}

// Strict comparisons of the clamp sub-gradient (inversion_utils.jl:24-28, 38-42).  The reference compares the
// DIVIDED quantities  dS/Δ  and  ±η₀H/Δ ; two raw values one ulp apart can round to the same quotient, which
// turns a strict inequality into a tie (this happens on structured inputs: a margin edge over a locally flat
// bed).  The kernels work with raw differences, so fp64 re-checks near-ties with true divisions (rare, divergent
// slow path) to take the same branch as the reference; fp32 is only ever compared within 1e-5 and skips it.
__device__ __forceinline__ bool gt_div(double x, double y, double d) {
    if (!(x > y)) return false;
    if ((x - y) > 1e-14 * fmax(fabs(x), fabs(y))) return true;
    return (x / d) > (y / d);
}
__device__ __forceinline__ bool gt_div(float x, float y, float) { return x > y; }

// Stage one (TX+2) x (TY+2) tile of max(H,0) and B in shared memory; zero outside the grid.
template <typename T>
__device__ __forceinline__ void load_cells(const GDesc<T>& d, int x0, int y0, const T* __restrict__ H,
                                           const T* __restrict__ B, T (*sH)[CW], T (*sB)[CW]) {
    for (int idx = threadIdx.x; idx < CW * CH; idx += NT) {
        int lx = idx % CW, ly = idx / CW;
        int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
        T h = T(0), b = T(0);
        if (gx >= 0 && gx < d.nx && gy >= 0 && gy < d.ny) {
            long long p = d.off + (long long)gy * d.ld + gx;
            h = __ldg(H + p);
            b = __ldg(B + p);
        }
        h = h > T(0) ? h : T(0);  // adjoint.jl:52
        sH[ly][lx] = h;
        sB[ly][lx] = surf_store<T>(b, h);
    }
}

// --------------------------------------------------------------------------------------------
// F1: dH = SIA2D(H)
// --------------------------------------------------------------------------------------------
template <typename T, bool CUBIC, bool AFIELD>
__global__ void __launch_bounds__(NT)
sia2d_rhs_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const T* __restrict__ H,
                 const T* __restrict__ B, const T* __restrict__ Af, T* __restrict__ dH, PhysDev<T> ph) {
    __shared__ T sH[CH][CW];
    __shared__ T sB[CH][CW];
    __shared__ T sD[NH][NW];

    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;

    load_cells<T>(d, x0, y0, H, B, sH, sB);
    __syncthreads();

    // Dual nodes: local (lx,ly) <-> global node (x0-1+lx, y0-1+ly); its cells are local
    // (lx..lx+1, ly..ly+1).
    for (int idx = threadIdx.x; idx < NW * NH; idx += NT) {
        int lx = idx % NW, ly = idx / NW;
        int a = x0 - 1 + lx, b = y0 - 1 + ly;
        T D = T(0);
        if (a >= 0 && a <= d.nx - 2 && b >= 0 && b <= d.ny - 2) {
            T h00 = sH[ly][lx], h10 = sH[ly][lx + 1], h01 = sH[ly + 1][lx], h11 = sH[ly + 1][lx + 1];
            T b00 = sB[ly][lx], b10 = sB[ly][lx + 1], b01 = sB[ly + 1][lx], b11 = sB[ly + 1][lx + 1];
            T sx0 = sdiff<T>(b10, b00, h10, h00), sx1 = sdiff<T>(b11, b01, h11, h01);  // diff_x(S) rows b, b+1
            T sy0 = sdiff<T>(b01, b00, h01, h00), sy1 = sdiff<T>(b11, b10, h11, h10);  // diff_y(S) cols a, a+1
            T gx = T(0.5) * (sx0 + sx1) * d.inv_dx;                               // avg_y(dSdx), adjoint.jl:60
            T gy = T(0.5) * (sy0 + sy1) * d.inv_dy;                               // avg_x(dSdy), :61
            T Hb = T(0.25) * ((h00 + h10) + (h01 + h11));                         // avg(H), :67
            T A = AFIELD ? __ldg(Af + d.off + (long long)b * d.ld + a) : d.A;
            T al, be, gA;
            node_diffusivity<T, CUBIC, false>(ph, A, Hb, gx * gx + gy * gy, D, al, be, gA);
        }
        sD[ly][lx] = D;
    }
    __syncthreads();

    const int tx = threadIdx.x & 31, tr = threadIdx.x >> 5;
    const int i = x0 + tx;
    if (i >= d.nx) return;
#pragma unroll
    for (int rr = 0; rr < TY / 8; ++rr) {
        int r = tr + rr * 8;
        int j = y0 + r;
        if (j >= d.ny) break;
        T out = T(0);
        if (i >= 1 && i <= d.nx - 2 && j >= 1 && j <= d.ny - 2) {
            const int cx = tx + 1, cy = r + 1;
            T hc = sH[cy][cx], hw = sH[cy][cx - 1], he = sH[cy][cx + 1], hs = sH[cy - 1][cx], hn = sH[cy + 1][cx];
            T bc = sB[cy][cx], bw = sB[cy][cx - 1], be = sB[cy][cx + 1], bs = sB[cy - 1][cx], bn = sB[cy + 1][cx];
            T d00 = sD[r][tx], d10 = sD[r][tx + 1], d01 = sD[r + 1][tx], d11 = sD[r + 1][tx + 1];
            // x-edges (i-1,j) and (i,j): Dx = avg_y(D) (adjoint.jl:96), clamped slope (:93)
            T cw = clamp_raw<T>(sdiff<T>(bc, bw, hc, hw), ph.eta0, hw, hc) * d.inv_dx;
            T ce = clamp_raw<T>(sdiff<T>(be, bc, he, hc), ph.eta0, hc, he) * d.inv_dx;
            T Fw = -(T(0.5) * (d00 + d01)) * cw;
            T Fe = -(T(0.5) * (d10 + d11)) * ce;
            // y-edges (i,j-1) and (i,j): Dy = avg_x(D) (:97)
            T cs = clamp_raw<T>(sdiff<T>(bc, bs, hc, hs), ph.eta0, hs, hc) * d.inv_dy;
            T cn = clamp_raw<T>(sdiff<T>(bn, bc, hn, hc), ph.eta0, hc, hn) * d.inv_dy;
            T Fs = -(T(0.5) * (d00 + d10)) * cs;
            T Fn = -(T(0.5) * (d01 + d11)) * cn;
            out = -((Fe - Fw) * d.inv_dx + (Fn - Fs) * d.inv_dy);  // adjoint.jl:528-533, 552-553
        }
        dH[d.off + (long long)j * d.ld + i] = out;
    }
}

// --------------------------------------------------------------------------------------------
// A1 + A2: discrete VJPs.
//   WRITE_H : out[i,j] = ((∂SIA/∂H)^T λ)[i,j]                      (adjoint.jl:99-148)
//   WRITE_S : partial[tile] = Σ_nodes-of-tile gA · D†               (adjoint.jl:235-250)
//             and, when vjpA != nullptr, vjpA[node] = gA · D†       (gridded A, target_utils.jl:163-173)
// Nodes on a tile seam belong to the tile holding cell (a+1,b+1)'s lower-left... precisely:
// node (a,b) is owned by the tile whose cell range contains (a,b).
// --------------------------------------------------------------------------------------------
template <typename T, bool CUBIC, bool AFIELD, bool WRITE_H, bool WRITE_S>
__global__ void __launch_bounds__(NT)
sia2d_vjp_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const T* __restrict__ lam,
                 const T* __restrict__ H, const T* __restrict__ B, const T* __restrict__ Af, T* __restrict__ out,
                 T* __restrict__ vjpA, double* __restrict__ partial, PhysDev<T> ph) {
    __shared__ T sH[CH][CW];
    __shared__ T sB[CH][CW];
    __shared__ T sL[CH][CW];   // λ with border cells and out-of-grid zeroed (λ_inn zero-extended)
    __shared__ T sD[NH][NW];   // D
    __shared__ T sA[NH][NW];   // α D†
    __shared__ T sP[NH][NW];   // β ∇Sx D†
    __shared__ T sQ[NH][NW];   // β ∇Sy D†
    __shared__ double sRed[NT / 32];

    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;

    load_cells<T>(d, x0, y0, H, B, sH, sB);
    for (int idx = threadIdx.x; idx < CW * CH; idx += NT) {
        int lx = idx % CW, ly = idx / CW;
        int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
        T l = T(0);
        if (gx >= 1 && gx <= d.nx - 2 && gy >= 1 && gy <= d.ny - 2) l = __ldg(lam + d.off + (long long)gy * d.ld + gx);
        sL[ly][lx] = l;
    }
    __syncthreads();

    double acc = 0.0;
    for (int idx = threadIdx.x; idx < NW * NH; idx += NT) {
        int lx = idx % NW, ly = idx / NW;
        int a = x0 - 1 + lx, b = y0 - 1 + ly;
        T D = T(0), aD = T(0), P = T(0), Q = T(0);
        if (a >= 0 && a <= d.nx - 2 && b >= 0 && b <= d.ny - 2) {
            T h00 = sH[ly][lx], h10 = sH[ly][lx + 1], h01 = sH[ly + 1][lx], h11 = sH[ly + 1][lx + 1];
            T b00 = sB[ly][lx], b10 = sB[ly][lx + 1], b01 = sB[ly + 1][lx], b11 = sB[ly + 1][lx + 1];
            T l00 = sL[ly][lx], l10 = sL[ly][lx + 1], l01 = sL[ly + 1][lx], l11 = sL[ly + 1][lx + 1];
            T sx0 = sdiff<T>(b10, b00, h10, h00), sx1 = sdiff<T>(b11, b01, h11, h01);
            T sy0 = sdiff<T>(b01, b00, h01, h00), sy1 = sdiff<T>(b11, b10, h11, h10);
            T gSx = T(0.5) * (sx0 + sx1) * d.inv_dx;
            T gSy = T(0.5) * (sy0 + sy1) * d.inv_dy;
            T Hb = T(0.25) * ((h00 + h10) + (h01 + h11));
            T A = AFIELD ? __ldg(Af + d.off + (long long)b * d.ld + a) : d.A;
            T al, be, gA;
            node_diffusivity<T, CUBIC, true>(ph, A, Hb, gSx * gSx + gSy * gSy, D, al, be, gA);
            // D† = avg_y_adjoint(-Fx† ∘ clamp(dSdx_e)) + avg_x_adjoint(-Fy† ∘ clamp(dSdy_e))
            // with Fx† = diff_x_adjoint(-λ_inn, Δx) = (λ̃[a+1,j] - λ̃[a,j]) / Δx      (adjoint.jl:99-104)
            // The four edges of this node are x-edges (a,b), (a,b+1) and y-edges (a,b), (a+1,b);
            // edges on a border row/column carry λ̃ = 0 on both ends and drop out by themselves.
            T cx0 = clamp_raw<T>(sx0, ph.eta0, h00, h10) * d.inv_dx;
            T cx1 = clamp_raw<T>(sx1, ph.eta0, h01, h11) * d.inv_dx;
            T cy0 = clamp_raw<T>(sy0, ph.eta0, h00, h01) * d.inv_dy;
            T cy1 = clamp_raw<T>(sy1, ph.eta0, h10, h11) * d.inv_dy;
            T fx0 = (l10 - l00) * d.inv_dx, fx1 = (l11 - l01) * d.inv_dx;
            T fy0 = (l01 - l00) * d.inv_dy, fy1 = (l11 - l10) * d.inv_dy;
            T Dadj = -T(0.5) * ((fx0 * cx0 + fx1 * cx1) + (fy0 * cy0 + fy1 * cy1));
            aD = al * Dadj;
            P = be * gSx * Dadj;
            Q = be * gSy * Dadj;
            if (WRITE_S) {
                // own the node iff its lower-left cell (a,b) lies inside this tile's cell range
                if (lx >= 1 && ly >= 1) {
                    T v = gA * Dadj;
                    acc += (double)v;
                    if (vjpA != nullptr) vjpA[d.off + (long long)b * d.ld + a] = v;
                }
            }
        }
        sD[ly][lx] = D;
        sA[ly][lx] = aD;
        sP[ly][lx] = P;
        sQ[ly][lx] = Q;
    }
    __syncthreads();

    if (WRITE_H) {
        const int tx = threadIdx.x & 31, tr = threadIdx.x >> 5;
        const int i = x0 + tx;
        const T e_dx = ph.eta0 * d.inv_dx, e_dy = ph.eta0 * d.inv_dy;
#pragma unroll
        for (int rr = 0; rr < TY / 8; ++rr) {
            int r = tr + rr * 8;
            int j = y0 + r;
            if (i < d.nx && j < d.ny) {
                const int cx = tx + 1, cy = r + 1;
                T hc = sH[cy][cx];
                T res = T(0);
                if (hc > T(0)) {  // adjoint.jl:148
                    T hw = sH[cy][cx - 1], he = sH[cy][cx + 1], hs = sH[cy - 1][cx], hn = sH[cy + 1][cx];
                    T bc = sB[cy][cx], bw = sB[cy][cx - 1], be = sB[cy][cx + 1], bs = sB[cy - 1][cx], bn = sB[cy + 1][cx];
                    T lc = sL[cy][cx], lw = sL[cy][cx - 1], le = sL[cy][cx + 1], ls = sL[cy - 1][cx], ln = sL[cy + 1][cx];
                    // nodes (i-1,j-1) (i,j-1) (i-1,j) (i,j)
                    T d00 = sD[r][tx], d10 = sD[r][tx + 1], d01 = sD[r + 1][tx], d11 = sD[r + 1][tx + 1];
                    T p00 = sP[r][tx], p10 = sP[r][tx + 1], p01 = sP[r + 1][tx], p11 = sP[r + 1][tx + 1];
                    T q00 = sQ[r][tx], q10 = sQ[r][tx + 1], q01 = sQ[r + 1][tx], q11 = sQ[r + 1][tx + 1];
                    // first term (adjoint.jl:123-127)
                    T t1 = T(0.25) * ((sA[r][tx] + sA[r][tx + 1]) + (sA[r + 1][tx] + sA[r + 1][tx + 1]));
                    t1 += (T(0.5) * (p00 + p01) - T(0.5) * (p10 + p11)) * d.inv_dx;  // diff_x_adjoint(avg_y_adjoint(βx D†))
                    t1 += (T(0.5) * (q00 + q10) - T(0.5) * (q01 + q11)) * d.inv_dy;  // diff_y_adjoint(avg_x_adjoint(βy D†))
                    // second term (adjoint.jl:130-144): ∂C = -F† ∘ D_edge through the clamp sub-gradient
                    // (inversion_utils.jl:22-29, 36-43; strict inequalities).
                    T t2 = T(0);
                    {   // west x-edge (i-1,j): this cell is the UPPER cell
                        T dS = sdiff<T>(bc, bw, hc, hw);
                        T dC = -((lc - lw) * d.inv_dx) * (T(0.5) * (d00 + d01));
                        T up = ph.eta0 * hc, lo = -(ph.eta0 * hw);
                        if (gt_div(up, dS, d.dx) && gt_div(dS, lo, d.dx)) t2 += dC * d.inv_dx;   // +∂dS[i-1]/Δx
                        if (gt_div(dS, up, d.dx)) t2 += e_dx * dC;
                    }
                    {   // east x-edge (i,j): this cell is the LOWER cell
                        T dS = sdiff<T>(be, bc, he, hc);
                        T dC = -((le - lc) * d.inv_dx) * (T(0.5) * (d10 + d11));
                        T up = ph.eta0 * he, lo = -(ph.eta0 * hc);
                        if (gt_div(up, dS, d.dx) && gt_div(dS, lo, d.dx)) t2 -= dC * d.inv_dx;   // -∂dS[i]/Δx
                        if (gt_div(lo, dS, d.dx)) t2 -= e_dx * dC;
                    }
                    {   // south y-edge (i,j-1): UPPER cell
                        T dS = sdiff<T>(bc, bs, hc, hs);
                        T dC = -((lc - ls) * d.inv_dy) * (T(0.5) * (d00 + d10));
                        T up = ph.eta0 * hc, lo = -(ph.eta0 * hs);
                        if (gt_div(up, dS, d.dy) && gt_div(dS, lo, d.dy)) t2 += dC * d.inv_dy;
                        if (gt_div(dS, up, d.dy)) t2 += e_dy * dC;
                    }
                    {   // north y-edge (i,j): LOWER cell
                        T dS = sdiff<T>(bn, bc, hn, hc);
                        T dC = -((ln - lc) * d.inv_dy) * (T(0.5) * (d01 + d11));
                        T up = ph.eta0 * hn, lo = -(ph.eta0 * hc);
                        if (gt_div(up, dS, d.dy) && gt_div(dS, lo, d.dy)) t2 -= dC * d.inv_dy;
                        if (gt_div(lo, dS, d.dy)) t2 -= e_dy * dC;
                    }
                    res = t1 + t2;
                }
                out[d.off + (long long)j * d.ld + i] = res;
            }
        }
    }
    if (WRITE_S) {
        double s = block_sum(acc, sRed);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}

// Second stage of the A2 reduction: one CTA per glacier sums its tiles' partials in a fixed order.
__global__ void __launch_bounds__(NT)
reduce_tiles_kernel(const int* __restrict__ tile_start, const double* __restrict__ partial, double* __restrict__ S) {
    __shared__ double sRed[NT / 32];
    int g = blockIdx.x;
    int t0 = tile_start[g], t1 = tile_start[g + 1];
    double acc = 0.0;
    for (int t = t0 + threadIdx.x; t < t1; t += NT) acc += partial[t];
    double s = block_sum(acc, sRed);
    if (threadIdx.x == 0) S[g] = s;
}

}  // namespace odinn
