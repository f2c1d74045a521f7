// Register-marching F1 / A1+A2 kernels (v2).
//
// The tiled v1 kernels were instruction-issue bound (ncu: 254 / 427 thread-instructions per cell, 84 % / 70 %
// issue-slot utilisation, 15 % / 11 % DRAM).  Here every dual node and every edge is evaluated exactly once:
//   * one warp owns a strip of 32 consecutive columns (lane = column, the contiguous axis, so every row access
//     is one coalesced 128 B / 256 B line) and marches over a chunk of rows;
//   * the 3x3 neighbourhood lives in registers: the previous row is carried, x-neighbours come from
//     warp shuffles (halo exchange), so there is no shared memory and no block barrier at all;
//   * lanes 0 and 31 are halo lanes: a strip produces 30 output columns.
// Work items (glacier, first column, row range) come from a table built at ensemble creation, so ragged
// ensembles are one launch.  Reference semantics and citations: see sia2d_kernels.cuh.
#pragma once
#include "sia2d_kernels.cuh"

namespace odinn {

constexpr int STRIP = 30;        // output columns per warp
constexpr int MARCH_WARPS = 8;   // warps per CTA
constexpr unsigned FULL = 0xffffffffu;

template <typename T>
__device__ __forceinline__ T shfl_dn(T v) { return __shfl_down_sync(FULL, v, 1); }
template <typename T>
__device__ __forceinline__ T shfl_up(T v) { return __shfl_up_sync(FULL, v, 1); }

// --------------------------------------------------------------------------------------------
// F1
// --------------------------------------------------------------------------------------------
template <typename T, bool CUBIC, bool AFIELD>
__global__ void __launch_bounds__(MARCH_WARPS * 32)
sia2d_rhs_march(const GDesc<T>* __restrict__ descs, const int4* __restrict__ items, int n_items,
                const T* __restrict__ H, const T* __restrict__ B, const T* __restrict__ Af, T* __restrict__ dH,
                PhysDev<T> ph) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<T> d = descs[it.x];
    const int i = it.y + lane, r0 = it.z, r1 = it.w;
    const int ic = min(max(i, 0), d.nx - 1);
    const T* Hp = H + d.off + ic;
    const T* Bp = B + d.off + ic;
    const T* Ap = AFIELD ? Af + d.off + min(ic, d.nx - 2) : nullptr;
    T* Op = dH + d.off + ic;
    const bool store_lane = (lane >= 1 && lane <= STRIP && i < d.nx);
    const bool col_inner = (i >= 1 && i <= d.nx - 2);
    const T eta0 = ph.eta0;

    // cell row r0-1
    int rc = max(r0 - 1, 0);
    T h = __ldg(Hp + (long long)rc * d.ld), b = __ldg(Bp + (long long)rc * d.ld);
    h = h > T(0) ? h : T(0);
    b = surf_store<T>(b, h);
    T hE = shfl_dn(h), bE = shfl_dn(b);
    T ex = sdiff<T>(bE, b, hE, h);
    T hx = h + hE;
    T Dp = T(0), Fy = T(0);

    int rn = min(r0, d.ny - 1);
    T h1n = __ldg(Hp + (long long)rn * d.ld), b1n = __ldg(Bp + (long long)rn * d.ld);

    for (int row = r0 - 1; row < r1; ++row) {
        // cell row row+1 (prefetched), then issue the prefetch of row+2
        T h1 = h1n, b1 = b1n;
        {
            int r2 = min(row + 2, d.ny - 1);
            h1n = __ldg(Hp + (long long)r2 * d.ld);
            b1n = __ldg(Bp + (long long)r2 * d.ld);
        }
        h1 = h1 > T(0) ? h1 : T(0);
        b1 = surf_store<T>(b1, h1);
        T hE1 = shfl_dn(h1), bE1 = shfl_dn(b1);
        T ex1 = sdiff<T>(bE1, b1, hE1, h1);
        T hx1 = h1 + hE1;
        T ey = sdiff<T>(b1, b, h1, h);
        T eyE = shfl_dn(ey);
        // node (i, row)
        T gx = T(0.5) * (ex + ex1) * d.inv_dx;
        T gy = T(0.5) * (ey + eyE) * d.inv_dy;
        T Hb = T(0.25) * (hx + hx1);
        T A = d.A;
        if (AFIELD) A = __ldg(Ap + (long long)min(max(row, 0), d.ny - 2) * d.ld);
        T D1, al, be, gA;
        node_diffusivity<T, CUBIC, false>(ph, A, Hb, gx * gx + gy * gy, D1, al, be, gA);
        T D1W = shfl_up(D1);
        // y-edge (i, row -> row+1)
        T Fy1 = -(T(0.5) * (D1W + D1)) * (clamp_raw<T>(ey, eta0, h, h1) * d.inv_dy);
        if (row >= r0) {  // warp-uniform
            T Fx = -(T(0.5) * (Dp + D1)) * (clamp_raw<T>(ex, eta0, h, hE) * d.inv_dx);
            T FxW = shfl_up(Fx);
            T out = -((Fx - FxW) * d.inv_dx + (Fy1 - Fy) * d.inv_dy);
            if (!(col_inner && row >= 1 && row <= d.ny - 2)) out = T(0);
            if (store_lane) Op[(long long)row * d.ld] = out;
        }
        h = h1; b = b1; hE = hE1; ex = ex1; hx = hx1; Dp = D1; Fy = Fy1;
    }
}

// --------------------------------------------------------------------------------------------
// A1 + A2 (see sia2d_vjp_kernel for the math; this is the same arithmetic, marched)
// --------------------------------------------------------------------------------------------
template <typename T, bool CUBIC, bool AFIELD, bool WRITE_H, bool WRITE_S>
__global__ void __launch_bounds__(MARCH_WARPS * 32)
sia2d_vjp_march(const GDesc<T>* __restrict__ descs, const int4* __restrict__ items, int n_items,
                const T* __restrict__ lam, const T* __restrict__ H, const T* __restrict__ B, const T* __restrict__ Af,
                T* __restrict__ out, T* __restrict__ vjpA, double* __restrict__ partial, PhysDev<T> ph) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<T> d = descs[it.x];
    const int i = it.y + lane, r0 = it.z, r1 = it.w;
    const int ic = min(max(i, 0), d.nx - 1);
    const T* Hp = H + d.off + ic;
    const T* Bp = B + d.off + ic;
    const T* Lp = lam + d.off + ic;
    const T* Ap = AFIELD ? Af + d.off + min(ic, d.nx - 2) : nullptr;
    T* Op = WRITE_H ? out + d.off + ic : nullptr;
    const bool store_lane = (lane >= 1 && lane <= STRIP && i < d.nx);
    const bool col_inner = (i >= 1 && i <= d.nx - 2);
    const bool node_col_ok = (i >= 0 && i <= d.nx - 2);
    const bool own_lane = (lane < STRIP);  // nodes i0 .. i0+29 are owned by this strip
    const T eta0 = ph.eta0;
    const T idx2 = d.inv_dx * d.inv_dx, idy2 = d.inv_dy * d.inv_dy;
    const T e_dx = eta0 * d.inv_dx, e_dy = eta0 * d.inv_dy;

    // cell row r0-1
    int rc = max(r0 - 1, 0);
    T h = __ldg(Hp + (long long)rc * d.ld), b = __ldg(Bp + (long long)rc * d.ld), l = __ldg(Lp + (long long)rc * d.ld);
    h = h > T(0) ? h : T(0);
    b = surf_store<T>(b, h);
    if (!(col_inner && (r0 - 1) >= 1 && (r0 - 1) <= d.ny - 2)) l = T(0);
    T hE = shfl_dn(h), bE = shfl_dn(b), lE = shfl_dn(l);
    T ex = sdiff<T>(bE, b, hE, h);
    T hx = h + hE;
    T fxr = lE - l;
    T px = fxr * clamp_raw<T>(ex, eta0, h, hE);
    T Dp = T(0), aDp = T(0), Pp = T(0), Qrow_p = T(0), yu_p = T(0);
    double acc = 0.0;

    int rn = min(r0, d.ny - 1);
    T h1n = __ldg(Hp + (long long)rn * d.ld), b1n = __ldg(Bp + (long long)rn * d.ld), l1n = __ldg(Lp + (long long)rn * d.ld);

    for (int row = r0 - 1; row < r1; ++row) {
        T h1 = h1n, b1 = b1n, l1 = l1n;
        {
            int r2 = min(row + 2, d.ny - 1);
            h1n = __ldg(Hp + (long long)r2 * d.ld);
            b1n = __ldg(Bp + (long long)r2 * d.ld);
            l1n = __ldg(Lp + (long long)r2 * d.ld);
        }
        h1 = h1 > T(0) ? h1 : T(0);
        b1 = surf_store<T>(b1, h1);
        if (!(col_inner && (row + 1) >= 1 && (row + 1) <= d.ny - 2)) l1 = T(0);  // λ_inn zero-extended
        T hE1 = shfl_dn(h1), bE1 = shfl_dn(b1), lE1 = shfl_dn(l1);
        // x-edge (i, row+1)
        T ex1 = sdiff<T>(bE1, b1, hE1, h1);
        T hx1 = h1 + hE1;
        T fxr1 = lE1 - l1;
        T px1 = fxr1 * clamp_raw<T>(ex1, eta0, h1, hE1);
        // y-edge (i, row -> row+1)
        T ey = sdiff<T>(b1, b, h1, h);
        T fyr = l1 - l;
        T py = fyr * clamp_raw<T>(ey, eta0, h, h1);
        T eyE = shfl_dn(ey), pyE = shfl_dn(py);
        // node (i, row)
        T gSx = T(0.5) * (ex + ex1) * d.inv_dx;
        T gSy = T(0.5) * (ey + eyE) * d.inv_dy;
        T Hb = T(0.25) * (hx + hx1);
        T A = d.A;
        if (AFIELD) A = __ldg(Ap + (long long)min(max(row, 0), d.ny - 2) * d.ld);
        T D1, al, be, gA;
        node_diffusivity<T, CUBIC, true>(ph, A, Hb, gSx * gSx + gSy * gSy, D1, al, be, gA);
        T Dadj = -T(0.5) * ((px + px1) * idx2 + (py + pyE) * idy2);
        const bool node_ok = node_col_ok && row >= 0 && row <= d.ny - 2;
        T bD = be * Dadj;
        T aD1 = node_ok ? al * Dadj : T(0);
        T P1 = node_ok ? bD * gSx : T(0);
        T Q1 = node_ok ? bD * gSy : T(0);
        if (WRITE_S) {
            if (node_ok && own_lane && row >= r0) {
                T v = gA * Dadj;
                acc += (double)v;
                if (vjpA != nullptr) vjpA[d.off + (long long)row * d.ld + i] = v;
            }
        }
        if (WRITE_H) {
            T D1W = shfl_up(D1), Q1W = shfl_up(Q1);
            T Qrow1 = Q1W + Q1;
            // y-edge sub-gradient (inversion_utils.jl:36-43): lower cell = row, upper cell = row+1
            T yl, yu1;
            {
                T dC = -(fyr * d.inv_dy) * (T(0.5) * (D1W + D1));
                T up = eta0 * h1, lo = -(eta0 * h);
                bool inside = gt_div(up, ey, d.dy) && gt_div(ey, lo, d.dy);
                T pass = inside ? dC * d.inv_dy : T(0);
                yl = -pass - (gt_div(lo, ey, d.dy) ? e_dy * dC : T(0));
                yu1 = pass + (gt_div(ey, up, d.dy) ? e_dy * dC : T(0));
            }
            if (row >= r0) {  // warp-uniform: output row `row`
                // x-edge sub-gradient (inversion_utils.jl:22-29): lower cell = i, upper cell = i+1
                T dC = -(fxr * d.inv_dx) * (T(0.5) * (Dp + D1));
                T up = eta0 * hE, lo = -(eta0 * h);
                bool inside = gt_div(up, ex, d.dx) && gt_div(ex, lo, d.dx);
                T pass = inside ? dC * d.inv_dx : T(0);
                T xl = -pass - (gt_div(lo, ex, d.dx) ? e_dx * dC : T(0));
                T xu = pass + (gt_div(ex, up, d.dx) ? e_dx * dC : T(0));
                T aDc = T(0.25) * (aDp + aD1);
                T Pc = (T(0.5) * d.inv_dx) * (Pp + P1);
                T ZW = shfl_up(aDc + Pc + xu);  // everything column i-1 sends to cell (i, row)
                T res = ZW + (aDc - Pc + xl) + (T(0.5) * d.inv_dy) * (Qrow_p - Qrow1) + (yl + yu_p);
                if (!(h > T(0))) res = T(0);  // adjoint.jl:148
                if (store_lane) Op[(long long)row * d.ld] = res;
            }
            Qrow_p = Qrow1;
            yu_p = yu1;
        }
        h = h1; b = b1; l = l1; hE = hE1; ex = ex1; hx = hx1; fxr = fxr1; px = px1;
        Dp = D1; aDp = aD1; Pp = P1;
    }
    if (WRITE_S) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL, acc, o);
        if (lane == 0) partial[item] = acc;
    }
}

// Second stage of the A2 reduction: one CTA per glacier sums its work items' partials in a fixed order.
__global__ void __launch_bounds__(NT)
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
