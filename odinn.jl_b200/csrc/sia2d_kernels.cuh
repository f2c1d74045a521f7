// Shared device helpers of the SIA2D stencil kernels (surface differences, flux clamp, strict comparisons).
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
// (The first, shared-memory tiled implementation of F1/A1/A2 lived here; it was instruction-bound --
//  profiles/r01_v1_tiled_* -- and was replaced by the register-marching kernels of sia2d_march.cuh.)
#pragma once
#include "common.cuh"

namespace odinn {

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

}  // namespace odinn
