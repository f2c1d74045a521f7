// Shared device-side types of the SIA2D hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace odinn {

// Tile of cells one CTA produces.  x is the contiguous axis (Julia column-major `Matrix`).
constexpr int TX = 32;
constexpr int TY = 16;
constexpr int NT = 256;  // threads per CTA: 32 x 8, each thread owns TY/8 rows

// Per-glacier descriptor.  Every device plane of the ensemble uses the same element offset
// `off` and pitch `ld` (a multiple of 32 elements, so rows start 128 B aligned).
template <typename T>
struct GDesc {
    long long off;
    int nx, ny, ld, tile0;  // tile0: index of this glacier's first tile in the tile table
    T dx, dy, inv_dx, inv_dy;
    T A;     // glacier-wide creep coefficient (cache.iceflow.A.value, ScalarCache)
    T temp;  // long-term air temperature fed to the A law (Laws.jl:348-358)
};

// params.physical folded into the constants the kernels use
// (src/models/target/target_utils.jl:3-19).
template <typename T>
struct PhysDev {
    T n, p, q;
    T Gam;   // Γ_noA = 2 (ρ g)^n / (n + 2)
    T Sl;    // S     = C (ρ g)^(p - q)
    T eta0;  // η₀
};


// Per-glacier controller state of the RDPK3Sp35 engine (rdpk.cu).  The fused stage epilogue of the F1 kernels reads the step h.
struct RkState {
    double t, tstop, dt, h, EEst;
    double err2, err3;      // PID history: 1 / EEst of the last two accepted steps
    double sk0, sk1;        // scratch of the initial-step algorithm (d0, d1)
    int last, accept, done;
    int steps, rejected;
};

// One RDPK3Sp35 (3S*+) stage fused into the epilogue of the F1 kernels (rdpk.cu): with k = SIA2D(S1) at the cell
//   S2 = S2in + d S1 ;   S1new = g1 S1 + g2 S2 + g3 u + (b h_g) k ;   est = est + (e h_g) k
// (RKF_FIRST: S1new = S1 + (b h_g) k, est = (e h_g) k -- the stage that starts a trial step from S1 = u), h_g the glacier's own step.
// S1new goes to a second plane (the neighbours still read S1); S2 and est are updated in place.  RKF_NORM (last stage): the pass also
// reduces  sum (est / (abstol + reltol max(|u|, |S1new|)))^2  per work item, so that neither the error plane nor a norm pass is needed.
enum { RKF_FIRST = 1, RKF_U = 2, RKF_WS2 = 4, RKF_WEST = 8, RKF_NORM = 16, RKF_LERP_ONLY = 32 };
// The combinations the scheme uses, as compile-time modes of the F1 kernels: first stage; stages without / with the g3 u term (both update
// S2 and est in place); last stage (g3 u term, error norm, neither S2 nor est written).
enum { RKM_NONE = 0, RKM_FIRST = 1, RKM_MID = 2, RKM_MID_U = 3, RKM_LAST = 4 };
inline int rk_mode_of_flags(int flags) {
    if (flags & RKF_FIRST) return RKM_FIRST;
    if (flags & RKF_NORM) return RKM_LAST;
    return (flags & RKF_U) ? RKM_MID_U : RKM_MID;
}
template <typename T>
struct RkFuse {
    const RkState* st;
    const T* S2in;
    T* S2out;
    T* est;
    const T* u;
    T g1, g2, g3, d;
    double b, e;        // multiplied by the glacier's h in double, as the elementwise stage kernels do
    T reltol, abstol;
    int flags;
};

// Programmatic dependent launch (see launch_pdl, launch.cuh): pdl_trigger lets the next kernel of the stream start its prologue,
// pdl_wait blocks until the previous kernel has completed and its writes are visible (no-ops in a plain launch).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__host__ __device__ inline int div_up(int a, int b) { return (a + b - 1) / b; }

template <typename T>
__device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T>
__device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }

__device__ __forceinline__ float tsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double tsqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float tpow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double tpow(double x, double y) { return pow(x, y); }

// Diffusivity at one dual-grid node and, when PARTIALS, its two partials and the θ-integrand.
//   D  = S H̄^{p-q+1} ∇S^{p-1} + A Γ H̄^{n+2} ∇S^{n-1}          (target_A.jl:16-30)
//   α  = ∂D/∂H̄                                                  (target_A.jl:32-46)
//   β  = (1/∇S) ∂D/∂∇S                                           (target_A.jl:48-62)
//   gA = Γ H̄^{n+2} ∇S^{n-1}  (∂A_spatial)                        (target_A.jl:71-72)
// CUBIC: n == 3 and C == 0, every power is integral and no sqrt is needed.
template <typename T, bool CUBIC, bool PARTIALS>
__device__ __forceinline__ void node_diffusivity(const PhysDev<T>& ph, T A, T Hb, T g2, T& D, T& alpha, T& beta,
                                                 T& gA) {
    if (CUBIC) {
        T H2 = Hb * Hb;
        T H4 = H2 * H2;
        T GH4 = ph.Gam * H4;
        gA = GH4 * Hb * g2;
        D = A * gA;
        if (PARTIALS) {
            alpha = T(5) * A * GH4 * g2;
            beta = T(2) * A * GH4 * Hb;
        }
    } else {
        T gS = tsqrt(g2);
        T pq = ph.p - ph.q;
        gA = ph.Gam * tpow(Hb, ph.n + T(2)) * tpow(gS, ph.n - T(1));
        T slide = (ph.Sl != T(0)) ? ph.Sl * tpow(Hb, pq + T(1)) * tpow(gS, ph.p - T(1)) : T(0);
        D = slide + A * gA;
        if (PARTIALS) {
            T a_s = T(0), b_s = T(0);
            if (ph.Sl != T(0)) {
                a_s = (pq + T(1)) * ph.Sl * tpow(Hb, pq) * tpow(gS, ph.p - T(1));
                b_s = ph.Sl * (ph.p - T(1)) * tpow(Hb, pq + T(1)) * tpow(gS, ph.p - T(3));
            }
            alpha = a_s + A * ph.Gam * (ph.n + T(2)) * tpow(Hb, ph.n + T(1)) * tpow(gS, ph.n - T(1));
            beta = b_s + A * ph.Gam * (ph.n - T(1)) * tpow(Hb, ph.n + T(2)) * tpow(gS, ph.n - T(3));
        }
    }
}

// Warp + CTA sum in a fixed order (bit-stable run to run).  Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* smem /* >= NT/32 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        for (int k = 0; k < nw; ++k) s += smem[k];
    }
    return s;
}

}  // namespace odinn
