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
