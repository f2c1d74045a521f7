// Kernels around the stencil: loss / adjoint seed / λ update of the reverse loop, per-glacier law for A.
//
// Reference semantics (ODINN.jl v1.1.0):
//   LossH(L2Sum)          src/losses/Losses.jl:116-152, 250-291   ℓ_j = Δt_j Σ mask (H_j - H_ref,j)² / (nx ny)
//   reverse (Euler) loop  src/inverse/SIA2D/gradient.jl:191-253   λ_{j-1} = λ_j + Δt_{j-1} VJP_H + ∂ℓ_j/∂H
//   LawA(nn, params)      src/laws/Laws.jl:323-386                A = minA + (maxA-minA)·NN([T]; θ)
//   MLP layout            src/models/trainable_components/ML_utils.jl:23-65 (Lux Dense: W[out×in] col-major, b[out])
#pragma once
#include "common.cuh"

namespace odinn {

//   partial = Σ W (H - Href)²        (W = mask / (nx ny), uploaded pre-divided)
//   lam_out = lam_in + dt·v + cseed·W·(H - Href)  when lam_out != nullptr   (gradient.jl:242, Losses.jl:270-291)
// One pass over the padded planes with 16-byte vector accesses (grid: chunks x glaciers): padding columns carry W = 0 and
// λ = v = 0, so they add nothing to the loss and stay zero in λ_out.  partial[glacier * gridDim.x + chunk].
template <typename T> struct LsVec;
template <> struct LsVec<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct LsVec<double> { typedef double2 type; static constexpr int N = 2; };
constexpr int LS_UNROLL = 4;
template <typename T> __device__ __forceinline__ void ls_unpack(const typename LsVec<T>::type& v, T* x);
template <> __device__ __forceinline__ void ls_unpack<float>(const float4& v, float* x) { x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
template <> __device__ __forceinline__ void ls_unpack<double>(const double2& v, double* x) { x[0] = v.x; x[1] = v.y; }
template <typename T> __device__ __forceinline__ typename LsVec<T>::type ls_pack(const T* x);
template <> __device__ __forceinline__ float4 ls_pack<float>(const float* x) { return make_float4(x[0], x[1], x[2], x[3]); }
template <> __device__ __forceinline__ double2 ls_pack<double>(const double* x) { return make_double2(x[0], x[1]); }

template <typename T>
__global__ void __launch_bounds__(NT)
loss_seed_vec_kernel(const GDesc<T>* __restrict__ descs, const T* __restrict__ H, const T* __restrict__ Href, const T* __restrict__ W,
                     const T* lam_in, const T* __restrict__ v, T* lam_out, double* __restrict__ partial, T dt, T cseed) {
    typedef typename LsVec<T>::type V;
    constexpr int N = LsVec<T>::N;
    __shared__ double sRed[NT / 32];
    const GDesc<T> d = descs[blockIdx.y];
    const long long nvec = (long long)d.ld * d.ny / N, base = d.off / N;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < LS_UNROLL; ++u) {
        const long long q = ((long long)blockIdx.x * LS_UNROLL + u) * NT + threadIdx.x;
        if (q < nvec) {
            const long long p = base + q;
            T h[N], r[N], w[N], l[N], vv[N], o[N];
            ls_unpack<T>(reinterpret_cast<const V*>(H)[p], h);
            ls_unpack<T>(reinterpret_cast<const V*>(Href)[p], r);
            ls_unpack<T>(reinterpret_cast<const V*>(W)[p], w);
            if (lam_out != nullptr) {
                if (lam_in) ls_unpack<T>(reinterpret_cast<const V*>(lam_in)[p], l);
                if (v) ls_unpack<T>(reinterpret_cast<const V*>(v)[p], vv);
            }
#pragma unroll
            for (int e = 0; e < N; ++e) {
                const T diff = h[e] - r[e];
                acc += (double)(w[e] * diff * diff);
                if (lam_out != nullptr) o[e] = (lam_in ? l[e] : T(0)) + dt * (v ? vv[e] : T(0)) + cseed * (w[e] * diff);
            }
            if (lam_out != nullptr) reinterpret_cast<V*>(lam_out)[p] = ls_pack<T>(o);
        }
    }
    double s = block_sum(acc, sRed);
    if (threadIdx.x == 0) partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = s;
}

// out[g] = (accumulate ? out[g] : 0) + scale * Σ_chunks partial[g][chunk]   (fixed order)
static __global__ void __launch_bounds__(NT)
reduce_chunks_scaled_kernel(const double* __restrict__ partial, int nchunk, double* __restrict__ out, double scale, int accumulate) {
    __shared__ double sRed[NT / 32];
    double acc = 0.0;
    for (int c = threadIdx.x; c < nchunk; c += NT) acc += partial[(long long)blockIdx.x * nchunk + c];
    double s = block_sum(acc, sRed);
    if (threadIdx.x == 0) out[blockIdx.x] = (accumulate ? out[blockIdx.x] : 0.0) + scale * s;
}

// out[g] = (accumulate ? out[g] : 0) + scale * Σ partial[start[g] .. start[g+1])   (fixed order, bit-stable)
static __global__ void __launch_bounds__(NT)
reduce_scaled_kernel(const int* __restrict__ start, const double* __restrict__ partial, double* __restrict__ out,
                     double scale, int accumulate) {
    __shared__ double sRed[NT / 32];
    int g = blockIdx.x;
    int t0 = start[g], t1 = start[g + 1];
    double acc = 0.0;
    for (int t = t0 + threadIdx.x; t < t1; t += NT) acc += partial[t];
    double s = block_sum(acc, sRed);
    if (threadIdx.x == 0) out[g] = (accumulate ? out[g] : 0.0) + scale * s;
}

// ---- per-glacier creep law  A_g = minA + (maxA - minA) · NN([T_g]; θ)  and its pullback ∂A_g/∂θ ----------

constexpr int MLP_MAX_LAYERS = 8;
constexpr int MLP_MAX_WIDTH = 64;
enum { ACT_IDENTITY = 0, ACT_SOFTPLUS = 1, ACT_SIGMOID = 2, ACT_TANH = 3, ACT_RELU = 4 };

struct MlpArch {
    int n_layers;                       // number of Dense layers
    int widths[MLP_MAX_LAYERS + 1];     // n_in, h1, ..., n_out
    int acts[MLP_MAX_LAYERS];
    int n_params;
};

__host__ __device__ inline double act_fwd(int a, double z) {
    switch (a) {
        case ACT_SOFTPLUS: return log1p(exp(-fabs(z))) + (z > 0 ? z : 0.0);  // NNlib.softplus, overflow-safe
        case ACT_SIGMOID: return 1.0 / (1.0 + exp(-z));
        case ACT_TANH: return tanh(z);
        case ACT_RELU: return z > 0 ? z : 0.0;
        default: return z;
    }
}
// derivative given pre-activation z and activation value y
__host__ __device__ inline double act_bwd(int a, double z, double y) {
    switch (a) {
        case ACT_SOFTPLUS: return 1.0 / (1.0 + exp(-z));
        case ACT_SIGMOID: return y * (1.0 - y);
        case ACT_TANH: return 1.0 - y * y;
        case ACT_RELU: return z > 0 ? 1.0 : 0.0;
        default: return 1.0;
    }
}

// One thread per glacier (the law is evaluated once per glacier and solve: callback_freq = 0, Laws.jl:346).
// Double precision regardless of the ensemble dtype.  J is [G x n_params], row-major.
static __global__ void law_A_nn_kernel(MlpArch arch, const double* __restrict__ theta, const double* __restrict__ temps,
                                int G, double minA, double maxA, double* __restrict__ A_out, double* __restrict__ J) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double a[MLP_MAX_LAYERS + 1][MLP_MAX_WIDTH];
    double z[MLP_MAX_LAYERS][MLP_MAX_WIDTH];
    a[0][0] = temps[g];
    int k = 0;
    for (int L = 0; L < arch.n_layers; ++L) {
        int ni = arch.widths[L], no = arch.widths[L + 1];
        const double* Wm = theta + k;            // W[o + i*no]  (vec(W), column-major out×in)
        const double* bv = theta + k + no * ni;
        for (int o = 0; o < no; ++o) {
            double s = bv[o];
            for (int i = 0; i < ni; ++i) s += Wm[o + i * no] * a[L][i];
            z[L][o] = s;
            a[L + 1][o] = act_fwd(arch.acts[L], s);
        }
        k += no * ni + no;
    }
    double y = a[arch.n_layers][0];
    A_out[g] = minA + (maxA - minA) * y;  // scale(), target_utils.jl:109-113
    if (J == nullptr) return;
    // backprop with output cotangent (maxA - minA)
    double gvec[MLP_MAX_WIDTH], gprev[MLP_MAX_WIDTH];
    gvec[0] = (maxA - minA);
    double* Jg = J + (long long)g * arch.n_params;
    for (int L = arch.n_layers - 1; L >= 0; --L) {
        int ni = arch.widths[L], no = arch.widths[L + 1];
        k -= no * ni + no;
        const double* Wm = theta + k;
        for (int i = 0; i < ni; ++i) gprev[i] = 0.0;
        for (int o = 0; o < no; ++o) {
            double dz = gvec[o] * act_bwd(arch.acts[L], z[L][o], a[L + 1][o]);
            Jg[k + no * ni + o] = dz;
            for (int i = 0; i < ni; ++i) {
                Jg[k + o + i * no] = dz * a[L][i];
                gprev[i] += Wm[o + i * no] * dz;
            }
        }
        for (int i = 0; i < ni; ++i) gvec[i] = gprev[i];
    }
}

// dθ[k] = Σ_g J[g,k] · S[g]   (glaciers in fixed order: aggregate∇θ, Model.jl:208-224)
static __global__ void law_pullback_kernel(const double* __restrict__ J, const double* __restrict__ S, int G, int n_params,
                                    double* __restrict__ dtheta) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_params) return;
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += J[(long long)g * n_params + k] * S[g];
    dtheta[k] = s;
}

template <typename T>
__global__ void set_A_kernel(GDesc<T>* descs, const double* __restrict__ A, int G) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) descs[g].A = (T)A[g];
}

}  // namespace odinn
