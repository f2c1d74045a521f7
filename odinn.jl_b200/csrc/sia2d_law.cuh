// Per-cell MLP laws: LawU (SIA2D_D_target, D = H̄·U) and LawY (SIA2D_D_hybrid_target), forward + partials + θ-pullback.
//
// Reference (ODINN.jl v1.1.0):
//   LawU f!            src/laws/Laws.jl:97-123            U[i,j] = post(NN(pre([H̄[i,j], ∇S[i,j]]); θ.U))
//   LawY f!            src/laws/Laws.jl:240-273           Y[i,j] = post(NN(pre([T, H̄[i,j]]); θ.Y))
//   pre / post         src/models/target/target_utils.jl:86-93 (post = max_NN·exp((y-1)/y)), :131-141 (pre = (x-m)/(M-m) - ½)
//   SIA2D_D_target     src/models/target/target_D_pure.jl:78-96 (D), :98-137 (α, β: central differences, δH = 1e-4,
//                      δ∇S = 1e-6; β is ∂D/∂|∇S|, NOT divided by |∇S| -- followed as is), :139-199 (∂D/∂θ, interpolation = :None)
//   SIA2D_D_hybrid     src/models/target/target_D_hybrid.jl:22-96 (D, α: one-sided difference δH = 1e-4, β analytic), :98-166 (∂D/∂θ)
//   contraction        src/inverse/SIA2D/adjoint.jl:235-250   ∂θ_k = Σ_ij (∂D/∂θ_k)[i,j]·D†[i,j]
//
// Three passes per evaluation (the law is compute-bound -- ~0.7 kflop per node and network evaluation for a 2-16-16-1
// network -- so the node planes' extra HBM traffic, 1-4 words per node, is not what limits it):
//   1. law_nodes_kernel   one thread per dual node: H̄, |∇S| from the 2x2 cells, the network, D (and α, β) -> node planes;
//   2. the marching stencil kernels in DFIELD mode (sia2d_march.cuh) consume D (α, β) and, for the θ-VJP, emit D†;
//   3. law_theta_kernel   one thread per node: back-propagation through the network weighted by D†·s, warp-shuffle +
//                         fixed-order block reduction into per-block partials, summed in tile order by
//                         law_theta_reduce_scaled (capi.cu) -- deterministic.
// The network is evaluated in fp64 whenever differences of it are taken (α, β: a 1e-6 step is below fp32 resolution of
// D) and for the pullback; the forward-only F1 evaluates it in the ensemble's precision.
#pragma once
#include "timeloop.cuh"

namespace odinn {

constexpr int LAW_MAX_WIDTH = 32;
constexpr int LAW_NT = 128;
enum { LAW_NONE = 0, LAW_U = 1, LAW_Y = 2 };

struct CellLaw {
    int kind;              // LAW_U | LAW_Y
    MlpArch arch;
    int prescale;          // inputs normalised with the bounds below
    double lo0, hi0, lo1, hi1;
    int postscale;         // y -> max_NN·exp((y-1)/y)
    double max_NN;
    double n_H, n_gS;      // LawY exponents (default n)
    double Gam, Sl, p, q;  // Γ_noA, S (target_utils.jl:3-19)
};

template <typename R> __device__ __forceinline__ R r_exp(R x);
template <> __device__ __forceinline__ float r_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double r_exp<double>(double x) { return exp(x); }
template <typename R> __device__ __forceinline__ R r_log1p(R x);
template <> __device__ __forceinline__ float r_log1p<float>(float x) { return log1pf(x); }
template <> __device__ __forceinline__ double r_log1p<double>(double x) { return log1p(x); }
template <typename R> __device__ __forceinline__ R r_tanh(R x);
template <> __device__ __forceinline__ float r_tanh<float>(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double r_tanh<double>(double x) { return tanh(x); }
template <typename R> __device__ __forceinline__ R r_pow(R x, R y);
template <> __device__ __forceinline__ float r_pow<float>(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double r_pow<double>(double x, double y) { return pow(x, y); }

template <typename R>
__device__ __forceinline__ R act_f(int a, R z) {
    switch (a) {
        case ACT_SOFTPLUS: return r_log1p<R>(r_exp<R>(-(z < R(0) ? -z : z))) + (z > R(0) ? z : R(0));
        case ACT_SIGMOID: return R(1) / (R(1) + r_exp<R>(-z));
        case ACT_TANH: return r_tanh<R>(z);
        case ACT_RELU: return z > R(0) ? z : R(0);
        default: return z;
    }
}

// Dense chain on a two-component input; θ (shared memory, precision R) in Lux layout [vec(W) col-major out×in; b] per layer.
// TAPE keeps pre-activations z and activations a of every layer for the backward pass.
template <typename R, bool TAPE>
__device__ __forceinline__ R mlp2_forward(const MlpArch& arch, const R* __restrict__ th, R x0, R x1, R (*a)[LAW_MAX_WIDTH],
                                          R (*z)[LAW_MAX_WIDTH]) {
    R cur[LAW_MAX_WIDTH], nxt[LAW_MAX_WIDTH];
    cur[0] = x0;
    cur[1] = x1;
    if (TAPE) { a[0][0] = x0; a[0][1] = x1; }
    int k = 0;
    for (int L = 0; L < arch.n_layers; ++L) {
        const int ni = arch.widths[L], no = arch.widths[L + 1];
        const R* W = th + k;
        const R* bv = th + k + no * ni;
        for (int o = 0; o < no; ++o) {
            R s = bv[o];
            for (int i = 0; i < ni; ++i) s += W[o + i * no] * cur[i];
            R y = act_f<R>(arch.acts[L], s);
            nxt[o] = y;
            if (TAPE) { z[L][o] = s; a[L + 1][o] = y; }
        }
        for (int o = 0; o < no; ++o) cur[o] = nxt[o];
        k += no * ni + no;
    }
    return cur[0];
}

template <typename R>
__device__ __forceinline__ R law_pre(const CellLaw& lw, int which, R x) {
    if (!lw.prescale) return x;
    const R lo = (R)(which ? lw.lo1 : lw.lo0), hi = (R)(which ? lw.hi1 : lw.hi0);
    return (x - lo) / (hi - lo) - R(0.5);
}
template <typename R>
__device__ __forceinline__ R law_post(const CellLaw& lw, R y) {
    return lw.postscale ? (R)lw.max_NN * r_exp<R>((y - R(1)) / y) : y;
}

// U(H̄, ∇S)  (LawU)   /   Y(T, H̄)  (LawY)
template <typename R>
__device__ __forceinline__ R law_eval(const CellLaw& lw, const R* th, R in0, R in1) {
    R y = mlp2_forward<R, false>(lw.arch, th, law_pre<R>(lw, 0, in0), law_pre<R>(lw, 1, in1), nullptr, nullptr);
    return law_post<R>(lw, y);
}

template <typename R>
__device__ __forceinline__ R hybrid_D(const CellLaw& lw, R Y, R Hb, R gS) {
    R d = Y * (R)lw.Gam * r_pow<R>(Hb, (R)lw.n_H + R(2)) * r_pow<R>(gS, (R)lw.n_gS - R(1));
    if (lw.Sl != 0.0) d += (R)lw.Sl * r_pow<R>(Hb, (R)(lw.p - lw.q) + R(1)) * r_pow<R>(gS, (R)lw.p - R(1));
    return d;
}

// H̄ and |∇S| at dual node (a, b) exactly as the forward recompute forms them (adjoint.jl:52-67)
template <typename T>
__device__ __forceinline__ void node_inputs(const GDesc<T>& d, const T* __restrict__ H, const T* __restrict__ B, int a, int b,
                                            double& Hb, double& gS) {
    const long long p = d.off + (long long)b * d.ld + a;
    const double h00 = fmax((double)__ldg(H + p), 0.0), h10 = fmax((double)__ldg(H + p + 1), 0.0);
    const double h01 = fmax((double)__ldg(H + p + d.ld), 0.0), h11 = fmax((double)__ldg(H + p + d.ld + 1), 0.0);
    const double s00 = (double)__ldg(B + p) + h00, s10 = (double)__ldg(B + p + 1) + h10;
    const double s01 = (double)__ldg(B + p + d.ld) + h01, s11 = (double)__ldg(B + p + d.ld + 1) + h11;
    const double gx = 0.5 * ((s10 - s00) + (s11 - s01)) / (double)d.dx;
    const double gy = 0.5 * ((s01 - s00) + (s11 - s10)) / (double)d.dy;
    Hb = 0.25 * (h00 + h10 + h01 + h11);
    gS = sqrt(gx * gx + gy * gy);
}

// Pass 1.  One thread per dual node of one tile.  R: precision of the network in the forward-only case.
template <typename T, typename R, bool PARTIALS>
__global__ void __launch_bounds__(LAW_NT)
law_nodes_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, CellLaw lw,
                 const double* __restrict__ theta, const T* __restrict__ H, const T* __restrict__ B, T* __restrict__ Dn,
                 T* __restrict__ Al, T* __restrict__ Be) {
    extern __shared__ __align__(16) unsigned char law_smem[];
    R* th = reinterpret_cast<R*>(law_smem);
    for (int k = threadIdx.x; k < lw.arch.n_params; k += LAW_NT) th[k] = (R)theta[k];
    __syncthreads();
    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    for (int c = threadIdx.x; c < TX * TY; c += LAW_NT) {
        const int a = x0 + (c % TX), b = y0 + (c / TX);
        if (a > d.nx - 2 || b > d.ny - 2) continue;
        double Hb, gS;
        node_inputs<T>(d, H, B, a, b, Hb, gS);
        const long long pn = d.off + (long long)b * d.ld + a;
        if (lw.kind == LAW_U) {
            const R U = law_eval<R>(lw, th, (R)Hb, (R)gS);
            Dn[pn] = (T)((R)Hb * U);                                           // target_D_pure.jl:78-96
            if (PARTIALS) {
                const double dH = 1e-4, dg = 1e-6;                             // target_D_pure.jl:105-137
                const double Dp = (double)law_eval<R>(lw, th, (R)(Hb + dH), (R)gS) * (Hb + dH);
                const double Dm = (double)law_eval<R>(lw, th, (R)(Hb - dH), (R)gS) * (Hb - dH);
                Al[pn] = (T)((Hb > 0.0 ? 1.0 : 0.0) * (Dp - Dm) / (2.0 * dH));
                const double Gp = (double)law_eval<R>(lw, th, (R)Hb, (R)(gS + dg)) * Hb;
                const double Gm = (double)law_eval<R>(lw, th, (R)Hb, (R)(gS - dg)) * Hb;
                Be[pn] = (T)((Gp - Gm) / (2.0 * dg));
            }
        } else {
            const R Tg = (R)d.temp;
            const R Y = law_eval<R>(lw, th, Tg, (R)Hb);
            Dn[pn] = (T)hybrid_D<R>(lw, Y, (R)Hb, (R)gS);                      // target_D_hybrid.jl:22-45
            if (PARTIALS) {
                const double dH = 1e-4;                                        // target_D_hybrid.jl:58-73
                const double pq = lw.p - lw.q;
                double noNN = (lw.n_H + 2.0) * (double)Y * lw.Gam * pow(Hb, lw.n_H + 1.0) * pow(gS, lw.n_gS - 1.0);
                if (lw.Sl != 0.0) noNN += (pq + 1.0) * lw.Sl * pow(Hb, pq) * pow(gS, lw.p - 1.0);
                const double Da = (double)hybrid_D<R>(lw, law_eval<R>(lw, th, Tg, (R)(Hb + dH)), (R)Hb, (R)gS);
                const double Db = (double)hybrid_D<R>(lw, Y, (R)Hb, (R)gS);
                Al[pn] = (T)(noNN + (Da - Db) / dH);
                double be = lw.Gam * (double)Y * (lw.n_gS - 1.0) * pow(Hb, lw.n_H + 2.0) * pow(gS, lw.n_gS - 3.0);
                if (lw.Sl != 0.0) be += lw.Sl * (lw.p - 1.0) * pow(Hb, pq + 1.0) * pow(gS, lw.p - 3.0);
                Be[pn] = (T)be;                                                // target_D_hybrid.jl:75-96
            }
        }
    }
}

// Pass 3.  ∂θ_k = Σ_nodes D†·s·post'(y)·∂NN/∂θ_k ,  s = H̄·[H̄ > 0] (LawU, nodes with H̄ == 0 skipped: target_D_pure.jl:169-171)
// or Γ H̄^{n_H+2} ∇S^{n_∇S-1} (LawY).  One block per tile; block_partial[block][k] written in full (fixed order).
template <typename T>
__global__ void __launch_bounds__(LAW_NT)
law_theta_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, CellLaw lw,
                 const double* __restrict__ theta, const T* __restrict__ H, const T* __restrict__ B,
                 const T* __restrict__ Dadj, double* __restrict__ block_partial) {
    extern __shared__ __align__(16) unsigned char law_smem[];
    double* th = reinterpret_cast<double*>(law_smem);                 // n_params
    double* wacc = th + lw.arch.n_params;                             // (LAW_NT/32) x n_params per-warp accumulators
    const int np = lw.arch.n_params;
    for (int k = threadIdx.x; k < np; k += LAW_NT) th[k] = theta[k];
    for (int k = threadIdx.x; k < (LAW_NT / 32) * np; k += LAW_NT) wacc[k] = 0.0;
    __syncthreads();
    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* my = wacc + warp * np;
    double a[MLP_MAX_LAYERS + 1][LAW_MAX_WIDTH], z[MLP_MAX_LAYERS][LAW_MAX_WIDTH];
    for (int c0 = 0; c0 < TX * TY; c0 += LAW_NT) {                     // all lanes iterate together (warp reductions below)
        const int c = c0 + threadIdx.x;
        const int an = x0 + (c % TX), bn = y0 + (c / TX);
        double w = 0.0;
        if (an <= d.nx - 2 && bn <= d.ny - 2) {
            double Hb, gS;
            node_inputs<T>(d, H, B, an, bn, Hb, gS);
            const double in0 = lw.kind == LAW_U ? Hb : (double)d.temp, in1 = lw.kind == LAW_U ? gS : Hb;
            const double y = mlp2_forward<double, true>(lw.arch, th, law_pre<double>(lw, 0, in0), law_pre<double>(lw, 1, in1), a, z);
            const double dpost = lw.postscale ? lw.max_NN * exp((y - 1.0) / y) / (y * y) : 1.0;
            double s;
            if (lw.kind == LAW_U) s = (Hb > 0.0) ? Hb : 0.0;
            else s = lw.Gam * pow(Hb, lw.n_H + 2.0) * pow(gS, lw.n_gS - 1.0);
            w = (double)__ldg(Dadj + d.off + (long long)bn * d.ld + an) * s * dpost;
            if (!(w == w)) w = 0.0;  // 0·inf from a degenerate node (H̄ = 0 with a negative exponent) contributes nothing
        }
        // backward through the chain with output cotangent w
        double g[LAW_MAX_WIDTH], gp[LAW_MAX_WIDTH];
        g[0] = w;
        int k = np;
        for (int L = lw.arch.n_layers - 1; L >= 0; --L) {
            const int ni = lw.arch.widths[L], no = lw.arch.widths[L + 1];
            k -= no * ni + no;
            const double* W = th + k;
            for (int i = 0; i < ni; ++i) gp[i] = 0.0;
            for (int o = 0; o < no; ++o) {
                const double dz = (w != 0.0) ? g[o] * act_bwd(lw.arch.acts[L], z[L][o], a[L + 1][o]) : 0.0;
                double rb = dz;                                         // bias gradient
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) rb += __shfl_down_sync(0xffffffffu, rb, sft);
                if (lane == 0) my[k + no * ni + o] += rb;
                for (int i = 0; i < ni; ++i) {
                    double rw = (w != 0.0) ? dz * a[L][i] : 0.0;        // weight gradient, vec(W) column-major
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) rw += __shfl_down_sync(0xffffffffu, rw, sft);
                    if (lane == 0) my[k + o + i * no] += rw;
                    gp[i] += (w != 0.0) ? W[o + i * no] * dz : 0.0;
                }
            }
            for (int i = 0; i < ni; ++i) g[i] = gp[i];
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < np; k += LAW_NT) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < LAW_NT / 32; ++wv) s += wacc[wv * np + k];
        block_partial[(long long)blockIdx.x * np + k] = s;
    }
}

}  // namespace odinn
