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
//   1. law_nodes_kernel   one thread per dual node: H̄, |∇S| from the 2x2 cells, the network, D (and α, β) -> node planes.
//                         The network runs register-resident (mlp2_regs: compile-time width bound, fully unrolled, LDS.128 weight
//                         quads); LawU's α, β are exact derivatives propagated as forward-mode tangents (one evaluation instead
//                         of the five of a central difference, no fp64 requirement);
//   2. the marching stencil kernels in DFIELD mode (sia2d_march.cuh) consume D (α, β) and, for the θ-VJP, emit D†;
//   3. law_theta_kernel   one thread per node back-propagates through the network weighted by D†·s and stages layer inputs and
//                         output cotangents in shared memory; the weight gradients are then small contractions over the 128
//                         nodes of a pass, one owner thread per parameter (fixed order), per-block partials summed in tile
//                         order by law_theta_reduce_scaled (capi.cu) -- deterministic.
// Precision: the ensemble's, except LawY's α (a one-sided difference of the network: fp64).
// Measured (tools/bench_cell_law.py, 500x500x64 glaciers, LawU 2-16-16-1, fp32): F1 2.4 ms, A1 4.3 ms, A1+A2 14.7 ms
// (first version, finite differences + shuffle reduction: 2.2 / 52.6 / 84.8 ms); fp64: 5.8 / 19.1 / 43.7 ms (5.6 / 51.1 / 82.4).
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

// Activation value and derivative from one exponential (softplus: the value is act_f's own expression, bit for bit).
template <typename R, bool DERIV>
__device__ __forceinline__ void act_fd(int a, R z, R& y, R& d) {
    switch (a) {
        case ACT_SOFTPLUS: {
            const R e = r_exp<R>(-(z < R(0) ? -z : z));
            y = r_log1p<R>(e) + (z > R(0) ? z : R(0));
            if (DERIV) d = (z >= R(0) ? R(1) : e) / (R(1) + e);   // sigmoid(z)
            break;
        }
        case ACT_SIGMOID: y = R(1) / (R(1) + r_exp<R>(-z)); if (DERIV) d = y * (R(1) - y); break;
        case ACT_TANH: y = r_tanh<R>(z); if (DERIV) d = R(1) - y * y; break;
        case ACT_RELU: y = z > R(0) ? z : R(0); if (DERIV) d = z > R(0) ? R(1) : R(0); break;
        default: y = z; if (DERIV) d = R(1); break;
    }
}

template <typename R> __device__ __forceinline__ void load_quad(const R* p, R* w);
template <> __device__ __forceinline__ void load_quad<float>(const float* p, float* w) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}
template <> __device__ __forceinline__ void load_quad<double>(const double* p, double* w) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
}

// Register-resident evaluation of the chain on a two-component input: MAXW is a compile-time bound on the layer widths, every
// loop over neurons is fully unrolled (guards on the actual widths), so the activations live in registers -- the run-time indexed
// arrays of mlp2_forward live in local memory.  TANGENT propagates the derivatives with respect to the two inputs alongside
// (forward mode: no tape), which replaces the four extra network evaluations of a central finite difference.  The summation order
// per neuron (bias first, inputs ascending) is that of mlp2_forward.
template <typename R, int MAXW, bool TANGENT>
__device__ __forceinline__ void mlp2_regs(const MlpArch& arch, const R* __restrict__ th, R x0, R x1, R& y, R& dy0, R& dy1) {
    R cur[MAXW], c0[MAXW], c1[MAXW];
#pragma unroll
    for (int i = 0; i < MAXW; ++i) { cur[i] = R(0); c0[i] = R(0); c1[i] = R(0); }
    cur[0] = x0; cur[1] = x1;
    c0[0] = R(1); c1[1] = R(1);
    int k = 0;
    for (int L = 0; L < arch.n_layers; ++L) {
        const int ni = arch.widths[L], no = arch.widths[L + 1];
        const R* W = th + k;
        const R* bv = W + no * ni;
        R s[MAXW], s0[MAXW], s1[MAXW];
#pragma unroll
        for (int o = 0; o < MAXW; ++o) { s[o] = o < no ? bv[o] : R(0); s0[o] = R(0); s1[o] = R(0); }
        const bool vec = ((no | k) & 3) == 0;  // whole quads of outputs, 16-byte aligned columns: one LDS.128 per four weights
#pragma unroll
        for (int i = 0; i < MAXW; ++i) {
            if (i < ni) {
                const R ci = cur[i], t0 = c0[i], t1 = c1[i];
                const R* Wi = W + i * no;   // column i of W (out x in, column-major): contiguous in o
#pragma unroll
                for (int o4 = 0; o4 < MAXW; o4 += 4) {
                    if (o4 < no) {           // (uniform: a 1-wide output layer costs one quad, not MAXW predicated slots)
                        R w4[4];
                        if (vec) load_quad<R>(Wi + o4, w4);
                        else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) w4[q] = (o4 + q < no) ? Wi[o4 + q] : R(0);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            s[o4 + q] += w4[q] * ci;
                            if (TANGENT) { s0[o4 + q] += w4[q] * t0; s1[o4 + q] += w4[q] * t1; }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int o4 = 0; o4 < MAXW; o4 += 4) {
            if (o4 < no) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int o = o4 + q;
                    R yv = R(0), dv = R(0);
                    if (o < no) act_fd<R, TANGENT>(arch.acts[L], s[o], yv, dv);
                    cur[o] = yv;
                    if (TANGENT) { c0[o] = dv * s0[o]; c1[o] = dv * s1[o]; }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) { cur[o4 + q] = R(0); if (TANGENT) { c0[o4 + q] = R(0); c1[o4 + q] = R(0); } }
            }
        }
        k += no * ni + no;
    }
    y = cur[0];
    dy0 = c0[0];
    dy1 = c1[0];
}

// U(H̄, ∇S)  (LawU)   /   Y(T, H̄)  (LawY)
template <typename R, int MAXW = LAW_MAX_WIDTH>
__device__ __forceinline__ R law_eval(const CellLaw& lw, const R* th, R in0, R in1) {
    R y, d0, d1;
    mlp2_regs<R, MAXW, false>(lw.arch, th, law_pre<R>(lw, 0, in0), law_pre<R>(lw, 1, in1), y, d0, d1);
    return law_post<R>(lw, y);
}
// U and its partials with respect to the two (un-normalised) inputs
template <typename R, int MAXW>
__device__ __forceinline__ R law_eval_grad(const CellLaw& lw, const R* th, R in0, R in1, R& dU0, R& dU1) {
    R y, d0, d1;
    mlp2_regs<R, MAXW, true>(lw.arch, th, law_pre<R>(lw, 0, in0), law_pre<R>(lw, 1, in1), y, d0, d1);
    const R U = law_post<R>(lw, y);
    const R dUdy = lw.postscale ? U / (y * y) : R(1);                       // d/dy max_NN exp((y-1)/y)
    const R sc0 = lw.prescale ? R(1) / (R)(lw.hi0 - lw.lo0) : R(1), sc1 = lw.prescale ? R(1) / (R)(lw.hi1 - lw.lo1) : R(1);
    dU0 = dUdy * d0 * sc0;
    dU1 = dUdy * d1 * sc1;
    return U;
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
template <typename T, typename R, bool PARTIALS, int MAXW>
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
            if (!PARTIALS) {
                const R U = law_eval<R, MAXW>(lw, th, (R)Hb, (R)gS);
                Dn[pn] = (T)((R)Hb * U);                                       // target_D_pure.jl:78-96
            } else {
                // α = ∂D/∂H̄, β = ∂D/∂|∇S| of D = H̄·U(H̄, |∇S|).  The reference takes central differences of the network
                // (δH = 1e-4, δ∇S = 1e-6, target_D_pure.jl:105-137), four extra evaluations whose truncation error is ~1e-13
                // and whose rounding noise is ~3e-10; here the exact derivatives ride along the forward evaluation
                // (forward-mode tangents): within that noise of the reference's values (measured on the oracle: 3e-10 / 7e-10
                // relative), one evaluation instead of five, and no fp64 requirement for an fp32 ensemble.
                R dU0, dU1;
                const R U = law_eval_grad<R, MAXW>(lw, th, (R)Hb, (R)gS, dU0, dU1);
                Dn[pn] = (T)((R)Hb * U);
                Al[pn] = (T)(Hb > 0.0 ? U + (R)Hb * dU0 : R(0));
                Be[pn] = (T)((R)Hb * dU1);
            }
        } else {
            const R Tg = (R)d.temp;
            const R Y = law_eval<R, MAXW>(lw, th, Tg, (R)Hb);
            Dn[pn] = (T)hybrid_D<R>(lw, Y, (R)Hb, (R)gS);                      // target_D_hybrid.jl:22-45
            if (PARTIALS) {
                const double dH = 1e-4;                                        // target_D_hybrid.jl:58-73
                const double pq = lw.p - lw.q;
                double noNN = (lw.n_H + 2.0) * (double)Y * lw.Gam * pow(Hb, lw.n_H + 1.0) * pow(gS, lw.n_gS - 1.0);
                if (lw.Sl != 0.0) noNN += (pq + 1.0) * lw.Sl * pow(Hb, pq) * pow(gS, lw.p - 1.0);
                const double Da = (double)hybrid_D<R>(lw, law_eval<R, MAXW>(lw, th, Tg, (R)(Hb + dH)), (R)Hb, (R)gS);
                const double Db = (double)hybrid_D<R>(lw, Y, (R)Hb, (R)gS);
                Al[pn] = (T)(noNN + (Da - Db) / dH);
                double be = lw.Gam * (double)Y * (lw.n_gS - 1.0) * pow(Hb, lw.n_H + 2.0) * pow(gS, lw.n_gS - 3.0);
                if (lw.Sl != 0.0) be += lw.Sl * (lw.p - 1.0) * pow(Hb, pq + 1.0) * pow(gS, lw.p - 3.0);
                Be[pn] = (T)be;                                                // target_D_hybrid.jl:75-96
            }
        }
    }
}

// interpolation = :Linear, step 1: every dual node adds  D†·s  (s as in pass 3) to the knots that bracket its (H̄, ∇S) [LawU] or its
// H̄ [LawY], split by the (bi)linear interpolation weights of Gridded(Linear()); inputs beyond the knot range are clamped to it
// (Interpolations.jl would throw; the reference dilates the range by 1.05).  Wlat: [n0 x max(n1, 1)] doubles of ONE glacier.
template <typename T>
__global__ void __launch_bounds__(LAW_NT)
law_lattice_scatter(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, CellLaw lw, const T* __restrict__ H,
                    const T* __restrict__ B, const T* __restrict__ Dadj, const double* __restrict__ knots, int n0, int n1,
                    double* __restrict__ Wlat) {
    extern __shared__ __align__(16) unsigned char law_smem[];
    double* kn = reinterpret_cast<double*>(law_smem);
    for (int k = threadIdx.x; k < n0 + n1; k += LAW_NT) kn[k] = knots[k];
    __syncthreads();
    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    auto locate = [&](const double* k, int nk, double x, int& a, double& w) {
        x = fmin(fmax(x, k[0]), k[nk - 1]);
        int lo = 0, hi = nk - 1;              // largest a in [0, nk-2] with k[a] <= x  (searchsorted(side = right) - 1, clipped)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (k[mid] <= x) lo = mid; else hi = mid; }
        a = lo;
        w = (x - k[a]) / (k[a + 1] - k[a]);
    };
    for (int c = threadIdx.x; c < TX * TY; c += LAW_NT) {
        const int an = x0 + (c % TX), bn = y0 + (c / TX);
        if (an > d.nx - 2 || bn > d.ny - 2) continue;
        double Hb, gS;
        node_inputs<T>(d, H, B, an, bn, Hb, gS);
        double sc;
        if (lw.kind == LAW_U) sc = (Hb > 0.0) ? Hb : 0.0;
        else sc = lw.Gam * pow(Hb, lw.n_H + 2.0) * pow(gS, lw.n_gS - 1.0);
        double v = (double)__ldg(Dadj + d.off + (long long)bn * d.ld + an) * sc;
        if (!(v == v) || v == 0.0) continue;
        int a, b = 0;
        double wa, wb = 0.0;
        locate(kn, n0, Hb, a, wa);
        if (n1 > 0) locate(kn + n0, n1, gS, b, wb);
        atomicAdd(Wlat + a + (long long)n0 * b, v * (1.0 - wa) * (1.0 - wb));
        atomicAdd(Wlat + a + 1 + (long long)n0 * b, v * wa * (1.0 - wb));
        if (n1 > 0) {
            atomicAdd(Wlat + a + (long long)n0 * (b + 1), v * (1.0 - wa) * wb);
            atomicAdd(Wlat + a + 1 + (long long)n0 * (b + 1), v * wa * wb);
        }
    }
}

// Pass 3.  ∂θ_k = Σ_nodes D†·s·post'(y)·∂NN/∂θ_k ,  s = H̄·[H̄ > 0] (LawU, nodes with H̄ == 0 skipped: target_D_pure.jl:169-171)
// or Γ H̄^{n_H+2} ∇S^{n_∇S-1} (LawY).  One block per tile; block_partial[block][k] written in full (fixed order).
//
// The weight gradient of a Dense layer over a batch of nodes is a small contraction  dW_L[o, i] = Σ_n dz_L[n, o] · a_L[n, i]:
// every thread back-propagates ONE node (register-resident, MAXW-unrolled) and leaves its layer inputs a and output cotangents dz
// in shared memory (node-minor layout, row pitch LAW_NT + 1: conflict-free both ways); then every thread OWNS a few parameters and
// runs the contraction over the LAW_NT nodes of the pass out of shared memory.  (The first version reduced every parameter with
// warp shuffles: 10 SHFL per parameter and warp, 3 210 per 32 nodes for the 2-16-16-1 network.)  Fixed summation order: bit-stable.
constexpr int LAW_PITCH = LAW_NT + 1;

// R: precision of the per-node forward / backward and of the contraction inside one pass (the ensemble's precision; the passes and
// tiles are accumulated in double).
//
// LATTICE (interpolation = :Linear, target_D_pure.jl:180-193 / target_D_hybrid.jl:136-166): the nodes are not the dual nodes of a
// tile but the knots of the interpolation lattice (n0 x n1 for LawU's (H̄, ∇S), n0 x 1 for LawY's H̄ with T fixed), and the weight of
// a knot is what law_lattice_scatter accumulated there -- sum over the cells of D†·s times the cell's (bi)linear interpolation
// weight -- times post'(y) at the knot: the contraction of the interpolated gradient tensor, reordered as a sum over knots.
template <typename T, typename R, int MAXW, bool LATTICE = false>
__global__ void __launch_bounds__(LAW_NT)
law_theta_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, CellLaw lw,
                 const double* __restrict__ theta, const T* __restrict__ H, const T* __restrict__ B,
                 const T* __restrict__ Dadj, double* __restrict__ block_partial, const double* __restrict__ knots = nullptr,
                 int n0 = 0, int n1 = 0, const double* __restrict__ Wlat = nullptr, int glacier = 0) {
    extern __shared__ __align__(16) unsigned char law_smem[];
    const int np = lw.arch.n_params, nl = lw.arch.n_layers;
    int NA = 0, NZ = 0;
    for (int L = 0; L < nl; ++L) { NA += lw.arch.widths[L]; NZ += lw.arch.widths[L + 1]; }
    double* accs = reinterpret_cast<double*>(law_smem);               // n_params accumulators (one owner thread each)
    int2* pmap = reinterpret_cast<int2*>(accs + np);                  // parameter -> (row of Zs, row of As or -1 for a bias)
    R* th = reinterpret_cast<R*>(pmap + np);                          // n_params (padded to a multiple of 4)
    R* As = th + ((np + 3) & ~3);                                     // [NA][LAW_PITCH] layer inputs
    R* Zs = As + (size_t)NA * LAW_PITCH;                              // [NZ][LAW_PITCH] act'(z), then dz
    for (int k = threadIdx.x; k < np; k += LAW_NT) { th[k] = (R)theta[k]; accs[k] = 0.0; }
    {
        int k = 0, aoff = 0, zoff = 0;
        for (int L = 0; L < nl; ++L) {
            const int ni = lw.arch.widths[L], no = lw.arch.widths[L + 1];
            for (int q = threadIdx.x; q < no * ni + no; q += LAW_NT) {
                if (q < no * ni) pmap[k + q] = make_int2(zoff + q % no, aoff + q / no);   // vec(W) column-major: o fastest
                else pmap[k + q] = make_int2(zoff + (q - no * ni), -1);
            }
            k += no * ni + no; aoff += ni; zoff += no;
        }
    }
    __syncthreads();
    const int2 tl = LATTICE ? make_int2(glacier, 0) : tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    const int n = threadIdx.x;
    const int n1e = n1 > 0 ? n1 : 1;
    for (int c0 = 0; c0 < TX * TY; c0 += LAW_NT) {
        const int c = c0 + n;
        int an = x0 + (c % TX), bn = y0 + (c / TX);
        bool in_grid = (an <= d.nx - 2 && bn <= d.ny - 2);
        double Hb = 0.0, gS = 0.0;
        long long lc = 0;
        if (LATTICE) {
            lc = (long long)blockIdx.x * (TX * TY) + c;   // knot index: a + n0 * b
            in_grid = lc < (long long)n0 * n1e;
            if (in_grid) {
                Hb = knots[lc % n0];
                gS = n1 > 0 ? knots[n0 + lc / n0] : 0.0;
            }
        } else if (in_grid) {
            node_inputs<T>(d, H, B, an, bn, Hb, gS);
        }
        const double in0 = lw.kind == LAW_U ? Hb : (double)d.temp, in1 = lw.kind == LAW_U ? gS : Hb;
        // ---- forward: layer inputs -> As, act'(z) -> Zs ----
        R cur[MAXW];
#pragma unroll
        for (int i = 0; i < MAXW; ++i) cur[i] = R(0);
        cur[0] = (R)law_pre<double>(lw, 0, in0);
        cur[1] = (R)law_pre<double>(lw, 1, in1);
        int k = 0, aoff = 0, zoff = 0;
        for (int L = 0; L < nl; ++L) {
            const int ni = lw.arch.widths[L], no = lw.arch.widths[L + 1];
            const R* W = th + k;
            const R* bv = W + no * ni;
            R sacc[MAXW];
#pragma unroll
            for (int o = 0; o < MAXW; ++o) sacc[o] = o < no ? bv[o] : R(0);
#pragma unroll
            for (int i = 0; i < MAXW; ++i) {
                if (i < ni) {
                    const R ci = cur[i];
                    As[(size_t)(aoff + i) * LAW_PITCH + n] = ci;
                    const R* Wi = W + i * no;
#pragma unroll
                    for (int o4 = 0; o4 < MAXW; o4 += 4) {
                        if (o4 < no) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (o4 + q < no) sacc[o4 + q] += Wi[o4 + q] * ci;
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < MAXW; ++o) {
                R yv = R(0), dv = R(0);
                if (o < no) {
                    act_fd<R, true>(lw.arch.acts[L], sacc[o], yv, dv);
                    Zs[(size_t)(zoff + o) * LAW_PITCH + n] = dv;
                }
                cur[o] = yv;
            }
            k += no * ni + no; aoff += ni; zoff += no;
        }
        const double y = (double)cur[0];
        double w = 0.0;
        if (in_grid) {
            const double dpost = lw.postscale ? lw.max_NN * exp((y - 1.0) / y) / (y * y) : 1.0;
            if (LATTICE) {
                w = Wlat[lc] * dpost;
            } else {
                double sc;
                if (lw.kind == LAW_U) sc = (Hb > 0.0) ? Hb : 0.0;
                else sc = lw.Gam * pow(Hb, lw.n_H + 2.0) * pow(gS, lw.n_gS - 1.0);
                w = (double)__ldg(Dadj + d.off + (long long)bn * d.ld + an) * sc * dpost;
            }
            if (!(w == w)) w = 0.0;  // 0·inf from a degenerate node (H̄ = 0 with a negative exponent) contributes nothing
        }
        // ---- backward: dz -> Zs ----
        R g[MAXW];
#pragma unroll
        for (int i = 0; i < MAXW; ++i) g[i] = R(0);
        g[0] = (R)w;
        for (int L = nl - 1; L >= 0; --L) {
            const int ni = lw.arch.widths[L], no = lw.arch.widths[L + 1];
            k -= no * ni + no; aoff -= ni; zoff -= no;
            const R* W = th + k;
            R dz[MAXW], gp[MAXW];
#pragma unroll
            for (int o = 0; o < MAXW; ++o) {
                dz[o] = R(0);
                if (o < no) {
                    R* zp = Zs + (size_t)(zoff + o) * LAW_PITCH + n;
                    dz[o] = (w != 0.0) ? g[o] * (*zp) : R(0);
                    *zp = dz[o];
                }
            }
#pragma unroll
            for (int i = 0; i < MAXW; ++i) {
                gp[i] = R(0);
                if (i < ni) {
                    const R* Wi = W + i * no;
#pragma unroll
                    for (int o4 = 0; o4 < MAXW; o4 += 4) {
                        if (o4 < no) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (o4 + q < no) gp[i] += Wi[o4 + q] * dz[o4 + q];
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < MAXW; ++i) g[i] = gp[i];
        }
        __syncthreads();
        // ---- contraction over the LAW_NT nodes of this pass: every thread owns parameters tid, tid + LAW_NT, ... ----
        for (int q = threadIdx.x; q < np; q += LAW_NT) {
            const int2 pm = pmap[q];
            const R* zr = Zs + (size_t)pm.x * LAW_PITCH;
            R sum = R(0);
            if (pm.y >= 0) {
                const R* ar = As + (size_t)pm.y * LAW_PITCH;
#pragma unroll 8
                for (int m = 0; m < LAW_NT; ++m) sum += zr[m] * ar[m];
            } else {
#pragma unroll 8
                for (int m = 0; m < LAW_NT; ++m) sum += zr[m];
            }
            accs[q] += (double)sum;
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < np; k += LAW_NT) block_partial[(long long)blockIdx.x * np + k] = accs[k];
}

}  // namespace odinn
