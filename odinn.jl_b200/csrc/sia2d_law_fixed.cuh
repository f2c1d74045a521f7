// Per-cell MLP laws, compile-time architecture: 2 -> H1 -> H2 -> 1 with softplus / softplus / sigmoid -- BASELINE config 4's network
// (two hidden layers of 16; `build_default_NN`-style chains of other shapes keep the generic kernels of sia2d_law.cuh).
//
// ncu of the generic kernels (profiles/r02_law_generic_ncu_summary.txt, 500 x 500 x 64 nodes, fp32): law_nodes_kernel executes ~3 200
// instructions per node for a network of 304 multiply-adds, and its first stall reason is `no_instruction` (5.1 stalled warps per
// issue): the run-time layer loop around 32-wide unrolled bodies with per-neuron activation switches neither fits the instruction
// cache nor the issue budget.  The theta pullback is worse: its weight-gradient contraction  dW[o, i] = sum_n dz[n, o] a[n, i]  reads
// two shared-memory words per multiply-add.  Here:
//   * widths and activations are template constants: straight-line code, weights as LDS.128 quads, no predicated slots;
//   * fp32 softplus / sigmoid through ex2.approx / lg2.approx / rcp.approx (absolute error < 1e-6 on the activations: inside the 5e-5
//     parity bound of the fp32 law path; fp64 keeps exp / log1p);
//   * the H2 x H1 weight-gradient contraction is REGISTER-TILED: a thread owns a 4 x 4 tile of dW2 and walks the nodes of its group
//     with two LDS.128 per 16 multiply-adds (the 128 threads are 16 tiles x 8 node groups); the 81 small parameters have one
//     owner thread each; accumulation over all passes of a block stays in registers; fixed summation order: bit-stable;
//   * one launch covers every glacier (block partials indexed by tile), one reduction launch sums the tiles of each glacier.
#pragma once
#include "sia2d_law.cuh"

namespace odinn {

constexpr int LF_PITCH = 20;   // floats per staged node row of 16: 80 bytes -> 16-byte aligned quads, conflict-free STS.128 / LDS.128

// softplus(z) and its derivative sigmoid(z); fp32: ex2 / lg2 / rcp approximations
__device__ __forceinline__ void lf_softplus(float z, float& y, float& d) {
    const float e = exp2f(-fabsf(z) * 1.4426950408889634f);          // exp(-|z|) in (0, 1]
    y = fmaxf(z, 0.0f) + __log2f(1.0f + e) * 0.6931471805599453f;
    d = __fdividef(z >= 0.0f ? 1.0f : e, 1.0f + e);
}
__device__ __forceinline__ void lf_softplus(double z, double& y, double& d) {
    const double e = exp(-fabs(z));
    y = log1p(e) + fmax(z, 0.0);
    d = (z >= 0.0 ? 1.0 : e) / (1.0 + e);
}
__device__ __forceinline__ void lf_sigmoid(float z, float& y, float& d) {
    y = __fdividef(1.0f, 1.0f + exp2f(-z * 1.4426950408889634f));
    d = y * (1.0f - y);
}
__device__ __forceinline__ void lf_sigmoid(double z, double& y, double& d) {
    y = 1.0 / (1.0 + exp(-z));
    d = y * (1.0 - y);
}

// theta layout (Lux): [vec(W1) (H1 x 2, col-major); b1; vec(W2) (H2 x H1); b2; vec(W3) (1 x H2); b3]
template <int H1, int H2>
struct LfLayout {
    static constexpr int W1 = 0, B1 = 2 * H1, W2 = B1 + H1, B2 = W2 + H1 * H2, W3 = B2 + H2, B3 = W3 + H2, NP = B3 + 1;
};

// Forward pass keeping what the backward pass needs: a1, d1 = act'(z1), a2, d2, y, d3.  TANGENT: derivatives of y with respect to the
// two inputs ride along (t0, t1).
template <typename R, int H1, int H2, bool TANGENT>
__device__ __forceinline__ void lf_forward(const R* __restrict__ th, R x0, R x1, R* a1, R* d1, R* a2, R* d2, R& y, R& d3, R& ty0, R& ty1) {
    typedef LfLayout<H1, H2> L;
    R u0[H1], u1[H1];   // tangents of layer 1 (TANGENT)
#pragma unroll
    for (int o4 = 0; o4 < H1; o4 += 4) {
        R wa[4], wb[4], bb[4];
        load_quad<R>(th + L::W1 + o4, wa);        // column 0 of W1
        load_quad<R>(th + L::W1 + H1 + o4, wb);   // column 1
        load_quad<R>(th + L::B1 + o4, bb);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const R z = bb[q] + wa[q] * x0 + wb[q] * x1;
            lf_softplus(z, a1[o4 + q], d1[o4 + q]);
            if (TANGENT) { u0[o4 + q] = d1[o4 + q] * wa[q]; u1[o4 + q] = d1[o4 + q] * wb[q]; }
        }
    }
    R z2[H2], v0[H2], v1[H2];
#pragma unroll
    for (int o4 = 0; o4 < H2; o4 += 4) load_quad<R>(th + L::B2 + o4, z2 + o4);
    if (TANGENT) {
#pragma unroll
        for (int o = 0; o < H2; ++o) { v0[o] = R(0); v1[o] = R(0); }
    }
#pragma unroll
    for (int i = 0; i < H1; ++i) {
#pragma unroll
        for (int o4 = 0; o4 < H2; o4 += 4) {
            R w[4];
            load_quad<R>(th + L::W2 + i * H2 + o4, w);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                z2[o4 + q] += w[q] * a1[i];
                if (TANGENT) { v0[o4 + q] += w[q] * u0[i]; v1[o4 + q] += w[q] * u1[i]; }
            }
        }
    }
    R z3 = th[L::B3], s0 = R(0), s1 = R(0);
#pragma unroll
    for (int o4 = 0; o4 < H2; o4 += 4) {
        R w[4];
        load_quad<R>(th + L::W3 + o4, w);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            lf_softplus(z2[o4 + q], a2[o4 + q], d2[o4 + q]);
            z3 += w[q] * a2[o4 + q];
            if (TANGENT) { s0 += w[q] * d2[o4 + q] * v0[o4 + q]; s1 += w[q] * d2[o4 + q] * v1[o4 + q]; }
        }
    }
    lf_sigmoid(z3, y, d3);
    if (TANGENT) { ty0 = d3 * s0; ty1 = d3 * s1; }
}

// ---- pass 1: node planes D (alpha, beta) -------------------------------------------------------------------------------------
template <typename T, int H1, int H2, bool PARTIALS>
__global__ void __launch_bounds__(LAW_NT)
law_nodes_fixed_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, CellLaw lw, const double* __restrict__ theta,
                       const T* __restrict__ H, const T* __restrict__ B, T* __restrict__ Dn, T* __restrict__ Al, T* __restrict__ Be) {
    typedef LfLayout<H1, H2> L;
    __shared__ __align__(16) T th[(L::NP + 3) & ~3];
    for (int k = threadIdx.x; k < L::NP; k += LAW_NT) th[k] = (T)theta[k];
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
        const bool isU = lw.kind == LAW_U;
        const T in0 = (T)(isU ? Hb : (double)d.temp), in1 = (T)(isU ? gS : Hb);
        T a1[H1], d1[H1], a2[H2], d2[H2], y, d3, t0, t1;
        lf_forward<T, H1, H2, PARTIALS>(th, law_pre<T>(lw, 0, in0), law_pre<T>(lw, 1, in1), a1, d1, a2, d2, y, d3, t0, t1);
        const T out = law_post<T>(lw, y);
        if (isU) {
            Dn[pn] = (T)((T)Hb * out);                                       // target_D_pure.jl:78-96
            if (PARTIALS) {   // exact derivatives (see law_nodes_kernel): alpha = dD/dHbar, beta = dD/d|gradS|
                const T dUdy = lw.postscale ? out / (y * y) : T(1);
                const T sc0 = lw.prescale ? T(1) / (T)(lw.hi0 - lw.lo0) : T(1), sc1 = lw.prescale ? T(1) / (T)(lw.hi1 - lw.lo1) : T(1);
                Al[pn] = (T)(Hb > 0.0 ? out + (T)Hb * dUdy * t0 * sc0 : T(0));
                Be[pn] = (T)((T)Hb * dUdy * t1 * sc1);
            }
        } else {
            Dn[pn] = (T)hybrid_D<T>(lw, out, (T)Hb, (T)gS);                   // target_D_hybrid.jl:22-45 (forward only: the partials
        }                                                                     //  of LawY are finite differences in fp64 -> generic kernel)
    }
}

// ---- pass 3: theta pullback ------------------------------------------------------------------------------------------------------
// block_partial[tile][k], k in Lux order.  LATTICE as in law_theta_kernel.
template <typename T, int H1, int H2, bool LATTICE>
__global__ void __launch_bounds__(LAW_NT)
law_theta_fixed_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, CellLaw lw, const double* __restrict__ theta,
                       const T* __restrict__ H, const T* __restrict__ B, const T* __restrict__ Dadj, double* __restrict__ block_partial,
                       const double* __restrict__ knots, int n0, int n1, const double* __restrict__ Wlat, int glacier) {
    static_assert(H1 == 16 && H2 == 16 && LAW_NT == 128, "tiling below: 16 tiles of 4 x 4, 8 node groups of 16 nodes");
    typedef LfLayout<H1, H2> L;
    extern __shared__ __align__(16) unsigned char lf_smem[];
    T* th = reinterpret_cast<T*>(lf_smem);                    // NP (padded to 4)
    T* sA1 = th + ((L::NP + 3) & ~3);                         // [128][LF_PITCH] a1
    T* sZ2 = sA1 + LAW_NT * LF_PITCH;                         // dz2
    T* sA2 = sZ2 + LAW_NT * LF_PITCH;                         // a2
    T* sZ1 = sA2 + LAW_NT * LF_PITCH;                         // dz1
    T* sX = sZ1 + LAW_NT * LF_PITCH;                          // [128][4]: x0, x1, dz3, -
    for (int k = threadIdx.x; k < L::NP; k += LAW_NT) th[k] = (T)theta[k];
    __syncthreads();
    const int2 tl = LATTICE ? make_int2(glacier, 0) : tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0t = (tl.y & 0xffff) * TX, y0t = (tl.y >> 16) * TY;
    const int n = threadIdx.x;
    const int n1e = n1 > 0 ? n1 : 1;
    // contraction ownership: tile (oi, ii) of dW2 and node group gq
    const int tq = n & 15, oi = tq >> 2, ii = tq & 3, gq = n >> 4;
    T acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = T(0);
    T small = T(0);   // one of the 81 small parameters (threads 0..80): dW1 (32), db1 (16), db2 (16), dW3 (16), db3 (1)
    for (int c0 = 0; c0 < TX * TY; c0 += LAW_NT) {
        const int c = c0 + n;
        const int an = x0t + (c % TX), bn = y0t + (c / TX);
        bool in_grid = (an <= d.nx - 2 && bn <= d.ny - 2);
        double Hb = 0.0, gS = 0.0;
        long long lc = 0;
        if (LATTICE) {
            lc = (long long)blockIdx.x * (TX * TY) + c;
            in_grid = lc < (long long)n0 * n1e;
            if (in_grid) { Hb = knots[lc % n0]; gS = n1 > 0 ? knots[n0 + lc / n0] : 0.0; }
        } else if (in_grid) {
            node_inputs<T>(d, H, B, an, bn, Hb, gS);
        }
        const bool isU = lw.kind == LAW_U;
        const T xi0 = law_pre<T>(lw, 0, (T)(isU ? Hb : (double)d.temp)), xi1 = law_pre<T>(lw, 1, (T)(isU ? gS : Hb));
        T a1[H1], d1[H1], a2[H2], d2[H2], y, d3, t0, t1;
        lf_forward<T, H1, H2, false>(th, xi0, xi1, a1, d1, a2, d2, y, d3, t0, t1);
        double w = 0.0;
        if (in_grid) {
            const double yd = (double)y;
            const double dpost = lw.postscale ? lw.max_NN * exp((yd - 1.0) / yd) / (yd * yd) : 1.0;
            if (LATTICE) {
                w = Wlat[lc] * dpost;
            } else {
                double sc;
                if (isU) sc = (Hb > 0.0) ? Hb : 0.0;
                else sc = lw.Gam * pow(Hb, lw.n_H + 2.0) * pow(gS, lw.n_gS - 1.0);
                w = (double)__ldg(Dadj + d.off + (long long)bn * d.ld + an) * sc * dpost;
            }
            if (!(w == w)) w = 0.0;
        }
        // ---- backward ----
        const T dz3 = (T)w * d3;
        T dz2[H2], g1[H1];
#pragma unroll
        for (int o4 = 0; o4 < H2; o4 += 4) {
            T w3[4];
            load_quad<T>(th + L::W3 + o4, w3);
#pragma unroll
            for (int q = 0; q < 4; ++q) dz2[o4 + q] = w3[q] * dz3 * d2[o4 + q];
        }
#pragma unroll
        for (int i = 0; i < H1; ++i) {
            T s = T(0);
#pragma unroll
            for (int o4 = 0; o4 < H2; o4 += 4) {
                T wq[4];
                load_quad<T>(th + L::W2 + i * H2 + o4, wq);
#pragma unroll
                for (int q = 0; q < 4; ++q) s += wq[q] * dz2[o4 + q];
            }
            g1[i] = s * d1[i];   // dz1
        }
        // ---- stage this node's rows (node-major, 16-byte quads) ----
        __syncthreads();   // (the previous pass's contraction is done with the buffers)
        auto stq = [&](T* base, const T* v) {
#pragma unroll
            for (int o4 = 0; o4 < 16; o4 += 4) {
                if (sizeof(T) == 4) *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + n * LF_PITCH + o4) =
                        make_float4((float)v[o4], (float)v[o4 + 1], (float)v[o4 + 2], (float)v[o4 + 3]);
                else {
                    double* p = reinterpret_cast<double*>(base) + n * LF_PITCH + o4;
                    *reinterpret_cast<double2*>(p) = make_double2((double)v[o4], (double)v[o4 + 1]);
                    *reinterpret_cast<double2*>(p + 2) = make_double2((double)v[o4 + 2], (double)v[o4 + 3]);
                }
            }
        };
        stq(sA1, a1); stq(sZ2, dz2); stq(sA2, a2); stq(sZ1, g1);
        sX[n * 4 + 0] = xi0; sX[n * 4 + 1] = xi1; sX[n * 4 + 2] = dz3;
        __syncthreads();
        // ---- contraction: dW2 tile over this thread's node group (16 nodes), then the small parameter over all 128 ----
#pragma unroll 4
        for (int m = 0; m < 16; ++m) {
            const int nd = gq + 8 * m;
            T zq[4], aq[4];
            load_quad<T>(sZ2 + nd * LF_PITCH + 4 * oi, zq);
            load_quad<T>(sA1 + nd * LF_PITCH + 4 * ii, aq);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int s_ = 0; s_ < 4; ++s_) acc[4 * r + s_] += zq[r] * aq[s_];
        }
        if (n < 81) {
            T s = T(0);
            if (n < 32) {            // dW1[o, j] = sum dz1[o] x_j       (k = o + H1 * j)
                const int o = n & 15, j = n >> 4;
                for (int m = 0; m < LAW_NT; ++m) s += sZ1[m * LF_PITCH + o] * sX[m * 4 + j];
            } else if (n < 48) {     // db1
                const int o = n - 32;
                for (int m = 0; m < LAW_NT; ++m) s += sZ1[m * LF_PITCH + o];
            } else if (n < 64) {     // db2
                const int o = n - 48;
                for (int m = 0; m < LAW_NT; ++m) s += sZ2[m * LF_PITCH + o];
            } else if (n < 80) {     // dW3[o] = sum dz3 a2[o]
                const int o = n - 64;
                for (int m = 0; m < LAW_NT; ++m) s += sX[m * 4 + 2] * sA2[m * LF_PITCH + o];
            } else {                 // db3
                for (int m = 0; m < LAW_NT; ++m) s += sX[m * 4 + 2];
            }
            small += s;
        }
    }
    // ---- block result: dW2 tiles summed over the 8 node groups (fixed order), small parameters straight ----
    __syncthreads();
    double* red = reinterpret_cast<double*>(sA1);   // reuse: [8][256] doubles = 16 KB <= the fp32 staging area (4 x 128 x 20 x 4 B = 40 KB)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s_ = 0; s_ < 4; ++s_) red[gq * 256 + (4 * oi + r) + H2 * (4 * ii + s_)] = (double)acc[4 * r + s_];   // W2[o, i] at o + H2 * i
    __syncthreads();
    double* bp = block_partial + (long long)blockIdx.x * L::NP;
    for (int k = n; k < 256; k += LAW_NT) {
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += red[g * 256 + k];
        bp[L::W2 + k] = s;
    }
    if (n < 32) bp[L::W1 + n] = (double)small;
    else if (n < 48) bp[L::B1 + (n - 32)] = (double)small;
    else if (n < 64) bp[L::B2 + (n - 48)] = (double)small;
    else if (n < 80) bp[L::W3 + (n - 64)] = (double)small;
    else if (n == 80) bp[L::B3] = (double)small;
}

template <typename T, int H1, int H2>
constexpr size_t lf_theta_smem() {
    return sizeof(T) * (((LfLayout<H1, H2>::NP + 3) & ~3) + 4 * LAW_NT * LF_PITCH + LAW_NT * 4);
}

// sums the block partials of every glacier's tiles: out[g][k] = (accumulate ? out : 0) + scale * sum_{tiles of g} partial[tile][k]
static __global__ void law_theta_reduce_all(const int* __restrict__ tile_start, const double* __restrict__ block_partial, int n_params,
                                            double* __restrict__ out, double scale, int accumulate) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (k >= n_params) return;
    double s = 0.0;
    for (int t = tile_start[g]; t < tile_start[g + 1]; ++t) s += block_partial[(long long)t * n_params + k];
    out[(long long)g * n_params + k] = (accumulate ? out[(long long)g * n_params + k] : 0.0) + scale * s;
}

}  // namespace odinn
