// Surface velocity, LossV and their discrete VJPs.
//
// Reference (ODINN.jl v1.1.0):
//   V = Huginn.V_from_H(simulation, H, t, θ)  [NOT IN TREE; shape fixed by adjoint.jl:268-350]: on the dual grid
//       Vx = -Dꜛ ∇Sx, Vy = -Dꜛ ∇Sy, stored in inn1 of an nx x ny matrix;  Dꜛ = Velocityꜛ (src/models/target/target_A.jl:94-108)
//   LossV(L2Sum; component :xy | :abs, scale_loss)   src/losses/Losses.jl:293-390, mask = V_ref > 0
//   VJP_λ_∂surface_V∂H_discrete                      src/inverse/SIA2D/adjoint.jl:268-350  (α = ∂Dꜛ/∂H̄, β = ∂Dꜛ/∂∇H, target_A.jl:110-141)
//   VJP_λ_∂surface_V∂θ_discrete                      adjoint.jl:352-413 with ∂A_spatialꜛ = Γꜛ H̄^{n+1} ∇S^{n-1} (target_A.jl:143-170)
//
// Two passes, no atomics: the node pass evaluates V, the loss, the cotangents and three node planes
//   cA = α sv,  cX = β ∇Sx sv + Dꜛ ∂Vx,  cY = β ∇Sy sv + Dꜛ ∂Vy,   sv = ∇Sx ∂Vx + ∇Sy ∂Vy
// the cell pass gathers  -(avg†(cA) + diff_x†(avg_y†(cX), Δx) + diff_y†(avg_x†(cY), Δy))  from the four nodes of a cell and adds
// it (weighted) to λ.  Evaluated at the tstops that hold velocity data only -- not on the per-step hot path.
#include <algorithm>
#include <utility>
#include <vector>

#include "ensemble.cuh"

namespace odinn {

struct VelConst {
    double GamUp;   // Γꜛ_noA = 2 (ρ g)^n / (n + 1)            (target_utils.jl:21-30)
    double Sl2;     // S (p - q + 2),  S = C (ρ g)^(p - q)       (target_A.jl:100-101)
    double n, p, q;
};

template <typename T>
struct VelNode {
    T gSx, gSy, Dup, alpha, beta, gAup;
};

// Node (i, j) of glacier d from the four cells (i..i+1, j..j+1); arithmetic in double (not a hot path).
template <typename T>
__device__ __forceinline__ void vel_node(const GDesc<T>& d, const T* __restrict__ H, const T* __restrict__ B, const VelConst& c,
                                         double A, int i, int j, double& gSx, double& gSy, double& Dup, double& alpha,
                                         double& beta, double& gAup) {
    const long long p00 = d.off + (long long)j * d.ld + i, p01 = p00 + d.ld;
    const double h00 = fmax((double)H[p00], 0.0), h10 = fmax((double)H[p00 + 1], 0.0);
    const double h01 = fmax((double)H[p01], 0.0), h11 = fmax((double)H[p01 + 1], 0.0);
    const double s00 = (double)B[p00] + h00, s10 = (double)B[p00 + 1] + h10, s01 = (double)B[p01] + h01, s11 = (double)B[p01 + 1] + h11;
    gSx = 0.5 * ((s10 - s00) + (s11 - s01)) * (double)d.inv_dx;   // avg_y(diff_x(S) / Δx)
    gSy = 0.5 * ((s01 - s00) + (s11 - s10)) * (double)d.inv_dy;   // avg_x(diff_y(S) / Δy)
    const double gS = sqrt(gSx * gSx + gSy * gSy);
    const double Hb = 0.25 * (h00 + h10 + h01 + h11);
    const double gn1 = pow(gS, c.n - 1.0), gn3 = pow(gS, c.n - 3.0);
    gAup = c.GamUp * pow(Hb, c.n + 1.0) * gn1;
    Dup = A * gAup;
    alpha = A * c.GamUp * (c.n + 1.0) * pow(Hb, c.n) * gn1;
    beta = A * c.GamUp * (c.n - 1.0) * pow(Hb, c.n + 1.0) * gn3;
    if (c.Sl2 != 0.0) {  // sliding term, exponents as written in target_A.jl:100-101, 118-119, 133-135
        const double pq = c.p - c.q;
        Dup += c.Sl2 * pow(Hb, pq + 1.0) * gn1;
        alpha += c.Sl2 * pow(Hb, pq) * gn1;
        beta += c.Sl2 * (c.p - 1.0) * pow(Hb, pq + 1.0) * gn3;
    }
}

// MODE 0: V only (writes Vx, Vy).  MODE 1: cotangents given in dVx, dVy.  MODE 2 / 3: LossV :xy / :abs against the reference planes.
// Node planes cA, cX, cY are written when cA != nullptr; partial[2 tile] = loss, partial[2 tile + 1] = Σ gAꜛ sv.
template <typename T, int MODE>
__global__ void __launch_bounds__(NT)
vel_node_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const T* __restrict__ H,
                const T* __restrict__ B, VelConst c, const T* __restrict__ inVx, const T* __restrict__ inVy,
                const T* __restrict__ inVabs, const T* __restrict__ Wv, T* __restrict__ outVx, T* __restrict__ outVy,
                T* __restrict__ cA, T* __restrict__ cX, T* __restrict__ cY, double* __restrict__ partial) {
    __shared__ double sRed[NT / 32];
    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    const int i = x0 + (threadIdx.x & 31), tr = threadIdx.x >> 5;
    double accL = 0.0, accS = 0.0;
#pragma unroll
    for (int rr = 0; rr < TY / 8; ++rr) {
        const int j = y0 + tr + rr * 8;
        if (i < d.nx && j < d.ny) {
            const long long p = d.off + (long long)j * d.ld + i;
            const bool node = (i < d.nx - 1 && j < d.ny - 1);
            double vx = 0.0, vy = 0.0, a = 0.0, bx = 0.0, by = 0.0;
            if (node) {
                double gSx, gSy, Dup, alpha, beta, gAup;
                vel_node<T>(d, H, B, c, (double)d.A, i, j, gSx, gSy, Dup, alpha, beta, gAup);
                vx = -Dup * gSx;
                vy = -Dup * gSy;
                if (MODE != 0) {
                    double dVx, dVy;
                    if (MODE == 1) {
                        dVx = (double)inVx[p];
                        dVy = (double)inVy[p];
                    } else {
                        const double w = (double)Wv[p];  // mask(V_ref > 0) / (nx ny [scale])
                        const double ex = vx - (double)inVx[p], ey = vy - (double)inVy[p];
                        if (MODE == 2) accL += w * (ex * ex + ey * ey);
                        else {
                            const double ev = sqrt(vx * vx + vy * vy) - (double)inVabs[p];
                            accL += w * ev * ev;
                        }
                        // :xy  ∂V{x,y} = 2 w (V{x,y} - V{x,y},ref);  :abs (Losses.jl:369-370)  ∂V (V{x,y} - V{x,y},ref) / (V - V_ref)
                        // with ∂V = 2 w (V - V_ref): the same value
                        dVx = 2.0 * w * ex;
                        dVy = 2.0 * w * ey;
                    }
                    const double sv = gSx * dVx + gSy * dVy;
                    accS += gAup * sv;
                    a = alpha * sv;
                    bx = beta * gSx * sv + Dup * dVx;
                    by = beta * gSy * sv + Dup * dVy;
                }
            }
            if (MODE == 0) {
                outVx[p] = (T)vx;
                outVy[p] = (T)vy;
            } else if (cA != nullptr) {
                cA[p] = (T)a;
                cX[p] = (T)bx;
                cY[p] = (T)by;
            }
        }
    }
    if (MODE != 0) {
        double sL = block_sum(accL, sRed);
        __syncthreads();
        double sS = block_sum(accS, sRed);
        if (threadIdx.x == 0) {
            partial[2 * blockIdx.x] = sL;
            partial[2 * blockIdx.x + 1] = sS;
        }
    }
}

// out[p] = (base ? base[p] : 0) + w * -(avg†(cA) + diff_x†(avg_y†(cX), Δx) + diff_y†(avg_x†(cY), Δy))[p]
template <typename T>
__global__ void __launch_bounds__(NT)
vel_cell_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const T* __restrict__ cA,
                const T* __restrict__ cX, const T* __restrict__ cY, const T* base, T* out, double w) {
    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    const int i = x0 + (threadIdx.x & 31), tr = threadIdx.x >> 5;
#pragma unroll
    for (int rr = 0; rr < TY / 8; ++rr) {
        const int j = y0 + tr + rr * 8;
        if (i < d.nx && j < d.ny) {
            const long long p = d.off + (long long)j * d.ld + i;
            // nodes (i-1, j-1), (i, j-1), (i-1, j), (i, j); a node exists for 0 <= ni < nx-1, 0 <= nj < ny-1
            const bool w_ = i >= 1, e_ = i < d.nx - 1, s_ = j >= 1, n_ = j < d.ny - 1;
            auto at = [&](const T* pl, int di, int dj) -> double { return (double)pl[p + di + (long long)dj * d.ld]; };
            double sa = 0.0, sx = 0.0, sy = 0.0;
            if (w_ && s_) { sa += at(cA, -1, -1); sx += at(cX, -1, -1); sy += at(cY, -1, -1); }
            if (e_ && s_) { sa += at(cA, 0, -1); sx -= at(cX, 0, -1); sy += at(cY, 0, -1); }
            if (w_ && n_) { sa += at(cA, -1, 0); sx += at(cX, -1, 0); sy -= at(cY, -1, 0); }
            if (e_ && n_) { sa += at(cA, 0, 0); sx -= at(cX, 0, 0); sy -= at(cY, 0, 0); }
            const double val = -(0.25 * sa + 0.5 * (double)d.inv_dx * sx + 0.5 * (double)d.inv_dy * sy);
            out[p] = (T)((base ? (double)base[p] : 0.0) + w * val);
        }
    }
}

// partial has 2 entries per tile: dst[g] (+)= scale * Σ partial[2 t + which]
__global__ void __launch_bounds__(NT)
vel_reduce_kernel(const int* __restrict__ start, const double* __restrict__ partial, int which, double* __restrict__ dst,
                  double scale, int accumulate) {
    __shared__ double sRed[NT / 32];
    const int g = blockIdx.x;
    double acc = 0.0;
    for (int t = start[g] + threadIdx.x; t < start[g + 1]; t += NT) acc += partial[2 * t + which];
    double s = block_sum(acc, sRed);
    if (threadIdx.x == 0) dst[g] = (accumulate ? dst[g] : 0.0) + scale * s;
}

static VelConst vel_const(const odinn_phys& p) {
    VelConst c;
    c.GamUp = 2.0 * std::pow(p.rho * p.g, p.n) / (p.n + 1.0);
    c.Sl2 = p.C * std::pow(p.rho * p.g, p.p - p.q) * (p.p - p.q + 2.0);
    c.n = p.n; c.p = p.p; c.q = p.q;
    return c;
}

static int vel_prepare(odinn_ensemble* e) {
    int rc;
    if (e->a_gridded || e->law_kind != 0) return fail(e, ODINN_ESTATE, "surface velocity is provided for glacier-wide A laws");
    if ((rc = ensure_plane(e, ODINN_FIELD_B))) return rc;
    for (int k = 0; k < 3; ++k)
        if ((rc = alloc_work_plane(e, &e->ext_dev[EXT_V_WORK0 + k]))) return rc;
    if (!e->ext_dev[EXT_V_PARTIAL]) ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_V_PARTIAL], sizeof(double) * 2 * (size_t)e->n_tiles));
    return sync_descs(e);
}

// Node pass (+ cell pass when lam_out) over tiles [t0, t0 + nt).  mode as in vel_node_kernel.
template <typename T>
static int vel_launch_t(odinn_ensemble* e, int mode, int t0, int nt, const void* H, const void* inVx, const void* inVy,
                        const void* inVabs, const void* Wv, void* outVx, void* outVy, bool planes) {
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const VelConst c = vel_const(e->phys);
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    T* cA = planes ? (T*)e->ext_dev[EXT_V_WORK0] : nullptr;
    T* cX = (T*)e->ext_dev[EXT_V_WORK1];
    T* cY = (T*)e->ext_dev[EXT_V_WORK2];
    double* partial = (double*)e->ext_dev[EXT_V_PARTIAL] + 2 * (size_t)t0;
#define VL(M) vel_node_kernel<T, M><<<nt, NT, 0, e->stream>>>(descs, e->d_tiles + t0, (const T*)H, B, c, (const T*)inVx, (const T*)inVy, \
                                                          (const T*)inVabs, (const T*)Wv, (T*)outVx, (T*)outVy, cA, cX, cY, partial)
    if (mode == 0) VL(0);
    else if (mode == 1) VL(1);
    else if (mode == 2) VL(2);
    else VL(3);
#undef VL
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}
static int vel_launch(odinn_ensemble* e, int mode, int t0, int nt, const void* H, const void* inVx, const void* inVy,
                      const void* inVabs, const void* Wv, void* outVx, void* outVy, bool planes) {
    return e->dtype == ODINN_F32 ? vel_launch_t<float>(e, mode, t0, nt, H, inVx, inVy, inVabs, Wv, outVx, outVy, planes)
                                 : vel_launch_t<double>(e, mode, t0, nt, H, inVx, inVy, inVabs, Wv, outVx, outVy, planes);
}
static int vel_cells(odinn_ensemble* e, int t0, int nt, const void* base, void* out, double w) {
    if (e->dtype == ODINN_F32)
        vel_cell_kernel<float><<<nt, NT, 0, e->stream>>>((const GDesc<float>*)e->d_descs, e->d_tiles + t0, (const float*)e->ext_dev[EXT_V_WORK0],
                                                        (const float*)e->ext_dev[EXT_V_WORK1], (const float*)e->ext_dev[EXT_V_WORK2],
                                                        (const float*)base, (float*)out, w);
    else
        vel_cell_kernel<double><<<nt, NT, 0, e->stream>>>((const GDesc<double>*)e->d_descs, e->d_tiles + t0, (const double*)e->ext_dev[EXT_V_WORK0],
                                                         (const double*)e->ext_dev[EXT_V_WORK1], (const double*)e->ext_dev[EXT_V_WORK2],
                                                         (const double*)base, (double*)out, w);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

static char* vref_plane(odinn_ensemble* e, int slot, int which) {  // which: 0 Vx_ref, 1 Vy_ref, 2 Vabs_ref, 3 Wv
    return (char*)e->ext_dev[EXT_V_REF] + ((size_t)slot * 4 + which) * (size_t)e->total * e->esize;
}

// Hook of the loss / reverse loops: velocity term of snapshot j (no-op without velocity data there or w == 0):
//   loss_dst[g] += w ℓ_V ;  lam += w ∂ℓ_V/∂H (when lam) ;  S_dst[g] += -w Σ gAꜛ sv (when S_dst)
int velocity_loss_term(odinn_ensemble* e, int j, const void* Hj, void* lam, double w, double* loss_dst, double* S_dst) {
    if (w == 0.0 || e->v_snap.empty()) return ODINN_OK;
    int slot = -1;
    for (size_t m = 0; m < e->v_snap.size(); ++m)
        if (e->v_snap[m] == j) slot = (int)m;
    if (slot < 0) return ODINN_OK;
    int rc = vel_prepare(e);
    if (rc) return rc;
    const int mode = e->lossV_component == 1 ? 3 : 2;
    if ((rc = vel_launch(e, mode, 0, e->n_tiles, Hj, vref_plane(e, slot, 0), vref_plane(e, slot, 1), vref_plane(e, slot, 2),
                         vref_plane(e, slot, 3), nullptr, nullptr, lam != nullptr)))
        return rc;
    const double* partial = (const double*)e->ext_dev[EXT_V_PARTIAL];
    if (loss_dst) {
        vel_reduce_kernel<<<e->G, NT, 0, e->stream>>>(e->d_tile_start, partial, 0, loss_dst, w, 1);
        ODINN_CHECK_LAUNCH(e);
    }
    if (S_dst) {
        vel_reduce_kernel<<<e->G, NT, 0, e->stream>>>(e->d_tile_start, partial, 1, S_dst, -w, 1);
        ODINN_CHECK_LAUNCH(e);
    }
    if (lam) return vel_cells(e, 0, e->n_tiles, lam, lam, w);
    return ODINN_OK;
}


// ---- velocity references interpolated at a quadrature node of the continuous adjoint (gradient.jl:289-301, 474-507) ----------

// out{0,1,2} = (1 - a) ref_a{Vx, Vy, Vabs} + a ref_b{...};  partial[2 tile] = #(Vabs > 0), partial[2 tile + 1] = sum_mask (Vx^2 + Vy^2)
template <typename T>
__global__ void __launch_bounds__(NT)
vq_lerp_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const T* __restrict__ ra, const T* __restrict__ rb,
               T* __restrict__ out, long long plane, double a, double* __restrict__ partial) {
    __shared__ double sRed[NT / 32];
    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    const int i = x0 + (threadIdx.x & 31), tr = threadIdx.x >> 5;
    double cnt = 0.0, ss = 0.0;
#pragma unroll
    for (int rr = 0; rr < TY / 8; ++rr) {
        const int j = y0 + tr + rr * 8;
        if (i < d.nx && j < d.ny) {
            const long long p = d.off + (long long)j * d.ld + i;
            double v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[c] = (1.0 - a) * (double)ra[c * plane + p] + a * (double)rb[c * plane + p];
                out[c * plane + p] = (T)v[c];
            }
            if (v[2] > 0.0) { cnt += 1.0; ss += v[0] * v[0] + v[1] * v[1]; }
        }
    }
    double s0 = block_sum(cnt, sRed);
    __syncthreads();
    double s1 = block_sum(ss, sRed);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s0; partial[2 * blockIdx.x + 1] = s1; }
}

// W = (Vabs > 0) / (nx ny [sqrt(mean_mask(Vx^2 + Vy^2))])      (Losses.jl:316, 327-331)
template <typename T>
__global__ void __launch_bounds__(NT)
vq_weight_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const T* __restrict__ Vabs, T* __restrict__ W,
                 const double* __restrict__ red, int G, int scale_loss) {
    const int2 tl = tiles[blockIdx.x];
    const GDesc<T> d = descs[tl.x];
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;
    const int i = x0 + (threadIdx.x & 31), tr = threadIdx.x >> 5;
    const double cnt = red[tl.x], ss = red[G + tl.x];
    const double sc = (scale_loss && cnt > 0.0) ? sqrt(ss / cnt) : 1.0;
    const double w = 1.0 / ((double)d.nx * (double)d.ny * sc);
#pragma unroll
    for (int rr = 0; rr < TY / 8; ++rr) {
        const int j = y0 + tr + rr * 8;
        if (i < d.nx && j < d.ny) {
            const long long p = d.off + (long long)j * d.ld + i;
            W[p] = ((double)Vabs[p] > 0.0) ? (T)w : T(0);
        }
    }
}

int velocity_theta_term_interp(odinn_ensemble* e, double tq, const double* t, int n_t, const void* H, double scale, double* S_dst) {
    if (scale == 0.0 || e->v_snap.empty()) return ODINN_OK;
    int rc = vel_prepare(e);
    if (rc) return rc;
    // data times in ascending order
    std::vector<std::pair<double, int>> data;
    for (size_t m = 0; m < e->v_snap.size(); ++m)
        if (e->v_snap[m] >= 0 && e->v_snap[m] < n_t) data.push_back({t[e->v_snap[m]], (int)m});
    if (data.empty()) return ODINN_OK;
    std::sort(data.begin(), data.end());
    const void *Vx, *Vy, *Vabs, *W;
    if (data.size() == 1) {  // "when there is only one reference velocity data we use a constant interpolator"
        const int m = data[0].second;
        Vx = vref_plane(e, m, 0); Vy = vref_plane(e, m, 1); Vabs = vref_plane(e, m, 2); W = vref_plane(e, m, 3);
    } else {
        size_t k = 0;
        while (k + 2 < data.size() && tq >= data[k + 1].first) ++k;
        double a = (tq - data[k].first) / (data[k + 1].first - data[k].first);
        a = std::min(1.0, std::max(0.0, a));
        if ((rc = alloc_work_plane(e, &e->ext_dev[EXT_VQ_WORK], 4))) return rc;
        if (!e->ext_dev[EXT_VQ_RED]) ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_VQ_RED], sizeof(double) * 2 * e->G));
        char* wk = (char*)e->ext_dev[EXT_VQ_WORK];
        const size_t pb = (size_t)e->total * e->esize;
        double* red = (double*)e->ext_dev[EXT_VQ_RED];
        double* partial = (double*)e->ext_dev[EXT_V_PARTIAL];
        if (e->dtype == ODINN_F32) {
            vq_lerp_kernel<float><<<e->n_tiles, NT, 0, e->stream>>>((const GDesc<float>*)e->d_descs, e->d_tiles, (const float*)vref_plane(e, data[k].second, 0),
                                                                   (const float*)vref_plane(e, data[k + 1].second, 0), (float*)wk, e->total, a, partial);
        } else {
            vq_lerp_kernel<double><<<e->n_tiles, NT, 0, e->stream>>>((const GDesc<double>*)e->d_descs, e->d_tiles, (const double*)vref_plane(e, data[k].second, 0),
                                                                    (const double*)vref_plane(e, data[k + 1].second, 0), (double*)wk, e->total, a, partial);
        }
        ODINN_CHECK_LAUNCH(e);
        vel_reduce_kernel<<<e->G, NT, 0, e->stream>>>(e->d_tile_start, partial, 0, red, 1.0, 0);
        ODINN_CHECK_LAUNCH(e);
        vel_reduce_kernel<<<e->G, NT, 0, e->stream>>>(e->d_tile_start, partial, 1, red + e->G, 1.0, 0);
        ODINN_CHECK_LAUNCH(e);
        if (e->dtype == ODINN_F32)
            vq_weight_kernel<float><<<e->n_tiles, NT, 0, e->stream>>>((const GDesc<float>*)e->d_descs, e->d_tiles, (const float*)(wk + 2 * pb), (float*)(wk + 3 * pb),
                                                                     red, e->G, e->lossV_scale_loss);
        else
            vq_weight_kernel<double><<<e->n_tiles, NT, 0, e->stream>>>((const GDesc<double>*)e->d_descs, e->d_tiles, (const double*)(wk + 2 * pb),
                                                                      (double*)(wk + 3 * pb), red, e->G, e->lossV_scale_loss);
        ODINN_CHECK_LAUNCH(e);
        Vx = wk; Vy = wk + pb; Vabs = wk + 2 * pb; W = wk + 3 * pb;
    }
    const int mode = e->lossV_component == 1 ? 3 : 2;
    if ((rc = vel_launch(e, mode, 0, e->n_tiles, H, Vx, Vy, Vabs, W, nullptr, nullptr, false))) return rc;
    vel_reduce_kernel<<<e->G, NT, 0, e->stream>>>(e->d_tile_start, (const double*)e->ext_dev[EXT_V_PARTIAL], 1, S_dst, -scale, 1);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

}  // namespace odinn

using namespace odinn;

#define VGUARD(e)                                                                                                        \
    if (!(e)) return fail(nullptr, ODINN_EARG, "null ensemble");                                                         \
    {                                                                                                                    \
        cudaError_t s_ = cudaSetDevice((e)->device);                                                                     \
        if (s_ != cudaSuccess) return fail((e), ODINN_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(s_));   \
    }

static int tiles_of(odinn_ensemble* e, int g, int& t0, int& nt) {
    if (g < 0 || g >= e->G) return fail(e, ODINN_EARG, "glacier index out of range");
    t0 = e->gl[g].tile0;
    nt = e->gl[g].ntx * e->gl[g].nty;
    return ODINN_OK;
}

extern "C" {

int odinn_surface_velocity(odinn_ensemble* e, int glacier, const void* H, int ldH, void* Vx, void* Vy, int ldV, double t) {
    (void)t;
    VGUARD(e);
    if (!H || !Vx || !Vy) return fail(e, ODINN_EARG, "null pointer");
    int rc, t0, nt;
    if ((rc = tiles_of(e, glacier, t0, nt)) || (rc = vel_prepare(e)) || (rc = ensure_plane(e, ODINN_FIELD_H))) return rc;
    if ((rc = odinn_upload(e, glacier, ODINN_FIELD_H, H, ldH))) return rc;
    void* oVx = e->ext_dev[EXT_V_WORK1];
    void* oVy = e->ext_dev[EXT_V_WORK2];
    if ((rc = vel_launch(e, 0, t0, nt, e->plane[ODINN_FIELD_H], nullptr, nullptr, nullptr, nullptr, oVx, oVy, false))) return rc;
    if ((rc = copy_plane_2d(e, glacier, oVx, Vx, ldV, false)) || (rc = copy_plane_2d(e, glacier, oVy, Vy, ldV, false))) return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_sia2d_vjp_surface_V(odinn_ensemble* e, int glacier, const void* dVx, const void* dVy, int ldV, const void* H, int ldH,
                              void* out_dH, int ldo, double* out_S, double t) {
    (void)t;
    VGUARD(e);
    if (!H || !dVx || !dVy) return fail(e, ODINN_EARG, "null pointer");
    int rc, t0, nt;
    if ((rc = tiles_of(e, glacier, t0, nt)) || (rc = vel_prepare(e)) || (rc = ensure_plane(e, ODINN_FIELD_H)) ||
        (rc = ensure_plane(e, ODINN_FIELD_LAMBDA)) || (rc = ensure_plane(e, ODINN_FIELD_VJP_H)) || (rc = ensure_plane(e, ODINN_FIELD_DH)))
        return rc;
    if ((rc = odinn_upload(e, glacier, ODINN_FIELD_H, H, ldH)) || (rc = odinn_upload(e, glacier, ODINN_FIELD_LAMBDA, dVx, ldV)) ||
        (rc = odinn_upload(e, glacier, ODINN_FIELD_DH, dVy, ldV)))
        return rc;
    if ((rc = vel_launch(e, 1, t0, nt, e->plane[ODINN_FIELD_H], e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_DH], nullptr, nullptr,
                         nullptr, nullptr, true)))
        return rc;
    if (out_dH) {
        if ((rc = vel_cells(e, t0, nt, nullptr, e->plane[ODINN_FIELD_VJP_H], 1.0))) return rc;
        if ((rc = copy_plane_2d(e, glacier, e->plane[ODINN_FIELD_VJP_H], out_dH, ldo, false))) return rc;
    }
    if (out_S) {
        vel_reduce_kernel<<<1, NT, 0, e->stream>>>(e->d_tile_start + glacier, (const double*)e->ext_dev[EXT_V_PARTIAL], 1, e->d_S + glacier, -1.0, 0);
        ODINN_CHECK_LAUNCH(e);
        ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_S + glacier, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    }
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    if (out_S) *out_S = e->h_S[0];
    return ODINN_OK;
}

int odinn_set_velocity_reference(odinn_ensemble* e, int glacier, int slot, int n_slots, int snapshot_index, const void* Vx_ref,
                                 const void* Vy_ref, const void* Vabs_ref, const void* Wv, int ld) {
    VGUARD(e);
    if (n_slots < 1 || slot < 0 || slot >= n_slots || snapshot_index < 0 || !Vx_ref || !Vy_ref || !Vabs_ref || !Wv)
        return fail(e, ODINN_EARG, "bad velocity reference arguments");
    if (glacier < 0 || glacier >= e->G) return fail(e, ODINN_EARG, "glacier index out of range");
    if ((int)e->v_snap.size() != n_slots) {
        if (e->ext_dev[EXT_V_REF]) cudaFree(e->ext_dev[EXT_V_REF]);
        e->ext_dev[EXT_V_REF] = nullptr;
        e->v_snap.assign(n_slots, -1);
    }
    int rc;
    if ((rc = alloc_work_plane(e, &e->ext_dev[EXT_V_REF], (size_t)4 * n_slots))) return rc;
    e->v_snap[slot] = snapshot_index;
    const void* src[4] = {Vx_ref, Vy_ref, Vabs_ref, Wv};
    for (int k = 0; k < 4; ++k)
        if ((rc = copy_plane_2d(e, glacier, vref_plane(e, slot, k), const_cast<void*>(src[k]), ld, true))) return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_set_loss_weights(odinn_ensemble* e, int n_t, const double* wH, const double* wV, int v_component) {
    VGUARD(e);
    if (n_t < 0 || (n_t > 0 && (!wH || !wV)) || (v_component != 0 && v_component != 1)) return fail(e, ODINN_EARG, "bad loss weights");
    e->loss_wH.assign(wH, wH + n_t);
    e->loss_wV.assign(wV, wV + n_t);
    e->lossV_component = v_component;
    return ODINN_OK;
}

}  // extern "C"

extern "C" int odinn_set_velocity_quadrature(odinn_ensemble* e, double theta_scale, int scale_loss) {
    VGUARD(e);
    e->lossV_theta_scale = theta_scale;
    e->lossV_scale_loss = scale_loss ? 1 : 0;
    return ODINN_OK;
}
