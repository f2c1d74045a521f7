// Shared-memory-resident forward solves of SMALL glaciers: one thread-block cluster per glacier, a whole range of
// tstop intervals (every sub-step and Runge-Kutta stage of it) in ONE launch.
//
// Why: the reference's real workloads are small grids (BASELINE config 1: one 128 x 128 glacier, config 3 / 4: 32 - 64
// glaciers of 100 - 400 px).  One F1 launch over such a grid lasts 2 - 4 us even when replayed from a CUDA graph, and a
// 5-year SSPRK3 run is 1440 of them: launch-latency bound at 1 - 2 % of what the marching kernels reach on a large ensemble
// (profiles/r01_v6_configs_f32.jsonl); the adaptive solvers add a host round trip per trial step.  Here the state never
// leaves the SMs between stages:
//   * the glacier is cut into CS row bands, one per CTA of the cluster; each CTA keeps B, the stage planes and the dual-node
//     diffusivity plane of its band (+ one halo row on each side) in shared memory;
//   * an RHS evaluation is two sweeps over the band -- nodes (D, once per node), then cells (edge fluxes, divergence, and the
//     Runge-Kutta update of the scheme as the epilogue) -- with 16-byte shared-memory accesses, four cells per thread;
//   * the first / last row of a new stage value is also stored into the neighbouring CTA's halo row through distributed shared
//     memory, and ONE cluster barrier (arrive.release / wait.acquire) per stage orders those stores before the next sweep;
//   * at every tstop the band is written to the snapshot plane (gradient.jl:73: the forward states the adjoint re-reads);
//   * the adaptive scheme reduces its error norm over the cluster in a fixed order (every CTA receives every partial through
//     DSMEM and forms the same sum), so every CTA takes the same accept / reject decision without a host round trip.
// Two kernels: sia2d_interval_cluster (Euler / SSPRK(3,3) with fixed sub-steps, odinn_solve_forward) and
// sia2d_rdpk_cluster (RDPK3Sp35 + PID controller, the reference's default integrator, odinn_solve_forward_adaptive).
// The arithmetic is that of sia2d_rhs_march (same raw sums, same clamp, same RK forms): reference semantics and citations
// in sia2d_kernels.cuh -- SIA2D! as restated at src/inverse/SIA2D/adjoint.jl:47-104, solve with tstops
// src/simulations/inversions/inversion_utils.jl:551-610, solver default src/inverse/AdjointTypes.jl:60.
#pragma once
#include <cooperative_groups.h>

#include "sia2d_march.cuh"

namespace odinn {

namespace cg = cooperative_groups;

#ifndef ODINN_CL_NT
#define ODINN_CL_NT 512
#endif
constexpr int CL_NT = ODINN_CL_NT;         // threads per CTA
constexpr int CL_PAD = 4;          // zero columns left of column 0 (the row pitch also leaves >= 4 right of column nx-1)
constexpr int CL_PLANES_FIXED = 5; // B, D, three rotating H planes
constexpr int CL_PLANES_RDPK = 8;  // B, D, three rotating planes (u, S1, S1'), S2, est, k1
constexpr int CL_PLANES_CA = 14;   // B, D, RDPK planes 2-7 (lambda), adjoint node planes 8-10, snapshots H_a, H_b and their interpolant H_t
constexpr int CL_PLANES_REV = 8;   // B, D, two rotating lambda planes, H_j, and the node planes alpha D+, beta dSx D+, beta dSy D+
constexpr int CL_MAX_CS = 16;

__host__ __device__ inline int cl_pitch(int nx) { return ((nx + 3) & ~3) + 2 * CL_PAD; }
__host__ __device__ inline int cl_band_rows(int ny, int cs) { return (ny + cs - 1) / cs; }
// bytes of dynamic shared memory one CTA needs for a glacier: n_planes planes of (R + 2) rows
inline size_t cl_smem_bytes(int nx, int ny, int cs, size_t esize, int n_planes) {
    return (size_t)n_planes * (size_t)(cl_band_rows(ny, cs) + 2) * (size_t)cl_pitch(nx) * esize;
}

// V consecutive cells of one row: the unit of work of a thread.  V = 4 (16-byte accesses in fp32); bands with fewer than 256 such
// items run with V = 2 so that more warps share the sweep (64 x 64 on 16 CTAs: 0.90 -> 0.76 us per stage).
template <typename T, int V> struct Vec { T v[V]; };
template <int V> __device__ __forceinline__ Vec<float, V> ldv(const float* p) {
    Vec<float, V> r;
    if constexpr (V == 4) { float4 q = *reinterpret_cast<const float4*>(p); r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w; }
    else if constexpr (V == 2) { float2 q = *reinterpret_cast<const float2*>(p); r.v[0] = q.x; r.v[1] = q.y; }
    else r.v[0] = *p;
    return r;
}
template <int V> __device__ __forceinline__ Vec<double, V> ldv(const double* p) {
    Vec<double, V> r;
    if constexpr (V == 4) {
        double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
    } else if constexpr (V == 2) { double2 a = *reinterpret_cast<const double2*>(p); r.v[0] = a.x; r.v[1] = a.y; }
    else r.v[0] = *p;
    return r;
}
template <int V> __device__ __forceinline__ void stv(float* p, const Vec<float, V>& q) {
    if constexpr (V == 4) *reinterpret_cast<float4*>(p) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]);
    else if constexpr (V == 2) *reinterpret_cast<float2*>(p) = make_float2(q.v[0], q.v[1]);
    else *p = q.v[0];
}
template <int V> __device__ __forceinline__ void stv(double* p, const Vec<double, V>& q) {
    if constexpr (V == 4) {
        *reinterpret_cast<double2*>(p) = make_double2(q.v[0], q.v[1]);
        *reinterpret_cast<double2*>(p + 2) = make_double2(q.v[2], q.v[3]);
    } else if constexpr (V == 2) *reinterpret_cast<double2*>(p) = make_double2(q.v[0], q.v[1]);
    else *p = q.v[0];
}

// sweep over (row, first column of an item) pairs of `nrows` rows, CL_NT items at a time, without a division per iteration
#define CL_SWEEP(ROW, COL0, NROWS)                                                                         \
    for (int ROW = t_row, cq_ = t_col, COL0 = t_col * V; ROW < (NROWS);                                    \
         cq_ += d_col, ROW += d_row, ROW += (cq_ >= Q), cq_ -= (cq_ >= Q) ? Q : 0, COL0 = cq_ * V)

// Synchronisation between the bands is ONE hardware cluster barrier per pass (barrier.cluster arrive.release / wait.acquire).  Two
// alternatives were measured and dropped (profiles/r02_cluster_sweep.txt, 128 x 128 on 16 CTAs, us per stage): a neighbour-only
// hand-shake through four mbarriers per CTA (remote mbarrier.arrive.release.cluster after a CTA barrier, try_wait.acquire.cluster)
// 1.21 vs 1.12; arriving early and evaluating the dual nodes that touch own rows only before the wait 1.17 vs 1.12 (0.90 vs 0.74 at
// 64 x 64).  A stage costs ~0.75 us even for a 64 x 64 grid: two dependent shared-memory sweeps, the release of the remote stores
// and the barrier -- not instruction issue (one, two or four cells per thread time the same there).
// Sub-gradient of the flux clamp on one edge (subgrad of sia2d_march.cuh without the warp vote: the sweeps here are not
// warp-convergent).  fp64 re-checks near-ties with true divisions to take the reference's branch (inversion_utils.jl:24-28, 38-42).
template <typename T, bool ETA1>
__device__ __forceinline__ void cl_subgrad(T dC, T e, T lo, T up, T delta, T eta0, T& to_lower, T& to_upper) {
    bool gt_lo = e > lo, lt_lo = lo > e, lt_up = up > e, gt_up = e > up;
    if (sizeof(T) == 8) {
        const T tol = T(1e-14) * (e < T(0) ? -e : e);
        const T d1 = e - lo, d2 = up - e;
        if ((d1 < tol && -d1 < tol) || (d2 < tol && -d2 < tol)) {
            const T qe = e / delta, ql = lo / delta, qu = up / delta;
            gt_lo = qe > ql; lt_lo = ql > qe; lt_up = qu > qe; gt_up = qe > qu;
        }
    }
    subgrad_cmp<T, ETA1>(dC, gt_lo, lt_lo, lt_up, gt_up, eta0, to_lower, to_upper);
}

// One CTA's band of one glacier.  Plane k of the carve-up is sm + k * plane; plane 0 is B, plane 1 is D (node row m of the band
// in local row m), the others belong to the scheme.  Local row l <-> grid row row0 - 1 + l (l = 0 and l = Rown + 1: halo rows).
template <typename T, bool CUBIC, bool ETA1, int V>
struct ClBand {
    int nx, ny, P, R, Q, row0, Rown, CS, rank, tid;
    int t_row, t_col, d_row, d_col;   // this thread's first (row, item) of a sweep and its stride
    size_t plane;
    T *sm, *sB, *sD, *sm_lo, *sm_hi;
    T eta0, hdx, hdy, kx, ky, A;
    PhysDev<T> ph;
    long long goff;
    int gld;
    int pAdj;   // first of the three node planes of the adjoint sweeps (alpha D+, beta dSx D+, beta dSy D+)

    __device__ __forceinline__ void init(cg::cluster_group& cluster, const GDesc<T>& d, const PhysDev<T>& phys, unsigned char* raw, int n_planes) {
        CS = (int)cluster.num_blocks();
        rank = (int)cluster.block_rank();
        tid = threadIdx.x;
        nx = d.nx; ny = d.ny;
        P = cl_pitch(nx); R = cl_band_rows(ny, CS); Q = (nx + V - 1) / V;
        row0 = min(rank * R, ny);
        Rown = min(row0 + R, ny) - row0;
        plane = (size_t)(R + 2) * P;
        sm = reinterpret_cast<T*>(raw);
        sB = sm;
        sD = sm + plane;
        // the same carve-up in the neighbours' shared memory (halo rows of the stage values)
        sm_lo = (rank > 0 && Rown > 0) ? cluster.map_shared_rank(sm, rank - 1) : nullptr;
        sm_hi = (rank + 1 < CS && row0 + Rown < ny) ? cluster.map_shared_rank(sm, rank + 1) : nullptr;
        ph = phys;
        eta0 = phys.eta0;
        hdx = T(0.5) * d.inv_dx; hdy = T(0.5) * d.inv_dy;
        kx = hdx * d.inv_dx; ky = hdy * d.inv_dy;   // ½/Δx², ½/Δy²
        A = d.A;
        goff = d.off; gld = d.ld;
        pAdj = 5;
        t_row = tid / Q; t_col = tid - t_row * Q; d_row = CL_NT / Q; d_col = CL_NT - d_row * Q;
        for (size_t k = tid; k < (size_t)n_planes * plane; k += CL_NT) sm[k] = T(0);
        __syncthreads();
    }
    __device__ __forceinline__ T* pl(int k) const { return sm + (size_t)k * plane; }

    // band rows row0-1 .. row1 (clipped to the grid) of a global plane into shared plane k (16-byte accesses: the rows of the
    // global planes are padded to 32 elements with zeros, a full quad is always readable)
    __device__ __forceinline__ void load(int k, const T* __restrict__ g) {
        T* dst = pl(k);
        const int Q4 = (nx + 3) >> 2;
        for (int q = tid; q < (Rown + 2) * Q4; q += CL_NT) {
            const int l = q / Q4, i0 = (q - l * Q4) << 2, j = row0 - 1 + l;
            if (Rown == 0 || j < 0 || j >= ny) continue;
            stv<4>(dst + (size_t)l * P + CL_PAD + i0, ldv<4>(g + goff + (long long)j * gld + i0));
        }
    }
    // own rows of shared plane k to one or two global planes (the row padding receives zeros)
    __device__ __forceinline__ void store(int k, T* __restrict__ g0, T* __restrict__ g1) {
        const T* src = pl(k);
        const int Q4 = (nx + 3) >> 2;
        for (int q = tid; q < Rown * Q4; q += CL_NT) {
            const int lr = q / Q4, i0 = (q - lr * Q4) << 2;
            const Vec<T, 4> v = ldv<4>(src + (size_t)(lr + 1) * P + CL_PAD + i0);
            const long long go = goff + (long long)(row0 + lr) * gld + i0;
            if (g0) stv<4>(g0 + go, v);
            if (g1) stv<4>(g1 + go, v);
        }
    }
    // an item of local row l of plane k; the first / last own row also goes to the neighbour's halo row
    __device__ __forceinline__ void put(int k, int l, size_t o, const Vec<T, V>& v) {
        stv<V>(pl(k) + o, v);
        if (l == 1 && sm_lo) stv<V>(sm_lo + (size_t)k * plane + o + (size_t)R * P, v);        // its local row R + 1
        if (l == Rown && sm_hi) stv<V>(sm_hi + (size_t)k * plane + o - (size_t)Rown * P, v);  // its local row 0
    }
    // ---- nodes: local node row m <-> grid node row row0 - 1 + m, between cell rows m and m + 1 ----
    __device__ __forceinline__ void nodes(const T* __restrict__ cur) {
        CL_SWEEP(m, a0, Rown + 1) {
            const int b = row0 - 1 + m;
            if (Rown == 0 || b < 0 || b > ny - 2) continue;
            const T* c0 = cur + (size_t)m * P + CL_PAD + a0;
            const T* c1 = c0 + P;
            const T* b0 = sB + (size_t)m * P + CL_PAD + a0;
            const T* b1 = b0 + P;
            const Vec<T, V> h0 = ldv<V>(c0), h1 = ldv<V>(c1), z0 = ldv<V>(b0), z1 = ldv<V>(b1);
            T h0e[V + 1], h1e[V + 1], s0e[V + 1], s1e[V + 1];
#pragma unroll
            for (int k = 0; k < V; ++k) { h0e[k] = h0.v[k]; h1e[k] = h1.v[k]; s0e[k] = z0.v[k]; s1e[k] = z1.v[k]; }
            h0e[V] = c0[V]; h1e[V] = c1[V]; s0e[V] = b0[V]; s1e[V] = b1[V];
#pragma unroll
            for (int k = 0; k < V + 1; ++k) {
                h0e[k] = fmx(h0e[k], T(0));                     // adjoint.jl:52
                h1e[k] = fmx(h1e[k], T(0));
                s0e[k] = surf_store<T>(s0e[k], h0e[k]);        // fp64: S = B + H rounded first (adjoint.jl:54); fp32: B kept
                s1e[k] = surf_store<T>(s1e[k], h1e[k]);
            }
            Vec<T, V> Dq;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const T ex0 = sdiff<T>(s0e[k + 1], s0e[k], h0e[k + 1], h0e[k]);
                const T ex1 = sdiff<T>(s1e[k + 1], s1e[k], h1e[k + 1], h1e[k]);
                const T ey0 = sdiff<T>(s1e[k], s0e[k], h1e[k], h0e[k]);
                const T ey1 = sdiff<T>(s1e[k + 1], s0e[k + 1], h1e[k + 1], h0e[k + 1]);
                const T u = (ex0 + ex1) * hdx, v = (ey0 + ey1) * hdy;   // ∇Sx, ∇Sy at the node (adjoint.jl:58-67)
                const T g2 = u * u + v * v;
                T Dn, al, be, gA;
                node_raw<T, CUBIC, false>(ph, A, (h0e[k] + h0e[k + 1]) + (h1e[k] + h1e[k + 1]), g2, Dn, al, be, gA);
                Dq.v[k] = (a0 + k <= nx - 2) ? Dn : T(0);
            }
            stv<V>(sD + (size_t)m * P + CL_PAD + a0, Dq);
        }
        __syncthreads();
    }

    // ---- cells of the band: ep(l, o, hc, f) with hc the item of `cur` and f = SIA2D(cur) on it (0 on the border) ----
    template <class Ep>
    __device__ __forceinline__ void cells(const T* __restrict__ cur, Ep&& ep) {
        CL_SWEEP(lr, i0, Rown) {
            const int l = lr + 1, j = row0 + lr;
            const size_t o = (size_t)l * P + CL_PAD + i0;
            const Vec<T, V> hc = ldv<V>(cur + o);
            Vec<T, V> f;
            if (j >= 1 && j <= ny - 2) {
                const Vec<T, V> hs = ldv<V>(cur + o - P), hn = ldv<V>(cur + o + P);
                const Vec<T, V> zc = ldv<V>(sB + o), zs = ldv<V>(sB + o - P), zn = ldv<V>(sB + o + P);
                const Vec<T, V> Ds = ldv<V>(sD + o - P), Dc = ldv<V>(sD + o);   // node rows j-1 and j, nodes i0 .. i0+V-1
                const T DsW = sD[o - P - 1], DcW = sD[o - 1];                   // node i0-1
                T he[V + 2], se[V + 2];
                he[0] = fmx(cur[o - 1], T(0));
                he[V + 1] = fmx(cur[o + V], T(0));
                se[0] = surf_store<T>(sB[o - 1], he[0]);
                se[V + 1] = surf_store<T>(sB[o + V], he[V + 1]);
#pragma unroll
                for (int k = 0; k < V; ++k) { he[k + 1] = fmx(hc.v[k], T(0)); se[k + 1] = surf_store<T>(zc.v[k], he[k + 1]); }
                // x-edge fluxes (raw): edge e joins cells i0-1+e and i0+e                      (adjoint.jl:93-97)
                T Fx[V + 1];
#pragma unroll
                for (int e = 0; e < V + 1; ++e) {
                    const T ex = sdiff<T>(se[e + 1], se[e], he[e + 1], he[e]);
                    const T up = ETA1 ? he[e + 1] : eta0 * he[e + 1], lo = ETA1 ? he[e] : eta0 * he[e];
                    const T Dsum = (e == 0) ? (DsW + DcW) : (Ds.v[e > 0 ? e - 1 : 0] + Dc.v[e > 0 ? e - 1 : 0]);
                    Fx[e] = Dsum * fmx(fmn(ex, up), -lo);
                }
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const T h = he[k + 1], s = se[k + 1];
                    const T hS = fmx(hs.v[k], T(0)), hN = fmx(hn.v[k], T(0));
                    const T sS = surf_store<T>(zs.v[k], hS), sN = surf_store<T>(zn.v[k], hN);
                    const T eyN = sdiff<T>(sN, s, hN, h), eyS = sdiff<T>(s, sS, h, hS);
                    const T eh = ETA1 ? h : eta0 * h, ehN = ETA1 ? hN : eta0 * hN, ehS = ETA1 ? hS : eta0 * hS;
                    const T DW_c = (k == 0) ? DcW : Dc.v[k > 0 ? k - 1 : 0], DW_s = (k == 0) ? DsW : Ds.v[k > 0 ? k - 1 : 0];
                    const T FyN = (DW_c + Dc.v[k]) * fmx(fmn(eyN, ehN), -eh);
                    const T FyS = (DW_s + Ds.v[k]) * fmx(fmn(eyS, eh), -ehS);
                    // dH = ½/Δx² ΔFx_raw + ½/Δy² ΔFy_raw  (see RhsMarch::step)
                    const T fv = kx * (Fx[k + 1] - Fx[k]) + ky * (FyN - FyS);
                    const int i = i0 + k;
                    f.v[k] = (i >= 1 && i <= nx - 2) ? fv : T(0);
                }
            } else {
#pragma unroll
                for (int k = 0; k < V; ++k) f.v[k] = T(0);   // border rows: dH = 0
            }
            ep(l, o, hc, f);
        }
    }
    // ---- discrete adjoint (A1 / A2), adjoint.jl:99-148, 235-250; oracle VJP_dSIA_dH_discrete / node_reduction_S ----
    // lambda is used zero-extended outside the interior cells (lambda_inn, adjoint.jl:99): with it every edge / node term that does
    // not exist on the border evaluates to zero by itself, so the sweeps carry no existence tests.
    __device__ __forceinline__ bool col_in(int i) const { return i >= 1 && i <= nx - 2; }
    __device__ __forceinline__ bool row_in(int j) const { return j >= 1 && j <= ny - 2; }

    // Node sweep.  MODE 0 (A1): planes 1, 5, 6, 7 <- D, alpha D+, beta dSx D+, beta dSy D+ for every node row of the band.
    // MODE 1 (A2): acc += scale * sum over the node rows this CTA OWNS (local rows 1 .. Rown) of  gA D+  (target_A.jl:71-72).
    template <int MODE>
    __device__ __forceinline__ void adj_nodes(const T* __restrict__ lam, const T* __restrict__ Hp, double& acc, double scale) {
        const T inv_dx = T(2) * hdx, inv_dy = T(2) * hdy;
        T* sP = pl(pAdj); T* sQx = pl(pAdj + 1); T* sQy = pl(pAdj + 2);
        CL_SWEEP(r, a0, MODE == 0 ? Rown + 1 : Rown) {
            const int m = MODE == 0 ? r : r + 1;
            const int b = row0 - 1 + m;
            if (Rown == 0 || b < 0 || b > ny - 2) continue;
            const size_t o0 = (size_t)m * P + CL_PAD + a0, o1 = o0 + P;
            const Vec<T, V> h0 = ldv<V>(Hp + o0), h1 = ldv<V>(Hp + o1), z0 = ldv<V>(sB + o0), z1 = ldv<V>(sB + o1);
            const Vec<T, V> q0 = ldv<V>(lam + o0), q1 = ldv<V>(lam + o1);
            T h0e[V + 1], h1e[V + 1], s0e[V + 1], s1e[V + 1], l0e[V + 1], l1e[V + 1];
#pragma unroll
            for (int k = 0; k < V; ++k) { h0e[k] = h0.v[k]; h1e[k] = h1.v[k]; s0e[k] = z0.v[k]; s1e[k] = z1.v[k]; l0e[k] = q0.v[k]; l1e[k] = q1.v[k]; }
            h0e[V] = Hp[o0 + V]; h1e[V] = Hp[o1 + V]; s0e[V] = sB[o0 + V]; s1e[V] = sB[o1 + V]; l0e[V] = lam[o0 + V]; l1e[V] = lam[o1 + V];
            const bool rb0 = row_in(b), rb1 = row_in(b + 1);
#pragma unroll
            for (int k = 0; k < V + 1; ++k) {
                h0e[k] = fmx(h0e[k], T(0));
                h1e[k] = fmx(h1e[k], T(0));
                s0e[k] = surf_store<T>(s0e[k], h0e[k]);
                s1e[k] = surf_store<T>(s1e[k], h1e[k]);
                const bool ci = col_in(a0 + k);
                l0e[k] = (rb0 && ci) ? l0e[k] : T(0);
                l1e[k] = (rb1 && ci) ? l1e[k] : T(0);
            }
            Vec<T, V> Dq, Pq, Qxq, Qyq;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const T ex0 = sdiff<T>(s0e[k + 1], s0e[k], h0e[k + 1], h0e[k]);
                const T ex1 = sdiff<T>(s1e[k + 1], s1e[k], h1e[k + 1], h1e[k]);
                const T ey0 = sdiff<T>(s1e[k], s0e[k], h1e[k], h0e[k]);
                const T ey1 = sdiff<T>(s1e[k + 1], s0e[k + 1], h1e[k + 1], h0e[k + 1]);
                const T u = (ex0 + ex1) * hdx, v = (ey0 + ey1) * hdy;
                const T g2 = u * u + v * v;
                T Dn, al, be, gA;
                node_raw<T, CUBIC, true>(ph, A, (h0e[k] + h0e[k + 1]) + (h1e[k] + h1e[k + 1]), g2, Dn, al, be, gA);
                // D+ = avg_y+(-Fx+ cx) + avg_x+(-Fy+ cy),  Fx+ = diff_x+(-lambda_inn)                       (adjoint.jl:99-104)
                const T X0 = -((l0e[k + 1] - l0e[k]) * inv_dx) * (clamp_raw<T>(ex0, eta0, h0e[k], h0e[k + 1]) * inv_dx);
                const T X1 = -((l1e[k + 1] - l1e[k]) * inv_dx) * (clamp_raw<T>(ex1, eta0, h1e[k], h1e[k + 1]) * inv_dx);
                const T Y0 = -((l1e[k] - l0e[k]) * inv_dy) * (clamp_raw<T>(ey0, eta0, h0e[k], h1e[k]) * inv_dy);
                const T Y1 = -((l1e[k + 1] - l0e[k + 1]) * inv_dy) * (clamp_raw<T>(ey1, eta0, h0e[k + 1], h1e[k + 1]) * inv_dy);
                const T Dd = T(0.5) * (X0 + X1) + T(0.5) * (Y0 + Y1);
                const bool ok = (a0 + k <= nx - 2);
                if (MODE == 0) {
                    Dq.v[k] = ok ? Dn : T(0);
                    Pq.v[k] = ok ? al * Dd : T(0);
                    Qxq.v[k] = ok ? be * u * Dd : T(0);
                    Qyq.v[k] = ok ? be * v * Dd : T(0);
                } else if (ok) {
                    acc += scale * (double)(gA * Dd);
                }
            }
            if (MODE == 0) {
                stv<V>(sD + o0, Dq);
                stv<V>(sP + o0, Pq);
                stv<V>(sQx + o0, Qxq);
                stv<V>(sQy + o0, Qyq);
            }
        }
        __syncthreads();
    }

    // Cell sweep of A1: ep(l, o, i0, lam_item (raw), H_item (raw), v) with v = (dSIA/dH)^T lambda on the item      (adjoint.jl:106-148)
    template <class Ep>
    __device__ __forceinline__ void adj_cells(const T* __restrict__ lam, const T* __restrict__ Hp, Ep&& ep) {
        const T inv_dx = T(2) * hdx, inv_dy = T(2) * hdy;
        const T* sP = pl(pAdj); const T* sQx = pl(pAdj + 1); const T* sQy = pl(pAdj + 2);
        CL_SWEEP(lr, i0, Rown) {
            const int l = lr + 1, j = row0 + lr;
            const size_t o = (size_t)l * P + CL_PAD + i0;
            const Vec<T, V> hc = ldv<V>(Hp + o), hs = ldv<V>(Hp + o - P), hn = ldv<V>(Hp + o + P);
            const Vec<T, V> zc = ldv<V>(sB + o), zs = ldv<V>(sB + o - P), zn = ldv<V>(sB + o + P);
            const Vec<T, V> lc = ldv<V>(lam + o), ls = ldv<V>(lam + o - P), ln = ldv<V>(lam + o + P);
            // node rows j-1 (S) and j (C), nodes i0-1 .. i0+V-1: index q <-> node column i0 - 1 + q
            T DS[V + 1], DC[V + 1], PS[V + 1], PC[V + 1], XS[V + 1], XC[V + 1], YS[V + 1], YC[V + 1];
            {
                const Vec<T, V> a = ldv<V>(sD + o - P), b = ldv<V>(sD + o), c = ldv<V>(sP + o - P), d = ldv<V>(sP + o);
                const Vec<T, V> e = ldv<V>(sQx + o - P), f = ldv<V>(sQx + o), g = ldv<V>(sQy + o - P), h = ldv<V>(sQy + o);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    DS[k + 1] = a.v[k]; DC[k + 1] = b.v[k]; PS[k + 1] = c.v[k]; PC[k + 1] = d.v[k];
                    XS[k + 1] = e.v[k]; XC[k + 1] = f.v[k]; YS[k + 1] = g.v[k]; YC[k + 1] = h.v[k];
                }
                DS[0] = sD[o - P - 1]; DC[0] = sD[o - 1]; PS[0] = sP[o - P - 1]; PC[0] = sP[o - 1];
                XS[0] = sQx[o - P - 1]; XC[0] = sQx[o - 1]; YS[0] = sQy[o - P - 1]; YC[0] = sQy[o - 1];
            }
            // cell row j, columns i0-1 .. i0+V: index q <-> column i0 - 1 + q
            T he[V + 2], se[V + 2], lt[V + 2];
            const bool rj = row_in(j), rs = row_in(j - 1), rn = row_in(j + 1);
            he[0] = fmx(Hp[o - 1], T(0)); he[V + 1] = fmx(Hp[o + V], T(0));
            se[0] = surf_store<T>(sB[o - 1], he[0]); se[V + 1] = surf_store<T>(sB[o + V], he[V + 1]);
            lt[0] = (rj && col_in(i0 - 1)) ? lam[o - 1] : T(0);
            lt[V + 1] = (rj && col_in(i0 + V)) ? lam[o + V] : T(0);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                he[k + 1] = fmx(hc.v[k], T(0));
                se[k + 1] = surf_store<T>(zc.v[k], he[k + 1]);
                lt[k + 1] = (rj && col_in(i0 + k)) ? lc.v[k] : T(0);
            }
            Vec<T, V> out;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const bool ci = col_in(i0 + k);
                const T hS = fmx(hs.v[k], T(0)), hN = fmx(hn.v[k], T(0));
                const T sS = surf_store<T>(zs.v[k], hS), sN = surf_store<T>(zn.v[k], hN);
                const T lS = (rs && ci) ? ls.v[k] : T(0), lN = (rn && ci) ? ln.v[k] : T(0);
                // first term: avg+(alpha D+) + diff_x+(avg_y+(beta dSx D+)) + diff_y+(avg_x+(beta dSy D+))           (adjoint.jl:106-127)
                T acc = T(0.25) * ((PS[k] + PS[k + 1]) + (PC[k] + PC[k + 1]));
                acc += (T(0.5) * (XS[k] + XC[k]) - T(0.5) * (XS[k + 1] + XC[k + 1])) * inv_dx;
                acc += (T(0.5) * (YS[k] + YS[k + 1]) - T(0.5) * (YC[k] + YC[k + 1])) * inv_dy;
                // second term: the four edges of the cell through the clamp sub-gradient                            (adjoint.jl:129-144)
                T lo_, up_;
                {   // east edge: this cell is the lower one
                    const T dC = -((lt[k + 2] - lt[k + 1]) * inv_dx) * (T(0.5) * (DS[k + 1] + DC[k + 1])) * inv_dx;
                    cl_subgrad<T, ETA1>(dC, sdiff<T>(se[k + 2], se[k + 1], he[k + 2], he[k + 1]), -(eta0 * he[k + 1]), eta0 * he[k + 2], T(1) / inv_dx, eta0, lo_, up_);
                    acc += lo_;
                }
                {   // west edge: this cell is the upper one
                    const T dC = -((lt[k + 1] - lt[k]) * inv_dx) * (T(0.5) * (DS[k] + DC[k])) * inv_dx;
                    cl_subgrad<T, ETA1>(dC, sdiff<T>(se[k + 1], se[k], he[k + 1], he[k]), -(eta0 * he[k]), eta0 * he[k + 1], T(1) / inv_dx, eta0, lo_, up_);
                    acc += up_;
                }
                {   // north edge (row j -> j+1): lower
                    const T dC = -((lN - lt[k + 1]) * inv_dy) * (T(0.5) * (DC[k] + DC[k + 1])) * inv_dy;
                    cl_subgrad<T, ETA1>(dC, sdiff<T>(sN, se[k + 1], hN, he[k + 1]), -(eta0 * he[k + 1]), eta0 * hN, T(1) / inv_dy, eta0, lo_, up_);
                    acc += lo_;
                }
                {   // south edge (row j-1 -> j): upper
                    const T dC = -((lt[k + 1] - lS) * inv_dy) * (T(0.5) * (DS[k] + DS[k + 1])) * inv_dy;
                    cl_subgrad<T, ETA1>(dC, sdiff<T>(se[k + 1], sS, he[k + 1], hS), -(eta0 * hS), eta0 * he[k + 1], T(1) / inv_dy, eta0, lo_, up_);
                    acc += up_;
                }
                out.v[k] = (hc.v[k] > T(0)) ? acc : T(0);   // adjoint.jl:148
            }
            ep(l, o, i0, lc, hc, out);
        }
    }

    // elementwise sweep over the own cells: ep(l, o)
    template <class Ep>
    __device__ __forceinline__ void own(Ep&& ep) {
        CL_SWEEP(lr, i0, Rown) {
            ep(lr + 1, (size_t)(lr + 1) * P + CL_PAD + i0);
        }
    }
};

// method: 0 Euler, 1 SSPRK(3,3) (Shu-Osher form, as forward_interval in capi.cu).  Intervals j0+1 .. j1 of the time grid t.
template <typename T, bool CUBIC, bool ETA1, int V>
__global__ void __launch_bounds__(CL_NT, 1)
sia2d_interval_cluster(const GDesc<T>* __restrict__ descs, const T* __restrict__ Hin, const T* __restrict__ Bg, T* __restrict__ Hout,
                       T* __restrict__ snap, long long plane_stride, const double* __restrict__ t, int j0, int j1, int nsub,
                       int method, PhysDev<T> ph) {
    extern __shared__ __align__(16) unsigned char cl_smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    ClBand<T, CUBIC, ETA1, V> bd;
    bd.init(cluster, descs[blockIdx.x / cluster.num_blocks()], ph, cl_smem_raw, CL_PLANES_FIXED);
    bd.load(0, Bg);
    bd.load(2, Hin);
    cluster.sync();  // every CTA of the cluster has initialised its planes before a neighbour stores into them

    // out = sa·u0 + sb·(cur + sdt·SIA2D(cur)) on the band; halo rows of `nxt` in the neighbours; cluster barrier
    auto stage = [&](int cur, int u0, int nxt, T sa, T sb, T sdt) {
        const T* pc = bd.pl(cur);
        const T* pu = bd.pl(u0);
        bd.nodes(pc);
        bd.cells(pc, [&](int l, size_t o, const Vec<T, V>& hc, const Vec<T, V>& f) {
            Vec<T, V> out;
#pragma unroll
            for (int k = 0; k < V; ++k) out.v[k] = sb * (hc.v[k] + sdt * f.v[k]);
            if (sa != T(0)) {
                const Vec<T, V> uq = ldv<V>(pu + o);
#pragma unroll
                for (int k = 0; k < V; ++k) out.v[k] = sa * uq.v[k] + out.v[k];
            }
            bd.put(nxt, l, o, out);
        });
        cluster.sync();
    };

    int a = 2, b = 3, c = 4;   // plane a: state u0; planes b, c: stage values
    const T third = (T)(1.0 / 3.0), twothirds = (T)(2.0 / 3.0);
    for (int j = j0 + 1; j <= j1; ++j) {
        const T h = (T)((t[j] - t[j - 1]) / nsub);
        for (int s = 0; s < nsub; ++s) {
            if (method == 0) {
                stage(a, a, b, T(0), T(1), h);                 // H + h f(H)
            } else {
                stage(a, a, b, T(0), T(1), h);                 // u1 = H + h f(H)
                stage(b, a, c, T(0.75), T(0.25), h);           // u2 = 3/4 H + 1/4 (u1 + h f(u1))
                stage(c, a, b, third, twothirds, h);           // H  = 1/3 H + 2/3 (u2 + h f(u2))
            }
            const int tmp = a; a = b; b = tmp;
        }
        bd.store(a, snap ? snap + (long long)j * plane_stride : nullptr, j == j1 ? Hout : nullptr);   // snapshot j (and the final state)
    }
}

// -------------------------------------------------------------------------------------------------------------------------
// RDPK3Sp35 + PID controller (the scheme of rdpk.cu / oracle integrate_rdpk3sp35, step for step), one cluster per glacier.
// -------------------------------------------------------------------------------------------------------------------------
struct RdpkCoef { double G1[4], G2[4], G3[4], D[4], B[5], E[5], C[6]; };

// per-glacier controller state carried between launches (a launch range ends at a mass-balance callback)
struct ClRkState {
    double t, dt;
    double lerr2, lerr3;   // PID history as LOGARITHMS of 1 / EEst of the last two accepted steps (one log + one exp per step instead of three pow)
    int steps, rejected, started, pad;
};

// Fixed-order sum over the cluster of up to two values per thread: every CTA receives every CTA's partial through DSMEM and adds
// them in rank order, so all CTAs hold the same bits.  Contains one cluster barrier (which also orders the halo stores of the pass).
__device__ __forceinline__ void cluster_sum2(cg::cluster_group& cluster, double& v0, double& v1, double (*red)[2][CL_MAX_CS], double* sRed,
                                             int& slot) {
    const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const double s0 = block_sum(v0, sRed);
    __syncthreads();
    const double s1 = block_sum(v1, sRed + CL_NT / 32);
    if (threadIdx.x == 0) {
        for (int r = 0; r < CS; ++r) {
            double(*rr)[2][CL_MAX_CS] = cluster.map_shared_rank(red, r);
            rr[slot][0][rank] = s0;
            rr[slot][1][rank] = s1;
        }
    }
    cluster.sync();
    double a0 = 0.0, a1 = 0.0;
    for (int r = 0; r < CS; ++r) { a0 += red[slot][0][r]; a1 += red[slot][1][r]; }
    v0 = a0; v1 = a1;
    slot ^= 1;
}

// The RDPK3Sp35 machinery of one cluster, shared by the forward solve and the reverse ODE of the continuous adjoint.
// rhs(plane, time, ep): evaluate f(time, plane k) on the band and call ep(l, o, item of the plane, item of f) for every own item.
// Planes 2, 3, 4 rotate (a = state at the start of the step, b / c = stage values, all with halo rows); 5, 6, 7 = S2, est, k1.
// one value (the error norm of a trial step)
__device__ __forceinline__ void cluster_sum1(cg::cluster_group& cluster, double& v0, double (*red)[2][CL_MAX_CS], double* sRed, int& slot) {
    const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const double s0 = block_sum(v0, sRed);
    if (threadIdx.x == 0)
        for (int r = 0; r < CS; ++r) cluster.map_shared_rank(red, r)[slot][0][rank] = s0;
    cluster.sync();
    double a0 = 0.0;
    for (int r = 0; r < CS; ++r) a0 += red[slot][0][r];
    v0 = a0;
    slot ^= 1;
}

template <typename T, bool CUBIC, bool ETA1, int V>
struct ClRdpk {
    static constexpr int pS2 = 5, pE = 6, pK1 = 7;
    ClBand<T, CUBIC, ETA1, V>& bd;
    cg::cluster_group& cluster;
    const RdpkCoef& cf;
    double (*red)[2][CL_MAX_CS];
    double* sRed;
    double* ctl;
    double reltol, abstol, dtmax, ncell;
    int max_steps;
    int a = 2, b = 3, c = 4;
    bool k1_valid = false;
    ClRkState s;
    int slot = 0, total = 0;

    __device__ __forceinline__ ClRdpk(ClBand<T, CUBIC, ETA1, V>& bd_, cg::cluster_group& cl_, const RdpkCoef& cf_, double (*red_)[2][CL_MAX_CS],
                                      double* sRed_, double* ctl_, double reltol_, double abstol_, double dtmax_, int max_steps_)
        : bd(bd_), cluster(cl_), cf(cf_), red(red_), sRed(sRed_), ctl(ctl_), reltol(reltol_), abstol(abstol_), dtmax(dtmax_),
          ncell((double)bd_.nx * (double)bd_.ny), max_steps(max_steps_) {}

    // first step size: dt0 > 0 as given, else OrdinaryDiffEq's ode_determine_initdt (Hairer-Wanner); sk = abstol + |u| reltol
    template <class Rhs>
    __device__ __forceinline__ void start(double t0, double dt0, Rhs&& rhs) {
        s.started = 1;
        s.t = t0;
        s.lerr2 = s.lerr3 = 0.0;
        s.steps = s.rejected = 0;
        if (dt0 > 0.0) {
            s.dt = fmin(dt0, dtmax);
            return;
        }
        double d0 = 0.0, d1 = 0.0;
        const T* pu = bd.pl(a);
        rhs(a, t0, [&](int, size_t o, const Vec<T, V>& hc, const Vec<T, V>& f) {
            stv<V>(bd.pl(pK1) + o, f);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const double sk = abstol + fabs((double)hc.v[k]) * reltol;
                const double r0 = (double)hc.v[k] / sk, r1 = (double)f.v[k] / sk;
                d0 += r0 * r0; d1 += r1 * r1;
            }
        });
        cluster_sum2(cluster, d0, d1, red, sRed, slot);
        d0 = sqrt(d0 / ncell); d1 = sqrt(d1 / ncell);
        const double h0 = fmin((d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1, dtmax);
        const T h0T = (T)(1.0 * h0);
        bd.own([&](int l, size_t o) {   // Euler probe u1 = u + h0 f0
            const Vec<T, V> uq = ldv<V>(pu + o), kq = ldv<V>(bd.pl(pK1) + o);
            Vec<T, V> x;
#pragma unroll
            for (int k = 0; k < V; ++k) x.v[k] = uq.v[k] + h0T * kq.v[k];
            bd.put(b, l, o, x);
        });
        cluster.sync();
        double d2 = 0.0, dz = 0.0;
        rhs(b, t0 + h0, [&](int, size_t o, const Vec<T, V>&, const Vec<T, V>& f) {
            const Vec<T, V> uq = ldv<V>(pu + o), kq = ldv<V>(bd.pl(pK1) + o);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const double sk = abstol + fabs((double)uq.v[k]) * reltol;
                const double r = (double)(f.v[k] - kq.v[k]) / sk;
                d2 += r * r;
            }
        });
        cluster_sum2(cluster, d2, dz, red, sRed, slot);
        d2 = sqrt(d2 / ncell) / h0;
        const double md = fmax(d1, d2);
        const double h1 = (md <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(10.0, -(2.0 + log10(md)) / 3.0);
        s.dt = fmin(fmin(100.0 * h0, h1), dtmax);
        k1_valid = true;
    }

    // the `while t < tstop` loop of integrate_rdpk3sp35: trial steps until the state (plane a) has landed on tstop
    template <class Rhs>
    __device__ __forceinline__ void advance(double tstop, Rhs&& rhs) {
        while (s.t < tstop) {
            if (++total > max_steps) break;
            // plan the step (rk_plan_step)
            double h = fmin(fmin(s.dt, dtmax), tstop - s.t);
            const bool last = (s.t + h >= tstop) || (tstop - (s.t + h) < 1e-14 * fmax(1.0, fabs(tstop)));
            if (last) h = tstop - s.t;
            const T* pu = bd.pl(a);
            {   // S1 = u + (B1 h) k1 ;  est = (E1 h) k1      with k1 = f(t, u) (re-evaluated after an accepted step: FSAL)
                const T bh = (T)(cf.B[0] * h), eh = (T)(cf.E[0] * h);
                if (!k1_valid) {
                    rhs(a, s.t, [&](int l, size_t o, const Vec<T, V>& hc, const Vec<T, V>& f) {
                        Vec<T, V> x, y;
#pragma unroll
                        for (int k = 0; k < V; ++k) { x.v[k] = hc.v[k] + bh * f.v[k]; y.v[k] = eh * f.v[k]; }
                        stv<V>(bd.pl(pK1) + o, f);
                        stv<V>(bd.pl(pE) + o, y);
                        bd.put(b, l, o, x);
                    });
                    k1_valid = true;
                } else {
                    bd.own([&](int l, size_t o) {
                        const Vec<T, V> uq = ldv<V>(pu + o), kq = ldv<V>(bd.pl(pK1) + o);
                        Vec<T, V> x, y;
#pragma unroll
                        for (int k = 0; k < V; ++k) { x.v[k] = uq.v[k] + bh * kq.v[k]; y.v[k] = eh * kq.v[k]; }
                        stv<V>(bd.pl(pE) + o, y);
                        bd.put(b, l, o, x);
                    });
                }
                cluster.sync();
            }
            int cur = b, nxt = c;
            double acc = 0.0;
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                // k = f(t + c h, S1);  S2 = S2in + d S1;  S1 = g1 S1 + g2 S2 + g3 u + (b h) k;  est += (e h) k       (rk_stage)
                const T g1 = (T)cf.G1[i], g2 = (T)cf.G2[i], g3 = (T)cf.G3[i], dd = (T)cf.D[i];
                const T bh = (T)(cf.B[i + 1] * h), eh = (T)(cf.E[i + 1] * h);
                const bool use_u = (cf.G3[i] != 0.0);
                rhs(cur, s.t + cf.C[i + 1] * h, [&](int l, size_t o, const Vec<T, V>& s1, const Vec<T, V>& f) {
                    const Vec<T, V> uq = ldv<V>(pu + o), er = ldv<V>(bd.pl(pE) + o);
                    const Vec<T, V> s2in = (i == 0) ? uq : ldv<V>(bd.pl(pS2) + o);
                    Vec<T, V> s2, x, y;
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        s2.v[k] = s2in.v[k] + dd * s1.v[k];
                        T v = g1 * s1.v[k] + g2 * s2.v[k];
                        if (use_u) v = v + g3 * uq.v[k];
                        x.v[k] = v + bh * f.v[k];
                        y.v[k] = er.v[k] + eh * f.v[k];
                    }
                    if (i < 3) stv<V>(bd.pl(pS2) + o, s2);
                    stv<V>(bd.pl(pE) + o, y);
                    bd.put(nxt, l, o, x);
                    if (i == 3) {   // error norm of the step: sum (est / (abstol + reltol max(|u|, |u_new|)))^2        (rk_sumsq)
#pragma unroll
                        for (int k = 0; k < V; ++k) {
                            const double m = fmax(fabs((double)uq.v[k]), fabs((double)x.v[k]));
                            const double r = (double)y.v[k] / (abstol + reltol * m);
                            acc += r * r;
                        }
                    }
                });
                cluster.sync();
                const int tmp = cur; cur = nxt; nxt = tmp;
            }
            cluster_sum1(cluster, acc, red, sRed, slot);
            // PID controller (rk_control): factor = e1^(b1/3) e2^(b2/3) e3^(b3/3) with e = 1 / EEst, evaluated through the logarithms by
            // thread 0 of every CTA on the same bits
            if (threadIdx.x == 0) {
                const double EEst = sqrt(acc / ncell);
                const double l1 = -log(fmax(EEst, 1e-300));
                double fac = exp((0.64 / 3.0) * l1 + (-0.31 / 3.0) * s.lerr2 + (0.04 / 3.0) * s.lerr3);
                fac = 1.0 + atan(fac - 1.0);
                ctl[0] = fac;
                ctl[1] = l1;
            }
            __syncthreads();
            const double fac = ctl[0], l1 = ctl[1];
            __syncthreads();
            s.steps++;
            if (fac >= 0.81) {
                s.t = last ? tstop : s.t + h;
                s.lerr3 = s.lerr2;
                s.lerr2 = l1;
                s.dt = h * fac;
                // u <- S1 (plane `cur` after the last swap); the old u plane becomes a stage plane
                const int old_a = a;
                a = cur;
                b = old_a;
                c = nxt;
                k1_valid = false;
            } else {
                s.rejected++;
                s.dt = h * fac;
                b = cur; c = nxt;   // (any two planes other than a)
                if (b == a || c == a) { b = (a == 2) ? 3 : 2; c = 9 - a - b; }
            }
        }
    }
};

template <typename T, bool CUBIC, bool ETA1, int V>
__global__ void __launch_bounds__(CL_NT, 1)
sia2d_rdpk_cluster(const GDesc<T>* __restrict__ descs, const T* __restrict__ Hin, const T* __restrict__ Bg, T* __restrict__ Hout,
                   T* __restrict__ snap, long long plane_stride, const double* __restrict__ t, int j0, int j1, ClRkState* __restrict__ states,
                   double reltol, double abstol, double dtmax, double dt0, int max_steps, RdpkCoef cf, PhysDev<T> ph) {
    extern __shared__ __align__(16) unsigned char cl_smem_raw[];
    __shared__ double red[2][2][CL_MAX_CS];
    __shared__ double sRed[2 * CL_NT / 32];
    __shared__ double ctl[4];
    cg::cluster_group cluster = cg::this_cluster();
    const int g = blockIdx.x / cluster.num_blocks();
    ClBand<T, CUBIC, ETA1, V> bd;
    bd.init(cluster, descs[g], ph, cl_smem_raw, CL_PLANES_RDPK);
    ClRdpk<T, CUBIC, ETA1, V> rk(bd, cluster, cf, red, sRed, ctl, reltol, abstol, dtmax, max_steps);
    bd.load(0, Bg);
    bd.load(rk.a, Hin);
    cluster.sync();

    // dH/dt = SIA2D(H): autonomous
    auto rhs = [&](int plane, double, auto&& ep) {
        const T* pc = bd.pl(plane);
        bd.nodes(pc);
        bd.cells(pc, ep);
    };
    rk.s = states[g];
    if (!rk.s.started) rk.start(t[j0], dt0, rhs);
    for (int j = j0 + 1; j <= j1; ++j) {
        rk.s.t = t[j - 1];
        rk.advance(t[j], rhs);
        bd.store(rk.a, snap ? snap + (long long)j * plane_stride : nullptr, j == j1 ? Hout : nullptr);
    }
    if (rk.total > max_steps) rk.s.started = -1;   // maxiters: reported by the launcher
    if (cluster.block_rank() == 0 && threadIdx.x == 0) states[g] = rk.s;
}

// -------------------------------------------------------------------------------------------------------------------------
// Discrete-adjoint reverse loop (gradient.jl:191-253 with LossH, Losses.jl:270-291), one cluster per glacier: steps jhi .. jlo+1 in
// ONE launch.  Per step j: H_j from the snapshot plane;  v = (dSIA/dH)^T lambda_j;  loss += w_j sum W (H_j - H_ref,j)^2;
// lambda_{j-1} = lambda_j + dt v + 2 w_j W (H_j - H_ref,j)  (halo rows to the neighbours, cluster barrier);
// S += dt * sum gA D+(lambda_{j-1}, H_j).  The loss and S accumulate per thread over all steps and are reduced once at the end.
// -------------------------------------------------------------------------------------------------------------------------
template <typename T, bool CUBIC, bool ETA1, int V>
__global__ void __launch_bounds__(CL_NT, 1)
sia2d_reverse_cluster(const GDesc<T>* __restrict__ descs, const T* __restrict__ Bg, const T* __restrict__ lam_in, T* __restrict__ lam_out,
                      const T* __restrict__ snap, const T* __restrict__ href, const T* __restrict__ wmask, long long plane_stride,
                      const double* __restrict__ t, const double* __restrict__ wH, int jhi, int jlo, double* __restrict__ loss_acc,
                      double* __restrict__ S_acc, PhysDev<T> ph) {
    extern __shared__ __align__(16) unsigned char cl_smem_raw[];
    __shared__ double red[2][2][CL_MAX_CS];
    __shared__ double sRed[2 * CL_NT / 32];
    cg::cluster_group cluster = cg::this_cluster();
    const int g = blockIdx.x / cluster.num_blocks();
    ClBand<T, CUBIC, ETA1, V> bd;
    bd.init(cluster, descs[g], ph, cl_smem_raw, CL_PLANES_REV);
    int la = 2, lb = 3;
    constexpr int pH = 4;
    bd.load(0, Bg);
    if (lam_in) bd.load(la, lam_in);   // (else lambda_k = 0, gradient.jl:140: the planes are zero-initialised)
    cluster.sync();

    double acc_loss = 0.0, acc_S = 0.0, dummy = 0.0;
    for (int j = jhi; j > jlo; --j) {
        const double dt = t[j] - t[j - 1], w = wH[j];
        const T dtT = (T)dt, cseed = (T)(2.0 * w);
        const long long pj = (long long)j * plane_stride;
        __syncthreads();                       // the A2 sweep of the previous step has finished reading the H plane
        bd.load(pH, snap + pj);
        __syncthreads();
        const T* lam = bd.pl(la);
        const T* Hp = bd.pl(pH);
        bd.template adj_nodes<0>(lam, Hp, dummy, 0.0);
        bd.adj_cells(lam, Hp, [&](int l, size_t o, int i0, const Vec<T, V>& lc, const Vec<T, V>& hc, const Vec<T, V>& v) {
            // H_ref and W of the item straight from the global planes (read once per cell and step)
            const long long go = bd.goff + (long long)(bd.row0 + l - 1) * bd.gld + i0;
            const Vec<T, V> hr = ldv<V>(href + pj + go), wm = ldv<V>(wmask + pj + go);
            Vec<T, V> x;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const T d = hc.v[k] - hr.v[k], wd = wm.v[k] * d;
                acc_loss += w * ((double)wd * (double)d);
                x.v[k] = lc.v[k] + dtT * v.v[k] + cseed * wd;
            }
            bd.put(lb, l, o, x);
        });
        cluster.sync();
        bd.template adj_nodes<1>(bd.pl(lb), Hp, acc_S, dt);   // dL/dtheta += dt VJP_theta(lambda_{j-1}, H_j)   (gradient.jl:245-249)
        const int tmp = la; la = lb; lb = tmp;
    }
    bd.store(la, lam_out, nullptr);
    int slot = 0;
    cluster_sum2(cluster, acc_loss, acc_S, red, sRed, slot);
    if (cluster.block_rank() == 0 && threadIdx.x == 0) {
        loss_acc[g] += acc_loss;
        S_acc[g] += acc_S;
    }
}

// -------------------------------------------------------------------------------------------------------------------------
// ContinuousAdjoint gradient with the reference's default reverse solve (gradient.jl:276-538, AdjointTypes.jl:53-66), one
// cluster per glacier, ONE launch: the reverse ODE  d(lambda)/d(tau) = (dSIA/dH)^T lambda  at  H_itp(-tau)  (linear interpolant of
// the forward snapshots) integrated with adaptive RDPK3Sp35 in tau = -t over the sorted union of the tstops and the Gauss-Legendre
// nodes; at a tstop the loss jump  lambda += dl_j/dH, loss += w_j l_j  (Losses.jl:270-291); at a quadrature node
// dL/dtheta-scalar += w_m sum gA D+(lambda, H_itp(t_m))  (gradient.jl:495-507).  LossH, glacier-wide A, discrete VJP flavour, no
// mass-balance callback (the host-driven engine of rdpk.cu covers the rest).  tab = t[n_t] | wH[n_t] | ev_tau[n_ev] | ev_code[n_ev]
// | qn[n_q] | qw[n_q];  ev_code = idx (+ 2^20 for a quadrature node).
// -------------------------------------------------------------------------------------------------------------------------
template <typename T, bool CUBIC, bool ETA1, int V>
__global__ void __launch_bounds__(CL_NT, 1)
sia2d_contadj_cluster(const GDesc<T>* __restrict__ descs, const T* __restrict__ Bg, T* __restrict__ lam_out, const T* __restrict__ snap,
                      const T* __restrict__ href, const T* __restrict__ wmask, long long plane_stride, const double* __restrict__ tab,
                      int n_t, int n_ev, int n_q, double reltol, double abstol, double dtmax, int max_steps, double* __restrict__ loss_acc,
                      double* __restrict__ S_acc, int* __restrict__ steps_out, RdpkCoef cf, PhysDev<T> ph) {
    extern __shared__ __align__(16) unsigned char cl_smem_raw[];
    __shared__ double red[2][2][CL_MAX_CS];
    __shared__ double sRed[2 * CL_NT / 32];
    __shared__ double ctl[4];
    cg::cluster_group cluster = cg::this_cluster();
    const int g = blockIdx.x / cluster.num_blocks();
    ClBand<T, CUBIC, ETA1, V> bd;
    bd.init(cluster, descs[g], ph, cl_smem_raw, CL_PLANES_CA);
    bd.pAdj = 8;
    ClRdpk<T, CUBIC, ETA1, V> rk(bd, cluster, cf, red, sRed, ctl, reltol, abstol, dtmax, max_steps);
    const double* t = tab;
    const double* wH = tab + n_t;
    const double* ev_tau = wH + n_t;
    const double* ev_code = ev_tau + n_ev;
    const double* qn = ev_code + n_ev;
    const double* qw = qn + n_q;
    int pHa = 11, pHb = 12;
    constexpr int pHt = 13;
    int jint = n_t - 2;   // the reverse solve is inside [t_jint, t_jint+1]: planes H_a = H(t_jint), H_b = H(t_jint+1)
    bd.load(0, Bg);
    bd.load(pHa, snap + (long long)jint * plane_stride);
    bd.load(pHb, snap + (long long)(jint + 1) * plane_stride);
    cluster.sync();

    double acc_loss = 0.0, acc_S = 0.0, dummy = 0.0;
    // H_t = (1 - a) H_a + a H_b on the band and its halo rows (rk_lerp)
    auto lerp_to = [&](double tt) {
        const T a1 = (T)((tt - t[jint]) / (t[jint + 1] - t[jint])), a0 = T(1) - a1;
        const T* Ha = bd.pl(pHa);
        const T* Hb = bd.pl(pHb);
        T* Ht = bd.pl(pHt);
        const int nq = (int)(((size_t)(bd.Rown + 2) * bd.P) >> 2);
        for (int q = threadIdx.x; q < nq; q += CL_NT) {
            const Vec<T, 4> x = ldv<4>(Ha + 4 * (size_t)q), y = ldv<4>(Hb + 4 * (size_t)q);
            Vec<T, 4> z;
#pragma unroll
            for (int k = 0; k < 4; ++k) z.v[k] = a0 * x.v[k] + a1 * y.v[k];
            stv<4>(Ht + 4 * (size_t)q, z);
        }
        __syncthreads();
    };
    auto rhs = [&](int plane, double tau, auto&& ep) {
        lerp_to(-tau);
        const T* pc = bd.pl(plane);
        const T* Ht = bd.pl(pHt);
        bd.template adj_nodes<0>(pc, Ht, dummy, 0.0);
        bd.adj_cells(pc, Ht, [&](int l, size_t o, int, const Vec<T, V>& lc, const Vec<T, V>&, const Vec<T, V>& v) { ep(l, o, lc, v); });
    };
    // effect_loss! at tstop j on the state plane (thickness term): lambda += 2 w_j W (H_j - H_ref,j), loss += w_j sum W (H_j - H_ref,j)^2
    auto loss_jump = [&](int j, int pHj) {
        const double w = wH[j];
        if (w == 0.0) return;
        const T cseed = (T)(2.0 * w);
        const long long pj = (long long)j * plane_stride;
        const T* Hj = bd.pl(pHj);
        const T* lam = bd.pl(rk.a);
        bd.own([&](int l, size_t o) {
            const int i0 = (int)(o - (size_t)l * bd.P) - CL_PAD;
            const long long go = bd.goff + (long long)(bd.row0 + l - 1) * bd.gld + i0;
            const Vec<T, V> hr = ldv<V>(href + pj + go), wm = ldv<V>(wmask + pj + go), hq = ldv<V>(Hj + o), lq = ldv<V>(lam + o);
            Vec<T, V> x;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const T d = hq.v[k] - hr.v[k], wd = wm.v[k] * d;
                acc_loss += w * ((double)wd * (double)d);
                x.v[k] = lq.v[k] + cseed * wd;
            }
            bd.put(rk.a, l, o, x);
        });
        cluster.sync();
        rk.k1_valid = false;
    };

    loss_jump(n_t - 1, pHb);                 // lambda(t_end) = dl/dH(t_end)   (gradient.jl:407-446)
    rk.start(ev_tau[0], 0.0, rhs);
    for (int i = 1; i < n_ev; ++i) {
        rk.s.t = ev_tau[i - 1];
        rk.advance(ev_tau[i], rhs);
        const int code = (int)ev_code[i], idx = code & ((1 << 20) - 1);
        if (code < (1 << 20)) {              // tstop idx: the loss jump; the solve continues in [t_{idx-1}, t_idx]
            loss_jump(idx, pHa);
            if (idx >= 1) {
                jint = idx - 1;
                const int tmp = pHa; pHa = pHb; pHb = tmp;   // H(t_idx) becomes the upper snapshot
                __syncthreads();
                bd.load(pHa, snap + (long long)jint * plane_stride);
                __syncthreads();
            }
        } else {                             // quadrature node idx
            lerp_to(qn[idx]);
            bd.template adj_nodes<1>(bd.pl(rk.a), bd.pl(pHt), acc_S, qw[idx]);
        }
    }
    bd.store(rk.a, lam_out, nullptr);
    cluster_sum2(cluster, acc_loss, acc_S, red, sRed, rk.slot);
    if (cluster.block_rank() == 0 && threadIdx.x == 0) {
        loss_acc[g] += acc_loss;
        S_acc[g] += acc_S;
        steps_out[g] = rk.total > max_steps ? -1 : rk.s.steps;
    }
}

}  // namespace odinn
