// Launchers of the per-cell law kernels (sia2d_law.cuh): node pass and theta pullback.
#include <cstdlib>

#include "launch.cuh"
#include "sia2d_law.cuh"
#include "sia2d_law_fixed.cuh"

namespace odinn {

static inline CellLaw* law_of(odinn_ensemble* e) { return static_cast<CellLaw*>(e->law_cfg); }

// The compile-time architecture of sia2d_law_fixed.cuh: 2 -> 16 -> 16 -> 1, softplus / softplus / sigmoid (ODINN_LAW_GENERIC=1: never).
static bool lf_match(const CellLaw& lw) {
    static const bool off = []() { const char* v = getenv("ODINN_LAW_GENERIC"); return v && v[0] == '1'; }();
    const MlpArch& a = lw.arch;
    return !off && a.n_layers == 3 && a.widths[0] == 2 && a.widths[1] == 16 && a.widths[2] == 16 && a.widths[3] == 1 &&
           a.acts[0] == ACT_SOFTPLUS && a.acts[1] == ACT_SOFTPLUS && a.acts[2] == ACT_SIGMOID;
}

static __global__ void law_theta_reduce_scaled(const double* __restrict__ block_partial, int n_tiles, int n_params,
                                        double* __restrict__ out, double scale, int accumulate) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_params) return;
    double s = 0.0;
    for (int t = 0; t < n_tiles; ++t) s += block_partial[(long long)t * n_params + k];  // tile order: deterministic
    out[k] = (accumulate ? out[k] : 0.0) + scale * s;
}

// Pass 1 over the tiles of glaciers [g0, g1) (g0 < 0: all): node planes D (and alpha, beta when partials).
template <typename T>
static int launch_law_nodes_t(odinn_ensemble* e, int g0, int g1, const void* H, bool partials) {
    const CellLaw lw = *law_of(e);
    int t0 = 0, nt = e->n_tiles;
    if (g0 >= 0) {
        t0 = e->gl[g0].tile0;
        nt = e->gl[g1 - 1].tile0 + e->gl[g1 - 1].ntx * e->gl[g1 - 1].nty - t0;
    }
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    int wmax = 0;
    for (int L = 0; L <= lw.arch.n_layers; ++L) wmax = std::max(wmax, lw.arch.widths[L]);
    const bool w16 = wmax <= 16;  // register-resident evaluator: compile-time width bound 16 or 32
#define LN(TT, RR, PP)                                                                                                         \
    do {                                                                                                                       \
        const size_t smem = sizeof(RR) * lw.arch.n_params;                                                                     \
        if (w16) law_nodes_kernel<TT, RR, PP, 16><<<nt, LAW_NT, smem, e->stream>>>(descs, e->d_tiles + t0, lw, e->d_law_theta, \
                     (const TT*)H, B, (TT*)e->lawD, PP ? (TT*)e->lawAl : nullptr, PP ? (TT*)e->lawBe : nullptr);               \
        else law_nodes_kernel<TT, RR, PP, 32><<<nt, LAW_NT, smem, e->stream>>>(descs, e->d_tiles + t0, lw, e->d_law_theta,     \
                     (const TT*)H, B, (TT*)e->lawD, PP ? (TT*)e->lawAl : nullptr, PP ? (TT*)e->lawBe : nullptr);               \
    } while (0)
    if (lf_match(lw) && (!partials || lw.kind == LAW_U)) {
        if (partials)
            law_nodes_fixed_kernel<T, 16, 16, true><<<nt, LAW_NT, 0, e->stream>>>(descs, e->d_tiles + t0, lw, e->d_law_theta, (const T*)H, B,
                                                                                 (T*)e->lawD, (T*)e->lawAl, (T*)e->lawBe);
        else
            law_nodes_fixed_kernel<T, 16, 16, false><<<nt, LAW_NT, 0, e->stream>>>(descs, e->d_tiles + t0, lw, e->d_law_theta, (const T*)H, B,
                                                                                  (T*)e->lawD, nullptr, nullptr);
        ODINN_CHECK_LAUNCH(e);
        return ODINN_OK;
    }
    if (!partials) LN(T, T, false);
    else if (lw.kind == LAW_U) LN(T, T, true);   // analytic partials ride along the forward evaluation: the ensemble's precision
    else LN(T, double, true);                    // LawY: one-sided difference of the network (target_D_hybrid.jl:58-73): fp64
#undef LN
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

int launch_law_nodes(odinn_ensemble* e, int g0, int g1, const void* H, bool partials) {
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_B)) || (rc = alloc_plane(e, &e->lawD))) return rc;
    if (partials && ((rc = alloc_plane(e, &e->lawAl)) || (rc = alloc_plane(e, &e->lawBe)))) return rc;
    if ((rc = sync_descs(e))) return rc;
    return e->dtype == ODINN_F32 ? launch_law_nodes_t<float>(e, g0, g1, H, partials) : launch_law_nodes_t<double>(e, g0, g1, H, partials);
}

// Pass 3 for glaciers [g0, g1): d_law_dtheta[g] = (accumulate ? old : 0) + scale * sum_nodes D_adj s dNN/dtheta.
template <typename T>
static int launch_law_theta_t(odinn_ensemble* e, int g0, int g1, const void* H, double scale, int accumulate) {
    const CellLaw lw = *law_of(e);
    const int np = lw.arch.n_params;
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    int NA = 0, NZ = 0, wmax = 0;
    for (int L = 0; L < lw.arch.n_layers; ++L) { NA += lw.arch.widths[L]; NZ += lw.arch.widths[L + 1]; }
    for (int L = 0; L <= lw.arch.n_layers; ++L) wmax = std::max(wmax, lw.arch.widths[L]);
    const size_t smem = (sizeof(double) + sizeof(int2)) * (size_t)np + sizeof(T) * (((size_t)np + 3) / 4 * 4 + (size_t)(NA + NZ) * LAW_PITCH);
    if (smem > 220 * 1024) return fail(e, ODINN_EARG, "per-cell law too large for the shared-memory pullback (reduce depth x width)");
    const bool w16 = wmax <= 16;
    // Opt in to > 48 KB dynamic shared memory.  The attribute is per DEVICE and handles on different GPUs may live in one process,
    // so it is set before every launch sequence (a host-side call, negligible beside these kernels) instead of being cached per process.
    if (smem > 48 * 1024) {
        if (w16) ODINN_CUDA(e, cudaFuncSetAttribute(law_theta_kernel<T, T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else ODINN_CUDA(e, cudaFuncSetAttribute(law_theta_kernel<T, T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int n0 = e->ext_int[2], n1 = e->ext_int[3];
    if (lf_match(lw)) {
        constexpr size_t fsmem = lf_theta_smem<T, 16, 16>();
        if (fsmem > 48 * 1024) {
            ODINN_CUDA(e, cudaFuncSetAttribute(law_theta_fixed_kernel<T, 16, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            ODINN_CUDA(e, cudaFuncSetAttribute(law_theta_fixed_kernel<T, 16, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        }
        if (n0 == 0) {
            // ONE launch over the tiles of glaciers [g0, g1), ONE reduction launch (block partials indexed by global tile)
            const int t0 = e->gl[g0].tile0, t1 = e->gl[g1 - 1].tile0 + e->gl[g1 - 1].ntx * e->gl[g1 - 1].nty;
            law_theta_fixed_kernel<T, 16, 16, false><<<t1 - t0, LAW_NT, fsmem, e->stream>>>(
                descs, e->d_tiles + t0, lw, e->d_law_theta, (const T*)H, B, (const T*)e->plane[ODINN_FIELD_VJP_A],
                e->d_law_partial + (size_t)t0 * np, nullptr, 0, 0, nullptr, 0);
            ODINN_CHECK_LAUNCH(e);
            law_theta_reduce_all<<<dim3(div_up(np, 128), g1 - g0), 128, 0, e->stream>>>(e->d_tile_start + g0, e->d_law_partial, np,
                                                                                      e->d_law_dtheta + (size_t)g0 * np, scale, accumulate);
            ODINN_CHECK_LAUNCH(e);
            return ODINN_OK;
        }
        const double* knots = (const double*)e->ext_dev[EXT_LAT_KNOTS];
        double* Wlat = (double*)e->ext_dev[EXT_LAT_W];
        const int nk = n0 * std::max(n1, 1), nb = div_up(nk, TX * TY);
        for (int g = g0; g < g1; ++g) {
            const int t0 = e->gl[g].tile0, nt = e->gl[g].ntx * e->gl[g].nty;
            ODINN_CUDA(e, cudaMemsetAsync(Wlat, 0, sizeof(double) * nk, e->stream));
            law_lattice_scatter<T><<<nt, LAW_NT, sizeof(double) * (n0 + n1), e->stream>>>(descs, e->d_tiles + t0, lw, (const T*)H, B,
                                                                                       (const T*)e->plane[ODINN_FIELD_VJP_A], knots, n0, n1, Wlat);
            ODINN_CHECK_LAUNCH(e);
            law_theta_fixed_kernel<T, 16, 16, true><<<nb, LAW_NT, fsmem, e->stream>>>(descs, nullptr, lw, e->d_law_theta, nullptr, nullptr, nullptr,
                                                                                     e->d_law_partial, knots, n0, n1, Wlat, g);
            ODINN_CHECK_LAUNCH(e);
            law_theta_reduce_scaled<<<div_up(np, 128), 128, 0, e->stream>>>(e->d_law_partial, nb, np, e->d_law_dtheta + (size_t)g * np, scale,
                                                                           accumulate);
            ODINN_CHECK_LAUNCH(e);
        }
        return ODINN_OK;
    }
    if (n0 > 0) {
        // interpolation = :Linear: scatter D†·s onto the knots (one lattice per glacier pass), then back-propagate the KNOTS
        const double* knots = (const double*)e->ext_dev[EXT_LAT_KNOTS];
        double* Wlat = (double*)e->ext_dev[EXT_LAT_W];
        const int nk = n0 * std::max(n1, 1);
        const int nb = div_up(nk, TX * TY);
        if (nb > e->max_tiles_per_glacier) return fail(e, ODINN_EARG, "interpolation lattice larger than the largest glacier: reduce n_interp_half");
        if (w16) { if (smem > 48 * 1024) ODINN_CUDA(e, cudaFuncSetAttribute(law_theta_kernel<T, T, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
        else if (smem > 48 * 1024) ODINN_CUDA(e, cudaFuncSetAttribute(law_theta_kernel<T, T, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int g = g0; g < g1; ++g) {
            const int t0 = e->gl[g].tile0, nt = e->gl[g].ntx * e->gl[g].nty;
            ODINN_CUDA(e, cudaMemsetAsync(Wlat, 0, sizeof(double) * nk, e->stream));
            law_lattice_scatter<T><<<nt, LAW_NT, sizeof(double) * (n0 + n1), e->stream>>>(descs, e->d_tiles + t0, lw, (const T*)H, B,
                                                                                       (const T*)e->plane[ODINN_FIELD_VJP_A], knots, n0, n1, Wlat);
            ODINN_CHECK_LAUNCH(e);
            if (w16)
                law_theta_kernel<T, T, 16, true><<<nb, LAW_NT, smem, e->stream>>>(descs, nullptr, lw, e->d_law_theta, nullptr, nullptr, nullptr,
                                                                                e->d_law_partial, knots, n0, n1, Wlat, g);
            else
                law_theta_kernel<T, T, 32, true><<<nb, LAW_NT, smem, e->stream>>>(descs, nullptr, lw, e->d_law_theta, nullptr, nullptr, nullptr,
                                                                                e->d_law_partial, knots, n0, n1, Wlat, g);
            ODINN_CHECK_LAUNCH(e);
            law_theta_reduce_scaled<<<div_up(np, 128), 128, 0, e->stream>>>(e->d_law_partial, nb, np, e->d_law_dtheta + (size_t)g * np, scale,
                                                                           accumulate);
            ODINN_CHECK_LAUNCH(e);
        }
        return ODINN_OK;
    }
    for (int g = g0; g < g1; ++g) {  // one glacier at a time: the block partials are [tiles of one glacier x n_theta]
        const int t0 = e->gl[g].tile0, nt = e->gl[g].ntx * e->gl[g].nty;
        if (w16)
            law_theta_kernel<T, T, 16><<<nt, LAW_NT, smem, e->stream>>>(descs, e->d_tiles + t0, lw, e->d_law_theta, (const T*)H, B,
                                                                    (const T*)e->plane[ODINN_FIELD_VJP_A], e->d_law_partial);
        else
            law_theta_kernel<T, T, 32><<<nt, LAW_NT, smem, e->stream>>>(descs, e->d_tiles + t0, lw, e->d_law_theta, (const T*)H, B,
                                                                    (const T*)e->plane[ODINN_FIELD_VJP_A], e->d_law_partial);
        ODINN_CHECK_LAUNCH(e);
        law_theta_reduce_scaled<<<div_up(np, 128), 128, 0, e->stream>>>(e->d_law_partial, nt, np, e->d_law_dtheta + (size_t)g * np,
                                                                       scale, accumulate);
        ODINN_CHECK_LAUNCH(e);
    }
    return ODINN_OK;
}

int launch_law_theta(odinn_ensemble* e, int g0, int g1, const void* H, double scale, int accumulate) {
    if (g0 < 0) { g0 = 0; g1 = e->G; }
    return e->dtype == ODINN_F32 ? launch_law_theta_t<float>(e, g0, g1, H, scale, accumulate)
                                 : launch_law_theta_t<double>(e, g0, g1, H, scale, accumulate);
}

}  // namespace odinn
