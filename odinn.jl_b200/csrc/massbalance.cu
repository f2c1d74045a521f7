// Mass-balance callback and its discrete VJP.
//
// Reference (ODINN.jl v1.1.0):
//   forward   mb_action! of the PeriodicCallback (src/simulations/inversions/inversion_utils.jl:498-517): at the end of every
//             step_MB window  S = B + H;  MB_timestep!;  apply_MB_mask!(H)  and the MB field is kept (MB_history)
//   reverse   VJP_λ_∂MB∂H(::DiscreteVJP, λ, H_preMB, ...) (src/inverse/SIA2D/VJPs.jl:107-151), applied to λ_j at the MB tstops
//             with H_preMB = H_j - MB (src/inverse/SIA2D/gradient.jl:201-207)
// The mass-balance model is TImodel1 (temperature index).  Its evaluation lives in Muninn / Sleipnir (NOT IN TREE); what the
// in-tree VJP fixes is reproduced here:  PDD = T + Γ (S - z_ref),  ∂MB/∂H = -DDF Γ [PDD >= 0] / (step_MB · 12) on
// MB_mask = (H > 0 ∧ MB < 0) ∨ (H > 10 ∧ MB >= 0),  MB clipped to -H where the ice would disappear (cotangent -λ there).
// ASSUMPTION (documented in DESIGN.md): MB = (acc_factor · snow - DDF · max(PDD, 0)) / (step_MB · 12) with glacier-wide
// snow for the window.  The climate scalars of every window do not depend on H, so the caller precomputes them per glacier
// (get_cumulative_climate!, VJPs.jl:112) and hands them over with odinn_set_mass_balance.
#include <vector>

#include "ensemble.cuh"

namespace odinn {

constexpr int MB_NPAR = 7;  // temp, gradient, ref_hgt, snow, DDF, acc_factor, scale = 1 / (step_MB * 12)

struct MbPar { double temp, grad, ref_hgt, snow, DDF, acc, scale; };

__device__ __forceinline__ double mb_value(const MbPar& c, double h, double b, bool& mask, bool& disappear, double& pdd) {
    pdd = c.temp + c.grad * ((b + h) - c.ref_hgt);
    double MB = (c.acc * c.snow - c.DDF * fmax(pdd, 0.0)) * c.scale;
    mask = (h > 0.0 && MB < 0.0) || (h > 10.0 && MB >= 0.0);
    disappear = false;
    if (!mask) return 0.0;
    if (h + MB < 0.0) { disappear = true; return -h; }
    return MB;
}

#define MB_COMMA ,
#define MB_TILE_LOOP(BODY)                                                        \
    const int2 tl = tiles[blockIdx.x];                                            \
    const GDesc<T> d = descs[tl.x];                                               \
    const MbPar c = par[tl.x];                                                    \
    const int x0 = (tl.y & 0xffff) * TX, y0 = (tl.y >> 16) * TY;                  \
    const int i = x0 + (threadIdx.x & 31), tr = threadIdx.x >> 5;                 \
    _Pragma("unroll") for (int rr = 0; rr < TY / 8; ++rr) {                       \
        const int j = y0 + tr + rr * 8;                                           \
        if (i < d.nx && j < d.ny) {                                               \
            const long long p = d.off + (long long)j * d.ld + i;                  \
            BODY                                                                  \
        }                                                                         \
    }

// H <- H + MB(H),  MBout <- MB
template <typename T>
__global__ void __launch_bounds__(NT)
mb_apply_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const MbPar* __restrict__ par,
                const T* __restrict__ B, T* __restrict__ H, T* __restrict__ MBout) {
    MB_TILE_LOOP({
        bool mask;
        bool gone;
        double pdd;
        const double h = (double)H[p];
        const double MB = mb_value(c MB_COMMA h MB_COMMA (double)B[p] MB_COMMA mask MB_COMMA gone MB_COMMA pdd);
        H[p] = (T)(h + MB);
        MBout[p] = (T)MB;
    })
}

// λ <- λ + VJP_λ_∂MB∂H(λ, H_preMB),  H_preMB = Hj - MB
template <typename T>
__global__ void __launch_bounds__(NT)
mb_adjoint_kernel(const GDesc<T>* __restrict__ descs, const int2* __restrict__ tiles, const MbPar* __restrict__ par,
                  const T* __restrict__ B, const T* __restrict__ Hj, const T* __restrict__ MBp, T* __restrict__ lam) {
    MB_TILE_LOOP({
        bool mask;
        bool gone;
        double pdd;
        const double h = (double)Hj[p] - (double)MBp[p];
        (void)mb_value(c MB_COMMA h MB_COMMA (double)B[p] MB_COMMA mask MB_COMMA gone MB_COMMA pdd);
        const double l = (double)lam[p];
        double out = 0.0;
        if (mask) out = -(c.DDF * (pdd < 0.0 ? 0.0 : c.grad * l)) * c.scale;
        if (gone) out = -l;
        lam[p] = (T)(l + out);
    })
}

static int mb_slot(const odinn_ensemble* e, int j) {
    for (size_t m = 0; m < e->mb_snap.size(); ++m)
        if (e->mb_snap[m] == j) return (int)m;
    return -1;
}
static char* mb_plane(odinn_ensemble* e, int m) { return (char*)e->ext_dev[EXT_MB] + (size_t)m * (size_t)e->total * e->esize; }
static const MbPar* mb_par(odinn_ensemble* e, int m) { return (const MbPar*)e->ext_dev[EXT_MB_PAR] + (size_t)m * e->G; }

// forward: the MB callback of snapshot j (no-op when none fires there).  Returns 1 in *applied when H was modified.
int mb_apply_step(odinn_ensemble* e, int j, void* H, int* applied) {
    if (applied) *applied = 0;
    const int m = mb_slot(e, j);
    if (m < 0) return ODINN_OK;
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_B)) || (rc = sync_descs(e))) return rc;
    if (e->dtype == ODINN_F32)
        mb_apply_kernel<float><<<e->n_tiles, NT, 0, e->stream>>>((const GDesc<float>*)e->d_descs, e->d_tiles, mb_par(e, m),
                                                                (const float*)e->plane[ODINN_FIELD_B], (float*)H, (float*)mb_plane(e, m));
    else
        mb_apply_kernel<double><<<e->n_tiles, NT, 0, e->stream>>>((const GDesc<double>*)e->d_descs, e->d_tiles, mb_par(e, m),
                                                                 (const double*)e->plane[ODINN_FIELD_B], (double*)H, (double*)mb_plane(e, m));
    ODINN_CHECK_LAUNCH(e);
    if (applied) *applied = 1;
    return ODINN_OK;
}

// reverse: λ_j += VJP_λ_∂MB∂H(λ_j, H_j - MB)   (gradient.jl:201-207)
int mb_adjoint_step(odinn_ensemble* e, int j, void* lam, const void* Hj) {
    const int m = mb_slot(e, j);
    if (m < 0) return ODINN_OK;
    int rc;
    if ((rc = sync_descs(e))) return rc;
    if (e->dtype == ODINN_F32)
        mb_adjoint_kernel<float><<<e->n_tiles, NT, 0, e->stream>>>((const GDesc<float>*)e->d_descs, e->d_tiles, mb_par(e, m),
                                                                  (const float*)e->plane[ODINN_FIELD_B], (const float*)Hj,
                                                                  (const float*)mb_plane(e, m), (float*)lam);
    else
        mb_adjoint_kernel<double><<<e->n_tiles, NT, 0, e->stream>>>((const GDesc<double>*)e->d_descs, e->d_tiles, mb_par(e, m),
                                                                   (const double*)e->plane[ODINN_FIELD_B], (const double*)Hj,
                                                                   (const double*)mb_plane(e, m), (double*)lam);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

}  // namespace odinn

using namespace odinn;

extern "C" {

int odinn_set_mass_balance(odinn_ensemble* e, int n_mb, const int* snapshot_index, const double* params) {
    if (!e) return fail(nullptr, ODINN_EARG, "null ensemble");
    {
        cudaError_t s_ = cudaSetDevice(e->device);
        if (s_ != cudaSuccess) return fail(e, ODINN_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(s_));
    }
    if (n_mb < 0 || (n_mb > 0 && (!snapshot_index || !params))) return fail(e, ODINN_EARG, "bad mass-balance arguments");
    for (int k : {(int)EXT_MB, (int)EXT_MB_PAR}) {
        if (e->ext_dev[k]) cudaFree(e->ext_dev[k]);
        e->ext_dev[k] = nullptr;
    }
    e->mb_snap.assign(snapshot_index, snapshot_index + n_mb);
    if (n_mb == 0) return ODINN_OK;
    int rc;
    if ((rc = alloc_work_plane(e, &e->ext_dev[EXT_MB], (size_t)n_mb))) return rc;
    static_assert(sizeof(MbPar) == sizeof(double) * MB_NPAR, "MbPar layout");
    const size_t bytes = sizeof(double) * MB_NPAR * (size_t)n_mb * e->G;
    ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_MB_PAR], bytes));
    ODINN_CUDA(e, cudaMemcpyAsync(e->ext_dev[EXT_MB_PAR], params, bytes, cudaMemcpyHostToDevice, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

/* MB field applied at MB step m (MB_history of the forward, inversion_utils.jl:509-510). */
int odinn_get_mass_balance(odinn_ensemble* e, int glacier, int m, void* host, int ld) {
    if (!e) return fail(nullptr, ODINN_EARG, "null ensemble");
    if (m < 0 || m >= (int)e->mb_snap.size() || !e->ext_dev[EXT_MB]) return fail(e, ODINN_ESTATE, "no such mass-balance step");
    if (glacier < 0 || glacier >= e->G || !host) return fail(e, ODINN_EARG, "bad arguments");
    int rc = copy_plane_2d(e, glacier, mb_plane(e, m), host, ld, false);
    if (rc) return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

}  // extern "C"
