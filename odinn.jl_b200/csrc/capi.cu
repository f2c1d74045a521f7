// C ABI of libodinn_b200.so (see include/odinn_b200.h for the reference interfaces replaced).
#include <cstdlib>
#include <type_traits>

#include "launch.cuh"
#include "sia2d_march2.cuh"  // (geometry constants of the work-item tables; no kernel is instantiated here)
#include "timeloop.cuh"
#include "sia2d_law.cuh"

namespace odinn {

thread_local std::string g_create_error;

int alloc_plane(odinn_ensemble* e, void** p, size_t n_planes) {
    if (*p) return ODINN_OK;
    size_t bytes = (size_t)e->total * e->esize * n_planes;
    cudaError_t st = cudaMalloc(p, bytes);
    if (st != cudaSuccess) {
        *p = nullptr;
        return fail(e, st == cudaErrorMemoryAllocation ? ODINN_ENOMEM : ODINN_ECUDA,
                    std::string("cudaMalloc of a device plane: ") + cudaGetErrorString(st));
    }
    ODINN_CUDA(e, cudaMemsetAsync(*p, 0, bytes, e->stream));
    return ODINN_OK;
}

int ensure_plane(odinn_ensemble* e, int field) {
    if (field < 0 || field >= ODINN_FIELD_COUNT_) return fail(e, ODINN_EARG, "bad field id");
    return alloc_plane(e, &e->plane[field]);
}

template <typename T>
static int sync_descs_t(odinn_ensemble* e) {
    // two tables: [0, G) the padded device layout, [G, 2G) the packed layout of the host-batch path
    // (ld = nx, glaciers back to back: exactly the bytes of the caller's matrices, so every copy is linear)
    std::vector<GDesc<T>> h(2 * e->G);
    for (int g = 0; g < 2 * e->G; ++g) {
        const bool packed = g >= e->G;
        const GlacierHost& s = e->gl[g % e->G];
        GDesc<T>& d = h[g];
        d.off = packed ? s.off_packed : s.off;
        d.nx = s.nx;
        d.ny = s.ny;
        d.ld = packed ? s.nx : s.ld;
        d.tile0 = s.tile0;
        d.dx = (T)s.dx;
        d.dy = (T)s.dy;
        d.inv_dx = (T)(1.0 / s.dx);
        d.inv_dy = (T)(1.0 / s.dy);
        d.A = (T)s.A;
        d.temp = (T)s.temp;
    }
    ODINN_CUDA(e, cudaMemcpyAsync(e->d_descs, h.data(), sizeof(GDesc<T>) * 2 * e->G, cudaMemcpyHostToDevice, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));  // h goes out of scope
    return ODINN_OK;
}

int sync_descs(odinn_ensemble* e) {
    if (!e->descs_dirty) return ODINN_OK;
    int rc = e->dtype == ODINN_F32 ? sync_descs_t<float>(e) : sync_descs_t<double>(e);
    if (rc == ODINN_OK) e->descs_dirty = false;
    return rc;
}

static void refresh_phys(odinn_ensemble* e) { e->cubic = (e->phys.n == 3.0 && e->phys.C == 0.0); }

static inline char* plane_ptr(odinn_ensemble* e, void* base, long long plane_index = 0) {
    return (char*)base + (size_t)plane_index * (size_t)e->total * e->esize;
}

// ---- per-cell laws (sia2d_law.cuh) --------------------------------------------------------------------------------

static inline CellLaw* law_of(odinn_ensemble* e) { return static_cast<CellLaw*>(e->law_cfg); }

static void law_refresh_phys(odinn_ensemble* e) {
    if (!e->law_cfg) return;
    CellLaw* lw = law_of(e);
    lw->Gam = 2.0 * std::pow(e->phys.rho * e->phys.g, e->phys.n) / (e->phys.n + 2.0);
    lw->Sl = e->phys.C * std::pow(e->phys.rho * e->phys.g, e->phys.p - e->phys.q);
    lw->p = e->phys.p;
    lw->q = e->phys.q;
}

// ---- F1 launch: out = SIA2D(Hin)   or, with a stage,  out = sa·U0 + sb·(Hin + sdt·SIA2D(Hin)) ----------------

__global__ void set_int_kernel(int* p, int v) { *p = v; }
__global__ void advance_interval_kernel(int* p) { *p += 1; }

// Glaciers [g0, g1); g0 < 0: whole ensemble.  `packed` selects the packed descriptor table (host-batch path).
static int launch_rhs_range(odinn_ensemble* e, int g0, int g1, const void* Hin, void* out, const Stage* st, bool packed) {
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_B))) return rc;
    if (e->a_gridded && (rc = ensure_plane(e, ODINN_FIELD_A))) return rc;
    if ((rc = sync_descs(e))) return rc;
    int i0 = 0, ni = e->n_items;
    if (g0 >= 0) {
        i0 = e->gl[g0].item0;
        ni = e->gl[g1 - 1].item0 + e->gl[g1 - 1].n_items - i0;
    }
    if (e->law_kind != LAW_NONE) {  // per-cell law: node pass, then the stencil in D-field mode (generic one-column kernels)
        if (packed) return fail(e, ODINN_ESTATE, "per-cell laws use the padded plane layout");
        if ((rc = launch_law_nodes(e, g0, g1, Hin, false))) return rc;
        return e->dtype == ODINN_F32 ? launch_rhs_law_t<float>(e, i0, ni, Hin, out, st) : launch_rhs_law_t<double>(e, i0, ni, Hin, out, st);
    }
    if (e->dtype == ODINN_F32 && e->march >= 2) return launch_rhs2(e, g0, g1, Hin, out, st, packed);
    return e->dtype == ODINN_F32 ? launch_rhs_t<float>(e, i0, ni, Hin, out, st, packed)
                                 : launch_rhs_t<double>(e, i0, ni, Hin, out, st, packed);
}
static int launch_rhs(odinn_ensemble* e, int g, const void* Hin, void* out, const Stage* st = nullptr) {
    return launch_rhs_range(e, g, g + 1, Hin, out, st, false);
}

// Glaciers [g0, g1); g0 < 0: whole ensemble.  S_dst: where the per-glacier sums go (d_S or an accumulator), scaled by `scale`.
// dH_out != nullptr: also dH_out <- SIA2D(H) -- fused into the A1+A2 pass where a fused kernel exists (fp32 two-column
// kernels, wH && wS), otherwise as a separate F1 launch.
static int launch_rhs_range(odinn_ensemble* e, int g0, int g1, const void* Hin, void* out, const Stage* st, bool packed);
static int launch_vjp_range(odinn_ensemble* e, int g0, int g1, const void* lam, const void* H, void* out, bool wH, bool wS,
                            double* S_dst, double scale, int accumulate, bool packed, void* dH_out = nullptr) {
    int rc;
    const bool fuse = dH_out && wH && wS && e->law_kind == LAW_NONE && !e->no_fuse &&
                      ((e->dtype == ODINN_F32 && (e->march == 2 || e->march == 4)) || (e->dtype == ODINN_F64 && e->cubic));
    if (dH_out && !fuse && (rc = launch_rhs_range(e, g0, g1, H, dH_out, nullptr, packed))) return rc;
    if (!fuse) dH_out = nullptr;
    if (!wH && !wS) return ODINN_OK;
    if ((rc = ensure_plane(e, ODINN_FIELD_B))) return rc;
    if (e->a_gridded && ((rc = ensure_plane(e, ODINN_FIELD_A)) || (wS && (rc = ensure_plane(e, ODINN_FIELD_VJP_A)))))
        return rc;
    if ((rc = sync_descs(e))) return rc;
    int i0 = 0, ni = e->n_items;
    if (g0 >= 0) {
        i0 = e->gl[g0].item0;
        ni = e->gl[g1 - 1].item0 + e->gl[g1 - 1].n_items - i0;
    }
    if (e->law_kind != LAW_NONE) {
        // node pass with partials -> stencil in D-field mode (D-adjoint plane when the theta-VJP is wanted) -> law pullback
        if (packed) return fail(e, ODINN_ESTATE, "per-cell laws use the padded plane layout");
        if (wS && (rc = ensure_plane(e, ODINN_FIELD_VJP_A))) return rc;
        if ((rc = launch_law_nodes(e, g0, g1, H, true))) return rc;
        rc = e->dtype == ODINN_F32 ? launch_vjp_law_t<float>(e, i0, ni, lam, H, out, wH, wS)
                                   : launch_vjp_law_t<double>(e, i0, ni, lam, H, out, wH, wS);
        if (rc) return rc;
        if (wS) return launch_law_theta(e, g0, g1, H, scale, accumulate);  // (S_dst is unused: the result is a vector per glacier)
        return ODINN_OK;
    }
    const bool two = (e->dtype == ODINN_F32 && e->march >= 2);
    const int* starts2 = e->d_item2_start;
    if (two) rc = launch_vjp2(e, g0, g1, lam, H, out, wH, wS, packed, dH_out, &starts2);
    else
        rc = e->dtype == ODINN_F32 ? launch_vjp_t<float>(e, i0, ni, lam, H, out, wH, wS, packed)
                                   : launch_vjp_t<double>(e, i0, ni, lam, H, out, wH, wS, packed, dH_out);
    if (rc) return rc;
    if (wS) {
        double* dst = S_dst ? S_dst : e->d_S;
        const int* starts = two ? starts2 : e->d_item_start;
        const double* part = (two && e->tma_partial_live) ? e->tma_partial_live : e->d_partial;
        if (g0 >= 0)
            reduce_scaled_kernel<<<g1 - g0, NT, 0, e->stream>>>(starts + g0, part, dst + g0, scale, accumulate);
        else
            reduce_scaled_kernel<<<e->G, NT, 0, e->stream>>>(starts, part, dst, scale, accumulate);
        ODINN_CHECK_LAUNCH(e);
    }
    return ODINN_OK;
}
static int launch_vjp(odinn_ensemble* e, int g, const void* lam, const void* H, void* out, bool wH, bool wS,
                      double* S_dst = nullptr, double scale = 1.0, int accumulate = 0) {
    return launch_vjp_range(e, g, g + 1, lam, H, out, wH, wS, S_dst, scale, accumulate, false);
}

// ---- continuous VJPs (sia2d_cont.cuh) ---------------------------------------------------------------------------

// out = (dSIA/dH)^T lam, continuous form (adjoint.jl:442-555).  g < 0: whole ensemble.
static int launch_vjpc(odinn_ensemble* e, int g, const void* lam, const void* H, void* out) {
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_B))) return rc;
    if (e->a_gridded && (rc = ensure_plane(e, ODINN_FIELD_A))) return rc;
    if ((rc = sync_descs(e))) return rc;
    int i0 = 0, ni = e->n_items;
    if (g >= 0) { i0 = e->gl[g].item0; ni = e->gl[g].n_items; }
    return e->dtype == ODINN_F32 ? launch_vjpc_t<float>(e, i0, ni, lam, H, out) : launch_vjpc_t<double>(e, i0, ni, lam, H, out);
}

// S[g] = Σ λ ⊙ pad(∇·(avg(∂A_spatial)·clamp(∇S)))  (adjoint.jl:582-662, glacier-wide law).  g < 0: whole ensemble.
static int launch_unitA_dot(odinn_ensemble* e, int g, const void* lam, const void* H, double* S_dst, double scale = 1.0,
                            int accumulate = 0) {
    // The Tullio chain of adjoint.jl:646-657 is linear in the tensor dD/dtheta and uses the CLAMPED edge slopes: transposed, it is the
    // discrete contraction sum_nodes dD/dtheta[node] D_adj[node] of adjoint.jl:250 (identical to rounding: tests/test_oracle_identities.py).
    // For a glacier-wide law that sum factorises and the unit-A forward + dot product below is the cheapest way to get it; for a gridded A
    // (one parameter per node) and for per-cell laws (a network gradient per node) the discrete A2 kernels already produce exactly it.
    if (e->a_gridded || e->law_kind != LAW_NONE)
        return launch_vjp_range(e, g, g + 1, lam, H, nullptr, false, true, S_dst, scale, accumulate, false);
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_B)) || (rc = alloc_plane(e, &e->work[0]))) return rc;
    if ((rc = sync_descs(e))) return rc;
    return e->dtype == ODINN_F32 ? launch_unitA_dot_t<float>(e, g, lam, H, S_dst, scale, accumulate)
                                 : launch_unitA_dot_t<double>(e, g, lam, H, S_dst, scale, accumulate);
}

static int copy2d_ptr(odinn_ensemble* e, int g, char* plane_base, bool dual, void* host, int ld, bool up,
                      cudaStream_t st) {
    if (g < 0 || g >= e->G) return fail(e, ODINN_EARG, "glacier index out of range");
    if (!host) return fail(e, ODINN_EARG, "null host pointer");
    const GlacierHost& s = e->gl[g];
    int w = dual ? s.nx - 1 : s.nx, h = dual ? s.ny - 1 : s.ny;
    if (ld < w) return fail(e, ODINN_EARG, "ld smaller than the number of rows");
    char* dev = plane_base + (size_t)s.off * e->esize;
    if (up)
        ODINN_CUDA(e, cudaMemcpy2DAsync(dev, (size_t)s.ld * e->esize, host, (size_t)ld * e->esize, (size_t)w * e->esize,
                                        h, cudaMemcpyHostToDevice, st));
    else
        ODINN_CUDA(e, cudaMemcpy2DAsync(host, (size_t)ld * e->esize, dev, (size_t)s.ld * e->esize, (size_t)w * e->esize,
                                        h, cudaMemcpyDeviceToHost, st));
    return ODINN_OK;
}

static int copy2d(odinn_ensemble* e, int g, int field, void* host, int ld, bool up, cudaStream_t st) {
    int rc = ensure_plane(e, field);
    if (rc) return rc;
    bool dual = (field == ODINN_FIELD_A || field == ODINN_FIELD_VJP_A);
    return copy2d_ptr(e, g, (char*)e->plane[field], dual, host, ld, up, st);
}

// ---- loss / seed ---------------------------------------------------------------------------------------------

// loss_dst[g] (+)= wloss · Σ W (H - Href)² ; optionally λ_out = λ_in + dt·v + cseed·W·(H - Href)
static int launch_loss_seed(odinn_ensemble* e, const void* H, const void* Href, const void* W, const void* lam_in,
                            const void* v, void* lam_out, double dt, double cseed, double* loss_dst, double wloss,
                            int accumulate) {
    int rc = sync_descs(e);
    if (rc) return rc;
    const int N = e->dtype == ODINN_F32 ? 4 : 2;
    long long max_vec = 0;
    for (int g = 0; g < e->G; ++g) max_vec = std::max(max_vec, (long long)e->gl[g].ld * e->gl[g].ny / N);
    const int nchunk = (int)((max_vec + (long long)NT * LS_UNROLL - 1) / ((long long)NT * LS_UNROLL));
    if (e->ext_int[1] < nchunk * e->G) {  // per-(glacier, chunk) partial sums of the loss
        if (e->ext_dev[EXT_LS_PARTIAL]) cudaFree(e->ext_dev[EXT_LS_PARTIAL]);
        e->ext_dev[EXT_LS_PARTIAL] = nullptr;
        ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_LS_PARTIAL], sizeof(double) * (size_t)nchunk * e->G));
        e->ext_int[1] = nchunk * e->G;
    }
    double* partial = (double*)e->ext_dev[EXT_LS_PARTIAL];
    const dim3 grid(nchunk, e->G);
    if (e->dtype == ODINN_F32)
        loss_seed_vec_kernel<float><<<grid, NT, 0, e->stream>>>((const GDesc<float>*)e->d_descs, (const float*)H, (const float*)Href,
                                                               (const float*)W, (const float*)lam_in, (const float*)v, (float*)lam_out,
                                                               partial, (float)dt, (float)cseed);
    else
        loss_seed_vec_kernel<double><<<grid, NT, 0, e->stream>>>((const GDesc<double>*)e->d_descs, (const double*)H, (const double*)Href,
                                                                (const double*)W, (const double*)lam_in, (const double*)v, (double*)lam_out,
                                                                partial, dt, cseed);
    ODINN_CHECK_LAUNCH(e);
    reduce_chunks_scaled_kernel<<<e->G, NT, 0, e->stream>>>(partial, nchunk, loss_dst, wloss, accumulate);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

int alloc_work_plane(odinn_ensemble* e, void** p, size_t n_planes) { return alloc_plane(e, p, n_planes); }
int copy_plane_2d(odinn_ensemble* e, int glacier, void* plane, void* host, int ld, bool to_device) {
    return copy2d_ptr(e, glacier, (char*)plane, false, host, ld, to_device, e->stream);
}
int vjp_planes(odinn_ensemble* e, const void* lam, const void* H, void* out, bool wH, bool wS, double* S_dst, double scale,
               int accumulate, bool continuous) {
    if (!continuous) return launch_vjp_range(e, -1, 0, lam, H, out, wH, wS, S_dst, scale, accumulate, false);
    int rc;
    if (wH && (rc = launch_vjpc(e, -1, lam, H, out))) return rc;
    if (wS && (rc = launch_unitA_dot(e, -1, lam, H, S_dst ? S_dst : e->d_S, scale, accumulate))) return rc;
    return ODINN_OK;
}
int loss_seed_planes(odinn_ensemble* e, const void* H, const void* Href, const void* W, const void* lam_in, const void* v,
                     void* lam_out, double dt, double cseed, double* loss_dst, double wloss, int accumulate) {
    return launch_loss_seed(e, H, Href, W, lam_in, v, lam_out, dt, cseed, loss_dst, wloss, accumulate);
}
int rhs_planes(odinn_ensemble* e, const void* Hin, void* out) { return launch_rhs(e, -1, Hin, out); }
bool pdl_enabled() {
    static const bool on = []() { const char* v = getenv("ODINN_PDL"); return v && v[0] == '1'; }();
    return on;
}
bool rhs_rk_fusable(const odinn_ensemble* e) { return e->law_kind == LAW_NONE; }
bool vjp_rk_fusable(const odinn_ensemble* e) {
    return e->law_kind == LAW_NONE && !e->a_gridded && ((e->dtype == ODINN_F32 && e->march >= 2) || (e->dtype == ODINN_F64 && e->cubic));
}
int vjp_planes_rk(odinn_ensemble* e, const void* S1in, const void* Ha, const void* Hb, void* S1out, const void* rkfuse, double c, double sign,
                  double ta, double tb, bool norm) {
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_B)) || (rc = sync_descs(e))) return rc;
    const bool two = (e->dtype == ODINN_F32);
    rc = two ? launch_vjp2_rk(e, S1in, Ha, Hb, S1out, rkfuse, c, sign, ta, tb) : launch_vjp_rk_t<double>(e, S1in, Ha, Hb, S1out, rkfuse, c, sign, ta, tb);
    if (rc || !norm) return rc;
    reduce_scaled_kernel<<<e->G, NT, 0, e->stream>>>(two ? e->d_item2_start : e->d_item_start, e->d_partial, e->d_S, 1.0, 0);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}
// S_dst[g] (+)= scale * sum gA D+(lam, H_itp) with H_itp = lerp(Ha, Hb) formed on load (the quadrature-node term of the continuous
// adjoint, gradient.jl:495-507); rkstate: the controller table the kernel takes each glacier's time from (every glacier sits on the node).
int vjp_planes_lerp_S(odinn_ensemble* e, const void* lam, const void* Ha, const void* Hb, const void* rkstate, double sign, double ta, double tb,
                      double* S_dst, double scale, int accumulate) {
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_B)) || (rc = sync_descs(e))) return rc;
    const bool two = (e->dtype == ODINN_F32);
    if (two) {
        RkFuse<float> f{};
        f.st = (const RkState*)rkstate;
        f.flags = RKF_FIRST | RKF_LERP_ONLY;
        rc = launch_vjp2_rk(e, lam, Ha, Hb, nullptr, &f, 0.0, sign, ta, tb, true);
    } else {
        RkFuse<double> f{};
        f.st = (const RkState*)rkstate;
        f.flags = RKF_FIRST | RKF_LERP_ONLY;
        rc = launch_vjp_rk_t<double>(e, lam, Ha, Hb, nullptr, &f, 0.0, sign, ta, tb, true);
    }
    if (rc) return rc;
    reduce_scaled_kernel<<<e->G, NT, 0, e->stream>>>(two ? e->d_item2_start : e->d_item_start, e->d_partial, S_dst, scale, accumulate);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}
int rhs_planes_rk(odinn_ensemble* e, const void* S1in, void* S1out, const void* rkfuse, bool norm) {
    Stage st{};
    st.rk = rkfuse;
    int rc = launch_rhs_range(e, -1, 0, S1in, S1out, &st, false);
    if (rc || !norm) return rc;
    const bool two = (e->dtype == ODINN_F32 && e->march >= 2);
    reduce_scaled_kernel<<<e->G, NT, 0, e->stream>>>(two ? e->d_item2_start : e->d_item_start, e->d_partial, e->d_S, 1.0, 0);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}
int reduce_tiles(odinn_ensemble* e, const double* tile_partial, double* dst, double scale, int accumulate) {
    reduce_scaled_kernel<<<e->G, NT, 0, e->stream>>>(e->d_tile_start, tile_partial, dst, scale, accumulate);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}
int prepare_snapshots(odinn_ensemble* e, int n_snap) {
    if (e->n_snap != n_snap) {
        if (e->snap) cudaFree(e->snap);
        e->snap = nullptr;
        e->n_snap = 0;
    }
    int rc = alloc_plane(e, &e->snap, n_snap);
    if (rc) return rc;
    e->n_snap = n_snap;
    return ODINN_OK;
}
void* snapshot_ptr(odinn_ensemble* e, int j) { return plane_ptr(e, e->snap, j); }

}  // namespace odinn

using namespace odinn;

#define GUARD(e)                                                   \
    if (!(e)) return fail(nullptr, ODINN_EARG, "null ensemble");   \
    {                                                              \
        cudaError_t _s = cudaSetDevice((e)->device);               \
        if (_s != cudaSuccess) return fail((e), ODINN_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_s)); \
    }

extern "C" {

int odinn_ensemble_create(int device, int dtype, int n_glaciers, const int* nx, const int* ny, const double* dx,
                          const double* dy, const odinn_phys* phys, odinn_ensemble** out) {
    if (!out) return fail(nullptr, ODINN_EARG, "out is null");
    *out = nullptr;
    if (n_glaciers <= 0 || !nx || !ny || !dx || !dy || !phys) return fail(nullptr, ODINN_EARG, "bad create arguments");
    if (dtype != ODINN_F32 && dtype != ODINN_F64) return fail(nullptr, ODINN_EARG, "dtype must be ODINN_F32 or ODINN_F64");
    int ndev = 0;
    cudaError_t st = cudaGetDeviceCount(&ndev);
    if (st != cudaSuccess || ndev == 0)
        return fail(nullptr, ODINN_ECUDA,
                    std::string("no CUDA device available (libodinn_b200 has no CPU fallback): ") + cudaGetErrorString(st));
    if (device < 0 || device >= ndev) return fail(nullptr, ODINN_EARG, "device index out of range");
    if ((st = cudaSetDevice(device)) != cudaSuccess)
        return fail(nullptr, ODINN_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(st));

    odinn_ensemble* e = new (std::nothrow) odinn_ensemble();
    if (!e) return fail(nullptr, ODINN_ENOMEM, "out of host memory");
    e->device = device;
    e->dtype = dtype;
    e->esize = dtype == ODINN_F32 ? 4 : 8;
    e->G = n_glaciers;
    e->phys = *phys;
    refresh_phys(e);
    e->gl.resize(n_glaciers);
    long long off = 0, offp = 0;
    int tile = 0;
    for (int g = 0; g < n_glaciers; ++g) {
        if (nx[g] < 3 || ny[g] < 3 || nx[g] > 65535 * TX || !(dx[g] > 0) || !(dy[g] > 0)) {
            delete e;
            return fail(nullptr, ODINN_EARG, "glacier grid must be at least 3x3 with positive spacing");
        }
        GlacierHost& s = e->gl[g];
        s.nx = nx[g];
        s.ny = ny[g];
        s.ld = div_up(nx[g], 32) * 32;
        s.off = off;
        s.off_packed = offp;
        offp += (long long)nx[g] * ny[g];
        s.dx = dx[g];
        s.dy = dy[g];
        s.A = 0.0;
        s.temp = 0.0;
        s.ntx = div_up(s.nx, TX);
        s.nty = div_up(s.ny, TY);
        s.tile0 = tile;
        tile += s.ntx * s.nty;
        e->max_tiles_per_glacier = std::max(e->max_tiles_per_glacier, s.ntx * s.nty);
        off += (long long)s.ld * s.ny;
        e->cells += (long long)s.nx * s.ny;
    }
    e->total = off;
    e->n_tiles = tile;
    if (dtype == ODINN_F32 && off + (long long)(ODINN_L2PF_ROWS + 8) * 65536 >= (1LL << 31)) {  // the fp32 A1+A2 kernels index planes with 32-bit element offsets
        delete e;
        return fail(nullptr, ODINN_EARG, "ensemble too large for one handle: a plane must stay below 2^31 elements (split the ensemble)");
    }

    std::vector<int2> tiles(tile);
    std::vector<int> tstart(n_glaciers + 1);
    for (int g = 0; g < n_glaciers; ++g) {
        const GlacierHost& s = e->gl[g];
        tstart[g] = s.tile0;
        for (int ty = 0; ty < s.nty; ++ty)
            for (int tx = 0; tx < s.ntx; ++tx) tiles[s.tile0 + ty * s.ntx + tx] = make_int2(g, (ty << 16) | tx);
    }
    tstart[n_glaciers] = tile;

    // Marching work items: strips of STRIP output columns x chunks of rows.  Shorter chunks for small
    // ensembles so that the grid still fills 148 SMs.
    std::vector<int4> items;
    std::vector<int> istart(n_glaciers + 1);
    // (Every chunk pays a warm-up step and re-reads its halo rows.  fp64 sweep at 500 x 500 x 256, profiles/r02_chunk_rows_sweep.txt: fused
    //  F1+A1+A2 step 0.627 ms with 32-row chunks, 0.589 with 64, 0.576 with 100, 0.591 with 167; F1 0.305 -> 0.268 ms.)
    for (int rows : {100, 64, 32, 16, 8}) {
        long long n = 0;
        for (int g = 0; g < n_glaciers; ++g) n += (long long)div_up(e->gl[g].nx, STRIP) * div_up(e->gl[g].ny, rows);
        e->chunk_rows = rows;
        if (n >= 148LL * 48) break;
    }
    if (const char* envr1 = getenv("ODINN_CHUNK_ROWS1")) {   // (tuning sweeps)
        if (atoi(envr1) >= 4) e->chunk_rows = atoi(envr1);
    }
    for (int g = 0; g < n_glaciers; ++g) {
        GlacierHost& s = e->gl[g];
        s.item0 = (int)items.size();
        istart[g] = s.item0;
        for (int r0 = 0; r0 < s.ny; r0 += e->chunk_rows)
            for (int st_ = 0; st_ < div_up(s.nx, STRIP); ++st_)
                items.push_back(make_int4(g, st_ * STRIP - 1, r0, std::min(r0 + e->chunk_rows, s.ny)));
        s.n_items = (int)items.size() - s.item0;
    }
    istart[n_glaciers] = (int)items.size();
    e->n_items = (int)items.size();

    // two-column strips (fp32): STRIP2 output columns per warp, first loaded column = 60k - 2 (even)
    std::vector<int4> items2;
    std::vector<int> istart2(n_glaciers + 1);
    {
        const char* env = getenv("ODINN_MARCH");
        if (env && (env[0] == '1' || env[0] == '2' || env[0] == '4')) e->march = env[0] - '0';
        const char* envf = getenv("ODINN_NO_FUSE");  // developer switch: F1 and A1+A2 as two launches even where a fused kernel exists
        e->no_fuse = envf && envf[0] == '1';
        const char* envr = getenv("ODINN_CHUNK_ROWS2");
        int forced = envr ? atoi(envr) : 0;
        for (int rows = 64; rows >= 8; rows /= 2) {
            long long n = 0;
            for (int g = 0; g < n_glaciers; ++g) n += (long long)div_up(e->gl[g].nx, STRIP2) * div_up(e->gl[g].ny, rows);
            e->chunk_rows2 = rows;
            if (n >= 148LL * 12) break;  // ~1 wave of warps; shorter chunks only add halo rows (sweep on BASELINE configs 3 / 4: 8 -> 16 rows, 24.7 -> 21.9 ms)
        }
        if (forced >= 4) e->chunk_rows2 = forced;
        for (int g = 0; g < n_glaciers; ++g) {
            GlacierHost& s = e->gl[g];
            if (s.nx & 1) e->all_nx_even = false;
            s.item20 = (int)items2.size();
            istart2[g] = s.item20;
            for (int r0 = 0; r0 < s.ny; r0 += e->chunk_rows2)
                for (int st_ = 0; st_ < div_up(s.nx, STRIP2); ++st_)
                    items2.push_back(make_int4(g, st_ * STRIP2 - 2, r0, std::min(r0 + e->chunk_rows2, s.ny)));
            s.n_items2 = (int)items2.size() - s.item20;
        }
        istart2[n_glaciers] = (int)items2.size();
        e->n_items2 = (int)items2.size();
    }
    // Big ensembles: a second table with row chunks of ~125 rows for the WHOLE-ensemble launches of F1, the fused SSPRK3 / Euler stage and
    // the A1 / A2 / reverse-step kernels (their per-item partial sums are then indexed by the long table's own start array).  Every chunk pays a warm-up step and re-reads its halo rows: at 500 x 500 x 256 the
    // SSPRK3 stage goes 0.197 -> 0.184 ms with 125-row chunks (profiles/r02_chunk_rows_sweep.txt).  Only where that still leaves two
    // waves of warps; the adaptive engines keep the short chunks (their partially active launches need the parallelism).
    std::vector<int4> items2L;
    std::vector<int> istart2L(n_glaciers + 1, 0);
    if (dtype == ODINN_F32 && !getenv("ODINN_CHUNK_ROWS2")) {
        for (int g = 0; g < n_glaciers; ++g) {
            const GlacierHost& s = e->gl[g];
            istart2L[g] = (int)items2L.size();
            const int nch = std::max(1, (s.ny + 62) / 125), rows = div_up(s.ny, nch);
            for (int r0 = 0; r0 < s.ny; r0 += rows)
                for (int st_ = 0; st_ < div_up(s.nx, STRIP2); ++st_) items2L.push_back(make_int4(g, st_ * STRIP2 - 2, r0, std::min(r0 + rows, s.ny)));
        }
        istart2L[n_glaciers] = (int)items2L.size();
        if ((long long)items2L.size() < 2LL * 148 * 16 || items2L.size() >= items2.size()) items2L.clear();
    }

#define CREATE_CUDA(call)                                                                     \
    do {                                                                                      \
        cudaError_t _st = (call);                                                             \
        if (_st != cudaSuccess) {                                                             \
            std::string m = std::string(#call) + ": " + cudaGetErrorString(_st);              \
            odinn_ensemble_destroy(e);                                                        \
            return fail(nullptr, ODINN_ECUDA, m);                                             \
        }                                                                                     \
    } while (0)
    CREATE_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaStreamCreateWithFlags(&e->copy_stream[0], cudaStreamNonBlocking));
    CREATE_CUDA(cudaStreamCreateWithFlags(&e->copy_stream[1], cudaStreamNonBlocking));
    size_t dsz = dtype == ODINN_F32 ? sizeof(GDesc<float>) : sizeof(GDesc<double>);
    CREATE_CUDA(cudaMalloc(&e->d_descs, dsz * n_glaciers * 2));
    CREATE_CUDA(cudaMalloc(&e->d_tiles, sizeof(int2) * tile));
    CREATE_CUDA(cudaMalloc(&e->d_tile_start, sizeof(int) * (n_glaciers + 1)));
    CREATE_CUDA(cudaMalloc(&e->d_partial, sizeof(double) * std::max(tile, std::max(e->n_items, e->n_items2))));
    CREATE_CUDA(cudaMalloc(&e->d_items2, sizeof(int4) * e->n_items2));
    CREATE_CUDA(cudaMalloc(&e->d_item2_start, sizeof(int) * (n_glaciers + 1)));
    CREATE_CUDA(cudaMemcpy(e->d_items2, items2.data(), sizeof(int4) * e->n_items2, cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMemcpy(e->d_item2_start, istart2.data(), sizeof(int) * (n_glaciers + 1), cudaMemcpyHostToDevice));
    if (!items2L.empty()) {
        CREATE_CUDA(cudaMalloc(&e->ext_dev[EXT_ITEMS2_LONG], sizeof(int4) * items2L.size()));
        CREATE_CUDA(cudaMemcpy(e->ext_dev[EXT_ITEMS2_LONG], items2L.data(), sizeof(int4) * items2L.size(), cudaMemcpyHostToDevice));
        e->ext_int[5] = (int)items2L.size();
        CREATE_CUDA(cudaMalloc(&e->ext_dev[EXT_ITEMS2_LONG_START], sizeof(int) * (n_glaciers + 1)));
        CREATE_CUDA(cudaMemcpy(e->ext_dev[EXT_ITEMS2_LONG_START], istart2L.data(), sizeof(int) * (n_glaciers + 1), cudaMemcpyHostToDevice));
    }
    CREATE_CUDA(cudaMalloc(&e->d_items, sizeof(int4) * e->n_items));
    CREATE_CUDA(cudaMalloc(&e->d_item_start, sizeof(int) * (n_glaciers + 1)));
    CREATE_CUDA(cudaMalloc(&e->d_S, sizeof(double) * n_glaciers * 4));  // S, Ssum, loss, A
    CREATE_CUDA(cudaMemset(e->d_S, 0, sizeof(double) * n_glaciers * 4));
    CREATE_CUDA(cudaMallocHost(&e->h_S, sizeof(double) * n_glaciers * 4));
    CREATE_CUDA(cudaMemcpy(e->d_tiles, tiles.data(), sizeof(int2) * tile, cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMemcpy(e->d_tile_start, tstart.data(), sizeof(int) * (n_glaciers + 1), cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMemcpy(e->d_items, items.data(), sizeof(int4) * e->n_items, cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMemcpy(e->d_item_start, istart.data(), sizeof(int) * (n_glaciers + 1), cudaMemcpyHostToDevice));
#undef CREATE_CUDA
    e->d_Ssum = e->d_S + n_glaciers;
    e->d_loss = e->d_S + 2 * n_glaciers;
    e->d_A = e->d_S + 3 * n_glaciers;
    *out = e;
    return ODINN_OK;
}

void odinn_ensemble_destroy(odinn_ensemble* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (int f = 0; f < ODINN_FIELD_COUNT_; ++f)
        if (e->plane[f]) cudaFree(e->plane[f]);
    void* ptrs[] = {e->d_descs, e->d_tiles, e->d_tile_start, e->d_partial, e->d_items, e->d_item_start, e->d_S,
                    e->snap, e->href, e->wmask, e->work[0], e->work[1], e->d_theta, e->d_J, e->d_dtheta, e->d_temps,
                    e->d_items2, e->d_item2_start, e->d_law_theta, e->lawD, e->lawAl, e->lawBe, e->d_law_partial,
                    e->d_law_dtheta, e->bpack, e->stage[0], e->stage[1], e->stage[2], e->stage[3], e->ad_plane[0],
                    e->ad_plane[1], e->ad_plane[2], e->ad_plane[3], e->ad_plane[4], e->ad_plane[5], e->d_ad_state, e->d_ad_dims};
    tma_cache_free(e);
    for (void* p : e->ext_dev)
        if (p) cudaFree(p);
    for (void* p : e->ext_host)
        if (p) cudaFreeHost(p);
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete law_of(e);
    if (e->h_S) cudaFreeHost(e->h_S);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->h_ad_active) cudaFreeHost(e->h_ad_active);
    if (e->fwd_graph_exec) cudaGraphExecDestroy(e->fwd_graph_exec);
    if (e->d_fwd_tab) cudaFree(e->d_fwd_tab);
    for (cudaEvent_t ev : e->ev_up) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_done) cudaEventDestroy(ev);
    for (int k = 0; k < 2; ++k)
        if (e->copy_stream[k]) cudaStreamDestroy(e->copy_stream[k]);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

const char* odinn_last_error(const odinn_ensemble* e) { return e ? e->err.c_str() : g_create_error.c_str(); }
int odinn_n_glaciers(const odinn_ensemble* e) { return e ? e->G : 0; }
int odinn_dtype_of(const odinn_ensemble* e) { return e ? e->dtype : -1; }
long long odinn_launch_count(const odinn_ensemble* e) { return e ? e->launches : 0; }
void* odinn_stream(odinn_ensemble* e) { return e ? (void*)e->stream : nullptr; }

int odinn_synchronize(odinn_ensemble* e) {
    GUARD(e);
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_upload(odinn_ensemble* e, int glacier, int field, const void* host, int ld) {
    GUARD(e);
    int rc = copy2d(e, glacier, field, const_cast<void*>(host), ld, true, e->stream);
    if (rc) return rc;
    if (field == ODINN_FIELD_B) e->bpack_dirty = true;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_download(odinn_ensemble* e, int glacier, int field, void* host, int ld) {
    GUARD(e);
    int rc = copy2d(e, glacier, field, host, ld, false, e->stream);
    if (rc) return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_set_A_scalar(odinn_ensemble* e, int glacier, double A) {
    GUARD(e);
    if (glacier < 0 || glacier >= e->G) return fail(e, ODINN_EARG, "glacier index out of range");
    e->gl[glacier].A = A;
    e->descs_dirty = true;
    return ODINN_OK;
}

int odinn_set_temperature(odinn_ensemble* e, int glacier, double T) {
    GUARD(e);
    if (glacier < 0 || glacier >= e->G) return fail(e, ODINN_EARG, "glacier index out of range");
    e->gl[glacier].temp = T;
    e->descs_dirty = true;
    return ODINN_OK;
}

int odinn_set_A_mode(odinn_ensemble* e, int gridded) {
    GUARD(e);
    e->a_gridded = gridded ? 1 : 0;
    return ODINN_OK;
}

int odinn_set_phys(odinn_ensemble* e, const odinn_phys* phys) {
    GUARD(e);
    if (!phys) return fail(e, ODINN_EARG, "phys is null");
    e->phys = *phys;
    refresh_phys(e);
    law_refresh_phys(e);
    return ODINN_OK;
}

// ---- per-call operators (host buffers) -----------------------------------------------------------------------

int odinn_sia2d_rhs(odinn_ensemble* e, int glacier, const void* H, int ldH, void* dH, int lddH, double t) {
    GUARD(e);
    (void)t;  // autonomous RHS: laws with callback_freq = 0 do not depend on t (Laws.jl:346)
    int rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_H, const_cast<void*>(H), ldH, true, e->stream))) return rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_DH))) return rc;
    if ((rc = launch_rhs(e, glacier, e->plane[ODINN_FIELD_H], e->plane[ODINN_FIELD_DH]))) return rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_DH, dH, lddH, false, e->stream))) return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_sia2d_vjp_H(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH, void* out,
                      int ldo, double t) {
    GUARD(e);
    (void)t;
    int rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_H, const_cast<void*>(H), ldH, true, e->stream))) return rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_LAMBDA, const_cast<void*>(lambda), ldl, true, e->stream))) return rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_VJP_H))) return rc;
    if ((rc = launch_vjp(e, glacier, e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_H], e->plane[ODINN_FIELD_VJP_H],
                         true, false)))
        return rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_VJP_H, out, ldo, false, e->stream))) return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_sia2d_vjp_theta(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                          double* out_S, double t) {
    GUARD(e);
    (void)t;
    if (!out_S) return fail(e, ODINN_EARG, "out_S is null");
    int rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_H, const_cast<void*>(H), ldH, true, e->stream))) return rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_LAMBDA, const_cast<void*>(lambda), ldl, true, e->stream))) return rc;
    if ((rc = launch_vjp(e, glacier, e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_H], nullptr, false, true)))
        return rc;
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S + glacier, e->d_S + glacier, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    *out_S = e->h_S[glacier];
    return ODINN_OK;
}

int odinn_sia2d_vjp_H_continuous(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                                 void* out, int ldo, double t) {
    GUARD(e);
    (void)t;
    int rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_H, const_cast<void*>(H), ldH, true, e->stream))) return rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_LAMBDA, const_cast<void*>(lambda), ldl, true, e->stream))) return rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_VJP_H))) return rc;
    if ((rc = launch_vjpc(e, glacier, e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_H], e->plane[ODINN_FIELD_VJP_H])))
        return rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_VJP_H, out, ldo, false, e->stream))) return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_sia2d_vjp_theta_continuous(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                                     double* out_S, double t) {
    GUARD(e);
    (void)t;
    if (!out_S) return fail(e, ODINN_EARG, "out_S is null");
    int rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_H, const_cast<void*>(H), ldH, true, e->stream))) return rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_LAMBDA, const_cast<void*>(lambda), ldl, true, e->stream))) return rc;
    if ((rc = launch_unitA_dot(e, glacier, e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_H], e->d_S))) return rc;
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S + glacier, e->d_S + glacier, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    *out_S = e->h_S[glacier];
    return ODINN_OK;
}

// ---- ensemble operators on resident planes ---------------------------------------------------------------------

int odinn_rhs_resident(odinn_ensemble* e) {
    GUARD(e);
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_H)) || (rc = ensure_plane(e, ODINN_FIELD_DH))) return rc;
    return launch_rhs(e, -1, e->plane[ODINN_FIELD_H], e->plane[ODINN_FIELD_DH]);
}

int odinn_vjp_resident(odinn_ensemble* e, int flags, double* S_out) {
    GUARD(e);
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_H)) || (rc = ensure_plane(e, ODINN_FIELD_LAMBDA))) return rc;
    if ((flags & 1) && (rc = ensure_plane(e, ODINN_FIELD_VJP_H))) return rc;
    if (flags & 4) {  // continuous flavour
        if ((flags & 1) && (rc = launch_vjpc(e, -1, e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_H], e->plane[ODINN_FIELD_VJP_H])))
            return rc;
        if ((flags & 2) && (rc = launch_unitA_dot(e, -1, e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_H], e->d_S)))
            return rc;
    } else {
        void* dH_out = nullptr;
        if (flags & 8) {  // also FIELD_DH <- SIA2D(FIELD_H), in the same pass where a fused kernel exists
            if ((rc = ensure_plane(e, ODINN_FIELD_DH))) return rc;
            dH_out = e->plane[ODINN_FIELD_DH];
        }
        rc = launch_vjp_range(e, -1, 0, e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_H], e->plane[ODINN_FIELD_VJP_H],
                              (flags & 1) != 0, (flags & 2) != 0, nullptr, 1.0, 0, false, dH_out);
        if (rc) return rc;
    }
    if ((flags & 2) && S_out) {
        ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_S, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
        ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
        memcpy(S_out, e->h_S, sizeof(double) * e->G);
    }
    return ODINN_OK;
}

int odinn_fwd_adj_batch_host(odinn_ensemble* e, const void* const* H, const void* const* lambda, void* const* dH,
                             void* const* vjpH, double* S) {
    GUARD(e);
    if (!H) return fail(e, ODINN_EARG, "H is null");
    const bool adj = (vjpH != nullptr) || (S != nullptr);
    if (adj && !lambda) return fail(e, ODINN_EARG, "lambda is required for the VJP outputs");
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_B))) return rc;
    // The call works on its own staging planes, so the resident planes (FIELD_H, ...) keep their contents.
    if ((rc = alloc_plane(e, &e->stage[0]))) return rc;
    if (adj && (rc = alloc_plane(e, &e->stage[1]))) return rc;
    if (dH && (rc = alloc_plane(e, &e->stage[2]))) return rc;
    if (vjpH && (rc = alloc_plane(e, &e->stage[3]))) return rc;
    if ((rc = sync_descs(e))) return rc;
    // Packed layout (ld = nx): the device planes hold exactly the caller's bytes, so every transfer is one linear
    // DMA instead of ny row copies.  The gridded-A field lives in the padded layout only -> padded (2-D copy) path.
    if (S && e->law_kind != LAW_NONE)
        return fail(e, ODINN_ESTATE, "with a per-cell law the theta-VJP is a vector per glacier: use odinn_law_cell_grad");
    const bool packed = !e->a_gridded && e->all_nx_even && e->law_kind == LAW_NONE;  // (the fp32 kernels need 8-byte aligned column pairs)
    if (packed && e->bpack_dirty) {
        if ((rc = alloc_plane(e, &e->bpack))) return rc;
        for (int g = 0; g < e->G; ++g) {
            const GlacierHost& s = e->gl[g];
            ODINN_CUDA(e, cudaMemcpy2DAsync((char*)e->bpack + (size_t)s.off_packed * e->esize, (size_t)s.nx * e->esize,
                                            (char*)e->plane[ODINN_FIELD_B] + (size_t)s.off * e->esize,
                                            (size_t)s.ld * e->esize, (size_t)s.nx * e->esize, s.ny,
                                            cudaMemcpyDeviceToDevice, e->stream));
        }
        e->bpack_dirty = false;
    }
    // chunks of ~2 M cells: H2D of chunk c+1, kernels of chunk c and D2H of chunk c-1 overlap
    // (A ramped schedule -- smaller first and last chunks to shorten the pipeline's fill and drain -- was measured slower, 5.17 vs
    // 5.37 G cell-steps/s: copies below ~32 MiB lose more PCIe efficiency than the shorter ends gain; profiles/r01_v6_sweep.txt.)
    const long long chunk_cells = e->batch_chunk_cells;
    std::vector<int> cstart{0};
    {
        long long acc = 0;
        for (int g = 0; g < e->G; ++g) {
            acc += (long long)e->gl[g].nx * e->gl[g].ny;
            if (acc >= chunk_cells && g + 1 < e->G) { cstart.push_back(g + 1); acc = 0; }
        }
        cstart.push_back(e->G);
    }
    const int nchunk = (int)cstart.size() - 1;
    while ((int)e->ev_up.size() < nchunk + 1) {
        cudaEvent_t a, b;
        ODINN_CUDA(e, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        ODINN_CUDA(e, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        e->ev_up.push_back(a);
        e->ev_done.push_back(b);
    }
    cudaStream_t up = e->copy_stream[0], dn = e->copy_stream[1];
    // the copy streams start after everything already queued on the compute stream (bpack, earlier calls)
    ODINN_CUDA(e, cudaEventRecord(e->ev_done[nchunk], e->stream));
    ODINN_CUDA(e, cudaStreamWaitEvent(up, e->ev_done[nchunk], 0));
    ODINN_CUDA(e, cudaStreamWaitEvent(dn, e->ev_done[nchunk], 0));
    auto xfer = [&](int g, int slot, void* host, bool to_dev, cudaStream_t st) -> int {
        if (!host) return fail(e, ODINN_EARG, "null host pointer in a batch array");
        const GlacierHost& s = e->gl[g];
        if (!packed) return copy2d_ptr(e, g, (char*)e->stage[slot], false, host, s.nx, to_dev, st);
        char* dev = (char*)e->stage[slot] + (size_t)s.off_packed * e->esize;
        size_t bytes = (size_t)s.nx * s.ny * e->esize;
        ODINN_CUDA(e, to_dev ? cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st)
                             : cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
        return ODINN_OK;
    };
    // Packed layout: glaciers whose host matrices are adjacent in memory (a batch held in one array) move as ONE copy per
    // plane and chunk -- 1 MiB copies reach 28 GB/s each way when both directions run, 64 MiB copies 49 GB/s
    // (tools/microbench/pcie.py on the B200 box).
    auto xfer_range = [&](int g0, int g1, int slot, void* const* host, bool to_dev, cudaStream_t st) -> int {
        if (!packed) {
            for (int g = g0; g < g1; ++g)
                if ((rc = xfer(g, slot, host[g], to_dev, st))) return rc;
            return ODINN_OK;
        }
        int r0 = g0;
        while (r0 < g1) {
            if (!host[r0]) return fail(e, ODINN_EARG, "null host pointer in a batch array");
            size_t bytes = (size_t)e->gl[r0].nx * e->gl[r0].ny * e->esize;
            int r1 = r0 + 1;
            while (r1 < g1 && host[r1] == (char*)host[r0] + bytes) {
                bytes += (size_t)e->gl[r1].nx * e->gl[r1].ny * e->esize;
                ++r1;
            }
            char* dev = (char*)e->stage[slot] + (size_t)e->gl[r0].off_packed * e->esize;
            ODINN_CUDA(e, to_dev ? cudaMemcpyAsync(dev, host[r0], bytes, cudaMemcpyHostToDevice, st)
                                 : cudaMemcpyAsync(host[r0], dev, bytes, cudaMemcpyDeviceToHost, st));
            r0 = r1;
        }
        return ODINN_OK;
    };
    for (int c = 0; c < nchunk; ++c) {
        const int g0 = cstart[c], g1 = cstart[c + 1];
        if ((rc = xfer_range(g0, g1, 0, const_cast<void* const*>(H), true, up))) return rc;
        if (adj && (rc = xfer_range(g0, g1, 1, const_cast<void* const*>(lambda), true, up))) return rc;
        ODINN_CUDA(e, cudaEventRecord(e->ev_up[c], up));
        ODINN_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_up[c], 0));
        if (!adj) {
            if (dH && (rc = launch_rhs_range(e, g0, g1, e->stage[0], e->stage[2], nullptr, packed))) return rc;
        } else if ((rc = launch_vjp_range(e, g0, g1, e->stage[1], e->stage[0], e->stage[3], vjpH != nullptr, S != nullptr,
                                          nullptr, 1.0, 0, packed, dH ? e->stage[2] : nullptr)))
            return rc;
        ODINN_CUDA(e, cudaEventRecord(e->ev_done[c], e->stream));
        ODINN_CUDA(e, cudaStreamWaitEvent(dn, e->ev_done[c], 0));
        if (dH && (rc = xfer_range(g0, g1, 2, dH, false, dn))) return rc;
        if (vjpH && (rc = xfer_range(g0, g1, 3, vjpH, false, dn))) return rc;
    }
    if (S) ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_S, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
    // join: the compute stream (the one callers time / synchronise) waits for both copy streams
    ODINN_CUDA(e, cudaEventRecord(e->ev_up[nchunk], dn));
    ODINN_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_up[nchunk], 0));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    if (S) memcpy(S, e->h_S, sizeof(double) * e->G);
    return ODINN_OK;
}

int odinn_set_batch_chunk(odinn_ensemble* e, long long cells) {
    GUARD(e);
    if (cells < 1) return fail(e, ODINN_EARG, "chunk size must be positive");
    e->batch_chunk_cells = cells;
    return ODINN_OK;
}

int odinn_host_register(odinn_ensemble* e, void* host, size_t bytes) {
    GUARD(e);
    if (!host || !bytes) return fail(e, ODINN_EARG, "null host range");
    ODINN_CUDA(e, cudaHostRegister(host, bytes, cudaHostRegisterPortable));
    return ODINN_OK;
}

int odinn_host_unregister(odinn_ensemble* e, void* host) {
    GUARD(e);
    ODINN_CUDA(e, cudaHostUnregister(host));
    return ODINN_OK;
}

// ---- on-device time loop ---------------------------------------------------------------------------------------

// One tstop interval: nsub sub-steps of the chosen scheme.  `tab` != nullptr: the launches are being captured into a CUDA graph
// and read their coefficients from the device table (h is then unused).
static int forward_interval(odinn_ensemble* e, int method, int nsub, double h, void*& Hs, void*& U1, void*& U2, const double* tab,
                            const int* interval) {
    int rc;
    for (int s = 0; s < nsub; ++s) {
        if (method == ODINN_EULER) {
            Stage s1{Hs, 0.0, 1.0, h, tab, interval};  // U1 = H + h f(H)
            if ((rc = launch_rhs(e, -1, Hs, U1, &s1))) return rc;
            std::swap(Hs, U1);
        } else {  // Shu-Osher SSPRK(3,3)
            Stage s1{Hs, 0.0, 1.0, h, tab, interval};                          // u1 = H + h f(H)
            Stage s2{Hs, 0.75, 0.25, h, tab ? tab + 3 : nullptr, interval};    // u2 = 3/4 H + 1/4 (u1 + h f(u1))
            Stage s3{Hs, 1.0 / 3.0, 2.0 / 3.0, h, tab ? tab + 6 : nullptr, interval};  // H = 1/3 H + 2/3 (u2 + h f(u2))  (in place over U0)
            if ((rc = launch_rhs(e, -1, Hs, U1, &s1))) return rc;
            if ((rc = launch_rhs(e, -1, U1, U2, &s2))) return rc;
            if ((rc = launch_rhs(e, -1, U2, Hs, &s3))) return rc;
        }
    }
    return ODINN_OK;
}

int odinn_solve_forward(odinn_ensemble* e, int method, int n_snap, const double* t, int nsub) {
    GUARD(e);
    if (n_snap < 1 || !t || nsub < 1) return fail(e, ODINN_EARG, "bad time grid");
    if (method != ODINN_EULER && method != ODINN_SSPRK3) return fail(e, ODINN_EARG, "unknown integration method");
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_H0)) || (rc = ensure_plane(e, ODINN_FIELD_H)) || (rc = ensure_plane(e, ODINN_FIELD_B))) return rc;
    if (e->a_gridded && (rc = ensure_plane(e, ODINN_FIELD_A))) return rc;
    if ((rc = alloc_plane(e, &e->work[0])) || (rc = alloc_plane(e, &e->work[1]))) return rc;
    if ((rc = prepare_snapshots(e, n_snap))) return rc;
    if ((rc = sync_descs(e))) return rc;
    const size_t pbytes = (size_t)e->total * e->esize;
    void* Hs = e->plane[ODINN_FIELD_H];  // current state; H and the work planes rotate by pointer
    void* U1 = e->work[0];
    void* U2 = e->work[1];

    // Small glaciers (every glacier of the ensemble fits the shared memory of one thread-block cluster): the state stays on the SMs
    // and ONE launch runs a whole range of intervals, writing the snapshots as it goes (sia2d_cluster.cuh).  A range ends where a
    // mass-balance callback fires (inversion_utils.jl:498-517): it is applied to the global plane between two launches.
    if (const int cs = cluster_plan(e, 0)) {
        const double* d_t = nullptr;
        if ((rc = upload_time_grid(e, t, n_snap, &d_t))) return rc;
        ODINN_CUDA(e, cudaMemcpyAsync(plane_ptr(e, e->snap, 0), e->plane[ODINN_FIELD_H0], pbytes, cudaMemcpyDeviceToDevice, e->stream));
        if (n_snap == 1) ODINN_CUDA(e, cudaMemcpyAsync(Hs, e->plane[ODINN_FIELD_H0], pbytes, cudaMemcpyDeviceToDevice, e->stream));
        const void* Hin = e->plane[ODINN_FIELD_H0];
        int j0 = 0;
        while (j0 < n_snap - 1) {
            int j1 = n_snap - 1;
            for (int m : e->mb_snap) if (m > j0 && m < j1) j1 = m;
            if ((rc = launch_interval_cluster(e, cs, method, nsub, j0, j1, Hin, Hs, e->snap, d_t))) return rc;
            int applied = 0;
            if ((rc = mb_apply_step(e, j1, Hs, &applied))) return rc;
            if (applied) ODINN_CUDA(e, cudaMemcpyAsync(plane_ptr(e, e->snap, j1), Hs, pbytes, cudaMemcpyDeviceToDevice, e->stream));
            Hin = Hs;
            j0 = j1;
        }
        return ODINN_OK;
    }
    ODINN_CUDA(e, cudaMemcpyAsync(Hs, e->plane[ODINN_FIELD_H0], pbytes, cudaMemcpyDeviceToDevice, e->stream));
    ODINN_CUDA(e, cudaMemcpyAsync(plane_ptr(e, e->snap, 0), Hs, pbytes, cudaMemcpyDeviceToDevice, e->stream));

    // The interval body is launch-bound for small and medium ensembles (a 128 x 128 glacier: 6 us per RHS launch against 2 us of
    // kernel): it is captured ONCE into a CUDA graph and replayed per interval; the step sizes come from a device table indexed by an
    // interval counter that the graph itself advances.  (Euler swaps H and the work plane every sub-step: even nsub only.)
    const char* nog = getenv("ODINN_NO_GRAPH");
    const bool use_graph = !(nog && nog[0] == '1') && e->law_kind == LAW_NONE && 
                           (method == ODINN_SSPRK3 || nsub % 2 == 0) && n_snap > 2;
    if (use_graph) {
        std::vector<double> tab((size_t)n_snap * 9, 0.0);
        for (int j = 1; j < n_snap; ++j) {
            const double h = (t[j] - t[j - 1]) / nsub;
            const double c[9] = {0.0, 1.0, h, 0.75, 0.25, h, 1.0 / 3.0, 2.0 / 3.0, h};
            std::copy(c, c + 9, tab.begin() + (size_t)j * 9);
        }
        if (e->fwd_tab_len < n_snap) {
            if (e->d_fwd_tab) cudaFree(e->d_fwd_tab);
            e->d_fwd_tab = nullptr;
            ODINN_CUDA(e, cudaMalloc(&e->d_fwd_tab, sizeof(double) * 9 * (size_t)n_snap + sizeof(int)));
            e->fwd_tab_len = n_snap;
            e->fwd_graph_key.clear();  // (the table pointer is baked into the graph)
        }
        int* d_interval = (int*)(e->d_fwd_tab + 9 * (size_t)e->fwd_tab_len);
        ODINN_CUDA(e, cudaMemcpyAsync(e->d_fwd_tab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, e->stream));
        ODINN_CUDA(e, cudaStreamSynchronize(e->stream));  // tab goes out of scope; also: nothing pending when the capture starts
        set_int_kernel<<<1, 1, 0, e->stream>>>(d_interval, 1);
        ODINN_CHECK_LAUNCH(e);
        // everything a captured launch bakes in: scheme, physics, layout switches, plane pointers
        std::string key;
        auto add = [&key](const void* p, size_t n) { key.append((const char*)p, n); };
        const void* ptrs[6] = {Hs, U1, U2, e->plane[ODINN_FIELD_B], e->plane[ODINN_FIELD_A], e->d_fwd_tab};
        add(&method, sizeof(method)); add(&nsub, sizeof(nsub)); add(&e->phys, sizeof(e->phys)); add(&e->a_gridded, sizeof(e->a_gridded));
        add(&e->march, sizeof(e->march)); add(ptrs, sizeof(ptrs));
        if (!e->fwd_graph_exec || key != e->fwd_graph_key) {
            if (e->fwd_graph_exec) { cudaGraphExecDestroy(e->fwd_graph_exec); e->fwd_graph_exec = nullptr; }
            const long long l0 = e->launches;
            cudaGraph_t graph = nullptr;
            ODINN_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            void *gH = Hs, *g1 = U1, *g2 = U2;
            rc = forward_interval(e, method, nsub, 0.0, gH, g1, g2, e->d_fwd_tab, d_interval);
            if (rc == ODINN_OK) advance_interval_kernel<<<1, 1, 0, e->stream>>>(d_interval);
            cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(e, ODINN_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&e->fwd_graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(e, ODINN_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
            e->fwd_graph_launches = (int)(e->launches - l0) + 1;
            e->launches = l0;
            e->fwd_graph_key = key;
        }
        for (int j = 1; j < n_snap; ++j) {
            ODINN_CUDA(e, cudaGraphLaunch(e->fwd_graph_exec, e->stream));
            e->launches += e->fwd_graph_launches;
            if ((rc = mb_apply_step(e, j, Hs, nullptr))) return rc;  // mass-balance callback at the end of its window (inversion_utils.jl:498-517)
            ODINN_CUDA(e, cudaMemcpyAsync(plane_ptr(e, e->snap, j), Hs, pbytes, cudaMemcpyDeviceToDevice, e->stream));
        }
        return ODINN_OK;
    }
    for (int j = 1; j < n_snap; ++j) {
        const double h = (t[j] - t[j - 1]) / nsub;
        if ((rc = forward_interval(e, method, nsub, h, Hs, U1, U2, nullptr, nullptr))) return rc;
        if ((rc = mb_apply_step(e, j, Hs, nullptr))) return rc;  // mass-balance callback at the end of its window (inversion_utils.jl:498-517)
        ODINN_CUDA(e, cudaMemcpyAsync(plane_ptr(e, e->snap, j), Hs, pbytes, cudaMemcpyDeviceToDevice, e->stream));
    }
    if (Hs != e->plane[ODINN_FIELD_H])  // leave the final state in FIELD_H
        ODINN_CUDA(e, cudaMemcpyAsync(e->plane[ODINN_FIELD_H], Hs, pbytes, cudaMemcpyDeviceToDevice, e->stream));
    return ODINN_OK;
}

int odinn_set_cluster_mode(odinn_ensemble* e, int mode) {
    GUARD(e);
    if (mode != -1 && mode != 0 && mode != 1 && mode != 2 && mode != 4 && mode != 8 && mode != 16)
        return fail(e, ODINN_EARG, "cluster mode must be -1 (automatic), 0 (off) or a cluster size 1, 2, 4, 8, 16");
    e->cluster_mode = mode;
    return ODINN_OK;
}

int odinn_get_snapshot(odinn_ensemble* e, int glacier, int j, void* host, int ld) {
    GUARD(e);
    if (!e->snap || j < 0 || j >= e->n_snap) return fail(e, ODINN_ESTATE, "no such snapshot (run odinn_solve_forward first)");
    int rc = copy2d_ptr(e, glacier, plane_ptr(e, e->snap, j), false, host, ld, false, e->stream);
    if (rc) return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_set_snapshot(odinn_ensemble* e, int glacier, int j, int n_snap, const void* host, int ld) {
    GUARD(e);
    if (n_snap < 1 || j < 0 || j >= n_snap) return fail(e, ODINN_EARG, "bad snapshot index");
    if (e->n_snap != n_snap) {
        if (e->snap) cudaFree(e->snap);
        e->snap = nullptr;
        e->n_snap = 0;
    }
    int rc;
    if ((rc = alloc_plane(e, &e->snap, n_snap))) return rc;
    e->n_snap = n_snap;
    if ((rc = copy2d_ptr(e, glacier, plane_ptr(e, e->snap, j), false, const_cast<void*>(host), ld, true, e->stream)))
        return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

int odinn_set_reference(odinn_ensemble* e, int glacier, int j, int n_snap, const void* Href, const void* W, int ld) {
    GUARD(e);
    if (n_snap < 1 || j < 0 || j >= n_snap) return fail(e, ODINN_EARG, "bad snapshot index");
    if (e->n_ref != n_snap) {
        if (e->href) cudaFree(e->href);
        if (e->wmask) cudaFree(e->wmask);
        e->href = e->wmask = nullptr;
        e->n_ref = 0;
    }
    int rc;
    if ((rc = alloc_plane(e, &e->href, n_snap)) || (rc = alloc_plane(e, &e->wmask, n_snap))) return rc;
    if (e->n_ref != n_snap) e->ref_has.assign((size_t)n_snap, 0);   // (planes are zero-filled: W = 0 where no data was given)
    e->n_ref = n_snap;
    e->ref_has[j] = 1;
    if ((rc = copy2d_ptr(e, glacier, plane_ptr(e, e->href, j), false, const_cast<void*>(Href), ld, true, e->stream)))
        return rc;
    if ((rc = copy2d_ptr(e, glacier, plane_ptr(e, e->wmask, j), false, const_cast<void*>(W), ld, true, e->stream)))
        return rc;
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    return ODINN_OK;
}

static int check_grad_state(odinn_ensemble* e, const double* t, int n_t) {
    if (!e->snap || !e->href) return fail(e, ODINN_ESTATE, "snapshots and reference data must be set first");
    if (e->n_snap != e->n_ref || n_t != e->n_snap)
        return fail(e, ODINN_ESTATE, "snapshot / reference / time counts differ");
    if (!t) return fail(e, ODINN_EARG, "t is null");
    return ODINN_OK;
}

int odinn_loss(odinn_ensemble* e, const double* t, int n_t, double* loss_out) {
    GUARD(e);
    int rc = check_grad_state(e, t, n_t);
    if (rc) return rc;
    if (!loss_out) return fail(e, ODINN_EARG, "loss_out is null");
    ODINN_CUDA(e, cudaMemsetAsync(e->d_loss, 0, sizeof(double) * e->G, e->stream));
    for (int j = 0; j < n_t; ++j) {  // Δt_H of the first data point is 0 (safe_slice, gradient.jl:146-149)
        const double wH = loss_weight_H(e, t, n_t, j), wV = loss_weight_V(e, n_t, j);
        if (wH != 0.0 && (rc = launch_loss_seed(e, plane_ptr(e, e->snap, j), plane_ptr(e, e->href, j), plane_ptr(e, e->wmask, j),
                                                nullptr, nullptr, nullptr, 0.0, 0.0, e->d_loss, wH, 1)))
            return rc;
        if ((rc = velocity_loss_term(e, j, plane_ptr(e, e->snap, j), nullptr, wV, e->d_loss, nullptr))) return rc;
    }
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_loss, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    memcpy(loss_out, e->h_S, sizeof(double) * e->G);
    return ODINN_OK;
}

int odinn_grad_discrete(odinn_ensemble* e, const double* t, int n_t, double* loss_out, double* Ssum_out) {
    GUARD(e);
    int rc = check_grad_state(e, t, n_t);
    if (rc) return rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_LAMBDA)) || (rc = ensure_plane(e, ODINN_FIELD_VJP_H))) return rc;
    if (e->a_gridded)
        return fail(e, ODINN_ESTATE, "odinn_grad_discrete supports glacier-wide A (use the per-call VJPs for gridded A)");
    if (e->law_kind != LAW_NONE)  // the theta-gradient of a per-cell law accumulates per glacier in d_law_dtheta
        ODINN_CUDA(e, cudaMemsetAsync(e->d_law_dtheta, 0, sizeof(double) * (size_t)e->G * e->law_n_theta, e->stream));
    const size_t pbytes = (size_t)e->total * e->esize;
    void* lam = e->plane[ODINN_FIELD_LAMBDA];
    void* vH = e->plane[ODINN_FIELD_VJP_H];
    ODINN_CUDA(e, cudaMemsetAsync(lam, 0, pbytes, e->stream));  // λ_k = 0 (gradient.jl:140)
    ODINN_CUDA(e, cudaMemsetAsync(e->d_loss, 0, sizeof(double) * e->G, e->stream));
    ODINN_CUDA(e, cudaMemsetAsync(e->d_Ssum, 0, sizeof(double) * e->G, e->stream));
    // Small ensembles, LossH only: the WHOLE reverse loop runs cluster-resident (sia2d_cluster.cuh: lambda in shared memory, H_j / H_ref /
    // W streamed from the snapshot planes, loss and S accumulated on the device) -- one launch per range of steps between two
    // mass-balance tstops instead of ~4 launches per saved step.
    bool only_H = true;
    for (int j = 0; j < n_t; ++j) only_H = only_H && loss_weight_V(e, n_t, j) == 0.0;
    if (const int cs = only_H ? cluster_plan(e, 2) : 0) {
        if ((rc = sync_descs(e))) return rc;
        std::vector<double> tw(2 * (size_t)n_t);
        for (int j = 0; j < n_t; ++j) { tw[j] = t[j]; tw[n_t + j] = loss_weight_H(e, t, n_t, j); }
        const double* d_tw = nullptr;
        if ((rc = upload_time_grid(e, tw.data(), 2 * n_t, &d_tw))) return rc;
        const void* lam_in = nullptr;   // lambda_k = 0 (gradient.jl:140)
        int jhi = n_t - 1;
        while (jhi >= 1) {
            int jlo = 0;
            for (int m : e->mb_snap) if (m < jhi && m > jlo) jlo = m;   // the next mass-balance tstop below jhi ends the range
            if ((rc = launch_reverse_cluster(e, cs, jhi, jlo, lam_in, lam, d_tw, d_tw + n_t))) return rc;
            // lambda_jlo += VJP_lambda_dMB/dH(lambda_jlo, H_jlo - MB) before step jlo is taken                      (gradient.jl:201-207)
            if (jlo >= 1 && (rc = mb_adjoint_step(e, jlo, lam, plane_ptr(e, e->snap, jlo)))) return rc;
            lam_in = lam;
            jhi = jlo;
        }
        const double wH0 = loss_weight_H(e, t, n_t, 0);
        void* H0s = plane_ptr(e, e->snap, 0);
        if (wH0 != 0.0 && (rc = launch_loss_seed(e, H0s, plane_ptr(e, e->href, 0), plane_ptr(e, e->wmask, 0), nullptr, nullptr, nullptr, 0.0,
                                                 0.0, e->d_loss, wH0, 1)))
            return rc;
        ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_loss, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
        ODINN_CUDA(e, cudaMemcpyAsync(e->h_S + e->G, e->d_Ssum, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
        ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
        if (loss_out) memcpy(loss_out, e->h_S, sizeof(double) * e->G);
        if (Ssum_out) memcpy(Ssum_out, e->h_S + e->G, sizeof(double) * e->G);
        return ODINN_OK;
    }
    // fp32 two-column kernels, glacier-wide A: the reverse time step is folded into the A1 pass (SEED variant of sia2d_vjp_march2:
    // lambda_{j-1} = lambda_j + dt VJP_H + dl_j/dH written to the other lambda plane, loss term reduced per strip): 9 words/cell per
    // saved step (A1+seed 6, A2 3) instead of 13 (A1 4, loss / seed 6, A2 3).  ODINN_NO_FUSE=1 keeps the three-pass form.
    const bool fused_seed = e->dtype == ODINN_F32 && e->march >= 2 && e->law_kind == LAW_NONE && !e->no_fuse;
    for (int j = n_t - 1; j >= 1; --j) {  // gradient.jl:191-253 (the j = 1 pass of the reference updates nothing)
        const double dt = t[j] - t[j - 1];
        void* Hj = plane_ptr(e, e->snap, j);
        const double wH = loss_weight_H(e, t, n_t, j), wV = loss_weight_V(e, n_t, j);
        // λ_j += VJP_λ_∂MB∂H(λ_j, H_j - MB) at the MB tstops                                (gradient.jl:201-207)
        if (j < n_t - 1 && (rc = mb_adjoint_step(e, j, lam, Hj))) return rc;  // (λ_k = 0 at the last snapshot: nothing to add)
        if (fused_seed && j < n_t - 1) {
            // λ_∂f∂H = VJP_H(λ_j, H_j);  ℓ += ℓ_j;  λ_{j-1} = λ_j + Δt_{j-1} λ_∂f∂H + ∂ℓ_j/∂H   in ONE pass        (:218-242)
            if ((rc = sync_descs(e))) return rc;
            const int* seed_starts = e->d_item2_start;
            if ((rc = launch_vjp2_seed(e, lam, Hj, plane_ptr(e, e->href, j), plane_ptr(e, e->wmask, j), vH, dt, 2.0 * wH, &seed_starts))) return rc;
            reduce_scaled_kernel<<<e->G, NT, 0, e->stream>>>(seed_starts, e->d_partial, e->d_loss, wH, 1);
            ODINN_CHECK_LAUNCH(e);
            std::swap(lam, vH);
        } else {
            // λ_∂f∂H = VJP_H(λ_j, H_j)                                                     (gradient.jl:235-237)
            if (j < n_t - 1) {
                if ((rc = launch_vjp(e, -1, lam, Hj, vH, true, false))) return rc;
            } else {
                ODINN_CUDA(e, cudaMemsetAsync(vH, 0, pbytes, e->stream));  // λ_k = 0 ⇒ VJP = 0
            }
            // ℓ += ℓ_j ;  λ_{j-1} = λ_j + Δt_{j-1} λ_∂f∂H + ∂ℓ_j/∂H,  ∂ℓ_j/∂H = 2 w_j W (H_j - H_ref,j), w_j = Δt_j for LossH  (:218-242)
            if ((rc = launch_loss_seed(e, Hj, plane_ptr(e, e->href, j), plane_ptr(e, e->wmask, j), lam, vH, lam, dt, 2.0 * wH,
                                       e->d_loss, wH, 1)))
                return rc;
        }
        // velocity term of LossV / LossHV: ℓ, ∂ℓ/∂H into λ_{j-1} and ∂ℓ/∂θ (Losses.jl:293-390)
        if ((rc = velocity_loss_term(e, j, Hj, lam, wV, e->d_loss, e->d_Ssum))) return rc;
        // dLdθ += Δt_{j-1} · VJP_θ(λ_{j-1}, H_j)                                        (:245-249)
        if ((rc = launch_vjp(e, -1, lam, Hj, nullptr, false, true, e->d_Ssum, dt, 1))) return rc;
    }
    if (lam != e->plane[ODINN_FIELD_LAMBDA])  // (the two λ planes alternate in the fused form: leave λ(t_0) in FIELD_LAMBDA)
        ODINN_CUDA(e, cudaMemcpyAsync(e->plane[ODINN_FIELD_LAMBDA], lam, pbytes, cudaMemcpyDeviceToDevice, e->stream));
    {
        // The j = 1 pass of the reference (snapshot 0 here) updates no λ but still adds its loss terms: ℓ += ℓ_1 and
        // dLdθ += ∂ℓ∂θ[1] (gradient.jl:218-232, 252).  With the default LossH weights w_0 = 0 (safe_slice); user weights
        // (odinn_set_loss_weights) may be nonzero there, and the forward / reverse losses must agree (gradient.jl:259).
        const double wH0 = loss_weight_H(e, t, n_t, 0), wV0 = loss_weight_V(e, n_t, 0);
        void* H0s = plane_ptr(e, e->snap, 0);
        if (wH0 != 0.0 && (rc = launch_loss_seed(e, H0s, plane_ptr(e, e->href, 0), plane_ptr(e, e->wmask, 0), nullptr, nullptr,
                                                 nullptr, 0.0, 0.0, e->d_loss, wH0, 1)))
            return rc;
        if ((rc = velocity_loss_term(e, 0, H0s, nullptr, wV0, e->d_loss, e->d_Ssum))) return rc;
    }
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_loss, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S + e->G, e->d_Ssum, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    if (loss_out) memcpy(loss_out, e->h_S, sizeof(double) * e->G);
    if (Ssum_out) memcpy(Ssum_out, e->h_S + e->G, sizeof(double) * e->G);
    return ODINN_OK;
}

// ---- per-cell laws -----------------------------------------------------------------------------------------------

int odinn_law_cell_nn_set(odinn_ensemble* e, int kind, int n_layers, const int* widths, const int* acts, const double* theta,
                          int n_theta, const double* prescale_bounds, double max_NN, double n_H, double n_gS) {
    GUARD(e);
    if (kind != LAW_U && kind != LAW_Y) return fail(e, ODINN_EARG, "law kind must be 1 (LawU) or 2 (LawY)");
    if (n_layers < 1 || n_layers > MLP_MAX_LAYERS || !widths || !acts || !theta) return fail(e, ODINN_EARG, "bad MLP description");
    CellLaw lw{};
    lw.kind = kind;
    lw.arch.n_layers = n_layers;
    int np = 0;
    for (int L = 0; L <= n_layers; ++L) {
        if (widths[L] < 1 || widths[L] > LAW_MAX_WIDTH) return fail(e, ODINN_EARG, "per-cell law width out of range (1..32)");
        lw.arch.widths[L] = widths[L];
    }
    for (int L = 0; L < n_layers; ++L) {
        if (acts[L] < ACT_IDENTITY || acts[L] > ACT_RELU) return fail(e, ODINN_EARG, "unknown activation code");
        lw.arch.acts[L] = acts[L];
        np += widths[L] * widths[L + 1] + widths[L + 1];
    }
    if (widths[0] != 2 || widths[n_layers] != 1) return fail(e, ODINN_EARG, "a per-cell law maps two inputs to one output");
    if (np != n_theta) return fail(e, ODINN_EARG, "theta length does not match the architecture");
    lw.arch.n_params = np;
    lw.prescale = prescale_bounds ? 1 : 0;
    if (prescale_bounds) { lw.lo0 = prescale_bounds[0]; lw.hi0 = prescale_bounds[1]; lw.lo1 = prescale_bounds[2]; lw.hi1 = prescale_bounds[3]; }
    lw.postscale = (max_NN > 0.0) ? 1 : 0;
    lw.max_NN = max_NN;
    lw.n_H = n_H > 0.0 ? n_H : e->phys.n;
    lw.n_gS = n_gS > 0.0 ? n_gS : e->phys.n;
    if (e->law_n_theta != np) {
        if (e->d_law_theta) cudaFree(e->d_law_theta);
        if (e->d_law_partial) cudaFree(e->d_law_partial);
        if (e->d_law_dtheta) cudaFree(e->d_law_dtheta);
        e->d_law_theta = e->d_law_partial = e->d_law_dtheta = nullptr;
        e->law_n_theta = 0;
        ODINN_CUDA(e, cudaMalloc(&e->d_law_theta, sizeof(double) * np));
        // block partials of the theta pullback: one row per tile of the WHOLE ensemble (the fixed-architecture kernel covers every glacier
        // in one launch); the generic kernel and the knot pass use the first max_tiles_per_glacier rows
        ODINN_CUDA(e, cudaMalloc(&e->d_law_partial, sizeof(double) * (size_t)np * std::max(e->n_tiles, e->max_tiles_per_glacier)));
        ODINN_CUDA(e, cudaMalloc(&e->d_law_dtheta, sizeof(double) * (size_t)np * e->G));
        ODINN_CUDA(e, cudaMemsetAsync(e->d_law_dtheta, 0, sizeof(double) * (size_t)np * e->G, e->stream));
        e->law_n_theta = np;
    }
    ODINN_CUDA(e, cudaMemcpyAsync(e->d_law_theta, theta, sizeof(double) * np, cudaMemcpyHostToDevice, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));  // theta is caller-owned
    if (!e->law_cfg) e->law_cfg = new CellLaw();
    *law_of(e) = lw;
    law_refresh_phys(e);
    e->law_kind = kind;
    return ODINN_OK;
}

int odinn_law_cell_interp_set(odinn_ensemble* e, int n0, const double* knots0, int n1, const double* knots1) {
    GUARD(e);
    if (n0 == 0) {  // exact per-node gradients (interpolation = :None)
        e->ext_int[2] = e->ext_int[3] = 0;
        return ODINN_OK;
    }
    if (e->law_kind == LAW_NONE) return fail(e, ODINN_ESTATE, "set a per-cell law first (odinn_law_cell_nn_set)");
    if (n0 < 2 || !knots0 || n1 < 0 || (n1 > 0 && (n1 < 2 || !knots1))) return fail(e, ODINN_EARG, "an interpolation axis needs at least 2 knots");
    if (e->law_kind == LAW_U && n1 == 0) return fail(e, ODINN_EARG, "LawU interpolates over (Hbar, gradS): two knot vectors");
    if (e->law_kind == LAW_Y && n1 != 0) return fail(e, ODINN_EARG, "LawY interpolates over Hbar only: one knot vector");
    for (int k = 1; k < n0; ++k) if (!(knots0[k] > knots0[k - 1])) return fail(e, ODINN_EARG, "knots must be strictly increasing");
    for (int k = 1; k < n1; ++k) if (!(knots1[k] > knots1[k - 1])) return fail(e, ODINN_EARG, "knots must be strictly increasing");
    if (e->ext_dev[EXT_LAT_KNOTS]) cudaFree(e->ext_dev[EXT_LAT_KNOTS]);
    if (e->ext_dev[EXT_LAT_W]) cudaFree(e->ext_dev[EXT_LAT_W]);
    e->ext_dev[EXT_LAT_KNOTS] = e->ext_dev[EXT_LAT_W] = nullptr;
    e->ext_int[2] = e->ext_int[3] = 0;
    ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_LAT_KNOTS], sizeof(double) * (n0 + n1)));
    ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_LAT_W], sizeof(double) * (size_t)n0 * std::max(n1, 1)));
    ODINN_CUDA(e, cudaMemcpyAsync(e->ext_dev[EXT_LAT_KNOTS], knots0, sizeof(double) * n0, cudaMemcpyHostToDevice, e->stream));
    if (n1 > 0)
        ODINN_CUDA(e, cudaMemcpyAsync((double*)e->ext_dev[EXT_LAT_KNOTS] + n0, knots1, sizeof(double) * n1, cudaMemcpyHostToDevice, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    {   // the knot pass writes one row of block partials per 512 knots: grow the partial buffer when the lattice has more blocks than
        // the largest glacier has tiles
        const int nb = div_up(n0 * std::max(n1, 1), TX * TY);
        if (nb > std::max(e->n_tiles, e->max_tiles_per_glacier)) {
            if (e->d_law_partial) cudaFree(e->d_law_partial);
            e->d_law_partial = nullptr;
            ODINN_CUDA(e, cudaMalloc(&e->d_law_partial, sizeof(double) * (size_t)e->law_n_theta * nb));
        }
        e->max_tiles_per_glacier = std::max(e->max_tiles_per_glacier, nb);
    }
    e->ext_int[2] = n0;
    e->ext_int[3] = n1;
    return ODINN_OK;
}

int odinn_law_cell_clear(odinn_ensemble* e) {
    GUARD(e);
    e->law_kind = LAW_NONE;
    return ODINN_OK;
}

int odinn_sia2d_vjp_theta_cell(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                               double* out_theta, int n_theta, double t) {
    GUARD(e);
    (void)t;
    if (e->law_kind == LAW_NONE) return fail(e, ODINN_ESTATE, "no per-cell law is set (odinn_law_cell_nn_set)");
    if (!out_theta || n_theta != e->law_n_theta) return fail(e, ODINN_EARG, "out_theta / n_theta do not match the law");
    int rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_H, const_cast<void*>(H), ldH, true, e->stream))) return rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_LAMBDA, const_cast<void*>(lambda), ldl, true, e->stream))) return rc;
    if ((rc = launch_vjp(e, glacier, e->plane[ODINN_FIELD_LAMBDA], e->plane[ODINN_FIELD_H], nullptr, false, true))) return rc;
    std::vector<double> tmp(n_theta);
    ODINN_CUDA(e, cudaMemcpyAsync(tmp.data(), e->d_law_dtheta + (size_t)glacier * n_theta, sizeof(double) * n_theta,
                                  cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    memcpy(out_theta, tmp.data(), sizeof(double) * n_theta);
    return ODINN_OK;
}

int odinn_law_cell_grad(odinn_ensemble* e, double* out, int n_theta) {
    GUARD(e);
    if (e->law_kind == LAW_NONE) return fail(e, ODINN_ESTATE, "no per-cell law is set (odinn_law_cell_nn_set)");
    if (!out || n_theta != e->law_n_theta) return fail(e, ODINN_EARG, "out / n_theta do not match the law");
    std::vector<double> tmp((size_t)n_theta * e->G);
    ODINN_CUDA(e, cudaMemcpyAsync(tmp.data(), e->d_law_dtheta, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    memcpy(out, tmp.data(), sizeof(double) * tmp.size());
    return ODINN_OK;
}

// ---- laws ------------------------------------------------------------------------------------------------------

int odinn_law_A_nn_apply(odinn_ensemble* e, int n_layers, const int* widths, const int* acts, const double* theta,
                         int n_theta, double* A_out) {
    GUARD(e);
    if (n_layers < 1 || n_layers > MLP_MAX_LAYERS || !widths || !acts || !theta)
        return fail(e, ODINN_EARG, "bad MLP description");
    MlpArch arch{};
    arch.n_layers = n_layers;
    int np = 0;
    for (int L = 0; L <= n_layers; ++L) {
        if (widths[L] < 1 || widths[L] > MLP_MAX_WIDTH) return fail(e, ODINN_EARG, "MLP width out of range (1..64)");
        arch.widths[L] = widths[L];
    }
    for (int L = 0; L < n_layers; ++L) {
        if (acts[L] < ACT_IDENTITY || acts[L] > ACT_RELU) return fail(e, ODINN_EARG, "unknown activation code");
        arch.acts[L] = acts[L];
        np += widths[L] * widths[L + 1] + widths[L + 1];
    }
    if (widths[0] != 1 || widths[n_layers] != 1) return fail(e, ODINN_EARG, "the A law maps one temperature to one A");
    if (np != n_theta) return fail(e, ODINN_EARG, "theta length does not match the architecture");
    arch.n_params = np;
    if (e->n_theta != np) {
        if (e->d_theta) cudaFree(e->d_theta);
        if (e->d_J) cudaFree(e->d_J);
        if (e->d_dtheta) cudaFree(e->d_dtheta);
        e->d_theta = e->d_J = e->d_dtheta = nullptr;
        e->n_theta = 0;
        ODINN_CUDA(e, cudaMalloc(&e->d_theta, sizeof(double) * np));
        ODINN_CUDA(e, cudaMalloc(&e->d_J, sizeof(double) * np * e->G));
        ODINN_CUDA(e, cudaMalloc(&e->d_dtheta, sizeof(double) * np));
        e->n_theta = np;
    }
    if (!e->d_temps) ODINN_CUDA(e, cudaMalloc(&e->d_temps, sizeof(double) * e->G));
    int rc = sync_descs(e);
    if (rc) return rc;
    std::vector<double> temps(e->G);
    for (int g = 0; g < e->G; ++g) temps[g] = e->gl[g].temp;
    ODINN_CUDA(e, cudaMemcpyAsync(e->d_temps, temps.data(), sizeof(double) * e->G, cudaMemcpyHostToDevice, e->stream));
    ODINN_CUDA(e, cudaMemcpyAsync(e->d_theta, theta, sizeof(double) * np, cudaMemcpyHostToDevice, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));  // temps / theta are caller- or stack-owned
    law_A_nn_kernel<<<div_up(e->G, 64), 64, 0, e->stream>>>(arch, (const double*)e->d_theta, (const double*)e->d_temps,
                                                           e->G, e->phys.minA, e->phys.maxA, e->d_A, (double*)e->d_J);
    ODINN_CHECK_LAUNCH(e);
    if (e->dtype == ODINN_F32)
        set_A_kernel<float><<<div_up(e->G, 128), 128, 0, e->stream>>>((GDesc<float>*)e->d_descs, e->d_A, e->G);
    else
        set_A_kernel<double><<<div_up(e->G, 128), 128, 0, e->stream>>>((GDesc<double>*)e->d_descs, e->d_A, e->G);
    ODINN_CHECK_LAUNCH(e);
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_A, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    for (int g = 0; g < e->G; ++g) e->gl[g].A = e->h_S[g];  // keep the host mirror in step (descs stay clean)
    if (A_out) memcpy(A_out, e->h_S, sizeof(double) * e->G);
    return ODINN_OK;
}

int odinn_law_A_nn_pullback(odinn_ensemble* e, const double* S, double* dtheta, int n_theta) {
    GUARD(e);
    if (!e->d_J || e->n_theta != n_theta) return fail(e, ODINN_ESTATE, "call odinn_law_A_nn_apply first");
    if (!dtheta) return fail(e, ODINN_EARG, "dtheta is null");
    const double* dS = e->d_Ssum;  // default: the sums left by odinn_grad_discrete
    if (S) {
        ODINN_CUDA(e, cudaMemcpyAsync(e->d_S, S, sizeof(double) * e->G, cudaMemcpyHostToDevice, e->stream));
        dS = e->d_S;
    }
    law_pullback_kernel<<<div_up(n_theta, 128), 128, 0, e->stream>>>((const double*)e->d_J, dS, e->G, n_theta,
                                                                     (double*)e->d_dtheta);
    ODINN_CHECK_LAUNCH(e);
    std::vector<double> tmp(n_theta);
    ODINN_CUDA(e, cudaMemcpyAsync(tmp.data(), e->d_dtheta, sizeof(double) * n_theta, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    memcpy(dtheta, tmp.data(), sizeof(double) * n_theta);
    return ODINN_OK;
}

}  // extern "C"
