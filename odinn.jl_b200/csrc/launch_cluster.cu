// Launchers of the cluster-resident forward solves (sia2d_cluster.cuh): eligibility, cluster size, launch attributes, and the
// host loop of the adaptive solve (launch ranges between mass-balance callbacks).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "launch.cuh"
#include "sia2d_cluster.cuh"

namespace odinn {

namespace {

constexpr size_t CL_SMEM_MAX = 227 * 1024 - 1024;   // (the adaptive kernel also holds ~1 KB of static shared memory)

size_t smem_for(const odinn_ensemble* e, int cs, int n_planes) {
    size_t m = 0;
    for (const GlacierHost& g : e->gl) m = std::max(m, cl_smem_bytes(g.nx, g.ny, cs, e->esize, n_planes));
    return m;
}

void fill_config(odinn_ensemble* e, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, size_t smem, int cs) {
    cfg = {};
    cfg.gridDim = dim3(cs * e->G);
    cfg.blockDim = dim3(CL_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = e->stream;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
}

template <typename K>
cudaError_t configure(K kernel, size_t smem, int cs) {
    cudaError_t st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st != cudaSuccess) return st;
    if (cs > 8) st = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return st;
}

// Cells per thread and sweep item: 4 unless the band of the largest glacier holds fewer than 256 quads, then 2.  ODINN_CLUSTER_V forces
// it (tuning).  The generic-exponent kernels exist for V = 4 only.
int choose_v(const odinn_ensemble* e, int cs) {
    static const int env = []() { const char* v = getenv("ODINN_CLUSTER_V"); return v ? atoi(v) : 0; }();
    if (!e->cubic) return 4;
    if (env == 2 || env == 4) return env;
    long long cells = 0;
    for (const GlacierHost& g : e->gl) cells = std::max(cells, (long long)cl_band_rows(g.ny, cs) * g.nx);
    return cells >= 2LL * CL_NT ? 4 : 2;
}

// Dispatch on (element type, n == 3 && C == 0, eta0 == 1, V):  F<T, CUBIC, ETA1, V>::run(args...)
template <template <typename, bool, bool, int> class F, typename... Args>
int dispatch(odinn_ensemble* e, int v, Args&&... args) {
    const bool eta1 = (e->phys.eta0 == 1.0);
#define DC(T, CUB, E1, VV) return F<T, CUB, E1, VV>::run(e, args...)
#define DV(T, E1) do { if (v == 2) DC(T, true, E1, 2); else DC(T, true, E1, 4); } while (0)
    if (e->dtype == ODINN_F32) {
        if (e->cubic) { if (eta1) DV(float, true); else DV(float, false); }
        else { if (eta1) DC(float, false, true, 4); else DC(float, false, false, 4); }
    } else {
        if (e->cubic) { if (eta1) DV(double, true); else DV(double, false); }
        else { if (eta1) DC(double, false, true, 4); else DC(double, false, false, 4); }
    }
#undef DV
#undef DC
}

template <typename T, bool CUBIC, bool ETA1, int V>
struct MaxClusters {
    static int run(odinn_ensemble* e, int kind, size_t smem, int cs, int* n) {
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        fill_config(e, cfg, at, smem, cs);
        cudaError_t st;
        if (kind == 0) {
            st = configure(sia2d_interval_cluster<T, CUBIC, ETA1, V>, smem, cs);
            if (st == cudaSuccess) st = cudaOccupancyMaxActiveClusters(n, sia2d_interval_cluster<T, CUBIC, ETA1, V>, &cfg);
        } else if (kind == 1) {
            st = configure(sia2d_rdpk_cluster<T, CUBIC, ETA1, V>, smem, cs);
            if (st == cudaSuccess) st = cudaOccupancyMaxActiveClusters(n, sia2d_rdpk_cluster<T, CUBIC, ETA1, V>, &cfg);
        } else if (kind == 2) {
            st = configure(sia2d_reverse_cluster<T, CUBIC, ETA1, V>, smem, cs);
            if (st == cudaSuccess) st = cudaOccupancyMaxActiveClusters(n, sia2d_reverse_cluster<T, CUBIC, ETA1, V>, &cfg);
        } else {
            st = configure(sia2d_contadj_cluster<T, CUBIC, ETA1, V>, smem, cs);
            if (st == cudaSuccess) st = cudaOccupancyMaxActiveClusters(n, sia2d_contadj_cluster<T, CUBIC, ETA1, V>, &cfg);
        }
        if (st != cudaSuccess) { cudaGetLastError(); *n = 0; }
        return ODINN_OK;
    }
};

template <typename T, bool CUBIC, bool ETA1, int V>
struct LaunchFixed {
    static int run(odinn_ensemble* e, int cs, size_t smem, int method, int nsub, int j0, int j1, const void* Hin, void* Hout, void* snap,
                   const double* d_t) {
        auto k = sia2d_interval_cluster<T, CUBIC, ETA1, V>;
        ODINN_CUDA(e, configure(k, smem, cs));   // (per device: cheap next to a launch that runs whole intervals)
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        fill_config(e, cfg, at, smem, cs);
        ODINN_CUDA(e, cudaLaunchKernelEx(&cfg, k, (const GDesc<T>*)e->d_descs, (const T*)Hin, (const T*)e->plane[ODINN_FIELD_B], (T*)Hout,
                                         (T*)snap, (long long)e->total, d_t, j0, j1, nsub, method == ODINN_EULER ? 0 : 1,
                                         make_phys<T>(e->phys)));
        e->launches++;
        return ODINN_OK;
    }
};

template <typename T, bool CUBIC, bool ETA1, int V>
struct LaunchRdpk {
    static int run(odinn_ensemble* e, int cs, size_t smem, int j0, int j1, const void* Hin, void* Hout, void* snap, const double* d_t,
                   ClRkState* states, double reltol, double abstol, double dtmax, double dt0, int max_steps, const RdpkCoef& cf) {
        auto k = sia2d_rdpk_cluster<T, CUBIC, ETA1, V>;
        ODINN_CUDA(e, configure(k, smem, cs));
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        fill_config(e, cfg, at, smem, cs);
        ODINN_CUDA(e, cudaLaunchKernelEx(&cfg, k, (const GDesc<T>*)e->d_descs, (const T*)Hin, (const T*)e->plane[ODINN_FIELD_B], (T*)Hout,
                                         (T*)snap, (long long)e->total, d_t, j0, j1, states, reltol, abstol, dtmax, dt0, max_steps, cf,
                                         make_phys<T>(e->phys)));
        e->launches++;
        return ODINN_OK;
    }
};

template <typename T, bool CUBIC, bool ETA1, int V>
struct LaunchReverse {
    static int run(odinn_ensemble* e, int cs, size_t smem, int jhi, int jlo, const void* lam_in, void* lam_out, const double* d_t,
                   const double* d_wH) {
        auto k = sia2d_reverse_cluster<T, CUBIC, ETA1, V>;
        ODINN_CUDA(e, configure(k, smem, cs));
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        fill_config(e, cfg, at, smem, cs);
        ODINN_CUDA(e, cudaLaunchKernelEx(&cfg, k, (const GDesc<T>*)e->d_descs, (const T*)e->plane[ODINN_FIELD_B], (const T*)lam_in, (T*)lam_out,
                                         (const T*)e->snap, (const T*)e->href, (const T*)e->wmask, (long long)e->total, d_t, d_wH, jhi, jlo,
                                         e->d_loss, e->d_Ssum, make_phys<T>(e->phys)));
        e->launches++;
        return ODINN_OK;
    }
};

template <typename T, bool CUBIC, bool ETA1, int V>
struct LaunchContAdj {
    static int run(odinn_ensemble* e, int cs, size_t smem, void* lam_out, const double* d_tab, int n_t, int n_ev, int n_q, double reltol,
                   double abstol, double dtmax, int max_steps, int* d_steps, const RdpkCoef& cf) {
        auto k = sia2d_contadj_cluster<T, CUBIC, ETA1, V>;
        ODINN_CUDA(e, configure(k, smem, cs));
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        fill_config(e, cfg, at, smem, cs);
        ODINN_CUDA(e, cudaLaunchKernelEx(&cfg, k, (const GDesc<T>*)e->d_descs, (const T*)e->plane[ODINN_FIELD_B], (T*)lam_out, (const T*)e->snap,
                                         (const T*)e->href, (const T*)e->wmask, (long long)e->total, d_tab, n_t, n_ev, n_q, reltol, abstol, dtmax,
                                         max_steps, e->d_loss, e->d_Ssum, d_steps, cf, make_phys<T>(e->phys)));
        e->launches++;
        return ODINN_OK;
    }
};

}  // namespace

// Cluster size the ensemble runs with, 0 when the cluster path does not apply: a glacier too large for the shared memory of a
// cluster, more glaciers than clusters that are co-resident (latency is what this path buys: an ensemble that needs several waves
// is a throughput problem, and there the marching kernels execute 2.5x fewer instructions per cell), gridded A, per-cell law,
// ODINN_CLUSTER=0 / odinn_set_cluster_mode(0).  Otherwise the largest size whose clusters are all co-resident.
int cluster_plan(odinn_ensemble* e, int kind) {
    static const int env = []() { const char* v = getenv("ODINN_CLUSTER"); return v ? atoi(v) : -1; }();   // 0: off; 1..16: this size
    const int forced = e->cluster_mode >= 0 ? e->cluster_mode : env;
    if (forced == 0 || e->law_kind != 0 || e->a_gridded) return 0;
    const int n_planes = kind == 0 ? CL_PLANES_FIXED : (kind == 1 ? CL_PLANES_RDPK : (kind == 2 ? CL_PLANES_REV : CL_PLANES_CA));
    const int sizes[5] = {16, 8, 4, 2, 1};
    for (int cs : sizes) {
        if (forced > 0 && cs != forced) continue;
        const size_t smem = smem_for(e, cs, n_planes);
        if (smem > CL_SMEM_MAX) continue;
        int n = 0;
        dispatch<MaxClusters>(e, choose_v(e, cs), kind, smem, cs, &n);
        if (n <= 0) continue;
        if (n >= e->G || forced > 0) return cs;
    }
    return 0;
}

int launch_interval_cluster(odinn_ensemble* e, int cs, int method, int nsub, int j0, int j1, const void* Hin, void* Hout, void* snap,
                            const double* d_t) {
    return dispatch<LaunchFixed>(e, choose_v(e, cs), cs, smem_for(e, cs, CL_PLANES_FIXED), method, nsub, j0, j1, Hin, Hout, snap, d_t);
}

// Steps jhi .. jlo+1 of the discrete-adjoint reverse loop in one launch (lam_in == nullptr: lambda = 0); the loss and S terms are ADDED to
// the handle's d_loss / d_Ssum.  d_t: device time grid, d_wH: device LossH weight per snapshot.
int launch_reverse_cluster(odinn_ensemble* e, int cs, int jhi, int jlo, const void* lam_in, void* lam_out, const double* d_t, const double* d_wH) {
    // (fp64: two cells per item -- the adjoint cell sweep holds ~100 values per item and spills at four)
    const int v = e->dtype == ODINN_F64 ? 2 : choose_v(e, cs);
    return dispatch<LaunchReverse>(e, v, cs, smem_for(e, cs, CL_PLANES_REV), jhi, jlo, lam_in, lam_out, d_t, d_wH);
}

// ContinuousAdjoint gradient, cluster-resident (sia2d_contadj_cluster): the event list is built here exactly as in
// grad_continuous_adaptive_t (rdpk.cu) -- stops in tau = -t: sort(unique(vcat(-reverse(tstops), -t_nodes))), gradient.jl:456.
// d_loss / d_Ssum must be zeroed by the caller; lambda(t_0) is left in FIELD_LAMBDA.
int grad_continuous_adaptive_cluster(odinn_ensemble* e, int cs, const double* t, int n_t, int n_q, const double* qn, const double* qw,
                                     double reltol, double abstol, double dtmax, int max_steps, int* steps_out) {
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_LAMBDA)) || (rc = ensure_plane(e, ODINN_FIELD_B))) return rc;
    struct Ev { double tau; int is_q; int idx; };
    std::vector<Ev> ev;
    for (int j = 0; j < n_t; ++j) ev.push_back({-t[j], 0, j});
    for (int m = 0; m < n_q; ++m) {
        bool dup = false;
        for (int j = 0; j < n_t; ++j) dup |= (qn[m] == t[j]);
        if (!dup) ev.push_back({-qn[m], 1, m});
    }
    std::stable_sort(ev.begin(), ev.end(), [](const Ev& a, const Ev& b) { return a.tau < b.tau; });
    const int n_ev = (int)ev.size();
    std::vector<double> tab;
    tab.reserve(2 * (size_t)n_t + 2 * (size_t)n_ev + 2 * (size_t)n_q);
    for (int j = 0; j < n_t; ++j) tab.push_back(t[j]);
    for (int j = 0; j < n_t; ++j) tab.push_back(loss_weight_H(e, t, n_t, j));
    for (const Ev& s : ev) tab.push_back(s.tau);
    for (const Ev& s : ev) tab.push_back((double)(s.idx + (s.is_q ? (1 << 20) : 0)));
    for (int m = 0; m < n_q; ++m) tab.push_back(qn[m]);
    for (int m = 0; m < n_q; ++m) tab.push_back(qw[m]);
    const double* d_tab = nullptr;
    if ((rc = upload_time_grid(e, tab.data(), (int)tab.size(), &d_tab))) return rc;
    if (!e->ext_dev[EXT_CL_STEPS]) ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_CL_STEPS], sizeof(int) * e->G));
    int* d_steps = (int*)e->ext_dev[EXT_CL_STEPS];
    RdpkCoef cf;
    rdpk_host_coefficients(cf.G1, cf.G2, cf.G3, cf.D, cf.B, cf.E, cf.C);
    const int v = e->dtype == ODINN_F64 ? 2 : choose_v(e, cs);
    if ((rc = dispatch<LaunchContAdj>(e, v, cs, smem_for(e, cs, CL_PLANES_CA), e->plane[ODINN_FIELD_LAMBDA], d_tab, n_t, n_ev, n_q, reltol, abstol,
                                      dtmax, max_steps, d_steps, cf)))
        return rc;
    std::vector<int> hs(e->G);
    ODINN_CUDA(e, cudaMemcpyAsync(hs.data(), d_steps, sizeof(int) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    for (int g = 0; g < e->G; ++g) {
        if (hs[g] < 0) return fail(e, ODINN_ESTATE, "rdpk3sp35: too many steps (maxiters)");
        if (steps_out) steps_out[g] = hs[g];
    }
    return ODINN_OK;
}

int upload_time_grid(odinn_ensemble* e, const double* t, int n_snap, const double** d_t) {
    if (e->ext_int[4] < n_snap) {
        if (e->ext_dev[EXT_CL_TIMES]) cudaFree(e->ext_dev[EXT_CL_TIMES]);
        e->ext_dev[EXT_CL_TIMES] = nullptr;
        e->ext_int[4] = 0;
        ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_CL_TIMES], sizeof(double) * (size_t)n_snap));
        e->ext_int[4] = n_snap;
    }
    ODINN_CUDA(e, cudaMemcpyAsync(e->ext_dev[EXT_CL_TIMES], t, sizeof(double) * (size_t)n_snap, cudaMemcpyHostToDevice, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));   // (t is the caller's pageable memory)
    *d_t = (const double*)e->ext_dev[EXT_CL_TIMES];
    return ODINN_OK;
}

// Adaptive forward solve with the reference's default integrator, cluster-resident: per glacier one cluster runs the whole
// `while t < tstop` loop of every interval on the device -- no host round trip per trial step (solve_forward_rdpk in rdpk.cu reads
// two integers back per step and launches ~15 kernels for it).  A launch range ends where a mass-balance callback fires.
int solve_forward_rdpk_cluster(odinn_ensemble* e, int cs, int n_snap, const double* t, double reltol, double abstol, double dt0,
                               int max_steps, int* steps_out, int* rejected_out) {
    int rc;
    const size_t pbytes = (size_t)e->total * e->esize;
    const double* d_t = nullptr;
    if ((rc = upload_time_grid(e, t, n_snap, &d_t))) return rc;
    if (!e->ext_dev[EXT_CL_RKSTATE]) ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_CL_RKSTATE], sizeof(ClRkState) * e->G));
    ClRkState* states = (ClRkState*)e->ext_dev[EXT_CL_RKSTATE];
    ODINN_CUDA(e, cudaMemsetAsync(states, 0, sizeof(ClRkState) * e->G, e->stream));
    RdpkCoef cf;
    rdpk_host_coefficients(cf.G1, cf.G2, cf.G3, cf.D, cf.B, cf.E, cf.C);
    void* Hs = e->plane[ODINN_FIELD_H];
    ODINN_CUDA(e, cudaMemcpyAsync(snapshot_ptr(e, 0), e->plane[ODINN_FIELD_H0], pbytes, cudaMemcpyDeviceToDevice, e->stream));
    if (n_snap == 1) ODINN_CUDA(e, cudaMemcpyAsync(Hs, e->plane[ODINN_FIELD_H0], pbytes, cudaMemcpyDeviceToDevice, e->stream));
    const double dtmax = std::fabs(t[n_snap - 1] - t[0]);
    const size_t smem = smem_for(e, cs, CL_PLANES_RDPK);
    const void* Hin = e->plane[ODINN_FIELD_H0];
    int j0 = 0;
    while (j0 < n_snap - 1) {
        int j1 = n_snap - 1;
        for (int m : e->mb_snap) if (m > j0 && m < j1) j1 = m;
        if ((rc = dispatch<LaunchRdpk>(e, choose_v(e, cs), cs, smem, j0, j1, Hin, Hs, snapshot_ptr(e, 0), d_t, states, reltol, abstol, dtmax, dt0, max_steps, cf)))
            return rc;
        int applied = 0;
        if ((rc = mb_apply_step(e, j1, Hs, &applied))) return rc;   // mass-balance callback at the end of its window (inversion_utils.jl:498-517)
        if (applied) ODINN_CUDA(e, cudaMemcpyAsync(snapshot_ptr(e, j1), Hs, pbytes, cudaMemcpyDeviceToDevice, e->stream));
        Hin = Hs;
        j0 = j1;
    }
    std::vector<ClRkState> hs(e->G);
    ODINN_CUDA(e, cudaMemcpyAsync(hs.data(), states, sizeof(ClRkState) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    for (int g = 0; g < e->G; ++g) {
        if (n_snap > 1 && hs[g].started < 0) return fail(e, ODINN_ESTATE, "rdpk3sp35: too many steps (maxiters)");
        if (steps_out) steps_out[g] = hs[g].steps;
        if (rejected_out) rejected_out[g] = hs[g].rejected;
    }
    return ODINN_OK;
}

}  // namespace odinn
