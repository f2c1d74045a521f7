// Host-side state behind an odinn_ensemble handle: glacier descriptors, device planes, tile table.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/odinn_b200.h"
#include "common.cuh"

struct GlacierHost {
    int nx, ny, ld;
    long long off;
    long long off_packed;  // element offset in the packed (ld = nx) layout of the host-batch path
    double dx, dy;
    double A;
    double temp;
    int tile0, ntx, nty;
    int item0, n_items;
    int item20, n_items2;  // work items of the two-column fp32 kernels (sia2d_march2.cuh)
};

struct odinn_ensemble {
    int device = 0;
    int dtype = ODINN_F64;
    int G = 0;
    size_t esize = 8;
    std::vector<GlacierHost> gl;
    long long total = 0;  // elements per plane (padded)
    long long cells = 0;  // Σ nx*ny
    odinn_phys phys{};
    bool cubic = false;   // n == 3 && C == 0
    int a_gridded = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    void* plane[ODINN_FIELD_COUNT_] = {nullptr};
    void* d_descs = nullptr;
    bool descs_dirty = true;
    int2* d_tiles = nullptr;
    int* d_tile_start = nullptr;
    int n_tiles = 0;
    int4* d_items = nullptr;      // marching work items {glacier, first column, row0, row1}
    int* d_item_start = nullptr;
    int n_items = 0;
    int chunk_rows = 32;
    int4* d_items2 = nullptr;     // two-column strips {glacier, first column (even), row0, row1}
    int* d_item2_start = nullptr;
    int n_items2 = 0;
    int chunk_rows2 = 32;
    int march = 4;                // fp32 kernel generation: 1 = one column per lane, 2 = two columns + f32x2, 4 = 2 + the fused step through the 2-D TMA ring
    void* tma_cache = nullptr;    // tensor maps per plane pointer, band work items (launch_march2.cu)
    double* tma_partial_live = nullptr;  // partial sums of the last fused launch when it went through the TMA variant (else nullptr)
    bool all_nx_even = true;
    int cluster_mode = -1;        // odinn_set_cluster_mode: -1 automatic, 0 off, else the forced cluster size (launch_cluster.cu)
    bool no_fuse = false;         // ODINN_NO_FUSE=1: never use the fused F1 + A1 + A2 kernel
    double* d_partial = nullptr;  // per-item / per-tile partial sums (two-stage, fixed-order reductions)
    double* d_S = nullptr;        // [4 x G]: S | Ssum | loss | A
    double* d_Ssum = nullptr;
    double* d_loss = nullptr;
    double* d_A = nullptr;
    double* h_S = nullptr;        // pinned mirror of d_S
    // on-device time loop / gradient state
    void* snap = nullptr;         // n_snap planes: forward snapshots H(t_j)        (gradient.jl:73,140)
    void* href = nullptr;         // n_ref planes: reference thickness H_ref(t_j)
    void* wmask = nullptr;        // n_ref planes: is_in_glacier mask / (nx ny)
    void* work[2] = {nullptr, nullptr};
    int n_snap = 0, n_ref = 0;
    std::vector<char> ref_has;    // [n_ref] snapshot j holds thickness data (odinn_set_reference was called for it): tH_ref of gradient.jl:79
    // law state (LawA(nn, params))
    void* d_theta = nullptr;
    void* d_J = nullptr;          // [G x n_theta] dA_g/dθ
    void* d_dtheta = nullptr;
    double* d_temps = nullptr;
    int n_theta = 0;
    // per-cell law state (LawU / LawY, sia2d_law.cuh)
    int law_kind = 0;             // LAW_NONE | LAW_U | LAW_Y
    void* law_cfg = nullptr;      // host copy of the CellLaw struct (opaque here)
    double* d_law_theta = nullptr;
    int law_n_theta = 0;
    void* lawD = nullptr;         // node planes: D, alpha, beta
    void* lawAl = nullptr;
    void* lawBe = nullptr;
    double* d_law_partial = nullptr;   // [max tiles per glacier x n_theta]
    double* d_law_dtheta = nullptr;    // [G x n_theta] last / accumulated theta-gradient per glacier
    int max_tiles_per_glacier = 0;
    long long launches = 0;
    std::string err;

    // host-batch path (odinn_fwd_adj_batch_host): packed copy of B, chunk events for the
    // H2D -> compute -> D2H pipeline over copy_stream[0] / stream / copy_stream[1]
    long long batch_chunk_cells = 8LL << 20;  // sweep profiles/r01_v6_sweep.txt: 2 Mi 4.94, 4 Mi 5.25, 8 Mi 5.37, 16 Mi 5.04, 32 Mi 4.13 G cell-steps/s
    void* bpack = nullptr;
    bool bpack_dirty = true;
    void* stage[4] = {nullptr, nullptr, nullptr, nullptr};  // H, lambda, dH, vjp_H of the host-batch call
    std::vector<cudaEvent_t> ev_up, ev_done;
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;

    // adaptive forward solve (adaptive.cu): k1..k4, trial state, stage input; per-glacier controller state
    void* ad_plane[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void* d_ad_state = nullptr;
    int* d_ad_dims = nullptr;     // [nx | ny | n_active]
    int* h_ad_active = nullptr;   // pinned
    // device / pinned-host allocations owned by the other translation units (freed by odinn_ensemble_destroy)
    void* ext_dev[32] = {nullptr};
    void* ext_host[4] = {nullptr};
    int ext_int[8] = {0};
    // CUDA graph of one tstop interval of the fixed-step forward loop (odinn_solve_forward) + its device table of step sizes
    cudaGraphExec_t fwd_graph_exec = nullptr;
    std::string fwd_graph_key;
    int fwd_graph_launches = 0;
    double* d_fwd_tab = nullptr;      // [fwd_tab_len x 9] stage coefficients, then one int: the interval counter
    int fwd_tab_len = 0;
    // loss configuration (odinn_set_loss_weights): per-snapshot multipliers of the thickness / velocity L2 terms; empty = LossH with Δt
    std::vector<double> loss_wH, loss_wV;
    int lossV_component = 0;          // 0 :xy, 1 :abs  (Losses.jl:318-325)
    double lossV_theta_scale = 0.0;   // continuous adjoint: multiplier of the quadrature-weighted dl_V/dtheta (1 LossV, `scaling` LossHV; gradient.jl:474-507)
    int lossV_scale_loss = 1;         // LossV.scale_loss (Losses.jl:327-331), needed to rebuild the weights of interpolated references
    std::vector<int> v_snap;          // snapshot index of every velocity-reference slot (velocity.cu)
    // mass balance (massbalance.cu): snapshot indices after which the MB callback fires, per-glacier climate scalars per MB step
    std::vector<int> mb_snap;
    std::vector<double> mb_params;    // [n_mb x G x MB_NPAR]
};
enum {  // ext_dev slots
    EXT_CA_HT = 0, EXT_CA_LAM1 = 3, EXT_CA_LAM2 = 4,                                                   // continuous adjoint (contadj.cu)
    EXT_MB = 8, EXT_MB_PAR = 9,                                                                      // mass balance (massbalance.cu)
    EXT_V_REF = 12, EXT_V_WORK0 = 13, EXT_V_WORK1 = 14, EXT_V_WORK2 = 15, EXT_V_PARTIAL = 16,          // surface velocity / LossV
    EXT_AD_PARTIAL = 17,                                                                            // adaptive solve (ext_int[0] = its length)
    EXT_LS_PARTIAL = 18,                                                                            // loss / seed pass (ext_int[1] = its length)
    EXT_RK_STATE = 19,                                                                              // RDPK3Sp35 per-glacier controller state (rdpk.cu)
    EXT_VQ_WORK = 20,                                                                               // 4 planes: velocity references interpolated at a quadrature node
    EXT_VQ_RED = 21,                                                                                // [2 G] mask count and sum of squares of those references
    EXT_CL_TIMES = 24, EXT_CL_RKSTATE = 25, EXT_CL_STEPS = 26,                                                                              // cluster-resident forward solve: time grid on the device (ext_int[4] = its length)
    EXT_ITEMS2_LONG_START = 28,                                                                     // [G + 1] first long item of every glacier (indexes the partial sums of a launch over the long table)
    EXT_ITEMS2_LONG = 27,                                                                           // long-chunk two-column work items for the F1 / stage launches of big ensembles (ext_int[5] = their number)
    EXT_LAT_KNOTS = 22, EXT_LAT_W = 23                                                              // law pullback with interpolation = :Linear: knots, knot weights (ext_int[2], [3] = n0, n1)
};

namespace odinn {

extern thread_local std::string g_create_error;

inline int fail(odinn_ensemble* e, int code, const std::string& msg) {
    if (e) e->err = msg;
    else g_create_error = msg;
    return code;
}

#define ODINN_CUDA(e, call)                                                                              \
    do {                                                                                                 \
        cudaError_t _st = (call);                                                                        \
        if (_st != cudaSuccess)                                                                          \
            return odinn::fail((e), ODINN_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_st));   \
    } while (0)

#define ODINN_CHECK_LAUNCH(e)                                                                            \
    do {                                                                                                 \
        cudaError_t _st = cudaGetLastError();                                                            \
        if (_st != cudaSuccess)                                                                          \
            return odinn::fail((e), ODINN_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(_st)); \
        (e)->launches++;                                                                                 \
    } while (0)

template <typename T>
inline PhysDev<T> make_phys(const odinn_phys& p) {
    PhysDev<T> d;
    d.n = (T)p.n;
    d.p = (T)p.p;
    d.q = (T)p.q;
    d.Gam = (T)(2.0 * std::pow(p.rho * p.g, p.n) / (p.n + 2.0));  // target_utils.jl:3-13
    d.Sl = (T)(p.C * std::pow(p.rho * p.g, p.p - p.q));           // target_utils.jl:15-19
    d.eta0 = (T)p.eta0;
    return d;
}

int ensure_plane(odinn_ensemble* e, int field);
int sync_descs(odinn_ensemble* e);
// shared with the other translation units of the library (adaptive.cu, ...)
int alloc_work_plane(odinn_ensemble* e, void** p, size_t n_planes = 1);      // zero-filled, no-op when *p is set
int rhs_planes(odinn_ensemble* e, const void* Hin, void* out);              // out <- SIA2D(Hin), whole ensemble, one launch
// One RDPK3Sp35 stage fused into F1 (rdpk.cu): S1out <- stage(S1in, k = SIA2D(S1in)) with the coefficients / planes of *rkfuse (RkFuse<T>);
// norm: the stage carries RKF_NORM -- the per-glacier sums of squares of the scaled error land in d_S.  Not for per-cell laws.
bool rhs_rk_fusable(const odinn_ensemble* e);
int rhs_planes_rk(odinn_ensemble* e, const void* S1in, void* S1out, const void* rkfuse, bool norm);
// The same for one stage of the continuous adjoint's reverse ODE: k = (dSIA/dH)^T S1in at H_itp = lerp(Ha, Hb) with the glacier's own
// weight a = (sign (t_g + c h_g) - ta) / (tb - ta); discrete VJP flavour, glacier-wide A (fp64: n = 3, C = 0).
bool vjp_rk_fusable(const odinn_ensemble* e);
int vjp_planes_rk(odinn_ensemble* e, const void* S1in, const void* Ha, const void* Hb, void* S1out, const void* rkfuse, double c, double sign,
                  double ta, double tb, bool norm);
int vjp_planes_lerp_S(odinn_ensemble* e, const void* lam, const void* Ha, const void* Hb, const void* rkstate, double sign, double ta, double tb,
                      double* S_dst, double scale, int accumulate);
int reduce_tiles(odinn_ensemble* e, const double* tile_partial, double* dst, double scale = 1.0, int accumulate = 0);  // dst[g] = Σ tiles of g
int prepare_snapshots(odinn_ensemble* e, int n_snap);
// A1 (wH: out <- (dSIA/dH)^T lam) and / or A2 (wS: S_dst[g] (+)= scale * S_g; nullptr -> the handle's d_S), discrete or continuous flavour
int vjp_planes(odinn_ensemble* e, const void* lam, const void* H, void* out, bool wH, bool wS, double* S_dst, double scale,
               int accumulate, bool continuous);
// one glacier's plane <-> host matrix (column-major, ld elements), on the handle's stream (asynchronous for pinned memory)
int copy_plane_2d(odinn_ensemble* e, int glacier, void* plane, void* host, int ld, bool to_device);
// velocity term of snapshot j in the loss / reverse loops (velocity.cu); no-op when there is no velocity data at j or w == 0
int velocity_loss_term(odinn_ensemble* e, int j, const void* Hj, void* lam, double w, double* loss_dst, double* S_dst);
// mass-balance callback of snapshot j / its adjoint (massbalance.cu); no-ops when no MB step fires at j
int mb_apply_step(odinn_ensemble* e, int j, void* H, int* applied);
int mb_adjoint_step(odinn_ensemble* e, int j, void* lam, const void* Hj);
// S_dst[g] += scale * dl_V/dtheta-scalar evaluated on the plane H against the velocity references interpolated linearly at time tq
// (one datum: constant; flat outside the data range) -- the quadrature-node term of the continuous adjoint (gradient.jl:289-301, 474-507)
int velocity_theta_term_interp(odinn_ensemble* e, double tq, const double* t, int n_t, const void* H, double scale, double* S_dst);
// cluster-resident solves of small glaciers (launch_cluster.cu).  kind: 0 fixed-step forward (Euler / SSPRK3), 1 RDPK3Sp35, 2 discrete-adjoint reverse loop, 3 continuous adjoint.
// cluster_plan: the cluster size the ensemble would run with (0: not eligible).  launch_interval_cluster: intervals j0+1 .. j1 of the
// time grid in one launch (snapshots j0+1 .. j1 and the final state written by the kernel).
int cluster_plan(odinn_ensemble* e, int kind);
int launch_interval_cluster(odinn_ensemble* e, int cs, int method, int nsub, int j0, int j1, const void* Hin, void* Hout, void* snap,
                            const double* d_t);
int solve_forward_rdpk_cluster(odinn_ensemble* e, int cs, int n_snap, const double* t, double reltol, double abstol, double dt0,
                               int max_steps, int* steps_out, int* rejected_out);
int launch_reverse_cluster(odinn_ensemble* e, int cs, int jhi, int jlo, const void* lam_in, void* lam_out, const double* d_t, const double* d_wH);
int grad_continuous_adaptive_cluster(odinn_ensemble* e, int cs, const double* t, int n_t, int n_q, const double* qn, const double* qw,
                                     double reltol, double abstol, double dtmax, int max_steps, int* steps_out);
void rdpk_host_coefficients(double G1[4], double G2[4], double G3[4], double D[4], double B[5], double E[5], double C[6]);
int upload_time_grid(odinn_ensemble* e, const double* t, int n_snap, const double** d_t);   // device copy of the tstops (EXT_CL_TIMES)
// adaptive forward solve with the reference's default integrator (rdpk.cu)
int solve_forward_rdpk(odinn_ensemble* e, int n_snap, const double* t, double reltol, double abstol, double dt0, int max_steps,
                       int* steps_out, int* rejected_out);
// Default LossH weight of snapshot j: ΔtH = diff(tH_ref) indexed by the position of t_j inside tH_ref -- the time since the PREVIOUS
// snapshot that holds thickness data, 0 for the first datum (safe_slice) and for tstops without data (gradient.jl:79-80, 144-149).
inline double loss_weight_H(const odinn_ensemble* e, const double* t, int n_t, int j) {
    if ((int)e->loss_wH.size() == n_t) return e->loss_wH[j];
    if ((int)e->ref_has.size() != n_t) return j > 0 ? t[j] - t[j - 1] : 0.0;
    if (!e->ref_has[j]) return 0.0;
    for (int p = j - 1; p >= 0; --p)
        if (e->ref_has[p]) return t[j] - t[p];
    return 0.0;
}
inline double loss_weight_V(const odinn_ensemble* e, int n_t, int j) { return (int)e->loss_wV.size() == n_t ? e->loss_wV[j] : 0.0; }
// loss_dst[g] (+)= wloss * sum W (H - Href)^2 ; optionally lam_out = lam_in + dt * v + cseed * W * (H - Href)
int loss_seed_planes(odinn_ensemble* e, const void* H, const void* Href, const void* W, const void* lam_in, const void* v,
                     void* lam_out, double dt, double cseed, double* loss_dst, double wloss, int accumulate);
void* snapshot_ptr(odinn_ensemble* e, int j);

}  // namespace odinn
