/*
 * odinn_b200.h -- C ABI of libodinn_b200.so: the B200 (sm_100a) SIA2D hot path of ODINN.jl.
 *
 * This is the drop-in boundary.  Every entry point names the reference (ODINN.jl v1.1.0,
 * commit 31dfbf2) interface it replaces as  file:line  relative to the reference tree.
 * Plain pointers and sizes only; no C++/torch types.  All matrices are Julia `Matrix`
 * layout: column-major, element (i,j) at  ptr[i + j*ld],  i = 0..nx-1 the fast ("x") axis,
 * ld >= nx in ELEMENTS.  Element type is the ensemble's dtype (Sleipnir.Float: Float64 by
 * default, Float32 build -- test/SIA2D_adjoint_utils.jl:22).
 *
 * Ownership: the caller owns host buffers for the duration of a call only; the library owns
 * every device buffer behind the handle.  A handle is bound to ONE CUDA device and ONE stream;
 * calls on one handle must not overlap; different handles are independent (one handle per
 * worker process = the reference's pmap worker, src/inverse/SIA2D/gradient.jl:9-10).
 *
 * Errors: every function returns 0 on success and a negative odinn_status otherwise; the
 * message is available from odinn_last_error().  The library never calls exit/abort.
 */
#ifndef ODINN_B200_H
#define ODINN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct odinn_ensemble odinn_ensemble;

enum odinn_dtype { ODINN_F32 = 0, ODINN_F64 = 1 };

enum odinn_method {
    ODINN_EULER = 0, ODINN_SSPRK3 = 1,  /* fixed sub-steps, odinn_solve_forward */
    ODINN_BS3 = 2,                      /* adaptive Bogacki-Shampine 3(2), odinn_solve_forward_adaptive */
    ODINN_RDPK3SP35 = 3                 /* adaptive RDPK3Sp35 + PID controller: the reference's default solver (src/inverse/AdjointTypes.jl:60) */
};

enum odinn_activation { ODINN_ACT_IDENTITY = 0, ODINN_ACT_SOFTPLUS = 1, ODINN_ACT_SIGMOID = 2, ODINN_ACT_TANH = 3, ODINN_ACT_RELU = 4 };

enum odinn_status {
    ODINN_OK = 0,
    ODINN_EARG = -1,   /* bad argument (null pointer, size mismatch, index out of range) */
    ODINN_ECUDA = -2,  /* a CUDA runtime call or kernel launch failed                    */
    ODINN_ESTATE = -3, /* call sequence error (e.g. gradient before forward solve)       */
    ODINN_ENOMEM = -4
};

/* Device-resident per-glacier planes that can be uploaded / downloaded. */
enum odinn_field {
    ODINN_FIELD_B = 0,       /* bedrock glacier.B                       (adjoint.jl:47)            */
    ODINN_FIELD_H = 1,       /* state / ice thickness                                              */
    ODINN_FIELD_DH = 2,      /* last RHS output                                                    */
    ODINN_FIELD_LAMBDA = 3,  /* adjoint variable lambda                                            */
    ODINN_FIELD_VJP_H = 4,   /* last dH-VJP output                                                 */
    ODINN_FIELD_A = 5,       /* gridded creep coefficient on the dual grid, (nx-1) x (ny-1)        */
    ODINN_FIELD_VJP_A = 6,   /* dual-grid integrand  (Gamma_noA Hbar^{n+2} gradS^{n-1}) * D_adj    */
    ODINN_FIELD_H0 = 7,      /* initial condition of the forward solve                             */
    ODINN_FIELD_COUNT_ = 8
};

/* params.physical + the scalar iceflow-cache entries read on the path
 * (src/models/target/target_utils.jl:3-29, src/models/target/target_A.jl:22,
 *  test/params_construction.jl:24-34). */
typedef struct odinn_phys {
    double rho;   /* ice density                    */
    double g;     /* gravity                        */
    double eta0;  /* flux-clamp factor  (adjoint.jl:92) */
    double n;     /* Glen exponent                  */
    double p;     /* Weertman exponents             */
    double q;
    double C;     /* sliding coefficient            */
    double minA;  /* law output bounds (Laws.jl:337-338, target_utils.jl:109-113) */
    double maxA;
} odinn_phys;

/* ---------------------------------------------------------------------------------------- */
/* Lifetime / state behind the boundary                                                      */
/* ---------------------------------------------------------------------------------------- */

/* Build an ensemble of n_glaciers independent glaciers on CUDA device `device`.
 * Replaces: the per-glacier state captured by `simulation` -- glacier.B, dx, dy, nx, ny
 * (src/inverse/SIA2D/adjoint.jl:39-49) and init_cache (src/simulations/inversions/
 * inversion_utils.jl:482-483).  One glacier == one pmap task of the reference. */
int odinn_ensemble_create(int device, int dtype, int n_glaciers, const int* nx, const int* ny,
                          const double* dx, const double* dy, const odinn_phys* phys,
                          odinn_ensemble** out);
void odinn_ensemble_destroy(odinn_ensemble* e);

/* Last error message of `e` (or of the calling thread's last failed create when e == NULL). */
const char* odinn_last_error(const odinn_ensemble* e);

int odinn_n_glaciers(const odinn_ensemble* e);
int odinn_dtype_of(const odinn_ensemble* e);
/* Count of CUDA kernels launched by this handle so far (bench.py's gpu_launches). */
long long odinn_launch_count(const odinn_ensemble* e);
int odinn_synchronize(odinn_ensemble* e);
/* The handle's cudaStream_t (as void*): every kernel and copy of this handle is issued on it, so device-side
 * timing (CUDA events) must be recorded on this stream. */
void* odinn_stream(odinn_ensemble* e);

/* Copy one glacier's plane host <-> device.  Dual-grid fields are (nx-1) x (ny-1). */
int odinn_upload(odinn_ensemble* e, int glacier, int field, const void* host, int ld);
int odinn_download(odinn_ensemble* e, int glacier, int field, void* host, int ld);

/* cache.iceflow.A.value as a glacier-wide scalar (ScalarCache, src/laws/Cache.jl:23-97). */
int odinn_set_A_scalar(odinn_ensemble* e, int glacier, double A);
/* Long-term air temperature of the glacier, the input of LawA(nn, params) (iAvgScalarTemp, Laws.jl:326). */
int odinn_set_temperature(odinn_ensemble* e, int glacier, double T);
/* Switch the ensemble between scalar A (0) and the gridded A field ODINN_FIELD_A (1)
 * (MatrixCache; LawA(params; scalar=false), src/laws/Laws.jl:430-454). */
int odinn_set_A_mode(odinn_ensemble* e, int gridded);
int odinn_set_phys(odinn_ensemble* e, const odinn_phys* phys);

/* ---------------------------------------------------------------------------------------- */
/* Reference-facing per-call operators (host buffers in, host buffers out)                   */
/* ---------------------------------------------------------------------------------------- */

/* dH <- SIA2D(H).  Replaces SIA2D_UDE!(dH, H, container, t) -> Huginn.SIA2D!
 * (src/simulations/inversions/inversion_utils.jl:691-699).  Writes all nx*ny entries of dH
 * (border = 0); never writes H. */
int odinn_sia2d_rhs(odinn_ensemble* e, int glacier, const void* H, int ldH, void* dH, int lddH, double t);

/* out <- (dSIA/dH)^T lambda.  Replaces VJP_lambda_dSIAdH(::DiscreteVJP, lambda, H, theta, simulation, t)
 * (src/inverse/SIA2D/VJPs.jl:2-5 -> src/inverse/SIA2D/adjoint.jl:31-151). */
int odinn_sia2d_vjp_H(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                      void* out, int ldo, double t);

/* *out_S <- sum_ij (Gamma_noA Hbar^{n+2} gradS^{n-1})[i,j] * D_adj[i,j]  -- the glacier-wide
 * contraction of VJP_lambda_dSIAdtheta(::DiscreteVJP, ...) (VJPs.jl:30-33 -> adjoint.jl:178-255,
 * target_A.jl:64-92): d_theta = (dA/dtheta) * S.  With the gridded A mode the per-node integrand is
 * left in ODINN_FIELD_VJP_A instead and *out_S is its sum. */
int odinn_sia2d_vjp_theta(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                          double* out_S, double t);

/* Continuous-adjoint flavours (differentiate-then-discretise; no flux clamp, no H > 0 mask, border 0).
 * Replace VJP_lambda_dSIAdH(::ContinuousVJP, ...) (src/inverse/SIA2D/VJPs.jl:7-10 -> adjoint.jl:442-555) and
 * VJP_lambda_dSIAdtheta(::ContinuousVJP, ...) (VJPs.jl:35-38 -> adjoint.jl:582-662).  As for the discrete flavour the
 * theta-VJP of a glacier-wide law returns the scalar S with d_theta = (dA/dtheta) * S; with the gridded A mode the per-node
 * integrand is left in ODINN_FIELD_VJP_A, with a per-cell law the vector is read with odinn_law_cell_grad (adjoint.jl:646-657 is
 * the transpose of the discrete contraction of adjoint.jl:250, so these two cases run the discrete A2 kernels). */
int odinn_sia2d_vjp_H_continuous(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                                 void* out, int ldo, double t);
int odinn_sia2d_vjp_theta_continuous(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                                     double* out_S, double t);

/* ---------------------------------------------------------------------------------------- */
/* Ensemble (batched) operators on device-resident planes                                    */
/* ---------------------------------------------------------------------------------------- */

/* FIELD_DH <- SIA2D(FIELD_H) for every glacier in one launch. */
int odinn_rhs_resident(odinn_ensemble* e);
/* flags bit0: FIELD_VJP_H <- (dSIA/dH)^T FIELD_LAMBDA ; bit1: per-glacier S (and FIELD_VJP_A) ;
 * bit2: use the continuous flavour instead of the discrete one ;
 * bit3 (discrete flavour): also FIELD_DH <- SIA2D(FIELD_H) -- the (lambda_dfdH, dH) pair the reference's VJP returns
 * (src/inverse/SIA2D/VJPs.jl:12-28); the A1 pass recomputes every forward intermediate, so with bits 0|1|3 the fp32
 * path is ONE fused F1 + A1 + A2 kernel (5 words/cell instead of 3 + 4).
 * S_out may be NULL; otherwise n_glaciers doubles (device->host read inside the call). */
int odinn_vjp_resident(odinn_ensemble* e, int flags, double* S_out);

/* One pmap batch through host buffers (the reference maps SIA2D_grad_batch! over glacier batches,
 * src/inverse/SIA2D/gradient.jl:9-10, 235-246): for every glacier g upload H[g] (and lambda[g]), run
 * F1 + A1 + A2, download dH[g], vjpH[g], S[g].  Any of dH / vjpH / S may be NULL to skip that
 * output; lambda may be NULL when only dH is requested.  All matrices use ld = nx.
 * Glaciers are processed in chunks: the host->device copies of chunk c+1, the kernels of chunk c and the
 * device->host copies of chunk c-1 overlap on three streams.  The transfers are asynchronous only for pinned
 * (page-locked) host memory; see odinn_host_register.  The resident planes are not touched. */
int odinn_fwd_adj_batch_host(odinn_ensemble* e, const void* const* H, const void* const* lambda,
                             void* const* dH, void* const* vjpH, double* S);
/* Cells per pipeline chunk of odinn_fwd_adj_batch_host (default 8 Mi cells). */
int odinn_set_batch_chunk(odinn_ensemble* e, long long cells);
/* Page-lock / release a caller-owned host range (a Julia Array's memory) so that transfers from it are true
 * asynchronous DMA.  Thin wrappers of cudaHostRegister / cudaHostUnregister. */
int odinn_host_register(odinn_ensemble* e, void* host, size_t bytes);
int odinn_host_unregister(odinn_ensemble* e, void* host);

/* ---------------------------------------------------------------------------------------- */
/* Device-resident time loop and gradient (H, lambda and the snapshots never leave HBM)      */
/* ---------------------------------------------------------------------------------------- */

/* Forward solve of every glacier from ODINN_FIELD_H0 over the common time grid t[0..n_snap-1] with `nsub`
 * fixed sub-steps per interval; the state at every t[j] is kept as snapshot j.  Replaces
 * simulate_iceflow_UDE! -> solve(ODEProblem(SIA2D_UDE!, H0, tspan; tstops)) with saveat = tstops
 * (src/simulations/inversions/inversion_utils.jl:551-572, 584-610, tstops :487-495).  The integrator is a
 * user parameter in the reference (params.solver.solver); offered here: explicit Euler and SSPRK(3,3), each
 * stage fused into the RHS kernel.  The launches of one tstop interval are captured once into a CUDA graph and replayed per
 * interval (step sizes from a device table, so non-uniform tstops need no re-capture); ODINN_NO_GRAPH=1 disables it. */
int odinn_solve_forward(odinn_ensemble* e, int method, int n_snap, const double* t, int nsub);
/* Small ensembles (the reference's usual workload: a few glaciers of 100 - 400 px, src/simulations/inversions/inversion_utils.jl:551-610
 * called once per glacier) run odinn_solve_forward shared-memory resident: one thread-block cluster per glacier keeps the state on the
 * SMs and ONE launch runs every sub-step and stage of a whole range of tstop intervals, writing the snapshots as it goes.  mode = -1
 * (default): automatic (used when every glacier fits the shared memory of a cluster of <= 16 CTAs and all clusters are co-resident,
 * glacier-wide A, no per-cell law); 0: never (marching kernels + CUDA graph); 1, 2, 4, 8, 16: this cluster size.  Same results either
 * way up to rounding.  odinn_solve_forward_adaptive(ODINN_RDPK3SP35) takes the same path: the whole adaptive loop of a glacier (stages,
 * error norm reduced over the cluster, PID controller, tstops) then runs on the device without a host round trip per trial step. */
int odinn_set_cluster_mode(odinn_ensemble* e, int mode);
/* Adaptive forward solve with tstops: replaces solve(ODEProblem(SIA2D_UDE!, H0, tspan; tstops), solver;
 * reltol, abstol, maxiters, saveat = tstops) (src/simulations/inversions/inversion_utils.jl:559-568; solver choice
 * src/parameters ... AdjointTypes.jl:60, test/test_grad_loss.jl:143).  method = ODINN_BS3: Bogacki-Shampine 3(2) with FSAL,
 * OrdinaryDiffEq's error norm sqrt(mean((err / (abstol + reltol max(|u|, |u_new|)))^2)) and an I controller.  Every glacier
 * advances with its OWN adaptive step (independent ODEs, one pmap task each in the reference); the controller runs on the
 * device.  dt0 <= 0: (t[1] - t[0]) / 16.  max_steps bounds the number of ensemble-wide trial steps (maxiters).
 * steps_out / rejected_out (n_glaciers ints, optional): trial steps and rejected steps per glacier.
 * method = ODINN_RDPK3SP35: the reference's default integrator (OrdinaryDiffEq's RDPK3Sp35, src/inverse/AdjointTypes.jl:60,
 * test/test_grad_loss.jl:143): 5-stage 3rd-order 3S*+ low-storage scheme of Ranocha et al. (2021), FSAL, 5 RHS per trial step, PID
 * step-size controller beta = (0.64, -0.31, 0.04) with the limiter 1 + atan(x - 1) and acceptance threshold 0.81, OrdinaryDiffEq's
 * automatic initial step when dt0 <= 0 (for BS3 dt0 <= 0 means (t[1] - t[0]) / 16).
 * Large ensembles (and every ensemble odinn_set_cluster_mode keeps off the cluster path) run RDPK3Sp35 with ONE launch per stage: the
 * stage update -- and the error norm in the last stage -- is the epilogue of the F1 launch that evaluates the stage's slope; glaciers that
 * have already landed on the tstop are skipped by the launches; for ensembles of <= 8 Mi cells the whole trial step is replayed from a
 * CUDA graph.  ODINN_RK_NO_FUSE=1 selects the separate elementwise stage passes, ODINN_RK_GRAPH=0 / 1 overrides the graph choice
 * (developer switches for A/B measurements; same results up to rounding).
 * Snapshots are kept as by odinn_solve_forward. */
int odinn_solve_forward_adaptive(odinn_ensemble* e, int method, int n_snap, const double* t, double reltol, double abstol,
                                 double dt0, int max_steps, int* steps_out, int* rejected_out);
int odinn_get_snapshot(odinn_ensemble* e, int glacier, int j, void* host, int ld);
/* Provide snapshot j from the host instead (e.g. a solution saved by OrdinaryDiffEq). */
int odinn_set_snapshot(odinn_ensemble* e, int glacier, int j, int n_snap, const void* host, int ld);
/* Reference thickness H_ref(t_j) and loss weights W = is_in_glacier(H_ref, distance) / (nx*ny)
 * (src/losses/Losses.jl:270-291 with normalization = prod(N), src/inverse/SIA2D/gradient.jl:161). */
int odinn_set_reference(odinn_ensemble* e, int glacier, int j, int n_snap, const void* Href, const void* W, int ld);

/* loss_out[g] = sum_j (t_j - t_{j-1}) * sum_cells W (H_j - H_ref,j)^2 : LossH(L2Sum) of
 * loss_iceflow_transient (src/simulations/inversions/inversion_utils.jl:383-461). */
int odinn_loss(odinn_ensemble* e, const double* t, int n_t, double* loss_out);

/* The DiscreteAdjoint reverse loop of SIA2D_grad_batch! (src/inverse/SIA2D/gradient.jl:191-253) for
 * LossH(L2Sum), glacier-wide A:  loss_out[g] as above and
 * Ssum_out[g] = sum_j dt_{j-1} * S_{g,j},  so that  dL/dtheta = sum_g (dA_g/dtheta) * Ssum_out[g]. */
int odinn_grad_discrete(odinn_ensemble* e, const double* t, int n_t, double* loss_out, double* Ssum_out);

/* ---- mass-balance callback ---- */

/* n_mb mass-balance steps: step m fires when the forward solve reaches snapshot snapshot_index[m] (the end of its step_MB window,
 * PeriodicCallback of src/simulations/inversions/inversion_utils.jl:498-517) and its VJP is applied to lambda at the same tstop
 * in odinn_grad_discrete (VJP_lambda_dMBdH(::DiscreteVJP, ...), src/inverse/SIA2D/VJPs.jl:107-151; gradient.jl:201-207).
 * params[(m * n_glaciers + g) * 7 + k], k = temp, gradient, ref_hgt, snow, DDF, acc_factor, scale: the cumulative climate of the
 * window for glacier g (get_cumulative_climate!, independent of H) and the TImodel1 coefficients; scale = 1 / (step_MB * 12).
 *   PDD = temp + gradient (B + H - ref_hgt);  MB = (acc_factor snow - DDF max(PDD, 0)) scale;  MB_mask = (H > 0 and MB < 0) or
 *   (H > 10 and MB >= 0);  MB = 0 outside the mask, -H where H + MB < 0;  H += MB.
 * (The evaluation of TImodel1 lives in Muninn, which is not vendored: the formula is the one the in-tree VJP differentiates.)
 * n_mb = 0 switches the callback off. */
int odinn_set_mass_balance(odinn_ensemble* e, int n_mb, const int* snapshot_index, const double* params);
/* The MB field applied at step m (cache.iceflow.MB_history). */
int odinn_get_mass_balance(odinn_ensemble* e, int glacier, int m, void* host, int ld);

/* ---- surface velocity and LossV ---- */

/* (Vx, Vy) <- V_from_H(H): Vx = -D_up grad_x S, Vy = -D_up grad_y S on the dual grid, stored in inn1 of nx x ny matrices (last row /
 * column 0).  Replaces Huginn.V_from_H(simulation, H, t, theta) at its call sites src/losses/Losses.jl:314, 358, with
 * D_up = Velocity_up of src/models/target/target_A.jl:94-108. */
int odinn_surface_velocity(odinn_ensemble* e, int glacier, const void* H, int ldH, void* Vx, void* Vy, int ldV, double t);
/* out_dH <- VJP_lambda_dsurface_VdH(::DiscreteVJP, dVx, dVy, H, ...) (src/inverse/SIA2D/adjoint.jl:268-350) and
 * *out_S <- the glacier-wide contraction of VJP_lambda_dsurface_Vdtheta (adjoint.jl:352-413, target_A.jl:143-170):
 * d_theta = (dA/dtheta) * S.  Either output may be NULL. */
int odinn_sia2d_vjp_surface_V(odinn_ensemble* e, int glacier, const void* dVx, const void* dVy, int ldV, const void* H, int ldH,
                              void* out_dH, int ldo, double* out_S, double t);
/* Reference surface velocity at snapshot `snapshot_index` (a tstop that holds velocity data, tV_ref in gradient.jl:118-121), kept in
 * slot `slot` of `n_slots`: Vx_ref, Vy_ref, Vabs_ref and the weights Wv = (Vabs_ref > 0) / (nx ny [* sqrt(mean(Vx_ref^2 + Vy_ref^2
 * over the mask)) when scale_loss]) (Losses.jl:316, 327-331), nx x ny matrices. */
int odinn_set_velocity_reference(odinn_ensemble* e, int glacier, int slot, int n_slots, int snapshot_index, const void* Vx_ref,
                                 const void* Vy_ref, const void* Vabs_ref, const void* Wv, int ld);
/* Loss type through per-snapshot multipliers of the two L2Sum terms: loss = sum_j wH[j] * L2Sum_H(H_j) + wV[j] * L2Sum_V(H_j).
 * LossH: wH = dt_H, wV = 0 (the default when no weights are set); LossV: wV = dt_V; LossHV: wH = dt_H^2, wV = scaling dt_V^2
 * (Losses.jl:250-268, 293-336, 391-409).  v_component: 0 = :xy, 1 = :abs.  n_t = 0 restores the default.
 * Used by odinn_loss and odinn_grad_discrete. */
int odinn_set_loss_weights(odinn_ensemble* e, int n_t, const double* wH, const double* wV, int v_component);

/* The ContinuousAdjoint branch of SIA2D_grad_batch! (src/inverse/SIA2D/gradient.jl:276-538; the reference's default
 * gradient, src/parameters/UDEparameters.jl:63) for LossH(L2Sum): linear interpolation H_itp(t) of the snapshots, reverse
 * ODE d(lambda)/d(tau) = VJP_H(lambda, H_itp(-tau)) with the loss jumps at the tstops t[0..n_t-1], and
 * Ssum_out[g] = sum_m q_weights[m] * S_g(lambda(q_nodes[m]), H_itp(q_nodes[m])) over the quadrature nodes (GaussQuadrature,
 * gradient.jl:560-566: Gauss-Legendre nodes / weights mapped to the time span, computed by the caller), so that
 * dL/dtheta = sum_g (dA_g/dtheta) * Ssum_out[g]; with a per-cell law the gradient accumulates per glacier (odinn_law_cell_grad).
 * continuous_vjp selects the VJP flavour used inside (gradient.jl:310-314).  The reverse solver (params.UDE.grad.solver in the
 * reference) is method = ODINN_EULER | ODINN_SSPRK3 with nsub fixed sub-steps between consecutive stops (tstops and nodes). */
int odinn_grad_continuous(odinn_ensemble* e, const double* t, int n_t, int n_quadrature, const double* q_nodes,
                          const double* q_weights, int continuous_vjp, int method, int nsub, double* loss_out,
                          double* Ssum_out);

/* The same branch with the reverse ODE solved as the reference solves it by default (gradient.jl:449-467 with the defaults of
 * src/inverse/AdjointTypes.jl:53-66): adaptive RDPK3Sp35 + PID controller in tau = -t, every glacier with its own step,
 * tstops_adjoint = sort(unique(-reverse(tstops) U -q_nodes)), reltol, abstol (reference default 1e-8 each), dtmax (1/12), maxiters.
 * Callbacks as upstream: lambda_1 = effect_loss!(t_end); the mass-balance PeriodicCallback (initial_affect: also at t_end, not at t_0)
 * lambda += VJP_lambda_dMBdH(lambda, H - MB) BEFORE the loss jump of the same tstop (CallbackSet order, gradient.jl:407-426); the
 * loss jumps use the weights of odinn_set_loss_weights, thickness AND velocity terms (gradient.jl:326-366); the velocity term's
 * dl/dtheta is quadrature-weighted (odinn_set_velocity_quadrature).  steps_out (optional): trial steps per glacier.
 * With the discrete VJP flavour and a glacier-wide A (fp64: n = 3, C = 0) a stage of the reverse ODE is ONE launch: H_itp is formed from
 * the two bracketing snapshots as the rows are loaded, A1 runs on it, the RDPK3Sp35 stage update and the error norm are the epilogue
 * (likewise the A2 pass at a quadrature node); small ensembles run the whole reverse solve cluster-resident (odinn_set_cluster_mode). */
int odinn_grad_continuous_adaptive(odinn_ensemble* e, const double* t, int n_t, int n_quadrature, const double* q_nodes,
                                   const double* q_weights, int continuous_vjp, double reltol, double abstol, double dtmax,
                                   int max_steps, double* loss_out, double* Ssum_out, int* steps_out);
/* Continuous adjoint with a velocity loss: dL/dtheta += sum_m q_weights[m] * theta_scale * dl_V/dtheta(H_itp(t_m), V_ref_itp(t_m)),
 * the references interpolated linearly over the data times (one datum: constant) and Delta_t = (1, 1) inside the quadrature
 * (gradient.jl:289-301, 474-507): theta_scale = 1 for LossV, LossHV.scaling for LossHV, 0 switches the term off.
 * scale_loss: LossV.scale_loss (Losses.jl:327-331), needed to rebuild the weights of the interpolated references. */
int odinn_set_velocity_quadrature(odinn_ensemble* e, double theta_scale, int scale_loss);

/* ---------------------------------------------------------------------------------------- */
/* Laws                                                                                      */
/* ---------------------------------------------------------------------------------------- */

/* LawA(nn, params) f! (src/laws/Laws.jl:348-358): for every glacier A_g = minA + (maxA-minA)*NN([T_g]; theta)
 * with a Lux Dense chain (widths[0..n_layers], acts[0..n_layers-1]; theta = [vec(W1); b1; vec(W2); b2; ...],
 * W out x in column-major).  Also evaluates the pullback dA_g/dtheta (p_VJP!, Laws.jl:359-362 /
 * precompute_all_VJPs_laws!, inversion_utils.jl:647-686) and keeps it on the device.  A_out (G doubles) optional. */
int odinn_law_A_nn_apply(odinn_ensemble* e, int n_layers, const int* widths, const int* acts, const double* theta,
                         int n_theta, double* A_out);
/* dtheta[k] = sum_g dA_g/dtheta_k * S[g]  (aggregate of the per-glacier gradients, Model.jl:208-224).
 * S == NULL uses the sums left on the device by odinn_grad_discrete. */
int odinn_law_A_nn_pullback(odinn_ensemble* e, const double* S, double* dtheta, int n_theta);

/* Per-cell MLP laws evaluated at every dual-grid node.
 *   kind 1: LawU  (src/laws/Laws.jl:97-123)   U = post(NN(pre([Hbar, gradS]); theta)),  D = Hbar * U
 *                 (SIA2D_D_target, src/models/target/target_D_pure.jl:78-199)
 *   kind 2: LawY  (src/laws/Laws.jl:240-273)  Y = post(NN(pre([T, Hbar]); theta)),
 *                 D = S Hbar^{p-q+1} gradS^{p-1} + Y Gamma Hbar^{n_H+2} gradS^{n_gS-1}
 *                 (SIA2D_D_hybrid_target, src/models/target/target_D_hybrid.jl:22-208); T from odinn_set_temperature.
 * widths[0] must be 2 and widths[n_layers] 1 (width <= 32).  prescale_bounds = {lo0, hi0, lo1, hi1} of pre()
 * (target_utils.jl:131-141) or NULL for raw inputs; max_NN > 0 enables post() = max_NN exp((y-1)/y)
 * (target_utils.jl:86-93); n_H, n_gS <= 0 default to the Glen exponent n.
 * While a per-cell law is set every RHS / VJP entry point of the handle (per-call, resident, time loop, reverse loop)
 * uses it instead of the A law.  LawU: the partials dD/dHbar and dD/dgradS (central differences of the network upstream,
 * target_D_pure.jl:105-137) are the exact derivatives, propagated alongside the forward evaluation -- within the rounding noise
 * of the reference's differences (3e-10 / 7e-10 relative on the oracle).  LawY: the reference's one-sided difference
 * (target_D_hybrid.jl:58-73), evaluated in fp64. */
int odinn_law_cell_nn_set(odinn_ensemble* e, int kind, int n_layers, const int* widths, const int* acts, const double* theta,
                          int n_theta, const double* prescale_bounds, double max_NN, double n_H, double n_gS);
int odinn_law_cell_clear(odinn_ensemble* e);
/* interpolation = :Linear of the law pullback (the default of SIA2D_D_hybrid_target, src/models/target/target_D_hybrid.jl:13, 136-166;
 * optional for SIA2D_D_target, target_D_pure.jl:180-193 with the lattice of src/laws/Laws.jl:140-168): the network gradient is taken at
 * the knots -- knots0 x knots1 = (Hbar, gradS) for LawU, knots0 = Hbar for LawY (n1 = 0) -- and interpolated (bi)linearly per cell
 * (Interpolations.jl Gridded(Linear())).  The knots are what the reference's create_interpolation (target_utils.jl:245-299) returns:
 * computed by the caller (feed_input_cache!, src/laws/Cache.jl:130-154, from the forward solve for LawU; from the current Hbar for
 * LawY), 2 * n_interp_half values, strictly increasing; inputs outside the knot range are clamped to it.  On the device the contraction
 * with D_adj is reordered as a sum over KNOTS: the cells scatter D_adj * s to their bracketing knots, then the knots are back-propagated
 * -- cost independent of |theta| x cells.  n0 = 0 restores the exact per-node gradient (interpolation = :None). */
int odinn_law_cell_interp_set(odinn_ensemble* e, int n0, const double* knots0, int n1, const double* knots1);
/* out_theta[k] = sum_ij (dD/dtheta_k)[i,j] * D_adj[i,j]: VJP_lambda_dSIAdtheta(::DiscreteVJP, ...) for a per-cell law
 * (adjoint.jl:235-250 with dDiffusivity/dtheta of target_D_pure.jl:139-199 / target_D_hybrid.jl:98-166,
 * interpolation = :None, i.e. the exact per-node network gradient). */
int odinn_sia2d_vjp_theta_cell(odinn_ensemble* e, int glacier, const void* lambda, int ldl, const void* H, int ldH,
                               double* out_theta, int n_theta, double t);
/* Per-glacier theta-gradients [n_glaciers x n_theta] left on the device by odinn_vjp_resident(flags & 2) (last VJP) or
 * accumulated by odinn_grad_discrete (sum_j dt_{j-1} VJP_theta). */
int odinn_law_cell_grad(odinn_ensemble* e, double* out, int n_theta);

#ifdef __cplusplus
}
#endif
#endif /* ODINN_B200_H */
