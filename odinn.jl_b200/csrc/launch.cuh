// Kernel launchers of libodinn_b200.so, one translation unit per kernel family so that the host logic (capi.cu) and the
// families compile in parallel and an edit of one does not rebuild the others:
//   launch_march2.cu            fp32 two-column marching kernels (sia2d_march2.cuh) + the 2-D TMA ring variant (sia2d_tma.cuh)
//   launch_march_f32.cu / _f64  one-column marching kernels (sia2d_march.cuh), continuous VJPs (sia2d_cont.cuh), D-field mode
//   launch_law.cu               per-cell law node pass and theta pullback (sia2d_law.cuh)
#pragma once
#include "ensemble.cuh"

// Template dispatch on (n == 3 && C == 0, gridded A, eta0 == 1).  ODINN_BENCH_ONLY (developer builds for kernel
// tuning) instantiates the benchmark configuration only; every other configuration then fails loudly.
#ifdef ODINN_BENCH_ONLY
#define ODINN_DISPATCH(L2)                                                                            \
    do {                                                                                              \
        if (e->cubic && !e->a_gridded) L2(true, false);                                               \
        else return fail(e, ODINN_ESTATE, "this is an ODINN_BENCH_ONLY build (n = 3, C = 0, scalar A only)"); \
    } while (0)
#define ODINN_ETA(L3, CUB, AF)                                                                        \
    do {                                                                                              \
        if (eta1) L3(CUB, AF, true);                                                                  \
        else return fail(e, ODINN_ESTATE, "this is an ODINN_BENCH_ONLY build (eta0 = 1 only)");       \
    } while (0)
#else
#define ODINN_DISPATCH(L2)                                                                            \
    do {                                                                                              \
        if (e->cubic) {                                                                               \
            if (e->a_gridded) L2(true, true); else L2(true, false);                                   \
        } else {                                                                                      \
            if (e->a_gridded) L2(false, true); else L2(false, false);                                 \
        }                                                                                             \
    } while (0)
#define ODINN_ETA(L3, CUB, AF) do { if (eta1) L3(CUB, AF, true); else L3(CUB, AF, false); } while (0)
#endif

namespace odinn {

// F1 with a fused Runge-Kutta stage:  out = sa U0 + sb (Hin + sdt SIA2D(Hin))
struct Stage {
    const void* U0;
    double sa, sb, sdt;
    const double* tab = nullptr;   // graph replay: device table of stage coefficients (offset to this stage) + interval counter
    const int* interval = nullptr;
    const void* rk = nullptr;      // RkFuse<T>*: an RDPK3Sp35 stage as the epilogue instead (U0, sa, sb, sdt unused; whole ensemble only).
                                   // With RKF_NORM the per-item partial sums of the error norm land in d_partial (two-column items in fp32).
};

int alloc_plane(odinn_ensemble* e, void** p, size_t n_planes = 1);

// Launch with programmatic dependent launch (PDL) allowed: the kernel may become resident while the previous kernel of the stream is
// still draining, and runs its prologue (work-item and descriptor loads, index arithmetic: everything that does not depend on the
// previous kernel's output) until its `griddepcontrol.wait`, which returns when the previous kernel has completed and flushed.  ONLY
// for kernels that execute pdl_wait() before their first access to data a preceding kernel may have written (the F1 marching kernels).
// Measured on the mid-size ensembles (configs 3 / 4, forward SSPRK3 replayed from the CUDA graph) and at the bench workload
// (profiles/r02_pdl_ab.txt): 0 - 1 % -- the launches are not bound by their start-up chain -- so it is OFF unless ODINN_PDL=1.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// fp32, two columns per lane.  g0 < 0: whole ensemble.  `packed`: packed descriptor table / packed B of the host-batch path.
int launch_rhs2(odinn_ensemble* e, int g0, int g1, const void* Hin, void* out, const Stage* st, bool packed);
// *starts_used: the per-glacier start table (device) that indexes the partial sums this launch wrote
int launch_vjp2(odinn_ensemble* e, int g0, int g1, const void* lam, const void* H, void* out, bool wH, bool wS, bool packed,
                void* dH_out, const int** starts_used);
void tma_cache_free(odinn_ensemble* e);
// A1 + reverse time step in one pass (fp32 two-column kernel, whole ensemble, glacier-wide A):
//   lam_new = lam + dt (dSIA/dH)^T lam + cseed W (H - Href);   partial sums of  sum W (H - Href)^2  per work item -> d_partial / d_item2_start
int launch_vjp2_seed(odinn_ensemble* e, const void* lam, const void* H, const void* Href, const void* W, void* lam_new, double dt, double cseed,
                     const int** starts_used);

// one RDPK3Sp35 stage of the continuous adjoint's reverse ODE in one pass: lerp(Ha, Hb) on load, A1, stage update as the epilogue
// (s_only: the A2 pass at a quadrature node with the same interpolation on load; per-item partial sums of S -> d_partial)
int launch_vjp2_rk(odinn_ensemble* e, const void* S1in, const void* Ha, const void* Hb, void* S1out, const void* rkfuse, double c, double sign,
                   double ta, double tb, bool s_only = false);
template <typename T> int launch_vjp_rk_t(odinn_ensemble* e, const void* S1in, const void* Ha, const void* Hb, void* S1out, const void* rkfuse,
                                          double c, double sign, double ta, double tb, bool s_only = false);

// one column per lane (fp32 generation 1 and fp64); items [i0, i0 + n_items)
template <typename T> int launch_rhs_t(odinn_ensemble* e, int i0, int n_items, const void* Hin, void* out, const Stage* st, bool packed);
template <typename T> int launch_rhs_law_t(odinn_ensemble* e, int i0, int n_items, const void* Hin, void* out, const Stage* st);
template <typename T> int launch_vjp_law_t(odinn_ensemble* e, int i0, int n_items, const void* lam, const void* H, void* out, bool wH, bool wS);
template <typename T> int launch_vjp_t(odinn_ensemble* e, int i0, int n_items, const void* lam, const void* H, void* out, bool wH, bool wS,
                                       bool packed, void* dH_out = nullptr);
template <typename T> int launch_vjpc_t(odinn_ensemble* e, int i0, int n_items, const void* lam, const void* H, void* out);
template <typename T> int launch_unitA_dot_t(odinn_ensemble* e, int g, const void* lam, const void* H, double* S_dst, double scale, int accumulate);

// per-cell laws: node pass (D, and alpha / beta when partials) and theta pullback for glaciers [g0, g1) (g0 < 0: all)
int launch_law_nodes(odinn_ensemble* e, int g0, int g1, const void* H, bool partials);
int launch_law_theta(odinn_ensemble* e, int g0, int g1, const void* H, double scale = 1.0, int accumulate = 0);

}  // namespace odinn
