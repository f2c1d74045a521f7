// Continuous-adjoint gradient on the device: the ContinuousAdjoint branch of SIA2D_grad_batch!
// (src/inverse/SIA2D/gradient.jl:276-538), the reference's DEFAULT gradient method (src/parameters/UDEparameters.jl:63).
//
//   H_itp(t)      linear interpolation of the forward snapshots over the tstops                    (gradient.jl:285-301)
//   λ(t_end)      = ∂ℓ/∂H(t_end)                      (effect_loss! applied by hand)              (:439-446)
//   dλ/dτ         = VJP_H(λ, H_itp(-τ)), τ = -t;  λ += ∂ℓ/∂H at every tstop (DiscreteCallback)   (:316-366, 449-470)
//   dL/dθ         = Σ_m w_m VJP_θ(λ(t_m), H_itp(t_m)) over the Gauss-Legendre nodes               (:305-306, 495-507)
//
// The reverse-ODE solver is a user parameter of the reference (params.UDE.grad.solver); here a fixed-step scheme (explicit
// Euler or SSPRK(3,3)) with `nsub` sub-steps between consecutive stops (tstops ∪ quadrature nodes), exactly the scheme of
// oracle/sia2d_numpy.py::loss_and_grad_continuous.  Either VJP flavour can be used inside (gradient.jl:310-314).
// λ, H_itp and the stage planes never leave HBM; the host only sequences kernels.
#include <vector>

#include "ensemble.cuh"

namespace odinn {

// Elementwise kernels over the padded planes of every glacier (grid: chunks x glaciers, 16-byte vector accesses; padding stays zero
// under these linear combinations) -- the same scheme as the adaptive solve's kernels.
template <typename T> struct CaVec;
template <> struct CaVec<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct CaVec<double> { typedef double2 type; static constexpr int N = 2; };
constexpr int CA_NT = 256;
constexpr int CA_UNROLL = 4;
__device__ __forceinline__ float4 ca_axpby(float a, const float4& x, float b, const float4& y) {
    return make_float4(a * x.x + b * y.x, a * x.y + b * y.y, a * x.z + b * y.z, a * x.w + b * y.w);
}
__device__ __forceinline__ double2 ca_axpby(double a, const double2& x, double b, const double2& y) {
    return make_double2(a * x.x + b * y.x, a * x.y + b * y.y);
}
__device__ __forceinline__ float4 ca_scale(float a, const float4& x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }
__device__ __forceinline__ double2 ca_scale(double a, const double2& x) { return make_double2(a * x.x, a * x.y); }

#define CA_VEC_LOOP(BODY)                                                                           \
    typedef typename CaVec<T>::type V;                                                              \
    const GDesc<T> d = descs[blockIdx.y];                                                           \
    const long long nvec = (long long)d.ld * d.ny / CaVec<T>::N, base = d.off / CaVec<T>::N;        \
    _Pragma("unroll") for (int u = 0; u < CA_UNROLL; ++u) {                                         \
        const long long q = ((long long)blockIdx.x * CA_UNROLL + u) * CA_NT + threadIdx.x;          \
        if (q < nvec) { const long long p = base + q; BODY }                                        \
    }

// Ht = (1 - a) Ha + a Hb
template <typename T>
__global__ void __launch_bounds__(CA_NT)
ca_lerp(const GDesc<T>* __restrict__ descs, const T* __restrict__ Ha, const T* __restrict__ Hb, T* __restrict__ Ht, T a) {
    CA_VEC_LOOP({ reinterpret_cast<V*>(Ht)[p] = ca_axpby(T(1) - a, reinterpret_cast<const V*>(Ha)[p], a, reinterpret_cast<const V*>(Hb)[p]); })
}

// out = sa U0 + sb (U + h V)      (one Shu-Osher stage of the reverse solve; out may alias U0 or U)
template <typename T>
__global__ void __launch_bounds__(CA_NT)
ca_stage(const GDesc<T>* __restrict__ descs, const T* U0, const T* U, const T* __restrict__ Vp, T* out, T sa, T sb, T h) {
    CA_VEC_LOOP({
        const V u = ca_axpby(T(1), reinterpret_cast<const V*>(U)[p], h, reinterpret_cast<const V*>(Vp)[p]);
        reinterpret_cast<V*>(out)[p] = (sa != T(0)) ? ca_axpby(sa, reinterpret_cast<const V*>(U0)[p], sb, u) : ca_scale(sb, u);
    })
}

template <typename T>
static int grad_continuous_t(odinn_ensemble* e, const double* t, int n_t, int n_q, const double* qn, const double* qw,
                             bool cont_vjp, int method, int nsub) {
    int rc;
    void** Htp = &e->ext_dev[EXT_CA_HT];
    void** U1p = &e->ext_dev[EXT_CA_LAM1];
    void** U2p = &e->ext_dev[EXT_CA_LAM2];
    if ((rc = alloc_work_plane(e, Htp)) || (rc = alloc_work_plane(e, U1p)) || (rc = alloc_work_plane(e, U2p))) return rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_LAMBDA)) || (rc = ensure_plane(e, ODINN_FIELD_VJP_H))) return rc;
    if ((rc = sync_descs(e))) return rc;
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    long long max_vec = 0;
    for (int g = 0; g < e->G; ++g) max_vec = std::max(max_vec, (long long)e->gl[g].ld * e->gl[g].ny / CaVec<T>::N);
    const dim3 egrid((unsigned)((max_vec + (long long)CA_NT * CA_UNROLL - 1) / ((long long)CA_NT * CA_UNROLL)), e->G);
    T* lam = (T*)e->plane[ODINN_FIELD_LAMBDA];
    T* V = (T*)e->plane[ODINN_FIELD_VJP_H];
    T *Ht = (T*)*Htp, *U1 = (T*)*U1p, *U2 = (T*)*U2p;
    const size_t pbytes = (size_t)e->total * e->esize;
    ODINN_CUDA(e, cudaMemsetAsync(lam, 0, pbytes, e->stream));
    ODINN_CUDA(e, cudaMemsetAsync(e->d_loss, 0, sizeof(double) * e->G, e->stream));
    ODINN_CUDA(e, cudaMemsetAsync(e->d_Ssum, 0, sizeof(double) * e->G, e->stream));
    if (e->law_kind != 0)
        ODINN_CUDA(e, cudaMemsetAsync(e->d_law_dtheta, 0, sizeof(double) * (size_t)e->G * e->law_n_theta, e->stream));

    auto H_itp = [&](double tt) -> int {  // Ht <- H_itp(tt)
        int j = 0;
        while (j + 2 < n_t && tt >= t[j + 1]) ++j;  // interval [t_j, t_{j+1}] with tt >= t_j (clipped to the last one)
        const double a = (tt - t[j]) / (t[j + 1] - t[j]);
        ca_lerp<T><<<egrid, CA_NT, 0, e->stream>>>(descs, (const T*)snapshot_ptr(e, j), (const T*)snapshot_ptr(e, j + 1), Ht, (T)a);
        ODINN_CHECK_LAUNCH(e);
        return ODINN_OK;
    };
    auto f = [&](double tt, const T* u) -> int {  // V <- VJP_H(u, H_itp(tt))
        int r = H_itp(tt);
        if (r) return r;
        return vjp_planes(e, u, Ht, V, true, false, nullptr, 1.0, 0, cont_vjp);
    };
    auto stage = [&](const T* U0, const T* U, T* out, double sa, double sb, double h) -> int {
        ca_stage<T><<<egrid, CA_NT, 0, e->stream>>>(descs, U0, U, V, out, (T)sa, (T)sb, (T)h);
        ODINN_CHECK_LAUNCH(e);
        return ODINN_OK;
    };

    // stops in descending time; at equal times the quadrature sample comes before the loss jump
    struct Ev { double t; int is_q; int idx; };
    std::vector<Ev> ev;
    for (int j = 0; j < n_t; ++j) ev.push_back({t[j], 0, j});
    for (int m = 0; m < n_q; ++m) ev.push_back({qn[m], 1, m});
    std::stable_sort(ev.begin(), ev.end(), [](const Ev& a, const Ev& b) { return a.t != b.t ? a.t > b.t : a.is_q > b.is_q; });

    bool started = false;
    double t_cur = 0.0;
    for (const Ev& s : ev) {
        if (started && s.t < t_cur) {
            const double h = (t_cur - s.t) / nsub;
            for (int k = 0; k < nsub; ++k) {
                const double ta = t_cur - k * h;
                if (method == ODINN_EULER) {
                    if ((rc = f(ta, lam)) || (rc = stage(lam, lam, lam, 0.0, 1.0, h))) return rc;
                } else {
                    if ((rc = f(ta, lam)) || (rc = stage(lam, lam, U1, 0.0, 1.0, h))) return rc;
                    if ((rc = f(ta - h, U1)) || (rc = stage(lam, U1, U2, 0.75, 0.25, h))) return rc;
                    if ((rc = f(ta - 0.5 * h, U2)) || (rc = stage(lam, U2, lam, 1.0 / 3.0, 2.0 / 3.0, h))) return rc;
                }
            }
        }
        started = true;
        t_cur = s.t;
        if (!s.is_q) {
            const int j = s.idx;
            // w_j = Δt_HV.H[ind-1] through safe_slice for LossH: 0 for the first data point (odinn_set_loss_weights overrides)
            const double wH = loss_weight_H(e, t, n_t, j), wV = loss_weight_V(e, n_t, j);
            // Callback order of the reference: interior tstops run CallbackSet(cb_adjoint_MB, cb_adjoint_loss) -- the mass-balance VJP
            // first (not at t_0: final_affect = false), then the loss jump; at t_end the loss is applied by hand to λ₁ BEFORE the solve
            // starts and the PeriodicCallback's initial_affect adds the MB term afterwards (gradient.jl:407-446).
            const bool at_end = (j == n_t - 1);
            if (!at_end && j != 0 && (rc = mb_adjoint_step(e, j, lam, snapshot_ptr(e, j)))) return rc;
            // ℓ += w_j Σ W (H_j - H_ref,j)² ;  λ += 2 w_j W (H_j - H_ref,j)       (Losses.jl:270-291)
            if (wH != 0.0 && (rc = loss_seed_planes(e, snapshot_ptr(e, j), (char*)e->href + (size_t)j * pbytes, (char*)e->wmask + (size_t)j * pbytes,
                                                    lam, nullptr, lam, 0.0, 2.0 * wH, e->d_loss, wH, 1)))
                return rc;
            // velocity term: ℓ and ∂ℓ/∂H at the tstop (its ∂ℓ/∂θ is quadrature-weighted, below)   (Losses.jl:293-390, gradient.jl:326-366)
            if ((rc = velocity_loss_term(e, j, snapshot_ptr(e, j), lam, wV, e->d_loss, nullptr))) return rc;
            if (at_end && (rc = mb_adjoint_step(e, j, lam, snapshot_ptr(e, j)))) return rc;
        } else {
            if ((rc = H_itp(s.t))) return rc;
            if ((rc = vjp_planes(e, lam, Ht, nullptr, false, true, e->d_Ssum, qw[s.idx], 1, cont_vjp))) return rc;
            // + w_m ∂ℓ/∂θ(t_m): velocity references interpolated at the node, Δt = (1, 1)   (gradient.jl:474-507)
            if (e->lossV_theta_scale != 0.0 &&
                (rc = velocity_theta_term_interp(e, s.t, t, n_t, Ht, e->lossV_theta_scale * qw[s.idx], e->d_Ssum)))
                return rc;
        }
    }
    return ODINN_OK;
}

}  // namespace odinn

using namespace odinn;

extern "C" int odinn_grad_continuous(odinn_ensemble* e, const double* t, int n_t, int n_quadrature, const double* q_nodes,
                                     const double* q_weights, int continuous_vjp, int method, int nsub, double* loss_out,
                                     double* Ssum_out) {
    if (!e) return fail(nullptr, ODINN_EARG, "null ensemble");
    {
        cudaError_t s_ = cudaSetDevice(e->device);
        if (s_ != cudaSuccess) return fail(e, ODINN_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(s_));
    }
    if (!e->snap || !e->href) return fail(e, ODINN_ESTATE, "snapshots and reference data must be set first");
    if (e->n_snap != e->n_ref || n_t != e->n_snap) return fail(e, ODINN_ESTATE, "snapshot / reference / time counts differ");
    if (!t || n_t < 2 || n_quadrature < 1 || !q_nodes || !q_weights || nsub < 1) return fail(e, ODINN_EARG, "bad continuous-adjoint arguments");
    if (method != ODINN_EULER && method != ODINN_SSPRK3) return fail(e, ODINN_EARG, "reverse solve: method must be ODINN_EULER or ODINN_SSPRK3");
    for (int m = 0; m < n_quadrature; ++m)
        if (!(q_nodes[m] >= t[0] && q_nodes[m] <= t[n_t - 1])) return fail(e, ODINN_EARG, "quadrature node outside the time span");
    if (e->a_gridded) return fail(e, ODINN_ESTATE, "odinn_grad_continuous supports glacier-wide A and per-cell laws");
    if (continuous_vjp && e->law_kind != 0) return fail(e, ODINN_ESTATE, "the continuous VJP flavour is provided for glacier-wide A laws");
    int rc = e->dtype == ODINN_F32 ? grad_continuous_t<float>(e, t, n_t, n_quadrature, q_nodes, q_weights, continuous_vjp != 0, method, nsub)
                                   : grad_continuous_t<double>(e, t, n_t, n_quadrature, q_nodes, q_weights, continuous_vjp != 0, method, nsub);
    if (rc) return rc;
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_loss, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S + e->G, e->d_Ssum, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    if (loss_out) memcpy(loss_out, e->h_S, sizeof(double) * e->G);
    if (Ssum_out) memcpy(Ssum_out, e->h_S + e->G, sizeof(double) * e->G);
    return ODINN_OK;
}
