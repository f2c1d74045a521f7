// Adaptive forward solve on the device: Bogacki-Shampine 3(2) with FSAL, per-glacier step-size control, tstops.
//
// Replaces: solve(ODEProblem(SIA2D_UDE!, H0, tspan; tstops), solver; reltol, abstol, saveat = tstops)
// of simulate_iceflow_UDE! (src/simulations/inversions/inversion_utils.jl:551-610; tstops assembly :487-495).  The
// integrator is a user parameter of the reference (params.solver.solver, default RDPK3Sp35 -- OrdinaryDiffEq, not in
// tree); the embedded pair offered here is BS3 with OrdinaryDiffEq's error norm  sqrt(mean((err / (abstol + reltol
// max(|u|, |u_new|)))^2))  and an I controller (safety 0.9, growth limits 0.2 .. 5), step by step the scheme of
// oracle/sia2d_numpy.py::solve_forward(method="bs3").
//
// Every glacier of the ensemble is an independent ODE (one pmap task in the reference), so every glacier carries its OWN
// (t, dt) on the device: the elementwise kernels read the step of their glacier from a state table, the accept / reject
// decision is taken by a controller kernel, and the host only reads back "how many glaciers are still inside the
// interval" (one int per step).  Glaciers that reached the tstop idle with h = 0 until the slowest one arrives.
#include "ensemble.cuh"

namespace odinn {

struct AdState {
    double t, b, dt, h, en;
    int last, truncated, accept, done;
    int steps, rejected;
};

__device__ __forceinline__ void ad_plan_step(AdState& s) {  // oracle: last = dt >= b - t; h = last ? b - t : dt
    s.last = s.dt >= (s.b - s.t);
    s.h = s.last ? (s.b - s.t) : s.dt;
    s.truncated = s.h < s.dt;
}

// New interval (a, b]: every glacier restarts from t = a with the step it carried over (dt0 on the very first interval).
__global__ void ad_begin_interval(AdState* st, int G, double a, double b, double dt0, int first) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    AdState s = st[g];
    if (first) {
        s.dt = dt0 > 0.0 ? dt0 : (b - a) / 16.0;
        s.steps = s.rejected = 0;
    }
    s.t = a;
    s.b = b;
    s.done = 0;
    s.accept = 0;
    ad_plan_step(s);
    st[g] = s;
}

// Elementwise kernels over the padded planes of every glacier: grid (chunks, glaciers), 16-byte vector accesses, four vectors in
// flight per thread and trip.  Rows are padded to 32 elements and every plane keeps its padding at zero, which these linear
// combinations preserve (and a zero error contribution), so the kernels need no (i, j) decomposition.  (The first version walked
// 32 x 16 tiles with scalar accesses: 1.6 ms of the 2.0 ms ensemble step at 500 x 500 x 256.)
template <typename T> struct AdVec;
template <> struct AdVec<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct AdVec<double> { typedef double2 type; static constexpr int N = 2; };
constexpr int AD_NT = 256;
constexpr int AD_UNROLL = 4;

template <typename T> __device__ __forceinline__ void ad_unpack(const typename AdVec<T>::type& v, T* x);
template <> __device__ __forceinline__ void ad_unpack<float>(const float4& v, float* x) { x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
template <> __device__ __forceinline__ void ad_unpack<double>(const double2& v, double* x) { x[0] = v.x; x[1] = v.y; }
template <typename T> __device__ __forceinline__ typename AdVec<T>::type ad_pack(const T* x);
template <> __device__ __forceinline__ float4 ad_pack<float>(const float* x) { return make_float4(x[0], x[1], x[2], x[3]); }
template <> __device__ __forceinline__ double2 ad_pack<double>(const double* x) { return make_double2(x[0], x[1]); }

#define AD_VEC_PROLOGUE                                                                             \
    typedef typename AdVec<T>::type V;                                                              \
    constexpr int N = AdVec<T>::N;                                                                  \
    const GDesc<T> d = descs[blockIdx.y];                                                           \
    const AdState s = st[blockIdx.y];                                                               \
    const long long nvec = (long long)d.ld * d.ny / N;                                              \
    const long long v0 = ((long long)blockIdx.x * AD_UNROLL) * AD_NT + threadIdx.x;                 \
    const long long base = d.off / N;
#define AD_LD(P, q) (reinterpret_cast<const V*>(P)[base + (q)])

// Y = H + c h_g K
template <typename T>
__global__ void __launch_bounds__(AD_NT)
ad_stage_input(const GDesc<T>* __restrict__ descs, const AdState* __restrict__ st, const T* __restrict__ H, const T* __restrict__ K,
               T* __restrict__ Y, double c) {
    AD_VEC_PROLOGUE
    const T ch = (T)(c * s.h);
#pragma unroll
    for (int u = 0; u < AD_UNROLL; ++u) {
        const long long q = v0 + (long long)u * AD_NT;
        if (q < nvec) {
            T h[N], k[N], y[N];
            ad_unpack<T>(AD_LD(H, q), h);
            ad_unpack<T>(AD_LD(K, q), k);
#pragma unroll
            for (int e = 0; e < N; ++e) y[e] = h[e] + ch * k[e];
            reinterpret_cast<V*>(Y)[base + q] = ad_pack<T>(y);
        }
    }
}

// Hn = H + h_g (2/9 k1 + 1/3 k2 + 4/9 k3)
template <typename T>
__global__ void __launch_bounds__(AD_NT)
ad_bs3_solution(const GDesc<T>* __restrict__ descs, const AdState* __restrict__ st, const T* __restrict__ H, const T* __restrict__ k1,
                const T* __restrict__ k2, const T* __restrict__ k3, T* __restrict__ Hn) {
    AD_VEC_PROLOGUE
    const T h = (T)s.h;
#pragma unroll
    for (int u = 0; u < AD_UNROLL; ++u) {
        const long long q = v0 + (long long)u * AD_NT;
        if (q < nvec) {
            T a[N], x1[N], x2[N], x3[N], y[N];
            ad_unpack<T>(AD_LD(H, q), a);
            ad_unpack<T>(AD_LD(k1, q), x1);
            ad_unpack<T>(AD_LD(k2, q), x2);
            ad_unpack<T>(AD_LD(k3, q), x3);
#pragma unroll
            for (int e = 0; e < N; ++e) y[e] = a[e] + h * (T(2.0 / 9.0) * x1[e] + T(1.0 / 3.0) * x2[e] + T(4.0 / 9.0) * x3[e]);
            reinterpret_cast<V*>(Hn)[base + q] = ad_pack<T>(y);
        }
    }
}

// partial[glacier * gridDim.x + chunk] = Σ (err / sc)²,  err = h (-5/72 k1 + 1/12 k2 + 1/9 k3 - 1/8 k4),  sc = abstol + reltol max(|H|, |Hn|)
template <typename T>
__global__ void __launch_bounds__(AD_NT)
ad_bs3_error(const GDesc<T>* __restrict__ descs, const AdState* __restrict__ st, const T* __restrict__ H, const T* __restrict__ Hn,
             const T* __restrict__ k1, const T* __restrict__ k2, const T* __restrict__ k3, const T* __restrict__ k4,
             double* __restrict__ partial, double reltol, double abstol) {
    __shared__ double sRed[AD_NT / 32];
    AD_VEC_PROLOGUE
    const T h = (T)s.h;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < AD_UNROLL; ++u) {
        const long long q = v0 + (long long)u * AD_NT;
        if (q < nvec) {
            T a[N], b[N], x1[N], x2[N], x3[N], x4[N];
            ad_unpack<T>(AD_LD(H, q), a);
            ad_unpack<T>(AD_LD(Hn, q), b);
            ad_unpack<T>(AD_LD(k1, q), x1);
            ad_unpack<T>(AD_LD(k2, q), x2);
            ad_unpack<T>(AD_LD(k3, q), x3);
            ad_unpack<T>(AD_LD(k4, q), x4);
#pragma unroll
            for (int e = 0; e < N; ++e) {
                const T err = h * (T(-5.0 / 72.0) * x1[e] + T(1.0 / 12.0) * x2[e] + T(1.0 / 9.0) * x3[e] - T(1.0 / 8.0) * x4[e]);
                const double sc = abstol + reltol * fmax(fabs((double)a[e]), fabs((double)b[e]));
                const double r = (double)err / sc;
                acc += r * r;
            }
        }
    }
    double sum = block_sum(acc, sRed);
    if (threadIdx.x == 0) partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = sum;
}

// sumsq[g] = Σ_chunks partial[g][chunk]   (fixed order)
__global__ void __launch_bounds__(AD_NT)
ad_reduce_chunks(const double* __restrict__ partial, int nchunk, double* __restrict__ sumsq) {
    __shared__ double sRed[AD_NT / 32];
    double acc = 0.0;
    for (int c = threadIdx.x; c < nchunk; c += AD_NT) acc += partial[(long long)blockIdx.x * nchunk + c];
    double sum = block_sum(acc, sRed);
    if (threadIdx.x == 0) sumsq[blockIdx.x] = sum;
}

// One thread per glacier: error norm -> accept / reject -> next step (the `while t < b` body of the oracle).
__global__ void ad_control(AdState* st, const double* __restrict__ sumsq, const int* __restrict__ nx, const int* __restrict__ ny,
                           int G, int* n_active) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    AdState s = st[g];
    if (s.done) {
        s.accept = 0;
        st[g] = s;
        return;
    }
    const double en = sqrt(sumsq[g] / ((double)nx[g] * (double)ny[g]));
    double fac = 0.9 * pow(1.0 / fmax(en, 1e-10), 1.0 / 3.0);
    fac = fmin(5.0, fmax(0.2, fac));
    s.en = en;
    if (en <= 1.0) {
        s.accept = 1;
        s.t = s.last ? s.b : s.t + s.h;
        // a step shortened only to land on the tstop must not shrink the proposal
        s.dt = (s.truncated && fac >= 1.0) ? fmax(s.h * fac, s.dt) : s.h * fac;
    } else {
        s.accept = 0;
        s.dt = s.h * fac;
        s.rejected++;
        atomicAdd(n_active + 1, 1);  // rejections of this trial step
    }
    s.steps++;
    if (s.t < s.b) {
        ad_plan_step(s);
        atomicAdd(n_active, 1);
    } else {
        s.done = 1;
        s.h = 0.0;
    }
    st[g] = s;
}

// accepted glaciers: H <- Hn, k1 <- k4 (FSAL)
template <typename T>
__global__ void __launch_bounds__(AD_NT)
ad_commit(const GDesc<T>* __restrict__ descs, const AdState* __restrict__ st, T* __restrict__ H, const T* __restrict__ Hn,
          T* __restrict__ k1, const T* __restrict__ k4) {
    AD_VEC_PROLOGUE
    if (!s.accept) return;
#pragma unroll
    for (int u = 0; u < AD_UNROLL; ++u) {
        const long long q = v0 + (long long)u * AD_NT;
        if (q < nvec) {
            reinterpret_cast<V*>(H)[base + q] = AD_LD(Hn, q);
            reinterpret_cast<V*>(k1)[base + q] = AD_LD(k4, q);
        }
    }
}

template <typename T>
static int solve_bs3_t(odinn_ensemble* e, int n_snap, const double* t, double reltol, double abstol, double dt0,
                       int max_steps, int* steps_out, int* rejected_out) {
    int rc;
    for (int k = 0; k < 6; ++k)
        if ((rc = alloc_work_plane(e, &e->ad_plane[k]))) return rc;
    if (!e->d_ad_state) {
        ODINN_CUDA(e, cudaMalloc(&e->d_ad_state, sizeof(AdState) * e->G));
        ODINN_CUDA(e, cudaMalloc(&e->d_ad_dims, sizeof(int) * 2 * e->G + 2 * sizeof(int)));
        ODINN_CUDA(e, cudaMallocHost(&e->h_ad_active, 2 * sizeof(int)));
        std::vector<int> dims(2 * e->G);
        for (int g = 0; g < e->G; ++g) { dims[g] = e->gl[g].nx; dims[e->G + g] = e->gl[g].ny; }
        ODINN_CUDA(e, cudaMemcpy(e->d_ad_dims, dims.data(), sizeof(int) * 2 * e->G, cudaMemcpyHostToDevice));
    }
    ODINN_CUDA(e, cudaMemsetAsync(e->d_ad_state, 0, sizeof(AdState) * e->G, e->stream));
    AdState* st = (AdState*)e->d_ad_state;
    const int* d_nx = e->d_ad_dims;
    const int* d_ny = e->d_ad_dims + e->G;
    int* d_active = e->d_ad_dims + 2 * e->G;
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const size_t pbytes = (size_t)e->total * e->esize;
    T* H = (T*)e->plane[ODINN_FIELD_H];
    T *k1 = (T*)e->ad_plane[0], *k2 = (T*)e->ad_plane[1], *k3 = (T*)e->ad_plane[2], *k4 = (T*)e->ad_plane[3];
    T *Hn = (T*)e->ad_plane[4], *Y = (T*)e->ad_plane[5];
    const int gb = (e->G + 127) / 128;
    long long max_vec = 0;
    for (int g = 0; g < e->G; ++g) max_vec = std::max(max_vec, (long long)e->gl[g].ld * e->gl[g].ny / AdVec<T>::N);
    const int nchunk = (int)((max_vec + (long long)AD_NT * AD_UNROLL - 1) / ((long long)AD_NT * AD_UNROLL));
    const dim3 egrid(nchunk, e->G);
    if (e->ext_int[0] < nchunk * e->G) {  // per-(glacier, chunk) partial sums of the error norm
        if (e->ext_dev[EXT_AD_PARTIAL]) cudaFree(e->ext_dev[EXT_AD_PARTIAL]);
        e->ext_dev[EXT_AD_PARTIAL] = nullptr;
        ODINN_CUDA(e, cudaMalloc(&e->ext_dev[EXT_AD_PARTIAL], sizeof(double) * (size_t)nchunk * e->G));
        e->ext_int[0] = nchunk * e->G;
    }
    double* ad_partial = (double*)e->ext_dev[EXT_AD_PARTIAL];

    ODINN_CUDA(e, cudaMemcpyAsync(H, e->plane[ODINN_FIELD_H0], pbytes, cudaMemcpyDeviceToDevice, e->stream));
    ODINN_CUDA(e, cudaMemcpyAsync(snapshot_ptr(e, 0), H, pbytes, cudaMemcpyDeviceToDevice, e->stream));
    if ((rc = rhs_planes(e, H, k1))) return rc;  // FSAL seed
    int total_steps = 0;
    for (int j = 1; j < n_snap; ++j) {
        ad_begin_interval<<<gb, 128, 0, e->stream>>>(st, e->G, t[j - 1], t[j], dt0, j == 1);
        ODINN_CHECK_LAUNCH(e);
        for (;;) {
            if (++total_steps > max_steps) return fail(e, ODINN_ESTATE, "bs3: too many steps (maxiters)");
            ad_stage_input<T><<<egrid, AD_NT, 0, e->stream>>>(descs, st, H, k1, Y, 0.5);
            ODINN_CHECK_LAUNCH(e);
            if ((rc = rhs_planes(e, Y, k2))) return rc;
            ad_stage_input<T><<<egrid, AD_NT, 0, e->stream>>>(descs, st, H, k2, Y, 0.75);
            ODINN_CHECK_LAUNCH(e);
            if ((rc = rhs_planes(e, Y, k3))) return rc;
            ad_bs3_solution<T><<<egrid, AD_NT, 0, e->stream>>>(descs, st, H, k1, k2, k3, Hn);
            ODINN_CHECK_LAUNCH(e);
            if ((rc = rhs_planes(e, Hn, k4))) return rc;
            ad_bs3_error<T><<<egrid, AD_NT, 0, e->stream>>>(descs, st, H, Hn, k1, k2, k3, k4, ad_partial, reltol, abstol);
            ODINN_CHECK_LAUNCH(e);
            ad_reduce_chunks<<<e->G, AD_NT, 0, e->stream>>>(ad_partial, nchunk, e->d_S);  // per-glacier Σ (fixed order)
            ODINN_CHECK_LAUNCH(e);
            ODINN_CUDA(e, cudaMemsetAsync(d_active, 0, 2 * sizeof(int), e->stream));
            ad_control<<<gb, 128, 0, e->stream>>>(st, e->d_S, d_nx, d_ny, e->G, d_active);
            ODINN_CHECK_LAUNCH(e);
            ODINN_CUDA(e, cudaMemcpyAsync(e->h_ad_active, d_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
            ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
            if (e->h_ad_active[1] == 0) {
                // every glacier accepted (the ones already at the tstop took h = 0: H_new == H and k4 == k1 bit for bit): the
                // commit is a pointer swap instead of a 4 words/cell copy
                std::swap(H, Hn);
                std::swap(k1, k4);
            } else {
                ad_commit<T><<<egrid, AD_NT, 0, e->stream>>>(descs, st, H, Hn, k1, k4);
                ODINN_CHECK_LAUNCH(e);
            }
            if (e->h_ad_active[0] == 0) break;
        }
        int mb_applied = 0;
        if ((rc = mb_apply_step(e, j, H, &mb_applied))) return rc;   // mass-balance callback at the end of its window
        if (mb_applied && (rc = rhs_planes(e, H, k1))) return rc;    // u was modified by the callback: the FSAL slope is stale
        ODINN_CUDA(e, cudaMemcpyAsync(snapshot_ptr(e, j), H, pbytes, cudaMemcpyDeviceToDevice, e->stream));
    }
    if ((void*)H != e->plane[ODINN_FIELD_H])  // leave the final state in FIELD_H
        ODINN_CUDA(e, cudaMemcpyAsync(e->plane[ODINN_FIELD_H], H, pbytes, cudaMemcpyDeviceToDevice, e->stream));
    if (steps_out || rejected_out) {
        std::vector<AdState> hs(e->G);
        ODINN_CUDA(e, cudaMemcpyAsync(hs.data(), st, sizeof(AdState) * e->G, cudaMemcpyDeviceToHost, e->stream));
        ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
        for (int g = 0; g < e->G; ++g) {
            if (steps_out) steps_out[g] = hs[g].steps;
            if (rejected_out) rejected_out[g] = hs[g].rejected;
        }
    }
    return ODINN_OK;
}

}  // namespace odinn

using namespace odinn;

extern "C" int odinn_solve_forward_adaptive(odinn_ensemble* e, int method, int n_snap, const double* t, double reltol,
                                            double abstol, double dt0, int max_steps, int* steps_out, int* rejected_out) {
    if (!e) return fail(nullptr, ODINN_EARG, "null ensemble");
    {
        cudaError_t s_ = cudaSetDevice(e->device);
        if (s_ != cudaSuccess) return fail(e, ODINN_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(s_));
    }
    if (method != ODINN_BS3 && method != ODINN_RDPK3SP35) return fail(e, ODINN_EARG, "adaptive solve: method must be ODINN_BS3 or ODINN_RDPK3SP35");
    if (n_snap < 1 || !t || !(reltol > 0.0) || !(abstol > 0.0) || max_steps < 1) return fail(e, ODINN_EARG, "bad adaptive-solve arguments");
    for (int j = 1; j < n_snap; ++j)
        if (!(t[j] > t[j - 1])) return fail(e, ODINN_EARG, "tstops must be strictly increasing");
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_H0)) || (rc = ensure_plane(e, ODINN_FIELD_H))) return rc;
    if ((rc = prepare_snapshots(e, n_snap))) return rc;
    if ((rc = sync_descs(e))) return rc;
    if (method == ODINN_RDPK3SP35) {
        if ((rc = ensure_plane(e, ODINN_FIELD_B))) return rc;
        // small ensembles: the whole adaptive loop of every glacier runs inside one thread-block cluster (sia2d_cluster.cuh)
        if (const int cs = cluster_plan(e, 1))
            return solve_forward_rdpk_cluster(e, cs, n_snap, t, reltol, abstol, dt0, max_steps, steps_out, rejected_out);
        return solve_forward_rdpk(e, n_snap, t, reltol, abstol, dt0, max_steps, steps_out, rejected_out);
    }
    return e->dtype == ODINN_F32 ? solve_bs3_t<float>(e, n_snap, t, reltol, abstol, dt0, max_steps, steps_out, rejected_out)
                                 : solve_bs3_t<double>(e, n_snap, t, reltol, abstol, dt0, max_steps, steps_out, rejected_out);
}
