// One-column marching launchers (included by launch_march_f32.cu / launch_march_f64.cu, which instantiate them for one T).
#include "launch.cuh"
#include "sia2d_march.cuh"
#include "sia2d_cont.cuh"
#include "timeloop.cuh"

namespace odinn {

template <typename T>
int launch_rhs_t(odinn_ensemble* e, int i0, int n_items, const void* Hin, void* out, const Stage* st, bool packed) {
    PhysDev<T> ph = make_phys<T>(e->phys);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs + (packed ? e->G : 0);
    const int4* items = e->d_items + i0;
    const T* H = (const T*)Hin;
    const T* B = (const T*)(packed ? e->bpack : e->plane[ODINN_FIELD_B]);
    const T* Af = (const T*)e->plane[ODINN_FIELD_A];
    T* dH = (T*)out;
    const bool eta1 = (e->phys.eta0 == 1.0);
    const T* U0 = st ? (const T*)st->U0 : nullptr;
    const T sa = st ? (T)st->sa : T(0), sb = st ? (T)st->sb : T(0), sdt = st ? (T)st->sdt : T(0);
    dim3 grid(div_up(n_items, MARCH_WARPS)), block(MARCH_WARPS * 32);
    const double* stab = st ? st->tab : nullptr;
    const int* sint = st ? st->interval : nullptr;
    cudaError_t lerr = cudaSuccess;
#define L(CUB, AF, E1, STG)                                                                                                               \
    lerr = launch_pdl(sia2d_rhs_march<T, CUB, AF, E1, STG, false, 0>, grid, block, e->stream, descs, items, n_items, H, B, Af, dH, ph, U0, \
                      sa, sb, sdt, T(0), 0, stab, sint, RkFuse<T>(), (double*)nullptr)
#define LRKM(CUB, AF, E1, M)                                                                                                              \
    lerr = launch_pdl(sia2d_rhs_march<T, CUB, AF, E1, false, false, M>, grid, block, e->stream, descs, items, n_items, H, B, Af, dH, ph,   \
                      (const T*)nullptr, T(0), T(0), T(0), T(0), 0, (const double*)nullptr, (const int*)nullptr, *(const RkFuse<T>*)st->rk, \
                      e->d_partial + i0)
#define LRK(CUB, AF, E1)                                                  \
    do {                                                                  \
        switch (rk_mode_of_flags(((const RkFuse<T>*)st->rk)->flags)) {    \
            case RKM_FIRST: LRKM(CUB, AF, E1, RKM_FIRST); break;          \
            case RKM_MID: LRKM(CUB, AF, E1, RKM_MID); break;              \
            case RKM_MID_U: LRKM(CUB, AF, E1, RKM_MID_U); break;          \
            default: LRKM(CUB, AF, E1, RKM_LAST); break;                  \
        }                                                                 \
    } while (0)
#define L3(CUB, AF, E1) do { if (st && st->rk) LRK(CUB, AF, E1); else if (st) L(CUB, AF, E1, true); else L(CUB, AF, E1, false); } while (0)
#define L2(CUB, AF) ODINN_ETA(L3, CUB, AF)
    ODINN_DISPATCH(L2);
#undef L2
#undef L3
#undef LRK
#undef LRKM
#undef L
    if (lerr != cudaSuccess) return fail(e, ODINN_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(lerr));
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

template <typename T>
int launch_rhs_law_t(odinn_ensemble* e, int i0, int n_items, const void* Hin, void* out, const Stage* st) {
    PhysDev<T> ph = make_phys<T>(e->phys);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const int4* items = e->d_items + i0;
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    const T* Dn = (const T*)e->lawD;
    const bool eta1 = (e->phys.eta0 == 1.0);
    const T* U0 = st ? (const T*)st->U0 : nullptr;
    const T sa = st ? (T)st->sa : T(0), sb = st ? (T)st->sb : T(0), sdt = st ? (T)st->sdt : T(0);
    dim3 grid(div_up(n_items, MARCH_WARPS)), block(MARCH_WARPS * 32);
#define LL(E1, STG) sia2d_rhs_march<T, true, true, E1, STG, true><<<grid, block, 0, e->stream>>>(descs, items, n_items, (const T*)Hin, B, Dn, (T*)out, ph, U0, sa, sb, sdt)
    if (eta1) { if (st) LL(true, true); else LL(true, false); }
    else { if (st) LL(false, true); else LL(false, false); }
#undef LL
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

template <typename T>
int launch_vjp_law_t(odinn_ensemble* e, int i0, int n_items, const void* lam, const void* H, void* out, bool wH, bool wS) {
    PhysDev<T> ph = make_phys<T>(e->phys);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const int4* items = e->d_items + i0;
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    const bool eta1 = (e->phys.eta0 == 1.0);
    T* vjpA = wS ? (T*)e->plane[ODINN_FIELD_VJP_A] : nullptr;
    dim3 grid(div_up(n_items, MARCH_WARPS)), block(MARCH_WARPS * 32);
#define LL(WH, WS, E1) sia2d_vjp_march<T, true, true, WH, WS, E1, true><<<grid, block, 0, e->stream>>>(descs, items, n_items, (const T*)lam, (const T*)H, B, (const T*)e->lawD, (T*)out, vjpA, e->d_partial + i0, ph, (const T*)e->lawAl, (const T*)e->lawBe)
#define LL2(E1) do { if (wH && wS) LL(true, true, E1); else if (wH) LL(true, false, E1); else LL(false, true, E1); } while (0)
    if (eta1) LL2(true); else LL2(false);
#undef LL2
#undef LL
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

template <typename T>
int launch_vjp_t(odinn_ensemble* e, int i0, int n_items, const void* lam_, const void* H_, void* out_, bool wH,
                        bool wS, bool packed, void* dH_out) {
    PhysDev<T> ph = make_phys<T>(e->phys);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs + (packed ? e->G : 0);
    const T* lam = (const T*)lam_;
    const T* H = (const T*)H_;
    const T* B = (const T*)(packed ? e->bpack : e->plane[ODINN_FIELD_B]);
    const T* Af = (const T*)e->plane[ODINN_FIELD_A];
    T* out = (T*)out_;
    T* vjpA = (wS && e->a_gridded) ? (T*)e->plane[ODINN_FIELD_VJP_A] : nullptr;
    double* partial = e->d_partial + i0;
    const int4* items = e->d_items + i0;
    const bool eta1 = (e->phys.eta0 == 1.0);
    dim3 grid(div_up(n_items, MARCH_WARPS)), block(MARCH_WARPS * 32);
#define L(CUB, AF, WH, WS, E1) \
    sia2d_vjp_march<T, CUB, AF, WH, WS, E1><<<grid, block, 0, e->stream>>>(descs, items, n_items, lam, H, B, Af, out, vjpA, partial, ph)
    // fused F1 + A1 + A2 (cubic form, fp64 only: the fp32 product path is the two-column kernel)
#define LF(AF, E1)                                                                                                            \
    do {                                                                                                                      \
        if constexpr (std::is_same<T, double>::value)                                                                         \
            sia2d_vjp_march<T, true, AF, true, true, E1, false, true><<<grid, block, 0, e->stream>>>(                         \
                descs, items, n_items, lam, H, B, Af, out, vjpA, partial, ph, nullptr, nullptr, (T*)dH_out);                  \
    } while (0)
#define L3(CUB, AF, E1)                         \
    do {                                        \
        if (dH_out && CUB) LF(AF, E1);              \
        else if (wH && wS) L(CUB, AF, true, true, E1);   \
        else if (wH) L(CUB, AF, true, false, E1);   \
        else L(CUB, AF, false, true, E1);           \
    } while (0)
#define L2(CUB, AF) ODINN_ETA(L3, CUB, AF)
    ODINN_DISPATCH(L2);
#undef L2
#undef L3
#undef LF
#undef L
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

// One RDPK3Sp35 stage of the continuous adjoint's reverse ODE in one pass (RKA variant of the cubic-form kernel: n = 3, C = 0,
// glacier-wide A); whole ensemble.  With RKF_NORM the per-item partial sums of the error norm land in d_partial.
template <typename T>
int launch_vjp_rk_t(odinn_ensemble* e, const void* S1in, const void* Ha, const void* Hb, void* S1out, const void* rkfuse, double c, double sign,
                    double ta, double tb, bool s_only) {
    PhysDev<T> ph = make_phys<T>(e->phys);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    const bool eta1 = (e->phys.eta0 == 1.0);
    const int n_items = e->n_items;
    dim3 grid(div_up(n_items, MARCH_WARPS)), block(MARCH_WARPS * 32);
#define LR(E1)                                                                                                                          \
    sia2d_vjp_march<T, true, false, true, false, E1, false, false, true><<<grid, block, 0, e->stream>>>(                                 \
        descs, e->d_items, n_items, (const T*)S1in, (const T*)Ha, B, nullptr, (T*)S1out, nullptr, e->d_partial, ph, nullptr, nullptr,   \
        nullptr, *(const RkFuse<T>*)rkfuse, (const T*)Hb, c, sign, ta, tb)
#define LQ(E1)                                                                                                                          \
    sia2d_vjp_march<T, true, false, false, true, E1, false, false, true><<<grid, block, 0, e->stream>>>(                                 \
        descs, e->d_items, n_items, (const T*)S1in, (const T*)Ha, B, nullptr, nullptr, nullptr, e->d_partial, ph, nullptr, nullptr,     \
        nullptr, *(const RkFuse<T>*)rkfuse, (const T*)Hb, c, sign, ta, tb)
    if (s_only) { if (eta1) LQ(true); else LQ(false); }
    else { if (eta1) LR(true); else LR(false); }
#undef LQ
#undef LR
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

template <typename T>
int launch_vjpc_t(odinn_ensemble* e, int i0, int n_items, const void* lam, const void* H, void* out) {
    PhysDev<T> ph = make_phys<T>(e->phys);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const int4* items = e->d_items + i0;
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    const T* Af = (const T*)e->plane[ODINN_FIELD_A];
    dim3 grid(div_up(n_items, MARCH_WARPS)), block(MARCH_WARPS * 32);
#define L(CUB, AF) sia2d_vjpc_march<T, CUB, AF><<<grid, block, 0, e->stream>>>(descs, items, n_items, (const T*)lam, (const T*)H, B, Af, (T*)out, ph)
    ODINN_DISPATCH(L);
#undef L
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

template <typename T>
int launch_unitA_dot_t(odinn_ensemble* e, int g, const void* lam, const void* H, double* S_dst, double scale, int accumulate) {
    odinn_phys p1 = e->phys;
    p1.C = 0.0;  // ∂D/∂A carries no sliding term (target_A.jl:71-72)
    PhysDev<T> ph = make_phys<T>(p1);
    const bool cubic = (p1.n == 3.0);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    int i0 = 0, ni = e->n_items, t0 = 0, nt = e->n_tiles;
    if (g >= 0) {
        i0 = e->gl[g].item0; ni = e->gl[g].n_items;
        t0 = e->gl[g].tile0; nt = e->gl[g].ntx * e->gl[g].nty;
    }
    const int4* items = e->d_items + i0;
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    T* scratch = (T*)e->work[0];
    const bool eta1 = (e->phys.eta0 == 1.0);
    dim3 grid(div_up(ni, MARCH_WARPS)), block(MARCH_WARPS * 32);
#define LU(CUB, E1) sia2d_rhs_march<T, CUB, false, E1, false><<<grid, block, 0, e->stream>>>(descs, items, ni, (const T*)H, B, nullptr, scratch, ph, nullptr, T(0), T(0), T(0), T(1), 1)
    if (cubic) { if (eta1) LU(true, true); else LU(true, false); }
    else { if (eta1) LU(false, true); else LU(false, false); }
#undef LU
    ODINN_CHECK_LAUNCH(e);
    dot_inner_kernel<T><<<nt, NT, 0, e->stream>>>(descs, e->d_tiles + t0, (const T*)lam, scratch, e->d_partial + t0);
    ODINN_CHECK_LAUNCH(e);
    if (g >= 0) reduce_scaled_kernel<<<1, NT, 0, e->stream>>>(e->d_tile_start + g, e->d_partial, S_dst + g, scale, accumulate);
    else reduce_scaled_kernel<<<e->G, NT, 0, e->stream>>>(e->d_tile_start, e->d_partial, S_dst, scale, accumulate);
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

#define ODINN_INSTANTIATE_MARCH(T)                                                                                                   \
    template int launch_rhs_t<T>(odinn_ensemble*, int, int, const void*, void*, const Stage*, bool);                                 \
    template int launch_rhs_law_t<T>(odinn_ensemble*, int, int, const void*, void*, const Stage*);                                   \
    template int launch_vjp_law_t<T>(odinn_ensemble*, int, int, const void*, const void*, void*, bool, bool);                        \
    template int launch_vjp_t<T>(odinn_ensemble*, int, int, const void*, const void*, void*, bool, bool, bool, void*);               \
    template int launch_vjpc_t<T>(odinn_ensemble*, int, int, const void*, const void*, void*);                                       \
    template int launch_vjp_rk_t<T>(odinn_ensemble*, const void*, const void*, const void*, void*, const void*, double, double, double, double, bool); \
    template int launch_unitA_dot_t<T>(odinn_ensemble*, int, const void*, const void*, double*, double, int);

}  // namespace odinn
