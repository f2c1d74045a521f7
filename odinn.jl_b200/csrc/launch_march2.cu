// Launchers of the fp32 two-column marching kernels (sia2d_march2.cuh).
#include "launch.cuh"
#include "sia2d_march2.cuh"

namespace odinn {

// fp32, two columns per lane (sia2d_march2.cuh)
int launch_rhs2(odinn_ensemble* e, int g0, int g1, const void* Hin, void* out, const Stage* st, bool packed) {
    PhysDev<float> ph = make_phys<float>(e->phys);
    const GDesc<float>* descs = (const GDesc<float>*)e->d_descs + (packed ? e->G : 0);
    int i0 = 0, n_items = e->n_items2;
    if (g0 >= 0) {
        i0 = e->gl[g0].item20;
        n_items = e->gl[g1 - 1].item20 + e->gl[g1 - 1].n_items2 - i0;
    }
    const int4* items = e->d_items2 + i0;
    const float* H = (const float*)Hin;
    const float* B = (const float*)(packed ? e->bpack : e->plane[ODINN_FIELD_B]);
    const float* Af = (const float*)e->plane[ODINN_FIELD_A];
    float* dH = (float*)out;
    const bool eta1 = (e->phys.eta0 == 1.0);
    const float* U0 = st ? (const float*)st->U0 : nullptr;
    const float sa = st ? (float)st->sa : 0.f, sb = st ? (float)st->sb : 0.f, sdt = st ? (float)st->sdt : 0.f;
    const double* stab = st ? st->tab : nullptr;
    const int* sint = st ? st->interval : nullptr;
    dim3 grid(div_up(n_items, MARCH2_WARPS)), block(MARCH2_WARPS * 32);
#define L(CUB, AF, E1, STG) \
    sia2d_rhs_march2<CUB, AF, E1, STG><<<grid, block, 0, e->stream>>>(descs, items, n_items, H, B, Af, dH, ph, U0, sa, sb, sdt, stab, sint)
#define L3(CUB, AF, E1) do { if (st) L(CUB, AF, E1, true); else L(CUB, AF, E1, false); } while (0)
#define L2(CUB, AF) ODINN_ETA(L3, CUB, AF)
    ODINN_DISPATCH(L2);
#undef L2
#undef L3
#undef L
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

// fp32, two columns per lane (sia2d_march2.cuh).  Partials are indexed by the two-column work items.
// dH_out != nullptr (with wH && wS): the fused F1 + A1 + A2 pass.
int launch_vjp2(odinn_ensemble* e, int g0, int g1, const void* lam_, const void* H_, void* out_, bool wH, bool wS,
                bool packed, void* dH_out) {
    PhysDev<float> ph = make_phys<float>(e->phys);
    const GDesc<float>* descs = (const GDesc<float>*)e->d_descs + (packed ? e->G : 0);
    int i0 = 0, n_items = e->n_items2;
    if (g0 >= 0) {
        i0 = e->gl[g0].item20;
        n_items = e->gl[g1 - 1].item20 + e->gl[g1 - 1].n_items2 - i0;
    }
    const float* lam = (const float*)lam_;
    const float* H = (const float*)H_;
    const float* B = (const float*)(packed ? e->bpack : e->plane[ODINN_FIELD_B]);
    const float* Af = (const float*)e->plane[ODINN_FIELD_A];
    float* out = (float*)out_;
    float* vjpA = (wS && e->a_gridded) ? (float*)e->plane[ODINN_FIELD_VJP_A] : nullptr;
    double* partial = e->d_partial + i0;
    const int4* items = e->d_items2 + i0;
    const bool eta1 = (e->phys.eta0 == 1.0);
    dim3 grid(div_up(n_items, MARCH2_WARPS)), block(MARCH2_WARPS * 32);
#define L(CUB, AF, WH, WS, E1) \
    sia2d_vjp_march2<CUB, AF, WH, WS, E1><<<grid, block, 0, e->stream>>>(descs, items, n_items, lam, H, B, Af, out, vjpA, partial, ph)
#define LF(CUB, AF, E1)                                                                                                  \
    sia2d_vjp_march2<CUB, AF, true, true, E1, true><<<grid, block, 0, e->stream>>>(descs, items, n_items, lam, H, B, Af,  \
                                                                                   out, vjpA, partial, ph, (float*)dH_out)
#define L3(CUB, AF, E1)                         \
    do {                                        \
        if (dH_out) LF(CUB, AF, E1);                \
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

}  // namespace odinn
