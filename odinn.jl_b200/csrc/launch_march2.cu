// Launchers of the fp32 two-column marching kernels (sia2d_march2.cuh) and of the 2-D TMA ring variant of the fused step
// (sia2d_tma.cuh).
#include <cstdlib>
#include <unordered_map>

#include "launch.cuh"
#include "sia2d_march2.cuh"
#include "sia2d_tma.cuh"

namespace odinn {

// ---- 2-D TMA ring variant: tensor maps (one per glacier and plane pointer) and band work items --------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmaCache {
    EncodeTiledFn encode = nullptr;
    std::unordered_map<const void*, CUtensorMap*> maps;   // plane base pointer -> device array of G tensor maps
    int4* d_bitems = nullptr;
    int* d_bstart = nullptr;                              // [G + 1] first partial slot of every glacier
    int n_bitems = 0;
    double* d_partial = nullptr;
    bool usable = false;
};
// Row chunks of the bands: 62 rows (+ 2 halo rows = 16 boxes of TMA_R = 4 rows), or ~100 when that still leaves two waves of CTAs -- every
// chunk pays two halo rows and a warm-up step (sweeps at 500x500x256, fused step: 62-row bands 0.2627 ms, 102 rows (5 chunks) 0.2570 ms,
// 126 rows (4 chunks) 0.2590 ms, 166 rows 0.266 ms, 250 rows 0.277 ms; profiles/r02_chunk_rows_sweep.txt).  ODINN_TMA_CHUNK_ROWS=<rows> overrides.
static int tma_chunk_rows(const odinn_ensemble* e) {
    const char* v = getenv("ODINN_TMA_CHUNK_ROWS");   // (read when the band table of an ensemble is built: the tests set it per ensemble)
    const int forced = v ? atoi(v) : 0;
    if (forced >= 6) return forced;
    long long ctas = 0;
    for (int g = 0; g < e->G; ++g) {
        const GlacierHost& s = e->gl[g];
        ctas += (long long)div_up(div_up(s.nx, STRIP2), TMA_NW) * std::max(1, (s.ny + 50) / 100);
    }
    return ctas >= 2LL * 148 * 5 ? 100 : 62;
}

void tma_cache_free(odinn_ensemble* e) {
    TmaCache* c = static_cast<TmaCache*>(e->tma_cache);
    if (!c) return;
    for (auto& kv : c->maps) cudaFree(kv.second);
    if (c->d_bitems) cudaFree(c->d_bitems);
    if (c->d_bstart) cudaFree(c->d_bstart);
    if (c->d_partial) cudaFree(c->d_partial);
    delete c;
    e->tma_cache = nullptr;
}

static int tma_prepare(odinn_ensemble* e, TmaCache*& c) {
    c = static_cast<TmaCache*>(e->tma_cache);
    if (c) return ODINN_OK;
    c = new TmaCache();
    e->tma_cache = c;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return ODINN_OK;  // (usable stays false: the caller falls back to the marching kernel)
    }
    c->encode = (EncodeTiledFn)fn;
    std::vector<int4> items;
    std::vector<int> start(e->G + 1, 0);
    const int chunk_rows = tma_chunk_rows(e);
    for (int g = 0; g < e->G; ++g) {
        const GlacierHost& s = e->gl[g];
        start[g] = (int)items.size() * TMA_NW;
        const int nstrips = div_up(s.nx, STRIP2), nbands = div_up(nstrips, TMA_NW);
        // row chunks of 4k - 2 rows (+ 2 halo rows = k boxes of TMA_R = 4 rows: no row is loaded that is not used), balanced so that
        // the last chunk of a glacier is not a sliver: 500 rows -> 7 x 66 + 38
        const int nch = std::max(1, (s.ny + chunk_rows / 2) / chunk_rows);
        const int rows = (div_up(s.ny, nch) + 2 + TMA_R - 1) / TMA_R * TMA_R - 2;
        for (int r0 = 0; r0 < s.ny; r0 += rows)
            for (int b = 0; b < nbands; ++b)
                items.push_back(make_int4(g, b * TMA_NW * STRIP2 - 2, r0, std::min(r0 + rows, s.ny)));
    }
    start[e->G] = (int)items.size() * TMA_NW;
    c->n_bitems = (int)items.size();
    ODINN_CUDA(e, cudaMalloc(&c->d_bitems, sizeof(int4) * items.size()));
    ODINN_CUDA(e, cudaMalloc(&c->d_bstart, sizeof(int) * start.size()));
    ODINN_CUDA(e, cudaMalloc(&c->d_partial, sizeof(double) * items.size() * TMA_NW));
    ODINN_CUDA(e, cudaMemcpy(c->d_bitems, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice));
    ODINN_CUDA(e, cudaMemcpy(c->d_bstart, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice));
    c->usable = true;
    return ODINN_OK;
}

// Device array of G tensor maps describing the plane at `base`: glacier g is an (nx x ny) fp32 matrix at element offset off[g] with
// row pitch ld[g]; box = TMA_BOXW columns x TMA_R rows, no swizzle, out-of-bounds elements read as zero.
static int tma_maps_for(odinn_ensemble* e, TmaCache* c, const void* base, const CUtensorMap** out) {
    auto it = c->maps.find(base);
    if (it != c->maps.end()) { *out = it->second; return ODINN_OK; }
    std::vector<CUtensorMap> h(e->G);
    for (int g = 0; g < e->G; ++g) {
        const GlacierHost& s = e->gl[g];
        const cuuint64_t dims[2] = {(cuuint64_t)s.nx, (cuuint64_t)s.ny};
        const cuuint64_t strides[1] = {(cuuint64_t)s.ld * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)TMA_BOXW, (cuuint32_t)TMA_R};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = c->encode(&h[g], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)((const float*)base + s.off), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(e, ODINN_ECUDA, "cuTensorMapEncodeTiled failed (plane offsets must be 16-byte aligned)");
    }
    CUtensorMap* d = nullptr;
    ODINN_CUDA(e, cudaMalloc(&d, sizeof(CUtensorMap) * e->G));
    ODINN_CUDA(e, cudaMemcpyAsync(d, h.data(), sizeof(CUtensorMap) * e->G, cudaMemcpyHostToDevice, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));  // (h goes out of scope)
    c->maps[base] = d;
    *out = d;
    return ODINN_OK;
}

// The fused F1 + A1 + A2 step through the TMA ring (whole ensemble, padded layout, n = 3, C = 0, glacier-wide A, eta0 = 1).
// *done = false when the variant does not apply (the caller then launches the marching kernel).
static int launch_fused_tma(odinn_ensemble* e, const float* lam, const float* H, const float* B, float* out, float* dH, bool* done,
                            const int** starts_used, double** partial_used) {
    *done = false;
    TmaCache* c = nullptr;
    int rc = tma_prepare(e, c);
    if (rc) return rc;
    if (!c->usable) return ODINN_OK;
    const CUtensorMap *mL, *mH, *mB;
    if ((rc = tma_maps_for(e, c, lam, &mL)) || (rc = tma_maps_for(e, c, H, &mH)) || (rc = tma_maps_for(e, c, B, &mB))) return rc;
    PhysDev<float> ph = make_phys<float>(e->phys);
    static const int variant = []() { const char* v = getenv("ODINN_TMA_VARIANT"); return v ? atoi(v) : 0; }();   // (tuning sweeps)
#define LT(ST, MC)                                                                                                                    \
    do {                                                                                                                              \
        if (tma_smem_bytes(ST) > 48 * 1024)                                                                                           \
            ODINN_CUDA(e, cudaFuncSetAttribute(sia2d_fused_tma<true, ST, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                               tma_smem_bytes(ST)));                                                                  \
        sia2d_fused_tma<true, ST, MC><<<c->n_bitems, TMA_NW * 32, tma_smem_bytes(ST), e->stream>>>(                                    \
            (const GDesc<float>*)e->d_descs, c->d_bitems, mL, mH, mB, out, dH, c->d_partial, ph);                                      \
    } while (0)
    if (variant == 1) LT(3, 6);
    else if (variant == 2) LT(3, 5);
    else if (variant == 3) LT(6, 5);
    else LT(4, 5);
#undef LT
    ODINN_CHECK_LAUNCH(e);
    *done = true;
    *starts_used = c->d_bstart;
    *partial_used = c->d_partial;
    return ODINN_OK;
}

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
    if (g0 < 0 && !(st && st->rk) && e->ext_int[5] > 0) {   // whole ensemble, no per-item partial sums: the long-chunk table (capi.cu)
        items = (const int4*)e->ext_dev[EXT_ITEMS2_LONG];
        n_items = e->ext_int[5];
    }
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
    cudaError_t lerr = cudaSuccess;
#define L(CUB, AF, E1, STG)                                                                                                           \
    lerr = launch_pdl(sia2d_rhs_march2<CUB, AF, E1, STG, 0>, grid, block, e->stream, descs, items, n_items, H, B, Af, dH, ph, U0, sa, sb, sdt, \
                      stab, sint, RkFuse<float>(), (double*)nullptr)
#define LRKM(CUB, AF, E1, M)                                                                                                          \
    lerr = launch_pdl(sia2d_rhs_march2<CUB, AF, E1, false, M>, grid, block, e->stream, descs, items, n_items, H, B, Af, dH, ph,       \
                      (const float*)nullptr, 0.f, 0.f, 0.f, (const double*)nullptr, (const int*)nullptr, *(const RkFuse<float>*)st->rk, e->d_partial)
#define LRK(CUB, AF, E1)                                                      \
    do {                                                                      \
        switch (rk_mode_of_flags(((const RkFuse<float>*)st->rk)->flags)) {    \
            case RKM_FIRST: LRKM(CUB, AF, E1, RKM_FIRST); break;              \
            case RKM_MID: LRKM(CUB, AF, E1, RKM_MID); break;                  \
            case RKM_MID_U: LRKM(CUB, AF, E1, RKM_MID_U); break;              \
            default: LRKM(CUB, AF, E1, RKM_LAST); break;                      \
        }                                                                     \
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

// fp32, two columns per lane (sia2d_march2.cuh).  Partials are indexed by the two-column work items.
// dH_out != nullptr (with wH && wS): the fused F1 + A1 + A2 pass.
int launch_vjp2(odinn_ensemble* e, int g0, int g1, const void* lam_, const void* H_, void* out_, bool wH, bool wS,
                bool packed, void* dH_out, const int** starts_used) {
    *starts_used = e->d_item2_start;
    if (e->march == 4 && dH_out && wH && wS && g0 < 0 && !packed && e->cubic && !e->a_gridded && e->phys.eta0 == 1.0) {
        bool done = false;
        double* part = nullptr;
        int rc = launch_fused_tma(e, (const float*)lam_, (const float*)H_, (const float*)e->plane[ODINN_FIELD_B], (float*)out_,
                                  (float*)dH_out, &done, starts_used, &part);
        if (rc) return rc;
        if (done) {  // the partial sums live in the variant's own buffer: hand them over through d_partial's reduce below
            e->tma_partial_live = part;
            return ODINN_OK;
        }
    }
    e->tma_partial_live = nullptr;
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
    if (g0 < 0 && e->ext_int[5] > 0) {   // whole (big) ensemble: the long-chunk table and its own start array (capi.cu)
        items = (const int4*)e->ext_dev[EXT_ITEMS2_LONG];
        n_items = e->ext_int[5];
        *starts_used = (const int*)e->ext_dev[EXT_ITEMS2_LONG_START];
    }
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


// One RDPK3Sp35 stage of the continuous adjoint's reverse ODE in one pass (RKA variant): S1out <- stage(S1in, k = (dSIA/dH)^T S1in at
// H_itp = lerp(Ha, Hb)); whole ensemble, glacier-wide A.  With RKF_NORM the per-item partial sums of the error norm land in d_partial.
// s_only: the A2 pass at a quadrature node instead (per-item partial sums of S at H_itp -> d_partial; rkfuse carries RKF_LERP_ONLY).
int launch_vjp2_rk(odinn_ensemble* e, const void* S1in, const void* Ha, const void* Hb, void* S1out, const void* rkfuse, double c, double sign,
                   double ta, double tb, bool s_only) {
    PhysDev<float> ph = make_phys<float>(e->phys);
    const GDesc<float>* descs = (const GDesc<float>*)e->d_descs;
    const int n_items = e->n_items2;
    const float* B = (const float*)e->plane[ODINN_FIELD_B];
    const bool eta1 = (e->phys.eta0 == 1.0);
    dim3 grid(div_up(n_items, MARCH2_WARPS)), block(MARCH2_WARPS * 32);
#define LR(CUB, E1)                                                                                                                       \
    sia2d_vjp_march2<CUB, false, true, false, E1, false, false, true><<<grid, block, 0, e->stream>>>(                                      \
        descs, e->d_items2, n_items, (const float*)S1in, (const float*)Ha, B, nullptr, (float*)S1out, nullptr, e->d_partial, ph, nullptr, \
        nullptr, nullptr, 0.f, 0.f, *(const RkFuse<float>*)rkfuse, (const float*)Hb, c, sign, ta, tb)
#define LQ(CUB, E1)                                                                                                                       \
    sia2d_vjp_march2<CUB, false, false, true, E1, false, false, true><<<grid, block, 0, e->stream>>>(                                      \
        descs, e->d_items2, n_items, (const float*)S1in, (const float*)Ha, B, nullptr, nullptr, nullptr, e->d_partial, ph, nullptr,       \
        nullptr, nullptr, 0.f, 0.f, *(const RkFuse<float>*)rkfuse, (const float*)Hb, c, sign, ta, tb)
    if (s_only) {
        if (e->cubic) { if (eta1) LQ(true, true); else LQ(true, false); }
        else { if (eta1) LQ(false, true); else LQ(false, false); }
    } else {
        if (e->cubic) { if (eta1) LR(true, true); else LR(true, false); }
        else { if (eta1) LR(false, true); else LR(false, false); }
    }
#undef LQ
#undef LR
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

int launch_vjp2_seed(odinn_ensemble* e, const void* lam_, const void* H_, const void* Href_, const void* W_, void* lam_new, double dt,
                     double cseed, const int** starts_used) {
    PhysDev<float> ph = make_phys<float>(e->phys);
    const GDesc<float>* descs = (const GDesc<float>*)e->d_descs;
    int n_items = e->n_items2;
    const int4* seed_items = e->d_items2;
    *starts_used = e->d_item2_start;
    if (e->ext_int[5] > 0) {   // big ensemble: the long-chunk table (capi.cu)
        seed_items = (const int4*)e->ext_dev[EXT_ITEMS2_LONG];
        n_items = e->ext_int[5];
        *starts_used = (const int*)e->ext_dev[EXT_ITEMS2_LONG_START];
    }
    const float* B = (const float*)e->plane[ODINN_FIELD_B];
    const bool eta1 = (e->phys.eta0 == 1.0);
    dim3 grid(div_up(n_items, MARCH2_WARPS)), block(MARCH2_WARPS * 32);
#define LS(CUB, E1)                                                                                                                      \
    sia2d_vjp_march2<CUB, false, true, false, E1, false, true><<<grid, block, 0, e->stream>>>(                                            \
        descs, seed_items, n_items, (const float*)lam_, (const float*)H_, B, nullptr, (float*)lam_new, nullptr, e->d_partial, ph, nullptr, \
        (const float*)Href_, (const float*)W_, (float)dt, (float)cseed)
    if (e->cubic) { if (eta1) LS(true, true); else LS(true, false); }
    else { if (eta1) LS(false, true); else LS(false, false); }
#undef LS
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

}  // namespace odinn
