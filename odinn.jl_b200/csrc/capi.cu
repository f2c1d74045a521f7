// C ABI of libodinn_b200.so (see include/odinn_b200.h for the reference interfaces replaced).
#include "ensemble.cuh"
#include "sia2d_kernels.cuh"
#include "sia2d_march.cuh"
#include <cstdlib>

namespace odinn {

thread_local std::string g_create_error;

int ensure_plane(odinn_ensemble* e, int field) {
    if (field < 0 || field >= ODINN_FIELD_COUNT_) return fail(e, ODINN_EARG, "bad field id");
    if (e->plane[field]) return ODINN_OK;
    void* p = nullptr;
    size_t bytes = (size_t)e->total * e->esize;
    ODINN_CUDA(e, cudaMalloc(&p, bytes));
    ODINN_CUDA(e, cudaMemsetAsync(p, 0, bytes, e->stream));
    e->plane[field] = p;
    return ODINN_OK;
}

template <typename T>
static int sync_descs_t(odinn_ensemble* e) {
    std::vector<GDesc<T>> h(e->G);
    for (int g = 0; g < e->G; ++g) {
        const GlacierHost& s = e->gl[g];
        GDesc<T>& d = h[g];
        d.off = s.off;
        d.nx = s.nx;
        d.ny = s.ny;
        d.ld = s.ld;
        d.tile0 = s.tile0;
        d.dx = (T)s.dx;
        d.dy = (T)s.dy;
        d.inv_dx = (T)(1.0 / s.dx);
        d.inv_dy = (T)(1.0 / s.dy);
        d.A = (T)s.A;
        d.temp = (T)s.temp;
    }
    ODINN_CUDA(e, cudaMemcpyAsync(e->d_descs, h.data(), sizeof(GDesc<T>) * e->G, cudaMemcpyHostToDevice, e->stream));
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

// ---- launches -----------------------------------------------------------------------------

template <typename T>
static int launch_rhs_t(odinn_ensemble* e, const int2* tiles, int n_tiles, const int4* items, int n_items) {
    PhysDev<T> ph = make_phys<T>(e->phys);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const T* H = (const T*)e->plane[ODINN_FIELD_H];
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    const T* Af = (const T*)e->plane[ODINN_FIELD_A];
    T* dH = (T*)e->plane[ODINN_FIELD_DH];
    const bool eta1 = (e->phys.eta0 == 1.0);
#define L(CUB, AF)                                                                                         \
    do {                                                                                                   \
        if (e->use_tiled)                                                                                  \
            sia2d_rhs_kernel<T, CUB, AF><<<n_tiles, NT, 0, e->stream>>>(descs, tiles, H, B, Af, dH, ph);   \
        else if (eta1)                                                                                     \
            sia2d_rhs_march<T, CUB, AF, true><<<div_up(n_items, MARCH_WARPS), MARCH_WARPS * 32, 0, e->stream>>>( \
                descs, items, n_items, H, B, Af, dH, ph);                                                  \
        else                                                                                               \
            sia2d_rhs_march<T, CUB, AF, false><<<div_up(n_items, MARCH_WARPS), MARCH_WARPS * 32, 0, e->stream>>>( \
                descs, items, n_items, H, B, Af, dH, ph);                                                  \
    } while (0)
    if (e->cubic) {
        if (e->a_gridded) L(true, true); else L(true, false);
    } else {
        if (e->a_gridded) L(false, true); else L(false, false);
    }
#undef L
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

// tile range [t0, t0+n_tiles): the whole ensemble or one glacier
static int launch_rhs(odinn_ensemble* e, int g) {
    int t0 = 0, n_tiles = e->n_tiles, i0 = 0, n_items = e->n_items;
    if (g >= 0) {
        t0 = e->gl[g].tile0;
        n_tiles = e->gl[g].ntx * e->gl[g].nty;
        i0 = e->gl[g].item0;
        n_items = e->gl[g].n_items;
    }
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_H)) || (rc = ensure_plane(e, ODINN_FIELD_B)) ||
        (rc = ensure_plane(e, ODINN_FIELD_DH)))
        return rc;
    if (e->a_gridded && (rc = ensure_plane(e, ODINN_FIELD_A))) return rc;
    if ((rc = sync_descs(e))) return rc;
    return e->dtype == ODINN_F32 ? launch_rhs_t<float>(e, e->d_tiles + t0, n_tiles, e->d_items + i0, n_items)
                                 : launch_rhs_t<double>(e, e->d_tiles + t0, n_tiles, e->d_items + i0, n_items);
}

template <typename T>
static int launch_vjp_t(odinn_ensemble* e, int t0, int n_tiles, int i0, int n_items, bool wH, bool wS) {
    PhysDev<T> ph = make_phys<T>(e->phys);
    const GDesc<T>* descs = (const GDesc<T>*)e->d_descs;
    const int2* tiles = e->d_tiles + t0;
    const T* lam = (const T*)e->plane[ODINN_FIELD_LAMBDA];
    const T* H = (const T*)e->plane[ODINN_FIELD_H];
    const T* B = (const T*)e->plane[ODINN_FIELD_B];
    const T* Af = (const T*)e->plane[ODINN_FIELD_A];
    T* out = (T*)e->plane[ODINN_FIELD_VJP_H];
    T* vjpA = (wS && e->a_gridded) ? (T*)e->plane[ODINN_FIELD_VJP_A] : nullptr;
    double* partial = e->d_partial + (e->use_tiled ? t0 : i0);
    const int4* items = e->d_items + i0;
    const bool eta1 = (e->phys.eta0 == 1.0);
#define L(CUB, AF, WH, WS)                                                                                           \
    do {                                                                                                             \
        if (e->use_tiled)                                                                                            \
            sia2d_vjp_kernel<T, CUB, AF, WH, WS><<<n_tiles, NT, 0, e->stream>>>(descs, tiles, lam, H, B, Af, out,    \
                                                                                vjpA, partial, ph);                  \
        else if (eta1)                                                                                               \
            sia2d_vjp_march<T, CUB, AF, WH, WS, true><<<div_up(n_items, MARCH_WARPS), MARCH_WARPS * 32, 0, e->stream>>>( \
                descs, items, n_items, lam, H, B, Af, out, vjpA, partial, ph);                                       \
        else                                                                                                         \
            sia2d_vjp_march<T, CUB, AF, WH, WS, false><<<div_up(n_items, MARCH_WARPS), MARCH_WARPS * 32, 0, e->stream>>>( \
                descs, items, n_items, lam, H, B, Af, out, vjpA, partial, ph);                                       \
    } while (0)
#define L2(CUB, AF)                        \
    do {                                   \
        if (wH && wS) L(CUB, AF, true, true);   \
        else if (wH) L(CUB, AF, true, false);   \
        else L(CUB, AF, false, true);           \
    } while (0)
    if (e->cubic) {
        if (e->a_gridded) L2(true, true); else L2(true, false);
    } else {
        if (e->a_gridded) L2(false, true); else L2(false, false);
    }
#undef L2
#undef L
    ODINN_CHECK_LAUNCH(e);
    return ODINN_OK;
}

// g < 0: whole ensemble
static int launch_vjp(odinn_ensemble* e, int g, bool wH, bool wS) {
    if (!wH && !wS) return ODINN_OK;
    int rc;
    if ((rc = ensure_plane(e, ODINN_FIELD_H)) || (rc = ensure_plane(e, ODINN_FIELD_B)) ||
        (rc = ensure_plane(e, ODINN_FIELD_LAMBDA)))
        return rc;
    if (wH && (rc = ensure_plane(e, ODINN_FIELD_VJP_H))) return rc;
    if (e->a_gridded && ((rc = ensure_plane(e, ODINN_FIELD_A)) || (wS && (rc = ensure_plane(e, ODINN_FIELD_VJP_A)))))
        return rc;
    if ((rc = sync_descs(e))) return rc;
    int t0 = 0, nt = e->n_tiles, i0 = 0, ni = e->n_items;
    if (g >= 0) {
        t0 = e->gl[g].tile0;
        nt = e->gl[g].ntx * e->gl[g].nty;
        i0 = e->gl[g].item0;
        ni = e->gl[g].n_items;
    }
    rc = e->dtype == ODINN_F32 ? launch_vjp_t<float>(e, t0, nt, i0, ni, wH, wS)
                               : launch_vjp_t<double>(e, t0, nt, i0, ni, wH, wS);
    if (rc) return rc;
    if (wS) {
        const int* starts = e->use_tiled ? e->d_tile_start : e->d_item_start;
        if (g >= 0)
            reduce_items_kernel<<<1, NT, 0, e->stream>>>(starts + g, e->d_partial, e->d_S + g);
        else
            reduce_items_kernel<<<e->G, NT, 0, e->stream>>>(starts, e->d_partial, e->d_S);
        ODINN_CHECK_LAUNCH(e);
    }
    return ODINN_OK;
}

static int copy2d(odinn_ensemble* e, int g, int field, void* host, int ld, bool up, cudaStream_t st) {
    if (g < 0 || g >= e->G) return fail(e, ODINN_EARG, "glacier index out of range");
    if (!host) return fail(e, ODINN_EARG, "null host pointer");
    const GlacierHost& s = e->gl[g];
    bool dual = (field == ODINN_FIELD_A || field == ODINN_FIELD_VJP_A);
    int w = dual ? s.nx - 1 : s.nx, h = dual ? s.ny - 1 : s.ny;
    if (ld < w) return fail(e, ODINN_EARG, "ld smaller than the number of rows");
    int rc = ensure_plane(e, field);
    if (rc) return rc;
    char* dev = (char*)e->plane[field] + (size_t)s.off * e->esize;
    if (up)
        ODINN_CUDA(e, cudaMemcpy2DAsync(dev, (size_t)s.ld * e->esize, host, (size_t)ld * e->esize, (size_t)w * e->esize,
                                        h, cudaMemcpyHostToDevice, st));
    else
        ODINN_CUDA(e, cudaMemcpy2DAsync(host, (size_t)ld * e->esize, dev, (size_t)s.ld * e->esize, (size_t)w * e->esize,
                                        h, cudaMemcpyDeviceToHost, st));
    return ODINN_OK;
}

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
    long long off = 0;
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
        s.dx = dx[g];
        s.dy = dy[g];
        s.A = 0.0;
        s.temp = 0.0;
        s.ntx = div_up(s.nx, TX);
        s.nty = div_up(s.ny, TY);
        s.tile0 = tile;
        tile += s.ntx * s.nty;
        off += (long long)s.ld * s.ny;
        e->cells += (long long)s.nx * s.ny;
    }
    e->total = off;
    e->n_tiles = tile;
    {
        const char* k = std::getenv("ODINN_KERNEL");
        e->use_tiled = (k && std::string(k) == "tiled") ? 1 : 0;
    }
    // Marching work items: strips of STRIP output columns x chunks of rows.  Shorter chunks for small
    // ensembles so that the grid still fills 148 SMs.
    std::vector<int4> items;
    std::vector<int> istart(n_glaciers + 1);
    for (int rows = 32; rows >= 8; rows /= 2) {
        long long n = 0;
        for (int g = 0; g < n_glaciers; ++g) n += (long long)div_up(e->gl[g].nx, STRIP) * div_up(e->gl[g].ny, rows);
        e->chunk_rows = rows;
        if (n >= 148LL * 48) break;
    }
    for (int g = 0; g < n_glaciers; ++g) {
        GlacierHost& s = e->gl[g];
        s.item0 = (int)items.size();
        istart[g] = s.item0;
        for (int r0 = 0; r0 < s.ny; r0 += e->chunk_rows)
            for (int st = 0; st < div_up(s.nx, STRIP); ++st)
                items.push_back(make_int4(g, st * STRIP - 1, r0, std::min(r0 + e->chunk_rows, s.ny)));
        s.n_items = (int)items.size() - s.item0;
    }
    istart[n_glaciers] = (int)items.size();
    e->n_items = (int)items.size();

    std::vector<int2> tiles(tile);
    std::vector<int> tstart(n_glaciers + 1);
    for (int g = 0; g < n_glaciers; ++g) {
        const GlacierHost& s = e->gl[g];
        tstart[g] = s.tile0;
        for (int ty = 0; ty < s.nty; ++ty)
            for (int tx = 0; tx < s.ntx; ++tx) tiles[s.tile0 + ty * s.ntx + tx] = make_int2(g, (ty << 16) | tx);
    }
    tstart[n_glaciers] = tile;

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
    CREATE_CUDA(cudaMalloc(&e->d_descs, dsz * n_glaciers));
    CREATE_CUDA(cudaMalloc(&e->d_tiles, sizeof(int2) * tile));
    CREATE_CUDA(cudaMalloc(&e->d_tile_start, sizeof(int) * (n_glaciers + 1)));
    CREATE_CUDA(cudaMalloc(&e->d_partial, sizeof(double) * std::max(tile, e->n_items)));
    CREATE_CUDA(cudaMalloc(&e->d_items, sizeof(int4) * e->n_items));
    CREATE_CUDA(cudaMalloc(&e->d_item_start, sizeof(int) * (n_glaciers + 1)));
    CREATE_CUDA(cudaMemcpy(e->d_items, items.data(), sizeof(int4) * e->n_items, cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMemcpy(e->d_item_start, istart.data(), sizeof(int) * (n_glaciers + 1), cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMalloc(&e->d_S, sizeof(double) * n_glaciers));
    CREATE_CUDA(cudaMallocHost(&e->h_S, sizeof(double) * n_glaciers));
    CREATE_CUDA(cudaMemcpy(e->d_tiles, tiles.data(), sizeof(int2) * tile, cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMemcpy(e->d_tile_start, tstart.data(), sizeof(int) * (n_glaciers + 1), cudaMemcpyHostToDevice));
#undef CREATE_CUDA
    *out = e;
    return ODINN_OK;
}

void odinn_ensemble_destroy(odinn_ensemble* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (int f = 0; f < ODINN_FIELD_COUNT_; ++f)
        if (e->plane[f]) cudaFree(e->plane[f]);
    if (e->d_descs) cudaFree(e->d_descs);
    if (e->d_tiles) cudaFree(e->d_tiles);
    if (e->d_tile_start) cudaFree(e->d_tile_start);
    if (e->d_partial) cudaFree(e->d_partial);
    if (e->d_items) cudaFree(e->d_items);
    if (e->d_item_start) cudaFree(e->d_item_start);
    if (e->d_S) cudaFree(e->d_S);
    if (e->h_S) cudaFreeHost(e->h_S);
    if (e->h_stage) cudaFreeHost(e->h_stage);
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
    return ODINN_OK;
}

int odinn_sia2d_rhs(odinn_ensemble* e, int glacier, const void* H, int ldH, void* dH, int lddH, double t) {
    GUARD(e);
    (void)t;  // autonomous RHS: laws with callback_freq = 0 do not depend on t (Laws.jl:346)
    int rc;
    if ((rc = copy2d(e, glacier, ODINN_FIELD_H, const_cast<void*>(H), ldH, true, e->stream))) return rc;
    if ((rc = launch_rhs(e, glacier))) return rc;
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
    if ((rc = launch_vjp(e, glacier, true, false))) return rc;
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
    if ((rc = launch_vjp(e, glacier, false, true))) return rc;
    ODINN_CUDA(e, cudaMemcpyAsync(e->h_S + glacier, e->d_S + glacier, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    *out_S = e->h_S[glacier];
    return ODINN_OK;
}

int odinn_rhs_resident(odinn_ensemble* e) {
    GUARD(e);
    return launch_rhs(e, -1);
}

int odinn_vjp_resident(odinn_ensemble* e, int flags, double* S_out) {
    GUARD(e);
    int rc = launch_vjp(e, -1, (flags & 1) != 0, (flags & 2) != 0);
    if (rc) return rc;
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
    bool adj = (vjpH != nullptr) || (S != nullptr);
    if (adj && !lambda) return fail(e, ODINN_EARG, "lambda is required for the VJP outputs");
    int rc;
    for (int g = 0; g < e->G; ++g) {
        if ((rc = copy2d(e, g, ODINN_FIELD_H, const_cast<void*>(H[g]), e->gl[g].nx, true, e->stream))) return rc;
        if (adj && (rc = copy2d(e, g, ODINN_FIELD_LAMBDA, const_cast<void*>(lambda[g]), e->gl[g].nx, true, e->stream)))
            return rc;
    }
    if (dH && (rc = launch_rhs(e, -1))) return rc;
    if (adj && (rc = launch_vjp(e, -1, vjpH != nullptr, S != nullptr))) return rc;
    for (int g = 0; g < e->G; ++g) {
        if (dH && (rc = copy2d(e, g, ODINN_FIELD_DH, dH[g], e->gl[g].nx, false, e->stream))) return rc;
        if (vjpH && (rc = copy2d(e, g, ODINN_FIELD_VJP_H, vjpH[g], e->gl[g].nx, false, e->stream))) return rc;
    }
    if (S) ODINN_CUDA(e, cudaMemcpyAsync(e->h_S, e->d_S, sizeof(double) * e->G, cudaMemcpyDeviceToHost, e->stream));
    ODINN_CUDA(e, cudaStreamSynchronize(e->stream));
    if (S) memcpy(S, e->h_S, sizeof(double) * e->G);
    return ODINN_OK;
}

}  // extern "C"
