// fp32 marching kernels fed by the bulk-copy engine (TMA, 1-D form): cp.async.bulk global -> shared + mbarrier.
//
// ncu of the LDG-fed two-column kernels (profiles/r01_v4_*): the issue slots are only 45 % (F1) busy and the top
// stall is long_scoreboard -- with the prefetch queue in registers a warp keeps 4 rows x 2 planes x 256 B in flight
// and 94-128 registers/thread cap the SM at 16-20 warps, i.e. ~40 KB in flight per SM, below what HBM3e needs
// (6.5 TB/s x ~1 us / 148 SMs ~ 45 KB).  Deeper register queues cost occupancy (sweep: profiles/r01_v4_sweep.txt).
// Here the rows are fetched by the copy engine into a per-warp shared-memory ring, so the bytes in flight are
// bounded by shared memory (up to ~200 KB / SM), not by registers:
//   * every warp owns its ring of NST stages x R rows x (H, B [, lambda] [, A] [, U0]); no CTA-wide barrier exists;
//   * lane 0 arms the stage's mbarrier with the byte count and issues one 256..272-byte bulk copy per row and plane;
//   * the warp waits on the stage's mbarrier, marches through its R rows (LDS.64 per plane), and after a __syncwarp
//     lane 0 refills the stage with the rows NST stages ahead;
//   * the arithmetic is the two-column packed-f32x2 step of sia2d_march2.cuh (RhsMarch2::compute / VjpMarch2::compute).
// Bulk copies need 16-byte aligned global addresses and sizes: a strip's 64 columns start at column 60k-2, so the
// copy fetches the aligned superset [60k-4, 60k+64) clipped to [0, ld) -- which requires the padded plane layout
// (ld and plane offsets multiples of 32 elements).  Ring columns that no copy ever writes are zero-filled once.
#pragma once
#include "sia2d_march2.cuh"

namespace odinn {

constexpr int BK_RP = 68;  // ring row pitch in floats (272 B)
#ifndef ODINN_BK_R
#define ODINN_BK_R 4       // rows per stage
#endif
#ifndef ODINN_BK_NST
#define ODINN_BK_NST 4     // stages per warp
#endif
#ifndef ODINN_BK_WARPS
#define ODINN_BK_WARPS 4
#endif
constexpr int BK_R = ODINN_BK_R, BK_NST = ODINN_BK_NST, BK_WARPS = ODINN_BK_WARPS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Per-warp ring: NST stages x NARR planes x R rows x BK_RP floats, fed by lane 0.
// Sequence entry q (q = 0 .. total-1) is cell row clamp(r0-1+q) of the cell planes and row clamp(r0-2+q) of the
// node / stage planes (the rows marching step `row = r0-2+q` consumes).
template <int NARR>
struct BulkRing {
    float* ring;
    uint64_t* bars;
    const float* src[NARR];     // plane base + glacier offset + first copied column
    int rshift[NARR];           // row = q + rshift, clamped to [0, rmax]
    int rmax[NARR];
    int ld, total, dcol;
    uint32_t row_bytes;

    __device__ __forceinline__ float* slot(int st, int arr, int j) const {
        return ring + ((st * NARR + arr) * BK_R + j) * BK_RP;
    }
    // lane 0 only
    __device__ __forceinline__ void issue(int k) const {
        const int st = k % BK_NST;
        const int q0 = k * BK_R;
        const int n = min(BK_R, total - q0);
        uint64_t* bar = bars + st;
        mbar_expect_tx(bar, (uint32_t)(n * NARR) * row_bytes);
#pragma unroll
        for (int j = 0; j < BK_R; ++j) {
            if (j < n) {
#pragma unroll
                for (int a = 0; a < NARR; ++a) {
                    const int r = min(max(q0 + j + rshift[a], 0), rmax[a]);
                    bulk_g2s(slot(st, a, j) + dcol, src[a] + (long long)r * ld, row_bytes, bar);
                }
            }
        }
    }
    __device__ __forceinline__ void wait(int k) const { mbar_wait(bars + (k % BK_NST), (uint32_t)((k / BK_NST) & 1)); }
};

template <int NARR>
__device__ __forceinline__ void bulk_ring_setup(BulkRing<NARR>& rg, unsigned char* smem_raw, int warp, int lane) {
    constexpr int RING_FLOATS = BK_NST * NARR * BK_R * BK_RP;
    rg.ring = reinterpret_cast<float*>(smem_raw) + warp * RING_FLOATS;
    rg.bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)BK_WARPS * RING_FLOATS * sizeof(float)) + warp * BK_NST;
    float4* z = reinterpret_cast<float4*>(rg.ring);
    for (int i = lane; i < RING_FLOATS / 4; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < BK_NST; ++s) mbar_init(rg.bars + s, 1);
    }
    fence_proxy_async();  // zero fill + barrier init visible to the copy engine
    __syncwarp();
}

template <int NARR>
constexpr size_t bulk_smem_bytes() {
    return (size_t)BK_WARPS * (BK_NST * NARR * BK_R * BK_RP * sizeof(float) + BK_NST * sizeof(uint64_t));
}

// --------------------------------------------------------------------------------------------
// F1
// --------------------------------------------------------------------------------------------
template <bool CUBIC, bool AFIELD, bool ETA1, bool STAGE>
__global__ void __launch_bounds__(BK_WARPS * 32)
sia2d_rhs_bulk(const GDesc<float>* __restrict__ descs, const int4* __restrict__ items, int n_items,
               const float* __restrict__ H, const float* __restrict__ B, const float* __restrict__ Af, float* dH,
               PhysDev<float> ph, const float* U0, float sa, float sb, float sdt) {
    constexpr int NARR = 2 + (AFIELD ? 1 : 0) + (STAGE ? 1 : 0);
    constexpr int IA = 2, IU = 2 + (AFIELD ? 1 : 0);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * BK_WARPS + warp;
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<float> d = descs[it.x];
    const int c0 = it.y + 2 * lane, r0 = it.z, r1 = it.w;

    BulkRing<NARR> rg;
    bulk_ring_setup<NARR>(rg, smem_raw, warp, lane);
    {
        const int a0 = it.y - 2;                       // multiple of 4
        const int gs = max(a0, 0), ge = min(a0 + BK_RP, d.ld);
        rg.ld = d.ld;
        rg.total = r1 - r0 + 2;
        rg.dcol = gs - a0;
        rg.row_bytes = (uint32_t)(ge - gs) * 4u;
        rg.src[0] = H + d.off + gs;  rg.rshift[0] = r0 - 1;  rg.rmax[0] = d.ny - 1;
        rg.src[1] = B + d.off + gs;  rg.rshift[1] = r0 - 1;  rg.rmax[1] = d.ny - 1;
        if (AFIELD) { rg.src[IA] = Af + d.off + gs; rg.rshift[IA] = r0 - 2; rg.rmax[IA] = d.ny - 2; }
        if (STAGE) { rg.src[IU] = U0 + d.off + gs; rg.rshift[IU] = r0 - 2; rg.rmax[IU] = d.ny - 1; }
    }
    const int nstage = (rg.total + BK_R - 1) / BK_R;
    if (lane == 0) {
        for (int k = 0; k < min(BK_NST, nstage); ++k) rg.issue(k);
    }

    RhsMarch2<CUBIC, AFIELD, ETA1, STAGE> m;
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.eta0 = ph.eta0;
    const float hdx = 0.5f * d.inv_dx, hdy = 0.5f * d.inv_dy;
    m.hdx = bc2(hdx);
    m.hdy = bc2(hdy);
    const bool inx = (c0 >= 1 && c0 <= d.nx - 2), iny = (c0 + 1 >= 1 && c0 + 1 <= d.nx - 2);
    m.kx = mk2(inx ? hdx * d.inv_dx : 0.0f, iny ? hdx * d.inv_dx : 0.0f);
    m.ky = mk2(inx ? hdy * d.inv_dy : 0.0f, iny ? hdy * d.inv_dy : 0.0f);
    m.A = bc2(d.A);
    const bool out_lane = (lane >= 1 && lane <= 30 && c0 >= 0);
    m.store_pair = out_lane && (c0 + 1 < d.nx);
    m.store_x = out_lane && (c0 + 1 == d.nx);
    const int ic = min(max(c0, 0), (d.nx - 1) & ~1);
    m.op = dH + d.off + ic + (long long)(r0 - 1) * d.ld;  // dereferenced for rows >= r0 only
    m.sa = bc2(sa);
    m.sb = bc2(sb);
    m.sdt = bc2(sdt);
    m.hraw = bc2(0.0f);
    m.Dp = bc2(0.0f);
    m.Fy = bc2(0.0f);

    const int lo = 2 + 2 * lane;  // ring column of the lane's pair
    for (int k = 0; k < nstage; ++k) {
        rg.wait(k);
        const int st = k % BK_NST;
        const int q0 = k * BK_R;
        const int row0 = r0 - 2 + q0;  // marching row of the stage's first entry
        const float* base = rg.slot(st, 0, 0) + lo;
        auto ld2 = [&](int arr, int j) { return *reinterpret_cast<const float2*>(base + (arr * BK_R + j) * BK_RP); };
        const bool plain = (q0 >= 2) && (q0 + BK_R <= rg.total) && (row0 >= 1) && (row0 + BK_R - 1 <= m.nym1 - 1);
        if (plain) {
#pragma unroll
            for (int j = 0; j < BK_R; ++j) {
                f2 An = AFIELD ? ld2(IA, j) : m.A;
                f2 u0 = STAGE ? ld2(IU, j) : bc2(0.0f);
                m.template compute<true, false>(row0 + j, ld2(0, j), ld2(1, j), u0, An);
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < BK_R; ++j) {
                const int q = q0 + j;
                if (q < rg.total) {
                    f2 hv = ld2(0, j), bv = ld2(1, j);
                    f2 An = AFIELD ? ld2(IA, j) : m.A;
                    f2 u0 = STAGE ? ld2(IU, j) : bc2(0.0f);
                    if (q == 0) {  // cell row r0-1: initial state
                        m.h = max2(hv, bc2(0.0f));
                        m.b = bv;
                        m.eh = ETA1 ? m.h : mul2(bc2(m.eta0), m.h);
                        f2 hE = east2(m.h), bE = east2(m.b);
                        m.ex = sdiff2(bE, m.b, hE, m.h);
                        m.hx = add2(m.h, hE);
                        m.ehE = ETA1 ? hE : mul2(bc2(m.eta0), hE);
                    } else if (q == 1) {
                        m.template compute<false, true>(row0 + j, hv, bv, u0, An);  // warm-up: node row r0-1, no output
                    } else {
                        m.template compute<true, true>(row0 + j, hv, bv, u0, An);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0 && k + BK_NST < nstage) rg.issue(k + BK_NST);
    }
}

// --------------------------------------------------------------------------------------------
// A1 + A2
// --------------------------------------------------------------------------------------------
template <bool CUBIC, bool AFIELD, bool WRITE_H, bool WRITE_S, bool ETA1>
__global__ void __launch_bounds__(BK_WARPS * 32)
sia2d_vjp_bulk(const GDesc<float>* __restrict__ descs, const int4* __restrict__ items, int n_items,
               const float* __restrict__ lam, const float* __restrict__ H, const float* __restrict__ B,
               const float* __restrict__ Af, float* __restrict__ out, float* __restrict__ vjpA,
               double* __restrict__ partial, PhysDev<float> ph) {
    constexpr int NARR = 3 + (AFIELD ? 1 : 0);
    constexpr int IA = 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * BK_WARPS + warp;
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<float> d = descs[it.x];
    const int c0 = it.y + 2 * lane, c1 = c0 + 1, r0 = it.z, r1 = it.w;

    BulkRing<NARR> rg;
    bulk_ring_setup<NARR>(rg, smem_raw, warp, lane);
    {
        const int a0 = it.y - 2;
        const int gs = max(a0, 0), ge = min(a0 + BK_RP, d.ld);
        rg.ld = d.ld;
        rg.total = r1 - r0 + 2;
        rg.dcol = gs - a0;
        rg.row_bytes = (uint32_t)(ge - gs) * 4u;
        rg.src[0] = H + d.off + gs;    rg.rshift[0] = r0 - 1;  rg.rmax[0] = d.ny - 1;
        rg.src[1] = B + d.off + gs;    rg.rshift[1] = r0 - 1;  rg.rmax[1] = d.ny - 1;
        rg.src[2] = lam + d.off + gs;  rg.rshift[2] = r0 - 1;  rg.rmax[2] = d.ny - 1;
        if (AFIELD) { rg.src[IA] = Af + d.off + gs; rg.rshift[IA] = r0 - 2; rg.rmax[IA] = d.ny - 2; }
    }
    const int nstage = (rg.total + BK_R - 1) / BK_R;
    if (lane == 0) {
        for (int k = 0; k < min(BK_NST, nstage); ++k) rg.issue(k);
    }

    VjpMarch2<CUBIC, AFIELD, WRITE_H, WRITE_S, ETA1> m;
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.eta0 = ph.eta0;
    const float hdx = 0.5f * d.inv_dx, hdy = 0.5f * d.inv_dy;
    m.hdx = bc2(hdx);
    m.hdy = bc2(hdy);
    m.nhx2 = bc2(-hdx * d.inv_dx);
    m.nhy2 = bc2(-hdy * d.inv_dy);
    m.qx = bc2(hdx * hdx);
    m.qy = bc2(hdy * hdy);
    m.A = bc2(d.A);
    m.lmask = mk2((c0 >= 1 && c0 <= d.nx - 2) ? 1.0f : 0.0f, (c1 >= 1 && c1 <= d.nx - 2) ? 1.0f : 0.0f);
    m.nodemask = mk2((c0 >= 0 && c0 <= d.nx - 2) ? 1.0f : 0.0f, (c1 >= 0 && c1 <= d.nx - 2) ? 1.0f : 0.0f);
    const bool out_lane = (lane >= 1 && lane <= 30 && c0 >= 0);
    m.store_pair = out_lane && (c1 < d.nx);
    m.store_x = out_lane && (c1 == d.nx);
    m.own_lane = (lane >= 1 && lane <= 30);
    m.vstore_pair = out_lane && (c1 <= d.nx - 2);
    m.vstore_x = out_lane && (c1 == d.nx - 1);
    const int ic = min(max(c0, 0), (d.nx - 1) & ~1);
    m.Ob = out;
    m.Vb = vjpA;
    m.Fb = nullptr;
    m.oout = (int)d.off + ic + (r0 - 1) * d.ld;
    m.Dp = m.aDp = m.Pp = m.Qrow_p = m.yu_p = m.acc = bc2(0.0f);

    const int lo = 2 + 2 * lane;
    for (int k = 0; k < nstage; ++k) {
        rg.wait(k);
        const int st = k % BK_NST;
        const int q0 = k * BK_R;
        const int row0 = r0 - 2 + q0;
        const float* base = rg.slot(st, 0, 0) + lo;
        auto ld2 = [&](int arr, int j) { return *reinterpret_cast<const float2*>(base + (arr * BK_R + j) * BK_RP); };
        // unmasked steps need: an output step on a full stage, λ row row+1 inner (row+1 <= ny-2), node row valid
        const bool plain = (q0 >= 2) && (q0 + BK_R <= rg.total) && (row0 >= 0) && (row0 + BK_R - 1 + 1 < m.nym1);
        if (plain) {
#pragma unroll
            for (int j = 0; j < BK_R; ++j) {
                f2 An = AFIELD ? ld2(IA, j) : m.A;
                m.template compute<true, false>(row0 + j, ld2(0, j), ld2(1, j), ld2(2, j), An);
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < BK_R; ++j) {
                const int q = q0 + j;
                if (q < rg.total) {
                    f2 hv = ld2(0, j), bv = ld2(1, j), lv = ld2(2, j);
                    f2 An = AFIELD ? ld2(IA, j) : m.A;
                    if (q == 0) {  // cell row r0-1
                        m.h = max2(hv, bc2(0.0f));
                        m.b = bv;
                        m.l = mul2(lv, m.lmask);
                        if (!(r0 >= 2 && r0 <= m.nym1)) m.l = bc2(0.0f);  // row r0-1 must be an inner row
                        m.eh = ETA1 ? m.h : mul2(bc2(m.eta0), m.h);
                        f2 hE = east2(m.h), bE = east2(m.b), lE = east2(m.l);
                        m.ex = sdiff2(bE, m.b, hE, m.h);
                        m.hx = add2(m.h, hE);
                        m.ehE = ETA1 ? hE : mul2(bc2(m.eta0), hE);
                        m.fxr = sub2(lE, m.l);
                        m.px = mul2(m.fxr, clamp2(m.ex, m.ehE, m.eh));
                    } else if (q == 1) {
                        m.template compute<false, true>(row0 + j, hv, bv, lv, An);
                    } else {
                        m.template compute<true, true>(row0 + j, hv, bv, lv, An);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0 && k + BK_NST < nstage) rg.issue(k + BK_NST);
    }

    if (WRITE_S) {
        double a = (double)m.acc.x + (double)m.acc.y;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
        if (lane == 0) partial[item] = a;
    }
}

}  // namespace odinn
