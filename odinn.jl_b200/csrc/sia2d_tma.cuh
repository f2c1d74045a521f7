// Fused F1 + A1 + A2 (fp32, n = 3, C = 0, glacier-wide A) with the inputs staged by 2-D TMA into a shared-memory ring.
//
// Same per-row arithmetic as sia2d_vjp_march2<WRITE_F> (VjpMarch2::compute_cubic, two columns per lane, packed f32x2) -- what
// changes is how the rows of lambda, H and B reach the registers:
//
//   * one CTA owns a BAND of TMA_NW adjacent strips (TMA_NW x 60 output columns + 2 halo columns on either side, + 2 more so that the box starts on a 16-byte boundary) over a chunk of
//     rows, and ONE elected thread issues `cp.async.bulk.tensor.2d` (tensor-map) loads of  TMA_BOXW columns x TMA_R rows  per
//     plane into a ring of TMA_STAGES stages -- one instruction per plane and stage instead of one LDG per warp, plane and row
//     (and instead of the 1-D `cp.async.bulk` per row of the retired bulk variant, whose address / ELECT / R2UR code cost 28
//     issue slots per 256-byte copy, profiles/r01_v4_sweep.txt);
//   * the strips of a band read their (overlapping) 64-column windows from shared memory, so the 4 halo columns between
//     neighbouring strips come from HBM once per band instead of once per strip, and a strip's row segment no longer straddles
//     an extra 32-byte sector (box rows are 752 bytes: 24 sectors for 180 useful columns vs 9 sectors for 60);
//   * out-of-grid columns and rows are zero-filled by the TMA unit (no clamped index arithmetic); every contribution they could
//     make is already masked by lmx / lmy / nodemask and the border-row logic of compute_cubic;
//   * the register prefetch queues (12 registers) and the L2-prefetch instructions of the marching kernel disappear: the ring
//     (TMA_STAGES - 1 boxes = 26 KB per CTA in flight) hides the DRAM latency.
//
// Ring protocol: full[s] is an mbarrier armed with the stage's byte count by the issuing thread and completed by the TMA unit;
// every strip warp waits on it (parity), consumes its TMA_R rows and bumps done[s]; the LAST warp to finish a stage re-arms the
// barrier and issues the box TMA_STAGES ahead into the same slot -- no producer warp, nobody ever waits for a free slot.
#pragma once
#include <cuda.h>  // CUtensorMap (type only: the encoder is obtained through cudaGetDriverEntryPoint, no libcuda link dependency)

#include "sia2d_march2.cuh"

namespace odinn {

#ifndef ODINN_TMA_NW
#define ODINN_TMA_NW 3
#endif
#ifndef ODINN_TMA_R
#define ODINN_TMA_R 4
#endif
constexpr int TMA_NW = ODINN_TMA_NW;                    // strips (warps) per CTA
// The innermost box coordinate must be a multiple of 16 bytes (c0 = -2 faults with "illegal instruction", c0 = -4 works:
// tools/microbench/tma_probe.cu), so a band's box starts 4 columns -- not 2 -- before its first output column: 4 + 180 + 4.
constexpr int TMA_BOXW = STRIP2 * TMA_NW + 8;           // box width in columns (188)
constexpr int TMA_R = ODINN_TMA_R;                      // rows per box
constexpr int TMA_PLANE_BYTES = (TMA_R * TMA_BOXW * 4 + 127) / 128 * 128;   // 3008 -> 3072: every plane of a stage starts 128-byte aligned
constexpr int TMA_PLANE_FLOATS = TMA_PLANE_BYTES / 4;
constexpr int TMA_BOX_BYTES = TMA_R * TMA_BOXW * 4;     // bytes one tensor-map copy delivers
constexpr int TMA_STAGE_FLOATS = 3 * TMA_PLANE_FLOATS;
constexpr int TMA_STAGE_BYTES = 3 * TMA_PLANE_BYTES;
constexpr int tma_smem_bytes(int stages) { return stages * TMA_STAGE_BYTES + stages * 8 + stages * 4 + 16; }
static_assert(TMA_PLANE_BYTES % 128 == 0 && (STRIP2 * TMA_NW) % 4 == 0, "TMA destinations must stay 128-byte aligned, box starts 16-byte aligned");
static_assert((TMA_BOXW * 4) % 16 == 0 && TMA_BOXW <= 256 && TMA_R <= 256, "box limits of cuTensorMapEncodeTiled");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tTMA_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra TMA_DONE;\n\tbra TMA_WAIT;\n\tTMA_DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// Band items: {glacier, first loaded column (180 b - 2), row0, row1}.  partial[(item) * TMA_NW + warp].
template <bool ETA1, int TMA_STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(TMA_NW * 32, MIN_CTAS)
sia2d_fused_tma(const GDesc<float>* __restrict__ descs, const int4* __restrict__ bitems, const CUtensorMap* __restrict__ mapsL,
                const CUtensorMap* __restrict__ mapsH, const CUtensorMap* __restrict__ mapsB, float* __restrict__ out,
                float* __restrict__ dH, double* __restrict__ partial, PhysDev<float> ph) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    float* ring = reinterpret_cast<float*>(tma_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(tma_smem + TMA_STAGES * TMA_STAGE_BYTES);
    int* done = reinterpret_cast<int*>(full + TMA_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int4 it = bitems[blockIdx.x];
    const GDesc<float> d = descs[it.x];
    const int r0 = it.z, r1 = it.w;
    const int nbox = (r1 - r0 + 2 + TMA_R - 1) / TMA_R;               // cell rows r0-1 .. r1
    const int nstrips = (d.nx + STRIP2 - 1) / STRIP2;
    const int nactive = min(TMA_NW, nstrips - (it.y + 2) / STRIP2);  // strips of this band that exist
    const CUtensorMap* mL = mapsL + it.x;
    const CUtensorMap* mH = mapsH + it.x;
    const CUtensorMap* mB = mapsB + it.x;
    const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(full);

    auto issue_box = [&](int b) {  // one thread: arm the slot's barrier, then one tensor-map copy per plane
        const int slot = b % TMA_STAGES;
        const uint32_t bar = full_s + 8 * slot, dst = ring_s + slot * TMA_STAGE_BYTES;
        const int row = r0 - 1 + b * TMA_R;
        mbar_expect_tx(bar, 3 * TMA_BOX_BYTES);
        tma_load_2d(dst, mL, it.y - 2, row, bar);   // (it.y - 2 = 180 b - 4: a multiple of 4 columns)
        tma_load_2d(dst + TMA_PLANE_BYTES, mH, it.y - 2, row, bar);
        tma_load_2d(dst + 2 * TMA_PLANE_BYTES, mB, it.y - 2, row, bar);
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(full_s + 8 * s, 1); done[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // the tensor maps live in global memory (written by cudaMemcpy): acquire them for the tensormap proxy
        asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(mL) : "memory");
        asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(mH) : "memory");
        asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(mB) : "memory");
        for (int b = 0; b < min(TMA_STAGES, nbox); ++b) issue_box(b);
    }
    __syncthreads();
    if (warp >= nactive) {
        if (lane == 0) partial[(long long)blockIdx.x * TMA_NW + warp] = 0.0;
        return;
    }

    // ---- per-strip setup (as in sia2d_vjp_march2) ----
    const int cbase = it.y + STRIP2 * warp;          // first loaded column of this strip (even)
    const int c0 = cbase + 2 * lane, c1 = c0 + 1;
    const int cmax = (d.nx - 1) & ~1;
    const int ic = min(max(c0, 0), cmax);
    VjpMarch2<true, false, true, true, ETA1, true> m;
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.eta0 = ph.eta0;
    const float hdx = 0.5f * d.inv_dx, hdy = 0.5f * d.inv_dy;
    m.nhx2 = bc2(-hdx * d.inv_dx);
    m.nhy2 = bc2(-hdy * d.inv_dy);
    m.A = bc2(d.A);
    m.lmask = mk2((c0 >= 1 && c0 <= d.nx - 2) ? 1.0f : 0.0f, (c1 >= 1 && c1 <= d.nx - 2) ? 1.0f : 0.0f);
    m.nodemask = mk2((c0 >= 0 && c0 <= d.nx - 2) ? 1.0f : 0.0f, (c1 >= 0 && c1 <= d.nx - 2) ? 1.0f : 0.0f);
    const bool out_lane = (lane >= 1 && lane <= 30 && c0 >= 0);
    // Padded layout only (ld > nx when nx is odd): the pair that straddles the last column stores its second element into the
    // padding column -- exactly 0 there (dH: lmx.y = lmy.y = 0; dSIA/dH^T lambda: the H > 0 mask of a zero padding cell), which is
    // what padding columns hold by invariant.  One store per output plane and row instead of a pair store + a scalar store.
    m.store_pair = out_lane && (c0 < d.nx);
    m.store_x = false;
    m.own_lane = (lane >= 1 && lane <= 30);
    m.vstore_pair = m.vstore_x = false;
    m.Ob = out; m.Fb = dH; m.Vb = nullptr;
    m.oout = (int)d.off + ic + (r0 - 1) * d.ld;   // dereferenced for rows >= r0 only
    {
        const float sK = sqrtf(ph.Gam * (1.0f / 1024.0f)), K2 = 2.0f * (ph.Gam * (1.0f / 1024.0f));
        m.hdxs = bc2(hdx * sK);
        m.hdys = bc2(hdy * sK);
        m.lmx = mul2(m.lmask, m.nhx2);
        m.lmy = mul2(m.lmask, m.nhy2);
        m.qx2 = bc2(K2 * (hdx * hdx));
        m.qy2 = bc2(K2 * (hdy * hdy));
    }
    m.Dp = m.aDp = m.Pp = m.Qp = m.yu_p = m.acc = m.Fyp = bc2(0.0f);

    const float* myring = ring + (2 + STRIP2 * warp + 2 * lane);   // (the box starts 2 columns before the band's first loaded column)
    auto ld_row = [&](const float* sp, int rr, f2& l1, f2& h1, f2& b1) {
        l1 = *reinterpret_cast<const float2*>(sp + rr * TMA_BOXW);
        h1 = *reinterpret_cast<const float2*>(sp + TMA_PLANE_FLOATS + rr * TMA_BOXW);
        b1 = *reinterpret_cast<const float2*>(sp + 2 * TMA_PLANE_FLOATS + rr * TMA_BOXW);
    };

    for (int b = 0; b < nbox; ++b) {
        const int slot = b % TMA_STAGES;
        mbar_wait(full_s + 8 * slot, (uint32_t)((b / TMA_STAGES) & 1));
        const float* sp = myring + slot * TMA_STAGE_FLOATS;
        const int i0 = r0 - 1 + b * TMA_R;   // cell row of the box's first row; step(row) consumes cell row row + 1
        if (b > 0 && i0 - 1 >= 1 && i0 + TMA_R - 2 <= d.ny - 3 && i0 + TMA_R - 1 <= r1) {
            // interior box: TMA_R unmasked output steps, rows i0-1 .. i0+TMA_R-2
#pragma unroll
            for (int rr = 0; rr < TMA_R; ++rr) {
                f2 l1, h1, b1;
                ld_row(sp, rr, l1, h1, b1);
                m.template compute_cubic<true, false>(i0 - 1 + rr, h1, b1, l1, m.A);
            }
        } else {
            for (int rr = 0; rr < TMA_R; ++rr) {
                const int i = i0 + rr;
                if (i > r1) break;
                f2 l1, h1, b1;
                ld_row(sp, rr, l1, h1, b1);
                if (b == 0 && rr == 0) {
                    // ---- cell row r0-1: the carried state of the first marching step ----
                    m.h = max2(h1, bc2(0.0f));
                    m.b = b1;
                    f2 l = mul2(l1, m.lmask);
                    if (!(r0 >= 2 && r0 <= m.nym1)) l = bc2(0.0f);  // row r0-1 must be an inner row
                    m.eh = ETA1 ? m.h : mul2(bc2(m.eta0), m.h);
                    f2 hE = east2(m.h), bE = east2(m.b);
                    m.ex = sdiff2(bE, m.b, hE, m.h);
                    m.hx = add2(m.h, hE);
                    m.ehE = ETA1 ? hE : mul2(bc2(m.eta0), hE);
                    m.cx = clamp2(m.ex, m.ehE, m.eh);
                    const f2 lx = mul2(l, m.nhx2);
                    m.ly = mul2(l, m.nhy2);
                    m.fx = sub2(east2(lx), lx);
                    m.px = mul2(m.fx, m.cx);
                } else if (i == r0) {
                    m.template compute_cubic<false, true>(i - 1, h1, b1, l1, m.A);
                } else {
                    m.template compute_cubic<true, true>(i - 1, h1, b1, l1, m.A);
                }
            }
        }
        // this warp is done with the slot; the last one to get here refills it with the box TMA_STAGES ahead
        __syncwarp();
        if (lane == 0) {
            const int prev = atomicAdd(&done[slot], 1);
            if (prev == nactive - 1) {
                done[slot] = 0;
                if (b + TMA_STAGES < nbox) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of the slot before the async-proxy writes
                    issue_box(b + TMA_STAGES);
                }
            }
        }
    }

    double a = m.own_lane ? (double)m.acc.x + (double)m.acc.y : 0.0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
    if (lane == 0) partial[(long long)blockIdx.x * TMA_NW + warp] = a;
}

}  // namespace odinn
