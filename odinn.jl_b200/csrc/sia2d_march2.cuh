// fp32 register-marching kernels, two columns per lane, packed f32x2 arithmetic (Blackwell FADD2/FMUL2/FFMA2).
//
// ncu of the one-column kernels (profiles/r01_v3_*): DRAM traffic == algorithmic bytes, but 67 (F1) / 121 (A1+A2)
// thread-instructions per cell keep the issue slots 70-74 % busy at 66 % / 52 % of the HBM roofline -- the kernels
// are instruction-issue bound.  Measured pipe rates on B200 (tools/microbench/pipes.cu): FFMA 3.6, FFMA2 1.8,
// FMNMX 2.0, SHFL 1.0 warp-instructions / clk / SM.  A packed FFMA2 costs ONE issue slot for two columns, and with
// two columns per lane only every second x-neighbour is in another lane, so shuffles, loads, stores and address
// arithmetic per cell halve as well.  Same operator as sia2d_march.cuh (F1: same operation order per column; the A1+A2 step in
// its cubic form shares node products and differs from the generic form in rounding only -- the parity tests cover every form).
//
// Geometry: a warp owns 64 consecutive columns  base .. base+63  (base even), lane l holds columns base+2l (.x)
// and base+2l+1 (.y); lanes 0 and 31 are halo lanes, so a strip produces the 60 columns base+2 .. base+61.
// Every row access of a warp is one 256-byte LDG.64 / STG.64.  Requires even ld and even plane offsets.
// When nx is odd the last pair straddles the row end and reads one padding element: every plane is allocated
// zero-filled and no kernel or copy ever writes a padding column (pair stores are suppressed there), so the
// value read is 0 and only feeds quantities that the column masks (kx, ky, lmask, nodemask) zero out.
#pragma once
#include "sia2d_march.cuh"

namespace odinn {

constexpr int STRIP2 = 60;       // output columns per warp
#ifndef ODINN_MARCH2_WARPS
#define ODINN_MARCH2_WARPS 4
#endif
constexpr int MARCH2_WARPS = ODINN_MARCH2_WARPS;
#ifndef ODINN_PF2_RHS
#define ODINN_PF2_RHS 4
#endif
#ifndef ODINN_PF2_VJP
#define ODINN_PF2_VJP 2
#endif

#ifndef ODINN_L2PF_ROWS
#define ODINN_L2PF_ROWS 8   // rows of L2 prefetch distance ahead of the register queue (0 = off; sweep: profiles/r01_v4_sweep.txt)
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// One prefetch instruction per row for ALL input planes of a strip: lane 9p+s (s = 0..8) owns the 32-byte sector s of
// the 256-byte (possibly misaligned: 9 sectors) row segment of plane p.  Replaces one prefetch + one 64-bit address add
// per plane and row (3 + 6 issue slots in the A1+A2 kernel) by one of each.
template <int NPLANES>
__device__ __forceinline__ const float* prefetch_lane_ptr(int lane, int col0, int cmax, long long off, const float* p0,
                                                          const float* p1, const float* p2) {
    const int pl = lane / 9, sec = lane - 9 * pl;
    if (pl >= NPLANES) return nullptr;
    const float* base = pl == 0 ? p0 : (pl == 1 ? p1 : p2);
    if (base == nullptr) return nullptr;
    return base + off + min(max(col0 + 8 * sec, 0), cmax);
}

#ifndef ODINN_VJP_RING_UNROLL
#define ODINN_VJP_RING_UNROLL 1
#endif
#ifndef ODINN_VJP2_MIN_CTAS
#define ODINN_VJP2_MIN_CTAS 4   // 4 CTAs x 4 warps per SM: caps the A1+A2 kernels at 128 registers
#endif
constexpr int VJP_RING_UNROLL = ODINN_VJP_RING_UNROLL;  // ring passes per main-loop iteration of the A1+A2 kernel

typedef float2 f2;
__device__ __forceinline__ f2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __ffma2_rn(b, mk2(-1.0f, -1.0f), a); }  // a - b, one FFMA2
__device__ __forceinline__ f2 max2(f2 a, f2 b) { return mk2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
__device__ __forceinline__ f2 min2(f2 a, f2 b) { return mk2(fminf(a.x, b.x), fminf(a.y, b.y)); }
__device__ __forceinline__ f2 neg2(f2 a) { return mk2(-a.x, -a.y); }
// value of the column to the east / west of each of the lane's two columns
__device__ __forceinline__ f2 east2(f2 v) { return mk2(v.y, __shfl_down_sync(FULL, v.x, 1)); }
__device__ __forceinline__ f2 west2(f2 v) { return mk2(__shfl_up_sync(FULL, v.y, 1), v.x); }
__device__ __forceinline__ f2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
// raw surface difference, fp32 form (see sdiff<float>): (b1 - b0) + (h1 - h0)
__device__ __forceinline__ f2 sdiff2(f2 b1, f2 b0, f2 h1, f2 h0) { return add2(sub2(b1, b0), sub2(h1, h0)); }
// clamp(e, -lo, up) = max(min(e, up), -lo)
__device__ __forceinline__ f2 clamp2(f2 e, f2 up, f2 lo) {
    return mk2(fmaxf(fminf(e.x, up.x), -lo.x), fmaxf(fminf(e.y, up.y), -lo.y));
}

template <bool CUBIC, bool PARTIALS>
__device__ __forceinline__ void node_raw2(const PhysDev<float>& ph, f2 A, f2 Hs, f2 g2, f2& D, f2& alpha, f2& beta, f2& gA) {
    if (CUBIC) {
        const float K = ph.Gam * (1.0f / 1024.0f);  // Γ/4^5
        f2 H2 = mul2(Hs, Hs);
        f2 H4 = mul2(H2, H2);
        f2 w = mul2(H4, Hs);
        if (!PARTIALS) {
            D = mul2(mul2(mul2(A, bc2(K)), g2), w);
            return;
        }
        f2 tg = mul2(bc2(K), g2);
        gA = mul2(tg, w);
        D = mul2(A, gA);
        alpha = mul2(mul2(bc2(20.0f), A), mul2(tg, H4));
        beta = mul2(mul2(mul2(bc2(2.0f), A), bc2(K)), w);
    } else {
        float Dx, ax, bx, gx, Dy, ay, by, gy;
        node_diffusivity<float, false, PARTIALS>(ph, A.x, 0.25f * Hs.x, g2.x, Dx, ax, bx, gx);
        node_diffusivity<float, false, PARTIALS>(ph, A.y, 0.25f * Hs.y, g2.y, Dy, ay, by, gy);
        D = mk2(Dx, Dy);
        if (PARTIALS) { alpha = mk2(ax, ay); beta = mk2(bx, by); gA = mk2(gx, gy); }
    }
}

// PF unmasked output steps in ring form (compile-time queue slots)
template <int PF, int K = 0, typename M>
__device__ __forceinline__ void ring_steps(M& m, int row) {
    if constexpr (K < PF) {
        m.template step<true, false, K>(row + K);
        ring_steps<PF, K + 1>(m, row);
    }
}

// --------------------------------------------------------------------------------------------
// F1 (see RhsMarch for the step structure; every quantity is a pair of adjacent columns)
// --------------------------------------------------------------------------------------------
// RK > 0: an RDPK3Sp35 stage as the epilogue (RkFuse, common.cuh) instead of the two-register stage of STAGE; the stage's shape is a
// compile-time mode (RKM_*, common.cuh) so that the epilogue is straight-line code: every lane loads and computes (the clamped pair
// index keeps the addresses inside the grid), only the stores and the norm accumulation are predicated.
template <bool CUBIC, bool AFIELD, bool ETA1, bool STAGE, int RK = 0>
struct RhsMarch2 {
    static constexpr bool RK_FIRST = RK == RKM_FIRST, RK_U = RK == RKM_MID_U || RK == RKM_LAST, RK_WR = RK == RKM_MID || RK == RKM_MID_U,
                          RK_NORM = RK == RKM_LAST;
    static constexpr int PF = ODINN_PF2_RHS;
    const float *hp, *bp, *ap, *up;
    const float* pfp;  // this lane's L2-prefetch sector (prefetch_lane_ptr), ODINN_L2PF_ROWS rows ahead of hp
    float* op;
    // RK: plane bases (warp-uniform; every plane shares the layout, so one element offset `oo` of the output row serves them all),
    // the glacier's (b h, e h), the operands of the output row (loaded at the top of the step) and the error-norm accumulator
    RkFuse<float> rk;
    const float* outb;
    const float* pfp2;  // second L2-prefetch sector pointer: est and u planes
    int oo;
    float rbh, reh;
    f2 r_s2, r_e, r_u, nacc;       // operands of the output row
    f2 nw;                         // 1 on the columns this lane owns (error-norm weights)
    f2 n_s2, n_e, n_u;             // ... of the next output row (loaded one step ahead: they have no deeper register queue)
    bool pf2_lane;
    int ld, nym1, ny2;
    float eta0;
    f2 hdx, hdy, kx, ky, A;  // kx, ky are zeroed on border / out-of-grid columns
    f2 sa, sb, sdt, hraw;
    bool store_pair, store_x, pf_lane;
    PhysDev<float> ph;
    f2 h, b, eh, ex, hx, ehE, Dp, Fy;
    f2 hq[PF], bq[PF];


    // SLOT < 0: the prefetch queue is shifted (hq[0] is always the next row).  SLOT >= 0: ring form for the
    // unrolled main loop -- the step consumes hq[SLOT] and refills the same slot, so no register moves.
    template <bool OUT, bool MASKED, int SLOT = -1>
    __device__ __forceinline__ void step(int row) {
        constexpr int RS = SLOT < 0 ? 0 : SLOT;        // slot read
        constexpr int WS = SLOT < 0 ? PF - 1 : SLOT;   // slot refilled
        f2 h1 = hq[RS], b1 = bq[RS];
        if (SLOT < 0) {
#pragma unroll
            for (int k = 0; k + 1 < PF; ++k) { hq[k] = hq[k + 1]; bq[k] = bq[k + 1]; }
        }
        if (MASKED) {
            int stp = (row + 1 + PF <= nym1) ? ld : 0;
            hp += stp;
            bp += stp;
        } else {
            hp += ld;
            bp += ld;
        }
        hq[WS] = ldg2(hp);
        bq[WS] = ldg2(bp);
        if (ODINN_L2PF_ROWS > 0) {
            if (MASKED) {
                pfp += (row + 1 + PF <= nym1) ? ld : 0;
            } else {
                pfp += ld;
                if (row + 1 + PF + ODINN_L2PF_ROWS <= nym1 && pf_lane) prefetch_l2(pfp);
            }
        }
        f2 u0 = bc2(0.0f);
        if (STAGE && OUT) {
            if (store_pair) u0 = ldg2(up);
            if (store_x) u0.x = __ldg(up);
        }
        if (RK) {
            if (ODINN_L2PF_ROWS > 0) {
                if (MASKED) {
                    pfp2 += (row + 1 + PF <= nym1) ? ld : 0;
                } else {
                    pfp2 += ld;
                    if (row + 1 + PF + ODINN_L2PF_ROWS <= nym1 && pf2_lane) prefetch_l2(pfp2);
                }
            }
            // operands of the NEXT output row (plain loads: S2 and est are rewritten in place by their owner; a row past the chunk
            // belongs to another warp and is never used; odd nx: the straddling pair reads the zero padding element)
            {
                const int on = oo + (MASKED ? ((row + 1 <= nym1) ? ld : 0) : ld);
                if (!RK_FIRST) {
                    r_s2 = n_s2; r_e = n_e;
                    n_s2 = *reinterpret_cast<const float2*>(rk.S2in + on);
                    n_e = *reinterpret_cast<const float2*>(rk.est + on);
                }
                if (RK_U) { r_u = n_u; n_u = *reinterpret_cast<const float2*>(rk.u + on); }
            }
        }
        f2 Anode = A;
        if (AFIELD) {
            Anode = ldg2(ap);
            if (MASKED) { if (row >= 0 && row < ny2) ap += ld; } else ap += ld;
        }
        compute<OUT, MASKED>(row, h1, b1, u0, Anode);
        if (STAGE) up += ld;
        if (RK) oo += ld;
    }

    // RDPK3Sp35 stage on the output row (operation order of rk_stage / rk_stage1_main in rdpk.cu); k = SIA2D(S1) of the pair.
    __device__ __forceinline__ void rk_epilogue(f2 k) {
        const f2 s1 = hraw;
        f2 s1n, er;
        if (RK_FIRST) {
            s1n = mk2(fmaf(rbh, k.x, s1.x), fmaf(rbh, k.y, s1.y));
            er = mk2(reh * k.x, reh * k.y);
        } else {
            const f2 s2 = mk2(fmaf(rk.d, s1.x, r_s2.x), fmaf(rk.d, s1.y, r_s2.y));
            f2 v = mk2(fmaf(rk.g2, s2.x, rk.g1 * s1.x), fmaf(rk.g2, s2.y, rk.g1 * s1.y));
            if (RK_U) v = mk2(fmaf(rk.g3, r_u.x, v.x), fmaf(rk.g3, r_u.y, v.y));
            s1n = mk2(fmaf(rbh, k.x, v.x), fmaf(rbh, k.y, v.y));
            er = mk2(fmaf(reh, k.x, r_e.x), fmaf(reh, k.y, r_e.y));
            if (RK_WR) {
                if (store_pair) *reinterpret_cast<float2*>(rk.S2out + oo) = s2;
                if (store_x) rk.S2out[oo] = s2.x;
            }
        }
        if (store_pair) *reinterpret_cast<float2*>(op) = s1n;
        if (store_x) *op = s1n.x;
        if (RK_FIRST || RK_WR) {
            if (store_pair) *reinterpret_cast<float2*>(rk.est + oo) = er;
            if (store_x) rk.est[oo] = er.x;
        }
        if (RK_NORM) {
            const float rx = __fdividef(er.x, fmaf(rk.reltol, fmaxf(fabsf(r_u.x), fabsf(s1n.x)), rk.abstol));
            const float ry = __fdividef(er.y, fmaf(rk.reltol, fmaxf(fabsf(r_u.y), fabsf(s1n.y)), rk.abstol));
            nacc.x = fmaf(nw.x * rx, rx, nacc.x);
            nacc.y = fmaf(nw.y * ry, ry, nacc.y);
        }
    }

    // One marching step given the cell row `row+1` (h1, b1), the stage operand u0 and the node coefficient A of node
    // row `row`; produces output row `row` at `op` and advances `op`.
    template <bool OUT, bool MASKED>
    __device__ __forceinline__ void compute(int row, f2 h1, f2 b1, f2 u0, f2 Anode) {
        const f2 hraw1 = h1;
        h1 = max2(h1, bc2(0.0f));                 // adjoint.jl:52
        f2 eh1 = ETA1 ? h1 : mul2(bc2(eta0), h1);
        f2 hE1 = east2(h1), bE1 = east2(b1);
        f2 ex1 = sdiff2(bE1, b1, hE1, h1);        // raw x-edge difference S[i+1]-S[i], row+1
        f2 hx1 = add2(h1, hE1);
        f2 ehE1 = ETA1 ? hE1 : mul2(bc2(eta0), hE1);
        f2 ey = sdiff2(b1, b, h1, h);             // raw y-edge difference S[j+1]-S[j]
        f2 eyE = east2(ey);
        f2 u = mul2(add2(ex, ex1), hdx), v = mul2(add2(ey, eyE), hdy);
        f2 g2 = fma2(v, v, mul2(u, u));
        f2 D1, al, be, gA;
        node_raw2<CUBIC, false>(ph, Anode, add2(hx, hx1), g2, D1, al, be, gA);
        f2 D1W = west2(D1);
        f2 Fy1 = mul2(add2(D1W, D1), clamp2(ey, eh1, eh));
        f2 Fx = mul2(add2(Dp, D1), clamp2(ex, ehE, eh));
        f2 FxW = west2(Fx);
        if (OUT) {
            f2 outv = fma2(ky, sub2(Fy1, Fy), mul2(kx, sub2(Fx, FxW)));
            if (MASKED) { if (row < 1 || row >= nym1) outv = bc2(0.0f); }
            if (STAGE) outv = fma2(sb, fma2(sdt, outv, hraw), mul2(sa, u0));
            if (RK) {
                rk_epilogue(outv);
            } else {
                if (store_pair) *reinterpret_cast<float2*>(op) = outv;
                if (store_x) *op = outv.x;  // (exclusive with store_pair: the pair straddling the last column, odd nx)
            }
        }
        op += ld;
        if (STAGE || RK) hraw = hraw1;
        h = h1; b = b1; eh = eh1; ex = ex1; hx = hx1; ehE = ehE1; Dp = D1; Fy = Fy1;
    }
};

template <bool CUBIC, bool AFIELD, bool ETA1, bool STAGE, int RK = 0>
__global__ void __launch_bounds__(MARCH2_WARPS * 32, RK == RKM_LAST ? 4 : 0)   // (the last RDPK stage would take 148 registers: 3 CTAs / SM)
sia2d_rhs_march2(const GDesc<float>* __restrict__ descs, const int4* __restrict__ items, int n_items,
                 const float* __restrict__ H, const float* __restrict__ B, const float* __restrict__ Af, float* dH,
                 PhysDev<float> ph, const float* U0, float sa, float sb, float sdt,
                 const double* __restrict__ stage_tab = nullptr, const int* __restrict__ interval = nullptr,
                 RkFuse<float> rkf = RkFuse<float>(), double* __restrict__ partial = nullptr) {
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * MARCH2_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<float> d = descs[it.x];
    const int c0 = it.y + 2 * lane, r0 = it.z, r1 = it.w;  // it.y even
    // pair index clamped into the grid (pairs start at even columns; the last pair may straddle nx when nx is odd)
    const int cmax = (d.nx - 1) & ~1;
    const int ic = min(max(c0, 0), cmax);
    RhsMarch2<CUBIC, AFIELD, ETA1, STAGE, RK> m;
    constexpr int PF = ODINN_PF2_RHS;
    // Everything above reads tables that no F1 kernel writes (a kernel that does write them -- set_A_kernel -- never triggers early, so
    // it has completed before this prologue starts); from here on the launch depends on the previous kernel of the stream.
    pdl_wait();
    if (RK) { if (rkf.st[it.x].done) return; }   // the glacier has landed on the tstop: nothing to integrate (rk_integrate commits by copy then)
    if (STAGE && stage_tab != nullptr) {  // graph replay: stage coefficients from the device table (see sia2d_rhs_march)
        const double* sp = stage_tab + (long long)(*interval) * 9;
        sa = (float)sp[0]; sb = (float)sp[1]; sdt = (float)sp[2];
    }
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.eta0 = ph.eta0;
    const float hdx = 0.5f * d.inv_dx, hdy = 0.5f * d.inv_dy;
    m.hdx = bc2(hdx);
    m.hdy = bc2(hdy);
    const bool inx = (c0 >= 1 && c0 <= d.nx - 2), iny = (c0 + 1 >= 1 && c0 + 1 <= d.nx - 2);
    m.kx = mk2(inx ? hdx * d.inv_dx : 0.0f, iny ? hdx * d.inv_dx : 0.0f);  // ½/Δx²
    m.ky = mk2(inx ? hdy * d.inv_dy : 0.0f, iny ? hdy * d.inv_dy : 0.0f);
    m.A = bc2(d.A);
    const bool out_lane = (lane >= 1 && lane <= 30 && c0 >= 0);
    m.store_pair = out_lane && (c0 + 1 < d.nx);
    m.store_x = out_lane && (c0 + 1 == d.nx);
    const int rc = max(r0 - 1, 0);
    m.hp = H + d.off + ic + (long long)rc * d.ld;
    m.bp = B + d.off + ic + (long long)rc * d.ld;
    {
        // L2 prefetch: H and B, plus U0 in a stage launch (9 lanes per plane, one 32-byte sector each).  U0 has no register prefetch
        // queue (it is read at the output row), so the L2 prefetch is what hides its DRAM latency: SSPRK3 stage 0.226 -> 0.197 ms.
        // (Skipping the U0 read in stages with sa == 0 was measured SLOWER -- 0.35 ms: the branch serialises the load in every stage.)
        const float* q = STAGE ? prefetch_lane_ptr<3>(lane, it.y, cmax, d.off, H, B, U0)
                         : RK  ? prefetch_lane_ptr<3>(lane, it.y, cmax, d.off, H, B, RK == RKM_FIRST ? nullptr : rkf.S2in)
                               : prefetch_lane_ptr<2>(lane, it.y, cmax, d.off, H, B, nullptr);
        m.pf_lane = q != nullptr;
        m.pfp = (m.pf_lane ? q : H + d.off) + (long long)rc * d.ld;  // advanced with hp below, then ODINN_L2PF_ROWS rows further
    }
    if (RK) {
        // the planes of the stage epilogue that are not prefetched above: est, and u when it is not the S2 input (stage 0 reads S2in = u)
        constexpr bool use_u = RK == RKM_MID_U || RK == RKM_LAST;
        const float* pu = (use_u && rkf.u != rkf.S2in) ? rkf.u : nullptr;
        const float* q2 = prefetch_lane_ptr<2>(lane, it.y, cmax, d.off, RK == RKM_FIRST ? nullptr : rkf.est, pu, nullptr);
        m.pf2_lane = q2 != nullptr;
        m.pfp2 = (m.pf2_lane ? q2 : H + d.off) + (long long)rc * d.ld;
        m.rk = rkf;
        const double hh = rkf.st[it.x].h;
        m.rbh = (float)(rkf.b * hh);
        m.reh = (float)(rkf.e * hh);
        m.oo = (int)d.off + ic + (r0 - 1) * d.ld;   // (fp32 planes stay below 2^31 elements: odinn_ensemble_create)
        m.nw = mk2((m.store_pair || m.store_x) ? 1.0f : 0.0f, m.store_pair ? 1.0f : 0.0f);
        m.n_s2 = m.n_e = m.n_u = bc2(0.0f);
        m.r_s2 = m.r_e = m.r_u = m.nacc = bc2(0.0f);
    }
    m.ap = AFIELD ? Af + d.off + ic + (long long)min(rc, d.ny - 2) * d.ld : nullptr;
    m.op = dH + d.off + ic + (long long)(r0 - 1) * d.ld;  // dereferenced for rows >= r0 only
    m.up = STAGE ? U0 + d.off + ic + (long long)(r0 - 1) * d.ld : nullptr;
    m.sa = bc2(sa);
    m.sb = bc2(sb);
    m.sdt = bc2(sdt);
    m.hraw = bc2(0.0f);

    // ---- cell row r0-1 ----
    {
        f2 hv = ldg2(m.hp), bv = ldg2(m.bp);
        m.h = max2(hv, bc2(0.0f));
        m.b = bv;
    }
    m.eh = ETA1 ? m.h : mul2(bc2(m.eta0), m.h);
    {
        f2 hE = east2(m.h), bE = east2(m.b);
        m.ex = sdiff2(bE, m.b, hE, m.h);
        m.hx = add2(m.h, hE);
        m.ehE = ETA1 ? hE : mul2(bc2(m.eta0), hE);
    }
    m.Dp = bc2(0.0f);
    m.Fy = bc2(0.0f);
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        if (r0 + k >= 1 && r0 + k <= m.nym1) { m.hp += d.ld; m.bp += d.ld; m.pfp += d.ld; }
        m.hq[k] = ldg2(m.hp);
        m.bq[k] = ldg2(m.bp);
        if (RK) { if (r0 + k >= 1 && r0 + k <= m.nym1) m.pfp2 += d.ld; }
    }
    m.pfp += (long long)ODINN_L2PF_ROWS * d.ld;
    if (RK) m.pfp2 += (long long)ODINN_L2PF_ROWS * d.ld;

    int row = r0 - 1;
    m.template step<false, true>(row);
    ++row;
    const int main_end = min(r1, d.ny - 1 - PF);
    for (; row < min(r1, 1); ++row) m.template step<true, true>(row);
    for (; row + PF <= main_end; row += PF) ring_steps<PF>(m, row);
    for (; row < main_end; ++row) m.template step<true, false>(row);
    for (; row < r1; ++row) m.template step<true, true>(row);
    if (RK) {
        if (RK == RKM_LAST) {   // (weights: only the lanes that own columns accumulated)
            double a = (double)m.nacc.x + (double)m.nacc.y;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
            if (lane == 0) partial[item] = a;
        }
    }
}

// --------------------------------------------------------------------------------------------
// A1 + A2 (see VjpMarch; every quantity is a pair of adjacent columns)
// --------------------------------------------------------------------------------------------
// Clamp sub-gradient for one column (see subgrad<> in sia2d_march.cuh), fp32 comparisons on raw differences.
// (a > b && c != d) ? v : 0  as  FSETP, FSETP.AND, FSEL  (the compiler's own choice is FSETP, FSEL, FSETP, FSEL)
__device__ __forceinline__ float sel_gt_ne(float v, float a, float b, float c, float d) {
#ifdef ODINN_NO_SELP_ASM
    return (a > b && c != d) ? v : 0.0f;
#else
    float r;
    asm("{\n\t.reg .pred p, q;\n\tsetp.gt.f32 p, %1, %2;\n\tsetp.neu.and.f32 q, %3, %4, p;\n\tselp.f32 %0, %5, 0f00000000, q;\n\t}"
        : "=f"(r) : "f"(a), "f"(b), "f"(c), "f"(d), "f"(v));
    return r;
#endif
}
// ((a > b) && (c != d)) ? 1.0f : 0.0f  as  FSETP + FSET.BF  (two ALU-pipe instructions; the product with the cotangent
// then rides on a packed FFMA2 instead of an FSEL per column)
__device__ __forceinline__ float mask_gt_ne(float a, float b, float c, float d) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\tset.neu.and.f32.f32 %0, %3, %4, p;\n\t}"
        : "=f"(r) : "f"(a), "f"(b), "f"(c), "f"(d));
    return r;
}
__device__ __forceinline__ float mask_gt(float a, float b) {
    float r;
    asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// Clamp sub-gradient (inversion_utils.jl:22-43) as weights: the cotangent dC of the clamped edge slope sends
// -ml * dC to the lower cell and +mu * dC to the upper cell (strict inequalities: ties send nothing).
template <bool ETA1>
__device__ __forceinline__ void submask1(float e, float lo, float up, float eta0, float& ml, float& mu) {
    if (ETA1) {
        ml = mask_gt_ne(up, e, e, lo);
        mu = mask_gt_ne(e, lo, e, up);
    } else {
        const float in = (up > e && e > lo) ? 1.0f : 0.0f;
        ml = in + ((lo > e) ? eta0 : 0.0f);
        mu = in + ((e > up) ? eta0 : 0.0f);
    }
}
template <bool ETA1>
__device__ __forceinline__ void submask2(f2 e, f2 lo, f2 up, float eta0, f2& ml, f2& mu) {
    submask1<ETA1>(e.x, lo.x, up.x, eta0, ml.x, mu.x);
    submask1<ETA1>(e.y, lo.y, up.y, eta0, ml.y, mu.y);
}

template <bool ETA1>
__device__ __forceinline__ void subgrad1(float dC, float e, float lo, float up, float eta0, float& to_lower, float& to_upper) {
    if (ETA1) {
        to_lower = sel_gt_ne(-dC, up, e, e, lo);
        to_upper = sel_gt_ne(dC, e, lo, e, up);
    } else {
        bool inside = (up > e) && (e > lo);
        float pass = inside ? dC : 0.0f;
        float edC = eta0 * dC;
        to_lower = -pass - ((lo > e) ? edC : 0.0f);
        to_upper = pass + ((e > up) ? edC : 0.0f);
    }
}
template <bool ETA1>
__device__ __forceinline__ void subgrad2(f2 dC, f2 e, f2 lo, f2 up, float eta0, f2& to_lower, f2& to_upper) {
    subgrad1<ETA1>(dC.x, e.x, lo.x, up.x, eta0, to_lower.x, to_upper.x);
    subgrad1<ETA1>(dC.y, e.y, lo.y, up.y, eta0, to_lower.y, to_upper.y);
}

// WRITE_F: the same pass also writes dH = SIA2D(H) (F1): every forward intermediate is recomputed here anyway, so the
// forward costs one more shuffle, ~8 packed operations and one store per row instead of a second pass over H and B.
// SEED (with WRITE_H): the reverse time step of the discrete adjoint folded into the A1 epilogue -- instead of storing
// v = (dSIA/dH)^T lambda the pass writes  lambda_new = lambda + dt v + cseed W (H - H_ref)  (gradient.jl:242 with the LossH seed of
// Losses.jl:270-291) to a SECOND lambda plane (the neighbours still read the old one) and accumulates the loss term sum W (H - H_ref)^2
// per strip: 5 reads + 1 write per cell instead of the 4 + 6 words of an A1 pass followed by a loss / seed pass.
// RKA (with WRITE_H): one RDPK3Sp35 stage of the continuous adjoint's reverse ODE  d(lambda)/d(tau) = (dSIA/dH)^T lambda  at H_itp(-tau)
// (gradient.jl:316-324) in ONE pass: H is the linear interpolant  la0 Ha + la1 Hb  of two forward snapshots formed when a row leaves the
// prefetch queues (no interpolated plane is ever written), and the stage update of RkFuse (common.cuh) is the epilogue -- instead of an
// interpolation pass (3 words), the A1 pass (4) and an elementwise stage pass (7-8): 9-10 words per cell and stage.
template <bool CUBIC, bool AFIELD, bool WRITE_H, bool WRITE_S, bool ETA1, bool WRITE_F = false, bool SEED = false, bool RKA = false>
struct VjpMarch2 {
    static constexpr int PF = ODINN_PF2_VJP;
    // Every plane shares one (offset, pitch) layout, so the kernel keeps 32-bit ELEMENT offsets and adds them to the
    // (warp-uniform) plane bases at each access -- one IMAD.WIDE per access, as a pointer bump would cost, but 3
    // registers instead of 14 for the 7 pointers (the kernel sits at the 128-register cap of 4 CTAs/SM).
    const float *Hb, *Bb, *Lb, *Ab;
    float *Ob, *Vb, *Fb;
    const float *Rb, *Wb;   // SEED: H_ref and W planes
    f2 sdt, scs, lossacc;   // SEED: dt, cseed (broadcast), loss accumulator
    // RKA: second snapshot plane + its queue, interpolation weights, the stage (planes, coefficients, flags), the glacier's (b h, e h),
    // the operands of the output row (raw S1, S2, est, u), error-norm accumulator and weights, two more L2-prefetch bases
    const float* H2b;
    f2 hq2[ODINN_PF2_VJP];
    f2 la0, la1;
    RkFuse<float> rk;
    float rbh, reh;
    f2 r_s1, r_s2, r_e, r_u, nacc, nw;
    const float *pf2, *pf3;
    const float* pfb;  // per lane: base of the plane this lane prefetches + its sector's column delta + ODINN_L2PF_ROWS rows
    int oin, oout, oa;  // offsets of the row being loaded / the row being written / the A-field node row
    int ld, nym1, ny2;
    float eta0;
    f2 hdx, hdy, nhx2, nhy2, qx, qy, A;
    f2 lmask;     // 1 on inner columns, 0 on border / out-of-grid columns (λ_inn zero-extension)
    f2 nodemask;  // 1 where the column carries a dual node (0 <= c <= nx-2)
    bool store_pair, store_x, own_lane, vstore_pair, vstore_x;
    PhysDev<float> ph;
    f2 h, b, l, eh, ex, hx, ehE, fxr, px, Dp, aDp, Pp, Qrow_p, yu_p, acc;
    f2 cx, Fyp;  // WRITE_F: clamped x-edge slope of the carried row (replaces the px carry), previous y-edge flux
    // CUBIC form (compute_cubic): folded constants and its own carried set
    f2 hdxs, hdys;   // ½/Δx·√K, ½/Δy·√K  with K = Γ/4⁵, so that u² + v² = K |∇S|²
    f2 lmx, lmy;     // lmask·(-½/Δx²), lmask·(-½/Δy²): λ is scaled once when it is loaded
    f2 qx2, qy2;     // 2K·¼/Δx², 2K·¼/Δy²
    f2 ly, fx, Qp;   // scaled λ row (y form), scaled Fx† of the carried row, Q of the previous node row
    f2 hq[PF], bq[PF], lq[PF];


    // Output of row `oout`: res = (dSIA/dH)^T lambda, or -- SEED -- the reverse time step built on it.
    // SEED operands of the row this step emits (raw lambda, raw H, H_ref, W), loaded at the TOP of the step so that the whole step
    // hides their latency (lambda and H were read two rows ago by this warp: L1 / L2 hits; H_ref and W are new DRAM streams,
    // L2-prefetched ODINN_L2PF_ROWS ahead by seed_loads).
    f2 s_l, s_h, s_r, s_w;
    bool seed_lane;   // store_pair || store_x
    const float* pfs;   // per lane: H_ref (lanes 0-8) or W (lanes 9-17) sector of the row ODINN_L2PF_ROWS ahead, or nullptr
    template <bool MASKED>
    __device__ __forceinline__ void seed_loads(int row) {
        if (!SEED) return;
        // (the pair that straddles the last column of an odd-nx grid reads the zero padding column: its second element contributes 0
        //  to the loss and is never stored; the packed layout, which has no padding, only takes even nx)
        if (seed_lane) { s_l = ldg2(Lb + oout); s_h = ldg2(Hb + oout); s_r = ldg2(Rb + oout); s_w = ldg2(Wb + oout); }
        if (ODINN_L2PF_ROWS > 0 && !MASKED && pfs != nullptr && row + ODINN_L2PF_ROWS <= nym1) prefetch_l2(pfs + oout);
    }
    __device__ __forceinline__ void emit(f2 res) {
        if (RKA) {   // res = k = (dSIA/dH)^T S1 at H_itp: the stage update of rk_stage / rk_stage1_main (rdpk.cu), straight-line
            const f2 k = res, s1 = r_s1;
            f2 s1n, er;
            if (rk.flags & RKF_FIRST) {
                s1n = mk2(fmaf(rbh, k.x, s1.x), fmaf(rbh, k.y, s1.y));
                er = mk2(reh * k.x, reh * k.y);
            } else {
                const f2 s2 = mk2(fmaf(rk.d, s1.x, r_s2.x), fmaf(rk.d, s1.y, r_s2.y));
                f2 v = mk2(fmaf(rk.g2, s2.x, rk.g1 * s1.x), fmaf(rk.g2, s2.y, rk.g1 * s1.y));
                if (rk.flags & RKF_U) v = mk2(fmaf(rk.g3, r_u.x, v.x), fmaf(rk.g3, r_u.y, v.y));
                s1n = mk2(fmaf(rbh, k.x, v.x), fmaf(rbh, k.y, v.y));
                er = mk2(fmaf(reh, k.x, r_e.x), fmaf(reh, k.y, r_e.y));
                if (rk.flags & RKF_WS2) {
                    if (store_pair) *reinterpret_cast<float2*>(rk.S2out + oout) = s2;
                    if (store_x) rk.S2out[oout] = s2.x;
                }
            }
            if (store_pair) *reinterpret_cast<float2*>(Ob + oout) = s1n;
            if (store_x) Ob[oout] = s1n.x;
            if (rk.flags & RKF_WEST) {
                if (store_pair) *reinterpret_cast<float2*>(rk.est + oout) = er;
                if (store_x) rk.est[oout] = er.x;
            }
            if (rk.flags & RKF_NORM) {
                const float rx = __fdividef(er.x, fmaf(rk.reltol, fmaxf(fabsf(r_u.x), fabsf(s1n.x)), rk.abstol));
                const float ry = __fdividef(er.y, fmaf(rk.reltol, fmaxf(fabsf(r_u.y), fabsf(s1n.y)), rk.abstol));
                nacc.x = fmaf(nw.x * rx, rx, nacc.x);
                nacc.y = fmaf(nw.y * ry, ry, nacc.y);
            }
            return;
        }
        if (SEED) {
            if (!seed_lane) return;
            const f2 df = sub2(s_h, s_r), wd = mul2(s_w, df);
            lossacc = fma2(wd, df, lossacc);
            res = fma2(sdt, res, fma2(scs, wd, s_l));
        }
        if (store_pair) *reinterpret_cast<float2*>(Ob + oout) = res;
        if (store_x) Ob[oout] = res.x;
    }

    template <bool OUT, bool MASKED, int SLOT = -1>
    __device__ __forceinline__ void step(int row) {
        constexpr int RS = SLOT < 0 ? 0 : SLOT;
        constexpr int WS = SLOT < 0 ? PF - 1 : SLOT;
        f2 h1 = hq[RS], b1 = bq[RS], l1 = lq[RS];
        if (RKA) h1 = fma2(la1, hq2[RS], mul2(la0, h1));   // H_itp = (1 - a) Ha + a Hb   (rk_lerp of rdpk.cu)
        if (SLOT < 0) {
#pragma unroll
            for (int k = 0; k + 1 < PF; ++k) { hq[k] = hq[k + 1]; bq[k] = bq[k + 1]; lq[k] = lq[k + 1]; if (RKA) hq2[k] = hq2[k + 1]; }
        }
        if (MASKED) oin += (row + 1 + PF <= nym1) ? ld : 0;
        else oin += ld;
        hq[WS] = ldg2(Hb + oin);
        if (RKA) hq2[WS] = ldg2(H2b + oin);
        bq[WS] = ldg2(Bb + oin);
        lq[WS] = ldg2(Lb + oin);
        if (ODINN_L2PF_ROWS > 0 && !MASKED) {
            if (row + 1 + PF + ODINN_L2PF_ROWS <= nym1) {
                prefetch_l2(pfb + oin);
                if (RKA) {
                    prefetch_l2(pf2 + oin);
                    if (pf3 != nullptr) prefetch_l2(pf3 + oin);
                }
            }
        }
        if (RKA && WRITE_H && OUT) {
            // operands of the row this step emits: every lane loads (the clamped pair index keeps the address inside the grid; plain loads:
            // S2 and est are rewritten in place by this thread)
            r_s1 = *reinterpret_cast<const float2*>(Lb + oout);
            if (!(rk.flags & RKF_FIRST)) {
                r_s2 = *reinterpret_cast<const float2*>(rk.S2in + oout);
                r_e = *reinterpret_cast<const float2*>(rk.est + oout);
            }
            if (rk.flags & (RKF_U | RKF_NORM)) r_u = *reinterpret_cast<const float2*>(rk.u + oout);
        }
        f2 Anode = A;
        if (AFIELD) {
            Anode = ldg2(Ab + oa);
            if (MASKED) { if (row >= 0 && row < ny2) oa += ld; } else oa += ld;
        }
        if (OUT) seed_loads<MASKED>(row);
        if (CUBIC) compute_cubic<OUT, MASKED>(row, h1, b1, l1, Anode);
        else compute<OUT, MASKED>(row, h1, b1, l1, Anode);
    }

    // n = 3, C = 0 form of compute(): the same operator with the constants folded and the per-node products shared.
    //   g2 = K|∇S|² (K folded into the slope scale), m = g2·Hs, z = Hs⁴·D†, zA = A z:
    //   D = Hs⁴·(A m),  α D† = 5 g2 zA,  β D† ∇S-part = 2K (Hs zA)·(raw slope sums),  ∂A_spatial ∘ D† = m z
    // λ is scaled by -½/Δ² (and the column mask) when loaded, so D† is a plain sum of four products and the clamp
    // cotangents need no further scaling; the clamp sub-gradient is applied as 0/1 weights (submask2) through packed FMAs.
    // Carried: h, b, eh, ex, hx, ehE, cx, fx, ly, px, Dp, aDp, Pp, Qp, yu_p, Fyp, acc.
    template <bool OUT, bool MASKED>
    __device__ __forceinline__ void compute_cubic(int row, f2 h1, f2 b1, f2 l1, f2 Anode) {
        f2 l1x = mul2(l1, lmx), l1y = mul2(l1, lmy);
        if (MASKED) { if (!(row >= 0 && row + 1 < nym1)) l1x = l1y = bc2(0.0f); }  // λ_inn zero-extended on border rows
        h1 = max2(h1, bc2(0.0f));
        f2 eh1 = ETA1 ? h1 : mul2(bc2(eta0), h1);
        f2 hE1 = east2(h1), bE1 = east2(b1), lE1x = east2(l1x);
        // x-edge (i→i+1, row+1)
        f2 ex1 = sdiff2(bE1, b1, hE1, h1);
        f2 hx1 = add2(h1, hE1);
        f2 ehE1 = ETA1 ? hE1 : mul2(bc2(eta0), hE1);
        f2 fx1 = sub2(lE1x, l1x);                         // -½/Δx² · Fx† (adjoint.jl:100)
        const f2 cx1 = clamp2(ex1, ehE1, eh1);
        f2 px1 = mul2(fx1, cx1);                          // (adjoint.jl:102)
        // y-edge (i, row→row+1)
        f2 ey = sdiff2(b1, b, h1, h);
        f2 fy = sub2(l1y, ly);
        const f2 cy = clamp2(ey, eh1, eh);
        f2 py = mul2(fy, cy);
        f2 eyE = east2(ey), pyE = east2(py);
        // node (i, row)
        f2 gxr = add2(ex, ex1), gyr = add2(ey, eyE);
        f2 u = mul2(gxr, hdxs), v = mul2(gyr, hdys);
        f2 g2 = fma2(v, v, mul2(u, u));
        f2 Hs = add2(hx, hx1);
        f2 H2 = mul2(Hs, Hs);
        f2 H4 = mul2(H2, H2);
        f2 mk = mul2(g2, Hs);
        f2 D1 = mul2(H4, mul2(mk, Anode));
        f2 Dadj = add2(add2(py, pyE), add2(px, px1));     // D† (adjoint.jl:102-104)
        Dadj = mul2(Dadj, nodemask);
        bool row_ok = true;
        if (MASKED) { row_ok = (row >= 0 && row < nym1); if (!row_ok) Dadj = bc2(0.0f); }
        f2 z = mul2(H4, Dadj);
        f2 zA = mul2(z, Anode);
        f2 aD1 = mul2(g2, zA);
        f2 bD = mul2(Hs, zA);
        f2 P1 = mul2(bD, gxr);
        f2 Q1 = mul2(bD, gyr);
        if (WRITE_S) {
            if (OUT) {
                if (AFIELD) {
                    f2 vS = mul2(mk, z);                  // ∂A_spatial ∘ D† (adjoint.jl:250)
                    acc = add2(acc, vS);                  // (halo lanes are dropped when the strip is reduced)
                    if (row_ok) {
                        if (vstore_pair) *reinterpret_cast<float2*>(Vb + oout) = vS;
                        if (vstore_x) Vb[oout] = vS.x;
                    }
                } else {
                    acc = fma2(mk, z, acc);
                }
            }
        }
        f2 D1W;
        if (WRITE_H || WRITE_F) D1W = west2(D1);
        if (WRITE_F) {
            // F1 with the operation order of RhsMarch2::compute
            f2 Fy1 = mul2(add2(D1W, D1), cy);
            f2 Fx = mul2(add2(Dp, D1), cx);
            f2 FxW = west2(Fx);
            if (OUT) {
                f2 outv = fma2(lmy, sub2(Fyp, Fy1), mul2(lmx, sub2(FxW, Fx)));
                if (MASKED) { if (row < 1 || row >= nym1) outv = bc2(0.0f); }
                if (store_pair) *reinterpret_cast<float2*>(Fb + oout) = outv;
                if (store_x) Fb[oout] = outv.x;
            }
            Fyp = Fy1;
        }
        if (WRITE_H) {
            f2 dCy = mul2(fy, add2(D1W, D1));             // ∂Cy/Δy = -Fy†·Dy/Δy
            f2 dCx = mul2(fx, add2(Dp, D1));
            f2 myl, myu, mxl, mxu;
            submask2<ETA1>(ey, neg2(eh), eh1, eta0, myl, myu);
            submask2<ETA1>(ex, neg2(eh), ehE, eta0, mxl, mxu);
            f2 SAW = fma2(sub2(Qp, Q1), qy2, mul2(add2(aDp, aD1), bc2(5.0f)));
            f2 SP = mul2(add2(Pp, P1), qx2);
            f2 ZW = west2(fma2(mxu, dCx, add2(SAW, SP)));  // everything column i-1 sends to cell (i, row)
            f2 yu1 = mul2(myu, dCy);
            if (OUT) {
                f2 own = fma2(myl, dCy, fma2(mxl, dCx, sub2(SP, SAW)));   // minus the cell's own share
                f2 res = sub2(ZW, sub2(own, yu_p));
                res = mul2(res, mk2(mask_gt(h.x, 0.0f), mask_gt(h.y, 0.0f)));  // adjoint.jl:148
                emit(res);
            }
            yu_p = yu1;
        }
        oout += ld;
        h = h1; b = b1; eh = eh1; ex = ex1; hx = hx1; ehE = ehE1; cx = cx1;
        fx = fx1; ly = l1y; px = px1;
        Dp = D1; aDp = aD1; Pp = P1; Qp = Q1;
    }

    // One marching step given the cell row `row+1` (h1, b1, raw λ row l1) and the node coefficient of node row `row`.
    template <bool OUT, bool MASKED>
    __device__ __forceinline__ void compute(int row, f2 h1, f2 b1, f2 l1, f2 Anode) {
        l1 = mul2(l1, lmask);
        if (MASKED) { if (!(row >= 0 && row + 1 < nym1)) l1 = bc2(0.0f); }  // λ_inn zero-extended on border rows
        h1 = max2(h1, bc2(0.0f));
        f2 eh1 = ETA1 ? h1 : mul2(bc2(eta0), h1);
        f2 hE1 = east2(h1), bE1 = east2(b1), lE1 = east2(l1);
        // x-edge (i→i+1, row+1)
        f2 ex1 = sdiff2(bE1, b1, hE1, h1);
        f2 hx1 = add2(h1, hE1);
        f2 ehE1 = ETA1 ? hE1 : mul2(bc2(eta0), hE1);
        f2 fxr1 = sub2(lE1, l1);                          // raw Fx† (adjoint.jl:100)
        const f2 cx1 = clamp2(ex1, ehE1, eh1);
        f2 px1 = mul2(fxr1, cx1);                         // raw Fx†·clamp(dSdx) (adjoint.jl:102)
        if (WRITE_F) px = mul2(fxr, cx);                  // (same product as the px1 of the previous step)
        // y-edge (i, row→row+1)
        f2 ey = sdiff2(b1, b, h1, h);
        f2 fyr = sub2(l1, l);
        const f2 cy = clamp2(ey, eh1, eh);
        f2 py = mul2(fyr, cy);
        f2 eyE = east2(ey), pyE = east2(py);
        // node (i, row)
        f2 gxr = add2(ex, ex1), gyr = add2(ey, eyE);
        f2 u = mul2(gxr, hdx), v = mul2(gyr, hdy);
        f2 D1, al, be, gA;
        node_raw2<CUBIC, true>(ph, Anode, add2(hx, hx1), fma2(v, v, mul2(u, u)), D1, al, be, gA);
        f2 Dadj = fma2(add2(py, pyE), nhy2, mul2(add2(px, px1), nhx2));  // D† (adjoint.jl:102-104)
        Dadj = mul2(Dadj, nodemask);
        bool row_ok = true;
        if (MASKED) { row_ok = (row >= 0 && row < nym1); if (!row_ok) Dadj = bc2(0.0f); }
        f2 bD = mul2(be, Dadj);
        f2 aD1 = mul2(al, Dadj);
        f2 P1 = mul2(bD, gxr);
        f2 Q1 = mul2(bD, gyr);
        if (WRITE_S) {
            if (OUT) {
                f2 vS = mul2(gA, Dadj);                   // ∂A_spatial ∘ D† (adjoint.jl:250)
                if (own_lane) acc = add2(acc, vS);
                if (AFIELD) {
                    if (row_ok) {
                        if (vstore_pair) *reinterpret_cast<float2*>(Vb + oout) = vS;
                        if (vstore_x) Vb[oout] = vS.x;
                    }
                }
            }
        }
        f2 D1W;
        if (WRITE_H || WRITE_F) D1W = west2(D1);
        if (WRITE_F) {
            // F1 with the operation order of RhsMarch2::compute (nhx2 = -kx, nhy2 = -ky: exact sign flips)
            f2 Fy1 = mul2(add2(D1W, D1), cy);
            f2 Fx = mul2(add2(Dp, D1), cx);
            f2 FxW = west2(Fx);
            if (OUT) {
                f2 outv = mul2(fma2(nhy2, sub2(Fyp, Fy1), mul2(nhx2, sub2(FxW, Fx))), lmask);
                if (MASKED) { if (row < 1 || row >= nym1) outv = bc2(0.0f); }
                if (store_pair) *reinterpret_cast<float2*>(Fb + oout) = outv;
                if (store_x) Fb[oout] = outv.x;
            }
            Fyp = Fy1;
            cx = cx1;
        }
        if (WRITE_H) {
            f2 Q1W = west2(Q1);
            f2 Qrow1 = add2(Q1W, Q1);
            f2 yl, yu1, xl, xu;
            {
                f2 dC = mul2(mul2(fyr, nhy2), add2(D1W, D1));  // ∂Cy/Δy = -Fy†·Dy/Δy
                subgrad2<ETA1>(dC, ey, neg2(eh), eh1, eta0, yl, yu1);
            }
            {
                f2 dC = mul2(mul2(fxr, nhx2), add2(Dp, D1));
                subgrad2<ETA1>(dC, ex, neg2(eh), ehE, eta0, xl, xu);
            }
            f2 aDc = mul2(bc2(0.25f), add2(aDp, aD1));
            f2 Pc = mul2(qx, add2(Pp, P1));
            f2 ZW = west2(add2(add2(aDc, Pc), xu));  // everything column i-1 sends to cell (i, row)
            if (OUT) {
                f2 res = add2(add2(add2(ZW, add2(sub2(aDc, Pc), xl)), mul2(qy, sub2(Qrow_p, Qrow1))), add2(yl, yu_p));
                if (!(h.x > 0.0f)) res.x = 0.0f;  // adjoint.jl:148
                if (!(h.y > 0.0f)) res.y = 0.0f;
                emit(res);
            }
            Qrow_p = Qrow1;
            yu_p = yu1;
        }
        oout += ld;
        h = h1; b = b1; l = l1; eh = eh1; ex = ex1; hx = hx1; ehE = ehE1; fxr = fxr1;
        if (!WRITE_F) px = px1;
        Dp = D1; aDp = aD1; Pp = P1;
    }
};

// RKA: H2 is the upper snapshot; the glacier's interpolation weight is a = (sign (t_g + lc h_g) - lta) / (ltb - lta) from its controller state.
template <bool CUBIC, bool AFIELD, bool WRITE_H, bool WRITE_S, bool ETA1, bool WRITE_F = false, bool SEED = false, bool RKA = false>
__global__ void __launch_bounds__(MARCH2_WARPS * 32, ODINN_VJP2_MIN_CTAS)
sia2d_vjp_march2(const GDesc<float>* __restrict__ descs, const int4* __restrict__ items, int n_items,
                 const float* __restrict__ lam, const float* __restrict__ H, const float* __restrict__ B,
                 const float* __restrict__ Af, float* __restrict__ out, float* __restrict__ vjpA,
                 double* __restrict__ partial, PhysDev<float> ph, float* __restrict__ dH = nullptr,
                 const float* __restrict__ Href = nullptr, const float* __restrict__ Wm = nullptr, float sdt = 0.0f, float scs = 0.0f,
                 RkFuse<float> rkf = RkFuse<float>(), const float* __restrict__ H2 = nullptr, double lc = 0.0, double lsign = 1.0,
                 double lta = 0.0, double ltb = 1.0) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * MARCH2_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int4 it = items[item];
    const GDesc<float> d = descs[it.x];
    const int c0 = it.y + 2 * lane, r0 = it.z, r1 = it.w;
    const int cmax = (d.nx - 1) & ~1;
    const int ic = min(max(c0, 0), cmax);
    VjpMarch2<CUBIC, AFIELD, WRITE_H, WRITE_S, ETA1, WRITE_F, SEED, RKA> m;
    constexpr int PF = ODINN_PF2_VJP;
    m.H2b = H2;
    m.pf2 = m.pf3 = nullptr;
    if (RKA) {
        const RkState st = rkf.st[it.x];
        // the glacier has landed on the stop: nothing to integrate (rk_integrate commits by copy then).  RKF_LERP_ONLY: the A2 pass at a
        // quadrature node (WRITE_S only: no stage, every glacier sits on the node) uses the interpolation on load alone.
        if (st.done && !(rkf.flags & RKF_LERP_ONLY)) return;
        const double tt = lsign * (st.t + lc * st.h);
        const float a1 = (float)((tt - lta) / (ltb - lta));
        m.la1 = bc2(a1);
        m.la0 = bc2(1.0f - a1);
        m.rk = rkf;
        m.rbh = (float)(rkf.b * st.h);
        m.reh = (float)(rkf.e * st.h);
        m.r_s1 = m.r_s2 = m.r_e = m.r_u = m.nacc = bc2(0.0f);
    }
    m.Rb = Href; m.Wb = Wm;
    m.sdt = bc2(sdt); m.scs = bc2(scs); m.lossacc = bc2(0.0f);
    m.s_l = m.s_h = m.s_r = m.s_w = bc2(0.0f);
    m.pfs = nullptr;
    m.seed_lane = false;
    m.ph = ph;
    m.ld = d.ld;
    m.nym1 = d.ny - 1;
    m.ny2 = d.ny - 2;
    m.eta0 = ph.eta0;
    const float hdx = 0.5f * d.inv_dx, hdy = 0.5f * d.inv_dy;
    m.hdx = bc2(hdx);
    m.hdy = bc2(hdy);
    m.nhx2 = bc2(-hdx * d.inv_dx);  // -½/Δx²
    m.nhy2 = bc2(-hdy * d.inv_dy);
    m.qx = bc2(hdx * hdx);          // ¼/Δx²
    m.qy = bc2(hdy * hdy);
    m.A = bc2(d.A);
    const int c1 = c0 + 1;
    m.lmask = mk2((c0 >= 1 && c0 <= d.nx - 2) ? 1.0f : 0.0f, (c1 >= 1 && c1 <= d.nx - 2) ? 1.0f : 0.0f);
    m.nodemask = mk2((c0 >= 0 && c0 <= d.nx - 2) ? 1.0f : 0.0f, (c1 >= 0 && c1 <= d.nx - 2) ? 1.0f : 0.0f);
    const bool out_lane = (lane >= 1 && lane <= 30 && c0 >= 0);
    m.store_pair = out_lane && (c1 < d.nx);
    m.store_x = out_lane && (c1 == d.nx);
    m.own_lane = (lane >= 1 && lane <= 30);
    m.seed_lane = m.store_pair || m.store_x;
    m.nw = mk2(m.seed_lane ? 1.0f : 0.0f, m.store_pair ? 1.0f : 0.0f);
    m.vstore_pair = out_lane && (c1 <= d.nx - 2);
    m.vstore_x = out_lane && (c1 == d.nx - 1);
    const int rc = max(r0 - 1, 0);
    m.Hb = H; m.Bb = B; m.Lb = lam; m.Ab = Af;
    m.Ob = out; m.Fb = dH; m.Vb = vjpA;
    const int o0 = (int)d.off + ic;  // (odinn_ensemble_create refuses planes of 2^31 elements or more)
    m.oin = o0 + rc * d.ld;
    m.oa = o0 + min(rc, d.ny - 2) * d.ld;
    m.oout = o0 + (r0 - 1) * d.ld;   // dereferenced for rows >= r0 only
    {
        // one prefetch instruction per row covers the three input planes: 10 lanes per plane, one 32-byte sector each
        // (the 256-byte row segment of a strip spans up to 9 sectors; lanes 30, 31 touch the next strip's first sectors)
        const int pl = min(lane / 10, 2), sec = lane - 10 * pl;
        const float* pb = pl == 0 ? H : (pl == 1 ? B : lam);
        m.pfb = pb + (min(max(it.y + 8 * sec, 0), cmax) - ic) + (long long)ODINN_L2PF_ROWS * d.ld;
        if (RKA) {   // H2, and the epilogue planes S2in, est (10 lanes each); u (when it is read and is not the S2 input) on a third instruction
            const float* pb2 = pl == 0 ? H2 : (pl == 1 ? ((rkf.flags & RKF_FIRST) ? nullptr : rkf.S2in) : ((rkf.flags & RKF_FIRST) ? nullptr : rkf.est));
            const long long dl = (min(max(it.y + 8 * sec, 0), cmax) - ic) + (long long)ODINN_L2PF_ROWS * d.ld;
            if (pb2 != nullptr) m.pf2 = pb2 + dl; else m.pf2 = H2 + dl;
            if ((rkf.flags & (RKF_U | RKF_NORM)) && rkf.u != rkf.S2in && lane < 10) m.pf3 = rkf.u + dl;
        }
        if (SEED && lane < 18) {
            const int sp = lane / 9, ss = lane - 9 * sp;
            m.pfs = (sp == 0 ? Href : Wm) + (min(max(it.y + 8 * ss, 0), cmax) - ic) + (long long)ODINN_L2PF_ROWS * d.ld;
        }
    }

    // ---- cell row r0-1 ----
    {
        f2 hv = ldg2(m.Hb + m.oin), bv = ldg2(m.Bb + m.oin), lv = ldg2(m.Lb + m.oin);
        if (RKA) hv = fma2(m.la1, ldg2(m.H2b + m.oin), mul2(m.la0, hv));
        m.h = max2(hv, bc2(0.0f));
        m.b = bv;
        m.l = mul2(lv, m.lmask);
        if (!(r0 >= 2 && r0 <= m.nym1)) m.l = bc2(0.0f);  // row r0-1 must be an inner row
    }
    m.eh = ETA1 ? m.h : mul2(bc2(m.eta0), m.h);
    {
        f2 hE = east2(m.h), bE = east2(m.b), lE = east2(m.l);
        m.ex = sdiff2(bE, m.b, hE, m.h);
        m.hx = add2(m.h, hE);
        m.ehE = ETA1 ? hE : mul2(bc2(m.eta0), hE);
        m.fxr = sub2(lE, m.l);
        m.cx = clamp2(m.ex, m.ehE, m.eh);
        m.px = mul2(m.fxr, m.cx);
    }
    if (CUBIC) {
        const float sK = sqrtf(ph.Gam * (1.0f / 1024.0f)), K2 = 2.0f * (ph.Gam * (1.0f / 1024.0f));
        m.hdxs = bc2(hdx * sK);
        m.hdys = bc2(hdy * sK);
        m.lmx = mul2(m.lmask, m.nhx2);
        m.lmy = mul2(m.lmask, m.nhy2);
        m.qx2 = bc2(K2 * (hdx * hdx));
        m.qy2 = bc2(K2 * (hdy * hdy));
        const f2 lx = mul2(m.l, m.nhx2);   // m.l is the masked λ row r0-1
        m.ly = mul2(m.l, m.nhy2);
        m.fx = sub2(east2(lx), lx);
        m.px = mul2(m.fx, m.cx);
        m.Qp = bc2(0.0f);
    }
    m.Dp = m.aDp = m.Pp = m.Qrow_p = m.yu_p = m.acc = m.Fyp = bc2(0.0f);
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        if (r0 + k >= 1 && r0 + k <= m.nym1) m.oin += d.ld;
        m.hq[k] = ldg2(m.Hb + m.oin);
        if (RKA) m.hq2[k] = ldg2(m.H2b + m.oin);
        m.bq[k] = ldg2(m.Bb + m.oin);
        m.lq[k] = ldg2(m.Lb + m.oin);
    }

    int row = r0 - 1;
    m.template step<false, true>(row);
    ++row;
    const int main_end = min(r1, d.ny - 1 - PF);
    for (; row < min(r1, 1); ++row) m.template step<true, true>(row);
#pragma unroll VJP_RING_UNROLL
    for (; row + PF <= main_end; row += PF) ring_steps<PF>(m, row);
    for (; row < main_end; ++row) m.template step<true, false>(row);
    for (; row < r1; ++row) m.template step<true, true>(row);

    if (RKA) {
        if (rkf.flags & RKF_NORM) {   // (weights: only the lanes that own columns accumulated)
            double a = (double)m.nacc.x + (double)m.nacc.y;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
            if (lane == 0) partial[item] = a;
        }
    }
    if (SEED) {   // (exclusive with WRITE_S: the strip's loss term goes where S would)
        double a = (double)m.lossacc.x + (double)m.lossacc.y;   // (only storing lanes accumulated)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
        if (lane == 0) partial[item] = a;
    }
    if (WRITE_S) {
        double a = m.own_lane ? (double)m.acc.x + (double)m.acc.y : 0.0;  // (the CUBIC form accumulates in the halo lanes too)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(FULL, a, s);
        if (lane == 0) partial[item] = a;
    }
}

}  // namespace odinn
