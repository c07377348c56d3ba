// The phase-1 machinery shared by the register-patch continuous_conv kernels k_cconv_lean (cconv_lean.cu) and
// k_cconv_apatch (cconv_apatch.cu): one warp per out point, lane = input channel, the trilinear patch of the point in
// registers.  See cconv_lean.cu for the reasoning behind each piece.
#pragma once
#include <type_traits>
#include "cconv_scatter.cuh"

namespace dmcf {

namespace lean {

static constexpr int kGatherSlots = 4;    // feature rows in flight per warp (ring of 128-byte slots, 512-byte aligned)
static constexpr int kMetaSlots = 40;     // int2 {byte offset of pair j+4 (or -1), base cell of pair j+1}
static constexpr int kHeadWords = 8;      // offsets of pairs 0..3
static constexpr int kWgtSlots = 36;      // 8 corner weights per pair
static constexpr int kRecWords = 2 * kMetaSlots + kHeadWords + 8 * kWgtSlots;  // 376 words per warp
static constexpr int kFilterSlots = 3;    // phase 2: filter k-quads in flight per warp (slots of 128 words): slot 0 is the
                                          // warp's gather ring, slots 1..2 the head of its record block
static constexpr int kScratchWords = kGatherSlots * 32 + kRecWords;            // per warp, both phases
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gptr));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
// patch row of one point: cell c goes STEP words further (STEP = cin x (points per tile + 1), a compile-time constant here)
template <int STEP, int K>
__device__ __forceinline__ void store_patch_row(float* pp, const float (&acc)[K]) {
#pragma unroll
    for (int c = 0; c < K; ++c) pp[c * STEP] = acc[c];
}

// next 128-byte slot of a 512-byte aligned ring of kGatherSlots slots
__device__ __forceinline__ uint32_t ring_next(uint32_t saddr) {
    return (saddr & ~(kGatherSlots * 128u - 1u)) | ((saddr + 128u) & (kGatherSlots * 128u - 1u));
}

// State of the walk over one chunk of compacted pair records.  Everything the current pair needs is already in
// registers (loaded while the previous pair was scattered), so no shared-memory latency sits on the per-pair chain.
struct Walk {
    const int2* m2;   // metadata entry of the NEXT pair
    const float* w;   // weights of the current pair
    int2 mm;          // metadata of the current pair j: {offset of pair j+4 or -1, base cell of pair j+1}
    float f;          // gathered feature of the current pair
    uint32_t sa;      // this lane's word in the ring slot of the current pair
    uint32_t sa_next; // ... and of the next pair
    float4 wa, wb;    // corner weights of the current pair
};

// Pairs of base cell C: scatter the 8 corner weights of the current pair while the feature / metadata / weights of the
// next one are fetched, and start the gather of pair j+4 into the slot of pair j.
template <class S, bool RELU, bool FX, int C>
__device__ __forceinline__ void merge_case(Walk& wk, float (&acc)[S::NACC], const char* fbase, int gate, float scale, float fc) {
    int b_next;
    do {
        float f = wk.f;
        const int2 mj = wk.mm;
        const uint32_t sj = wk.sa;
        cp_wait<kGatherSlots - 2>();  // the gather of pair j+1 has landed
        wk.sa = wk.sa_next;
        wk.f = lds_f32(wk.sa);
        wk.sa_next = ring_next(wk.sa);
        wk.mm = *wk.m2;
        ++wk.m2;
        if (RELU) f = fmaxf(f, 0.0f);
        if (FX) f = fmaf(f, scale, fc);
        S::template scatter<C>(acc, wk.wa, wk.wb, f);
        if ((mj.x | gate) >= 0) cp_async4(sj, fbase + (unsigned)mj.x);
        cp_commit();
        wk.w += 8;
        wk.wa = *reinterpret_cast<const float4*>(wk.w);
        wk.wb = *reinterpret_cast<const float4*>(wk.w + 4);
        b_next = mj.y;
    } while (b_next == C);
}

// The chunk is ordered by base cell (checked when it is built), so a cell that occurs in the chunk (`present`, a warp-
// uniform mask) is the current one when the merge reaches it: cells without pairs cost one uniform bit test.
template <class S, bool RELU, bool FX, int C0, int N>
__device__ __forceinline__ void merge_group(Walk& wk, unsigned present, float (&acc)[S::NACC], const char* fbase, int gate, float scale, float fc) {
    if constexpr (N > 0) {
        if (present & (1u << (C0 & 31))) merge_case<S, RELU, FX, C0>(wk, acc, fbase, gate, scale, fc);
        merge_group<S, RELU, FX, C0 + 1, N - 1>(wk, present, acc, fbase, gate, scale, fc);
    }
}

// One sweep over all base cells in ascending order; present[w] = bit mask of the cells 32w..32w+31 that occur in the chunk.
template <class S, bool RELU, bool FX, int W = 0>
__device__ __forceinline__ void sweep(Walk& wk, const unsigned (&present)[S::NBW], float (&acc)[S::NACC], const char* fbase,
                                      int gate, float scale, float fc) {
    if constexpr (W < S::NBW) {
        constexpr int N = S::NB - 32 * W < 32 ? S::NB - 32 * W : 32;
        if (S::NBW == 1 || present[W] != 0u) merge_group<S, RELU, FX, 32 * W, N>(wk, present[W], acc, fbase, gate, scale, fc);
        sweep<S, RELU, FX, W + 1>(wk, present, acc, fbase, gate, scale, fc);
    }
}

// Scatter policies: where the 8 corner weights of a pair with base cell C go.
// Full trilinear patch: acc[cell] for every filter cell (k_cconv_lean).
template <int KZ, int KY, int KX>
struct FullPatch {
    using G = FilterGrid<KZ, KY, KX>;
    static constexpr int NACC = G::K, NB = G::NB, NBW = (G::NB + 31) / 32;
    template <int C>
    static __device__ __forceinline__ void scatter(float (&acc)[NACC], const float4& wa, const float4& wb, float f) {
        scatter_case<KZ, KY, KX, 0, KZ, C>(acc, wa, wb, f);
    }
};

// Antisymmetric filters F[rev(cell)] = -F[cell] (rev = all three cell coordinates mirrored = linear index K-1-cell):
//   sum_cell patch[cell] . F[cell] = sum_{cell < K/2} (patch[cell] - patch[K-1-cell]) . F[cell],
// so only the folded half patch acc[t] = patch[t] - patch[K-1-t] is kept: half the registers (k_cconv_apatch).
template <int KZ, int KY, int KX>
struct AntiPatch {
    using G = FilterGrid<KZ, KY, KX>;
    static_assert(G::K % 2 == 0, "antisymmetric filter grids have an even number of cells");
    static constexpr int NACC = G::K / 2, NB = G::NB, NBW = (G::NB + 31) / 32;
    template <int CELL>
    static __device__ __forceinline__ void corner(float (&acc)[NACC], float w, float f) {
        if constexpr (CELL < NACC) acc[CELL] = fmaf(w, f, acc[CELL]);
        else acc[G::K - 1 - CELL] = fmaf(-w, f, acc[G::K - 1 - CELL]);
    }
    template <int C>
    static __device__ __forceinline__ void scatter(float (&acc)[NACC], const float4& wa, const float4& wb, float f) {
        constexpr int x0 = C % G::NBX, y0 = (C / G::NBX) % G::NBY, z0 = C / (G::NBX * G::NBY);
        constexpr int sx = 1, sy = KX, sz = KY * KX;
        constexpr int c0 = (z0 * KY + y0) * KX + x0;
        corner<c0>(acc, wa.x, f);
        if constexpr (KX > 1) corner<c0 + sx>(acc, wa.y, f);
        if constexpr (KY > 1) corner<c0 + sy>(acc, wa.z, f);
        if constexpr (KX > 1 && KY > 1) corner<c0 + sx + sy>(acc, wa.w, f);
        if constexpr (KZ > 1) {
            corner<c0 + sz>(acc, wb.x, f);
            if constexpr (KX > 1) corner<c0 + sz + sx>(acc, wb.y, f);
            if constexpr (KY > 1) corner<c0 + sz + sy>(acc, wb.z, f);
            if constexpr (KX > 1 && KY > 1) corner<c0 + sz + sx + sy>(acc, wb.w, f);
        }
    }
};

// The per-warp scratch of phase 1 (shared memory) and the per-lane constants of the gathers.
struct WarpCtx {
    int2* meta;       // [kMetaSlots]
    int* metai;       // the same as ints
    int* head;        // [kHeadWords]
    float* wgt;       // [kWgtSlots][8]
    const char* fbase;  // feature base + this lane's channel
    int gate;         // 0, or the sign bit for lanes beyond cin (they start no gathers)
    int stride_b;     // feature row stride in bytes
    int lane;
    unsigned lt_mask;
    uint32_t sa;      // this lane's word of the current gather slot (shared-window address; ring 512-byte aligned)

    __device__ __forceinline__ void init(float* ring, float* wrec, const ConvParams& p, int lane_) {
        lane = lane_;
        meta = reinterpret_cast<int2*>(wrec);
        metai = reinterpret_cast<int*>(wrec);
        head = metai + 2 * kMetaSlots;
        wgt = wrec + 2 * kMetaSlots + kHeadWords;
        const bool lane_ci = lane < p.cin;
        lt_mask = (1u << lane) - 1u;
        stride_b = (int)p.inp_stride * 4;
        gate = lane_ci ? 0 : (int)0x80000000;
        fbase = reinterpret_cast<const char*>(p.inp_feat) + 4 * (lane_ci ? lane : p.cin - 1);
        asm volatile("" : "+l"(fbase));  // keep base + lane offset as ONE 64-bit register: a gather address is one 64-bit add
        sa = (uint32_t)__cvta_generic_to_shared(ring + lane);
    }
};

// All pairs of one out point: chunks of 32 raw records -> base form -> compacted, ordered records in the warp's scratch
// -> merge walk into acc.  `cur` holds the raw records of the point's first chunk on entry (loaded by the caller, so
// that they are in flight early).  Returns the normaliser contribution of this lane.
template <class S, bool RELU, bool FX>
__device__ __forceinline__ float point_patch(const ConvParams& p, WarpCtx& cx, PairRec cur, int64_t rs, int64_t re, float ox,
                                             float oy, float oz, float fc, float (&acc)[S::NACC]) {
    using G = typename S::G;
    const int lane = cx.lane;
    float norm_acc = 0.0f;
#pragma unroll 1
    for (int64_t c0 = rs;; c0 += 32) {
        const bool last = c0 + 32 >= re;
        // ---- this chunk: raw records -> base form ----
        int row = cur.row;
        norm_acc += cur.norm;
        int b = 0;
        float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
        if (row >= 0) {
            int bx, by, bz;
            float xl, xh, yl, yh, zl, zh;
            base_axis(G::KX_, cur.g.i0 & 0xff, cur.g.wx0, cur.g.wx1, bx, xl, xh);
            base_axis(G::KY_, (cur.g.i0 >> 8) & 0xff, cur.g.wy0, cur.g.wy1, by, yl, yh);
            base_axis(G::KZ_, (cur.g.i0 >> 16) & 0xff, cur.g.wz0, cur.g.wz1, bz, zl, zh);
            b = (bz * G::NBY + by) * G::NBX + bx;
            wa = make_float4(xl * yl * zl, xh * yl * zl, xl * yh * zl, xh * yh * zl);
            wb = make_float4(xl * yl * zh, xh * yl * zh, xl * yh * zh, xh * yh * zh);
        }
        // The walk needs the chunk ordered by base cell, dropped pairs last.  dmcf_cconv_prepare writes its records in
        // that order; geometry evaluated in-kernel arrives in neighbour-list order.  Check, and sort if needed (warp
        // bitonic sort of (cell, lane), then pull the fields from the source lane).
        const unsigned cellkey = row >= 0 ? (unsigned)b : 0xffffffu;
        const unsigned prevkey = __shfl_up_sync(0xffffffffu, cellkey, 1);
        if (!__all_sync(0xffffffffu, lane == 0 || prevkey <= cellkey)) {
            unsigned key = (cellkey << 5) | (unsigned)lane;
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const unsigned other = __shfl_xor_sync(0xffffffffu, key, j);
                    const bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
                    const unsigned mn = min(key, other), mx = max(key, other);
                    key = (up == lower) ? mn : mx;
                }
            }
            const int src = key & 31;
            row = __shfl_sync(0xffffffffu, row, src);
            b = __shfl_sync(0xffffffffu, b, src);
            wa.x = __shfl_sync(0xffffffffu, wa.x, src); wa.y = __shfl_sync(0xffffffffu, wa.y, src);
            wa.z = __shfl_sync(0xffffffffu, wa.z, src); wa.w = __shfl_sync(0xffffffffu, wa.w, src);
            wb.x = __shfl_sync(0xffffffffu, wb.x, src); wb.y = __shfl_sync(0xffffffffu, wb.y, src);
            wb.z = __shfl_sync(0xffffffffu, wb.z, src); wb.w = __shfl_sync(0xffffffffu, wb.w, src);
        }
        const bool valid = row >= 0;
        unsigned present[S::NBW];
#pragma unroll
        for (int w = 0; w < S::NBW; ++w)
            present[w] = __reduce_or_sync(0xffffffffu, (valid && (b >> 5) == w) ? 1u << (b & 31) : 0u);
        const unsigned active = __ballot_sync(0xffffffffu, valid);
        const int cnt = __popc(active);
        __syncwarp();  // previous chunk fully consumed
        // pair at compacted position pos: offset -> entry pos-4 (.x), base cell -> entry pos-1 (.y); the first four
        // offsets go to `head`
        auto put_meta = [&](int pos, int off, int bb) {
            if (pos >= kGatherSlots) cx.metai[2 * (pos - kGatherSlots)] = off; else cx.head[pos] = off;
            if (pos >= 1) cx.metai[2 * (pos - 1) + 1] = bb; else cx.head[4] = bb;
        };
        if (valid) {
            const int pos = __popc(active & cx.lt_mask);
            put_meta(pos, row * cx.stride_b, b);
            *reinterpret_cast<float4*>(cx.wgt + pos * 8) = wa;
            *reinterpret_cast<float4*>(cx.wgt + pos * 8 + 4) = wb;
        }
        if (lane < 8) put_meta(cnt + lane, -1, S::NB);  // eight null pairs: no gather, cell NB = end of chunk
        __syncwarp();
        // ---- raw records of this point's next chunk: in flight during the walk ----
        if (!last) cur = pair_record(p, c0 + 32 + lane, c0 + 32 + lane < re, ox, oy, oz);
        // ---- walk the chunk: merge over the base cells ----
        if (cnt > 0) {
            Walk wk;
            {
                // gathers of pairs 0..3 into the ring slots following the current one (every earlier gather of this
                // warp has been consumed: null pairs never start one)
                const int4 h = *reinterpret_cast<const int4*>(cx.head);
                uint32_t s = cx.sa;
                if ((h.x | cx.gate) >= 0) cp_async4(s, cx.fbase + (unsigned)h.x);
                cp_commit(); s = ring_next(s);
                if ((h.y | cx.gate) >= 0) cp_async4(s, cx.fbase + (unsigned)h.y);
                cp_commit(); s = ring_next(s);
                if ((h.z | cx.gate) >= 0) cp_async4(s, cx.fbase + (unsigned)h.z);
                cp_commit(); s = ring_next(s);
                if ((h.w | cx.gate) >= 0) cp_async4(s, cx.fbase + (unsigned)h.w);
                cp_commit();
                cp_wait<kGatherSlots - 1>();  // pair 0 has landed
                wk.sa = cx.sa;
                wk.sa_next = ring_next(cx.sa);
                wk.f = lds_f32(cx.sa);
                wk.mm = cx.meta[0];
                wk.m2 = cx.meta + 1;
                wk.w = cx.wgt;
                wk.wa = *reinterpret_cast<const float4*>(cx.wgt);
                wk.wb = *reinterpret_cast<const float4*>(cx.wgt + 4);
            }
            sweep<S, RELU, FX>(wk, present, acc, cx.fbase, cx.gate, p.feat_scale, fc);
            cx.sa = wk.sa;
        }
        if (last) break;
    }
    return norm_acc;
}

// ---- multi-pair phase 1 for narrow inputs (cin <= 8) ----------------------------------------------------------------
// With few input channels the lane = channel layout of point_patch leaves most lanes idle and still pays the ~28-instruction
// walk for every pair (measured: the 4 -> 32 and the 24 -> 32 cross-scale convs of Liquid3d cost the same 2 ms).  Here the 32
// lanes are SL pair slots x CP = 32 / SL channels: one step scatters SL pairs at once, every lane into ITS OWN column of a
// per-warp shared-memory patch accs[cell][ch * SL + slot] (bank = a permutation of the lane id: conflict free, no atomics; no
// sort, no indirect branch, no per-pair control flow).  The chunks of a point alternate between two register sets (static
// indices, no register moves behind loads in flight): while the steps of chunk k run, the features of chunk k+1 and the raw
// records of chunk k+2 are in flight.  When the point is complete the SL slot words of every (cell, channel) -- contiguous, one
// LDS.128 -- are summed in a fixed order (deterministic) into the CTA's patch tile and zeroed for the warp's next point.
template <class G, int SL, int MT>
__device__ __forceinline__ float point_patch_mp(const ConvParams& p, int lane, const PairRec& first, int64_t rs, int64_t re,
                                                float ox, float oy, float oz, float fc, bool fx, float* accs, float* patch, int m) {
    static_assert(SL == 4 || SL == 8, "pair slots per step");
    constexpr int CP = 32 / SL, NT = 32 / SL;
    constexpr unsigned FULL = 0xffffffffu;
    const int slot = lane / CP, ch = lane % CP;
    const int col = ch * SL + slot;  // this lane's column of the slot patch
    const bool ch_ok = ch < p.cin;
    const char* fbase = reinterpret_cast<const char*>(p.inp_feat) + 4 * (ch_ok ? ch : 0);
    const int stride_b = (int)p.inp_stride * 4;
    const bool relu = p.relu_input != 0;
    const float scale = p.feat_scale;
    float norm_acc = 0.0f;

    // two chunk states (even / odd chunks of the point): raw records, base form (cell of corner 0, byte offset of the feature
    // row or -1, 8 corner weights) of this lane's pair, and the features of this lane's slot for the NT steps
    PairRec raw[2];
    int c000[2], off[2], nrange[2];
    float4 wa[2], wb[2];
    float f[2][NT];

    auto to_base = [&](const PairRec& r, int& c0, int& o, float4& a, float4& b) {
        c0 = 0; o = -1;
        a = make_float4(0.f, 0.f, 0.f, 0.f); b = a;
        if (r.row >= 0) {
            int bx, by, bz;
            float xl, xh, yl, yh, zl, zh;
            base_axis(G::KX_, r.g.i0 & 0xff, r.g.wx0, r.g.wx1, bx, xl, xh);
            base_axis(G::KY_, (r.g.i0 >> 8) & 0xff, r.g.wy0, r.g.wy1, by, yl, yh);
            base_axis(G::KZ_, (r.g.i0 >> 16) & 0xff, r.g.wz0, r.g.wz1, bz, zl, zh);
            c0 = (bz * G::KY_ + by) * G::KX_ + bx;
            a = make_float4(xl * yl * zl, xh * yl * zl, xl * yh * zl, xh * yh * zl);
            b = make_float4(xl * yl * zh, xh * yl * zh, xl * yh * zh, xh * yh * zh);
            o = r.row * stride_b;
        }
    };
    // features of a chunk: step t serves pairs t*SL .. t*SL + SL - 1, this lane the pair of its slot
    auto load_feats = [&](int o_lane, int nsteps, float (&ff)[NT]) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int o = __shfl_sync(FULL, o_lane, t * SL + slot);
            ff[t] = (t < nsteps && ch_ok && o >= 0) ? __ldg(reinterpret_cast<const float*>(fbase + (unsigned)o)) : 0.0f;
        }
    };
    auto steps = [&](int c0_lane, const float4& a4, const float4& b4, const float (&ff)[NT], int nsteps) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (t < nsteps) {  // warp uniform
                const int q = t * SL + slot;
                const int c0 = __shfl_sync(FULL, c0_lane, q);
                const float w0 = __shfl_sync(FULL, a4.x, q), w1 = __shfl_sync(FULL, a4.y, q);
                const float w2 = __shfl_sync(FULL, a4.z, q), w3 = __shfl_sync(FULL, a4.w, q);
                const float w4 = __shfl_sync(FULL, b4.x, q), w5 = __shfl_sync(FULL, b4.y, q);
                const float w6 = __shfl_sync(FULL, b4.z, q), w7 = __shfl_sync(FULL, b4.w, q);
                float v = ff[t];
                if (relu) v = fmaxf(v, 0.0f);
                if (fx) v = fmaf(v, scale, fc);
                float* a = accs + c0 * 32 + col;
                constexpr int sx = 32, sy = 32 * G::KX_, sz = 32 * G::KY_ * G::KX_;
                a[0] = fmaf(w0, v, a[0]);
                if constexpr (G::KX_ > 1) a[sx] = fmaf(w1, v, a[sx]);
                if constexpr (G::KY_ > 1) a[sy] = fmaf(w2, v, a[sy]);
                if constexpr (G::KX_ > 1 && G::KY_ > 1) a[sx + sy] = fmaf(w3, v, a[sx + sy]);
                if constexpr (G::KZ_ > 1) {
                    a[sz] = fmaf(w4, v, a[sz]);
                    if constexpr (G::KX_ > 1) a[sz + sx] = fmaf(w5, v, a[sz + sx]);
                    if constexpr (G::KY_ > 1) a[sz + sy] = fmaf(w6, v, a[sz + sy]);
                    if constexpr (G::KX_ > 1 && G::KY_ > 1) a[sz + sx + sy] = fmaf(w7, v, a[sz + sx + sy]);
                }
            }
        }
    };
    auto chunk_len = [&](int64_t c0) {
        const int64_t n = re - c0;
        return (int)(n < 0 ? 0 : (n < 32 ? n : 32));
    };
    // chunk c0 lives in state P; prepares chunk c0 + 32 in state 1 - P (needs its raw records, issued one chunk earlier), issues
    // the raw records of chunk c0 + 64 into raw[P] (chunk c0's are dead: it is in base form), then runs the steps of chunk c0.
    // Returns true when chunk c0 was the last one.
    auto half = [&](auto P_, int64_t c0) {
        constexpr int P = decltype(P_)::value, Q = 1 - P;
        const bool last = c0 + 32 >= re;  // warp uniform
        if (!last) {
            to_base(raw[Q], c000[Q], off[Q], wa[Q], wb[Q]);
            norm_acc += raw[Q].norm;
            nrange[Q] = chunk_len(c0 + 32);
            if (c0 + 64 < re) raw[P] = pair_record(p, c0 + 64 + lane, c0 + 64 + lane < re, ox, oy, oz);
            load_feats(off[Q], (nrange[Q] + SL - 1) / SL, f[Q]);
        }
        steps(c000[P], wa[P], wb[P], f[P], (nrange[P] + SL - 1) / SL);
        return last;
    };

    raw[0] = first;
    raw[1] = first;  // placeholder, overwritten before use
    to_base(raw[0], c000[0], off[0], wa[0], wb[0]);
    norm_acc += raw[0].norm;
    nrange[0] = chunk_len(rs);
    if (rs + 32 < re) raw[1] = pair_record(p, rs + 32 + lane, rs + 32 + lane < re, ox, oy, oz);
    load_feats(off[0], (nrange[0] + SL - 1) / SL, f[0]);
#pragma unroll 1
    for (int64_t c0 = rs;; c0 += 64) {
        if (half(std::integral_constant<int, 0>{}, c0)) break;
        if (half(std::integral_constant<int, 1>{}, c0 + 32)) break;
    }
    // ---- slot words -> patch row: lane handles (cell, channel) = idx / CP, idx % CP; its SL slot words are contiguous ----
    __syncwarp();
#pragma unroll 2
    for (int idx = lane; idx < G::K * CP; idx += 32) {
        const int cell = idx / CP, c = idx % CP;
        float4* w4 = reinterpret_cast<float4*>(accs + cell * 32 + c * SL);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 a = w4[0];
        w4[0] = z4;
        float v = (a.x + a.y) + (a.z + a.w);
        if constexpr (SL == 8) {
            const float4 b = w4[1];
            w4[1] = z4;
            v += (b.x + b.y) + (b.z + b.w);
        }
        if (c < p.cin) patch[patchq_index<MT>(m, cell * p.cin + c)] = v;
    }
    __syncwarp();
    return norm_acc;
}

}  // namespace lean

}  // namespace dmcf
