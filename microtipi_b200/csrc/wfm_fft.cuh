// wfm_fft.cuh -- batched in-register / shared-memory complex FFT engine for sm_100a.
//
// Replaces JTransforms' DoubleFFT_2D / FloatFFT_2D.complexForward (call sites
// WideFieldModel.java:319-321, 604-605, 917-918, 1241-1242 and the Float twins 248-249,
// 476-477, 781-782, 1099-1100): unnormalised forward transform, kernel e^{-2 pi i jk/N}.
//
// One transform of length N is computed by T = N/E threads, each holding E complex values in
// registers.  N = R1*R2*R3 (R3 == 1 for a two-stage plan); every stage is a set of radix-R
// butterflies done entirely in registers, with one shared-memory exchange between stages
// (decimation in frequency, in place, so a stage reads and writes the same smem cells and needs a
// single barrier).  Thread -> butterfly maps are chosen so that
//     input  slot v[u*R1 + r] = x[(t + T*u) + (N/R1)*r]
//     output slot v[u*RL + r] = X[(t + T*u) + (N/RL)*r]        (RL = last radix)
// i.e. consecutive threads own consecutive indices on both sides -> coalesced global access
// with no extra reordering pass.
//
// Two shared-memory layouts:
//   RowLayout  -- the transform runs along the contiguous axis; each transform has a private,
//                 padded row (padding chosen by tools/bank_conflicts.py: conflict-free for fp64).
//   ColLayout  -- C transforms (adjacent columns) run side by side, column index innermost, so
//                 every access of 8 adjacent lanes is one 128-byte wavefront by construction.
#pragma once

#include "wfm_platform.cuh"

#define WFM_DEVI __device__ __forceinline__

namespace wfm {

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };
template <typename T> using cx = typename Vec2<T>::type;

template <typename T> WFM_DEVI cx<T> mkc(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename C> WFM_DEVI C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> WFM_DEVI C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// fp32 mode: a complex add / subtract is ONE packed instruction on sm_100 (add.rn.f32x2 -> FADD2); the butterflies
// are 52 of the 56 arithmetic instructions of a radix-8 transform
#ifndef WFM_F32X2
#define WFM_F32X2 1
#endif
#if WFM_F32X2 && !defined(WFM_EMU)
template <> WFM_DEVI float2 cadd<float2>(float2 a, float2 b) {
    unsigned long long ra, rb, rr;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rr) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rr));
    return r;
}
template <> WFM_DEVI float2 csub<float2>(float2 a, float2 b) {
    unsigned long long ra, rb, rr;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rr) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rr));
    return r;
}
#endif
template <typename C> WFM_DEVI C cmul(C a, C w) { C r; r.x = a.x * w.x - a.y * w.y; r.y = a.x * w.y + a.y * w.x; return r; }
// x * (-i)
template <typename C> WFM_DEVI C mul_neg_i(C a) { C r; r.x = a.y; r.y = -a.x; return r; }

// ---- multiply by W16^M = exp(-2 pi i M/16), M in [0,8), all folded at compile time ---------
template <typename T, int M> struct MulW16 {
    static WFM_DEVI cx<T> run(cx<T> a) {
        if constexpr (M == 0) {
            return a;
        } else if constexpr (M >= 4) {
            return mul_neg_i(MulW16<T, M - 4>::run(a));
        } else if constexpr (M == 2) {
            const T s = (T)0.70710678118654752440084436210484903928;
            return mkc<T>((a.x + a.y) * s, (a.y - a.x) * s);
        } else {
            const T c = (M == 1) ? (T)0.92387953251128675612818318939678828682 : (T)0.38268343236508977172845998403039886676;
            const T s = (M == 1) ? (T)0.38268343236508977172845998403039886676 : (T)0.92387953251128675612818318939678828682;
            return mkc<T>(a.x * c + a.y * s, a.y * c - a.x * s);
        }
    }
};

// ---- radix-R DFT in registers, natural-order in and out (radix-2 DIT recursion) ------------
template <typename T, int R> struct Dft;

template <typename T, int R, int K> struct DftCombine {
    static WFM_DEVI void run(cx<T> (&v)[R], const cx<T> (&e)[R / 2], const cx<T> (&o)[R / 2]) {
        const cx<T> t = MulW16<T, K * (16 / R)>::run(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        if constexpr (K + 1 < R / 2) DftCombine<T, R, K + 1>::run(v, e, o);
    }
};

template <typename T> struct Dft<T, 1> {
    static WFM_DEVI void run(cx<T> (&)[1]) {}
};
template <typename T> struct Dft<T, 2> {
    static WFM_DEVI void run(cx<T> (&v)[2]) {
        const cx<T> a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <typename T, int R> struct Dft {
    static_assert(R == 4 || R == 8 || R == 16, "radix");
    static WFM_DEVI void run(cx<T> (&v)[R]) {
        cx<T> e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
        Dft<T, R / 2>::run(e);
        Dft<T, R / 2>::run(o);
        DftCombine<T, R, 0>::run(v, e, o);
    }
};

// Radix-8 DFT whose inputs 2..5 are structurally zero (the pupil occupies the legs {0, 1, 6, 7} of the first
// stage when it fits the central half of the frequency axis, see PipeCfg / "narrow" kernels):
//   v[k] = a0 + i^k a6 + W8^k a1 + W8^-k a7 = p[k mod 4] + c_k (a1 + a7) - i s_k (a1 - a7)
// 32 additions + 4 multiplications instead of 48 + 8.
template <typename T> WFM_DEVI void dft8_in0167(cx<T> (&v)[8]) {
    const T h = (T)0.70710678118654752440084436210484903928;
    const cx<T> a0 = v[0], a1 = v[1], a6 = v[6], a7 = v[7];
    const cx<T> p0 = cadd(a0, a6), p2 = csub(a0, a6);
    const cx<T> p1 = mkc<T>(a0.x - a6.y, a0.y + a6.x);     // a0 + i a6
    const cx<T> p3 = mkc<T>(a0.x + a6.y, a0.y - a6.x);     // a0 - i a6
    const cx<T> S = cadd(a1, a7), D = csub(a1, a7);
    const cx<T> U = mkc<T>(h * S.x, h * S.y);               // h S
    const cx<T> V = mkc<T>(h * D.y, -(h * D.x));            // -i h D
    const cx<T> W = mkc<T>(D.y, -D.x);                      // -i D
    const cx<T> A = cadd(U, V), B = csub(V, U);
    v[0] = cadd(p0, S); v[4] = csub(p0, S);
    v[2] = cadd(p2, W); v[6] = csub(p2, W);
    v[1] = cadd(p1, A); v[5] = csub(p1, A);
    v[3] = cadd(p3, B); v[7] = csub(p3, B);
}

// Radix-16 with inputs 4..11 zero: its even and odd halves are radix-8 transforms with inputs 2..5 zero.
template <typename T> WFM_DEVI void dft16_in_narrow(cx<T> (&v)[16]) {
    cx<T> e[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    dft8_in0167<T>(e);
    dft8_in0167<T>(o);
    DftCombine<T, 16, 0>::run(v, e, o);
}

// ---- plans ---------------------------------------------------------------------------------
template <int N_, int E_, int R1_, int R2_, int R3_> struct PlanBase {
    static constexpr int N = N_, E = E_, R1 = R1_, R2 = R2_, R3 = R3_;
    static_assert(R1 * R2 * R3 == N, "factorisation");
    static_assert(E % R1 == 0 && E % R2 == 0 && E % R3 == 0, "radix must divide E");
    static constexpr int T = N / E;                  // threads per transform
    static constexpr int S1 = N / R1;                // stage-1 leg stride
    static constexpr bool THREE = (R3 > 1);
    static constexpr int RL = THREE ? R3 : R2;       // last radix
    static constexpr int SL = N / RL;                // output leg stride
};

template <int N> struct Plan;
template <> struct Plan<16> : PlanBase<16, 8, 4, 4, 1> {};     // half-length transform of the 32-point real rows (wfm_conv.cuh)
template <> struct Plan<32> : PlanBase<32, 8, 8, 4, 1> {};
template <> struct Plan<64> : PlanBase<64, 8, 8, 8, 1> {};
template <> struct Plan<128> : PlanBase<128, 8, 8, 4, 4> {};
template <> struct Plan<256> : PlanBase<256, 8, 8, 8, 4> {};
#if defined(WFM_PLAN512_W32)   /* experiment: one warp per transform (T = 32), two radix-8 butterflies per lane and stage */
template <> struct Plan<512> : PlanBase<512, 16, 8, 8, 8> {};
#elif defined(WFM_PLAN512_E16)   /* experiment: one warp per transform (T = 32), 16 values per thread */
template <> struct Plan<512> : PlanBase<512, 16, 16, 8, 4> {};
#else
template <> struct Plan<512> : PlanBase<512, 8, 8, 8, 8> {};
#endif
// 1024 = 8*8*16: two radix-8 butterflies per thread in the twiddled stages, the radix-16 butterfly last (no twiddles
// behind it: fewest live registers).  Measured at 1024^2 x 64 fp64 against 16*8*8 (-DWFM_PLAN1024_A): 0.923 vs 0.948 ms per
// step at two resident CTAs, 1.05 vs 1.27 ms at three (profiles/r02g_other_shapes.md).
#ifdef WFM_PLAN1024_A
template <> struct Plan<1024> : PlanBase<1024, 16, 16, 8, 8> {};
#else
template <> struct Plan<1024> : PlanBase<1024, 16, 8, 8, 16> {};
#endif
template <> struct Plan<2048> : PlanBase<2048, 16, 16, 16, 8> {};

// Row padding shifts (PA, PB): pad(i) = i + (i >> PA) + (i >> PB), 0 disables a term.
template <int N, int ESZ> struct RowPad;
template <> struct RowPad<16, 16> { static constexpr int PA = 0, PB = 0; };
template <> struct RowPad<32, 16> { static constexpr int PA = 3, PB = 4; };
template <> struct RowPad<64, 16> { static constexpr int PA = 3, PB = 0; };
template <> struct RowPad<128, 16> { static constexpr int PA = 0, PB = 4; };
template <> struct RowPad<256, 16> { static constexpr int PA = 3, PB = 6; };
#ifdef WFM_PLAN512_E16
template <> struct RowPad<512, 16> { static constexpr int PA = 3, PB = 6; };
#else
template <> struct RowPad<512, 16> { static constexpr int PA = 0, PB = 6; };
#endif
#ifdef WFM_PLAN1024_A
template <> struct RowPad<1024, 16> { static constexpr int PA = 0, PB = 6; };
#else
template <> struct RowPad<1024, 16> { static constexpr int PA = 0, PB = 7; };
#endif
template <> struct RowPad<2048, 16> { static constexpr int PA = 0, PB = 7; };
template <> struct RowPad<16, 8> { static constexpr int PA = 0, PB = 0; };
template <> struct RowPad<32, 8> { static constexpr int PA = 3, PB = 4; };
template <> struct RowPad<64, 8> { static constexpr int PA = 3, PB = 5; };
template <> struct RowPad<128, 8> { static constexpr int PA = 2, PB = 0; };
template <> struct RowPad<256, 8> { static constexpr int PA = 0, PB = 4; };
template <> struct RowPad<512, 8> { static constexpr int PA = 0, PB = 6; };
#ifdef WFM_PLAN1024_A
template <> struct RowPad<1024, 8> { static constexpr int PA = 0, PB = 6; };
#else
template <> struct RowPad<1024, 8> { static constexpr int PA = 0, PB = 7; };
#endif
template <> struct RowPad<2048, 8> { static constexpr int PA = 4, PB = 8; };

// fp32 rows (8-byte cells, 16 lanes per 128-byte pass): no shift padding serves both the stage-2 exchange (two stage-1
// blocks per pass) and the last-stage loads without replays (best 1.33x: 25 % of the shared wavefronts of the fp32
// pipelines were replays, profiles/r02i).  A per-block offset  off(k1) = A (k1 & 1) + B ((k1 >> 1) & 1) + C (k1 >> 2)  cells
// together with 16-byte last-stage loads (two adjacent cells per LDS.128) is conflict free for every access of the
// engine (tools/bank_conflicts.py, score_off).  A = 0 selects the shift padding.
template <int N, int ESZ> struct RowOff { static constexpr int A = 0, B = 0, C = 0; };
#ifndef WFM_NO_F32_ROWOFF
// (only off(k1) mod 16 matters for the banks; multiples of 16 are added so that off() is non-decreasing -- blocks must not overlap)
template <> struct RowOff<128, 8> { static constexpr int A = 4, B = 8, C = 2 + 16; };            // 0 4 8 12 18 22 26 30
template <> struct RowOff<256, 8> { static constexpr int A = 4, B = 8, C = 2 + 16; };
template <> struct RowOff<512, 8> { static constexpr int A = 8, B = 2 + 16, C = 4 + 32; };       // 0 8 18 26 36 44 54 62
template <> struct RowOff<1024, 8> { static constexpr int A = 2, B = 4, C = 8; };                // 0 2 4 ... 14
#endif

template <typename T, int N> struct RowLayout {
    using P = RowPad<N, (int)sizeof(cx<T>)>;
    using O = RowOff<N, (int)sizeof(cx<T>)>;
    static constexpr bool BLOCK_OFF = (O::A != 0);
    static constexpr int S1 = Plan<N>::S1;
    __host__ __device__ static constexpr int pad_c(int i) {
        if (BLOCK_OFF) {
            const int k = i / S1;
            return i + O::A * (k & 1) + O::B * ((k >> 1) & 1) + O::C * (k >> 2);
        }
        return i + (P::PA ? (i >> P::PA) : 0) + (P::PB ? (i >> P::PB) : 0);
    }
    static constexpr bool blocks_disjoint() {
        for (int k = 1; k < N / S1; ++k) if (pad_c(k * S1) < pad_c(k * S1 - 1) + 1) return false;
        return true;
    }
    static_assert(blocks_disjoint(), "row layout: the stage-1 blocks overlap");
    // cells per transform (even with block offsets: every private row then starts 16-byte aligned)
    static constexpr int LEN = BLOCK_OFF ? ((pad_c(N - 1) + 2) & ~1) : pad_c(N - 1) + 1;
    // the last stage may read its adjacent cells two at a time (16-byte loads): cell indices of a butterfly's first leg are even
    static constexpr bool VEC_LAST = BLOCK_OFF;
    static WFM_DEVI int at(int i) { return pad_c(i); }
    __host__ __device__ static constexpr int at_c(int i) { return pad_c(i); }
    // Index arithmetic the engine may rely on for blocks of B consecutive indices starting at multiples of B
    // (B = N/R1): at(k*B + b) == at(b) + k*at_c(B) for b < B, and at(base + j) == at(base) + j while base + j stays
    // inside its block.  Both hold when every padding term steps exactly once per block, i.e. 2^shift == B
    // (constant inside a block, and (k*B + b) >> shift == k + (b >> shift)).
    template <int B> __host__ __device__ static constexpr bool affine() {
        return BLOCK_OFF ? (B == S1) : ((P::PA == 0 || (1 << P::PA) == B) && (P::PB == 0 || (1 << P::PB) == B));
    }
    static constexpr int UNIT = 1;                    // distance between consecutive indices of a block
};

// Column layout: cell = (i + (i >> SH)) * C + column.  With SH = log2(N/R1) the stage-3 reads of a
// narrow tile (C = 4 fp64 columns, two index values per 128-byte wavefront) are conflict free too
// (tools/bank_conflicts.py, score_cols).
template <int C, int SH> struct ColLayout {
    __host__ __device__ static constexpr int pad_c(int i) { return i + (i >> SH); }
    static WFM_DEVI int at(int i) { return pad_c(i) * C; }
    __host__ __device__ static constexpr int at_c(int i) { return pad_c(i) * C; }
    template <int B> __host__ __device__ static constexpr bool affine() { return (1 << SH) == B; }
    static constexpr int UNIT = C;
    static constexpr bool VEC_LAST = false;
};
__host__ __device__ constexpr int ilog2_c(int v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }

// ---- barrier scopes --------------------------------------------------------------------------
// A column tile is transposed through shared memory by all warps of the CTA: CTA-wide barrier.
struct CtaSync {
#ifdef WFM_PROBE_NO_CTA_SYNC      /* timing probe only (races, wrong results): what do the exchange barriers of the column items cost? */
    static WFM_DEVI void sync(int) {}
#else
    static WFM_DEVI void sync(int) { __syncthreads(); }
#endif
};
// A row transform is private to its TT threads: a named barrier over those warps when TT >= 64,
// a warp barrier when the transform fits in one warp.  (ncu on the first pipeline revision:
// barrier stalls dominated with 16-warp CTA barriers; profiles/r01b_*.)
template <int TT> struct RowSync {
    static WFM_DEVI void sync(int slot) {
#ifdef WFM_PROBE_NO_ROW_SYNC      /* timing probe only (races, wrong results): what do the 2-warp barriers of the row transforms cost? */
        (void)slot;
        return;
#endif
        if constexpr (TT >= 64) {
#ifdef WFM_EMU
            emu::named_barrier(1 + slot, TT);
#else
            asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(TT) : "memory");
#endif
        } else {
            __syncwarp();
        }
    }
};

// Powers w^1 .. w^(R-1) of a base twiddle with a dependency depth of log2(R) complex products instead of R-2
// (w^k = w^hb(k) * w^(k-hb(k)), hb = highest power of two <= k): the transforms are bound by dependency latency,
// and the sequential chain was the longest one of a stage.
// It keeps more powers live, so it is only used where the register budget allows (the Jacobian pipeline at 80
// registers: 0.4245 -> 0.418 ms; the PSF pipeline at 64 registers spills with it: 0.445 -> 0.455 ms).
template <typename T, int R> WFM_DEVI void twiddle_powers(cx<T> (&pw)[R], const cx<T> w) {
    pw[0] = mkc<T>((T)1, (T)0);
    pw[1] = w;
#pragma unroll
    for (int k = 2; k < R; ++k) {
        int hb = 1;
        while (hb * 2 <= k) hb *= 2;
        pw[k] = (hb == k) ? cmul(pw[k / 2], pw[k / 2]) : cmul(pw[hb], pw[k - hb]);
    }
}

// ---- the engine ----------------------------------------------------------------------------
// v   : E register values of this thread (slot convention above)
// sm  : base of this transform's shared cells (RowLayout: private row; ColLayout: smem + column)
// t   : thread index inside the transform, [0, T)
// tw2 : stage-2 base twiddles in SHARED memory, tw2[d] = W_N^(R1*d), d in [0, R3) (a strided read of
//       tw was an 8-way bank conflict: 13 % of all shared wavefronts in profiles/r01d_*)
// tw  : W_N table in SHARED memory, tw[m] = exp(-2 pi i m / N), m in [0, N).  Only the base twiddle
//       of each butterfly is read (one LDS per butterfly); its powers w^2..w^(R-1) are formed by
//       complex multiplication in registers.  (ncu on the first revision showed the LSU data pipe
//       95 % busy with 2/3 of its wavefronts spent on per-leg twiddle loads; profiles/r01a_*.)
// S   : barrier scope (CtaSync / RowSync); every thread inside that scope must call this together.  The caller
// must place a barrier between the end of one call and the start of the next one that reuses sm.
struct NoHook { WFM_DEVI void operator()() const {} };

// Hook: callable run once by every thread right after the first exchange barrier (used by the
// pipelines to claim the next work item while two thirds of the transform are still ahead).
// SPARSE1: the caller guarantees that the middle half of the stage-1 legs of every butterfly is zero (legs 2..5
// for R1 == 8, 4..11 for R1 == 16); their slots in v are ignored.
// TWTREE: 0 = sequential chain w, w^2, ... (2 live twiddles), 1 = tree (depth log2 R, up to R/2 live),
//         2 = two interleaved chains stepping by w^2 (depth R/2, 3 live)
// Hook2: callable run once by every thread right after the SECOND exchange barrier of a three-stage plan, i.e. before
// the last (twiddle-free, lowest register pressure) stage: the place to put loads for the NEXT transform in flight.
// TWTAB: bit 0 = the stage-1 twiddles w^k come from the shared table, tw[(k-1)*S1 + b] = W_N^(b*k), k in [1, R1) (row k = 1
//        is the base table); bit 1 = the stage-2 twiddles likewise, tw2[(k-1)*R3 + d3] = W_N^(R1*d3*k), k in [1, R2)
//        (the R3 lanes of a butterfly group read adjacent entries: one wavefront per load).  No power chains: fewer FP64 instructions, shorter
//        dependency chains, more LDS.
template <typename T, class P, class L, class S, class Hook = NoHook, bool SPARSE1 = false, int TWTREE = 0, class Hook2 = NoHook,
          int TWTAB = 0>
WFM_DEVI void fft_inplace(cx<T> (&v)[P::E], cx<T>* sm, const int t, const cx<T>* tw, const cx<T>* tw2,
                          const int sync_id, const Hook& hook = Hook(), const Hook2& hook2 = Hook2()) {
    static_assert(!SPARSE1 || P::R1 == 8 || P::R1 == 16, "sparse first stage: radix 8 or 16");
    constexpr int E = P::E, R1 = P::R1, R2 = P::R2, R3 = P::R3, TT = P::T, S1 = P::S1;
    // AFF: the layout is affine over the S1-blocks (see RowLayout::affine): every stage then needs ONE address per
    // butterfly, its legs are compile-time offsets.  (Without it the compiler kept one address register per leg for
    // most exchanges: cuobjdump on the 512-point pipelines.)
#ifdef WFM_NO_AFFINE
    constexpr bool AFF = false;
#else
    constexpr bool AFF = L::template affine<S1>();
#endif
    constexpr int UNIT = L::UNIT;
    // stage 1: radix R1 over legs of stride S1, twiddle W_N^(b*k1), scatter to cell k1*S1 + b
#pragma unroll
    for (int u = 0; u < E / R1; ++u) {
        cx<T> a[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) a[r] = v[u * R1 + r];
        if constexpr (SPARSE1 && R1 == 8) dft8_in0167<T>(reinterpret_cast<cx<T>(&)[8]>(a));
        else if constexpr (SPARSE1 && R1 == 16) dft16_in_narrow<T>(reinterpret_cast<cx<T>(&)[16]>(a));
        else Dft<T, R1>::run(a);
        const int b = t + TT * u;
        cx<T>* const s1 = sm + L::at(b);               // AFF: leg k of this butterfly lives at s1[k * LS1]
        auto cell1 = [&](int k) -> cx<T>& { return AFF ? s1[L::at_c(k * S1)] : sm[L::at(k * S1 + b)]; };   // (= k * LS1 for the shift paddings)
        cell1(0) = a[0];
        if constexpr (TWTAB & 1) {
#pragma unroll
            for (int k = 1; k < R1; ++k) cell1(k) = cmul(a[k], tw[(k - 1) * S1 + b]);
        } else if constexpr (TWTREE == 1) {
            cx<T> pw[R1];
            twiddle_powers<T, R1>(pw, tw[b]);
#pragma unroll
            for (int k = 1; k < R1; ++k) cell1(k) = cmul(a[k], pw[k]);
        } else if constexpr (TWTREE == 2) {
            const cx<T> w = tw[b];
            const cx<T> w2 = cmul(w, w);
            cx<T> wo = w, we = w2;
#pragma unroll
            for (int k = 1; k < R1; k += 2) {
                cell1(k) = cmul(a[k], wo);
                if (k + 1 < R1) cell1(k + 1) = cmul(a[k + 1], we);
                if (k + 2 < R1) wo = cmul(wo, w2);
                if (k + 3 < R1) we = cmul(we, w2);
            }
        } else {
            const cx<T> w = tw[b];
            cx<T> wk = w;
#pragma unroll
            for (int k = 1; k < R1; ++k) {
                cell1(k) = cmul(a[k], wk);
                if (k + 1 < R1) wk = cmul(wk, w);
            }
        }
    }
    S::sync(sync_id);
    hook();
    if constexpr (P::THREE) {
        // stage 2: inside block k1, radix R2 over legs of stride R3, twiddle W_N^(R1*d3*k2)
#pragma unroll
        for (int u = 0; u < E / R2; ++u) {
            const int b = t + TT * u;
            const int k1 = b / R3, d3 = b % R3;
            const int base = k1 * S1 + d3;
            cx<T>* const s2 = sm + L::at(base);            // AFF: legs r*R3 further on inside block k1
            auto cell2 = [&](int r) -> cx<T>& { return AFF ? s2[r * R3 * UNIT] : sm[L::at(base + r * R3)]; };
            cx<T> a[R2];
#pragma unroll
            for (int r = 0; r < R2; ++r) a[r] = cell2(r);
            Dft<T, R2>::run(a);
#ifdef WFM_FAKE_NO_X2   /* timing experiment only (wrong results): what would the step cost without the second exchange? */
            const cx<T> w = tw2[d3];
            cx<T> wk = w;
            v[u * R2] = a[0];
#pragma unroll
            for (int k = 1; k < R2; ++k) {
                v[u * R2 + k] = cmul(a[k], wk);
                if (k + 1 < R2) wk = cmul(wk, w);
            }
        }
#else
            cell2(0) = a[0];
            // compact table tw2[d] = W_N^(R1*d): adjacent lanes, adjacent cells
            if constexpr (TWTAB & 2) {
#pragma unroll
                for (int k = 1; k < R2; ++k) cell2(k) = cmul(a[k], tw2[(k - 1) * R3 + d3]);
            } else if constexpr (TWTREE == 1) {
                cx<T> pw[R2];
                twiddle_powers<T, R2>(pw, tw2[d3]);
#pragma unroll
                for (int k = 1; k < R2; ++k) cell2(k) = cmul(a[k], pw[k]);
            } else if constexpr (TWTREE == 2) {
                const cx<T> w = tw2[d3];
                const cx<T> w2 = cmul(w, w);
                cx<T> wo = w, we = w2;
#pragma unroll
                for (int k = 1; k < R2; k += 2) {
                    cell2(k) = cmul(a[k], wo);
                    if (k + 1 < R2) cell2(k + 1) = cmul(a[k + 1], we);
                    if (k + 2 < R2) wo = cmul(wo, w2);
                    if (k + 3 < R2) we = cmul(we, w2);
                }
            } else {
                const cx<T> w = tw2[d3];
                cx<T> wk = w;
#pragma unroll
                for (int k = 1; k < R2; ++k) {
                    cell2(k) = cmul(a[k], wk);
                    if (k + 1 < R2) wk = cmul(wk, w);
                }
            }
        }
        S::sync(sync_id);
#endif
        hook2();
        // stage 3: radix R3 over adjacent cells; butterfly b = k1 + R1*k2 -> X[b + (N/R3)*r]
#pragma unroll
        for (int u = 0; u < E / R3; ++u) {
            const int b = t + TT * u;
            const int k1 = b % R1, k2 = b / R1;
            const int base = k1 * S1 + k2 * R3;
            cx<T> a[R3];
#ifdef WFM_FAKE_NO_X2
#pragma unroll
            for (int r = 0; r < R3; ++r) a[r] = v[u * R3 + r];
            (void)base;
#else
            const cx<T>* const s3 = sm + L::at(base);      // AFF: R3 adjacent cells of block k1
            if constexpr (AFF && L::VEC_LAST && sizeof(cx<T>) == 8 && R3 % 2 == 0) {
                const float4* const s3v = reinterpret_cast<const float4*>(s3);      // two cells per 16-byte load
#pragma unroll
                for (int r = 0; r < R3; r += 2) {
                    const float4 x = s3v[r / 2];
                    a[r] = mkc<T>((T)x.x, (T)x.y);
                    a[r + 1] = mkc<T>((T)x.z, (T)x.w);
                }
            } else {
#pragma unroll
                for (int r = 0; r < R3; ++r) a[r] = AFF ? s3[r * UNIT] : sm[L::at(base + r)];
            }
#endif
            Dft<T, R3>::run(a);
#pragma unroll
            for (int r = 0; r < R3; ++r) v[u * R3 + r] = a[r];
        }
    } else {
        // last stage of a two-stage plan: radix R2 over adjacent cells of block k1 = b
#pragma unroll
        for (int u = 0; u < E / R2; ++u) {
            const int b = t + TT * u;
            const int base = b * S1;
            const cx<T>* const s2 = sm + L::at(base);
            cx<T> a[R2];
#pragma unroll
            for (int r = 0; r < R2; ++r) a[r] = AFF ? s2[r * UNIT] : sm[L::at(base + r)];
            Dft<T, R2>::run(a);
#pragma unroll
            for (int r = 0; r < R2; ++r) v[u * R2 + r] = a[r];
        }
    }
}

}  // namespace wfm
