// wfm_conv.cuh -- FFT convolution data term on the device ("next" row f1).
//
// Restates the role of TiPi's mitiv.conv.WeightedConvolutionCost as microTiPi drives it
// (PSF_Estimation.java:147-150 build/setPSF(obj,off)/setData/setWeights, :157,206
// computeCostAndGradient(1.0, psf, gcost, true)): the *object* is the kernel of the operator and
// the microscope PSF h is the variable,
//     cost = alpha * 1/2 * sum_k w_k * ((obj (*) h)_k - y_k)^2        (periodic 3-D convolution, offset 0)
//     grad = alpha * corr(obj, w * (obj (*) h - y))                   (same shape as h; this is the q of apply_J_*)
// TiPi's source is not in the reference tree (un-vendored, un-pinned): PARITY UNPINNED -- the
// semantics above are the documented ones and are checked against the oracle and by finite differences.
//
// Every array here is real, so the transforms work on the HALF spectrum: the x pass reads a real row of Nx samples
// as Nx/2 complex numbers z[j] = h[2j] + i h[2j+1], transforms them with the half-length plan and untangles the
// result into the Nx/2+1 non-redundant coefficients (the mirror element comes through the row's shared cells).
// The work volume is V[Nz][Ny][P] complex, P = Nx/2 + 8 (entries kx = 0..Nx/2, then zero padding up to a
// multiple of the column tile), i.e. half the bytes of a complex volume on every later pass:
//   x pass : rows (real <-> half spectrum), RowLayout, one transform per T-thread group
//   y pass : columns inside a plane, ColLayout tiles of CW adjacent kx
//   z pass : columns of the [Nz][Ny*P] matrix, ColLayout tiles of CW adjacent entries
// Inverse transforms use IFFT(v) = conj(FFT(conj(v)))/Ntot; the conjugations, the spectral products
// with X = FFT3(obj), the residual, the weights, the cost reduction and the final real part are all
// fused into the load / store of the neighbouring passes, so one evaluation is 12 sweeps of the half volume.
// (k_conv_rows, the full complex x pass, is kept for wfm_get_mtf, which returns the whole spectrum.)
#pragma once
#include "wfm_kernels.cuh"

namespace wfm {

enum ConvLoad { CL_CPLX = 0, CL_REAL = 1 };
enum ConvStore {
    CS_CPLX = 0,        // V = v
    CS_MULX_CONJ = 1,   // V = conj(v * X)                    (end of FFT3(h): spectrum product, set up the inverse)
    CS_RESID = 2,       // r = Re(v)/Ntot - y; cost += w r^2; V = (w r, 0)   (end of the inverse)
    CS_MULCX_CONJ = 3,  // V = conj(v * conj(X))              (end of FFT3(w r))
    CS_GRAD = 4,        // g = alpha * Re(v)/Ntot  -> real array           (end of the last inverse)
    CS_SPECTRUM = 5     // X = v                                (FFT3(obj) at set_object time)
};

// ---- z-slab sharding across GPUs -----------------------------------------------------------------------------
// Real-space volumes (h, data, weights, residual, gradient) and the x / y passes live in Z-SLABS: device g holds
// planes [z0_g, z0_g + nzl_g).  The z pass needs every plane of a (ky, kx) column, so it runs in PENCIL layout:
// device g holds ALL Nz planes of the rows ky in [y0_g, y0_g + nyl_g), Vp_g[z][ky - y0_g][kx].  The all-to-all
// transposes between the two layouts are not separate collectives: the y pass stores every output row straight into
// the pencil volume of the device that owns its ky (CW adjacent kx = one 128-byte NVLink store per row and tile),
// and the fused z pass stores every output plane straight into the slab volume of the device that owns its z.
// n items over `parts` owners, the first n % parts owners hold one more (the split of wfm_create_multi).
struct SplitMap {
    int n, parts;
    __host__ __device__ int base() const { return n / parts; }
    __host__ __device__ int rem() const { return n % parts; }
    __host__ __device__ int first(int o) const { return o * base() + (o < rem() ? o : rem()); }
    __host__ __device__ int count(int o) const { return base() + (o < rem() ? 1 : 0); }
    __host__ __device__ int owner(int i) const {
        const int b = base(), r = rem(), cut = r * (b + 1);
        return i < cut ? i / (b + 1) : r + (i - cut) / b;
    }
};
template <typename T> struct ConvPeers {
    cx<T>* vol[WFM_MAX_RANKS];   // destination volume of every device as mapped here (pencil volumes for the y pass,
                                 // slab volumes for the z pass)
    SplitMap split;              // how the scattered axis (ky for the y pass, z for the z pass) is divided
    int src_first;               // this device's first index on the OTHER axis (its z0 for the y pass, its y0 for the z pass)
    int ny_full;                 // Ny of the whole volume (slab rows per plane)
};

template <typename T> struct ConvArgs {
    cx<T>* V;             // work volume
    const T* real_in;     // CL_REAL source
    const cx<T>* X;       // spectrum of the object
    cx<T>* Xout;          // CS_SPECTRUM destination
    const T* y; const T* w;   // data, weights (w may be NULL = 1)
    T* grad;              // CS_GRAD destination
    double* cost_part;    // [gridDim.x] per-CTA partial sums of w r^2
    const cx<T>* tw;      // twiddles of this pass's length
    const cx<T>* twn;     // r2c / c2r rows: W_Nx^k, k < Nx/2
    T* resid;             // c2r CS_RESID destination: w * r as a real volume
    int nx, ny, nz;
    double inv_ntot, alpha;
    int clear_grad;       // CS_GRAD: 1 = overwrite, 0 = accumulate (TiPi's `clr` flag)
    // z-sharded data term (wfm_conv_create_multi): where the y pass / the fused z pass deliver their output
    ConvPeers<T> peers;
};

template <typename T, int STORE>
WFM_DEVI void conv_store(const ConvArgs<T>& a, size_t idx, cx<T> v, double& cost_acc) {
    if constexpr (STORE == CS_CPLX) {
        a.V[idx] = v;
    } else if constexpr (STORE == CS_SPECTRUM) {
        a.Xout[idx] = v;
    } else if constexpr (STORE == CS_MULX_CONJ) {
        const cx<T> p = cmul(v, a.X[idx]);
        a.V[idx] = mkc<T>(p.x, -p.y);
    } else if constexpr (STORE == CS_MULCX_CONJ) {
        const cx<T> x = a.X[idx];
        const cx<T> p = cmul(v, mkc<T>(x.x, -x.y));
        a.V[idx] = mkc<T>(p.x, -p.y);
    } else if constexpr (STORE == CS_RESID) {
        const double r = (double)v.x * a.inv_ntot - (double)a.y[idx];
        const double wv = a.w ? (double)a.w[idx] : 1.0;
        cost_acc += wv * r * r;
        a.V[idx] = mkc<T>((T)(wv * r), (T)0);
    } else {   // CS_GRAD
        const T g = (T)(a.alpha * ((double)v.x * a.inv_ntot));
        a.grad[idx] = a.clear_grad ? g : (T)(a.grad[idx] + g);
    }
}

// sum over the CTA of each thread's cost contribution -> cost_part[blockIdx linear]
WFM_DEVI void conv_cost_reduce(double v, double* out) {
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        const int nw = (blockDim.x + 31) / 32;
        for (int i = 0; i < nw; ++i) s += red[i];
        *out = s;
    }
}

// ---- x pass: one row per TT-thread group, RB rows per CTA -------------------------------------------
template <typename T, int N, int LOAD, int STORE>
__global__ void __launch_bounds__(RowCfg<N>::THREADS) k_conv_rows(ConvArgs<T> a) {
    using P = Plan<N>;
    using L = RowLayout<T, N>;
    constexpr int RB = RowCfg<N>::RB, TT = P::T, E = P::E;
    WFM_DYN_SMEM(cx<T>, cells);
    cx<T>* tw_s = cells + RB * L::LEN;
    for (int i = threadIdx.x; i < N; i += RowCfg<N>::THREADS) tw_s[i] = a.tw[i];
    if (threadIdx.x < P::R3) tw_s[N + threadIdx.x] = a.tw[P::R1 * threadIdx.x];
    __syncthreads();
    const int slot = threadIdx.x / TT, t = threadIdx.x % TT;
    const size_t nrows = (size_t)a.ny * a.nz;
    const size_t row = (size_t)blockIdx.x * RB + slot;
    const bool valid = row < nrows;
    const size_t base = (valid ? row : 0) * N;
    cx<T> v[E];
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r) {
            const int x = (t + TT * u) + P::S1 * r;
            if constexpr (LOAD == CL_REAL) v[u * P::R1 + r] = mkc<T>(valid ? a.real_in[base + x] : (T)0, (T)0);
            else v[u * P::R1 + r] = valid ? a.V[base + x] : mkc<T>((T)0, (T)0);
        }
    fft_inplace<T, P, L, RowSync<TT>>(v, cells + slot * L::LEN, t, tw_s, tw_s + N, slot);
    double cost_acc = 0.0;
    if (valid) {
#pragma unroll
        for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
            for (int r = 0; r < P::RL; ++r)
                conv_store<T, STORE>(a, base + (t + TT * u) + P::SL * r, v[u * P::RL + r], cost_acc);
    }
    if constexpr (STORE == CS_RESID) conv_cost_reduce(cost_acc, &a.cost_part[blockIdx.x]);
}

// ---- x pass on real data -------------------------------------------------------------------------------
__host__ __device__ constexpr int conv_pitch(int nx) { return nx / 2 + 8; }

template <int M> struct ConvRowSmem {
    template <typename T> static constexpr size_t bytes() {
        return sizeof(cx<T>) * ((size_t)RowCfg<M>::RB * RowLayout<T, M>::LEN + 2 * (size_t)M + 16);
    }
};

// real row -> half spectrum.  STORE: CS_CPLX (-> V) or CS_SPECTRUM (-> Xout)
//   Z = FFT_M(z), z[j] = h[2j] + i h[2j+1];  H[k] = 1/2 [(Z[k] + conj Z[M-k]) - i W_N^k (Z[k] - conj Z[M-k])], k < M;
//   H[M] = Re Z[0] - Im Z[0]
template <typename T, int N, int STORE>
__global__ void __launch_bounds__(RowCfg<N / 2>::THREADS) k_conv_rows_r2c(ConvArgs<T> a) {
    constexpr int M = N / 2;
    using P = Plan<M>;
    using L = RowLayout<T, M>;
    constexpr int RB = RowCfg<M>::RB, TT = P::T, E = P::E;
    WFM_DYN_SMEM(cx<T>, cells);
    cx<T>* tw_s = cells + RB * L::LEN;
    cx<T>* twn_s = tw_s + M + 16;
    const int slot = threadIdx.x / TT, t = threadIdx.x % TT;
    const size_t nrows = (size_t)a.ny * a.nz;
    const size_t row = (size_t)blockIdx.x * RB + slot;
    const bool valid = row < nrows;
    const cx<T>* src = reinterpret_cast<const cx<T>*>(a.real_in + (valid ? row : 0) * N);
    cx<T> v[E];                                        // the row's loads go out first: the twiddle copy overlaps them
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r)
            v[u * P::R1 + r] = valid ? src[(t + TT * u) + P::S1 * r] : mkc<T>((T)0, (T)0);
    for (int i = threadIdx.x; i < M; i += RowCfg<M>::THREADS) { tw_s[i] = a.tw[i]; twn_s[i] = a.twn[i]; }
    if (threadIdx.x < P::R3) tw_s[M + threadIdx.x] = a.tw[P::R1 * threadIdx.x];
    __syncthreads();
    cx<T>* sm = cells + slot * L::LEN;
    fft_inplace<T, P, L, RowSync<TT>>(v, sm, t, tw_s, tw_s + M, slot);
    RowSync<TT>::sync(slot);                           // the last stage has read the cells: reuse them for the mirror
    // Z[k] goes to cell (M - k) mod M of the UNPADDED row, so that the partner reads cell k: both accesses are runs of
    // consecutive cells across lanes (descending / ascending) -- conflict free.  (Storing at the padded cell of k and
    // reading the padded cell of M - k made every 8-lane group straddle a padding step: two wavefronts per read, 17 % of
    // the shared wavefronts of this kernel, ncu profiles/r01n.)
#pragma unroll
    for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
#ifdef WFM_CONV_MIRROR_PADDED   /* A/B knob: the round-1 exchange */
        for (int r = 0; r < P::RL; ++r) sm[L::at((t + TT * u) + P::SL * r)] = v[u * P::RL + r];
#else
        for (int r = 0; r < P::RL; ++r) sm[(M - ((t + TT * u) + P::SL * r)) & (M - 1)] = v[u * P::RL + r];
#endif
    RowSync<TT>::sync(slot);
    if (!valid) return;
    cx<T>* out = (STORE == CS_SPECTRUM ? a.Xout : a.V) + row * (size_t)conv_pitch(N);
#pragma unroll
    for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
        for (int r = 0; r < P::RL; ++r) {
            const int k = (t + TT * u) + P::SL * r;
#ifdef WFM_CONV_MIRROR_PADDED
            const cx<T> Z = v[u * P::RL + r], Zm = sm[L::at((M - k) & (M - 1))], w = twn_s[k];
#else
            const cx<T> Z = v[u * P::RL + r], Zm = sm[k], w = twn_s[k];
#endif
            const T sx = Z.x + Zm.x, sy = Z.y - Zm.y;              // Z + conj(Zm)
            const T dx = Z.x - Zm.x, dy = Z.y + Zm.y;              // Z - conj(Zm)
            const T px = w.x * dy + w.y * dx, py = w.y * dy - w.x * dx;   // -i w (Z - conj Zm)
            out[k] = mkc<T>((T)0.5 * (sx + px), (T)0.5 * (sy + py));
            if (k == 0) out[M] = mkc<T>(Z.x - Z.y, (T)0);
        }
    for (int i = t; i < 7; i += TT) out[M + 1 + i] = mkc<T>((T)0, (T)0);   // padding up to the pitch
}

// half spectrum -> real row, fed with V = conj(G) (the inverse passes run forward transforms on conjugates):
//   c[k] = (V[k] + conj V[M-k]) - i W_N^k (V[k] - conj V[M-k]);  w = FFT_M(c);  g[2j] = Re w[j], g[2j+1] = -Im w[j]
// STORE: CS_RESID (r = g/Ntot - y; cost += w r^2; resid = w r) or CS_GRAD (grad = alpha g/Ntot)
template <typename T, int N, int STORE>
__global__ void __launch_bounds__(RowCfg<N / 2>::THREADS) k_conv_rows_c2r(ConvArgs<T> a) {
    constexpr int M = N / 2;
    using P = Plan<M>;
    using L = RowLayout<T, M>;
    constexpr int RB = RowCfg<M>::RB, TT = P::T, E = P::E;
    WFM_DYN_SMEM(cx<T>, cells);
    cx<T>* tw_s = cells + RB * L::LEN;
    cx<T>* twn_s = tw_s + M + 16;
    for (int i = threadIdx.x; i < M; i += RowCfg<M>::THREADS) { tw_s[i] = a.tw[i]; twn_s[i] = a.twn[i]; }
    if (threadIdx.x < P::R3) tw_s[M + threadIdx.x] = a.tw[P::R1 * threadIdx.x];
    __syncthreads();
    const int slot = threadIdx.x / TT, t = threadIdx.x % TT;
    const size_t nrows = (size_t)a.ny * a.nz;
    const size_t row = (size_t)blockIdx.x * RB + slot;
    const bool valid = row < nrows;
    const cx<T>* in = a.V + (valid ? row : 0) * (size_t)conv_pitch(N);
    cx<T> v[E];
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r) {
            const int k = (t + TT * u) + P::S1 * r;
            cx<T> c = mkc<T>((T)0, (T)0);
            if (valid) {
                const cx<T> V = in[k], Vm = in[M - k], w = twn_s[k];
                const T sx = V.x + Vm.x, sy = V.y - Vm.y;
                const T dx = V.x - Vm.x, dy = V.y + Vm.y;
                c = mkc<T>(sx + (w.x * dy + w.y * dx), sy + (w.y * dy - w.x * dx));
            }
            v[u * P::R1 + r] = c;
        }
    fft_inplace<T, P, L, RowSync<TT>>(v, cells + slot * L::LEN, t, tw_s, tw_s + M, slot);
    double cost_acc = 0.0;
    if (valid) {
        const size_t base = row * N;
#pragma unroll
        for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
            for (int r = 0; r < P::RL; ++r) {
                const int j = (t + TT * u) + P::SL * r;
                const cx<T> wv = v[u * P::RL + r];
                const double g0 = (double)wv.x * a.inv_ntot, g1 = -(double)wv.y * a.inv_ntot;
                const size_t idx = base + 2 * (size_t)j;
                if constexpr (STORE == CS_RESID) {
                    const cx<T> yv = *reinterpret_cast<const cx<T>*>(a.y + idx);
                    cx<T> wt = mkc<T>((T)1, (T)1);
                    if (a.w) wt = *reinterpret_cast<const cx<T>*>(a.w + idx);
                    const double r0 = g0 - (double)yv.x, r1 = g1 - (double)yv.y;
                    cost_acc += (double)wt.x * r0 * r0 + (double)wt.y * r1 * r1;
                    *reinterpret_cast<cx<T>*>(a.resid + idx) = mkc<T>((T)((double)wt.x * r0), (T)((double)wt.y * r1));
                } else {
                    cx<T> g = mkc<T>((T)(a.alpha * g0), (T)(a.alpha * g1));
                    if (!a.clear_grad) {
                        const cx<T> old = *reinterpret_cast<const cx<T>*>(a.grad + idx);
                        g = mkc<T>(old.x + g.x, old.y + g.y);
                    }
                    *reinterpret_cast<cx<T>*>(a.grad + idx) = g;
                }
            }
    }
    if constexpr (STORE == CS_RESID) conv_cost_reduce(cost_acc, &a.cost_part[blockIdx.x]);
}

// ---- y / z pass: CW adjacent columns of length LEN, sample stride `stride` elements ---------------------
// tile -> base: tiles_per_outer tiles of CW columns inside each of `nouter` blocks of `outer_stride` elements
template <typename T, int LEN> struct ConvColCfg {
    static constexpr int TT = Plan<LEN>::T;
#ifndef WFM_CONV_CW_LONG
#define WFM_CONV_CW_LONG 8   /* columns per tile when the transform takes >= 64 threads: full 128-byte lines (eval_fg 3.53 -> 3.29 ms against 4 columns) */
#endif
    static constexpr int CW = (sizeof(T) == 8) ? (TT >= 64 ? (Plan<LEN>::E == 8 ? WFM_CONV_CW_LONG : 4) : 8) : 8;   // (E = 16 plans: 4, registers)
    static constexpr int THREADS = CW * TT;
    static constexpr int SH = ilog2_c(Plan<LEN>::S1);
    using ColL = ColLayout<CW, SH>;
    static constexpr int CELLS = CW * (ColL::pad_c(LEN - 1) + 1);
    static constexpr size_t SMEM = sizeof(cx<T>) * (size_t)(CELLS + LEN + 16);
    // k_conv_cols_zz took 88-92 registers = 2 CTAs per SM (ncu, profiles/r01n_conv_ncu_full.md: warps active 24 %,
    // DRAM 42 %); three resident CTAs where the tile allows it
#ifndef WFM_CONV_ZZ_PREFETCH
#define WFM_CONV_ZZ_PREFETCH 1
#endif
#ifndef WFM_CONV_ZZ_MINB
#define WFM_CONV_ZZ_MINB 3
#endif
    static constexpr int MINB_ZZ = (THREADS <= 256 && LEN <= 256) ? WFM_CONV_ZZ_MINB : 1;
};

// SCATTER (y pass of a z-sharded volume, STORE == CS_CPLX): output row ky of local plane zl goes to the pencil
// volume of the device that owns ky, at [z0 + zl][ky - y0_owner][kx].
template <typename T, int LEN, int STORE, bool SCATTER = false>
__global__ void __launch_bounds__(ConvColCfg<T, LEN>::THREADS) k_conv_cols(ConvArgs<T> a, size_t stride, int tiles_per_outer,
                                                                            size_t outer_stride) {
    using P = Plan<LEN>;
    using Cfg = ConvColCfg<T, LEN>;
    using L = typename Cfg::ColL;
    constexpr int CW = Cfg::CW, TT = P::T, E = P::E;
    WFM_DYN_SMEM(cx<T>, cells);
    cx<T>* tw_s = cells + Cfg::CELLS;
    const int c = threadIdx.x % CW, t = threadIdx.x / CW;
    const size_t base = (size_t)(blockIdx.x / tiles_per_outer) * outer_stride + (size_t)(blockIdx.x % tiles_per_outer) * CW + c;
    cx<T> v[E];                                        // the tile's loads go out first: the twiddle copy overlaps them
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r) v[u * P::R1 + r] = a.V[base + (size_t)((t + TT * u) + P::S1 * r) * stride];
    for (int i = threadIdx.x; i < LEN; i += Cfg::THREADS) tw_s[i] = a.tw[i];
    if (threadIdx.x < P::R3) tw_s[LEN + threadIdx.x] = a.tw[P::R1 * threadIdx.x];
    __syncthreads();
    fft_inplace<T, P, L, CtaSync>(v, cells + c, t, tw_s, tw_s + LEN, 0);
    double cost_acc = 0.0;
    if constexpr (SCATTER) {
        static_assert(STORE == CS_CPLX, "scatter stores plain values");
        const int zl = blockIdx.x / tiles_per_outer;
        const size_t kx = (size_t)(blockIdx.x % tiles_per_outer) * CW + c;
        const size_t zg = (size_t)(a.peers.src_first + zl);
#pragma unroll
        for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
            for (int r = 0; r < P::RL; ++r) {
                const int ky = (t + TT * u) + P::SL * r;
                const int o = a.peers.split.owner(ky);
                const size_t nyl = (size_t)a.peers.split.count(o), kyl = (size_t)(ky - a.peers.split.first(o));
                a.peers.vol[o][(zg * nyl + kyl) * stride + kx] = v[u * P::RL + r];      // `stride` = row pitch of the volume
            }
    } else {
#pragma unroll
        for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
            for (int r = 0; r < P::RL; ++r)
                conv_store<T, STORE>(a, base + (size_t)((t + TT * u) + P::SL * r) * stride, v[u * P::RL + r], cost_acc);
        if constexpr (STORE == CS_RESID) conv_cost_reduce(cost_acc, &a.cost_part[blockIdx.x]);
    }
}

// ---- z pass there and back: FFT_z, spectral product, conjugate, FFT_z again without leaving the SM ---------
// MUL = CS_MULX_CONJ: V <- FFT_z(conj(FFT_z(V) * X));  CS_MULCX_CONJ: V <- FFT_z(conj(FFT_z(V) * conj(X))).
// Saves one write + read of the work volume per transform pair; the only extra cost is one exchange through the
// tile's shared cells (the first transform leaves its output in output-slot order, the second wants input-slot order).
// SCATTER (z-sharded volume): the kernel works on this device's pencil volume [Nz][nyl][P]; output plane z of
// pencil row kyl goes to the slab volume of the device that owns z, at [z - z0_owner][y0 + kyl][kx].
template <typename T, int LEN, int MUL, bool SCATTER = false>
__global__ void __launch_bounds__(ConvColCfg<T, LEN>::THREADS, ConvColCfg<T, LEN>::MINB_ZZ) k_conv_cols_zz(ConvArgs<T> a, size_t stride, int tiles_per_outer,
                                                                               size_t outer_stride) {
    using P = Plan<LEN>;
    using Cfg = ConvColCfg<T, LEN>;
    using L = typename Cfg::ColL;
    constexpr int CW = Cfg::CW, TT = P::T, E = P::E;
    WFM_DYN_SMEM(cx<T>, cells);
    cx<T>* tw_s = cells + Cfg::CELLS;
    const int c = threadIdx.x % CW, t = threadIdx.x / CW;
    const size_t base = (size_t)(blockIdx.x / tiles_per_outer) * outer_stride + (size_t)(blockIdx.x % tiles_per_outer) * CW + c;
    cx<T>* sm = cells + c;
    cx<T> v[E];                                        // the tile's loads go out first: the twiddle copy overlaps them
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r) v[u * P::R1 + r] = a.V[base + (size_t)((t + TT * u) + P::S1 * r) * stride];
#if WFM_CONV_ZZ_PREFETCH
    // the tile of X = FFT3(obj) the spectral product will read after the first transform: DRAM -> L2 now
    for (int k = threadIdx.x; k < LEN; k += Cfg::THREADS)
        wfm_prefetch_l2(&a.X[base - c + (size_t)k * stride], (unsigned)(CW * sizeof(cx<T>)));
#endif
    for (int i = threadIdx.x; i < LEN; i += Cfg::THREADS) tw_s[i] = a.tw[i];
    if (threadIdx.x < P::R3) tw_s[LEN + threadIdx.x] = a.tw[P::R1 * threadIdx.x];
    __syncthreads();
    fft_inplace<T, P, L, CtaSync>(v, sm, t, tw_s, tw_s + LEN, 0);
    __syncthreads();                                   // the last stage has read the cells
#pragma unroll
    for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
        for (int r = 0; r < P::RL; ++r) {
            const int k = (t + TT * u) + P::SL * r;
            cx<T> x = a.X[base + (size_t)k * stride];
            if constexpr (MUL == CS_MULCX_CONJ) x.y = -x.y;
            const cx<T> p = cmul(v[u * P::RL + r], x);
            sm[L::at(k)] = mkc<T>(p.x, -p.y);
        }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r) v[u * P::R1 + r] = sm[L::at((t + TT * u) + P::S1 * r)];
    __syncthreads();                                   // everybody holds its inputs: the cells may be overwritten
    fft_inplace<T, P, L, CtaSync>(v, sm, t, tw_s, tw_s + LEN, 0);
    if constexpr (SCATTER) {
        const size_t pitch = (size_t)conv_pitch(a.nx);
        const size_t cell = (size_t)(blockIdx.x % tiles_per_outer) * CW + c;     // (kyl, kx) inside the pencil plane
        const size_t kyl = cell / pitch, kx = cell - kyl * pitch;
        const size_t row = ((size_t)a.peers.src_first + kyl) * pitch + kx;       // the same entry inside a slab plane
        const size_t plane = (size_t)a.peers.ny_full * pitch;
#pragma unroll
        for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
            for (int r = 0; r < P::RL; ++r) {
                const int z = (t + TT * u) + P::SL * r;
                const int o = a.peers.split.owner(z);
                a.peers.vol[o][(size_t)(z - a.peers.split.first(o)) * plane + row] = v[u * P::RL + r];
            }
    } else {
#pragma unroll
        for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
            for (int r = 0; r < P::RL; ++r) a.V[base + (size_t)((t + TT * u) + P::SL * r) * stride] = v[u * P::RL + r];
    }
}

// fixed-order sum of the per-CTA partials: cost = alpha/2 * sum
// (1024 threads, four independent partial sums each: one CTA of 256 threads with a single dependent chain took 32 us)
__global__ void __launch_bounds__(1024) k_conv_cost_final(const double* __restrict__ part, int n, double alpha,
                                                          double* __restrict__ out) {
    __shared__ double red[32];
    const int bd = blockDim.x;
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
    int i = threadIdx.x;
    for (; i + 3 * bd < n; i += 4 * bd) { x0 += part[i]; x1 += part[i + bd]; x2 += part[i + 2 * bd]; x3 += part[i + 3 * bd]; }
    for (; i < n; i += bd) x0 += part[i];
    double x = (x0 + x1) + (x2 + x3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x / 32); ++w) s += red[w];
        out[0] = 0.5 * alpha * s;
    }
}

}  // namespace wfm
