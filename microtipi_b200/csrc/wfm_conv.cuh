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
// 3-D FFTs are separable passes over a complex work volume V[Nz][Ny][Nx]:
//   x pass : rows, RowLayout, 4 rows per 256-thread CTA, named barrier per row (same engine as the PSF path)
//   y pass : columns inside a plane, ColLayout tiles of CW adjacent x
//   z pass : columns of the [Nz][Npix] matrix, ColLayout tiles of CW adjacent pixels
// Inverse transforms use IFFT(v) = conj(FFT(conj(v)))/Ntot; the conjugations, the spectral products
// with X = FFT3(obj), the residual, the weights, the cost reduction and the final real part are all
// fused into the load / store of the neighbouring passes, so one evaluation is 12 sweeps of the volume.
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

template <typename T> struct ConvArgs {
    cx<T>* V;             // work volume
    const T* real_in;     // CL_REAL source
    const cx<T>* X;       // spectrum of the object
    cx<T>* Xout;          // CS_SPECTRUM destination
    const T* y; const T* w;   // data, weights (w may be NULL = 1)
    T* grad;              // CS_GRAD destination
    double* cost_part;    // [gridDim.x] per-CTA partial sums of w r^2
    const cx<T>* tw;      // twiddles of this pass's length
    int nx, ny, nz;
    double inv_ntot, alpha;
    int clear_grad;       // CS_GRAD: 1 = overwrite, 0 = accumulate (TiPi's `clr` flag)
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

// ---- y / z pass: CW adjacent columns of length LEN, sample stride `stride` elements ---------------------
// tile -> base: tiles_per_outer tiles of CW columns inside each of `nouter` blocks of `outer_stride` elements
template <typename T, int LEN> struct ConvColCfg {
    static constexpr int TT = Plan<LEN>::T;
    static constexpr int CW = (sizeof(T) == 8) ? (TT >= 64 ? 4 : 8) : 8;
    static constexpr int THREADS = CW * TT;
    static constexpr int SH = ilog2_c(Plan<LEN>::S1);
    using ColL = ColLayout<CW, SH>;
    static constexpr int CELLS = CW * (ColL::pad_c(LEN - 1) + 1);
    static constexpr size_t SMEM = sizeof(cx<T>) * (size_t)(CELLS + LEN + 16);
};

template <typename T, int LEN, int STORE>
__global__ void __launch_bounds__(ConvColCfg<T, LEN>::THREADS) k_conv_cols(ConvArgs<T> a, size_t stride, int tiles_per_outer,
                                                                            size_t outer_stride) {
    using P = Plan<LEN>;
    using Cfg = ConvColCfg<T, LEN>;
    using L = typename Cfg::ColL;
    constexpr int CW = Cfg::CW, TT = P::T, E = P::E;
    WFM_DYN_SMEM(cx<T>, cells);
    cx<T>* tw_s = cells + Cfg::CELLS;
    for (int i = threadIdx.x; i < LEN; i += Cfg::THREADS) tw_s[i] = a.tw[i];
    if (threadIdx.x < P::R3) tw_s[LEN + threadIdx.x] = a.tw[P::R1 * threadIdx.x];
    __syncthreads();
    const int c = threadIdx.x % CW, t = threadIdx.x / CW;
    const size_t base = (size_t)(blockIdx.x / tiles_per_outer) * outer_stride + (size_t)(blockIdx.x % tiles_per_outer) * CW + c;
    cx<T> v[E];
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r) v[u * P::R1 + r] = a.V[base + (size_t)((t + TT * u) + P::S1 * r) * stride];
    fft_inplace<T, P, L, CtaSync>(v, cells + c, t, tw_s, tw_s + LEN, 0);
    double cost_acc = 0.0;
#pragma unroll
    for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
        for (int r = 0; r < P::RL; ++r)
            conv_store<T, STORE>(a, base + (size_t)((t + TT * u) + P::SL * r) * stride, v[u * P::RL + r], cost_acc);
    if constexpr (STORE == CS_RESID) conv_cost_reduce(cost_acc, &a.cost_part[blockIdx.x]);
}

// fixed-order sum of the per-CTA partials: cost = alpha/2 * sum
__global__ void k_conv_cost_final(const double* __restrict__ part, int n, double alpha, double* __restrict__ out) {
    __shared__ double red[8];
    double x = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) x += part[i];
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
