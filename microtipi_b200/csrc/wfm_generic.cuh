// wfm_generic.cuh -- any-N fallback of the PSF / Jacobian path.
//
// The reference builds `new DoubleFFT_2D(Nx, Ny)` for ANY Nx == Ny (WFM:319, 604, 917, 1241; JTransforms switches to
// mixed-radix / Bluestein plans off the powers of two).  The pipelines of wfm_kernels.cuh are radix-2/4/8/16 plans
// for N in {32 ... 2048}; every other N in [2, 4096] runs here: the same pruned row-column decomposition, with each
// 1-D transform written as the DFT sum it is, in double precision, over exact twiddles W_N^m (m = (n*k) mod N from
// the N-entry table the host computes in long double).  O(N) work per output instead of O(log N): a functional
// path, not a fast one -- every BASELINE shape is a power of two and takes the pipelines.
//
//   computePsf (WFM:280-350):  S = rho*cis(phi + defoc*psi) on the active rows x columns
//                              T[ky][xi]  = sum_j  W^(y_j*ky)  S[j][xi]          (columns: active rows in, all ky out)
//                              a[ky][kx]  = sum_i  W^(x_i*kx)  T[ky][i]          (rows: active columns in, all kx out)
//                              cpx = conj(a), psf = |a|^2 * PSFnorm
//   apply_J_*  (WFM:883-965):  U[y][i]    = sum_x  W^(x*x_i)   conj(a)[y][x]*q[y][x]   (rows: all x in, active kx out)
//                              B[j][i]    = sum_y  W^(y*y_j)   U[y][i]                 (columns: all y in, active ky out)
//                              jin / J from B on the pupil strip -> the same Gj / Gm images, reduced by k_jac_reduce
#pragma once
#include "wfm_kernels.cuh"

namespace wfm {

#define WFM_GEN_BX 32
#define WFM_GEN_BY 8

// pupil synthesis on the active rows x columns of one plane per blockIdx.z  (WFM:311-316)
__global__ void k_gen_synth(double2* __restrict__ S, const double* __restrict__ rho, const double* __restrict__ phi,
                            const double* __restrict__ psi, const int* __restrict__ act_x, const int* __restrict__ act_y,
                            int nax, int nay, Geom g, int p0, int single) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nax || j >= nay) return;
    const int pl = p0 + blockIdx.z;
    const int in = act_x[i] + g.N * act_y[j];
    const double r = rho[in];
    double2 v = make_double2(0.0, 0.0);
    if (r != 0.0) {
        const double s = defoc_scale_dev(g.z0 + pl, g.nz_global, g.dz);
        const double ph = __dadd_rn(phi[in], __dmul_rn(s, psi[in]));
        double sn, cs;
        sincos(ph, &sn, &cs);
        v = make_double2(__dmul_rn(r, cs), __dmul_rn(r, sn));
        if (single) v = make_double2((double)(float)v.x, (double)(float)v.y);   // cast to float before the transform (WFM:243-245)
    }
    S[((size_t)blockIdx.z * nay + j) * nax + i] = v;
}

// DFT along the SLOW axis of [K][NC] matrices (one per blockIdx.z):
//   out[m][c] = sum_k W_N^(src(k) * frq(m)) in[k][c],   src / frq = index lists (NULL: identity)
__global__ void k_gen_dft_slow(double2* __restrict__ out, const double2* __restrict__ in, const double2* __restrict__ tw,
                               int N, int K, const int* __restrict__ src, int M, const int* __restrict__ frq, int NC) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= NC || m >= M) return;
    const unsigned f = (unsigned)(frq ? frq[m] : m);
    const double2* ip = in + (size_t)blockIdx.z * K * NC + c;
    double ax = 0.0, ay = 0.0;
    for (int k = 0; k < K; ++k) {
        const unsigned s = (unsigned)(src ? src[k] : k);
        const double2 w = tw[(s * f) % (unsigned)N];
        const double2 v = ip[(size_t)k * NC];
        ax += v.x * w.x - v.y * w.y;
        ay += v.x * w.y + v.y * w.x;
    }
    out[((size_t)blockIdx.z * M + m) * NC + c] = make_double2(ax, ay);
}

// PSF row pass + fused store (WFM:323-328): a[ky][kx] = sum_i W^(x_i*kx) T[ky][i]; cpx = conj(a), psf = |a|^2 * PSFnorm
template <typename T>
__global__ void k_gen_psf_rows(cx<T>* __restrict__ cpx, T* __restrict__ psf, const double2* __restrict__ Tm,
                               const double2* __restrict__ tw, const int* __restrict__ act_x, int nax, Geom g, int p0) {
    const int kx = blockIdx.x * blockDim.x + threadIdx.x, ky = blockIdx.y * blockDim.y + threadIdx.y;
    const int N = g.N;
    if (kx >= N || ky >= N) return;
    const double2* ip = Tm + ((size_t)blockIdx.z * N + ky) * nax;
    double ax = 0.0, ay = 0.0;
    for (int i = 0; i < nax; ++i) {
        const double2 w = tw[((unsigned)act_x[i] * (unsigned)kx) % (unsigned)N];
        const double2 v = ip[i];
        ax += v.x * w.x - v.y * w.y;
        ay += v.x * w.y + v.y * w.x;
    }
    const size_t o = ((size_t)(p0 + blockIdx.z) * N + ky) * N + kx;
    const T re = (T)ax, im = (T)ay;                                  // fp32 mode: the transform's result as a float (WFM:250-256)
    cpx[o] = mkc<T>(re, -im);
    if constexpr (sizeof(T) == 8) psf[o] = (T)__dmul_rn(__dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im)), g.psf_norm);
    else psf[o] = (T)__fmul_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)), (T)g.psf_norm);
}

// Jacobian row pass with the fused load Aq = conj(a)*q (WFM:907-914): U[y][i] = sum_x W^(x*x_i) Aq[y][x]
template <typename T>
__global__ void k_gen_jac_rows(double2* __restrict__ U, const cx<T>* __restrict__ cpx, const T* __restrict__ q,
                               const double2* __restrict__ tw, const int* __restrict__ act_x, int nax, Geom g, int p0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int N = g.N;
    if (i >= nax || y >= N) return;
    const size_t base = ((size_t)(p0 + blockIdx.z) * N + y) * N;
    const unsigned f = (unsigned)act_x[i];
    double ax = 0.0, ay = 0.0;
    for (int x = 0; x < N; ++x) {
        const cx<T> a = cpx[base + x];
        const T qv = q[base + x];
        const double vx = (double)(T)(a.x * qv), vy = (double)(T)(a.y * qv);   // the product in the handle's precision
        const double2 w = tw[((unsigned)x * f) % (unsigned)N];
        ax += vx * w.x - vy * w.y;
        ay += vx * w.y + vy * w.x;
    }
    U[((size_t)blockIdx.z * N + y) * nax + i] = make_double2(ax, ay);
}

// masked trig products on the pupil strip (WFM:925-928, 1253, 607-611): B[j][i] -> Gj / Gm at cell (ky = y_j, xi = i)
__global__ void k_gen_jac_trig(double* __restrict__ Gj, double* __restrict__ Gm, const double2* __restrict__ B, Strip st,
                               const int* __restrict__ act_y, int nax, int nay, int pitch, int C, Geom g, int p0,
                               int last_plane_only, int single) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nax || j >= nay) return;
    const int pl = p0 + blockIdx.z;
    const size_t cell = strip_cell(act_y[j], i, g.N, C);
    const unsigned fl = st.flags[cell];
    const int iz = g.z0 + pl;
    const bool mod_plane = (Gm != nullptr) && (!last_plane_only || iz == g.nz_global - 1);
    if (!(fl & (mod_plane ? 3u : 1u))) return;
    double2 b = B[((size_t)blockIdx.z * nay + j) * nax + i];
    if (single) { b.x = (double)(float)b.x; b.y = (double)(float)b.y; }   // fp32 mode: the float transform's output (WFM:781-802)
    const double s = defoc_scale_dev(iz, g.nz_global, g.dz);
    const double ph = __dadd_rn(st.phi[cell], __dmul_rn(s, st.psi[cell]));
    double sn, cs;
    sincos(ph, &sn, &cs);
    const size_t o = (size_t)pl * g.N * pitch + cell;
    if (fl & 1u) Gj[o] = st.rho[cell] * (b.x * sn + b.y * cs);
    if (mod_plane) Gm[o] = b.x * cs - b.y * sn;
}

}  // namespace wfm
