// wfm_kernels.cuh -- CUDA kernels of the widefield PSF path (sm_100a).
//
// WFM = /root/reference/src/microTiPi/epifluorescence/WideFieldModel.java ("para" branches).
//
// Data layout in HBM (TiPi first-index-fastest, see include/wfm_b200.h):
//   rho, phi, psi : double[Npix]            pupil modulus / phase / defocus function
//   mask, support : uint8[Npix]             maskPupil; support = pixels that can be non-zero
//   Z             : double[nzern][Npix]     orthonormal Zernike basis
//   cpx           : cx<T>[nzl][Npix]        conj(FFT2(A_z))            (WFM:325-326)
//   psf           : T[nzl][Npix]            |a|^2 * PSFnorm           (WFM:327)
//   T1 ring       : cx<T>[ring][N][pitch]   column-pass output of the PSF transform (active columns only)
//   T2 ring       : cx<T>[ring][N][pitch]   row-pass output of the adjoint transform (active kx only)
//   Gj, Gm        : double[nzl][N][pitch]   per-plane Jacobian integrands on the compact pupil strip
//
// Pruning (quirk Q6): the pupil is zero outside `support`, so the forward transform only
// row-transforms the `nay` rows that intersect it, and the adjoint transform only keeps /
// column-transforms the `nax` columns that intersect it.  Results are identical.
#pragma once
#include "wfm_fft.cuh"

#include <stdint.h>

namespace wfm {

#define WFM_MAX_COEF 128

struct Coefs { double v[WFM_MAX_COEF]; };

// ---- plane geometry ------------------------------------------------------------------------
struct Geom {
    int N;          // Nx == Ny (WFM:158)
    int nz_global;  // Nz of the whole stack: drives PSFnorm and the wrap rule
    int z0;         // first plane of this slab
    int nzl;        // planes in this handle: the slab, or nbatch stacked models of nzm planes each
    int nzm;        // planes per model (== nzl unless the handle is a batch of independent models)
    double dz;
    double psf_norm;  // 1/(Nx*Ny*Nz)  WFM:284
};

// defoc_scale = DEUXPI*(iz - Nz)*dz or DEUXPI*iz*dz, strict '>' (WFM:302-309); evaluated left to
// right without fused multiply-add, like the JVM.
WFM_DEVI double defoc_scale_dev(int iz, int Nz, double dz) {
    const int zi = (iz > Nz / 2) ? iz - Nz : iz;
    return __dmul_rn(__dmul_rn(6.283185307179586, (double)zi), dz);
}
// defoc = (iz - Nz)*dz or iz*dz (WFM:1220-1229)
WFM_DEVI double defoc_depth_dev(int iz, int Nz, double dz) {
    const int zi = (iz > Nz / 2) ? iz - Nz : iz;
    return __dmul_rn((double)zi, dz);
}
WFM_DEVI int kappa_dev(int n, int N) { return (n > N / 2) ? n - N : n; }
// ---- pupil trig ------------------------------------------------------------------------------------------
// sin / cos of ph = phi + defoc_scale*psi (|ph| up to ~1e3 rad) for every active pixel of every plane, in both
// pipelines.  CUDA's sincos(double) spends ~30 FP64 instructions plus ~30 UMOV / IMAD that materialise its 64-bit
// polynomial coefficients (4.3 % of all issued instructions of k_psf_pipeline, profiles/r01p).  Here: e^{i ph} =
// e^{i 2 pi k/64} * e^{i r} with k from a 64-entry table in shared memory (exact values, computed by the host in long
// double) and |r| <= pi/64, where degree-9 / degree-8 Taylor polynomials are exact to 1e-20; the reduction
// r = ph - n*(2 pi/64) is a two-term Cody-Waite with a 33-bit head (n*C1 exact for |n| < 2^20), so r carries
// no rounding error of its own.  ~19 FP64 instructions and one LDS.128.  |ph| >= 2^20*(2 pi/64) falls back to sincos().
#ifndef WFM_TABLE_CIS
#define WFM_TABLE_CIS 0   /* measured neutral against sincos() (0.8224 / 0.8290 vs 0.8204 / 0.8178 ms per step, A/B/A/B): kept as a knob */
#endif
#define WFM_CIS_ENTRIES 64
WFM_DEVI void wfm_cis(double x, const double2* __restrict__ tab, double* s, double* c) {
    const double t = x * 10.185916357881302;                 // 64 / (2 pi)
    if (!(fabs(t) < 1048576.0)) { sincos(x, s, c); return; }
    const double n = rint(t);
    // 2 pi/64 = C1 + C2, C1 = the leading 33 bits
    double r = fma(-n, 0x1.921fb54400000p-4 /* 0.09817477042088285 */, x);
    r = fma(-n, 0x1.0b4611a626331p-38 /* 3.79818781656637e-12 */, r);
    const double r2 = r * r;
    double ps = fma(r2, 2.7557319223985893e-06, -1.984126984126984e-04);     // 1/9!, -1/7!
    ps = fma(ps, r2, 8.333333333333333e-03);                                  // 1/5!
    ps = fma(ps, r2, -1.6666666666666666e-01);                                // -1/3!
    const double sr = fma(ps * r2, r, r);
    double pc = fma(r2, 2.48015873015873e-05, -1.388888888888889e-03);       // 1/8!, -1/6!
    pc = fma(pc, r2, 4.1666666666666664e-02);                                 // 1/4!
    pc = fma(pc, r2, -0.5);
    const double cr = fma(pc, r2, 1.0);
    const double2 e = tab[(int)n & (WFM_CIS_ENTRIES - 1)];                    // (cos, sin) of 2 pi k/64
    *c = fma(e.x, cr, -(e.y * sr));
    *s = fma(e.x, sr, e.y * cr);
}
#ifdef WFM_FAKE_TRIG   /* profiling experiment only: how much of an item is the trig? */
#define WFM_SINCOS(x, s, c) do { *(s) = (x) * 0.5; *(c) = 1.0 - (x); } while (0)
#elif WFM_TABLE_CIS
#define WFM_SINCOS(x, s, c) wfm_cis((x), cis_s, (s), (c))
#else
#define WFM_SINCOS(x, s, c) sincos((x), (s), (c))
#endif

// ---- launch shapes --------------------------------------------------------------------------
// Row kernels: RB transforms per CTA, 256 threads.  Column kernels: C adjacent columns per CTA.
template <int N> struct RowCfg {
    static constexpr int T = Plan<N>::T;
    static constexpr int RB = (256 / T) < N ? (256 / T) : N;
    static constexpr int THREADS = RB * T;
};
template <typename T, int N> struct ColCfg {
    static constexpr int TT = Plan<N>::T;
    // 8 fp64 (16 fp32) adjacent columns = one 128-byte wavefront; more columns when the transform
    // is short so that a CTA has at least 128 threads; fewer when shared memory would overflow.
    static constexpr int BASE = (sizeof(T) == 8) ? 8 : 16;
    static constexpr int WANT = (128 / TT) > BASE ? (128 / TT) : BASE;
    static constexpr int FIT = (int)((200 * 1024) / (sizeof(cx<T>) * N));
    static constexpr int C0 = WANT < FIT ? WANT : FIT;
    static constexpr int C = C0 >= 32 ? 32 : (C0 >= 16 ? 16 : (C0 >= 8 ? 8 : 4));
    static constexpr int THREADS = C * TT;
    static constexpr size_t SMEM = sizeof(cx<T>) * (size_t)N * C;
};

// strip cell of pixel (ky, compact column xi): tile-major, see "Pupil strip" below
WFM_DEVI size_t strip_cell(int ky, int xi, int N, int C) { return ((size_t)(xi / C) * N + ky) * C + (xi % C); }

// ================================================================================================
// pupil construction (elementwise; arithmetic mirrors the JVM: no FMA contraction)
// ================================================================================================

// computeMaskPupil() WFM:1374-1406
__global__ void k_mask_pupil(uint8_t* __restrict__ map, uint8_t* __restrict__ mask, int N, double dxy,
                             double radius /* NA/lambda */) {
    const int in = blockIdx.x * blockDim.x + threadIdx.x;
    if (in >= N * N) return;
    const int nx = in % N, ny = in / N;
    const double s0 = 1.0 / dxy / (double)N;          // Math.pow(1/dxy/N, 2)
    const double scale = __dmul_rn(s0, s0);
    const double iy = (double)(ny < N - ny ? ny : N - ny);
    const double ix = (double)(nx < N - nx ? nx : N - nx);
    const double ry = __dmul_rn(__dmul_rn(iy, iy), scale);
    const double rx = __dmul_rn(__dmul_rn(ix, ix), scale);
    const uint8_t m = (__dadd_rn(rx, ry) < __dmul_rn(radius, radius)) ? 1 : 0;
    map[in] = m;
    mask[in] = m;
}

// computeDefocus() WFM:1452-1499: only mapPupil pixels are touched.
__global__ void k_compute_defocus(double* __restrict__ psi, uint8_t* __restrict__ mask,
                                  const uint8_t* __restrict__ map, int N, double dxy, double lambda_ni,
                                  double deltaX, double deltaY) {
    const int in = blockIdx.x * blockDim.x + threadIdx.x;
    if (in >= N * N) return;
    if (!map[in]) return;
    const int nx = in % N, ny = in / N;
    const double scale = 1.0 / __dmul_rn((double)N, dxy);
    const double ay = __dsub_rn(__dmul_rn(scale, (double)kappa_dev(ny, N)), deltaY);
    const double ax = __dsub_rn(__dmul_rn(scale, (double)kappa_dev(nx, N)), deltaX);
    const double ry = __dmul_rn(ay, ay), rx = __dmul_rn(ax, ax);
    const double q = __dsub_rn(__dsub_rn(__dmul_rn(lambda_ni, lambda_ni), rx), ry);
    if (q < 0.0) { psi[in] = 0.0; mask[in] = 0; }
    else { psi[in] = __dsqrt_rn(q); mask[in] = 1; }
}

// setPhase() WFM:1625-1649: phi = sum_n Z[in + (n+off)*Npix]*alpha_n on maskPupil, else 0
__global__ void k_set_phase(double* __restrict__ phi, const double* __restrict__ Z,
                            const uint8_t* __restrict__ mask, Coefs alpha, int n, int off, int npix) {
    const int in = blockIdx.x * blockDim.x + threadIdx.x;
    wfm_grid_dep_trigger();
    if (in >= npix) return;
    double acc = 0.0;
    if (mask[in]) {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(Z[in + (size_t)(k + off) * npix], alpha.v[k]));
    }
    phi[in] = acc;
}
// setPhase() when phi is the only pupil array that changed since the strip was packed (the steady state of the
// optimiser loop): one thread per SUPPORT cell (8.7 % of the pixels) writes phi and its strip copy -- same sums, term by
// term; off the support phi is zero already (every earlier setPhase left it so) -- and the step needs no k_pack_strip.
__global__ void k_set_phase_cells(double* __restrict__ phi, double* __restrict__ s_phi, const double* __restrict__ Z,
                                  const uint8_t* __restrict__ mask, Coefs alpha, int n, int off, int npix,
                                  const int* __restrict__ cell_list, const int* __restrict__ in_list, int ncells) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    wfm_grid_dep_trigger();
    if (li >= ncells) return;
    const int in = in_list[li];
    double acc = 0.0;
    if (mask[in]) {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(Z[in + (size_t)(k + off) * npix], alpha.v[k]));
    }
    phi[in] = acc;
    s_phi[cell_list[li]] = acc;
}

// setModulus() WFM:1588-1610: rho = sum_n Z[in + n*Npix]*beta_n*betaNorm on maskPupil, else 0
__global__ void k_set_modulus(double* __restrict__ rho, const double* __restrict__ Z,
                              const uint8_t* __restrict__ mask, Coefs beta, int n, double beta_norm, int npix) {
    const int in = blockIdx.x * blockDim.x + threadIdx.x;
    if (in >= npix) return;
    double acc = 0.0;
    if (mask[in]) {
        for (int k = 0; k < n; ++k)
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(Z[in + (size_t)k * npix], beta.v[k]), beta_norm));
    }
    rho[in] = acc;
}

// Batch variants (handles created by wfm_create_batch: nbatch independent models that share optics and basis, each
// with its own coefficient vectors -- BASELINE config 5).  blockIdx.y = model; coefficients come from device tables;
// the arithmetic is that of the single-model kernels above, term by term.
__global__ void k_set_phase_b(double* __restrict__ phi, const double* __restrict__ Z, const uint8_t* __restrict__ mask,
                              const double* __restrict__ alpha_tab, int n, int off, int npix) {
    const int in = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t mo = (size_t)blockIdx.y * npix;
    wfm_grid_dep_trigger();
    if (in >= npix) return;
    double acc = 0.0;
    if (mask[mo + in]) {
        const double* al = alpha_tab + (size_t)blockIdx.y * n;
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(Z[in + (size_t)(k + off) * npix], al[k]));
    }
    phi[mo + in] = acc;
}
__global__ void k_set_modulus_b(double* __restrict__ rho, const double* __restrict__ Z, const uint8_t* __restrict__ mask,
                                const double* __restrict__ beta_tab, const double* __restrict__ bpar, int n, int npix) {
    const int in = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t mo = (size_t)blockIdx.y * npix;
    if (in >= npix) return;
    double acc = 0.0;
    if (mask[mo + in]) {
        const double* be = beta_tab + (size_t)blockIdx.y * n;
        const double bn = bpar[4 * blockIdx.y + 3];          // 1/|beta| of this model
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(Z[in + (size_t)k * npix], be[k]), bn));
    }
    rho[mo + in] = acc;
}
// par[b] = {ni/lambda, deltaX, deltaY, -}
__global__ void k_compute_defocus_b(double* __restrict__ psi, uint8_t* __restrict__ mask, const uint8_t* __restrict__ map,
                                    int N, double dxy, const double* __restrict__ par) {
    const int in = blockIdx.x * blockDim.x + threadIdx.x;
    if (in >= N * N) return;
    if (!map[in]) return;
    const size_t mo = (size_t)blockIdx.y * N * N;
    const double lambda_ni = par[4 * blockIdx.y], deltaX = par[4 * blockIdx.y + 1], deltaY = par[4 * blockIdx.y + 2];
    const int nx = in % N, ny = in / N;
    const double scale = 1.0 / __dmul_rn((double)N, dxy);
    const double ay = __dsub_rn(__dmul_rn(scale, (double)kappa_dev(ny, N)), deltaY);
    const double ax = __dsub_rn(__dmul_rn(scale, (double)kappa_dev(nx, N)), deltaX);
    const double ry = __dmul_rn(ay, ay), rx = __dmul_rn(ax, ax);
    const double q = __dsub_rn(__dsub_rn(__dmul_rn(lambda_ni, lambda_ni), rx), ry);
    if (q < 0.0) { psi[mo + in] = 0.0; mask[mo + in] = 0; }
    else { psi[mo + in] = __dsqrt_rn(q); mask[mo + in] = 1; }
}

// ================================================================================================
// Zernike basis on the device: Zernike.zernikeArray (Zernike.java:119-288) + Gram-Schmidt (WFM:196)
// ================================================================================================
#define WFM_ZERN_MAXS 16   // radial polynomial terms per mode: supports radial degree n <= 30

struct ZernMode {
    int n, m;
    int kind;          // 0: m == 0, 1: cosine (Noll index even), 2: sine (Noll index odd)
    double norm;       // sqrt(n+1) or sqrt(2(n+1))          Zernike.java:222,245,267
    double R[WFM_ZERN_MAXS];   // (-1)^s (n-s)!/(s!(p-s)!(q-s)!)  Zernike.java:70-90
};

// One thread per pixel, all modes.  r = sqrt(kx^2+ky^2), theta = atan2(ky, kx) in FFT order
// (TiPi MathUtils.fftDist1D / fftAngle1D -- source unavailable, ASSUMED).
__global__ void k_zernike_modes(double* __restrict__ Z, const ZernMode* __restrict__ modes, int nzern,
                                int nmax, int N, double radius_px) {
    const int in = blockIdx.x * blockDim.x + threadIdx.x;
    const int npix = N * N;
    if (in >= npix) return;
    const double kx = (double)kappa_dev(in % N, N), ky = (double)kappa_dev(in / N, N);
    const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)));
    const bool inside = r < radius_px;                                  // strict, Zernike.java:146
    double rP[2 * WFM_ZERN_MAXS];
    rP[0] = inside ? 1.0 : 0.0;
    rP[1] = inside ? __ddiv_rn(r, radius_px) : 0.0;                     // Zernike.java:150
    for (int k = 2; k <= nmax; ++k) rP[k] = __dmul_rn(rP[k - 1], rP[1]);  // Zernike.java:171,205
    const double theta = atan2(ky, kx);
    Z[in] = inside ? 1.0 : 0.0;                                         // piston, Zernike.java:149
    for (int nz = 1; nz < nzern; ++nz) {
        const ZernMode md = modes[nz];
        double zr = 0.0;
        for (int s = (md.n - md.m) / 2; s >= 0; --s) zr = __dadd_rn(zr, __dmul_rn(md.R[s], rP[md.n - 2 * s]));
        double val = __dmul_rn(md.norm, zr);
        if (md.kind == 1) val = __dmul_rn(val, cos(__dmul_rn((double)md.m, theta)));
        else if (md.kind == 2) val = __dmul_rn(val, sin(__dmul_rn((double)md.m, theta)));
        Z[in + (size_t)nz * npix] = val;
    }
}

#define WFM_DOT_THREADS 256
// partial[b] = sum over the block's grid-stride slice of a[i]*b[i]   (warp shuffle, then block)
__global__ void __launch_bounds__(WFM_DOT_THREADS) k_dot_partial(const double* __restrict__ a,
                                                                const double* __restrict__ b, int n,
                                                                double* __restrict__ partial) {
    __shared__ double red[WFM_DOT_THREADS / 32];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += a[i] * b[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0;
        for (int w = 0; w < WFM_DOT_THREADS / 32; ++w) x += red[w];
        partial[blockIdx.x] = x;
    }
}
// mode 0: zk -= dot*zj (Gram-Schmidt projection).  mode 1: zk *= 1/sqrt(dot) (normalisation).
// dot = fixed-order sum of the partials, recomputed by every thread's block leader.
__global__ void k_gs_update(double* __restrict__ zk, const double* __restrict__ zj,
                            const double* __restrict__ partial, int nparts, int n, int mode) {
    __shared__ double dot_s;
    if (threadIdx.x == 0) {
        double x = 0.0;
        for (int b = 0; b < nparts; ++b) x += partial[b];
        dot_s = x;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mode == 0) zk[i] = zk[i] - dot_s * zj[i];
    else zk[i] = zk[i] * (1.0 / sqrt(dot_s));
}

// counter-based splitmix64 uniform(-1,1)  (oracle: splitmix64_uniform)
template <typename T>
__global__ void k_fill_uniform(T* __restrict__ out, uint64_t seed, uint64_t first, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t z = (first + i) * 0x9E3779B97F4A7C15ull + seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    out[i] = (T)__dsub_rn(__dmul_rn(2.0, u), 1.0);
}

// ArrayUtils.roll(psf) (BlindDeconvJob.java:100): the PSF is computed with its origin at voxel (0,0,0); callers
// that want it centred shift every axis by half its length: out[(i + n/2) mod n] = in[i].
template <typename T>
__global__ void k_roll3(T* __restrict__ out, const T* __restrict__ in, int nx, int ny, int nz) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t vox = (size_t)nx * ny * nz;
    if (i >= vox) return;
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((size_t)nx * ny));
    const int xo = (x + nx / 2) % nx, yo = (y + ny / 2) % ny, zo = (z + nz / 2) % nz;
    out[xo + (size_t)nx * (yo + (size_t)ny * zo)] = in[i];
}

// ================================================================================================
// Pupil strip: rho, phi, psi and the mask/support flags gathered into the compact [N][pitch] layout
// of the active columns, so that a column tile reads them as contiguous, independent loads.
// ================================================================================================
// Strip cells are TILE-MAJOR: the C columns of a pipeline column tile are adjacent, then ky, then the
// tile: cell(ky, xi) = ((xi / C) * N + ky) * C + xi % C.  A column item then reads / writes one
// contiguous 16*C*(lanes/C)-byte run per warp access instead of touching one cache line per row
// (ncu, profiles/r01d_*: the 171-column pass cost as many L1 wavefronts as the 512-row pass).
struct Strip {
    const double* rho; const double* phi; const double* psi;   // [pitch/C][N][C]
    const uint8_t* flags;                                        // bit 0: maskPupil, bit 1: support
};
__global__ void k_pack_strip(double* __restrict__ s_rho, double* __restrict__ s_phi, double* __restrict__ s_psi,
                             uint8_t* __restrict__ s_flags, const double* __restrict__ rho,
                             const double* __restrict__ phi, const double* __restrict__ psi,
                             const uint8_t* __restrict__ mask, const uint8_t* __restrict__ support,
                             const int* __restrict__ act_x, int N, int nax, int pitch, int C) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    wfm_grid_dep_trigger();
    wfm_grid_dep_wait();
    if (cell >= N * pitch) return;
    const size_t mo = (size_t)blockIdx.y * N * N, so = (size_t)blockIdx.y * N * pitch;   // model of a batch handle
    const int tile = cell / (N * C), rem = cell % (N * C);
    const int ky = rem / C, xi = tile * C + rem % C;
    double r = 0.0, f = 0.0, p = 0.0;
    uint8_t fl = 0;
    if (xi < nax) {
        const int in = act_x[xi] + N * ky;
        if (support[in]) {
            fl = (uint8_t)(2u | (mask[mo + in] ? 1u : 0u));
            r = rho[mo + in]; f = phi[mo + in]; p = psi[mo + in];
        }
    }
    s_rho[so + cell] = r; s_phi[so + cell] = f; s_psi[so + cell] = p; s_flags[so + cell] = fl;
}

// ================================================================================================
// Persistent two-stage pipelines.
//
// Both directions of the path are "row transform -> pruned intermediate -> column transform":
//   computePsf (WFM:280-350):   A: pupil synthesis + FFT_x of the active rows  -> T ring
//                               B: FFT_y of every column + conj(a), |a|^2 store (HBM stream out)
//   apply_J_*  (WFM:883-965..): A: conj(a)*q load (HBM stream in) + FFT_x, keep active kx -> T ring
//                               B: FFT_y of the active columns + masked trig products -> integrands
// One persistent kernel per direction: CTAs pull work items from a global queue whose order
// interleaves A-items of plane p with B-items of plane p-LAG; B(p) waits on a per-plane counter
// until all A(p) items are published, A(p) waits until the ring slot's previous tenant B(p-RING)
// is finished.  An item only ever waits for items that were dequeued before it, by CTAs that are
// therefore running: no deadlock, no co-residency requirement.  The intermediate lives in a small
// ring (RING planes) that stays L2-resident instead of making an HBM round trip.
// ================================================================================================
struct PipeCtl {
    unsigned* queue;   // [1] next item
    unsigned* err;     // [1] set when a dependency wait times out
    unsigned* done;    // [1] CTAs that have left the work loop (the last one clears the counters)
    unsigned* cntA;    // [nzl] published A-items per plane
    unsigned* cntB;    // [nzl] finished  B-items per plane
    int ring;          // planes in the intermediate ring
    int lag;           // B(p - lag) is queued next to A(p)
    int nA, nB;        // items per plane
    int roles;         // bit 0: run A items, bit 1: run B items (both set in production)
    int nzm;           // planes per model (batch handles), else the slab's plane count
    int ring_stride;   // elements per ring slot (N * pitch)
};

// fp64 column tile of the pipelines for N >= 256: 4 columns (64-byte segments of conj(a)) give
// 256-thread CTAs at N = 512, i.e. four independent barrier domains per SM instead of two.
#ifndef WFM_PIPE_C64
#define WFM_PIPE_C64 4
#endif
#ifndef WFM_PIPE_C32
#define WFM_PIPE_C32 4   /* 256-thread CTAs as in fp64: 0.592 -> 0.561 ms/step at 512^2 fp32 */
#endif
// Row items double-buffer their global loads in registers: the next row's loads are issued before
// the current row's transform, so a row group always has a row in flight (needs ~48 more registers).
#ifndef WFM_ROW_PREFETCH
#define WFM_ROW_PREFETCH 0
#endif
#ifndef WFM_PIPE_KR
#define WFM_PIPE_KR 4
#endif
// The Jacobian pipeline (80 registers) forms its twiddle powers as a tree of depth log2(R) (wfm_fft.cuh).
#ifndef WFM_JAC_TW_TREE
#define WFM_JAC_TW_TREE 1
#endif
#ifndef WFM_PSF_TW_TREE
#define WFM_PSF_TW_TREE 0
#endif
#ifndef WFM_PSF_TWTAB
#define WFM_PSF_TWTAB 0
#endif
#ifndef WFM_JAC_TWTAB
#define WFM_JAC_TWTAB 0
#endif
// the same knobs for the fp32 pipelines (half the shared-memory wavefronts per element: the table reads may pay there)
#ifndef WFM_PSF_TWTAB32
#define WFM_PSF_TWTAB32 0
#endif
#ifndef WFM_JAC_TWTAB32
#define WFM_JAC_TWTAB32 0
#endif
// Consumed ring data is dropped from L2 without write-back (wfm_discard_l2): the ring (44 planes, 63 MB at 512^2 fp64)
// outlives its L2 residency between two tenants of a slot, so without this every ring byte is eventually written to
// DRAM (ncu: 0.31 GB of DRAM writes per Jacobian launch that stores 0.05 GB of integrands).
#ifndef WFM_RING_DISCARD
#define WFM_RING_DISCARD 1
#endif
// Row items of the Jacobian bulk-prefetch their next row of conj(a) and q into L2 (TMA prefetch).
#ifndef WFM_L2_PREFETCH
#define WFM_L2_PREFETCH 1
#endif
template <typename T, int N> struct PipeCfg {
    using P = Plan<N>;
    static constexpr int C = (N >= 256) ? (sizeof(T) == 8 ? WFM_PIPE_C64 : WFM_PIPE_C32) : ColCfg<T, N>::C;   // columns per B-item == rows per A-item
    static_assert(P::T < 64 || C <= 15, "one named barrier per row transform");
    static constexpr int TT = P::T;
    static constexpr int THREADS = C * TT;
    // a row item gives every TT-thread group KR consecutive-in-time rows: the CTA-wide barrier and the
    // queue claim at the item boundary are paid once per C*KR rows
    static constexpr int KR = ((N / C) % WFM_PIPE_KR == 0 && N >= 256) ? WFM_PIPE_KR : 1;
    static constexpr int ROWS_PER_ITEM = C * KR;
    // Jacobian row items at 1024 and up: 2 rows per group (0.462 -> 0.425 ms per launch at 1024^2 x 64, the PSF row items
    // prefer 4: 0.439 vs 0.449 ms)
#ifdef WFM_PIPE_KR_JAC
    static constexpr int KR_JAC = WFM_PIPE_KR_JAC;
#else
    static constexpr int KR_JAC = (N >= 1024 && KR > 2) ? 2 : KR;
#endif
    static constexpr int ROWS_PER_ITEM_JAC = C * KR_JAC;
    static constexpr int SH = ilog2_c(P::S1);
    using ColL = ColLayout<C, SH>;
    static constexpr int ROWLEN = RowLayout<T, N>::LEN;
    static constexpr int COLLEN = ColL::pad_c(N - 1) + 1;
    static constexpr int CELLS = C * (ROWLEN > COLLEN ? ROWLEN : COLLEN);
    // WFM_{PSF,JAC}_TWTAB: twiddle powers read from shared tables instead of formed by multiplication (fft_inplace TWTAB)
    static constexpr int TWTAB_PSF = sizeof(T) == 8 ? (WFM_PSF_TWTAB) : (WFM_PSF_TWTAB32);
    static constexpr int TWTAB_JAC = sizeof(T) == 8 ? (WFM_JAC_TWTAB) : (WFM_JAC_TWTAB32);
    static constexpr int TWTAB_ANY = TWTAB_PSF | TWTAB_JAC;
    // stage-2 twiddles: base entries (R3 <= 16), or every power k in [1, R2) when tabulated
    static constexpr int TW2 = (TWTAB_ANY & 2) ? (((P::R2 - 1) * P::R3 + 1) & ~1) : 16;   // even: what follows stays 16-byte aligned
    // the engine reads tw[b] for b < N/R1 only (stage-1 base twiddles): the shared copy holds just those S1 entries
    // (tabulated: R1 - 1 rows of S1 entries, row k-1 = W_N^(b k))
    static constexpr int TW1 = (TWTAB_ANY & 1) ? (((P::R1 - 1) * P::S1 + 1) & ~1) : P::S1;
    static constexpr size_t SMEM = sizeof(cx<T>) * (size_t)(CELLS + TW1 + TW2) + sizeof(int) * (size_t)N +
                                   sizeof(double2) * WFM_CIS_ENTRIES;         // + the e^{i 2 pi k/64} table of wfm_cis
    // Jacobian row items: conj(a) rows arrive by bulk-async copy (TMA) in a per-group landing buffer, one row ahead
    // (JAC_TMA); enabled where the landing buffers still fit beside MINB_JAC resident CTAs
#ifndef WFM_JAC_TMA
#define WFM_JAC_TMA 0   /* measured slower than register loads at 512^2 fp64 (0.842-0.92 vs 0.810 ms per step, profiles/r02g_tma_variants.md): kept as a knob */
#endif
    // WFM_JAC_TMA: 0 = register loads; 1 = conj(a) rows by bulk copy, q by register loads one row ahead; 2 = conj(a) AND q
    // rows by bulk copy (12 KB of landing buffer per row group at 512^2 fp64: two resident CTAs instead of three)
    static constexpr bool JAC_TMA_Q = (WFM_JAC_TMA == 2);
    static constexpr size_t LANDING = (sizeof(cx<T>) + (JAC_TMA_Q ? sizeof(T) : 0)) * (size_t)C * N + 64;
    // resident CTAs per SM the register allocation is tuned for: 1024 threads (64 registers each)
    static constexpr int BY_THREADS = 1024 / THREADS < 1 ? 1 : (1024 / THREADS > 8 ? 8 : 1024 / THREADS);
    static constexpr int BY_SMEM = (int)((220 * 1024) / SMEM) < 1 ? 1 : (int)((220 * 1024) / SMEM);
    // 16 fp64 complex values per thread are 64 registers of data alone: at three resident 256-thread CTAs (85 registers)
    // the 1024-point kernels spill hundreds of bytes per thread (measured 1.05-1.27 vs 0.92 ms per step at 1024^2 x 64)
    static constexpr int BY_REGS = (sizeof(T) == 8 && P::E >= 16) ? (512 / THREADS < 1 ? 1 : 512 / THREADS) : 8;
#ifdef WFM_PIPE_MINB
    static constexpr int MINB = WFM_PIPE_MINB;
#else
    static constexpr int MINB0 = BY_THREADS < BY_SMEM ? BY_THREADS : BY_SMEM;
    static constexpr int MINB = MINB0 < BY_REGS ? MINB0 : BY_REGS;
#endif
    // The Jacobian pipeline spills at 64 registers (hoisted strip offsets + pointwise part); it is
    // faster with ~85 registers and three quarters of the CTAs (measured 0.410 vs 0.432 ms at 512^2).
#ifdef WFM_PIPE_MINB_JAC
    static constexpr int MINB_JAC = WFM_PIPE_MINB_JAC;
#else
    static constexpr int MINB_JAC0 = (sizeof(T) == 8 && MINB >= 4) ? (3 * MINB) / 4 : MINB;
    static constexpr int MINB_JAC = (JAC_TMA_Q && N == 512 && sizeof(T) == 8) ? 2 : MINB_JAC0;
#endif
    // The narrow Jacobian kernel (4 instead of 8 strip offsets and output predicates per thread) would fit 64
    // registers with 8 bytes of spill, but the full CTA count measured slower (0.841 vs 0.832 ms/step, A/B/A/B).
#ifdef WFM_PIPE_MINB_JAC_NARROW
    template <bool NARROW> static constexpr int minb_jac() { return NARROW ? WFM_PIPE_MINB_JAC_NARROW : MINB_JAC; }
#else
    template <bool NARROW> static constexpr int minb_jac() { return MINB_JAC; }
#endif
    // (N <= 512: at 1024 the landing buffers would cost a resident CTA)
    static constexpr bool JAC_TMA = WFM_JAC_TMA && N >= 256 && N <= 512 &&
                                    (SMEM + LANDING + 1024) * (size_t)MINB_JAC <= 233472;
    static constexpr size_t SMEM_JAC = SMEM + (JAC_TMA ? LANDING : 0);
};

// type 0 = A, 1 = B, -1 = done; model = plane / nzm (batch handles: which model's pupil; 0 otherwise);
// ringoff = (plane % ring) * ring_stride, the plane's slot in the intermediate ring, in elements
struct PipeItem { int type; int plane; int sub; int model; int ringoff; };

WFM_DEVI PipeItem pipe_decode(unsigned idx, int P, const PipeCtl& c) {
    PipeItem it;
    it.model = 0; it.ringoff = 0;
    const int lag = c.lag < P ? c.lag : P;
    const unsigned headA = (unsigned)lag * c.nA;
    const unsigned per = (unsigned)(c.nA + c.nB);
    const unsigned mid = (unsigned)(P - lag) * per;
    if (idx < headA) { it.type = 0; it.plane = (int)(idx / c.nA); it.sub = (int)(idx % c.nA); return it; }
    unsigned r = idx - headA;
    if (r < mid) {
        const int ph = (int)(r / per);
        const int w = (int)(r % per);
        if (w < c.nB) { it.type = 1; it.plane = ph; it.sub = w; }
        else { it.type = 0; it.plane = lag + ph; it.sub = w - c.nB; }
        return it;
    }
    r -= mid;
    if (r < (unsigned)lag * c.nB) { it.type = 1; it.plane = (P - lag) + (int)(r / c.nB); it.sub = (int)(r % c.nB); return it; }
    it.type = -1; it.plane = 0; it.sub = 0;
    return it;
}

#ifndef WFM_WAIT_LIMIT_NS
#define WFM_WAIT_LIMIT_NS 20000000000ull
#endif
// thread 0 polls a counter (L2, volatile) until it reaches `target`; everybody then passes a barrier
WFM_DEVI void pipe_wait(const unsigned* cnt, unsigned target, unsigned* err) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (*(volatile const unsigned*)cnt < target) {
            // a dependency is produced by items dequeued earlier, whose CTAs are running: microseconds.  The limit is
            // wall-clock time (WFM_WAIT_LIMIT_NS, 20 s), read every 4096 polls: an error flag, never a hang
            if ((++spins & 4095u) == 0u) {
                const unsigned long long now = wfm_now_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > WFM_WAIT_LIMIT_NS || *(volatile unsigned*)err) { *(volatile unsigned*)err = 1u; break; }
            }
            WFM_SPIN_PAUSE();
        }
        __threadfence();
    }
    __syncthreads();
}
// CTA barrier (all stores of the item issued), then thread 0 fences at GPU scope and publishes --
// the arrive half of a cooperative-groups grid barrier.
WFM_DEVI void pipe_signal(unsigned* cnt) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(cnt, 1u); }
}
// The same, with the item's consumed ring lines dropped from L2 (wfm_discard_l2) by the first warp between the barrier
// and the publication: drop(lane) issues lane's share; the warp barrier + thread 0's fence order them before the atomic.
template <class Drop> WFM_DEVI void pipe_signal_drop(unsigned* cnt, const Drop& drop) {
    __syncthreads();
    if (threadIdx.x < 32) {
        drop((int)threadIdx.x);
        __syncwarp();
        if (threadIdx.x == 0) { __threadfence(); atomicAdd(cnt, 1u); }
    }
}

// Self-cleaning control block: the last CTA to leave the work loop zeroes the queue and the per-plane
// counters, so the next launch needs no memset (every other CTA has already stopped touching them).
WFM_DEVI void pipe_finish(const PipeCtl& c, int P) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(c.done, 1u) == gridDim.x - 1) {
            for (int p = 0; p < P; ++p) { c.cntA[p] = 0u; c.cntB[p] = 0u; }
            *c.queue = 0u;
            *c.done = 0u;
            __threadfence();
        }
    }
}

// Work-item queue of a persistent CTA.  Thread 0 claims item i+1 LATE in item i (start of the last
// row of a row item / after the first stage of a column item): the atomic's latency is still hidden
// behind the rest of the item, but a CTA holds two claims only for a fraction of the item, so the
// window of claimed-but-unstarted queue entries -- which sets the lag B(p) must trail A(p) by, hence
// the ring size -- is ~1.3 items per CTA instead of 2.  At claim time thread 0 also probes the
// item's dependency counter once (acquire): in steady state it is already met.
struct PipeQueue {
    // shared: per slot {type, plane, sub, model, dependency already satisfied, ring offset}.  Thread 0 decodes the
    // queue index when it claims it (late in the previous item), so the other threads start an item with six
    // shared-memory reads instead of the integer divisions of pipe_decode (measured: 0.8445 -> 0.8335 ms/step for one
    // division), and no thread computes plane % ring: that modulo and the 64-bit slot offset sat in every ROW of the
    // row items (the compiler rematerialised them per row at 64 registers: ~60 of ~430 instructions, cuobjdump).
    int* s;
    int cur;
    bool pre;      // (thread 0) next item already claimed during this item
    static constexpr int SLOT = 6;
    WFM_DEVI static bool probe(const PipeItem& it, const PipeCtl& c) {
        const unsigned* cnt; unsigned target;
        if (it.type == 0) {
            if (it.plane < c.ring || !(c.roles & 2)) return true;
            cnt = &c.cntB[it.plane - c.ring]; target = (unsigned)c.nB;
        } else {
            if (!(c.roles & 1)) return true;
            cnt = &c.cntA[it.plane]; target = (unsigned)c.nA;
        }
        const bool ok = *(volatile const unsigned*)cnt >= target;
        if (ok) __threadfence();
        return ok;
    }
    WFM_DEVI void claim(int slot, const PipeCtl& c, int P) {
        const unsigned idx = atomicAdd(c.queue, 1u);
        const PipeItem it = pipe_decode(idx, P, c);
        int* d = s + slot * SLOT;
        d[0] = it.type; d[1] = it.plane; d[2] = it.sub; d[3] = it.type < 0 ? 0 : it.plane / c.nzm;
        d[4] = (it.type < 0 || probe(it, c)) ? 1 : 0;
        d[5] = it.type < 0 ? 0 : (it.plane % c.ring) * c.ring_stride;
    }
    WFM_DEVI void init(int* smem12, const PipeCtl& c, int P) {
        s = smem12; cur = 0; pre = false;
        if (threadIdx.x == 0) claim(0, c, P);
        __syncthreads();
    }
    // the current item; `ready` tells whether its dependency was already observed as met
    WFM_DEVI PipeItem take(const PipeCtl&, int, bool& ready) {
        pre = false;
        const int* d = s + cur * SLOT;
        PipeItem it;
        it.type = d[0]; it.plane = d[1]; it.sub = d[2]; it.model = d[3]; it.ringoff = d[5];
        ready = d[4] != 0;
        return it;
    }
    // claim the next item (thread 0, once per item; called from inside the item and again, as a
    // no-op, before the item is published)
    WFM_DEVI bool prefetch(const PipeCtl& c, int P) {
        if (threadIdx.x == 0 && !pre) { claim(cur ^ 1, c, P); pre = true; return true; }
        return false;
    }
    // (thread 0, right after a claim) the item just claimed
    WFM_DEVI PipeItem peek_next() const {
        const int* d = s + (cur ^ 1) * SLOT;
        PipeItem it;
        it.type = d[0]; it.plane = d[1]; it.sub = d[2]; it.model = d[3]; it.ringoff = d[5];
        return it;
    }
    WFM_DEVI void advance() { cur ^= 1; }
};

// What an item must wait for before it may touch the ring slot (NULL counter: nothing).
struct PipeDep { const unsigned* cnt; unsigned target; unsigned* err; };
// hook for fft_inplace: claim the next item after the first stage of a column item
struct NoClaimAction { WFM_DEVI void operator()(const PipeItem&) const {} };
template <class OnClaim = NoClaimAction> struct PipePrefetchHook {
    PipeQueue& qu; const PipeCtl& ctl; int P; OnClaim on_claim;
    WFM_DEVI void operator()() const { if (qu.prefetch(ctl, P)) on_claim(qu.peek_next()); }
};
WFM_DEVI void pipe_wait(const PipeDep& d) { if (d.cnt) pipe_wait(d.cnt, d.target, d.err); }

// (slot, t) of a thread of a row item: which row transform of the CTA it belongs to and its index inside it.
// With two warps per transform (TT = 64) and four transforms per CTA the two warps of a transform are w and
// w + 4, i.e. they live on the same SM sub-partition (warp w is scheduled by sub-partition w % 4): the scheduler
// that parks one of them at the exchange barrier is the one running its partner.
#ifndef WFM_ROW_SAME_SMSP
#define WFM_ROW_SAME_SMSP 0
#endif
template <int C, int TT> WFM_DEVI void row_thread_map(int& slot, int& t) {
    if constexpr (WFM_ROW_SAME_SMSP && TT == 64 && C == 4) {
        const int w = threadIdx.x >> 5;
        slot = w & 3;
        t = ((w >> 2) << 5) + (threadIdx.x & 31);
    } else {
        slot = threadIdx.x / TT;
        t = threadIdx.x % TT;
    }
}

// ================================================================================================
// computePsf()  WFM:280-350 (fp32: 209-278)
//
// a = FFT2(A) is evaluated columns first, rows second: the pupil is zero outside `nax` active
// columns, so pass A transforms only those (FFT along y, pupil synthesis fused into the load), and
// pass B -- the big one, all N rows -- runs along the contiguous axis: every warp stores full,
// consecutive 512-byte runs of conj(a) and 256-byte runs of psf straight from registers.
// ================================================================================================
template <typename T> struct PsfArgs {
    Geom g;
    Strip st;           // pupil in strip layout
    const int* inv_x;   // [N]   column -> compact index or -1
    int nax;
    int pitch;          // nax rounded up to a multiple of the column tile
    const cx<T>* tw;    // W_N table (global; copied to shared memory once per CTA)
    const double2* cis; // [64] (cos, sin)(2 pi k/64) for wfm_cis
    cx<T>* T1;          // ring: [ring][N][pitch]
    cx<T>* cpx;
    T* psf;
};

// A-item: active columns xi0 .. xi0+C-1 of plane pl.  A = rho*exp(i(phi + defoc_scale*psi)) is
// synthesised in the load (WFM:311-316; sincos only where rho != 0, quirk Q6), then FFT along y.
// NARROW: every active row and column lies in [0, N/4) u [3N/4, N) -- the first-stage legs 2..5 (of 8) of both
// passes are structurally zero, so they are neither loaded nor transformed (fft_inplace SPARSE1).
// The same holds for the outputs the Jacobian needs: only the last-stage legs r < R/4 or r >= R - R/4 can hit the
// support, so the others are neither stored (row items) nor multiplied out (column items).
template <int R, bool NARROW> WFM_DEVI constexpr bool leg_live(int r) { return !NARROW || r < R / 4 || r >= R - R / 4; }

template <typename T, int N, bool NARROW>
WFM_DEVI void psf_cols_item(const PsfArgs<T>& a, int pl, int sub, int bm, int ringoff, cx<T>* cells, const cx<T>* tw_s,
                            const double2* cis_s, const PipeDep& dep, PipeQueue& qu, const PipeCtl& ctl) {
    using P = Plan<N>;
    constexpr int C = PipeCfg<T, N>::C, TT = P::T, E = P::E;
    using L = typename PipeCfg<T, N>::ColL;
    const int c = threadIdx.x % C, t = threadIdx.x / C;
    // bm: model of a batch handle (0 otherwise): its strip follows
    const int ssub = sub + bm * (a.pitch / C);         // the previous model's, i.e. pitch/C tiles further on
    const double s = defoc_scale_dev(a.g.z0 + (pl - bm * a.g.nzm), a.g.nz_global, a.g.dz);
    double rho[E];
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r)
            if (leg_live<P::R1, NARROW>(r))
                rho[u * P::R1 + r] = __ldg(&a.st.rho[((size_t)ssub * N + (t + TT * u) + P::S1 * r) * C + c]);
    cx<T> v[E];
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u) {
#pragma unroll
        for (int r = 0; r < P::R1; ++r) {
            const int e = u * P::R1 + r;
            cx<T> val = mkc<T>((T)0, (T)0);
            if (leg_live<P::R1, NARROW>(r)) {
                if (rho[e] != 0.0) {
                    const size_t cell = ((size_t)ssub * N + (t + TT * u) + P::S1 * r) * C + c;
                    const double ph = __dadd_rn(__ldg(&a.st.phi[cell]), __dmul_rn(s, __ldg(&a.st.psi[cell])));
                    double sn, cs;
                    WFM_SINCOS(ph, &sn, &cs);
                    val = mkc<T>((T)__dmul_rn(rho[e], cs), (T)__dmul_rn(rho[e], sn));
                }
            }
            v[e] = val;
        }
    }
    fft_inplace<T, P, L, CtaSync, PipePrefetchHook<>, NARROW, WFM_PSF_TW_TREE, NoHook, PipeCfg<T, N>::TWTAB_PSF>(
        v, cells + c, t, tw_s, tw_s + PipeCfg<T, N>::TW1, 0, PipePrefetchHook<>{qu, ctl, a.g.nzl, NoClaimAction{}});
    pipe_wait(dep);                                   // ring slot free? (its previous tenant's row items are done)
    cx<T>* dst = a.T1 + (size_t)ringoff + (size_t)sub * N * C + c;
#pragma unroll
    for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
        for (int r = 0; r < P::RL; ++r) __stcg(&dst[(size_t)((t + TT * u) + P::SL * r) * C], v[u * P::RL + r]);
}

// B-item: ROWS_PER_ITEM rows of plane pl (each TT-thread group walks KR of them): FFT along x
// (inactive columns are zero), then the fused streaming store of conj(a) and |a|^2*PSFnorm
// (WFM:323-328) as full contiguous rows.  No CTA-wide barrier inside.
template <typename T, int N, bool NARROW>
WFM_DEVI void psf_rows_item(const PsfArgs<T>& a, int pl, int sub, int ringoff, cx<T>* cells, const cx<T>* tw_s,
                            const int* invx_s, PipeQueue& qu, const PipeCtl& ctl) {
    using P = Plan<N>;
    using L = RowLayout<T, N>;
    using Cfg = PipeCfg<T, N>;
    constexpr int C = Cfg::C, TT = P::T, E = P::E;
    int slot, t;
    row_thread_map<C, TT>(slot, t);
    const T norm = (T)a.g.psf_norm;
    int xis[E];                                        // strip offset of column x (without the ky term) or -1
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r) {
            if (!leg_live<P::R1, NARROW>(r)) { xis[u * P::R1 + r] = -1; continue; }
            const int xi = invx_s[(t + TT * u) + P::S1 * r];
            xis[u * P::R1 + r] = xi >= 0 ? (xi / C) * N * C + xi % C : -1;
        }
#if WFM_ROW_PREFETCH
    cx<T> nv[E];
    {
        const cx<T>* src0 = a.T1 + (size_t)ringoff + (size_t)(sub * Cfg::ROWS_PER_ITEM + slot) * C;
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (leg_live<P::R1, NARROW>(e % P::R1)) nv[e] = (xis[e] >= 0) ? __ldcg(&src0[xis[e]]) : mkc<T>((T)0, (T)0);
    }
#endif
#pragma unroll 1
    for (int kk = 0; kk < Cfg::KR; ++kk) {
        if (kk == Cfg::KR - 1) qu.prefetch(ctl, a.g.nzl);             // claim the next item behind the last row
        const int ky = sub * Cfg::ROWS_PER_ITEM + kk * C + slot;     // N % ROWS_PER_ITEM == 0
#ifdef WFM_PROBE_PSF_L1LD          /* timing probe only (wrong results): ring loads that hit the L1 */
        const cx<T>* src = a.T1 + (size_t)slot * C;
#else
        const cx<T>* src = a.T1 + (size_t)ringoff + (size_t)ky * C;
#endif
        cx<T> v[E];
#if WFM_ROW_PREFETCH
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = leg_live<P::R1, NARROW>(e % P::R1) ? nv[e] : mkc<T>((T)0, (T)0);
        if (kk + 1 < Cfg::KR) {
            const cx<T>* nsrc = src + (size_t)C * C;
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (leg_live<P::R1, NARROW>(e % P::R1)) nv[e] = (xis[e] >= 0) ? __ldcg(&nsrc[xis[e]]) : mkc<T>((T)0, (T)0);
        }
#else
#pragma unroll
        for (int e = 0; e < E; ++e)
#ifdef WFM_PROBE_PSF_L1LD
            v[e] = (leg_live<P::R1, NARROW>(e % P::R1) && xis[e] >= 0) ? __ldg(&src[xis[e]]) : mkc<T>((T)0, (T)0);
#else
            v[e] = (leg_live<P::R1, NARROW>(e % P::R1) && xis[e] >= 0) ? __ldcg(&src[xis[e]]) : mkc<T>((T)0, (T)0);
#endif
#endif
        fft_inplace<T, P, L, RowSync<TT>, NoHook, NARROW, WFM_PSF_TW_TREE, NoHook, PipeCfg<T, N>::TWTAB_PSF>(v, cells + slot * L::LEN, t, tw_s, tw_s + PipeCfg<T, N>::TW1, slot);
#ifdef WFM_PROBE_PSF_L2ST          /* timing probe only (wrong results): stores that never reach DRAM */
        const size_t base = (size_t)N * (blockIdx.x * C + slot);
#else
        const size_t base = (size_t)pl * N * N + (size_t)N * ky;
#endif
#pragma unroll
        for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
            for (int r = 0; r < P::RL; ++r) {
                const int kx = (t + TT * u) + P::SL * r;
                const cx<T> val = v[u * P::RL + r];
                __stcs(&a.cpx[base + kx], mkc<T>(val.x, -val.y));     // store conjugate of A (WFM:326)
                if constexpr (sizeof(T) == 8)
                    __stcs(&a.psf[base + kx], (T)__dmul_rn(__dadd_rn(__dmul_rn(val.x, val.x), __dmul_rn(val.y, val.y)), norm));
                else
                    __stcs(&a.psf[base + kx], (T)__fmul_rn(__fadd_rn(__fmul_rn(val.x, val.x), __fmul_rn(val.y, val.y)), norm));
            }
        if (kk + 1 < Cfg::KR) RowSync<TT>::sync(slot);    // the next row's stage-1 stores reuse the cells
    }
}

// shared twiddle tables of a pipeline CTA, from the global W_N table (tw[m] = W_N^m)
template <typename T, int N> WFM_DEVI void pipe_fill_twiddles(cx<T>* tw_s, cx<T>* tw2_s, const cx<T>* __restrict__ tw) {
    using Cfg = PipeCfg<T, N>;
    using P = Plan<N>;
    for (int i = threadIdx.x; i < ((Cfg::TWTAB_ANY & 1) ? (P::R1 - 1) * P::S1 : P::S1); i += Cfg::THREADS) {
        const int k = i / P::S1 + 1, b = i % P::S1;          // row k-1: W_N^(b k); row 0 is the base table
        tw_s[i] = tw[(b * k) % N];
    }
    if constexpr (Cfg::TWTAB_ANY & 2) {
        for (int i = threadIdx.x; i < (P::R2 - 1) * P::R3; i += Cfg::THREADS) {
            const int k = i / P::R3 + 1, d3 = i % P::R3;
            tw2_s[i] = tw[(P::R1 * d3 * k) % N];
        }
    } else {
        if (threadIdx.x < P::R3) tw2_s[threadIdx.x] = tw[P::R1 * threadIdx.x];
    }
}

template <typename T, int N, bool NARROW>
__global__ void __launch_bounds__(PipeCfg<T, N>::THREADS, PipeCfg<T, N>::MINB) k_psf_pipeline(PsfArgs<T> a, PipeCtl ctl) {
    using Cfg = PipeCfg<T, N>;
    WFM_DYN_SMEM(cx<T>, cells);
    cx<T>* tw_s = cells + Cfg::CELLS;
    cx<T>* tw2_s = tw_s + Cfg::TW1;
    int* invx_s = reinterpret_cast<int*>(tw2_s + Cfg::TW2);
    double2* cis_s = reinterpret_cast<double2*>(invx_s + N);
    for (int i = threadIdx.x; i < N; i += Cfg::THREADS) invx_s[i] = a.inv_x[i];
    pipe_fill_twiddles<T, N>(tw_s, tw2_s, a.tw);
    if (threadIdx.x < WFM_CIS_ENTRIES) cis_s[threadIdx.x] = a.cis[threadIdx.x];
    wfm_grid_dep_trigger();  // the next kernel's CTAs may take the place of ours as we exit
    wfm_grid_dep_wait();     // the tables above are constants; everything below depends on the previous kernel
    __shared__ int s_queue[2 * PipeQueue::SLOT];
    PipeQueue qu;
    const int P = a.g.nzl;
    qu.init(s_queue, ctl, P);
    for (;;) {
        bool ready;
        PipeItem it = qu.take(ctl, P, ready);
        if (it.type < 0) break;
#ifdef WFM_RING_MOD_IN_ITEM   /* A/B knob: the slot offset computed by every thread, as before */
        it.ringoff = (it.plane % ctl.ring) * ctl.ring_stride;
#endif
        if (it.type == 0) {
            if (ctl.roles & 1) {
                PipeDep dep;
                dep.cnt = ready ? nullptr : &ctl.cntB[it.plane - ctl.ring]; dep.target = (unsigned)ctl.nB; dep.err = ctl.err;
                psf_cols_item<T, N, NARROW>(a, it.plane, it.sub, it.model, it.ringoff, cells, tw_s, cis_s, dep, qu, ctl);
            }
            qu.prefetch(ctl, P);
            pipe_signal(&ctl.cntA[it.plane]);
        } else {
            if (ctl.roles & 2) {
                if (!ready) pipe_wait(&ctl.cntA[it.plane], ctl.nA, ctl.err);
                psf_rows_item<T, N, NARROW>(a, it.plane, it.sub, it.ringoff, cells, tw_s, invx_s, qu, ctl);
#if WFM_RING_DISCARD == 1
                // rows [sub*RPI, (sub+1)*RPI) of every column tile have been read by all groups: drop their lines from L2
                if constexpr ((sizeof(cx<T>) * Cfg::C * Cfg::ROWS_PER_ITEM) % 128 == 0) {
                    __syncthreads();
                    constexpr int LPT = (int)(sizeof(cx<T>) * Cfg::C * Cfg::ROWS_PER_ITEM / 128);     // lines per tile
                    const int ntiles = a.pitch / Cfg::C;
                    const char* base = reinterpret_cast<const char*>(a.T1 + (size_t)it.ringoff + (size_t)it.sub * Cfg::ROWS_PER_ITEM * Cfg::C);
                    for (int i = threadIdx.x; i < ntiles * LPT; i += Cfg::THREADS)
                        wfm_discard_l2(base + (size_t)(i / LPT) * (sizeof(cx<T>) * (size_t)N * Cfg::C) + (size_t)(i % LPT) * 128);
                }
#endif
            }
            qu.prefetch(ctl, P);
#if WFM_RING_DISCARD == 2
            if constexpr ((sizeof(cx<T>) * Cfg::C * Cfg::ROWS_PER_ITEM) % 128 == 0) {
                // rows [sub*RPI, (sub+1)*RPI) of every column tile have been read: drop their lines from L2
                constexpr int LPT = (int)(sizeof(cx<T>) * Cfg::C * Cfg::ROWS_PER_ITEM / 128);     // lines per tile
                const int nl = (a.pitch / Cfg::C) * LPT;
                const char* base = reinterpret_cast<const char*>(a.T1 + (size_t)it.ringoff + (size_t)it.sub * Cfg::ROWS_PER_ITEM * Cfg::C);
                pipe_signal_drop(&ctl.cntB[it.plane], [&](int lane) {
                    if (!(ctl.roles & 2)) return;
                    for (int i = lane; i < nl; i += 32)
                        wfm_discard_l2(base + (size_t)(i / LPT) * (sizeof(cx<T>) * (size_t)N * Cfg::C) + (size_t)(i % LPT) * 128);
                });
            } else {
                pipe_signal(&ctl.cntB[it.plane]);
            }
#else
            pipe_signal(&ctl.cntB[it.plane]);
#endif
        }
        qu.advance();
    }
    pipe_finish(ctl, P);
}

// ================================================================================================
// apply_J_phase / apply_J_defocus / apply_J_modulus   WFM:883-965, 1201-1288, 566-683
// ================================================================================================
template <typename T> struct JacArgs {
    Geom g;
    const cx<T>* cpx;
    const T* q;
    Strip st;          // pupil in strip layout
    const int* inv_x;  // [N]
    int nax;
    int pitch;         // nax rounded up to a multiple of the column tile
    const cx<T>* tw;
    const double2* cis; // [64] (cos, sin)(2 pi k/64) for wfm_cis
    cx<T>* T2;         // ring: [ring][N][pitch]
    double* Gj;        // [nzl][N][pitch]  jin  = rho*(B_re sin ph + B_im cos ph)  on maskPupil
    double* Gm;        // [nzl][N][pitch]  J    = B_re cos ph - B_im sin ph        on the support (or NULL)
    int last_plane_only;  // quirk Q1 compat mode for the modulus Jacobian
};

// Claim action of the Jacobian pipeline (thread 0, late in the previous item): when the item just claimed is a row
// item, its first row of every group -- C consecutive rows of conj(a) and of q -- starts its way from DRAM to L2 now
// (TMA bulk prefetch), so that the first load of the item is an L2 hit like the later rows, whose prefetch is issued
// one row ahead by their own group.  (Probe -DWFM_PROBE_JAC_L2: row loads that never miss L2 are worth 6 % of the kernel.)
#ifndef WFM_CLAIM_PREFETCH
#define WFM_CLAIM_PREFETCH 1
#endif
template <typename T, int N> struct JacClaimPrefetch {
    const cx<T>* cpx; const T* q;
    WFM_DEVI void operator()(const PipeItem& it) const {
#if WFM_CLAIM_PREFETCH && WFM_L2_PREFETCH
        using Cfg = PipeCfg<T, N>;
        if (it.type != 0) return;
        const size_t base = (size_t)it.plane * N * N + (size_t)N * (it.sub * Cfg::ROWS_PER_ITEM_JAC);
#pragma unroll
        for (int r = 0; r < Cfg::C; ++r) {
            wfm_prefetch_l2(&cpx[base + (size_t)N * r], (unsigned)(N * sizeof(cx<T>)));
            wfm_prefetch_l2(&q[base + (size_t)N * r], (unsigned)(N * sizeof(T)));
        }
#endif
    }
};

// A-item: ROWS_PER_ITEM rows of plane pl (each TT-thread group walks KR of them).  Aq = conj(a)*q
// fused into the streaming load (WFM:907-914), FFT along x, keep the active kx only.
// Hook of the row transforms of the TMA variant: right after the first exchange barrier every thread of the group has
// consumed the landing buffer (its stage-1 loads fed the stores that precede the barrier), so the group's elected
// thread posts the next row's bulk copy into it; the copy then has the rest of the transform to arrive.
struct RowBulkLoadHook {
    void* dst; const void* src; unsigned bytes; uint64_t* bar; bool issue;
    void* dst2; const void* src2; unsigned bytes2;     // second copy on the same barrier phase (the q row), bytes2 = 0: none
    WFM_DEVI void operator()() const {
        if (!issue) return;
        wfm_mbar_expect(bar, bytes + bytes2);
        wfm_bulk_copy(dst, src, bytes, bar);
        if (bytes2) wfm_bulk_copy(dst2, src2, bytes2, bar);
        wfm_mbar_complete_emu(bar);
    }
};
// Second hook of the same transforms (before the last stage): the NEXT row's q values start their way into registers,
// so that their L2 latency is covered by the last stage, the stores and the row barrier of this row.
#ifndef WFM_JAC_Q_AHEAD
#define WFM_JAC_Q_AHEAD 1
#endif
template <typename T, int E, int R1, int TT, int S1> struct RowQAheadHook {
    T (&qn)[E]; const T* src; int t; bool issue;
    WFM_DEVI void operator()() const {
        if (!issue) return;
#pragma unroll
        for (int u = 0; u < E / R1; ++u)
#pragma unroll
            for (int r = 0; r < R1; ++r) qn[u * R1 + r] = __ldcs(&src[(t + TT * u) + S1 * r]);
    }
};

template <typename T, int N, bool NARROW>
WFM_DEVI void jac_rows_item(const JacArgs<T>& a, int pl, int sub, int ringoff, cx<T>* cells, const cx<T>* tw_s,
                            const int* invx_s, const PipeDep& dep, PipeQueue& qu, const PipeCtl& ctl,
                            cx<T>* landing, uint64_t* mbar, unsigned& row_phase) {
    using P = Plan<N>;
    using L = RowLayout<T, N>;
    using Cfg = PipeCfg<T, N>;
    constexpr int C = Cfg::C, TT = P::T, E = P::E;
    int slot, t;
    row_thread_map<C, TT>(slot, t);
    T* const landq = reinterpret_cast<T*>(landing + (size_t)C * N);      // (JAC_TMA_Q) q rows land behind the conj(a) rows
    if constexpr (Cfg::JAC_TMA) {
        // first row of the item: nothing of this group is in flight any more (its previous item is complete), so the
        // landing buffer is free; the copy is an L2 hit when the claim-time prefetch was in time
        const size_t base0 = (size_t)pl * N * N + (size_t)N * (sub * Cfg::ROWS_PER_ITEM_JAC + slot);
        RowBulkLoadHook first{landing + (size_t)slot * N, &a.cpx[base0], (unsigned)(N * sizeof(cx<T>)), &mbar[slot], t == 0,
                              landq + (size_t)slot * N, &a.q[base0], Cfg::JAC_TMA_Q ? (unsigned)(N * sizeof(T)) : 0u};
        first();
    }
    pipe_wait(dep);                                    // ring slot free? (rarely taken: probed at claim time)
    int xis[E];                                        // strip offset of column kx (without the y term) or -1
#pragma unroll
    for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
        for (int r = 0; r < P::RL; ++r) {
            if (!leg_live<P::RL, NARROW>(r)) { xis[u * P::RL + r] = -1; continue; }
            const int xi = invx_s[(t + TT * u) + P::SL * r];
            xis[u * P::RL + r] = xi >= 0 ? (xi / C) * N * C + xi % C : -1;
        }
#if WFM_ROW_PREFETCH
    cx<T> nc[E];
    T nq[E];
    {
        const size_t base0 = (size_t)pl * N * N + (size_t)N * (sub * Cfg::ROWS_PER_ITEM_JAC + slot);
#pragma unroll
        for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
            for (int r = 0; r < P::R1; ++r) {
                const int x = (t + TT * u) + P::S1 * r;
                nc[u * P::R1 + r] = __ldcs(&a.cpx[base0 + x]);
                nq[u * P::R1 + r] = __ldcs(&a.q[base0 + x]);
            }
    }
#endif
    T qn[E];                                           // (TMA variant) the next row's q, loaded one row ahead
#pragma unroll
    for (int e = 0; e < E; ++e) qn[e] = (T)0;
#pragma unroll 1
    for (int kk = 0; kk < Cfg::KR_JAC; ++kk) {
        if (kk == Cfg::KR_JAC - 1 && qu.prefetch(ctl, a.g.nzl))           // claim the next item behind the last row
            JacClaimPrefetch<T, N>{a.cpx, a.q}(qu.peek_next());
        const int y = sub * Cfg::ROWS_PER_ITEM_JAC + kk * C + slot;      // N % ROWS_PER_ITEM == 0
#if defined(WFM_PROBE_JAC_L1)      /* timing probes only (wrong results): what if the row loads never left the SM / the L2? */
        const size_t base = 0;
#elif defined(WFM_PROBE_JAC_L2)
        const size_t base = (size_t)N * (blockIdx.x * C + slot);
#else
        const size_t base = (size_t)pl * N * N + (size_t)N * y;
#endif
#if WFM_L2_PREFETCH && !defined(WFM_PROBE_JAC_L1) && !defined(WFM_PROBE_JAC_L2)
        if (t == 0 && kk + 1 < Cfg::KR_JAC) {              // this group's next row: DRAM -> L2 while this row is transformed
            if constexpr (!Cfg::JAC_TMA) wfm_prefetch_l2(&a.cpx[base + (size_t)N * C], (unsigned)(N * sizeof(cx<T>)));
            if constexpr (!Cfg::JAC_TMA_Q) wfm_prefetch_l2(&a.q[base + (size_t)N * C], (unsigned)(N * sizeof(T)));
        }
#endif
        cx<T> v[E];
#if WFM_ROW_PREFETCH
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = mkc<T>(nc[e].x * nq[e], nc[e].y * nq[e]);
        if (kk + 1 < Cfg::KR_JAC) {                        // next row in flight during this row's transform
            const size_t nb = base + (size_t)N * C;
#pragma unroll
            for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
                for (int r = 0; r < P::R1; ++r) {
                    const int x = (t + TT * u) + P::S1 * r;
                    nc[u * P::R1 + r] = __ldcs(&a.cpx[nb + x]);
                    nq[u * P::R1 + r] = __ldcs(&a.q[nb + x]);
                }
        }
#else
        if constexpr (Cfg::JAC_TMA) {
            // q straight from global (its row was prefetched into L2 one row ahead): already in registers for every row
            // but the first of an item (RowQAheadHook); conj(a) from the landing buffer
            T qv[E];
#pragma unroll
            for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
                for (int r = 0; r < P::R1; ++r)
                    if constexpr (!Cfg::JAC_TMA_Q)
                        qv[u * P::R1 + r] = (WFM_JAC_Q_AHEAD && kk > 0) ? qn[u * P::R1 + r] : __ldcs(&a.q[base + (t + TT * u) + P::S1 * r]);
            wfm_mbar_wait(&mbar[slot], row_phase);
            ++row_phase;
            const cx<T>* land = landing + (size_t)slot * N;
            if constexpr (Cfg::JAC_TMA_Q) {
#pragma unroll
                for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
                    for (int r = 0; r < P::R1; ++r) qv[u * P::R1 + r] = landq[(size_t)slot * N + (t + TT * u) + P::S1 * r];
            }
#pragma unroll
            for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
                for (int r = 0; r < P::R1; ++r) {
                    const cx<T> av = land[(t + TT * u) + P::S1 * r];
                    v[u * P::R1 + r] = mkc<T>(av.x * qv[u * P::R1 + r], av.y * qv[u * P::R1 + r]);
                }
        } else {
#pragma unroll
            for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
                for (int r = 0; r < P::R1; ++r) {
                    const int x = (t + TT * u) + P::S1 * r;
#if defined(WFM_PROBE_JAC_L1) || defined(WFM_PROBE_JAC_L2)
                    const cx<T> av = __ldg(&a.cpx[base + x]);
                    const T qv = __ldg(&a.q[base + x]);
#else
                    const cx<T> av = __ldcs(&a.cpx[base + x]);
                    const T qv = __ldcs(&a.q[base + x]);
#endif
                    v[u * P::R1 + r] = mkc<T>(av.x * qv, av.y * qv);
                }
        }
#endif
        if constexpr (Cfg::JAC_TMA) {
            const bool more = kk + 1 < Cfg::KR_JAC;
            RowBulkLoadHook hk{landing + (size_t)slot * N, &a.cpx[base + (more ? (size_t)N * C : 0)],
                               (unsigned)(N * sizeof(cx<T>)), &mbar[slot], more && t == 0,
                               landq + (size_t)slot * N, &a.q[base + (more ? (size_t)N * C : 0)],
                               Cfg::JAC_TMA_Q ? (unsigned)(N * sizeof(T)) : 0u};
            using QHook = RowQAheadHook<T, E, P::R1, TT, P::S1>;
            QHook qh{qn, &a.q[base + (more ? (size_t)N * C : 0)], t, WFM_JAC_Q_AHEAD && more && !Cfg::JAC_TMA_Q};
            fft_inplace<T, P, L, RowSync<TT>, RowBulkLoadHook, false, WFM_JAC_TW_TREE, QHook, PipeCfg<T, N>::TWTAB_JAC>(v, cells + slot * L::LEN, t, tw_s,
                                                                                                           tw_s + PipeCfg<T, N>::TW1, slot, hk, qh);
        } else {
            fft_inplace<T, P, L, RowSync<TT>, NoHook, false, WFM_JAC_TW_TREE, NoHook, PipeCfg<T, N>::TWTAB_JAC>(v, cells + slot * L::LEN, t, tw_s, tw_s + PipeCfg<T, N>::TW1, slot);
        }
        cx<T>* dst = a.T2 + (size_t)ringoff + (size_t)y * C;
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (leg_live<P::RL, NARROW>(e % P::RL) && xis[e] >= 0) __stcg(&dst[xis[e]], v[e]);
        if (kk + 1 < Cfg::KR_JAC) RowSync<TT>::sync(slot);
    }
}

// B-item: active columns xi0 .. xi0+C-1 of plane pl: FFT along y, then the masked trig products
// shared by the three Jacobians, written per plane (summed over z by k_jac_reduce in fixed order):
//   jin = rho*(B_re sin ph + B_im cos ph)   on maskPupil   (WFM:925-928, 1253)
//   J   = B_re cos ph - B_im sin ph         on the support (WFM:607-611)
template <typename T, int N, bool NARROW>
WFM_DEVI void jac_cols_item(const JacArgs<T>& a, int pl, int sub, int bm, int ringoff, cx<T>* cells, const cx<T>* tw_s,
                            const double2* cis_s, PipeQueue& qu, const PipeCtl& ctl) {
    using P = Plan<N>;
    constexpr int C = PipeCfg<T, N>::C, TT = P::T, E = P::E;
    using L = typename PipeCfg<T, N>::ColL;
    const int c = threadIdx.x % C, t = threadIdx.x / C;
    const int xi = sub * C + c;
    const bool colvalid = xi < a.nax;
    const size_t tbase = (size_t)sub * N * C + c;      // this thread's column inside the tile-major strip
    // bm: model of a batch handle (0 otherwise)
    const size_t sbase = tbase + (size_t)bm * N * a.pitch;   // the same column in that model's pupil strip
    const cx<T>* src = a.T2 + (size_t)ringoff + tbase;
    cx<T> v[E];
#pragma unroll
    for (int u = 0; u < E / P::R1; ++u)
#pragma unroll
        for (int r = 0; r < P::R1; ++r) {
            const int y = (t + TT * u) + P::S1 * r;
            v[u * P::R1 + r] = colvalid ? __ldcg(&src[(size_t)y * C]) : mkc<T>((T)0, (T)0);
        }
    // the flags of this thread's output cells: independent loads, in flight during the transform
    unsigned fl = 0;
#pragma unroll
    for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
        for (int r = 0; r < P::RL; ++r)
            if (leg_live<P::RL, NARROW>(r))
                fl |= (unsigned)__ldg(&a.st.flags[sbase + (size_t)((t + TT * u) + P::SL * r) * C]) << (2 * (u * P::RL + r));
    using ClaimHook = PipePrefetchHook<JacClaimPrefetch<T, N>>;
    fft_inplace<T, P, L, CtaSync, ClaimHook, false, WFM_JAC_TW_TREE, NoHook, PipeCfg<T, N>::TWTAB_JAC>(
        v, cells + c, t, tw_s, tw_s + PipeCfg<T, N>::TW1, 0, ClaimHook{qu, ctl, a.g.nzl, JacClaimPrefetch<T, N>{a.cpx, a.q}});
#if WFM_RING_DISCARD == 1
    {   // the tile has been read by every thread (the transform's barriers are behind us): drop its lines from L2
        const char* tile = reinterpret_cast<const char*>(a.T2 + (size_t)ringoff + (size_t)sub * N * C);
        constexpr int LINES = (int)(sizeof(cx<T>) * (size_t)N * C / 128);
        for (int i = threadIdx.x; i < LINES; i += PipeCfg<T, N>::THREADS) wfm_discard_l2(tile + (size_t)i * 128);
    }
#endif
    const int iz = a.g.z0 + (pl - bm * a.g.nzm);
    const double s = defoc_scale_dev(iz, a.g.nz_global, a.g.dz);
    const bool mod_plane = (a.Gm != nullptr) && (!a.last_plane_only || iz == a.g.nz_global - 1);
    const unsigned want = mod_plane ? 3u : 1u;         // mask bit, plus the support bit when J is needed
    const size_t obase = (size_t)pl * N * a.pitch + tbase;
#pragma unroll
    for (int u = 0; u < E / P::RL; ++u)
#pragma unroll
        for (int r = 0; r < P::RL; ++r) {
            const int e = u * P::RL + r;
            if (!leg_live<P::RL, NARROW>(r)) continue;
            const unsigned f = (fl >> (2 * e)) & want;
            if (!f) continue;
            const size_t row = (size_t)((t + TT * u) + P::SL * r) * C;
            const size_t cell = sbase + row;
            const double ph = __dadd_rn(__ldg(&a.st.phi[cell]), __dmul_rn(s, __ldg(&a.st.psi[cell])));
            const double rho = __ldg(&a.st.rho[cell]);
            double sn, cs;
            WFM_SINCOS(ph, &sn, &cs);
            const double br = (double)v[e].x, bi = (double)v[e].y;
            if (f & 1u) a.Gj[obase + row] = rho * (br * sn + bi * cs);
            if (mod_plane) a.Gm[obase + row] = br * cs - bi * sn;
        }
}

template <typename T, int N, bool NARROW>
__global__ void __launch_bounds__(PipeCfg<T, N>::THREADS, PipeCfg<T, N>::template minb_jac<NARROW>()) k_jac_pipeline(JacArgs<T> a, PipeCtl ctl) {
    using Cfg = PipeCfg<T, N>;
    WFM_DYN_SMEM(cx<T>, cells);
    cx<T>* tw_s = cells + Cfg::CELLS;
    cx<T>* tw2_s = tw_s + Cfg::TW1;
    int* invx_s = reinterpret_cast<int*>(tw2_s + Cfg::TW2);
    double2* cis_s = reinterpret_cast<double2*>(invx_s + N);
    for (int i = threadIdx.x; i < N; i += Cfg::THREADS) invx_s[i] = a.inv_x[i];
    pipe_fill_twiddles<T, N>(tw_s, tw2_s, a.tw);
    if (threadIdx.x < WFM_CIS_ENTRIES) cis_s[threadIdx.x] = a.cis[threadIdx.x];
    // landing buffers of the row items (bulk-async copies of conj(a) rows) and their barriers
    cx<T>* landing = reinterpret_cast<cx<T>*>(cis_s + WFM_CIS_ENTRIES);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(landing) + (Cfg::JAC_TMA ? Cfg::LANDING - 64 : 0));
    unsigned row_phase = 0;                      // rows this thread's group has received so far (= barrier phase)
    if constexpr (Cfg::JAC_TMA) {
        if (threadIdx.x < Cfg::C) wfm_mbar_init(&mbar[threadIdx.x], 1);
        wfm_mbar_init_fence();
    }
    wfm_grid_dep_trigger();  // the next kernel's CTAs may take the place of ours as we exit
    wfm_grid_dep_wait();     // the tables above are constants; everything below depends on the previous kernel
    __shared__ int s_queue[2 * PipeQueue::SLOT];
    PipeQueue qu;
    const int P = a.g.nzl;
    qu.init(s_queue, ctl, P);
    for (;;) {
        bool ready;
        PipeItem it = qu.take(ctl, P, ready);
        if (it.type < 0) break;
#ifdef WFM_RING_MOD_IN_ITEM   /* A/B knob: the slot offset computed by every thread, as before */
        it.ringoff = (it.plane % ctl.ring) * ctl.ring_stride;
#endif
        if (it.type == 0) {
            if (ctl.roles & 1) {
                PipeDep dep;
                dep.cnt = ready ? nullptr : &ctl.cntB[it.plane - ctl.ring]; dep.target = (unsigned)ctl.nB; dep.err = ctl.err;
                jac_rows_item<T, N, NARROW>(a, it.plane, it.sub, it.ringoff, cells, tw_s, invx_s, dep, qu, ctl, landing, mbar, row_phase);
            }
            if (qu.prefetch(ctl, P)) JacClaimPrefetch<T, N>{a.cpx, a.q}(qu.peek_next());
            pipe_signal(&ctl.cntA[it.plane]);
        } else {
            if (ctl.roles & 2) {
                if (!ready) pipe_wait(&ctl.cntA[it.plane], ctl.nA, ctl.err);
                jac_cols_item<T, N, NARROW>(a, it.plane, it.sub, it.model, it.ringoff, cells, tw_s, cis_s, qu, ctl);
            }
            if (qu.prefetch(ctl, P)) JacClaimPrefetch<T, N>{a.cpx, a.q}(qu.peek_next());
#if WFM_RING_DISCARD == 2
            {   // the column tile has been read by every thread: drop its lines from L2
                const char* tile = reinterpret_cast<const char*>(a.T2 + (size_t)it.ringoff + (size_t)it.sub * N * Cfg::C);
                constexpr int LINES = (int)(sizeof(cx<T>) * (size_t)N * Cfg::C / 128);
                pipe_signal_drop(&ctl.cntB[it.plane], [&](int lane) {
                    if (!(ctl.roles & 2)) return;
                    for (int i = lane; i < LINES; i += 32) wfm_discard_l2(tile + (size_t)i * 128);
                });
            }
#else
            pipe_signal(&ctl.cntB[it.plane]);
#endif
        }
        qu.advance();
    }
    pipe_finish(ctl, P);
}

// ---- contraction of the per-plane integrands with the basis: warp-then-block reductions --------
// The basis is gathered once (per basis / support change) into the order of the support-cell list,
// so that the reduction reads it as contiguous rows: Zs[k][li] = Z[k][in_list[li]].
__global__ void k_pack_basis(double* __restrict__ Zs, const double* __restrict__ Z, const int* __restrict__ in_list,
                             int ncells, int nzern, int npix) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= ncells) return;
    const int in = in_list[li];
    for (int k = 0; k < nzern; ++k) Zs[(size_t)k * ncells + li] = Z[(size_t)k * npix + in];
}

struct ReduceArgs {
    Geom g;
    const double* Gj; const double* Gm; int pitch;
    const double* Zs;        // [nzern][ncells] basis in cell-list order
    const double* psi;
    const uint8_t* flags;    // strip flags (bit 0 maskPupil, bit 1 support)
    const int* cell_list;    // [ncells] tile-major strip cells that lie on the support
    const int* in_list;      // [ncells] their pixel index kx + N*ky
    int ncells;
    int nphase, nmod, phase_off;
    unsigned kinds;
    int last_plane_only;
    double dxy, lambda_ni, deltaX, deltaY;
    const double* bpar;      // batch handles: [nbatch][4] = {ni/lambda, deltaX, deltaY, 1/|beta|} per model, else NULL
    int cpm;                 // plane chunks per model: grid.y = nbatch * cpm
    int batches;             // batches of WFM_RED_PLANES planes per chunk (1 .. WFM_RED_BATCHES), chosen per launch
    double* block_part;      // [nbatch][cpm][nblocks][glen]
    int glen;                // 3 + nphase + nmod
};

#define WFM_RED_THREADS 256
#define WFM_RED_CHUNK 8
#ifndef WFM_RED_PLANES
#define WFM_RED_PLANES 16   // loads a thread keeps in flight at once; 4 and 8 measured slower
#endif
#ifndef WFM_RED_BATCHES
#define WFM_RED_BATCHES 4   // most batches of WFM_RED_PLANES a thread sums one after the other: a CTA owns a chunk of up to
#endif                      // 64 planes, so that 512^2 x 256 is ONE wave of 360 CTAs (was 2.4 waves of 1440); small stacks
                            // take fewer batches per chunk so that the grid still fills the GPU (launch_jac_reduce)
#define WFM_RED_CHUNK_PLANES (WFM_RED_PLANES * WFM_RED_BATCHES)

// One thread per support cell and per chunk of WFM_RED_PLANES planes.  Sums the planes of the chunk
// in fixed order (gP = sum jin, gD = sum defoc*jin, gM = sum J), then forms the glen dot products
// WFM_RED_CHUNK at a time: shuffle-reduce inside each warp, then across the warps of the block
// through shared memory.
__global__ void __launch_bounds__(WFM_RED_THREADS) k_jac_reduce(ReduceArgs a) {
    __shared__ double red[WFM_RED_THREADS / 32][WFM_RED_CHUNK];
    const int N = a.g.N;
    const size_t img = (size_t)N * a.pitch;
    const int li = blockIdx.x * WFM_RED_THREADS + threadIdx.x;
    const bool sup = li < a.ncells;                    // every listed cell lies on the support
    const size_t cell = sup ? (size_t)a.cell_list[li] : 0;      // (index lists: constant since the last basis change)
    const int in = sup ? a.in_list[li] : 0;
    wfm_grid_dep_trigger();
    wfm_grid_dep_wait();                               // Gj / Gm come from the pipeline kernel before us
    const int bm = blockIdx.y / a.cpm;                 // model of a batch handle (0 otherwise); chunks never straddle models
    const int chunk_planes = WFM_RED_PLANES * a.batches;
    const int zc0 = (blockIdx.y - bm * a.cpm) * chunk_planes;          // first plane of the chunk inside its model
    const int pend = (zc0 + chunk_planes < a.g.nzm) ? bm * a.g.nzm + zc0 + chunk_planes : (bm + 1) * a.g.nzm;
    const bool m = sup && (a.flags[(size_t)bm * img + cell] & 1u);
    double gP = 0.0, gD = 0.0, gM = 0.0;
#pragma unroll 1
    for (int bt = 0; bt < a.batches; ++bt) {
        const int zl0 = zc0 + bt * WFM_RED_PLANES;
        const int p0 = bm * a.g.nzm + zl0;
        if (p0 >= pend) break;
        const int p1 = (p0 + WFM_RED_PLANES < pend) ? p0 + WFM_RED_PLANES : pend;
        if (m) {
            // all loads of the batch in flight at once (the kernel is latency-bound), then the fixed-order sums
            double jin[WFM_RED_PLANES];
#pragma unroll
            for (int k = 0; k < WFM_RED_PLANES; ++k) jin[k] = (p0 + k < p1) ? __ldcs(&a.Gj[(size_t)(p0 + k) * img + cell]) : 0.0;
#pragma unroll
            for (int k = 0; k < WFM_RED_PLANES; ++k) {
                gP += jin[k];
                gD += defoc_depth_dev(a.g.z0 + zl0 + k, a.g.nz_global, a.g.dz) * jin[k];
            }
        }
        if (sup && (a.kinds & 4u)) {
            double jm[WFM_RED_PLANES];
#pragma unroll
            for (int k = 0; k < WFM_RED_PLANES; ++k)
                jm[k] = (p0 + k < p1 && (!a.last_plane_only || a.g.z0 + zl0 + k == a.g.nz_global - 1))
                            ? __ldcs(&a.Gm[(size_t)(p0 + k) * img + cell]) : 0.0;
#pragma unroll
            for (int k = 0; k < WFM_RED_PLANES; ++k) gM += jm[k];
        }
    }
    // defocus weights: idef = 1/psi on maskPupil (WFM:1251); rx, ry of the prologue WFM:1040-1061
    double wD = 0.0, rx = 0.0, ry = 0.0;
    const double lambda_ni = a.bpar ? a.bpar[4 * bm] : a.lambda_ni;
    if (m && (a.kinds & 1u)) {
        const double scale = 1.0 / ((double)N * a.dxy);
        wD = gD * (1.0 / a.psi[(size_t)bm * N * N + in]);
        rx = (double)kappa_dev(in % N, N) * scale - (a.bpar ? a.bpar[4 * bm + 1] : a.deltaX);
        ry = (double)kappa_dev(in / N, N) * scale - (a.bpar ? a.bpar[4 * bm + 2] : a.deltaY);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* out = a.block_part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * a.glen;
    for (int j0 = 0; j0 < a.glen; j0 += WFM_RED_CHUNK) {
        double acc[WFM_RED_CHUNK];
#pragma unroll
        for (int jj = 0; jj < WFM_RED_CHUNK; ++jj) {
            const int j = j0 + jj;
            double val = 0.0;
            if (j < a.glen && sup) {
                if (j < 3) {
                    if (a.kinds & 1u) val = (j == 0) ? wD * lambda_ni : (j == 1 ? wD * rx : wD * ry);
                } else if (j < 3 + a.nphase) {
                    if ((a.kinds & 2u) && m) val = gP * a.Zs[(size_t)(j - 3 + a.phase_off) * a.ncells + li];
                } else {
                    if (a.kinds & 4u) val = gM * a.Zs[(size_t)(j - 3 - a.nphase) * a.ncells + li];
                }
            }
            acc[jj] = val;
        }
#pragma unroll
        for (int jj = 0; jj < WFM_RED_CHUNK; ++jj) {
            double x = acc[jj];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            acc[jj] = x;
        }
        if (lane == 0) {
#pragma unroll
            for (int jj = 0; jj < WFM_RED_CHUNK; ++jj) red[warp][jj] = acc[jj];
        }
        __syncthreads();
        if (threadIdx.x < WFM_RED_CHUNK && j0 + (int)threadIdx.x < a.glen) {
            double x = 0.0;
#pragma unroll
            for (int w = 0; w < WFM_RED_THREADS / 32; ++w) x += red[w][threadIdx.x];
            out[j0 + threadIdx.x] = x;
        }
        __syncthreads();
    }
}

// Final fixed-order sum over blocks and the scale factors of each Jacobian:
//   defocus: t = -2pi*PSFnorm*jin (WFM:1253), d = sum t*{lambda_ni, rx, ry}*defoc/psi (1258-1260,1278-1280)
//   phase  : g[k] = -2*PSFnorm*sum jin*Z (WFM:937)
//   modulus: 2*PSFnorm*sum J*Z_k * (1-(beta_k*NBeta)^2)*NBeta (WFM:674)
// ---- gradient exchange over peer memory (NVLink / NVSwitch), fused into k_jac_final ------------------------
// One process per GPU (torchrun): every rank owns a landing buffer, mapped into every peer through CUDA IPC
// (wfm_exchange_export / _connect).  A rank stores its partial K-vector into slot [rank] of EVERY peer's buffer
// straight from k_jac_final, fences at system scope and raises its flag on every peer; the last CTA of the kernel
// then waits for the world's flags on its own buffer and adds the slots in rank order -- every rank gets the same,
// deterministic sum.  This replaces the one NCCL all-reduce of the step (~15-27 us of launch + protocol latency
// for 112 bytes) by one NVLink store round (~2-4 us).  Buffers are double-buffered by call parity: a rank can be at
// most one call ahead of a peer (it needs that peer's flag of call e before it can leave call e).
#define WFM_MAX_RANKS 16
struct XchgArgs {
    double* slots[WFM_MAX_RANKS];     // landing buffer of every rank as mapped here: [2][world][glen]
    unsigned* flags[WFM_MAX_RANKS];   // flag words of every rank as mapped here:     [2][world]
    unsigned* ticket;                 // [1] local: CTAs of k_jac_final that have pushed their component
    unsigned* err;                    // [1] local: set when a peer's flag does not arrive
    int rank, world;                  // world == 0: no exchange (plain handle)
    unsigned epoch;                   // number of this call, from 1
};
WFM_DEVI void wfm_fence_system() {
#ifdef WFM_EMU
    __threadfence();
#else
    __threadfence_system();
#endif
}

#define WFM_FINAL_THREADS 128
// One CTA per gradient component: strided partial sums, warp shuffle, then across warps.
// Batch handles: blockIdx.y = model, nblocks partials per model, beta / 1/|beta| from the device tables.
__global__ void __launch_bounds__(WFM_FINAL_THREADS) k_jac_final(const double* __restrict__ block_part, int nblocks,
                                                                int glen, int nphase, double psf_norm, Coefs beta,
                                                                double nbeta, unsigned kinds,
                                                                const double* __restrict__ beta_tab, int nmod,
                                                                const double* __restrict__ bpar,
                                                                double* __restrict__ grad, XchgArgs xc) {
    __shared__ double red[WFM_FINAL_THREADS / 32];
    const int j = blockIdx.x;
    double x = 0.0;
    wfm_grid_dep_wait();
    block_part += (size_t)blockIdx.y * nblocks * glen;
    grad += (size_t)blockIdx.y * glen;
    for (int b = threadIdx.x; b < nblocks; b += WFM_FINAL_THREADS) x += block_part[(size_t)b * glen + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    double out = 0.0;
    if (threadIdx.x == 0) {
        x = 0.0;
        for (int w = 0; w < WFM_FINAL_THREADS / 32; ++w) x += red[w];
        if (j < 3) {
            if (kinds & 1u) out = -6.283185307179586 * psf_norm * x;
        } else if (j < 3 + nphase) {
            if (kinds & 2u) out = -2.0 * psf_norm * x;
        } else {
            if (kinds & 4u) {
                if (beta_tab) nbeta = bpar[4 * blockIdx.y + 3];
                const double bk = (beta_tab ? beta_tab[(size_t)blockIdx.y * nmod + (j - 3 - nphase)] : beta.v[j - 3 - nphase]) * nbeta;
                out = 2.0 * psf_norm * x * (1.0 - bk * bk) * nbeta;
            }
        }
    }
    if (xc.world <= 0) {
        if (threadIdx.x == 0) grad[j] = out;
        return;
    }
    // ---- cross-rank sum over peer memory (single-model handles: gridDim.y == 1) ----
    __shared__ int s_last;
    const unsigned half = xc.epoch & 1u;
    if (threadIdx.x == 0) {
        const size_t slot = ((size_t)half * xc.world + xc.rank) * glen + j;
        for (int r = 0; r < xc.world; ++r) *(volatile double*)&xc.slots[r][slot] = out;   // NVLink stores (own slot included)
        wfm_fence_system();
        s_last = (atomicAdd(xc.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
        if (s_last) { *(volatile unsigned*)xc.ticket = 0u; wfm_fence_system(); }
    }
    __syncthreads();
    if (!s_last) return;
    // last CTA of the grid: every component of this rank has been pushed and fenced at system scope
    if ((int)threadIdx.x < xc.world) {
        const int r = threadIdx.x;
        *(volatile unsigned*)&xc.flags[r][half * xc.world + xc.rank] = xc.epoch;           // raise my flag on rank r
        const volatile unsigned* mine = xc.flags[xc.rank] + half * xc.world;
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while ((int)(mine[r] - xc.epoch) < 0) {                                            // ... and wait for rank r's flag here
            if ((++spins & 4095u) == 0u) {                                                 // wall-clock limit: an error, never a hang
                const unsigned long long now = wfm_now_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 3 * WFM_WAIT_LIMIT_NS) { *(volatile unsigned*)xc.err = 1u; break; }
            }
            WFM_SPIN_PAUSE();
        }
        wfm_fence_system();
    }
    __syncthreads();
    const volatile double* land = xc.slots[xc.rank] + (size_t)half * xc.world * glen;
    for (int k = threadIdx.x; k < glen; k += WFM_FINAL_THREADS) {
        double sum = 0.0;
        for (int r = 0; r < xc.world; ++r) sum += land[(size_t)r * glen + k];              // fixed rank order
        grad[k] = sum;
    }
}

// Multi-device handles (one process drives all GPUs): the partial K-vectors of the devices have landed in slots[dev][glen]
// on the first device (stored there over NVLink by each device's k_jac_final); add them in device order.
__global__ void k_sum_slots(const double* __restrict__ slots, int nparts, int glen, double* __restrict__ grad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= glen) return;
    double sum = 0.0;
    for (int p = 0; p < nparts; ++p) sum += slots[(size_t)p * glen + j];
    grad[j] = sum;
}

}  // namespace wfm
