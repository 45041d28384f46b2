// wfm_conv_api.inl -- C ABI of the device-side FFT-convolution data term (included by wfm_api.cu).
// See wfm_conv.cuh for the semantics and the pass structure.
#include "wfm_conv.cuh"

struct wfm_conv {
    int nx = 0, ny = 0, nz = 0;
    int precision = WFM_F64;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    DevBuf V, X, y, w, R, hdev, gdev, cost_part, cost_dev, twx, twz, twh, twn;
    bool lite = false;                  // full complex work volume only (wfm_get_mtf)
    bool have_obj = false, have_data = false, have_w = false;
    std::string err;
    // ---- z-sharded data term (wfm_conv_create_multi) ----------------------------------------------------------
    // A PARENT (parts non-empty) owns one child per device; a CHILD holds the z-slab [z0, z0 + nz) of the real-space
    // volumes (its nz is the slab's) plus the pencil volume Vp / Xp = all nz_all planes of its rows [y0, y0 + nyl).
    std::vector<wfm_conv*> parts;
    wfm_conv* parent = nullptr;
    int part = 0, nz_all = 0, z0 = 0, y0 = 0, nyl = 0;
    DevBuf Vp, Xp;
    cudaEvent_t ev_pass = nullptr;      // "my scatter pass has been queued" (cross-device barrier, one per child)
    bool multi() const { return !parts.empty(); }
    size_t pvox() const { return (size_t)pitch() * nyl * nz_all; }   // entries of the pencil volume
    size_t vox() const { return (size_t)nx * ny * nz; }
    int pitch() const { return conv_pitch(nx); }
    size_t hvox() const { return (size_t)pitch() * ny * nz; }      // entries of the half-spectrum volume
    size_t esz() const { return precision == WFM_F64 ? 8 : 4; }
    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
};

namespace {

// dst[m] = exp(-2 pi i m / period), m < n
int conv_upload_twiddles(wfm_conv* c, DevBuf& dst, int n, int period = 0) {
    if (period == 0) period = n;
    std::vector<double2> t(n);
    for (int m = 0; m < n; ++m) {
        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)period;
        t[m].x = (double)cosl(a); t[m].y = (double)(-sinl(a));
    }
    WFM_CK(c, dst.ensure(sizeof(double2) * n));
    WFM_CK(c, cudaMemcpy(dst.p, t.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
    return WFM_OK;
}

template <typename T, int N, int LOAD, int STORE> int conv_rows(wfm_conv* c, ConvArgs<T> a) {
    auto kfn = &k_conv_rows<T, N, LOAD, STORE>;
    const size_t smem = sizeof(cx<T>) * ((size_t)RowCfg<N>::RB * RowLayout<T, N>::LEN + N + 16);
    if (smem > 48 * 1024) WFM_CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t nrows = (size_t)a.ny * a.nz;
    const unsigned grid = (unsigned)((nrows + RowCfg<N>::RB - 1) / RowCfg<N>::RB);
    a.tw = (const cx<T>*)c->twx.p;
    WFM_LAUNCH(kfn, dim3(grid), dim3(RowCfg<N>::THREADS), smem, c->stream, a);
    WFM_CK_LAUNCH(c, "k_conv_rows");
    return WFM_OK;
}

template <typename T, int N, int STORE> int conv_rows_r2c(wfm_conv* c, ConvArgs<T> a) {
    constexpr int M = N / 2;
    auto kfn = &k_conv_rows_r2c<T, N, STORE>;
    const size_t smem = ConvRowSmem<M>::template bytes<T>();
    if (smem > 48 * 1024) WFM_CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t nrows = (size_t)a.ny * a.nz;
    const unsigned grid = (unsigned)((nrows + RowCfg<M>::RB - 1) / RowCfg<M>::RB);
    a.tw = (const cx<T>*)c->twh.p; a.twn = (const cx<T>*)c->twn.p;
    WFM_LAUNCH(kfn, dim3(grid), dim3(RowCfg<M>::THREADS), smem, c->stream, a);
    WFM_CK_LAUNCH(c, "k_conv_rows_r2c");
    return WFM_OK;
}
template <typename T, int N, int STORE> int conv_rows_c2r(wfm_conv* c, ConvArgs<T> a, int* nparts) {
    constexpr int M = N / 2;
    auto kfn = &k_conv_rows_c2r<T, N, STORE>;
    const size_t smem = ConvRowSmem<M>::template bytes<T>();
    if (smem > 48 * 1024) WFM_CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t nrows = (size_t)a.ny * a.nz;
    const unsigned grid = (unsigned)((nrows + RowCfg<M>::RB - 1) / RowCfg<M>::RB);
    if (nparts) *nparts = (int)grid;
    a.tw = (const cx<T>*)c->twh.p; a.twn = (const cx<T>*)c->twn.p;
    WFM_LAUNCH(kfn, dim3(grid), dim3(RowCfg<M>::THREADS), smem, c->stream, a);
    WFM_CK_LAUNCH(c, "k_conv_rows_c2r");
    return WFM_OK;
}

// axis 1: columns inside each plane (length ny == nx == N); axis 2: columns along z (length nz).  `pitch` = entries
// per row of the volume (nx for the full complex volume, conv_pitch(nx) for the half spectrum).
template <typename T, int LEN, int STORE> int conv_cols(wfm_conv* c, ConvArgs<T> a, int axis, int pitch) {
    using Cfg = ConvColCfg<T, LEN>;
    auto kfn = &k_conv_cols<T, LEN, STORE>;
    if (Cfg::SMEM > 48 * 1024) WFM_CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const size_t npix = (size_t)pitch * a.ny;
    size_t stride, outer_stride; int tiles_per_outer; size_t nouter;
    if (axis == 1) { stride = pitch; tiles_per_outer = pitch / Cfg::CW; outer_stride = npix; nouter = a.nz; a.tw = (const cx<T>*)c->twx.p; }
    else { stride = npix; tiles_per_outer = (int)(npix / Cfg::CW); outer_stride = 0; nouter = 1; a.tw = (const cx<T>*)c->twz.p; }
    const unsigned grid = (unsigned)(nouter * tiles_per_outer);
    WFM_LAUNCH(kfn, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, c->stream, a, stride, tiles_per_outer, outer_stride);
    WFM_CK_LAUNCH(c, "k_conv_cols");
    return WFM_OK;
}

// y pass of a z-sharded volume: the outputs land in the pencil volumes of their owners (ConvArgs::peers)
template <typename T, int LEN> int conv_cols_scatter(wfm_conv* c, ConvArgs<T> a, int pitch) {
    using Cfg = ConvColCfg<T, LEN>;
    auto kfn = &k_conv_cols<T, LEN, CS_CPLX, true>;
    if (Cfg::SMEM > 48 * 1024) WFM_CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const size_t npix = (size_t)pitch * a.ny;
    a.tw = (const cx<T>*)c->twx.p;
    const unsigned grid = (unsigned)((size_t)a.nz * (pitch / Cfg::CW));
    WFM_LAUNCH(kfn, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, c->stream, a, (size_t)pitch, pitch / Cfg::CW, npix);
    WFM_CK_LAUNCH(c, "k_conv_cols (scatter)");
    return WFM_OK;
}
// fused z pass on this device's pencil volume; the outputs land in the slab volumes of their owners
template <typename T, int LEN, int MUL> int conv_cols_zz_scatter(wfm_conv* c, ConvArgs<T> a, int pitch) {
    using Cfg = ConvColCfg<T, LEN>;
    auto kfn = &k_conv_cols_zz<T, LEN, MUL, true>;
    if (Cfg::SMEM > 48 * 1024) WFM_CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const size_t npix = (size_t)pitch * a.ny;                 // a.ny = rows of the pencil volume (nyl)
    a.tw = (const cx<T>*)c->twz.p;
    WFM_LAUNCH(kfn, dim3((unsigned)(npix / Cfg::CW)), dim3(Cfg::THREADS), Cfg::SMEM, c->stream, a, npix, (int)(npix / Cfg::CW), (size_t)0);
    WFM_CK_LAUNCH(c, "k_conv_cols_zz (scatter)");
    return WFM_OK;
}

#define WFM_CONV_SWITCH(len, CALL)                                              \
    switch (len) {                                                              \
        case 32: { constexpr int L_ = 32; return CALL; }                        \
        case 64: { constexpr int L_ = 64; return CALL; }                        \
        case 128: { constexpr int L_ = 128; return CALL; }                      \
        case 256: { constexpr int L_ = 256; return CALL; }                      \
        case 512: { constexpr int L_ = 512; return CALL; }                      \
        case 1024: { constexpr int L_ = 1024; return CALL; }                    \
        case 2048: { constexpr int L_ = 2048; return CALL; }                    \
        default: return c->fail(WFM_ERR_UNSUPPORTED, "unsupported FFT length %d", (int)(len)); \
    }

template <int LOAD, int STORE> int conv_rows_n(wfm_conv* c, const ConvArgs<double>& a) {
    WFM_CONV_SWITCH(c->nx, (conv_rows<double, L_, LOAD, STORE>(c, a)))
}
// z pass there and back in one kernel (axis 2 only)
template <typename T, int LEN, int MUL> int conv_cols_zz(wfm_conv* c, ConvArgs<T> a, int pitch) {
    using Cfg = ConvColCfg<T, LEN>;
    auto kfn = &k_conv_cols_zz<T, LEN, MUL>;
    if (Cfg::SMEM > 48 * 1024) WFM_CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const size_t npix = (size_t)pitch * a.ny;
    a.tw = (const cx<T>*)c->twz.p;
    WFM_LAUNCH(kfn, dim3((unsigned)(npix / Cfg::CW)), dim3(Cfg::THREADS), Cfg::SMEM, c->stream, a, npix, (int)(npix / Cfg::CW), (size_t)0);
    WFM_CK_LAUNCH(c, "k_conv_cols_zz");
    return WFM_OK;
}
template <int MUL> int conv_cols_zz_n(wfm_conv* c, const ConvArgs<double>& a, int pitch) {
    WFM_CONV_SWITCH(c->nz, (conv_cols_zz<double, L_, MUL>(c, a, pitch)))
}
int conv_cols_scatter_n(wfm_conv* c, const ConvArgs<double>& a, int pitch) {
    WFM_CONV_SWITCH(c->ny, (conv_cols_scatter<double, L_>(c, a, pitch)))
}
template <int MUL> int conv_cols_zz_scatter_n(wfm_conv* c, const ConvArgs<double>& a, int pitch) {
    WFM_CONV_SWITCH(c->nz_all, (conv_cols_zz_scatter<double, L_, MUL>(c, a, pitch)))
}
template <int STORE> int conv_cols_n(wfm_conv* c, const ConvArgs<double>& a, int axis, int pitch) {
    WFM_CONV_SWITCH(axis == 1 ? c->ny : c->nz, (conv_cols<double, L_, STORE>(c, a, axis, pitch)))
}
template <int STORE> int conv_cols_zall_n(wfm_conv* c, const ConvArgs<double>& a, int pitch) {      // z pass of length nz_all
    WFM_CONV_SWITCH(c->nz_all, (conv_cols<double, L_, STORE>(c, a, 2, pitch)))
}
template <int STORE> int conv_r2c_n(wfm_conv* c, const ConvArgs<double>& a) {
    WFM_CONV_SWITCH(c->nx, (conv_rows_r2c<double, L_, STORE>(c, a)))
}
template <int STORE> int conv_c2r_n(wfm_conv* c, const ConvArgs<double>& a, int* nparts) {
    WFM_CONV_SWITCH(c->nx, (conv_rows_c2r<double, L_, STORE>(c, a, nparts)))
}

ConvArgs<double> conv_args(wfm_conv* c) {
    ConvArgs<double> a;
    memset(&a, 0, sizeof(a));
    a.V = (double2*)c->V.p; a.X = (const double2*)c->X.p; a.Xout = (double2*)c->X.p;
    a.y = (const double*)c->y.p; a.w = c->have_w ? (const double*)c->w.p : nullptr;
    a.nx = c->nx; a.ny = c->ny; a.nz = c->nz;
    a.inv_ntot = 1.0 / ((double)c->nx * (double)c->ny * (double)c->nz);
    a.alpha = 1.0; a.clear_grad = 1;
    return a;
}

}  // namespace

namespace wfm_multi {
int conv_destroy(wfm_conv* p);
int conv_scatter_host(wfm_conv* p, const void* host, DevBuf wfm_conv::*buf);
int conv_set_object(wfm_conv* p, const void* obj_host);
int conv_cost_and_gradient_host(wfm_conv* p, double alpha, const void* h_host, void* grad_host, int clr, double* cost);
int eval_fg(wfm_model* h, wfm_conv* p, int param, int n, double alpha, double* cost, double* grad_out, unsigned kinds);
}  // namespace wfm_multi

extern "C" {

static int conv_create_impl(wfm_conv** out, int nx, int ny, int nz, int precision, int device, bool lite, int slab_of = 0);

int wfm_conv_create(wfm_conv** out, int nx, int ny, int nz, int precision, int device) {
    return conv_create_impl(out, nx, ny, nz, precision, device, false);
}

// lite: only the work volume and the twiddles (3-D transform helper of wfm_get_mtf)
// slab_of > 0: a child of a z-sharded data term -- nz is the slab's plane count, slab_of the length of the z transform
static int conv_create_impl(wfm_conv** out, int nx, int ny, int nz, int precision, int device, bool lite, int slab_of) {
    if (!out) { g_create_error = "out is NULL"; return WFM_ERR_INVALID_ARG; }
    *out = nullptr;
    if (nx != ny) { g_create_error = "Nx should equal Ny"; return WFM_ERR_INVALID_ARG; }
    const int nz_fft = slab_of > 0 ? slab_of : nz;
    if (!supported_n(nx) || !supported_n(nz_fft) || nz < 1) { g_create_error = "Nx and Nz must be powers of two in [32, 2048]"; return WFM_ERR_UNSUPPORTED; }
    if (precision != WFM_F64) { g_create_error = "the convolution data term is fp64 only in this revision"; return WFM_ERR_UNSUPPORTED; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        g_create_error = "no CUDA device available (this library has no CPU fallback)"; return WFM_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "bad device index"; return WFM_ERR_INVALID_ARG; }
    wfm_conv* c = new (std::nothrow) wfm_conv();
    if (!c) { g_create_error = "out of host memory"; return WFM_ERR_NOMEM; }
    c->nx = nx; c->ny = ny; c->nz = nz; c->nz_all = nz_fft; c->precision = precision; c->device = device; c->lite = lite;
    auto bail = [&](int code, const char* what) { g_create_error = what; wfm_conv_destroy(c); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(WFM_ERR_CUDA, "cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(WFM_ERR_CUDA, "cudaStreamCreate failed");
    c->stream = c->own_stream;
    const size_t vox = c->vox();
    if (c->V.ensure(16 * (lite ? vox : c->hvox())) || c->cost_dev.ensure(8) || c->cost_part.ensure(8 * ((size_t)ny * nz + 8)) ||
        (!lite && (c->X.ensure(16 * c->hvox()) || c->y.ensure(8 * vox) || c->R.ensure(8 * vox))))
        return bail(WFM_ERR_NOMEM, "device allocation failed");
    if (conv_upload_twiddles(c, c->twx, nx) != WFM_OK || conv_upload_twiddles(c, c->twz, nz_fft) != WFM_OK ||
        conv_upload_twiddles(c, c->twh, nx / 2) != WFM_OK || conv_upload_twiddles(c, c->twn, nx / 2, nx) != WFM_OK)
        return bail(WFM_ERR_CUDA, "twiddle upload failed");
    *out = c;
    return WFM_OK;
}

int wfm_conv_destroy(wfm_conv* c) {
    if (!c) return WFM_OK;
    if (c->multi()) return wfm_multi::conv_destroy(c);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->V, &c->X, &c->y, &c->w, &c->R, &c->hdev, &c->gdev, &c->cost_part, &c->cost_dev, &c->twx, &c->twz,
                      &c->twh, &c->twn})
        b->release();
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return WFM_OK;
}

const char* wfm_conv_last_error(const wfm_conv* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int wfm_conv_set_stream(wfm_conv* c, void* s) {
    if (!c) return WFM_ERR_INVALID_ARG;
    if (c->multi()) return c->fail(WFM_ERR_UNSUPPORTED, "a z-sharded data term runs every device on its own stream");
    WFM_ENTER(c);
    cudaStreamSynchronize(c->stream);
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return WFM_OK;
}

// fdata.setPSF(obj, off) with off = {0,0,0} (PSF_Estimation.java:145,148): X = FFT3(obj)
int wfm_conv_set_object(wfm_conv* c, const void* obj_host) {
    if (!c) return WFM_ERR_INVALID_ARG;
    if (!obj_host) return c->fail(WFM_ERR_INVALID_ARG, "object is NULL");
    if (c->multi()) return wfm_multi::conv_set_object(c, obj_host);
    WFM_ENTER(c);
    WFM_CK(c, c->hdev.ensure(8 * c->vox()));
    WFM_CK(c, cudaMemcpyAsync(c->hdev.p, obj_host, 8 * c->vox(), cudaMemcpyHostToDevice, c->stream));
    ConvArgs<double> a = conv_args(c);
    a.real_in = (const double*)c->hdev.p;
    int rc = conv_r2c_n<CS_CPLX>(c, a); if (rc) return rc;
    rc = conv_cols_n<CS_CPLX>(c, a, 1, c->pitch()); if (rc) return rc;
    rc = conv_cols_n<CS_SPECTRUM>(c, a, 2, c->pitch()); if (rc) return rc;
    WFM_CK(c, cudaStreamSynchronize(c->stream));
    c->have_obj = true;
    return WFM_OK;
}

int wfm_conv_set_data(wfm_conv* c, const void* y_host) {          // fdata.setData(data)  PSF_Estimation.java:149
    if (!c) return WFM_ERR_INVALID_ARG;
    if (!y_host) return c->fail(WFM_ERR_INVALID_ARG, "data is NULL");
    if (c->multi()) {
        int rc = wfm_multi::conv_scatter_host(c, y_host, &wfm_conv::y);
        if (!rc) { c->have_data = true; for (wfm_conv* k : c->parts) k->have_data = true; }
        return rc;
    }
    WFM_ENTER(c);
    WFM_CK(c, cudaMemcpyAsync(c->y.p, y_host, 8 * c->vox(), cudaMemcpyHostToDevice, c->stream));
    WFM_CK(c, cudaStreamSynchronize(c->stream));
    c->have_data = true;
    return WFM_OK;
}

int wfm_conv_set_weights(wfm_conv* c, const void* w_host) {       // fdata.setWeights(weights, true)  :150
    if (!c) return WFM_ERR_INVALID_ARG;
    if (c->multi()) {
        if (!w_host) { c->have_w = false; for (wfm_conv* k : c->parts) k->have_w = false; return WFM_OK; }
        int rc = wfm_multi::conv_scatter_host(c, w_host, &wfm_conv::w);
        if (!rc) { c->have_w = true; for (wfm_conv* k : c->parts) k->have_w = true; }
        return rc;
    }
    WFM_ENTER(c);
    if (!w_host) { c->have_w = false; return WFM_OK; }
    WFM_CK(c, c->w.ensure(8 * c->vox()));
    WFM_CK(c, cudaMemcpyAsync(c->w.p, w_host, 8 * c->vox(), cudaMemcpyHostToDevice, c->stream));
    WFM_CK(c, cudaStreamSynchronize(c->stream));
    c->have_w = true;
    return WFM_OK;
}

// computeCostAndGradient(alpha, x = psf, gx, clr)  PSF_Estimation.java:157,206 -- device-resident, asynchronous.
// cost_dev (device double, may be NULL -> the handle's own slot) receives alpha/2 * sum w r^2.
int wfm_conv_cost_and_gradient_dev(wfm_conv* c, double alpha, const void* h_dev, void* grad_dev, int clr, double* cost_dev) {
    if (!c) return WFM_ERR_INVALID_ARG;
    if (!h_dev || !grad_dev) return c->fail(WFM_ERR_INVALID_ARG, "h_dev / grad_dev is NULL");
    if (c->multi() || c->parent) return c->fail(WFM_ERR_UNSUPPORTED, "z-sharded data term: use the host-buffer call or wfm_eval_fg");
    if (!c->have_obj || !c->have_data) return c->fail(WFM_ERR_STATE, "object and data must be set first");
    WFM_ENTER(c);
    ConvArgs<double> a = conv_args(c);
    a.real_in = (const double*)h_dev; a.grad = (double*)grad_dev; a.alpha = alpha; a.clear_grad = clr ? 1 : 0;
    a.cost_part = (double*)c->cost_part.p;
    a.resid = (double*)c->R.p;
    const int P = c->pitch();
    int rc, nparts = 0;
    // H = FFT3(h) (half spectrum)
    if ((rc = conv_r2c_n<CS_CPLX>(c, a))) return rc;
    if ((rc = conv_cols_n<CS_CPLX>(c, a, 1, P))) return rc;
    // r = IFFT3(H X) - y; cost; R = w r      (z forward, product, z inverse fused: the volume stays on the SM)
    if ((rc = conv_cols_zz_n<CS_MULX_CONJ>(c, a, P))) return rc;
    if ((rc = conv_cols_n<CS_CPLX>(c, a, 1, P))) return rc;
    if ((rc = conv_c2r_n<CS_RESID>(c, a, &nparts))) return rc;
    {
        auto kfin = &k_conv_cost_final;
        WFM_LAUNCH(kfin, dim3(1), dim3(1024), 0, c->stream, (const double*)c->cost_part.p, nparts, alpha,
                   cost_dev ? cost_dev : (double*)c->cost_dev.p);
        WFM_CK_LAUNCH(c, "k_conv_cost_final");
    }
    // W = FFT3(w r)
    a.real_in = (const double*)c->R.p;
    if ((rc = conv_r2c_n<CS_CPLX>(c, a))) return rc;
    if ((rc = conv_cols_n<CS_CPLX>(c, a, 1, P))) return rc;
    // grad = alpha * IFFT3(W conj(X))
    if ((rc = conv_cols_zz_n<CS_MULCX_CONJ>(c, a, P))) return rc;
    if ((rc = conv_cols_n<CS_CPLX>(c, a, 1, P))) return rc;
    if ((rc = conv_c2r_n<CS_GRAD>(c, a, nullptr))) return rc;
    return WFM_OK;
}

// host-buffer variant: copies h in, gradient and cost out (synchronous)
int wfm_conv_cost_and_gradient(wfm_conv* c, double alpha, const void* h_host, void* grad_host, int clr, double* cost) {
    if (!c) return WFM_ERR_INVALID_ARG;
    if (!h_host || !grad_host || !cost) return c->fail(WFM_ERR_INVALID_ARG, "h / grad / cost is NULL");
    if (c->multi()) {
        if (!c->have_obj || !c->have_data) return c->fail(WFM_ERR_STATE, "object and data must be set first");
        return wfm_multi::conv_cost_and_gradient_host(c, alpha, h_host, grad_host, clr, cost);
    }
    WFM_ENTER(c);
    const size_t bytes = 8 * c->vox();
    WFM_CK(c, c->hdev.ensure(bytes));
    WFM_CK(c, c->gdev.ensure(bytes));
    WFM_CK(c, cudaMemcpyAsync(c->hdev.p, h_host, bytes, cudaMemcpyHostToDevice, c->stream));
    if (!clr) WFM_CK(c, cudaMemcpyAsync(c->gdev.p, grad_host, bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = wfm_conv_cost_and_gradient_dev(c, alpha, c->hdev.p, c->gdev.p, clr, nullptr); if (rc) return rc;
    WFM_CK(c, cudaMemcpyAsync(grad_host, c->gdev.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    WFM_CK(c, cudaMemcpyAsync(cost, c->cost_dev.p, 8, cudaMemcpyDeviceToHost, c->stream));
    WFM_CK(c, cudaStreamSynchronize(c->stream));
    return WFM_OK;
}

// One COMPUTE_FG step of PSF_Estimation.fitPSF (PSF_Estimation.java:202-217) entirely on the device:
//   pupil.setParam(x) -> pupil.computePsf() -> fcost = fdata.computeCostAndGradient(1.0, psf, gcost, true)
//   -> gX = pupil.apply_Jacobian(gcost, x.getSpace())
// Only x (n doubles) goes to the device and {cost, gX} come back.  param: WFM_DEFOCUS / WFM_PHASE / WFM_MODULUS.
int wfm_eval_fg(wfm_model* h, wfm_conv* c, int param, const double* x, int n, double alpha, double* cost, double* grad_out) {
    if (!h || !c) return WFM_ERR_INVALID_ARG;
    if (!cost || !grad_out) return h->fail(WFM_ERR_INVALID_ARG, "cost / grad_out is NULL");
    if (h->multi() != c->multi()) return h->fail(WFM_ERR_INVALID_ARG, "model and data term must both be multi-device handles, or neither");
    if (h->multi()) {                 // z-slabs on every device, transposes over NVLink peer memory (wfm_conv_multi.inl)
        if (h->precision != WFM_F64) return h->fail(WFM_ERR_UNSUPPORTED, "wfm_eval_fg is fp64 only in this revision");
        if (h->N != c->nx || h->nz_global != c->nz) return h->fail(WFM_ERR_INVALID_ARG, "model and data term must have the same shape");
        int rcm; unsigned km;
        switch (param) {
            case WFM_DEFOCUS: rcm = x ? wfm_set_defocus(h, x, n) : WFM_OK; km = WFM_J_DEFOCUS; if (!x) n = h->ndefocus; break;
            case WFM_PHASE: rcm = x ? wfm_set_phase(h, x, n) : WFM_OK; km = WFM_J_PHASE; break;
            case WFM_MODULUS: rcm = x ? wfm_set_modulus(h, x, n) : WFM_OK; km = WFM_J_MODULUS; break;
            default: return h->fail(WFM_ERR_INVALID_ARG, "DoubleShapedVector param does not belong to any space");
        }
        if (rcm) return rcm;
        return wfm_multi::eval_fg(h, c, param, n, alpha, cost, grad_out, km);
    }
    if (h->precision != WFM_F64) return h->fail(WFM_ERR_UNSUPPORTED, "wfm_eval_fg is fp64 only in this revision");
    if (h->N != c->nx || h->nz_global != c->nz || h->z0 != 0 || h->nzl != h->nz_global || h->nbatch != 1)
        return h->fail(WFM_ERR_INVALID_ARG, "model and data term must have the same (unsharded) shape");
    if (h->device != c->device) return h->fail(WFM_ERR_INVALID_ARG, "model and data term live on different devices");
    int rc;
    unsigned kinds;
    // x == NULL: the caller has already taken the setParam(x) step through the matching setter (the host mirrors do,
    // so that their parameterCoefs / ni / deltaX / deltaY stay in step, WFM:412-422, 1516-1531)
    switch (param) {                                                           // WFM:412-422
        case WFM_DEFOCUS: rc = x ? wfm_set_defocus(h, x, n) : WFM_OK; kinds = WFM_J_DEFOCUS; if (!x) n = h->ndefocus; break;
        case WFM_PHASE: rc = x ? wfm_set_phase(h, x, n) : WFM_OK; kinds = WFM_J_PHASE; break;
        case WFM_MODULUS: rc = x ? wfm_set_modulus(h, x, n) : WFM_OK; kinds = WFM_J_MODULUS; break;
        default: return h->fail(WFM_ERR_INVALID_ARG, "DoubleShapedVector param does not belong to any space");
    }
    if (rc) return rc;
    WFM_ENTER(h);
    if ((rc = compute_psf_impl(h))) return rc;
    WFM_CK(h, c->gdev.ensure(8 * c->vox()));                // (before the stream swap: no early return while it is swapped)
    cudaStream_t saved = c->stream;
    c->stream = h->stream;                                  // one stream: the steps are ordered
    rc = wfm_conv_cost_and_gradient_dev(c, alpha, h->psf.p, c->gdev.p, 1, nullptr);
    c->stream = saved;
    if (rc) return h->fail(rc, "%s", c->err.c_str());
    if ((rc = wfm_apply_jacobian_dev(h, kinds, c->gdev.p, (double*)h->grad.p))) return rc;
    std::vector<double> g(h->glen());
    WFM_CK(h, cudaMemcpyAsync(g.data(), h->grad.p, 8 * g.size(), cudaMemcpyDeviceToHost, h->stream));
    WFM_CK(h, cudaMemcpyAsync(cost, c->cost_dev.p, 8, cudaMemcpyDeviceToHost, h->stream));
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    if ((rc = check_pipeline(h))) return rc;
    const int off = (param == WFM_DEFOCUS) ? 0 : (param == WFM_PHASE ? 3 : 3 + h->nphase);
    const int len = (param == WFM_DEFOCUS) ? n : (param == WFM_PHASE ? h->nphase : h->nmod);
    memcpy(grad_out, g.data() + off, 8 * (size_t)len);
    return WFM_OK;
}

// getMtf() WFM:1807-1828 as intended: FFT3 of the PSF (DoubleFFT_3D.complexForward on the zero-imaginary copy).
int wfm_get_mtf(wfm_model* h, void* out_host) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "getMtf (the 3-D transform crosses the z-slabs)");
    if (!out_host) return h->fail(WFM_ERR_INVALID_ARG, "output pointer is NULL");
    if (h->precision != WFM_F64) return h->fail(WFM_ERR_UNSUPPORTED, "getMtf is fp64 only in this revision");
    if (h->z0 != 0 || h->nzl != h->nz_global) return h->fail(WFM_ERR_UNSUPPORTED, "getMtf needs the whole stack on one handle");
    if (!supported_n(h->nz_global)) return h->fail(WFM_ERR_UNSUPPORTED, "getMtf needs Nz to be a power of two in [32, 2048]");
    WFM_ENTER(h);
    int rc = compute_psf_impl(h); if (rc) return rc;
    wfm_conv* c = nullptr;
    rc = conv_create_impl(&c, h->N, h->N, h->nz_global, WFM_F64, h->device, true);
    if (rc) return h->fail(rc, "%s", g_create_error.c_str());
    c->stream = h->stream;
    ConvArgs<double> a = conv_args(c);
    a.real_in = (const double*)h->psf.p;
    a.Xout = (double2*)c->V.p;
    if (!(rc = conv_rows_n<CL_REAL, CS_CPLX>(c, a)) && !(rc = conv_cols_n<CS_CPLX>(c, a, 1, c->nx)) &&
        !(rc = conv_cols_n<CS_CPLX>(c, a, 2, c->nx))) {
        cudaError_t e = cudaMemcpyAsync(out_host, c->V.p, 16 * c->vox(), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = h->fail(WFM_ERR_CUDA, "MTF copy failed: %s", cudaGetErrorString(e));
    } else {
        h->fail(rc, "%s", c->err.c_str());
    }
    c->stream = c->own_stream;
    wfm_conv_destroy(c);
    return rc ? rc : check_pipeline(h);
}

}  // extern "C"
