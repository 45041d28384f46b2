// wfm_conv_multi.inl -- the FFT-convolution data term sharded by z-slab across the GPUs of one box
// (SURVEY.md section 8, row f1 "across GPUs"; included by wfm_api.cu after wfm_conv_api.inl and wfm_multi.inl).
//
// PSF_Estimation.fitPSF evaluates cost + gradient once per iteration (PSF_Estimation.java:202-217).  With the PSF
// model sharded by z-slab (wfm_create_multi) the data term has to follow, or the inner loop falls back to host
// buffers and PCIe.  Layouts and the exchange are described in wfm_conv.cuh (ConvPeers): real-space volumes and the
// x / y passes in z-slabs, the z pass in pencils, and the two transposes done by the STORES of the neighbouring
// passes over NVLink peer memory -- the y-pass epilogue writes pencils, the fused z-pass epilogue writes slabs.
//
// One host thread per device queues that device's passes on the stream of the model's child handle; between a
// scatter pass and its consumer the threads meet at a host barrier (all events recorded), then every stream waits
// for every other device's event.  No kernel ever spins on another device.
#include <condition_variable>
#include <mutex>

namespace wfm_multi {

// reusable host barrier for the per-device threads (std::barrier is C++20)
struct HostBarrier {
    explicit HostBarrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned gen = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != gen; });
    }
    std::mutex m_; std::condition_variable cv_; int n_, count_ = 0; unsigned gen_ = 0;
};

// per-device threads over the children of a conv parent
int conv_parallel(wfm_conv* p, const std::function<int(int)>& fn) {
    const int n = (int)p->parts.size();
    std::vector<int> rc(n, WFM_OK);
    if (n == 1) rc[0] = fn(0);
    else {
        std::vector<std::thread> th;
        th.reserve(n);
        for (int i = 0; i < n; ++i) th.emplace_back([&, i]() { rc[i] = fn(i); });
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < n; ++i)
        if (rc[i]) { p->err = p->parts[i]->err; return rc[i]; }
    return WFM_OK;
}

// Every device has queued its scatter pass: record, meet, make this stream wait for all the others.  Always two
// rendezvous, whatever fails: a device whose launch failed still takes part, so nobody is left waiting.
int device_barrier(wfm_conv* p, int i, HostBarrier& hb) {
    wfm_conv* c = p->parts[i];
    int rc = WFM_OK;
    cudaError_t e = cudaEventRecord(c->ev_pass, c->stream);
    if (e != cudaSuccess) rc = c->fail(WFM_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e));
    hb.wait();                         // every event of this round has been recorded
    for (int o = 0; o < (int)p->parts.size() && !rc; ++o) {
        if (o == i) continue;
        e = cudaStreamWaitEvent(c->stream, p->parts[o]->ev_pass, 0);
        if (e != cudaSuccess) rc = c->fail(WFM_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
    }
    hb.wait();                         // nobody re-records its event before everybody has queued its waits
    return rc;
}

ConvArgs<double> child_args(wfm_conv* p, wfm_conv* c) {
    ConvArgs<double> a = conv_args(c);
    a.nz = c->nz;                                            // slab planes
    a.inv_ntot = 1.0 / ((double)c->nx * (double)c->ny * (double)c->nz_all);
    a.peers.ny_full = c->ny;
    return a;
}
void to_pencils(wfm_conv* p, wfm_conv* c, ConvArgs<double>& a) {      // y-pass epilogue -> pencil volumes
    for (size_t o = 0; o < p->parts.size(); ++o) a.peers.vol[o] = (double2*)p->parts[o]->Vp.p;
    a.peers.split = SplitMap{c->ny, (int)p->parts.size()};
    a.peers.src_first = c->z0;
}
void to_slabs(wfm_conv* p, wfm_conv* c, ConvArgs<double>& a) {        // z-pass epilogue -> slab volumes
    for (size_t o = 0; o < p->parts.size(); ++o) a.peers.vol[o] = (double2*)p->parts[o]->V.p;
    a.peers.split = SplitMap{c->nz_all, (int)p->parts.size()};
    a.peers.src_first = c->y0;
}

// The stages below take the running status `rc`: once it is set they queue nothing more but still keep every
// rendezvous, so that the per-device threads stay in step.

// slab of a real volume -> half spectrum in pencil layout (x pass, y pass with scatter); then the barrier
void forward_to_pencils(wfm_conv* p, int i, const double* real_slab, HostBarrier& hb, int& rc) {
    wfm_conv* c = p->parts[i];
    if (!rc) {
        ConvArgs<double> a = child_args(p, c);
        a.real_in = real_slab;
        rc = conv_r2c_n<CS_CPLX>(c, a);
        if (!rc) { to_pencils(p, c, a); rc = conv_cols_scatter_n(c, a, c->pitch()); }
    }
    const int rb = device_barrier(p, i, hb);
    if (!rc) rc = rb;
}

// pencils: FFT_z, product with X (or conj X), conj, FFT_z; back to the slabs; barrier; y pass (local)
template <int MUL> void through_z_and_back(wfm_conv* p, int i, HostBarrier& hb, int& rc) {
    wfm_conv* c = p->parts[i];
    if (!rc) {
        ConvArgs<double> a = child_args(p, c);
        a.V = (double2*)c->Vp.p; a.X = (const double2*)c->Xp.p;
        a.ny = c->nyl; a.nz = c->nz_all;
        to_slabs(p, c, a);
        a.peers.ny_full = c->ny;
        rc = conv_cols_zz_scatter_n<MUL>(c, a, c->pitch());
    }
    const int rb = device_barrier(p, i, hb);
    if (!rc) rc = rb;
    if (!rc) {
        ConvArgs<double> b = child_args(p, c);
        rc = conv_cols_n<CS_CPLX>(c, b, 1, c->pitch());
    }
}

// cost + gradient of one child; h_slab / grad_slab are device pointers on the child's device.  `rc`: status so far
// (e.g. of the caller's uploads); four rendezvous rounds in every case.
int child_cost_and_gradient(wfm_conv* p, int i, double alpha, const double* h_slab, double* grad_slab, int clr, HostBarrier& hb,
                            int rc = WFM_OK) {
    wfm_conv* c = p->parts[i];
    DeviceScope scope(c->device);
    forward_to_pencils(p, i, h_slab, hb, rc);
    through_z_and_back<CS_MULX_CONJ>(p, i, hb, rc);
    if (!rc) {
        ConvArgs<double> a = child_args(p, c);
        a.cost_part = (double*)c->cost_part.p; a.resid = (double*)c->R.p;
        int nparts = 0;
        rc = conv_c2r_n<CS_RESID>(c, a, &nparts);
        if (!rc) {
            auto kfin = &k_conv_cost_final;
            WFM_LAUNCH(kfin, dim3(1), dim3(1024), 0, c->stream, (const double*)c->cost_part.p, nparts, alpha, (double*)c->cost_dev.p);
        }
    }
    forward_to_pencils(p, i, (const double*)c->R.p, hb, rc);
    through_z_and_back<CS_MULCX_CONJ>(p, i, hb, rc);
    if (!rc) {
        ConvArgs<double> g = child_args(p, c);
        g.grad = grad_slab; g.alpha = alpha; g.clear_grad = clr ? 1 : 0;
        rc = conv_c2r_n<CS_GRAD>(c, g, nullptr);
    }
    return rc;
}

int conv_destroy(wfm_conv* p) {
    for (wfm_conv* c : p->parts) {
        { DeviceScope s(c->device); c->Vp.release(); c->Xp.release(); if (c->ev_pass) cudaEventDestroy(c->ev_pass); c->ev_pass = nullptr; }
        wfm_conv_destroy(c);
    }
    p->parts.clear();
    delete p;
    return WFM_OK;
}

// enable peer access between every pair of the listed devices (stores of the scatter passes)
int enable_all_peers(const int* devices, int n, std::string& err) {
    for (int i = 0; i < n; ++i) {
        DeviceScope s(devices[i]);
        for (int o = 0; o < n; ++o) {
            if (o == i) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[o]) != cudaSuccess || !can) {
                err = "the z-sharded data term needs peer access between all its devices (NVLink / NVSwitch)";
                return WFM_ERR_UNSUPPORTED;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[o], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { err = "cudaDeviceEnablePeerAccess failed"; return WFM_ERR_CUDA; }
            cudaGetLastError();
        }
    }
    return WFM_OK;
}

}  // namespace wfm_multi

extern "C" {

// WeightedConvolutionCost.build(space) over n_dev devices: the same z-slab split as wfm_create_multi.
int wfm_conv_create_multi(wfm_conv** out, int nx, int ny, int nz, int precision, const int* devices, int n_dev) {
    if (!out) { g_create_error = "out is NULL"; return WFM_ERR_INVALID_ARG; }
    *out = nullptr;
    if (!devices || n_dev < 1 || n_dev > WFM_MAX_RANKS) { g_create_error = "bad device list"; return WFM_ERR_INVALID_ARG; }
    if (nx != ny) { g_create_error = "Nx should equal Ny"; return WFM_ERR_INVALID_ARG; }
    if (!supported_n(nx) || !supported_n(nz)) { g_create_error = "Nx and Nz must be powers of two in [32, 2048]"; return WFM_ERR_UNSUPPORTED; }
    if (precision != WFM_F64) { g_create_error = "the convolution data term is fp64 only in this revision"; return WFM_ERR_UNSUPPORTED; }
    if (nz < n_dev || ny < n_dev) { g_create_error = "fewer planes / rows than devices"; return WFM_ERR_INVALID_ARG; }
    for (int i = 0; i < n_dev; ++i)
        for (int k = 0; k < i; ++k)
            if (devices[i] == devices[k]) { g_create_error = "a device appears twice in the device list"; return WFM_ERR_INVALID_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { g_create_error = "no CUDA device available (this library has no CPU fallback)"; return WFM_ERR_CUDA; }
    for (int i = 0; i < n_dev; ++i) if (devices[i] < 0 || devices[i] >= ndev) { g_create_error = "bad device index"; return WFM_ERR_INVALID_ARG; }
    {
        std::string perr;
        const int rc = wfm_multi::enable_all_peers(devices, n_dev, perr);
        if (rc) { g_create_error = perr; return rc; }
    }
    wfm_conv* p = new (std::nothrow) wfm_conv();
    if (!p) { g_create_error = "out of host memory"; return WFM_ERR_NOMEM; }
    p->nx = nx; p->ny = ny; p->nz = nz; p->nz_all = nz; p->precision = precision; p->device = devices[0];
    const SplitMap zs{nz, n_dev}, ys{ny, n_dev};
    for (int i = 0; i < n_dev; ++i) {
        wfm_conv* c = nullptr;
        // the child is a plain handle over its slab; the z twiddles are re-made for the full length below
        int rc = conv_create_impl(&c, nx, ny, zs.count(i), precision, devices[i], false, /*slab_of=*/nz);
        if (rc) { wfm_multi::conv_destroy(p); return rc; }
        c->parent = p; c->part = i; c->nz_all = nz; c->z0 = zs.first(i); c->y0 = ys.first(i); c->nyl = ys.count(i);
        p->parts.push_back(c);
        DeviceScope s(devices[i]);
        if (c->Vp.ensure(16 * c->pvox()) != cudaSuccess || c->Xp.ensure(16 * c->pvox()) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_pass, cudaEventDisableTiming) != cudaSuccess) {
            wfm_multi::conv_destroy(p); g_create_error = "device allocation failed"; return WFM_ERR_NOMEM;
        }
    }
    *out = p;
    return WFM_OK;
}

int wfm_conv_parts(const wfm_conv* c) { return c ? (int)c->parts.size() : 0; }

}  // extern "C"

namespace wfm_multi {

// copy the slabs of a host volume to the children, one thread per device
int conv_scatter_host(wfm_conv* p, const void* host, DevBuf wfm_conv::*buf) {
    return conv_parallel(p, [&](int i) {
        wfm_conv* c = p->parts[i];
        DeviceScope s(c->device);
        const size_t plane = 8 * (size_t)c->nx * c->ny;
        WFM_CK(c, (c->*buf).ensure(plane * c->nz));
        WFM_CK(c, cudaMemcpyAsync((c->*buf).p, (const char*)host + plane * c->z0, plane * c->nz, cudaMemcpyHostToDevice, c->stream));
        WFM_CK(c, cudaStreamSynchronize(c->stream));
        return (int)WFM_OK;
    });
}

int conv_set_object(wfm_conv* p, const void* obj_host) {
    HostBarrier hb((int)p->parts.size());
    int rc = conv_parallel(p, [&](int i) {
        wfm_conv* c = p->parts[i];
        DeviceScope s(c->device);
        const size_t plane = 8 * (size_t)c->nx * c->ny;
        cudaError_t e = c->hdev.ensure(plane * c->nz);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->hdev.p, (const char*)obj_host + plane * c->z0, plane * c->nz, cudaMemcpyHostToDevice, c->stream);
        int r = e == cudaSuccess ? (int)WFM_OK : c->fail(WFM_ERR_CUDA, "object upload failed: %s", cudaGetErrorString(e));
        forward_to_pencils(p, i, (const double*)c->hdev.p, hb, r);
        if (r) return r;
        ConvArgs<double> a = child_args(p, c);                  // X = FFT_z on the pencils (no product, no way back)
        a.V = (double2*)c->Vp.p; a.Xout = (double2*)c->Xp.p; a.ny = c->nyl; a.nz = c->nz_all;
        r = conv_cols_zall_n<CS_SPECTRUM>(c, a, c->pitch());
        if (r) return r;
        WFM_CK(c, cudaStreamSynchronize(c->stream));
        c->have_obj = true;
        return (int)WFM_OK;
    });
    if (!rc) p->have_obj = true;
    return rc;
}

// host buffers: h and the gradient are whole volumes; the slabs travel over their own PCIe links
int conv_cost_and_gradient_host(wfm_conv* p, double alpha, const void* h_host, void* grad_host, int clr, double* cost) {
    HostBarrier hb((int)p->parts.size());
    std::vector<double> part(p->parts.size(), 0.0);
    int rc = conv_parallel(p, [&](int i) {
        wfm_conv* c = p->parts[i];
        DeviceScope s(c->device);
        const size_t plane = 8 * (size_t)c->nx * c->ny, bytes = plane * c->nz;
        cudaError_t e = c->hdev.ensure(bytes);
        if (e == cudaSuccess) e = c->gdev.ensure(bytes);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->hdev.p, (const char*)h_host + plane * c->z0, bytes, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && !clr) e = cudaMemcpyAsync(c->gdev.p, (const char*)grad_host + plane * c->z0, bytes, cudaMemcpyHostToDevice, c->stream);
        int r = e == cudaSuccess ? (int)WFM_OK : c->fail(WFM_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
        r = child_cost_and_gradient(p, i, alpha, (const double*)c->hdev.p, (double*)c->gdev.p, clr, hb, r);
        if (r) return r;
        WFM_CK(c, cudaMemcpyAsync((char*)grad_host + plane * c->z0, c->gdev.p, bytes, cudaMemcpyDeviceToHost, c->stream));
        WFM_CK(c, cudaMemcpyAsync(&part[i], c->cost_dev.p, 8, cudaMemcpyDeviceToHost, c->stream));
        WFM_CK(c, cudaStreamSynchronize(c->stream));
        return (int)WFM_OK;
    });
    if (rc) return rc;
    double s = 0.0;
    for (double v : part) s += v;                               // device order: deterministic
    *cost = s;
    return WFM_OK;
}

// One COMPUTE_FG step over the devices of a multi model + multi data term (same device list, same z split).
int eval_fg(wfm_model* h, wfm_conv* p, int param, int n, double alpha, double* cost, double* grad_out, unsigned kinds) {
    const int nd = (int)h->parts.size();
    if ((int)p->parts.size() != nd) return h->fail(WFM_ERR_INVALID_ARG, "model and data term are spread over different device lists");
    for (int i = 0; i < nd; ++i) {
        const wfm_model* m = h->parts[i]; const wfm_conv* c = p->parts[i];
        if (m->device != c->device || m->z0 != c->z0 || m->nzl != c->nz)
            return h->fail(WFM_ERR_INVALID_ARG, "model and data term are spread over different device lists / slabs");
    }
    if (!p->have_obj || !p->have_data) return h->fail(WFM_ERR_STATE, "object and data must be set first");
    const int L = h->parts[0]->glen();
    double* slots = (double*)h->xslots.p;
    HostBarrier hb(nd);
    std::vector<double> part(nd, 0.0);
    std::vector<cudaStream_t> saved(nd);
    for (int i = 0; i < nd; ++i) { saved[i] = p->parts[i]->stream; p->parts[i]->stream = h->parts[i]->stream; }   // one stream per device
    int rc = conv_parallel(p, [&](int i) {
        wfm_model* m = h->parts[i]; wfm_conv* c = p->parts[i];
        DeviceScope s(c->device);
        int r = compute_psf_impl(m);
        if (r) c->err = m->err;
        if (!r && c->gdev.ensure(8 * (size_t)c->nx * c->ny * c->nz) != cudaSuccess) r = c->fail(WFM_ERR_NOMEM, "device allocation failed");
        r = child_cost_and_gradient(p, i, alpha, (const double*)m->psf.p, (double*)c->gdev.p, 1, hb, r);
        if (r) return r;
        double* target = h->peer_direct[i] ? slots + (size_t)i * L : (double*)m->grad.p;
        r = wfm_apply_jacobian_dev(m, kinds, c->gdev.p, target);
        if (r) { c->err = m->err; return r; }
        if (!h->peer_direct[i]) WFM_CK(c, cudaMemcpyPeerAsync(slots + (size_t)i * L, h->device, m->grad.p, m->device, 8 * (size_t)L, m->stream));
        WFM_CK(c, cudaEventRecord(h->part_done[i], m->stream));
        WFM_CK(c, cudaMemcpyAsync(&part[i], c->cost_dev.p, 8, cudaMemcpyDeviceToHost, m->stream));
        return (int)WFM_OK;
    });
    for (int i = 0; i < nd; ++i) p->parts[i]->stream = saved[i];
    if (rc) return h->fail(rc, "%s", p->err.c_str());
    wfm_model* m0 = h->parts[0];
    DeviceScope s(m0->device);
    for (int i = 1; i < nd; ++i) WFM_CK(h, cudaStreamWaitEvent(m0->stream, h->part_done[i], 0));
    auto ksum = &k_sum_slots;
    WFM_LAUNCH(ksum, dim3((L + 127) / 128), dim3(128), 0, m0->stream, (const double*)slots, nd, L, (double*)m0->grad.p);
    WFM_CK_LAUNCH(h, "k_sum_slots");
    std::vector<double> g(L);
    WFM_CK(h, cudaMemcpyAsync(g.data(), m0->grad.p, 8 * (size_t)L, cudaMemcpyDeviceToHost, m0->stream));
    for (int i = 0; i < nd; ++i) {
        DeviceScope si(h->parts[i]->device);
        WFM_CK(h, cudaStreamSynchronize(h->parts[i]->stream));
        int r = check_pipeline(h->parts[i]); if (r) { h->err = h->parts[i]->err; return r; }
    }
    double csum = 0.0;
    for (double v : part) csum += v;
    *cost = csum;
    const int off = (param == WFM_DEFOCUS) ? 0 : (param == WFM_PHASE ? 3 : 3 + h->nphase);
    const int len = (param == WFM_DEFOCUS) ? n : (param == WFM_PHASE ? h->nphase : h->nmod);
    memcpy(grad_out, g.data() + off, 8 * (size_t)len);
    return WFM_OK;
}

}  // namespace wfm_multi
