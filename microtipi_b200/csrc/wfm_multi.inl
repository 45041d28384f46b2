// wfm_multi.inl -- multi-GPU behind the C ABI (included by wfm_api.cu).
//
// The reference's only parallel runtime is a thread pool inside ONE process (WFM:287-288, one Callable per
// plane, 291-333) and its caller is one JVM thread (PSF_Estimation.java:202-217).  A drop-in therefore has to
// drive the 8 GPUs of a box from that one caller:
//
//   wfm_create_multi(devices, n_dev)   one z-slab child handle per device (contiguous slabs, the first nz % n_dev
//                                      devices hold one extra plane); every child runs on its own stream
//   setters                            broadcast to every child (each device rebuilds its own pupil: Npix work)
//   wfm_get_psf / _cpx_psf             every device copies its slab straight to its offset of the caller's array,
//                                      one host thread per device -> n_dev PCIe links in parallel
//   wfm_apply_j_*                      q is split the same way; the n_dev partial K-vectors are added in device order
//   wfm_multi_apply_jacobian_dev       device-resident q slabs; every device's k_jac_final stores its partial K-vector
//                                      into a slot on the first device over NVLink (peer mapping), events order the
//                                      first device's k_sum_slots behind them -- no host round trip, no NCCL
//
// and, for the one-process-per-GPU layout (torchrun), the same exchange across processes over CUDA IPC:
//   wfm_exchange_export / _connect     every rank's landing buffer is mapped into every peer; wfm_apply_jacobian_dev
//                                      then returns the sum over the ranks (k_jac_final, XchgArgs).
#include <thread>

namespace wfm_multi {

int unsupported(wfm_model* h, const char* what) {
    return h->fail(WFM_ERR_UNSUPPORTED, "not available on a multi-device handle: %s", what);
}

// fn(child) on every child, in device order; the calls only queue work, so the devices run concurrently
int broadcast(wfm_model* h, const std::function<int(wfm_model*)>& fn) {
    for (wfm_model* c : h->parts) {
        const int rc = fn(c);
        if (rc) { h->err = c->err; return rc; }
    }
    return WFM_OK;
}

// fn(part index) on one host thread per device (bulk copies: n_dev PCIe links at once)
int parallel(wfm_model* h, const std::function<int(int)>& fn) {
    const int n = (int)h->parts.size();
    std::vector<int> rc(n, WFM_OK);
    if (n == 1) {
        rc[0] = fn(0);
    } else {
        std::vector<std::thread> th;
        th.reserve(n);
        for (int i = 0; i < n; ++i) th.emplace_back([&, i]() { rc[i] = fn(i); });
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < n; ++i)
        if (rc[i]) { h->err = h->parts[i]->err; return rc[i]; }
    return WFM_OK;
}

// the parent mirrors what the validations of the entry points read
int sync_meta(wfm_model* h) {
    const wfm_model* c = h->parts[0];
    h->nphase = c->nphase; h->nmod = c->nmod; h->nzern = c->nzern; h->radial = c->radial; h->ndefocus = c->ndefocus;
    h->have_optics = c->have_optics; h->have_rho = c->have_rho; h->modulus_mode = c->modulus_mode;
    h->lambda_ni = c->lambda_ni; h->deltaX = c->deltaX; h->deltaY = c->deltaY; h->ni = c->ni;
    h->pstate = 0;
    return WFM_OK;
}

int compute_psf(wfm_model* h) {
    return broadcast(h, [](wfm_model* c) { DeviceScope s(c->device); return compute_psf_impl(c); });
}

// getPsf() / get_cpxPsf(): compute where dirty (all devices queue their pipelines first), then every device copies
// its slab to its own offset of the caller's array.
int get_stack(wfm_model* h, void* out, bool cpx, bool async) {
    if (!out) return h->fail(WFM_ERR_INVALID_ARG, "output pointer is NULL");
    int rc = compute_psf(h); if (rc) return rc;
    const size_t plane = (size_t)h->npix() * h->esz() * (cpx ? 2 : 1);
    if (async) {
        if (cpx) return unsupported(h, "asynchronous read-back of cpxPsf");
        return broadcast(h, [&](wfm_model* c) {
            const size_t off = plane * (size_t)(c->z0 - h->z0);
            return wfm_get_psf_async(c, (char*)out + off);
        });
    }
    return parallel(h, [&](int i) {
        wfm_model* c = h->parts[i];
        const size_t off = plane * (size_t)(c->z0 - h->z0);
        return cpx ? wfm_get_cpx_psf(c, (char*)out + off) : wfm_get_psf(c, (char*)out + off);
    });
}

int wait_transfers(wfm_model* h) { return broadcast(h, [](wfm_model* c) { return wfm_wait_transfers(c); }); }
int synchronize(wfm_model* h) { return broadcast(h, [](wfm_model* c) { return wfm_synchronize(c); }); }

// host q: split by slab, one host thread per device (H2D of the slab, Jacobian, D2H of the partial K-vector);
// the partial vectors are added in device order on the host (deterministic)
int apply_host(wfm_model* h, unsigned kinds, const void* q_host, std::vector<double>& g) {
    const int L = h->parts[0]->glen();
    const int n = (int)h->parts.size();
    std::vector<std::vector<double>> part(n);
    const size_t plane = (size_t)h->npix() * h->esz();
    int rc = parallel(h, [&](int i) {
        wfm_model* c = h->parts[i];
        DeviceScope s(c->device);
        const size_t off = plane * (size_t)(c->z0 - h->z0);
        part[i].resize(L);
        return apply_host_single(c, kinds, (const char*)q_host + off, part[i].data());
    });
    if (rc) return rc;
    g.assign(L, 0.0);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < L; ++k) g[k] += part[i][k];
    return WFM_OK;
}

int kernel_times(wfm_model* h, double* ms, uint64_t* counts) {       // the slowest device per kernel group
    for (int k = 0; k < WFM_KERNEL_IDS; ++k) { ms[k] = 0.0; counts[k] = 0; }
    for (wfm_model* c : h->parts) {
        double m[WFM_KERNEL_IDS]; uint64_t n[WFM_KERNEL_IDS];
        int rc = wfm_get_kernel_times(c, m, n, WFM_KERNEL_IDS);
        if (rc) { h->err = c->err; return rc; }
        for (int k = 0; k < WFM_KERNEL_IDS; ++k) if (m[k] > ms[k]) { ms[k] = m[k]; counts[k] = n[k]; }
    }
    return WFM_OK;
}

int destroy(wfm_model* h) {
    for (wfm_model* c : h->parts) wfm_destroy(c);
    h->parts.clear();
    {
        DeviceScope s(h->device);
        h->xslots.release();
        for (cudaEvent_t e : h->part_done) if (e) cudaEventDestroy(e);
    }
    delete h;
    return WFM_OK;
}

}  // namespace wfm_multi

extern "C" {

// new WideFieldModel(...) over n_dev devices of one box (device_list, n_dev).
int wfm_create_multi(wfm_model** out, int nx, int ny, int nz, double dxy, double dz, int precision, const int* devices,
                     int n_dev) {
    if (!out) { g_create_error = "out is NULL"; return WFM_ERR_INVALID_ARG; }
    *out = nullptr;
    if (!devices || n_dev < 1 || n_dev > WFM_MAX_RANKS) { g_create_error = "bad device list"; return WFM_ERR_INVALID_ARG; }
    if (nz < n_dev) { g_create_error = "fewer z-planes than devices"; return WFM_ERR_INVALID_ARG; }
    for (int i = 0; i < n_dev; ++i)
        for (int k = 0; k < i; ++k)
            if (devices[i] == devices[k]) { g_create_error = "a device appears twice in the device list"; return WFM_ERR_INVALID_ARG; }
    wfm_model* h = new (std::nothrow) wfm_model();
    if (!h) { g_create_error = "out of host memory"; return WFM_ERR_NOMEM; }
    h->N = nx; h->nz_global = nz; h->z0 = 0; h->nzl = nz; h->nzm = nz; h->nbatch = 1; h->dxy = dxy; h->dz = dz;
    h->precision = precision; h->device = devices[0];
    const int base = nz / n_dev, rem = nz % n_dev;
    int z0 = 0;
    for (int i = 0; i < n_dev; ++i) {
        const int nzl = base + (i < rem ? 1 : 0);
        wfm_model* c = nullptr;
        const int rc = wfm_create_slab(&c, nx, ny, nz, z0, nzl, dxy, dz, precision, devices[i]);
        if (rc != WFM_OK) { wfm_multi::destroy(h); return rc; }          // g_create_error set by the child
        c->siblings = n_dev;                                             // (the children share the host's cores when they stage)
        h->parts.push_back(c);
        h->part_z0.push_back(z0);
        z0 += nzl;
    }
    // landing buffer of the partial gradient vectors on the first device + peer access towards it
    {
        DeviceScope s(h->device);
        if (h->xslots.ensure(8 * (size_t)n_dev * (3 + 2 * WFM_MAX_COEF)) != cudaSuccess) {
            wfm_multi::destroy(h); g_create_error = "device allocation failed"; return WFM_ERR_NOMEM;
        }
    }
    h->peer_direct.assign(n_dev, 0);
    h->part_done.assign(n_dev, nullptr);
    h->peer_direct[0] = 1;
    for (int i = 0; i < n_dev; ++i) {
        DeviceScope s(devices[i]);
        if (cudaEventCreateWithFlags(&h->part_done[i], cudaEventDisableTiming) != cudaSuccess) {
            wfm_multi::destroy(h); g_create_error = "cudaEventCreate failed"; return WFM_ERR_CUDA;
        }
        if (i == 0) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) h->peer_direct[i] = 1;
            cudaGetLastError();                                            // (clear "already enabled")
        }
    }
    *out = h;
    return WFM_OK;
}

int wfm_multi_parts(const wfm_model* h) { return h ? (int)h->parts.size() : 0; }

int wfm_multi_part_info(const wfm_model* h, int part, int* device, int* z0, int* nz_local) {
    if (!h || part < 0 || part >= (int)h->parts.size()) return WFM_ERR_INVALID_ARG;
    const wfm_model* c = h->parts[part];
    if (device) *device = c->device;
    if (z0) *z0 = c->z0;
    if (nz_local) *nz_local = c->nzl;
    return WFM_OK;
}

// Borrowed child handle of one device (device-resident use: wfm_device_psf, wfm_fill_uniform, ...).  It belongs to the
// parent: do not destroy it, and do not call its setters (the parent broadcasts them).
int wfm_multi_part(wfm_model* h, int part, wfm_model** child) {
    if (!h || !child || part < 0 || part >= (int)h->parts.size()) return WFM_ERR_INVALID_ARG;
    *child = h->parts[part];
    return WFM_OK;
}

// Device-resident Jacobians of a multi handle: q_dev[i] = device pointer of part i's slab of q (on device i);
// grad_dev = 3 + nPhase + nModulus doubles ON THE FIRST DEVICE, complete (summed over the devices in device order)
// when the first device's stream reaches the end of this call's work.  Asynchronous.
int wfm_multi_apply_jacobian_dev(wfm_model* h, unsigned kinds, const void* const* q_dev, double* grad_dev) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (!h->multi()) return h->fail(WFM_ERR_INVALID_ARG, "not a multi-device handle");
    if (!q_dev || !grad_dev) return h->fail(WFM_ERR_INVALID_ARG, "q_dev / grad_dev is NULL");
    const int n = (int)h->parts.size();
    const int L = h->parts[0]->glen();
    double* slots = (double*)h->xslots.p;
    for (int i = 0; i < n; ++i) {
        wfm_model* c = h->parts[i];
        if (!q_dev[i]) return h->fail(WFM_ERR_INVALID_ARG, "q_dev[%d] is NULL", i);
        DeviceScope s(c->device);
        // k_jac_final of device i stores its partial K-vector straight into slot i on the first device (NVLink peer
        // store); without peer access it lands in the child's own buffer and is copied across
        double* target = h->peer_direct[i] ? slots + (size_t)i * L : (double*)c->grad.p;
        int rc = wfm_apply_jacobian_dev(c, kinds, q_dev[i], target);
        if (rc) { h->err = c->err; return rc; }
        if (!h->peer_direct[i])
            WFM_CK(h, cudaMemcpyPeerAsync(slots + (size_t)i * L, h->device, c->grad.p, c->device, 8 * (size_t)L, c->stream));
        WFM_CK(h, cudaEventRecord(h->part_done[i], c->stream));
    }
    wfm_model* c0 = h->parts[0];
    DeviceScope s(c0->device);
    for (int i = 1; i < n; ++i) WFM_CK(h, cudaStreamWaitEvent(c0->stream, h->part_done[i], 0));
    auto ksum = &k_sum_slots;
    WFM_LAUNCH(ksum, dim3((L + 127) / 128), dim3(128), 0, c0->stream, (const double*)slots, n, L, grad_dev);
    WFM_CK_LAUNCH(h, "k_sum_slots");
    return WFM_OK;
}

// ---- cross-process exchange over CUDA IPC peer memory ------------------------------------------------------------
// Call order on every rank: wfm_exchange_export (allocates this rank's landing buffer, returns its 64-byte IPC
// handle) -> all-gather the handles by any means (torch.distributed, MPI, a file) -> wfm_exchange_connect.  From then
// on wfm_apply_jacobian_dev leaves the SUM over the ranks in grad_dev on every rank.  Every rank must issue the same
// sequence of Jacobian calls (the exchange is a collective).
int wfm_exchange_export(wfm_model* h, int world, void* handle_out) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "wfm_exchange_export (a multi-device handle already sums over its devices)");
    if (!handle_out || world < 1 || world > WFM_MAX_RANKS) return h->fail(WFM_ERR_INVALID_ARG, "bad world size / handle pointer");
    if (h->nbatch > 1) return h->fail(WFM_ERR_UNSUPPORTED, "the gradient exchange needs a single-model handle");
    WFM_ENTER(h);
    wfm_exchange_close(h);
    Exchange* x = new (std::nothrow) Exchange();
    if (!x) return h->fail(WFM_ERR_NOMEM, "out of host memory");
    x->world = world;
    x->glen_cap = 3 + 2 * WFM_MAX_COEF;
    x->slot_bytes = sizeof(double) * 2 * (size_t)world * x->glen_cap;
    const size_t bytes = x->slot_bytes + sizeof(unsigned) * 2 * (size_t)world;
    if (cudaMalloc(&x->base, bytes) != cudaSuccess) { delete x; return h->fail(WFM_ERR_NOMEM, "device allocation failed"); }
    cudaError_t e = cudaMemset(x->base, 0, bytes);
    if (e == cudaSuccess) e = x->local.ensure(2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(x->local.p, 0, 2 * sizeof(unsigned));
    cudaIpcMemHandle_t ipc;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&ipc, x->base);
    if (e != cudaSuccess) {
        cudaFree(x->base); x->local.release(); delete x;
        return h->fail(WFM_ERR_CUDA, "exchange buffer set-up failed: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == WFM_EXCHANGE_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &ipc, sizeof(ipc));
    h->xchg = x;
    return WFM_OK;
}

int wfm_exchange_connect(wfm_model* h, int rank, int world, const void* handles) {
    if (!h) return WFM_ERR_INVALID_ARG;
    Exchange* x = h->xchg;
    if (!x) return h->fail(WFM_ERR_STATE, "wfm_exchange_export has not been called");
    if (!handles || world != x->world || rank < 0 || rank >= world) return h->fail(WFM_ERR_INVALID_ARG, "bad rank / world / handles");
    WFM_ENTER(h);
    x->rank = rank;
    x->mapped.assign(world, nullptr);
    for (int r = 0; r < world; ++r) {
        if (r == rank) { x->mapped[r] = x->base; continue; }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, (const char*)handles + (size_t)r * sizeof(ipc), sizeof(ipc));
        const cudaError_t e = cudaIpcOpenMemHandle(&x->mapped[r], ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            x->mapped[r] = nullptr;
            return h->fail(WFM_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
        }
    }
    x->epoch = 0;
    x->connected = true;
    return WFM_OK;
}

// 1 when the last exchanges all completed, WFM_ERR_INTERNAL when a peer's flag did not arrive (checked after a sync)
int wfm_exchange_status(wfm_model* h) {
    if (!h) return WFM_ERR_INVALID_ARG;
    Exchange* x = h->xchg;
    if (!x || !x->connected) return 0;
    WFM_ENTER(h);
    unsigned w[2] = {0, 0};
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    WFM_CK(h, cudaMemcpy(w, x->local.p, sizeof(w), cudaMemcpyDeviceToHost));
    if (w[1]) return h->fail(WFM_ERR_INTERNAL, "gradient exchange: a peer's flag did not arrive");
    return 1;
}

int wfm_exchange_close(wfm_model* h) {
    if (!h) return WFM_ERR_INVALID_ARG;
    Exchange* x = h->xchg;
    if (!x) return WFM_OK;
    DeviceScope s(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < (int)x->mapped.size(); ++r)
        if (x->mapped[r] && x->mapped[r] != x->base) cudaIpcCloseMemHandle(x->mapped[r]);
    if (x->base) cudaFree(x->base);
    x->local.release();
    delete x;
    h->xchg = nullptr;
    return WFM_OK;
}

}  // extern "C"
