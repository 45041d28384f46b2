/*
 * wfm_b200.h -- C ABI of the B200-native widefield PSF model + Jacobians.
 *
 * This is the drop-in boundary for ONE path of jplumail/microTiPi: the pupil ->
 * per-z defocus -> 2-D FFT -> |.|^2 pipeline of WideFieldModel.computePsf and its
 * adjoint pipeline apply_J_phase / apply_J_defocus / apply_J_modulus.  The
 * reference has no FFI seam of its own (it is pure Java); every entry point below
 * names the Java member it replaces so that a `WideFieldModel`-compatible Java
 * class can bind it through Panama FFM or JNI.
 *
 *   WFM = /root/reference/src/microTiPi/epifluorescence/WideFieldModel.java
 *   MM  = /root/reference/src/microTiPi/microscopy/MicroscopeModel.java
 *
 * Conventions
 *   - Plain C types only.  The caller owns every host pointer; the library copies
 *     in/out before returning.  Device memory is owned by the handle.
 *   - Every function returns an int status: WFM_OK (0) or a negative wfm_status.
 *     Nothing aborts or throws across the ABI.  wfm_last_error() gives the text.
 *     WFM_ERR_INVALID_ARG corresponds to the reference's IllegalArgumentException
 *     sites (WFM:159,407,420,1514,1530,1592,1629).
 *   - A handle is NOT thread-safe (the reference's callers are single threaded,
 *     PSF_Estimation.java:200-251); distinct handles may be used concurrently.
 *   - Layout is TiPi's first-index-fastest: pixel in = ix + Nx*iy (WFM:385);
 *     psf  flat index ix + Nx*(iy + Ny*izl)               (MM:73-76)
 *     cpx  flat index c + 2*(ix + Nx*(iy + Ny*izl)), c=0 re, c=1 im   (WFM:170,341)
 *     Z    flat index in + k*Npix, k < nzern                          (WFM:1605)
 *     where izl is the plane index inside the handle's z-slab [z0, z0+nz_local).
 *   - precision F64: psf/cpx/q are double.  precision F32 (`single=true`): psf/cpx/q
 *     are float, pupil trig and all Jacobian reductions stay double (WFM:243-245,
 *     791-802).  Gradient outputs are always double.
 *   - There is no CPU fallback: every compute entry point runs CUDA kernels on the
 *     handle's device and fails with WFM_ERR_CUDA if it cannot.
 */
#ifndef WFM_B200_H
#define WFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define WFM_API __declspec(dllexport)
#else
#define WFM_API __attribute__((visibility("default")))
#endif

typedef struct wfm_model wfm_model; /* opaque */

typedef enum wfm_status {
    WFM_OK = 0,
    WFM_ERR_INVALID_ARG = -1, /* IllegalArgumentException in the reference */
    WFM_ERR_UNSUPPORTED = -2, /* e.g. Nx not a supported power of two */
    WFM_ERR_STATE = -3,       /* a required setter has not been called yet */
    WFM_ERR_CUDA = -4,
    WFM_ERR_NOMEM = -5,
    WFM_ERR_INTERNAL = -6
} wfm_status;

typedef enum wfm_precision { WFM_F64 = 0, WFM_F32 = 1 } wfm_precision;

/* WFM:113-123  DEFOCUS=0, PHASE=1, MODULUS=2 */
typedef enum wfm_param { WFM_DEFOCUS = 0, WFM_PHASE = 1, WFM_MODULUS = 2 } wfm_param;

/* Jacobian selection bits for wfm_apply_jacobian_dev / wfm_apply_j_all */
#define WFM_J_DEFOCUS 1u
#define WFM_J_PHASE 2u
#define WFM_J_MODULUS 4u

/* Quirk Q1: the live fp64 apply_J_modulus keeps only the last plane
 * (WFM:662-675); the intended behaviour sums over z (WFM:710-726). */
typedef enum wfm_modulus_mode {
    WFM_MODULUS_INTENDED = 0,
    WFM_MODULUS_REFERENCE_LAST_PLANE = 1
} wfm_modulus_mode;

/* ---- life cycle -------------------------------------------------------------------- */

/* new WideFieldModel(shape, ...) state holder: WFM:154-172 + MM:62-78.
 * Whole stack on one device.  Fails with WFM_ERR_INVALID_ARG if nx != ny (WFM:158). */
WFM_API int wfm_create(wfm_model** out, int nx, int ny, int nz, double dxy, double dz,
                       int precision, int device);

/* z-slab shard of a global stack: planes [z0, z0+nz_local) of nz_global.
 * PSFnorm = 1/(Nx*Ny*nz_global) and the iz > Nz/2 wrap use the GLOBAL Nz (WFM:284,302). */
WFM_API int wfm_create_slab(wfm_model** out, int nx, int ny, int nz_global, int z0, int nz_local,
                            double dxy, double dz, int precision, int device);

/* Batch of `nbatch` INDEPENDENT models of the same shape on one handle (BASELINE config 5: parameter estimation
 * over many bead PSFs; SURVEY.md 8 b3 asks for a batch dimension rather than nbatch handles).  The models share
 * optics (wfm_set_optics), basis and support; each has its own phase / modulus / defocus vector, hence its own
 * rho, phi, psi, maskPupil.  One computePsf() / apply_J_*() pass runs the planes of ALL models through a single
 * pipeline launch.  Layouts gain a leading (slowest) model index: psf [nbatch][Nz][Ny][Nx], cpx likewise,
 * rho/phi/psi/mask [nbatch][Npix], gradients [nbatch][3 + nPhase + nModulus].
 * On a batch handle wfm_set_phase / _modulus / _defocus give every model the same vector; wfm_batch_set_* take one
 * row per model.  wfm_apply_j_* (single-model outputs), wfm_set_pupil_arrays, the rolled PSF, the MTF and
 * wfm_eval_fg return WFM_ERR_UNSUPPORTED / WFM_ERR_INVALID_ARG on a batch handle. */
WFM_API int wfm_create_batch(wfm_model** out, int nx, int ny, int nz, int nbatch, double dxy, double dz,
                             int precision, int device);
WFM_API int wfm_batch_size(const wfm_model* h);

/* The whole stack over n_dev GPUs of one box, driven by ONE host thread like the reference's callers
 * (PSF_Estimation.java:202-217): contiguous z-slabs, one per device (the first nz % n_dev devices hold one extra
 * plane).  The handle answers every entry point of a plain one with the reference layouts: setters are
 * broadcast, wfm_get_psf / wfm_get_cpx_psf copy every slab from its device straight to its offset of the caller's
 * array (one PCIe link per device, in parallel), wfm_apply_j_* split q the same way and add the per-device partial
 * vectors in device order.  Not available: wfm_set_stream, the rolled PSF / MTF / wfm_eval_fg (they cross slabs), the
 * batch calls; device-resident entry points go through the children (wfm_multi_part). */
WFM_API int wfm_create_multi(wfm_model** out, int nx, int ny, int nz, double dxy, double dz, int precision,
                             const int* devices, int n_dev);
WFM_API int wfm_multi_parts(const wfm_model* h);                 /* devices of a multi handle, 0 for a plain handle */
WFM_API int wfm_multi_part_info(const wfm_model* h, int part, int* device, int* z0, int* nz_local);
/* Borrowed child handle of one device, for device-resident use (wfm_device_psf, wfm_fill_uniform, wfm_synchronize).
 * It belongs to the parent: never destroy it and never call its setters. */
WFM_API int wfm_multi_part(wfm_model* h, int part, wfm_model** child);
/* Device-resident Jacobians of a multi handle.  q_dev[i]: part i's slab of q on device i.  grad_dev: 3 + nPhase +
 * nModulus doubles on the FIRST device, summed over the devices in device order: every device's k_jac_final stores its
 * partial vector into a slot on the first device over NVLink (peer mapping) and the first device adds the slots once
 * the others' events have fired -- no host round trip, no NCCL.  Asynchronous (first device's stream). */
WFM_API int wfm_multi_apply_jacobian_dev(wfm_model* h, unsigned kinds, const void* const* q_dev, double* grad_dev);

/* One process per GPU (torchrun): the same sum across PROCESSES, through CUDA IPC peer memory instead of an NCCL
 * all-reduce of ~14 doubles.  Every rank: wfm_exchange_export (allocates its landing buffer, returns the 64-byte IPC
 * handle) -> all-gather the handles by any means -> wfm_exchange_connect(rank, world, handles[world][64]).  From then
 * on wfm_apply_jacobian_dev leaves the sum over ALL ranks in grad_dev on every rank (fixed rank order: identical
 * bits everywhere): k_jac_final stores the partial vector into every peer's buffer over NVLink, raises a flag per
 * peer and adds the slots when the world's flags have arrived.  It is a collective: every rank issues the same
 * sequence of Jacobian calls.  wfm_exchange_status: 1 = connected and healthy, 0 = not connected, < 0 = a peer's
 * flag did not arrive (time-out, never a hang).  Close every rank's exchange before destroying the handles. */
#define WFM_EXCHANGE_HANDLE_BYTES 64
WFM_API int wfm_exchange_export(wfm_model* h, int world, void* handle_out);
WFM_API int wfm_exchange_connect(wfm_model* h, int rank, int world, const void* handles);
WFM_API int wfm_exchange_status(wfm_model* h);
WFM_API int wfm_exchange_close(wfm_model* h);

WFM_API int wfm_destroy(wfm_model* h);

/* Text of the last error on this handle (or of the last failed wfm_create* when h == NULL). */
WFM_API const char* wfm_last_error(const wfm_model* h);

/* Run the handle's work on a caller-owned CUDA stream (cudaStream_t passed as void*);
 * NULL restores the handle's own stream. */
WFM_API int wfm_set_stream(wfm_model* h, void* cuda_stream);
WFM_API int wfm_synchronize(wfm_model* h);
/* Device-side ordering against another stream (no host stall): wfm_wait_stream makes the handle's stream wait for
 * everything queued on `cuda_stream` so far (e.g. the producer of q); wfm_fence_stream makes `cuda_stream` wait for
 * everything queued on the handle's stream so far (e.g. before a collective on the gradient it just wrote). */
WFM_API int wfm_wait_stream(wfm_model* h, void* cuda_stream);
WFM_API int wfm_fence_stream(wfm_model* h, void* cuda_stream);

/* ---- pupil construction ------------------------------------------------------------- */

/* NA, lambda, ni of the constructor (WFM:161-166) followed by computeMaskPupil()
 * (WFM:1374-1406): sets mapPupil = maskPupil = disk of radius NA/lambda.  Marks the PSF dirty. */
WFM_API int wfm_set_optics(wfm_model* h, double NA, double lambda, double ni);

/* The (Gram-Schmidt orthonormalised) Zernike basis Z of computeZernike() (WFM:194-197):
 * nzern planes of Npix doubles.  `radial` selects the phase-mode offset (n+1 vs n+3,
 * WFM:1640-1644).  The library copies Z to the device. */
WFM_API int wfm_set_basis(wfm_model* h, const double* Z, int nzern, int radial);

/* Build the basis on the device instead: Zernike.zernikeArray (Zernike.java:119-288) +
 * in-order Gram-Schmidt (call site WFM:196), radius_px = NA/lambda*dxy*Nx (WFM:195). */
WFM_API int wfm_build_basis(wfm_model* h, int nzern, int radial);
WFM_API int wfm_get_basis(wfm_model* h, double* Z_out, int nzern);

/* setPhase(DoubleShapedVector) WFM:1625-1649.  n + offset must be <= nzern. */
WFM_API int wfm_set_phase(wfm_model* h, const double* alpha, int n);
/* setModulus(DoubleShapedVector) WFM:1588-1610. */
WFM_API int wfm_set_modulus(wfm_model* h, const double* beta, int n);
/* setDefocus(DoubleShapedVector) WFM:1510-1534 + computeDefocus() 1452-1499.
 * n == 3: {ni/lambda, deltaX, deltaY}; n == 1: {ni/lambda}; n == 2 is rejected
 * (it indexes element 2 of a length-2 vector in the reference, quirk Q4). */
WFM_API int wfm_set_defocus(wfm_model* h, const double* defoc, int n);

/* Batch handles: setPhase / setModulus / setDefocus of every model, table[nbatch][n] (row b = model b). */
WFM_API int wfm_batch_set_phase(wfm_model* h, const double* alpha_table, int n);
WFM_API int wfm_batch_set_modulus(wfm_model* h, const double* beta_table, int n);
WFM_API int wfm_batch_set_defocus(wfm_model* h, const double* defoc_table, int n);

/* Escape hatch for "identical synthetic pupils": load rho, phi, psi (Npix doubles each) and
 * maskPupil (Npix bytes) verbatim.  Any pointer may be NULL to keep the current array. */
WFM_API int wfm_set_pupil_arrays(wfm_model* h, const double* rho, const double* phi,
                                 const double* psi, const uint8_t* mask);

WFM_API int wfm_set_modulus_mode(wfm_model* h, int mode);

/* getRho/getPhi/getPsi/getMaskPupil WFM:1673,1713,1723,1784 (host copies). */
WFM_API int wfm_get_rho(wfm_model* h, double* out);
WFM_API int wfm_get_phi(wfm_model* h, double* out);
WFM_API int wfm_get_psi(wfm_model* h, double* out);
WFM_API int wfm_get_mask(wfm_model* h, uint8_t* out);

/* ---- the hot path ------------------------------------------------------------------- */

/* computePsf() WFM:206-396: no-op when the PSF is valid (PState > 0, WFM:207). */
WFM_API int wfm_compute_psf(wfm_model* h);
/* freeMem() WFM:1970-1974: PState = 0 (device buffers are kept for reuse). */
WFM_API int wfm_invalidate(wfm_model* h);
/* 1 when psf/cpxPsf are valid (PState), 0 when dirty. */
WFM_API int wfm_psf_state(const wfm_model* h);

/* getPsf() WFM:1798-1804 / get_cpxPsf() WFM:1856-1861: compute if dirty, then copy the
 * slab to host memory (double or float according to the handle's precision). */
WFM_API int wfm_get_psf(wfm_model* h, void* out_host);
WFM_API int wfm_get_cpx_psf(wfm_model* h, void* out_host);
/* getPsf() with the device->host copy queued on the handle's second stream: returns at once, so that the copy
 * overlaps the host->device copy of the next q (PCIe is full duplex).  out_host must be pinned
 * (wfm_host_alloc) and stay valid until wfm_wait_transfers() returns; the next computePsf() is ordered
 * after the copy.  wfm_wait_transfers blocks until every queued read-back has landed. */
WFM_API int wfm_get_psf_async(wfm_model* h, void* out_host);
WFM_API int wfm_wait_transfers(wfm_model* h);
/* "next" row f4.  ArrayUtils.roll(pupil.getPsf()) of BlindDeconvJob.java:100: the PSF with its origin moved from
 * voxel (0,0,0) to the centre of the volume, out[(i + n/2) mod n] = in[i] on every axis; host copy and
 * device-resident variant (whole stack on one handle).  getMtf() WFM:1807-1828 as intended (the reference loop
 * never terminates, quirk Q8): the unnormalised 3-D DFT of the PSF, interleaved complex (2,Nx,Ny,Nz); fp64,
 * Nz a power of two in [32, 2048]. */
WFM_API int wfm_get_psf_rolled(wfm_model* h, void* out_host);
WFM_API int wfm_roll_psf_dev(wfm_model* h, void* out_dev);
WFM_API int wfm_get_mtf(wfm_model* h, void* out_host);
/* Device-resident views of the same arrays (valid until the next setter / destroy). */
WFM_API int wfm_device_psf(wfm_model* h, void** dev_ptr);
WFM_API int wfm_device_cpx_psf(wfm_model* h, void** dev_ptr);

/* apply_J_phase WFM:738-1021, apply_J_defocus WFM:1029-1369, apply_J_modulus WFM:429-730.
 * q_host: gradient w.r.t. the PSF voxels of this slab, same shape/precision as psf.
 * out: n doubles; n must equal nPhase / 1 or 3 / nModulus.  If the PSF is dirty it is
 * recomputed first (quirk Q5: the reference would dereference a null cpxPsf). */
WFM_API int wfm_apply_j_phase(wfm_model* h, const void* q_host, double* out, int n);
WFM_API int wfm_apply_j_defocus(wfm_model* h, const void* q_host, double* out, int n);
WFM_API int wfm_apply_j_modulus(wfm_model* h, const void* q_host, double* out, int n);
/* apply_Jacobian(grad, xspace) WFM:399-409 with the space identified by its flag. */
WFM_API int wfm_apply_jacobian(wfm_model* h, int param, const void* q_host, double* out, int n);
/* One FFT pass for all three (they differ only after the FFT: WFM:928 vs 1253 vs 610). */
WFM_API int wfm_apply_j_all(wfm_model* h, const void* q_host, double* out_defocus3,
                            double* out_phase, double* out_modulus);

/* Batch handles: the selected Jacobians (WFM_J_* bits) of every model in one pass.  q_host [nbatch][Nz][Ny][Nx];
 * out [nbatch][3 + nPhase + nModulus] doubles, each row [defocus(3) | phase | modulus]. */
WFM_API int wfm_batch_apply_jacobian(wfm_model* h, unsigned kinds, const void* q_host, double* out);

/* Device-resident variant: q_dev and grad_dev are device pointers.  grad_dev receives
 * 3 + nPhase + nModulus doubles laid out [defocus(3) | phase | modulus]; entries of kinds not
 * selected are zero.  For a z-slab handle the values are this slab's PARTIAL sums, ready for
 * one sum-allreduce across ranks.  Asynchronous on the handle's stream.  On a batch handle grad_dev receives
 * nbatch such rows. */
WFM_API int wfm_apply_jacobian_dev(wfm_model* h, unsigned kinds, const void* q_dev, double* grad_dev);
WFM_API int wfm_grad_length(const wfm_model* h);

/* ---- "next" row f1: the FFT-convolution data term on the device ------------------------ */
/* TiPi mitiv.conv.WeightedConvolutionCost as PSF_Estimation drives it (PSF_Estimation.java:147-150,157,206):
 * the OBJECT is the kernel of the operator, the microscope PSF h is the variable,
 *     cost = alpha/2 * sum w * (obj (*) h - y)^2,   grad = alpha * corr(obj, w * (obj (*) h - y))
 * (periodic 3-D convolution at the data shape, offset {0,0,0}).  TiPi's source is not in the reference tree:
 * these semantics are restated, parity unpinned.  fp64, nx == ny, nx and nz powers of two in [32, 2048]. */
typedef struct wfm_conv wfm_conv;
WFM_API int wfm_conv_create(wfm_conv** out, int nx, int ny, int nz, int precision, int device);
/* The data term sharded by z-slab over n_dev GPUs (same device list and split as wfm_create_multi): real-space volumes
 * and the x / y passes in slabs, the z pass in pencils; the two transposes are done by the stores of the y pass and of
 * the fused z pass over NVLink peer memory (no separate all-to-all).  Answers wfm_conv_set_object / _set_data /
 * _set_weights / wfm_conv_cost_and_gradient (whole host volumes; every slab uses its own PCIe link) and wfm_eval_fg
 * together with a wfm_create_multi model: the inner loop of the PSF fit then scales over the box with only the
 * parameter vector crossing PCIe.  Needs peer access between all the devices. */
WFM_API int wfm_conv_create_multi(wfm_conv** out, int nx, int ny, int nz, int precision, const int* devices, int n_dev);
WFM_API int wfm_conv_parts(const wfm_conv* c);
WFM_API int wfm_conv_destroy(wfm_conv* c);
WFM_API const char* wfm_conv_last_error(const wfm_conv* c);
WFM_API int wfm_conv_set_stream(wfm_conv* c, void* cuda_stream);
WFM_API int wfm_conv_set_object(wfm_conv* c, const void* obj_host);     /* fdata.setPSF(obj, off)   :148 */
WFM_API int wfm_conv_set_data(wfm_conv* c, const void* data_host);      /* fdata.setData(data)      :149 */
WFM_API int wfm_conv_set_weights(wfm_conv* c, const void* w_host);      /* fdata.setWeights(w,true) :150; NULL = 1 */
/* fdata.computeCostAndGradient(alpha, psf, gcost, clr) :157,206 -- host buffers (synchronous) ... */
WFM_API int wfm_conv_cost_and_gradient(wfm_conv* c, double alpha, const void* h_host, void* grad_host, int clr,
                                       double* cost);
/* ... and device-resident (asynchronous on the handle's stream; cost_dev may be NULL). */
WFM_API int wfm_conv_cost_and_gradient_dev(wfm_conv* c, double alpha, const void* h_dev, void* grad_dev, int clr,
                                           double* cost_dev);
/* One COMPUTE_FG step of PSF_Estimation.fitPSF (PSF_Estimation.java:202-217) entirely on the device:
 * setParam(x) -> computePsf() -> computeCostAndGradient(alpha, psf, gcost, true) -> apply_Jacobian(gcost, space).
 * Only x (n doubles) crosses to the device; the cost and the n gradient doubles come back.  x == NULL: the caller
 * has already taken the setParam step through wfm_set_defocus / _phase / _modulus (what the host mirrors do, so
 * that their parameterCoefs stay in step); the chain then starts at computePsf(). */
WFM_API int wfm_eval_fg(wfm_model* h, wfm_conv* c, int param, const double* x, int n, double alpha, double* cost,
                        double* grad_out);

/* ---- utilities ---------------------------------------------------------------------- */

/* Counter-based splitmix64 uniform(-1,1) fill of a device array: element i
 * gets u(seed, first_index + i), written as double or float. */
WFM_API int wfm_fill_uniform(wfm_model* h, void* dev_ptr, int precision, uint64_t seed,
                             uint64_t first_index, uint64_t count);

/* Pinned host memory for the host<->device copies of the entry points above. */
WFM_API int wfm_host_alloc(void** out, size_t bytes);
WFM_API int wfm_host_free(void* p);

/* Introspection used by the benchmark and the tests. */
WFM_API int wfm_get_info(const wfm_model* h, int* nx, int* ny, int* nz_global, int* z0, int* nz_local,
                         int* precision, int* nzern, int* nphase, int* nmodulus);
/* Number of active pupil rows / columns the pruned FFT passes visit. */
WFM_API int wfm_active_extent(const wfm_model* h, int* n_active_x, int* n_active_y);
/* Per-kernel device timing with CUDA events on the handle's stream (off by default).  While on,
 * every launch group below is bracketed by two events; wfm_get_kernel_times synchronises and
 * returns the accumulated milliseconds and launch-group counts since profiling was switched on. */
#define WFM_K_PSF 0        /* k_psf_pipeline: pupil synthesis, row FFT, column FFT, conj(a) + |a|^2 store */
#define WFM_K_JAC 1        /* k_jac_pipeline: conj(a)*q load, row FFT, column FFT, masked trig products   */
#define WFM_K_JAC_REDUCE 2 /* k_jac_reduce + k_jac_final: sum over z, Zernike / defocus contractions     */
#define WFM_K_SETTERS 3    /* k_set_phase                                                                  */
#define WFM_KERNEL_IDS 4
WFM_API int wfm_set_profiling(wfm_model* h, int on);
WFM_API int wfm_get_kernel_times(wfm_model* h, double* ms_out, uint64_t* counts_out, int n);
/* Total kernel launches issued by this library since load (all handles). */
WFM_API uint64_t wfm_launch_count(void);
WFM_API const char* wfm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WFM_B200_H */
