/*
 * ps21.h -- C ABI of libps21.so, the B200 (sm_100a) implementation of 21CMMC's power-spectrum
 * likelihood hot path.
 *
 * The reference has no FFI on this path: the boundary it crosses is the Python call
 *   powerbox.tools.get_power(box, boxlength, bins, bin_ave=False, get_variance=False,
 *                            log_bins, k_weights)
 * made once per lightcone chunk / coeval box from
 *   Likelihood1DPowerLightcone.compute_power   src/py21cmmc/likelihood.py:734-800 (call :782-790)
 *   Likelihood1DPowerCoeval.compute_power      src/py21cmmc/likelihood.py:433-498 (call :480-488)
 * inside reduce_data (:833-865, :639-654), followed by the k^3/2pi^2 scaling (:863, :652) and
 * the chi^2 of computeLikelihood (:584-607).  The entry points below are what a ctypes binding in
 * those methods calls instead.
 *
 * Conventions: return 0 on success, a negative ps21 error code otherwise; ps21_last_error()
 * returns a thread-local message.  All pointers are plain host or device pointers as documented;
 * no torch types cross this boundary.  A plan is bound to one CUDA device, owns its workspace and
 * is not thread-safe.  Calls taking a `stream` are asynchronous on that CUDA stream (cudaStream_t
 * passed as void*; NULL = legacy default stream); the *_host variants synchronise before returning.
 */
#ifndef PS21_H
#define PS21_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS21_OK 0
#define PS21_ERR_INVALID (-1)     /* bad argument / unsupported configuration */
#define PS21_ERR_CUDA (-2)        /* a CUDA runtime call failed */
#define PS21_ERR_NOMEM (-3)       /* workspace allocation failed */
#define PS21_ERR_STATE (-4)       /* call order violated (e.g. lnl before set_data) */

#define PS21_MODE_SPHERICAL 0     /* 1-D |k| bins: get_power(..., res_ndim=None) */
#define PS21_MODE_CYLINDRICAL 1   /* (k_perp bin, k_par) : get_power(..., res_ndim=2) */

#define PS21_PATH_AUTO 0          /* fused kernels when the shape allows, else the generic two-pass ones, else general */
#define PS21_PATH_GENERAL 1       /* three-pass mixed-radix path (any size) */
#define PS21_PATH_FUSED 2         /* fail plan creation unless the fused path applies */
#define PS21_PATH_TWOPASS 3       /* generic two-pass kernels (any axes that factor into radices <= 32 / primes <= 13);
                                     fail plan creation when the shape is not supported */

typedef struct ps21_plan ps21_plan;

/*
 * Geometry + binning of one "box" (a lightcone that is cut into nchunks line-of-sight chunks, or a
 * coeval cube with nchunks = 1).  Everything floating point that decides a bin index is computed
 * by the caller with the very NumPy expressions powerbox uses and passed in explicitly
 * (SURVEY.md Appendix A steps 2, 4, 6); the device never re-derives an edge.
 */
typedef struct ps21_desc {
    int32_t n0, n1;               /* cells on the two sky-plane axes (HII_DIM) */
    int32_t n2_total;             /* cells on the last axis of every input box = row stride in floats */
    int32_t nchunks;              /* spectra per box; chunk c covers [chunk_start[c], chunk_end[c]) of axis 2
                                     (likelihood.py:836-852) */
    const int32_t* chunk_start;   /* [nchunks] */
    const int32_t* chunk_end;     /* [nchunks] */
    const double* volume;         /* [nchunks] V = L0*L1*L2_c in Mpc^3 (get_power's V) */
    const double* k0sq;           /* [n0] k_x^2, unshifted FFT order (np.fft.fftfreq order) */
    const double* k1sq;           /* [n1] k_y^2 */
    const double* k2sq;           /* [sum_c nz_c] k_z^2 of each chunk, concatenated in chunk order */
    int32_t nbins;                /* n_psbins */
    const double* edges;          /* [nchunks][nbins+1] bin edges (|k|, or k_perp when cylindrical) */
    const double* kcentres;       /* [nchunks][nbins] bin centres (likelihood.py:792-797); used for k^3/2pi^2 */
    int32_t ignore_kperp_zero;    /* likelihood.py:775-776 */
    int32_t ignore_kpar_zero;     /* likelihood.py:777-778 */
    int32_t ignore_k_zero;        /* likelihood.py:779-780 */
    int32_t mode;                 /* PS21_MODE_* */
    int32_t device;               /* CUDA device ordinal */
    int32_t path;                 /* PS21_PATH_* */
    int64_t workspace_bytes;      /* cap for the half-spectrum scratch (0 = default: 1/6 of device memory, <= 24 GiB);
                                     the scratch grows on demand up to the cap */
} ps21_desc;

typedef struct ps21_info {
    int32_t nshapes;              /* distinct chunk lengths */
    int32_t path_of_chunk0;       /* PS21_PATH_GENERAL / _FUSED / _TWOPASS actually chosen for chunk 0 */
    int64_t out_per_box;          /* doubles of power per box: sum_c nbins_out_c */
    int64_t workspace_bytes;      /* scratch currently allocated */
    int64_t alg_bytes_per_box;    /* algorithmic HBM bytes per box: sum_c 4N_c + 16 N_h,c */
    int64_t launches_last_call;   /* kernels launched by the last *_batched / *_host call */
    double  dominant_ms_last;     /* reserved */
} ps21_info;

const char* ps21_last_error(void);
int ps21_version(void);

int ps21_plan_create(const ps21_desc* desc, ps21_plan** plan_out);
int ps21_plan_destroy(ps21_plan* plan);
int ps21_plan_info(const ps21_plan* plan, ps21_info* info);

/* Per-kernel timing with CUDA events on the launching stream.  Kinds: 0 offset estimate, 1 first FFT pass
 * (z, or z + part of y), 2 middle pass (general path only), 3 last FFT pass fused with |delta_k|^2 +
 * histogram, 4 finalize (power, Delta^2, chi^2).  ps21_plan_profile synchronises, fills ms_by_kind[5] /
 * launches_by_kind[5] with the totals since the last query and resets them. */
int ps21_plan_set_profiling(ps21_plan* plan, int32_t on);
int ps21_plan_profile(ps21_plan* plan, double* ms_by_kind, int64_t* launches_by_kind);

/* CUDA graphs: with on != 0 the kernel sequence of a ps21_power_batched / ps21_lnl_batched call is captured once per
 * distinct argument set (buffers, batch) and replayed with a single cudaGraphLaunch afterwards (the launches run on a
 * stream owned by the plan, ordered after / before the caller's stream by events).  Callers that pass new buffers on
 * every call fall back to direct launches after 16 captures without a replay.  Off by default. */
int ps21_plan_set_graphs(ps21_plan* plan, int32_t on);
int ps21_plan_graph_stats(const ps21_plan* plan, int64_t* captures, int64_t* replays);

/* Frees everything that regrows on demand (half-spectrum scratch, partial sums, staging and result buffers, graphs);
 * the plan stays usable.  Synchronises the device. */
int ps21_plan_release_workspace(ps21_plan* plan);

/* offsets[c] = first double of chunk c inside one box's power record; offsets[nchunks] = out_per_box.
 * Spherical: nbins per chunk.  Cylindrical: nbins * (nz_c/2+1), row-major [bin][k_par index 0..nz_c/2]. */
int ps21_plan_layout(const ps21_plan* plan, int64_t* offsets);

/* Summed integer weights of every output bin (bincount(indx, weights=k_weights)[1:-1] of powerbox's
 * _get_binweights); host array of out_per_box int64.  Depends on geometry only. */
int ps21_plan_counts(const ps21_plan* plan, int64_t* counts);

/* Replaces the get_power loop of reduce_data for `batch` boxes resident on the device.
 * d_boxes: float32, box b at d_boxes + b*batch_stride, C-order (n0, n1, n2_total).
 * d_power: [batch][out_per_box] float64; empty bins are NaN (0/0), as in the reference. */
int ps21_power_batched(ps21_plan* plan, const float* d_boxes, int64_t batch, int64_t batch_stride,
                       double* d_power, void* stream);

/* Data side of computeLikelihood (likelihood.py:584-604), walker independent, HOST arrays.
 * nsets data sets; box b is compared with set (b % nsets) (coeval: one set per redshift).
 * data / noise: [nsets][nchunks][nbins]: data spline and noise spline evaluated at the bin centres
 * (noise may be NULL = 0).  kmask: [nsets][nchunks][nbins] 1 where min_k <= k <= max_k. */
int ps21_set_data(ps21_plan* plan, int32_t nsets, const double* data, const double* noise,
                  const uint8_t* kmask, double model_uncertainty, int32_t error_on_model);

/* reduce_data + computeLikelihood fused, spherical mode only.
 * d_lnl: [batch] float64 (device).  d_delta (may be NULL): [batch][nchunks][nbins] Delta^2 = P k^3/2pi^2.
 * d_power (may be NULL): [batch][nchunks][nbins]. */
int ps21_lnl_batched(ps21_plan* plan, const float* d_boxes, int64_t batch, int64_t batch_stride,
                     double* d_lnl, double* d_delta, double* d_power, void* stream);

/* Same two calls with HOST buffers (pageable or pinned): the library stages boxes through the
 * device in slices, overlapping H2D copies with kernels, and synchronises before returning. */
int ps21_power_host(ps21_plan* plan, const float* h_boxes, int64_t batch, int64_t batch_stride,
                    double* h_power);
int ps21_lnl_host(ps21_plan* plan, const float* h_boxes, int64_t batch, int64_t batch_stride,
                  double* h_lnl, double* h_delta, double* h_power);

/* The same with one HOST pointer per box (each a C-contiguous (n0, n1, n2_total) float32 array owned by its walker,
 * e.g. the LightCone.brightness_temp arrays of the contexts of one emcee step, likelihood.py:834-856): no gather into
 * one buffer on the host.  Pinned (ps21_host_alloc / ps21_host_register) buffers copy at the PCIe line rate. */
int ps21_power_host_ptrs(ps21_plan* plan, const float* const* h_box_ptrs, int64_t batch, double* h_power);
int ps21_lnl_host_ptrs(ps21_plan* plan, const float* const* h_box_ptrs, int64_t batch, double* h_lnl,
                       double* h_delta, double* h_power);

/* Page-locked host memory for walker buffers (thin wrappers over cudaHostAlloc / cudaHostRegister, so that the
 * Python host layer needs no CUDA binding of its own). */
int ps21_host_alloc(void** ptr, int64_t bytes);
int ps21_host_free(void* ptr);
int ps21_host_register(void* ptr, int64_t bytes);
int ps21_host_unregister(void* ptr);

/* computeLikelihood on already reduced models: h_delta [batch][nchunks][nbins] (host) -> h_lnl [batch]. */
int ps21_lnl_from_delta(ps21_plan* plan, const double* h_delta, int64_t batch, double* h_lnl);

/* Plan-free chi^2 of computeLikelihood (likelihood.py:601-604) for models reduced elsewhere: nrec rows of n
 * Delta^2 values against data/noise/mask of length n; the row is cut into nseg segments (one per chunk /
 * redshift, lengths seg_len[]) and lnl = sum_seg -0.5 * sum_k (...).  HOST arrays; noise may be NULL. */
int ps21_chi2_host(int32_t device, int64_t nrec, int64_t n, const double* delta, const double* data,
                   const double* noise, const uint8_t* mask, int64_t nseg, const int64_t* seg_len,
                   double model_uncertainty, int32_t error_on_model, double* lnl);

/* HERA-style consumer of the spectra: one (band, field) block of
 * Likelihood1DPowerLightconeUpper.computeLikelihood, likelihood.py:2286-2328.  For every one of nrec models:
 * after = wf[nk][nkwf] . model[nkwf];  err = sqrt(var + (frac*after)^2 + mean_err^2) where mean_err is the
 * window-weighted model error (model_err may be NULL);  L = 0.5 + 0.5 erf((limit - after)/(sqrt2 err)), values
 * <= 0 clamped to 1e-50;  lnl[r] += nansum(log L).  HOST arrays; lnl is accumulated into (caller zeroes it). */
int ps21_upper_limit_host(int32_t device, int64_t nrec, int64_t nk, int64_t nkwf, const double* model,
                          const double* wf, const double* limit, const double* var, const double* model_err,
                          double frac, double* lnl);

/* Same window-function matvec and error model, but for measured band powers instead of limits: the Gaussian
 * branch of Likelihood1DPowerCoeval.computeLikelihood, likelihood.py:517-583:
 * lnl[r] += -0.5 * sum_k (after - value)^2 / (var + (frac*after)^2 + mean_err^2)   (NaN propagates, np.sum). */
int ps21_windowed_gauss_host(int32_t device, int64_t nrec, int64_t nk, int64_t nkwf, const double* model,
                             const double* wf, const double* value, const double* var, const double* model_err,
                             double frac, double* lnl);

#ifdef __cplusplus
}
#endif
#endif /* PS21_H */
