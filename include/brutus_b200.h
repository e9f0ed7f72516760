/*
 * brutus_b200.h -- C ABI of libbrutus_b200.so: the B200 (sm_100a) implementation of the
 * brute-force photometric likelihood sweep of joshspeagle/brutus.
 *
 * The reference has no FFI: its seams are plain Python callables.  Each entry
 * point below names the reference interface it stands in for (paths relative to the reference
 * tree).  Plain pointers and sizes only; no exceptions cross the boundary; every call returns
 * an int status (0 = OK, negative = error, see BF_E_*), and bf_last_error() gives the message.
 *
 * Threading: a handle is NOT re-entrant (calls on one handle must be serialised); distinct handles are
 * independent.  A handle drives one CUDA device (bf_create) or several devices of this process
 * (bf_create_multi: stars are sharded over them inside the batch calls, one host thread and stream per
 * device); with one process per GPU (torchrun) each process holds a one-device handle that joins an NCCL
 * process group (bf_nccl_init).  Inputs are never modified.  All host buffers are caller-owned.
 */
#ifndef BRUTUS_B200_H
#define BRUTUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_OK 0
#define BF_E_INVALID (-1)   /* bad argument (message says which)                          */
#define BF_E_CUDA (-2)      /* CUDA runtime error                                          */
#define BF_E_NOGRID (-3)    /* sweep requested before bf_set_grid                           */
#define BF_E_CAPACITY (-4)  /* reserved */
#define BF_E_THRESH (-5)    /* init_thresh > ltol_subthresh (ValueError at brutus/fitting.py:691-693) */
#define BF_E_NOMEM (-6)

#define BF_MAX_FILT 16      /* bands per grid supported by the compiled kernels */

/* grid memory layouts accepted by bf_set_grid */
#define BF_LAYOUT_C 0       /* (Nmodel, Nfilt, 3) C order: what load_models returns, brutus/utils.py:588-605 */
#define BF_LAYOUT_F 1       /* same shape, Fortran order: what _fit builds, brutus/fitting.py:1964 */

/* arithmetic the kernels compute in */
#define BF_PRECISION_F32 0  /* throughput path */
#define BF_PRECISION_F64 1  /* verification path (the reference computes in float64) */

typedef struct bf_handle bf_handle;

/* Fit options: the keyword arguments of loglike (brutus/fitting.py:579-585) that reach the kernels,
 * plus wt_thresh of lnpost (brutus/fitting.py:824, :988-991). */
typedef struct bf_options {
    double avlim[2];        /* default (0, 20)      */
    double av_gauss[2];     /* default (0, 1e6)     */
    double rvlim[2];        /* default (1, 8)       */
    double rv_gauss[2];     /* default (3.32, 0.18) */
    double ltol;            /* default 3e-2         */
    double ltol_subthresh;  /* default 1e-2         */
    double init_thresh;     /* default 5e-3         */
    double wt_thresh;       /* default 1e-3 (bf_sweep_batch only) */
    double select_slack;    /* default 0.5 (bf_sweep_batch only): the sweep keeps as selection candidates the
                               models whose provisional lnprob is within ln(wt_thresh) - select_slack of the
                               running maximum.  Results do not depend on it: if the final max(lnprob) falls
                               more than select_slack below the provisional one the star is redone with every
                               model as a candidate (bf_stats.fallbacks) */
    int32_t dim_prior;      /* default 1            */
    int32_t max_iter;       /* cap on mag / flux loop iterations (reference: unbounded); 0 = 64 mag, 1000 flux;
                               stars stopped by it are counted in bf_stats.unconverged */
    int32_t apply_parallax_clip; /* 1: lnpost's rough parallax prior (fitting.py:976-980) is applied
                                    before thresholding; 0 mimics lnpost(parallax=None) */
    int32_t skip_d2h;       /* bf_sweep_batch: leave the records on the device (device-only timing) */
} bf_options;

/* Per-call statistics (device times from CUDA events on the handle's stream). */
typedef struct bf_stats {
    double ms_device;       /* first kernel -> last kernel of the call                */
    double ms_magfit;       /* sum over launches of the full-grid magnitude-fit sweep  */
    double ms_flux;         /* survivor flux-space refinement + final lnprob           */
    double ms_select;       /* candidate expansion / re-fit, threshold, compaction, record kernels */
    int64_t kernel_launches;
    int64_t magfit_launches;
    int64_t magfit_star_passes; /* stars x full-grid passes done by those launches         */
    int64_t resweeps;       /* stars whose speculated mag-iteration count was wrong    */
    int64_t candidates;     /* candidate records the sweep appended ((star, model) pairs that may matter)   */
    int64_t fallbacks;      /* stars redone with every model as a candidate (see select_slack)      */
    int64_t survivors;      /* total models that survived the cull (brutus/fitting.py:758-759)  */
    int64_t selected;       /* total models that passed wt_thresh                     */
    int64_t h2d_bytes, d2h_bytes;
    double ms_post;         /* bf_fit_batch: prior integration, evidence, resampling kernels */
    int64_t selected2;      /* bf_fit_batch: models that passed lnpost's second threshold     */
    int64_t clipped;        /* bf_fit_batch: stars whose second selection was cut to nsel_max */
    int64_t fixups;         /* records refined as likely survivors that the exact cull rejected (redone)   */
    int64_t flux_more_launches; /* flux iterations beyond the ones the sweep runs itself (whole-pool passes) */
    int64_t regroups;       /* star groups split because their candidate records overflowed the pool      */
    int64_t unconverged;    /* stars whose mag or flux loop was stopped by the iteration cap (max_iter)     */
    int64_t host_syncs;     /* times the host waited for the device inside the call (control read-backs)    */
} bf_stats;

void bf_default_options(bf_options* opt);

/* Lifetime.  device = CUDA ordinal; precision = BF_PRECISION_*. */
int bf_create(int device, int precision, bf_handle** out);
int bf_destroy(bf_handle* h);

/* ---- several GPUs -----------------------------------------------------
 * The reference fits objects strictly one after another (`for i in range(Ndata)`, brutus/fitting.py:1980) and no
 * object depends on another, so the catalogue is sharded into contiguous star ranges, the grid is replicated,
 * and nothing is exchanged in the hot loop.
 *
 * One process, n devices: bf_create_multi builds one engine per device and an NCCL communicator over them
 * (ncclCommInitAll).  bf_set_grid then copies the grid to the first device once and replicates it with ONE
 * ncclBroadcast; bf_sweep_batch / bf_fit_batch give device d the stars [d*N/n, (d+1)*N/n) and return one result
 * in catalogue order.  The posterior's random numbers are keyed by the catalogue index of a star, so results do
 * not depend on n.  bf_loglike_full and bf_get_seds run on the first device.  NCCL is loaded at run time
 * (dlopen "libnccl.so.2"): single-device handles do not need it. */
int bf_create_multi(const int* devices, int ndev, int precision, bf_handle** out);
int bf_num_devices(const bf_handle* h);

/* One process per GPU: rank 0 calls bf_nccl_unique_id and hands the 128 bytes to every rank (any channel: the
 * Python side uses a TCP socket on MASTER_ADDR); every rank then calls bf_nccl_init on its one-device handle
 * (ncclCommInitRank).  bf_set_grid_bcast: `coeffs` is read on rank `root` only (NULL elsewhere); one
 * ncclBroadcast, every rank re-tiles its copy.  bf_bcast_host replicates a small host buffer (labels, priors)
 * and bf_allreduce_max reduces n doubles in place (device timings are reported as the maximum over ranks; it is
 * also a barrier).  With a single rank these are no-ops. */
int bf_nccl_unique_id(void* out128);
int bf_nccl_init(bf_handle* h, const void* id128, int rank, int world);
int bf_set_grid_bcast(bf_handle* h, const float* coeffs, int64_t nmodel, int32_t nfilt, int32_t layout, int32_t root);
int bf_bcast_host(bf_handle* h, void* buf, int64_t bytes, int root);
int bf_allreduce_max(bf_handle* h, double* vals, int32_t n);
const char* bf_last_error(const bf_handle* h); /* h may be NULL: last error of a failed bf_create */

/* Stage the SED grid in HBM once.  Replaces BruteForce.__init__'s self.models (brutus/fitting.py:1139)
 * and the per-star copies `np.array(self.models, order='F')` / `mag_coeffs[:, mask, :]`
 * (:1964, :714).  coeffs: float32, nmodel*nfilt*3 values in `layout`; re-tiled on the device to
 * [coef][band][model].  bf_set_grid_device takes a DEVICE pointer (e.g. the target of an NCCL
 * broadcast) in BF_LAYOUT_C or _F. */
int bf_set_grid(bf_handle* h, const float* coeffs, int64_t nmodel, int32_t nfilt, int32_t layout);
int bf_set_grid_device(bf_handle* h, const void* d_coeffs, int64_t nmodel, int32_t nfilt, int32_t layout);

/* Optional model label columns used by lnprior_ext (brutus/fitting.py:1995-2009):
 * labels[l*nmodel + i], float64. */
int bf_set_labels(bf_handle* h, const double* labels, int32_t nlabel);

/* av_init / rv_init of loglike (brutus/fitting.py:583, :700-703): where the magnitude fit of each model starts
 * (float64 [nmodel] each).  Staged once; in effect for every later bf_loglike_full on the handle until cleared with
 * (NULL, NULL) or until the grid is replaced.  The default -- the prior means av_gauss[0], rv_gauss[0] -- is what
 * BruteForce.fit always uses (:1983-1992), so the batch calls ignore these arrays. */
int bf_set_init(bf_handle* h, const double* av_init, const double* rv_init);

/* B1: one star, full-length outputs -- the contract of
 *   loglike(data, data_err, data_mask, mag_coeffs, ..., return_vals=True)  (brutus/fitting.py:579-820)
 * flux/err: nfilt float64; mask: nfilt uint8 (NOT modified; mask_clean_out receives the cleaned mask
 * the reference writes in place at :709).  parallax/parallax_err: NaN = not provided (:750-751).
 * Outputs (caller-allocated float64): lnl, chi2, scale, av, rv [nmodel]; icov [nmodel*9] or NULL.
 * diag (may be NULL): [0]=Ndim [1]=mag iterations [2]=flux iterations [3]=survivors of the cull. */
int bf_loglike_full(bf_handle* h, const double* flux, const double* err, const uint8_t* mask,
                    double parallax, double parallax_err, const bf_options* opt,
                    double* lnl, double* chi2, double* scale, double* av, double* rv, double* icov,
                    uint8_t* mask_clean_out, int64_t* diag);

/* Compacted records of the selected models, in pinned host memory OWNED BY THE LIBRARY: valid until
 * the next bf_sweep_batch on the handle or bf_destroy.  rows is a [nrows][stride] matrix of
 * float32 (BF_PRECISION_F32 handles) or float64 (BF_PRECISION_F64): row 0 lnl (incl. lnprior_ext),
 * 1 scale, 2 av, 3 chi2, 4 rv, 5..10 the unique entries (ss, sa, sr, aa, ar, rr) of the symmetric
 * precision matrix icov_sar (brutus/fitting.py:563-574).  Only the first nrows rows are filled. */
typedef struct bf_records {
    int64_t n;                /* records delivered                        */
    int64_t stride;           /* elements between consecutive rows         */
    int32_t elem_size;        /* 4 or 8                                    */
    int32_t nrows;            /* BF_REC_BASIC, BF_REC_FIT or BF_REC_FULL   */
    const int32_t* model_idx; /* [n], ascending within each star           */
    const void* rows;
} bf_records;
#define BF_REC_BASIC 3   /* lnl, scale, av        */
#define BF_REC_FIT 5     /* + chi2, rv            */
#define BF_REC_FULL 11   /* + icov_sar            */

/* B2: many stars, compacted outputs -- the per-star body of BruteForce._fit
 * (brutus/fitting.py:1980-2009: loglike + lnprior_ext) fused with lnpost's first stage (:976-991:
 * rough parallax prior, -1e300 clean-up, selection lnprob > max + ln wt_thresh).
 *   flux, err      [nstar*nfilt] float64;  mask [nstar*nfilt] uint8
 *   parallax, parallax_err [nstar] float64 (NaN = none); either may be NULL (= all NaN)
 *   ext_mean, ext_std [nstar*nlabel] float64 or NULL (lnprior_ext; needs bf_set_labels)
 *   record_rows    BF_REC_*: how much of each record to compute and ship
 * per-star outputs (each may be NULL except offsets): ndim [nstar] int32, n_iter [nstar*2] int32
 *   (mag, flux loop iterations), n_surv [nstar] int64 (survivors of the cull), max_lnprob [nstar]
 *   float64, offsets [nstar+1] int64: records of star s are [offsets[s], offsets[s+1]) of *out.
 * Stars are processed in batches; the device->host copy of one batch's records (copy stream, pinned
 * memory) overlaps the next batch's kernels. */
int bf_sweep_batch(bf_handle* h, int64_t nstar, const double* flux, const double* err,
                   const uint8_t* mask, const double* parallax, const double* parallax_err,
                   const double* ext_mean, const double* ext_std, const bf_options* opt,
                   int32_t record_rows, int32_t* ndim, int32_t* n_iter, int64_t* n_surv,
                   double* max_lnprob, int64_t* offsets, bf_records* out);

/* ---- device-side posterior: lnpost after its first selection, evidence, resampling --------------------
 * (brutus/fitting.py:999-1107 and :2012-2061) */

/* Parameters of the default Galactic prior, gal_lnprior (brutus/pdf.py:476-486), plus the two frame
 * constants astropy's Galactocentric frame supplies at brutus/pdf.py:630-635. */
typedef struct bf_gal_params {
    double R_solar, Z_solar, R_thin, Z_thin, Rs_thin, R_thick, Z_thick, f_thick, Rs_thick;
    double Rs_halo, q_halo_ctr, q_halo_inf, r_q_halo, eta_halo, f_halo;
    double feh_thin, feh_thin_sigma, feh_thick, feh_thick_sigma, feh_halo, feh_halo_sigma;
    double max_age, min_age, feh_age_ctr, feh_age_scale, nsigma_from_max_age, max_sigma, min_sigma;
    double galcen_distance, z_sun;   /* 8.122 kpc, 0.0208 kpc */
} bf_gal_params;
void bf_default_gal_params(bf_gal_params* g);

/* Keyword arguments of lnpost / _fit that reach the posterior kernels. */
typedef struct bf_post_options {
    int32_t nmc_prior;      /* Nmc_prior (fit() default 50); must be >= 1                          */
    int32_t ndraws;         /* Ndraws, default 250                                                 */
    uint64_t seed;          /* keys the counter-based generator (stands in for `rstate`)            */
    int32_t use_gal_prior;  /* 1: built-in gal_lnprior with `gal`; 0: flat distance prior           */
    int32_t reserved;
    int64_t star_base;      /* catalogue index of the first star of this call: the generator is keyed by
                               (seed, star_base + s, model, draw), so a batched or star-sharded caller gets
                               the same numbers as one call over the whole catalogue */
    int64_t nsel_max;       /* lnpost's memory clip (brutus/fitting.py:969-970, :1029-1036): when a star's second
                               selection holds more than nsel_max = int(mem_lim / Nmc_prior / 4e-4) models, only
                               the nsel_max with the largest lnlike + lnprior are kept; 0 = no clip */
    bf_gal_params gal;
    /* test hooks (NULL in production): host-supplied random numbers so that the device result can be
     * compared draw for draw with a NumPy restatement.
     *   z_override [nmodel][3][nmc_prior] standard normals, used for every star (the reference draws
     *              z.reshape(Nsel, 3, Nmc), brutus/utils.py:897)
     *   u_override [nstar][2][ndraws] uniforms: [0] the model draw (:2040), [1] the MC pick (:2052) */
    const double* z_override;
    const double* u_override;
} bf_post_options;
void bf_default_post_options(bf_post_options* o);

/* Static per-model inputs of lnpost, staged once: lnprior (the `lnprior` grid of fit(),
 * brutus/fitting.py:1004), and the label columns 'feh' / 'loga' the Galactic prior uses
 * (brutus/pdf.py:669, :694).  Each is float64 [nmodel] or NULL (0 / label absent). */
int bf_set_model_priors(bf_handle* h, const double* lnprior, const double* feh, const double* loga);

/* Posterior draws of every star, [nstar*ndraws] each (cov_sar: [nstar*ndraws*9]): the 13-tuple
 * BruteForce._fit yields per object (brutus/fitting.py:2059-2061).  Like bf_records, the arrays live in pinned
 * host memory OWNED BY THE LIBRARY (the device writes them over PCIe at full link speed, no staging copy):
 * bf_fit_batch fills in the pointers; they stay valid until the next bf_fit_batch on the handle or bf_destroy. */
typedef struct bf_draws {
    const int32_t* model_idx;   /* sidxs  (-99 where a star has no selected model) */
    const double *scale, *av, *rv, *cov_sar, *lnprob, *dist, *red, *dred, *logwt;
} bf_draws;

/* The per-star body of BruteForce._fit end to end on the device (brutus/fitting.py:1980-2061):
 * bf_sweep_batch's work, then lnpost (priors at the MLE, second threshold, covariances, Monte Carlo
 * integration over the Galactic and parallax priors), the evidence, chi2min and the resampling.
 * Only ndraws samples per star cross PCIe.  Not applied: the 3-D dust prior (no map bundled: flat
 * A(V) prior, as fit(dustfile=None) :1396-1398).
 *   coords [nstar*2] float64 Galactic (l, b) in degrees, or NULL with use_gal_prior = 0
 * outputs: ndim [nstar] (incl. +1 for a finite parallax, :2030), n_iter [nstar*2], nsel [nstar] (size of
 * the second selection), levid, chi2min [nstar]; any of ndim / n_iter / nsel may be NULL. */
int bf_fit_batch(bf_handle* h, int64_t nstar, const double* flux, const double* err, const uint8_t* mask,
                 const double* parallax, const double* parallax_err, const double* coords,
                 const double* ext_mean, const double* ext_std, const bf_options* opt,
                 const bf_post_options* post, int32_t* ndim, int32_t* n_iter, int64_t* nsel,
                 double* levid, double* chi2min, bf_draws* out);

/* Reddened SEDs of (model, Av, Rv) samples from the staged grid -- get_seds / _get_seds
 * (brutus/utils.py:1089-1159, :286-347), the grid-touching part of photometric_offsets (:1268-1271).
 *   n       number of samples
 *   idx     [n] model index of each sample, or NULL for the identity (n == nmodel: the whole grid)
 *   av, rv  [n] float64
 *   return_flux  0: magnitudes; 1: flux densities 10^(-0.4 m), reddening vectors scaled by -0.4 ln10 F
 * outputs (caller-allocated float64, [n*nfilt], C order; rvecs / drvecs may be NULL). */
int bf_get_seds(bf_handle* h, int64_t n, const int32_t* idx, const double* av, const double* rv,
                int32_t return_flux, double* seds, double* rvecs, double* drvecs);

/* The device part of photometric_offsets (brutus/utils.py:1225-1400) for nobj objects with nsamps posterior samples
 * each, on the first device of the handle:
 *   seds [nobj*nsamps*nfilt]   flux-density SEDs of the samples, get_seds(models[idxs], av=reds, rv=dreds,
 *                              return_flux=True) / dists^2 (:1268-1271)
 *   wt   [nfilt][nobj*nsamps]  for every band b with mask_fit[b] != 0: exp(lnl - logsumexp(lnl)) over the samples of
 *                              an object, lnl = phot_loglike(phot * old_offsets, err * old_offsets, mask without
 *                              band b, seds, dim_prior) (:1162-1222, :1299-1309); zeros for the other bands
 * phot, err [nobj*nfilt] float64; mask [nobj*nfilt], mask_fit [nfilt] uint8; idxs int32, reds, dreds, dists float64
 * [nobj*nsamps]; old_offsets [nfilt] or NULL (ones).  What is left to the caller is the bootstrap (:1311-1333), which
 * consumes the caller's random state. */
int bf_offsets_weights(bf_handle* h, int64_t nobj, int32_t nsamps, const double* phot, const double* err,
                       const uint8_t* mask, const int32_t* idxs, const double* reds, const double* dreds,
                       const double* dists, const double* old_offsets, const uint8_t* mask_fit, int32_t dim_prior,
                       double* seds, double* wt);

/* Statistics of the most recent bf_loglike_full / bf_sweep_batch / bf_fit_batch call on this handle. */
int bf_get_stats(const bf_handle* h, bf_stats* out);

/* Per-kernel device times (CUDA events around every launch) accumulated since the last bf_get_trace on this
 * handle, one "name launches ms" line per kernel; empty unless the handle was created with BRUTUS_B200_TRACE=1
 * in the environment.  The reference's only instrumentation is a wall-clock mean per object
 * (brutus/fitting.py:1717-1731).  The string is owned by the handle and valid until the next call. */
const char* bf_get_trace(bf_handle* h);

/* Benchmark hygiene: overwrite a 512 MB scratch buffer so that nothing of the previous step stays
 * in the 126 MB L2 (B200_PROFILING.md, "Timing hygiene"). */
int bf_flush_l2(bf_handle* h);

/* Build / device introspection. */
int bf_device_count(void);
const char* bf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BRUTUS_B200_H */
