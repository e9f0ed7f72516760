/*
 * oracle/loglike_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, float64, scalar-loop restatement of the reference's brute-force photometric
 * likelihood path, used (a) as the checker for the CUDA kernels in tests/ and smoke(), and
 * (b) as the CPU baseline timed by bench.py.  Nothing under brutus_b200/ may link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against golden vectors that
 * tests/gen_golden.py produced by running the unmodified reference (numba/NumPy) in the build
 * container; tests/test_oracle.py::test_live_reference re-runs the reference when it is present.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * The arithmetic order of the reference is kept (divisions by the variance, pow(10, x), the
 * same accumulation order over bands) so that agreement is at the 1e-12 level.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double avlim[2];        /* brutus/fitting.py:580 */
    double av_gauss[2];     /* brutus/fitting.py:580 */
    double rvlim[2];        /* brutus/fitting.py:581 */
    double rv_gauss[2];     /* brutus/fitting.py:581 */
    double ltol;            /* brutus/fitting.py:583 */
    double ltol_subthresh;  /* brutus/fitting.py:583 */
    double init_thresh;     /* brutus/fitting.py:583 */
    int32_t dim_prior;      /* brutus/fitting.py:583 */
    int32_t max_iter;       /* safety cap (the reference loops are unbounded); 0 = 1000 */
} ref_options;

/* brutus/utils.py:286-347 (_get_seds) for one model; coeff = (Nb,3) float32, promoted to f64 */
static void get_seds_one(const float *coeff, int nb, double av, double rv, int return_flux,
                         double *sed, double *rvec, double *drvec) {
    const double fac = -0.4 * log(10.);
    for (int j = 0; j < nb; j++) {
        double mags = coeff[3 * j + 0], r0 = coeff[3 * j + 1], dr = coeff[3 * j + 2];
        drvec[j] = dr;
        rvec[j] = r0 + rv * dr;
        sed[j] = mags + av * rvec[j];
        if (return_flux) {
            sed[j] = pow(10., -0.4 * sed[j]);
            rvec[j] *= fac * sed[j];
            drvec[j] *= fac * sed[j];
        }
    }
}

/* brutus/fitting.py:430-576 (_get_sed_mle) for one model.  icov is 3x3 row-major. */
static void get_sed_mle_one(const double *data, const double *tot_var, const float *coeff, int nb,
                            double av, double rv, const ref_options *o, double *models,
                            double *rvecs, double *drvecs, double *scale_out, double *icov,
                            double *resid) {
    const double av_reg = 0.05, rv_reg = 0.1; /* :433 */
    get_seds_one(coeff, nb, av, rv, 1, models, rvecs, drvecs); /* :506 */
    double s_num = 0., s_den = 0.;
    for (int j = 0; j < nb; j++) { /* :512-515 */
        s_num += models[j] * data[j] / tot_var[j];
        s_den += models[j] * models[j] / tot_var[j];
    }
    double scale = s_num / s_den; /* :516 */
    if (scale <= 1e-20) scale = 1e-20; /* :517-518 */
    double sr_mix = 0., sa_mix = 0., a_den = 0., r_den = 0., ar_mix = 0.;
    const double Av_varinv = 1. / (o->av_gauss[1] * o->av_gauss[1]);
    const double Rv_varinv = 1. / (o->rv_gauss[1] * o->rv_gauss[1]);
    const double a_den_reg = 1. / (av_reg * av_reg), r_den_reg = 1. / (rv_reg * rv_reg);
    for (int j = 0; j < nb; j++) { /* :527-553 */
        double models_int = pow(10., -0.4 * (double)coeff[3 * j + 0]);
        double reddening = models[j] - models_int;
        models[j] = models[j] * scale;
        resid[j] = data[j] - models[j];
        sr_mix += drvecs[j] * ((models[j] - resid[j]) / tot_var[j]);
        sa_mix += rvecs[j] * ((models[j] - resid[j]) / tot_var[j]);
        rvecs[j] = rvecs[j] * scale;
        drvecs[j] = drvecs[j] * scale;
        reddening *= scale;
        ar_mix += drvecs[j] * ((reddening - resid[j]) / tot_var[j]);
        a_den += rvecs[j] * rvecs[j] / tot_var[j];
        r_den += drvecs[j] * drvecs[j] / tot_var[j];
    }
    a_den += Av_varinv; /* :556-561 */
    r_den += Rv_varinv;
    a_den += a_den_reg;
    r_den += r_den_reg;
    icov[0] = s_den; icov[4] = a_den; icov[8] = r_den; /* :566-574 */
    icov[1] = icov[3] = sa_mix;
    icov[2] = icov[6] = sr_mix;
    icov[5] = icov[7] = ar_mix;
    *scale_out = scale;
}

/* brutus/utils.py:130-176 (_chisquare_logpdf, loc=0, scale=1) */
static double chisquare_logpdf(double x, double df) {
    if (x <= 0.) return -INFINITY;
    double ans = -log(pow(2., df / 2.) * tgamma(df / 2.));
    return ans + (df / 2. - 1.) * log(x) - x / 2.;
}

/*
 * brutus/fitting.py:579-820 (loglike, return_vals=True) for ONE star.
 *
 *   data, err   (nfilt)   float64 flux densities and errors
 *   mask_io     (nfilt)   uint8; cleaned in place like the reference does (:708-709)
 *   coeffs      (nmodel, nfilt, 3) float32, C order
 *   parallax, parallax_err: NaN = not provided (:750-751)
 *   outputs lnl, chi2, scale, av, rv (nmodel), icov (nmodel*9), all float64
 *   diag[0]=Ndim, diag[1]=mag-loop iterations, diag[2]=flux-loop iterations, diag[3]=survivors
 *   surv_out: optional (nmodel) uint8 flag, 1 where the model survived the cull (:758-759)
 *   lnlp_out: optional (nmodel) cull statistic lnl_p (:747-756); the survivors are lnl_p > max + ln(init_thresh)
 *             (tests: float32 kernels may flip models within rounding of that threshold)
 * returns 0, or -1 on the ValueError of :691-693, -2 on allocation failure.
 */
static int loglike_impl(const double *data, const double *err, uint8_t *mask_io, int nfilt,
                        const float *coeffs, int64_t nmodel, const ref_options *o, double parallax,
                        double parallax_err, const double *av_init, const double *rv_init, double *lnl,
                        double *chi2, double *scale, double *av, double *rv, double *icov, int64_t *diag,
                        uint8_t *surv_out, double *lnlp_out) {
    if (o->init_thresh > o->ltol_subthresh) return -1; /* :691-693 */
    const int max_iter = o->max_iter > 0 ? o->max_iter : 1000;

    /* clean data (:708-710) */
    int nb = 0;
    int band[64];
    if (nfilt > 64) return -3;
    for (int j = 0; j < nfilt; j++) {
        int clean = isfinite(data[j]) && isfinite(err[j]) && (err[j] > 0.);
        if (!clean) mask_io[j] = 0;
        if (mask_io[j]) band[nb++] = j;
    }
    const int Ndim = nb;

    /* sub-select clean observations (:713-716) */
    double flux[64], tot_var[64], mags[64], mags_var[64];
    for (int k = 0; k < nb; k++) {
        flux[k] = data[band[k]];
        tot_var[k] = err[band[k]] * err[band[k]];
    }
    /* magnitudes (:722-725) */
    {
        const double c2 = (2.5 / log(10.)) * (2.5 / log(10.));
        for (int k = 0; k < nb; k++) {
            mags[k] = -2.5 * log10(flux[k]);
            mags_var[k] = c2 * tot_var[k] / (flux[k] * flux[k]);
            if (!isfinite(mags[k])) { mags[k] = 0.; mags_var[k] = 1e50; }
        }
    }

    /* working arrays, (nmodel, nb) like the reference */
    size_t nn = (size_t)nmodel * (size_t)(nb > 0 ? nb : 1);
    float *mco = (float *)malloc(sizeof(float) * nn * 3);
    double *resid = (double *)malloc(sizeof(double) * nn);
    double *rvecs = (double *)malloc(sizeof(double) * nn);
    double *drvecs = (double *)malloc(sizeof(double) * nn);
    double *wk = (double *)malloc(sizeof(double) * (size_t)nmodel * 6);
    if (!mco || !resid || !rvecs || !drvecs || !wk) {
        free(mco); free(resid); free(rvecs); free(drvecs); free(wk);
        return -2;
    }
    double *dav = wk, *drv = wk + nmodel, *logwt = wk + 2 * nmodel;
    double *s_den = wk + 3 * nmodel, *rp_den = wk + 4 * nmodel, *srp_mix = wk + 5 * nmodel;

    /* mcoeffs = mag_coeffs[:, mask, :] (:714) */
    for (int64_t i = 0; i < nmodel; i++)
        for (int k = 0; k < nb; k++)
            for (int c = 0; c < 3; c++)
                mco[((size_t)i * nb + k) * 3 + c] = coeffs[((size_t)i * nfilt + band[k]) * 3 + c];

    /* initial magnitudes at (av_init, rv_init), by default the prior means (:700-703, :728-733) */
    for (int64_t i = 0; i < nmodel; i++) {
        av[i] = av_init ? av_init[i] : 0. + o->av_gauss[0];
        rv[i] = rv_init ? rv_init[i] : 0. + o->rv_gauss[0];
        double sed[64];
        get_seds_one(mco + (size_t)i * nb * 3, nb, av[i], rv[i], 0, sed, rvecs + (size_t)i * nb,
                     drvecs + (size_t)i * nb);
        for (int k = 0; k < nb; k++) resid[(size_t)i * nb + k] = mags[k] - sed[k];
    }

    /* ---- _optimize_fit_mag (brutus/fitting.py:34-271), stepsize = 1 (:734) ---- */
    const double mtol = 2.5 * o->ltol; /* :732 */
    const double avmin = o->avlim[0], avmax = o->avlim[1];
    const double rvmin = o->rvlim[0], rvmax = o->rvlim[1];
    const double Av_mean = o->av_gauss[0], Rv_mean = o->rv_gauss[0];
    const double Av_varinv = 1. / (o->av_gauss[1] * o->av_gauss[1]);
    const double Rv_varinv = 1. / (o->rv_gauss[1] * o->rv_gauss[1]);
    const double log_init_thresh = log(o->init_thresh);
    for (int64_t i = 0; i < nmodel; i++) { /* :158-164 */
        double sd = 0., rp = 0., sm = 0.;
        const double *dr = drvecs + (size_t)i * nb;
        for (int k = 0; k < nb; k++) {
            sd += 1. / mags_var[k];
            rp += dr[k] * dr[k] / mags_var[k];
            sm += dr[k] / mags_var[k];
        }
        s_den[i] = sd; rp_den[i] = rp; srp_mix[i] = sm;
    }
    int n_mag = 0;
    while (1) { /* :173 */
        n_mag++;
        for (int64_t i = 0; i < nmodel; i++) {
            double *rs = resid + (size_t)i * nb, *rvv = rvecs + (size_t)i * nb;
            const double *dr = drvecs + (size_t)i * nb;
            const double stepsize = 1.;
            /* solve for Av (:176-204) */
            double a_den = 0., sa_mix = 0., resid_s = 0., resid_a = 0.;
            for (int k = 0; k < nb; k++) {
                a_den += rvv[k] * rvv[k] / mags_var[k];
                sa_mix += rvv[k] / mags_var[k];
                resid_s += rs[k] / mags_var[k];
                resid_a += rs[k] * rvv[k] / mags_var[k];
            }
            resid_a += (Av_mean - av[i]) * Av_varinv;
            a_den += Av_varinv;
            double sa_idet = 1. / (s_den[i] * a_den - sa_mix * sa_mix);
            double d = sa_idet * (s_den[i] * resid_a - sa_mix * resid_s);
            d = d * stepsize;
            if (d < avmin - av[i]) d = avmin - av[i];
            if (d > avmax - av[i]) d = avmax - av[i];
            dav[i] = d;
            av[i] = av[i] + d;
            for (int k = 0; k < nb; k++) rs[k] = rs[k] - d * rvv[k];
            /* solve for Rv (:206-237) */
            double resid_r = 0.;
            resid_s = 0.;
            double r_den = rp_den[i] * av[i] * av[i];
            double sr_mix = srp_mix[i] * av[i];
            for (int k = 0; k < nb; k++) {
                resid_s += rs[k] / mags_var[k];
                resid_r += rs[k] * dr[k] / mags_var[k];
            }
            resid_r = resid_r * av[i];
            resid_r += (Rv_mean - rv[i]) * Rv_varinv;
            r_den += Rv_varinv;
            double sr_idet = 1. / (s_den[i] * r_den - sr_mix * sr_mix);
            double e = sr_idet * (s_den[i] * resid_r - sr_mix * resid_s);
            e = e * stepsize;
            if (e < rvmin - rv[i]) e = rvmin - rv[i];
            if (e > rvmax - rv[i]) e = rvmax - rv[i];
            drv[i] = e;
            rv[i] = rv[i] + e;
            for (int k = 0; k < nb; k++) {
                rs[k] = rs[k] - av[i] * e * dr[k];
                rvv[k] = rvv[k] + e * dr[k];
            }
            /* :240-243 */
            double c2 = 0.;
            for (int k = 0; k < nb; k++) c2 += rs[k] * rs[k] / mags_var[k];
            logwt[i] = -0.5 * c2;
        }
        double max_logwt = -1e300; /* :246-249 */
        for (int64_t i = 0; i < nmodel; i++)
            if (logwt[i] > max_logwt) max_logwt = logwt[i];
        double errv = -1e300; /* :252-260 */
        for (int64_t i = 0; i < nmodel; i++) {
            if (logwt[i] > max_logwt + log_init_thresh) {
                double a = fabs(dav[i]), b = fabs(drv[i]);
                if (a > errv) errv = a;
                if (b > errv) errv = b;
            }
        }
        if (errv < mtol) break; /* :263 */
        if (n_mag >= max_iter) break;
    }

    /* _get_sed_mle for every model (:267), then the cull statistics (:745-756) */
    double *lnl_p = logwt; /* reuse */
    int have_par = isfinite(parallax) && isfinite(parallax_err);
    {
        double models[64];
        for (int64_t i = 0; i < nmodel; i++) {
            double *rs = resid + (size_t)i * nb;
            get_sed_mle_one(flux, tot_var, mco + (size_t)i * nb * 3, nb, av[i], rv[i], o, models,
                            rvecs + (size_t)i * nb, drvecs + (size_t)i * nb, &scale[i],
                            icov + (size_t)i * 9, rs);
            double c2 = 0.;
            for (int k = 0; k < nb; k++) c2 += rs[k] * rs[k] / tot_var[k];
            chi2[i] = c2;
            lnl[i] = -0.5 * c2;
            lnl_p[i] = lnl[i];
            if (have_par) {
                double par = sqrt(scale[i]);
                double chi2_p = (par - parallax) * (par - parallax) / (parallax_err * parallax_err);
                lnl_p[i] = lnl[i] - 0.5 * chi2_p;
            }
        }
    }
    /* threshold (:758-759) */
    double lmax = -INFINITY;
    for (int64_t i = 0; i < nmodel; i++)
        if (lnl_p[i] > lmax) lmax = lnl_p[i];
    const double lthr = lmax + log(o->init_thresh);
    int64_t nsel = 0;
    int64_t *sel = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nmodel > 0 ? nmodel : 1));
    if (!sel) { free(mco); free(resid); free(rvecs); free(drvecs); free(wk); return -2; }
    for (int64_t i = 0; i < nmodel; i++) {
        int keep = lnl_p[i] > lthr;
        if (keep) sel[nsel++] = i;
        if (surv_out) surv_out[i] = (uint8_t)keep;
        if (lnlp_out) lnlp_out[i] = lnl_p[i];
    }

    /* ---- flux-space iteration on the survivors (:778-803) ---- */
    double *step = dav, *lnl_old = drv; /* reuse (indexed by survivor slot) */
    double *lnl_new = s_den, *chi2_new = rp_den, *scale_new = srp_mix;
    double *icov_new = (double *)malloc(sizeof(double) * 9 * (size_t)(nsel > 0 ? nsel : 1));
    if (!icov_new) { free(sel); free(mco); free(resid); free(rvecs); free(drvecs); free(wk); return -2; }
    for (int64_t q = 0; q < nsel; q++) { step[q] = 1.; lnl_old[q] = -1e300; }
    const double rescaling = 1.2;
    const double ln_sub = log(o->ltol_subthresh);
    double lerr = 1e300;
    int n_flux = 0;
    while (lerr > o->ltol && nsel > 0) {
        n_flux++;
        /* _optimize_fit_flux (brutus/fitting.py:274-427) */
        for (int64_t q = 0; q < nsel; q++) {
            int64_t i = sel[q];
            double *rs = resid + (size_t)i * nb, *rvv = rvecs + (size_t)i * nb;
            double *dr = drvecs + (size_t)i * nb;
            double a_num = 0., a_den = 0., r_num = 0., r_den = 0.;
            for (int k = 0; k < nb; k++) { /* :387-389 */
                a_num += rvv[k] * rs[k] / tot_var[k];
                a_den += rvv[k] * rvv[k] / tot_var[k];
            }
            a_num += (Av_mean - av[i]) * Av_varinv;
            a_den += Av_varinv;
            double d = a_num / a_den;
            d *= step[q];
            for (int k = 0; k < nb; k++) { /* :396-398 */
                r_num += dr[k] * rs[k] / tot_var[k];
                r_den += dr[k] * dr[k] / tot_var[k];
            }
            r_num += (Rv_mean - rv[i]) * Rv_varinv;
            r_den += Rv_varinv;
            double e = r_num / r_den;
            e *= step[q];
            if (d < avmin - av[i]) d = avmin - av[i]; /* :404-420 */
            if (d > avmax - av[i]) d = avmax - av[i];
            av[i] += d;
            if (e < rvmin - rv[i]) e = rvmin - rv[i];
            if (e > rvmax - rv[i]) e = rvmax - rv[i];
            rv[i] += e;
            double models[64];
            get_sed_mle_one(flux, tot_var, mco + (size_t)i * nb * 3, nb, av[i], rv[i], o, models,
                            rvv, dr, &scale_new[q], icov_new + (size_t)q * 9, rs); /* :423 */
            double c2 = 0.;
            for (int k = 0; k < nb; k++) c2 += rs[k] * rs[k] / tot_var[k]; /* :792 */
            chi2_new[q] = c2;
            lnl_new[q] = -0.5 * c2; /* :795 */
        }
        /* stopping criterion (:798-799) */
        double mx = -INFINITY;
        for (int64_t q = 0; q < nsel; q++)
            if (lnl_new[q] > mx) mx = lnl_new[q];
        lerr = -INFINITY;
        for (int64_t q = 0; q < nsel; q++)
            if (lnl_new[q] > mx + ln_sub) {
                double dl = fabs(lnl_new[q] - lnl_old[q]);
                if (dl > lerr) lerr = dl;
            }
        for (int64_t q = 0; q < nsel; q++) { /* :802-803 */
            if (lnl_new[q] < lnl_old[q]) step[q] /= rescaling;
            lnl_old[q] = lnl_new[q];
        }
        if (n_flux >= max_iter) break;
    }
    /* scatter back (:806-810) */
    if (n_flux > 0) {
        double sumlog = 0.;
        for (int k = 0; k < nb; k++) sumlog += log(tot_var[k]);
        const double cst = -0.5 * (Ndim * log(2. * M_PI) + sumlog);
        for (int64_t q = 0; q < nsel; q++) {
            int64_t i = sel[q];
            lnl[i] = lnl_new[q] + cst;
            chi2[i] = chi2_new[q];
            scale[i] = scale_new[q];
            memcpy(icov + (size_t)i * 9, icov_new + (size_t)q * 9, sizeof(double) * 9);
        }
    }
    /* dimensionality prior (:813-815) */
    if (o->dim_prior)
        for (int64_t i = 0; i < nmodel; i++) lnl[i] = chisquare_logpdf(chi2[i], (double)(Ndim - 3));

    if (diag) { diag[0] = Ndim; diag[1] = n_mag; diag[2] = n_flux; diag[3] = nsel; }
    free(icov_new); free(sel); free(mco); free(resid); free(rvecs); free(drvecs); free(wk);
    return 0;
}

int brutus_ref_loglike(const double *data, const double *err, uint8_t *mask_io, int nfilt,
                       const float *coeffs, int64_t nmodel, const ref_options *o, double parallax,
                       double parallax_err, double *lnl, double *chi2, double *scale, double *av,
                       double *rv, double *icov, int64_t *diag, uint8_t *surv_out, double *lnlp_out) {
    return loglike_impl(data, err, mask_io, nfilt, coeffs, nmodel, o, parallax, parallax_err, NULL, NULL, lnl,
                        chi2, scale, av, rv, icov, diag, surv_out, lnlp_out);
}

/* the same with the caller's per-model av_init / rv_init (nmodel each; NULL = the prior mean, :700-703) */
int brutus_ref_loglike_init(const double *data, const double *err, uint8_t *mask_io, int nfilt,
                            const float *coeffs, int64_t nmodel, const ref_options *o, double parallax,
                            double parallax_err, const double *av_init, const double *rv_init, double *lnl,
                            double *chi2, double *scale, double *av, double *rv, double *icov,
                            int64_t *diag, uint8_t *surv_out, double *lnlp_out) {
    return loglike_impl(data, err, mask_io, nfilt, coeffs, nmodel, o, parallax, parallax_err, av_init, rv_init,
                        lnl, chi2, scale, av, rv, icov, diag, surv_out, lnlp_out);
}

/*
 * The O(Nmodel) operations that follow loglike in the per-star loop and that the CUDA path fuses
 * (SURVEY.md section 8 row a-7):
 *   - external Gaussian label priors added to lnlike   (brutus/fitting.py:1995-2009)
 *   - rough parallax prior in scale space               (brutus/fitting.py:976-982,
 *                                                        brutus/pdf.py:178-222, :225-260)
 *   - non-finite -> -1e300                              (brutus/fitting.py:983-985)
 *   - first selection lnprob > max + ln(wt_thresh)      (brutus/fitting.py:988-991)
 * lnl is updated in place (as the reference does); lnprob and sel_flag are outputs.
 * labels: (nlabel, nmodel) float64 column-major per label; ext_mean/ext_std (nlabel), a label
 * is skipped unless isfinite(mean) and std > 0 (:1999).  have_parallax = 0 mimics parallax=None.
 * returns the number selected.
 */
int64_t brutus_ref_select(int64_t nmodel, double *lnl, const double *scale, const double *icov,
                          int have_parallax, double parallax, double parallax_err, int nlabel,
                          const double *labels, const double *ext_mean, const double *ext_std,
                          double wt_thresh, double *lnprob, uint8_t *sel_flag) {
    for (int l = 0; l < nlabel; l++) {
        if (!(isfinite(ext_mean[l]) && ext_std[l] > 0.)) continue;
        double ivar = 1. / (ext_std[l] * ext_std[l]);
        double cst = log(2. * M_PI * ext_std[l] * ext_std[l]);
        const double *lab = labels + (size_t)l * nmodel;
        for (int64_t i = 0; i < nmodel; i++) {
            double c2 = (lab[i] - ext_mean[l]) * (lab[i] - ext_mean[l]);
            c2 *= ivar;
            lnl[i] += -0.5 * (c2 + cst);
        }
    }
    int apply = have_parallax && isfinite(parallax) && isfinite(parallax_err) &&
                (parallax / parallax_err > 4.);
    double s_mean = 0., s_std = 0.;
    if (apply) { /* brutus/pdf.py:252-256 */
        double pm = parallax > 0. ? parallax : 0., pe = parallax_err;
        s_mean = pm * pm + pe * pe;
        s_std = sqrt(2 * pe * pe * pe * pe + 4 * pm * pm * pe * pe);
    }
    double mx = -INFINITY;
    for (int64_t i = 0; i < nmodel; i++) {
        double lp = lnl[i];
        if (apply) {
            double serr = 1. / sqrt(fabs(icov[(size_t)i * 9]));
            double svar_tot = s_std * s_std + serr * serr;
            double c2 = (scale[i] - s_mean) * (scale[i] - s_mean) / svar_tot;
            double lnorm = log(2. * M_PI * svar_tot);
            lp = lnl[i] + -0.5 * (c2 + lnorm);
        }
        if (!isfinite(lp)) lp = -1e300;
        lnprob[i] = lp;
        if (lp > mx) mx = lp;
    }
    double thr = log(wt_thresh) + mx;
    int64_t n = 0;
    for (int64_t i = 0; i < nmodel; i++) {
        sel_flag[i] = (uint8_t)(lnprob[i] > thr);
        n += sel_flag[i];
    }
    return n;
}

/*
 * Many stars, OpenMP over stars (one workspace per thread): the "all host cores" CPU baseline.
 * Only per-star summaries are kept: best[6*s + {0..5}] = (argmax lnl, max lnl, chi2, scale, av,
 * rv at the argmax) and diag[4*s ..].  flux/err (nstar,nfilt), mask (nstar,nfilt) uint8.
 */
int brutus_ref_loglike_batch(int nstar, const double *flux, const double *err, const uint8_t *mask,
                             int nfilt, const float *coeffs, int64_t nmodel, const ref_options *o,
                             const double *parallax, const double *parallax_err, double *best,
                             int64_t *diag, int nthreads) {
    int rc_all = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * (size_t)nmodel * 14);
#pragma omp for schedule(dynamic, 1)
        for (int s = 0; s < nstar; s++) {
            if (!buf) { rc_all = -2; continue; }
            uint8_t m[64];
            memcpy(m, mask + (size_t)s * nfilt, (size_t)nfilt);
            double *lnl = buf, *chi2 = buf + nmodel, *sc = buf + 2 * nmodel, *av = buf + 3 * nmodel,
                   *rv = buf + 4 * nmodel, *icov = buf + 5 * nmodel;
            int rc = brutus_ref_loglike(flux + (size_t)s * nfilt, err + (size_t)s * nfilt, m, nfilt,
                                        coeffs, nmodel, o, parallax ? parallax[s] : NAN,
                                        parallax_err ? parallax_err[s] : NAN, lnl, chi2, sc, av, rv,
                                        icov, diag + 4 * (size_t)s, NULL, NULL);
            if (rc) { rc_all = rc; continue; }
            int64_t k = 0;
            for (int64_t i = 1; i < nmodel; i++)
                if (lnl[i] > lnl[k]) k = i;
            double *b = best + 6 * (size_t)s;
            b[0] = (double)k; b[1] = lnl[k]; b[2] = chi2[k]; b[3] = sc[k]; b[4] = av[k]; b[5] = rv[k];
        }
        free(buf);
    }
    return rc_all;
}

int brutus_ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
