// api.cu -- C ABI (include/brutus_b200.h), host orchestration and the band-count-independent
// kernels of the brute-force likelihood sweep.  See DESIGN.md for the pipeline:
//
//   prepare star rows (host, float64)  ->  k_magfit: the full-grid sweep (speculated iteration count);
//   writes per-star maxima and a 1-bit candidate map, nothing else  ->  verify / re-sweep mispredicted
//   stars  ->  k_cand_scan + k_expand: candidate map -> ordered candidate records  ->  k_refit: exact
//   re-fit + exact cull of the candidates  ->  k_flux / k_flux_ctl on the survivors until every star
//   converges (convergence decided on the device)  ->  k_final: lnlike, lnprob, per-star max  ->
//   either full-length outputs (B1, bf_loglike_full) or threshold + ordered compaction + k_records
//   (B2, bf_sweep_batch).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include <cub/device/device_segmented_radix_sort.cuh>

#include "../../include/brutus_b200.h"
#include "common.cuh"
#include "posterior.cuh"

namespace bf {

// ---- launch tables, one per band count (inst.cu) ---------------------------------------------------
#define BF_DECL(n) const KTable<float>* ktable_f32_##n(); const KTable<double>* ktable_f64_##n();
BF_DECL(1) BF_DECL(2) BF_DECL(3) BF_DECL(4) BF_DECL(5) BF_DECL(6) BF_DECL(7) BF_DECL(8)
BF_DECL(9) BF_DECL(10) BF_DECL(11) BF_DECL(12) BF_DECL(13) BF_DECL(14) BF_DECL(15) BF_DECL(16)
#undef BF_DECL
template <typename T> const KTable<T>* get_ktable(int nb);
#define BF_CASE(n, sfx) case n: return ktable_##sfx##_##n();
#define BF_ALL(sfx) BF_CASE(1, sfx) BF_CASE(2, sfx) BF_CASE(3, sfx) BF_CASE(4, sfx) BF_CASE(5, sfx) \
    BF_CASE(6, sfx) BF_CASE(7, sfx) BF_CASE(8, sfx) BF_CASE(9, sfx) BF_CASE(10, sfx) BF_CASE(11, sfx) \
    BF_CASE(12, sfx) BF_CASE(13, sfx) BF_CASE(14, sfx) BF_CASE(15, sfx) BF_CASE(16, sfx)
template <> const KTable<float>* get_ktable<float>(int nb) { switch (nb) { BF_ALL(f32) } return nullptr; }
template <> const KTable<double>* get_ktable<double>(int nb) { switch (nb) { BF_ALL(f64) } return nullptr; }

// =================================================================================================
// band-count-independent kernels
// =================================================================================================

// grid re-tiling: user layout (C or Fortran order of (nmodel, nfilt, 3)) -> the coefficient-major copy
// [coef][band][npad] read by the sweep and the model-major copy [npad][row_stride] read by the gathers
__global__ void k_retile(const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ rows,
                         int64_t nmodel, int64_t npad, int nfilt, int rs, int layout) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    for (int k = 3 * nfilt; k < rs; k++) rows[i * rs + k] = 0.f;
    for (int j = 0; j < nfilt; j++)
        for (int c = 0; c < 3; c++) {
            // the padding models (i >= nmodel) replicate the last real model: they can then take part in the
            // sweep's per-star max-reductions unmasked (a duplicate never changes a maximum); only their
            // candidate bits are suppressed
            const int64_t ii = i < nmodel ? i : nmodel - 1;
            float v = layout == BF_LAYOUT_C ? src[(ii * nfilt + j) * 3 + c]
                                            : src[((int64_t)c * nfilt + j) * nmodel + ii];
            // A band the caller masks out is neutralised by zero star weights, which only works on finite
            // coefficients: sentinel (NaN / inf) entries are stored as 0.  The reference slices masked bands
            // away (brutus/fitting.py:714) and never reads them; a sentinel in a band that IS used makes the
            // reference return NaN for that model, here it is fitted with a zero coefficient.
            if (!isfinite(v)) v = 0.f;
            dst[((int64_t)c * nfilt + j) * npad + i] = v;
            rows[i * rs + c * nfilt + j] = v;
        }
}

template <typename T> __global__ void k_convert_labels(const double* src, T* dst, int64_t nmodel, int64_t npad, int nlabel) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    // padding entries replicate the last real model, like the grid's (k_retile)
    for (int l = 0; l < nlabel; l++) dst[(int64_t)l * npad + i] = (T)src[(int64_t)l * nmodel + (i < nmodel ? i : nmodel - 1)];
}

// get_seds / _get_seds (brutus/utils.py:286-347) for n (model, Av, Rv) samples; one thread per (sample, band)
template <typename T>
__global__ void k_get_seds(const float* __restrict__ rows, int rs, int nfilt, int64_t n, const int* __restrict__ idx,
                           const double* __restrict__ av, const double* __restrict__ rv, int flux,
                           double* __restrict__ seds, double* __restrict__ rvecs, double* __restrict__ drvecs) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nfilt) return;
    const int64_t i = t / nfilt;
    const int j = (int)(t - i * nfilt);
    const float* __restrict__ row = rows + (int64_t)(idx ? idx[i] : i) * rs;
    const T mag0 = (T)row[j], r0 = (T)row[nfilt + j], dr = (T)row[2 * nfilt + j];
    T dv = dr;                                   // :336
    T rvec = r0 + (T)rv[i] * dr;                 // :337
    T sed = mag0 + (T)av[i] * rvec;              // :338
    if (flux) {                                  // :341-345
        sed = Num<T>::exp2(T(-kC2) * sed);
        rvec *= T(kFac) * sed;
        dv *= T(kFac) * sed;
    }
    seds[t] = (double)sed;
    if (rvecs) rvecs[t] = (double)rvec;
    if (drvecs) drvecs[t] = (double)dv;
}

// photometric_offsets, the per-sample part (brutus/utils.py:1268-1271 and :1293-1309): one thread per posterior
// sample (object o, sample k).  The sample's SED in flux on the staged grid (_get_seds :286-347, then / dist^2),
// and for every band b that is being fitted the log-likelihood of the sample with band b left out --
// phot_loglike(phot * old_offsets, err * old_offsets, mask without b, seds) (:1162-1222) -- summed over the
// bands in their order, float64.
template <typename T>
__global__ void __launch_bounds__(256) k_offsets_lnl(const float* __restrict__ rows, int rs, int nfilt, int64_t nobj, int nsamps,
                                                     const int* __restrict__ idx, const double* __restrict__ av,
                                                     const double* __restrict__ rv, const double* __restrict__ dist,
                                                     const double* __restrict__ phot, const double* __restrict__ var,
                                                     const uint8_t* __restrict__ mask, const uint8_t* __restrict__ mask_fit,
                                                     int dim_prior, double* __restrict__ seds, double* __restrict__ lnl) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = nobj * nsamps;
    if (t >= n) return;
    const int64_t o = t / nsamps;
    const float* __restrict__ row = rows + (int64_t)idx[t] * rs;
    const double d = dist[t], id2 = 1.0 / (d * d);
    double sed[BF_MAX_FILT];
#pragma unroll 1
    for (int j = 0; j < nfilt; j++) {
        const T mag0 = (T)row[j], r0 = (T)row[nfilt + j], dr = (T)row[2 * nfilt + j];
        const T rvec = r0 + (T)rv[t] * dr;                       // :337
        const T m = mag0 + (T)av[t] * rvec;                      // :338
        sed[j] = (double)Num<T>::exp2(T(-kC2) * m) * id2;        // :343, :1271
        seds[t * nfilt + j] = sed[j];
    }
    const double* __restrict__ ph = phot + o * nfilt;            // phot * old_offsets
    const double* __restrict__ vr = var + o * nfilt;             // (err * old_offsets)^2
    const uint8_t* __restrict__ mk = mask + o * nfilt;
    for (int b = 0; b < nfilt; b++) {
        if (!mask_fit[b]) continue;
        double chi2 = 0., lnvar = 0.;
        int ndim = 0;
        for (int j = 0; j < nfilt; j++) {
            if (!mk[j] || j == b) continue;
            const double r = ph[j] - sed[j];
            chi2 += r * r / vr[j];
            lnvar += log(vr[j]);
            ndim++;
        }
        double l;
        if (dim_prior) {                                         // chi2 distribution with Ndim - 3 dof (:1215-1219)
            const double a = 0.5 * (ndim - 3);
            const double xl = (a - 1.) == 0. ? 0. : (a - 1.) * log(chi2);   // scipy.special.xlogy
            l = xl - 0.5 * chi2 - lgamma(a) - 0.6931471805599453 * a;
        } else {
            l = -0.5 * chi2 - 0.5 * (ndim * 1.8378770664093453 + lnvar);
        }
        lnl[(int64_t)b * n + t] = l;
    }
}

// weights of the samples of one object for one fitted band: exp(lnl - logsumexp(lnl)) over the object's samples
// (brutus/utils.py:1308-1309; scipy's logsumexp: a non-finite maximum is replaced by 0).  One warp per (band, object).
__global__ void __launch_bounds__(256) k_offsets_wt(const double* __restrict__ lnl, const uint8_t* __restrict__ mask_fit,
                                                    int nfilt, int64_t nobj, int nsamps, double* __restrict__ wt) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nfilt * nobj) return;
    const int b = (int)(w / nobj);
    if (!mask_fit[b]) return;
    const double* __restrict__ x = lnl + w * nsamps;
    double m = -CUDART_INF;
    for (int k = lane; k < nsamps; k += 32) m = fmax(m, x[k]);
    for (int off = 16; off; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (!isfinite(m)) m = 0.;   // scipy; a NaN entry (fmax skips it) still turns the sum, hence every weight, into NaN
    double s = 0.;
    for (int k = lane; k < nsamps; k += 32) s += exp(x[k] - m);
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const double lse = log(s) + m;
    for (int k = lane; k < nsamps; k += 32) wt[w * nsamps + k] = exp(x[k] - lse);
}

template <typename T>
__global__ void k_reset_red(typename Enc<T>::U* red, const int* list, int nlist, unsigned mask) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlist * kNumRed) return;
    int s = t / kNumRed, k = t % kNumRed;
    if (mask >> k & 1u) red[(int64_t)list[s] * kNumRed + k] = Enc<T>::enc(Num<T>::neg_inf());
}

// block-wide exclusive scan helper: returns the exclusive prefix of v within the CTA (1024 threads)
// and the CTA total through `total`
__device__ __forceinline__ int block_exscan_1024(int v, int* s_w, int& total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = s_w[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, w, o);
            if (threadIdx.x >= o) w += y;
        }
        s_w[threadIdx.x] = w;
    }
    __syncthreads();
    int incl = x + ((threadIdx.x >> 5) ? s_w[(threadIdx.x >> 5) - 1] : 0);
    total = s_w[31];
    __syncthreads();
    return incl - v;
}

// one CTA per star: exclusive scan of the popcounts of the star's bitmap words -> wpre; total -> nbits.
// Run on the candidate map (candidate counts, when a group has to be split) and, after k_sel cleared the
// bits of the candidates that failed the selection, on the selection map: wpre[word] + popc(lower bits) is
// then the position of a selected model within its star, in ascending model order.
__global__ void __launch_bounds__(1024) k_cand_scan(const uint32_t* __restrict__ cand, int* __restrict__ wpre,
                                                    const int* __restrict__ list, int64_t nwords,
                                                    int64_t* __restrict__ nbits) {
    __shared__ int s_w[32];
    const int slot = list[blockIdx.x];
    const uint32_t* c = cand + (int64_t)slot * nwords;
    int* w = wpre + (int64_t)slot * nwords;
    int carry = 0;
    for (int64_t b = 0; b < nwords; b += 1024) {
        int64_t t = b + threadIdx.x;
        int v = t < nwords ? __popc(c[t]) : 0;
        int tot;
        int ex = block_exscan_1024(v, s_w, tot);
        if (t < nwords) w[t] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) nbits[slot] = carry;
}

// ---- passes over the n records the sweep appended --------------------------------------------------------
template <typename T> struct PassParams {
    PoolArrays<T> pool;
    int64_t n;
    const T* stars;
    int* star_int;
    typename Enc<T>::U* red;
    DevOpts<T> o;
    int* fixlist;      // k_cull: records refined as likely survivors that are not survivors
    int* nfix;
    // k_final
    int64_t npad;
    const T* labels;   // [nlabel][npad]
    const T* ext;      // [batch][nlabel][3]
    int nlabel;
    // k_sel
    uint32_t* cand;
    int64_t nwords;
};

// a record is live if its star has not been swept again since (another iteration count, or every model
// as a candidate)
template <typename T> __device__ __forceinline__ bool rec_live(const PassParams<T>& p, int tag) {
    return tag_epoch(tag) == (p.star_int[tag_slot(tag) * SI_COUNT + SI_EPOCH] & 0xff);
}

// The exact cull (brutus/fitting.py:758-759) against the final per-star maximum of lnl_p, the survivor
// counts, and the convergence reductions of the flux iterations the sweep already ran (:798-799):
//   "lerr <= ltol"  <=>  max{lnl_new_i : |lnl_new_i - lnl_old_i| > ltol} <= max lnl_new + ln(ltol_subthresh)
template <typename T> __global__ void __launch_bounds__(kTile) k_cull(const PassParams<T> p) {
    __shared__ StarAgg<T, 2> agg;
    const int which[2] = {RED_FL, RED_FB};
    agg.init();
    __syncthreads();
    int64_t lo, hi;
    pass_range(p.n, lo, hi);
    for (int64_t base = lo; base < hi; base += kPassStep) {
        int tag[kPassU], slot[kPassU];
        bool live[kPassU], surv[kPassU];
        T lp[kPassU], lnew[kPassU], lold[kPassU], thr[kPassU];
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            const int64_t q = base + u * kTile + threadIdx.x;
            tag[u] = q < hi ? p.pool.sflag[q] : -1;
        }
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            const int64_t q = base + u * kTile + threadIdx.x;
            live[u] = tag[u] != -1 && rec_live(p, tag[u]);
            slot[u] = live[u] ? tag_slot(tag[u]) : -1;
            lp[u] = live[u] ? p.pool.lp[q] : T(0);
            lnew[u] = live[u] ? p.pool.lold[q] : T(0);
            lold[u] = live[u] ? p.pool.lprev[q] : T(0);
            thr[u] = live[u] ? Enc<T>::dec(p.red[(int64_t)slot[u] * kNumRed + RED_LP]) + p.o.ln_init : T(0);
        }
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            const int64_t q = base + u * kTile + threadIdx.x;
            surv[u] = live[u] && lp[u] > thr[u];
            T v[2] = {Num<T>::neg_inf(), Num<T>::neg_inf()};
            if (surv[u]) {
                v[0] = (lnew[u] == lnew[u]) ? lnew[u] : Num<T>::neg_inf();
                v[1] = (tabs(lnew[u] - lold[u]) > p.o.ltol) ? v[0] : Num<T>::neg_inf();
                p.pool.sflag[q] = tag[u] | kFlagSurv << 24;
            }
            // refined as a likely survivor, but not a survivor: onto the fix-up list (one atomic per warp: per-record
            // atomics on the single counter cost more than the rest of the pass)
            const bool fix = !surv[u] && live[u] && (tag_flags(tag[u]) & kFlagFluxed);
            const unsigned fb = __ballot_sync(0xffffffffu, fix);
            if (fb) {
                const int lane = threadIdx.x & 31, lead = __ffs(fb) - 1;
                int at = 0;
                if (lane == lead) at = atomicAdd(p.nfix, __popc(fb));
                at = __shfl_sync(0xffffffffu, at, lead);
                if (fix) p.fixlist[at + __popc(fb & ((1u << lane) - 1u))] = (int)q;
            }
            agg.add(p.red, which, p.star_int + SI_NSURV, SI_COUNT, slot[u], surv[u], v, surv[u]);
        }
    }
    agg.flush(p.red, which, p.star_int + SI_NSURV, SI_COUNT);
}

// Survivors of the stars whose flux loop continues (SI_ACTIVE after the first k_flux_ctl): their pool indices, for
// k_flux_more.  A light walk over the record tags; one atomic per warp.
template <typename T> __global__ void __launch_bounds__(kTile) k_flux_list(const PassParams<T> p, int* list, int* nlist) {
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * kPassStep; base < p.n; base += (int64_t)gridDim.x * kPassStep) {
        int tag[kPassU];
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            const int64_t q = base + u * kTile + threadIdx.x;
            tag[u] = q < p.n ? p.pool.sflag[q] : 0;
        }
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            bool act = false;
            if (tag_flags(tag[u]) & kFlagSurv) {
                const int* si = p.star_int + tag_slot(tag[u]) * SI_COUNT;
                act = tag_epoch(tag[u]) == (si[SI_EPOCH] & 0xff) && si[SI_ACTIVE] != 0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, act);
            if (bal) {
                int at = 0;
                if (lane == 0) at = atomicAdd(nlist, __popc(bal));
                at = __shfl_sync(0xffffffffu, at, 0);
                if (act) list[at + __popc(bal & ((1u << lane) - 1u))] = (int)(base + u * kTile + threadIdx.x);
            }
        }
    }
}

// Convergence control of the flux loops, on the device (one CTA): "lerr > ltol" (:781, :798-799) restated
// on the two max-reductions of the `nit` iterations that just ran (first: the ones inside the sweep).
// *any_out = 1 if some star of the group needs another iteration.
template <typename T>
__global__ void __launch_bounds__(1024) k_flux_ctl(int* star_int, typename Enc<T>::U* red, const int* list, int nlist,
                                                   int first, int nit, int max_iter, T ln_sub, int* any_out) {
    int any = 0;
    for (int t = threadIdx.x; t < nlist; t += blockDim.x) {
        const int slot = list[t];
        int* si = star_int + slot * SI_COUNT;
        if (first) {   // start of the star's flux loop (:778-781)
            si[SI_ACTIVE] = si[SI_NSURV] > 0 ? 1 : 0;
            si[SI_NFLUX] = 0;
        }
        if (si[SI_ACTIVE]) {
            typename Enc<T>::U* r = red + (int64_t)slot * kNumRed;
            const int nf = si[SI_NFLUX] + nit;
            si[SI_NFLUX] = nf;
            const bool more = Enc<T>::dec(r[RED_FB]) > Enc<T>::dec(r[RED_FL]) + ln_sub;
            if (more && nf < max_iter) any = 1;
            else si[SI_ACTIVE] = 0;
            r[RED_FL] = Enc<T>::enc(Num<T>::neg_inf());
            r[RED_FB] = Enc<T>::enc(Num<T>::neg_inf());
        }
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) *any_out = any;
}

// lnlike / lnprob of every candidate from its final (chi2, scale, s_den), per-star max(lnprob) (:990)
template <typename T> __global__ void __launch_bounds__(kTile) k_final(const PassParams<T> p) {
    __shared__ StarAgg<T, 1> agg;
    const int which[1] = {RED_LNP};
    agg.init();
    __syncthreads();
    int64_t lo, hi;
    pass_range(p.n, lo, hi);
    for (int64_t base = lo; base < hi; base += kPassStep) {
        int tag[kPassU], slot[kPassU];
        bool live[kPassU];
        T chi2[kPassU], sden[kPassU], scale[kPassU];
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            const int64_t q = base + u * kTile + threadIdx.x;
            tag[u] = q < hi ? p.pool.sflag[q] : -1;
        }
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            const int64_t q = base + u * kTile + threadIdx.x;
            live[u] = tag[u] != -1 && rec_live(p, tag[u]);
            slot[u] = live[u] ? tag_slot(tag[u]) : -1;
            chi2[u] = live[u] ? p.pool.chi2[q] : T(1);
            sden[u] = live[u] ? p.pool.sden[q] : T(1);
            scale[u] = live[u] ? p.pool.scale[q] : T(1);
        }
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            const int64_t q = base + u * kTile + threadIdx.x;
            T lp[1] = {Num<T>::neg_inf()};
            if (live[u]) {
                const T* __restrict__ srow = p.stars + (int64_t)slot[u] * kStarStride;
                T ext = T(0);
                if (p.nlabel > 0)
                    ext = ext_prior<T>(p.labels, p.ext + (int64_t)slot[u] * p.nlabel * 3, p.nlabel, p.npad, p.pool.model[q]);
                T lnl;
                lnl_lnprob<T>(chi2[u], sden[u], scale[u], (tag_flags(tag[u]) & kFlagSurv) != 0, srow, p.o.dim_prior, ext, lnl, lp[0]);
                p.pool.lnl[q] = lnl;
                p.pool.lnprob[q] = lp[0];
            }
            agg.add(p.red, which, nullptr, 0, slot[u], live[u], lp, false);
        }
    }
    agg.flush(p.red, which, nullptr, 0);
}

// brutus/fitting.py:990-991: lnprob > max(lnprob) + ln(wt_thresh).  Selected records are flagged; the bit of a
// candidate that fails is cleared, which turns the sweep's candidate map into the selection map.
template <typename T> __global__ void __launch_bounds__(kTile) k_sel(const PassParams<T> p) {
    const int64_t hi = p.n;
    {   // one step per CTA, in pool order: the CTAs in flight stay within a narrow window of every array
        const int64_t base = (int64_t)blockIdx.x * kPassStep;
        int tag[kPassU];
        T lnp[kPassU], thr[kPassU];
        bool live[kPassU];
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            const int64_t q = base + u * kTile + threadIdx.x;
            tag[u] = q < hi ? p.pool.sflag[q] : -1;
            lnp[u] = q < hi ? p.pool.lnprob[q] : T(0);
        }
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            live[u] = tag[u] != -1 && rec_live(p, tag[u]);
            thr[u] = live[u] ? Enc<T>::dec(p.red[(int64_t)tag_slot(tag[u]) * kNumRed + RED_LNP]) + p.o.ln_wt : T(0);
        }
#pragma unroll
        for (int u = 0; u < kPassU; u++) {
            if (!live[u]) continue;
            const int64_t q = base + u * kTile + threadIdx.x;
            if (lnp[u] > thr[u]) {
                p.pool.sflag[q] = tag[u] | kFlagSel << 24;
            } else {
                const int i = p.pool.model[q];
                atomicAnd(&p.cand[(int64_t)tag_slot(tag[u]) * p.nwords + (i >> 5)], ~(1u << (i & 31)));
            }
        }
    }
}

// O = element type of the outputs: double for the full-length B1 arrays (the reference returns float64),
// T for the compacted B2 records (no point shipping more bits than were computed).
template <typename T, typename O> struct OutParams {
    PoolArrays<T> pool;
    int64_t n;
    const int* star_int;
    // compacted records: rows of a [11][ld] matrix: lnl, scale, av, chi2, rv, icov(ss, sa, sr, aa, ar, rr); only the
    // first `nrows` are produced (3, 5 or 11).  Full-length: separate arrays indexed by model, icov 9 per model.
    O *o_lnl, *o_chi2, *o_scale, *o_av, *o_rv, *o_icov;
    int64_t ld;
    int nrows;
    int* o_idx;    // model index of each record
    int* o_star;   // optional: star slot of each record (device posterior)
    const uint32_t* sel;   // selection map and its per-word prefix
    const int* wpre;
    int64_t nwords;
    const int64_t* base;   // [batch] first output position of each star
    int* ord;              // [nsel] pool index of the t-th selected record in (star, model) order
    int64_t nsel;
};

// selected records -> their (star, model)-ordered position: rank of the model's bit in the selection map.
// ord[t] = pool index of the t-th selected record.  The posterior reads the pool through it (no copy of the
// records); the records-out path gathers through it so that its 12 output streams are written coalesced.
template <typename T> __global__ void __launch_bounds__(kTile) k_ord(const OutParams<T, T> p) {
    const PoolArrays<T>& pl = p.pool;
    const int64_t hi = p.n;
    // one step per CTA, in pool order: the CTAs in flight stay within a narrow window of every array
    const int64_t base = (int64_t)blockIdx.x * kPassStep;
    int tag[kPassU], model[kPassU];
#pragma unroll
    for (int u = 0; u < kPassU; u++) {
        const int64_t q = base + u * kTile + threadIdx.x;
        tag[u] = q < hi ? pl.sflag[q] : 0;
        model[u] = q < hi ? pl.model[q] : 0;
    }
#pragma unroll
    for (int u = 0; u < kPassU; u++) {
        const int slot = tag_slot(tag[u]);
        if ((tag_flags(tag[u]) & kFlagSel) && tag_epoch(tag[u]) == (p.star_int[slot * SI_COUNT + SI_EPOCH] & 0xff)) {
            const int64_t w = (int64_t)slot * p.nwords + (model[u] >> 5);
            const int64_t t = p.base[slot] + p.wpre[w] + __popc(p.sel[w] & ((1u << (model[u] & 31)) - 1u));
            p.ord[t] = (int)(base + u * kTile + threadIdx.x);
        }
    }
}

// compacted, ordered records of the selection: one thread per output position, gathering from the pool
template <typename T> __global__ void __launch_bounds__(kTile) k_out(const OutParams<T, T> p) {
    const PoolArrays<T>& pl = p.pool;
    const int64_t o = (int64_t)blockIdx.x * kTile + threadIdx.x;
    if (o >= p.nsel) return;
    const int64_t q = p.ord[o];
    p.o_idx[o] = pl.model[q];
    if (p.o_star) p.o_star[o] = tag_slot(pl.sflag[q]);
    const T lnl = pl.lnl[q], sc = pl.scale[q], av = pl.av[q];
    if (p.nrows <= 3) { p.o_lnl[o] = lnl; p.o_scale[o] = sc; p.o_av[o] = av; return; }
    const T chi2 = pl.chi2[q], rv = pl.rv[q];
    if (p.nrows <= 5) { p.o_lnl[o] = lnl; p.o_scale[o] = sc; p.o_av[o] = av; p.o_chi2[o] = chi2; p.o_rv[o] = rv; return; }
    const T ss = pl.sden[q], sa = pl.isa[q], sr = pl.isr[q], aa = pl.iaa[q], ar = pl.iar[q], rr = pl.irr[q];
    p.o_lnl[o] = lnl; p.o_scale[o] = sc; p.o_av[o] = av; p.o_chi2[o] = chi2; p.o_rv[o] = rv;
    T* w6 = p.o_icov + o;
    w6[0] = ss; w6[p.ld] = sa; w6[2 * p.ld] = sr; w6[3 * p.ld] = aa; w6[4 * p.ld] = ar; w6[5 * p.ld] = rr;
}

// every model of one star (B1): record -> position `model` of the full-length float64 arrays
template <typename T> __global__ void __launch_bounds__(kTile) k_out_full(const OutParams<T, double> p) {
    const int64_t q = (int64_t)blockIdx.x * kTile + threadIdx.x;
    if (q >= p.n) return;
    const int tag = p.pool.sflag[q];
    const int slot = tag_slot(tag);
    if (tag_epoch(tag) != (p.star_int[slot * SI_COUNT + SI_EPOCH] & 0xff)) return;
    const PoolArrays<T>& pl = p.pool;
    const int64_t t = pl.model[q];
    p.o_lnl[t] = (double)pl.lnl[q];
    p.o_chi2[t] = (double)pl.chi2[q];
    p.o_scale[t] = (double)pl.scale[q];
    p.o_av[t] = (double)pl.av[q];
    p.o_rv[t] = (double)pl.rv[q];
    if (!p.o_icov) return;
    double* w = p.o_icov + t * 9;
    const double ss = (double)pl.sden[q], sa = (double)pl.isa[q], sr = (double)pl.isr[q], aa = (double)pl.iaa[q],
                 ar = (double)pl.iar[q], rr = (double)pl.irr[q];
    w[0] = ss; w[1] = sa; w[2] = sr; w[3] = sa; w[4] = aa; w[5] = ar; w[6] = sr; w[7] = ar; w[8] = rr;
}

// in-place exclusive scan of n ints by one CTA; total -> *tot
__global__ void __launch_bounds__(1024) k_scan_blocks(int* v, int64_t n, int64_t* tot) {
    // exclusive scan of n block counts in place, one CTA, 8 consecutive entries per thread and round (the scan of the
    // 578 k block counts of a 1 000-star batch took 0.5 ms at one entry per thread: 565 rounds of two barriers each)
    constexpr int kPer = 8;
    __shared__ int s_w[32];
    int64_t carry = 0;
    for (int64_t b = 0; b < n; b += 1024 * kPer) {
        const int64_t t0 = b + (int64_t)threadIdx.x * kPer;
        int x[kPer];
        if (t0 + kPer <= n) {
            const int4 a = *reinterpret_cast<const int4*>(v + t0), c = *reinterpret_cast<const int4*>(v + t0 + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = c.x; x[5] = c.y; x[6] = c.z; x[7] = c.w;
        } else {
#pragma unroll
            for (int k = 0; k < kPer; k++) x[k] = t0 + k < n ? v[t0 + k] : 0;
        }
        int sum = 0;
#pragma unroll
        for (int k = 0; k < kPer; k++) { const int y = x[k]; x[k] = sum; sum += y; }   // exclusive within the thread
        int total;
        const int ex = block_exscan_1024(sum, s_w, total);
        const int base = (int)carry + ex;
        if (t0 + kPer <= n) {
            *reinterpret_cast<int4*>(v + t0) = make_int4(base + x[0], base + x[1], base + x[2], base + x[3]);
            *reinterpret_cast<int4*>(v + t0 + 4) = make_int4(base + x[4], base + x[5], base + x[6], base + x[7]);
        } else {
#pragma unroll
            for (int k = 0; k < kPer; k++)
                if (t0 + k < n) v[t0 + k] = base + x[k];
        }
        carry += total;
    }
    if (threadIdx.x == 0) *tot = carry;
}

// =================================================================================================
// host side
// =================================================================================================
static thread_local std::string g_create_error;

struct EngineBase {
    int device = 0;
    int precision = BF_PRECISION_F32;
    std::string err;
    bf_stats stats{};
    virtual ~EngineBase() {}
    virtual int set_grid(const float* co, int64_t nmodel, int nfilt, int layout, bool on_device) = 0;
    virtual int set_labels(const double* labels, int nlabel) = 0;
    virtual int flush_l2() = 0;
    virtual int loglike_full(const double* flux, const double* errv, const uint8_t* mask, double par,
                             double perr, const bf_options* opt, double* lnl, double* chi2,
                             double* scale, double* av, double* rv, double* icov,
                             uint8_t* mask_out, int64_t* diag) = 0;
    virtual int sweep_batch(int64_t nstar, const double* flux, const double* errv, const uint8_t* mask,
                            const double* par, const double* perr, const double* ext_mean,
                            const double* ext_std, const bf_options* opt, int record_rows, int32_t* ndim,
                            int32_t* n_iter, int64_t* n_surv, double* max_lnprob, int64_t* offsets,
                            bf_records* out) = 0;
    virtual int set_model_priors(const double* lnprior, const double* feh, const double* loga) = 0;
    virtual int set_init(const double* av_init, const double* rv_init) = 0;
    virtual int get_seds(int64_t n, const int32_t* idx, const double* av, const double* rv, int flux, double* seds,
                         double* rvecs, double* drvecs) = 0;
    virtual int offsets_weights(int64_t nobj, int nsamps, const double* phot, const double* errv, const uint8_t* mask,
                                const int32_t* idxs, const double* reds, const double* dreds, const double* dists,
                                const double* old_offsets, const uint8_t* mask_fit, int dim_prior, double* seds,
                                double* wt) = 0;
    virtual int fit_batch(int64_t nstar, const double* flux, const double* errv, const uint8_t* mask,
                          const double* par, const double* perr, const double* coords, const double* ext_mean,
                          const double* ext_std, const bf_options* opt, const bf_post_options* po, int32_t* ndim,
                          int32_t* n_iter, int64_t* nsel, double* levid, double* chi2min, bf_draws* out) = 0;
    virtual const char* get_trace() = 0;
    // multi-device calls: the draws of this engine's stars go to [off, off + nstar) of a shared pinned arena
    virtual void set_ext_arena(char* arena, size_t ntot, size_t off) = 0;
    virtual int nfilt_() const = 0;
    virtual int nlabel_() const = 0;
    virtual const char* records_arena(int64_t* cap) = 0;
};

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            char b_[512];                                                                      \
            snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            err = b_;                                                                          \
            return BF_E_CUDA;                                                                  \
        }                                                                                      \
    } while (0)

// Small control read-backs (per-star maxima, counters) do NOT go through the copy engines: a D2H
// cudaMemcpyAsync on the compute stream queues behind the record copies in flight on the copy stream
// (hundreds of MB each), which serialised the kernels of batch g+1 with the record D2H of batch g.
// Instead a tiny kernel stores the words into mapped pinned host memory (SM-issued PCIe writes).
__global__ void k_publish(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst_host, size_t nwords) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nwords) dst_host[t] = src[t];
}

// host array in mapped pinned memory (cudaHostAllocMapped; with UVA the same pointer is valid in kernels)
template <typename P> struct PinVec {
    P* p = nullptr;
    size_t n = 0;
    PinVec() = default;
    PinVec(const PinVec&) = delete;
    PinVec& operator=(const PinVec&) = delete;
    ~PinVec() { release(); }
    cudaError_t resize(size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaHostAlloc((void**)&p, want * sizeof(P), cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) { n = want; std::memset(p, 0, want * sizeof(P)); }
        return e;
    }
    P* data() { return p; }
    const P* data() const { return p; }
    P& operator[](size_t i) { return p[i]; }
    const P& operator[](size_t i) const { return p[i]; }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

template <typename P> struct DevBuf {
    P* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }   // temporaries on error paths (CK returns early) do not leak
    cudaError_t ensure(size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(P));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// Host-side preparation of one star row (float64): the data clean-up and magnitude conversion of
// loglike (brutus/fitting.py:706-725) in the normalised form the kernels consume (common.cuh).
struct StarPrep {
    double row[kStarStride];
    int ndim;
    uint8_t clean[kMaxFilt];
};

static void prep_star(const double* flux, const double* errv, const uint8_t* mask, int nfilt, double par,
                      double perr, int apply_clip, double slack, StarPrep& sp) {
    std::memset(sp.row, 0, sizeof sp.row);
    const double c2 = (2.5 / std::log(10.)) * (2.5 / std::log(10.));
    int ndim = 0, npos = 0;
    double msum = 0., sumlog = 0.;
    double m[kMaxFilt];
    bool pos[kMaxFilt];
    for (int j = 0; j < nfilt; j++) {
        bool cl = mask[j] && std::isfinite(flux[j]) && std::isfinite(errv[j]) && errv[j] > 0.;  // :708-709
        sp.clean[j] = cl;
        pos[j] = false;
        if (!cl) continue;
        ndim++;
        sumlog += std::log(errv[j] * errv[j]);
        if (flux[j] > 0.) {  // finite magnitude (:722-724)
            m[j] = -2.5 * std::log10(flux[j]);
            pos[j] = true;
            msum += m[j];
            npos++;
        }
    }
    const double mbar = npos ? msum / npos : 0.;
    double S = 0.;
    for (int j = 0; j < nfilt; j++) {
        if (!sp.clean[j]) continue;
        const double sig = errv[j];
        sp.row[SR_AL + j] = flux[j] / sig;
        if (pos[j]) {
            sp.row[SR_CM + j] = m[j] - mbar;
            double u = (flux[j] * flux[j]) / (c2 * sig * sig);  // 1 / mags_var (:723)
            sp.row[SR_U + j] = u;
            S += u;
            sp.row[SR_BE + j] = flux[j] / sig;
        } else {
            // non-positive flux: ignored by the magnitude fit (variance 1e50, :725) but a real band
            // of the flux-space fit; the reference magnitude for its model flux is mbar.
            sp.row[SR_BE + j] = std::pow(10., -0.4 * mbar) / sig;
        }
    }
    double* sc = sp.row + SR_SC;
    sc[SC_MBAR] = mbar;
    // No band with positive flux: the reference's magnitude fit then runs on weights of 1e-50 (:725), against which
    // the prior precisions dominate completely -- Av and Rv stay at the prior means to ~1e-38 and the loop stops after
    // one iteration.  With every weight 0 the 2x2 solves would be 0/0; S = 1 (weights still 0) gives exactly that
    // behaviour: zero steps, one iteration.
    sc[SC_S] = S > 0. ? S : 1.;
    // Exactly one band with positive flux: the reference's 2x2 systems are singular up to weights of 1e-50 (the
    // determinant S a - b^2 is pure rounding noise next to S/sigma_Av^2) and its magnitude-space step is whatever the
    // cancellation leaves -- in practice zero, one iteration.  That is made explicit here: the band's weight is
    // dropped from the magnitude fit as well (Av, Rv start the flux phase at the prior means).
    if (npos == 1) {
        for (int j = 0; j < nfilt; j++) sp.row[SR_U + j] = 0.;
        sc[SC_S] = 1.;
    }
    const bool have = std::isfinite(par) && std::isfinite(perr);  // :750-751
    sc[SC_PAR] = have ? par : 0.;
    sc[SC_PIVAR] = have ? 1. / (perr * perr) : 0.;
    const double k = ndim - 3;
    sc[SC_LNORM] = -(0.5 * k * std::log(2.) + std::lgamma(0.5 * k));  // brutus/utils.py:170
    sc[SC_KHM1] = 0.5 * k - 1.;
    sc[SC_GCONST] = -0.5 * (ndim * std::log(2. * M_PI) + sumlog);     // brutus/fitting.py:806-807
    const bool sp_apply = apply_clip && have && (par / perr > 4.);     // brutus/pdf.py:209
    sc[SC_SPAPPLY] = sp_apply ? 1. : 0.;
    if (sp_apply) {  // brutus/pdf.py:252-256
        double pm = par > 0. ? par : 0.;
        sc[SC_SMEAN] = pm * pm + perr * perr;
        sc[SC_SVAR] = 2 * perr * perr * perr * perr + 4 * pm * pm * perr * perr;
    }
    sc[SC_SLACK] = slack;
    sp.ndim = ndim;
}

template <typename T> struct Engine : EngineBase {
    using U = typename Enc<T>::U;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr;
    cudaEvent_t ev_rec[2] = {nullptr, nullptr}, ev_cp[2] = {nullptr, nullptr};
    char* ext_arena = nullptr;  // shared draw arena of a multi-device handle (not owned)
    size_t ext_ntot = 0, ext_off = 0;   // arena size in draws; this engine's first star
    void set_ext_arena(char* a, size_t ntot, size_t off) override { ext_arena = a; ext_ntot = ntot; ext_off = off; }
    const char* records_arena(int64_t* cap) override { *cap = arena_cap; return arena; }
    int nfilt_() const override { return nfilt; }
    int nlabel_() const override { return nlabel; }
    char* draw_arena = nullptr; // pinned host memory holding the posterior draws of the last bf_fit_batch
    size_t draw_cap = 0;        // in draws (nstar * ndraws)
    char* arena = nullptr;      // pinned host memory holding the records of the last bf_sweep_batch
    int64_t arena_cap = 0;
    int64_t nmodel = 0, npad = 0, nwords = 0;
    int nfilt = 0, nlabel = 0, rs = 0;
    int batch_cap = 0;
    int64_t pool_cap = 0;
    int max_flux = 1000;         // cap on flux-loop iterations (make_opts)
    const KTable<T>* kt = nullptr;

    DevBuf<float> d_grid, d_rows;
    DevBuf<T> d_labels, d_stars, d_ext, d_poolT;
    DevBuf<int> d_star_int, d_list, d_wpre, d_poolI, d_ctr, d_blk;
    DevBuf<uint32_t> d_cand;
    DevBuf<int64_t> d_ncand, d_base, d_tot;
    DevBuf<U> d_red, d_probe;
    DevBuf<double> d_out;
    DevBuf<char> d_flush, d_stage[2];
    // device posterior (bf_fit_batch)
    cudaEvent_t evP0 = nullptr, evP1 = nullptr;
    bool have_prior[3] = {false, false, false};   // lnprior, feh, loga staged?
    DevBuf<T> d_av_init, d_rv_init;               // per-model start of the magnitude fit (bf_set_init), [npad]
    bool have_init = false, use_init = false;     // staged? / in effect for the launches of the current call (B1 only)
    bool presweep_cfg = true, presweep = true;    // sweeps in two launches (see process_group; BRUTUS_B200_PRESWEEP=0 turns it
                                                  // off for A/B measurements); off for B1, where every model is a candidate
    DevBuf<T> d_lnprior, d_feh, d_loga, d_lnp1, d_lnp2, d_lnb1, d_keys, d_keys_sorted, d_clip;
    DevBuf<int> d_seg;
    DevBuf<char> d_cubtmp;
    DevBuf<GalStar<T>> d_gstar;
    DevBuf<int> d_rstar, d_nsel2, d_sel2, d_oidx, d_ord;
    DevBuf<int64_t> d_off2;
    DevBuf<double> d_cdf, d_ptot, d_odbl;
    PinVec<int> h_nsel2;

    // host staging, all in pinned memory: what the hot loop uploads (star rows, lists, per-star ints) is copied
    // by the DMA engine straight from these buffers, and the control read-backs land in them (k_publish)
    PinVec<T> h_stars, h_ext;
    PinVec<int> h_list, h_int, h_ctr;
    PinVec<int64_t> h_base, h_ncand;
    PinVec<U> h_red, h_probe;
    std::vector<int> h_kpred;

    // ---- optional per-kernel timing (BRUTUS_B200_TRACE=1): CUDA events around every launch, summed per kernel
    // name over a call; bf_get_trace() returns the table.  Off by default: recording events costs nothing on the
    // device but the bookkeeping is not free on the host.
    bool trace_on = false;
    std::vector<cudaEvent_t> trace_pool;
    struct TraceRec { const char* name; size_t e0, e1; };
    std::vector<TraceRec> trace_recs;
    std::map<std::string, std::pair<double, int64_t>> trace_acc;
    size_t trace_next = 0;
    cudaEvent_t trace_event() {
        if (trace_next == trace_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); trace_pool.push_back(e); }
        return trace_pool[trace_next++];
    }
    struct TraceScope {
        Engine* e; size_t i0 = 0; const char* name;
        TraceScope(Engine* eng, const char* nm) : e(eng), name(nm) {
            if (e->trace_on) { cudaEvent_t ev = e->trace_event(); i0 = e->trace_next - 1; cudaEventRecord(ev, e->stream); }
        }
        ~TraceScope() {
            if (e->trace_on) { cudaEvent_t ev = e->trace_event(); cudaEventRecord(ev, e->stream); e->trace_recs.push_back({name, i0, e->trace_next - 1}); }
        }
    };
    void trace_collect() {   // end of a call: the stream is idle
        if (!trace_on) return;
        cudaStreamSynchronize(stream);
        for (const TraceRec& r : trace_recs) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, trace_pool[r.e0], trace_pool[r.e1]) == cudaSuccess) {
                auto& a = trace_acc[r.name];
                a.first += ms; a.second += 1;
            }
        }
        trace_recs.clear();
        trace_next = 0;
    }
    std::string trace_text;
    const char* get_trace() override {
        trace_text.clear();
        char b[256];
        for (const auto& kv : trace_acc) {
            snprintf(b, sizeof b, "%-24s %8lld launches %12.3f ms\n", kv.first.c_str(), (long long)kv.second.second, kv.second.first);
            trace_text += b;
        }
        trace_acc.clear();
        return trace_text.c_str();
    }
#define TRACE(name) TraceScope trace_scope_(this, name)

    // device -> mapped pinned host, on the compute stream, without touching a copy engine
    void publish(void* host_dst, const void* dev_src, size_t bytes) {
        TRACE("k_publish");
        const size_t nw = bytes / 4;
        if (!nw) return;
        k_publish<<<(unsigned)((nw + 255) / 256), 256, 0, stream>>>((const uint32_t*)dev_src, (uint32_t*)host_dst, nw);
        stats.kernel_launches++;
    }

    // d_ctr: [0..1] records appended to the pool (64 bit), [2] fix-up list length, [3] length of the list of
    // active survivors (k_flux_more), [4..] "any star active" flags
    enum { CTR_POOL = 0, CTR_NFIX = 2, CTR_NLIST = 3, CTR_ANY = 4, CTR_COUNT = 16 };
    static constexpr int kShipBatch = 256;   // stars per sub-batch of the records-out path (D2H pipelining)

    ~Engine() override {
        cudaSetDevice(device);
        d_grid.release(); d_rows.release(); d_labels.release(); d_stars.release(); d_ext.release();
        d_poolT.release(); d_star_int.release(); d_list.release(); d_wpre.release(); d_poolI.release();
        d_ctr.release(); d_blk.release(); d_cand.release();
        d_probe.release(); d_ncand.release(); d_base.release(); d_tot.release(); d_red.release(); d_out.release(); d_flush.release();
        d_lnprior.release(); d_feh.release(); d_loga.release(); d_lnp1.release(); d_lnp2.release(); d_gstar.release();
        d_av_init.release(); d_rv_init.release();
        d_lnb1.release(); d_keys.release(); d_keys_sorted.release(); d_clip.release(); d_seg.release(); d_cubtmp.release();
        d_ord.release(); d_rstar.release(); d_nsel2.release(); d_sel2.release(); d_oidx.release(); d_off2.release(); d_cdf.release();
        d_ptot.release(); d_odbl.release(); h_nsel2.release();
        for (cudaEvent_t e : trace_pool) cudaEventDestroy(e);
        if (evP0) cudaEventDestroy(evP0);
        if (evP1) cudaEventDestroy(evP1);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (evA) cudaEventDestroy(evA);
        if (evB) cudaEventDestroy(evB);
        for (int k = 0; k < 2; k++) {
            d_stage[k].release();
            if (ev_rec[k]) cudaEventDestroy(ev_rec[k]);
            if (ev_cp[k]) cudaEventDestroy(ev_cp[k]);
        }
        if (arena) cudaFreeHost(arena);
        if (draw_arena) cudaFreeHost(draw_arena);
        h_stars.release(); h_ext.release(); h_list.release(); h_int.release(); h_ctr.release(); h_base.release();
        h_ncand.release(); h_red.release(); h_probe.release();
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }

    int init() {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
        CK(cudaEventCreate(&evA)); CK(cudaEventCreate(&evB));
        CK(cudaEventCreate(&evP0)); CK(cudaEventCreate(&evP1));
        CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CK(cudaEventCreateWithFlags(&ev_rec[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_cp[k], cudaEventDisableTiming));
        }
        if (const char* e = getenv("BRUTUS_B200_TRACE")) trace_on = atoi(e) != 0;
        if (const char* e = getenv("BRUTUS_B200_PRESWEEP")) presweep_cfg = presweep = atoi(e) != 0;
        CK(h_ctr.resize(CTR_COUNT));
        CK(d_ctr.ensure(CTR_COUNT));
        CK(d_tot.ensure(2));
        return BF_OK;
    }

    PoolArrays<T> pool() {
        PoolArrays<T> q;
        q.model = d_poolI.p; q.sflag = d_poolI.p + pool_cap;
        T* b = d_poolT.p;
        T** f[kPoolReals] = {&q.av, &q.rv, &q.chi2, &q.scale, &q.sden, &q.lp, &q.eta, &q.lold, &q.lprev,
                             &q.isa, &q.isr, &q.iaa, &q.iar, &q.irr, &q.lnl, &q.lnprob};
        for (int k = 0; k < kPoolReals; k++) *f[k] = b + (size_t)k * pool_cap;
        return q;
    }
    int* fixlist() { return d_poolI.p + (size_t)kPoolInts * pool_cap; }

    int set_grid(const float* co, int64_t nm, int nf, int layout, bool on_device) override {
        CK(cudaSetDevice(device));
        if (!co || nm <= 0 || nf <= 0) { err = "bf_set_grid: null grid or non-positive shape"; return BF_E_INVALID; }
        if (nf > kMaxFilt) { err = "bf_set_grid: nfilt exceeds BF_MAX_FILT (16)"; return BF_E_INVALID; }
        if (nm > (int64_t)1000000000) { err = "bf_set_grid: nmodel too large"; return BF_E_INVALID; }
        if (layout != BF_LAYOUT_C && layout != BF_LAYOUT_F) { err = "bf_set_grid: unknown layout"; return BF_E_INVALID; }
        kt = get_ktable<T>(nf);
        if (!kt) { err = "bf_set_grid: no kernels compiled for this band count"; return BF_E_INVALID; }
        nmodel = nm; nfilt = nf;
        npad = (nm + kTile - 1) / kTile * kTile;
        nwords = npad / 32;
        rs = row_stride(nf);
        nlabel = 0;
        have_prior[0] = have_prior[1] = have_prior[2] = false;
        have_init = false;
        const size_t nval = (size_t)nm * nf * 3;
        CK(d_grid.ensure((size_t)3 * nf * npad));
        CK(d_rows.ensure((size_t)rs * npad));
        const float* src = co;
        DevBuf<float> tmp;
        if (!on_device) {
            CK(tmp.ensure(nval));
            CK(cudaMemcpyAsync(tmp.p, co, nval * sizeof(float), cudaMemcpyHostToDevice, stream));
            src = tmp.p;
            stats.h2d_bytes += nval * sizeof(float);
        }
        k_retile<<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(src, d_grid.p, d_rows.p, nmodel, npad, nfilt, rs, layout);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        tmp.release();
        // star batch: bounded by the candidate maps (2 x 4 B per 32 models per star)
        size_t fr = 0, tot = 0;
        CK(cudaMemGetInfo(&fr, &tot));
        // Star batch.  Large batches amortise the host round trips of a batch (iteration-count verification,
        // survivor / selection counts) and the tails of its kernels: 1024 stars measured +3.5 % over 256.  The
        // records-out path works in sub-batches of kShipBatch so that the D2H of one sub-batch overlaps the
        // kernels of the next (run_catalogue's batch_limit).
        int want = 1024;
        if (const char* e = getenv("BRUTUS_B200_BATCH")) want = std::max(1, std::min(4096, atoi(e)));
        const size_t per_star = (size_t)8 * nwords;
        batch_cap = (int)std::max<size_t>(1, std::min<size_t>(want, (fr / 8) / per_star));
        // candidate pool: kPoolInts + 1 ints + kPoolReals T per record, at most ~1/4 of the free memory
        const size_t rec = (kPoolInts + 1) * sizeof(int) + kPoolReals * sizeof(T);
        int64_t cap = (int64_t)std::min<size_t>((fr / 4) / rec, (size_t)1 << 30);
        if (const char* e = getenv("BRUTUS_B200_POOL")) cap = std::max<int64_t>(1, atoll(e));
        // at least one star with every model as a candidate, swept twice (a stale set of records plus the live one)
        pool_cap = std::max<int64_t>(2 * npad, std::min<int64_t>(cap, (int64_t)batch_cap * npad + npad));
        CK(d_stars.ensure((size_t)batch_cap * kStarStride));
        CK(d_star_int.ensure((size_t)batch_cap * SI_COUNT));
        CK(d_list.ensure((size_t)batch_cap));
        CK(d_cand.ensure((size_t)batch_cap * nwords));
        CK(d_wpre.ensure((size_t)batch_cap * nwords));
        CK(d_ncand.ensure((size_t)batch_cap));
        CK(d_base.ensure((size_t)batch_cap));
        CK(d_red.ensure((size_t)batch_cap * kNumRed));
        CK(d_probe.ensure((size_t)batch_cap * 2 * kProbeIter));
        CK(d_poolI.ensure((size_t)(kPoolInts + 1) * pool_cap));
        CK(d_poolT.ensure((size_t)kPoolReals * pool_cap));
        CK(d_blk.ensure((size_t)(pool_cap / kTile + 2)));
        CK(h_stars.resize((size_t)batch_cap * kStarStride));
        CK(h_int.resize((size_t)batch_cap * SI_COUNT));
        CK(h_list.resize(batch_cap));
        h_kpred.assign(batch_cap, 0);
        CK(h_red.resize((size_t)batch_cap * kNumRed));
        CK(h_probe.resize((size_t)batch_cap * 2 * kProbeIter));
        CK(h_ncand.resize(batch_cap));
        CK(h_base.resize(batch_cap));
        return BF_OK;
    }

    int set_labels(const double* labels, int nl) override {
        CK(cudaSetDevice(device));
        if (!kt) { err = "bf_set_labels: call bf_set_grid first"; return BF_E_NOGRID; }
        if (nl < 0 || (nl > 0 && !labels)) { err = "bf_set_labels: bad arguments"; return BF_E_INVALID; }
        nlabel = nl;
        if (nl == 0) return BF_OK;
        DevBuf<double> tmp;
        CK(tmp.ensure((size_t)nl * nmodel));
        CK(cudaMemcpyAsync(tmp.p, labels, (size_t)nl * nmodel * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(d_labels.ensure((size_t)nl * npad));
        k_convert_labels<T><<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(tmp.p, d_labels.p, nmodel, npad, nl);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        tmp.release();
        CK(d_ext.ensure((size_t)batch_cap * nl * 3));
        CK(h_ext.resize((size_t)batch_cap * nl * 3));
        return BF_OK;
    }

    int flush_l2() override {
        CK(cudaSetDevice(device));
        const size_t n = (size_t)512 << 20;  // 4x the 126 MB L2
        CK(d_flush.ensure(n));
        CK(cudaMemsetAsync(d_flush.p, 0, n, stream));
        CK(cudaStreamSynchronize(stream));
        return BF_OK;
    }

    int make_opts(const bf_options* opt, DevOpts<T>& o, int& max_iter) {
        if (opt->init_thresh > opt->ltol_subthresh) {
            err = "The initial threshold must be smaller than or equal to the final threshold applied to be useful!";
            return BF_E_THRESH;
        }
        o.Abar = (T)opt->av_gauss[0]; o.PA = (T)(1. / (opt->av_gauss[1] * opt->av_gauss[1]));
        o.Rbar = (T)opt->rv_gauss[0]; o.PR = (T)(1. / (opt->rv_gauss[1] * opt->rv_gauss[1]));
        o.avmin = (T)opt->avlim[0]; o.avmax = (T)opt->avlim[1];
        o.rvmin = (T)opt->rvlim[0]; o.rvmax = (T)opt->rvlim[1];
        o.mtol = (T)(2.5 * opt->ltol);
        o.ln_init = (T)(opt->init_thresh > 0 ? std::log(opt->init_thresh) : -INFINITY);
        o.ltol = (T)opt->ltol;
        o.ln_sub = (T)std::log(opt->ltol_subthresh);
        o.ln_wt = (T)(opt->wt_thresh > 0 ? std::log(opt->wt_thresh) : -INFINITY);
        o.dim_prior = opt->dim_prior;
        // The reference's loops are unbounded (brutus/fitting.py:173-264, :781-803).  Caps, as a safety net only: 64 mag
        // iterations (each costs sweeps of the whole grid), 1000 flux iterations (cheap: they revisit a list of
        // survivors); opt->max_iter > 0 sets both.  Stars that hit a cap are counted in stats.unconverged.
        max_iter = opt->max_iter > 0 ? opt->max_iter : 64;
        max_flux = opt->max_iter > 0 ? opt->max_iter : 1000;
        return BF_OK;
    }

    // CTAs of a pass over the pool: exactly one wave -- what the device keeps resident of this kernel -- each CTA taking
    // a contiguous range of records (1.6 waves of a fixed 148 x 8 grid left the second wave's SMs idle for 20 % of the pass)
    template <typename K> unsigned pass_ctas(int64_t n, K kernel) {
        int per_sm = 0, sms = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTile, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        const int64_t wave = std::min<int64_t>((int64_t)sms * per_sm, kPassCtas);
        return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + kPassStep - 1) / kPassStep, wave));
    }

    void phase_begin() { cudaEventRecord(evA, stream); }
    cudaError_t sync() { stats.host_syncs++; return cudaStreamSynchronize(stream); }
    double phase_end() {  // synchronises the stream
        stats.host_syncs++;
        cudaEventRecord(evB, stream);
        cudaEventSynchronize(evB);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evA, evB);
        return ms;
    }

    // ---- phase 0: speculate the mag-iteration count of each star from a 1/32 subsample of the grid ----
    int probe_k(int ns, const DevOpts<T>& o, int max_iter) {
        const int64_t ntile = npad / kTile;
        if (ntile < 128 || max_iter < 2) return BF_OK;   // small grids: a re-sweep is cheaper than a probe
        ProbeParams<T> pp;
        pp.grid = d_grid.p; pp.npad = npad; pp.nmodel = nmodel; pp.stars = d_stars.p; pp.nstar = ns;
        pp.av_init = use_init ? d_av_init.p : nullptr; pp.rv_init = use_init ? d_rv_init.p : nullptr;
        pp.tile_stride = 32; pp.o = o; pp.out = d_probe.p;   // 1/32 of the grid: measured best (16: +0.65 ms of probe; 64: more re-sweeps)
        if (const char* e = getenv("BRUTUS_B200_PROBE_STRIDE")) pp.tile_stride = std::max(1, atoi(e));
        for (size_t k = 0; k < (size_t)ns * 2 * kProbeIter; k++) h_probe[k] = Enc<T>::enc(-std::numeric_limits<T>::infinity());
        CK(cudaMemcpyAsync(d_probe.p, h_probe.data(), (size_t)ns * 2 * kProbeIter * sizeof(U), cudaMemcpyHostToDevice, stream));
        phase_begin();
        { TRACE("k_kprobe"); kt->kprobe(pp, stream); }
        stats.kernel_launches++;
        CK(cudaGetLastError());
        publish(h_probe.data(), d_probe.p, (size_t)ns * 2 * kProbeIter * sizeof(U));
        stats.ms_select += phase_end();
        for (int s = 0; s < ns; s++) {
            int k = 1;
            for (; k <= kProbeIter; k++) {
                const U* r = &h_probe[((size_t)s * kProbeIter + (k - 1)) * 2];
                if (!(Enc<T>::dec(r[1]) > Enc<T>::dec(r[0]) + o.ln_init)) break;   // err < tol at iteration k
            }
            // The probe only chooses between 1 and 2 for the first full sweep.  A sweep with K iterations tells
            // whether the FULL grid converged at K-1 and at K, and the reference stops at the FIRST converged
            // iteration, so K is only ever raised two at a time from iterations already known not to have
            // converged (process_group): starting a star at K >= 3 on the word of a subsample would leave its
            // early iterations unverified.
            h_int[s * SI_COUNT + SI_KSPEC] = std::min(std::min(k, 2), max_iter);
            h_kpred[s] = k;   // the subsample's prediction steers how far a later sweep reaches (never past known + 2)
        }
        return BF_OK;
    }

    // ---- the stars (slots) [g0, g1) of the uploaded batch, from the sweep to the selection map ----------------
    // 1. the fused sweep (k_sweep) with verified speculation of the mag-iteration count: stars whose count was
    //    wrong are swept again (their earlier records go stale: SI_EPOCH);
    // 2. exact cull + convergence test of the flux iterations the sweep ran (k_cull, k_flux_ctl), the rare
    //    fix-ups and further flux iterations (k_fixup, k_flux_more);
    // 3. lnlike / lnprob / per-star maximum (k_final), first selection of lnpost (k_sel), and the scan of the
    //    selection map that orders the output (k_cand_scan).
    // *fits = false: the group's records did not fit in the pool; h_ncand[g0..g1) then hold the stars'
    // candidate counts for regrouping.  On success h_ncand hold the selection counts, *nrec the pool fill.
    int process_group(int ns, int g0, int g1, const DevOpts<T>& o, int max_iter, std::vector<char>& exact,
                      int* n_mag, bool* fits, int64_t* nrec) {
        const PoolArrays<T> pl = pool();
        const int ng = g1 - g0;
        const int nit_first = std::min(2, max_flux);
        bool redo_all = true;
        while (redo_all) {
            redo_all = false;
            for (int s = g0; s < g1; s++) {
                int* si = &h_int[s * SI_COUNT];
                si[SI_EPOCH] = 0; si[SI_NSURV] = 0; si[SI_ACTIVE] = 0; si[SI_NFLUX] = 0;
            }
            CK(cudaMemsetAsync(d_ctr.p, 0, CTR_COUNT * sizeof(int), stream));
            // ---- 1. sweep ----
            std::vector<int> lst(ng);
            for (int k = 0; k < ng; k++) lst[k] = g0 + k;
            bool first_pass = true;
            int64_t count = 0;
            while (!lst.empty()) {
                const int nl = (int)lst.size();
                for (int k = 0; k < nl; k++) h_list[k] = lst[k];
                CK(cudaMemcpyAsync(d_list.p, h_list.data(), (size_t)nl * sizeof(int), cudaMemcpyHostToDevice, stream));
                CK(cudaMemcpyAsync(d_star_int.p, h_int.data(), (size_t)ns * SI_COUNT * sizeof(int), cudaMemcpyHostToDevice, stream));
                { TRACE("k_reset_red"); k_reset_red<T><<<(nl * kNumRed + 255) / 256, 256, 0, stream>>>(d_red.p, d_list.p, nl, (1u << kNumRed) - 1u); }
                stats.kernel_launches++;
                if (!first_pass) stats.resweeps += nl;
                SweepParams<T> sp;
                sp.grid = d_grid.p; sp.npad = npad; sp.nmodel = nmodel; sp.stars = d_stars.p;
                sp.star_int = d_star_int.p; sp.list = d_list.p; sp.nlist = nl; sp.o = o; sp.red = d_red.p;
                sp.cand = d_cand.p; sp.nwords = nwords; sp.labels = d_labels.p; sp.ext = d_ext.p; sp.nlabel = nlabel;
                sp.pool = pl; sp.pool_cap = pool_cap; sp.pool_count = (unsigned long long*)(d_ctr.p + CTR_POOL);
                sp.nit_first = nit_first;
                sp.av_init = use_init ? d_av_init.p : nullptr; sp.rv_init = use_init ? d_rv_init.p : nullptr;
                // A candidate is flagged against the maxima known when its CTA starts, so the sweep is issued in two
                // launches: a strided subsample of the model tiles first, then the rest -- every CTA of the second
                // launch starts from the maxima of the first, whatever the order of the grid (an ordered lattice keeps
                // its best models in a few tiles).  On a small grid a wave of CTAs covers every tile of a star chunk
                // and nothing is published in time: there the grid is swept twice, first for the maxima alone
                // (cheap: the grid is L2-resident and the work is in the candidates' dense phase).  With every model a
                // candidate anyway (B1, a star redone after a fallback) one plain launch does.
                const int64_t ntile = npad / kTile;
                int nlaunch = 1;
                phase_begin();
                if (!presweep || ntile < 2) {
                    sp.tile_mode = 0; sp.tile_S = 1; sp.maxima_only = 0;
                    { TRACE("k_sweep"); CK((cudaError_t)kt->sweep(sp, stream)); }
                } else if (ntile < 1024) {
                    nlaunch = 2;
                    sp.tile_mode = 0; sp.tile_S = 1; sp.maxima_only = 1;
                    { TRACE("k_sweep_maxima"); CK((cudaError_t)kt->sweep(sp, stream)); }
                    sp.maxima_only = 0;
                    { TRACE("k_sweep"); CK((cudaError_t)kt->sweep(sp, stream)); }
                } else {
                    nlaunch = 2;
                    sp.tile_S = (int)std::min<int64_t>(32, std::max<int64_t>(2, ntile / 128)); sp.maxima_only = 0;
                    sp.tile_mode = 1;
                    { TRACE("k_sweep"); CK((cudaError_t)kt->sweep(sp, stream)); }
                    sp.tile_mode = 2;
                    { TRACE("k_sweep"); CK((cudaError_t)kt->sweep(sp, stream)); }
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(evB, stream));
                stats.kernel_launches += nlaunch; stats.magfit_launches += nlaunch; stats.magfit_star_passes += nl;
                publish(h_red.data(), d_red.p, (size_t)ns * kNumRed * sizeof(U));
                publish(h_ctr.data(), d_ctr.p, CTR_COUNT * sizeof(int));
                CK(sync());
                {
                    float ms = 0.f;
                    CK(cudaEventElapsedTime(&ms, evA, evB));
                    stats.ms_magfit += ms;
                }
                stats.d2h_bytes += (size_t)ns * kNumRed * sizeof(U) + CTR_COUNT * sizeof(int);
                std::memcpy(&count, &h_ctr[CTR_POOL], sizeof(int64_t));
                std::vector<int> next;
                for (int k = 0; k < nl; k++) {
                    const int s = lst[k];
                    int& ksp = h_int[s * SI_COUNT + SI_KSPEC];
                    if (n_mag) n_mag[s] = ksp;
                    if (exact[s]) continue;
                    const U* r = &h_red[(size_t)s * kNumRed];
                    // brutus/fitting.py:252-263 restated on the two max-reductions.  Invariant: every iteration below
                    // ksp - 1 is already known NOT to have converged on the full grid (first sweep at ksp <= 2, then
                    // +2 only after a sweep that saw ksp - 1 and ksp unconverged), so the first converged iteration
                    // among {ksp - 1, ksp} is the reference's stopping iteration.
                    const bool conv_prev = ksp >= 2 && !(Enc<T>::dec(r[RED_B0]) > Enc<T>::dec(r[RED_L0]) + o.ln_init);
                    const bool conv_last = !(Enc<T>::dec(r[RED_B1]) > Enc<T>::dec(r[RED_L1]) + o.ln_init);
                    if (conv_prev) {           // the reference would have stopped one iteration earlier
                        ksp -= 1; exact[s] = 1; next.push_back(s);
                    } else if (!conv_last && ksp < max_iter) {
                        // next sweep: iterations (ksp, ksp + 1) if the probe expects convergence at ksp + 1, else
                        // (ksp + 1, ksp + 2); either way contiguous with what is known
                        ksp = std::min(h_kpred[s] == ksp + 1 ? ksp + 1 : ksp + 2, max_iter); next.push_back(s);
                    } else {
                        if (!conv_last) stats.unconverged++;   // the cap on mag iterations stopped the loop
                        exact[s] = 1;
                    }
                }
                for (int s : next) h_int[s * SI_COUNT + SI_EPOCH] += 1;   // their records so far are stale
                lst.swap(next);
                first_pass = false;
            }
            for (int k = 0; k < ng; k++) h_list[k] = g0 + k;
            CK(cudaMemcpyAsync(d_list.p, h_list.data(), (size_t)ng * sizeof(int), cudaMemcpyHostToDevice, stream));
            if (count > pool_cap) {   // records were dropped: report the candidate counts, the caller splits the group
                { TRACE("k_cand_scan"); k_cand_scan<<<ng, 1024, 0, stream>>>(d_cand.p, d_wpre.p, d_list.p, nwords, d_ncand.p); }
                stats.kernel_launches++;
                CK(cudaGetLastError());
                publish(h_ncand.data() + g0, d_ncand.p + g0, (size_t)ng * sizeof(int64_t));
                CK(sync());
                *fits = false;
                return BF_OK;
            }
            const int64_t n = count;
            const unsigned nblk = (unsigned)((n + kTile - 1) / kTile);
            PassParams<T> pp{};
            pp.pool = pl; pp.n = n; pp.stars = d_stars.p; pp.star_int = d_star_int.p; pp.red = d_red.p; pp.o = o;
            pp.fixlist = fixlist(); pp.nfix = d_ctr.p + CTR_NFIX;
            pp.npad = npad; pp.labels = d_labels.p; pp.ext = d_ext.p; pp.nlabel = nlabel;
            pp.cand = d_cand.p; pp.nwords = nwords;
            // ---- 2. exact cull, flux-loop control ----
            phase_begin();
            { TRACE("k_reset_red"); k_reset_red<T><<<(ng * kNumRed + 255) / 256, 256, 0, stream>>>(d_red.p, d_list.p, ng, (1u << RED_FL) | (1u << RED_FB) | (1u << RED_LNP)); }
            if (n > 0) { TRACE("k_cull"); k_cull<T><<<pass_ctas(n, k_cull<T>), kTile, 0, stream>>>(pp); }
            { TRACE("k_flux_ctl"); k_flux_ctl<T><<<1, 1024, 0, stream>>>(d_star_int.p, d_red.p, d_list.p, ng, 1, nit_first, max_flux, o.ln_sub, d_ctr.p + CTR_ANY); }
            stats.kernel_launches += 3;
            CK(cudaGetLastError());
            publish(h_ctr.data(), d_ctr.p, CTR_COUNT * sizeof(int));
            stats.ms_select += phase_end();
            const int nfix = h_ctr[CTR_NFIX];
            int any = h_ctr[CTR_ANY];
            stats.candidates += n;
            stats.fixups += nfix;
            phase_begin();
            RecParams<T> rp{};
            rp.rows = d_rows.p; rp.stars = d_stars.p; rp.star_int = d_star_int.p; rp.o = o; rp.pool = pl; rp.red = d_red.p;
            rp.av_init = use_init ? d_av_init.p : nullptr; rp.rv_init = use_init ? d_rv_init.p : nullptr;
            if (nfix > 0) {
                rp.n = nfix; rp.list = fixlist();
                { TRACE("k_fixup"); kt->fixup(rp, stream); }
                stats.kernel_launches++;
                CK(cudaGetLastError());
            }
            // ---- further flux-space iterations for the stars that have not converged (:781-803); the per-star
            // convergence test runs on the device, the host only polls "anything still active?".  The first
            // extra iteration walks the whole pool and lists the active stars' survivors (in the memory of the
            // fix-up list, consumed by now); later ones visit that list only. ----
            int done_iter = nit_first;
            rp.n = n; rp.list = fixlist(); rp.nlist = d_ctr.p + CTR_NLIST;
            if (any && done_iter < max_flux) {
                { TRACE("k_flux_list"); k_flux_list<T><<<pass_ctas(n, k_flux_list<T>), kTile, 0, stream>>>(pp, fixlist(), d_ctr.p + CTR_NLIST); }
                stats.kernel_launches++;
            }
            while (any && done_iter < max_flux) {
                int any_slot = CTR_ANY;
                for (int r = 0; r < 3 && done_iter < max_flux; r++) {
                    { TRACE("k_flux_more"); kt->flux_more(rp, stream); }
                    any_slot = CTR_ANY + 1 + r;
                    { TRACE("k_flux_ctl"); k_flux_ctl<T><<<1, 1024, 0, stream>>>(d_star_int.p, d_red.p, d_list.p, ng, 0, 1, max_flux, o.ln_sub, d_ctr.p + any_slot); }
                    stats.kernel_launches += 2;
                    stats.flux_more_launches++;
                    done_iter += 1;
                }
                CK(cudaGetLastError());
                publish(h_ctr.data(), d_ctr.p, CTR_COUNT * sizeof(int));
                CK(sync());
                any = h_ctr[any_slot];
            }
            // ---- 3. lnlike / lnprob, per-star maximum, first selection, selection map scan ----
            if (n > 0) {
                { TRACE("k_final"); k_final<T><<<pass_ctas(n, k_final<T>), kTile, 0, stream>>>(pp); }
                { TRACE("k_sel"); k_sel<T><<<(unsigned)((n + kPassStep - 1) / kPassStep), kTile, 0, stream>>>(pp); }
                stats.kernel_launches += 2;
            }
            { TRACE("k_cand_scan"); k_cand_scan<<<ng, 1024, 0, stream>>>(d_cand.p, d_wpre.p, d_list.p, nwords, d_ncand.p); }
            stats.kernel_launches++;
            CK(cudaGetLastError());
            publish(h_ncand.data() + g0, d_ncand.p + g0, (size_t)ng * sizeof(int64_t));
            publish(h_red.data(), d_red.p, (size_t)ns * kNumRed * sizeof(U));
            publish(h_int.data(), d_star_int.p, (size_t)ns * SI_COUNT * sizeof(int));
            stats.ms_flux += phase_end();
            stats.d2h_bytes += (size_t)ns * (kNumRed * sizeof(U) + SI_COUNT * sizeof(int)) + ng * sizeof(int64_t);
            // ---- was the sweep's candidate set a superset of the selection?  It is whenever the final
            // max(lnprob) did not fall more than `slack` below the provisional one; otherwise redo the group
            // with every model of those stars as a candidate (rare: counted in stats.fallbacks) ----
            for (int s = g0; s < g1; s++) {
                const double sl = (double)h_stars[(size_t)s * kStarStride + SR_SC + SC_SLACK];
                if (!std::isfinite(sl)) continue;
                const double M = (double)Enc<T>::dec(h_red[(size_t)s * kNumRed + RED_LNP]);
                const double M0 = (double)Enc<T>::dec(h_red[(size_t)s * kNumRed + RED_M0]);
                if (!(M >= M0 - sl)) {
                    stats.fallbacks++;
                    h_stars[(size_t)s * kStarStride + SR_SC + SC_SLACK] = (T)INFINITY;
                    redo_all = true;
                }
            }
            if (redo_all) {
                CK(cudaMemcpyAsync(d_stars.p, h_stars.data(), (size_t)ns * kStarStride * sizeof(T), cudaMemcpyHostToDevice, stream));
                continue;
            }
            for (int s = g0; s < g1; s++) {
                stats.survivors += h_int[s * SI_COUNT + SI_NSURV];
                if (h_int[s * SI_COUNT + SI_NFLUX] >= max_flux && max_flux > 2) stats.unconverged++;   // stopped by the cap
            }
            *nrec = n;
        }
        *fits = true;
        return BF_OK;
    }

    int upload_stars(int ns) {
        CK(cudaMemcpyAsync(d_stars.p, h_stars.data(), (size_t)ns * kStarStride * sizeof(T), cudaMemcpyHostToDevice, stream));
        stats.h2d_bytes += (size_t)ns * (kStarStride * sizeof(T) + (SI_COUNT + 1) * sizeof(int));
        if (nlabel > 0) {
            CK(cudaMemcpyAsync(d_ext.p, h_ext.data(), (size_t)ns * nlabel * 3 * sizeof(T), cudaMemcpyHostToDevice, stream));
            stats.h2d_bytes += (size_t)ns * nlabel * 3 * sizeof(T);
        }
        return BF_OK;
    }

    int fill_rows(int ns, const double* flux, const double* errv, const uint8_t* mask, const double* par,
                  const double* perr, const double* ext_mean, const double* ext_std, int apply_clip, double slack,
                  int max_iter, int32_t* ndim_out, uint8_t* mask_out) {
        StarPrep sp;
        for (int s = 0; s < ns; s++) {
            prep_star(flux + (size_t)s * nfilt, errv + (size_t)s * nfilt, mask + (size_t)s * nfilt, nfilt,
                      par ? par[s] : NAN, perr ? perr[s] : NAN, apply_clip, slack, sp);
            for (int k = 0; k < kStarStride; k++) h_stars[(size_t)s * kStarStride + k] = (T)sp.row[k];
            if (sp.ndim < 4) {   // Ndim - 3 degrees of freedom (brutus/fitting.py:815); BruteForce refuses such objects (:1413-1420)
                err = "fewer than 4 bands of acceptable photometry: the fit is degenerate (brutus/fitting.py:1413-1420)";
                return BF_E_INVALID;
            }
            int* si = &h_int[s * SI_COUNT];
            for (int k = 0; k < SI_COUNT; k++) si[k] = 0;
            si[SI_NDIM] = sp.ndim;
            // initial speculation: 2 mag iterations (what the reference needs in the common case)
            si[SI_KSPEC] = std::min(2, max_iter);
            h_kpred[s] = 0;
            if (ndim_out) ndim_out[s] = sp.ndim;
            if (mask_out) std::memcpy(mask_out + (size_t)s * nfilt, sp.clean, nfilt);
            for (int l = 0; l < nlabel; l++) {
                T* x = &h_ext[((size_t)s * nlabel + l) * 3];
                double mu = ext_mean ? ext_mean[(size_t)s * nlabel + l] : NAN;
                double sd = ext_std ? ext_std[(size_t)s * nlabel + l] : NAN;
                if (std::isfinite(mu) && sd > 0.) {  // brutus/fitting.py:1999
                    x[0] = (T)mu; x[1] = (T)(1. / (sd * sd)); x[2] = (T)std::log(2. * M_PI * sd * sd);
                } else { x[0] = x[1] = x[2] = T(0); }
            }
        }
        return BF_OK;
    }

    int loglike_full(const double* flux, const double* errv, const uint8_t* mask, double par, double perr,
                     const bf_options* opt, double* lnl, double* chi2, double* scale, double* av, double* rv,
                     double* icov, uint8_t* mask_out, int64_t* diag) override {
        CK(cudaSetDevice(device));
        if (!kt) { err = "bf_loglike_full: no grid (call bf_set_grid)"; return BF_E_NOGRID; }
        if (!flux || !errv || !mask || !opt || !lnl || !chi2 || !scale || !av || !rv) { err = "bf_loglike_full: null argument"; return BF_E_INVALID; }
        stats = bf_stats{};
        DevOpts<T> o; int max_iter;
        int rc = make_opts(opt, o, max_iter);
        if (rc) return rc;
        const int saved_labels = nlabel;
        nlabel = 0;  // loglike itself applies no label priors
        int32_t nd;
        // slack = +inf: every model is a candidate, so the pool holds a record of every model
        rc = fill_rows(1, flux, errv, mask, &par, &perr, nullptr, nullptr, 0, INFINITY, max_iter, &nd, mask_out);
        if (rc) { nlabel = saved_labels; return rc; }
        rc = upload_stars(1);
        int nm = 0;
        std::vector<char> exact(1, 0);
        CK(cudaEventRecord(ev0, stream));
        use_init = have_init;   // av_init / rv_init of loglike (brutus/fitting.py:700-703): this entry point only
        presweep = false;
        if (!rc) rc = probe_k(1, o, max_iter);
        int64_t tot = 0;
        bool fits = true;
        if (!rc) rc = process_group(1, 0, 1, o, max_iter, exact, &nm, &fits, &tot);
        use_init = false;
        presweep = presweep_cfg;
        nlabel = saved_labels;
        if (rc) return rc;
        if (!fits) { err = "bf_loglike_full: candidate pool too small for one star"; return BF_E_NOMEM; }
        if (tot < nmodel) { err = "internal: full-length pass did not keep every model"; return BF_E_INVALID; }
        const size_t per = icov ? 14 : 5;
        CK(d_out.ensure((size_t)nmodel * per));
        OutParams<T, double> op{};
        op.pool = pool(); op.n = tot; op.star_int = d_star_int.p;
        double* b = d_out.p;
        op.o_lnl = b; op.o_chi2 = b + nmodel; op.o_scale = b + 2 * nmodel; op.o_av = b + 3 * nmodel; op.o_rv = b + 4 * nmodel;
        op.o_icov = icov ? b + 5 * nmodel : nullptr;
        phase_begin();
        { TRACE("k_out_full"); k_out_full<T><<<(unsigned)((tot + kTile - 1) / kTile), kTile, 0, stream>>>(op); }
        stats.kernel_launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev1, stream));
        stats.ms_select += phase_end();
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        stats.ms_device = ms;
        const size_t nb = (size_t)nmodel * sizeof(double);
        CK(cudaMemcpy(lnl, op.o_lnl, nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(chi2, op.o_chi2, nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(scale, op.o_scale, nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(av, op.o_av, nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rv, op.o_rv, nb, cudaMemcpyDeviceToHost));
        if (icov) CK(cudaMemcpy(icov, op.o_icov, nb * 9, cudaMemcpyDeviceToHost));
        stats.d2h_bytes += nb * per;
        if (diag) { diag[0] = nd; diag[1] = nm; diag[2] = h_int[SI_NFLUX]; diag[3] = h_int[SI_NSURV]; }
        trace_collect();
        return BF_OK;
    }

    // ---- library-owned pinned result arena: [11 rows of T x cap][idx int32 x cap] ----
    int ensure_arena(int64_t need, int64_t written) {
        if (need <= arena_cap) return BF_OK;
        int64_t ncap = std::max<int64_t>(std::max<int64_t>(need + need / 4, 2 * arena_cap), (int64_t)1 << 20);
        CK(cudaStreamSynchronize(copy_stream));
        char* na = nullptr;
        CK(cudaHostAlloc((void**)&na, (size_t)ncap * (sizeof(int) + 11 * sizeof(T)), cudaHostAllocPortable));
        if (arena && written > 0) {
            std::memcpy(na + (size_t)11 * ncap * sizeof(T), arena + (size_t)11 * arena_cap * sizeof(T),
                        (size_t)written * sizeof(int));
            for (int r = 0; r < 11; r++)
                std::memcpy(na + (size_t)r * ncap * sizeof(T), arena + (size_t)r * arena_cap * sizeof(T),
                            (size_t)written * sizeof(T));
        }
        if (arena) cudaFreeHost(arena);
        arena = na;
        arena_cap = ncap;
        return BF_OK;
    }

    // ---- the catalogue pipeline shared by bf_sweep_batch and bf_fit_batch: star batches -> groups of stars
    // whose candidate records fit in the pool -> process_group -> ordered records of the first selection in a
    // device staging buffer; `consume` then ships them (B2) or integrates the priors over them (device posterior).
    struct GroupCtx {
        int64_t s0;          // catalogue index of slot 0 of the batch
        int ns, g0, g1;      // stars in the batch; slots [g0, g1) of this group
        int64_t nsel_tot;    // records of the group (first selection)
        const int* ord;      // [nsel_tot] pool index of each selected record, in (star, model) order
        int buf;             // staging buffer used (records-out path only)
        T* rows;             // [11][nsel_tot]
        int* idx;            // [nsel_tot] model index
        DevOpts<T> o;
    };

    template <typename Consumer>
    int run_catalogue(int64_t nstar, const double* flux, const double* errv, const uint8_t* mask,
                      const double* par, const double* perr, const double* ext_mean, const double* ext_std,
                      const bf_options* opt, int record_rows, bool want_rows, int32_t* ndim, int32_t* n_iter,
                      int64_t* n_surv, double* max_lnprob, int64_t* offsets, int batch_limit, Consumer&& consume) {
        DevOpts<T> o; int max_iter;
        int rc = make_opts(opt, o, max_iter);
        if (rc) return rc;
        const double slack = opt->select_slack;
        std::vector<int> nm(batch_cap);
        std::vector<char> exact(batch_cap);
        int grp = 0;
        offsets[0] = 0;
        const int bstep = std::max(1, std::min(batch_cap, batch_limit));
        for (int64_t s0 = 0; s0 < nstar; s0 += bstep) {
            const int ns = (int)std::min<int64_t>(bstep, nstar - s0);
            rc = fill_rows(ns, flux + (size_t)s0 * nfilt, errv + (size_t)s0 * nfilt, mask + (size_t)s0 * nfilt,
                           par ? par + s0 : nullptr, perr ? perr + s0 : nullptr,
                           ext_mean ? ext_mean + (size_t)s0 * nlabel : nullptr,
                           ext_std ? ext_std + (size_t)s0 * nlabel : nullptr, opt->apply_parallax_clip, slack, max_iter,
                           ndim ? ndim + s0 : nullptr, nullptr);
            if (rc) return rc;
            rc = upload_stars(ns);
            if (rc) return rc;
            CK(cudaEventRecord(ev0, stream));
            for (int s = 0; s < ns; s++) exact[s] = 0;
            rc = probe_k(ns, o, max_iter);
            if (rc) return rc;
            // groups of consecutive stars whose candidate records fit in the pool together: the whole batch
            // first; a group that overflows is split by its (now known) candidate counts
            std::deque<std::pair<int, int>> groups;
            groups.emplace_back(0, ns);
            while (!groups.empty()) {
                const int g0 = groups.front().first, g1 = groups.front().second;
                groups.pop_front();
                const int ng = g1 - g0;
                int64_t nrec = 0;
                bool fits = true;
                rc = process_group(ns, g0, g1, o, max_iter, exact, nm.data(), &fits, &nrec);
                if (rc) return rc;
                if (!fits) {
                    if (ng == 1) { err = "internal: candidate pool too small for one star"; return BF_E_NOMEM; }
                    stats.regroups++;
                    // headroom: stale records of re-swept stars are gone now (the counts are exact), what remains
                    // is the run-to-run variation of the candidate superset
                    std::vector<std::pair<int, int>> parts;
                    int a = g0;
                    int64_t cnt = 0;
                    for (int s = g0; s < g1; s++) {
                        const int64_t need = h_ncand[s] + h_ncand[s] / 16 + 64;
                        if (s > a && cnt + need > pool_cap) { parts.emplace_back(a, s); a = s; cnt = 0; }
                        cnt += need;
                    }
                    parts.emplace_back(a, g1);
                    if (parts.size() == 1) { parts.clear(); parts.emplace_back(g0, g0 + ng / 2); parts.emplace_back(g0 + ng / 2, g1); }
                    for (size_t k = parts.size(); k-- > 0;) groups.push_front(parts[k]);
                    continue;
                }
                int64_t nsel_tot = 0;
                for (int s = g0; s < g1; s++) {
                    if (n_iter) { n_iter[2 * (s0 + s)] = nm[s]; n_iter[2 * (s0 + s) + 1] = h_int[s * SI_COUNT + SI_NFLUX]; }
                    if (n_surv) n_surv[s0 + s] = h_int[s * SI_COUNT + SI_NSURV];
                    if (max_lnprob) {
                        T v = Enc<T>::dec(h_red[(size_t)s * kNumRed + RED_LNP]);
                        max_lnprob[s0 + s] = (v <= Num<T>::kNegBig) ? -1e300 : (double)v;
                    }
                    h_base[s] = nsel_tot;
                    offsets[s0 + s + 1] = offsets[s0 + s] + h_ncand[s];
                    nsel_tot += h_ncand[s];
                }
                stats.selected += nsel_tot;
                // ---- ordered records of the selection into a device staging buffer ----
                GroupCtx gc{};
                gc.s0 = s0; gc.ns = ns; gc.g0 = g0; gc.g1 = g1; gc.nsel_tot = nsel_tot; gc.o = o;
                if (nsel_tot > 0) {
                    phase_begin();
                    CK(cudaMemcpyAsync(d_base.p + g0, h_base.data() + g0, (size_t)ng * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
                    CK(d_ord.ensure((size_t)nsel_tot));
                    OutParams<T, T> op{};
                    op.pool = pool(); op.n = nrec; op.star_int = d_star_int.p;
                    op.sel = d_cand.p; op.wpre = d_wpre.p; op.nwords = nwords; op.base = d_base.p;
                    op.ord = d_ord.p; op.nsel = nsel_tot;
                    { TRACE("k_ord"); k_ord<T><<<(unsigned)((nrec + kPassStep - 1) / kPassStep), kTile, 0, stream>>>(op); }
                    stats.kernel_launches++;
                    gc.ord = d_ord.p;
                    if (want_rows) {   // records-out: the ordered [11][nsel_tot] matrix + model indices in a staging buffer
                        const int buf = grp & 1;
                        grp++;
                        CK(cudaStreamWaitEvent(stream, ev_cp[buf], 0));  // staging buffer free again?
                        CK(d_stage[buf].ensure((size_t)nsel_tot * (sizeof(int) + 11 * sizeof(T))));
                        T* b = (T*)d_stage[buf].p;                                   // [11][nsel_tot] rows, then idx
                        op.o_idx = (int*)(d_stage[buf].p + (size_t)11 * nsel_tot * sizeof(T));
                        op.o_star = nullptr;
                        op.ld = nsel_tot; op.nrows = record_rows;
                        op.o_lnl = b; op.o_scale = b + nsel_tot; op.o_av = b + 2 * nsel_tot; op.o_chi2 = b + 3 * nsel_tot;
                        op.o_rv = b + 4 * nsel_tot; op.o_icov = b + 5 * nsel_tot;
                        { TRACE("k_out"); k_out<T><<<(unsigned)((nsel_tot + kTile - 1) / kTile), kTile, 0, stream>>>(op); }
                        stats.kernel_launches++;
                        CK(cudaEventRecord(ev_rec[buf], stream));
                        gc.buf = buf; gc.rows = b; gc.idx = op.o_idx;
                    }
                    CK(cudaGetLastError());
                    CK(cudaEventRecord(evB, stream));
                }
                rc = consume(gc);
                if (rc) return rc;
                if (nsel_tot > 0) {
                    CK(cudaEventSynchronize(evB));
                    float msr = 0.f;
                    CK(cudaEventElapsedTime(&msr, evA, evB));
                    stats.ms_select += msr;
                }
            }
            CK(cudaEventRecord(ev1, stream));
            CK(cudaEventSynchronize(ev1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev0, ev1));
            stats.ms_device += ms;
        }
        return BF_OK;
    }

    int sweep_batch(int64_t nstar, const double* flux, const double* errv, const uint8_t* mask,
                    const double* par, const double* perr, const double* ext_mean, const double* ext_std,
                    const bf_options* opt, int record_rows, int32_t* ndim, int32_t* n_iter, int64_t* n_surv,
                    double* max_lnprob, int64_t* offsets, bf_records* out) override {
        CK(cudaSetDevice(device));
        if (!kt) { err = "bf_sweep_batch: no grid (call bf_set_grid)"; return BF_E_NOGRID; }
        if (nstar < 0 || (nstar > 0 && (!flux || !errv || !mask)) || !opt || !offsets || !out) { err = "bf_sweep_batch: null argument"; return BF_E_INVALID; }
        if (record_rows != 3 && record_rows != 5 && record_rows != 11) { err = "bf_sweep_batch: record_rows must be 3, 5 or 11"; return BF_E_INVALID; }
        stats = bf_stats{};
        int64_t written = 0;
        // asynchronous D2H of each group's records on the copy stream: overlaps the next group's / batch's kernels
        auto ship = [&](GroupCtx& g) -> int {
            if (g.nsel_tot > 0 && !opt->skip_d2h) {
                int rc = ensure_arena(written + g.nsel_tot, written);
                if (rc) return rc;
                CK(cudaStreamWaitEvent(copy_stream, ev_rec[g.buf], 0));
                CK(cudaMemcpyAsync(arena + (size_t)11 * arena_cap * sizeof(T) + (size_t)written * sizeof(int),
                                   g.idx, (size_t)g.nsel_tot * sizeof(int), cudaMemcpyDeviceToHost, copy_stream));
                // one plain 1-D copy per record row (each 100+ MB): the pitched 2-D copy ran at
                // ~42 GB/s on the PCIe Gen5 link, 1-D copies reach the measured ~57 GB/s
                for (int r = 0; r < record_rows; r++)
                    CK(cudaMemcpyAsync(arena + ((size_t)r * arena_cap + (size_t)written) * sizeof(T),
                                       g.rows + (size_t)r * g.nsel_tot, (size_t)g.nsel_tot * sizeof(T),
                                       cudaMemcpyDeviceToHost, copy_stream));
                CK(cudaEventRecord(ev_cp[g.buf], copy_stream));
                stats.d2h_bytes += (size_t)g.nsel_tot * (sizeof(int) + record_rows * sizeof(T));
            }
            written += g.nsel_tot;
            return BF_OK;
        };
        int ship_batch = kShipBatch;
        if (const char* e = getenv("BRUTUS_B200_SHIP_BATCH")) ship_batch = std::max(1, atoi(e));
        int rc = run_catalogue(nstar, flux, errv, mask, par, perr, ext_mean, ext_std, opt, record_rows, true, ndim,
                               n_iter, n_surv, max_lnprob, offsets, opt->skip_d2h ? batch_cap : ship_batch, ship);
        if (rc) return rc;
        CK(cudaStreamSynchronize(copy_stream));
        trace_collect();
        out->n = opt->skip_d2h ? 0 : written;
        out->stride = arena_cap;
        out->elem_size = (int32_t)sizeof(T);
        out->nrows = record_rows;
        out->model_idx = arena ? (const int32_t*)(arena + (size_t)11 * arena_cap * sizeof(T)) : nullptr;
        out->rows = (const void*)arena;
        return BF_OK;
    }

    // ---- get_seds on the staged grid (bf_get_seds) ----
    int get_seds(int64_t n, const int32_t* idx, const double* av, const double* rv, int flux, double* seds,
                 double* rvecs, double* drvecs) override {
        CK(cudaSetDevice(device));
        if (!kt) { err = "bf_get_seds: no grid (call bf_set_grid)"; return BF_E_NOGRID; }
        if (n < 0 || (n > 0 && (!av || !rv || !seds))) { err = "bf_get_seds: null argument"; return BF_E_INVALID; }
        if (!idx && n != nmodel) { err = "bf_get_seds: without idx, n must equal nmodel"; return BF_E_INVALID; }
        if (n == 0) return BF_OK;
        if (idx)
            for (int64_t i = 0; i < n; i++)
                if (idx[i] < 0 || idx[i] >= nmodel) { err = "bf_get_seds: model index out of range"; return BF_E_INVALID; }
        const int nout = 1 + (rvecs ? 1 : 0) + (drvecs ? 1 : 0);
        DevBuf<double> d_in, d_o;
        DevBuf<int> d_ix;
        CK(d_in.ensure((size_t)2 * n));
        CK(d_o.ensure((size_t)nout * n * nfilt));
        CK(cudaMemcpyAsync(d_in.p, av, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_in.p + n, rv, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (idx) {
            CK(d_ix.ensure((size_t)n));
            CK(cudaMemcpyAsync(d_ix.p, idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        double* o_s = d_o.p;
        double* o_r = rvecs ? d_o.p + (size_t)n * nfilt : nullptr;
        double* o_d = drvecs ? d_o.p + (size_t)(nout - 1) * n * nfilt : nullptr;
        const int64_t tot = n * nfilt;
        k_get_seds<T><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(d_rows.p, rs, nfilt, n, idx ? d_ix.p : nullptr, d_in.p,
                                                                       d_in.p + n, flux, o_s, o_r, o_d);
        CK(cudaGetLastError());
        stats.kernel_launches++;
        CK(cudaMemcpyAsync(seds, o_s, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (rvecs) CK(cudaMemcpyAsync(rvecs, o_r, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (drvecs) CK(cudaMemcpyAsync(drvecs, o_d, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(sync());
        d_in.release(); d_o.release(); d_ix.release();
        return BF_OK;
    }

    // ---- photometric_offsets: SEDs of the posterior samples and the leave-one-band-out weights (bf_offsets_weights) ----
    int offsets_weights(int64_t nobj, int nsamps, const double* phot, const double* errv, const uint8_t* mask,
                        const int32_t* idxs, const double* reds, const double* dreds, const double* dists,
                        const double* old_offsets, const uint8_t* mask_fit, int dim_prior, double* seds,
                        double* wt) override {
        CK(cudaSetDevice(device));
        if (!kt) { err = "bf_offsets_weights: no grid (call bf_set_grid)"; return BF_E_NOGRID; }
        if (nobj < 0 || nsamps < 1) { err = "bf_offsets_weights: nobj must be >= 0 and nsamps >= 1"; return BF_E_INVALID; }
        if (nobj == 0) return BF_OK;
        if (!phot || !errv || !mask || !idxs || !reds || !dreds || !dists || !mask_fit || !seds || !wt) {
            err = "bf_offsets_weights: null argument"; return BF_E_INVALID;
        }
        const int64_t n = nobj * nsamps;
        for (int64_t i = 0; i < n; i++)
            if (idxs[i] < 0 || idxs[i] >= nmodel) { err = "bf_offsets_weights: model index out of range"; return BF_E_INVALID; }
        // per-object inputs with the previous offsets applied (brutus/utils.py:1303-1304)
        std::vector<double> hp((size_t)nobj * nfilt), hv((size_t)nobj * nfilt);
        for (int64_t o = 0; o < nobj; o++)
            for (int j = 0; j < nfilt; j++) {
                const double f = old_offsets ? old_offsets[j] : 1.0;
                const double e = errv[o * nfilt + j] * f;
                hp[(size_t)o * nfilt + j] = phot[o * nfilt + j] * f;
                hv[(size_t)o * nfilt + j] = e * e;
            }
        DevBuf<double> d_in, d_obj, d_o;
        DevBuf<int> d_ix;
        DevBuf<uint8_t> d_mk;
        CK(d_in.ensure((size_t)3 * n));
        CK(d_obj.ensure((size_t)2 * nobj * nfilt));
        CK(d_ix.ensure((size_t)n));
        CK(d_mk.ensure((size_t)nobj * nfilt + nfilt));
        CK(d_o.ensure((size_t)n * nfilt * 3));   // seds [n][nfilt], lnl [nfilt][n], wt [nfilt][n]
        CK(cudaMemcpyAsync(d_in.p, reds, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_in.p + n, dreds, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_in.p + 2 * n, dists, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_obj.p, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_obj.p + hp.size(), hv.data(), hv.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_ix.p, idxs, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_mk.p, mask, (size_t)nobj * nfilt, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_mk.p + (size_t)nobj * nfilt, mask_fit, (size_t)nfilt, cudaMemcpyHostToDevice, stream));
        double* o_s = d_o.p;
        double* o_l = d_o.p + (size_t)n * nfilt;
        double* o_w = d_o.p + (size_t)2 * n * nfilt;
        const uint8_t* d_mfit = d_mk.p + (size_t)nobj * nfilt;
        CK(cudaMemsetAsync(o_w, 0, (size_t)n * nfilt * sizeof(double), stream));   // bands that are not fitted: zeros
        k_offsets_lnl<T><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_rows.p, rs, nfilt, nobj, nsamps, d_ix.p, d_in.p,
                                                                        d_in.p + n, d_in.p + 2 * n, d_obj.p,
                                                                        d_obj.p + hp.size(), d_mk.p, d_mfit, dim_prior, o_s, o_l);
        const int64_t nwarp = (int64_t)nfilt * nobj;
        k_offsets_wt<<<(unsigned)((nwarp * 32 + 255) / 256), 256, 0, stream>>>(o_l, d_mfit, nfilt, nobj, nsamps, o_w);
        CK(cudaGetLastError());
        stats.kernel_launches += 2;
        CK(cudaMemcpyAsync(seds, o_s, (size_t)n * nfilt * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(wt, o_w, (size_t)n * nfilt * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(sync());
        stats.h2d_bytes += (size_t)n * (3 * sizeof(double) + sizeof(int)) + (size_t)nobj * nfilt * 17;
        stats.d2h_bytes += (size_t)2 * n * nfilt * sizeof(double);
        return BF_OK;
    }

    // ---- per-model start of the magnitude fit: av_init / rv_init of loglike (bf_set_init) ----
    int set_init(const double* av_init, const double* rv_init) override {
        CK(cudaSetDevice(device));
        if (!av_init && !rv_init) { have_init = false; return BF_OK; }
        if (!kt) { err = "bf_set_init: call bf_set_grid first"; return BF_E_NOGRID; }
        if (!av_init || !rv_init) { err = "bf_set_init: give both av_init and rv_init, or neither"; return BF_E_INVALID; }
        for (int64_t i = 0; i < nmodel; i++)
            if (!std::isfinite(av_init[i]) || !std::isfinite(rv_init[i])) { err = "bf_set_init: non-finite initial value"; return BF_E_INVALID; }
        const double* src[2] = {av_init, rv_init};
        DevBuf<T>* dst[2] = {&d_av_init, &d_rv_init};
        DevBuf<double> tmp;
        CK(tmp.ensure((size_t)nmodel));
        for (int k = 0; k < 2; k++) {
            CK(cudaMemcpyAsync(tmp.p, src[k], (size_t)nmodel * sizeof(double), cudaMemcpyHostToDevice, stream));
            CK(dst[k]->ensure((size_t)npad));
            k_convert_labels<T><<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(tmp.p, dst[k]->p, nmodel, npad, 1);
            CK(cudaGetLastError());
            CK(sync());
        }
        tmp.release();
        have_init = true;
        return BF_OK;
    }

    // ---- static per-model priors / labels of lnpost (bf_set_model_priors) ----
    int set_model_priors(const double* lnprior, const double* feh, const double* loga) override {
        CK(cudaSetDevice(device));
        if (!kt) { err = "bf_set_model_priors: call bf_set_grid first"; return BF_E_NOGRID; }
        const double* src[3] = {lnprior, feh, loga};
        DevBuf<T>* dst[3] = {&d_lnprior, &d_feh, &d_loga};
        DevBuf<double> tmp;
        for (int k = 0; k < 3; k++) {
            have_prior[k] = src[k] != nullptr;
            if (!src[k]) continue;
            CK(tmp.ensure((size_t)nmodel));
            CK(cudaMemcpyAsync(tmp.p, src[k], (size_t)nmodel * sizeof(double), cudaMemcpyHostToDevice, stream));
            CK(dst[k]->ensure((size_t)npad));
            k_convert_labels<T><<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(tmp.p, dst[k]->p, nmodel, npad, 1);
            CK(cudaGetLastError());
            CK(sync());
            stats.h2d_bytes += (size_t)nmodel * sizeof(double);
        }
        tmp.release();
        return BF_OK;
    }

    // constants of the Galactic prior (brutus/pdf.py:476-749) in the form posterior.cuh consumes
    GalDev<T> make_gal(const bf_post_options* po) const {
        GalDev<T> G{};
        const bf_gal_params& g = po->gal;
        G.use = po->use_gal_prior ? 1 : 0;
        G.has_feh = have_prior[1] ? 1 : 0;
        G.has_age = have_prior[2] ? 1 : 0;
        G.same_rs = g.Rs_thin == g.Rs_thick ? 1 : 0;
        G.Rs_thin2 = (T)(g.Rs_thin * g.Rs_thin); G.Rs_thick2 = (T)(g.Rs_thick * g.Rs_thick); G.Rs_halo2 = (T)(g.Rs_halo * g.Rs_halo);
        G.R_solar = (T)g.R_solar; G.aZ_solar = (T)std::fabs(g.Z_solar);
        G.iR_thin = (T)(1. / g.R_thin); G.iZ_thin = (T)(1. / g.Z_thin);
        G.iR_thick = (T)(1. / g.R_thick); G.iZ_thick = (T)(1. / g.Z_thick);
        G.ln_f_thick = (T)std::log(g.f_thick); G.ln_f_halo = (T)std::log(g.f_halo);
        G.rq2 = (T)(g.r_q_halo * g.r_q_halo); G.irq = (T)(1. / g.r_q_halo);
        G.q_inf = (T)g.q_halo_inf; G.dq = (T)(g.q_halo_inf - g.q_halo_ctr); G.eta = (T)g.eta_halo;
        const double rp_s = std::sqrt(g.R_solar * g.R_solar + g.Z_solar * g.Z_solar + g.r_q_halo * g.r_q_halo);
        const double q_s = g.q_halo_inf - (g.q_halo_inf - g.q_halo_ctr) * std::exp(1. - rp_s / g.r_q_halo);
        G.ln_Reff_solar = (T)(0.5 * std::log(g.R_solar * g.R_solar + (g.Z_solar / q_s) * (g.Z_solar / q_s) + g.Rs_halo * g.Rs_halo));
        const double mu[3] = {g.feh_thin, g.feh_thick, g.feh_halo};
        const double sg[3] = {g.feh_thin_sigma, g.feh_thick_sigma, g.feh_halo_sigma};
        for (int x = 0; x < 3; x++) {
            G.feh_mu[x] = (T)mu[x];
            G.feh_isig2[x] = (T)(1. / (sg[x] * sg[x]));
            G.feh_lnorm[x] = (T)std::log(2. * M_PI * sg[x] * sg[x]);
            // age prior of the component: truncated normal on [min_age, max_age] (brutus/pdf.py:455-470)
            const double am = (g.max_age - g.min_age) / (1. + std::exp((mu[x] - g.feh_age_ctr) / g.feh_age_scale)) + g.min_age;
            double as = (g.max_age - am) / g.nsigma_from_max_age;
            as = std::min(std::max(as, g.min_sigma), g.max_sigma);
            const double a = (g.min_age - am) / as, b = (g.max_age - am) / as;
            G.age_mu[x] = (T)am;
            G.age_isig[x] = (T)(1. / as);
            G.age_lnden[x] = (T)(std::log(as / 2.) + std::log(std::erf(b / std::sqrt(2.)) - std::erf(a / std::sqrt(2.))));
        }
        G.min_age = (T)g.min_age; G.max_age = (T)g.max_age;
        const double l2e = 1.4426950408889634;
        G.l2_iR_thin = (T)(-l2e / g.R_thin); G.l2_iZ_thin = (T)(-l2e / g.Z_thin);
        G.l2_iR_thick = (T)(-l2e / g.R_thick); G.l2_iZ_thick = (T)(-l2e / g.Z_thick);
        G.l2_ln_f_thick = (T)(l2e * std::log(g.f_thick)); G.l2_irq = (T)(-l2e / g.r_q_halo);
        G.l2_halo = (T)(l2e * ((double)G.eta * (double)G.ln_Reff_solar + (double)G.ln_f_halo));
        G.nh_eta = (T)(-0.5 * g.eta_halo);
        return G;
    }

    // ---- the per-star body of BruteForce._fit on the device (brutus/fitting.py:1980-2061) ----
    int fit_batch(int64_t nstar, const double* flux, const double* errv, const uint8_t* mask, const double* par,
                  const double* perr, const double* coords, const double* ext_mean, const double* ext_std,
                  const bf_options* opt, const bf_post_options* po, int32_t* ndim, int32_t* n_iter, int64_t* nsel,
                  double* levid, double* chi2min, bf_draws* out) override {
        CK(cudaSetDevice(device));
        if (!kt) { err = "bf_fit_batch: no grid (call bf_set_grid)"; return BF_E_NOGRID; }
        if (nstar < 0 || (nstar > 0 && (!flux || !errv || !mask)) || !opt || !po || !out || !levid || !chi2min) { err = "bf_fit_batch: null argument"; return BF_E_INVALID; }
        if (po->nmc_prior < 1 || po->ndraws < 1) { err = "bf_fit_batch: nmc_prior and ndraws must be >= 1"; return BF_E_INVALID; }
        if (po->use_gal_prior && !coords) { err = "`coord` must be provided if using the default Galactic model prior."; return BF_E_INVALID; }
        // pinned arena for the draws: [8 double arrays | cov 9 doubles | idx int32] x ntot draws; owned by this
        // engine, or the shared arena of a multi-device handle (this engine then fills its stars' slice)
        size_t ntot = (size_t)std::max<int64_t>(nstar, 1) * po->ndraws, aoff = 0;
        char* abase = nullptr;
        if (ext_arena) {
            abase = ext_arena; ntot = ext_ntot; aoff = ext_off;
        } else {
            if (ntot > draw_cap) {
                if (draw_arena) cudaFreeHost(draw_arena);
                draw_arena = nullptr; draw_cap = 0;
                CK(cudaHostAlloc((void**)&draw_arena, ntot * (17 * sizeof(double) + sizeof(int32_t)) + 64, cudaHostAllocPortable));
                draw_cap = ntot;
            }
            abase = draw_arena;
        }
        double* const hd = (double*)abase;                     // 8 arrays, then cov
        int32_t* const hidx = (int32_t*)(abase + ntot * 17 * sizeof(double));
        stats = bf_stats{};
        const int nd = po->ndraws, nmc = po->nmc_prior;
        const GalDev<T> G = make_gal(po);
        // test hooks: host-supplied normals / uniforms
        DevBuf<double> d_zov, d_uov;
        if (po->z_override) {
            CK(d_zov.ensure((size_t)nmodel * 3 * nmc));
            CK(cudaMemcpyAsync(d_zov.p, po->z_override, (size_t)nmodel * 3 * nmc * sizeof(double), cudaMemcpyHostToDevice, stream));
        }
        CK(d_gstar.ensure((size_t)batch_cap));
        CK(d_nsel2.ensure((size_t)batch_cap));
        CK(d_off2.ensure((size_t)batch_cap + 1));
        CK(d_ptot.ensure((size_t)batch_cap));
        CK(d_oidx.ensure((size_t)batch_cap * nd));
        CK(d_odbl.ensure((size_t)batch_cap * nd * 17 + 2 * (size_t)batch_cap));
        CK(h_nsel2.resize(batch_cap));
        if (po->u_override) CK(d_uov.ensure((size_t)batch_cap * 2 * nd));
        std::vector<GalStar<T>> h_gstar(batch_cap);
        std::vector<int64_t> h_off2(batch_cap + 1);
        std::vector<int64_t> offsets(nstar + 1);
        int64_t last_s0 = -1;
        auto post = [&](GroupCtx& g) -> int {
            const int ng = g.g1 - g.g0;
            if (g.s0 != last_s0) {   // first group of a batch: per-star geometry, random-number overrides
                last_s0 = g.s0;
                const double st = po->gal.z_sun / po->gal.galcen_distance, ct = std::sqrt(1. - st * st);
                for (int s = 0; s < g.ns; s++) {
                    GalStar<T> q{};
                    if (coords) {
                        const double l = coords[2 * (g.s0 + s)] * M_PI / 180., b = coords[2 * (g.s0 + s) + 1] * M_PI / 180.;
                        const double cl = std::cos(l), sl = std::sin(l), cb = std::cos(b), sb = std::sin(b);
                        q.ax = (T)(cb * cl * ct + sb * st); q.ay = (T)(cb * sl); q.az = (T)(sb * ct - cb * cl * st);
                        q.x0 = (T)(-po->gal.galcen_distance * ct); q.z0 = (T)(po->gal.galcen_distance * st);
                    }
                    h_gstar[s] = q;
                }
                CK(cudaMemcpyAsync(d_gstar.p, h_gstar.data(), (size_t)g.ns * sizeof(GalStar<T>), cudaMemcpyHostToDevice, stream));
                if (po->u_override)
                    CK(cudaMemcpyAsync(d_uov.p, po->u_override + (size_t)g.s0 * 2 * nd, (size_t)g.ns * 2 * nd * sizeof(double), cudaMemcpyHostToDevice, stream));
            }
            const int64_t n1 = g.nsel_tot;
            PostParams<T> pp{};
            if (n1 > 0) CK(d_rstar.ensure((size_t)n1));
            pp.ord = g.ord; pp.pool = pool(); pp.rstar = d_rstar.p; pp.n1 = n1;
            pp.lnprior = have_prior[0] ? d_lnprior.p : nullptr;
            pp.feh = have_prior[1] ? d_feh.p : nullptr;
            pp.loga = have_prior[2] ? d_loga.p : nullptr;
            pp.G = G; pp.gstar = d_gstar.p; pp.stars = d_stars.p; pp.red = d_red.p; pp.ln_wt = g.o.ln_wt;
            pp.avmin = g.o.avmin; pp.avmax = g.o.avmax; pp.rvmin = g.o.rvmin; pp.rvmax = g.o.rvmax;
            pp.nmc = nmc; pp.ndraws = nd; pp.seed = po->seed; pp.star_base = po->star_base + g.s0;
            pp.zov = po->z_override ? d_zov.p : nullptr;
            pp.uov = po->u_override ? d_uov.p : nullptr;
            pp.g0 = g.g0;
            pp.nsel2 = d_nsel2.p; pp.off2 = d_off2.p; pp.tot = d_ptot.p; pp.blk = d_blk.p;
            double* ob = d_odbl.p;
            const size_t per = (size_t)batch_cap * nd;
            pp.o_idx = d_oidx.p;
            pp.o_scale = ob; pp.o_av = ob + per; pp.o_rv = ob + 2 * per; pp.o_lnprob = ob + 3 * per; pp.o_dist = ob + 4 * per;
            pp.o_red = ob + 5 * per; pp.o_dred = ob + 6 * per; pp.o_logwt = ob + 7 * per; pp.o_cov = ob + 8 * per;
            pp.o_levid = ob + 17 * per; pp.o_chi2min = pp.o_levid + batch_cap;
            CK(cudaEventRecord(evP0, stream));
            CK(cudaMemsetAsync(d_nsel2.p + g.g0, 0, (size_t)ng * sizeof(int), stream));
            int64_t n2 = 0;
            if (n1 > 0) {
                CK(d_lnp1.ensure((size_t)n1)); CK(d_lnp2.ensure((size_t)n1)); CK(d_sel2.ensure((size_t)n1)); CK(d_cdf.ensure((size_t)n1));
                CK(d_lnb1.ensure((size_t)n1));
                pp.lnp1 = d_lnp1.p; pp.lnp2 = d_lnp2.p; pp.sel2 = d_sel2.p; pp.cdf = d_cdf.p; pp.lnb1 = d_lnb1.p;
                const unsigned nb1 = (unsigned)((n1 + kTile - 1) / kTile);
                { TRACE("k_post_mle"); k_post_mle<T><<<nb1, kTile, 0, stream>>>(pp); }
                { TRACE("k_post_count"); k_post_count<T><<<nb1, kTile, 0, stream>>>(pp); }
                { TRACE("k_scan_blocks"); k_scan_blocks<<<1, 1024, 0, stream>>>(d_blk.p, nb1, d_tot.p); }
                stats.kernel_launches += 3;
                CK(cudaGetLastError());
            }
            publish(h_nsel2.data() + g.g0, d_nsel2.p + g.g0, (size_t)ng * sizeof(int));
            CK(sync());
            h_off2[g.g0] = 0;
            for (int s = g.g0; s < g.g1; s++) {
                h_off2[s + 1] = h_off2[s] + h_nsel2[s];
                if (nsel) nsel[g.s0 + s] = h_nsel2[s];
            }
            n2 = h_off2[g.g1];
            // ---- lnpost's memory clip (:1029-1036): stars whose second selection exceeds nsel_max keep their
            // nsel_max best models by lnlike + lnprior (one segmented radix sort over the over-full stars) ----
            if (po->nsel_max > 0 && n2 > 0) {
                std::vector<int> seg;   // begin, end, slot per over-full star
                for (int s = g.g0; s < g.g1; s++)
                    if (h_nsel2[s] > po->nsel_max) { seg.push_back((int)h_off2[s]); seg.push_back((int)h_off2[s + 1]); seg.push_back(s); }
                const int nseg = (int)seg.size() / 3;
                if (nseg > 0) {
                    if (n2 >= ((int64_t)1 << 31)) { err = "bf_fit_batch: second selection too large for the memory clip"; return BF_E_NOMEM; }
                    std::vector<int> hs(3 * (size_t)nseg);
                    for (int k = 0; k < nseg; k++) { hs[k] = seg[3 * k]; hs[nseg + k] = seg[3 * k + 1]; hs[2 * nseg + k] = seg[3 * k + 2]; }
                    CK(d_seg.ensure(hs.size()));
                    CK(cudaMemcpyAsync(d_seg.p, hs.data(), hs.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
                    CK(d_keys.ensure((size_t)n2)); CK(d_keys_sorted.ensure((size_t)n2)); CK(d_clip.ensure((size_t)batch_cap));
                    pp.n2 = n2; pp.keys = d_keys.p;
                    CK(cudaMemcpyAsync(d_off2.p + g.g0, h_off2.data() + g.g0, (size_t)(ng + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
                    const unsigned nb1 = (unsigned)((n1 + kTile - 1) / kTile);
                    { TRACE("k_post_write"); k_post_write<T><<<nb1, kTile, 0, stream>>>(pp); }
                    k_post_keys<T><<<(unsigned)((n2 + 255) / 256), 256, 0, stream>>>(pp);
                    size_t tmp_bytes = 0;
                    CK(cub::DeviceSegmentedRadixSort::SortKeysDescending(nullptr, tmp_bytes, d_keys.p, d_keys_sorted.p, (int)n2, nseg,
                                                                         d_seg.p, d_seg.p + nseg, 0, (int)sizeof(T) * 8, stream));
                    CK(d_cubtmp.ensure(tmp_bytes + 16));
                    CK(cub::DeviceSegmentedRadixSort::SortKeysDescending(d_cubtmp.p, tmp_bytes, d_keys.p, d_keys_sorted.p, (int)n2, nseg,
                                                                         d_seg.p, d_seg.p + nseg, 0, (int)sizeof(T) * 8, stream));
                    k_fill<T><<<(batch_cap + 255) / 256, 256, 0, stream>>>(d_clip.p, batch_cap, -std::numeric_limits<T>::infinity());
                    k_post_thr<T><<<(nseg + 255) / 256, 256, 0, stream>>>(d_keys_sorted.p, d_seg.p, d_seg.p + 2 * nseg, nseg, po->nsel_max, d_clip.p);
                    pp.clip_thr = d_clip.p;
                    CK(cudaMemsetAsync(d_nsel2.p + g.g0, 0, (size_t)ng * sizeof(int), stream));
                    { TRACE("k_post_count"); k_post_count<T><<<nb1, kTile, 0, stream>>>(pp); }
                    { TRACE("k_scan_blocks"); k_scan_blocks<<<1, 1024, 0, stream>>>(d_blk.p, nb1, d_tot.p); }
                    stats.kernel_launches += 7;
                    CK(cudaGetLastError());
                    publish(h_nsel2.data() + g.g0, d_nsel2.p + g.g0, (size_t)ng * sizeof(int));
                    CK(sync());
                    for (int s = g.g0; s < g.g1; s++) {
                        h_off2[s + 1] = h_off2[s] + h_nsel2[s];
                        if (nsel) nsel[g.s0 + s] = h_nsel2[s];
                    }
                    n2 = h_off2[g.g1];
                    stats.clipped += nseg;
                }
            }
            stats.selected2 += n2;
            pp.n2 = n2;
            CK(cudaMemcpyAsync(d_off2.p + g.g0, h_off2.data() + g.g0, (size_t)(ng + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
            if (n2 > 0) {
                const unsigned nb1 = (unsigned)((n1 + kTile - 1) / kTile);
                { TRACE("k_post_write"); k_post_write<T><<<nb1, kTile, 0, stream>>>(pp); }
                if (pp.zov) { TRACE("k_post_mc"); k_post_mc<T, true, false><<<(unsigned)((n2 + kTile - 1) / kTile), kTile, 0, stream>>>(pp); }
                else if (G.same_rs && G.use) { TRACE("k_post_mc"); k_post_mc<T, false, true><<<(unsigned)((n2 + kTile - 1) / kTile), kTile, 0, stream>>>(pp); }
                else { TRACE("k_post_mc"); k_post_mc<T, false, false><<<(unsigned)((n2 + kTile - 1) / kTile), kTile, 0, stream>>>(pp); }
                stats.kernel_launches += 2;
            }
            { TRACE("k_post_cdf"); k_post_cdf<T><<<ng, 1024, 0, stream>>>(pp); }
            const int dthreads = std::min(kTile, (nd + 31) / 32 * 32);   // the kernel strides over the draws
            if (pp.zov) { TRACE("k_post_draw"); k_post_draw<T, true><<<ng, dthreads, 0, stream>>>(pp); }
            else { TRACE("k_post_draw"); k_post_draw<T, false><<<ng, dthreads, 0, stream>>>(pp); }
            stats.kernel_launches += 2;
            CK(cudaGetLastError());
            CK(cudaEventRecord(evP1, stream));
            // ---- ndraws samples per star back to the caller's arrays ----
            const size_t off = (size_t)g.g0 * nd, cnt = (size_t)ng * nd, dst = (aoff + (size_t)(g.s0 + g.g0)) * nd;   // aoff: first star of this engine in the arena
            CK(cudaMemcpyAsync(hidx + dst, pp.o_idx + off, cnt * sizeof(int), cudaMemcpyDeviceToHost, stream));
            double* hdst[8];
            for (int k = 0; k < 8; k++) hdst[k] = hd + (size_t)k * ntot;   // scale, av, rv, lnprob, dist, red, dred, logwt
            for (int k = 0; k < 8; k++)
                CK(cudaMemcpyAsync(hdst[k] + dst, ob + k * per + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(hd + 8 * ntot + dst * 9, pp.o_cov + off * 9, cnt * 9 * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(levid + g.s0 + g.g0, pp.o_levid + g.g0, (size_t)ng * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(chi2min + g.s0 + g.g0, pp.o_chi2min + g.g0, (size_t)ng * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(sync());
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, evP0, evP1));
            stats.ms_post += ms;
            stats.d2h_bytes += cnt * (sizeof(int) + 17 * sizeof(double)) + 2 * (size_t)ng * sizeof(double);
            return BF_OK;
        };
        int rc = run_catalogue(nstar, flux, errv, mask, par, perr, ext_mean, ext_std, opt, 11, false, ndim, n_iter,
                               nullptr, nullptr, offsets.data(), batch_cap, post);
        d_zov.release(); d_uov.release();
        if (rc) return rc;
        trace_collect();
        out->model_idx = hidx;
        out->scale = hd; out->av = hd + ntot; out->rv = hd + 2 * ntot; out->lnprob = hd + 3 * ntot; out->dist = hd + 4 * ntot;
        out->red = hd + 5 * ntot; out->dred = hd + 6 * ntot; out->logwt = hd + 7 * ntot; out->cov_sar = hd + 8 * ntot;
        if (ndim && par && perr)   // the parallax counts as one more datum (brutus/fitting.py:2028-2030)
            for (int64_t s = 0; s < nstar; s++)
                if (std::isfinite(par[s]) && std::isfinite(perr[s])) ndim[s] += 1;
        return BF_OK;
    }
};

}  // namespace bf

// =================================================================================================
// C ABI
// =================================================================================================
// ---- NCCL, resolved at run time --------------------------------------------------------------------------
// The library does not link NCCL: single-GPU use must work where NCCL is absent, and a process that also
// imports PyTorch already has a libnccl.so.2 loaded (dlopen by soname then returns that same copy instead of
// mixing two versions).  Only the multi-device entry points need it.
#include <dlfcn.h>
#include <nccl.h>
#include <thread>

namespace bf {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api(std::string& err) {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (api.lib) {
#define BF_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name)
            BF_SYM(GetUniqueId, "ncclGetUniqueId"); BF_SYM(CommInitRank, "ncclCommInitRank");
            BF_SYM(CommInitAll, "ncclCommInitAll"); BF_SYM(CommDestroy, "ncclCommDestroy");
            BF_SYM(Broadcast, "ncclBroadcast"); BF_SYM(AllReduce, "ncclAllReduce");
            BF_SYM(GroupStart, "ncclGroupStart"); BF_SYM(GroupEnd, "ncclGroupEnd");
            BF_SYM(GetErrorString, "ncclGetErrorString"); BF_SYM(GetVersion, "ncclGetVersion");
#undef BF_SYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.Broadcast ||
                !api.AllReduce || !api.GroupStart || !api.GroupEnd) { dlclose(api.lib); api.lib = nullptr; }
        }
    }
    if (!api.lib) { err = "NCCL (libnccl.so.2) could not be loaded: multi-GPU entry points are unavailable"; return nullptr; }
    return &api;
}
}  // namespace bf

// One handle = one engine per device.  bf_create: one device.  bf_create_multi: several devices of this
// process, an NCCL communicator per device (ncclCommInitAll), one host thread per device in the batch calls.
struct bf_handle {
    std::vector<bf::EngineBase*> eng;
    std::vector<ncclComm_t> comms;      // in-process group (bf_create_multi), or the one rank of a process group
    int rank = 0, world = 1;            // process group (bf_nccl_init); in-process handles keep (0, 1)
    std::string err;
    bf_stats stats{};
    // combined results of the multi-device batch calls (pinned, owned by the handle)
    char* draw_arena = nullptr; size_t draw_cap = 0;
    char* rec_arena = nullptr; int64_t rec_cap = 0;
    std::vector<int64_t> last_split;    // stars [last_split[d], last_split[d+1]) went to device d in the last call
    bf::EngineBase* e0() const { return eng[0]; }
};

#define HCK(h, call)                                                                           \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            char b_[512];                                                                      \
            snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            (h)->err = b_;                                                                     \
            return BF_E_CUDA;                                                                  \
        }                                                                                      \
    } while (0)
#define HNCCL(h, api, call)                                                                    \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess) {                                                               \
            (h)->err = std::string("NCCL: ") + #call + ": " + ((api)->GetErrorString ? (api)->GetErrorString(r_) : "error"); \
            return BF_E_CUDA;                                                                  \
        }                                                                                      \
    } while (0)

static bf::EngineBase* make_engine(int device, int precision, std::string& err, int* rc_out) {
    bf::EngineBase* eng = nullptr;
    if (precision == BF_PRECISION_F32) eng = new bf::Engine<float>();
    else if (precision == BF_PRECISION_F64) eng = new bf::Engine<double>();
    else { err = "bf_create: unknown precision"; *rc_out = BF_E_INVALID; return nullptr; }
    eng->device = device;
    eng->precision = precision;
    int rc = precision == BF_PRECISION_F32 ? static_cast<bf::Engine<float>*>(eng)->init()
                                           : static_cast<bf::Engine<double>*>(eng)->init();
    if (rc) { err = eng->err; delete eng; *rc_out = rc; return nullptr; }
    *rc_out = BF_OK;
    return eng;
}

// contiguous star ranges, sizes differing by at most one
static void split_stars(int64_t nstar, int nd, std::vector<int64_t>& b) {
    b.assign(nd + 1, 0);
    const int64_t base = nstar / nd, extra = nstar % nd;
    for (int d = 0; d < nd; d++) b[d + 1] = b[d] + base + (d < extra ? 1 : 0);
}

extern "C" {

void bf_default_options(bf_options* o) {
    if (!o) return;
    o->avlim[0] = 0.; o->avlim[1] = 20.;
    o->av_gauss[0] = 0.; o->av_gauss[1] = 1e6;
    o->rvlim[0] = 1.; o->rvlim[1] = 8.;
    o->rv_gauss[0] = 3.32; o->rv_gauss[1] = 0.18;
    o->ltol = 3e-2; o->ltol_subthresh = 1e-2; o->init_thresh = 5e-3; o->wt_thresh = 1e-3;
    o->select_slack = 0.5;
    o->dim_prior = 1; o->max_iter = 0; o->apply_parallax_clip = 1; o->skip_d2h = 0;
}

int bf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* bf_version(void) { return "brutus_b200 0.2.0 (sm_100a)"; }

int bf_create_multi(const int* devices, int ndev, int precision, bf_handle** out) {
    if (!out) { bf::g_create_error = "bf_create: null out pointer"; return BF_E_INVALID; }
    *out = nullptr;
    if (!devices || ndev < 1) { bf::g_create_error = "bf_create_multi: need at least one device"; return BF_E_INVALID; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        bf::g_create_error = std::string("bf_create: no CUDA device available (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return BF_E_CUDA;
    }
    for (int d = 0; d < ndev; d++) {
        if (devices[d] < 0 || devices[d] >= n) { bf::g_create_error = "bf_create: device ordinal out of range"; return BF_E_INVALID; }
        for (int k = 0; k < d; k++)
            if (devices[k] == devices[d]) { bf::g_create_error = "bf_create_multi: duplicate device"; return BF_E_INVALID; }
    }
    bf_handle* h = new bf_handle();
    for (int d = 0; d < ndev; d++) {
        int rc = BF_OK;
        bf::EngineBase* eng = make_engine(devices[d], precision, bf::g_create_error, &rc);
        if (!eng) { for (auto* x : h->eng) delete x; delete h; return rc; }
        h->eng.push_back(eng);
    }
    if (ndev > 1) {   // the communicators the grid broadcast runs on
        bf::NcclApi* api = bf::nccl_api(bf::g_create_error);
        ncclResult_t r = ncclSuccess;
        if (api) {
            h->comms.resize(ndev);
            r = api->CommInitAll(h->comms.data(), ndev, devices);
        }
        if (!api || r != ncclSuccess) {
            if (api) bf::g_create_error = std::string("ncclCommInitAll: ") + (api->GetErrorString ? api->GetErrorString(r) : "error");
            for (auto* x : h->eng) delete x;
            delete h;
            return BF_E_CUDA;
        }
    }
    *out = h;
    return BF_OK;
}

int bf_create(int device, int precision, bf_handle** out) { return bf_create_multi(&device, 1, precision, out); }

int bf_num_devices(const bf_handle* h) { return h ? (int)h->eng.size() : 0; }

int bf_destroy(bf_handle* h) {
    if (!h) return BF_OK;
    if (!h->comms.empty()) {
        std::string e;
        if (bf::NcclApi* api = bf::nccl_api(e))
            for (ncclComm_t c : h->comms) if (c) api->CommDestroy(c);
    }
    for (auto* x : h->eng) delete x;
    if (h->draw_arena) cudaFreeHost(h->draw_arena);
    if (h->rec_arena) cudaFreeHost(h->rec_arena);
    delete h;
    return BF_OK;
}

const char* bf_last_error(const bf_handle* h) {
    if (!h) return bf::g_create_error.c_str();
    if (!h->err.empty()) return h->err.c_str();
    for (auto* x : h->eng) if (!x->err.empty()) return x->err.c_str();
    return "";
}

// ---- process groups: one process per GPU (torchrun), NCCL inside the library -------------------------------
int bf_nccl_unique_id(void* out128) {
    if (!out128) return BF_E_INVALID;
    bf::NcclApi* api = bf::nccl_api(bf::g_create_error);
    if (!api) return BF_E_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) { bf::g_create_error = "ncclGetUniqueId failed"; return BF_E_CUDA; }
    std::memcpy(out128, &id, sizeof id);
    return BF_OK;
}

int bf_nccl_init(bf_handle* h, const void* id128, int rank, int world) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return BF_E_INVALID;
    if (h->eng.size() != 1) { h->err = "bf_nccl_init: a process-group handle drives exactly one device"; return BF_E_INVALID; }
    h->err.clear();
    bf::NcclApi* api = bf::nccl_api(h->err);
    if (!api) return BF_E_CUDA;
    HCK(h, cudaSetDevice(h->e0()->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    h->comms.assign(1, nullptr);
    HNCCL(h, api, api->CommInitRank(&h->comms[0], world, id, rank));
    h->rank = rank; h->world = world;
    return BF_OK;
}

// host buffer of `bytes` bytes: on return every rank holds rank `root`'s content (staged through the device)
int bf_bcast_host(bf_handle* h, void* buf, int64_t bytes, int root) {
    if (!h || bytes < 0 || (bytes > 0 && !buf)) return BF_E_INVALID;
    if (h->world == 1 || bytes == 0) return BF_OK;
    h->err.clear();
    bf::NcclApi* api = bf::nccl_api(h->err);
    if (!api) return BF_E_CUDA;
    HCK(h, cudaSetDevice(h->e0()->device));
    void* d = nullptr;
    HCK(h, cudaMalloc(&d, (size_t)bytes));
    if (h->rank == root) HCK(h, cudaMemcpy(d, buf, (size_t)bytes, cudaMemcpyHostToDevice));
    ncclResult_t r = api->Broadcast(d, d, (size_t)bytes, ncclChar, root, h->comms[0], 0);
    cudaError_t e = cudaStreamSynchronize(0);
    if (r == ncclSuccess && e == cudaSuccess && h->rank != root) e = cudaMemcpy(buf, d, (size_t)bytes, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (r != ncclSuccess) { h->err = std::string("ncclBroadcast: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"); return BF_E_CUDA; }
    HCK(h, e);
    return BF_OK;
}

// element-wise maximum over the ranks of n doubles, in place (also a barrier): device-timed results of a
// sharded run are reported as the maximum over ranks
int bf_allreduce_max(bf_handle* h, double* vals, int32_t n) {
    if (!h || n < 0 || (n > 0 && !vals)) return BF_E_INVALID;
    if (h->world == 1 || n == 0) return BF_OK;
    h->err.clear();
    bf::NcclApi* api = bf::nccl_api(h->err);
    if (!api) return BF_E_CUDA;
    HCK(h, cudaSetDevice(h->e0()->device));
    double* d = nullptr;
    HCK(h, cudaMalloc((void**)&d, (size_t)n * sizeof(double)));
    HCK(h, cudaMemcpy(d, vals, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    ncclResult_t r = api->AllReduce(d, d, (size_t)n, ncclDouble, ncclMax, h->comms[0], 0);
    cudaError_t e = cudaStreamSynchronize(0);
    if (r == ncclSuccess && e == cudaSuccess) e = cudaMemcpy(vals, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"); return BF_E_CUDA; }
    HCK(h, e);
    return BF_OK;
}

// The grid goes host -> device once (first device of the handle / rank `root` of the process group), then ONE
// ncclBroadcast replicates it GPU to GPU over NVLink; every device re-tiles its copy.
static int set_grid_everywhere(bf_handle* h, const float* coeffs, int64_t nmodel, int32_t nfilt, int32_t layout, int root) {
    h->err.clear();
    const int nd = (int)h->eng.size();
    if (nd == 1 && h->world == 1) return h->e0()->set_grid(coeffs, nmodel, nfilt, layout, false);
    if (nmodel <= 0 || nfilt <= 0 || nfilt > BF_MAX_FILT) { h->err = "bf_set_grid: bad shape"; return BF_E_INVALID; }
    bf::NcclApi* api = bf::nccl_api(h->err);
    if (!api) return BF_E_CUDA;
    const size_t nval = (size_t)nmodel * nfilt * 3;
    const bool have_src = nd > 1 || h->rank == root;
    if (have_src && !coeffs) { h->err = "bf_set_grid: null grid on the root"; return BF_E_INVALID; }
    std::vector<float*> buf(nd, nullptr);
    std::vector<cudaStream_t> st(nd, nullptr);
    int rc = BF_OK;
    auto cleanup = [&]() {
        for (int d = 0; d < nd; d++) {
            cudaSetDevice(h->eng[d]->device);
            if (st[d]) cudaStreamDestroy(st[d]);
            if (buf[d]) cudaFree(buf[d]);
        }
    };
    for (int d = 0; d < nd && rc == BF_OK; d++) {
        if (cudaSetDevice(h->eng[d]->device) != cudaSuccess || cudaMalloc((void**)&buf[d], nval * sizeof(float)) != cudaSuccess ||
            cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking) != cudaSuccess) { h->err = "bf_set_grid: device allocation failed"; rc = BF_E_NOMEM; }
    }
    if (rc == BF_OK && have_src) {
        cudaSetDevice(h->eng[0]->device);
        if (cudaMemcpyAsync(buf[0], coeffs, nval * sizeof(float), cudaMemcpyHostToDevice, st[0]) != cudaSuccess) { h->err = "bf_set_grid: H2D copy failed"; rc = BF_E_CUDA; }
    }
    if (rc == BF_OK) {
        ncclResult_t r = api->GroupStart();
        for (int d = 0; d < nd && r == ncclSuccess; d++) {
            cudaSetDevice(h->eng[d]->device);
            r = api->Broadcast(buf[d], buf[d], nval, ncclFloat, nd > 1 ? 0 : root, h->comms[nd > 1 ? d : 0], st[d]);
        }
        ncclResult_t r2 = api->GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) { h->err = std::string("ncclBroadcast: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"); rc = BF_E_CUDA; }
    }
    for (int d = 0; d < nd && rc == BF_OK; d++) {
        cudaSetDevice(h->eng[d]->device);
        if (cudaStreamSynchronize(st[d]) != cudaSuccess) { h->err = "bf_set_grid: broadcast failed"; rc = BF_E_CUDA; break; }
        rc = h->eng[d]->set_grid(buf[d], nmodel, nfilt, layout, true);
    }
    cleanup();
    return rc;
}

int bf_set_grid(bf_handle* h, const float* coeffs, int64_t nmodel, int32_t nfilt, int32_t layout) {
    if (!h) return BF_E_INVALID;
    if (h->world > 1) { h->err = "bf_set_grid: this handle belongs to a process group, use bf_set_grid_bcast"; return BF_E_INVALID; }
    return set_grid_everywhere(h, coeffs, nmodel, nfilt, layout, 0);
}

int bf_set_grid_bcast(bf_handle* h, const float* coeffs, int64_t nmodel, int32_t nfilt, int32_t layout, int32_t root) {
    if (!h) return BF_E_INVALID;
    return set_grid_everywhere(h, coeffs, nmodel, nfilt, layout, root);
}

int bf_set_grid_device(bf_handle* h, const void* d_coeffs, int64_t nmodel, int32_t nfilt, int32_t layout) {
    if (!h) return BF_E_INVALID;
    if (h->eng.size() != 1) { h->err = "bf_set_grid_device: single-device handles only"; return BF_E_INVALID; }
    h->err.clear();
    return h->e0()->set_grid((const float*)d_coeffs, nmodel, nfilt, layout, true);
}

int bf_set_labels(bf_handle* h, const double* labels, int32_t nlabel) {
    if (!h) return BF_E_INVALID;
    h->err.clear();
    for (auto* e : h->eng) { int rc = e->set_labels(labels, nlabel); if (rc) return rc; }
    return BF_OK;
}

int bf_loglike_full(bf_handle* h, const double* flux, const double* err, const uint8_t* mask, double parallax,
                    double parallax_err, const bf_options* opt, double* lnl, double* chi2, double* scale,
                    double* av, double* rv, double* icov, uint8_t* mask_clean_out, int64_t* diag) {
    if (!h) return BF_E_INVALID;
    h->err.clear();
    int rc = h->e0()->loglike_full(flux, err, mask, parallax, parallax_err, opt, lnl, chi2, scale, av, rv, icov,
                                   mask_clean_out, diag);
    h->stats = h->e0()->stats;
    return rc;
}

// statistics of a multi-device call: counts add up, device times are the slowest device's
static void merge_stats(bf_handle* h) {
    bf_stats t = h->eng[0]->stats;
    for (size_t d = 1; d < h->eng.size(); d++) {
        const bf_stats& s = h->eng[d]->stats;
        t.ms_device = std::max(t.ms_device, s.ms_device); t.ms_magfit = std::max(t.ms_magfit, s.ms_magfit);
        t.ms_flux = std::max(t.ms_flux, s.ms_flux); t.ms_select = std::max(t.ms_select, s.ms_select);
        t.ms_post = std::max(t.ms_post, s.ms_post);
        t.kernel_launches += s.kernel_launches; t.magfit_launches += s.magfit_launches;
        t.magfit_star_passes += s.magfit_star_passes; t.resweeps += s.resweeps; t.candidates += s.candidates;
        t.fallbacks += s.fallbacks; t.survivors += s.survivors; t.selected += s.selected;
        t.h2d_bytes += s.h2d_bytes; t.d2h_bytes += s.d2h_bytes; t.selected2 += s.selected2; t.clipped += s.clipped;
        t.fixups += s.fixups; t.flux_more_launches += s.flux_more_launches; t.regroups += s.regroups;
        t.unconverged += s.unconverged; t.host_syncs = std::max(t.host_syncs, s.host_syncs);
    }
    h->stats = t;
}

int bf_sweep_batch(bf_handle* h, int64_t nstar, const double* flux, const double* err, const uint8_t* mask,
                   const double* parallax, const double* parallax_err, const double* ext_mean,
                   const double* ext_std, const bf_options* opt, int32_t record_rows, int32_t* ndim,
                   int32_t* n_iter, int64_t* n_surv, double* max_lnprob, int64_t* offsets, bf_records* out) {
    if (!h) return BF_E_INVALID;
    h->err.clear();
    const int nd = (int)h->eng.size();
    if (nd == 1 || nstar < nd) {
        int rc = h->e0()->sweep_batch(nstar, flux, err, mask, parallax, parallax_err, ext_mean, ext_std, opt,
                                      record_rows, ndim, n_iter, n_surv, max_lnprob, offsets, out);
        h->stats = h->e0()->stats;
        split_stars(nstar, 1, h->last_split);
        return rc;
    }
    if (!flux || !err || !mask || !opt || !offsets || !out) { h->err = "bf_sweep_batch: null argument"; return BF_E_INVALID; }
    // stars sharded contiguously over the devices, one host thread each, no collective
    const int nfilt = h->e0()->nfilt_(), nlabel = h->e0()->nlabel_();
    split_stars(nstar, nd, h->last_split);
    const std::vector<int64_t>& b = h->last_split;
    std::vector<int> rcs(nd, BF_OK);
    std::vector<bf_records> recs(nd);
    std::vector<std::vector<int64_t>> offs(nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++) {
        offs[d].assign((size_t)(b[d + 1] - b[d]) + 1, 0);
        th.emplace_back([&, d]() {
            const int64_t lo = b[d], n = b[d + 1] - b[d];
            rcs[d] = h->eng[d]->sweep_batch(n, flux + lo * nfilt, err + lo * nfilt, mask + lo * nfilt,
                                            parallax ? parallax + lo : nullptr, parallax_err ? parallax_err + lo : nullptr,
                                            ext_mean ? ext_mean + lo * nlabel : nullptr, ext_std ? ext_std + lo * nlabel : nullptr,
                                            opt, record_rows, ndim ? ndim + lo : nullptr, n_iter ? n_iter + 2 * lo : nullptr,
                                            n_surv ? n_surv + lo : nullptr, max_lnprob ? max_lnprob + lo : nullptr,
                                            offs[d].data(), &recs[d]);
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < nd; d++) if (rcs[d]) return rcs[d];
    merge_stats(h);
    // gather: one CSR over the whole catalogue, records concatenated in star order in the handle's arena
    int64_t total = 0;
    offsets[0] = 0;
    for (int d = 0; d < nd; d++) {
        for (int64_t s = 0; s < b[d + 1] - b[d]; s++) offsets[b[d] + s + 1] = total + offs[d][s + 1];
        total += recs[d].n;
    }
    const size_t es = (size_t)recs[0].elem_size;
    if (total > h->rec_cap) {
        if (h->rec_arena) cudaFreeHost(h->rec_arena);
        h->rec_arena = nullptr; h->rec_cap = 0;
        HCK(h, cudaHostAlloc((void**)&h->rec_arena, (size_t)total * (sizeof(int) + 11 * es) + 64, cudaHostAllocPortable));
        h->rec_cap = total;
    }
    const int64_t cap = std::max<int64_t>(h->rec_cap, 1);
    if (total > 0) {
        th.clear();
        int64_t at = 0;
        for (int d = 0; d < nd; d++) {
            const int64_t n = recs[d].n, dst = at;
            at += n;
            if (n == 0) continue;
            th.emplace_back([&, d, n, dst]() {
                std::memcpy(h->rec_arena + (size_t)11 * cap * es + (size_t)dst * sizeof(int), recs[d].model_idx, (size_t)n * sizeof(int));
                for (int r = 0; r < record_rows; r++)
                    std::memcpy(h->rec_arena + ((size_t)r * cap + (size_t)dst) * es,
                                (const char*)recs[d].rows + (size_t)r * recs[d].stride * es, (size_t)n * es);
            });
        }
        for (auto& t : th) t.join();
    }
    out->n = total; out->stride = cap; out->elem_size = (int32_t)es; out->nrows = record_rows;
    out->model_idx = h->rec_arena ? (const int32_t*)(h->rec_arena + (size_t)11 * cap * es) : nullptr;
    out->rows = h->rec_arena;
    return BF_OK;
}

void bf_default_gal_params(bf_gal_params* g) {   /* defaults of gal_lnprior, brutus/pdf.py:476-486 */
    if (!g) return;
    g->R_solar = 8.2; g->Z_solar = 0.025; g->R_thin = 2.6; g->Z_thin = 0.3; g->Rs_thin = 2.0;
    g->R_thick = 2.0; g->Z_thick = 0.9; g->f_thick = 0.04; g->Rs_thick = 2.0;
    g->Rs_halo = 2.0; g->q_halo_ctr = 0.2; g->q_halo_inf = 0.8; g->r_q_halo = 6.0; g->eta_halo = 4.2; g->f_halo = 0.005;
    g->feh_thin = -0.2; g->feh_thin_sigma = 0.3; g->feh_thick = -0.7; g->feh_thick_sigma = 0.4;
    g->feh_halo = -1.6; g->feh_halo_sigma = 0.5;
    g->max_age = 13.8; g->min_age = 0.; g->feh_age_ctr = -0.5; g->feh_age_scale = 0.5;
    g->nsigma_from_max_age = 2.; g->max_sigma = 4.; g->min_sigma = 1.;
    g->galcen_distance = 8.122; g->z_sun = 0.0208;
}

void bf_default_post_options(bf_post_options* o) {
    if (!o) return;
    o->nmc_prior = 50; o->ndraws = 250; o->seed = 0; o->use_gal_prior = 1; o->reserved = 0; o->star_base = 0;
    o->nsel_max = 0;
    bf_default_gal_params(&o->gal);
    o->z_override = nullptr; o->u_override = nullptr;
}

int bf_set_model_priors(bf_handle* h, const double* lnprior, const double* feh, const double* loga) {
    if (!h) return BF_E_INVALID;
    h->err.clear();
    for (auto* e : h->eng) { int rc = e->set_model_priors(lnprior, feh, loga); if (rc) return rc; }
    return BF_OK;
}

int bf_set_init(bf_handle* h, const double* av_init, const double* rv_init) {
    if (!h) return BF_E_INVALID;
    h->err.clear();
    return h->e0()->set_init(av_init, rv_init);   // bf_loglike_full runs on the first device
}

int bf_fit_batch(bf_handle* h, int64_t nstar, const double* flux, const double* err, const uint8_t* mask,
                 const double* parallax, const double* parallax_err, const double* coords,
                 const double* ext_mean, const double* ext_std, const bf_options* opt,
                 const bf_post_options* post, int32_t* ndim, int32_t* n_iter, int64_t* nsel,
                 double* levid, double* chi2min, bf_draws* out) {
    if (!h) return BF_E_INVALID;
    h->err.clear();
    const int nd = (int)h->eng.size();
    if (nd == 1 || nstar < nd) {
        h->e0()->set_ext_arena(nullptr, 0, 0);
        int rc = h->e0()->fit_batch(nstar, flux, err, mask, parallax, parallax_err, coords, ext_mean, ext_std, opt, post,
                                    ndim, n_iter, nsel, levid, chi2min, out);
        h->stats = h->e0()->stats;
        split_stars(nstar, 1, h->last_split);
        return rc;
    }
    if (!flux || !err || !mask || !opt || !post || !out || !levid || !chi2min) { h->err = "bf_fit_batch: null argument"; return BF_E_INVALID; }
    if (post->ndraws < 1 || post->nmc_prior < 1) { h->err = "bf_fit_batch: nmc_prior and ndraws must be >= 1"; return BF_E_INVALID; }
    if (post->u_override) { h->err = "bf_fit_batch: u_override is a single-device test hook"; return BF_E_INVALID; }
    // one shared pinned arena: every device writes the draws of its stars straight into its slice
    const size_t ntot = (size_t)nstar * post->ndraws;
    if (ntot > h->draw_cap) {
        if (h->draw_arena) cudaFreeHost(h->draw_arena);
        h->draw_arena = nullptr; h->draw_cap = 0;
        HCK(h, cudaHostAlloc((void**)&h->draw_arena, ntot * (17 * sizeof(double) + sizeof(int32_t)) + 64, cudaHostAllocPortable));
        h->draw_cap = ntot;
    }
    const int nfilt = h->e0()->nfilt_(), nlabel = h->e0()->nlabel_();
    split_stars(nstar, nd, h->last_split);
    const std::vector<int64_t>& b = h->last_split;
    std::vector<int> rcs(nd, BF_OK);
    std::vector<bf_draws> dr(nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++) {
        th.emplace_back([&, d]() {
            const int64_t lo = b[d], n = b[d + 1] - b[d];
            bf_post_options po = *post;
            po.star_base = post->star_base + lo;   // the generator is keyed by the catalogue index: results do not depend on the sharding
            h->eng[d]->set_ext_arena(h->draw_arena, ntot, (size_t)lo);
            rcs[d] = h->eng[d]->fit_batch(n, flux + lo * nfilt, err + lo * nfilt, mask + lo * nfilt,
                                          parallax ? parallax + lo : nullptr, parallax_err ? parallax_err + lo : nullptr,
                                          coords ? coords + 2 * lo : nullptr,
                                          ext_mean ? ext_mean + lo * nlabel : nullptr, ext_std ? ext_std + lo * nlabel : nullptr,
                                          opt, &po, ndim ? ndim + lo : nullptr, n_iter ? n_iter + 2 * lo : nullptr,
                                          nsel ? nsel + lo : nullptr, levid + lo, chi2min + lo, &dr[d]);
            h->eng[d]->set_ext_arena(nullptr, 0, 0);
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < nd; d++) if (rcs[d]) return rcs[d];
    merge_stats(h);
    *out = dr[0];   // every engine reports the same arena pointers
    return BF_OK;
}

int bf_get_seds(bf_handle* h, int64_t n, const int32_t* idx, const double* av, const double* rv, int32_t return_flux,
                double* seds, double* rvecs, double* drvecs) {
    if (!h) return BF_E_INVALID;
    h->err.clear();
    return h->e0()->get_seds(n, idx, av, rv, return_flux, seds, rvecs, drvecs);
}

int bf_offsets_weights(bf_handle* h, int64_t nobj, int32_t nsamps, const double* phot, const double* err, const uint8_t* mask,
                       const int32_t* idxs, const double* reds, const double* dreds, const double* dists,
                       const double* old_offsets, const uint8_t* mask_fit, int32_t dim_prior, double* seds, double* wt) {
    if (!h) return BF_E_INVALID;
    h->err.clear();
    return h->e0()->offsets_weights(nobj, nsamps, phot, err, mask, idxs, reds, dreds, dists, old_offsets, mask_fit, dim_prior,
                                    seds, wt);
}

int bf_flush_l2(bf_handle* h) {
    if (!h) return BF_E_INVALID;
    for (auto* e : h->eng) { int rc = e->flush_l2(); if (rc) return rc; }
    return BF_OK;
}

const char* bf_get_trace(bf_handle* h) { return h ? h->e0()->get_trace() : ""; }

int bf_get_stats(const bf_handle* h, bf_stats* out) {
    if (!h || !out) return BF_E_INVALID;
    *out = h->stats;
    return BF_OK;
}

}  // extern "C"
