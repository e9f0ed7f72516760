// api.cu -- C ABI (include/brutus_b200.h), host orchestration and the band-count-independent
// kernels of the brute-force likelihood sweep.  See DESIGN.md for the pipeline:
//
//   prepare star rows (host, float64)  ->  k_magfit (full grid, speculated iteration count)
//   -> verify / re-sweep mispredicted stars -> cull (count, scan, ordered write) -> k_flux on the
//   survivor pool until every star converges -> scatter -> k_lnprob (+max) -> either full-length
//   outputs (B1, bf_loglike_full) or threshold + ordered compaction + k_records (B2, bf_sweep_batch).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/brutus_b200.h"
#include "common.cuh"

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

// grid re-tiling: user layout (C or Fortran order of (nmodel, nfilt, 3)) -> [coef][band][npad]
__global__ void k_retile(const float* __restrict__ src, float* __restrict__ dst, int64_t nmodel,
                         int64_t npad, int nfilt, int layout) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    for (int j = 0; j < nfilt; j++)
        for (int c = 0; c < 3; c++) {
            float v = 0.f;
            if (i < nmodel)
                v = layout == BF_LAYOUT_C ? src[(i * nfilt + j) * 3 + c]
                                          : src[((int64_t)c * nfilt + j) * nmodel + i];
            dst[((int64_t)c * nfilt + j) * npad + i] = v;
        }
}

template <typename T> __global__ void k_convert_labels(const double* src, T* dst, int64_t nmodel, int64_t npad, int nlabel) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    for (int l = 0; l < nlabel; l++) dst[(int64_t)l * npad + i] = i < nmodel ? (T)src[(int64_t)l * nmodel + i] : T(0);
}

template <typename T>
__global__ void k_reset_red(typename Enc<T>::U* red, const int* list, int nlist, unsigned mask) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlist * kNumRed) return;
    int s = t / kNumRed, k = t % kNumRed;
    if (mask >> k & 1u) red[(int64_t)list[s] * kNumRed + k] = Enc<T>::enc(Num<T>::neg_inf());
}

// lnl_p of the cull, same expression as kernels_nb.cuh::cull_lnl
template <typename T> __device__ __forceinline__ T cull_lnl2(T chi2, T s, const T* __restrict__ srow) {
    T dp = Num<T>::sqrt_fast(s) - srow[SR_SC + SC_PAR];
    return T(-0.5) * fma(dp * dp, srow[SR_SC + SC_PIVAR], chi2);
}

enum { FLAG_CULL = 0, FLAG_SELECT = 1 };

template <typename T> struct FlagParams {
    const T* stars;
    StateArrays<T> st;
    const typename Enc<T>::U* red;
    DevOpts<T> o;
    int64_t npad, nmodel;
    const int* list;
    int nlist, ntile;
    int* cnt;            // [batch][ntile]: per-tile count, then (after k_scan_tiles) exclusive offset
    const int64_t* base; // [batch]: first record of each star in the pool
    int *out_model, *out_star;
};

template <typename T, int MODE>
__device__ __forceinline__ bool eval_flag(const FlagParams<T>& p, int slot, int64_t i) {
    if (i >= p.nmodel) return false;
    const int64_t off = (int64_t)slot * p.npad + i;
    if (MODE == FLAG_CULL) {
        // brutus/fitting.py:758-759: lnl_p > max(lnl_p) + ln(init_thresh)
        const T* srow = p.stars + (int64_t)slot * kStarStride;
        T lmax = Enc<T>::dec(p.red[(int64_t)slot * kNumRed + RED_LP]);
        return cull_lnl2(p.st.chi2[off], p.st.scale[off], srow) > lmax + p.o.ln_init;
    } else {
        // brutus/fitting.py:990-991: lnprob > max(lnprob) + ln(wt_thresh)
        T lmax = Enc<T>::dec(p.red[(int64_t)slot * kNumRed + RED_LNP]);
        return p.st.lnprob[off] > lmax + p.o.ln_wt;
    }
}

template <typename T, int MODE> __global__ void __launch_bounds__(kTile) k_count(const FlagParams<T> p) {
    const int slot = p.list[blockIdx.y];
    const int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x;
    int n = __syncthreads_count(eval_flag<T, MODE>(p, slot, i));
    if (threadIdx.x == 0) p.cnt[(int64_t)slot * p.ntile + blockIdx.x] = n;
}

// one CTA per star: in-place exclusive scan of the per-tile counts; total -> tot[slot]
__global__ void __launch_bounds__(1024) k_scan_tiles(int* cnt, const int* list, int ntile, int64_t* tot) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int slot = list[blockIdx.x];
    int* c = cnt + (int64_t)slot * ntile;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b = 0; b < ntile; b += 1024) {
        int t = b + threadIdx.x;
        int v = t < ntile ? c[t] : 0;
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
        int carry = s_carry;
        int incl = x + ((threadIdx.x >> 5) ? s_w[(threadIdx.x >> 5) - 1] : 0);
        if (t < ntile) c[t] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) tot[slot] = s_carry;
}

// ordered write: records of a star are contiguous and ascending in model index
template <typename T, int MODE> __global__ void __launch_bounds__(kTile) k_write(const FlagParams<T> p) {
    __shared__ int s_w[kTile / 32];
    const int slot = p.list[blockIdx.y];
    const int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x;
    const bool f = eval_flag<T, MODE>(p, slot, i);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_w[w] = __popc(bal);
    __syncthreads();
    if (f) {
        int pre = __popc(bal & ((1u << lane) - 1u));
        for (int k = 0; k < w; k++) pre += s_w[k];
        int64_t pos = p.base[slot] + p.cnt[(int64_t)slot * p.ntile + blockIdx.x] + pre;
        p.out_model[pos] = (int)i;
        p.out_star[pos] = slot;
    }
}

// start of the flux loop: stepsize 1, lnl_old = -1e300 (brutus/fitting.py:778-779)
template <typename T> __global__ void k_flux_init(PoolArrays<T> pool, int64_t n, StateArrays<T> st, int64_t npad) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int64_t off = (int64_t)pool.star[q] * npad + pool.model[q];
    pool.av[q] = st.av[off];
    pool.rv[q] = st.rv[off];
    pool.eta[q] = T(1);
    pool.lold[q] = Num<T>::kNegBig;
}

// brutus/fitting.py:808-810: survivors' results replace the mag-fit values.  s_den > 0 always, so its
// sign bit marks "survived the cull" (needed for the Gaussian constant when dim_prior is off, :806-808).
template <typename T> __global__ void k_scatter(PoolArrays<T> pool, int64_t n, StateArrays<T> st, int64_t npad) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int64_t off = (int64_t)pool.star[q] * npad + pool.model[q];
    st.chi2[off] = pool.chi2[q];
    st.scale[off] = pool.scale[q];
    st.sden[off] = -pool.sden[q];
    st.av[off] = pool.av[q];
    st.rv[off] = pool.rv[q];
}

template <typename T> struct LnprobParams {
    const T* stars;
    StateArrays<T> st;
    typename Enc<T>::U* red;
    DevOpts<T> o;
    int64_t npad, nmodel;
    const int* list;
    const T* labels;   // [nlabel][npad]
    const T* ext;      // [batch][nlabel][3] = (mean, 1/std^2, ln(2 pi std^2)); ivar = 0 -> inactive
    int nlabel;
};

// lnlike as loglike returns it (brutus/fitting.py:806-815, brutus/utils.py:130-176), the external label
// priors (:1995-2009), lnpost's rough parallax prior (:976-982, brutus/pdf.py:178-222) and the -1e300
// clean-up (:983-985); per-star max(lnprob) (:990).
template <typename T> __global__ void __launch_bounds__(kTile) k_lnprob(const LnprobParams<T> p) {
    __shared__ typename Enc<T>::U s_max;
    const int slot = p.list[blockIdx.y];
    const int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x;
    if (threadIdx.x == 0) s_max = Enc<T>::enc(Num<T>::neg_inf());
    __syncthreads();
    T v = Num<T>::neg_inf();
    if (i < p.nmodel) {
        const T* __restrict__ srow = p.stars + (int64_t)slot * kStarStride;
        const int64_t off = (int64_t)slot * p.npad + i;
        const T chi2 = p.st.chi2[off];
        const T sd = p.st.sden[off];
        T lnl;
        if (p.o.dim_prior) {
            lnl = chi2 <= T(0) ? Num<T>::neg_inf()
                               : srow[SR_SC + SC_LNORM] + srow[SR_SC + SC_KHM1] * Num<T>::log(chi2) - T(0.5) * chi2;
        } else {
            lnl = T(-0.5) * chi2 + (sd < T(0) ? srow[SR_SC + SC_GCONST] : T(0));
        }
        for (int l = 0; l < p.nlabel; l++) {
            const T* x = p.ext + ((int64_t)slot * p.nlabel + l) * 3;
            if (x[1] > T(0)) {
                T d = p.labels[(int64_t)l * p.npad + i] - x[0];
                lnl += T(-0.5) * (d * d * x[1] + x[2]);
            }
        }
        T lp = lnl;
        if (srow[SR_SC + SC_SPAPPLY] != T(0)) {
            T svar = srow[SR_SC + SC_SVAR] + T(1) / tabs(sd);
            T d = p.st.scale[off] - srow[SR_SC + SC_SMEAN];
            lp = lnl + T(-0.5) * (Num<T>::div(d * d, svar) + Num<T>::log(T(6.283185307179586) * svar));
        }
        if (!Num<T>::finite(lp)) lp = Num<T>::kNegBig;
        p.st.lnl[off] = lnl;
        p.st.lnprob[off] = lp;
        v = lp;
    }
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, Enc<T>::enc(v));
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(&p.red[(int64_t)slot * kNumRed + RED_LNP], s_max);
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

template <typename P> struct DevBuf {
    P* p = nullptr;
    size_t n = 0;
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
                      double perr, int apply_clip, StarPrep& sp) {
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
    sc[SC_S] = S;
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
    sp.ndim = ndim;
}

template <typename T> struct Engine : EngineBase {
    using U = typename Enc<T>::U;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr;
    cudaEvent_t ev_rec[2] = {nullptr, nullptr}, ev_cp[2] = {nullptr, nullptr};
    char* arena = nullptr;      // pinned host memory holding the records of the last bf_sweep_batch
    int64_t arena_cap = 0;
    int64_t nmodel = 0, npad = 0;
    int nfilt = 0, ntile = 0, nlabel = 0;
    int batch_cap = 0;
    int64_t pool_cap = 0;
    const KTable<T>* kt = nullptr;

    DevBuf<float> d_grid;
    DevBuf<T> d_labels, d_stars, d_ext, d_state, d_poolT;
    DevBuf<int> d_star_int, d_list, d_cnt, d_poolI;
    DevBuf<int64_t> d_tot, d_base;
    DevBuf<U> d_red;
    DevBuf<double> d_out;
    DevBuf<char> d_flush, d_stage[2];

    std::vector<T> h_stars, h_ext;
    std::vector<int> h_int, h_list;
    std::vector<U> h_red;
    std::vector<int64_t> h_tot, h_base;

    ~Engine() override {
        cudaSetDevice(device);
        d_grid.release(); d_labels.release(); d_stars.release(); d_ext.release(); d_state.release();
        d_poolT.release(); d_star_int.release(); d_list.release(); d_cnt.release(); d_poolI.release();
        d_tot.release(); d_base.release(); d_red.release(); d_out.release(); d_flush.release();
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
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }

    int init() {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
        CK(cudaEventCreate(&evA)); CK(cudaEventCreate(&evB));
        CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CK(cudaEventCreateWithFlags(&ev_rec[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_cp[k], cudaEventDisableTiming));
        }
        return BF_OK;
    }

    StateArrays<T> state() {
        StateArrays<T> s;
        const size_t n = (size_t)batch_cap * npad;
        s.chi2 = d_state.p; s.scale = d_state.p + n; s.sden = d_state.p + 2 * n; s.av = d_state.p + 3 * n;
        s.rv = d_state.p + 4 * n; s.lnl = d_state.p + 5 * n; s.lnprob = d_state.p + 6 * n;
        return s;
    }
    PoolArrays<T> pool() {
        PoolArrays<T> q;
        q.model = d_poolI.p; q.star = d_poolI.p + pool_cap;
        T* b = d_poolT.p;
        q.av = b; q.rv = b + pool_cap; q.eta = b + 2 * pool_cap; q.lold = b + 3 * pool_cap;
        q.chi2 = b + 4 * pool_cap; q.scale = b + 5 * pool_cap; q.sden = b + 6 * pool_cap;
        return q;
    }

    int set_grid(const float* co, int64_t nm, int nf, int layout, bool on_device) override {
        CK(cudaSetDevice(device));
        if (!co || nm <= 0 || nf <= 0) { err = "bf_set_grid: null grid or non-positive shape"; return BF_E_INVALID; }
        if (nf > kMaxFilt) { err = "bf_set_grid: nfilt exceeds BF_MAX_FILT (16)"; return BF_E_INVALID; }
        if (nm > (int64_t)2000000000) { err = "bf_set_grid: nmodel too large"; return BF_E_INVALID; }
        if (layout != BF_LAYOUT_C && layout != BF_LAYOUT_F) { err = "bf_set_grid: unknown layout"; return BF_E_INVALID; }
        kt = get_ktable<T>(nf);
        if (!kt) { err = "bf_set_grid: no kernels compiled for this band count"; return BF_E_INVALID; }
        nmodel = nm; nfilt = nf;
        npad = (nm + kTile - 1) / kTile * kTile;
        ntile = (int)(npad / kTile);
        nlabel = 0;
        const size_t nval = (size_t)nm * nf * 3;
        CK(d_grid.ensure((size_t)3 * nf * npad));
        const float* src = co;
        DevBuf<float> tmp;
        if (!on_device) {
            CK(tmp.ensure(nval));
            CK(cudaMemcpyAsync(tmp.p, co, nval * sizeof(float), cudaMemcpyHostToDevice, stream));
            src = tmp.p;
            stats.h2d_bytes += nval * sizeof(float);
        }
        k_retile<<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(src, d_grid.p, nmodel, npad, nfilt, layout);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        tmp.release();
        // size the per-(star, model) state for a star batch: at most 256 stars and ~1/5 of free HBM
        size_t fr = 0, tot = 0;
        CK(cudaMemGetInfo(&fr, &tot));
        const size_t per_star = (size_t)7 * sizeof(T) * npad;
        size_t budget = std::min<size_t>(fr / 5, (size_t)24 << 30);
        batch_cap = (int)std::max<size_t>(1, std::min<size_t>(256, budget / per_star));
        pool_cap = std::max<int64_t>(npad, std::min<int64_t>((int64_t)64 << 20, (int64_t)batch_cap * npad));
        CK(d_state.ensure((size_t)7 * batch_cap * npad));
        CK(d_stars.ensure((size_t)batch_cap * kStarStride));
        CK(d_star_int.ensure((size_t)batch_cap * SI_COUNT));
        CK(d_list.ensure((size_t)batch_cap));
        CK(d_cnt.ensure((size_t)batch_cap * ntile));
        CK(d_tot.ensure((size_t)batch_cap));
        CK(d_base.ensure((size_t)batch_cap));
        CK(d_red.ensure((size_t)batch_cap * kNumRed));
        CK(d_poolI.ensure((size_t)2 * pool_cap));
        CK(d_poolT.ensure((size_t)7 * pool_cap));
        h_stars.resize((size_t)batch_cap * kStarStride);
        h_int.resize((size_t)batch_cap * SI_COUNT);
        h_list.resize(batch_cap);
        h_red.resize((size_t)batch_cap * kNumRed);
        h_tot.resize(batch_cap);
        h_base.resize(batch_cap);
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
        h_ext.resize((size_t)batch_cap * nl * 3);
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
        o.ln_init = (T)std::log(opt->init_thresh);
        o.ltol = (T)opt->ltol;
        o.ln_sub = (T)std::log(opt->ltol_subthresh);
        o.ln_wt = (T)(opt->wt_thresh > 0 ? std::log(opt->wt_thresh) : -INFINITY);
        o.dim_prior = opt->dim_prior;
        max_iter = opt->max_iter > 0 ? opt->max_iter : 64;
        return BF_OK;
    }

    void phase_begin() { cudaEventRecord(evA, stream); }
    double phase_end() {  // call only right before/after a stream sync
        cudaEventRecord(evB, stream);
        cudaEventSynchronize(evB);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evA, evB);
        return ms;
    }

    // Runs the whole pipeline up to and including k_lnprob for `ns` stars already described by
    // h_stars/h_int (slots 0..ns-1).  On return the state arrays hold final values.
    int run_fit(int ns, const DevOpts<T>& o, int max_iter, int* n_mag, int* n_flux, int64_t* n_surv) {
        const StateArrays<T> st = state();
        const PoolArrays<T> pl = pool();
        // initial speculation: 2 mag iterations (what the reference needs in the common case)
        for (int s = 0; s < ns; s++) {
            h_int[s * SI_COUNT + SI_KSPEC] = std::min(2, max_iter);
            h_int[s * SI_COUNT + SI_ACTIVE] = 0;
            h_list[s] = s;
        }
        CK(cudaMemcpyAsync(d_stars.p, h_stars.data(), (size_t)ns * kStarStride * sizeof(T), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_star_int.p, h_int.data(), (size_t)ns * SI_COUNT * sizeof(int), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_list.p, h_list.data(), (size_t)ns * sizeof(int), cudaMemcpyHostToDevice, stream));
        stats.h2d_bytes += (size_t)ns * (kStarStride * sizeof(T) + (SI_COUNT + 1) * sizeof(int));
        if (nlabel > 0) {
            CK(cudaMemcpyAsync(d_ext.p, h_ext.data(), (size_t)ns * nlabel * 3 * sizeof(T), cudaMemcpyHostToDevice, stream));
            stats.h2d_bytes += (size_t)ns * nlabel * 3 * sizeof(T);
        }
        CK(cudaEventRecord(ev0, stream));
        k_reset_red<T><<<(ns * kNumRed + 255) / 256, 256, 0, stream>>>(d_red.p, d_list.p, ns, 0xffu);
        stats.kernel_launches++;

        // ---- magnitude fit with verified speculation of the iteration count ----
        std::vector<int> lst(h_list.begin(), h_list.begin() + ns);
        std::vector<char> exact(ns, 0);
        bool first_pass = true;
        while (!lst.empty()) {
            const int nl = (int)lst.size();
            if (!first_pass) {
                for (int k = 0; k < nl; k++) h_list[k] = lst[k];
                CK(cudaMemcpyAsync(d_list.p, h_list.data(), (size_t)nl * sizeof(int), cudaMemcpyHostToDevice, stream));
                CK(cudaMemcpyAsync(d_star_int.p, h_int.data(), (size_t)ns * SI_COUNT * sizeof(int), cudaMemcpyHostToDevice, stream));
                k_reset_red<T><<<(nl * kNumRed + 255) / 256, 256, 0, stream>>>(d_red.p, d_list.p, nl, 0x1fu);
                stats.kernel_launches++;
                stats.resweeps += nl;
            }
            SweepParams<T> sp;
            sp.grid = d_grid.p; sp.npad = npad; sp.nmodel = nmodel; sp.stars = d_stars.p;
            sp.star_int = d_star_int.p; sp.list = d_list.p; sp.nlist = nl; sp.o = o; sp.st = st; sp.red = d_red.p;
            phase_begin();
            kt->magfit(sp, stream);
            CK(cudaGetLastError());
            CK(cudaEventRecord(evB, stream));
            stats.kernel_launches++; stats.magfit_launches++; stats.magfit_star_passes += nl;
            CK(cudaMemcpyAsync(h_red.data(), d_red.p, (size_t)ns * kNumRed * sizeof(U), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            {
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, evA, evB));
                stats.ms_magfit += ms;
            }
            stats.d2h_bytes += (size_t)ns * kNumRed * sizeof(U);
            std::vector<int> next;
            for (int k = 0; k < nl; k++) {
                const int s = lst[k];
                int& ksp = h_int[s * SI_COUNT + SI_KSPEC];
                n_mag[s] = ksp;
                if (exact[s]) continue;
                const U* r = &h_red[(size_t)s * kNumRed];
                // brutus/fitting.py:252-263 restated on the two max-reductions
                const bool conv_prev = ksp >= 2 && !(Enc<T>::dec(r[RED_B0]) > Enc<T>::dec(r[RED_L0]) + o.ln_init);
                const bool conv_last = !(Enc<T>::dec(r[RED_B1]) > Enc<T>::dec(r[RED_L1]) + o.ln_init);
                if (conv_prev) {           // the reference would have stopped one iteration earlier
                    ksp -= 1; exact[s] = 1; next.push_back(s);
                } else if (!conv_last && ksp < max_iter) {
                    ksp = std::min(ksp + 2, max_iter); next.push_back(s);
                }
            }
            lst.swap(next);
            first_pass = false;
        }

        // ---- cull (brutus/fitting.py:758-768): count, scan, ordered write into the survivor pool ----
        for (int s = 0; s < ns; s++) h_list[s] = s;
        CK(cudaMemcpyAsync(d_list.p, h_list.data(), (size_t)ns * sizeof(int), cudaMemcpyHostToDevice, stream));
        FlagParams<T> fp;
        fp.stars = d_stars.p; fp.st = st; fp.red = d_red.p; fp.o = o; fp.npad = npad; fp.nmodel = nmodel;
        fp.list = d_list.p; fp.nlist = ns; fp.ntile = ntile; fp.cnt = d_cnt.p; fp.base = d_base.p;
        fp.out_model = pl.model; fp.out_star = pl.star;
        phase_begin();
        k_count<T, FLAG_CULL><<<dim3(ntile, ns), kTile, 0, stream>>>(fp);
        k_scan_tiles<<<ns, 1024, 0, stream>>>(d_cnt.p, d_list.p, ntile, d_tot.p);
        stats.kernel_launches += 2;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_tot.data(), d_tot.p, (size_t)ns * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
        stats.ms_select += phase_end();
        for (int s = 0; s < ns; s++) { n_surv[s] = h_tot[s]; n_flux[s] = 0; stats.survivors += h_tot[s]; }

        // groups of consecutive stars whose survivors fit in the pool together
        int g0 = 0;
        while (g0 < ns) {
            int g1 = g0;
            int64_t tot = 0;
            while (g1 < ns && (g1 == g0 || tot + h_tot[g1] <= pool_cap)) { h_base[g1] = tot; tot += h_tot[g1]; g1++; }
            const int ng = g1 - g0;
            if (tot > pool_cap) { err = "internal: survivor pool too small"; return BF_E_NOMEM; }
            CK(cudaMemcpyAsync(d_base.p + g0, h_base.data() + g0, (size_t)ng * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
            fp.list = d_list.p + g0; fp.nlist = ng;
            phase_begin();
            k_write<T, FLAG_CULL><<<dim3(ntile, ng), kTile, 0, stream>>>(fp);
            k_flux_init<T><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(pl, tot, st, npad);
            stats.kernel_launches += 2;
            CK(cudaGetLastError());
            // ---- flux-space iterations until every star of the group converges (:781-803) ----
            std::vector<int> act;
            for (int s = g0; s < g1; s++) { h_int[s * SI_COUNT + SI_ACTIVE] = 1; act.push_back(s); }
            bool first = true;
            while (!act.empty()) {
                const int na = (int)act.size();
                for (int k = 0; k < na; k++) h_list[k] = act[k];
                // d_list is reused for the active list; the group list is restored afterwards
                CK(cudaMemcpyAsync(d_list.p, h_list.data(), (size_t)na * sizeof(int), cudaMemcpyHostToDevice, stream));
                CK(cudaMemcpyAsync(d_star_int.p, h_int.data(), (size_t)ns * SI_COUNT * sizeof(int), cudaMemcpyHostToDevice, stream));
                k_reset_red<T><<<(na * kNumRed + 255) / 256, 256, 0, stream>>>(d_red.p, d_list.p, na, (1u << RED_FL) | (1u << RED_FB));
                FluxParams<T> xp;
                xp.grid = d_grid.p; xp.npad = npad; xp.nmodel = nmodel; xp.stars = d_stars.p;
                xp.star_int = d_star_int.p; xp.o = o; xp.pool = pl; xp.nsv = tot; xp.red = d_red.p;
                xp.nit = first ? std::min(2, max_iter) : 1;
                kt->flux(xp, stream);
                stats.kernel_launches += 2;
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(h_red.data(), d_red.p, (size_t)ns * kNumRed * sizeof(U), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                stats.d2h_bytes += (size_t)ns * kNumRed * sizeof(U);
                std::vector<int> next;
                for (int k = 0; k < na; k++) {
                    const int s = act[k];
                    n_flux[s] += xp.nit;
                    const U* r = &h_red[(size_t)s * kNumRed];
                    // "lerr > ltol" (:781, :798-799) restated on the two max-reductions
                    const bool more = Enc<T>::dec(r[RED_FB]) > Enc<T>::dec(r[RED_FL]) + o.ln_sub;
                    if (more && n_flux[s] < max_iter) next.push_back(s);
                    else h_int[s * SI_COUNT + SI_ACTIVE] = 0;
                }
                act.swap(next);
                first = false;
            }
            k_scatter<T><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(pl, tot, st, npad);
            stats.kernel_launches++;
            CK(cudaGetLastError());
            stats.ms_flux += phase_end();
            for (int s = 0; s < ns; s++) h_list[s] = s;
            CK(cudaMemcpyAsync(d_list.p, h_list.data(), (size_t)ns * sizeof(int), cudaMemcpyHostToDevice, stream));
            g0 = g1;
        }

        // ---- lnlike / lnprob for every model, per-star max ----
        LnprobParams<T> lp;
        lp.stars = d_stars.p; lp.st = st; lp.red = d_red.p; lp.o = o; lp.npad = npad; lp.nmodel = nmodel;
        lp.list = d_list.p; lp.labels = d_labels.p; lp.ext = d_ext.p; lp.nlabel = nlabel;
        k_lnprob<T><<<dim3(ntile, ns), kTile, 0, stream>>>(lp);
        stats.kernel_launches++;
        CK(cudaGetLastError());
        return BF_OK;
    }

    int fill_rows(int ns, const double* flux, const double* errv, const uint8_t* mask, const double* par,
                  const double* perr, const double* ext_mean, const double* ext_std, int apply_clip,
                  int32_t* ndim_out, uint8_t* mask_out) {
        StarPrep sp;
        for (int s = 0; s < ns; s++) {
            prep_star(flux + (size_t)s * nfilt, errv + (size_t)s * nfilt, mask + (size_t)s * nfilt, nfilt,
                      par ? par[s] : NAN, perr ? perr[s] : NAN, apply_clip, sp);
            for (int k = 0; k < kStarStride; k++) h_stars[(size_t)s * kStarStride + k] = (T)sp.row[k];
            h_int[s * SI_COUNT + SI_NDIM] = sp.ndim;
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
        fill_rows(1, flux, errv, mask, &par, &perr, nullptr, nullptr, 0, &nd, mask_out);
        int nm, nf; int64_t nsv;
        rc = run_fit(1, o, max_iter, &nm, &nf, &nsv);
        nlabel = saved_labels;
        if (rc) return rc;
        const size_t per = icov ? 14 : 5;
        CK(d_out.ensure((size_t)nmodel * per));
        RecordParams<T, double> rp{};
        rp.grid = d_grid.p; rp.npad = npad; rp.nmodel = nmodel; rp.stars = d_stars.p; rp.o = o; rp.st = state();
        rp.sel_model = nullptr; rp.sel_star = nullptr; rp.nrec = 0; rp.star_slot = 0; rp.ld = 0; rp.nrows = 11; rp.o_idx = nullptr;
        double* b = d_out.p;
        rp.o_lnl = b; rp.o_chi2 = b + nmodel; rp.o_scale = b + 2 * nmodel; rp.o_av = b + 3 * nmodel; rp.o_rv = b + 4 * nmodel;
        rp.o_icov = icov ? b + 5 * nmodel : nullptr;
        phase_begin();
        kt->records_full(rp, stream);
        stats.kernel_launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev1, stream));
        stats.ms_select += phase_end();
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        stats.ms_device = ms;
        const size_t nb = (size_t)nmodel * sizeof(double);
        CK(cudaMemcpy(lnl, rp.o_lnl, nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(chi2, rp.o_chi2, nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(scale, rp.o_scale, nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(av, rp.o_av, nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rv, rp.o_rv, nb, cudaMemcpyDeviceToHost));
        if (icov) CK(cudaMemcpy(icov, rp.o_icov, nb * 9, cudaMemcpyDeviceToHost));
        stats.d2h_bytes += nb * per;
        if (diag) { diag[0] = nd; diag[1] = nm; diag[2] = nf; diag[3] = nsv; }
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

    int sweep_batch(int64_t nstar, const double* flux, const double* errv, const uint8_t* mask,
                    const double* par, const double* perr, const double* ext_mean, const double* ext_std,
                    const bf_options* opt, int record_rows, int32_t* ndim, int32_t* n_iter, int64_t* n_surv,
                    double* max_lnprob, int64_t* offsets, bf_records* out) override {
        CK(cudaSetDevice(device));
        if (!kt) { err = "bf_sweep_batch: no grid (call bf_set_grid)"; return BF_E_NOGRID; }
        if (nstar < 0 || (nstar > 0 && (!flux || !errv || !mask)) || !opt || !offsets || !out) { err = "bf_sweep_batch: null argument"; return BF_E_INVALID; }
        if (record_rows != 3 && record_rows != 5 && record_rows != 11) { err = "bf_sweep_batch: record_rows must be 3, 5 or 11"; return BF_E_INVALID; }
        stats = bf_stats{};
        DevOpts<T> o; int max_iter;
        int rc = make_opts(opt, o, max_iter);
        if (rc) return rc;
        const StateArrays<T> st = state();
        const PoolArrays<T> pl = pool();
        std::vector<int> nm(batch_cap), nf(batch_cap);
        std::vector<int64_t> nsv(batch_cap);
        int64_t written = 0;
        int grp = 0;
        offsets[0] = 0;
        for (int64_t s0 = 0; s0 < nstar; s0 += batch_cap) {
            const int ns = (int)std::min<int64_t>(batch_cap, nstar - s0);
            fill_rows(ns, flux + (size_t)s0 * nfilt, errv + (size_t)s0 * nfilt, mask + (size_t)s0 * nfilt,
                      par ? par + s0 : nullptr, perr ? perr + s0 : nullptr,
                      ext_mean ? ext_mean + (size_t)s0 * nlabel : nullptr,
                      ext_std ? ext_std + (size_t)s0 * nlabel : nullptr, opt->apply_parallax_clip,
                      ndim ? ndim + s0 : nullptr, nullptr);
            rc = run_fit(ns, o, max_iter, nm.data(), nf.data(), nsv.data());
            if (rc) return rc;
            // ---- first selection of lnpost (brutus/fitting.py:988-991), ordered compaction ----
            FlagParams<T> fp;
            fp.stars = d_stars.p; fp.st = st; fp.red = d_red.p; fp.o = o; fp.npad = npad; fp.nmodel = nmodel;
            fp.list = d_list.p; fp.nlist = ns; fp.ntile = ntile; fp.cnt = d_cnt.p; fp.base = d_base.p;
            fp.out_model = pl.model; fp.out_star = pl.star;
            phase_begin();
            k_count<T, FLAG_SELECT><<<dim3(ntile, ns), kTile, 0, stream>>>(fp);
            k_scan_tiles<<<ns, 1024, 0, stream>>>(d_cnt.p, d_list.p, ntile, d_tot.p);
            stats.kernel_launches += 2;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(h_tot.data(), d_tot.p, (size_t)ns * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(h_red.data(), d_red.p, (size_t)ns * kNumRed * sizeof(U), cudaMemcpyDeviceToHost, stream));
            stats.ms_select += phase_end();
            for (int s = 0; s < ns; s++) {
                if (n_iter) { n_iter[2 * (s0 + s)] = nm[s]; n_iter[2 * (s0 + s) + 1] = nf[s]; }
                if (n_surv) n_surv[s0 + s] = nsv[s];
                if (max_lnprob) {
                    T v = Enc<T>::dec(h_red[(size_t)s * kNumRed + RED_LNP]);
                    max_lnprob[s0 + s] = (v <= Num<T>::kNegBig) ? -1e300 : (double)v;
                }
                offsets[s0 + s + 1] = offsets[s0 + s] + h_tot[s];
                stats.selected += h_tot[s];
            }
            // groups of stars whose records fit in the pool: k_write + k_records into a device staging
            // buffer, then an asynchronous D2H on the copy stream that overlaps the next batch's compute
            int g0 = 0;
            while (g0 < ns) {
                int g1 = g0;
                int64_t tot = 0;
                while (g1 < ns && (g1 == g0 || tot + h_tot[g1] <= pool_cap)) { h_base[g1] = tot; tot += h_tot[g1]; g1++; }
                const int ng = g1 - g0;
                if (tot > 0) {
                    const int buf = grp & 1;
                    grp++;
                    CK(cudaStreamWaitEvent(stream, ev_cp[buf], 0));  // staging buffer free again?
                    CK(cudaMemcpyAsync(d_base.p + g0, h_base.data() + g0, (size_t)ng * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
                    fp.list = d_list.p + g0; fp.nlist = ng;
                    phase_begin();
                    k_write<T, FLAG_SELECT><<<dim3(ntile, ng), kTile, 0, stream>>>(fp);
                    CK(d_stage[buf].ensure((size_t)tot * (sizeof(int) + 11 * sizeof(T))));
                    RecordParams<T, T> rp{};
                    rp.grid = d_grid.p; rp.npad = npad; rp.nmodel = nmodel; rp.stars = d_stars.p; rp.o = o; rp.st = st;
                    rp.sel_model = pl.model; rp.sel_star = pl.star; rp.nrec = tot; rp.star_slot = 0;
                    T* b = (T*)d_stage[buf].p;                                   // [11][tot] rows, then idx
                    rp.o_idx = (int*)(d_stage[buf].p + (size_t)11 * tot * sizeof(T));
                    rp.ld = tot; rp.nrows = record_rows;
                    rp.o_lnl = b; rp.o_scale = b + tot; rp.o_av = b + 2 * tot; rp.o_chi2 = b + 3 * tot;
                    rp.o_rv = b + 4 * tot; rp.o_icov = b + 5 * tot;
                    kt->records(rp, stream);
                    stats.kernel_launches += 2;
                    CK(cudaGetLastError());
                    CK(cudaEventRecord(ev_rec[buf], stream));
                    CK(cudaEventRecord(evB, stream));
                    if (!opt->skip_d2h) {
                    rc = ensure_arena(written + tot, written);
                    if (rc) return rc;
                    CK(cudaStreamWaitEvent(copy_stream, ev_rec[buf], 0));
                    CK(cudaMemcpyAsync(arena + (size_t)11 * arena_cap * sizeof(T) + (size_t)written * sizeof(int),
                                       rp.o_idx, (size_t)tot * sizeof(int), cudaMemcpyDeviceToHost, copy_stream));
                    CK(cudaMemcpy2DAsync(arena + (size_t)written * sizeof(T),
                                         (size_t)arena_cap * sizeof(T), b, (size_t)tot * sizeof(T),
                                         (size_t)tot * sizeof(T), record_rows, cudaMemcpyDeviceToHost, copy_stream));
                    CK(cudaEventRecord(ev_cp[buf], copy_stream));
                    stats.d2h_bytes += (size_t)tot * (sizeof(int) + record_rows * sizeof(T));
                    }
                    CK(cudaEventSynchronize(evB));
                    float msr = 0.f;
                    CK(cudaEventElapsedTime(&msr, evA, evB));
                    stats.ms_select += msr;
                }
                written += tot;
                g0 = g1;
            }
            CK(cudaEventRecord(ev1, stream));
            CK(cudaEventSynchronize(ev1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev0, ev1));
            stats.ms_device += ms;
        }
        CK(cudaStreamSynchronize(copy_stream));
        out->n = opt->skip_d2h ? 0 : written;
        out->stride = arena_cap;
        out->elem_size = (int32_t)sizeof(T);
        out->nrows = record_rows;
        out->model_idx = arena ? (const int32_t*)(arena + (size_t)11 * arena_cap * sizeof(T)) : nullptr;
        out->rows = (const void*)arena;
        return BF_OK;
    }
};

}  // namespace bf

// =================================================================================================
// C ABI
// =================================================================================================
struct bf_handle {
    bf::EngineBase* eng;
};

extern "C" {

void bf_default_options(bf_options* o) {
    if (!o) return;
    o->avlim[0] = 0.; o->avlim[1] = 20.;
    o->av_gauss[0] = 0.; o->av_gauss[1] = 1e6;
    o->rvlim[0] = 1.; o->rvlim[1] = 8.;
    o->rv_gauss[0] = 3.32; o->rv_gauss[1] = 0.18;
    o->ltol = 3e-2; o->ltol_subthresh = 1e-2; o->init_thresh = 5e-3; o->wt_thresh = 1e-3;
    o->dim_prior = 1; o->max_iter = 0; o->apply_parallax_clip = 1; o->skip_d2h = 0;
}

int bf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* bf_version(void) { return "brutus_b200 0.1.0 (sm_100a)"; }

int bf_create(int device, int precision, bf_handle** out) {
    if (!out) { bf::g_create_error = "bf_create: null out pointer"; return BF_E_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        bf::g_create_error = std::string("bf_create: no CUDA device available (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return BF_E_CUDA;
    }
    if (device < 0 || device >= n) { bf::g_create_error = "bf_create: device ordinal out of range"; return BF_E_INVALID; }
    bf::EngineBase* eng = nullptr;
    if (precision == BF_PRECISION_F32) eng = new bf::Engine<float>();
    else if (precision == BF_PRECISION_F64) eng = new bf::Engine<double>();
    else { bf::g_create_error = "bf_create: unknown precision"; return BF_E_INVALID; }
    eng->device = device;
    eng->precision = precision;
    int rc = precision == BF_PRECISION_F32 ? static_cast<bf::Engine<float>*>(eng)->init()
                                           : static_cast<bf::Engine<double>*>(eng)->init();
    if (rc) { bf::g_create_error = eng->err; delete eng; return rc; }
    *out = new bf_handle{eng};
    return BF_OK;
}

int bf_destroy(bf_handle* h) {
    if (!h) return BF_OK;
    delete h->eng;
    delete h;
    return BF_OK;
}

const char* bf_last_error(const bf_handle* h) { return h ? h->eng->err.c_str() : bf::g_create_error.c_str(); }

int bf_set_grid(bf_handle* h, const float* coeffs, int64_t nmodel, int32_t nfilt, int32_t layout) {
    if (!h) return BF_E_INVALID;
    return h->eng->set_grid(coeffs, nmodel, nfilt, layout, false);
}

int bf_set_grid_device(bf_handle* h, const void* d_coeffs, int64_t nmodel, int32_t nfilt, int32_t layout) {
    if (!h) return BF_E_INVALID;
    return h->eng->set_grid((const float*)d_coeffs, nmodel, nfilt, layout, true);
}

int bf_set_labels(bf_handle* h, const double* labels, int32_t nlabel) {
    if (!h) return BF_E_INVALID;
    return h->eng->set_labels(labels, nlabel);
}

int bf_loglike_full(bf_handle* h, const double* flux, const double* err, const uint8_t* mask, double parallax,
                    double parallax_err, const bf_options* opt, double* lnl, double* chi2, double* scale,
                    double* av, double* rv, double* icov, uint8_t* mask_clean_out, int64_t* diag) {
    if (!h) return BF_E_INVALID;
    return h->eng->loglike_full(flux, err, mask, parallax, parallax_err, opt, lnl, chi2, scale, av, rv, icov,
                                mask_clean_out, diag);
}

int bf_sweep_batch(bf_handle* h, int64_t nstar, const double* flux, const double* err, const uint8_t* mask,
                   const double* parallax, const double* parallax_err, const double* ext_mean,
                   const double* ext_std, const bf_options* opt, int32_t record_rows, int32_t* ndim,
                   int32_t* n_iter, int64_t* n_surv, double* max_lnprob, int64_t* offsets, bf_records* out) {
    if (!h) return BF_E_INVALID;
    return h->eng->sweep_batch(nstar, flux, err, mask, parallax, parallax_err, ext_mean, ext_std, opt,
                               record_rows, ndim, n_iter, n_surv, max_lnprob, offsets, out);
}

int bf_flush_l2(bf_handle* h) {
    if (!h) return BF_E_INVALID;
    return h->eng->flush_l2();
}

int bf_get_stats(const bf_handle* h, bf_stats* out) {
    if (!h || !out) return BF_E_INVALID;
    *out = h->eng->stats;
    return BF_OK;
}

}  // extern "C"
