// common.cuh -- shared definitions for the brutus_b200 CUDA kernels (sm_100a).
//
// Data layout in HBM:
//   grid   float32 [3][NB][npad]        coefficient-major, then band, model axis contiguous
//                                        (c = 0: mag0 @ 1 kpc, 1: R0, 2: dR/dRv; brutus/utils.py:293-298)
//   stars  T [batch][kStarStride]       per-star normalised photometry, see StarRow below
//   state  T [batch][npad] x 7          chi2, scale, s_den, av, rv, lnl, lnprob per (star, model)
//   red    U [batch][kNumRed]           per-star max-reductions, order-preserving unsigned encoding
//   pool   survivor records (SoA)       models that survive the cull, contiguous per star
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace bf {

constexpr int kMaxFilt = 16;
constexpr int kTile = 256;       // models per CTA (= threads per CTA) in the per-model kernels
constexpr int kStarChunk = 32;   // stars looped over by one CTA of the mag-fit sweep

// ---- star row (units of T) -------------------------------------------------------------------
// cm  : m_j - mbar, centred observed magnitude (0 for masked / non-positive-flux bands)
// u   : mag-space weight 1/var(m_j) (brutus/fitting.py:723); 0 for masked / non-positive flux (:725)
// al  : d_j / sigma_j                    (signed S/N; 0 for masked)
// be  : d*_j / sigma_j, d* = d_j if d_j > 0 else 10^(-0.4 mbar) (reference flux for the band)
constexpr int SR_CM = 0, SR_U = kMaxFilt, SR_AL = 2 * kMaxFilt, SR_BE = 3 * kMaxFilt;
constexpr int SR_SC = 4 * kMaxFilt;
enum StarScalar {
    SC_MBAR = 0,   // mean magnitude of the positive-flux bands
    SC_S,          // sum_j u_j                       (s_den of brutus/fitting.py:158-162)
    SC_PAR,        // parallax (0 if none)
    SC_PIVAR,      // 1/parallax_err^2 (0 if none)    (brutus/fitting.py:750-756)
    SC_LNORM,      // -ln(2^(k/2) Gamma(k/2)), k = Ndim-3   (brutus/utils.py:170)
    SC_KHM1,       // k/2 - 1
    SC_GCONST,     // -0.5 (Ndim ln 2pi + sum ln sigma^2)   (brutus/fitting.py:806-807)
    SC_SPAPPLY,    // 1 if the rough scale-parallax prior applies (brutus/pdf.py:209)
    SC_SMEAN,      // brutus/pdf.py:255
    SC_SVAR,       // s_std^2, brutus/pdf.py:256
    SC_COUNT
};
constexpr int kStarStride = SR_SC + 16;

// per-star ints
enum StarInt { SI_NDIM = 0, SI_KSPEC, SI_ACTIVE, SI_COUNT = 4 };

// per-star reductions
enum Red {
    RED_L0 = 0,  // max logwt at mag iteration kspec-1          (brutus/fitting.py:246-249)
    RED_B0,      // max logwt over models with max(|dAv|,|dRv|) >= tol at iteration kspec-1
    RED_L1,      // same two at iteration kspec
    RED_B1,
    RED_LP,      // max lnl_p after the mag fit                 (brutus/fitting.py:758)
    RED_FL,      // max lnl_new in the current flux iteration   (brutus/fitting.py:798)
    RED_FB,      // max lnl_new over survivors with |lnl_new - lnl_old| > ltol
    RED_LNP,     // max lnprob                                  (brutus/fitting.py:990)
    kNumRed
};

// ---- order-preserving float -> unsigned encoding (so atomicMax works on floats) ---------------
template <typename T> struct Enc;
template <> struct Enc<float> {
    using U = unsigned int;
    __host__ __device__ static U enc(float f) {
#ifdef __CUDA_ARCH__
        U u = __float_as_uint(f);
#else
        U u; memcpy(&u, &f, 4);
#endif
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
    __host__ __device__ static float dec(U e) {
        U u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
#ifdef __CUDA_ARCH__
        return __uint_as_float(u);
#else
        float f; memcpy(&f, &u, 4); return f;
#endif
    }
};
template <> struct Enc<double> {
    using U = unsigned long long;
    __host__ __device__ static U enc(double f) {
#ifdef __CUDA_ARCH__
        U u = (U)__double_as_longlong(f);
#else
        U u; memcpy(&u, &f, 8);
#endif
        return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    }
    __host__ __device__ static double dec(U e) {
        U u = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
#ifdef __CUDA_ARCH__
        return __longlong_as_double((long long)u);
#else
        double f; memcpy(&f, &u, 8); return f;
#endif
    }
};

// ---- numeric helpers, float = throughput path, double = verification path ----------------------
template <typename T> struct Num;
template <> struct Num<float> {
    static constexpr float kNegBig = -3.0e38f;   // stands in for the reference's -1e300 sentinel
    __device__ static float exp2(float x) { return exp2f(x); }          // MUFU.EX2
    __device__ static float log(float x) { return __logf(x); }          // MUFU.LG2 * ln2
    __device__ static float sqrt(float x) { return sqrtf(x); }
    __device__ static float sqrt_fast(float x) { return x * rsqrtf(x); }   // x > 0 (scale >= 1e-20)
    __device__ static float max(float a, float b) { return fmaxf(a, b); }  // FMNMX
    __device__ static float min(float a, float b) { return fminf(a, b); }
    __device__ static float div_fast(float a, float b) { return __fdividef(a, b); }
    __device__ static float div(float a, float b) { return a / b; }
    __device__ static float neg_inf() { return -CUDART_INF_F; }
    __device__ static bool finite(float x) { return isfinite(x); }
};
template <> struct Num<double> {
    static constexpr double kNegBig = -1.0e300;
    __device__ static double exp2(double x) { return ::exp2(x); }
    __device__ static double log(double x) { return ::log(x); }
    __device__ static double sqrt(double x) { return ::sqrt(x); }
    __device__ static double sqrt_fast(double x) { return ::sqrt(x); }
    __device__ static double max(double a, double b) { return a > b ? a : b; }
    __device__ static double min(double a, double b) { return a < b ? a : b; }
    __device__ static double div_fast(double a, double b) { return a / b; }
    __device__ static double div(double a, double b) { return a / b; }
    __device__ static double neg_inf() { return -CUDART_INF; }
    __device__ static bool finite(double x) { return isfinite(x); }
};

template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T tabs(T a) { return a < T(0) ? -a : a; }

// ---- fit options on the device ------------------------------------------------------------------
template <typename T> struct DevOpts {
    T Abar, PA, Rbar, PR;          // prior means and precisions   (brutus/fitting.py:146-148)
    T avmin, avmax, rvmin, rvmax;  // clip limits                   (:143-144)
    T mtol;                        // 2.5 * ltol                    (:732)
    T ln_init;                     // ln(init_thresh)               (:150, :758)
    T ltol, ln_sub;                // (:781, :780)
    T ln_wt;                       // ln(wt_thresh)                 (:990)
    int dim_prior;                 // (:813)
};

constexpr double kC2 = 1.3287712379549449;    // 0.4 * log2(10): 10^(0.4 x) = 2^(kC2 x)
constexpr double kFac = -0.9210340371976184;  // -0.4 * ln(10), brutus/utils.py:328

// ---- kernel parameter blocks ----------------------------------------------------------------------
template <typename T> struct StateArrays {
    T *chi2, *scale, *sden, *av, *rv, *lnl, *lnprob;   // each [batch][npad]
};

template <typename T> struct SweepParams {
    const float* grid;
    int64_t npad, nmodel;
    const T* stars;          // [batch][kStarStride]
    const int* star_int;     // [batch][SI_COUNT]
    const int* list;         // star slots processed by this launch
    int nlist;
    DevOpts<T> o;
    StateArrays<T> st;
    typename Enc<T>::U* red; // [batch][kNumRed]
};

template <typename T> struct PoolArrays {
    int *model, *star;                       // [cap]
    T *av, *rv, *eta, *lold, *chi2, *scale, *sden;
};

template <typename T> struct FluxParams {
    const float* grid;
    int64_t npad, nmodel;
    const T* stars;
    const int* star_int;
    DevOpts<T> o;
    PoolArrays<T> pool;
    int64_t nsv;
    int nit;                 // flux iterations executed by this launch (2 first, then 1)
    typename Enc<T>::U* red;
};

// O = element type of the outputs: double for the full-length B1 arrays (the reference returns
// float64), T for the compacted B2 records (no point shipping more bits than were computed).
template <typename T, typename O> struct RecordParams {
    const float* grid;
    int64_t npad, nmodel;
    const T* stars;
    DevOpts<T> o;
    StateArrays<T> st;
    // mode A (compacted records): nrec records (sel_model/sel_star)
    const int* sel_model;
    const int* sel_star;
    int64_t nrec;
    // mode B (full-length, one star slot): sel_model == nullptr
    int star_slot;
    // outputs.  Mode A: rows of a [11][ld] matrix: lnl, scale, av, chi2, rv, icov(ss,sa,sr,aa,ar,rr);
    // only the first `nrows` are produced (3, 5 or 11).  Mode B: separate arrays, icov 9 per model.
    O *o_lnl, *o_chi2, *o_scale, *o_av, *o_rv, *o_icov;
    int64_t ld;
    int nrows;
    int* o_idx;   // mode A: copy of sel_model next to the rows (the pool is reused by the next batch)
};

// Kernel launchers instantiated once per band count (inst.cu, -DBF_NB=n).
template <typename T> struct KTable {
    void (*magfit)(const SweepParams<T>&, cudaStream_t);
    void (*flux)(const FluxParams<T>&, cudaStream_t);
    void (*records)(const RecordParams<T, T>&, cudaStream_t);          // compacted records (B2)
    void (*records_full)(const RecordParams<T, double>&, cudaStream_t); // full-length float64 (B1)
};

}  // namespace bf
