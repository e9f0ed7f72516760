// common.cuh -- shared definitions for the brutus_b200 CUDA kernels (sm_100a).
//
// Data layout in HBM:
//   grid   float32 [3][NB][npad]        coefficient-major, then band, model axis contiguous: the full-grid
//                                        sweep reads it fully coalesced (thread = model)
//                                        (c = 0: mag0 @ 1 kpc, 1: R0, 2: dR/dRv; brutus/utils.py:293-298)
//   rows   float32 [npad][row_stride]   the same coefficients model-major ([c][band] inside a row, row padded
//                                        to 16 B): per-candidate gathers touch 3-4 sectors instead of 3*NB
//   stars  T [batch][kStarStride]       per-star normalised photometry, see the star row below
//   cand   u32 [batch][npad/32]         candidate bitmap written by the sweep (superset of everything
//                                        that can survive the cull or pass the selection threshold)
//   red    U [batch][kNumRed]           per-star max-reductions, order-preserving unsigned encoding
//   pool   candidate records (SoA)      one record per candidate (star, model), ascending (star, model)
//
// Nothing of size O(Nmodel x stars) other than the 1-bit candidate map is ever written.
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
    SC_SLACK,      // selection-candidate slack of the sweep (+inf: every model is a candidate)
    SC_COUNT
};
constexpr int kStarStride = SR_SC + 16;

// per-star ints
enum StarInt { SI_NDIM = 0, SI_KSPEC, SI_ACTIVE, SI_NFLUX, SI_COUNT = 4 };

// per-star reductions
enum Red {
    RED_L0 = 0,  // max logwt at mag iteration kspec-1          (brutus/fitting.py:246-249)
    RED_B0,      // max logwt over models with max(|dAv|,|dRv|) >= tol at iteration kspec-1
    RED_L1,      // same two at iteration kspec
    RED_B1,
    RED_LP,      // max lnl_p after the mag fit                 (brutus/fitting.py:758)
    RED_M0,      // max provisional lnprob (mag-fit values for every model)
    RED_FL,      // max lnl_new in the current flux iteration   (brutus/fitting.py:798)
    RED_FB,      // max lnl_new over survivors with |lnl_new - lnl_old| > ltol
    RED_LNP,     // max lnprob                                  (brutus/fitting.py:990)
    RED_P1,      // max of lnlike + priors at the MLE over the first selection   (brutus/fitting.py:1015)
    RED_P2,      // max final lnprob over the second selection                    (brutus/fitting.py:2034)
    RED_NCHI,    // max of -chi2 (incl. the parallax term) over the second selection (:2030, :2035)
    kNumRed
};
constexpr int kSweepRed = 6;     // RED_L0 .. RED_M0 are produced by the sweep

// candidate record flags
constexpr int kFlagSurv = 1;     // survived the cull (brutus/fitting.py:758-759)

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


// ---- two-lane values --------------------------------------------------------------------------------
template <typename T> struct P2;
template <> struct __align__(8) P2<float> { unsigned long long v; };
template <> struct __align__(16) P2<double> { double x, y; };

__device__ __forceinline__ P2<float> mk2(float a, float b) {
    P2<float> r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ P2<double> mk2(double a, double b) { P2<double> r; r.x = a; r.y = b; return r; }
__device__ __forceinline__ float lo2(P2<float> p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); return a; }
__device__ __forceinline__ float hi2(P2<float> p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); return b; }
__device__ __forceinline__ double lo2(P2<double> p) { return p.x; }
__device__ __forceinline__ double hi2(P2<double> p) { return p.y; }
template <typename T> __device__ __forceinline__ P2<T> bc2(T a) { return mk2(a, a); }   // folds into a .F32 broadcast operand
__device__ __forceinline__ P2<float> fma2(P2<float> a, P2<float> b, P2<float> c) {
    P2<float> r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r;
}
__device__ __forceinline__ P2<float> mul2(P2<float> a, P2<float> b) {
    P2<float> r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
}
__device__ __forceinline__ P2<float> add2(P2<float> a, P2<float> b) {
    P2<float> r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
}
__device__ __forceinline__ P2<float> sub2(P2<float> a, P2<float> b) {
    P2<float> r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
}
__device__ __forceinline__ P2<double> fma2(P2<double> a, P2<double> b, P2<double> c) { return mk2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ P2<double> mul2(P2<double> a, P2<double> b) { return mk2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ P2<double> add2(P2<double> a, P2<double> b) { return mk2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ P2<double> sub2(P2<double> a, P2<double> b) { return mk2(a.x - b.x, a.y - b.y); }
template <typename T> __device__ __forceinline__ T hsum2(P2<T> p) { return lo2(p) + hi2(p); }
// pair (2p, 2p+1) of a star-row array (shared or global memory, 8/16-byte aligned)
__device__ __forceinline__ P2<float> ld2(const float* p) { P2<float> r; r.v = *reinterpret_cast<const unsigned long long*>(p); return r; }
__device__ __forceinline__ P2<double> ld2(const double* p) { return mk2(p[0], p[1]); }


// ---- per-star aggregation inside a CTA of candidate records ------------------------------------------
// Records are sorted by star, so most CTAs see one or two stars.  Values for the star of the CTA's first
// record are combined in shared memory and published by ONE global atomic per CTA; records of other
// stars (CTAs that straddle a star boundary) fall back to warp-level / per-thread global atomics.
// Without this, the per-star global atomics of a few million consecutive records serialise on one or
// two L2 addresses.  All threads of the CTA must call these (they contain __syncthreads()).
template <typename T>
__device__ __forceinline__ void cta_star_max(typename Enc<T>::U* red, int which, int slot, bool have, T v) {
    using U = typename Enc<T>::U;
    __shared__ U s_m;
    __shared__ int s_slotA;
    if (threadIdx.x == 0) { s_m = Enc<T>::enc(Num<T>::neg_inf()); s_slotA = slot; }
    __syncthreads();
    const int slotA = s_slotA;
    const bool mine = have && slot == slotA;
    T w = warp_max((mine && v == v) ? v : Num<T>::neg_inf());
    if ((threadIdx.x & 31) == 0 && w > Num<T>::neg_inf()) atomicMax(&s_m, Enc<T>::enc(w));
    if (have && slot != slotA && v == v) atomicMax(&red[(int64_t)slot * kNumRed + which], Enc<T>::enc(v));
    __syncthreads();
    if (threadIdx.x == 0 && slotA >= 0) atomicMax(&red[(int64_t)slotA * kNumRed + which], s_m);
    __syncthreads();
}

__device__ __forceinline__ void cta_star_count(int* cnt, int slot, bool flag) {
    __shared__ int s_slotA2;
    if (threadIdx.x == 0) s_slotA2 = slot;
    __syncthreads();
    const int slotA = s_slotA2;
    const int n = __syncthreads_count(flag && slot == slotA);
    if (flag && slot != slotA) atomicAdd(&cnt[slot], 1);
    if (threadIdx.x == 0 && n > 0) atomicAdd(&cnt[slotA], n);
}

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
constexpr double kCandMargin = 1e-2;          // the sweep's cull-candidate test is this much looser than
                                              // the exact test applied when candidates are re-fitted

__host__ __device__ constexpr int row_stride(int nb) { return (3 * nb + 3) / 4 * 4; }

// lnlike as loglike returns it (brutus/fitting.py:806-815, brutus/utils.py:130-176) plus the external
// label priors `ext` (:1995-2009), and lnprob = lnlike + lnpost's rough parallax prior (:976-982,
// brutus/pdf.py:178-222) with the -1e300 clean-up (:983-985).
template <typename T>
__device__ __forceinline__ void lnl_lnprob(T chi2, T sden, T scale, bool surv, const T* __restrict__ srow,
                                           int dim_prior, T ext, T& lnl, T& lp) {
    if (dim_prior) {
        lnl = chi2 <= T(0) ? Num<T>::neg_inf()
                           : srow[SR_SC + SC_LNORM] + srow[SR_SC + SC_KHM1] * Num<T>::log(chi2) - T(0.5) * chi2;
    } else {
        lnl = T(-0.5) * chi2 + (surv ? srow[SR_SC + SC_GCONST] : T(0));
    }
    lnl += ext;
    lp = lnl;
    if (srow[SR_SC + SC_SPAPPLY] != T(0)) {
        // approximate reciprocals (float: MUFU.RCP, 1-2 ulp): this term only enters lnprob, i.e. the selection
        // threshold and max_lnprob, never the reported lnl / chi2
        T svar = srow[SR_SC + SC_SVAR] + Num<T>::div_fast(T(1), tabs(sden));
        T d = scale - srow[SR_SC + SC_SMEAN];
        lp = lnl + T(-0.5) * (Num<T>::div_fast(d * d, svar) + Num<T>::log(T(6.283185307179586) * svar));
    }
    if (!Num<T>::finite(lp)) lp = Num<T>::kNegBig;
}

// sum over the active label constraints of -0.5 ((label - mean)^2 / std^2 + ln 2 pi std^2)
// ext_row: [nlabel][3] = (mean, 1/std^2, ln(2 pi std^2)); ivar = 0 -> inactive (brutus/fitting.py:1999)
template <typename T>
__device__ __forceinline__ T ext_prior(const T* __restrict__ labels, const T* __restrict__ ext_row, int nlabel,
                                       int64_t npad, int64_t i) {
    T acc = T(0);
    for (int l = 0; l < nlabel; l++) {
        const T* x = ext_row + l * 3;
        if (x[1] > T(0)) {
            T d = labels[(int64_t)l * npad + i] - x[0];
            acc += T(-0.5) * (d * d * x[1] + x[2]);
        }
    }
    return acc;
}

// ---- kernel parameter blocks ----------------------------------------------------------------------
template <typename T> struct SweepParams {
    const float* grid;        // [3][NB][npad]
    int64_t npad, nmodel;
    const T* stars;           // [batch][kStarStride]
    const int* star_int;      // [batch][SI_COUNT]
    const int* list;          // star slots processed by this launch
    int nlist;
    DevOpts<T> o;
    typename Enc<T>::U* red;  // [batch][kNumRed]
    uint32_t* cand;           // [batch][nwords]
    int64_t nwords;
    const T* labels;          // [nlabel][npad]
    const T* ext;             // [batch][nlabel][3]
    int nlabel;
};

template <typename T> struct PoolArrays {
    int *model, *star, *flag;                // [cap]
    T *av, *rv, *chi2, *scale, *sden, *lnl, *lnprob;
};

// compact working set of the flux loops: one entry per survivor of the cull, any order
template <typename T> struct SurvArrays {
    int *q, *model, *star;                   // q = index of the survivor's candidate record
    T *av, *rv, *eta, *lold, *chi2, *scale, *sden;
};

// re-fit of the candidates (exact mag fit of the (star, model) pairs flagged by the sweep)
template <typename T> struct RefitParams {
    const float* rows;        // [npad][row_stride]
    const T* stars;
    const int* star_int;
    DevOpts<T> o;
    PoolArrays<T> pool;
    int64_t ncand;
    const typename Enc<T>::U* red;
    SurvArrays<T> sv;         // survivors are appended here (CTA-granular, any order)
    int* nsv;                 // [1] total survivors appended
    int* nsurv;               // [batch] per-star survivor count
};

template <typename T> struct FluxParams {
    const float* rows;
    const T* stars;
    const int* star_int;
    DevOpts<T> o;
    SurvArrays<T> sv;
    int64_t nsv;
    int nit;                 // flux iterations executed by this launch (2 first, then 1)
    typename Enc<T>::U* red;
};

// O = element type of the outputs: double for the full-length B1 arrays (the reference returns
// float64), T for the compacted B2 records (no point shipping more bits than were computed).
template <typename T, typename O> struct RecordParams {
    const float* rows;
    const T* stars;
    DevOpts<T> o;
    PoolArrays<T> pool;
    // mode A (compacted records): nrec selected pool entries sel_q[0..nrec)
    // mode B (full-length, one star whose candidates are all models): sel_q == nullptr, nrec = nmodel
    const int* sel_q;
    int64_t nrec;
    // outputs.  Mode A: rows of a [11][ld] matrix: lnl, scale, av, chi2, rv, icov(ss,sa,sr,aa,ar,rr);
    // only the first `nrows` are produced (3, 5 or 11).  Mode B: separate arrays, icov 9 per model.
    O *o_lnl, *o_chi2, *o_scale, *o_av, *o_rv, *o_icov;
    int64_t ld;
    int nrows;
    int* o_idx;   // mode A: model index of each record
    int* o_star;  // mode A, optional: star slot of each record (device posterior)
};

// Kernel launchers instantiated once per band count (inst.cu, -DBF_NB=n).
// iteration-count probe (k_kprobe)
constexpr int kProbeIter = 4;
template <typename T> struct ProbeParams {
    const float* grid;
    int64_t npad, nmodel;
    const T* stars;
    int nstar;                    // slots 0..nstar-1
    int tile_stride;
    DevOpts<T> o;
    typename Enc<T>::U* out;      // [nstar][2 * kProbeIter]: (L_k, B_k)
};

template <typename T> struct KTable {
    void (*kprobe)(const ProbeParams<T>&, cudaStream_t);
    void (*magfit)(const SweepParams<T>&, cudaStream_t);
    void (*refit)(const RefitParams<T>&, cudaStream_t);
    void (*flux)(const FluxParams<T>&, cudaStream_t);
    void (*records)(const RecordParams<T, T>&, cudaStream_t);          // compacted records (B2)
    void (*records_full)(const RecordParams<T, double>&, cudaStream_t); // full-length float64 (B1)
};

}  // namespace bf
