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
//                                        that can survive the cull or pass the selection threshold); the
//                                        bits of candidates that fail the selection are cleared later, and
//                                        the rank of a bit gives a record its (star, model)-ordered position
//   red    U [batch][kNumRed]           per-star max-reductions, order-preserving unsigned encoding
//   pool   candidate records (SoA)      one COMPLETE record per candidate (star, model), appended by the
//                                        sweep itself in groups of <= 32 (one warp flush), in no particular
//                                        order: fitted values, flux-loop state and the precision matrix
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
// star rows staged in shared memory use a padded stride: rows stay 16-byte aligned (the sweep reads them
// with broadcast LDS.128) and lanes that read rows of DIFFERENT stars (the dense phase) spread over the banks
constexpr int kStarSmem = kStarStride + 4;

// per-star ints
enum StarInt { SI_NDIM = 0, SI_KSPEC, SI_ACTIVE, SI_NFLUX, SI_EPOCH, SI_NSURV, SI_COUNT = 8 };

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

// candidate record tag (PoolArrays::sflag): star slot | sweep epoch of the star << 16 | flags << 24.
// A star that is swept again (other iteration count) bumps its epoch; records of older epochs are stale.
constexpr int kFlagFluxed = 1;   // the sweep ran the first flux iterations on it (it was a likely survivor)
constexpr int kFlagSurv = 2;     // survived the cull (brutus/fitting.py:758-759)
constexpr int kFlagSel = 4;      // passed lnpost's first selection (brutus/fitting.py:988-991)
__host__ __device__ constexpr int tag_slot(int t) { return t & 0xffff; }
__host__ __device__ constexpr int tag_epoch(int t) { return (t >> 16) & 0xff; }
__host__ __device__ constexpr int tag_flags(int t) { return (t >> 24) & 0xff; }
__host__ __device__ constexpr int make_tag(int slot, int epoch, int flags) { return slot | (epoch & 0xff) << 16 | flags << 24; }

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


// ---- per-star aggregation over candidate records --------------------------------------------------------
// The pool is appended by warp flushes of the sweep: consecutive records belong to the ~32 stars of one
// star chunk, in short runs (a flush holds a handful of stars).  Per-record -- or even per-run -- global
// atomics on the per-star maxima would serialise on a few dozen L2 addresses (measured: 7 ms per 1 000
// stars for the cull pass alone).  Instead every CTA walks a long CONTIGUOUS range of records and combines
// in three levels: lanes of a run with one redux (float), runs into a small per-CTA table in shared memory,
// and the table into global memory once at the end of the range.
constexpr int kAggSlots = 64;
constexpr int kPassCtas = 148 * 8;   // CTAs of a pass over the pool (each takes a contiguous range of records)

// lanes [start, end) holding the same star slot as this lane (records of a star are contiguous in a flush)
__device__ __forceinline__ unsigned run_mask(int slot, int lane, int& start) {
    const int prev = __shfl_up_sync(0xffffffffu, slot, 1);
    const unsigned starts = __ballot_sync(0xffffffffu, lane == 0 || prev != slot);
    start = 31 - __clz(starts & (0xffffffffu >> (31 - lane)));
    const unsigned above = lane == 31 ? 0u : (starts & (0xffffffffu << (lane + 1)));
    const unsigned upto = above ? ((1u << (__ffs(above) - 1)) - 1u) : 0xffffffffu;
    return upto & ~((1u << start) - 1u);
}

template <typename T, int NR> struct StarAgg {
    using U = typename Enc<T>::U;
    int tag[kAggSlots];
    int cnt[kAggSlots];
    U v[NR][kAggSlots];
    __device__ __forceinline__ void init() {
        for (int t = threadIdx.x; t < kAggSlots; t += blockDim.x) {
            tag[t] = -1; cnt[t] = 0;
#pragma unroll
            for (int k = 0; k < NR; k++) v[k][t] = U(0);   // 0 encodes below every value
        }
    }
    // table entry of `slot`, or -1 when its hash bucket belongs to another star (then: global atomics).
    // Plain reads first: almost every call finds its star already in the table and its maximum unchanged, and
    // shared-memory atomics are what bounds these passes otherwise.
    __device__ __forceinline__ int claim(int slot) {
        const int h = slot & (kAggSlots - 1);
        const int cur = *(volatile int*)&tag[h];
        if (cur == slot) return h;
        if (cur != -1) return -1;
        const int old = atomicCAS(&tag[h], -1, slot);
        return (old == -1 || old == slot) ? h : -1;
    }
    // call with every lane of the warp; `which[k]` are the RED_* indices of the NR maxima
    __device__ __forceinline__ void add(U* red, const int (&which)[NR], int* cnt_glob, int cnt_stride, int slot,
                                        bool have, const T (&val)[NR], bool count_flag) {
        const int lane = threadIdx.x & 31;
        int start;
        const unsigned mask = run_mask(slot, lane, start);
        U m[NR];
#pragma unroll
        for (int k = 0; k < NR; k++) m[k] = (have && val[k] == val[k]) ? Enc<T>::enc(val[k]) : U(0);
        warp_combine(m, mask);
        const int nc = __popc(__ballot_sync(0xffffffffu, count_flag) & mask);
        if (lane != start || slot < 0) return;
        const int h = claim(slot);
#pragma unroll
        for (int k = 0; k < NR; k++) {
            if (m[k] == U(0)) continue;
            if (h >= 0) { if (m[k] > *(volatile U*)&v[k][h]) atomicMax(&v[k][h], m[k]); }
            else atomicMax(&red[(int64_t)slot * kNumRed + which[k]], m[k]);
        }
        if (cnt_glob && nc > 0) {
            if (h >= 0) atomicAdd(&cnt[h], nc);
            else atomicAdd(&cnt_glob[(int64_t)slot * cnt_stride], nc);
        }
    }
    __device__ __forceinline__ void flush(U* red, const int (&which)[NR], int* cnt_glob, int cnt_stride) {
        __syncthreads();
        for (int t = threadIdx.x; t < kAggSlots; t += blockDim.x) {
            const int slot = tag[t];
            if (slot < 0) continue;
#pragma unroll
            for (int k = 0; k < NR; k++)
                if (v[k][t] != U(0)) atomicMax(&red[(int64_t)slot * kNumRed + which[k]], v[k][t]);
            if (cnt_glob && cnt[t] > 0) atomicAdd(&cnt_glob[(int64_t)slot * cnt_stride], cnt[t]);
        }
    }
    // float: one redux per maximum; double (verification path): shuffle over the run
    __device__ __forceinline__ static void warp_combine(unsigned int (&m)[NR], unsigned mask) {
#pragma unroll
        for (int k = 0; k < NR; k++) m[k] = __reduce_max_sync(mask, m[k]);
    }
    __device__ __forceinline__ static void warp_combine(unsigned long long (&m)[NR], unsigned mask) {
#pragma unroll
        for (int k = 0; k < NR; k++) {
            const unsigned long long x0 = m[k];
            unsigned long long x = x0;
            for (int j = 0; j < 32; j++) {
                const unsigned long long y = __shfl_sync(0xffffffffu, x0, j);
                if ((mask >> j & 1u) && y > x) x = y;
            }
            m[k] = x;
        }
    }
};

// The passes are streams of small dependent loads (tag -> star -> fields); a thread works on kPassU records at
// a time so that their loads overlap (one record per thread left the passes at a quarter of the HBM bandwidth).
constexpr int kPassU = 4;
constexpr int kPassStep = kTile * kPassU;

// the contiguous range of records [lo, hi) of this CTA, in whole kPassStep-record steps
__device__ __forceinline__ void pass_range(int64_t n, int64_t& lo, int64_t& hi) {
    int64_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + kPassStep - 1) / kPassStep * kPassStep;
    lo = (int64_t)blockIdx.x * per;
    hi = lo + per < n ? lo + per : n;
    if (lo > hi) lo = hi;
}

// Same, for kernels whose records ARE sorted by star (the posterior works on the ordered selection), so most
// CTAs see one or two stars: values for the star of the CTA's first record are combined in shared memory and
// published by ONE global atomic per CTA.  All threads of the CTA must call these (they contain __syncthreads()).
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
// Candidate pool (SoA, one entry per (star, model) the sweep flagged, unordered).
template <typename T> struct PoolArrays {
    int *model;       // model index
    int *sflag;       // make_tag(star slot, epoch, flags)
    T *av, *rv, *chi2, *scale, *sden;   // current fit (flux-phase values once kFlagFluxed)
    T *lp;            // cull statistic lnl_p at the magnitude-fit (Av, Rv)      (brutus/fitting.py:747-756)
    T *eta, *lold, *lprev;              // flux-loop state: stepsize, lnl_new of the last two iterations (:778-803)
    T *isa, *isr, *iaa, *iar, *irr;     // precision matrix icov_sar at the current fit; ss = sden (:563-574)
    T *lnl, *lnprob;  // written by k_final; until then, for fluxed records: the magnitude-fit (Av, Rv)
};
constexpr int kPoolInts = 2, kPoolReals = 16;

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
    PoolArrays<T> pool;
    int64_t pool_cap;
    unsigned long long* pool_count;   // [1] records appended so far (keeps counting past pool_cap)
    int nit_first;            // flux iterations the sweep runs on likely survivors: min(2, max_iter)
    const T* av_init;         // [npad] per-model start of the magnitude fit (brutus/fitting.py:700-703), or both null:
    const T* rv_init;         // the prior means (every caller but bf_loglike_full after bf_set_init)
    int tile_mode;            // which model tiles the launch covers: 0 all, 1 the multiples of tile_S, 2 the others
    int tile_S;               // >= 2 when tile_mode != 0
    int maxima_only;          // 1: only the per-star maxima are produced (no candidates, no records)
};

// passes over the n records of the pool that need the model's coefficients again (both rare)
template <typename T> struct RecParams {
    const float* rows;        // [npad][row_stride]
    const T* stars;
    const int* star_int;
    DevOpts<T> o;
    PoolArrays<T> pool;
    int64_t n;                // records of the pool (k_flux_more) / entries of `list` (k_fixup)
    const int* list;          // k_fixup: pool indices to redo; k_flux_more: the survivors of the active stars
    const int* nlist;         // k_flux_more: length of the list (device)
    typename Enc<T>::U* red;
    const T* av_init;         // as in SweepParams
    const T* rv_init;
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
    const T* av_init;             // as in SweepParams
    const T* rv_init;
};

template <typename T> struct KTable {
    void (*kprobe)(const ProbeParams<T>&, cudaStream_t);
    int (*sweep)(const SweepParams<T>&, cudaStream_t);      // returns a cudaError_t (dynamic shared memory opt-in)
    void (*fixup)(const RecParams<T>&, cudaStream_t);
    void (*flux_more)(const RecParams<T>&, cudaStream_t);
};

}  // namespace bf
