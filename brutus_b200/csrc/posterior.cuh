// posterior.cuh -- the part of the reference's `lnpost` that follows the first selection, the evidence
// and the posterior resampling of `BruteForce._fit`, on the device:
//
//   k_post_mle    priors at the MLE: lnlike + lnprior[model] + Galactic prior(1/sqrt(scale))   brutus/fitting.py:999-1010
//   k_post_count / k_post_write   second threshold lnp > max + ln(wt_thresh), ordered compaction  :1012-1026
//   k_post_mc     covariance = inverse of icov_sar, regularised until positive definite           :1038-1065
//                 Nmc draws of (s, Av, Rv), Galactic + parallax prior at each, in-bounds mask,
//                 lnp += logsumexp - ln(Neff)                                                      :1068-1101
//   k_post_cdf    evidence logsumexp(lnp) and the cumulative weights of the selected models        :2032-2040
//   k_post_draw   Ndraws models by inverse CDF, then one of the model's Nmc (dist, Av, Rv) draws   :2040-2053
//
// The reference draws its normals / uniforms from the caller's numpy RandomState; here they come from a
// counter-based generator (Philox4x32-10) keyed by (seed, star, model, draw), so results do not depend on
// batching or on the number of GPUs.  Parity with the reference is therefore distributional -- except in
// the test mode where the host supplies the normals and uniforms (bf_post_options.z_override /
// u_override): then every number is comparable with the NumPy restatement draw for draw.
#pragma once
#include "common.cuh"

namespace bf {

// ---- Philox4x32 (Salmon et al. 2011) ------------------------------------------------------------------
// R = 10 rounds is the standard generator (resampling uniforms); R = 7, the fewest rounds that pass
// BigCrush in the paper, feeds the Monte Carlo normals, whose loop is bound by instruction issue.
#ifndef BF_MC_POLY_ANGLES
#define BF_MC_POLY_ANGLES 2
#endif
constexpr int kMcPolyAngles = BF_MC_POLY_ANGLES;   // Box-Muller angles (of 3 per block) evaluated by polynomial

template <int R>
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in (0, 1) from 32 random bits: the top 23 bits become the mantissa of a float in [1, 2) (no
// int->float conversion, which would occupy the transcendental pipe the Monte Carlo loop is short of)
__device__ __forceinline__ float u01(uint32_t x) { return __uint_as_float(0x3f800000u | (x >> 9)) - 0.99999994f; }
// uniform in [0, 1) with 53 bits
__device__ __forceinline__ double u01d(uint32_t a, uint32_t b) {
    return (double)((((uint64_t)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}

// ---- math in T -----------------------------------------------------------------------------------------
__device__ __forceinline__ float pexp(float x) { return __expf(x); }
__device__ __forceinline__ double pexp(double x) { return exp(x); }
__device__ __forceinline__ float plog(float x) { return __logf(x); }
__device__ __forceinline__ double plog(double x) { return log(x); }
// float: the approximate MUFU forms (2 ulp), no IEEE slow paths; the Monte Carlo noise is ~1e-1 relative
__device__ __forceinline__ float psqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double psqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float prsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double prsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float prcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double prcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float pexp10(float x) { return exp2f(x * 3.3219281f); }
__device__ __forceinline__ double pexp10(double x) { return exp10(x); }

__device__ __forceinline__ float pex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double pex2(double x) { return exp2(x); }
__device__ __forceinline__ float plg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double plg2(double x) { return log2(x); }
// lane-wise transcendental of a pair (MUFU is scalar: the two halves of the register pair are fed one by one)
#define BF_PAIR_FN(name, fn) \
    template <typename T> __device__ __forceinline__ P2<T> name(P2<T> v) { return mk2(fn(lo2(v)), fn(hi2(v))); }
BF_PAIR_FN(sqrt2, psqrt)
BF_PAIR_FN(rsqrt2, prsqrt)
BF_PAIR_FN(rcp2, prcp)
BF_PAIR_FN(ex22, pex2)
BF_PAIR_FN(lg22, plg2)
BF_PAIR_FN(abs2, tabs)
#undef BF_PAIR_FN
constexpr double kLog2e = 1.4426950408889634, kLn2 = 0.6931471805599453;

// sine and cosine of an angle uniform on the circle, from 16 random bits, WITHOUT the transcendental unit: the
// low 14 bits place theta in [-pi/4, pi/4), the top 2 bits rotate by a multiple of pi/2 (swap + signs).  Both
// series are evaluated together with packed FMAs (|error| < 4e-6).  The Monte Carlo loop is bound by the
// transcendental unit, not by the FMA pipe.
__device__ __forceinline__ void sincos_bits(uint32_t h, float& sn, float& cs) {
    // [1, 2) from 14 mantissa bits, then theta = (pi/2) (f - 1.5)
    const float f = __uint_as_float(0x3f800000u | ((h & 0x3fffu) << 9));
    const float th = fmaf(f, 1.5707963267948966f, -2.356194490192345f);
    const float t = th * th;
    // sin(th) / th = 1 - t/6 + t^2/120 - t^3/5040 ;  cos(th) = 1 - t/2 + t^2/24 - t^3/720
    const P2<float> tt = bc2(t);
    P2<float> pq = fma2(tt, mk2(-1.98412698e-4f, -1.38888889e-3f), mk2(8.33333333e-3f, 4.16666667e-2f));
    pq = fma2(pq, tt, mk2(-1.66666667e-1f, -0.5f));
    pq = fma2(pq, tt, bc2(1.f));
    const float s0 = lo2(pq) * th, c0 = hi2(pq);
    const uint32_t q = h >> 14;                                   // 0..3: rotate by q pi/2
    const bool sw = (q & 1u) != 0u;
    const float s1 = sw ? c0 : s0, c1 = sw ? s0 : c0;             // q odd: (sin, cos) <- (cos, -sin)
    sn = __uint_as_float(__float_as_uint(s1) ^ ((q >> 1) << 31));
    cs = __uint_as_float(__float_as_uint(c1) ^ (((q ^ (q >> 1)) & 1u) << 31));
}

// six standard normals for draws (2k, 2k + 1) of (star, model): ONE Philox block.  Each 32-bit word feeds one
// Box-Muller pair: its high half is the radius uniform, (h + 1/2) / 2^16 (the normals reach 4.7 sigma), its low
// half the angle.  The fourth word is not used.  The generator, not the prior, was the largest single item of
// the Monte Carlo loop when every draw had a block of its own.  Two of the three angles go through the packed
// polynomial, one through MUFU.SIN/COS: that balances the transcendental unit against instruction issue.
template <typename T>
__device__ __forceinline__ void normals6(uint64_t seed, uint64_t star, uint32_t model, uint32_t k, T (&za)[3], T (&zb)[3]) {
    uint32_t r[4];
    philox4x32<7>(model, k, (uint32_t)star, (uint32_t)(star >> 32) ^ 0x4D435A31u, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    float n[6];
#pragma unroll
    for (int w = 0; w < 3; w++) {
        // [1, 2) from the mantissa bits, minus (1 - 2^-17): exact, no int->float conversion
        const float u = __uint_as_float(0x3f800000u | ((r[w] >> 16) << 7)) - 0.99999237060546875f;
        const float rad = psqrt(-1.3862943611198906f * plg2(u));          // sqrt(-2 ln u)
        float sn, cs;
        if (w < kMcPolyAngles) {
            sincos_bits(r[w] & 0xffffu, sn, cs);
        } else {
            const float v = __uint_as_float(0x3f800000u | ((r[w] & 0xffffu) << 7)) - 1.f;
            __sincosf(6.2831853f * v, &sn, &cs);
        }
        n[2 * w] = rad * cs; n[2 * w + 1] = rad * sn;
    }
    za[0] = (T)n[0]; za[1] = (T)n[1]; za[2] = (T)n[2];
    zb[0] = (T)n[3]; zb[1] = (T)n[4]; zb[2] = (T)n[5];
}

// ---- the Galactic prior (brutus/pdf.py:476-749) --------------------------------------------------------
// constants prepared on the host (api.cu, make_gal) from bf_gal_params
template <typename T> struct GalDev {
    int use;                                  // 0: no distance prior
    int has_feh, has_age;
    int same_rs;                              // Rs_thin == Rs_thick (the defaults): one square root serves both disks
    T Rs_thin2, Rs_thick2, Rs_halo2;          // smoothing radii squared          (:301, :363)
    T R_solar, aZ_solar;
    T iR_thin, iZ_thin, iR_thick, iZ_thick;   // inverse scale lengths            (:303-304)
    T ln_f_thick, ln_f_halo;                  // (:647, :654)
    T rq2, irq, q_inf, dq, eta, ln_Reff_solar;  // halo oblateness + normalisation (:358-374)
    T feh_mu[3], feh_isig2[3], feh_lnorm[3];  // thin, thick, halo                (:380-408)
    T age_mu[3], age_isig[3], age_lnden[3];   // truncated normals                (:455-470, brutus/utils.py:232-284)
    T min_age, max_age;
    // the same constants in the base-2 form the Monte Carlo loop uses (gal_prior2): multiplied out on the host, in
    // float64, instead of once per pair of draws
    T l2_iR_thin, l2_iZ_thin, l2_iR_thick, l2_iZ_thick;   // -log2(e) x inverse scale lengths
    T l2_ln_f_thick, l2_irq, l2_halo, nh_eta;             // log2(e) ln f_thick; -log2(e) / r_q; log2(e) (eta ln Reff_sun + ln f_halo); -eta / 2
};
// heliocentric -> galactocentric, linear in the distance d: x = d ax + x0, y = d ay, z = d az + z0
template <typename T> struct GalStar { T ax, ay, az, x0, z0; };
// exp(log-prior) of the model's labels in each component; 1 when the label is absent
template <typename T> struct ModelW { T f[3], g[3]; };

template <typename T>
__device__ __forceinline__ void model_weights(const GalDev<T>& G, const T* __restrict__ feh,
                                              const T* __restrict__ loga, int64_t i, ModelW<T>& w) {
#pragma unroll
    for (int x = 0; x < 3; x++) { w.f[x] = T(1); w.g[x] = T(1); }
    if (!G.use) return;
    if (G.has_feh) {
        const T v = feh[i];
#pragma unroll
        for (int x = 0; x < 3; x++) {
            const T d = G.feh_mu[x] - v;
            w.f[x] = pexp(T(-0.5) * (d * d * G.feh_isig2[x] + G.feh_lnorm[x]));
        }
    }
    if (G.has_age) {
        const T age = pexp10(loga[i] - T(9));                                   // Gyr (:694)
        const bool out = age < G.min_age || age > G.max_age;                    // brutus/utils.py:275-277
#pragma unroll
        for (int x = 0; x < 3; x++) {
            const T xi = (age - G.age_mu[x]) * G.age_isig[x];
            w.g[x] = out ? T(0) : pexp(T(-0.9189385332046727) - T(0.5) * xi * xi - G.age_lnden[x]);
        }
    }
}

// ln prior = 2 ln d + ln sum_X n_X + ln(sum f_X n_X / sum n_X) + ln(sum g_X n_X / sum n_X)
// with n_X the thin / thick / halo number densities (:622-745).  The three logsumexp of the reference share
// the same exponentials; they are taken relative to the halo term, which dominates wherever the disks
// underflow and is never more than ~e^10 below them, so no running maximum is needed:
//   ln prior = ln n_halo + ln( s1 s2 / (s0 s) ),   s = 1/d^2 (the scale),  s_k = sums of n_X / n_halo.
// `s` is the scale factor (parallax^2), `rs` = 1/sqrt(s) = the distance in kpc.
template <typename T>
__device__ __forceinline__ T gal_lnprior(const GalDev<T>& G, const GalStar<T>& gs, const ModelW<T>& w, T s, T d) {
    if (!G.use) return T(0);
    const T x = fma(d, gs.ax, gs.x0), y = d * gs.ay, z = fma(d, gs.az, gs.z0);
    const T R2 = x * x + y * y, aZ = tabs(z);
    const T Rthin = psqrt(R2 + G.Rs_thin2);
    const T Rthick = G.same_rs ? Rthin : psqrt(R2 + G.Rs_thick2);
    const T lt = -((Rthin - G.R_solar) * G.iR_thin + (aZ - G.aZ_solar) * G.iZ_thin);
    const T lk = -((Rthick - G.R_solar) * G.iR_thick + (aZ - G.aZ_solar) * G.iZ_thick) + G.ln_f_thick;
    const T rp = psqrt(R2 + z * z + G.rq2);
    const T q = G.q_inf - G.dq * pexp(T(1) - rp * G.irq);
    const T zq = z * prcp(q);
    const T lh = -G.eta * (T(0.5) * plog(R2 + zq * zq + G.Rs_halo2) - G.ln_Reff_solar) + G.ln_f_halo;
    const T nt = pexp(lt - lh), nk = pexp(lk - lh);
    const T s0 = nt + nk + T(1);
    const T s1 = w.f[0] * nt + w.f[1] * nk + w.f[2];
    const T s2 = w.g[0] * nt + w.g[1] * nk + w.g[2];
    return lh + plog(s1 * s2 * prcp(s0 * s));
}

// ---- parameters shared by the posterior kernels ------------------------------------------------------------
template <typename T> struct PostParams {
    // records of the first selection, read in place from the candidate pool: ord[t] = pool index of the t-th
    // selected record in (star, model) order (k_ord); no copy of the records is made for the posterior
    const int* ord;
    PoolArrays<T> pool;
    int* rstar;              // [n1] star slot per record (written by k_post_mle)
    int64_t n1;
    // static per-model priors / labels ([npad], any may be null)
    const T* lnprior;
    const T* feh;
    const T* loga;
    GalDev<T> G;
    const GalStar<T>* gstar; // [batch]
    const T* stars;          // star rows (parallax, 1/parallax_err^2)
    typename Enc<T>::U* red; // [batch][kNumRed]
    T ln_wt;
    T avmin, avmax, rvmin, rvmax;
    // second selection
    T* lnp1;                 // [n1]
    T* lnb1;                 // [n1] lnlike + lnprior: the ranking key of the memory clip (:1024, :1029-1036)
    const T* clip_thr;       // [batch] smallest key kept by the clip (-inf: star not clipped), or null
    T* keys;                 // [n2] keys of the second selection, gathered for the segmented sort
    int* blk;                // per 256-record block counts -> offsets
    int* nsel2;              // [batch]
    int* sel2;               // [n2] -> record index t
    int64_t n2;
    const int64_t* off2;     // [batch+1] star -> first entry of sel2 (slots g0..g1 of the current group)
    T* lnp2;                 // [n2]
    double* cdf;             // [n2]
    double* tot;             // [batch] sum of the weights
    // Monte Carlo
    int nmc, ndraws;
    uint64_t seed;
    int64_t star_base;       // catalogue index of slot 0 (keys the generator)
    const double* zov;       // [nmodel][3][nmc] or null
    const double* uov;       // [batch][2][ndraws] (slot-indexed) or null
    // outputs, [batch][ndraws] (slot-indexed)
    int g0;                  // first slot of the group
    int* o_idx;
    double *o_scale, *o_av, *o_rv, *o_cov, *o_lnprob, *o_dist, *o_red, *o_dred, *o_logwt;
    double *o_levid, *o_chi2min;   // [batch]
};

// (:999-1010)
template <typename T> __global__ void __launch_bounds__(kTile) k_post_mle(const PostParams<T> p) {
    const int64_t t = (int64_t)blockIdx.x * kTile + threadIdx.x;
    const bool in = t < p.n1;
    int slot = -1;
    T lp = Num<T>::neg_inf();
    if (in) {
        const int64_t q = p.ord[t];
        slot = tag_slot(p.pool.sflag[q]);
        p.rstar[t] = slot;
        const int64_t i = p.pool.model[q];
        ModelW<T> w;
        model_weights<T>(p.G, p.feh, p.loga, i, w);
        const T scale = p.pool.scale[q];
        const T base = p.pool.lnl[q] + (p.lnprior ? p.lnprior[i] : T(0));
        lp = base + gal_lnprior<T>(p.G, p.gstar[slot], w, scale, prsqrt(scale));
        p.lnp1[t] = lp;
        p.lnb1[t] = base;
    }
    cta_star_max<T>(p.red, RED_P1, slot, in, lp);
}

template <typename T> __device__ __forceinline__ bool post_flag(const PostParams<T>& p, int64_t t, int& slot) {
    slot = -1;
    if (t >= p.n1) return false;
    slot = p.rstar[t];
    if (p.clip_thr && !(p.lnb1[t] >= p.clip_thr[slot])) return false;                       // (:1029-1036)
    return p.lnp1[t] > Enc<T>::dec(p.red[(int64_t)slot * kNumRed + RED_P1]) + p.ln_wt;   // (:1013-1016)
}

template <typename T> __global__ void __launch_bounds__(kTile) k_post_count(const PostParams<T> p) {
    const int64_t t = (int64_t)blockIdx.x * kTile + threadIdx.x;
    int slot;
    const bool f = post_flag(p, t, slot);
    const int n = __syncthreads_count(f);
    if (threadIdx.x == 0) p.blk[blockIdx.x] = n;
    cta_star_count(p.nsel2, slot, f);
}

template <typename T> __global__ void __launch_bounds__(kTile) k_post_write(const PostParams<T> p) {
    __shared__ int s_w[kTile / 32];
    const int64_t t = (int64_t)blockIdx.x * kTile + threadIdx.x;
    int slot;
    const bool f = post_flag(p, t, slot);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_w[w] = __popc(bal);
    __syncthreads();
    if (f) {
        int pre = __popc(bal & ((1u << lane) - 1u));
        for (int k = 0; k < w; k++) pre += s_w[k];
        p.sel2[p.blk[blockIdx.x] + pre] = (int)t;
    }
}

// memory clip (:1029-1036): keys of the second selection, and the threshold of every over-full star from its
// keys sorted in descending order (segment k of the sort = star seg_slot[k], starting at seg_begin[k])
template <typename T> __global__ void k_post_keys(const PostParams<T> p) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < p.n2) p.keys[u] = p.lnb1[p.sel2[u]];
}
template <typename T>
__global__ void k_post_thr(const T* __restrict__ sorted, const int* __restrict__ seg_begin, const int* __restrict__ seg_slot,
                           int nseg, int64_t nsel_max, T* __restrict__ thr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nseg) thr[seg_slot[k]] = sorted[(int64_t)seg_begin[k] + nsel_max - 1];
}
template <typename T> __global__ void k_fill(T* p, int n, T v) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) p[k] = v;
}

// Covariance of (s, Av, Rv) in RELATIVE scale units (s' = s / scale, so that the matrix is O(1) whatever
// the star's flux level): inverse of icov_sar by the adjugate (brutus/utils.py:71-114), regularised until
// positive definite exactly as brutus/fitting.py:1042-1065 (the tests there are invariant under the
// rescaling), then its Cholesky factor (brutus/utils.py:893-894).  float64 throughout: 3x3 algebra is
// negligible next to the Nmc prior evaluations.
struct Cov3 {
    double c[6];   // cov' (00, 01, 02, 11, 12, 22), relative units
    double L[6];   // Cholesky factor (00, 10, 11, 20, 21, 22)
};

__device__ __forceinline__ bool inv_sym3(const double (&a)[6], double (&c)[6]) {
    const double a00 = a[0], a01 = a[1], a02 = a[2], a11 = a[3], a12 = a[4], a22 = a[5];
    const double m00 = a11 * a22 - a12 * a12, m01 = a02 * a12 - a01 * a22, m02 = a01 * a12 - a02 * a11;
    const double m11 = a00 * a22 - a02 * a02, m12 = a01 * a02 - a00 * a12, m22 = a00 * a11 - a01 * a01;
    // determinant as the mean of the three row expansions (brutus/utils.py:100)
    const double det = ((a00 * m00 + a01 * m01 + a02 * m02) + (a01 * m01 + a11 * m11 + a12 * m12) +
                        (a02 * m02 + a12 * m12 + a22 * m22)) * (1.0 / 3.0);
    const double id = 1.0 / det;
    c[0] = m00 * id; c[1] = m01 * id; c[2] = m02 * id; c[3] = m11 * id; c[4] = m12 * id; c[5] = m22 * id;
    // positive definite?  (np.linalg.eigvals(cov) > 0 for a symmetric matrix <=> Sylvester)
    const double d2 = c[0] * c[3] - c[1] * c[1];
    const double d3 = c[0] * (c[3] * c[5] - c[4] * c[4]) - c[1] * (c[1] * c[5] - c[4] * c[2]) + c[2] * (c[1] * c[4] - c[3] * c[2]);
    return c[0] > 0. && d2 > 0. && d3 > 0.;
}

template <typename T>
__device__ __forceinline__ void make_cov(const PoolArrays<T>& pl, int64_t q, double scale, Cov3& cv) {
    double a[6];
    a[0] = (double)pl.sden[q] * scale * scale;
    a[1] = (double)pl.isa[q] * scale;
    a[2] = (double)pl.isr[q] * scale;
    a[3] = (double)pl.iaa[q];
    a[4] = (double)pl.iar[q];
    a[5] = (double)pl.irr[q];
    bool ok = inv_sym3(a, cv.c);
    const double iw2 = 1.0 / (0.02 * 0.02);   // 2 % Gaussian prior (:1044)
    double count = 1.0;
    for (int it = 0; it < 80 && !ok; it++) {
        const bool i1 = cv.c[0] <= 0., i2 = cv.c[3] <= 0., i3 = cv.c[5] <= 0.;
        const bool s1 = i1 || (!i2 && !i3), s2 = i2 || (!i1 && !i3), s3 = i3 || (!i1 && !i2);   // (:1054-1056)
        if (s1) a[0] += count * iw2;        // count / (scale * width)^2 in absolute units
        if (s2) a[3] += count * iw2;
        if (s3) a[5] += count * iw2;
        ok = inv_sym3(a, cv.c);
        count *= 2.0;
    }
    const double* c = cv.c;
    const double l00 = sqrt(c[0] > 0. ? c[0] : 0.);
    const double l10 = l00 > 0. ? c[1] / l00 : 0., l20 = l00 > 0. ? c[2] / l00 : 0.;
    const double t11 = c[3] - l10 * l10;
    const double l11 = sqrt(t11 > 0. ? t11 : 0.);
    const double l21 = l11 > 0. ? (c[4] - l20 * l10) / l11 : 0.;
    const double t22 = c[5] - l20 * l20 - l21 * l21;
    cv.L[0] = l00; cv.L[1] = l10; cv.L[2] = l11; cv.L[3] = l20; cv.L[4] = l21; cv.L[5] = sqrt(t22 > 0. ? t22 : 0.);
}

// one Monte Carlo draw of a selected model and its log-prior (:1070-1095)
template <typename T> struct McCtx {
    T scale, av, rv;
    T L[6];
    ModelW<T> w;
    GalStar<T> gs;
    T par, pivar, lnorm_par;   // parallax prior (brutus/pdf.py:144-175); pivar = 0: none
    T par_c2, par_k2;          // the same in log2 units: lg2 prior += par_c2 (p - par)^2 + par_k2
    uint32_t model;
    uint64_t star;
};

// ---- two Monte Carlo draws at a time ---------------------------------------------------------------------
// k_post_mc is bound by instruction issue (ncu: 184 instructions and 17 MUFU per draw, issue 78 %, XU 71 %), so
// the floating-point work of draws (j, j+1) is done with packed FP32 (FFMA2 / FMUL2 / FADD2: one issue slot
// for both draws); the transcendentals stay scalar.  Same formula as gal_lnprior above, evaluated in base 2 and returned in two
// parts, prior = arg * 2^lg2h: the accumulation over draws needs 2^(lg2 prior - reference) = arg * 2^(lg2h - reference),
// one exponential and no logarithm.
// T = double runs the same code on plain pairs.
// SAME: the production configuration known at compile time -- the Galactic prior in use with Rs_thin == Rs_thick (the
// defaults).  As run-time tests the second pair of square roots is predicated off but still issued, and the branch
// around the prior costs its not-taken initialisations and splits the schedule of the loop.
template <typename T, bool SAME = false>
__device__ __forceinline__ void gal_prior2(const GalDev<T>& G, const GalStar<T>& gs, const ModelW<T>& w, P2<T> s, P2<T> d,
                                           P2<T>& lg2h, P2<T>& arg) {
    const P2<T> x = fma2(d, bc2(gs.ax), bc2(gs.x0)), y = mul2(d, bc2(gs.ay)), z = fma2(d, bc2(gs.az), bc2(gs.z0));
    const P2<T> R2 = fma2(x, x, mul2(y, y));
    const P2<T> dz = sub2(abs2(z), bc2(G.aZ_solar));
    const P2<T> Rthin = sqrt2(add2(R2, bc2(G.Rs_thin2)));
    const P2<T> Rthick = (SAME || G.same_rs) ? Rthin : sqrt2(add2(R2, bc2(G.Rs_thick2)));
    // log2 of the disk densities relative to the solar neighbourhood
    const P2<T> lt = fma2(sub2(Rthin, bc2(G.R_solar)), bc2(G.l2_iR_thin), mul2(dz, bc2(G.l2_iZ_thin)));
    const P2<T> lk = fma2(sub2(Rthick, bc2(G.R_solar)), bc2(G.l2_iR_thick), fma2(dz, bc2(G.l2_iZ_thick), bc2(G.l2_ln_f_thick)));
    const P2<T> rp = sqrt2(add2(fma2(z, z, R2), bc2(G.rq2)));
    const P2<T> q = fma2(ex22(fma2(rp, bc2(G.l2_irq), bc2(T(kLog2e)))), bc2(-G.dq), bc2(G.q_inf));
    const P2<T> zq = mul2(z, rcp2(q));
    // log2 of the halo density: -eta/2 log2(Reff^2) + eta log2(Reff_solar) + log2 f_halo
    const P2<T> lh = fma2(lg22(add2(fma2(zq, zq, R2), bc2(G.Rs_halo2))), bc2(G.nh_eta), bc2(G.l2_halo));
    const P2<T> nt = ex22(sub2(lt, lh)), nk = ex22(sub2(lk, lh));
    const P2<T> s0 = add2(add2(nt, nk), bc2(T(1)));
    const P2<T> s1 = fma2(nt, bc2(w.f[0]), fma2(nk, bc2(w.f[1]), bc2(w.f[2])));
    const P2<T> s2 = fma2(nt, bc2(w.g[0]), fma2(nk, bc2(w.g[1]), bc2(w.g[2])));
    arg = mul2(mul2(s1, s2), rcp2(mul2(s0, s)));
    lg2h = lh;
}

// draws j and j + 1 of a selected model: (s, Av, Rv), in-bounds flags and log-priors (:1070-1095).  The second
// lane is computed even when j + 1 == nmc (the caller ignores it).
template <typename T> struct McPair {
    P2<T> s, a, r;
    P2<T> lg2h, arg; // the prior at the two draws = arg * 2^lg2h (meaningless for a lane that is out of bounds)
    bool inb[2];
    // the log-prior of lane k the way the reference holds it: -1e300 out of bounds (:1093), -inf for NaN (:1095 nan_to_num
    // of exp) -- only the resampling kernel and the exact fallback of the accumulation need this form
    __device__ T lnp(int k) const {
        const T l = (k ? hi2(lg2h) + plg2(hi2(arg)) : lo2(lg2h) + plg2(lo2(arg))) * T(kLn2);
        return !inb[k] ? Num<T>::kNegBig : (l != l ? Num<T>::neg_inf() : l);
    }
};

template <typename T, bool ZOV, bool SAME = false>
__device__ __forceinline__ void mc_pair(const PostParams<T>& p, const McCtx<T>& c, int j, McPair<T>& o) {
    T za[3], zb[3];
    if (ZOV) {
        const int jb = j + 1 < p.nmc ? j + 1 : j;
        const double* zz = p.zov + (size_t)c.model * 3 * p.nmc;
#pragma unroll
        for (int k = 0; k < 3; k++) { za[k] = (T)zz[k * p.nmc + j]; zb[k] = (T)zz[k * p.nmc + jb]; }
    } else {
        normals6<T>(p.seed, c.star, c.model, (uint32_t)(j >> 1), za, zb);
    }
    const P2<T> z0 = mk2(za[0], zb[0]), z1 = mk2(za[1], zb[1]), z2 = mk2(za[2], zb[2]);
    o.s = mul2(bc2(c.scale), fma2(bc2(c.L[0]), z0, bc2(T(1))));
    o.a = fma2(bc2(c.L[2]), z1, fma2(bc2(c.L[1]), z0, bc2(c.av)));
    o.r = fma2(bc2(c.L[5]), z2, fma2(bc2(c.L[4]), z1, fma2(bc2(c.L[3]), z0, bc2(c.rv))));
    const T sv[2] = {lo2(o.s), hi2(o.s)}, av[2] = {lo2(o.a), hi2(o.a)}, rv[2] = {lo2(o.r), hi2(o.r)};
#pragma unroll
    for (int k = 0; k < 2; k++)   // (:1090-1092)
        o.inb[k] = sv[k] >= T(1e-20) && av[k] >= p.avmin && av[k] <= p.avmax && rv[k] >= p.rvmin && rv[k] <= p.rvmax;
    // out-of-bounds lanes may carry a negative scale: give them a harmless one, their result is discarded
    const P2<T> sc = mk2(o.inb[0] ? sv[0] : c.scale, o.inb[1] ? sv[1] : c.scale);
    const P2<T> dist = rsqrt2(sc), par = mul2(sc, dist);
    P2<T> lp = bc2(T(0));
    o.arg = bc2(T(1));
    if (SAME || p.G.use) gal_prior2<T, SAME>(p.G, c.gs, c.w, sc, dist, lp, o.arg);
    if (c.pivar > T(0)) {
        const P2<T> d = sub2(par, bc2(c.par));
        lp = fma2(mul2(d, d), bc2(c.par_c2), add2(lp, bc2(c.par_k2)));
    }
    o.lg2h = lp;
}

template <typename T>
__device__ __forceinline__ void mc_setup(const PostParams<T>& p, int64_t t, int slot, McCtx<T>& c, Cov3& cv) {
    const int64_t q = p.ord[t];
    const int64_t i = p.pool.model[q];
    c.scale = p.pool.scale[q];
    c.av = p.pool.av[q];
    c.rv = p.pool.rv[q];
    make_cov<T>(p.pool, q, (double)c.scale, cv);
#pragma unroll
    for (int k = 0; k < 6; k++) c.L[k] = (T)cv.L[k];
    model_weights<T>(p.G, p.feh, p.loga, i, c.w);
    c.gs = p.gstar[slot];
    const T* __restrict__ srow = p.stars + (int64_t)slot * kStarStride;
    c.par = srow[SR_SC + SC_PAR];
    c.pivar = srow[SR_SC + SC_PIVAR];
    c.lnorm_par = c.pivar > T(0) ? T(1.8378770664093453) - plog(c.pivar) : T(0);
    c.par_c2 = T(-0.5 * kLog2e) * c.pivar;
    c.par_k2 = T(-0.5 * kLog2e) * c.lnorm_par;
    c.model = (uint32_t)i;
    c.star = (uint64_t)(p.star_base + slot);
}

// log-sum-exp accumulator over the draws; -inf terms contribute nothing, kNegBig terms (out of bounds,
// the reference's -1e300) contribute exp(kNegBig - max) = 0 unless every term is kNegBig
template <typename T> struct Lse {
    T m, s;
    __device__ Lse() : m(Num<T>::neg_inf()), s(T(0)) {}
    __device__ void add(T x) {
        if (!(x > Num<T>::neg_inf())) return;
        if (x > m) { s = s * pexp(m - x) + T(1); m = x; }
        else s += pexp(x - m);
    }
    __device__ T value() const { return s > T(0) ? m + plog(s) : Num<T>::neg_inf(); }
};

// (:1038-1106) one thread per model of the second selection.  ZOV: normals supplied by the host (test mode)
template <typename T, bool ZOV, bool SAME> __global__ void __launch_bounds__(kTile, 4) k_post_mc(const PostParams<T> p) {
    const int64_t u = (int64_t)blockIdx.x * kTile + threadIdx.x;
    const bool in = u < p.n2;
    int slot = -1;
    T lnp = Num<T>::neg_inf(), nchi = Num<T>::neg_inf();
    if (in) {
        const int64_t t = p.sel2[u];
        slot = p.rstar[t];
        McCtx<T> c;
        Cov3 cv;
        mc_setup<T>(p, t, slot, c, cv);
        // logsumexp over the draws against a FIXED reference, the prior at the MLE itself: one ex2 and one add per
        // draw, no running maximum.  The prior moves by far less than the 2^+-126 range of float32 over draws a few
        // sigma from the MLE; if the sum nevertheless leaves the range the model is redone the careful way.
        T ref2;
        {
            const P2<T> s0 = bc2(c.scale), d0 = rsqrt2(s0);
            P2<T> l0 = bc2(T(0)), a0 = bc2(T(1));
            if (SAME || p.G.use) gal_prior2<T, SAME>(p.G, c.gs, c.w, s0, d0, l0, a0);
            if (c.pivar > T(0)) {
                const P2<T> d = sub2(mul2(s0, d0), bc2(c.par));
                l0 = fma2(mul2(d, d), bc2(c.par_c2), add2(l0, bc2(c.par_k2)));
            }
            ref2 = lo2(l0) + plg2(lo2(a0));
            if (!Num<T>::finite(ref2)) ref2 = T(0);
        }
        T sum0 = T(0), sum1 = T(0);
        int neff = 0;
        for (int j = 0; j < p.nmc; j += 2) {
            McPair<T> m;
            mc_pair<T, ZOV, SAME>(p, c, j, m);
            const P2<T> e = mul2(m.arg, ex22(sub2(m.lg2h, bc2(ref2))));
            const bool in1 = m.inb[1] && j + 1 < p.nmc;
            // max(NaN, 0) = 0: a NaN prior counts as exp(-inf) like in the reference (:1095)
            sum0 += m.inb[0] ? Num<T>::max(lo2(e), T(0)) : T(0);
            sum1 += in1 ? Num<T>::max(hi2(e), T(0)) : T(0);
            neff += (m.inb[0] ? 1 : 0) + (in1 ? 1 : 0);
        }
        const T sum = sum0 + sum1;
        T lse;
        if (sum > T(0) && Num<T>::finite(sum)) {
            lse = (ref2 + plg2(sum)) * T(kLn2);
        } else {   // every term underflowed (or none is finite): running-maximum accumulation, as the reference's logsumexp
            Lse<T> acc;
            for (int j = 0; j < p.nmc; j += 2) {
                McPair<T> m;
                mc_pair<T, ZOV, SAME>(p, c, j, m);
                acc.add(m.lnp(0));
                if (j + 1 < p.nmc) acc.add(m.lnp(1));
            }
            lse = acc.value();
        }
        const int64_t q = p.ord[t];
        const int64_t i = p.pool.model[q];
        lnp = p.pool.lnl[q] + (p.lnprior ? p.lnprior[i] : T(0));             // lnlike + lnprior (:1024)
        lnp = neff > 0 ? lnp + lse - plog((T)neff) : Num<T>::kNegBig;  // (:1098-1100; Neff = 0 -> +inf -> -1e300)
        if (!Num<T>::finite(lnp) || lnp < Num<T>::kNegBig) lnp = Num<T>::kNegBig;   // (:1103-1105)
        p.lnp2[u] = lnp;
        // chi2 with the parallax term (:2025-2030), for chi2min
        T chi2 = p.pool.chi2[q];
        if (c.pivar > T(0)) { const T d = psqrt(c.scale) - c.par; chi2 += d * d * c.pivar; }
        nchi = -chi2;
    }
    cta_star_max<T>(p.red, RED_P2, slot, in, lnp);
    cta_star_max<T>(p.red, RED_NCHI, slot, in, nchi);
}

// block-wide inclusive scan of doubles (1024 threads); returns the inclusive prefix, total through `total`
__device__ __forceinline__ double block_incscan_1024(double v, double* s_w, double& total) {
    double x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = s_w[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double y = __shfl_up_sync(0xffffffffu, w, o);
            if (threadIdx.x >= o) w += y;
        }
        s_w[threadIdx.x] = w;
    }
    __syncthreads();
    const double incl = x + ((threadIdx.x >> 5) ? s_w[(threadIdx.x >> 5) - 1] : 0.);
    total = s_w[31];
    __syncthreads();
    return incl;
}

// one CTA per star of the group: cumulative weights exp(lnp - max) of its selected models, evidence (:2032-2039)
template <typename T> __global__ void __launch_bounds__(1024) k_post_cdf(const PostParams<T> p) {
    __shared__ double s_w[32];
    const int slot = p.g0 + blockIdx.x;
    const int64_t lo = p.off2[slot], hi = p.off2[slot + 1];
    const T M = Enc<T>::dec(p.red[(int64_t)slot * kNumRed + RED_P2]);
    double carry = 0.;
    for (int64_t b = lo; b < hi; b += 1024) {
        const int64_t u = b + threadIdx.x;
        const double w = u < hi ? exp((double)p.lnp2[u] - (double)M) : 0.;
        double tot;
        const double inc = block_incscan_1024(w, s_w, tot);
        if (u < hi) p.cdf[u] = carry + inc;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        p.tot[slot] = carry;
        const bool bad = !(M > Num<T>::kNegBig);
        p.o_levid[slot] = bad ? -1e300 + log(carry > 0. ? carry : 1.) : (double)M + log(carry);
        const T nchi = Enc<T>::dec(p.red[(int64_t)slot * kNumRed + RED_NCHI]);
        p.o_chi2min[slot] = -(double)nchi;
    }
}

// one CTA per star, one thread per posterior draw (:2040-2061)
template <typename T, bool ZOV> __global__ void __launch_bounds__(kTile) k_post_draw(const PostParams<T> p) {
    const int slot = p.g0 + blockIdx.x;
    const int64_t lo = p.off2[slot], hi = p.off2[slot + 1];
    for (int d = threadIdx.x; d < p.ndraws; d += blockDim.x) {
    const int64_t o = (int64_t)slot * p.ndraws + d;
    if (hi <= lo) { p.o_idx[o] = -99; continue; }
    double u1, u2;
    if (p.uov) {
        u1 = p.uov[((size_t)slot * 2 + 0) * p.ndraws + d];
        u2 = p.uov[((size_t)slot * 2 + 1) * p.ndraws + d];
    } else {
        uint32_t r[4];
        const uint64_t star = (uint64_t)(p.star_base + slot);
        philox4x32<10>((uint32_t)d, 0x44524157u, (uint32_t)star, (uint32_t)(star >> 32) ^ 0x52455331u, (uint32_t)p.seed,
                       (uint32_t)(p.seed >> 32), r);
        u1 = u01d(r[0], r[1]);
        u2 = u01d(r[2], r[3]);
    }
    // numpy's choice: cdf /= cdf[-1]; searchsorted(cdf, u, side='right') = first entry with cdf > u
    const double tot = p.tot[slot];
    int64_t a = lo, b = hi;
    while (a < b) {
        const int64_t mid = (a + b) >> 1;
        if (p.cdf[mid] / tot > u1) b = mid; else a = mid + 1;
    }
    const int64_t u = a < hi ? a : hi - 1;
    const int64_t t = p.sel2[u];
    McCtx<T> c;
    Cov3 cv;
    mc_setup<T>(p, t, slot, c, cv);
    const double sc = (double)c.scale;
    p.o_idx[o] = p.pool.model[p.ord[t]];
    p.o_scale[o] = sc;
    p.o_av[o] = (double)c.av;
    p.o_rv[o] = (double)c.rv;
    double* cc = p.o_cov + o * 9;
    cc[0] = cv.c[0] * sc * sc; cc[1] = cc[3] = cv.c[1] * sc; cc[2] = cc[6] = cv.c[2] * sc;
    cc[4] = cv.c[3]; cc[5] = cc[7] = cv.c[4]; cc[8] = cv.c[5];
    const T lnp = p.lnp2[u];
    p.o_lnprob[o] = lnp <= Num<T>::kNegBig ? -1e300 : (double)lnp;
    // pick one of the model's Nmc draws with probability ~ exp(lnp_mc) (:2050-2053)
    Lse<T> acc;
    for (int j = 0; j < p.nmc; j += 2) {   // the same paired evaluation as k_post_mc: identical log-priors
        McPair<T> m;
        mc_pair<T, ZOV>(p, c, j, m);
        acc.add(m.lnp(0));
        if (j + 1 < p.nmc) acc.add(m.lnp(1));
    }
    const T m = acc.m;
    const double W = (double)acc.s;
    double run = 0.;
    int pick = -1;
    T ps = c.scale, pa = c.av, pr = c.rv, pl = Num<T>::neg_inf();
    for (int j0 = 0; j0 < p.nmc && pick < 0; j0 += 2) {
        McPair<T> mp;
        mc_pair<T, ZOV>(p, c, j0, mp);
        const T sv[2] = {lo2(mp.s), hi2(mp.s)}, av[2] = {lo2(mp.a), hi2(mp.a)}, rv[2] = {lo2(mp.r), hi2(mp.r)};
        const T lv[2] = {mp.lnp(0), mp.lnp(1)};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int j = j0 + k;
            if (j >= p.nmc || pick >= 0) continue;
            // all draws at -inf: exp(-inf - -inf) is NaN in the reference too; fall through to the last draw
            if (lv[k] > Num<T>::neg_inf()) run += (double)pexp(lv[k] - m);
            if ((W > 0. && run / W > u2) || j == p.nmc - 1) { pick = j; ps = sv[k]; pa = av[k]; pr = rv[k]; pl = lv[k]; }
        }
    }
    p.o_dist[o] = 1. / sqrt((double)ps);
    p.o_red[o] = (double)pa;
    p.o_dred[o] = (double)pr;
    p.o_logwt[o] = pl <= Num<T>::kNegBig ? -1e300 : (double)pl;
    }
}

}  // namespace bf
