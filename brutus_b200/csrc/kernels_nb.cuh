// kernels_nb.cuh -- the band-count-templated kernels of the likelihood sweep.
//
// One thread owns one model; every sum over bands is a register accumulation (no shuffles).
// The arithmetic is the reference's (brutus/fitting.py:34-576) rewritten in
// per-star normalised, centred units so that float32 keeps ~1e-7 absolute accuracy on O(1) values:
//
//   centred magnitude residual  e'_j = e_j - c,  e_j = m_j - (mu_j + A r_j),  c = mbar - bbar_i
//   flux ratio                  F_j / d*_j = 10^(0.4 e_j) = g_j E,  g_j = 2^(kC2 e'_j), E = 2^(kC2 c)
//   sigma-normalised model      M_j/sigma_j = shat g_j be_j,  shat = s E,  residual t_j = al_j - shat g_j be_j
//
// so  scale s = (sum g be al / sum (g be)^2) / E   and   chi2 = sum t_j^2   (fitting.py:510-518, :745).
//
// The kernels are bound by FP32 instruction issue, not by HBM, so the band axis is
// processed two bands at a time with Blackwell's packed FP32 instructions (PTX fma/mul/add.rn.f32x2 ->
// SASS FFMA2/FMUL2/FADD2: one issue slot for two lanes' worth of work): bands (2p, 2p+1) live in one
// 64-bit register pair, per-band sums are kept as (even, odd) partial sums and folded once.  An odd
// band count is padded with a band of zero weight.  The float64 instantiation (verification path) uses
// the same code with a plain two-element struct.
#pragma once
#include "common.cuh"

namespace bf {

// Star-independent per-model quantities, hoisted out of the star loop.
constexpr int kRefitTile = kTile;  // threads per CTA of k_refit (128 measured neutral: finer survivor appends scatter the flux gathers)
constexpr int kFluxTile = 64;    // threads per CTA of k_flux (see the kernel)
template <typename T, int NB> struct ModelRegs {
    static constexpr int NP = (NB + 1) / 2;
    P2<T> ncb[NP];  // -(b_j - bbar), b_j = mu_j + Abar r0_j  (model magnitudes at the prior-mean reddening)
    P2<T> r0[NP];   // R_j + Rbar D_j                         (brutus/utils.py:337-338 at rv = rv_gauss[0])
    P2<T> D[NP];    // dR/dRv
    T bbar;
};

template <typename T, int NB>
__device__ __forceinline__ void finish_model(const T (&mu)[NB + 1], const T (&R)[NB + 1], const T (&D)[NB + 1],
                                             const DevOpts<T>& o, ModelRegs<T, NB>& m) {
    constexpr int NP = ModelRegs<T, NB>::NP;
    T r0[2 * NP], cb[2 * NP], Dd[2 * NP];
    T sum = T(0);
#pragma unroll
    for (int j = 0; j < 2 * NP; j++) {
        if (j < NB) {
            Dd[j] = D[j];
            r0[j] = fma(o.Rbar, D[j], R[j]);
            cb[j] = fma(o.Abar, r0[j], mu[j]);
            sum += cb[j];
        } else {
            Dd[j] = r0[j] = cb[j] = T(0);   // padding band: zero model terms, zero star weights
        }
    }
    m.bbar = sum * (T(1) / T(NB));
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const T c0 = m.bbar - cb[2 * p];
        const T c1 = (2 * p + 1 < NB) ? m.bbar - cb[2 * p + 1] : T(0);
        m.ncb[p] = mk2(c0, c1);
        m.r0[p] = mk2(r0[2 * p], r0[2 * p + 1]);
        m.D[p] = mk2(Dd[2 * p], Dd[2 * p + 1]);
    }
}

// from the coefficient-major grid (fully coalesced: thread = model)
template <typename T, int NB>
__device__ __forceinline__ void load_model(const float* __restrict__ grid, int64_t npad, int64_t i,
                                           const DevOpts<T>& o, ModelRegs<T, NB>& m) {
    T mu[NB + 1], R[NB + 1], D[NB + 1];
#pragma unroll
    for (int j = 0; j < NB; j++) {
        mu[j] = (T)__ldg(grid + (int64_t)(0 * NB + j) * npad + i);
        R[j] = (T)__ldg(grid + (int64_t)(1 * NB + j) * npad + i);
        D[j] = (T)__ldg(grid + (int64_t)(2 * NB + j) * npad + i);
    }
    finish_model<T, NB>(mu, R, D, o, m);
}

// from the model-major copy of the grid (3-4 sectors per model: the per-candidate gathers)
template <typename T, int NB>
__device__ __forceinline__ void load_model_row(const float* __restrict__ rows, int64_t i, const DevOpts<T>& o,
                                               ModelRegs<T, NB>& m) {
    constexpr int RS = row_stride(NB);
    float v[RS];
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(rows + i * RS);
#pragma unroll
    for (int k = 0; k < RS / 4; k++) {
        float4 t = __ldg(p4 + k);
        v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
    T mu[NB + 1], R[NB + 1], D[NB + 1];
#pragma unroll
    for (int j = 0; j < NB; j++) { mu[j] = (T)v[j]; R[j] = (T)v[NB + j]; D[j] = (T)v[2 * NB + j]; }
    finish_model<T, NB>(mu, R, D, o, m);
}

// Result of the flux-space MLE at fixed (A, rho): brutus/fitting.py:430-576 (_get_sed_mle), normalised.
template <typename T, int NB> struct Mle {
    P2<T> gb[(NB + 1) / 2];  // g_j be_j
    T shat;       // s E (after the 1e-20 floor on s)
    T s;          // scale                                  (:516-518)
    T E;          // 2^(kC2 c)
    T den;        // sum (g be)^2  (s_den = E^2 den, :515)
    T chi2;       // sum t^2                                 (:745 / :792)
};

template <typename T, int NB>
__device__ __forceinline__ void mle_from_resid(const P2<T> (&e)[(NB + 1) / 2], T c, const T* __restrict__ srow,
                                               Mle<T, NB>& r) {
    constexpr int NP = (NB + 1) / 2;
    P2<T> num = bc2(T(0)), den = bc2(T(0));
    const P2<T> k2 = bc2(T(kC2));
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const P2<T> x = mul2(e[p], k2);
        const T g0 = Num<T>::exp2(lo2(x));
        const T g1 = (2 * p + 1 < NB) ? Num<T>::exp2(hi2(x)) : T(0);
        const P2<T> gb = mul2(mk2(g0, g1), ld2(srow + SR_BE + 2 * p));
        r.gb[p] = gb;
        num = fma2(gb, ld2(srow + SR_AL + 2 * p), num);
        den = fma2(gb, gb, den);
    }
    // shat = num/den may use the approximate reciprocal: chi2 is stationary in shat at the MLE
    // (sum t_j gb_j = 0), so a 2-ulp error in shat enters chi2 only at second order.
    const T E = Num<T>::exp2(T(kC2) * c);
    const T Einv = Num<T>::exp2(T(-kC2) * c);
    const T dens = hsum2(den);
    T shat = Num<T>::div_fast(hsum2(num), dens);
    T s = shat * Einv;
    const bool floor_s = s <= T(1e-20);  // brutus/fitting.py:517-518
    s = floor_s ? T(1e-20) : s;
    shat = floor_s ? T(1e-20) * E : shat;
    P2<T> chi = bc2(T(0));
    const P2<T> nsh = bc2(-shat);
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const P2<T> t = fma2(nsh, r.gb[p], ld2(srow + SR_AL + 2 * p));
        chi = fma2(t, t, chi);
    }
    r.shat = shat; r.s = s; r.E = E; r.den = dens; r.chi2 = hsum2(chi);
}

// lnl_p of the cull (brutus/fitting.py:747-756): -chi2/2 - (sqrt(s) - parallax)^2 / (2 parallax_err^2)
template <typename T>
__device__ __forceinline__ T cull_lnl(T chi2, T s, const T* __restrict__ srow) {
    T dp = Num<T>::sqrt_fast(s) - srow[SR_SC + SC_PAR];
    return T(-0.5) * fma(dp * dp, srow[SR_SC + SC_PIVAR], chi2);
}

// One (Av, Rv) update of the magnitude fit for one (model, star): brutus/fitting.py:174-243.
// Carries gs = sum e' u between iterations; gs2 and the post-update gs follow algebraically from
// the updates (e' -= dA r; e' -= A dR D) instead of being re-summed over the bands.
template <typename T, int NB>
__device__ __forceinline__ void mag_iter(const ModelRegs<T, NB>& m, const DevOpts<T>& o,
                                         const P2<T> (&u)[(NB + 1) / 2], T S, T c, T Q, T Tm,
                                         P2<T> (&e)[(NB + 1) / 2], P2<T> (&r)[(NB + 1) / 2], T& A, T& rho,
                                         T& gs, T& ell, T& delta) {
    constexpr int NP = (NB + 1) / 2;
    // --- solve for Av (:176-204) ---
    P2<T> a2 = bc2(T(0)), b2 = bc2(T(0)), ga2 = bc2(T(0));
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const P2<T> ru = mul2(r[p], u[p]);
        a2 = fma2(ru, r[p], a2);
        b2 = add2(b2, ru);
        ga2 = fma2(ru, e[p], ga2);
    }
    const T a = hsum2(a2) + o.PA, b = hsum2(b2);
    const T ga = fma(o.Abar - A, o.PA, hsum2(ga2));
    T dA = Num<T>::div_fast(S * ga - b * gs, S * a - b * b);
    dA = Num<T>::max(dA, o.avmin - A);
    dA = Num<T>::min(dA, o.avmax - A);
    A += dA;
    // --- solve for Rv (:206-237) ---
    const T gs2 = fma(-dA, b, gs);             // sum (e' - dA r) u
    const P2<T> ndA = bc2(-dA);
    P2<T> gr2 = bc2(T(0));
#pragma unroll
    for (int p = 0; p < NP; p++) {
        e[p] = fma2(ndA, r[p], e[p]);
        gr2 = fma2(mul2(e[p], u[p]), m.D[p], gr2);
    }
    const T gr = fma(hsum2(gr2), A, (o.Rbar - rho) * o.PR);
    const T q = fma(Q * A, A, o.PR);
    const T tt = Tm * A;
    T dR = Num<T>::div_fast(S * gr - tt * gs2, S * q - tt * tt);
    dR = Num<T>::max(dR, o.rvmin - rho);
    dR = Num<T>::min(dR, o.rvmax - rho);
    rho += dR;
    // --- update residuals / reddening vector, chi2 in magnitudes (:235-243) ---
    const T AdR = A * dR;
    gs = fma(-AdR, Tm, gs2);                   // sum (e' - A dR D) u
    const P2<T> nAdR = bc2(-AdR), dR2 = bc2(dR);
    P2<T> chi = bc2(T(0));
#pragma unroll
    for (int p = 0; p < NP; p++) {
        e[p] = fma2(nAdR, m.D[p], e[p]);
        r[p] = fma2(dR2, m.D[p], r[p]);
        chi = fma2(mul2(e[p], u[p]), e[p], chi);
    }
    // logwt uses the un-centred residual e = e' + c (reference quirk)
    ell = T(-0.5) * (hsum2(chi) + c * (T(2) * gs + c * S));
    delta = Num<T>::max(tabs(dA), tabs(dR));
}

// warp-wide max.  float: one CREDUX.MAX.F32 (sm_100a redux.sync on f32; NaN inputs are ignored);
// double: shuffle butterfly.
__device__ __forceinline__ float warp_max_fast(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ double warp_max_fast(double v) { return warp_max((v == v) ? v : -CUDART_INF); }

// initial residuals (brutus/fitting.py:728-733) and the star-weighted model sums of :158-164
template <typename T, int NB>
__device__ __forceinline__ void mag_init(const ModelRegs<T, NB>& m, const T* __restrict__ srow,
                                         P2<T> (&u)[(NB + 1) / 2], P2<T> (&e)[(NB + 1) / 2],
                                         P2<T> (&r)[(NB + 1) / 2], T& Q, T& Tm, T& gs) {
    constexpr int NP = (NB + 1) / 2;
    P2<T> Q2 = bc2(T(0)), T2 = bc2(T(0)), g2 = bc2(T(0));
#pragma unroll
    for (int p = 0; p < NP; p++) {
        u[p] = ld2(srow + SR_U + 2 * p);
        e[p] = add2(ld2(srow + SR_CM + 2 * p), m.ncb[p]);
        r[p] = m.r0[p];
        const P2<T> Du = mul2(m.D[p], u[p]);
        Q2 = fma2(Du, m.D[p], Q2);
        T2 = add2(T2, Du);
        g2 = fma2(e[p], u[p], g2);
    }
    Q = hsum2(Q2); Tm = hsum2(T2); gs = hsum2(g2);
}

// The whole magnitude-space fit of one (model, star) pair: initial residuals (brutus/fitting.py:728-733),
// `kspec` iterations of _optimize_fit_mag (:173-264), leaving the centred residuals in e.  (l0, b0) and
// (l1, b1) are the reduction inputs of iterations kspec-1 and kspec: logwt, and logwt where the step
// max(|dAv|, |dRv|) is still >= tol (else -inf).  Shared by the sweep and the candidate re-fit so that
// both evaluate bit-identical arithmetic.
template <typename T, int NB>
__device__ __forceinline__ void magfit_one(const ModelRegs<T, NB>& m, const DevOpts<T>& o,
                                           const T* __restrict__ srow, int kspec, T c,
                                           P2<T> (&e)[(NB + 1) / 2], T& A, T& rho, T& l0, T& b0, T& l1, T& b1) {
    constexpr int NP = (NB + 1) / 2;
    const T ninf = Num<T>::neg_inf();
    const T S = srow[SR_SC + SC_S];
    A = o.Abar; rho = o.Rbar;
    P2<T> u[NP], r[NP];
    T Q, Tm, gs;
    mag_init<T, NB>(m, srow, u, e, r, Q, Tm, gs);
    T ell = T(0), delta = T(0);
    l0 = ninf; b0 = ninf;
    if (kspec == 2) {   // the common case, fully unrolled
        mag_iter<T, NB>(m, o, u, S, c, Q, Tm, e, r, A, rho, gs, ell, delta);
        l0 = ell;
        b0 = (delta >= o.mtol) ? l0 : ninf;
        mag_iter<T, NB>(m, o, u, S, c, Q, Tm, e, r, A, rho, gs, ell, delta);
    } else {
        for (int k = 1; k <= kspec; k++) {
            mag_iter<T, NB>(m, o, u, S, c, Q, Tm, e, r, A, rho, gs, ell, delta);
            if (k == kspec - 1) {
                l0 = ell;
                b0 = (delta >= o.mtol) ? l0 : ninf;
            }
        }
    }
    l1 = ell;
    b1 = (delta >= o.mtol) ? l1 : ninf;
}

// =================================================================================================
// Kernel 0: iteration-count probe.  Runs kProbeIter mag iterations on every `tile_stride`-th model
// tile and records, per star and iteration k, the two maxima the reference's stopping rule needs
// (:246-263).  The host turns them into the speculated iteration count of the full sweep, which
// verifies it on the whole grid (so a wrong guess costs a re-sweep, never a wrong answer).
// =================================================================================================
template <typename T, int NB>
__global__ void __launch_bounds__(kTile) k_kprobe(const ProbeParams<T> p) {
    using U = typename Enc<T>::U;
    constexpr int NP = (NB + 1) / 2;
    __shared__ __align__(16) T s_star[kStarChunk][kStarStride];
    __shared__ U s_red[kStarChunk][2 * kProbeIter];
    const int first = blockIdx.y * kStarChunk;
    const int nst = min(kStarChunk, p.nstar - first);
    for (int t = threadIdx.x; t < nst * kStarStride; t += kTile)
        s_star[t / kStarStride][t % kStarStride] = p.stars[(int64_t)first * kStarStride + t];
    for (int t = threadIdx.x; t < kStarChunk * 2 * kProbeIter; t += kTile)
        s_red[t / (2 * kProbeIter)][t % (2 * kProbeIter)] = Enc<T>::enc(Num<T>::neg_inf());
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * p.tile_stride * kTile + threadIdx.x;   // < npad; padding replicates a real model
    const DevOpts<T> o = p.o;
    ModelRegs<T, NB> m;
    load_model<T, NB>(p.grid, p.npad, i, o, m);
    const int lane = threadIdx.x & 31;
    const T ninf = Num<T>::neg_inf();
    T acc[2 * kProbeIter];
#pragma unroll
    for (int k = 0; k < 2 * kProbeIter; k++) acc[k] = ninf;
    for (int s = 0; s < nst; s++) {
        const T* __restrict__ srow = s_star[s];
        const T S = srow[SR_SC + SC_S];
        const T c = srow[SR_SC + SC_MBAR] - m.bbar;
        T A = o.Abar, rho = o.Rbar;
        P2<T> u[NP], e[NP], r[NP];
        T Q, Tm, gs;
        mag_init<T, NB>(m, srow, u, e, r, Q, Tm, gs);
#pragma unroll
        for (int k = 0; k < kProbeIter; k++) {
            T ell, delta;
            mag_iter<T, NB>(m, o, u, S, c, Q, Tm, e, r, A, rho, gs, ell, delta);
            T l = ell;
            T b = (delta >= o.mtol) ? l : ninf;
            l = warp_max_fast(l);
            b = warp_max_fast(b);
            if (lane == s) { acc[2 * k] = l; acc[2 * k + 1] = b; }
        }
    }
    if (lane < nst) {
#pragma unroll
        for (int k = 0; k < 2 * kProbeIter; k++)
            if (acc[k] == acc[k]) atomicMax(&s_red[lane][k], Enc<T>::enc(acc[k]));
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nst * 2 * kProbeIter; t += kTile)
        atomicMax(&p.out[(int64_t)first * 2 * kProbeIter + t], s_red[t / (2 * kProbeIter)][t % (2 * kProbeIter)]);
}

// =================================================================================================
// Kernel 1: full-grid magnitude-space fit (brutus/fitting.py:728-741 -> _optimize_fit_mag :34-271,
// then _get_sed_mle :267, the cull statistic :745-756 and a provisional lnprob) for a list of stars.
// grid = (model tiles, star chunks); each thread keeps its model in registers and loops over the
// chunk's stars, whose rows sit in shared memory (broadcast reads).
//
// Nothing per (model, star) is stored except ONE BIT: whether the pair can still matter, i.e. whether
// it may survive the cull (lnl_p > max + ln init_thresh) or pass lnpost's first selection
// (lnprob > max + ln wt_thresh).  Both tests are relative to per-star maxima that are only known after
// the sweep, so they are evaluated against a running maximum (warp-local maximum combined with the
// per-star global maximum published by the CTAs that already finished): a running maximum never
// exceeds the final one, hence the flagged set is a superset; the exact tests are re-applied when
// the flagged pairs are re-fitted (k_refit).
//
// The number of mag iterations applied to every model of a star is a grid-wide decision in the
// reference (:246-263).  It is speculated here (SI_KSPEC) and verified afterwards from two plain
// max-reductions per iteration:  "err < tol"  <=>  max{logwt_i : max(|dAv_i|,|dRv_i|) >= tol} <=
// max logwt + ln(init_thresh).
// Per-star reductions: lane s of every warp keeps the warp's maxima for star s of the chunk
// (kStarChunk == 32), so the star loop contains no shared-memory atomics.
// =================================================================================================
template <typename T, int NB>
// up to 8 bands: capped at 80 registers (3 CTAs = 24 warps per SM); left to itself ptxas takes 94 and the
// kernel loses a third of its warps (measured 15.8 -> 16.4 ms per 1 000 stars)
__global__ void __launch_bounds__(kTile, (NB <= 8 ? 3 : 2)) k_magfit(const SweepParams<T> p) {
    using U = typename Enc<T>::U;
    constexpr int NP = (NB + 1) / 2;
    static_assert(kStarChunk == 32, "lane <-> star mapping of the reductions");
    __shared__ __align__(16) T s_star[kStarChunk][kStarStride];
    __shared__ int s_slot[kStarChunk];
    __shared__ int s_kspec[kStarChunk];
    __shared__ T s_snap[kStarChunk][2];
    __shared__ U s_red[kStarChunk][kSweepRed];

    const int first = blockIdx.y * kStarChunk;
    const int nst = min(kStarChunk, p.nlist - first);
    for (int t = threadIdx.x; t < nst * kStarStride; t += kTile) {
        int s = t / kStarStride, k = t - s * kStarStride;
        s_star[s][k] = p.stars[(int64_t)p.list[first + s] * kStarStride + k];
    }
    for (int t = threadIdx.x; t < nst; t += kTile) {
        int slot = p.list[first + t];
        s_slot[t] = slot;
        s_kspec[t] = p.star_int[slot * SI_COUNT + SI_KSPEC];
        // running per-star maxima published so far (benign race: any value <= the final maximum is valid)
        const volatile U* rr = p.red + (int64_t)slot * kNumRed;
        s_snap[t][0] = Enc<T>::dec(rr[RED_LP]);
        s_snap[t][1] = Enc<T>::dec(rr[RED_M0]);
    }
    for (int t = threadIdx.x; t < kStarChunk * kSweepRed; t += kTile)
        s_red[t / kSweepRed][t % kSweepRed] = Enc<T>::enc(Num<T>::neg_inf());
    __syncthreads();

    const int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x;  // npad is a multiple of kTile
    const bool valid = i < p.nmodel;
    const DevOpts<T> o = p.o;
    ModelRegs<T, NB> m;
    load_model<T, NB>(p.grid, p.npad, i, o, m);
    const int lane = threadIdx.x & 31;
    const int64_t word = i >> 5;
    const T ninf = Num<T>::neg_inf();
    const T ln_init_c = o.ln_init - T(kCandMargin);
    T acc0 = ninf, acc1 = ninf, acc2 = ninf, acc3 = ninf, acc4 = ninf, acc5 = ninf;  // lane s <-> star s

    for (int s = 0; s < nst; s++) {
        const T* __restrict__ srow = s_star[s];
        const T c = srow[SR_SC + SC_MBAR] - m.bbar;
        T A, rho, l0, b0, l1, b1;
        P2<T> e[NP];
        magfit_one<T, NB>(m, o, srow, s_kspec[s], c, e, A, rho, l0, b0, l1, b1);
        // --- _get_sed_mle at the fitted (Av, Rv) (:267) and the cull statistic (:745-756) ---
        Mle<T, NB> r4;
        mle_from_resid<T, NB>(e, c, srow, r4);
        T lp = cull_lnl(r4.chi2, r4.s, srow);
        // --- provisional lnlike / lnprob from the mag-fit values (final for every non-survivor) ---
        const int slot = s_slot[s];
        T ext = T(0);
        if (p.nlabel > 0) ext = ext_prior<T>(p.labels, p.ext + (int64_t)slot * p.nlabel * 3, p.nlabel, p.npad, i);
        T lnl0, lq;
        lnl_lnprob<T>(r4.chi2, r4.den * r4.E * r4.E, r4.s, false, srow, o.dim_prior, ext, lnl0, lq);
        // no `valid` masks on the reductions: padding models replicate the last real model (k_retile)
        l0 = warp_max_fast(l0); b0 = warp_max_fast(b0);
        l1 = warp_max_fast(l1); b1 = warp_max_fast(b1);
        const T lpm = warp_max_fast(lp);
        const T lqm = warp_max_fast(lq);
        if (lane == s) { acc0 = l0; acc1 = b0; acc2 = l1; acc3 = b1; acc4 = lpm; acc5 = lqm; }
        // --- candidate bit ---
        const T thr1 = Num<T>::max(s_snap[s][0], lpm) + ln_init_c;
        const T thr2 = Num<T>::max(s_snap[s][1], lqm) + o.ln_wt - srow[SR_SC + SC_SLACK];
        const bool cand = valid && (lp > thr1 || lq > thr2);
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) p.cand[(int64_t)slot * p.nwords + word] = bal;
    }
    // NaN maxima (every lane NaN) must not poison the unsigned-encoded atomics
    if (lane < nst) {
        if (acc0 == acc0) atomicMax(&s_red[lane][0], Enc<T>::enc(acc0));
        if (acc1 == acc1) atomicMax(&s_red[lane][1], Enc<T>::enc(acc1));
        if (acc2 == acc2) atomicMax(&s_red[lane][2], Enc<T>::enc(acc2));
        if (acc3 == acc3) atomicMax(&s_red[lane][3], Enc<T>::enc(acc3));
        if (acc4 == acc4) atomicMax(&s_red[lane][4], Enc<T>::enc(acc4));
        if (acc5 == acc5) atomicMax(&s_red[lane][5], Enc<T>::enc(acc5));
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nst * kSweepRed; t += kTile) {
        int s = t / kSweepRed, k = t % kSweepRed;
        const int map[kSweepRed] = {RED_L0, RED_B0, RED_L1, RED_B1, RED_LP, RED_M0};
        atomicMax(&p.red[(int64_t)s_slot[s] * kNumRed + map[k]], s_red[s][k]);
    }
}

// =================================================================================================
// Kernel 2: exact re-fit of the candidates.  One thread per candidate record (star, model): repeats the
// sweep's arithmetic for the pair, applies the exact cull test (:758-759) against the now final
// per-star maximum, initialises the record and appends the survivors to the compact flux working set
// (stepsize 1, lnl_old = -1e300, :778-779).
// =================================================================================================
template <typename T, int NB>
// up to 8 bands: 64 registers -> 4 CTAs per SM: the kernel waits on gathers (ncu: long_scoreboard 6.7 per issue),
// occupancy helps; more bands need more registers (3 CTAs up to 12 bands, 2 beyond, to stay clear of spills)
__global__ void __launch_bounds__(kRefitTile, (NB <= 8 ? 4 : (NB <= 12 ? 3 : 2)) * kTile / kRefitTile) k_refit(const RefitParams<T> p) {
    constexpr int NP = (NB + 1) / 2;
    const int64_t q = (int64_t)blockIdx.x * kRefitTile + threadIdx.x;
    const bool inr = q < p.ncand;
    bool surv = false;
    int slot = -1, model = 0;
    T Av = T(0), Rv = T(0), chi2 = T(0), scale = T(0), sden = T(0);
    if (inr) {
        slot = p.pool.star[q];
        const int64_t i = p.pool.model[q];
        const DevOpts<T> o = p.o;
        const T* __restrict__ srow = p.stars + (int64_t)slot * kStarStride;
        ModelRegs<T, NB> m;
        load_model_row<T, NB>(p.rows, i, o, m);
        const T c = srow[SR_SC + SC_MBAR] - m.bbar;
        T A, rho, l0, b0, l1, b1;
        P2<T> e[NP];
        magfit_one<T, NB>(m, o, srow, p.star_int[slot * SI_COUNT + SI_KSPEC], c, e, A, rho, l0, b0, l1, b1);
        Mle<T, NB> r4;
        mle_from_resid<T, NB>(e, c, srow, r4);
        const T lp = cull_lnl(r4.chi2, r4.s, srow);
        surv = lp > Enc<T>::dec(p.red[(int64_t)slot * kNumRed + RED_LP]) + o.ln_init;
        p.pool.av[q] = A;
        p.pool.rv[q] = rho;
        p.pool.chi2[q] = chi2 = r4.chi2;
        p.pool.scale[q] = scale = r4.s;
        p.pool.sden[q] = sden = r4.den * r4.E * r4.E;
        p.pool.flag[q] = surv ? kFlagSurv : 0;
        Av = A; Rv = rho; model = (int)i;
    }
    // append the survivors to the compact flux working set: one global atomic per CTA
    __shared__ int s_w[kRefitTile / 32];
    __shared__ int s_base;
    const unsigned bal = __ballot_sync(0xffffffffu, surv);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_w[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < kRefitTile / 32; k++) n += s_w[k];
        s_base = n ? atomicAdd(p.nsv, n) : 0;
    }
    __syncthreads();
    if (surv) {
        int pos = s_base + __popc(bal & ((1u << lane) - 1u));
        for (int k = 0; k < w; k++) pos += s_w[k];
        p.sv.q[pos] = (int)q;
        p.sv.model[pos] = model;
        p.sv.star[pos] = slot;
        p.sv.av[pos] = Av;
        p.sv.rv[pos] = Rv;
        p.sv.eta[pos] = T(1);                 // stepsize 1, lnl_old = -1e300 (:778-779)
        p.sv.lold[pos] = Num<T>::kNegBig;
        p.sv.chi2[pos] = chi2;
        p.sv.scale[pos] = scale;
        p.sv.sden[pos] = sden;
    }
    cta_star_count(p.nsurv, slot, surv);
}

// residuals at an arbitrary (A, rho): e'_j = cm_j - cb_j - (A r_j - Abar r0_j), r_j = r0_j + (rho - Rbar) D_j
template <typename T, int NB>
__device__ __forceinline__ void resid_at(const ModelRegs<T, NB>& m, const DevOpts<T>& o,
                                         const T* __restrict__ srow, T A, T rho, P2<T> (&e)[(NB + 1) / 2],
                                         P2<T> (&r)[(NB + 1) / 2]) {
    constexpr int NP = (NB + 1) / 2;
    const P2<T> drho = bc2(rho - o.Rbar), A2 = bc2(A), nAbar = bc2(-o.Abar);
#pragma unroll
    for (int p = 0; p < NP; p++) {
        r[p] = fma2(drho, m.D[p], m.r0[p]);
        const P2<T> red = fma2(A2, r[p], mul2(nAbar, m.r0[p]));
        e[p] = sub2(add2(ld2(srow + SR_CM + 2 * p), m.ncb[p]), red);
    }
}

// =================================================================================================
// Kernel 3: flux-space refinement of the survivors (brutus/fitting.py:778-803 with
// _optimize_fit_flux :274-427).  One thread per survivor; `nit` iterations are executed in
// registers, and the convergence reductions are recorded for the last one:
//   "lerr <= ltol"  <=>  max{lnl_new_i : |lnl_new_i - lnl_old_i| > ltol} <= max lnl_new + ln(ltol_subthresh)
// Stars whose loop has converged (SI_ACTIVE == 0, decided on the device by k_flux_ctl) are skipped.
// =================================================================================================
template <typename T, int NB>
// 80 registers -> 24 warps per SM (16 at the natural 90 registers; 64 registers spill and measured slower).
// CTAs of kFluxTile = 64 threads: every thread waits on a two-level gather and then runs ~650 dependent
// instructions, and the per-star reductions at the end need CTA-wide barriers, so with 256-thread CTAs the
// warps that finish early idle at the barrier while their registers stay allocated (ncu: issue 26 %, DRAM
// 22 %).  Measured per 1 000 stars: 256 threads 8.6 ms, 128: 7.7, 64: 7.4, 32: 8.5 (global atomics per CTA).
__global__ void __launch_bounds__(kFluxTile, (NB <= 8 ? 3 : 2) * kTile / kFluxTile) k_flux(const FluxParams<T> p) {
    constexpr int NP = (NB + 1) / 2;
    const int64_t t = (int64_t)blockIdx.x * kFluxTile + threadIdx.x;
    const bool inrange = t < p.nsv;
    int slot = inrange ? p.sv.star[t] : -1;
    const bool act = inrange && p.star_int[slot * SI_COUNT + SI_ACTIVE] != 0;
    const DevOpts<T> o = p.o;
    T v1 = Num<T>::neg_inf(), v2 = Num<T>::neg_inf();
    if (act) {
        const int64_t i = p.sv.model[t];
        const T* __restrict__ srow = p.stars + (int64_t)slot * kStarStride;
        ModelRegs<T, NB> m;
        load_model_row<T, NB>(p.rows, i, o, m);
        const T c = srow[SR_SC + SC_MBAR] - m.bbar;
        T A = p.sv.av[t], rho = p.sv.rv[t], eta = p.sv.eta[t], lold = p.sv.lold[t];
        P2<T> e[NP], r[NP];
        Mle<T, NB> r4;
        resid_at<T, NB>(m, o, srow, A, rho, e, r);
        mle_from_resid<T, NB>(e, c, srow, r4);
        T lnew = lold;
        for (int it = 0; it < p.nit; it++) {
            // one (dAv, dRv) step from the current model / residuals (:385-420)
            P2<T> an = bc2(T(0)), ad = bc2(T(0)), rn = bc2(T(0)), rd = bc2(T(0));
            const P2<T> sh = bc2(r4.shat);
#pragma unroll
            for (int pp = 0; pp < NP; pp++) {
                const P2<T> Ms = mul2(sh, r4.gb[pp]);                        // M_j / sigma_j
                const P2<T> tj = sub2(ld2(srow + SR_AL + 2 * pp), Ms);       // resid_j / sigma_j
                const P2<T> rM = mul2(r[pp], Ms), DM = mul2(m.D[pp], Ms);
                an = fma2(rM, tj, an);
                ad = fma2(rM, rM, ad);
                rn = fma2(DM, tj, rn);
                rd = fma2(DM, DM, rd);
            }
            T dA = Num<T>::div(fma(T(kFac), hsum2(an), (o.Abar - A) * o.PA), fma(T(kFac * kFac), hsum2(ad), o.PA)) * eta;
            T dR = Num<T>::div(fma(T(kFac), hsum2(rn), (o.Rbar - rho) * o.PR), fma(T(kFac * kFac), hsum2(rd), o.PR)) * eta;
            dA = tmax(dA, o.avmin - A);
            dA = tmin(dA, o.avmax - A);
            A += dA;
            dR = tmax(dR, o.rvmin - rho);
            dR = tmin(dR, o.rvmax - rho);
            rho += dR;
            resid_at<T, NB>(m, o, srow, A, rho, e, r);
            mle_from_resid<T, NB>(e, c, srow, r4);          // :423
            lnew = T(-0.5) * r4.chi2;                       // :792-795
            if (it == p.nit - 1) {
                v1 = (lnew == lnew) ? lnew : Num<T>::neg_inf();
                v2 = (tabs(lnew - lold) > o.ltol) ? v1 : Num<T>::neg_inf();
            }
            if (lnew < lold) eta = eta / T(1.2);            // :802
            lold = lnew;                                    // :803
        }
        p.sv.av[t] = A;
        p.sv.rv[t] = rho;
        p.sv.eta[t] = eta;
        p.sv.lold[t] = lold;
        p.sv.chi2[t] = r4.chi2;
        p.sv.scale[t] = r4.s;
        p.sv.sden[t] = r4.den * r4.E * r4.E;
    }
    // per-star reductions, one pair of global atomics per CTA and star
    if (__syncthreads_or(act)) {
        cta_star_max<T>(p.red, RED_FL, slot, act, v1);
        cta_star_max<T>(p.red, RED_FB, slot, act, v2);
    }
}

// =================================================================================================
// Kernel 4: output records.  Recomputes _get_sed_mle (brutus/fitting.py:502-576) at the final
// (Av, Rv) of each requested candidate to produce the full precision matrix icov_sar.
// Mode A: compacted records of the selected candidates (6 unique icov entries, element type T);
// mode B: every model of one star (pool entry q == model q; 9 entries, float64: what loglike returns).
// =================================================================================================
template <typename T, int NB, typename O>
__global__ void __launch_bounds__(kTile) k_records(const RecordParams<T, O> p) {
    constexpr int NP = (NB + 1) / 2;
    const int64_t t = (int64_t)blockIdx.x * kTile + threadIdx.x;
    if (t >= p.nrec) return;
    const bool modeA = p.sel_q != nullptr;
    const int64_t q = modeA ? p.sel_q[t] : t;
    const int slot = p.pool.star[q];
    const int64_t i = p.pool.model[q];
    const DevOpts<T> o = p.o;
    const T* __restrict__ srow = p.stars + (int64_t)slot * kStarStride;
    const T A = p.pool.av[q], rho = p.pool.rv[q];
    if (modeA) {
        p.o_idx[t] = (int)i;
        if (p.o_star) p.o_star[t] = slot;
        p.o_lnl[t] = (O)p.pool.lnl[q];
        p.o_scale[t] = (O)p.pool.scale[q];
        p.o_av[t] = (O)A;
        if (p.nrows > 3) {
            p.o_chi2[t] = (O)p.pool.chi2[q];
            p.o_rv[t] = (O)rho;
        }
        if (p.nrows <= 5) return;
    } else {
        p.o_lnl[t] = (O)p.pool.lnl[q];
        p.o_chi2[t] = (O)p.pool.chi2[q];
        p.o_scale[t] = (O)p.pool.scale[q];
        p.o_av[t] = (O)A;
        p.o_rv[t] = (O)rho;
        if (!p.o_icov) return;
    }
    ModelRegs<T, NB> m;
    load_model_row<T, NB>(p.rows, i, o, m);
    const T c = srow[SR_SC + SC_MBAR] - m.bbar;
    P2<T> e[NP], r[NP];
    Mle<T, NB> r4;
    resid_at<T, NB>(m, o, srow, A, rho, e, r);
    mle_from_resid<T, NB>(e, c, srow, r4);
    // cross terms (:526-561) in sigma-normalised units; see the header comment and DESIGN.md
    P2<T> sa = bc2(T(0)), sr = bc2(T(0)), ar = bc2(T(0)), aden = bc2(T(0)), rden = bc2(T(0));
    const P2<T> sh = bc2(r4.shat), kA = bc2(T(kC2) * A), one = bc2(T(1));
#pragma unroll
    for (int pp = 0; pp < NP; pp++) {
        const P2<T> Ms = mul2(sh, r4.gb[pp]);
        const P2<T> tj = sub2(ld2(srow + SR_AL + 2 * pp), Ms);
        const P2<T> x = mul2(kA, r[pp]);
        const P2<T> h = mk2(Num<T>::exp2(lo2(x)), Num<T>::exp2(hi2(x)));   // F0_j / F_j = 10^(0.4 A r_j)   (:529-530)
        const P2<T> mmr = sub2(Ms, tj);                                     // (models - resid)/sigma         (:539-542)
        sa = fma2(mul2(r[pp], r4.gb[pp]), mmr, sa);
        sr = fma2(mul2(m.D[pp], r4.gb[pp]), mmr, sr);
        const P2<T> DM = mul2(m.D[pp], Ms), rM = mul2(r[pp], Ms);
        ar = fma2(DM, sub2(mul2(Ms, sub2(one, h)), tj), ar);                // drvecs (reddening - resid)/var (:550-551)
        aden = fma2(rM, rM, aden);
        rden = fma2(DM, DM, rden);
    }
    const O f = (O)kFac, E = (O)r4.E;
    const O ss = (O)r4.den * E * E;
    const O dsa = f * E * (O)hsum2(sa), dsr = f * E * (O)hsum2(sr);
    const O dar = f * (O)hsum2(ar);
    const O daa = f * f * (O)hsum2(aden) + (O)o.PA + (O)(1. / (0.05 * 0.05));
    const O drr = f * f * (O)hsum2(rden) + (O)o.PR + (O)(1. / (0.1 * 0.1));
    if (modeA) {
        O* w = p.o_icov + t;
        w[0] = ss; w[p.ld] = dsa; w[2 * p.ld] = dsr; w[3 * p.ld] = daa; w[4 * p.ld] = dar; w[5 * p.ld] = drr;
    } else {
        O* w = p.o_icov + t * 9;
        w[0] = ss; w[1] = dsa; w[2] = dsr; w[3] = dsa; w[4] = daa; w[5] = dar; w[6] = dsr; w[7] = dar; w[8] = drr;
    }
}

// ---- launchers -------------------------------------------------------------------------------------
template <typename T, int NB> void launch_kprobe(const ProbeParams<T>& p, cudaStream_t st) {
    const int64_t ntile = p.npad / kTile;
    dim3 grid((unsigned)((ntile + p.tile_stride - 1) / p.tile_stride), (unsigned)((p.nstar + kStarChunk - 1) / kStarChunk));
    k_kprobe<T, NB><<<grid, kTile, 0, st>>>(p);
}
template <typename T, int NB> void launch_magfit(const SweepParams<T>& p, cudaStream_t st) {
    dim3 grid((unsigned)(p.npad / kTile), (unsigned)((p.nlist + kStarChunk - 1) / kStarChunk));
    k_magfit<T, NB><<<grid, kTile, 0, st>>>(p);
}
template <typename T, int NB> void launch_refit(const RefitParams<T>& p, cudaStream_t st) {
    if (p.ncand <= 0) return;
    k_refit<T, NB><<<(unsigned)((p.ncand + kRefitTile - 1) / kRefitTile), kRefitTile, 0, st>>>(p);
}
template <typename T, int NB> void launch_flux(const FluxParams<T>& p, cudaStream_t st) {
    if (p.nsv <= 0) return;
    k_flux<T, NB><<<(unsigned)((p.nsv + kFluxTile - 1) / kFluxTile), kFluxTile, 0, st>>>(p);
}
template <typename T, int NB, typename O> void launch_records(const RecordParams<T, O>& p, cudaStream_t st) {
    if (p.nrec <= 0) return;
    k_records<T, NB, O><<<(unsigned)((p.nrec + kTile - 1) / kTile), kTile, 0, st>>>(p);
}

}  // namespace bf
