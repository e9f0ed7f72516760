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
#pragma once
#include "common.cuh"

namespace bf {

// Star-independent per-model quantities, hoisted out of the star loop.
template <typename T, int NB> struct ModelRegs {
    T cb[NB];   // b_j - bbar, b_j = mu_j + Abar r0_j  (model magnitudes at the prior-mean reddening)
    T r0[NB];   // R_j + Rbar D_j                      (brutus/utils.py:337-338 at rv = rv_gauss[0])
    T D[NB];    // dR/dRv
    T bbar;
};

template <typename T, int NB>
__device__ __forceinline__ void load_model(const float* __restrict__ grid, int64_t npad, int64_t i,
                                           const DevOpts<T>& o, ModelRegs<T, NB>& m) {
    T sum = T(0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
        T mu = (T)__ldg(grid + (int64_t)(0 * NB + j) * npad + i);
        T R = (T)__ldg(grid + (int64_t)(1 * NB + j) * npad + i);
        T D = (T)__ldg(grid + (int64_t)(2 * NB + j) * npad + i);
        m.D[j] = D;
        m.r0[j] = fma(o.Rbar, D, R);
        m.cb[j] = fma(o.Abar, m.r0[j], mu);
        sum += m.cb[j];
    }
    m.bbar = sum * (T(1) / T(NB));
#pragma unroll
    for (int j = 0; j < NB; j++) m.cb[j] -= m.bbar;
}

// Result of the flux-space MLE at fixed (A, rho): brutus/fitting.py:430-576 (_get_sed_mle), normalised.
template <typename T, int NB> struct Mle {
    T gb[NB];     // g_j be_j
    T shat;       // s E (after the 1e-20 floor on s)
    T s;          // scale                                  (:516-518)
    T E;          // 2^(kC2 c)
    T den;        // sum (g be)^2  (s_den = E^2 den, :515)
    T chi2;       // sum t^2                                 (:745 / :792)
};

template <typename T, int NB>
__device__ __forceinline__ void mle_from_resid(const T (&e)[NB], T c, const T* __restrict__ srow,
                                               Mle<T, NB>& r) {
    T num = T(0), den = T(0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
        T g = Num<T>::exp2(T(kC2) * e[j]);
        T gb = g * srow[SR_BE + j];
        r.gb[j] = gb;
        num = fma(gb, srow[SR_AL + j], num);
        den = fma(gb, gb, den);
    }
    // shat = num/den may use the approximate reciprocal: chi2 is stationary in shat at the MLE
    // (sum t_j gb_j = 0), so a 2-ulp error in shat enters chi2 only at second order.
    const T E = Num<T>::exp2(T(kC2) * c);
    const T Einv = Num<T>::exp2(T(-kC2) * c);
    T shat = Num<T>::div_fast(num, den);
    T s = shat * Einv;
    const bool floor_s = s <= T(1e-20);  // brutus/fitting.py:517-518
    s = floor_s ? T(1e-20) : s;
    shat = floor_s ? T(1e-20) * E : shat;
    T chi2 = T(0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
        T t = fma(-shat, r.gb[j], srow[SR_AL + j]);
        chi2 = fma(t, t, chi2);
    }
    r.shat = shat; r.s = s; r.E = E; r.den = den; r.chi2 = chi2;
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
                                         const T* __restrict__ srow, T S, T c, T Q, T Tm, T (&e)[NB],
                                         T (&r)[NB], T& A, T& rho, T& gs, T& ell, T& delta) {
    // --- solve for Av (:176-204) ---
    T a = o.PA, b = T(0), ga = (o.Abar - A) * o.PA;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        T ru = r[j] * srow[SR_U + j];
        a = fma(ru, r[j], a);
        b += ru;
        ga = fma(ru, e[j], ga);
    }
    T dA = Num<T>::div_fast(S * ga - b * gs, S * a - b * b);
    dA = Num<T>::max(dA, o.avmin - A);
    dA = Num<T>::min(dA, o.avmax - A);
    A += dA;
    // --- solve for Rv (:206-237) ---
    const T gs2 = fma(-dA, b, gs);             // sum (e' - dA r) u
    T gr = T(0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
        e[j] = fma(-dA, r[j], e[j]);
        gr = fma(e[j] * srow[SR_U + j], m.D[j], gr);
    }
    gr = fma(gr, A, (o.Rbar - rho) * o.PR);
    const T q = fma(Q * A, A, o.PR);
    const T tt = Tm * A;
    T dR = Num<T>::div_fast(S * gr - tt * gs2, S * q - tt * tt);
    dR = Num<T>::max(dR, o.rvmin - rho);
    dR = Num<T>::min(dR, o.rvmax - rho);
    rho += dR;
    // --- update residuals / reddening vector, chi2 in magnitudes (:235-243) ---
    const T AdR = A * dR;
    gs = fma(-AdR, Tm, gs2);                   // sum (e' - A dR D) u
    T chi = T(0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
        e[j] = fma(-AdR, m.D[j], e[j]);
        r[j] = fma(dR, m.D[j], r[j]);
        chi = fma(e[j] * srow[SR_U + j], e[j], chi);
    }
    // logwt uses the un-centred residual e = e' + c (reference quirk)
    ell = T(-0.5) * (chi + c * (T(2) * gs + c * S));
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


// =================================================================================================
// Kernel 0: iteration-count probe.  Runs kProbeIter mag iterations on every `tile_stride`-th model
// tile and records, per star and iteration k, the two maxima the reference's stopping rule needs
// (:246-263).  The host turns them into the speculated iteration count of the full sweep, which
// verifies it on the whole grid (so a wrong guess costs a re-sweep, never a wrong answer).
// =================================================================================================
template <typename T, int NB>
__global__ void __launch_bounds__(kTile) k_kprobe(const ProbeParams<T> p) {
    using U = typename Enc<T>::U;
    __shared__ T s_star[kStarChunk][kStarStride];
    __shared__ U s_red[kStarChunk][2 * kProbeIter];
    const int first = blockIdx.y * kStarChunk;
    const int nst = min(kStarChunk, p.nstar - first);
    for (int t = threadIdx.x; t < nst * kStarStride; t += kTile)
        s_star[t / kStarStride][t % kStarStride] = p.stars[(int64_t)first * kStarStride + t];
    for (int t = threadIdx.x; t < kStarChunk * 2 * kProbeIter; t += kTile)
        s_red[t / (2 * kProbeIter)][t % (2 * kProbeIter)] = Enc<T>::enc(Num<T>::neg_inf());
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * p.tile_stride * kTile + threadIdx.x;
    const bool valid = i < p.nmodel;
    const DevOpts<T> o = p.o;
    ModelRegs<T, NB> m;
    load_model<T, NB>(p.grid, p.npad, valid ? i : 0, o, m);
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
        T e[NB], r[NB];
        T Q = T(0), Tm = T(0), gs = T(0);
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const T u = srow[SR_U + j];
            e[j] = srow[SR_CM + j] - m.cb[j];
            r[j] = m.r0[j];
            T Du = m.D[j] * u;
            Q = fma(Du, m.D[j], Q);
            Tm += Du;
            gs = fma(e[j], u, gs);
        }
#pragma unroll
        for (int k = 0; k < kProbeIter; k++) {
            T ell, delta;
            mag_iter<T, NB>(m, o, srow, S, c, Q, Tm, e, r, A, rho, gs, ell, delta);
            T l = valid ? ell : ninf;
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

// The whole magnitude-space fit of one (model, star) pair: initial residuals (brutus/fitting.py:728-733),
// `kspec` iterations of _optimize_fit_mag (:173-264), leaving the centred residuals in e.  (l0, b0) and
// (l1, b1) are the reduction inputs of iterations kspec-1 and kspec: logwt, and logwt where the step
// max(|dAv|, |dRv|) is still >= tol (else -inf).  Shared by the sweep and the candidate re-fit so that
// both evaluate bit-identical arithmetic.
template <typename T, int NB>
__device__ __forceinline__ void magfit_one(const ModelRegs<T, NB>& m, const DevOpts<T>& o,
                                           const T* __restrict__ srow, int kspec, T c, T (&e)[NB], T& A,
                                           T& rho, T& l0, T& b0, T& l1, T& b1) {
    const T ninf = Num<T>::neg_inf();
    const T S = srow[SR_SC + SC_S];
    A = o.Abar; rho = o.Rbar;
    T r[NB];
    T Q = T(0), Tm = T(0), gs = T(0);
    // brutus/fitting.py:158-164 (rp_den, srp_mix) and the initial residuals (:733)
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const T u = srow[SR_U + j];
        e[j] = srow[SR_CM + j] - m.cb[j];
        r[j] = m.r0[j];
        T Du = m.D[j] * u;
        Q = fma(Du, m.D[j], Q);
        Tm += Du;
        gs = fma(e[j], u, gs);
    }
    T ell = T(0), delta = T(0);
    l0 = ninf; b0 = ninf;
    if (kspec == 2) {   // the common case, fully unrolled
        mag_iter<T, NB>(m, o, srow, S, c, Q, Tm, e, r, A, rho, gs, ell, delta);
        l0 = ell;
        b0 = (delta >= o.mtol) ? l0 : ninf;
        mag_iter<T, NB>(m, o, srow, S, c, Q, Tm, e, r, A, rho, gs, ell, delta);
    } else {
        for (int k = 1; k <= kspec; k++) {
            mag_iter<T, NB>(m, o, srow, S, c, Q, Tm, e, r, A, rho, gs, ell, delta);
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
__global__ void __launch_bounds__(kTile) k_magfit(const SweepParams<T> p) {
    using U = typename Enc<T>::U;
    static_assert(kStarChunk == 32, "lane <-> star mapping of the reductions");
    __shared__ T s_star[kStarChunk][kStarStride];
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
        T e[NB];
        magfit_one<T, NB>(m, o, srow, s_kspec[s], c, e, A, rho, l0, b0, l1, b1);
        // --- _get_sed_mle at the fitted (Av, Rv) (:267) and the cull statistic (:745-756) ---
        Mle<T, NB> r4;
        mle_from_resid<T, NB>(e, c, srow, r4);
        T lp = cull_lnl(r4.chi2, r4.s, srow);
        // --- provisional lnlike / lnprob from the mag-fit values (final for every non-survivor) ---
        const int slot = s_slot[s];
        T ext = T(0);
        if (p.nlabel > 0 && valid) ext = ext_prior<T>(p.labels, p.ext + (int64_t)slot * p.nlabel * 3, p.nlabel, p.npad, i);
        T lnl0, lq;
        lnl_lnprob<T>(r4.chi2, r4.den * r4.E * r4.E, r4.s, false, srow, o.dim_prior, ext, lnl0, lq);
        l0 = valid ? l0 : ninf; b0 = valid ? b0 : ninf;
        l1 = valid ? l1 : ninf; b1 = valid ? b1 : ninf;
        const T lpv = valid ? lp : ninf;
        const T lqv = valid ? lq : ninf;
        l0 = warp_max_fast(l0); b0 = warp_max_fast(b0);
        l1 = warp_max_fast(l1); b1 = warp_max_fast(b1);
        const T lpm = warp_max_fast(lpv);
        const T lqm = warp_max_fast(lqv);
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

// Model coefficients of one model from the model-major copy of the grid (3-4 sectors per model).
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
    T sum = T(0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
        T mu = (T)v[j], R = (T)v[NB + j], D = (T)v[2 * NB + j];
        m.D[j] = D;
        m.r0[j] = fma(o.Rbar, D, R);
        m.cb[j] = fma(o.Abar, m.r0[j], mu);
        sum += m.cb[j];
    }
    m.bbar = sum * (T(1) / T(NB));
#pragma unroll
    for (int j = 0; j < NB; j++) m.cb[j] -= m.bbar;
}

// =================================================================================================
// Kernel 2: exact re-fit of the candidates.  One thread per candidate record (star, model): repeats the
// sweep's arithmetic for the pair, applies the exact cull test (:758-759) against the now final
// per-star maximum, initialises the record (stepsize 1, lnl_old = -1e300, :778-779) and appends the
// survivors to the flux work list.
// =================================================================================================
template <typename T, int NB>
__global__ void __launch_bounds__(kTile) k_refit(const RefitParams<T> p) {
    const int64_t q = (int64_t)blockIdx.x * kTile + threadIdx.x;
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
        T e[NB];
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
    __shared__ int s_w[kTile / 32];
    __shared__ int s_base;
    const unsigned bal = __ballot_sync(0xffffffffu, surv);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_w[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < kTile / 32; k++) n += s_w[k];
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
                                         const T* __restrict__ srow, T A, T rho, T (&e)[NB], T (&r)[NB]) {
    const T drho = rho - o.Rbar;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        r[j] = fma(drho, m.D[j], m.r0[j]);
        T red = fma(A, r[j], -o.Abar * m.r0[j]);
        e[j] = (srow[SR_CM + j] - m.cb[j]) - red;
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
__global__ void __launch_bounds__(kTile) k_flux(const FluxParams<T> p) {
    const int64_t t = (int64_t)blockIdx.x * kTile + threadIdx.x;
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
        T e[NB], r[NB];
        Mle<T, NB> r4;
        resid_at<T, NB>(m, o, srow, A, rho, e, r);
        mle_from_resid<T, NB>(e, c, srow, r4);
        T lnew = lold;
        for (int it = 0; it < p.nit; it++) {
            // one (dAv, dRv) step from the current model / residuals (:385-420)
            T an = T(0), ad = T(0), rn = T(0), rd = T(0);
#pragma unroll
            for (int j = 0; j < NB; j++) {
                T Ms = r4.shat * r4.gb[j];                 // M_j / sigma_j
                T tj = srow[SR_AL + j] - Ms;               // resid_j / sigma_j
                T rM = r[j] * Ms, DM = m.D[j] * Ms;
                an = fma(rM, tj, an);
                ad = fma(rM, rM, ad);
                rn = fma(DM, tj, rn);
                rd = fma(DM, DM, rd);
            }
            T dA = Num<T>::div(fma(T(kFac), an, (o.Abar - A) * o.PA), fma(T(kFac * kFac), ad, o.PA)) * eta;
            T dR = Num<T>::div(fma(T(kFac), rn, (o.Rbar - rho) * o.PR), fma(T(kFac * kFac), rd, o.PR)) * eta;
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
    T e[NB], r[NB];
    Mle<T, NB> r4;
    resid_at<T, NB>(m, o, srow, A, rho, e, r);
    mle_from_resid<T, NB>(e, c, srow, r4);
    // cross terms (:526-561) in sigma-normalised units; see the header comment and DESIGN.md
    T sa = T(0), sr = T(0), ar = T(0), aden = T(0), rden = T(0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
        T Ms = r4.shat * r4.gb[j];
        T tj = srow[SR_AL + j] - Ms;
        T h = Num<T>::exp2(T(kC2) * A * r[j]);     // F0_j / F_j = 10^(0.4 A r_j)   (:529-530)
        T mmr = Ms - tj;                           // (models - resid)/sigma         (:539-542)
        sa = fma(r[j] * r4.gb[j], mmr, sa);
        sr = fma(m.D[j] * r4.gb[j], mmr, sr);
        T DM = m.D[j] * Ms, rM = r[j] * Ms;
        ar = fma(DM, fma(Ms, T(1) - h, -tj), ar);  // drvecs (reddening - resid)/var (:550-551)
        aden = fma(rM, rM, aden);
        rden = fma(DM, DM, rden);
    }
    const O f = (O)kFac, E = (O)r4.E;
    const O ss = (O)r4.den * E * E;
    const O dsa = f * E * (O)sa, dsr = f * E * (O)sr;
    const O dar = f * (O)ar;
    const O daa = f * f * (O)aden + (O)o.PA + (O)(1. / (0.05 * 0.05));
    const O drr = f * f * (O)rden + (O)o.PR + (O)(1. / (0.1 * 0.1));
    if (modeA) {
        O* w = p.o_icov + t;
        w[0] = ss; w[p.ld] = dsa; w[2 * p.ld] = dsr; w[3 * p.ld] = daa; w[4 * p.ld] = dar; w[5 * p.ld] = drr;
    } else {
        O* w = p.o_icov + t * 9;
        w[0] = ss; w[1] = dsa; w[2] = dsr; w[3] = dsa; w[4] = daa; w[5] = dar; w[6] = dsr; w[7] = dar; w[8] = drr;
    }
}

// ---- launchers -------------------------------------------------------------------------------------
template <typename T, int NB> void launch_magfit(const SweepParams<T>& p, cudaStream_t st) {
    dim3 grid((unsigned)(p.npad / kTile), (unsigned)((p.nlist + kStarChunk - 1) / kStarChunk));
    k_magfit<T, NB><<<grid, kTile, 0, st>>>(p);
}
template <typename T, int NB> void launch_kprobe(const ProbeParams<T>& p, cudaStream_t st) {
    const int64_t ntile = p.npad / kTile;
    dim3 grid((unsigned)((ntile + p.tile_stride - 1) / p.tile_stride), (unsigned)((p.nstar + kStarChunk - 1) / kStarChunk));
    k_kprobe<T, NB><<<grid, kTile, 0, st>>>(p);
}
template <typename T, int NB> void launch_refit(const RefitParams<T>& p, cudaStream_t st) {
    if (p.ncand <= 0) return;
    k_refit<T, NB><<<(unsigned)((p.ncand + kTile - 1) / kTile), kTile, 0, st>>>(p);
}
template <typename T, int NB> void launch_flux(const FluxParams<T>& p, cudaStream_t st) {
    if (p.nsv <= 0) return;
    k_flux<T, NB><<<(unsigned)((p.nsv + kTile - 1) / kTile), kTile, 0, st>>>(p);
}
template <typename T, int NB, typename O> void launch_records(const RecordParams<T, O>& p, cudaStream_t st) {
    if (p.nrec <= 0) return;
    k_records<T, NB, O><<<(unsigned)((p.nrec + kTile - 1) / kTile), kTile, 0, st>>>(p);
}

}  // namespace bf
