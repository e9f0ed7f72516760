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
template <typename T, int NB> struct ModelRegs {
    static constexpr int NP = (NB + 1) / 2;
    P2<T> ncb[NP];  // -(b_j - bbar), b_j = mu_j + Abar r0_j  (model magnitudes at the prior-mean reddening)
    P2<T> r0[NP];   // R_j + Rbar D_j                         (brutus/utils.py:337-338 at rv = rv_gauss[0])
    P2<T> D[NP];    // dR/dRv
    T bbar;
    T A0, R0;       // expansion point of the model terms = where the magnitude fit starts: the prior means (Abar, Rbar)
                    // unless the caller of loglike supplied av_init / rv_init (brutus/fitting.py:700-703)
};

template <typename T, int NB>
__device__ __forceinline__ void finish_model(const T (&mu)[NB + 1], const T (&R)[NB + 1], const T (&D)[NB + 1],
                                             T A0, T R0, ModelRegs<T, NB>& m) {
    constexpr int NP = ModelRegs<T, NB>::NP;
    T r0[2 * NP], cb[2 * NP], Dd[2 * NP];
    T sum = T(0);
#pragma unroll
    for (int j = 0; j < 2 * NP; j++) {
        if (j < NB) {
            Dd[j] = D[j];
            r0[j] = fma(R0, D[j], R[j]);
            cb[j] = fma(A0, r0[j], mu[j]);
            sum += cb[j];
        } else {
            Dd[j] = r0[j] = cb[j] = T(0);   // padding band: zero model terms, zero star weights
        }
    }
    m.bbar = sum * (T(1) / T(NB));
    m.A0 = A0; m.R0 = R0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const T c0 = m.bbar - cb[2 * p];
        const T c1 = (2 * p + 1 < NB) ? m.bbar - cb[2 * p + 1] : T(0);
        m.ncb[p] = mk2(c0, c1);
        m.r0[p] = mk2(r0[2 * p], r0[2 * p + 1]);
        m.D[p] = mk2(Dd[2 * p], Dd[2 * p + 1]);
    }
}

// floats per model in the CTA's shared-memory tile of raw coefficients ([c][band], like `rows`): a multiple
// of 4 whose quarter is odd, so that the 16-byte accesses of 8 consecutive models hit 8 different bank groups
__host__ __device__ constexpr int tile_stride(int nb) { return (row_stride(nb) / 4) % 2 == 0 ? row_stride(nb) + 4 : row_stride(nb); }

// ModelRegs from a model's 3 NB raw coefficients v = [mu | R | D]
template <typename T, int NB>
__device__ __forceinline__ void model_from_coeffs(const float (&v)[row_stride(NB)], T A0, T R0, ModelRegs<T, NB>& m) {
    T mu[NB + 1], R[NB + 1], D[NB + 1];
#pragma unroll
    for (int j = 0; j < NB; j++) { mu[j] = (T)v[j]; R[j] = (T)v[NB + j]; D[j] = (T)v[2 * NB + j]; }
    finish_model<T, NB>(mu, R, D, A0, R0, m);
}

// a model's coefficients from the coefficient-major grid (fully coalesced: thread = model)
template <int NB>
__device__ __forceinline__ void load_coeffs(const float* __restrict__ grid, int64_t npad, int64_t i, float (&v)[row_stride(NB)]) {
#pragma unroll
    for (int k = 0; k < row_stride(NB); k++) v[k] = k < 3 * NB ? __ldg(grid + (int64_t)k * npad + i) : 0.f;
}

template <typename T, int NB>
__device__ __forceinline__ void load_model(const float* __restrict__ grid, int64_t npad, int64_t i,
                                           T A0, T R0, ModelRegs<T, NB>& m) {
    float v[row_stride(NB)];
    load_coeffs<NB>(grid, npad, i, v);
    model_from_coeffs<T, NB>(v, A0, R0, m);
}

// from a contiguous, 16-byte aligned row of row_stride(NB) floats: the model-major copy of the grid in HBM
// (3-4 sectors per model: the per-record gathers) or the CTA's tile in shared memory
template <typename T, int NB>
__device__ __forceinline__ void load_model_row(const float* __restrict__ row, T A0, T R0, ModelRegs<T, NB>& m) {
    constexpr int RS = row_stride(NB);
    float v[RS];
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int k = 0; k < RS / 4; k++) {
        float4 t = p4[k];
        v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
    model_from_coeffs<T, NB>(v, A0, R0, m);
}

// The expansion point of model i: the caller's per-model (av_init, rv_init) when the kernel is instantiated for them
// (INIT: bf_loglike_full after bf_set_init), else the prior means, which then stay compile-time aliases of o.Abar /
// o.Rbar (no extra registers on the throughput path).
template <typename T, bool INIT>
__device__ __forceinline__ void init_point(const DevOpts<T>& o, const T* __restrict__ av_init, const T* __restrict__ rv_init,
                                           int64_t i, T& A0, T& R0) {
    if (INIT) { A0 = av_init[i]; R0 = rv_init[i]; }
    else { A0 = o.Abar; R0 = o.Rbar; }
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
    A = m.A0; rho = m.R0;
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
template <typename T, int NB, bool INIT>
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
    {
        T A0, R0;
        init_point<T, INIT>(o, p.av_init, p.rv_init, i, A0, R0);
        load_model<T, NB>(p.grid, p.npad, i, A0, R0, m);
    }
    const int lane = threadIdx.x & 31;
    const T ninf = Num<T>::neg_inf();
    T acc[2 * kProbeIter];
#pragma unroll
    for (int k = 0; k < 2 * kProbeIter; k++) acc[k] = ninf;
    for (int s = 0; s < nst; s++) {
        const T* __restrict__ srow = s_star[s];
        const T S = srow[SR_SC + SC_S];
        const T c = srow[SR_SC + SC_MBAR] - m.bbar;
        T A = m.A0, rho = m.R0;
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

// residuals at an arbitrary (A, rho): e'_j = cm_j - cb_j - (A r_j - A0 r0_j), r_j = r0_j + (rho - R0) D_j
template <typename T, int NB>
__device__ __forceinline__ void resid_at(const ModelRegs<T, NB>& m, const DevOpts<T>& o,
                                         const T* __restrict__ srow, T A, T rho, P2<T> (&e)[(NB + 1) / 2],
                                         P2<T> (&r)[(NB + 1) / 2]) {
    constexpr int NP = (NB + 1) / 2;
    const P2<T> drho = bc2(rho - m.R0), A2 = bc2(A), nAbar = bc2(-m.A0);
#pragma unroll
    for (int p = 0; p < NP; p++) {
        r[p] = fma2(drho, m.D[p], m.r0[p]);
        const P2<T> red = fma2(A2, r[p], mul2(nAbar, m.r0[p]));
        e[p] = sub2(add2(ld2(srow + SR_CM + 2 * p), m.ncb[p]), red);
    }
}

// One iteration of the flux-space refinement of a survivor (brutus/fitting.py:784-803 with
// _optimize_fit_flux :385-420 and _get_sed_mle :423): a (dAv, dRv) step from the current model / residuals
// (r, r4), then the MLE at the new (A, rho).  Returns lnl_new = -chi2/2 (:792-795).  Shared by the sweep's
// dense phase and k_flux_more, so an iteration gives the same bits wherever it runs.
template <typename T, int NB>
__device__ __forceinline__ T flux_step(const ModelRegs<T, NB>& m, const DevOpts<T>& o, const T* __restrict__ srow,
                                       T c, T eta, T& A, T& rho, P2<T> (&e)[(NB + 1) / 2], P2<T> (&r)[(NB + 1) / 2],
                                       Mle<T, NB>& r4) {
    constexpr int NP = (NB + 1) / 2;
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
    return T(-0.5) * r4.chi2;
}

// Off-diagonal and reddening entries of icov_sar at the current fit (brutus/fitting.py:526-574), from the
// reddening vector r and the MLE r4 at (A, rho): ic = (sa, sr, aa, ar, rr); ss = r4.den E^2 is the caller's.
template <typename T, int NB>
__device__ __forceinline__ void icov_terms(const ModelRegs<T, NB>& m, const DevOpts<T>& o, const T* __restrict__ srow,
                                           T A, const P2<T> (&r)[(NB + 1) / 2], const Mle<T, NB>& r4, T (&ic)[5]) {
    constexpr int NP = (NB + 1) / 2;
    // cross terms in sigma-normalised units; see the header comment and DESIGN.md
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
    const T f = T(kFac), E = r4.E;
    ic[0] = f * E * hsum2(sa);
    ic[1] = f * E * hsum2(sr);
    ic[2] = f * f * hsum2(aden) + o.PA + T(1. / (0.05 * 0.05));
    ic[3] = f * hsum2(ar);
    ic[4] = f * f * hsum2(rden) + o.PR + T(1. / (0.1 * 0.1));
}

template <typename T>
__device__ __forceinline__ void store_fit(const PoolArrays<T>& pl, int64_t q, T A, T rho, T chi2, T s, T sden, const T (&ic)[5]) {
    pl.av[q] = A; pl.rv[q] = rho; pl.chi2[q] = chi2; pl.scale[q] = s; pl.sden[q] = sden;
    pl.isa[q] = ic[0]; pl.isr[q] = ic[1]; pl.iaa[q] = ic[2]; pl.iar[q] = ic[3]; pl.irr[q] = ic[4];
}

// =================================================================================================
// Kernel 1: the fused sweep.  grid = (model tiles, star chunks); each thread keeps its model in registers
// and loops over the chunk's stars, whose rows sit in shared memory (broadcast reads).
//
// MAIN LOOP, every (model, star): the magnitude-space fit (brutus/fitting.py:728-741 -> _optimize_fit_mag
// :34-271), _get_sed_mle at the fitted (Av, Rv) (:267), the cull statistic (:745-756) and a provisional
// lnprob.  Per star it produces six max-reductions (lane s of every warp keeps the warp's maxima for star s
// of the chunk, kStarChunk == 32, so the star loop contains no shared-memory atomics) and ONE BIT per pair:
// whether the pair can still matter, i.e. whether it may survive the cull (lnl_p > max + ln init_thresh) or
// pass lnpost's first selection (lnprob > max + ln wt_thresh).  Both tests are relative to per-star maxima
// that are only known after the sweep, so they are evaluated against a running maximum (warp-local maximum
// combined with the per-star global maximum published by the CTAs that already finished): a running maximum
// never exceeds the final one, hence the flagged set is a superset; the exact tests are applied to the
// flagged pairs' records afterwards (k_cull, k_sel).
//
// The number of mag iterations applied to every model of a star is a grid-wide decision in the reference
// (:246-263).  It is speculated (SI_KSPEC) and verified afterwards from two plain max-reductions per
// iteration:  "err < tol"  <=>  max{logwt_i : max(|dAv_i|,|dRv_i|) >= tol} <= max logwt + ln(init_thresh).
//
// DENSE PHASE, flagged pairs only (~15 % on the locus mock, clustered): a flagged lane pushes (star, lane,
// Av, Rv) into its WARP's ring buffer in shared memory; whenever 32 entries are queued the warp processes
// them with all 32 lanes busy: the entry's model comes from the CTA's shared-memory tile of coefficients,
// its star row from shared memory (no global gathers, no divergence).  For each entry: the flux MLE at the
// magnitude fit; for likely survivors of the cull the first `nit_first` iterations of the flux-space
// refinement (:778-803) -- the reference always runs at least two, and two is what most stars need; the
// precision matrix icov_sar (:526-574); and the complete record is appended to the candidate pool.  What
// used to be four passes of latency-bound gathers over the candidates (re-fit, flux loop, scatter, records)
// is arithmetic on data already on the SM.
// =================================================================================================
#ifndef BF_FLUSH_STARS
#define BF_FLUSH_STARS 4
#endif
constexpr int kFlushStars = BF_FLUSH_STARS;        // the CTA meets (one barrier) every kFlushStars stars to drain its queue
constexpr int kQueue = kTile * (kFlushStars + 1);  // ring-buffer entries per CTA: < kTile left over + kTile per star pushed

template <typename T, int NB> struct SweepSmem {
    static constexpr int RS = tile_stride(NB);
    static constexpr size_t off_star = 0;
    static constexpr size_t off_tile = off_star + sizeof(T) * kStarChunk * kStarSmem;
    static constexpr size_t off_qav = off_tile + sizeof(float) * kTile * RS;
    static constexpr size_t off_qrv = off_qav + sizeof(T) * kQueue;
    static constexpr size_t off_qlp = off_qrv + sizeof(T) * kQueue;
    static constexpr size_t off_qkey = off_qlp + sizeof(T) * kQueue;
    static constexpr size_t off_red = off_qkey + sizeof(uint32_t) * kQueue;   // [warp][star][kSweepRed] of T
    static constexpr size_t off_snap = off_red + sizeof(T) * (kTile / 32) * kStarChunk * kSweepRed;
    static constexpr size_t off_int = off_snap + sizeof(T) * kStarChunk * 2;
    static constexpr size_t off_bal = off_int + sizeof(int) * (kStarChunk * 3 + 4 + 2 * (kTile / 32) + 4);   // + queue tail, head, warp counts, tile
    static constexpr size_t bytes = off_bal + sizeof(uint32_t) * kStarChunk * (kTile / 32);             // candidate words [star][warp]
};

// The dense phase for entries [head, head + n) of the CTA's ring buffer (n <= kTile): one entry per thread, the
// records appended to the pool as one contiguous block.  All threads of the CTA must call it (it contains
// barriers); warps whose 32 threads are all beyond n only keep the barriers company.
template <typename T, int NB, bool INIT>
__device__ __forceinline__ void dense_flush(const SweepParams<T>& p, const DevOpts<T>& o, const T* __restrict__ s_star,
                                            const float* __restrict__ s_tile, const int* __restrict__ s_tag,
                                            const T* __restrict__ q_av, const T* __restrict__ q_rv,
                                            const T* __restrict__ q_lp, const uint32_t* __restrict__ q_key, int* s_head,
                                            int n, unsigned long long* s_base, const int* s_tileid) {
    constexpr int NP = (NB + 1) / 2;
    constexpr int RS = tile_stride(NB);
    const bool act = (int)threadIdx.x < n;
    const int head = *s_head;                                   // ring position of the first entry (< kQueue)
    int qi = head + (act ? (int)threadIdx.x : 0);               // idle threads shadow entry 0 (results discarded)
    if (qi >= kQueue) qi -= kQueue;
    const uint32_t key = q_key[qi];
    T A = q_av[qi], rho = q_rv[qi];
    // the cull statistic exactly as the main loop computed it: the exact cull test (k_cull) compares it with the
    // per-star maximum of the same quantity, so survival never hinges on the rounding of a second evaluation
    const T lp = q_lp[qi];
    if (threadIdx.x == 0) *s_base = atomicAdd(p.pool_count, (unsigned long long)n);   // one atomic per flush
    __syncthreads();   // every thread has read the head and its entry (the ring slots may be overwritten); s_base is visible
    if (threadIdx.x == 0) {
        const int h = head + n;
        *s_head = h >= kQueue ? h - kQueue : h;
    }
    if ((int)(threadIdx.x & ~31u) < n) {
    const int64_t q = (int64_t)*s_base + threadIdx.x;
    const int s = key & 31, ml = (key >> 5) & 255;
    const bool fluxed = ((key >> 13) & 1u) != 0u && p.nit_first > 0;
    const PoolArrays<T>& pl = p.pool;
    const bool wr = act && q < p.pool_cap;
    if (wr) {   // what is known already goes out first (coalesced stores of the warp's records): fewer live registers below
        pl.model[q] = *s_tileid * kTile + ml;
        pl.sflag[q] = s_tag[s] | (fluxed ? kFlagFluxed << 24 : 0);
        pl.lp[q] = lp;
        pl.lnl[q] = A; pl.lnprob[q] = rho;   // the magnitude fit, kept for k_fixup until k_final overwrites it
    }
    const T* __restrict__ srow = s_star + s * kStarSmem;
    ModelRegs<T, NB> m;
    {
        T A0, R0;
        init_point<T, INIT>(o, p.av_init, p.rv_init, (int64_t)*s_tileid * kTile + ml, A0, R0);
        load_model_row<T, NB>(s_tile + ml * RS, A0, R0, m);
    }
    const T c = srow[SR_SC + SC_MBAR] - m.bbar;
    P2<T> e[NP], r[NP];
    Mle<T, NB> r4;
    resid_at<T, NB>(m, o, srow, A, rho, e, r);
    mle_from_resid<T, NB>(e, c, srow, r4);
    T eta = T(1), lold = Num<T>::kNegBig, lprev = Num<T>::kNegBig;   // stepsize 1, lnl_old = -1e300 (:778-779)
    if (fluxed) {
        for (int it = 0; it < p.nit_first; it++) {
            const T lnew = flux_step<T, NB>(m, o, srow, c, eta, A, rho, e, r, r4);
            lprev = lold;
            if (lnew < lold) eta = eta / T(1.2);            // :802
            lold = lnew;                                    // :803
        }
    }
    T ic[5];
    icov_terms<T, NB>(m, o, srow, A, r, r4, ic);
    if (wr) {
        store_fit<T>(pl, q, A, rho, r4.chi2, r4.s, r4.den * r4.E * r4.E, ic);
        pl.eta[q] = eta; pl.lold[q] = lold; pl.lprev[q] = lprev;
    }
    }
    __syncthreads();   // s_base may be rewritten by the next flush; the new head is visible
}

// up to 8 bands: capped at 80 registers (3 CTAs = 24 warps per SM); left to itself ptxas takes more and the
// kernel loses a third of its warps (measured 15.8 -> 16.4 ms per 1 000 stars in round 1).  The float64
// instantiation (verification) takes what it needs.
template <typename T, int NB> constexpr int sweep_min_ctas() { return sizeof(T) == 8 ? 1 : (NB <= 8 ? 3 : 2); }

template <typename T, int NB, bool INIT>
__global__ void __launch_bounds__(kTile, (sweep_min_ctas<T, NB>())) k_sweep(const SweepParams<T> p) {
    using U = typename Enc<T>::U;
    using SM = SweepSmem<T, NB>;
    constexpr int NP = (NB + 1) / 2;
    constexpr int RS = SM::RS;
    static_assert(kStarChunk == 32, "lane <-> star mapping of the reductions");
    extern __shared__ __align__(16) unsigned char smem[];
    T* s_star = reinterpret_cast<T*>(smem + SM::off_star);                 // [32][kStarSmem]
    float* s_tile = reinterpret_cast<float*>(smem + SM::off_tile);         // [256][RS]
    T* s_qav = reinterpret_cast<T*>(smem + SM::off_qav);                   // [kQueue]: the CTA's candidate queue
    T* s_qrv = reinterpret_cast<T*>(smem + SM::off_qrv);
    T* s_qlp = reinterpret_cast<T*>(smem + SM::off_qlp);
    uint32_t* s_qkey = reinterpret_cast<uint32_t*>(smem + SM::off_qkey);
    T* s_red = reinterpret_cast<T*>(smem + SM::off_red);                   // [32][8][kSweepRed]: per-star, per-warp maxima
    T* s_snap = reinterpret_cast<T*>(smem + SM::off_snap);                 // [32][2]
    int* s_slot = reinterpret_cast<int*>(smem + SM::off_int);              // [32]
    int* s_kspec = s_slot + kStarChunk;
    int* s_tag = s_kspec + kStarChunk;
    int* s_qtot = s_tag + kStarChunk;                                      // entries pushed so far
    int* s_head = s_qtot + 1;                                              // ring position of the oldest queued entry
    unsigned long long* s_base = reinterpret_cast<unsigned long long*>(s_qtot + 2);   // pool position of the flush
    int* s_wc = s_qtot + 4;                                                // [2][8]: candidates per warp since the last meeting
    int* s_tileid = s_wc + 2 * (kTile / 32);                               // the CTA's model tile (kept out of the registers)
    uint32_t* s_bal = reinterpret_cast<uint32_t*>(smem + SM::off_bal);     // [32][8]: the candidate map words of this CTA

    const int first = blockIdx.y * kStarChunk;
    const int nst = min(kStarChunk, p.nlist - first);
    for (int t = threadIdx.x; t < nst * kStarStride; t += kTile) {
        int s = t / kStarStride, k = t - s * kStarStride;
        s_star[s * kStarSmem + k] = p.stars[(int64_t)p.list[first + s] * kStarStride + k];
    }
    if (threadIdx.x == 0) { *s_qtot = 0; *s_head = 0; }
    if (threadIdx.x < 2 * (kTile / 32)) s_wc[threadIdx.x] = 0;
    for (int t = threadIdx.x; t < nst; t += kTile) {
        int slot = p.list[first + t];
        s_slot[t] = slot;
        s_kspec[t] = p.star_int[slot * SI_COUNT + SI_KSPEC];
        s_tag[t] = make_tag(slot, p.star_int[slot * SI_COUNT + SI_EPOCH], 0);
        // running per-star maxima published so far (benign race: any value <= the final maximum is valid)
        const volatile U* rr = p.red + (int64_t)slot * kNumRed;
        s_snap[2 * t] = Enc<T>::dec(rr[RED_LP]);
        s_snap[2 * t + 1] = Enc<T>::dec(rr[RED_M0]);
        // a maxima-only launch flags nothing: thresholds of +inf (no test in the star loop)
        if (p.maxima_only) s_snap[2 * t] = s_snap[2 * t + 1] = -Num<T>::neg_inf();
    }

    // The model tile of this CTA.  A sweep is two launches: first the tiles that are multiples of tile_S -- a strided
    // subsample of the grid -- then the rest, whose CTAs all start from the per-star maxima the first launch published
    // (s_snap above): the running thresholds are tight from the first CTA on, whatever the order of the grid.  Small
    // grids (a CTA wave covers every tile of a star chunk, so nothing is published in time) are instead swept twice,
    // first in maxima-only mode (tile_mode 0 both times).
    const int tile = p.tile_mode == 0 ? (int)blockIdx.x
                   : p.tile_mode == 1 ? (int)blockIdx.x * p.tile_S
                                      : (int)blockIdx.x + (int)blockIdx.x / (p.tile_S - 1) + 1;
    if (threadIdx.x == 0) *s_tileid = tile;                 // read back where it is needed again (dense phase, map)
    const int64_t i = (int64_t)tile * kTile + threadIdx.x;  // npad is a multiple of kTile
    // padding models exist in the last tile only; a CTA-uniform count keeps the per-star test to one compare
    const int nvalid = (int)(p.nmodel - (int64_t)tile * kTile < (int64_t)kTile ? p.nmodel - (int64_t)tile * kTile : (int64_t)kTile);
    // the bound is re-read from shared memory in the star loop (one LDS + compare): held as a predicate across the
    // loop it is spilled through a register to local memory
    if (threadIdx.x == 0) s_tileid[1] = nvalid;
    const DevOpts<T> o = p.o;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    float* tile_w = s_tile + wrp * 32 * RS;                         // this warp's 32 models
    {   // coefficients: coalesced from the coefficient-major grid into the tile (conflict-free 16-byte stores)
        float v[row_stride(NB)];
        load_coeffs<NB>(p.grid, p.npad, i, v);
        float4* dst = reinterpret_cast<float4*>(tile_w + lane * RS);
#pragma unroll
        for (int k = 0; k < row_stride(NB) / 4; k++) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
    __syncthreads();
    T A0, R0;
    init_point<T, INIT>(o, p.av_init, p.rv_init, i, A0, R0);   // i < npad: the init arrays are padded like the grid
    ModelRegs<T, NB> m;
    load_model_row<T, NB>(tile_w + lane * RS, A0, R0, m);
    const T ninf = Num<T>::neg_inf();
    const T ln_init_c = o.ln_init - T(kCandMargin);
    T* w_red = s_red + wrp * kSweepRed;                            // the warp's maxima for star s: [s][wrp][kSweepRed]
    int pend = 0;                                                   // queued entries not yet processed (CTA-uniform)

    for (int s = 0; s < nst; s++) {
        const T* __restrict__ srow = s_star + s * kStarSmem;
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
        {   // warp-uniform values to a warp-uniform address: every lane stores (no branch, no lane test)
            T* w = w_red + s * (kTile / 32) * kSweepRed;
            w[0] = l0; w[1] = b0; w[2] = l1; w[3] = b1; w[4] = lpm; w[5] = lqm;
        }
        // --- candidate bit, queue ---
        const T thr1 = Num<T>::max(s_snap[2 * s], lpm) + ln_init_c;
        const T thr2 = Num<T>::max(s_snap[2 * s + 1], lqm) + o.ln_wt - srow[SR_SC + SC_SLACK];
        const bool likely = lp > thr1;                              // may survive the cull
        const bool cand = (int)threadIdx.x < reinterpret_cast<const volatile int*>(s_tileid)[1] && (likely || lq > thr2);
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        s_bal[s * (kTile / 32) + wrp] = bal;                        // written to the map after the star loop, a sector per star
        const int par = (s / kFlushStars) & 1;                      // parity of the warp-count buffer in use
        if (bal) {   // push the warp's candidates of this star into the CTA's queue: one shared-memory atomic per warp
            int base = 0;
            if (lane == 0) {   // inline PTX: the compiler would wrap a one-lane atomicAdd in its own warp aggregation
                const int n = __popc(bal);
                // the address is made to depend on a lane id ptxas cannot tie to the branch (it is 0 here): with a
                // provably uniform address ptxas wraps each atomic in a warp aggregation (vote, find-leader, popc,
                // shuffle: ~12 instructions) that a single active lane does not need
                uint32_t lid;
                asm volatile("mov.u32 %0, %%laneid;" : "=r"(lid));
                const uint32_t a_tot = (uint32_t)__cvta_generic_to_shared(s_qtot) + (lid << 2);
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"(a_tot), "r"(n) : "memory");
                if (base < kQueue && base + n >= kQueue)                      // the push that wraps the ring
                    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a_tot), "r"(-kQueue) : "memory");
                if (kFlushStars == 1) s_wc[par * (kTile / 32) + wrp] = n;
                else asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a_tot + 16u + 4u * (par * (kTile / 32) + wrp)), "r"(n) : "memory");
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cand) {
                unsigned lt_mask;
                asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
                int pos = base + __popc(bal & lt_mask);
                if (pos >= kQueue) pos -= kQueue;
                s_qav[pos] = A; s_qrv[pos] = rho; s_qlp[pos] = lp;
                s_qkey[pos] = (uint32_t)s | (uint32_t)threadIdx.x << 5 | (likely ? 1u << 13 : 0u);
            }
        } else if (kFlushStars == 1 && lane == 0) s_wc[par * (kTile / 32) + wrp] = 0;
        // The CTA meets every kFlushStars stars: the warps' candidates of those stars are then adjacent in the queue
        // (and later in the pool: runs of a star's records several times longer than a warp alone would give, which is
        // what the pool passes and the ordered gather want), and every full kTile of queued entries is run through the
        // dense phase by all 256 threads, one entry each.  The queue fill is derived from per-warp counts (double
        // buffered), so every thread sees the same value whatever the faster warps push next.
        if (((s + 1) & (kFlushStars - 1)) == 0 || s == nst - 1) {
            __syncthreads();
            const int4 c0 = *reinterpret_cast<const int4*>(s_wc + par * (kTile / 32));
            const int4 c1 = *reinterpret_cast<const int4*>(s_wc + par * (kTile / 32) + 4);
            pend += c0.x + c0.y + c0.z + c0.w + c1.x + c1.y + c1.z + c1.w;
            // the other buffer was read by everybody before this barrier: clear this warp's slot for its next use
            if (kFlushStars > 1 && lane == 0) s_wc[(par ^ 1) * (kTile / 32) + wrp] = 0;
            if (pend >= kTile) {
                do {
                    dense_flush<T, NB, INIT>(p, o, s_star, s_tile, s_tag, s_qav, s_qrv, s_qlp, s_qkey, s_head, kTile, s_base, s_tileid);
                    pend -= kTile;
                } while (pend >= kTile);
                load_model_row<T, NB>(tile_w + lane * RS, A0, R0, m);   // this thread's own model again
            }
        }
    }
    if (pend > 0) dense_flush<T, NB, INIT>(p, o, s_star, s_tile, s_tag, s_qav, s_qrv, s_qlp, s_qkey, s_head, pend, s_base, s_tileid);
    __syncthreads();
    // the candidate map: the CTA's 8 words of a star are one 32-byte sector
    if (!p.maxima_only)
        for (int t = threadIdx.x; t < nst * (kTile / 32); t += kTile)
            p.cand[(int64_t)s_slot[t / (kTile / 32)] * p.nwords + (int64_t)*s_tileid * (kTile / 32) + t % (kTile / 32)] = s_bal[t];
    // combine the warps' maxima and publish them; NaN maxima (every lane NaN) must not poison the
    // unsigned-encoded atomics
    for (int t = threadIdx.x; t < nst * kSweepRed; t += kTile) {
        T v = ninf;
        bool any = false;
#pragma unroll
        for (int w = 0; w < kTile / 32; w++) {
            const T x = s_red[((t / kSweepRed) * (kTile / 32) + w) * kSweepRed + t % kSweepRed];
            if (x == x) { v = any ? Num<T>::max(v, x) : x; any = true; }
        }
        const int s = t / kSweepRed, k = t % kSweepRed;
        const int map[kSweepRed] = {RED_L0, RED_B0, RED_L1, RED_B1, RED_LP, RED_M0};
        if (any) atomicMax(&p.red[(int64_t)s_slot[s] * kNumRed + map[k]], Enc<T>::enc(v));
    }
}

// =================================================================================================
// Kernel 2 (rare): records the sweep refined as likely survivors that the exact cull test (k_cull, against
// the final per-star maximum) rejected -- pairs flagged while the running maximum was still low.  A
// non-survivor keeps its magnitude-fit values (brutus/fitting.py:805-810 scatters survivors only): redo the
// MLE and icov at the magnitude fit kept in the record.
// =================================================================================================
template <typename T, int NB, bool INIT>
__global__ void __launch_bounds__(kTile) k_fixup(const RecParams<T> p) {
    constexpr int NP = (NB + 1) / 2;
    const int64_t t = (int64_t)blockIdx.x * kTile + threadIdx.x;
    if (t >= p.n) return;
    const int64_t q = p.list[t];
    const PoolArrays<T>& pl = p.pool;
    const int tag = pl.sflag[q];
    const int slot = tag_slot(tag);
    const DevOpts<T> o = p.o;
    const T* __restrict__ srow = p.stars + (int64_t)slot * kStarStride;
    ModelRegs<T, NB> m;
    {
        const int64_t i = pl.model[q];
        T A0, R0;
        init_point<T, INIT>(o, p.av_init, p.rv_init, i, A0, R0);
        load_model_row<T, NB>(p.rows + i * row_stride(NB), A0, R0, m);
    }
    const T c = srow[SR_SC + SC_MBAR] - m.bbar;
    const T A = pl.lnl[q], rho = pl.lnprob[q];
    P2<T> e[NP], r[NP];
    Mle<T, NB> r4;
    resid_at<T, NB>(m, o, srow, A, rho, e, r);
    mle_from_resid<T, NB>(e, c, srow, r4);
    T ic[5];
    icov_terms<T, NB>(m, o, srow, A, r, r4, ic);
    store_fit<T>(pl, q, A, rho, r4.chi2, r4.s, r4.den * r4.E * r4.E, ic);
    pl.sflag[q] = tag & ~(kFlagFluxed << 24);
}

// =================================================================================================
// Kernel 3 (rare): one more flux iteration for the survivors of the stars whose loop has not converged after
// the iterations the sweep ran (SI_ACTIVE, decided on the device by k_flux_ctl), with the convergence
// reductions of that iteration:
//   "lerr <= ltol"  <=>  max{lnl_new_i : |lnl_new_i - lnl_old_i| > ltol} <= max lnl_new + ln(ltol_subthresh)
// It visits the list of those survivors that k_flux_list (api.cu) extracted from the pool.
// =================================================================================================
template <typename T, int NB, bool INIT>
__global__ void __launch_bounds__(kTile) k_flux_more(const RecParams<T> p) {
    constexpr int NP = (NB + 1) / 2;
    __shared__ StarAgg<T, 2> agg;
    const int which[2] = {RED_FL, RED_FB};
    agg.init();
    __syncthreads();
    const PoolArrays<T>& pl = p.pool;
    const int64_t n = (int64_t)*p.nlist;
    for (int64_t base = (int64_t)blockIdx.x * kTile; base < n; base += (int64_t)gridDim.x * kTile) {
        const int64_t t = base + threadIdx.x;
        int64_t q = -1;
        int slot = -1;
        bool act = false;
        if (t < n) {
            q = p.list[t];
            const int tag = pl.sflag[q];
            const int sl = tag_slot(tag);
            if (p.star_int[sl * SI_COUNT + SI_ACTIVE] != 0) { act = true; slot = sl; }
        }
        T v[2] = {Num<T>::neg_inf(), Num<T>::neg_inf()};
        if (act) {
            const DevOpts<T> o = p.o;
            const T* __restrict__ srow = p.stars + (int64_t)slot * kStarStride;
            ModelRegs<T, NB> m;
            {
                const int64_t i = pl.model[q];
                T A0, R0;
                init_point<T, INIT>(o, p.av_init, p.rv_init, i, A0, R0);
                load_model_row<T, NB>(p.rows + i * row_stride(NB), A0, R0, m);
            }
            const T c = srow[SR_SC + SC_MBAR] - m.bbar;
            T A = pl.av[q], rho = pl.rv[q], eta = pl.eta[q], lold = pl.lold[q];
            P2<T> e[NP], r[NP];
            Mle<T, NB> r4;
            resid_at<T, NB>(m, o, srow, A, rho, e, r);
            mle_from_resid<T, NB>(e, c, srow, r4);
            const T lnew = flux_step<T, NB>(m, o, srow, c, eta, A, rho, e, r, r4);
            v[0] = (lnew == lnew) ? lnew : Num<T>::neg_inf();
            v[1] = (tabs(lnew - lold) > o.ltol) ? v[0] : Num<T>::neg_inf();
            pl.lprev[q] = lold;
            if (lnew < lold) eta = eta / T(1.2);            // :802
            pl.eta[q] = eta;
            pl.lold[q] = lnew;                              // :803
            T ic[5];
            icov_terms<T, NB>(m, o, srow, A, r, r4, ic);
            store_fit<T>(pl, q, A, rho, r4.chi2, r4.s, r4.den * r4.E * r4.E, ic);
        }
        if (__any_sync(0xffffffffu, act)) agg.add(p.red, which, nullptr, 0, slot, act, v, false);
    }
    agg.flush(p.red, which, nullptr, 0);
}

// ---- launchers -------------------------------------------------------------------------------------
template <typename T, int NB> void launch_kprobe(const ProbeParams<T>& p, cudaStream_t st) {
    const int64_t ntile = p.npad / kTile;
    dim3 grid((unsigned)((ntile + p.tile_stride - 1) / p.tile_stride), (unsigned)((p.nstar + kStarChunk - 1) / kStarChunk));
    if (p.av_init) k_kprobe<T, NB, true><<<grid, kTile, 0, st>>>(p);
    else k_kprobe<T, NB, false><<<grid, kTile, 0, st>>>(p);
}
template <typename T, int NB, bool INIT> int launch_sweep_(const SweepParams<T>& p, cudaStream_t st) {
    static bool configured[64] = {};   // per device: dynamic shared memory above 48 KB needs an opt-in
    int dev = 0;
    cudaGetDevice(&dev);
    constexpr size_t bytes = SweepSmem<T, NB>::bytes;
    if (dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_sweep<T, NB, INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    const int64_t ntile = p.npad / kTile, nsub = p.tile_mode == 0 ? ntile : (ntile + p.tile_S - 1) / p.tile_S;
    const int64_t nx = p.tile_mode == 2 ? ntile - nsub : nsub;
    if (nx <= 0) return (int)cudaSuccess;
    dim3 grid((unsigned)nx, (unsigned)((p.nlist + kStarChunk - 1) / kStarChunk));
    k_sweep<T, NB, INIT><<<grid, kTile, bytes, st>>>(p);
    return (int)cudaSuccess;
}
template <typename T, int NB> int launch_sweep(const SweepParams<T>& p, cudaStream_t st) {
    return p.av_init ? launch_sweep_<T, NB, true>(p, st) : launch_sweep_<T, NB, false>(p, st);
}
template <typename T, int NB> void launch_fixup(const RecParams<T>& p, cudaStream_t st) {
    if (p.n <= 0) return;
    if (p.av_init) k_fixup<T, NB, true><<<(unsigned)((p.n + kTile - 1) / kTile), kTile, 0, st>>>(p);
    else k_fixup<T, NB, false><<<(unsigned)((p.n + kTile - 1) / kTile), kTile, 0, st>>>(p);
}
template <typename T, int NB> void launch_flux_more(const RecParams<T>& p, cudaStream_t st) {
    if (p.n <= 0) return;
    const int64_t ctas = (p.n + kTile - 1) / kTile;   // p.n: upper bound of the list length, the kernel reads *nlist
    if (p.av_init) k_flux_more<T, NB, true><<<(unsigned)(ctas < kPassCtas ? ctas : kPassCtas), kTile, 0, st>>>(p);
    else k_flux_more<T, NB, false><<<(unsigned)(ctas < kPassCtas ? ctas : kPassCtas), kTile, 0, st>>>(p);
}

}  // namespace bf
