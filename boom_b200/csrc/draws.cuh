// draws.cuh -- per-observation latent draws of the auxiliary-mixture samplers (device side).
//
// What the reference does per observation, and where:
//   logit, small sample  BinomialLogitCltDataImputer::impute_small_sample
//                        (Models/Glm/PosteriorSamplers/BinomialLogitDataImputer.cpp:134-152):
//                        rtrun_logit_mt (distributions/trun_logit.cpp:163-174) then
//                        NormalMixtureApproximation::unmix (NormalMixtureApproximation.cpp:280-290)
//   logit, CLT           impute_large_sample (BinomialLogitDataImputer.cpp:155-210)
//   Poisson              PoissonDataImputer::impute (PoissonDataImputer.cpp:36-98) and
//                        unmix_poisson_augmented_data (poisson_mixture_approximation_table.cpp:44-61)
// The random stream is Philox4x32-10 keyed by (seed, iteration, GLOBAL row, slot), so the
// draws are identical for any sharding of the rows over GPUs.
#pragma once
#include <cstdint>

namespace boomgpu {

constexpr int kMaxLogitK = 16;
constexpr double kLnSqrt2Pi = 0.918938533204672741780329736406;
constexpr double kTwoPi = 6.283185307179586476925286766559;

// ---- Philox4x32-10 ------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}

// (k + 1/2) 2^-52 for the 52-bit integer k = hi : top 20 bits of lo -- exact, strictly inside (0, 1).  The 52 bits are dropped
// into the mantissa of a double in [1, 2) and 1 - 2^-53 is subtracted: one exact DADD instead of a 64-bit integer -> double
// conversion, an add and a multiply (same value, bit for bit, as the oracle's ((double)k + 0.5) * 2^-52).
__device__ __forceinline__ double bits_to_open01(uint32_t hi, uint32_t lo) {
  const double d = __hiloint2double((int)(0x3ff00000u | (hi >> 12)), (int)__funnelshift_r(lo, hi, 12));
  return d - 0x1.fffffffffffffp-1;
}

// seed / iteration key the stream; rk holds the ten Philox round keys (k0 + r 0x9E3779B9, k1 + r 0xBB67AE85) so that the
// kernels read them as constants instead of re-deriving them per observation
struct RngKey { uint64_t seed, iteration; uint32_t rk[20]; };

__host__ __device__ inline void philox_round_keys(RngKey &key) {
  uint32_t k0 = (uint32_t)key.seed, k1 = (uint32_t)(key.seed >> 32);
  for (int r = 0; r < 10; ++r) { key.rk[2 * r] = k0; key.rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

__device__ __forceinline__ void uniform_pair(const RngKey &key, uint64_t row, uint32_t slot, double &u0, double &u1) {
  U4 c;
  c.x = (uint32_t)row; c.y = (uint32_t)(row >> 32);
  c.z = (uint32_t)key.iteration;
  c.w = (((uint32_t)(key.iteration >> 32) & 0xFFFFu) << 16) | (slot & 0xFFFFu);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ key.rk[2 * r]; n.y = lo1; n.z = hi0 ^ c.w ^ key.rk[2 * r + 1]; n.w = lo0;
    c = n;
  }
  u0 = bits_to_open01(c.x, c.y);
  u1 = bits_to_open01(c.z, c.w);
}

// ---- mixtures -------------------------------------------------------------------------
// Full-precision mixture (global memory, owned by the context): the FP64 statement of unmix and the CLT branch read it.
struct LogitMixture {
  int K;
  double mu[kMaxLogitK];
  double sigma[kMaxLogitK];
  double inv_sigma[kMaxLogitK];
  double weights[kMaxLogitK];
  double lconst[kMaxLogitK];     // log w - ln sqrt(2 pi) - log sigma
  double inv_sigsq[kMaxLogitK];  // 1 / (sigma * sigma)
};

// What the per-trial hot loop needs, passed by value as a kernel parameter (constant bank: every lane reads the
// same entry): single-precision log2-density coefficients for the certified selection below, and 1/sigma^2.
constexpr float kLog2e = 1.4426950408889634f;
struct LogitHot {
  int K;
  float lconst2[kMaxLogitK];     // log2(e) * lconst
  float hs2[kMaxLogitK];         // -0.5 * log2(e) / sigma^2
  float mu_c[kMaxLogitK];        // mu[k] - center
  double center;                 // = mu[0]: residuals are centred in FP64 before they are rounded to FP32
  double inv_sigsq[kMaxLogitK];
  double mu_d[kMaxLogitK];       // FP64 component means and log(1 / sigma^2): what the Poisson statistics add per draw
  double logw[kMaxLogitK];
  // scale mixtures (every component mean equal -- the logit table): lconst2 shifted so that its maximum is 0, and the
  // index of the widest component.  exp2(hs2 d^2 + l0) then cannot overflow and needs no per-draw maximum.
  int zero_mean, wide;
  float l0[kMaxLogitK];
};

// Poisson table in global memory (per-row nu makes the lookups divergent).
struct PoissonTable {
  int ntab;
  const int64_t *nu;
  const int32_t *offset;
  const double *mu;
  const double *inv_sigma;
  const double *lconst;
  const double *sigma;
  const float *mu_f;       // single-precision copies for the certified selection: mu[k] - mu[first component of the entry]
  const float *lconst2_f;
  const float *hs2_f;
  const double *inv_sigsq;  // 1 / sigma^2 and its log, per component
  const double *logw;
  const int32_t *dense;     // dense[nu] = entry index (or -1) for nu < dense_n: no search for the common small counts
  int dense_n;
  int64_t gaussian_cutoff;
  int e1;  // entry index of nu == 1 (every row uses it)
};

// The table entries of the small counts (nu < 32: K = 10 or 4 components each), staged in SHARED memory by kernels that
// have room for them.  Per-row nu makes the lookups divergent; in global memory every draw walks a chain of dependent
// loads (dense index -> offsets -> components -> selected component): ~16 k cycles per 32-row slice at C2.
constexpr int kPoisSmemNu = 32, kPoisSmemComps = 320;
struct PoissonSmem {
  int off[kPoisSmemNu + 1];              // components of nu: off[nu] .. off[nu + 1] - 1 (empty: nu not in the table)
  float mu_f[kPoisSmemComps], lconst2_f[kPoisSmemComps], hs2_f[kPoisSmemComps];
  double center[kPoisSmemNu];            // mu of the entry's first component (the FP32 residual is centred on it)
  double mu[kPoisSmemComps], inv_sigsq[kPoisSmemComps], logw[kPoisSmemComps];
};

// NormalMixtureApproximation::unmix given its uniform.  The reference normalises the
// probabilities before rmulti_mt draws v ~ U(0, sum); selecting on the unnormalised
// cumulative sums with v = U * sum is the same event.
template <class F>
__device__ __forceinline__ int unmix_generic(int K, double unif, F lp_of) {
  double lp[kMaxLogitK];
  double mx = -1e300;
#pragma unroll
  for (int s = 0; s < kMaxLogitK; ++s) {
    if (s < K) { lp[s] = lp_of(s); mx = fmax(mx, lp[s]); }
  }
  double tot = 0;
#pragma unroll
  for (int s = 0; s < kMaxLogitK; ++s) {
    if (s < K) { lp[s] = exp(lp[s] - mx); tot += lp[s]; }
  }
  double v = unif * tot, cs = 0;
  int k = K - 1;
  bool found = false;
#pragma unroll
  for (int s = 0; s < kMaxLogitK; ++s) {
    if (s < K) {
      cs += lp[s];
      if (!found && v <= cs) { k = s; found = true; }
    }
  }
  return k;
}

// Certified single-precision selection.  The component chosen by unmix depends on the FP64 probabilities only
// through the comparisons  U * sum_j q_j <= q_0 + ... + q_k.  Evaluating the q_k in FP32 (ex2.approx: K MUFU instead
// of K FP64 exp) perturbs every cumulative sum by less than ~(2e-6 + 3e-7 |log2 q_max|) * sum: the residual is
// centred in FP64 before it is rounded, so what is lost is 2^-24 of its distance to the components, i.e. a relative
// 2^-22 of the exponents.  Whenever U * sum is farther than margin = 4e-5 (1 + |log2 q_max| / 20) * sum from EVERY
// boundary, the FP32 decision therefore equals the FP64 one; otherwise (probability ~ 2 (K-1) 4e-5 per draw) the
// caller falls back to the FP64 statement.  The result is always the indicator the FP64 rule gives, so the bit-exact
// indicator counts against the oracle survive.
constexpr float kUnmixMargin = 4e-5f;

// r_c: residual minus the mixture's centre, rounded to FP32 by the caller; mu_of(s) is relative to the same centre.
// KMAX: compile-time bound on K (K == KMAX exactly unless KMAX is one of the open-ended sizes 4 and kMaxLogitK)
template <int KMAX, class FMU, class FL, class FH>
__device__ __forceinline__ bool unmix_certified(int K, float r_c, double unif, FMU mu_of, FL lconst2_of, FH hs2_of, int &kout) {
  constexpr bool kOpen = (KMAX == kMaxLogitK || KMAX == 4);
  float q[KMAX];
  float mx = -3.0e38f;
#pragma unroll
  for (int s = 0; s < KMAX; ++s) {
    if (kOpen && s >= K) { q[s] = -3.0e38f; continue; }
    const float dlt = r_c - mu_of(s);
    q[s] = fmaf(hs2_of(s), dlt * dlt, lconst2_of(s));
    mx = fmaxf(mx, q[s]);
  }
  float tot = 0.f;
#pragma unroll
  for (int s = 0; s < KMAX; ++s) {
    q[s] = exp2f(q[s] - mx);
    tot += q[s];
  }
  const float v = (float)unif * tot;
  float cs = 0.f, margin = 3.0e38f;
  int k = K - 1;
  bool found = false;
#pragma unroll
  for (int s = 0; s < KMAX - 1; ++s) {
    if (kOpen && s >= K - 1) break;
    cs += q[s];
    const float dd = v - cs;
    margin = fminf(margin, fabsf(dd));
    if (!found && dd <= 0.f) { k = s; found = true; }
  }
  kout = k;
  return margin > kUnmixMargin * tot * fmaf(0.05f, fabsf(mx), 1.f);
}

// The same certificate for a SCALE mixture (all means equal: the 9-component logit table), K known at compile time.
// log2 q_s = hs2_s d^2 + l0_s with l0 <= 0, so every exponent is <= 0 (no maximum to find, no overflow) and the largest one
// is at least the widest component's: |log2 q_max| <= |log2 q_wide| bounds the error scale of the certificate.  The
// indicator is the number of cumulative sums below U sum q -- the first k with U sum q <= q_0 + .. + q_k of the FP64 rule.
// 2^x by the SFU alone (MUFU.EX2, ~2 ulp, denormal results flushed to zero): exp2f() wraps the same instruction in a range
// test and two scalings that the exponents here (<= 0, and a flushed tail term is a zero probability) do not need
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int K>
__device__ __forceinline__ bool unmix_certified_scale(const LogitHot &h, float r_c, double unif, int &kout) {
  const float d2 = r_c * r_c;
  float cs[K];
#pragma unroll
  for (int s = 0; s < K; ++s) cs[s] = ex2_approx(fmaf(h.hs2[s], d2, h.l0[s]));
#pragma unroll
  for (int s = 1; s < K; ++s) cs[s] += cs[s - 1];
  const float tot = cs[K - 1];
  const float v = (float)unif * tot;
  int k = 0;
  float margin = 3.0e38f;
#pragma unroll
  for (int s = 0; s < K - 1; ++s) {
    const float dd = v - cs[s];
    k += dd > 0.f ? 1 : 0;
    margin = fminf(margin, fabsf(dd));
  }
  kout = k;
  const float qw = fmaf(h.hs2[K - 1], d2, h.l0[K - 1]);   // host orders nothing: `wide` is checked to be K - 1 when zero_mean is set
  return d2 < 2000.f && margin > kUnmixMargin * tot * fmaf(-0.05f, qw, 1.f);
}

// the FP64 statements of the selection, out of line: taken by ~1e-3 of the draws
__device__ __noinline__ int unmix_logit_fp64(const LogitMixture *__restrict__ m, double resid, double unif) {
  return unmix_generic(m->K, unif, [&](int s) {
    double x = (resid - __ldg(m->mu + s)) * __ldg(m->inv_sigma + s);
    return __ldg(m->lconst + s) - 0.5 * x * x;
  });
}
__device__ __noinline__ int unmix_table_fp64(const double *__restrict__ mu, const double *__restrict__ inv_sigma,
                                             const double *__restrict__ lconst, int K, double resid, double unif) {
  return unmix_generic(K, unif, [&](int s) {
    double x = (resid - __ldg(mu + s)) * __ldg(inv_sigma + s);
    return __ldg(lconst + s) - 0.5 * x * x;
  });
}

__device__ __forceinline__ int unmix_logit(const LogitHot &h, const LogitMixture *__restrict__ m, double resid, double unif) {
  int k;
  bool ok;
  auto mu_of = [&](int s) { return h.mu_c[s]; };
  auto lc_of = [&](int s) { return h.lconst2[s]; };
  auto hs_of = [&](int s) { return h.hs2[s]; };
  const float r_c = (float)(resid - h.center);
  if (h.K == 9 && h.zero_mean) ok = unmix_certified_scale<9>(h, r_c, unif, k);
  else if (h.K == 9) ok = unmix_certified<9>(9, r_c, unif, mu_of, lc_of, hs_of, k);
  else ok = unmix_certified<kMaxLogitK>(h.K, r_c, unif, mu_of, lc_of, hs_of, k);
  if (!ok) k = unmix_logit_fp64(m, resid, unif);
  return k;
}

__device__ __forceinline__ double rtrun_logit(double eta, bool success, double unif) {
  double c = __drcp_rn(1.0 + exp(eta));  // plogis(0 - eta)
  double u = success ? c + (1.0 - c) * unif : c * unif;
  u = fmin(u, 1.0 - 0x1p-53);
  u = fmax(u, 2.2250738585072014e-308);
  return log(u * __drcp_rn(1.0 - u)) + eta;
}

// ---- branch-free FP64 elementary functions for the Bernoulli fast path ------------------------------------------
// libdevice's exp / log / reciprocal carry range checks that end basic blocks; with the argument ranges known
// here they are unnecessary, and straight-line code lets the scheduler interleave the two observations a lane owns.

// exp(x) for |x| <= 708: Cody-Waite reduction by ln 2, degree-13 Taylor polynomial on |r| <= ln2 / 2 (truncation
// 4e-18), scaling by 2^k through the exponent field (the result is a normal number over the whole range).
// The coefficients live in constant memory: a DFMA takes one operand straight from the constant bank, whereas a literal
// FP64 constant costs two UMOV to assemble in a uniform register (38 of the 434 instructions of a Bernoulli draw).
static __constant__ double kExpC[17] = {
    1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10,   // log2 e, -ln2_hi, -ln2_lo (fdlibm split)
    1.6059043836821613e-10, 2.08767569878680990e-09, 2.50521083854417188e-08, 2.75573192239858907e-07,   // 1/13! .. 1/10!
    2.75573192239858907e-06, 2.48015873015873016e-05, 1.98412698412698413e-04, 1.38888888888888894e-03,  // 1/9! .. 1/6!
    8.33333333333333322e-03, 4.16666666666666644e-02, 1.66666666666666657e-01, 0.5, 1.0, 1.0};
__device__ __forceinline__ double exp_nobranch(double x) {
  const double kd = rint(x * kExpC[0]);
  double r = fma(kd, kExpC[1], x);
  r = fma(kd, kExpC[2], r);
  double p = kExpC[3];
#pragma unroll
  for (int i = 4; i < 17; ++i) p = fma(p, r, kExpC[i]);
  const int k = (int)kd;
  return __hiloint2double(__double2hiint(p) + k * 1048576, __double2loint(p));
}

// 1 / d for a positive normal d away from the exponent limits: MUFU seed + two Newton steps (<= 1 ulp)
__device__ __forceinline__ double rcp_nobranch(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

static __constant__ double kLogC[9] = {1.531383769920937332e-01, 2.222219843214978396e-01, 3.999999999940941908e-01,   // Lg6, Lg4, Lg2
                                       1.479819860511658591e-01, 1.818357216161805012e-01, 2.857142874366239149e-01,   // Lg7, Lg5, Lg3
                                       6.666666666666735130e-01, 6.93147180369123816490e-01, 1.90821492927058770002e-10};  // Lg1, ln2_hi, ln2_lo
// log(v) for a positive normal v: fdlibm e_log.c (m in [sqrt(1/2), sqrt 2), s = f / (2 + f), Lg1..Lg7), < 1 ulp
__device__ __forceinline__ double log_nobranch(double v) {
  int hi = __double2hiint(v);
  const int lo = __double2loint(v);
  int e = (hi >> 20) - 1023;
  hi &= 0x000fffff;
  const int adj = (hi + 0x95f64) & 0x100000;   // mantissa above sqrt 2: halve it, bump the exponent
  e += adj >> 20;
  const double m = __hiloint2double(hi | (adj ^ 0x3ff00000), lo);
  const double f = m - 1.0;
  const double s = f * rcp_nobranch(2.0 + f);
  const double z = s * s, w = z * z;
  const double t1 = w * fma(w, fma(w, kLogC[0], kLogC[1]), kLogC[2]);
  const double t2 = z * fma(w, fma(w, fma(w, kLogC[3], kLogC[4]), kLogC[5]), kLogC[6]);
  const double R = t1 + t2, hfsq = 0.5 * f * f, dk = (double)e;
  return fma(dk, kLogC[7], f - (hfsq - fma(s, hfsq + R, dk * kLogC[8])));
}

// rtrun_logit for |eta| < 600, algebraically: with E = exp(eta), c = 1 / (1 + E),
//   success: u = c + (1 - c) U,  u / (1 - u) = (1 + E U) / (E (1 - U))   =>  z = log((1 + E U) / (1 - U))
//   failure: u = c U,            u / (1 - u) = U / (1 + E - U)           =>  z = eta + log(U / (1 + E - U))
// one exp, one reciprocal, one log (trun_logit.cpp:163-174 evaluates three quotients); U in (0, 1) strictly.
__device__ __forceinline__ double rtrun_logit_fast(double eta, bool success, double unif) {
  const double E = exp_nobranch(eta);
  const double num = success ? fma(E, unif, 1.0) : unif;
  const double den = success ? 1.0 - unif : (1.0 + E) - unif;
  const double lg = log_nobranch(num * rcp_nobranch(den));
  return success ? lg : lg + eta;
}

// log(1 - Phi(a)) and the hazard phi(a) / (1 - Phi(a)) through erfcx (stable in both tails)
__device__ __forceinline__ double normal_hazard(double a) {
  return 0.7978845608028654 / erfcx(a * 0.7071067811865476);  // sqrt(2/pi) / erfcx(a / sqrt 2)
}

__device__ __forceinline__ void trun_norm_moments(double mu, double sigma, bool positive, double &mean, double &var) {
  double sigsq = sigma * sigma;
  if (positive) {
    double alpha = (0.0 - mu) / sigma;
    double r = normal_hazard(alpha);
    mean = mu + sigma * r;
    var = sigsq * (1 - r * (r - alpha));
  } else {
    double beta = (0.0 - mu) / sigma;
    double r = normal_hazard(-beta);
    mean = mu - sigma * r;
    var = sigsq * (1 - beta * r - r * r);
  }
  var = fmax(var, 0.0);
}

// Binomial(n, p) from one uniform: chop-down search outward from the mode (see oracle).
__device__ inline int64_t binomial_from_uniform(int64_t n, double p, double unif) {
  if (n <= 0 || p <= 0.0) return 0;
  if (p >= 1.0) return n;
  double q = 1.0 - p, dn = (double)n;
  int64_t m = (int64_t)floor((dn + 1.0) * p);
  if (m > n) m = n;
  double dm = (double)m;
  double fm = exp(lgamma(dn + 1.0) - lgamma(dm + 1.0) - lgamma(dn - dm + 1.0) + dm * log(p) + (dn - dm) * log(q));
  double u = unif - fm;
  if (u <= 0) return m;
  double r = p / q, fu = fm, fd = fm;
  int64_t ku = m, kd = m;
  for (;;) {
    bool moved = false;
    if (ku < n) {
      fu *= r * (double)(n - ku) / (double)(ku + 1);
      ++ku; moved = true;
      u -= fu;
      if (u <= 0) return ku;
    }
    if (kd > 0) {
      fd *= (double)kd / (r * (double)(n - kd + 1));
      --kd; moved = true;
      u -= fd;
      if (u <= 0) return kd;
    }
    if (!moved) return m;
  }
}

__device__ inline void multinomial_from_uniforms(int64_t n, int K, const double *prob, const double *unif, int64_t *out) {
  double p_tot = 0;
  for (int k = 0; k < K; ++k) { p_tot += prob[k]; out[k] = 0; }
  if (n == 0) return;
  for (int k = 0; k < K - 1; ++k) {
    double pp = prob[k] / p_tot;
    pp = fmin(pp, 1.0);
    if (!(pp > 0.0)) pp = 0.0;
    out[k] = binomial_from_uniform(n, pp, unif[k]);
    n -= out[k];
    if (n <= 0) return;
    p_tot -= prob[k];
  }
  out[K - 1] = n;
}

// CLT branch, kept out of line: Bernoulli data never take it.
__device__ __noinline__ void logit_impute_large(const LogitMixture *__restrict__ mp, double ntrials, double y, double eta,
                                                const RngKey &key, uint64_t row, double &sum, double &info) {
  const LogitMixture &m = *mp;
  const int K = m.K;
  double p0[kMaxLogitK], p1[kMaxLogitK], un0[8], un1[8];
  int64_t N0[kMaxLogitK], N1[kMaxLogitK];
  double neg = 1.0 / (1.0 + exp(eta)), pos = 1.0 / (1.0 + exp(-eta));
  double s0 = 0, s1 = 0;
  for (int k = 0; k < K; ++k) {
    double a = (0.0 - eta) / m.sigma[k];
    p0[k] = m.weights[k] / neg * normcdf(a);
    p1[k] = m.weights[k] / pos * normcdf(-a);
  }
  for (int k = 0; k < K; ++k) { s0 += p0[k]; s1 += p1[k]; }
  for (int k = 0; k < K; ++k) { p0[k] /= s0; p1[k] /= s1; }
  for (int b = 0; b < 4; ++b) {
    uniform_pair(key, row, b, un0[2 * b], un0[2 * b + 1]);
    uniform_pair(key, row, 4 + b, un1[2 * b], un1[2 * b + 1]);
  }
  multinomial_from_uniforms((int64_t)(ntrials - y), K, p0, un0, N0);
  multinomial_from_uniforms((int64_t)y, K, p1, un1, N1);
  double mean = 0, var = 0, w = 0;
  for (int k = 0; k < K; ++k) {
    int64_t tot = N0[k] + N1[k];
    if (tot == 0) continue;
    double sigsq = m.sigma[k] * m.sigma[k], sig4 = sigsq * sigsq;
    w += (double)tot / sigsq;
    double tm, tv;
    if (N0[k] > 0) {
      trun_norm_moments(eta, m.sigma[k], false, tm, tv);
      mean += (double)N0[k] * tm / sigsq;
      var += (double)N0[k] * tv / sig4;
    }
    if (N1[k] > 0) {
      trun_norm_moments(eta, m.sigma[k], true, tm, tv);
      mean += (double)N1[k] * tm / sigsq;
      var += (double)N1[k] * tv / sig4;
    }
  }
  double u0, u1;
  uniform_pair(key, row, 8, u0, u1);
  double zn = sqrt(-2.0 * log(u0)) * cos(kTwoPi * u1);
  sum = mean + sqrt(var) * zn;
  info = w;
}

// What lives in global memory for the out-of-line paths: the FP64 mixture and a copy of the hot-loop constants.
struct LogitMixtureDev { LogitMixture full; LogitHot hot; };

// One trial: truncated-logistic draw + indicator.  Shared by the Bernoulli fast path (h in the constant bank) and
// the general path (h in global memory).
__device__ __forceinline__ void logit_trial(const LogitHot &h, const LogitMixture *__restrict__ m, double eta, bool success,
                                            double u0, double u1, double &sum, double &info) {
  const double z = rtrun_logit(eta, success, u0);
  const int k = unmix_logit(h, m, z - eta, u1);
  const double cw = h.inv_sigsq[k];
  info += cw;
  sum += z * cw;
}

// Everything that is not a Bernoulli observation: validation, the per-trial loop for 1 < n_i <= clt_threshold
// (BinomialLogitDataImputer.cpp:134-152), the CLT branch (:155-210).  Out of line so that the Bernoulli path keeps
// its registers; reads the mixture from global memory.
__device__ __noinline__ bool logit_impute_general(const LogitMixtureDev *__restrict__ md, int clt_threshold, double ntrials, double y,
                                                  double eta, uint64_t seed, uint64_t iteration, uint64_t row, double *sum_out,
                                                  double *info_out) {
  double sum = 0, info = 0;
  *sum_out = 0; *info_out = 0;
  if (!(y <= ntrials) || y < 0 || ntrials < 0 || !isfinite(eta)) return false;
  RngKey key;
  key.seed = seed; key.iteration = iteration;
  philox_round_keys(key);
  if (ntrials > (double)clt_threshold) {
    if (md->full.K > 9) return false;  // slot layout of the CLT branch: K - 1 <= 8 conditional binomials per side
    logit_impute_large(&md->full, ntrials, y, eta, key, row, sum, info);
  } else {
    const int nt = (int)ceil(ntrials), ns = (int)ceil(y);   // i < ntrials, i < y for integer i (the reference's loop bounds are doubles)
    for (int i = 0; i < nt; ++i) {
      double u0, u1;
      uniform_pair(key, row, (uint32_t)i, u0, u1);
      logit_trial(md->hot, &md->full, eta, i < ns, u0, u1, sum, info);
    }
  }
  *sum_out = sum; *info_out = info;
  return true;
}

// the Bernoulli fast path applies: one trial, a 0/1 response, |eta| < 600 (so that nothing in it can overflow)
__device__ __forceinline__ bool logit_is_bernoulli_fast(double ntrials, double y, double eta) {
  return ntrials == 1.0 && (y == 0.0 || y == 1.0) && fabs(eta) < 600.0;
}

// R Bernoulli observations of one lane, written operation by operation across the R so that each stage is R
// independent instruction streams in ONE basic block (the scheduler interleaves them: ILP R on the dependent FP64
// chains of exp / reciprocal / log).  The FP64 fallbacks of the certified selection come after the straight-line part.
template <int R>
__device__ __forceinline__ void logit_bernoulli_draws(const LogitHot &h, const LogitMixtureDev *__restrict__ md, const double (&eta)[R],
                                                      const double (&y)[R], const RngKey &key, const uint64_t (&row)[R],
                                                      double (&sum)[R], double (&info)[R]) {
  double u0[R], u1[R], z[R];
  int k[R];
  bool ok[R];
#pragma unroll
  for (int j = 0; j < R; ++j) uniform_pair(key, row[j], 0u, u0[j], u1[j]);
#pragma unroll
  for (int j = 0; j < R; ++j) z[j] = rtrun_logit_fast(eta[j], y[j] != 0.0, u0[j]);
  auto mu_of = [&](int s) { return h.mu_c[s]; };
  auto lc_of = [&](int s) { return h.lconst2[s]; };
  auto hs_of = [&](int s) { return h.hs2[s]; };
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const float r_c = (float)((z[j] - eta[j]) - h.center);
    if (h.K == 9 && h.zero_mean) ok[j] = unmix_certified_scale<9>(h, r_c, u1[j], k[j]);
    else if (h.K == 9) ok[j] = unmix_certified<9>(9, r_c, u1[j], mu_of, lc_of, hs_of, k[j]);
    else ok[j] = unmix_certified<kMaxLogitK>(h.K, r_c, u1[j], mu_of, lc_of, hs_of, k[j]);
  }
#pragma unroll
  for (int j = 0; j < R; ++j) {
    if (!ok[j]) k[j] = unmix_logit_fp64(&md->full, z[j] - eta[j], u1[j]);
    info[j] = h.inv_sigsq[k[j]];
    sum[j] = z[j] * info[j];
  }
}

// BinomialLogitCltDataImputer::impute.  Returns false on invalid input (y > n, negative, NaN eta).
__device__ __forceinline__ bool logit_impute(const LogitHot &h, const LogitMixtureDev *__restrict__ md, int clt_threshold,
                                             double ntrials, double y, double eta, const RngKey &key, uint64_t row,
                                             double &sum, double &info) {
  if (logit_is_bernoulli_fast(ntrials, y, eta)) {   // Bernoulli: one trial, straight line
    double u0, u1;
    uniform_pair(key, row, 0u, u0, u1);
    const double z = rtrun_logit_fast(eta, y != 0.0, u0);
    const int k = unmix_logit(h, &md->full, z - eta, u1);
    info = h.inv_sigsq[k];
    sum = z * info;
    return true;
  }
  return logit_impute_general(md, clt_threshold, ntrials, y, eta, key.seed, key.iteration, row, &sum, &info);
}

// ---- probit sibling ---------------------------------------------------------------------
// N(eta, 1) truncated to z > 0 (positive) or z < 0 by inversion from one uniform, on the tail that does not cancel:
//   z > 0:  z = eta - Phi^-1(u Phi(eta));   z < 0:  z = eta + Phi^-1(u Phi(-eta))
// (BinomialProbitDataImputer.cpp:55-66 calls rtrun_norm_mt, distributions/trun_norm.cpp:36-110 -- rejection samplers: the
// same law from another stream).  Beyond 30 standard deviations the exponential limit of the normal tail is used.
__device__ __forceinline__ double rtrun_norm_unit(double eta, bool positive, double unif) {
  const double m = positive ? eta : -eta;
  double t = (m < -30.0) ? -log(unif) / (-m) : m - normcdfinv(unif * normcdf(m));
  t = fmax(t, 0.0);
  return positive ? t : -t;
}

// BinomialProbitDataImputer::impute (BinomialProbitDataImputer.cpp:31-68): the sum of the n_i latent normals of an
// observation -- y_i of them positive, n_i - y_i negative -- drawn one by one, or in one normal draw from the truncated
// moments when a side has more than clt_threshold members.  Philox slots as in the oracle (bo_probit_impute).
__device__ __noinline__ bool probit_impute(int clt_threshold, double ntrials, double nsuccess, double eta, uint64_t seed,
                                           uint64_t iteration, uint64_t row, double *sum_z) {
  *sum_z = 0;
  const long long n = llround(ntrials), y = llround(nsuccess);
  if (y < 0 || n < 0 || y > n || !isfinite(eta)) return false;
  RngKey key;
  key.seed = seed; key.iteration = iteration;
  philox_round_keys(key);
  double ans = 0, mean, var, u0, u1;
  uint32_t slot = 0;
  if (y > clt_threshold) {
    trun_norm_moments(eta, 1.0, true, mean, var);
    uniform_pair(key, row, slot++, u0, u1);
    ans += (double)y * mean + sqrt((double)y * var) * (sqrt(-2.0 * log(u0)) * cos(kTwoPi * u1));
  } else {
    for (long long i = 0; i < y; ++i) { uniform_pair(key, row, slot++, u0, u1); ans += rtrun_norm_unit(eta, true, u0); }
  }
  if (n - y > clt_threshold) {
    trun_norm_moments(eta, 1.0, false, mean, var);
    uniform_pair(key, row, slot++, u0, u1);
    ans += (double)(n - y) * mean + sqrt((double)(n - y) * var) * (sqrt(-2.0 * log(u0)) * cos(kTwoPi * u1));
  } else {
    for (long long i = 0; i < n - y; ++i) { uniform_pair(key, row, slot++, u0, u1); ans += rtrun_norm_unit(eta, false, u0); }
  }
  *sum_z = ans;
  return true;
}

// ---- Student-t sibling (TRegressionSampler) ----------------------------------
// Gamma(shape, rate) on the Philox stream.  The reference's TDataImputer::impute (Models/Glm/PosteriorSamplers/
// TDataImputer.cpp:25-29) calls rgamma_mt (distributions/Rmath_dist.cpp:72-74 -> Bmath/rgamma.cpp, Ahrens-Dieter GD / GS):
// the same law from another stream.  Here Marsaglia & Tsang's (2000) method: attempt k consumes block k of the row -- its
// first uniform gives the normal deviate by inversion, its second decides acceptance (>= 95 % accept for shape >= 1).
// A shape below 1 is drawn as Gamma(shape + 1) U^(1/shape) with U from block 0xFFFF.  Slots as in the oracle (bo_rgamma).
constexpr uint32_t kGammaBoostSlot = 0xFFFFu;
constexpr int kGammaMaxAttempts = 4096;
// attempts first_attempt, first_attempt + 1, ... of the rejection loop; the accepted d v (before the rate) or -1
__device__ __noinline__ double rgamma_attempts(double d, double c, int first_attempt, uint64_t seed, uint64_t iteration, uint64_t row) {
  RngKey key;
  key.seed = seed; key.iteration = iteration;
  philox_round_keys(key);
  double u0, u1;
  for (int k = first_attempt; k < kGammaMaxAttempts; ++k) {
    uniform_pair(key, row, (uint32_t)k, u0, u1);
    const double x = normcdfinv(u0);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (log(u1) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
  }
  return -1.0;
}
__device__ inline bool rgamma_philox(double shape, double rate, const RngKey &key, uint64_t row, double *out) {
  *out = 0;
  if (!(shape > 0) || !(rate > 0) || !isfinite(shape) || !isfinite(rate)) return false;
  const double a = shape < 1.0 ? shape + 1.0 : shape;
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double g = rgamma_attempts(d, c, 0, key.seed, key.iteration, row);
  if (g < 0) return false;
  if (shape < 1.0) {
    double u0, u1;
    uniform_pair(key, row, kGammaBoostSlot, u0, u1);
    g *= exp(log(u0) / shape);
  }
  *out = g / rate;
  return true;
}

// The same draw for R observations of one lane with shape >= 1 (nu >= 1: every Student-t model one meets): attempt 0 of all
// R rows in straight line -- independent dependency chains the scheduler interleaves, 95-98 % of them accept -- and the
// rejected rows finish out of line from attempt 1.  Same attempts, same blocks, same result as rgamma_philox.
template <int R>
__device__ __forceinline__ bool rgamma_philox_rows(double shape, const double (&rate)[R], const RngKey &key, const uint64_t (&rows)[R],
                                                   double (&out)[R]) {
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double g[R];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    double u0, u1;
    uniform_pair(key, rows[j], 0u, u0, u1);
    const double x = normcdfinv(u0);
    const double v = 1.0 + c * x, v3 = v * v * v;
    // branch-free logarithms (positive normal arguments; v <= 0 is rejected before its logarithm is looked at)
    const bool accept = v > 0.0 && log_nobranch(u1) < 0.5 * x * x + d - d * v3 + d * log_nobranch(fmax(v3, 1e-300));
    g[j] = accept ? d * v3 : -1.0;
  }
#pragma unroll
  for (int j = 0; j < R; ++j) {
    if (g[j] < 0) g[j] = rgamma_attempts(d, c, 1, key.seed, key.iteration, rows[j]);
    ok = ok && g[j] >= 0 && rate[j] > 0 && isfinite(rate[j]);
    out[j] = ok ? g[j] / rate[j] : 0.0;
  }
  return ok;
}

// log density of the Student t observation y = mu + sigma t_nu at standardised residual delta, WITHOUT the terms that
// depend on (sigma, nu) alone: -(nu + 1)/2 log1p(delta^2 / nu).  The caller adds n (lgamma((nu+1)/2) - lgamma(nu/2) -
// log(nu pi)/2 - log sigma)  (dstudent, distributions/student_fix.cpp:28-41 over Bmath dt).
__device__ __forceinline__ double student_log_kernel(double delta, double nu) { return -0.5 * (nu + 1.0) * log1p(delta * delta / nu); }

// ---- Poisson --------------------------------------------------------------------------
__device__ __forceinline__ int poisson_table_find(const PoissonTable &t, int64_t nu) {
  if (nu < (int64_t)t.dense_n) return nu >= 0 ? __ldg(t.dense + nu) : -1;
  int lo = 0, hi = t.ntab - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    int64_t v = __ldg(t.nu + mid);
    if (v == nu) {
      while (mid > 0 && __ldg(t.nu + mid - 1) == nu) --mid;
      return mid;
    }
    if (v < nu) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// unmix against table entry `entry` (global memory); logw = log(1 / sigma_k^2)
__device__ __forceinline__ bool unmix_poisson(const PoissonTable &t, int entry, double resid, double unif,
                                              double &mu, double &weight, double &logw, int &kout) {
  if (entry < 0) return false;
  int a = __ldg(t.offset + entry), K = __ldg(t.offset + entry + 1) - a;
  if (K > kMaxLogitK) return false;
  int k;
  auto mu_of = [&](int s) { return __ldg(t.mu_f + a + s); };
  auto lc_of = [&](int s) { return __ldg(t.lconst2_f + a + s); };
  auto hs_of = [&](int s) { return __ldg(t.hs2_f + a + s); };
  bool ok;
  const float r_c = (float)(resid - __ldg(t.mu + a));
  if (K == 10) ok = unmix_certified<10>(10, r_c, unif, mu_of, lc_of, hs_of, k);
  else if (K <= 4) ok = unmix_certified<4>(K, r_c, unif, mu_of, lc_of, hs_of, k);
  else ok = unmix_certified<kMaxLogitK>(K, r_c, unif, mu_of, lc_of, hs_of, k);
  if (!ok) k = unmix_table_fp64(t.mu + a, t.inv_sigma + a, t.lconst + a, K, resid, unif);
  mu = __ldg(t.mu + a + k);
  weight = __ldg(t.inv_sigsq + a + k);
  logw = __ldg(t.logw + a + k);
  kout = k;
  return true;
}

// the nu = 1 entry (every observation draws from it) from the constant bank
__device__ __forceinline__ void unmix_poisson_ext(const LogitHot &h, const PoissonTable &t, double resid, double unif,
                                                  double &mu, double &weight, double &logw, int &kout) {
  int k;
  bool ok;
  auto mu_of = [&](int s) { return h.mu_c[s]; };
  auto lc_of = [&](int s) { return h.lconst2[s]; };
  auto hs_of = [&](int s) { return h.hs2[s]; };
  const float r_c = (float)(resid - h.center);
  if (h.K == 10) ok = unmix_certified<10>(10, r_c, unif, mu_of, lc_of, hs_of, k);
  else ok = unmix_certified<kMaxLogitK>(h.K, r_c, unif, mu_of, lc_of, hs_of, k);
  if (!ok) {
    const int a = __ldg(t.offset + t.e1);
    k = unmix_table_fp64(t.mu + a, t.inv_sigma + a, t.lconst + a, h.K, resid, unif);
  }
  mu = h.mu_d[k]; weight = h.inv_sigsq[k]; logw = h.logw[k]; kout = k;
}

// cooperative fill of the shared-memory copy (all threads of the CTA; the caller synchronises afterwards)
__device__ __forceinline__ void poisson_smem_fill(PoissonSmem *ps, const PoissonTable &t, int tid, int nthreads) {
  if (tid == 0) {
    int pos = 0;
    for (int nu = 0; nu < kPoisSmemNu; ++nu) {
      ps->off[nu] = pos;
      const int e = nu < t.dense_n ? __ldg(t.dense + nu) : -1;
      if (e >= 0) {
        const int a = __ldg(t.offset + e), K = __ldg(t.offset + e + 1) - a;
        if (K <= kMaxLogitK && pos + K <= kPoisSmemComps) { ps->center[nu] = __ldg(t.mu + a); pos += K; }
      }
    }
    ps->off[kPoisSmemNu] = pos;
  }
  __syncthreads();
  for (int nu = tid; nu < kPoisSmemNu; nu += nthreads) {
    const int K = ps->off[nu + 1] - ps->off[nu];
    if (K > 0) {
      const int a = __ldg(t.offset + __ldg(t.dense + nu)), b = ps->off[nu];
      for (int k = 0; k < K; ++k) {
        ps->mu_f[b + k] = __ldg(t.mu_f + a + k); ps->lconst2_f[b + k] = __ldg(t.lconst2_f + a + k); ps->hs2_f[b + k] = __ldg(t.hs2_f + a + k);
        ps->mu[b + k] = __ldg(t.mu + a + k); ps->inv_sigsq[b + k] = __ldg(t.inv_sigsq + a + k); ps->logw[b + k] = __ldg(t.logw + a + k);
      }
    }
  }
}

// unmix against the shared-memory copy of entry nu (nu < kPoisSmemNu, entry present); the FP64 fallback reads global memory
__device__ __forceinline__ bool unmix_poisson_smem(const PoissonSmem *ps, const PoissonTable &t, int nu, double resid, double unif,
                                                   double &mu, double &weight, double &logw, int &kout) {
  const int b = ps->off[nu], K = ps->off[nu + 1] - b;
  if (K <= 0) return false;
  int k;
  auto mu_of = [&](int s) { return ps->mu_f[b + s]; };
  auto lc_of = [&](int s) { return ps->lconst2_f[b + s]; };
  auto hs_of = [&](int s) { return ps->hs2_f[b + s]; };
  bool ok;
  const float r_c = (float)(resid - ps->center[nu]);
  if (K == 10) ok = unmix_certified<10>(10, r_c, unif, mu_of, lc_of, hs_of, k);
  else if (K <= 4) ok = unmix_certified<4>(K, r_c, unif, mu_of, lc_of, hs_of, k);
  else ok = unmix_certified<kMaxLogitK>(K, r_c, unif, mu_of, lc_of, hs_of, k);
  if (!ok) {
    const int a = __ldg(t.offset + __ldg(t.dense + nu));
    k = unmix_table_fp64(t.mu + a, t.inv_sigma + a, t.lconst + a, K, resid, unif);
  }
  mu = ps->mu[b + k]; weight = ps->inv_sigsq[b + k]; logw = ps->logw[b + k];
  kout = k;
  return true;
}

struct PoissonLatent { double z_int, mu_int, w_int, z_ext, mu_ext, w_ext, lw_int, lw_ext; int k_int, k_ext; };

// The extreme-eta statement of PoissonDataImputer::impute (PoissonDataImputer.cpp:55-79), out of line.
__device__ __noinline__ double poisson_zext_extreme(double eta, double delta, double e1) {
  if (delta > 0) {
    double err = -log(e1);
    double a = log(delta), b = -err - eta;
    if (a < b) { double tmp = a; a = b; b = tmp; }
    return -(a + log1p(exp(b - a)));
  }
  return eta + (-log(e1));
}

// PoissonDataImputer::impute.  rc: 0 ok, 1 nu missing from the table, 2 invalid input.
//   tau = E * Beta(y, 1) = E * U^(1/y) = E * exp(log U / y);   z_int = -log tau = -(log E + log U / y)
//   z_ext = -log(delta + Exp(1) / exp(eta)),  delta = E - tau
// For exposures and eta in the ordinary range every elementary function is the branch-free kind of the logit path.
__device__ __forceinline__ int poisson_impute(const LogitHot &ext, const PoissonTable &t, int64_t y, double exposure, double eta,
                                              const RngKey &key, uint64_t row, PoissonLatent &o, const PoissonSmem *ps = nullptr) {
  o.z_int = o.mu_int = o.w_int = o.lw_int = 0; o.k_int = o.k_ext = -1;
  if (y < 0 || !(exposure >= 0) || !isfinite(eta)) return 2;
  double ua0, ua1, ub0, ub1;
  uniform_pair(key, row, 0, ua0, ua1);
  uniform_pair(key, row, 1, ub0, ub1);
  const bool ordinary = fabs(eta) < 600 && exposure > 1e-280 && exposure < 1e280;
  double tau = 0.0, z_int = 0.0, z_ext;
  if (ordinary) {
    if (y > 0) {
      const double t1 = log_nobranch(ua0) * rcp_nobranch((double)y);
      tau = exposure * exp_nobranch(t1);
      z_int = -(log_nobranch(exposure) + t1);
    }
    const double delta = exposure - tau;
    const double e1 = -log_nobranch(ua1);
    z_ext = -log_nobranch(fma(exp_nobranch(-eta), e1, delta));
  } else {
    if (y > 0) { tau = exposure * pow(ua0, 1.0 / (double)y); z_int = -log(tau); }
    const double delta = exposure - tau;
    const double e1 = -log(ua1);
    z_ext = fabs(eta) < 600 ? -log(delta + (1.0 / exp(eta)) * e1) : poisson_zext_extreme(eta, delta, e1);
  }
  unmix_poisson_ext(ext, t, z_ext - eta, ub0, o.mu_ext, o.w_ext, o.lw_ext, o.k_ext);
  o.z_ext = z_ext;
  if (y > 0) {
    o.z_int = z_int;
    if (y >= t.gaussian_cutoff) {
      o.mu_int = -log((double)y);
      o.w_int = 1.0 / (1.0 / (double)y);
      o.lw_int = log(o.w_int);
    } else if (ps != nullptr && y < kPoisSmemNu && ps->off[y + 1] > ps->off[y]) {
      unmix_poisson_smem(ps, t, (int)y, z_int - eta, ub1, o.mu_int, o.w_int, o.lw_int, o.k_int);
    } else {
      if (!unmix_poisson(t, poisson_table_find(t, y), z_int - eta, ub1, o.mu_int, o.w_int, o.lw_int, o.k_int)) return 1;
    }
  }
  return 0;
}

// ---- log densities (value only) ---------------------------------------------------------
// dbinom(y; n, p) on the log scale.  The reference uses R's saddle-point form
// (Bmath/dbinom.cpp:61-99); lgamma + x log p + (n-x) log q agrees to ~1e-15 absolute per term.
__device__ __forceinline__ double dbinom_log(double x, double n, double eta) {
  // log p = -log1p(exp(-eta)), log q = -log1p(exp(eta)), stable for any eta
  double lp = eta > 0 ? -log1p(exp(-eta)) : eta - log1p(exp(eta));
  double lq = eta > 0 ? -eta - log1p(exp(-eta)) : -log1p(exp(eta));
  double lc = (x == 0 || x == n) ? 0.0 : lgamma(n + 1.0) - lgamma(x + 1.0) - lgamma(n - x + 1.0);
  double t1 = x == 0 ? 0.0 : x * lp;
  double t2 = (n - x) == 0 ? 0.0 : (n - x) * lq;
  return lc + t1 + t2;
}

__device__ __forceinline__ double dpois_log(double x, double lambda) {
  if (lambda == 0) return x == 0 ? 0.0 : -INFINITY;
  if (x == 0) return -lambda;
  return x * log(lambda) - lambda - lgamma(x + 1.0);
}

}  // namespace boomgpu
