// kernels.cuh -- CUDA kernels of the auxiliary-mixture Gibbs hot path, written for sm_100a.
//
//   (fused_tma_kernel    p <= 64 : the single-pass kernel in use, see fused_tma.cuh)
//   fused_small_kernel   p <= 64 : its first version, kept for X that TMA cannot describe (odd leading dimension,
//                        unaligned adopted pointer).  cp.async-staged row chunks in shared memory,
//                        eta = x'beta, latent draw, weighted SYRK on FP64 DMMA with the whole upper
//                        triangle of X'WX resident in registers.  HBM-bound (8 n (p+2) bytes).
//   impute_rows_kernel   p  > 64 : pass 1, warp-per-row GEMV + latent draw -> (w_i, s_i).  HBM-bound.
//   syrk_dmma_kernel     p  > 64 : pass 2, split-K weighted SYRK  X' diag(w) X  (upper triangle) and
//                        X's on FP64 DMMA (mma.sync m8n8k4), X tiles staged by TMA tensor tiles
//                        (cp.async.bulk.tensor.2d + mbarrier) through a 6-stage ring.  FP64-pipe-bound.
//   syrk_rdiag_kernel    p  > 64 and p not a multiple of 128: the regions of the ragged last column block, in strip forms
//                        whose work follows the block's width (same ring, same partial tiles).
//   panel_dmma_kernel    active-set statistics: X' diag(w) X_A for the included columns, one pass over X.
//   residual_*_kernel, student_loglike_kernel   the Student-t sibling's observed-data log likelihood (nu draw).
//   reduce_* kernels     deterministic (fixed order) reduction of the per-CTA partials.
//
// The reference equivalent of all of this is the per-observation loop
// Models/PosteriorSamplers/Imputer.hpp:175-180 with SufficientStatistics::update
// (Models/Glm/PosteriorSamplers/BinomialLogitAuxmixSampler.cpp:61-67) /
// WeightedRegSuf::add_data (Models/Glm/WeightedRegressionModel.cpp:161-169) inside.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "draws.cuh"

namespace boomgpu {

// kLogitLL / kPoissonLL: no draw; the row's "latent" is its log-likelihood curvature, so the same kernels return
// log likelihood, gradient and (minus) Hessian in one pass (BinomialLogitModel.cpp:140-180, PoissonRegressionModel.cpp:56-95)
// kProbit: the probit sibling (BinomialProbitSpikeSlabSampler): weight n_i, weighted value = sum of the latent normals
// kStudentT: the Student-t sibling (TRegressionSampler): weight w_i ~ Gamma((nu+1)/2, (nu + delta_i^2)/2), weighted value w_i y_i
enum Model : int { kLogit = 0, kPoisson = 1, kSupplied = 2, kLogitLL = 3, kPoissonLL = 4, kProbit = 5, kStudentT = 6 };

struct RowData {
  const double *X;
  int64_t ldx;
  int64_t n;
  int p;
  const double *y;         // logit: successes
  const double *ntrials;   // logit
  const int64_t *yi;       // poisson counts
  const double *exposure;  // poisson
  const double *w_in;      // supplied latents (kSupplied)
  const double *s_in;
  uint64_t row_offset;
};

struct RowOut {           // optional per-row outputs (test hooks); any may be null
  double *w;              // weight  (logit: information;      poisson: w_int + w_ext)
  double *s;              // weighted value (logit: sum;       poisson: w_int r_int + w_ext r_ext)
  double *out6;
  int32_t *k2;
};

struct DrawParams {
  LogitHot hot;             // constant bank: the logit mixture
  LogitHot ext;             // constant bank: the Poisson table's nu = 1 entry (every observation draws from it)
  const LogitMixtureDev *mix;  // global memory (FP64 fallback of the selection, general / CLT path)
  PoissonTable tab;
  RngKey key;
  int clt_threshold;
  double log_alpha;         // kLogitLL: eta = x'beta - log_alpha (BinomialLogitModel.cpp:168)
  double t_inv_sigma, t_nu; // kStudentT: 1 / sigma and the tail thickness of TRegressionModel
};

// ---- small helpers ------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One observation: draw (or read) the latent summary.  Returns weight / weighted value of the
// single rank-1 update the row contributes, plus the scalar statistics WeightedRegSuf keeps.
struct RowLatent { double w, s, yWy, sumlogw, count; };

// The per-observation inputs besides x: loaded early (before waiting on staged rows) to hide their latency.
struct RowObs { double y, aux; int64_t yi; };   // logit: (successes, trials); Poisson: (yi, exposure); supplied: (w, s)

template <int MODEL>
__device__ __forceinline__ RowObs load_obs(const RowData &d, int64_t i) {
  RowObs o;
  o.y = 0; o.aux = 0; o.yi = 0;
  if (MODEL == kLogit || MODEL == kLogitLL || MODEL == kProbit) { o.y = __ldg(d.y + i); o.aux = __ldg(d.ntrials + i); }
  else if (MODEL == kPoisson || MODEL == kPoissonLL) { o.yi = __ldg(d.yi + i); o.aux = __ldg(d.exposure + i); }
  else if (MODEL == kStudentT) { o.y = __ldg(d.y + i); }
  else { o.y = __ldg(d.w_in + i); o.aux = __ldg(d.s_in + i); }
  return o;
}

template <int MODEL>
__device__ __forceinline__ RowLatent impute_row(const RowData &d, const DrawParams &prm, const RowOut &out, const RowObs &obs,
                                                int64_t i, double eta, int *err, const PoissonSmem *tab_s = nullptr) {
  RowLatent r;
  r.w = r.s = r.yWy = r.sumlogw = 0; r.count = 1;
  if (MODEL == kLogit) {
    double sum, info;
    bool ok = logit_impute(prm.hot, prm.mix, prm.clt_threshold, obs.aux, obs.y, eta, prm.key, d.row_offset + (uint64_t)i, sum, info);
    if (!ok) { atomicOr(err, 2); sum = 0; info = 0; }
    r.w = info; r.s = sum;
  } else if (MODEL == kPoisson) {
    PoissonLatent o;
    int rc = poisson_impute(prm.ext, prm.tab, obs.yi, obs.aux, eta, prm.key, d.row_offset + (uint64_t)i, o, tab_s);
    if (rc) {
      atomicOr(err, rc == 1 ? 1 : 2);
    } else {
      double re = o.z_ext - o.mu_ext;
      r.w = o.w_ext; r.s = o.w_ext * re; r.yWy = o.w_ext * re * re; r.sumlogw = o.lw_ext;
      if (obs.yi > 0) {
        double ri = o.z_int - o.mu_int;
        r.w += o.w_int; r.s += o.w_int * ri; r.yWy += o.w_int * ri * ri; r.sumlogw += o.lw_int;
        r.count = 2;
      }
      if (out.out6) {
        double *q = out.out6 + 6 * i;
        q[0] = o.z_int; q[1] = o.mu_int; q[2] = o.w_int; q[3] = o.z_ext; q[4] = o.mu_ext; q[5] = o.w_ext;
      }
      if (out.k2) { out.k2[2 * i] = o.k_int; out.k2[2 * i + 1] = o.k_ext; }
    }
  } else if (MODEL == kProbit) {
    double sz;
    const bool ok = probit_impute(prm.clt_threshold, obs.aux, obs.y, eta, prm.key.seed, prm.key.iteration, d.row_offset + (uint64_t)i, &sz);
    if (!ok) { atomicOr(err, 2); sz = 0; }
    r.w = ok ? obs.aux : 0.0;    // refresh_xtx: xtx += n_i x x' (BinomialProbitSpikeSlabSampler.cpp:72-78)
    r.s = sz;
  } else if (MODEL == kStudentT) {
    // TRegressionSampler::impute_latent_data (TRegressionSampler.cpp:128-143): weight | residual, then
    // WeightedRegSuf::add_data(x, y, weight) and the weight model's GammaSuf (sum, sum of logs: the same two scalars)
    const double delta = (obs.y - eta) * prm.t_inv_sigma;
    double w;
    const bool ok = isfinite(delta) && rgamma_philox(0.5 * (prm.t_nu + 1.0), 0.5 * (prm.t_nu + delta * delta), prm.key, d.row_offset + (uint64_t)i, &w);
    if (!ok) { atomicOr(err, 2); w = 0; }
    r.w = w; r.s = w * obs.y; r.yWy = w * obs.y * obs.y; r.sumlogw = ok ? log(w) : 0.0;
  } else if (MODEL == kLogitLL) {
    // w = n p q (so that -X'WX is the Hessian), s = y - n p (X's is the gradient), the log density rides in the yWy slot
    const double e = eta - prm.log_alpha;
    const double pr = 1.0 / (1.0 + exp(-e));
    r.w = obs.aux * pr * (1.0 - pr);
    r.s = obs.y - obs.aux * pr;
    r.yWy = dbinom_log(obs.y, obs.aux, e);
    if (!(obs.y <= obs.aux) || obs.y < 0 || !isfinite(eta)) atomicOr(err, 2);
  } else if (MODEL == kPoissonLL) {
    const double lambda = exp(eta);
    r.w = lambda;                                  // the reference's Hessian weight has no exposure (PoissonRegressionModel.cpp:84)
    r.s = (double)obs.yi - obs.aux * lambda;
    r.yWy = dpois_log((double)obs.yi, obs.aux * lambda);
    if (obs.yi < 0 || !(obs.aux >= 0) || !isfinite(eta)) atomicOr(err, 2);
  } else {
    r.w = obs.y; r.s = obs.aux;
  }
  if (out.w) out.w[i] = r.w;
  if (out.s) out.s[i] = r.s;
  return r;
}

template <int MODEL>
__device__ __forceinline__ RowLatent impute_row(const RowData &d, const DrawParams &prm, const RowOut &out, int64_t i, double eta,
                                                int *err) {
  return impute_row<MODEL>(d, prm, out, load_obs<MODEL>(d, i), i, eta, err);
}

// =============================================================================================
// Fused single-pass kernel, p <= 64.
//   NB  = ceil(p / 8) column blocks (8 x 8 DMMA atoms), NA = NB (NB+1) / 2 upper atoms per warp.
//   R   = rows per chunk (256 when NB <= 4, else 128); thread r < R draws row r of the chunk.
// Shared memory row stride LDS = 8 NB + 4 doubles: LDS = 4 (mod 8) makes the DMMA fragment loads
// (4 rows x 8 columns per instruction) conflict free; the eta loop rotates its start column by
// (r >> 2) & 3 to be conflict free with the same stride.
// Per-CTA partial: [P8*P8 tile | P8 xty | 8 scalars], reduced by reduce_partials_kernel (fused_tma.cuh).
// =============================================================================================
constexpr int kSmallThreads = 256;
__host__ __device__ constexpr int small_rows(int nb) { return nb <= 4 ? 256 : 128; }
__host__ __device__ constexpr int small_lds(int nb) { return 8 * nb + 4; }
__host__ __device__ constexpr size_t small_smem_bytes(int nb) {
  return sizeof(double) * (2 * (size_t)small_rows(nb) * small_lds(nb) + 2 * small_rows(nb) + 8 * nb + 64);
}
__host__ __device__ constexpr int64_t small_partial_len(int nb) { return 64 * nb * nb + 8 * nb + 8; }

template <int NB, int MODEL>
__global__ void __launch_bounds__(kSmallThreads, (NB <= 3) ? 2 : 1)
fused_small_kernel(RowData d, DrawParams prm, RowOut out, const double *__restrict__ beta, double *__restrict__ partials,
                   int *err, uint32_t div_magic, int vec2) {
  constexpr int R = small_rows(NB);
  constexpr int LDS = small_lds(NB);
  constexpr int P8 = 8 * NB;
  constexpr int NA = NB * (NB + 1) / 2;
  constexpr int RW = R / 8;  // rows per warp in the SYRK phase
  extern __shared__ __align__(128) double smem[];
  double *xs0 = smem;
  double *xs1 = smem + R * LDS;
  double *w_s = smem + 2 * R * LDS;
  double *s_s = w_s + R;
  double *beta_s = s_s + R;
  double *red_s = beta_s + P8;  // 64 doubles of scratch

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int p = d.p;
  const int64_t nchunks = (d.n + R - 1) / R;

  // zero both stages once (pad columns are never written by the copies) and stage beta
  for (int e = tid; e < 2 * R * LDS; e += kSmallThreads) smem[e] = 0.0;
  if (tid < P8) beta_s[tid] = tid < p ? beta[tid] : 0.0;
  __syncthreads();

  auto load_chunk = [&](int64_t chunk, double *xs) {
    const int64_t row0 = chunk * R;
    const int64_t valid_rows = min((int64_t)R, d.n - row0);
    const double *src0 = d.X + row0 * d.ldx;
    if (vec2) {
      const int total = R * p / 2;
      for (int q = tid; q < total; q += kSmallThreads) {
        int e = 2 * q;
        int i = __umulhi((uint32_t)e, div_magic);   // vec2 implies p >= 2
        int j = e - i * p;
        double *dst = xs + i * LDS + j;
        if (i < valid_rows) cp_async16(dst, src0 + (int64_t)i * d.ldx + j);
        else { dst[0] = 0.0; dst[1] = 0.0; }
      }
    } else {
      const int total = R * p;
      for (int e = tid; e < total; e += kSmallThreads) {
        int i = p == 1 ? e : __umulhi((uint32_t)e, div_magic);   // 2^32 / 1 + 1 does not fit the magic
        int j = e - i * p;
        double *dst = xs + i * LDS + j;
        if (i < valid_rows) cp_async8(dst, src0 + (int64_t)i * d.ldx + j);
        else dst[0] = 0.0;
      }
    }
    cp_async_commit();
  };

  double c[NA][2];
#pragma unroll
  for (int a = 0; a < NA; ++a) { c[a][0] = 0.0; c[a][1] = 0.0; }
  double xty_acc = 0.0;
  double sc_count = 0, sc_ywy = 0, sc_sumw = 0, sc_sumlogw = 0;

  // xty mapping: thread -> (column j, row group g)
  constexpr int G = kSmallThreads / P8;
  const int xj = tid % P8, xg = tid / P8;

  int64_t chunk = blockIdx.x;
  int stage = 0;
  if (chunk < nchunks) load_chunk(chunk, xs0);
  for (; chunk < nchunks; chunk += gridDim.x) {
    double *xs = stage ? xs1 : xs0;
    cp_async_wait_all();
    __syncthreads();
    if (chunk + gridDim.x < nchunks) load_chunk(chunk + gridDim.x, stage ? xs0 : xs1);

    // ---- draw phase: thread r owns row r of the chunk
    if (tid < R) {
      const int64_t i = chunk * R + tid;
      double wv = 0, sv = 0;
      if (i < d.n) {
        const double *xr = xs + tid * LDS;
        const int rot = (tid >> 2) & 3;
        double eta = 0;
#pragma unroll 8
        for (int j = 0; j < P8; ++j) {
          int jj = j + rot;
          jj = jj >= P8 ? jj - P8 : jj;
          eta = fma(xr[jj], beta_s[jj], eta);
        }
        RowLatent r = impute_row<MODEL>(d, prm, out, i, eta, err);
        wv = r.w; sv = r.s;
        sc_count += r.count; sc_ywy += r.yWy; sc_sumw += r.w; sc_sumlogw += r.sumlogw;
      }
      w_s[tid] = wv;
      s_s[tid] = sv;
    }
    __syncthreads();

    // ---- weighted SYRK on DMMA: warp wid owns rows [wid*RW, wid*RW + RW) of the chunk
#pragma unroll 2
    for (int kk = 0; kk < RW / 4; ++kk) {
      const int row = wid * RW + kk * 4 + (lane & 3);
      const double wv = w_s[row];
      const double *xr = xs + row * LDS + (lane >> 2);
      double xa[NB], xw[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) { xa[b] = xr[8 * b]; xw[b] = xa[b] * wv; }
      int a = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj) { dmma884(c[a][0], c[a][1], xw[bi], xa[bj]); ++a; }
    }
    // ---- X's: thread (xj, xg) sums rows xg, xg+G, ...
    if (xg < G) {
#pragma unroll 4
      for (int r = xg; r < R; r += G) xty_acc = fma(s_s[r], xs[r * LDS + xj], xty_acc);
    }
    stage ^= 1;
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- CTA reduction in a fixed order (deterministic), then one partial per CTA
  double *tile = smem;  // P8*P8 doubles, reuse stage memory
  for (int e = tid; e < P8 * P8; e += kSmallThreads) tile[e] = 0.0;
  double *xty_s = smem + P8 * P8;  // G * P8 doubles
  if (xg < G) xty_s[xg * P8 + xj] = xty_acc;
  __syncthreads();
  for (int w = 0; w < 8; ++w) {
    if (wid == w) {
      int a = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj) {
          double *t = tile + (8 * bi + (lane >> 2)) * P8 + 8 * bj + 2 * (lane & 3);
          t[0] += c[a][0];
          t[1] += c[a][1];
          ++a;
        }
    }
    __syncthreads();
  }
  double *my = partials + (int64_t)blockIdx.x * small_partial_len(NB);
  for (int e = tid; e < P8 * P8; e += kSmallThreads) my[e] = tile[e];
  if (tid < P8) {
    double s = 0;
    for (int g = 0; g < G; ++g) s += xty_s[g * P8 + tid];
    my[P8 * P8 + tid] = s;
  }
  // scalars: warp shuffle then the 8 warp values in order
  double v0 = warp_sum(sc_count), v1 = warp_sum(sc_ywy), v2 = warp_sum(sc_sumw), v3 = warp_sum(sc_sumlogw);
  if (lane == 0) { red_s[wid * 4 + 0] = v0; red_s[wid * 4 + 1] = v1; red_s[wid * 4 + 2] = v2; red_s[wid * 4 + 3] = v3; }
  __syncthreads();
  if (tid < 4) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += red_s[w * 4 + tid];
    my[P8 * P8 + P8 + tid] = s;
  }
}

// =============================================================================================
// Pass 1 for p > 64: a warp owns 32 consecutive rows; eta by a warp-cooperative dot product
// with coalesced 16-byte loads, then lane r draws row r.  Writes w_i and s_i (n doubles each)
// and per-CTA scalar partials.
// =============================================================================================
constexpr int kImputeThreads = 256;

template <int MODEL, bool VEC2>
__global__ void __launch_bounds__(kImputeThreads)
impute_rows_kernel(RowData d, DrawParams prm, RowOut out, const double *__restrict__ beta, double *__restrict__ w_buf,
                   double *__restrict__ s_buf, double *__restrict__ scalar_partials, int *err) {
  extern __shared__ __align__(128) double smem[];
  double *beta_s = smem;  // p (+1) doubles
  __shared__ double red_s[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int p = d.p;
  for (int j = tid; j < p + 1; j += kImputeThreads) beta_s[j] = j < p ? beta[j] : 0.0;
  __syncthreads();

  double sc_count = 0, sc_ywy = 0, sc_sumw = 0, sc_sumlogw = 0;
  const int64_t ngroups = (d.n + 31) / 32;
  const int warps_per_grid = gridDim.x * (kImputeThreads / 32);
  for (int64_t g = (int64_t)blockIdx.x * (kImputeThreads / 32) + wid; g < ngroups; g += warps_per_grid) {
    const int64_t row0 = g * 32;
    double my_eta = 0;
    const int nrows = (int)min((int64_t)32, d.n - row0);
    for (int r0 = 0; r0 < nrows; r0 += 4) {
      double acc[4] = {0, 0, 0, 0};
      const double *xr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xr[u] = d.X + (row0 + min(r0 + u, nrows - 1)) * d.ldx;  // clamped rows are discarded
      if (VEC2) {
        const double2 *b2 = reinterpret_cast<const double2 *>(beta_s);
        const int p2 = p >> 1;
#pragma unroll 2
        for (int j = lane; j < p2; j += 32) {
          double2 xv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) xv[u] = __ldg(reinterpret_cast<const double2 *>(xr[u]) + j);
          const double2 bv = b2[j];
#pragma unroll
          for (int u = 0; u < 4; ++u) { acc[u] = fma(xv[u].x, bv.x, acc[u]); acc[u] = fma(xv[u].y, bv.y, acc[u]); }
        }
        if ((p & 1) && lane == 0) {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fma(__ldg(xr[u] + p - 1), beta_s[p - 1], acc[u]);
        }
      } else {
#pragma unroll 2
        for (int j = lane; j < p; j += 32) {
          const double bv = beta_s[j];
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fma(__ldg(xr[u] + j), bv, acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double e = warp_sum(acc[u]);
        if (lane == r0 + u) my_eta = e;
      }
    }
    const int64_t i = row0 + lane;
    double wv = 0, sv = 0;
    if (i < d.n) {
      RowLatent r = impute_row<MODEL>(d, prm, out, i, my_eta, err);
      wv = r.w; sv = r.s;
      sc_count += r.count; sc_ywy += r.yWy; sc_sumw += r.w; sc_sumlogw += r.sumlogw;
      w_buf[i] = wv;
      s_buf[i] = sv;
    }
  }
  double v0 = warp_sum(sc_count), v1 = warp_sum(sc_ywy), v2 = warp_sum(sc_sumw), v3 = warp_sum(sc_sumlogw);
  if (lane == 0) { red_s[wid * 4 + 0] = v0; red_s[wid * 4 + 1] = v1; red_s[wid * 4 + 2] = v2; red_s[wid * 4 + 3] = v3; }
  __syncthreads();
  if (tid < 4) {
    double s = 0;
    for (int w = 0; w < kImputeThreads / 32; ++w) s += red_s[w * 4 + tid];
    scalar_partials[(int64_t)blockIdx.x * 4 + tid] = s;
  }
}

// Pass 1 when beta is sparse (spike-and-slab: beta is EXACTLY zero off the included set, GlmCoefs.cpp:304-309, and the
// included set is ~20 of 500 columns at C3): lane r owns row r of the warp's 32 rows and reads only the included columns --
// one 32-byte sector per column and row (fewer when included columns are neighbours: the sector stays in L1) instead of
// the whole 8 p byte row.  eta is the same sum with the zero terms left out.  nnz indices / values sit in shared memory.
template <int MODEL>
__global__ void __launch_bounds__(kImputeThreads)
impute_rows_gather_kernel(RowData d, DrawParams prm, RowOut out, const double *__restrict__ beta, const int *__restrict__ nz_idx, int nnz,
                          double *__restrict__ w_buf, double *__restrict__ s_buf, double *__restrict__ scalar_partials, int *err) {
  extern __shared__ __align__(128) double smem[];
  double *b_s = smem;                                   // nnz values
  int *i_s = reinterpret_cast<int *>(smem + nnz);       // nnz column indices
  __shared__ double red_s[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int j = tid; j < nnz; j += kImputeThreads) { const int c = nz_idx[j]; i_s[j] = c; b_s[j] = beta[c]; }
  __syncthreads();

  double sc_count = 0, sc_ywy = 0, sc_sumw = 0, sc_sumlogw = 0;
  const int64_t ngroups = (d.n + 31) / 32;
  const int warps_per_grid = gridDim.x * (kImputeThreads / 32);
  for (int64_t g = (int64_t)blockIdx.x * (kImputeThreads / 32) + wid; g < ngroups; g += warps_per_grid) {
    const int64_t i = g * 32 + lane;
    if (i < d.n) {
      const RowObs obs = load_obs<MODEL>(d, i);
      const double *xr = d.X + i * d.ldx;
      double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
      int j = 0;
      for (; j + 4 <= nnz; j += 4) {
        const double x0 = __ldg(xr + i_s[j]), x1 = __ldg(xr + i_s[j + 1]), x2 = __ldg(xr + i_s[j + 2]), x3 = __ldg(xr + i_s[j + 3]);
        e0 = fma(x0, b_s[j], e0); e1 = fma(x1, b_s[j + 1], e1); e2 = fma(x2, b_s[j + 2], e2); e3 = fma(x3, b_s[j + 3], e3);
      }
      for (; j < nnz; ++j) e0 = fma(__ldg(xr + i_s[j]), b_s[j], e0);
      const double eta = (e0 + e1) + (e2 + e3);
      RowLatent r = impute_row<MODEL>(d, prm, out, obs, i, eta, err);
      sc_count += r.count; sc_ywy += r.yWy; sc_sumw += r.w; sc_sumlogw += r.sumlogw;
      w_buf[i] = r.w;
      s_buf[i] = r.s;
    }
  }
  double v0 = warp_sum(sc_count), v1 = warp_sum(sc_ywy), v2 = warp_sum(sc_sumw), v3 = warp_sum(sc_sumlogw);
  if (lane == 0) { red_s[wid * 4 + 0] = v0; red_s[wid * 4 + 1] = v1; red_s[wid * 4 + 2] = v2; red_s[wid * 4 + 3] = v3; }
  __syncthreads();
  if (tid < 4) {
    double s = 0;
    for (int w = 0; w < kImputeThreads / 32; ++w) s += red_s[w * 4 + tid];
    scalar_partials[(int64_t)blockIdx.x * 4 + tid] = s;
  }
}

// =============================================================================================
// Pass 2 for p > 64: split-K weighted SYRK on FP64 DMMA.
//
//   Output regions are 128 x 128 blocks (I <= J) of the upper triangle; a region is cut into
//   32 x 32 "units" of 4 x 4 DMMA atoms.  A CTA = 8 consumer warps (two units each) + 1 producer
//   warp, and handles one (k-slice, region) pair: its rows are streamed through a 6-stage ring of
//   KB = 16 row tiles written by TMA bulk copies (one 1-D cp.async.bulk per row and panel, row
//   stride 132 doubles = 4 (mod 8), so fragment loads are conflict free), full/empty mbarriers.
//   Diagonal regions only compute the 10 units on or above the diagonal (and the diagonal units
//   skip their 6 lower atoms) and also produce X's for their 128 columns: one extra DMMA per
//   A fragment with B = [s, 0, ..., 0].
//   Per-CTA partial: 128 x 128 tile (+ 128 xty for diagonal regions), reduced by
//   reduce_syrk_kernel in k-slice order (deterministic).
// =============================================================================================
// Consumer layout: 8 warps x 2 units (default) or 16 warps x 1 unit, plus a producer warpgroup that hands its registers
// over with setmaxnreg.  Measured (profiles/README.md): 16 x 1 is no faster at p = 500 (8.38 vs 8.32 ms per 1 M rows) and
// slower at p = 4000 (254 vs 239 ms per 0.5 M rows: one more fragment load per DMMA); repeating the arithmetic of every
// stage on the same data scales the time linearly, so the kernel is bound by its own DMMA / LDS / DMUL stream (tensor
// pipe 89 % active), not by the supply of tiles.
#ifndef BOOMGPU_SYRK_WARPS
#define BOOMGPU_SYRK_WARPS 8
#endif
constexpr int kSyrkConsumerWarps = BOOMGPU_SYRK_WARPS;
constexpr int kSyrkProducerWarps = 4;   // a full warpgroup, so that setmaxnreg can move its registers to the consumers
constexpr int kSyrkThreads = 32 * (kSyrkConsumerWarps + kSyrkProducerWarps);
#ifndef BOOMGPU_SYRK_KB
#define BOOMGPU_SYRK_KB 16
#endif
#ifndef BOOMGPU_SYRK_STAGES
#define BOOMGPU_SYRK_STAGES 6
#endif
constexpr int kSyrkKB = BOOMGPU_SYRK_KB;        // rows per stage
constexpr int kSyrkStages = BOOMGPU_SYRK_STAGES;
constexpr int kSyrkPanelLd = 132;  // doubles
constexpr int kSyrkStageDoubles = 2 * kSyrkKB * kSyrkPanelLd + 2 * kSyrkKB;  // panels A, B, then w[KB], s[KB]
constexpr size_t kSyrkSmemBytes = sizeof(double) * kSyrkStages * kSyrkStageDoubles + 8 * 2 * kSyrkStages + 64;
constexpr int64_t kSyrkTileLen = 128 * 128 + 128;

struct SyrkUnit { int8_t ui, uj, flags; };  // flags: 1 valid, 2 diagonal unit, 4 xty duty, 8 xty duty when unit (ui, ui+1) is cut off by p
// [region type: 0 off-diagonal, 1 diagonal, 2 / 3 the same when the region's LAST unit column is ragged (p mod 128 in
// (96, 128): the p = 500 case), where the work is dealt so that every SM sub-partition gets its share of the cut units]
struct SyrkUnitTable { SyrkUnit u[4][kSyrkConsumerWarps][2]; };

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D TMA tile load: box of the tensor map at (c0 = column, c1 = row) -> shared memory, completes on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

struct SyrkItem { int16_t I, J, kind; int16_t pad; };   // kind 0: off-diagonal region (I, J); 1: diagonal region (I, I); 2: the pair (I, I), (J, J)
struct SyrkParams {
  const double *X;
  int64_t ldx;
  int64_t n;
  int p;
  const double *w;   // n rounded up to kSyrkKB, zero padded
  const double *s;
  int nblk;          // ceil(p / 128)
  int nregions;      // nblk (nblk + 1) / 2
  int ksplit;
  int64_t rows_per_slice;  // multiple of kSyrkKB
  double *partials;  // [ksplit][nregions][kSyrkTileLen]
  int diag_form;     // option "syrk_diag": 0 strip form for whole diagonal regions (default), 1 unit form everywhere
  int filter;        // profiling aid (option "syrk_filter"): 0 all regions, 1 off-diagonal regions only, 2 diagonal only
  int order;         // option "syrk_order": 1 off-diagonal regions first, diagonal regions last, 0 k-slice major,
                     // 2 k-slice major over UNIFORM work items (off-diagonal regions + PAIRS of diagonal regions), see SyrkItem
  int skip_ragged_diag;           // bit 0: the ragged last diagonal region, bit 1: the off-diagonal regions of the ragged last column
                                  // block are computed by syrk_rdiag_kernel: their CTAs of this grid exit
  const SyrkItem *items;          // order 2: the work items of one k-slice, in launch order (device memory)
  int nitems;
};

// Order 2.  A diagonal region costs about half an off-diagonal one, which is why order 1 schedules them last -- but then the
// diagonal CTAs re-read every panel long after the off-diagonal CTAs of the same rows went through, and even perfect sharing
// in L2 cannot get below 2 x X of DRAM traffic (measured: 2.6 x).  Here TWO diagonal regions (I, I) and (I + 1, I + 1) are
// one work item: it loads the same two panels as the off-diagonal region (I, I + 1) and computes 2 x 136 atoms against 256,
// so all CTAs of a k-slice cost the same, are handed out together (k-slice major) and walk the same rows at the same pace:
// a panel tile is fetched from DRAM by the first CTA that needs it and found in L2 by the others.


// CTA -> (k-slice, column blocks I <= J).  The hardware hands CTAs to the SMs in blockIdx order as they free up, so the END of
// the grid decides how long the last SMs idle: a diagonal region costs about half an off-diagonal one (136 of 256 atoms), so
// all off-diagonal work (k-slice major: the regions of one k-slice run together and share their panels in L2) goes first and
// the cheap diagonal CTAs fill the tail (ncu, C3: SM-idle share of the launch 3.0 % with the k-slice major order).
__device__ __forceinline__ void syrk_block_to_work(int b, const SyrkParams &prm, int &kslice, int &I, int &J, int &kind) {
  const int nblk = prm.nblk;
  if (prm.order == 2) {
    kslice = b / prm.nitems;
    const SyrkItem it = prm.items[b - kslice * prm.nitems];
    I = it.I; J = it.J; kind = it.kind;
    return;
  }
  kind = -1;
  if (prm.order == 0) {
    kslice = b / prm.nregions;
    int i = 0, rem = b - kslice * prm.nregions;
    while (rem >= nblk - i) { rem -= nblk - i; ++i; }
    I = i; J = i + rem;
    return;
  }
  const int noff = prm.nregions - nblk;
  const int off_total = prm.ksplit * noff;
  if (b < off_total) {
    kslice = b / noff;
    int i = 0, rem = b - kslice * noff;
    while (rem >= nblk - 1 - i) { rem -= nblk - 1 - i; ++i; }
    I = i; J = i + 1 + rem;
  } else {
    b -= off_total;
    kslice = b / nblk;
    I = J = b - kslice * nblk;
  }
}

__device__ __forceinline__ void region_to_blocks(int region, int nblk, int &I, int &J) {
  // regions enumerated row by row over the upper triangle: (0,0),(0,1)...(0,nblk-1),(1,1),...
  int i = 0, rem = region;
  while (rem >= nblk - i) { rem -= nblk - i; ++i; }
  I = i; J = i + rem;
}

struct SyrkWarpCtx {
  double *smem;
  uint64_t *full_bar, *empty_bar;
  int nstages_total, lane, panelB_off;
  SyrkUnit u0, u1;
  int mmax0, nmax0, mmax1, nmax1;
  double *tile;
};

// The k loop of one consumer warp.  T0/T1: unit type (0 none, 1 full 4x4 atoms, 2 diagonal unit: the 10
// atoms with n >= m); DUTY: which unit (1 or 2; 0 none) also accumulates X's for its row block.
// Two full units of one warp always share their A fragments (same row block).
// NM (1..4): how many 8-column atoms of the warp's LAST unit lie inside X.  p that is not a multiple of 32 leaves the last
// unit column ragged (p = 500: 3 of 4 atoms); only a warp's last unit can be ragged (the unit table lists a warp's units
// left to right), a full unit is then cut in n only, a diagonal unit in m and n.  Compile-time so that the cut atoms are
// neither loaded nor multiplied (they used to be computed and masked at the store: 3 % of the DMMAs at p = 500).
template <int T0, int T1, int DUTY, int NM>
__device__ __forceinline__ void syrk_consume(const SyrkWarpCtx &wc) {
  constexpr int M0 = (T1 == 0 && T0 == 2) ? NM : 4, N0 = (T1 == 0) ? NM : 4;   // atom limits of unit 0
  constexpr int M1 = (T1 == 2) ? NM : 4, N1 = NM;                               // ... of unit 1
  const int lane = wc.lane;
  double c0[4][4][2], c1[4][4][2], cx[4][2];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
#pragma unroll
    for (int n = 0; n < 4; ++n) { c0[m][n][0] = c0[m][n][1] = 0.0; c1[m][n][0] = c1[m][n][1] = 0.0; }
    cx[m][0] = cx[m][1] = 0.0;
  }
  const int a0_off = 32 * wc.u0.ui + (lane >> 2), b0_off = wc.panelB_off + 32 * wc.u0.uj + (lane >> 2);
  const int a1_off = 32 * wc.u1.ui + (lane >> 2), b1_off = wc.panelB_off + 32 * wc.u1.uj + (lane >> 2);
  constexpr bool kSameA = (T0 == 1 && T1 == 1);

  for (int it = 0; it < wc.nstages_total; ++it) {
    const int s = it % kSyrkStages;
    const uint32_t phase = (it / kSyrkStages) & 1;
    mbar_wait(wc.full_bar + s, phase);
    const double *stage = wc.smem + s * kSyrkStageDoubles;
    const double *w_s = stage + 2 * kSyrkKB * kSyrkPanelLd;
    const double *s_s = w_s + kSyrkKB;
    if (T0 != 0) {
#pragma unroll
      for (int kk = 0; kk < kSyrkKB / 4; ++kk) {
        const int row = kk * 4 + (lane & 3);
        const double wv = w_s[row];
        const double *xr = stage + row * kSyrkPanelLd;
        double a[4], aw[4], b[4];
#pragma unroll
        for (int m = 0; m < M0; ++m) { a[m] = xr[a0_off + 8 * m]; aw[m] = a[m] * wv; }
#pragma unroll
        for (int n = 0; n < N0; ++n) b[n] = xr[b0_off + 8 * n];
#pragma unroll
        for (int m = 0; m < M0; ++m)
#pragma unroll
          for (int n = 0; n < N0; ++n)
            if (T0 == 1 || n >= m) dmma884(c0[m][n][0], c0[m][n][1], aw[m], b[n]);
        if (DUTY == 1) {  // X's: four DFMA on the A fragments already in registers (a DMMA with B = [s, 0, ..] wasted 7/8 of it)
          const double sv = s_s[row];
#pragma unroll
          for (int m = 0; m < M0; ++m) cx[m][0] = fma(a[m], sv, cx[m][0]);
        }
        if (T1 != 0) {
          if (!kSameA) {
#pragma unroll
            for (int m = 0; m < M1; ++m) { a[m] = xr[a1_off + 8 * m]; aw[m] = a[m] * wv; }
          }
#pragma unroll
          for (int n = 0; n < N1; ++n) b[n] = xr[b1_off + 8 * n];
#pragma unroll
          for (int m = 0; m < M1; ++m)
#pragma unroll
            for (int n = 0; n < N1; ++n)
              if (T1 == 1 || n >= m) dmma884(c1[m][n][0], c1[m][n][1], aw[m], b[n]);
          if (DUTY == 2) {
            const double sv = s_s[row];
#pragma unroll
            for (int m = 0; m < M1; ++m) cx[m][0] = fma(a[m], sv, cx[m][0]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(wc.empty_bar + s);
  }

  // ===== epilogue: fragments -> this CTA's partial tile (row major 128 x 128, then 128 xty)
  double *tile = wc.tile;
  auto store_unit = [&](const SyrkUnit &u, double (&cc)[4][4][2], bool dg, int mmax, int nmax) {
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        if ((dg && n < m) || m >= mmax || n >= nmax) continue;
        const int r = 32 * u.ui + 8 * m + (lane >> 2);
        const int cidx = 32 * u.uj + 8 * n + 2 * (lane & 3);
        *reinterpret_cast<double2 *>(tile + r * 128 + cidx) = make_double2(cc[m][n][0], cc[m][n][1]);
      }
  };
  if (T0 != 0) store_unit(wc.u0, c0, T0 == 2, wc.mmax0, wc.nmax0);
  if (T1 != 0) store_unit(wc.u1, c1, T1 == 2, wc.mmax1, wc.nmax1);
  if (DUTY != 0) {
    const int ui = DUTY == 1 ? wc.u0.ui : wc.u1.ui;
    const int mmax = DUTY == 1 ? wc.mmax0 : wc.mmax1;
    // lane holds the rows = lane & 3 (mod 4) of column 32 ui + 8 m + (lane >> 2): sum the four row classes
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      cx[m][0] += __shfl_xor_sync(0xffffffffu, cx[m][0], 1);
      cx[m][0] += __shfl_xor_sync(0xffffffffu, cx[m][0], 2);
    }
    if ((lane & 3) == 0) {
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (m < mmax) tile[128 * 128 + 32 * ui + 8 * m + (lane >> 2)] = cx[m][0];
    }
  }
}

// The k loop of one consumer warp in a WHOLE diagonal region (all 128 columns inside X), strip form.  In 8 x 8 atoms the
// region is a 16 x 16 grid of which the 136 atoms on or above the diagonal are wanted; warp W owns atom rows W and 15 - W:
// (16 - W) + (W + 1) = 17 atoms for every warp, so the four SM sub-partitions carry 34 DMMA per k-step each (the unit
// form deals 36 / 32 / 32 / 36 and leaves the two-diagonal-unit warps on the critical path with 16 fragment loads and
// 8 DMUL per 20 DMMA).  The A fragment of atom row r is the B fragment of atom column r, so a k-step loads the B fragments
// of columns W .. 15 once and nothing else; X's for atom rows W and 15 - W rides on the same fragments (two DFMA).
template <int W>
__device__ __forceinline__ void syrk_consume_strip(const SyrkWarpCtx &wc) {
  constexpr int R1 = W, R2 = 15 - W, N1 = 16 - R1, N2 = 16 - R2;
  const int lane = wc.lane;
  double c1[N1][2], c2[N2][2], cx1 = 0.0, cx2 = 0.0;
#pragma unroll
  for (int n = 0; n < N1; ++n) c1[n][0] = c1[n][1] = 0.0;
#pragma unroll
  for (int n = 0; n < N2; ++n) c2[n][0] = c2[n][1] = 0.0;
  const int off = lane >> 2;
  for (int it = 0; it < wc.nstages_total; ++it) {
    const int s = it % kSyrkStages;
    const uint32_t phase = (it / kSyrkStages) & 1;
    mbar_wait(wc.full_bar + s, phase);
    const double *stage = wc.smem + s * kSyrkStageDoubles;
    const double *w_s = stage + 2 * kSyrkKB * kSyrkPanelLd;
    const double *s_s = w_s + kSyrkKB;
#pragma unroll
    for (int kk = 0; kk < kSyrkKB / 4; ++kk) {
      const int row = kk * 4 + (lane & 3);
      const double wv = w_s[row], sv = s_s[row];
      const double *xr = stage + row * kSyrkPanelLd + off;
      double b[N1];
#pragma unroll
      for (int n = 0; n < N1; ++n) b[n] = xr[8 * (R1 + n)];
      const double a1 = b[0] * wv, a2 = b[R2 - R1] * wv;
#pragma unroll
      for (int n = 0; n < N1; ++n) dmma884(c1[n][0], c1[n][1], a1, b[n]);
#pragma unroll
      for (int n = 0; n < N2; ++n) dmma884(c2[n][0], c2[n][1], a2, b[R2 - R1 + n]);
      cx1 = fma(b[0], sv, cx1);
      cx2 = fma(b[R2 - R1], sv, cx2);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(wc.empty_bar + s);
  }
  double *tile = wc.tile;
#pragma unroll
  for (int n = 0; n < N1; ++n)
    *reinterpret_cast<double2 *>(tile + (8 * R1 + off) * 128 + 8 * (R1 + n) + 2 * (lane & 3)) = make_double2(c1[n][0], c1[n][1]);
#pragma unroll
  for (int n = 0; n < N2; ++n)
    *reinterpret_cast<double2 *>(tile + (8 * R2 + off) * 128 + 8 * (R2 + n) + 2 * (lane & 3)) = make_double2(c2[n][0], c2[n][1]);
  cx1 += __shfl_xor_sync(0xffffffffu, cx1, 1); cx1 += __shfl_xor_sync(0xffffffffu, cx1, 2);
  cx2 += __shfl_xor_sync(0xffffffffu, cx2, 1); cx2 += __shfl_xor_sync(0xffffffffu, cx2, 2);
  if ((lane & 3) == 0) {
    tile[128 * 128 + 8 * R1 + off] = cx1;
    tile[128 * 128 + 8 * R2 + off] = cx2;
  }
}

// Two diagonal regions in one CTA (work item kind 2): the strip form on panel A for region (I, I) and on panel B for region
// (J, J), k-step by k-step on the same stage.  68 DMMA per k-step and SM sub-partition against the 64 of an off-diagonal
// region.  A ragged last region (p = 500: 116 of 128 columns) runs through the same code: TMA delivers zeros beyond p, the
// 23 atoms they make (1 % of a k-slice's work) are computed and never read by the reduction.
template <int W>
__device__ __forceinline__ void syrk_consume_strip2(const SyrkWarpCtx &wc, double *tile2) {
  constexpr int R1 = W, R2 = 15 - W, N1 = 16 - R1, N2 = 16 - R2;
  const int lane = wc.lane;
  double cA1[N1][2], cA2[N2][2], cB1[N1][2], cB2[N2][2], cxA1 = 0.0, cxA2 = 0.0, cxB1 = 0.0, cxB2 = 0.0;
#pragma unroll
  for (int n = 0; n < N1; ++n) cA1[n][0] = cA1[n][1] = cB1[n][0] = cB1[n][1] = 0.0;
#pragma unroll
  for (int n = 0; n < N2; ++n) cA2[n][0] = cA2[n][1] = cB2[n][0] = cB2[n][1] = 0.0;
  const int off = lane >> 2;
  for (int it = 0; it < wc.nstages_total; ++it) {
    const int s = it % kSyrkStages;
    const uint32_t phase = (it / kSyrkStages) & 1;
    mbar_wait(wc.full_bar + s, phase);
    const double *stage = wc.smem + s * kSyrkStageDoubles;
    const double *w_s = stage + 2 * kSyrkKB * kSyrkPanelLd;
    const double *s_s = w_s + kSyrkKB;
#pragma unroll
    for (int kk = 0; kk < kSyrkKB / 4; ++kk) {
      const int row = kk * 4 + (lane & 3);
      const double wv = w_s[row], sv = s_s[row];
      {
        const double *xr = stage + row * kSyrkPanelLd + off;
        double b[N1];
#pragma unroll
        for (int n = 0; n < N1; ++n) b[n] = xr[8 * (R1 + n)];
        const double a1 = b[0] * wv, a2 = b[R2 - R1] * wv;
#pragma unroll
        for (int n = 0; n < N1; ++n) dmma884(cA1[n][0], cA1[n][1], a1, b[n]);
#pragma unroll
        for (int n = 0; n < N2; ++n) dmma884(cA2[n][0], cA2[n][1], a2, b[R2 - R1 + n]);
        cxA1 = fma(b[0], sv, cxA1);
        cxA2 = fma(b[R2 - R1], sv, cxA2);
      }
      {
        const double *xr = stage + kSyrkKB * kSyrkPanelLd + row * kSyrkPanelLd + off;
        double b[N1];
#pragma unroll
        for (int n = 0; n < N1; ++n) b[n] = xr[8 * (R1 + n)];
        const double a1 = b[0] * wv, a2 = b[R2 - R1] * wv;
#pragma unroll
        for (int n = 0; n < N1; ++n) dmma884(cB1[n][0], cB1[n][1], a1, b[n]);
#pragma unroll
        for (int n = 0; n < N2; ++n) dmma884(cB2[n][0], cB2[n][1], a2, b[R2 - R1 + n]);
        cxB1 = fma(b[0], sv, cxB1);
        cxB2 = fma(b[R2 - R1], sv, cxB2);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(wc.empty_bar + s);
  }
  auto store = [&](double *tile, double (&c1)[N1][2], double (&c2)[N2][2], double cx1, double cx2) {
#pragma unroll
    for (int n = 0; n < N1; ++n)
      *reinterpret_cast<double2 *>(tile + (8 * R1 + off) * 128 + 8 * (R1 + n) + 2 * (lane & 3)) = make_double2(c1[n][0], c1[n][1]);
#pragma unroll
    for (int n = 0; n < N2; ++n)
      *reinterpret_cast<double2 *>(tile + (8 * R2 + off) * 128 + 8 * (R2 + n) + 2 * (lane & 3)) = make_double2(c2[n][0], c2[n][1]);
    cx1 += __shfl_xor_sync(0xffffffffu, cx1, 1); cx1 += __shfl_xor_sync(0xffffffffu, cx1, 2);
    cx2 += __shfl_xor_sync(0xffffffffu, cx2, 1); cx2 += __shfl_xor_sync(0xffffffffu, cx2, 2);
    if ((lane & 3) == 0) {
      tile[128 * 128 + 8 * R1 + off] = cx1;
      tile[128 * 128 + 8 * R2 + off] = cx2;
    }
  };
  store(wc.tile, cA1, cA2, cxA1, cxA2);
  store(tile2, cB1, cB2, cxB1, cxB2);
}

template <int NM>
__device__ __forceinline__ void syrk_dispatch_nm(int role, const SyrkWarpCtx &wc) {
  if (kSyrkConsumerWarps == 8) {
    switch (role) {
      case 0:   syrk_consume<0, 0, 0, 4>(wc); break;
      case 100: syrk_consume<1, 0, 0, NM>(wc); break;
      case 101: syrk_consume<1, 0, 1, NM>(wc); break;
      case 110: syrk_consume<1, 1, 0, NM>(wc); break;
      case 200: syrk_consume<2, 0, 0, NM>(wc); break;
      case 201: syrk_consume<2, 0, 1, NM>(wc); break;
      case 220: syrk_consume<2, 2, 0, NM>(wc); break;
      case 221: syrk_consume<2, 2, 1, NM>(wc); break;
      case 222: syrk_consume<2, 2, 2, NM>(wc); break;
      default: __trap();
    }
  } else {   // one unit per warp: the two-unit roles are not instantiated (they would set the kernel's register count)
    switch (role) {
      case 0:   syrk_consume<0, 0, 0, 4>(wc); break;
      case 100: syrk_consume<1, 0, 0, NM>(wc); break;
      case 101: syrk_consume<1, 0, 1, NM>(wc); break;
      case 200: syrk_consume<2, 0, 0, NM>(wc); break;
      case 201: syrk_consume<2, 0, 1, NM>(wc); break;
      default: __trap();
    }
  }
}
__device__ __forceinline__ void syrk_dispatch_role(int role, int nm, const SyrkWarpCtx &wc) {
  switch (nm) {
    case 4: syrk_dispatch_nm<4>(role, wc); break;
    case 3: syrk_dispatch_nm<3>(role, wc); break;
    case 2: syrk_dispatch_nm<2>(role, wc); break;
    case 1: syrk_dispatch_nm<1>(role, wc); break;
    default: __trap();
  }
}

__global__ void __launch_bounds__(kSyrkThreads, 1)
syrk_dmma_kernel(const __grid_constant__ CUtensorMap xmap, SyrkParams prm, SyrkUnitTable table) {
  extern __shared__ __align__(128) double smem[];
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kSyrkStages * kSyrkStageDoubles);
  uint64_t *empty_bar = full_bar + kSyrkStages;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int kslice, I, J, kind;
  syrk_block_to_work((int)blockIdx.x, prm, kslice, I, J, kind);
  const bool pair = kind == 2;           // two diagonal regions (I, I) and (J, J) on the two panels of this CTA
  const int region = pair ? I * prm.nblk - I * (I - 1) / 2
                          : I * prm.nblk - I * (I - 1) / 2 + (J - I);   // row-by-row index over the upper triangle (region_to_blocks)
  const bool diag = (I == J);
  if (prm.filter && (prm.filter == 1) == diag) return;   // profiling aid: results are incomplete on purpose
  if (prm.skip_ragged_diag && !pair && J == prm.nblk - 1 && ((prm.p + 7) & ~7) - 128 * J < 128 &&
      (prm.skip_ragged_diag & (diag ? 1 : 2)))
    return;   // syrk_rdiag_kernel's regions
  const int64_t row_begin = (int64_t)kslice * prm.rows_per_slice;
  const int64_t row_end = min(prm.n, row_begin + prm.rows_per_slice);
  const int nstages_total = row_end > row_begin ? (int)((row_end - row_begin + kSyrkKB - 1) / kSyrkKB) : 0;

  if (tid == 0) {
    for (int s = 0; s < kSyrkStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, kSyrkConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // ===== producer role: ONE TMA tile per panel and stage.  The box is {132 columns, 16 rows}: four columns wider
  // than the panel, so the tile lands with the 132-double row pitch the fragment loads want (4 mod 8: conflict free)
  // without any per-row copy; columns beyond p and rows beyond n arrive as zeros.  (Per-row bulk copies kept the TMA
  // unit at ~34 requests per stage, one per ~55 cycles: 13 % of the consumers' samples were waits for data.)
  constexpr uint32_t kPanelBytes = kSyrkKB * kSyrkPanelLd * sizeof(double);
  const uint32_t stage_bytes = (diag ? 1u : 2u) * kPanelBytes + 2 * kSyrkKB * 8;
  auto produce = [&](int it) {
    const int s = it % kSyrkStages;
    const uint32_t phase = (it / kSyrkStages) & 1;
    mbar_wait(empty_bar + s, phase ^ 1);
    double *stage = smem + s * kSyrkStageDoubles;
    const int64_t r0 = row_begin + (int64_t)it * kSyrkKB;
    if (lane == 0) {
      mbar_expect_tx(full_bar + s, stage_bytes);
      tma_load_2d(stage, &xmap, 128 * I, (int)r0, full_bar + s);
      if (!diag) tma_load_2d(stage + kSyrkKB * kSyrkPanelLd, &xmap, 128 * J, (int)r0, full_bar + s);
      tma_bulk_g2s(stage + 2 * kSyrkKB * kSyrkPanelLd, prm.w + r0, kSyrkKB * 8, full_bar + s);
      tma_bulk_g2s(stage + 2 * kSyrkKB * kSyrkPanelLd + kSyrkKB, prm.s + r0, kSyrkKB * 8, full_bar + s);
    }
  };
  if (wid >= kSyrkConsumerWarps) {
    // producer warp(group): never competes for the DMMA pipe; as a full warpgroup it also hands its registers over
    if (kSyrkConsumerWarps == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 24;\n");
    if (wid == kSyrkConsumerWarps) {
      for (int it = 0; it < nstages_total; ++it) produce(it);
    }
    return;
  }
  if (kSyrkConsumerWarps == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
  // what the 16 consumer warps take must fit what the 4 producer warps gave back to the CTA pool:
  // 16 (112 - 96) <= 4 (96 - 24); asking for more (120) spins in USETMAXREG.TRY_ALLOC for ever
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");

  // ===== consumer role: resolved ONCE per warp into compile-time unit types so that the k loop
  // carries no predicates (a predicated mma.sync costs a WARPSYNC + branch per instruction).
  const int P8 = (prm.p + 7) & ~7;
  const int remJ = P8 - 128 * J;   // columns (whole 8-column atoms) of column block J inside X
  if (pair) {
    SyrkWarpCtx wc;
    wc.smem = smem; wc.full_bar = full_bar; wc.empty_bar = empty_bar; wc.nstages_total = nstages_total;
    wc.lane = lane;
    wc.panelB_off = kSyrkKB * kSyrkPanelLd;
    wc.tile = prm.partials + ((int64_t)kslice * prm.nregions + region) * kSyrkTileLen;
    double *tile2 = prm.partials + ((int64_t)kslice * prm.nregions + (J * prm.nblk - J * (J - 1) / 2)) * kSyrkTileLen;
    switch (wid) {
      case 0: syrk_consume_strip2<0>(wc, tile2); break;
      case 1: syrk_consume_strip2<1>(wc, tile2); break;
      case 2: syrk_consume_strip2<2>(wc, tile2); break;
      case 3: syrk_consume_strip2<3>(wc, tile2); break;
      case 4: syrk_consume_strip2<4>(wc, tile2); break;
      case 5: syrk_consume_strip2<5>(wc, tile2); break;
      case 6: syrk_consume_strip2<6>(wc, tile2); break;
      default: syrk_consume_strip2<7>(wc, tile2); break;
    }
    return;
  }
  if (kSyrkConsumerWarps == 8 && diag && remJ >= 128 && prm.diag_form == 0) {   // whole diagonal region: strip form
    SyrkWarpCtx wc;
    wc.smem = smem; wc.full_bar = full_bar; wc.empty_bar = empty_bar; wc.nstages_total = nstages_total;
    wc.lane = lane;
    wc.panelB_off = 0;
    wc.tile = prm.partials + ((int64_t)kslice * prm.nregions + region) * kSyrkTileLen;
    switch (wid) {
      case 0: syrk_consume_strip<0>(wc); break;
      case 1: syrk_consume_strip<1>(wc); break;
      case 2: syrk_consume_strip<2>(wc); break;
      case 3: syrk_consume_strip<3>(wc); break;
      case 4: syrk_consume_strip<4>(wc); break;
      case 5: syrk_consume_strip<5>(wc); break;
      case 6: syrk_consume_strip<6>(wc); break;
      default: syrk_consume_strip<7>(wc); break;
    }
    return;
  }
  const int ttype = (diag ? 1 : 0) + ((remJ > 96 && remJ < 128) ? 2 : 0);
  SyrkUnit u0 = table.u[ttype][wid][0];
  SyrkUnit u1 = table.u[ttype][wid][1];
  // 8-column atoms of each unit that lie inside X (<= 0: the whole unit is cut off by p)
  const int mmax0 = min(4, (P8 - 128 * I - 32 * u0.ui) / 8), nmax0 = min(4, (P8 - 128 * J - 32 * u0.uj) / 8);
  const int mmax1 = min(4, (P8 - 128 * I - 32 * u1.ui) / 8), nmax1 = min(4, (P8 - 128 * J - 32 * u1.uj) / 8);
  const bool v0 = (u0.flags & 1) && mmax0 > 0 && nmax0 > 0, v1 = (u1.flags & 1) && mmax1 > 0 && nmax1 > 0;
  // X's for row block ui: the full unit (ui, ui+1) carries it; when p cuts that unit off, the diagonal unit does
  const bool duty0 = v0 && ((u0.flags & 4) || ((u0.flags & 8) && P8 - 128 * I - 32 * (u0.ui + 1) <= 0));
  const bool duty1 = v1 && ((u1.flags & 4) || ((u1.flags & 8) && P8 - 128 * I - 32 * (u1.ui + 1) <= 0));
  const int t0 = v0 ? ((u0.flags & 2) ? 2 : 1) : 0;
  const int t1 = v1 ? ((u1.flags & 2) ? 2 : 1) : 0;   // the unit table lists units so that v1 implies v0
  const int duty = duty0 ? 1 : (duty1 ? 2 : 0);

  SyrkWarpCtx wc;
  wc.smem = smem; wc.full_bar = full_bar; wc.empty_bar = empty_bar; wc.nstages_total = nstages_total;
  wc.lane = lane;
  wc.panelB_off = diag ? 0 : kSyrkKB * kSyrkPanelLd;
  wc.u0 = u0; wc.u1 = u1;
  wc.mmax0 = mmax0; wc.nmax0 = nmax0; wc.mmax1 = mmax1; wc.nmax1 = nmax1;
  wc.tile = prm.partials + ((int64_t)kslice * prm.nregions + region) * kSyrkTileLen;

  // atoms of the warp's last unit inside X (1..4); earlier units of the warp are whole
  const int nm = t1 ? nmax1 : (t0 ? nmax0 : 4);
  syrk_dispatch_role(t0 * 100 + t1 * 10 + duty, nm, wc);
}

// =============================================================================================
// The RAGGED diagonal region (the last column block when p is not a multiple of 128; the only region when 64 < p < 128) in
// strip form over its A = ceil(cols / 8) < 16 atom columns.  The 128 x 128 machinery above deals units of 32 x 32 to 8 warps x 2
// slots, so a region with few columns keeps most warps idle while the CTA still walks its rows at the pace of a full region
// (p = 70 cost what p = 128 costs: 0.19 of the FP64 peak).  Here warp W owns atom rows W and A - 1 - W -- A + 1 atoms for every
// busy warp, nothing loaded or multiplied beyond column 8 A -- so the region's time follows its size.  Same ring, same producer,
// same partial-tile layout as syrk_dmma_kernel (one CTA per k-slice); the main grid's CTAs of this region exit at once.
// =============================================================================================
template <int W, int A>
__device__ __forceinline__ void syrk_consume_rstrip(const SyrkWarpCtx &wc) {
  constexpr int R1 = W, R2 = A - 1 - W;
  constexpr bool kBusy = R1 <= R2, kTwo = R1 < R2;
  constexpr int N1 = kBusy ? A - R1 : 1, N2 = kTwo ? A - R2 : 1;
  const int lane = wc.lane;
  double c1[N1][2], c2[N2][2], cx1 = 0.0, cx2 = 0.0;
#pragma unroll
  for (int n = 0; n < N1; ++n) c1[n][0] = c1[n][1] = 0.0;
#pragma unroll
  for (int n = 0; n < N2; ++n) c2[n][0] = c2[n][1] = 0.0;
  const int off = lane >> 2;
  for (int it = 0; it < wc.nstages_total; ++it) {
    const int s = it % kSyrkStages;
    const uint32_t phase = (it / kSyrkStages) & 1;
    mbar_wait(wc.full_bar + s, phase);
    if (kBusy) {
      const double *stage = wc.smem + s * kSyrkStageDoubles;
      const double *w_s = stage + 2 * kSyrkKB * kSyrkPanelLd;
      const double *s_s = w_s + kSyrkKB;
#pragma unroll
      for (int kk = 0; kk < kSyrkKB / 4; ++kk) {
        const int row = kk * 4 + (lane & 3);
        const double wv = w_s[row], sv = s_s[row];
        const double *xr = stage + row * kSyrkPanelLd + off;
        double b[N1];
#pragma unroll
        for (int n = 0; n < N1; ++n) b[n] = xr[8 * (R1 + n)];
        const double a1 = b[0] * wv;
#pragma unroll
        for (int n = 0; n < N1; ++n) dmma884(c1[n][0], c1[n][1], a1, b[n]);
        cx1 = fma(b[0], sv, cx1);
        if (kTwo) {
          const double a2 = b[R2 - R1] * wv;
#pragma unroll
          for (int n = 0; n < N2; ++n) dmma884(c2[n][0], c2[n][1], a2, b[R2 - R1 + n]);
          cx2 = fma(b[R2 - R1], sv, cx2);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(wc.empty_bar + s);
  }
  if (!kBusy) return;
  double *tile = wc.tile;
#pragma unroll
  for (int n = 0; n < N1; ++n)
    *reinterpret_cast<double2 *>(tile + (8 * R1 + off) * 128 + 8 * (R1 + n) + 2 * (lane & 3)) = make_double2(c1[n][0], c1[n][1]);
  cx1 += __shfl_xor_sync(0xffffffffu, cx1, 1); cx1 += __shfl_xor_sync(0xffffffffu, cx1, 2);
  if ((lane & 3) == 0) tile[128 * 128 + 8 * R1 + off] = cx1;
  if (kTwo) {
#pragma unroll
    for (int n = 0; n < N2; ++n)
      *reinterpret_cast<double2 *>(tile + (8 * R2 + off) * 128 + 8 * (R2 + n) + 2 * (lane & 3)) = make_double2(c2[n][0], c2[n][1]);
    cx2 += __shfl_xor_sync(0xffffffffu, cx2, 1); cx2 += __shfl_xor_sync(0xffffffffu, cx2, 2);
    if ((lane & 3) == 0) tile[128 * 128 + 8 * R2 + off] = cx2;
  }
}

// An off-diagonal region (I, last) whose column block has only AB < 16 atom columns: warp W owns atom rows 2 W and 2 W + 1 of the
// region and all AB columns -- 2 AB atoms for every warp whatever AB is (the unit form leaves warps idle and runs at the pace of a
// full region: 32 atoms on the busiest warp).
template <int AB>
__device__ __forceinline__ void syrk_consume_cstrip(const SyrkWarpCtx &wc, int W) {
  const int lane = wc.lane;
  double c0[AB][2], c1[AB][2];
#pragma unroll
  for (int n = 0; n < AB; ++n) c0[n][0] = c0[n][1] = c1[n][0] = c1[n][1] = 0.0;
  const int off = lane >> 2;
  for (int it = 0; it < wc.nstages_total; ++it) {
    const int s = it % kSyrkStages;
    const uint32_t phase = (it / kSyrkStages) & 1;
    mbar_wait(wc.full_bar + s, phase);
    const double *stage = wc.smem + s * kSyrkStageDoubles;
    const double *w_s = stage + 2 * kSyrkKB * kSyrkPanelLd;
#pragma unroll
    for (int kk = 0; kk < kSyrkKB / 4; ++kk) {
      const int row = kk * 4 + (lane & 3);
      const double wv = w_s[row];
      const double *xa = stage + row * kSyrkPanelLd + off + 16 * W;
      const double *xb = stage + kSyrkKB * kSyrkPanelLd + row * kSyrkPanelLd + off;
      const double a0 = xa[0] * wv, a1 = xa[8] * wv;
#pragma unroll
      for (int n = 0; n < AB; ++n) {
        const double b = xb[8 * n];
        dmma884(c0[n][0], c0[n][1], a0, b);
        dmma884(c1[n][0], c1[n][1], a1, b);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(wc.empty_bar + s);
  }
  double *tile = wc.tile;
#pragma unroll
  for (int n = 0; n < AB; ++n) {
    *reinterpret_cast<double2 *>(tile + (16 * W + off) * 128 + 8 * n + 2 * (lane & 3)) = make_double2(c0[n][0], c0[n][1]);
    *reinterpret_cast<double2 *>(tile + (16 * W + 8 + off) * 128 + 8 * n + 2 * (lane & 3)) = make_double2(c1[n][0], c1[n][1]);
  }
}

// grid: k-slice major over the column blocks I = 0 .. nblk - 1 of the ragged last block column: I = nblk - 1 is the diagonal
// region (strip form), I < nblk - 1 the off-diagonal regions (column-strip form; only when prm.skip_ragged_diag has bit 1)
template <int A>
__global__ void __launch_bounds__(kSyrkThreads, 1)
syrk_rdiag_kernel(const __grid_constant__ CUtensorMap xmap, SyrkParams prm) {
  extern __shared__ __align__(128) double smem[];
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kSyrkStages * kSyrkStageDoubles);
  uint64_t *empty_bar = full_bar + kSyrkStages;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool with_offdiag = (prm.skip_ragged_diag & 2) != 0;
  const int per = with_offdiag ? prm.nblk : 1;
  const int kslice = (int)blockIdx.x / per, J = prm.nblk - 1;
  const int I = with_offdiag ? (int)blockIdx.x - kslice * per : J;
  const bool diag = I == J;
  if (diag && !(prm.skip_ragged_diag & 1)) return;
  const int region = I * prm.nblk - I * (I - 1) / 2 + (J - I);
  const int64_t row_begin = (int64_t)kslice * prm.rows_per_slice;
  const int64_t row_end = min(prm.n, row_begin + prm.rows_per_slice);
  const int nstages_total = row_end > row_begin ? (int)((row_end - row_begin + kSyrkKB - 1) / kSyrkKB) : 0;
  if (tid == 0) {
    for (int s = 0; s < kSyrkStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, kSyrkConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  constexpr uint32_t kPanelBytes = kSyrkKB * kSyrkPanelLd * sizeof(double);
  const uint32_t kStageBytes = (diag ? 1u : 2u) * kPanelBytes + 2 * kSyrkKB * 8;
  if (wid >= kSyrkConsumerWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (wid == kSyrkConsumerWarps) {
      for (int it = 0; it < nstages_total; ++it) {
        const int s = it % kSyrkStages;
        const uint32_t phase = (it / kSyrkStages) & 1;
        mbar_wait(empty_bar + s, phase ^ 1);
        double *stage = smem + s * kSyrkStageDoubles;
        const int64_t r0 = row_begin + (int64_t)it * kSyrkKB;
        if (lane == 0) {
          mbar_expect_tx(full_bar + s, kStageBytes);
          tma_load_2d(stage, &xmap, 128 * I, (int)r0, full_bar + s);
          if (!diag) tma_load_2d(stage + kSyrkKB * kSyrkPanelLd, &xmap, 128 * J, (int)r0, full_bar + s);
          tma_bulk_g2s(stage + 2 * kSyrkKB * kSyrkPanelLd, prm.w + r0, kSyrkKB * 8, full_bar + s);
          tma_bulk_g2s(stage + 2 * kSyrkKB * kSyrkPanelLd + kSyrkKB, prm.s + r0, kSyrkKB * 8, full_bar + s);
        }
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
  SyrkWarpCtx wc;
  wc.smem = smem; wc.full_bar = full_bar; wc.empty_bar = empty_bar; wc.nstages_total = nstages_total;
  wc.lane = lane;
  wc.panelB_off = 0;
  wc.tile = prm.partials + ((int64_t)kslice * prm.nregions + region) * kSyrkTileLen;
  if (!diag) { syrk_consume_cstrip<A>(wc, wid); return; }
  switch (wid) {
    case 0: syrk_consume_rstrip<0, A>(wc); break;
    case 1: syrk_consume_rstrip<1, A>(wc); break;
    case 2: syrk_consume_rstrip<2, A>(wc); break;
    case 3: syrk_consume_rstrip<3, A>(wc); break;
    case 4: syrk_consume_rstrip<4, A>(wc); break;
    case 5: syrk_consume_rstrip<5, A>(wc); break;
    case 6: syrk_consume_rstrip<6, A>(wc); break;
    default: syrk_consume_rstrip<7, A>(wc); break;
  }
}

// =============================================================================================
// Active-set statistics: G = X' diag(w) X_A for a small column set A (the included variables), plus
// diag_j = sum_i w_i x_ij^2 and X's.  One sweep over the inclusion indicators reads of X'WX only the columns in the
// current model and the diagonal (BinomialLogitSpikeSlabSampler.cpp:88-117,180-222 select sub-blocks of suf().xtx()), so
// p (|A| + 2) numbers replace the p^2 of the full SYRK: n p (2 |A| + 4) flops instead of n p (p + 1) -- at C3 (p = 500,
// |A| ~ 21) the step becomes bound by reading X once.
//   CTA = (k-slice, 128-column block I of X); per stage one TMA tile of X (132 x 16) and one of X_A (B panel, NBA 8-column
//   atoms + 4 pad columns, from the gathered matrix boomgpu_select_columns builds), w and s by bulk copies, through the same
//   mbarrier ring as the SYRK.  Consumer warp q owns 16 columns of the block (2 A atoms) x all NBA B atoms.
//   Partial per CTA: [128 x 128 (NBA * 8 columns used) | diag 128 | xty 128], reduced over k-slices in order.
// =============================================================================================
constexpr int64_t kPanelTileLen = 128 * 128 + 256;
struct PanelParams {
  int64_t n;
  int p, nblk, ka8;        // ka8 = 8 NBA: columns of the B panel (zero padded)
  const double *w, *s;
  int ksplit;
  int64_t rows_per_slice;
  double *partials;        // [ksplit][nblk][kPanelTileLen]
};

template <int NBA, int BLD>
__device__ __forceinline__ void panel_consume_b(const SyrkWarpCtx &wc, int wq) {
  const int lane = wc.lane;
  double c[2][NBA][2], dg[2] = {0.0, 0.0}, cx[2] = {0.0, 0.0};
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NBA; ++n) c[m][n][0] = c[m][n][1] = 0.0;
  const int a_off = 16 * wq + (lane >> 2);
  for (int it = 0; it < wc.nstages_total; ++it) {
    const int s = it % kSyrkStages;
    const uint32_t phase = (it / kSyrkStages) & 1;
    mbar_wait(wc.full_bar + s, phase);
    const double *stage = wc.smem + s * kSyrkStageDoubles;
    const double *w_s = stage + 2 * kSyrkKB * kSyrkPanelLd;
    const double *s_s = w_s + kSyrkKB;
#pragma unroll
    for (int kk = 0; kk < kSyrkKB / 4; ++kk) {
      const int row = kk * 4 + (lane & 3);
      const double wv = w_s[row], sv = s_s[row];
      const double *xr = stage + row * kSyrkPanelLd;
      double a[2], aw[2], b[NBA];
#pragma unroll
      for (int m = 0; m < 2; ++m) { a[m] = xr[a_off + 8 * m]; aw[m] = a[m] * wv; }
#pragma unroll
      for (int n = 0; n < NBA; ++n) b[n] = stage[wc.panelB_off + row * BLD + (lane >> 2) + 8 * n];   // pitch 8 NBA + 4: 4 mod 8, conflict free
#pragma unroll
      for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int n = 0; n < NBA; ++n) dmma884(c[m][n][0], c[m][n][1], aw[m], b[n]);
        dg[m] = fma(aw[m], a[m], dg[m]);
        cx[m] = fma(a[m], sv, cx[m]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(wc.empty_bar + s);
  }
  double *tile = wc.tile;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
#pragma unroll
    for (int n = 0; n < NBA; ++n)
      *reinterpret_cast<double2 *>(tile + (16 * wq + 8 * m + (lane >> 2)) * 128 + 8 * n + 2 * (lane & 3)) = make_double2(c[m][n][0], c[m][n][1]);
    dg[m] += __shfl_xor_sync(0xffffffffu, dg[m], 1); dg[m] += __shfl_xor_sync(0xffffffffu, dg[m], 2);
    cx[m] += __shfl_xor_sync(0xffffffffu, cx[m], 1); cx[m] += __shfl_xor_sync(0xffffffffu, cx[m], 2);
    if ((lane & 3) == 0) {
      tile[128 * 128 + 16 * wq + 8 * m + (lane >> 2)] = dg[m];
      tile[128 * 128 + 128 + 16 * wq + 8 * m + (lane >> 2)] = cx[m];
    }
  }
}

template <int NBA>
__global__ void __launch_bounds__(kSyrkThreads, 1)
panel_dmma_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap amap, PanelParams prm) {
  extern __shared__ __align__(128) double smem[];
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kSyrkStages * kSyrkStageDoubles);
  uint64_t *empty_bar = full_bar + kSyrkStages;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int kslice = blockIdx.x / prm.nblk;
  const int I = blockIdx.x - kslice * prm.nblk;
  const int64_t row_begin = (int64_t)kslice * prm.rows_per_slice;
  const int64_t row_end = min(prm.n, row_begin + prm.rows_per_slice);
  const int nstages_total = row_end > row_begin ? (int)((row_end - row_begin + kSyrkKB - 1) / kSyrkKB) : 0;
  if (tid == 0) {
    for (int s = 0; s < kSyrkStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, kSyrkConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  // the B tile: box {8 NBA + 4, 16} of the gathered matrix: it lands with its own row pitch 8 NBA + 4 doubles (4 mod 8)
  constexpr int kBLd = 8 * NBA + 4;
  constexpr uint32_t kABytes = kSyrkKB * kSyrkPanelLd * sizeof(double), kBBytes = kSyrkKB * kBLd * sizeof(double);
  if (wid >= kSyrkConsumerWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (wid == kSyrkConsumerWarps) {
      for (int it = 0; it < nstages_total; ++it) {
        const int s = it % kSyrkStages;
        const uint32_t phase = (it / kSyrkStages) & 1;
        mbar_wait(empty_bar + s, phase ^ 1);
        double *stage = smem + s * kSyrkStageDoubles;
        const int64_t r0 = row_begin + (int64_t)it * kSyrkKB;
        if (lane == 0) {
          mbar_expect_tx(full_bar + s, kABytes + kBBytes + 2 * kSyrkKB * 8);
          tma_load_2d(stage, &xmap, 128 * I, (int)r0, full_bar + s);
          tma_load_2d(stage + kSyrkKB * kSyrkPanelLd, &amap, 0, (int)r0, full_bar + s);
          tma_bulk_g2s(stage + 2 * kSyrkKB * kSyrkPanelLd, prm.w + r0, kSyrkKB * 8, full_bar + s);
          tma_bulk_g2s(stage + 2 * kSyrkKB * kSyrkPanelLd + kSyrkKB, prm.s + r0, kSyrkKB * 8, full_bar + s);
        }
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
  SyrkWarpCtx wc;
  wc.smem = smem; wc.full_bar = full_bar; wc.empty_bar = empty_bar; wc.nstages_total = nstages_total;
  wc.lane = lane;
  wc.panelB_off = kSyrkKB * kSyrkPanelLd;
  wc.tile = prm.partials + ((int64_t)kslice * prm.nblk + I) * kPanelTileLen;
  panel_consume_b<NBA, kBLd>(wc, wid);
}

// G[j][a] (row major p x ka8), diag[j], xty[j] <- sums of the partial tiles over k-slices (fixed order)
__global__ void reduce_panel_kernel(PanelParams prm, double *__restrict__ G, double *__restrict__ diag, double *__restrict__ xty) {
  const int p = prm.p, ka8 = prm.ka8;
  const int64_t per_blk = 128 * (int64_t)(ka8 + 2);
  const int64_t total = (int64_t)prm.nblk * per_blk;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int I = (int)(e / per_blk);
    const int t = (int)(e - (int64_t)I * per_blk);
    const int r = t / (ka8 + 2), cidx = t - r * (ka8 + 2);
    const int j = 128 * I + r;
    if (j >= p) continue;
    const int64_t src = cidx < ka8 ? (int64_t)r * 128 + cidx : (int64_t)128 * 128 + (cidx - ka8) * 128 + r;
    double sum = 0;
    for (int k = 0; k < prm.ksplit; ++k) sum += prm.partials[((int64_t)k * prm.nblk + I) * kPanelTileLen + src];
    if (cidx < ka8) G[(int64_t)j * ka8 + cidx] = sum;
    else if (cidx == ka8) diag[j] = sum;
    else xty[j] = sum;
  }
}

// Sums partial tiles over k-slices (fixed order) into the p x p matrix (both triangles) and xty.
__global__ void reduce_syrk_kernel(SyrkParams prm, double *__restrict__ suf) {
  const int p = prm.p;
  const int64_t total = (int64_t)prm.nregions * kSyrkTileLen;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int region = (int)(e / kSyrkTileLen);
    const int t = (int)(e - (int64_t)region * kSyrkTileLen);
    int I, J;
    region_to_blocks(region, prm.nblk, I, J);
    int a, b = 0;
    bool is_xty = t >= 128 * 128;
    if (is_xty) {
      if (I != J) continue;
      a = 128 * I + (t - 128 * 128);
      if (a >= p) continue;
    } else {
      a = 128 * I + t / 128;
      b = 128 * J + t % 128;
      if (a >= p || b >= p || a > b) continue;
    }
    double sum = 0;
    for (int k = 0; k < prm.ksplit; ++k) sum += prm.partials[((int64_t)k * prm.nregions + region) * kSyrkTileLen + t];
    if (is_xty) {
      suf[(int64_t)p * p + a] = sum;
    } else {
      suf[a + (int64_t)b * p] = sum;
      suf[b + (int64_t)a * p] = sum;
    }
  }
}

// sums per-CTA scalar partials (4 per CTA) into suf[p*p+p .. +4)
__global__ void reduce_scalars_kernel(const double *__restrict__ partials, int nparts, double *__restrict__ dst) {
  if (threadIdx.x < 4) {
    double s = 0;
    for (int c = 0; c < nparts; ++c) s += partials[(int64_t)c * 4 + threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

// =============================================================================================
// Log likelihood (value only): warp per row, block partials, fixed-order final sum.
// =============================================================================================
template <int MODEL>
__global__ void __launch_bounds__(256) loglike_kernel(RowData d, const double *__restrict__ beta, double *__restrict__ partials) {
  extern __shared__ __align__(128) double smem[];
  double *beta_s = smem;
  __shared__ double red_s[8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int j = tid; j < d.p; j += 256) beta_s[j] = beta[j];
  __syncthreads();
  double acc = 0;
  const int warps_per_grid = gridDim.x * 8;
  for (int64_t i = (int64_t)blockIdx.x * 8 + wid; i < d.n; i += warps_per_grid) {
    const double *xr = d.X + i * d.ldx;
    double e = 0;
    for (int j = lane; j < d.p; j += 32) e = fma(__ldg(xr + j), beta_s[j], e);
    e = warp_sum(e);
    if (lane == 0) {
      if (MODEL == kLogit) acc += dbinom_log(d.y[i], d.ntrials[i], e);
      else acc += dpois_log((double)d.yi[i], d.exposure[i] * exp(e));
    }
  }
  if (lane == 0) red_s[wid] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += red_s[w];
    partials[blockIdx.x] = s;
  }
}

// Student-t sibling: residuals e_i = y_i - x_i'beta.  Two forms: a warp per row (wide rows), and for p <= 128 G lanes per row
// with several row groups in flight (at p = 16 the warp-per-row form leaves half the lanes idle and serialises one load latency
// per row: 3 ms per 25 M rows; a lane per row thrashes L1 with its 128-byte stride: 2.9 ms).
__global__ void __launch_bounds__(256) residual_kernel(RowData d, const double *__restrict__ beta, double *__restrict__ resid) {
  extern __shared__ __align__(128) double smem[];
  double *beta_s = smem;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int j = tid; j < d.p; j += 256) beta_s[j] = beta[j];
  __syncthreads();
  const int warps_per_grid = gridDim.x * 8;
  for (int64_t i = (int64_t)blockIdx.x * 8 + wid; i < d.n; i += warps_per_grid) {
    const double *xr = d.X + i * d.ldx;
    double e = 0;
    for (int j = lane; j < d.p; j += 32) e = fma(__ldg(xr + j), beta_s[j], e);
    e = warp_sum(e);
    if (lane == 0) resid[i] = d.y[i] - e;
  }
}

// G lanes per row (G = the power of two >= min(p, 32)): a warp instruction reads 32 / G consecutive rows -- one contiguous
// stretch of X when ldx == p -- and the group's partial products meet in log2(G) shuffles; U row groups in flight per warp.
template <int G>
__global__ void __launch_bounds__(256) residual_rows_kernel(RowData d, const double *__restrict__ beta, double *__restrict__ resid) {
  extern __shared__ __align__(128) double smem[];
  double *beta_s = smem;
  const int tid = threadIdx.x, lane = tid & 31, p = d.p;
  for (int j = tid; j < p; j += 256) beta_s[j] = beta[j];
  __syncthreads();
  constexpr int RPW = 32 / G, U = 8;      // rows per warp instruction, row groups in flight
  const int sub = lane / G, col0 = lane % G;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (tid >> 5), nwarps = (int64_t)gridDim.x * 8;
  for (int64_t base = warp * (RPW * U); base < d.n; base += nwarps * (RPW * U)) {
    double e[U];
    const double *xr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = min(base + u * RPW + sub, d.n - 1);   // rows beyond n re-read the last row; their result is not stored
      xr[u] = d.X + i * d.ldx;
      e[u] = 0.0;
    }
    for (int j = col0; j < p; j += G) {   // the U loads of one trip are independent: all in flight together
      const double b = beta_s[j];
      double x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = __ldg(xr[u] + j);
#pragma unroll
      for (int u = 0; u < U; ++u) e[u] = fma(x[u], b, e[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) e[u] += __shfl_xor_sync(0xffffffffu, e[u], o);
      const int64_t i = base + u * RPW + sub;
      if (col0 == 0 && i < d.n) resid[i] = __ldg(d.y + i) - e[u];
    }
  }
}

// ... and the part of the observed-data log likelihood that depends on the rows, from the stored residuals:
// sum_i -(nu + 1)/2 log(1 + e_i^2 / (nu sigma^2)).  The slice sampler on nu (TRegressionSampler::draw_nu_given_observed_data,
// TRegressionSampler.cpp:173-176 over TRegressionModel::log_likelihood, TRegression.cpp:74-86) evaluates it several times per
// draw with beta and sigma fixed: 8 n bytes per evaluation instead of a pass over X.  Four independent chains per thread; fixed
// assignment of rows to threads and fixed-order partials (deterministic).  log(1 + x) by the branch-free logarithm: its
// absolute error per term is that of log1p, and the terms are summed.
__global__ void __launch_bounds__(256) student_loglike_kernel(const double *__restrict__ resid, int64_t n, double inv_sigma, double nu,
                                                              double *__restrict__ partials) {
  __shared__ double red_s[8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double inv_nu = 1.0 / nu;
  const int64_t stride = (int64_t)gridDim.x * 256;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i0 = (int64_t)blockIdx.x * 256 + tid; i0 < n; i0 += 4 * stride) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) {
        const double delta = __ldg(resid + i) * inv_sigma;
        const double arg = fma(delta * delta, inv_nu, 1.0);
        acc[u] += arg < 1e300 ? log_nobranch(arg) : log(arg);   // overflow / NaN residuals take the library's special cases
      }
    }
  }
  double a = -0.5 * (nu + 1.0) * ((acc[0] + acc[1]) + (acc[2] + acc[3]));
  a = warp_sum(a);
  if (lane == 0) red_s[wid] = a;
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += red_s[w];
    partials[blockIdx.x] = s;
  }
}

// fixed-order sum of up to a few thousand partials by one warp: lane l sums its contiguous chunk, the lanes combine in a fixed tree
__global__ void reduce_sum_warp_kernel(const double *__restrict__ partials, int nparts, double *__restrict__ dst) {
  const int lane = threadIdx.x;
  const int chunk = (nparts + 31) / 32;
  double s = 0;
  for (int c = lane * chunk; c < min(nparts, (lane + 1) * chunk); ++c) s += partials[c];
  s = warp_sum(s);
  if (lane == 0) dst[0] = s;
}

// X's alone (the probit sibling: X'WX does not change from one iteration to the next, only X'z does): thread t owns column
// pairs t, t + 256, ... (coalesced 16-byte loads along a row), a CTA walks its block of rows; per-CTA partials, fixed-order sum.
constexpr int kXtsThreads = 256;
constexpr int kXtsMaxPairs = 32;   // column pairs per thread: p <= 2 * 256 * 32 = 16384
template <int PAIRS>
__global__ void __launch_bounds__(kXtsThreads) xts_kernel(const double *__restrict__ X, int64_t ldx, int64_t n, int p,
                                                           const double *__restrict__ s, double *__restrict__ partials) {
  const int tid = threadIdx.x;
  const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
  double2 acc[PAIRS];
#pragma unroll
  for (int q = 0; q < PAIRS; ++q) acc[q] = make_double2(0.0, 0.0);
  const int p2 = (p + 1) >> 1;   // ldx is even and the pad column (if p is odd) holds zeros or is never stored
  for (int64_t i = r0; i < r1; ++i) {
    const double sv = __ldg(s + i);
    const double2 *xr = reinterpret_cast<const double2 *>(X + i * ldx);
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) {
      const int c = tid + q * kXtsThreads;
      if (c < p2) {
        const double2 x = __ldg(xr + c);
        acc[q].x = fma(x.x, sv, acc[q].x);
        acc[q].y = fma(x.y, sv, acc[q].y);
      }
    }
  }
  double *my = partials + (int64_t)blockIdx.x * (2 * (int64_t)p2);
#pragma unroll
  for (int q = 0; q < PAIRS; ++q) {
    const int c = tid + q * kXtsThreads;
    if (c < p2) { my[2 * c] = acc[q].x; my[2 * c + 1] = acc[q].y; }
  }
}
// out_i = w_i x_ij: the right-hand side of column j of X'WX (boomgpu_weighted_column)
__global__ void weight_column_kernel(const double *__restrict__ X, int64_t ldx, int64_t n, int j, const double *__restrict__ w,
                                     double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __ldg(w + i) * __ldg(X + i * ldx + j);
}
__global__ void reduce_xts_kernel(const double *__restrict__ partials, int nparts, int p, double *__restrict__ xty) {
  const int p2x2 = 2 * ((p + 1) >> 1);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x) {
    double sum = 0;
    for (int c = 0; c < nparts; ++c) sum += partials[(int64_t)c * p2x2 + j];
    xty[j] = sum;
  }
}

// present[v] = 1 iff some row has count y == v, 0 <= v < len (benign same-value races).  What the host needs to know which
// entries NormalMixtureApproximationTable::approximate would have to add for this data (NormalMixtureApproximation.cpp:472-532).
__global__ void counts_present_kernel(const int64_t *__restrict__ y, int64_t n, unsigned char *__restrict__ present, int64_t len) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = __ldg(y + i);
    if (v >= 0 && v < len) present[v] = 1;
  }
}

// out (n x ldo, pad columns zero) = the k listed columns of X: the included-variable design matrix X_gamma that the
// chunk log posteriors of the composite sampler walk (BinomialLogitCompositeSpikeSlabSampler.cpp:34-74 selects the same
// columns from every observation on every evaluation)
__global__ void select_columns_kernel(const double *__restrict__ X, int64_t ldx, int64_t n, const int *__restrict__ cols, int k,
                                      double *__restrict__ out, int ldo) {
  const int64_t total = n * ldo;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / ldo;
    const int j = (int)(e - i * ldo);
    out[e] = j < k ? __ldg(X + i * ldx + __ldg(cols + j)) : 0.0;
  }
}

__global__ void reduce_sum_kernel(const double *__restrict__ partials, int nparts, double *__restrict__ dst) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0;
    for (int c = 0; c < nparts; ++c) s += partials[c];
    dst[0] = s;
  }
}

}  // namespace boomgpu
