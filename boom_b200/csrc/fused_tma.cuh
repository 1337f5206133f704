// fused_tma.cuh -- the single-pass kernel for p <= 64 (configs C1, C2, C5): TMA-fed, warp-autonomous.
//
//   One persistent CTA per SM, NW warps, no block-wide barrier in the steady state.  Every warp owns a ring of
//   S slices in shared memory; a slice is 32 RPL consecutive observations x (8 NB + 2) doubles, written by ONE
//   cp.async.bulk.tensor.2d (TMA tile of the row-major X; the box is wider than p, so the pad columns arrive as
//   zeros, and rows beyond n arrive as zeros too) that completes on the slice's mbarrier.  The warp that consumes a
//   slice also re-arms it (lane 0 issues the copy S slices ahead right after the warp's last read), so there is no
//   producer warp and no "empty" barrier.
//   Per slice, lane r owns observations r, r + 32, ... (RPL of them: independent dependency chains the scheduler
//   interleaves):  eta = x_r . beta with 16-byte loads from shared memory (conflict free by the row pitch alone, see
//   tma_padw; beta comes from the constant bank with compile-time indices), the latent draw (Philox keyed by the
//   global row; y_r / n_r were fetched one slice ahead), then the warp accumulates the slice's rank-1 updates with
//   FP64 DMMA (m8n8k4; A = w_k x_k fragments, B = x_k fragments straight from the slice; (w_k, s_k) sit in the two pad
//   columns of row k) into the upper triangle of X'WX held in registers (NB (NB + 1) / 2 atoms), and X'Wz with NB
//   DFMA per 4 rows.  beta arrives as a kernel parameter.
//   Epilogue: the warps add their fragments into one shared tile (warp-tile parallel, fixed order per element), the CTA
//   writes one partial, and the LAST CTA to finish (a counter in global memory, __threadfence on both sides) sums the
//   partials in CTA order -- one thread per output element, 32 loads in flight -- and writes the statistics to the
//   device buffer and to the host-mapped copy: one launch per Gibbs step, deterministic (no floating point atomics).
//   Option single_launch = 0 keeps the separate reduce_partials_kernel.
//   Wide tiles (NB >= 5) also exist in a 12-warp form whose accumulators live in tensor memory between DMMA phases (PARK, below).
//
//   Algorithmic traffic: 8 (p + 2) bytes per observation, read once.
//   Reference equivalent: Imputer.hpp:175-180 over BinomialLogitAuxmixSampler.cpp:61-97 /
//   PoissonRegressionAuxMixSampler.cpp:58-81.
#pragma once
#include <cuda.h>

#include "kernels.cuh"

namespace boomgpu {

// Tile shape per NB.  A slice is 32 RPL consecutive observations x (8 NB + 4) doubles; lane r owns the RPL observations
// r, r + 32, ... of the slice.  Measured at p = 16 / 25 M rows: (12 warps, RPL 1) 0.893 ms = (8 warps, RPL 2) 0.894 ms:
// the shared FP64 / DMMA pipe (48 % busy) and the issue slots (45 %) bound it, not latency, so RPL stays 1 (finer slices
// balance short data sets better).
// (tuning knobs for the narrow tiles: -DBOOMGPU_TMA_NW_SMALL=.. -DBOOMGPU_TMA_S_SMALL=.. -DBOOMGPU_TMA_RPL_SMALL=..)
// Round 2, after the draw diet (fewer instructions, same dependent chains): the kernel is bound by the LATENCY of a lane's
// serial chain (Philox rounds -> exp -> reciprocal -> log -> selection), not by issue slots or the FP64 pipe, so two
// observations per lane (two independent chains the scheduler interleaves) now pay: p = 16 / 25 M rows 0.854 -> 0.749 ms
// with (12 warps, RPL 2); (16 warps, RPL 1) spills at 128 registers and is slower (0.902).
#ifndef BOOMGPU_TMA_NW_SMALL
#define BOOMGPU_TMA_NW_SMALL 12
#endif
#ifndef BOOMGPU_TMA_S_SMALL
#define BOOMGPU_TMA_S_SMALL 2
#endif
#ifndef BOOMGPU_TMA_RPL_SMALL
#define BOOMGPU_TMA_RPL_SMALL 2
#endif
#ifndef BOOMGPU_TMA_NW_3
#define BOOMGPU_TMA_NW_3 12
#endif
#ifndef BOOMGPU_TMA_S_3
#define BOOMGPU_TMA_S_3 2
#endif
#ifndef BOOMGPU_TMA_RPL_3
#define BOOMGPU_TMA_RPL_3 1
#endif
#ifndef BOOMGPU_TMA_NW_4
#define BOOMGPU_TMA_NW_4 10
#endif
#ifndef BOOMGPU_TMA_S_4
#define BOOMGPU_TMA_S_4 1
#endif
#ifndef BOOMGPU_TMA_RPL_4
#define BOOMGPU_TMA_RPL_4 2     // p = 32 / 8 M rows: 0.747 -> 0.688 ms (NB = 3 prefers RPL 1: 0.594 vs 0.611)
#endif
__host__ __device__ constexpr int tma_rpl(int nb) { return nb <= 2 ? BOOMGPU_TMA_RPL_SMALL : (nb == 3 ? BOOMGPU_TMA_RPL_3 : (nb == 4 ? BOOMGPU_TMA_RPL_4 : 1)); }
// wide tiles (40 < p <= 64): 8 warps with a SINGLE slice each beat 6 warps with two (p = 64: 1.59 vs 2.15 ms per 8 M rows):
// the per-slice DMMA work is long enough that the other warp of the sub-partition covers the reload bubble.
#ifndef BOOMGPU_TMA_NW_WIDE
#define BOOMGPU_TMA_NW_WIDE 8
#endif
#ifndef BOOMGPU_TMA_S_WIDE
#define BOOMGPU_TMA_S_WIDE 1
#endif
__host__ __device__ constexpr int tma_warps(int nb) { return nb <= 2 ? BOOMGPU_TMA_NW_SMALL : (nb == 3 ? BOOMGPU_TMA_NW_3 : (nb == 4 ? BOOMGPU_TMA_NW_4 : BOOMGPU_TMA_NW_WIDE)); }
__host__ __device__ constexpr int tma_stages(int nb) { return nb <= 2 ? BOOMGPU_TMA_S_SMALL : (nb == 3 ? BOOMGPU_TMA_S_3 : (nb == 4 ? BOOMGPU_TMA_S_4 : BOOMGPU_TMA_S_WIDE)); }
// Row pitch of a slice in shared memory: 8 NB + 2 doubles = 16 NB + 4 words.  (a) lane r reads row r with 16-byte loads:
// a quarter warp's rows start 4 r (NB even) or 20 r (NB odd) words apart mod 32 -- eight distinct 4-bank groups, conflict
// free with NO per-lane rotation of the column order, so beta comes straight from the constant bank with compile-time
// indices.  (b) the four rows of a DMMA k-step are taken two apart (rows 8 t + u + 2 k, see the k loop): 2 pitches =
// 32 NB + 8 words = 8 mod 32, so the 16 lanes of a fragment load (4 rows x 4 columns x 8 bytes) cover 32 distinct banks.
// The two pad columns of row r hold (w_r, s_r).
__host__ __device__ constexpr int tma_padw(int nb) { return 8 * nb + 2; }
__host__ __device__ constexpr int tma_slice_rows(int nb) { return 32 * tma_rpl(nb); }
__host__ __device__ constexpr int tma_slice_doubles(int nb) { return tma_slice_rows(nb) * tma_padw(nb); }
__host__ __device__ constexpr size_t tma_smem_bytes(int nb, int nw, bool park = false) {
  return sizeof(double) * ((size_t)nw * tma_stages(nb) * tma_slice_doubles(nb) + 8 * nb + 64 + (park ? 4 * 32 * nw : 0)) +
         sizeof(uint64_t) * nw * tma_stages(nb) + 128;
}
__host__ __device__ constexpr size_t tma_smem_bytes(int nb) { return tma_smem_bytes(nb, tma_warps(nb)); }
__host__ __device__ constexpr int64_t tma_partial_len(int nb) { return 64 * nb * nb + 8 * nb + 8; }

// beta travels as a kernel parameter (p <= 64: 512 bytes of the constant bank): no host->device copy per step
struct BetaParam { double b[64]; };

// Sum of the per-CTA partials in CTA order (deterministic), by warp `warp` of `nwarps`: one warp per output element, lanes
// stride over the CTAs, then a fixed shuffle tree.  Partials are read around L1 (__ldcg): in the single-launch form they
// were written by other CTAs of the same grid.
__device__ __forceinline__ void reduce_partials_body(const double *__restrict__ partials, int nparts, int nb, int p,
                                                     double *__restrict__ suf, double *__restrict__ host_out, const int *err, int warp,
                                                     int nwarps, int lane, bool writes_flag) {
  const int P8 = 8 * nb;
  const int64_t plen = 64 * (int64_t)nb * nb + 8 * nb + 8;
  const int ntri = p * (p + 1) / 2;
  const int total = ntri + p + 4;
  for (int e = warp; e < total; e += nwarps) {
    int a = 0, b = 0;
    int64_t src;
    if (e < ntri) {
      // e -> (a, b), a <= b, rows of the upper triangle laid end to end
      int rem = e;
      while (rem >= p - a) { rem -= p - a; ++a; }
      b = a + rem;
      src = (int64_t)a * P8 + b;
    } else if (e < ntri + p) {
      src = (int64_t)P8 * P8 + (e - ntri);
    } else {
      src = (int64_t)P8 * P8 + P8 + (e - ntri - p);
    }
    double s = 0;
    for (int cta = lane; cta < nparts; cta += 32) s += __ldcg(partials + cta * plen + src);
    s = warp_sum(s);
    if (lane == 0) {
      if (e < ntri) {
        suf[a + (int64_t)b * p] = s;
        suf[b + (int64_t)a * p] = s;
        if (host_out) { host_out[a + (int64_t)b * p] = s; host_out[b + (int64_t)a * p] = s; }
      } else {
        suf[(int64_t)p * p + (e - ntri)] = s;
        if (host_out) host_out[(int64_t)p * p + (e - ntri)] = s;
      }
    }
  }
  if (host_out && writes_flag) host_out[(int64_t)p * p + p + 4] = (double)__ldcg(err);
}

// what the last CTA of the single-launch form needs to finish the step
struct TailParams {
  double *suf;             // device: [p*p | p | 4]
  double *host_out;        // host-mapped copy (or null)
  unsigned int *counter;   // CTAs that have written their partial; reset by the last one
};

// The last CTA of a single-launch step: one THREAD per output element, the partials summed in CTA order (deterministic)
// with 32 independent loads in flight (a warp per element serialises ~20 elements x 5 dependent L2 round trips on every
// warp: 28 us at C1; this form is ~3 us).  Partials are read around L1: other CTAs of the same grid wrote them.
template <int NB>
__device__ __forceinline__ void sum_partials_tail(const double *__restrict__ partials, int nparts, int p, const TailParams &tail,
                                                  const int *err, int tid, int nthreads) {
  constexpr int P8 = 8 * NB;
  {
    const int ntri = p * (p + 1) / 2, total = ntri + p + 4;
    constexpr int64_t plen = tma_partial_len(NB);
    for (int e = tid; e < total; e += nthreads) {
      int a = 0, b = 0;
      int64_t src;
      if (e < ntri) {
        int rem = e;
        while (rem >= p - a) { rem -= p - a; ++a; }
        b = a + rem;
        src = (int64_t)a * P8 + b;
      } else if (e < ntri + p) {
        src = (int64_t)P8 * P8 + (e - ntri);
      } else {
        src = (int64_t)P8 * P8 + P8 + (e - ntri - p);
      }
      double s = 0;
      int cta = 0;
      for (; cta + 32 <= nparts; cta += 32) {
        double v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = __ldcg(partials + (int64_t)(cta + u) * plen + src);
#pragma unroll
        for (int u = 0; u < 32; ++u) s += v[u];
      }
      {   // the remainder (< 32 partials), again with every load in flight before the first add; same order of additions
        double v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = cta + u < nparts ? __ldcg(partials + (int64_t)(cta + u) * plen + src) : 0.0;
#pragma unroll
        for (int u = 0; u < 32; ++u) s += v[u];
      }
      if (e < ntri) {
        tail.suf[a + (int64_t)b * p] = s;
        tail.suf[b + (int64_t)a * p] = s;
        if (tail.host_out) { tail.host_out[a + (int64_t)b * p] = s; tail.host_out[b + (int64_t)a * p] = s; }
      } else {
        tail.suf[(int64_t)p * p + (e - ntri)] = s;
        if (tail.host_out) tail.host_out[(int64_t)p * p + (e - ntri)] = s;
      }
    }
    if (tail.host_out && tid == 0) tail.host_out[(int64_t)p * p + p + 4] = (double)__ldcg(err);
  }
}

// ---- accumulators parked in tensor memory (wide tiles, option small_variant = 4 / 5) -------------------------------------
// The register-resident triangle of a wide tile (NB = 7: 28 atoms = 112 registers) is dead weight during a warp's draw phase
// and caps the CTA at 8 warps x 255 registers, while the draw phase is a long dependent chain per lane that only MORE warps
// hide (ncu, C2: tensor pipe 52 % active, issue slots 25 %).  Blackwell's tensor memory (256 KB per SM, idle here: FP64 has
// no tcgen05.mma) takes the triangle between the DMMA phases: tcgen05.st after a slice's k-steps, tcgen05.ld before the next
// slice's -- 32x32b shape, lane = thread, one 32-bit column per register, each warp in its own lane quadrant
// (32 (warp mod 4)) and column slot, so no two warps ever touch the same cell.  The kernel then fits 10 or 12 warps.
constexpr int kParkColsPerWarp = 160;   // >= 4 * 36 (NB = 8), a multiple of 16; three warps per lane quadrant: 480 of 512 columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&u)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
               "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]), "r"(u[10]),
               "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&u)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                 "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr)
               : "memory");
}
template <int NA>
__device__ __forceinline__ void tmem_park(uint32_t taddr, const double (&c)[NA][2]) {
  constexpr int NR = 4 * NA, NCH = (NR + 15) / 16;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    uint32_t u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = 16 * ch + i;   // register r: atom r / 4, component (r % 4) / 2, low / high word r % 2
      u[i] = r < NR ? (uint32_t)((r & 1) ? __double2hiint(c[r / 4][(r % 4) / 2]) : __double2loint(c[r / 4][(r % 4) / 2])) : 0u;
    }
    tmem_st16(taddr + 16 * ch, u);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
template <int NA>
__device__ __forceinline__ void tmem_unpark(uint32_t taddr, double (&c)[NA][2]) {
  constexpr int NR = 4 * NA, NCH = (NR + 15) / 16;
  uint32_t u[NCH][16];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) tmem_ld16(taddr + 16 * ch, u[ch]);
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = 4 * a + 2 * j;
      c[a][j] = __hiloint2double((int)u[(r + 1) / 16][(r + 1) % 16], (int)u[r / 16][r % 16]);
    }
}

template <int NB, int MODEL, int NWT = tma_warps(NB), bool PARK = false>
__global__ void __launch_bounds__(32 * NWT, 1)
fused_tma_kernel(const __grid_constant__ CUtensorMap xmap, RowData d, DrawParams prm, RowOut out, const __grid_constant__ BetaParam beta,
                 double *__restrict__ partials, int *err, TailParams tail) {
  constexpr int NW = NWT;
  constexpr int S = tma_stages(NB);
  constexpr int RPL = tma_rpl(NB);
  constexpr int ROWS = 32 * RPL;
  constexpr int PADW = tma_padw(NB);
  constexpr int SLICE = tma_slice_doubles(NB);
  constexpr int P8 = 8 * NB;
  constexpr int NA = NB * (NB + 1) / 2;
  constexpr uint32_t kSliceBytes = SLICE * sizeof(double);
  static_assert(!PARK || (NW <= 12 && 4 * NA <= kParkColsPerWarp), "tensor-memory parking: at most three warps per lane quadrant");
  extern __shared__ __align__(128) double smem[];
  double *ring = smem;                                  // [NW][S][SLICE]
  double *beta_s = smem + (size_t)NW * S * SLICE;       // P8
  double *red_s = beta_s + P8;                          // 64
  double *sc_s = red_s + 64;                            // parked form only: [4][32 NW] per-thread scalar statistics
  uint64_t *bars = reinterpret_cast<uint64_t *>(sc_s + (PARK ? 4 * 32 * NW : 0));  // [NW][S]
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int p = d.p;
  if (tid == 0) {
    for (int i = 0; i < NW * S; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (PARK && wid == 0) {   // the whole tensor memory of the SM (one CTA per SM): 512 columns x 128 lanes x 32 bits
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  __syncthreads();
  // this warp's cells: lanes 32 (wid mod 4) .. + 31 (the only ones it may address), columns [slot * 160, + 4 NA)
  const uint32_t my_tmem = PARK ? tmem_base_s + ((uint32_t)(32 * (wid & 3)) << 16) + (uint32_t)((wid >> 2) * kParkColsPerWarp) : 0u;

  // slices of ROWS rows are dealt to (CTA, warp): the k-th slice of this warp is ((blockIdx + k grid) NW + wid)
  const int64_t nslices = (d.n + ROWS - 1) / ROWS;
  const int64_t stride = (int64_t)gridDim.x * NW;
  const int64_t first = (int64_t)blockIdx.x * NW + wid;
  double *my_ring = ring + (size_t)wid * S * SLICE;
  uint64_t *my_bars = bars + wid * S;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int64_t q = first + s * stride;
      if (q < nslices) {
        mbar_expect_tx(my_bars + s, kSliceBytes);
        tma_load_2d(my_ring + s * SLICE, &xmap, 0, (int)(q * ROWS), my_bars + s);
      }
    }
  }

  double c[NA][2];
#pragma unroll
  for (int a = 0; a < NA; ++a) { c[a][0] = 0.0; c[a][1] = 0.0; }
  double xty_acc[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) xty_acc[b] = 0.0;
  double sc_count = 0, sc_ywy = 0, sc_sumw = 0, sc_sumlogw = 0;
  if (PARK) {
    tmem_park<NA>(my_tmem, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) sc_s[k * 32 * NW + tid] = 0.0;
  }

  int slot = 0;
  uint32_t phase = 0;
  // y_i / n_i (or exposure, or supplied latents) are fetched one slice ahead: a global load per lane whose latency
  // the previous slice's arithmetic covers
  RowObs obs_next[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    obs_next[j].y = 0; obs_next[j].aux = 0; obs_next[j].yi = 0;
    const int64_t i0 = first * ROWS + 32 * j + lane;
    if (first < nslices && i0 < d.n) obs_next[j] = load_obs<MODEL>(d, i0);
  }
  for (int64_t q = first; q < nslices; q += stride) {
    int64_t i[RPL];
    bool valid[RPL];
    RowObs obs[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      i[j] = q * ROWS + 32 * j + lane;
      valid[j] = i[j] < d.n;
      obs[j] = obs_next[j];
      const int64_t in = (q + stride) * ROWS + 32 * j + lane;
      if (in < d.n) obs_next[j] = load_obs<MODEL>(d, in);
    }
    mbar_wait(my_bars + slot, phase);
    double *xs = my_ring + slot * SLICE;

    // ---- lane r: eta of its RPL observations, then their latent draws
    double eta[RPL];
    {
      double e0[RPL], e1[RPL];
#pragma unroll
      for (int j = 0; j < RPL; ++j) { e0[j] = 0; e1[j] = 0; }
#pragma unroll
      for (int col = 0; col < P8; col += 2) {
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          const double2 x = *reinterpret_cast<const double2 *>(xs + (32 * j + lane) * PADW + col);
          e0[j] = fma(x.x, beta.b[col], e0[j]);        // beta: kernel parameter, constant-bank operand
          e1[j] = fma(x.y, beta.b[col + 1], e1[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < RPL; ++j) eta[j] = e0[j] + e1[j];
    }
    double wv[RPL], sv[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) { wv[j] = 0; sv[j] = 0; }
    bool done = false;
    if (MODEL == kLogit && RPL > 1) {
      // all of this warp's observations are Bernoulli with a moderate eta: RPL straight-line draws per lane, interleaved
      bool fast = true;
#pragma unroll
      for (int j = 0; j < RPL; ++j) fast = fast && valid[j] && logit_is_bernoulli_fast(obs[j].aux, obs[j].y, eta[j]);
      if (__all_sync(0xffffffffu, fast)) {
        uint64_t rows[RPL];
        double ys[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) { rows[j] = d.row_offset + (uint64_t)i[j]; ys[j] = obs[j].y; }
        logit_bernoulli_draws<RPL>(prm.hot, prm.mix, eta, ys, prm.key, rows, sv, wv);
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          sc_count += 1.0; sc_sumw += wv[j];
          if (out.w) out.w[i[j]] = wv[j];
          if (out.s) out.s[i[j]] = sv[j];
        }
        done = true;
      }
    }
    if (MODEL == kStudentT && RPL > 1 && prm.t_nu >= 1.0) {
      // the Student-t weights of this lane's RPL observations: first attempts interleaved (draws.cuh: rgamma_philox_rows)
      bool all_valid = true;
#pragma unroll
      for (int j = 0; j < RPL; ++j) all_valid = all_valid && valid[j];
      if (__all_sync(0xffffffffu, all_valid)) {
        uint64_t rows[RPL];
        double rate[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          rows[j] = d.row_offset + (uint64_t)i[j];
          const double delta = (obs[j].y - eta[j]) * prm.t_inv_sigma;
          rate[j] = 0.5 * (prm.t_nu + delta * delta);
        }
        if (!rgamma_philox_rows<RPL>(0.5 * (prm.t_nu + 1.0), rate, prm.key, rows, wv)) atomicOr(err, 2);
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          sv[j] = wv[j] * obs[j].y;
          sc_count += 1.0; sc_ywy += sv[j] * obs[j].y; sc_sumw += wv[j]; sc_sumlogw += wv[j] > 1e-300 ? log_nobranch(wv[j]) : 0.0;
          if (out.w) out.w[i[j]] = wv[j];
          if (out.s) out.s[i[j]] = sv[j];
        }
        done = true;
      }
    }
    if (!done) {
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        if (valid[j]) {
          RowLatent r = impute_row<MODEL>(d, prm, out, obs[j], i[j], eta[j], err);
          wv[j] = r.w; sv[j] = r.s;
          sc_count += r.count; sc_ywy += r.yWy; sc_sumw += r.w; sc_sumlogw += r.sumlogw;
        }
      }
    }
    // (w_r, s_r) go to the two pad columns of row r: the k-steps below read them back with one 16-byte load
#pragma unroll
    for (int j = 0; j < RPL; ++j)
      *reinterpret_cast<double2 *>(xs + (32 * j + lane) * PADW + P8) = make_double2(wv[j], sv[j]);
    __syncwarp();

    // ---- the warp's rank-1 updates: DMMA k-steps of 4 rows
    if (PARK) {
      // nothing of the draw phase stays in registers across the k-steps: the scalar statistics go to this thread's cells
      sc_s[0 * 32 * NW + tid] += sc_count; sc_s[1 * 32 * NW + tid] += sc_ywy;
      sc_s[2 * 32 * NW + tid] += sc_sumw;  sc_s[3 * 32 * NW + tid] += sc_sumlogw;
      sc_count = 0; sc_ywy = 0; sc_sumw = 0; sc_sumlogw = 0;
      tmem_unpark<NA>(my_tmem, c);
    }
#pragma unroll(NB <= 2 ? 8 : (NB <= 4 ? 4 : (PARK ? 1 : 2)))
    for (int kk = 0; kk < ROWS / 4; ++kk) {
      const int row = 8 * (kk >> 1) + (kk & 1) + 2 * (lane & 3);   // the k-step's four rows, two apart (bank layout: tma_padw)
      const double2 ws = *reinterpret_cast<const double2 *>(xs + row * PADW + P8);
      const double wk = ws.x, sk = ws.y;
      const double *xr = xs + row * PADW + (lane >> 2);
      if (PARK) {   // register diet: the A fragment of one atom row at a time
        double xa[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          xa[b] = xr[8 * b];
          xty_acc[b] = fma(xa[b], sk, xty_acc[b]);
        }
        int a = 0;
#pragma unroll
        for (int bi = 0; bi < NB; ++bi) {
          const double xwi = xa[bi] * wk;
#pragma unroll
          for (int bj = bi; bj < NB; ++bj) { dmma884(c[a][0], c[a][1], xwi, xa[bj]); ++a; }
        }
      } else {
        double xa[NB], xw[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          xa[b] = xr[8 * b];
          xw[b] = xa[b] * wk;
          xty_acc[b] = fma(xa[b], sk, xty_acc[b]);
        }
        int a = 0;
#pragma unroll
        for (int bi = 0; bi < NB; ++bi)
#pragma unroll
          for (int bj = bi; bj < NB; ++bj) { dmma884(c[a][0], c[a][1], xw[bi], xa[bj]); ++a; }
      }
    }

    if (PARK) tmem_park<NA>(my_tmem, c);
    // ---- re-arm the slot S slices ahead (all lanes are done with it; the generic-proxy writes of (w, s) are
    // ordered before the async-proxy overwrite)
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      const int64_t qn = q + (int64_t)S * stride;
      if (qn < nslices) {
        mbar_expect_tx(my_bars + slot, kSliceBytes);
        tma_load_2d(my_ring + slot * SLICE, &xmap, 0, (int)(qn * ROWS), my_bars + slot);
      }
    }
    if (++slot == S) { slot = 0; phase ^= 1; }
  }

  // ---- CTA reduction in warp order (deterministic), one partial per CTA
  if (PARK) tmem_unpark<NA>(my_tmem, c);
  __syncthreads();  // every issued copy has been consumed: the ring is free
  if (PARK && wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base_s) : "memory");
  constexpr int TILE = P8 * P8 + P8;    // [P8 x P8 | X'Wz P8]
  // X'Wz: lanes with the same (lane >> 2) hold the rows = lane & 3 (mod 4) of the same columns
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    xty_acc[b] += __shfl_xor_sync(0xffffffffu, xty_acc[b], 1);
    xty_acc[b] += __shfl_xor_sync(0xffffffffu, xty_acc[b], 2);
  }
  double *my = partials + (int64_t)blockIdx.x * tma_partial_len(NB);
  constexpr bool kWarpTiles = (size_t)NW * TILE <= (size_t)NW * S * SLICE;   // every warp's fragments fit the (free) ring
  if (kWarpTiles) {
    // each warp lays its fragments into its own tile, then every thread sums one element over the warps in warp order:
    // two block barriers instead of one per warp (the epilogue is most of the kernel when n is small: C1)
    for (int e = tid; e < NW * TILE; e += 32 * NW) smem[e] = 0.0;
    __syncthreads();
    double *tile = smem + (size_t)wid * TILE;
    int a = 0;
#pragma unroll
    for (int bi = 0; bi < NB; ++bi)
#pragma unroll
      for (int bj = bi; bj < NB; ++bj) {
        *reinterpret_cast<double2 *>(tile + (8 * bi + (lane >> 2)) * P8 + 8 * bj + 2 * (lane & 3)) = make_double2(c[a][0], c[a][1]);
        ++a;
      }
    if ((lane & 3) == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) tile[P8 * P8 + 8 * b + (lane >> 2)] = xty_acc[b];
    }
    __syncthreads();
    for (int e = tid; e < TILE; e += 32 * NW) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += smem[(size_t)w * TILE + e];
      my[e] = s;
    }
  } else {
    double *tile = smem;                  // P8 * P8
    double *xty_s = smem + P8 * P8;       // P8
    for (int e = tid; e < TILE; e += 32 * NW) smem[e] = 0.0;
    __syncthreads();
    for (int w = 0; w < NW; ++w) {
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
        if ((lane & 3) == 0) {
#pragma unroll
          for (int b = 0; b < NB; ++b) xty_s[8 * b + (lane >> 2)] += xty_acc[b];
        }
      }
      __syncthreads();
    }
    for (int e = tid; e < TILE; e += 32 * NW) my[e] = smem[e];
  }
  if (PARK) {
    sc_count += sc_s[0 * 32 * NW + tid]; sc_ywy += sc_s[1 * 32 * NW + tid];
    sc_sumw += sc_s[2 * 32 * NW + tid];  sc_sumlogw += sc_s[3 * 32 * NW + tid];
  }
  double v0 = warp_sum(sc_count), v1 = warp_sum(sc_ywy), v2 = warp_sum(sc_sumw), v3 = warp_sum(sc_sumlogw);
  if (lane == 0) { red_s[wid * 4 + 0] = v0; red_s[wid * 4 + 1] = v1; red_s[wid * 4 + 2] = v2; red_s[wid * 4 + 3] = v3; }
  __syncthreads();
  if (tid < 4) {
    double s = 0;
    for (int w = 0; w < NW; ++w) s += red_s[w * 4 + tid];
    my[P8 * P8 + P8 + tid] = s;
  }

  // ---- single launch: the CTA that writes the LAST partial sums them all (in CTA order: deterministic) into the packed
  // statistics -- no second kernel, no launch gap (a third of the C1 iteration)
  if (tail.counter == nullptr) return;
  __shared__ unsigned int ticket_s;
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket_s = atomicAdd(tail.counter, 1u);
  __syncthreads();
  if (ticket_s != gridDim.x - 1) return;
  __threadfence();
  sum_partials_tail<NB>(partials, (int)gridDim.x, d.p, tail, err, tid, 32 * NW);
  if (tid == 0) *tail.counter = 0u;
}

// =============================================================================================
// Wide tiles (40 < p <= 64), warp-specialised.  In fused_tma_kernel every warp alternates between its draw phase (a long
// dependent chain per lane: latency bound) and its DMMA phase (28-36 atoms per k-step: pipe bound), and with the 8 warps
// that fit (the triangle alone is 112-144 registers) the FP64 pipe idles through every draw phase: 50 % busy at C2.
// Here the roles are split, per SM sub-partition m (0..3), around a ring of S 32-row slices:
//   accumulate warps m and 4 + m   own HALF of the triangle each (atoms a = h mod 2: 14-18 atoms, 56-72 registers) and do
//                                  nothing but DMMA k-steps on the ring's slices.  Two of them, because ONE warp cannot
//                                  keep the pipe busy: a warp that has issued a DMMA waits out the issue interval (the NOP
//                                  after every DMMA in the SASS) and its DMUL / DFMA / LDS go in series with it -- measured
//                                  53 % of the pipe with one accumulate warp per sub-partition.
//   draw warps 8 + m and 12 + m    take turns preparing the ring's slices: eta, latent draw, (w, s) into the pad columns.
// Slot k mod S: full[k] (TMA bytes landed) -> draw warp -> drawn[k] -> both accumulate warps -> empty[k] (two arrivals) ->
// accumulate warp m re-arms full and issues the TMA load of slice k + S.  No block-wide barrier in the steady state.
// 16 warps x 128 registers.
// =============================================================================================
constexpr int kWsAccWarps = 8, kWsRings = 4, kWsWarps = 16;
__host__ __device__ constexpr int ws_stages(int nb) { return nb <= 6 ? 4 : 3; }
__host__ __device__ constexpr size_t ws_smem_bytes(int nb) {
  return sizeof(double) * ((size_t)kWsRings * ws_stages(nb) * 32 * tma_padw(nb) + 64) +
         sizeof(uint64_t) * 3 * kWsRings * ws_stages(nb) + 128 + sizeof(PoissonSmem);
}

// The accumulate warp's loop (see fused_ws_kernel): atoms a with a mod 2 == HALF of the triangle, X's for the column blocks b
// with b mod 2 == HALF; HALF 0 also re-arms the ring.
template <int NB, int HALF>
__device__ __forceinline__ void ws_accumulate(const CUtensorMap &xmap, double *my_ring, uint64_t *my_full, uint64_t *my_drawn,
                                              uint64_t *my_empty, int64_t first, int64_t stride, int64_t nslices, int lane,
                                              double (&c)[(NB * (NB + 1) / 2 + 1) / 2][2], double (&xty_acc)[NB]) {
  constexpr int S = ws_stages(NB);
  constexpr int PADW = tma_padw(NB);
  constexpr int SLICE = 32 * PADW;
  constexpr int P8 = 8 * NB;
  constexpr int NH = (NB * (NB + 1) / 2 + 1) / 2;
  constexpr uint32_t kSliceBytes = SLICE * sizeof(double);
#pragma unroll
  for (int a = 0; a < NH; ++a) { c[a][0] = 0.0; c[a][1] = 0.0; }
#pragma unroll
  for (int b = 0; b < NB; ++b) xty_acc[b] = 0.0;
  if (HALF == 0 && lane == 0) {
#pragma unroll
    for (int s0 = 0; s0 < S; ++s0) {
      const int64_t q = first + s0 * stride;
      if (q < nslices) {
        mbar_expect_tx(my_full + s0, kSliceBytes);
        tma_load_2d(my_ring + s0 * SLICE, &xmap, 0, (int)(q * 32), my_full + s0);
      }
    }
  }
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t q = first; q < nslices; q += stride) {
    mbar_wait(my_drawn + slot, phase);
    const double *xs = my_ring + slot * SLICE;
#pragma unroll 2
    for (int kk = 0; kk < 8; ++kk) {
      const int row = 8 * (kk >> 1) + (kk & 1) + 2 * (lane & 3);   // the k-step's four rows, two apart (bank layout: tma_padw)
      const double2 ws = *reinterpret_cast<const double2 *>(xs + row * PADW + P8);
      const double *xr = xs + row * PADW + (lane >> 2);
      double xa[NB], xw[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        xa[b] = xr[8 * b];
        xw[b] = xa[b] * ws.x;
        if ((b & 1) == HALF) xty_acc[b] = fma(xa[b], ws.y, xty_acc[b]);
      }
      int a = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj) {
          if ((a & 1) == HALF) dmma884(c[a >> 1][0], c[a >> 1][1], xw[bi], xa[bj]);
          ++a;
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(my_empty + slot);
    if (HALF == 0) {
      // both halves are done with the slot: re-arm it S slices ahead (the draw warp's generic-proxy writes of (w, s) are
      // ordered before the async-proxy overwrite by the fence)
      mbar_wait(my_empty + slot, phase);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      if (lane == 0) {
        const int64_t qn = q + (int64_t)S * stride;
        if (qn < nslices) {
          mbar_expect_tx(my_full + slot, kSliceBytes);
          tma_load_2d(my_ring + slot * SLICE, &xmap, 0, (int)(qn * 32), my_full + slot);
        }
      }
    }
    if (++slot == S) { slot = 0; phase ^= 1; }
  }
}

template <int NB, int MODEL>
__global__ void __launch_bounds__(32 * kWsWarps, 1)
fused_ws_kernel(const __grid_constant__ CUtensorMap xmap, RowData d, DrawParams prm, RowOut out, const __grid_constant__ BetaParam beta,
                double *__restrict__ partials, int *err, TailParams tail) {
  constexpr int S = ws_stages(NB);
  constexpr int PADW = tma_padw(NB);
  constexpr int SLICE = 32 * PADW;
  constexpr int P8 = 8 * NB;
  constexpr int NA = NB * (NB + 1) / 2;
  constexpr int NH = (NA + 1) / 2;                                 // atoms per accumulate warp
  constexpr uint32_t kSliceBytes = SLICE * sizeof(double);
  extern __shared__ __align__(128) double smem[];
  double *ring = smem;                                             // [4][S][SLICE]
  double *red_s = smem + (size_t)kWsRings * S * SLICE;             // 64
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(red_s + 64);   // [4][S]
  uint64_t *drawn_bar = full_bar + kWsRings * S;                   // [4][S]
  uint64_t *empty_bar = drawn_bar + kWsRings * S;                  // [4][S]
  PoissonSmem *tab_s = reinterpret_cast<PoissonSmem *>(empty_bar + kWsRings * S + 2);   // 16-byte aligned

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m = wid & 3;                 // sub-partition = ring this warp works on
  const bool is_acc = wid < kWsAccWarps;
  const int half = (wid >> 2) & 1;       // accumulate warps: which half of the triangle; draw warps: whose turn
  if (tid == 0) {
    for (int i = 0; i < kWsRings * S; ++i) { mbar_init(full_bar + i, 1); mbar_init(drawn_bar + i, 1); mbar_init(empty_bar + i, 2); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (MODEL == kPoisson) poisson_smem_fill(tab_s, prm.tab, tid, 32 * kWsWarps);   // small-count table entries -> shared memory
  __syncthreads();

  // 32-row slices are dealt to (CTA, ring): the k-th slice of ring m is (blockIdx + k grid) 4 + m
  const int64_t nslices = (d.n + 31) / 32;
  const int64_t stride = (int64_t)gridDim.x * kWsRings;
  const int64_t first = (int64_t)blockIdx.x * kWsRings + m;
  double *my_ring = ring + (size_t)m * S * SLICE;
  uint64_t *my_full = full_bar + m * S, *my_drawn = drawn_bar + m * S, *my_empty = empty_bar + m * S;

  double c[NH][2];
  double xty_acc[NB];
  double sc_count = 0, sc_ywy = 0, sc_sumw = 0, sc_sumlogw = 0;

  if (is_acc) {
    // ===== accumulate warp (HALF at compile time: a run-time test on the atom index would issue every DMMA in both warps)
    if (half == 0) ws_accumulate<NB, 0>(xmap, my_ring, my_full, my_drawn, my_empty, first, stride, nslices, lane, c, xty_acc);
    else ws_accumulate<NB, 1>(xmap, my_ring, my_full, my_drawn, my_empty, first, stride, nslices, lane, c, xty_acc);
  } else {
    // ===== draw warp: every second slice of ring m
    RowObs obs_next;
    obs_next.y = 0; obs_next.aux = 0; obs_next.yi = 0;
    {
      const int64_t q0 = first + half * stride;
      const int64_t i0 = q0 * 32 + lane;
      if (q0 < nslices && i0 < d.n) obs_next = load_obs<MODEL>(d, i0);
    }
    int64_t k = half;
    for (int64_t q = first + half * stride; q < nslices; q += 2 * stride, k += 2) {
      const int slot = (int)(k % S);
      const uint32_t phase = (uint32_t)((k / S) & 1);
      const int64_t i = q * 32 + lane;
      const bool valid = i < d.n;
      const RowObs obs = obs_next;
      const int64_t in = (q + 2 * stride) * 32 + lane;
      if (in < d.n) obs_next = load_obs<MODEL>(d, in);
      mbar_wait(my_full + slot, phase);
      double *xs = my_ring + slot * SLICE;
      double e0 = 0, e1 = 0;
#pragma unroll
      for (int col = 0; col < P8; col += 2) {
        const double2 x = *reinterpret_cast<const double2 *>(xs + lane * PADW + col);
        e0 = fma(x.x, beta.b[col], e0);
        e1 = fma(x.y, beta.b[col + 1], e1);
      }
      double wv = 0, sv = 0;
      if (valid) {
        RowLatent r = impute_row<MODEL>(d, prm, out, obs, i, e0 + e1, err, MODEL == kPoisson ? tab_s : nullptr);
        wv = r.w; sv = r.s;
        sc_count += r.count; sc_ywy += r.yWy; sc_sumw += r.w; sc_sumlogw += r.sumlogw;
      }
      *reinterpret_cast<double2 *>(xs + lane * PADW + P8) = make_double2(wv, sv);
      __syncwarp();
      if (lane == 0) mbar_arrive(my_drawn + slot);   // release: the (w, s) stores above are visible to the waiters
    }
  }

  // ---- CTA reduction: the eight accumulate warps in warp order (deterministic), one partial per CTA
  __syncthreads();  // every issued copy has been consumed: the ring is free
  constexpr int TILE = P8 * P8 + P8;
  double *tile = smem;
  double *xty_s = smem + P8 * P8;
  for (int e = tid; e < TILE; e += 32 * kWsWarps) smem[e] = 0.0;
  if (is_acc) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      xty_acc[b] += __shfl_xor_sync(0xffffffffu, xty_acc[b], 1);
      xty_acc[b] += __shfl_xor_sync(0xffffffffu, xty_acc[b], 2);
    }
  }
  __syncthreads();
  for (int w = 0; w < kWsAccWarps; ++w) {
    if (wid == w) {
      int a = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj) {
          if ((a & 1) == half) {
            double *t = tile + (8 * bi + (lane >> 2)) * P8 + 8 * bj + 2 * (lane & 3);
            t[0] += c[a >> 1][0];
            t[1] += c[a >> 1][1];
          }
          ++a;
        }
      if ((lane & 3) == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if ((b & 1) == half) xty_s[8 * b + (lane >> 2)] += xty_acc[b];
      }
    }
    __syncthreads();
  }
  double *my = partials + (int64_t)blockIdx.x * tma_partial_len(NB);
  for (int e = tid; e < TILE; e += 32 * kWsWarps) my[e] = smem[e];
  double v0 = warp_sum(sc_count), v1 = warp_sum(sc_ywy), v2 = warp_sum(sc_sumw), v3 = warp_sum(sc_sumlogw);
  if (lane == 0) { red_s[wid * 4 + 0] = v0; red_s[wid * 4 + 1] = v1; red_s[wid * 4 + 2] = v2; red_s[wid * 4 + 3] = v3; }
  __syncthreads();
  if (tid < 4) {
    double sum = 0;
    for (int w = kWsAccWarps; w < kWsWarps; ++w) sum += red_s[w * 4 + tid];
    my[P8 * P8 + P8 + tid] = sum;
  }
  if (tail.counter == nullptr) return;
  __shared__ unsigned int ticket_s;
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket_s = atomicAdd(tail.counter, 1u);
  __syncthreads();
  if (ticket_s != gridDim.x - 1) return;
  __threadfence();
  sum_partials_tail<NB>(partials, (int)gridDim.x, d.p, tail, err, tid, 32 * kWsWarps);
  if (tid == 0) *tail.counter = 0u;
}

// One warp per output element: lanes stride over the per-CTA partials, then a fixed shuffle tree.
// suf layout [p*p | p | 4]; writes both triangles.  Partial layout: [P8*P8 tile | P8 | 8].
// host_out (optional): a host-mapped pinned buffer of suf_len + 1 doubles that receives the statistics and, in its last
// slot, the device-side validation flag -- the synchronous steps then need no device->host copy call at all.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double *__restrict__ partials, int nparts, int nb, int p,
                                                             double *__restrict__ suf, double *__restrict__ host_out,
                                                             const int *__restrict__ err) {
  reduce_partials_body(partials, nparts, nb, p, suf, host_out, err, (blockIdx.x * blockDim.x + threadIdx.x) >> 5,
                       (gridDim.x * blockDim.x) >> 5, threadIdx.x & 31, blockIdx.x == 0 && threadIdx.x == 0);
}

}  // namespace boomgpu
