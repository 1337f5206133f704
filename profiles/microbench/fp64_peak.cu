// FP64 pipe microbenchmark for B200 (sm_100a): measures the roofline
// denominators that MEASURED_PEAKS.json does not carry (it only has HBM copy
// and bf16): DFMA issue rate, DMMA (mma.sync f64) issue rate in each PTX shape,
// whether the two overlap, and a read-only HBM stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CH>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double c[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_dmma884(double *out, int iters, double a, double b) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_dmma1688(double *out, int iters, double a, double b) {
  double c[CH][4];
  double av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
  for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma1688(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_dmma16816(double *out, int iters, double a, double b) {
  double c[CH][4];
  double av[8], bv[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) bv[i] = b + i;
#pragma unroll
  for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma16816(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// even warps DMMA, odd warps DFMA: do the two pipes overlap?
template <int CH>
__global__ void k_mixed(double *out, int iters, double a, double b, int dfma_per_dmma_iter) {
  int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double c[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters * dfma_per_dmma_iter; ++it) {
#pragma unroll
      for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i];
  } else {
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < CH; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 exp/log throughput (the latent draws are transcendental-heavy)
__global__ void k_exp(double *out, int iters, double a) {
  double x = a + threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; ++it) { s += exp(x); x += 1e-6; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_log(double *out, int iters, double a) {
  double x = a + 1.0 + threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; ++it) { s += log(x); x += 1e-6; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_expf(float *out, int iters, float a) {
  float x = a + threadIdx.x * 1e-3f, s = 0;
  for (int it = 0; it < iters; ++it) { s += __expf(x); x += 1e-6f; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// read-only stream: sum of a big buffer (what the imputer pass does to X)
__global__ void k_read(const double2 *__restrict__ x, size_t n2, double *out) {
  double s = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 v0 = x[i], v1 = x[i + stride], v2 = x[i + 2 * stride], v3 = x[i + 3 * stride];
    s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
  }
  for (; i < n2; i += stride) { double2 v = x[i]; s += v.x + v.y; }
  if (s == 1.2345) out[0] = s;
}

template <class F>
float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 1024));
  const int iters = 20000;
  for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM
    int threads = 128, blocks = sms * (wps / 4);
    {
      float ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 8 * iters * (double)blocks * threads;
      printf("{\"test\": \"dfma\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", wps, ms, fl / ms * 1e-9);
    }
    {
      float ms = timeit([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32);
      printf("{\"test\": \"dmma_m8n8k4\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", wps, ms, fl / ms * 1e-9);
    }
    {
      float ms = timeit([&] { k_dmma1688<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 1024 * 8 * iters * (double)blocks * (threads / 32);
      printf("{\"test\": \"dmma_m16n8k8\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", wps, ms, fl / ms * 1e-9);
    }
    {
      float ms = timeit([&] { k_dmma16816<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 2048 * 8 * iters * (double)blocks * (threads / 32);
      printf("{\"test\": \"dmma_m16n8k16\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", wps, ms, fl / ms * 1e-9);
    }
  }
  {
    // mixed at 16 warps/SM: 8 DMMA warps + 8 DFMA warps; dfma warps do r x iterations
    int threads = 128, blocks = sms * 4;
    for (int r = 0; r <= 8; r += 2) {
      float ms = timeit([&] { k_mixed<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, r); });
      double fl_mma = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 64);
      double fl_fma = 2.0 * 8 * iters * r * (double)blocks * (threads / 2);
      printf("{\"test\": \"mixed\", \"dfma_ratio\": %d, \"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}\n",
             r, ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    }
  }
  {
    int threads = 256, blocks = sms * 8, it2 = 2000;
    float ms = timeit([&] { k_exp<<<blocks, threads>>>(out, it2, 0.5); });
    printf("{\"test\": \"exp_f64\", \"ms\": %.3f, \"gexp_per_s\": %.1f}\n", ms, (double)it2 * blocks * threads / ms * 1e-6);
    ms = timeit([&] { k_log<<<blocks, threads>>>(out, it2, 0.5); });
    printf("{\"test\": \"log_f64\", \"ms\": %.3f, \"glog_per_s\": %.1f}\n", ms, (double)it2 * blocks * threads / ms * 1e-6);
    ms = timeit([&] { k_expf<<<blocks, threads>>>((float *)out, it2, 0.5f); });
    printf("{\"test\": \"expf_fast_f32\", \"ms\": %.3f, \"gexp_per_s\": %.1f}\n", ms, (double)it2 * blocks * threads / ms * 1e-6);
  }
  {
    size_t bytes = (size_t)8 << 30;  // 8 GiB >> L2
    double2 *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    for (int bps = 4; bps <= 16; bps *= 2) {
      float ms = timeit([&] { k_read<<<sms * bps, 256>>>(buf, bytes / 16, out); });
      printf("{\"test\": \"hbm_read_only\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"gbs\": %.1f}\n", bps, ms, bytes / ms * 1e-6);
    }
    CK(cudaFree(buf));
  }
  return 0;
}
