// Micro-benchmark (measurement tool, not product): FP64 FMA-pipe and FP64 tensor (DMMA,
// mma.sync.m8n8k4 / m16n8k4 .f64) throughput on this GPU.  Gives the FP64 roofline denominator
// (MEASURED_PEAKS.json has only HBM and bf16).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
  double a[16];
  const double x = 1.0000001, y = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i * 0.5 + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double* out, int iters) {
  double c[8][2];
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, int iters) {
  double c[4][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - 1e-9 * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 3 * i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA stream with R independent DADDs per DMMA interleaved: does the FP64 FMA pipe share issue / execution resources
// with the FP64 tensor path?  (qgemm.cu forms its operand combinations with DADDs inside the main loop.)
template <int R>
__global__ void k_dmma_dadd(double* out, int iters) {
  double c[8][2], s[8];
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; s[i] = 0.5 * i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
      for (int r = 0; r < R; ++r) asm volatile("add.f64 %0, %0, %1;" : "+d"(s[(i + r) & 7]) : "d"(a));
    }
  }
  double t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += c[i][0] + c[i][1] + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <class F>
double timeit(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 4096;
  for (int tpb : {256, 512, 1024}) {
    for (int bps : {1, 2, 4}) {
      if (tpb * bps > 2048) continue;
      const int grid = sms * bps;
      double ms = timeit([&] { k_dfma<<<grid, tpb>>>(out, iters); }, 5);
      double fl = 2.0 * 16 * iters * (double)grid * tpb;
      printf("{\"kernel\":\"dfma\",\"tpb\":%d,\"ctas_per_sm\":%d,\"tflops\":%.2f}\n", tpb, bps, fl / ms * 1e-9);
      ms = timeit([&] { k_dmma884<<<grid, tpb>>>(out, iters); }, 5);
      fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)grid * (tpb / 32);
      printf("{\"kernel\":\"dmma_m8n8k4\",\"tpb\":%d,\"ctas_per_sm\":%d,\"tflops\":%.2f}\n", tpb, bps, fl / ms * 1e-9);
      ms = timeit([&] { k_dmma16816<<<grid, tpb>>>(out, iters); }, 5);
      fl = 2.0 * 16 * 8 * 16 * 4 * iters * (double)grid * (tpb / 32);
      printf("{\"kernel\":\"dmma_m16n8k16\",\"tpb\":%d,\"ctas_per_sm\":%d,\"tflops\":%.2f}\n", tpb, bps, fl / ms * 1e-9);
    }
  }
  for (int bps : {1, 2}) {
    const int tpb = 256, grid = sms * bps;
    double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)grid * (tpb / 32);
    double ms = timeit([&] { k_dmma_dadd<0><<<grid, tpb>>>(out, iters); }, 5);
    printf("{\"kernel\":\"dmma+0dadd\",\"ctas_per_sm\":%d,\"dmma_tflops\":%.2f}\n", bps, fl / ms * 1e-9);
    ms = timeit([&] { k_dmma_dadd<1><<<grid, tpb>>>(out, iters); }, 5);
    printf("{\"kernel\":\"dmma+1dadd\",\"ctas_per_sm\":%d,\"dmma_tflops\":%.2f}\n", bps, fl / ms * 1e-9);
    ms = timeit([&] { k_dmma_dadd<2><<<grid, tpb>>>(out, iters); }, 5);
    printf("{\"kernel\":\"dmma+2dadd\",\"ctas_per_sm\":%d,\"dmma_tflops\":%.2f}\n", bps, fl / ms * 1e-9);
    ms = timeit([&] { k_dmma_dadd<4><<<grid, tpb>>>(out, iters); }, 5);
    printf("{\"kernel\":\"dmma+4dadd\",\"ctas_per_sm\":%d,\"dmma_tflops\":%.2f}\n", bps, fl / ms * 1e-9);
  }
  printf("{\"sms\":%d,\"clock_khz\":%d,\"name\":\"%s\"}\n", sms, p.clockRate, p.name);
  return 0;
}
