// FP64 tensor-core (DMMA) throughput probe for sm_100a: mma.sync m8n8k4 / m16n8k8 / m16n8k16, f64 accumulate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu && tools/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int ACC>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed) {
  double a[8], b[4], c[ACC][4];
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = seed * (threadIdx.x * 3 + i);
  for (int k = 0; k < ACC; ++k)
    for (int i = 0; i < 4; ++i) c[k][i] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ACC; ++k) {
      if constexpr (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a[0]), "d"(b[0]));
      } else if constexpr (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
                       "d"(b[2]), "d"(b[3]));
      }
    }
  }
  double s = 0.0;
  for (int k = 0; k < ACC; ++k)
    for (int i = 0; i < 4; ++i) s += c[k][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// plain DFMA for comparison: 8 independent chains per thread
__global__ void __launch_bounds__(256) probe_dfma(double* out, int iters, double seed) {
  double c[8], a = seed * threadIdx.x, b = seed + 1.0;
  for (int i = 0; i < 8; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fma(a, b, c[i]);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  const int iters = 20000;
  const char* names[3] = {"m8n8k4", "m16n8k8", "m16n8k16"};
  const double flop[3] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
  for (int cps = 1; cps <= 4; cps *= 2) {       // CTAs of 256 threads per SM: 8, 16, 32 warps
    const int grid = sms * cps;
    float ms;
#define RUN(S, A)                                                                                         \
    ms = time_ms([&] { probe<S, A><<<grid, 256>>>(out, iters, 1e-9); });                                  \
    printf("%-9s acc=%d warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s  (%.1f cycles per MMA per SM sub-partition at 1.965 GHz)\n", names[S], A, 8 * cps, ms, \
           flop[S] * A * (double)iters * grid * 8 / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / ((double)iters * A * 2 * cps));
    RUN(0, 1) RUN(0, 4) RUN(1, 1) RUN(1, 4) RUN(2, 1) RUN(2, 4)
    ms = time_ms([&] { probe_dfma<<<grid, 256>>>(out, iters * 4, 1e-9); });
    printf("DFMA x8   warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", 8 * cps, ms, 2.0 * 8 * iters * 4.0 * grid * 256 / (ms * 1e-3) / 1e12);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
