// Stand-alone probe for the tcgen05 building blocks of datapar_tc.cu (run on the GPU box; not part of the library):
// checks the SWIZZLE_NONE descriptors in the three operand-major combinations the kernel uses, the M=64 accumulator
// layout in TMEM, and prints rough MMA / TMEM-load timings.  `tc_probe --mixed` tries a fp16 x bf16 MMA (illegal).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu && ./tools/tc_probe
#include <cuda_fp16.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "../eeyore_b200/csrc/tc05.cuh"

using namespace eb::tc;

struct ProbeSmem {
  alignas(128) uint16_t A[128 * 64];   // 16 KB
  alignas(128) uint16_t B[128 * 64];   // 16 KB (128 x 64 when used as H1, 64 x 64 as weights)
  alignas(128) uint16_t ones[128];     // 256 B of bf16 1.0: an N=8, K=16 K-major operand whatever the k step
  alignas(8) unsigned long long bar[2];
  uint32_t tmem_base;
};

__device__ bool wait_or_timeout(void* bar, uint32_t parity, int* flag) {
  for (long it = 0; it < 20000000L; ++it)
    if (mbar_try_wait(bar, parity)) return true;
  *flag = 1;
  return false;
}

// mode 1: D[128x64] = A(K-major 128x64) * B(K-major 64x64)^T
// mode 2: D[128x64] = A(K-major 128x64) * B(MN-major: buffer [o=64][i=64], N = i, K = o)
// mode 3: D[64x72]  = A(MN-major: buffer [r=128][o=64], M = o, K = r) * B(MN-major buffer [r=128][i=64]) ; + ones -> cols 64..71
__global__ void __launch_bounds__(256, 1) probe_kernel(int mode, const float* gA, const float* gB, float* out, int* flag,
                                                       long long* cycles, int swap_mn) {
  extern __shared__ __align__(128) unsigned char raw[];
  ProbeSmem& s = *reinterpret_cast<ProbeSmem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // global row-major -> core-matrix layout
  const int a_rows = 128, a_cols = 64;
  const int b_rows = (mode == 3 || mode == 5) ? 128 : 64, b_cols = 64;
  const uint32_t csA = 2048, csB = (mode == 3 || mode == 5) ? 2048 : 1024;
  for (int e = tid; e < a_rows * a_cols; e += 256) {
    const int r = e / a_cols, c = e % a_cols;
    *reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s.A) + cm_off(r, c, csA)) =
        (mode == 6) ? __half_as_ushort(__float2half_rn(gA[e])) : (uint16_t)(__float_as_uint(gA[e]) >> 16);
  }
  for (int e = tid; e < b_rows * b_cols; e += 256) {
    const int r = e / b_cols, c = e % b_cols;
    *reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s.B) + cm_off(r, c, csB)) = (uint16_t)(__float_as_uint(gB[e]) >> 16);
  }
  if (tid < 128) s.ones[tid] = 0x3F80;
  if (tid == 0) {
    mbar_init(&s.bar[0], 1);
    mbar_init(&s.bar[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&s.tmem_base);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = s.tmem_base;
  const uint32_t aA = smem_u32(s.A), aB = smem_u32(s.B), aO = smem_u32(s.ones);
  long long t0 = 0, t1 = 0;
  if (warp == 0 && elect_one()) {
    t0 = clock64();
    if (mode == 1) {
      const uint32_t id = idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k)
        mma_bf16(tm, smem_desc(aA + k * 2 * csA, csA, 128), smem_desc(aB + k * 2 * csB, csB, 128), id, k > 0);
    } else if (mode == 2) {
      const uint32_t id = idesc_bf16(128, 64, 0, 1);
      for (int k = 0; k < 4; ++k) {
        const uint64_t bd = swap_mn ? smem_desc(aB + k * 256, csB, 128) : smem_desc(aB + k * 256, 128, csB);
        mma_bf16(tm, smem_desc(aA + k * 2 * csA, csA, 128), bd, id, k > 0);
      }
    } else if (mode == 3) {
      const uint32_t id = idesc_bf16(64, 64, 1, 1), id1 = idesc_bf16(64, 8, 1, 0);
      const uint64_t od = smem_desc(aO, 128, 128);
      for (int k = 0; k < 8; ++k) {
        const uint64_t ad = swap_mn ? smem_desc(aA + k * 256, csA, 128) : smem_desc(aA + k * 256, 128, csA);
        const uint64_t bd = swap_mn ? smem_desc(aB + k * 256, csB, 128) : smem_desc(aB + k * 256, 128, csB);
        mma_bf16(tm, ad, bd, id, k > 0);
        mma_bf16(tm + 64, ad, od, id1, k > 0);
      }
    } else if (mode == 6) {  // mixed formats: A fp16 (K-major), B bf16 (K-major) -- raises "illegal instruction" on B200
      const uint32_t id = idesc_f16kind(128, 64, 0, 0, 0, 1);
      for (int k = 0; k < 4; ++k)
        mma_bf16(tm, smem_desc(aA + k * 2 * csA, csA, 128), smem_desc(aB + k * 2 * csB, csB, 128), id, k > 0);
    } else if (mode == 4 || mode == 5) {  // timing only: 96 back-to-back MMAs
      const uint32_t id = (mode == 4) ? idesc_bf16(128, 64, 0, 0) : idesc_bf16(64, 64, 1, 1);
      const uint64_t a4 = smem_desc(aA, csA, 128), b4 = smem_desc(aB, 1024, 128);
      const uint64_t a5 = smem_desc(aA, 128, csA), b5 = smem_desc(aB, 128, 2048);
#pragma unroll 8
      for (int k = 0; k < 96; ++k) {
        if (mode == 4) mma_bf16(tm, desc_advance(a4, (k & 3) * 2 * csA), desc_advance(b4, (k & 3) * 2 * 1024), id, k > 0);
        else mma_bf16(tm, desc_advance(a5, (k & 7) * 256), desc_advance(b5, (k & 7) * 256), id, k > 0);
      }
    }
    mma_commit(&s.bar[0]);
  }
  const bool ok = wait_or_timeout(&s.bar[0], 0, flag);
  if (t0 != 0) { t1 = clock64(); cycles[0] = t1 - t0; }
  fence_after_sync();
  if (ok) {
    // dump: thread = TMEM lane (32 * (warp % 4) + lane); warps 0-3 columns 0..47, warps 4-7 columns 48..95
    const int q = warp & 3, half = warp >> 2;
    const uint32_t base = tm + ((uint32_t)(32 * q) << 16) + 48 * half;
    long long c0 = clock64();
    uint32_t v[32], w[16];
    tmem_ld32(base, v);
    tmem_ld16(base + 32, w);
    tmem_ld_wait();
    long long c1 = clock64();
    if (tid == 0) cycles[1] = c1 - c0;
    const int tl = 32 * q + lane;
    for (int j = 0; j < 32; ++j) out[tl * 96 + 48 * half + j] = __uint_as_float(v[j]);
    for (int j = 0; j < 16; ++j) out[tl * 96 + 48 * half + 32 + j] = __uint_as_float(w[j]);
    if (mode == 1) {  // tcgen05.st round trip through columns 128 + 32 * half ..
      uint32_t u[32], z[32];
      for (int j = 0; j < 32; ++j) u[j] = (uint32_t)(tid * 1000 + j);
      tmem_st32(tm + ((uint32_t)(32 * q) << 16) + 128 + 32 * half, u);
      tmem_st_wait();
      tmem_ld32(tm + ((uint32_t)(32 * q) << 16) + 128 + 32 * half, z);
      tmem_ld_wait();
      int badst = 0;
      for (int j = 0; j < 32; ++j) badst += (z[j] != u[j]);
      if (badst) atomicAdd(flag + 1, badst);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

static int run_mode(int mode, int swap_mn) {
  const int a_cols = 64;
  const int b_rows = (mode == 3 || mode == 5) ? 128 : 64, b_cols = 64;
  std::vector<float> A(128 * a_cols), B(b_rows * b_cols);
  srand(1234 + mode);
  for (auto& v : A) v = (float)(rand() % 9 - 4);
  for (auto& v : B) v = (float)(rand() % 9 - 4);
  float *dA, *dB, *dO;
  int* dF;
  long long* dC;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dO, 128 * 96 * 4);
  cudaMalloc(&dF, 8);
  cudaMalloc(&dC, 16);
  cudaMemset(dF, 0, 8);
  cudaMemset(dO, 0, 128 * 96 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProbeSmem));
  probe_kernel<<<1, 256, sizeof(ProbeSmem)>>>(mode, dA, dB, dO, dF, dC, swap_mn);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mode %d swap %d: CUDA error %s\n", mode, swap_mn, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> O(128 * 96);
  int flag = 0, flag2[2] = {0, 0};
  long long cyc[2];
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(flag2, dF, 8, cudaMemcpyDeviceToHost);
  flag = flag2[0];
  if (mode == 1) printf("tcgen05.st/ld round trip: %d mismatches\n", flag2[1]);
  cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
  if (flag) {
    printf("mode %d swap %d: TIMEOUT waiting for the MMA commit\n", mode, swap_mn);
    return 3;
  }
  if (mode == 4 || mode == 5) {
    printf("mode %d: 96 MMAs issue->complete %lld cycles (%.1f per MMA); TMEM ld of 48 cols x 256 threads: %lld cycles\n", mode,
           cyc[0], cyc[0] / 96.0, cyc[1]);
    return 0;
  }
  long bad = 0, total = 0;
  auto report = [&](int lane, int col, float exp) {
    const float got = O[lane * 96 + col];
    ++total;
    if (got != exp) {
      if (bad < 6) printf("   mismatch lane %d col %d: got %g expected %g\n", lane, col, got, exp);
      ++bad;
    }
  };
  if (mode == 1 || mode == 6) {
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 64; ++n) {
        float acc = 0;
        for (int k = 0; k < 64; ++k) acc += A[r * 64 + k] * B[n * 64 + k];
        report(r, n, acc);
      }
  } else if (mode == 2) {
    for (int r = 0; r < 128; ++r)
      for (int i = 0; i < 64; ++i) {
        float acc = 0;
        for (int o = 0; o < 64; ++o) acc += A[r * 64 + o] * B[o * 64 + i];
        report(r, i, acc);
      }
  } else {
    for (int o = 0; o < 64; ++o) {
      const int lane = (o & 15) + 32 * (o >> 4);
      for (int i = 0; i < 64; ++i) {
        float acc = 0;
        for (int r = 0; r < 128; ++r) acc += A[r * 64 + o] * B[r * 64 + i];
        report(lane, i, acc);
      }
      float sum = 0;
      for (int r = 0; r < 128; ++r) sum += A[r * 64 + o];
      for (int j = 64; j < 72; ++j) report(lane, j, sum);
    }
  }
  printf("mode %d swap %d: %ld / %ld mismatches; MMA issue->complete %lld cycles, TMEM ld %lld cycles  %s\n", mode, swap_mn, bad,
         total, cyc[0], cyc[1], bad ? "FAIL" : "OK");
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dF); cudaFree(dC);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "--mixed"))   // separate process: the illegal instruction kills the CUDA context
    return run_mode(6, 0);
  int rc = 0;
  rc |= run_mode(1, 0);
  rc |= run_mode(2, 0);
  rc |= run_mode(2, 1);
  rc |= run_mode(3, 0);
  rc |= run_mode(3, 1);
  run_mode(4, 0);
  run_mode(5, 0);
  return rc;
}
