// Micro-benchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2, new on sm_100) on the B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, float w, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 m = make_float2(w, w);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) {  // 2 scalar FFMA per pair
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i].x) : "f"(w), "f"(acc[(i + 1) & 7].y));
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i].y) : "f"(w), "f"(acc[(i + 1) & 7].x));
        } else {  // one packed FFMA2 per pair
          acc[i] = ffma2(acc[i], m, acc[(i + 1) & 7]);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, int sms) {
  const int iters = 4096, blocks = sms * 8;
  kern<MODE><<<blocks, 256>>>(out, 0.999f, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<MODE><<<blocks, 256>>>(out, 0.999f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)blocks * 256 * iters * 8 * 8 * 2;  // scalar FMAs (lanes)
  printf("{\"bench\": \"%s\", \"ms\": %.3f, \"tfma_per_s\": %.2f, \"fma_per_clk_per_sm_at_1965MHz\": %.1f}\n", name, ms, fma / ms / 1e9,
         fma / (ms * 1e-3) / 1.965e9 / sms);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  run<0>("ffma_scalar", out, sms);
  run<1>("ffma2_packed", out, sms);
  run<0>("ffma_scalar", out, sms);
  run<1>("ffma2_packed", out, sms);
  return 0;
}
