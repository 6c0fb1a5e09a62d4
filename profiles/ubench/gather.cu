// Micro-benchmark: how fast can a B200 SM gather 112-byte voxel records with the access shape of the renderer's forward?
//
// The lane-group forward hands every contributing sample to 8 lanes; lane cj (0..6; lane 7 idles) loads float4 cj of each of the
// sample's 8 corner records: 4 (x, y) columns x 2 z-adjacent records, 8 LDG.128 per lane, four samples per warp instruction.
// This kernel issues exactly that and nothing else (two LOP3 per record keep all 16 bytes of every load alive), at three footprints:
//   l1   every CTA gathers from the same 24 KB of records   -> every line is an L1 hit: the L1 data-pipe rate for this shape
//   l2   a 32 MB volume (L2-resident, little L1 reuse)       -> L2 -> L1 rate for scattered 112-byte records
//   hbm  a 1.9 GB volume, random cells                       -> no locality at all
// and with 16 / 24 / 32 resident warps per SM to show the latency side.  Reported: requested bytes (7 lanes x 16 B x 8 records per
// sample) per second and per clock per SM -- the denominator for "fraction of the SM-side ceiling" in DESIGN.md 4.5.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather.bin gather.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix32(unsigned h) {
  h ^= h >> 16, h *= 0x7FEB352Du, h ^= h >> 15, h *= 0x846CA68Bu, h ^= h >> 16;
  return h;
}

// grid dims of the record volume: [W][D][H] records of 28 floats (112 B), like the renderer's feature storage
struct Vol {
  const float4* feat;
  int W, D, H;
  unsigned mask;  // cells are drawn from [0, mask]^3 (+1 for the neighbouring groups): mask + 3 <= dim
};

template <int LOADS_IN_FLIGHT>
__global__ void __launch_bounds__(128) gather_kernel(Vol v, int iters, float* out) {
  const int lane = threadIdx.x & 31, ms = lane >> 3, cj = lane & 7;
  const bool role_ok = cj < 7;
  const unsigned warp = blockIdx.x * 4u + (threadIdx.x >> 5);
  const unsigned stride4 = 7;  // float4s per record
  unsigned acc = 0u;
  for (int it = 0; it < iters; ++it) {
    // one pseudo-random cell per lane group and iteration (the four groups of a warp stay within a few cells of each other,
    // like the samples of a marching step of an 8x4 pixel tile)
    const unsigned h = mix32(warp * 0x9E3779B9u + (unsigned)it);
    const unsigned x = (h & v.mask) + (ms & 1), y = ((h >> 10) & v.mask) + (ms >> 1), z = (h >> 20) & v.mask;  // x + 1, y + 1, z + 1 < dim
    const unsigned base = (x * v.D + y) * v.H + z;
    const unsigned dy = v.H, dx = v.D * v.H;
    const unsigned rec[8] = {base, base + 1, base + dy, base + dy + 1, base + dx, base + dx + 1, base + dx + dy, base + dx + dy + 1};
    float4 q[8];
    if (role_ok) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4* p = v.feat + (size_t)rec[k] * stride4 + cj;
        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q[k].x), "=f"(q[k].y), "=f"(q[k].z), "=f"(q[k].w) : "l"(p));
      }
      // all four components are consumed (ptxas narrows a vector load whose upper components are dead): two LOP3 per record
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc ^= __float_as_uint(q[k].x) ^ __float_as_uint(q[k].y);
        acc ^= __float_as_uint(q[k].z) ^ __float_as_uint(q[k].w);
      }
    }
  }
  if (acc == 0x12345678u) out[0] = 1.0f;
}

static void run(const char* name, const float4* feat, int dim, unsigned mask, int ctas_per_sm, int sms, float* out) {
  const int iters = 2048, blocks = sms * ctas_per_sm, W = dim, D = dim, H = dim;
  Vol v{feat, W, D, H, mask};
  gather_kernel<8><<<blocks, 128>>>(v, 64, out);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  gather_kernel<8><<<blocks, 128>>>(v, iters, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double samples = (double)blocks * 4 * 4 * iters;  // warps x lane groups x iterations
  const double bytes = samples * 8 * 112;
  printf("{\"bench\": \"%s\", \"volume\": [%d, %d, %d], \"warps_per_sm\": %d, \"ms\": %.3f, \"samples_per_s\": %.3e, \"gather_TB_per_s\": %.2f, "
         "\"bytes_per_clk_per_sm_at_1965MHz\": %.1f}\n",
         name, W, D, H, ctas_per_sm * 4, ms, samples / (ms * 1e-3), bytes / (ms * 1e-3) / 1e12, bytes / (ms * 1e-3) / 1.965e9 / sms);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t big = (size_t)258 * 258 * 258 * 112;  // 1.9 GB: the size of the c3 feature volume
  float4* feat;
  float* out;
  if (cudaMalloc(&feat, big) != cudaSuccess || cudaMalloc(&out, 4) != cudaSuccess) return 1;
  cudaMemset(feat, 0, big);
  for (int ctas : {4, 6, 8}) {
    run("l1", feat, 6, 3u, ctas, sms, out);       // 216 records = 24 KB: L1-resident
    run("l2", feat, 66, 63u, ctas, sms, out);     // 32 MB: L2-resident
    run("hbm", feat, 258, 255u, ctas, sms, out);  // 1.9 GB
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
