// r3d_comm.cu -- in-switch (NVLS) sum all-reduce of the grid gradient over NVLink 5 / NVSwitch.
//
// The path's only exchange is the sum over ranks of the dense grid gradient (1.95 GB at 256^3 /
// degree 2).  A ring all-reduce moves 2*(n-1)/n * bytes per GPU and direction; with NVSwitch multicast objects the
// switch itself can reduce: every rank owns 1/n of the buffer, pulls the SUM of all replicas of its slice with
// multimem.ld_reduce (the switch reads the n replicas and returns one reduced value) and pushes the result to all
// replicas with multimem.st.  Per GPU and direction that is ~1x the bytes instead of 1.75x.
//
// `multicast_ptr` is the multicast virtual address of a buffer that every rank allocated symmetrically and bound to one
// multicast object (PyTorch: torch.distributed._symmetric_memory.empty + rendezvous -> handle.multicast_ptr).  The caller
// brackets the launch with cross-rank barriers on the same stream (all replicas complete before, all slices broadcast
// after); this function only enqueues the reduction kernel.
#include "r3d_host.h"

namespace r3d {

__global__ void __launch_bounds__(512) multimem_allreduce_kernel(float* __restrict__ mc, long long vec_begin, long long vec_end) {
  // 16 bytes per access; UNROLL independent multimem loads in flight per thread before the stores
  constexpr int UNROLL = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = vec_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < vec_end; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const float* p = mc + 4 * (i + u * stride);
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                   : "l"(p)
                   : "memory");
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float* p = mc + 4 * (i + u * stride);
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z),
                   "f"(v[u].w)
                   : "memory");
    }
  }
  for (; i < vec_end; i += stride) {
    float4 v;
    const float* p = mc + 4 * i;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_multimem_all_reduce(void* multicast_ptr, int64_t num_floats, int32_t rank, int32_t world_size, int32_t num_blocks,
                                       void* cuda_stream) {
  if (!multicast_ptr) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_all_reduce: multicast pointer is NULL (no NVLS multicast mapping)");
  if (world_size < 1 || rank < 0 || rank >= world_size) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_all_reduce: bad rank/world_size");
  if (num_floats < 0 || (num_floats % 4) != 0 || !aligned16(multicast_ptr))
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_all_reduce: buffer must be 16-byte aligned and a multiple of 4 floats");
  const long long vecs = num_floats / 4;
  const long long per = (vecs + world_size - 1) / world_size;
  const long long begin = per * rank, end = begin + per < vecs ? begin + per : vecs;
  if (begin >= end) return R3D_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = num_blocks > 0 ? num_blocks : sms * 2;
  const long long need = (end - begin + 511) / 512;
  if (blocks > need) blocks = (int)need;
  multimem_allreduce_kernel<<<blocks, 512, 0, static_cast<cudaStream_t>(cuda_stream)>>>(static_cast<float*>(multicast_ptr), begin, end);
  return check_launch("r3d_multimem_all_reduce");
}
