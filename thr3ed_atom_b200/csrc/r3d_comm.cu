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

#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------------------------------
// Fused exchange + optimizer: reduce-scatter -> shard-local Adam -> all-gather in ONE kernel (replaces
// the all-reduce followed by torch.optim.Adam.step of reference modules/trainers.py:339-341, :242-245).
//
// Rank r owns slice r of the flat parameter vector.  For every 16 bytes of its slice it
//   1. pulls the SUM over all ranks of the gradient  (multimem.ld_reduce on the gradient's multicast address: the NVSwitch
//      reads the n replicas and returns one reduced value),
//   2. applies Adam with ITS shard of the optimizer state (exp_avg / exp_avg_sq exist once per box, 1/n per GPU),
//   3. pushes the updated PARAMETERS to every replica      (multimem.st on the parameters' multicast address).
// The reduced gradient is never written anywhere; per GPU and direction ~1x the gradient bytes cross NVLink, the dense
// 7-streams-per-element optimizer pass over the whole grid disappears from every GPU, and so do 7/8 of its state.
// The caller brackets the launch with cross-rank barriers (all gradients complete before; all parameters updated after).
// ---------------------------------------------------------------------------------------------------------------------
// summed gradient and both moments of a 16-byte group all zero (voxels no ray of any rank has reached since training began): the Adam
// update is the identity (m' = v' = 0, p' = p - step * 0 / eps = p), so the group needs no store and no broadcast
__device__ __forceinline__ bool still(const float4& G, const float4& M, const float4& V) {
  return G.x == 0.0f && G.y == 0.0f && G.z == 0.0f && G.w == 0.0f && M.x == 0.0f && M.y == 0.0f && M.z == 0.0f && M.w == 0.0f && V.x == 0.0f &&
         V.y == 0.0f && V.z == 0.0f && V.w == 0.0f;
}

template <int UNROLL>
__global__ void __launch_bounds__(512) multimem_adam_kernel(const float* __restrict__ grad_mc, float* __restrict__ param_mc,
                                                            const float* __restrict__ param_local, float* __restrict__ m, float* __restrict__ v,
                                                            long long vec_begin, long long vec_end, float step, float b1, float b2, float eps,
                                                            float bc2_sqrt, float gscale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto update = [&](const float4& G, float4& P, float4& M, float4& V) {
#define R3D_ADAM1(c)                                       \
  {                                                        \
    const float gg = G.c * gscale;                         \
    M.c = fmaf(b1, M.c, (1.0f - b1) * gg);                 \
    V.c = fmaf(b2, V.c, (1.0f - b2) * gg * gg);            \
    P.c -= step * (M.c / (sqrtf(V.c) / bc2_sqrt + eps));   \
  }
    R3D_ADAM1(x) R3D_ADAM1(y) R3D_ADAM1(z) R3D_ADAM1(w)
#undef R3D_ADAM1
  };
  long long i = vec_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < vec_end; i += UNROLL * stride) {
    float4 G[UNROLL], P[UNROLL], M[UNROLL], V[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long j = i + u * stride;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(G[u].x), "=f"(G[u].y), "=f"(G[u].z), "=f"(G[u].w)
                   : "l"(grad_mc + 4 * j)
                   : "memory");
      P[u] = *reinterpret_cast<const float4*>(param_local + 4 * j);
      M[u] = reinterpret_cast<const float4*>(m)[j - vec_begin];
      V[u] = reinterpret_cast<const float4*>(v)[j - vec_begin];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long j = i + u * stride;
      if (still(G[u], M[u], V[u])) continue;  // identity update: nothing to store, nothing to broadcast
      update(G[u], P[u], M[u], V[u]);
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(param_mc + 4 * j), "f"(P[u].x), "f"(P[u].y),
                   "f"(P[u].z), "f"(P[u].w)
                   : "memory");
      reinterpret_cast<float4*>(m)[j - vec_begin] = M[u];
      reinterpret_cast<float4*>(v)[j - vec_begin] = V[u];
    }
  }
  for (; i < vec_end; i += stride) {
    float4 G, P, M, V;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(G.x), "=f"(G.y), "=f"(G.z), "=f"(G.w)
                 : "l"(grad_mc + 4 * i)
                 : "memory");
    P = *reinterpret_cast<const float4*>(param_local + 4 * i);
    M = reinterpret_cast<const float4*>(m)[i - vec_begin];
    V = reinterpret_cast<const float4*>(v)[i - vec_begin];
    if (still(G, M, V)) continue;
    update(G, P, M, V);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(param_mc + 4 * i), "f"(P.x), "f"(P.y), "f"(P.z), "f"(P.w)
                 : "memory");
    reinterpret_cast<float4*>(m)[i - vec_begin] = M;
    reinterpret_cast<float4*>(v)[i - vec_begin] = V;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same fused exchange + optimizer over plain peer-to-peer loads / stores (no multicast object needed).  Rank r reads the
// gradient of its slice from every replica (its own from HBM, the others over NVLink), adds them in rank order, applies
// Adam with its shard of the state and writes the new parameters into every replica.  Per GPU and direction
// 2 * (n-1)/n x the slice-sum of bytes cross NVLink: at n = 2 that is 1.0x the gradient bytes, where the in-switch version
// moves 1.5x (with two ranks the switch pulls the caller's own replica out and back, and multicasts the parameters back to
// their sender) -- so this is the two-rank path; from n = 4 on the in-switch version moves fewer bytes.
// ---------------------------------------------------------------------------------------------------------------------
struct PeerPtrs {
  const float* grad[8];
  float* param[8];
};

template <int UNROLL>
__global__ void __launch_bounds__(512) peer_adam_kernel(const PeerPtrs pp, int world, int rank, float* __restrict__ m, float* __restrict__ v,
                                                        long long vec_begin, long long vec_end, float step, float b1, float b2, float eps,
                                                        float bc2_sqrt, float gscale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto ld_sys = [](const float* p) {  // the peers' gradients were completed before the caller's barrier: system-scope load, never a stale line
    float4 r;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
  };
  long long i = vec_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < vec_end; i += UNROLL * stride) {
    float4 G[UNROLL], P[UNROLL], M[UNROLL], V[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < vec_end) {
        G[u] = ld_sys(pp.grad[0] + 4 * j);
        for (int r = 1; r < world; ++r) {
          const float4 g = ld_sys(pp.grad[r] + 4 * j);
          G[u].x += g.x, G[u].y += g.y, G[u].z += g.z, G[u].w += g.w;
        }
        P[u] = *reinterpret_cast<const float4*>(pp.param[rank] + 4 * j);
        M[u] = reinterpret_cast<const float4*>(m)[j - vec_begin];
        V[u] = reinterpret_cast<const float4*>(v)[j - vec_begin];
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < vec_end && !still(G[u], M[u], V[u])) {
#define R3D_ADAM1(c)                                                \
  {                                                                 \
    const float gg = G[u].c * gscale;                               \
    M[u].c = fmaf(b1, M[u].c, (1.0f - b1) * gg);                    \
    V[u].c = fmaf(b2, V[u].c, (1.0f - b2) * gg * gg);               \
    P[u].c -= step * (M[u].c / (sqrtf(V[u].c) / bc2_sqrt + eps));   \
  }
        R3D_ADAM1(x) R3D_ADAM1(y) R3D_ADAM1(z) R3D_ADAM1(w)
#undef R3D_ADAM1
        for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(pp.param[r] + 4 * j) = P[u];
        reinterpret_cast<float4*>(m)[j - vec_begin] = M[u];
        reinterpret_cast<float4*>(v)[j - vec_begin] = V[u];
      }
    }
  }
}

// slice of rank `rank`, in float4 units: the same partition for the all-reduce and for the fused optimizer
static void slice_of(long long vecs, int rank, int world, long long& begin, long long& end) {
  const long long per = (vecs + world - 1) / world;
  begin = per * rank;
  end = begin + per < vecs ? begin + per : vecs;
  if (begin > vecs) begin = end = vecs;
}

}  // namespace r3d

using namespace r3d;

extern "C" int64_t r3d_multimem_shard_floats(int64_t num_floats, int32_t world_size) {
  if (num_floats < 0 || (num_floats % 4) != 0 || world_size < 1) return -1;
  const long long vecs = num_floats / 4;
  return 4 * ((vecs + world_size - 1) / world_size);
}

extern "C" int r3d_multimem_adam_step(void* grad_multicast_ptr, void* param_multicast_ptr, const float* param_local, float* exp_avg_shard,
                                      float* exp_avg_sq_shard, int64_t num_floats, int32_t rank, int32_t world_size, float lr, float beta1,
                                      float beta2, float eps, float bias_correction1, float bias_correction2, float grad_scale,
                                      int32_t num_blocks, void* cuda_stream) {
  if (!grad_multicast_ptr || !param_multicast_ptr)
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_adam_step: multicast pointer is NULL (no NVLS multicast mapping)");
  if (!param_local || !exp_avg_shard || !exp_avg_sq_shard) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_adam_step: NULL argument");
  if (world_size < 1 || rank < 0 || rank >= world_size) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_adam_step: bad rank/world_size");
  if (num_floats < 0 || (num_floats % 4) != 0 || !aligned16(grad_multicast_ptr) || !aligned16(param_multicast_ptr) || !aligned16(param_local) ||
      !aligned16(exp_avg_shard) || !aligned16(exp_avg_sq_shard))
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_adam_step: buffers must be 16-byte aligned and a multiple of 4 floats");
  if (!(bias_correction1 > 0.f) || !(bias_correction2 > 0.f)) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_adam_step: bias corrections must be positive");
  long long begin, end;
  slice_of(num_floats / 4, rank, world_size, begin, end);
  if (begin >= end) return R3D_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = num_blocks > 0 ? num_blocks : sms * 2;
  const long long need = (end - begin + 511) / 512;
  if (blocks > need) blocks = (int)need;
  // tuning hook: $R3D_MULTIMEM_UNROLL = independent 16-byte multimem loads in flight per thread (default 4)
  static const int unroll = [] {
    const char* e = getenv("R3D_MULTIMEM_UNROLL");
    return e ? atoi(e) : 4;
  }();
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const float step = lr / bias_correction1, bc2s = sqrtf(bias_correction2);
  const float* gmc = static_cast<const float*>(grad_multicast_ptr);
  float* pmc = static_cast<float*>(param_multicast_ptr);
  if (unroll >= 8)
    multimem_adam_kernel<8><<<blocks, 512, 0, st>>>(gmc, pmc, param_local, exp_avg_shard, exp_avg_sq_shard, begin, end, step, beta1, beta2, eps, bc2s, grad_scale);
  else if (unroll <= 2)
    multimem_adam_kernel<2><<<blocks, 512, 0, st>>>(gmc, pmc, param_local, exp_avg_shard, exp_avg_sq_shard, begin, end, step, beta1, beta2, eps, bc2s, grad_scale);
  else
    multimem_adam_kernel<4><<<blocks, 512, 0, st>>>(gmc, pmc, param_local, exp_avg_shard, exp_avg_sq_shard, begin, end, step, beta1, beta2, eps, bc2s, grad_scale);
  return check_launch("r3d_multimem_adam_step");
}

extern "C" int r3d_peer_adam_step(const void* const* grad_ptrs, void* const* param_ptrs, float* exp_avg_shard, float* exp_avg_sq_shard,
                                  int64_t num_floats, int32_t rank, int32_t world_size, float lr, float beta1, float beta2, float eps,
                                  float bias_correction1, float bias_correction2, float grad_scale, int32_t num_blocks, void* cuda_stream) {
  if (!grad_ptrs || !param_ptrs || !exp_avg_shard || !exp_avg_sq_shard) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_peer_adam_step: NULL argument");
  if (world_size < 1 || world_size > 8 || rank < 0 || rank >= world_size)
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_peer_adam_step: bad rank/world_size (1..8 ranks)");
  if (num_floats < 0 || (num_floats % 4) != 0 || !aligned16(exp_avg_shard) || !aligned16(exp_avg_sq_shard))
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_peer_adam_step: buffers must be 16-byte aligned and a multiple of 4 floats");
  PeerPtrs pp{};
  for (int r = 0; r < world_size; ++r) {
    if (!grad_ptrs[r] || !param_ptrs[r] || !aligned16(grad_ptrs[r]) || !aligned16(param_ptrs[r]))
      return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_peer_adam_step: replica %d: NULL or unaligned pointer", r);
    pp.grad[r] = static_cast<const float*>(grad_ptrs[r]);
    pp.param[r] = static_cast<float*>(param_ptrs[r]);
  }
  if (!(bias_correction1 > 0.f) || !(bias_correction2 > 0.f)) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_peer_adam_step: bias corrections must be positive");
  long long begin, end;
  slice_of(num_floats / 4, rank, world_size, begin, end);
  if (begin >= end) return R3D_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = num_blocks > 0 ? num_blocks : sms * 2;
  const long long need = (end - begin + 511) / 512;
  if (blocks > need) blocks = (int)need;
  const float step = lr / bias_correction1, bc2s = sqrtf(bias_correction2);
  peer_adam_kernel<4><<<blocks, 512, 0, static_cast<cudaStream_t>(cuda_stream)>>>(pp, world_size, rank, exp_avg_shard, exp_avg_sq_shard, begin, end, step,
                                                                                  beta1, beta2, eps, bc2s, grad_scale);
  return check_launch("r3d_peer_adam_step");
}

extern "C" int r3d_multimem_all_reduce(void* multicast_ptr, int64_t num_floats, int32_t rank, int32_t world_size, int32_t num_blocks,
                                       void* cuda_stream) {
  if (!multicast_ptr) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_all_reduce: multicast pointer is NULL (no NVLS multicast mapping)");
  if (world_size < 1 || rank < 0 || rank >= world_size) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_all_reduce: bad rank/world_size");
  if (num_floats < 0 || (num_floats % 4) != 0 || !aligned16(multicast_ptr))
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_multimem_all_reduce: buffer must be 16-byte aligned and a multiple of 4 floats");
  long long begin, end;
  slice_of(num_floats / 4, rank, world_size, begin, end);
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
