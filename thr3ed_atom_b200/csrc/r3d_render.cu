// r3d_render.cu -- fused forward and backward kernels of the SH-voxel-grid renderer (sm_100a).
//
// Kernels in this file
//   render_fwd_group_kernel  forward, default: per-ray maths; every contributing sample is handed to a group of lanes that
//                            gathers its 8 corner records straight from global memory (DUAL: also the band-0 image)
//   render_fwd_coop_kernel   forward, staged variant (A/B): warp-cooperative gather through shared memory (cp.async / TMA)
//   render_fwd_kernel        forward, one thread per ray end to end (band-0 "diffuse" renders, unpadded layouts, A/B)
//   render_bwd_coop_kernel   backward, default: per-ray maths + warp-cooperative, cell-merged scatter; MASK: marches by the
//                            forward's contribution ballots instead of repeating the density probe; DUAL: specular + diffuse
//   render_bwd_kernel        backward, one thread per ray end to end (A/B)
//   mark_touched_kernel      measurement helper (unique voxels referenced by a batch)
// R3dRenderConfig.variant bits (A/B only, -DR3D_AB_VARIANTS): 1 per-ray backward, 2 per-ray forward, 4 TMA-staged forward,
//   8 cp.async-staged forward, 32 warp-specialised forward (+64 sorted, +256/512 L2/L1 prefetch, +1024 dual-stream consumer),
//   128 warp-specialised backward, 2048 lane-group + cell-sorted publish, 4096/8192 lane-group + next-sample prefetch,
//   16384 lane-group + register look-ahead, 32768 two-kernel forward (r3d_fwd_split.cuh; $R3D_SPLIT_MODE picks the gather
//   flavour), 65536 lane-group + density quad volume.  $R3D_FWD_TMA=1|2: TMA density bricks.
//
// Common structure.  One thread owns one ray and marches it front to back.  Per sample it
//   1. forms the sample position exactly as the reference does (sample.py:54-67),
//   2. applies the strict inside-AABB test (voxels.py:252-274 / process.py:80-84),
//   3. gathers the 8 corner densities (a [W][D][H] fp32 volume: 4 B/voxel, L2 resident up to 256^3+),
//      interpolates, scales and activates them (voxels.py:292-309),
//   4. ONLY IF the sample's density is non-zero needs the 8 corner SH records; each record is contracted with
//      the ray's SH basis first (the contraction is linear, so it commutes with the trilinear weights:
//      3 accumulators instead of 3*(deg+1)^2), then sigmoid and compositing
//      (spherical_harmonics.py:86-116, accumulate.py:43-88).
// A sample with sigma == 0 has alpha == 0 exactly, hence weight 0 and no effect on colour, depth,
// acc or the transmittance, and (ReLU' = 0) no gradient: skipping its feature traffic is exact.
// Likewise a ray whose transmittance reached exactly 0 can stop.
//
// The backward kernels re-march the ray (nothing of size N*S is saved except the optional sample cache of
// (sigmoid(raw), sigma) records written by the forward), rebuild alpha/T, and use the forward's outputs for the
// suffix sums:
//   dL/dsigma_i = delta_i * ( T_{i+1} q_i - sum_{j>i} w_j q_j ),   sum_{j>i} = Total - prefix_i
//   q_i = g_c . sigmoid(raw_i) + g_d z_i + g_a,  Total = g_c . C_fg + g_d depth + g_a acc
// (SURVEY.md A.6).  Gradients are scattered with 128-bit vector reductions (red.global.add.v4.f32).
//
// The cooperative kernels exist because of what the profiles showed: the path is bound inside the
// SM -- L1 data pipe (one wavefront per distinct 128-byte line per request), issue slots -- not by HBM; they make
// consecutive lanes cover consecutive bytes of ONE voxel record, merge samples that share an interpolation cell
// (backward) and keep all 32 lanes on contributing samples (forward).
#include "r3d_host.h"

#include <cuda.h>  // CUtensorMap (type and enums only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

// resident CTAs per SM the cooperative kernels are compiled for (register budget = 65536 / (128 * blocks)).
// Measured at c3 on the B200: backward 7.95 ms at 4 CTAs/SM, 7.45 ms at 5 (96 registers, a few bytes of spill);
// forward unchanged between 4 and 5.
#ifndef R3D_FWD_BLOCKS
#define R3D_FWD_BLOCKS 4
#endif
// forward lane-group kernel: software-pipelined record loads (two register buffers)
#ifndef R3D_FWD_PIPE
#define R3D_FWD_PIPE 0  // measured at c3: 5.44 ms pipelined vs 4.41 ms (4 CTAs/SM), 5.40 vs 5.29 ms (3 CTAs/SM)
#endif
// forward lane-group kernel: packed FFMA2 in the record accumulation (bit-identical; measured at c3: 4.43 -> 4.37 ms)
#ifndef R3D_FWD_FFMA2
#define R3D_FWD_FFMA2 1
#endif
#ifndef R3D_BWD_BLOCKS
#define R3D_BWD_BLOCKS 5
#endif
// mask-path backward: request the next marching step's per-sample record one step ahead
#ifndef R3D_BWD_PREFETCH
#define R3D_BWD_PREFETCH 1
#endif
// backward of the single-pass specular + diffuse render: at 5 CTAs/SM (96 registers) it spills 172 bytes; measured
// 12.52 vs 12.90 ms (c3 step) and 3.07 vs 3.29 ms (32768-ray trainer step) in favour of 4 CTAs/SM (128 registers)
#ifndef R3D_BWD_DUAL_BLOCKS
#define R3D_BWD_DUAL_BLOCKS 4
#endif

namespace r3d {

struct OutP {
  float* __restrict__ colour;
  float* __restrict__ depth;
  float* __restrict__ acc;
  float* __restrict__ disparity;
  float4* __restrict__ cache;  // optional [S][N] (sigmoid(raw) rgb, sigma) of every sigma != 0 sample, for the backward
  float* __restrict__ colour_diffuse;  // single-pass specular + diffuse render: band-0 image of the same samples
  float4* __restrict__ cache_diffuse;  // optional [S][N] (sigmoid(raw_diffuse) rgb, -)
  unsigned* __restrict__ mask;         // optional [S][warps of the launch]: ballot of the contributing rays per marching step
};

struct BwdP {
  const float* __restrict__ colour;  // saved forward outputs
  const float* __restrict__ depth;
  const float* __restrict__ acc;
  const float* __restrict__ g_colour;  // upstream grads (nullable)
  const float* __restrict__ g_depth;
  const float* __restrict__ g_acc;
  const float* __restrict__ g_disp;
  float* __restrict__ gdens;  // outputs (nullable)
  float* __restrict__ gfeat;
  const float4* __restrict__ cache;  // optional sample cache written by the forward pass
  const float* __restrict__ colour_diffuse;    // single-pass specular + diffuse render: saved band-0 image,
  const float* __restrict__ g_colour_diffuse;  //   its upstream gradient (nullable)
  const float4* __restrict__ cache_diffuse;    //   and its per-sample records
  const unsigned* __restrict__ mask;           // optional contribution ballots written by the forward (see OutP)
};

struct RayCtx {
  Ray r;
  float dnorm;
  DepthGen dg;
  int i_lo, i_hi;
};

__device__ __forceinline__ void setup_ray(const GridP& g, const RaysP& rp, const CfgP& c, long long ray, RayCtx& s,
                                          float& vx, float& vy, float& vz) {
  s.r = load_ray(rp, ray);
  const Ray& r = s.r;
  // ||d||: accumulate.py:55 / process.py:53
  s.dnorm = sqrtf(fmaf(r.dz, r.dz, fmaf(r.dy, r.dy, r.dx * r.dx)));
  vx = __fdiv_rn(r.dx, s.dnorm), vy = __fdiv_rn(r.dy, s.dnorm), vz = __fdiv_rn(r.dz, s.dnorm);
  float near = c.near, far = c.far;
  if (rp.bounds) {
    near = __ldg(rp.bounds + 2 * ray), far = __ldg(rp.bounds + 2 * ray + 1);
  } else if (c.flags & R3D_FLAG_OPTIMIZED_SAMPLING) {
    reference_aabb_bounds(g, r, c.near, c.far, near, far);
  }
  s.dg.near = near, s.dg.far = far;
  s.dg.S = c.S, s.dg.half = c.S / 2;
  s.dg.step = c.S > 1 ? __fdiv_rn(1.0f, (float)(c.S - 1)) : 0.0f;
  s.dg.perturb = (c.flags & R3D_FLAG_PERTURB) != 0;
  s.dg.jit = c.jitter ? c.jitter + (size_t)ray * c.S : nullptr;
  s.dg.key = ray_rng_key(c.seed_lo, c.seed_hi, ray);
  sample_range(g, r, near, far, c.S, s.i_lo, s.i_hi);
}

__device__ __forceinline__ void ldg256(const float* __restrict__ p, float* v) {
  asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
      : "l"(p));
}

// ---- per-corner SH record contraction (forward) -------------------------------------------------
template <int DEG, int VEC>
__device__ __forceinline__ void corner_radiance(const float* __restrict__ rec, const float (&Y)[(DEG + 1) * (DEG + 1)],
                                                bool diffuse, float w, float& r, float& g, float& b) {
  constexpr int K = (DEG + 1) * (DEG + 1);
  constexpr int F = 3 * K;
  if (DEG > 0 && diffuse) {  // SH band 0 only (process.py:59-63)
    const float wy = w * Y[0];
    r = fmaf(wy, __ldg(rec), r);
    g = fmaf(wy, __ldg(rec + K), g);
    b = fmaf(wy, __ldg(rec + 2 * K), b);
    return;
  }
  constexpr int NV = (F + 7) / 8 * 2;  // float4 slots, rounded so that both vector widths fit
  float v[NV * 4];
  if constexpr (VEC == 8) {
    // Blackwell 256-bit global load (LDG.E.256): one request per 32-byte sector of the record
#pragma unroll
    for (int j = 0; j < (F + 7) / 8; ++j) ldg256(rec + 8 * j, &v[8 * j]);
  } else if constexpr (VEC == 4) {
#pragma unroll
    for (int j = 0; j < (F + 3) / 4; ++j) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(rec) + j);
      v[4 * j] = q.x, v[4 * j + 1] = q.y, v[4 * j + 2] = q.z, v[4 * j + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < F; ++e) v[e] = __ldg(rec + e);
  }
  float sr = 0.f, sg = 0.f, sb = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    sr = fmaf(Y[k], v[k], sr);
    sg = fmaf(Y[k], v[K + k], sg);
    sb = fmaf(Y[k], v[2 * K + k], sb);
  }
  r = fmaf(w, sr, r), g = fmaf(w, sg, g), b = fmaf(w, sb, b);
}

template <int DEG, int VEC>
__device__ __forceinline__ void gather_radiance(const GridP& g, const Cell& c, const float (&Y)[(DEG + 1) * (DEG + 1)],
                                                bool diffuse, float& rr, float& rg, float& rb) {
  rr = rg = rb = 0.f;
#pragma unroll
  for (int ix = 0; ix < 2; ++ix)
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
      const float wxy = c.wx[ix] * c.wy[iy];
      const size_t col = (size_t)(c.ox[ix] + c.oy[iy]);
#pragma unroll
      for (int iz = 0; iz < 2; ++iz) {
        const float w = wxy * c.wz[iz];
        corner_radiance<DEG, VEC>(g.feat + (col + c.oz[iz]) * (size_t)g.stride, Y, diffuse, w, rr, rg, rb);
      }
    }
}

// shared-memory reads at a precomputed 32-bit shared address (the member sweep of the cooperative backward forms its
// addresses as base - index * stride: one IMAD per read)
__device__ __forceinline__ float lds_f32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lds_v2(unsigned a, float& x, float& y) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void lds_v4(unsigned a, float& x, float& y, float& z, float& w) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a) : "memory");
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- per-corner gradient scatter (backward) -----------------------------------------------------
// d coeff[ch][k] = Y_k * d raw_ch (SH is linear), times the corner's trilinear weight.
template <int DEG, int VEC>
__device__ __forceinline__ void corner_scatter(float* __restrict__ rec, const float (&Y)[(DEG + 1) * (DEG + 1)],
                                               bool diffuse, float w, const float (&draw)[3]) {
  constexpr int K = (DEG + 1) * (DEG + 1);
  constexpr int F = 3 * K;
  const float wd[3] = {w * draw[0], w * draw[1], w * draw[2]};
  if (DEG > 0 && diffuse) {
    atomicAdd(rec, wd[0] * Y[0]);
    atomicAdd(rec + K, wd[1] * Y[0]);
    atomicAdd(rec + 2 * K, wd[2] * Y[0]);
    return;
  }
  if constexpr (VEC != 0) {
    constexpr int NV = (F + 3) / 4;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float q[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const int e = 4 * j + l;
        q[l] = (e < F) ? wd[e / K] * Y[e % K] : 0.0f;
      }
      red_add_v4(rec + 4 * j, q[0], q[1], q[2], q[3]);
    }
  } else {
#pragma unroll
    for (int e = 0; e < F; ++e) atomicAdd(rec + e, wd[e / K] * Y[e % K]);
  }
}

// lane layout shared by the warp-cooperative kernels: LPR lanes per voxel record, one float4 each
template <int DEG>
struct CoopShape {
  static constexpr int K = (DEG + 1) * (DEG + 1);
  static constexpr int F = 3 * K;
  static constexpr int NV = (F + 3) / 4;                              // float4s of a record that carry data
  static constexpr int LPR = NV <= 1 ? 1 : (NV <= 4 ? 4 : (NV <= 8 ? 8 : 16));  // lanes per record
  static constexpr int CPP = (32 / LPR) < 8 ? (32 / LPR) : 8;        // corners per pass
  static constexpr int PASSES = 8 / CPP;
  static constexpr int PROW = (NV % 2) ? 4 * NV : 4 * NV + 4;         // row stride: odd multiple of 4 floats => conflict-free
  static constexpr int WROW = 12;                                     // 8 used; 12 keeps 128-bit stores conflict-free
  // Row of the backward's published weights: [0..7] the 8 corner weights ordered so that the PASSES weights one lane needs
  // are adjacent (one vector read), [8..15] weight * dL/dsigma_pre per corner (the density gradient needs no second
  // table), 4 floats of padding (stride 20 floats = 5 x 16 bytes, odd: conflict-free 128-bit stores).
  static constexpr int WROW2 = 20;
  __host__ __device__ static constexpr int wpos(int k) { return (k % CPP) * PASSES + (k / CPP); }
};

// =================================================================================================
// forward
// =================================================================================================
template <int DEG, int VEC>
__global__ void __launch_bounds__(128) render_fwd_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  if (ray < 0) return;

  RayCtx s;
  float vx, vy, vz;
  setup_ray(g, rp, c, ray, s, vx, vy, vz);
  constexpr int K = (DEG + 1) * (DEG + 1);
  float Y[K];
  sh_basis<DEG>(vx, vy, vz, Y);
  const bool diffuse = (c.flags & R3D_FLAG_DIFFUSE) != 0;
  const Ray& r = s.r;

  float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, acc = 0.f;
  if (s.i_lo <= s.i_hi) {
    DepthMarch dm;
    dm.start(s.dg, s.i_lo);
    float z = dm.next(s.dg, s.i_lo);
    for (int i = s.i_lo; i <= s.i_hi; ++i) {
      const bool last = (i == c.S - 1);
      const float zn = last ? 0.0f : dm.next(s.dg, i + 1);
      const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));  // sample.py:67
      const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
      const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
      if (inside_aabb(g, px, py, pz)) {
        Cell cell;
        make_cell_inside(g, px, py, pz, cell);
        float dpost;
        const float sigma = density_post(g.post, density_pre_interp(g, cell), dpost);
        if (sigma != 0.0f) {
          // accumulate.py:49-55, :24-28
          const float delta = last ? __fmul_rn(kInfinity, s.dnorm) : __fmul_rn(__fsub_rn(zn, z), s.dnorm);
          const float alpha = 1.0f - exp_neg(sigma * delta);
          const float w = alpha * T;
          float rr, rg, rb;
          gather_radiance<DEG, VEC>(g, cell, Y, diffuse, rr, rg, rb);
          const float sr = sigmoidf_(rr), sg = sigmoidf_(rg), sb = sigmoidf_(rb);
          if (out.cache) out.cache[(size_t)i * rp.n + ray] = make_float4(sr, sg, sb, sigma);
          cr = fmaf(w, sr, cr);
          cg = fmaf(w, sg, cg);
          cb = fmaf(w, sb, cb);
          dep = fmaf(w, z, dep);
          acc += w;
          T *= (1.0f - alpha);
          if (T == 0.0f) break;  // every later weight is alpha*0 = 0 exactly
        }
      }
      z = zn;
    }
  }
  if (c.flags & R3D_FLAG_WHITE_BKGD) {  // accumulate.py:77-81
    const float bg = 1.0f - acc;
    cr += bg, cg += bg, cb += bg;
  }
  out.colour[3 * ray] = cr, out.colour[3 * ray + 1] = cg, out.colour[3 * ray + 2] = cb;
  out.depth[ray] = dep;
  out.acc[ray] = acc;
  if (out.disparity) {  // accumulate.py:85-88; 0/0 stays NaN through torch.maximum
    const float ratio = __fdiv_rn(dep, acc);
    const float m = (ratio != ratio) ? ratio : fmaxf(kZeroPlus, ratio);
    out.disparity[ray] = __fdiv_rn(1.0f, m);
  }
}


#ifdef R3D_AB_VARIANTS  // staged forward variants (cp.async / TMA bulk copies through shared memory): measurement only
// =================================================================================================
// forward, warp-cooperative gather (default for the padded layouts, non-diffuse)
//
// Profile of the thread-per-ray gather (profiles/r01_v0_ncu_full_summary.md, l1tex__data_pipe_lsu_wavefronts 94 %):
// the L1 data pipe spends one wavefront per distinct 128-byte line a request touches; with one ray per lane a 32-lane
// LDG.128 touches ~10 voxel records to deliver 16 bytes from each, 56 times per marching step.  Here the warp first
// finds the distinct interpolation cells among its contributing samples (__match_any_sync), copies each such cell's 8
// corner records global -> shared, consecutive lanes covering consecutive 16-byte pieces of one record
// (a record costs one or two wavefronts instead of one per lane per piece), and then every ray reads its cell's
// records from shared memory.  The maths per ray is unchanged (same order of operations as render_fwd_kernel).
// =================================================================================================
template <int DEG>
struct FwdStageShape {
  using S = CoopShape<DEG>;
  static constexpr int SLOTS = DEG >= 3 ? 4 : 8;       // distinct cells staged per round (static smem <= 48 KB)
  static constexpr int REC = 4 * S::NV;                // floats staged per record
  static constexpr int SLOT = 8 * REC + 4;             // (SLOT / 4) odd: 128-bit reads of different slots hit different banks
  static constexpr int PER_SLOT = 8 * S::NV;           // float4 copies per cell
};

template <int DEG>
struct alignas(16) FwdStageSmem {  // one per warp: keep every warp's tables 16-byte aligned
  using H = FwdStageShape<DEG>;
  float rec[H::SLOTS * H::SLOT];
  int vox[H::SLOTS * 8];
  unsigned long long mbar;  // completion barrier of the TMA (cp.async.bulk) staging variant
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) with mbarrier completion ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int DEG, bool TMA>
__global__ void __launch_bounds__(128, DEG >= 3 ? 3 : R3D_FWD_BLOCKS) render_fwd_coop_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out) {
  using H = FwdStageShape<DEG>;
  using S = CoopShape<DEG>;
  constexpr int K = S::K, NV = S::NV;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) FwdStageSmem<DEG> smem_all[4];
  FwdStageSmem<DEG>& sm = smem_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  unsigned tma_phase = 0;
  if constexpr (TMA) {
    if (lane == 0) mbar_init(&sm.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }

  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  bool alive = ray >= 0;
  RayCtx s;
  float Y[K];
  s.i_lo = 1, s.i_hi = 0;
  if (alive) {
    float vx, vy, vz;
    setup_ray(g, rp, c, ray, s, vx, vy, vz);
    sh_basis<DEG>(vx, vy, vz, Y);
  }
  const Ray& r = s.r;
  bool marching = alive && s.i_lo <= s.i_hi;
  int lo = marching ? s.i_lo : 0x7fffffff, hi = marching ? s.i_hi : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(FULL, lo, o));
    hi = max(hi, __shfl_xor_sync(FULL, hi, o));
  }

  // staging role of this lane: float4 `cj` of corner `pass * CPP + cq`
  const int cq = lane / S::LPR, cj = lane % S::LPR;
  const bool role_ok = (cq < S::CPP) && (cj < S::NV);

  float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, acc = 0.f;
  float z = 0.f;
  bool have_z = false;
  DepthMarch dm;
  dm.bm = dm.bc = 0.f;
  for (int i = lo; i <= hi; ++i) {
    // ---- per-lane: position, inside test, cell, density ----
    bool contributes = false;
    float wc[8];
    int vox[8];
    float sigma = 0.f, zn = 0.f;
    bool last = false;
    unsigned long long key = 0ull;  // 64-bit: voxel offset (up to 2^31) + the low corner's validity pattern
    const bool mine = marching && i >= s.i_lo && i <= s.i_hi;
    if (mine) {
      if (!have_z) dm.start(s.dg, i), z = dm.next(s.dg, i), have_z = true;
      last = (i == c.S - 1);
      zn = last ? 0.0f : dm.next(s.dg, i + 1);
      const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
      const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
      const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
      if (inside_aabb(g, px, py, pz)) {
        Cell cell;
        make_cell_inside(g, px, py, pz, cell);
        float dpost;
        sigma = density_post(g.post, density_pre_interp(g, cell), dpost);
        if (sigma != 0.0f) {
          contributes = true;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
            wc[k] = cell.wx[ix] * cell.wy[iy] * cell.wz[iz];
            vox[k] = cell.ox[ix] + cell.oy[iy] + cell.oz[iz];
          }
          key = (unsigned long long)(unsigned)vox[0] | ((unsigned long long)((cell.wx[0] != 0.f) | ((cell.wy[0] != 0.f) << 1) | ((cell.wz[0] != 0.f) << 2)) << 32);
        }
      }
    }
    const unsigned act = __ballot_sync(FULL, contributes);
    if (act != 0u) {
      // ---- group the contributing samples by cell; every distinct cell gets a staging slot ----
      unsigned peers = 0u;
      if (contributes) peers = __match_any_sync(act, key);
      const int leader_lane = contributes ? (__ffs(peers) - 1) : 0;
      const bool leader = contributes && leader_lane == lane;
      const unsigned leaders = __ballot_sync(FULL, leader);
      const int num_cells = __popc(leaders);
      const int my_slot = __popc(leaders & ((1u << leader_lane) - 1u));
      float rr = 0.f, rg = 0.f, rb = 0.f;
      for (int sb = 0; sb < num_cells; sb += H::SLOTS) {
        const int n = min(H::SLOTS, num_cells - sb);
        if (leader && my_slot >= sb && my_slot < sb + H::SLOTS) {
          int* v = sm.vox + (my_slot - sb) * 8;
          *reinterpret_cast<int4*>(v) = make_int4(vox[0], vox[1], vox[2], vox[3]);
          *reinterpret_cast<int4*>(v + 4) = make_int4(vox[4], vox[5], vox[6], vox[7]);
        }
        __syncwarp();
        // ---- stage n cells x 8 records with cp.async (all copies of the round in flight, one wait).  Fixed lane
        //      roles (LPR lanes per record, one 16-byte piece each, CPP records per request): a request reads whole
        //      records with consecutive lanes and lands them contiguously in shared memory.
        if constexpr (TMA) {
          // one TMA bulk copy per record (REC*4 bytes, 16-byte aligned on both sides), completion counted on the
          // warp's mbarrier: the staging traffic goes L2 -> shared memory without passing through the LSU data pipe
          constexpr unsigned kRecBytes = H::REC * 4;
          if (lane == 0) mbar_expect_tx(&sm.mbar, (unsigned)n * 8u * kRecBytes);
          __syncwarp();
          for (int q = lane; q < n * 8; q += 32) {
            const int slot = q >> 3, corner = q & 7;
            tma_bulk_g2s(sm.rec + slot * H::SLOT + corner * H::REC,
                         g.feat + (size_t)(unsigned)sm.vox[slot * 8 + corner] * (size_t)(unsigned)g.stride, kRecBytes, &sm.mbar);
          }
          mbar_wait(&sm.mbar, tma_phase);
          tma_phase ^= 1u;
        } else {
          if (role_ok) {
            for (int slot = 0; slot < n; ++slot) {
#pragma unroll
              for (int pass = 0; pass < S::PASSES; ++pass) {
                const int corner = pass * S::CPP + cq;
                cp_async16(sm.rec + slot * H::SLOT + corner * H::REC + 4 * cj,
                           g.feat + (size_t)(unsigned)sm.vox[slot * 8 + corner] * (size_t)(unsigned)g.stride + 4 * cj);
              }
            }
          }
          cp_async_wait_all();
        }
        __syncwarp();
        // ---- every ray whose cell is staged contracts its 8 corner records with its own SH basis ----
        if (contributes && my_slot >= sb && my_slot < sb + H::SLOTS) {
          const float* base = sm.rec + (my_slot - sb) * H::SLOT;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float* rec = base + k * H::REC;
            float v[4 * NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
              const float4 q4 = *reinterpret_cast<const float4*>(rec + 4 * j);
              v[4 * j] = q4.x, v[4 * j + 1] = q4.y, v[4 * j + 2] = q4.z, v[4 * j + 3] = q4.w;
            }
            float sr = 0.f, sg = 0.f, sbl = 0.f;
#pragma unroll
            for (int kk = 0; kk < K; ++kk) {
              sr = fmaf(Y[kk], v[kk], sr);
              sg = fmaf(Y[kk], v[K + kk], sg);
              sbl = fmaf(Y[kk], v[2 * K + kk], sbl);
            }
            rr = fmaf(wc[k], sr, rr), rg = fmaf(wc[k], sg, rg), rb = fmaf(wc[k], sbl, rb);
          }
        }
        if constexpr (TMA) fence_proxy_async_smem();  // generic-proxy reads above, async-proxy writes next
        __syncwarp();  // the slots are overwritten by the next round / the next marching step
      }
      if (contributes) {
        const float delta = last ? __fmul_rn(kInfinity, s.dnorm) : __fmul_rn(__fsub_rn(zn, z), s.dnorm);
        const float alpha = 1.0f - exp_neg(sigma * delta);
        const float w = alpha * T;
        const float sr = sigmoidf_(rr), sg = sigmoidf_(rg), sb2 = sigmoidf_(rb);
        if (out.cache) out.cache[(size_t)i * rp.n + ray] = make_float4(sr, sg, sb2, sigma);
        cr = fmaf(w, sr, cr);
        cg = fmaf(w, sg, cg);
        cb = fmaf(w, sb2, cb);
        dep = fmaf(w, z, dep);
        acc += w;
        T *= (1.0f - alpha);
        if (T == 0.0f) marching = false;  // every later weight is alpha*0 = 0 exactly
      }
    }
    if (mine) z = zn;
  }
  if (!alive) return;
  if (c.flags & R3D_FLAG_WHITE_BKGD) {
    const float bg = 1.0f - acc;
    cr += bg, cg += bg, cb += bg;
  }
  out.colour[3 * ray] = cr, out.colour[3 * ray + 1] = cg, out.colour[3 * ray + 2] = cb;
  out.depth[ray] = dep;
  out.acc[ray] = acc;
  if (out.disparity) {
    const float ratio = __fdiv_rn(dep, acc);
    const float m = (ratio != ratio) ? ratio : fmaxf(kZeroPlus, ratio);
    out.disparity[ray] = __fdiv_rn(1.0f, m);
  }
}

#endif  // R3D_AB_VARIANTS

// =================================================================================================
// forward, lane-group gather (default for the padded layouts, non-diffuse)
//
// Profile of render_fwd_coop_kernel (profiles/r01_v3_ncu_full_summary.md): 1.08 G shared-memory wavefronts per launch =
// 75 % of all SM cycles.  Every contributing ray pulled its 8 x 112-byte records out of shared memory by itself
// (56 LDS.128 per marching step, ~half the lanes idle because their sample has sigma == 0, and rays of the same cell
// reading the same bytes again), on top of the staging writes.  Here a contributing sample is instead handed to a GROUP
// of LPR consecutive lanes (8 at degree 2): lane `cj` of the group loads float4 `cj` of each of the 8 corner records
// straight from global memory (one coalesced 112-byte read per record and group, 32/LPR samples per request), applies
// the sample's 8 trilinear weights and its ray's SH basis to its 4 elements, and the group reduces the three channel
// sums with 4 packed shuffles.  Nothing of the record passes through shared memory; what does is per SAMPLE, not per
// record: 8 weights + 8 voxel offsets published by the owning lane, one float4 of the ray's (per-kernel constant)
// expanded SH table per group lane, and the 3 results back.  All 32 lanes work on contributing samples only.
// =================================================================================================
template <int DEG>
struct FwdGroupShape {
  using S = CoopShape<DEG>;
  static constexpr int LPR = S::LPR;        // lanes per sample (1, 4, 8, 16 for degree 0..3)
  static constexpr int MPI = 32 / LPR;      // samples per iteration of a warp
  static constexpr int YROW = 4 * S::NV;    // expanded SH row: Y[e % K] for record element e < F, 0 for the pad
};

template <int DEG, bool DUAL>
struct alignas(16) FwdGroupSmem {  // one per warp
  using H = FwdGroupShape<DEG>;
  float Y[32 * H::YROW];  // row per lane (= ray); written once per kernel
  float W[32 * 8];        // rows per contributing sample of the current marching step (rank order)
  unsigned V[32 * 8];     // corner record indices in float4 units
  float R[32 * 4];        // raw radiance (r, g, b, -) per contributing sample
  float R2[DUAL ? 32 * 4 : 4];  // raw band-0 ("diffuse") radiance of the same samples (single-pass specular + diffuse render)
  int src[32];            // owning lane of each rank
};

// DUAL: also produce the band-0 ("diffuse", process.py:59-63) image of the same samples (reference trainer:
// modules/trainers.py:306-330 renders every batch twice).  raw_diffuse[ch] = C0 * coeff[ch][0] is the k = 0 element of
// the record the specular render interpolates anyway: the lanes that hold elements 0, K, 2K hand it over before the SH
// weighting, everything else (samples, density, weights, depth, acc) is shared.
// SORT: publish the contributing samples grouped by interpolation cell instead of in lane order.  Consecutive slots are
// gathered by DIFFERENT lane groups in the SAME load instruction, and lanes that name the same address are served by
// one L1 wavefront: a cell shared by several samples of a marching step then crosses the L1 data pipe once per
// instruction instead of once per sample.
// PF: software prefetch.  The gather of a marching step waits for its slowest sector, and ~8 % of a step's sectors are first
// touches that come from DRAM (L2 hit rate 46 %): every lane-group iteration then costs a full DRAM latency although HBM is
// 90 % idle.  The NEXT sample's position is known one step ahead (its depth is already computed for this sample's interval),
// so each lane requests the lines of the cell it will land in -- 4 (x, y) columns of two z-adjacent records = 224 contiguous
// bytes each -- into L2 (PF >= 1) and, with PF == 2, the 8 density words into L1, a whole marching step before they are
// needed.  The cell is computed without the reference's exact rounding: a prefetch is a hint, results cannot change.
//
// TMA: density bricks through the Tensor Memory Accelerator.  The probe is the other global-memory round trip of a marching
// step (8 scattered 4-byte loads per lane, ~400 cycles on an L2 hit) and on trained grids -- mostly empty space -- it is the
// whole forward.  The next sample's depth is already known when a step starts, so the warp computes the cells its 32 rays
// will land in at step i + 1, takes their bounding box (redux.sync min / max) and one elected lane issues ONE
// cp.async.bulk.tensor.3d copy (SASS UTMALDG) of that kTmaBox brick of the density volume into a double-buffered
// shared-memory slot of the warp (1.1 KB), completion on an mbarrier.  At step i + 1 the 8 corner densities are 8 LDS at
// immediate offsets from one base: no address arithmetic, no clamping, no validity logic -- out-of-range box elements are
// zero-filled by the TMA unit, which IS grid_sample's zero padding (voxels.py:296-303) -- and the position / cell arithmetic
// of step i + 1 is the arithmetic the look-ahead already did.  The same box coordinates drive one
// cp.async.bulk.prefetch.tensor.4d (SASS UTMAPF) of the feature records into L2, replacing the per-lane prefetch of PF
// (which cost 15 % more instructions than it saved).  Boxes that do not fit (1 % of the steps at the BASELINE shapes, every
// step of an incoherent ray batch) and the first step of a ray fall back to the direct loads; values are bit-identical
// either way (same 8 numbers, same order of accumulation).
constexpr int kTmaBoxX = 6, kTmaBoxY = 6, kTmaBoxZ = 8;  // voxels; z is the contiguous axis: 8 floats = 32 bytes per box row

__device__ __forceinline__ void tma_load_box3(void* smem_dst, const CUtensorMap* map, int cz, int cy, int cx, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(map), "r"(cz), "r"(cy), "r"(cx), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_box4(const CUtensorMap* map, int c0, int cz, int cy, int cx) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(cz), "r"(cy), "r"(cx) : "memory");
}
__device__ __forceinline__ void tma_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  for (unsigned spin = 0;; ++spin) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 24)) __trap();  // a protocol bug must be a CUDA error, never a hung GPU
  }
}

template <int DEG, bool DUAL, bool SORT, int PF, bool TMA, bool LA = false, bool DQ = false>
__device__ __forceinline__ void fwd_group_body(const GridP& g, const RaysP& rp, const CfgP& c, const OutP& out, const CUtensorMap* dmap,
                                               const CUtensorMap* fmap) {
  using H = FwdGroupShape<DEG>;
  using S = CoopShape<DEG>;
  constexpr int K = S::K, F = S::F, NV = S::NV, LPR = H::LPR, MPI = H::MPI;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) FwdGroupSmem<DEG, DUAL> smem_all[4];
  FwdGroupSmem<DEG, DUAL>& sm = smem_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  constexpr int kBoxFloats = kTmaBoxX * kTmaBoxY * kTmaBoxZ;
  __shared__ __align__(128) float dbox_all[TMA ? 4 * 2 * kBoxFloats : 1];
  __shared__ __align__(8) unsigned long long dbar_all[TMA ? 8 : 1];
  float* const dbox = dbox_all + (TMA ? (threadIdx.x >> 5) * 2 * kBoxFloats : 0);
  unsigned long long* const dbar = dbar_all + (TMA ? (threadIdx.x >> 5) * 2 : 0);
  if constexpr (TMA) {
    if (lane == 0) {
      tma_mbar_init(&dbar[0], 1), tma_mbar_init(&dbar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }

  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  const bool alive = ray >= 0;
  RayCtx s;
  s.i_lo = 1, s.i_hi = 0;
  {
    float Y[K];
#pragma unroll
    for (int k = 0; k < K; ++k) Y[k] = 0.f;
    if (alive) {
      float vx, vy, vz;
      setup_ray(g, rp, c, ray, s, vx, vy, vz);
      sh_basis<DEG>(vx, vy, vz, Y);
    }
    float* Yrow = sm.Y + lane * H::YROW;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float q4[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) q4[l] = (4 * j + l < F) ? Y[(4 * j + l) % K] : 0.0f;
      *reinterpret_cast<float4*>(Yrow + 4 * j) = make_float4(q4[0], q4[1], q4[2], q4[3]);
    }
  }
  const Ray& r = s.r;
  bool marching = alive && s.i_lo <= s.i_hi;
  int lo = marching ? s.i_lo : 0x7fffffff, hi = marching ? s.i_hi : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(FULL, lo, o));
    hi = max(hi, __shfl_xor_sync(FULL, hi, o));
  }
  __syncwarp();  // SH table visible to the warp

  // group role of this lane: float4 `cj` of the 8 corner records of sample `base + ms`
  const int ms = lane / LPR, cj = lane % LPR;
  const bool role_ok = cj < NV;
  const unsigned stride4 = (unsigned)g.stride >> 2;  // record stride in float4s (the layout is a multiple of 4 floats)
  const unsigned long long feat_lane = reinterpret_cast<unsigned long long>(g.feat) + 16ull * (unsigned)cj;
  // Where this lane's channel sums end up after the group reduction (slot 0/1/2 = r/g/b of sm.R, -1 = nowhere), and, at
  // degree 2 (K = 9: float4s 2 and 4 straddle a channel boundary), the lane it trades with in the second exchange.
  int out_slot = -1, trade_lane = lane;
  if constexpr (DEG == 1) out_slot = role_ok ? cj : -1;
  if constexpr (DEG == 3) out_slot = (role_ok && (cj & 3) == 0) ? (cj >> 2) : -1;
  if constexpr (DEG == 2) {
    out_slot = cj == 0 ? 0 : (cj == 3 ? 1 : (cj == 5 ? 2 : -1));
    const int x = (cj == 0 || cj == 2) ? 2 : ((cj == 3 || cj == 4) ? 7 : ((cj == 5 || cj == 6) ? 3 : 0));
    trade_lane = lane ^ x;
  }
  const bool split1 = DEG == 2 && cj == 2, split2 = DEG == 2 && cj == 4, odd = (cj & 1) != 0;
  // DUAL: element ch * K (coefficient k = 0 of channel ch) sits in float4 (ch * K) / 4, component (ch * K) % 4
  int dslot = -1, dcomp = 0;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch)
    if (DUAL && DEG > 0 && cj == (ch * K) / 4) dslot = ch, dcomp = (ch * K) % 4;
  float cdr = 0.f, cdg = 0.f, cdb = 0.f;

  float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, acc = 0.f;
  float z = 0.f;
  bool have_z = false;
  DepthMarch dm;
  dm.bm = dm.bc = 0.f;
  // TMA look-ahead state: this lane's next sample (inside test + cell), the warp's next brick (origin, validity), the slot
  // it lands in and the mbarrier phases of the two slots
  bool have_next = false, nx_inside = false, nx_box_ok = false;
  CellQ nx_cq;
  nx_cq.ix = nx_cq.iy = nx_cq.iz = 0;
  int nx_bx = 0, nx_by = 0, nx_bz = 0, cur = 0;
  unsigned phase = 0u;
  Cell la_cell;      // LA: the next sample's cell and its 8 corner densities (requested one step ahead)
  float la_d[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) la_d[k] = 0.f;
  la_cell.ox[0] = la_cell.ox[1] = la_cell.oy[0] = la_cell.oy[1] = la_cell.oz[0] = la_cell.oz[1] = 0;
  la_cell.wx[0] = la_cell.wx[1] = la_cell.wy[0] = la_cell.wy[1] = la_cell.wz[0] = la_cell.wz[1] = 0.f;
  for (int i = lo; i <= hi; ++i) {
    // ---- per-lane: position, inside test, cell, density ----
    bool contributes = false;
    float sigma = 0.f, zn = 0.f;
    bool last = false;
    Cell cell;
    const bool mine = marching && i >= s.i_lo && i <= s.i_hi;
    if constexpr (LA) {
      // register look-ahead: the 8 density loads of sample i + 1 are issued before the gather of sample i and consumed one
      // iteration later, so the probe's global round trip overlaps the gather instead of preceding it
      if (mine) {
        if (!have_z) dm.start(s.dg, i), z = dm.next(s.dg, i), have_z = true;
        last = (i == c.S - 1);
        zn = last ? 0.0f : dm.next(s.dg, i + 1);
        bool inside;
        if (have_next) {
          inside = nx_inside, cell = la_cell;
          if (inside) {
            float sacc = 0.0f;
#pragma unroll
            for (int ix = 0; ix < 2; ++ix)
#pragma unroll
              for (int iy = 0; iy < 2; ++iy) {
                const float wxy = cell.wx[ix] * cell.wy[iy];
                float v0 = la_d[4 * ix + 2 * iy], v1 = la_d[4 * ix + 2 * iy + 1];
                if (g.pre == R3D_PRE_ABS) v0 = fabsf(v0), v1 = fabsf(v1);
                sacc = fmaf(wxy * cell.wz[0], v0, sacc);
                sacc = fmaf(wxy * cell.wz[1], v1, sacc);
              }
            float dpost;
            sigma = density_post(g.post, sacc * (g.pre == R3D_PRE_ABS ? fabsf(g.dscale) : g.dscale), dpost);
            contributes = sigma != 0.0f;
          }
        } else {
          const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
          const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
          const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
          if (inside_aabb(g, px, py, pz)) {
            make_cell_inside(g, px, py, pz, cell);
            float dpost;
            sigma = density_post(g.post, density_pre_interp<true>(g, cell), dpost);
            contributes = sigma != 0.0f;
          }
        }
      }
      have_next = mine && !last && (i + 1 <= s.i_hi);
      nx_inside = false;
      if (have_next) {
        const float qx = __fadd_rn(r.ox, __fmul_rn(r.dx, zn));
        const float qy = __fadd_rn(r.oy, __fmul_rn(r.dy, zn));
        const float qz = __fadd_rn(r.oz, __fmul_rn(r.dz, zn));
        nx_inside = inside_aabb(g, qx, qy, qz);
        if (nx_inside) {
          make_cell_inside(g, qx, qy, qz, la_cell);
#pragma unroll
          for (int ix = 0; ix < 2; ++ix)
#pragma unroll
            for (int iy = 0; iy < 2; ++iy) {
              const unsigned col = (unsigned)(la_cell.ox[ix] + la_cell.oy[iy]);
              la_d[4 * ix + 2 * iy] = __ldg(g.dens + (col + (unsigned)la_cell.oz[0]));
              la_d[4 * ix + 2 * iy + 1] = __ldg(g.dens + (col + (unsigned)la_cell.oz[1]));
            }
        }
      }
    } else if constexpr (DQ) {
      // density quad volume (r3d_device.cuh): the 8 corner densities are two 16-byte loads, no clamps; the clamped offsets /
      // zeroed weights of the cell are formed only for samples that go on to gather records
      if (mine) {
        if (!have_z) dm.start(s.dg, i), z = dm.next(s.dg, i), have_z = true;
        last = (i == c.S - 1);
        zn = last ? 0.0f : dm.next(s.dg, i + 1);
        const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
        const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
        const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
        if (inside_aabb(g, px, py, pz)) {
          CellQ cq;
          make_cell_q(g, px, py, pz, cq);
          float dpost;
          sigma = density_post(g.post, density_pre_interp_q(g, cq), dpost);
          contributes = sigma != 0.0f;
          if (contributes) cell_from_q(g, cq, cell);
        }
      }
    } else if constexpr (!TMA) {
      if (mine) {
        if (!have_z) dm.start(s.dg, i), z = dm.next(s.dg, i), have_z = true;
        last = (i == c.S - 1);
        zn = last ? 0.0f : dm.next(s.dg, i + 1);
        const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
        const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
        const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
        if (inside_aabb(g, px, py, pz)) {
          make_cell_inside(g, px, py, pz, cell);
          float dpost;
          sigma = density_post(g.post, density_pre_interp<true>(g, cell), dpost);
          contributes = sigma != 0.0f;
        }
      }
    } else {
      // ---- this step's brick (requested one step ago) ----
      const bool box_now = nx_box_ok;
      if (box_now) {
        tma_mbar_wait(&dbar[cur], (phase >> cur) & 1u);
        phase ^= 1u << cur;
      }
      if (mine) {
        if (!have_z) dm.start(s.dg, i), z = dm.next(s.dg, i), have_z = true;
        last = (i == c.S - 1);
        zn = last ? 0.0f : dm.next(s.dg, i + 1);
        bool inside;
        CellQ cq;
        if (have_next) {  // position, inside test and cell of this sample were formed by the previous step's look-ahead
          inside = nx_inside, cq = nx_cq;
        } else {
          const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
          const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
          const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
          inside = inside_aabb(g, px, py, pz);
          if (inside) make_cell_q(g, px, py, pz, cq);
        }
        if (inside) {
          float pre;
          if (box_now && have_next) {
            // 8 corner densities of the cell from the warp's brick: one base, immediate offsets, zero-filled outside the grid
            const float* b0 = dbox + cur * kBoxFloats + ((cq.ix - nx_bx) * kTmaBoxY + (cq.iy - nx_by)) * kTmaBoxZ + (cq.iz - nx_bz);
            const float* b1 = b0 + kTmaBoxY * kTmaBoxZ;
            const bool ab = g.pre == R3D_PRE_ABS;
            float sacc = 0.0f;
            float wxy = cq.wx[0] * cq.wy[0];
            float v0 = b0[0], v1 = b0[1];
            if (ab) v0 = fabsf(v0), v1 = fabsf(v1);
            sacc = fmaf(wxy * cq.wz[0], v0, sacc), sacc = fmaf(wxy * cq.wz[1], v1, sacc);
            wxy = cq.wx[0] * cq.wy[1];
            v0 = b0[kTmaBoxZ], v1 = b0[kTmaBoxZ + 1];
            if (ab) v0 = fabsf(v0), v1 = fabsf(v1);
            sacc = fmaf(wxy * cq.wz[0], v0, sacc), sacc = fmaf(wxy * cq.wz[1], v1, sacc);
            wxy = cq.wx[1] * cq.wy[0];
            v0 = b1[0], v1 = b1[1];
            if (ab) v0 = fabsf(v0), v1 = fabsf(v1);
            sacc = fmaf(wxy * cq.wz[0], v0, sacc), sacc = fmaf(wxy * cq.wz[1], v1, sacc);
            wxy = cq.wx[1] * cq.wy[1];
            v0 = b1[kTmaBoxZ], v1 = b1[kTmaBoxZ + 1];
            if (ab) v0 = fabsf(v0), v1 = fabsf(v1);
            sacc = fmaf(wxy * cq.wz[0], v0, sacc), sacc = fmaf(wxy * cq.wz[1], v1, sacc);
            pre = sacc * (ab ? fabsf(g.dscale) : g.dscale);
            float dpost;
            sigma = density_post(g.post, pre, dpost);
            contributes = sigma != 0.0f;
            if (contributes) cell_from_q(g, cq, cell);
          } else {
            cell_from_q(g, cq, cell);
            float dpost;
            sigma = density_post(g.post, density_pre_interp<true>(g, cell), dpost);
            contributes = sigma != 0.0f;
          }
        }
      }
      // ---- look-ahead: where do the warp's rays land at step i + 1?  One brick request for all of them. ----
      have_next = mine && !last && (i + 1 <= s.i_hi);
      nx_inside = false;
      if (have_next) {
        const float qx = __fadd_rn(r.ox, __fmul_rn(r.dx, zn));
        const float qy = __fadd_rn(r.oy, __fmul_rn(r.dy, zn));
        const float qz = __fadd_rn(r.oz, __fmul_rn(r.dz, zn));
        nx_inside = inside_aabb(g, qx, qy, qz);
        if (nx_inside) make_cell_q(g, qx, qy, qz, nx_cq);
      }
      const bool want = have_next && nx_inside;
      const unsigned any = __ballot_sync(FULL, want);
      cur ^= 1;
      nx_box_ok = false;
      if (any != 0u) {
        const int big = 0x3fffffff;
        const int mnx = __reduce_min_sync(FULL, want ? nx_cq.ix : big), mxx = __reduce_max_sync(FULL, want ? nx_cq.ix : -big);
        const int mny = __reduce_min_sync(FULL, want ? nx_cq.iy : big), mxy = __reduce_max_sync(FULL, want ? nx_cq.iy : -big);
        const int mnz = __reduce_min_sync(FULL, want ? nx_cq.iz : big), mxz = __reduce_max_sync(FULL, want ? nx_cq.iz : -big);
        // a cell needs voxels i0 and i0 + 1 on every axis; the box origin along the contiguous axis must keep every box row
        // 16-byte aligned in global memory (TMA moves 16-byte granules): z origin = a multiple of 4 floats
        const int oz = mnz & ~3;
        if (mxx - mnx + 2 <= kTmaBoxX && mxy - mny + 2 <= kTmaBoxY && mxz - oz + 2 <= kTmaBoxZ) {
          nx_box_ok = true, nx_bx = mnx, nx_by = mny, nx_bz = oz;
          if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot's last generic-proxy reads (two steps ago) precede the copy
            tma_mbar_expect_tx(&dbar[cur], (unsigned)(kBoxFloats * sizeof(float)));
            tma_load_box3(dbox + cur * kBoxFloats, dmap, oz, mny, mnx, &dbar[cur]);
            if (fmap != nullptr) tma_prefetch_box4(fmap, 0, oz, mny, mnx);
          }
        }
      }
    }
    if constexpr (PF != 0) {
      if (mine && !last) {
        const float qx = fmaf(r.dx, zn, r.ox), qy = fmaf(r.dy, zn, r.oy), qz = fmaf(r.dz, zn, r.oz);
        if (inside_aabb(g, qx, qy, qz)) {
          const float gx = (fmaf(qx, g.ns[0], g.nb[0]) + 1.0f) * (0.5f * (float)g.W) - 0.5f;
          const float gy = (fmaf(qy, g.ns[1], g.nb[1]) + 1.0f) * (0.5f * (float)g.D) - 0.5f;
          const float gz = (fmaf(qz, g.ns[2], g.nb[2]) + 1.0f) * (0.5f * (float)g.H) - 0.5f;
          const int x0 = min(max((int)floorf(gx), 0), max(g.W - 2, 0)), y0 = min(max((int)floorf(gy), 0), max(g.D - 2, 0));
          const int z0 = min(max((int)floorf(gz), 0), max(g.H - 2, 0));
          const unsigned v00 = (unsigned)((x0 * g.D + y0) * g.H + z0);
          const unsigned dy = (unsigned)(g.D > 1 ? g.H : 0), dx = (unsigned)(g.W > 1 ? g.D * g.H : 0);
          const unsigned cols[4] = {v00, v00 + dy, v00 + dx, v00 + dx + dy};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const char* col = reinterpret_cast<const char*>(g.feat) + 16ull * ((unsigned long long)cols[q] * stride4);
            // records z0 and z0 + 1 of the column: 2 * stride floats, at most three 128-byte lines
            asm volatile("prefetch.global.L2 [%0];" ::"l"(col));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(col + 4 * g.stride));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(col + 8 * g.stride - 4));
            if constexpr (PF == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(g.dens + cols[q]));  // z0 and z0 + 1 share a sector almost always
          }
        }
      }
    }
    const unsigned act = __ballot_sync(FULL, contributes);
    if (out.mask && lane == 0) out.mask[(size_t)i * (gridDim.x * 4u) + (blockIdx.x * 4u + (threadIdx.x >> 5))] = act;
    if (act != 0u) {
      const int total = __popc(act);
      int rank = __popc(act & ((1u << lane) - 1u));
      if constexpr (SORT) {
        // slot = (samples of cells whose first lane precedes this cell's first lane) + (position among the cell's lanes);
        // the low and the high voxel of the cell (corners 0 and 7) identify it
        unsigned peers = 0u;
        if (contributes)
          peers = __match_any_sync(act, (unsigned long long)(unsigned)(cell.ox[0] + cell.oy[0] + cell.oz[0]) |
                                            ((unsigned long long)(unsigned)(cell.ox[1] + cell.oy[1] + cell.oz[1]) << 32));
        const int leader = contributes ? (__ffs(peers) - 1) : lane;
        int scan = (contributes && leader == lane) ? __popc(peers) : 0;  // group size at the group's first lane
        const int own = scan;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(FULL, scan, o);
          if (lane >= o) scan += up;
        }
        rank = __shfl_sync(FULL, scan - own, leader) + __popc(peers & ((1u << lane) - 1u));
      }
      // ---- publish the sample: 8 corner weights, 8 corner record indices (in float4s; the launcher checked that
      //      they fit 32 bits), owning lane ----
      if (contributes) {
        float wc[8];
        unsigned rec4[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
          wc[k] = cell.wx[ix] * cell.wy[iy] * cell.wz[iz];
          rec4[k] = (unsigned)(cell.ox[ix] + cell.oy[iy] + cell.oz[iz]) * stride4;
        }
        float* Wrow = sm.W + rank * 8;
        *reinterpret_cast<float4*>(Wrow) = make_float4(wc[0], wc[1], wc[2], wc[3]);
        *reinterpret_cast<float4*>(Wrow + 4) = make_float4(wc[4], wc[5], wc[6], wc[7]);
        unsigned* Vrow = sm.V + rank * 8;
        *reinterpret_cast<uint4*>(Vrow) = make_uint4(rec4[0], rec4[1], rec4[2], rec4[3]);
        *reinterpret_cast<uint4*>(Vrow + 4) = make_uint4(rec4[4], rec4[5], rec4[6], rec4[7]);
        sm.src[rank] = lane;
      }
      __syncwarp();
      // ---- lane groups: MPI samples per iteration.  issue() puts the 8 record loads of a sample in flight, finish()
      //      applies weights and SH basis and reduces over the group; the loads of iteration n + 1 are issued before
      //      iteration n is finished (two register buffers), so only the first load latency of a step is exposed.
      auto issue = [&](int m, float4(&q)[8]) {
        if (m < total && role_ok) {
          const uint4 v0 = *reinterpret_cast<const uint4*>(sm.V + m * 8);
          const uint4 v1 = *reinterpret_cast<const uint4*>(sm.V + m * 8 + 4);
          const unsigned vk[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // address = lane base + 16 * record index
            unsigned long long addr;
            asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(addr) : "r"(vk[k]), "l"(feat_lane));
            q[k] = __ldg(reinterpret_cast<const float4*>(addr));
          }
        }
      };
      auto finish = [&](int m, const float4(&q)[8]) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < total && role_ok) {
          const float4 w0 = *reinterpret_cast<const float4*>(sm.W + m * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(sm.W + m * 8 + 4);
          const float4 y4 = *reinterpret_cast<const float4*>(sm.Y + sm.src[m] * H::YROW + 4 * cj);
          const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#if R3D_FWD_FFMA2
          {  // packed fp32 FMA: two record elements per issue slot (same products, same order per element)
            float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              a01 = ffma2(make_float2(q[k].x, q[k].y), wk[k], a01);
              a23 = ffma2(make_float2(q[k].z, q[k].w), wk[k], a23);
            }
            a = make_float4(a01.x, a01.y, a23.x, a23.y);
          }
#else
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a.x = fmaf(wk[k], q[k].x, a.x), a.y = fmaf(wk[k], q[k].y, a.y);
            a.z = fmaf(wk[k], q[k].z, a.z), a.w = fmaf(wk[k], q[k].w, a.w);
          }
#endif
          if constexpr (DUAL && DEG > 0) {  // band-0 radiance: C0 * interpolated coeff[ch][0] (Y[0] = C0 for every ray)
            if (dslot >= 0) sm.R2[m * 4 + dslot] = 0.28209479177387814f * (dcomp == 0 ? a.x : (dcomp == 1 ? a.y : (dcomp == 2 ? a.z : a.w)));
          }
          a.x *= y4.x, a.y *= y4.y, a.z *= y4.z, a.w *= y4.w;  // the pad element has Y = 0
        }
        // ---- channel sums of the group.  The record is channel-major (coeff[ch][k] = rec[ch * K + k]). ----
        if constexpr (DEG == 0) {  // one lane per sample, elements 0..2 are the three channels
          if (m < total) *reinterpret_cast<float4*>(sm.R + m * 4) = a;
        } else {
          const float u01 = a.x + a.y, u23 = a.z + a.w;
          float v;
          if constexpr (DEG == 2) {
            // float4 2 = {r8 | g0 g1 g2}, float4 4 = {g7 g8 | b0 b1}; every other float4 lies inside one channel.
            // A = part of the lane's first channel, B = part of its second one (0 for unsplit lanes).
            const float A = split1 ? a.x : (split2 ? u01 : u01 + u23);
            const float B = split1 ? a.y + u23 : (split2 ? u23 : 0.0f);
            // exchange 1 (xor 1): even lanes take the neighbour's A, odd lanes the neighbour's B:
            //   lane 0: r(0,1)   lane 3: g = B2 + A3   lane 5: b = B4 + A5;  lanes 2, 4, 6 keep their own A
            const float x = __shfl_xor_sync(FULL, odd ? A : B, 1);
            v = (split1 || split2) ? A : A + x;
            // exchange 2: r = lane 0 + lane 2, g = lane 3 + lane 4, b = lane 5 + lane 6
            v += __shfl_sync(FULL, v, trade_lane);
          } else {
            v = u01 + u23;
            if constexpr (DEG == 3) {  // 4 lanes per channel
              v += __shfl_xor_sync(FULL, v, 1);
              v += __shfl_xor_sync(FULL, v, 2);
            }
          }
          if (m < total && out_slot >= 0) sm.R[m * 4 + out_slot] = v;
        }
      };
#if R3D_FWD_PIPE
      {
        float4 qa[8], qb[8];
        int base = 0;
        issue(ms, qa);
        while (true) {  // `base`, `total` are warp-uniform: every lane takes the same path
          if (base + MPI < total) issue(base + MPI + ms, qb);
          finish(base + ms, qa);
          base += MPI;
          if (base >= total) break;
          if (base + MPI < total) issue(base + MPI + ms, qa);
          finish(base + ms, qb);
          base += MPI;
          if (base >= total) break;
        }
      }
#else
      for (int base = 0; base < total; base += MPI) {
        float4 q[8];
        issue(base + ms, q);
        finish(base + ms, q);
      }
#endif
      __syncwarp();
      if (contributes) {
        const float4 raw = *reinterpret_cast<const float4*>(sm.R + rank * 4);
        const float delta = last ? __fmul_rn(kInfinity, s.dnorm) : __fmul_rn(__fsub_rn(zn, z), s.dnorm);
        const float alpha = 1.0f - exp_neg(sigma * delta);
        const float w = alpha * T;
        const float sr = sigmoidf_(raw.x), sg = sigmoidf_(raw.y), sb2 = sigmoidf_(raw.z);
        if (out.cache) out.cache[(size_t)i * rp.n + ray] = make_float4(sr, sg, sb2, sigma);
        cr = fmaf(w, sr, cr);
        cg = fmaf(w, sg, cg);
        cb = fmaf(w, sb2, cb);
        if constexpr (DUAL) {
          // degree 0 has no higher bands: the diffuse radiance is the specular one
          const float4 raw2 = DEG > 0 ? *reinterpret_cast<const float4*>(sm.R2 + rank * 4) : raw;
          const float dr = sigmoidf_(raw2.x), dg_ = sigmoidf_(raw2.y), db = sigmoidf_(raw2.z);
          if (out.cache_diffuse) out.cache_diffuse[(size_t)i * rp.n + ray] = make_float4(dr, dg_, db, 0.0f);
          cdr = fmaf(w, dr, cdr);
          cdg = fmaf(w, dg_, cdg);
          cdb = fmaf(w, db, cdb);
        }
        dep = fmaf(w, z, dep);
        acc += w;
        T *= (1.0f - alpha);
        if (T == 0.0f) marching = false;  // every later weight is alpha*0 = 0 exactly
      }
    }
    if (mine) z = zn;
  }
  if (!alive) return;
  if (c.flags & R3D_FLAG_WHITE_BKGD) {
    const float bg = 1.0f - acc;
    cr += bg, cg += bg, cb += bg;
    cdr += bg, cdg += bg, cdb += bg;
  }
  out.colour[3 * ray] = cr, out.colour[3 * ray + 1] = cg, out.colour[3 * ray + 2] = cb;
  if constexpr (DUAL) out.colour_diffuse[3 * ray] = cdr, out.colour_diffuse[3 * ray + 1] = cdg, out.colour_diffuse[3 * ray + 2] = cdb;
  out.depth[ray] = dep;
  out.acc[ray] = acc;
  if (out.disparity) {
    const float ratio = __fdiv_rn(dep, acc);
    const float m = (ratio != ratio) ? ratio : fmaxf(kZeroPlus, ratio);
    out.disparity[ray] = __fdiv_rn(1.0f, m);
  }
}

template <int DEG, bool DUAL, bool SORT = false, int PF = 0>
__global__ void __launch_bounds__(128, DEG >= 3 ? 3 : R3D_FWD_BLOCKS) render_fwd_group_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out) {
  fwd_group_body<DEG, DUAL, SORT, PF, false>(g, rp, c, out, nullptr, nullptr);
}

#ifdef R3D_AB_VARIANTS
// the same kernel probing the density quad volume
template <int DEG, bool DUAL>
__global__ void __launch_bounds__(128, DEG >= 3 ? 3 : R3D_FWD_BLOCKS) render_fwd_group_dq_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out) {
  fwd_group_body<DEG, DUAL, false, 0, false, false, true>(g, rp, c, out, nullptr, nullptr);
}
// the same kernel with the density loads of sample i + 1 issued one step ahead (register look-ahead)
template <int DEG, bool DUAL>
__global__ void __launch_bounds__(128, DEG >= 3 ? 3 : R3D_FWD_BLOCKS) render_fwd_group_la_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out) {
  fwd_group_body<DEG, DUAL, false, 0, false, true>(g, rp, c, out, nullptr, nullptr);
}
// the same kernel with TMA density bricks (+ TMA L2 prefetch of the feature bricks when `feat_prefetch` is set): measured
// slower than the direct loads, kept in the measurement build
template <int DEG, bool DUAL>
__global__ void __launch_bounds__(128, DEG >= 3 ? 3 : R3D_FWD_BLOCKS)
    render_fwd_group_tma_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out, const __grid_constant__ CUtensorMap dens_map,
                                const __grid_constant__ CUtensorMap feat_map, const int feat_prefetch) {
  fwd_group_body<DEG, DUAL, false, 0, true>(g, rp, c, out, &dens_map, feat_prefetch ? &feat_map : nullptr);
}
#endif

#ifdef R3D_AB_VARIANTS
}  // namespace r3d
#include "r3d_fwd_ws.cuh"
#include "r3d_fwd_split.cuh"
namespace r3d {
#endif

// Per-ray upstream gradients folded into the three numbers the march needs: g_c (3), g_d, g_a, plus
// Total = sum_i w_i q_i rebuilt from the forward outputs.  Returns false when the ray receives no gradient.
struct RayGrad {
  float gc[3], gd, ga, total;
  float gcd[3];  // upstream gradient of the band-0 ("diffuse") image of a single-pass specular + diffuse render
};
__device__ __forceinline__ bool load_ray_grad(const BwdP& b, const CfgP& c, long long ray, RayGrad& rg) {
  float* gc = rg.gc;
  gc[0] = gc[1] = gc[2] = 0.f;
  float* gcd = rg.gcd;
  gcd[0] = gcd[1] = gcd[2] = 0.f;
  if (b.g_colour_diffuse) gcd[0] = __ldg(b.g_colour_diffuse + 3 * ray), gcd[1] = __ldg(b.g_colour_diffuse + 3 * ray + 1), gcd[2] = __ldg(b.g_colour_diffuse + 3 * ray + 2);
  float gd = 0.f, ga = 0.f;
  if (b.g_colour) gc[0] = __ldg(b.g_colour + 3 * ray), gc[1] = __ldg(b.g_colour + 3 * ray + 1), gc[2] = __ldg(b.g_colour + 3 * ray + 2);
  if (b.g_depth) gd = __ldg(b.g_depth + ray);
  if (b.g_acc) ga = __ldg(b.g_acc + ray);
  const float acc_f = __ldg(b.acc + ray), dep_f = __ldg(b.depth + ray);
  if (b.g_disp) {
    // disparity = 1 / max(eps, depth/acc) (accumulate.py:85-88): acc/depth when depth/acc > eps
    const float gdisp = __ldg(b.g_disp + ray);
    if (gdisp != 0.0f) {
      const float ratio = __fdiv_rn(dep_f, acc_f);
      if (ratio != ratio) {
        gd = ga = ratio;  // the reference back-propagates NaN through 0/0
      } else if (ratio > kZeroPlus) {
        const float disp = __fdiv_rn(1.0f, ratio);
        gd = fmaf(gdisp, -disp * disp / acc_f, gd);
        ga = fmaf(gdisp, disp / acc_f, ga);
      }
    }
  }
  float cfr = __ldg(b.colour + 3 * ray), cfg_ = __ldg(b.colour + 3 * ray + 1), cfb = __ldg(b.colour + 3 * ray + 2);
  float dfr = 0.f, dfg = 0.f, dfb = 0.f;
  if (b.g_colour_diffuse) dfr = __ldg(b.colour_diffuse + 3 * ray), dfg = __ldg(b.colour_diffuse + 3 * ray + 1), dfb = __ldg(b.colour_diffuse + 3 * ray + 2);
  if (c.flags & R3D_FLAG_WHITE_BKGD) {
    const float bg = 1.0f - acc_f;
    cfr -= bg, cfg_ -= bg, cfb -= bg;
    ga -= (gc[0] + gc[1] + gc[2]);  // d(1 - acc)/d acc on every channel
    if (b.g_colour_diffuse) {
      dfr -= bg, dfg -= bg, dfb -= bg;
      ga -= (gcd[0] + gcd[1] + gcd[2]);
    }
  }
  rg.gd = gd, rg.ga = ga;
  rg.total = fmaf(gc[0], cfr, fmaf(gc[1], cfg_, fmaf(gc[2], cfb, fmaf(gd, dep_f, ga * acc_f))));
  rg.total = fmaf(gcd[0], dfr, fmaf(gcd[1], dfg, fmaf(gcd[2], dfb, rg.total)));
  return !(gc[0] == 0.f && gc[1] == 0.f && gc[2] == 0.f && gd == 0.f && ga == 0.f && gcd[0] == 0.f && gcd[1] == 0.f && gcd[2] == 0.f);
}

#ifdef R3D_AB_VARIANTS
}  // namespace r3d
#include "r3d_bwd_ws.cuh"
namespace r3d {
#endif

#ifdef R3D_AB_VARIANTS  // thread-per-ray backward: measurement only
// =================================================================================================
// backward
// =================================================================================================
template <int DEG, int VEC>
__global__ void __launch_bounds__(128) render_bwd_kernel(const GridP g, const RaysP rp, const CfgP c, const BwdP b) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  if (ray < 0) return;

  RayGrad rgd;
  if (!load_ray_grad(b, c, ray, rgd)) return;
  const float gc[3] = {rgd.gc[0], rgd.gc[1], rgd.gc[2]};
  const float gd = rgd.gd, ga = rgd.ga, total = rgd.total;

  RayCtx s;
  float vx, vy, vz;
  setup_ray(g, rp, c, ray, s, vx, vy, vz);
  constexpr int K = (DEG + 1) * (DEG + 1);
  float Y[K];
  sh_basis<DEG>(vx, vy, vz, Y);
  const bool diffuse = (c.flags & R3D_FLAG_DIFFUSE) != 0;
  const Ray& r = s.r;
  const float dmul = (g.pre == R3D_PRE_ABS) ? fabsf(g.dscale) : g.dscale;

  float T = 1.0f, prefix = 0.f;
  if (s.i_lo > s.i_hi) return;
  // |q_i| <= |g_c|_1 + |g_d| z_max + |g_a|  (z is monotone and non-negative along the marched range)
  const float qmax = fabsf(gc[0]) + fabsf(gc[1]) + fabsf(gc[2]) + fabsf(gd) * fmaxf(fabsf(s.dg.near), fabsf(s.dg.far)) + fabsf(ga);
  DepthMarch dm;
  dm.start(s.dg, s.i_lo);
  float z = dm.next(s.dg, s.i_lo);
  for (int i = s.i_lo; i <= s.i_hi; ++i) {
    const bool last = (i == c.S - 1);
    const float zn = last ? 0.0f : dm.next(s.dg, i + 1);
    const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
    const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
    const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
    if (inside_aabb(g, px, py, pz)) {
      Cell cell;
      make_cell_inside(g, px, py, pz, cell);
      float dpost;
      const float sigma = density_post(g.post, density_pre_interp(g, cell), dpost);
      if (sigma != 0.0f || dpost != 0.0f) {
        const float delta = last ? __fmul_rn(kInfinity, s.dnorm) : __fmul_rn(__fsub_rn(zn, z), s.dnorm);
        const float alpha = 1.0f - exp_neg(sigma * delta);
        const float w = alpha * T;
        const float Tn = T * (1.0f - alpha);
        float sr, sg, sb;
        if (b.cache && sigma != 0.0f) {
          const float4 cv = __ldg(b.cache + (size_t)i * rp.n + ray);
          sr = cv.x, sg = cv.y, sb = cv.z;
        } else {
          float rr, rg, rb;
          gather_radiance<DEG, VEC>(g, cell, Y, diffuse, rr, rg, rb);
          sr = sigmoidf_(rr), sg = sigmoidf_(rg), sb = sigmoidf_(rb);
        }
        const float q = fmaf(gc[0], sr, fmaf(gc[1], sg, fmaf(gc[2], sb, fmaf(gd, z, ga))));
        prefix = fmaf(w, q, prefix);
        // sum_{j>i} w_j q_j = total - prefix.  It is empty for the last sample and once the transmittance is exactly 0.
        // The subtraction carries the rounding residue of two O(|total|) numbers (~6e-8 |total|); the true value is
        // bounded by T_{i+1} * max|q| (the later weights sum to at most T_{i+1}, sigmoid <= 1), so it is clamped to
        // that bound: deep inside opaque matter the residue would otherwise dwarf the (vanishing) true gradient.
        float suffix = 0.0f;
        if (!last && Tn != 0.0f) {
          const float bound = Tn * qmax;
          suffix = fminf(fmaxf(total - prefix, -bound), bound);
        }
        const float dsigma = delta * (Tn * q - suffix);
        const float dpre = dsigma * dpost * dmul;
        const float draw[3] = {w * gc[0] * sr * (1.0f - sr), w * gc[1] * sg * (1.0f - sg), w * gc[2] * sb * (1.0f - sb)};
        const bool feat_grad = b.gfeat && (draw[0] != 0.f || draw[1] != 0.f || draw[2] != 0.f);
#pragma unroll
        for (int ix = 0; ix < 2; ++ix)
#pragma unroll
          for (int iy = 0; iy < 2; ++iy) {
            const float wxy = cell.wx[ix] * cell.wy[iy];
            const size_t col = (size_t)(cell.ox[ix] + cell.oy[iy]);
#pragma unroll
            for (int iz = 0; iz < 2; ++iz) {
              const float wc = wxy * cell.wz[iz];
              if (wc == 0.0f) continue;  // out-of-range (zero padding) or exactly-on-plane corner
              const size_t vox = col + cell.oz[iz];
              if (b.gdens && dpre != 0.0f) {
                float gv = wc * dpre;
                if (g.pre == R3D_PRE_ABS) {
                  const float v = __ldg(g.dens + vox);
                  gv = (v > 0.f) ? gv : ((v < 0.f) ? -gv : 0.0f);  // d|x|/dx = sign(x), 0 at 0 (torch.abs)
                }
                atomicAdd(b.gdens + vox, gv);
              }
              if (feat_grad) corner_scatter<DEG, VEC>(b.gfeat + vox * (size_t)g.stride, Y, diffuse, wc, draw);
            }
          }
        T = Tn;
        if (T == 0.0f) break;
      }
    }
    z = zn;
  }
}


#endif  // R3D_AB_VARIANTS

// =================================================================================================
// backward, warp-cooperative scatter (default)
//
// Profile of the thread-per-ray scatter above (profiles/r01_v0_ncu_full_summary.md): the L1 data pipe issues one
// wavefront per distinct 128-byte line a request touches, and a 32-lane REDG.128 touches ~16 of them (each lane its
// own voxel record); 56 such requests per marching step make the kernel L1/L2-reduction bound.  Here the scatter is
// transposed through shared memory instead.  Per marching step of a warp (32 rays of an 8x4 pixel tile):
//   1. every lane does the per-ray maths of its sample (position, cell, density, alpha, T, q, dL/dsigma, dL/draw); the
//      per-sample sigmoid(raw)/sigma come from the forward's sample cache when there is one (no second gather),
//   2. lanes publish their 8 corner weights, corner voxel offsets, dL/dsigma_pre and the product table
//      P[e] = dL/draw[ch(e)] * Y[k(e)] (the SH part of the chain rule) in shared memory,
//   3. lanes whose samples fall in the same interpolation cell are grouped (__match_any_sync); for every distinct
//      cell the warp forms sum_members w[m][corner] * P[m][e] with LPR lanes per voxel record (one float4 each) and
//      issues ONE 128-bit reduction per lane: a record is one coalesced, fully used line instead of 32 scattered
//      lanes, and samples sharing a cell are summed before they reach L2.
// =================================================================================================
template <int DEG>
struct alignas(16) CoopSmem {
  using S = CoopShape<DEG>;
  float P[32 * S::PROW + 64];  // +64: lanes whose float4 index is past the record still read in bounds
  float W[32 * S::WROW2];
  int V[32 * S::WROW];
};

// DUAL: backward of the single-pass specular + diffuse render: the band-0 image adds g_cd . sigmoid(raw_d_i) to q_i and
// d raw_d[ch] * Y[0] to element ch * K of the product row; nothing else changes (same samples, same weights).
// MASK: the forward left its per-step contribution ballots (BwdP::mask) and per-sample records (BwdP::cache, which hold
// sigma) and the density post-activation is ReLU (d sigma / d pre = 1 wherever sigma != 0): a sample contributes to the
// gradient exactly if it contributed to the image, so the march takes sigma from the cache and needs neither the inside test
// nor the 8-corner density gather, and steps in which no ray of the warp contributed are skipped outright.
template <int DEG, int VEC, bool DUAL, bool MASK>
__global__ void __launch_bounds__(128, DEG >= 3 ? 3 : (DUAL ? R3D_BWD_DUAL_BLOCKS : R3D_BWD_BLOCKS)) render_bwd_coop_kernel(const GridP g, const RaysP rp, const CfgP c, const BwdP b) {
  using S = CoopShape<DEG>;
  constexpr int K = S::K, F = S::F;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) CoopSmem<DEG> smem_all[4];
  CoopSmem<DEG>& sm = smem_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;

  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);

  // ---- per-ray state; lanes without work stay in the loop (the scatter is a warp-wide collective) ----
  RayGrad rgd;
  bool alive = (ray >= 0) && load_ray_grad(b, c, ray, rgd);
  RayCtx s;
  float Y[K];
  float qmax = 0.f;
  s.i_lo = 1, s.i_hi = 0;
  if (alive) {
    float vx, vy, vz;
    setup_ray(g, rp, c, ray, s, vx, vy, vz);
    sh_basis<DEG>(vx, vy, vz, Y);
    qmax = fabsf(rgd.gc[0]) + fabsf(rgd.gc[1]) + fabsf(rgd.gc[2]) + fabsf(rgd.gd) * fmaxf(fabsf(s.dg.near), fabsf(s.dg.far)) + fabsf(rgd.ga);
    if constexpr (DUAL) qmax += fabsf(rgd.gcd[0]) + fabsf(rgd.gcd[1]) + fabsf(rgd.gcd[2]);
    alive = s.i_lo <= s.i_hi;
  }
  const bool diffuse = (c.flags & R3D_FLAG_DIFFUSE) != 0;
  const float dmul = (g.pre == R3D_PRE_ABS) ? fabsf(g.dscale) : g.dscale;
  const Ray& r = s.r;

  // warp-uniform loop bounds
  int lo = alive ? s.i_lo : 0x7fffffff, hi = alive ? s.i_hi : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(FULL, lo, o));
    hi = max(hi, __shfl_xor_sync(FULL, hi, o));
  }

  // cooperative-phase role of this lane: float4 `cj` of corner `pass * CPP + cq`
  const int cq = lane / S::LPR, cj = lane % S::LPR;
  const bool role_ok = (cq < S::CPP) && (cj < S::NV);
  // in a band-0-only (diffuse) render only the float4s holding a k = 0 coefficient (elements 0, K, 2K) carry gradient
  const bool band_ok = !(DEG > 0 && diffuse) || (cj == 0) || (cj == K / 4) || (cj == (2 * K) / 4);
  const unsigned ustride = (unsigned)g.stride;
  // per-lane bases into the published tables (the member sweep adds m * row stride)
  const unsigned w_top = (unsigned)__cvta_generic_to_shared(sm.W + 31 * S::WROW2 + S::PASSES * (cq % S::CPP));
  const unsigned wd_top = (unsigned)__cvta_generic_to_shared(sm.W + 31 * S::WROW2 + 8 + (lane & 7));
  const unsigned p_top = (unsigned)__cvta_generic_to_shared(sm.P + 31 * S::PROW + 4 * cj);
  const unsigned mask_stride = gridDim.x * 4u, mask_warp = blockIdx.x * 4u + (threadIdx.x >> 5);
  unsigned fmask_next = 0u;
  float4 cv_next = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (MASK) {
    if (lo <= hi) {
      fmask_next = __ldg(b.mask + (size_t)lo * mask_stride + mask_warp);
#if R3D_BWD_PREFETCH
      if (alive && ((fmask_next >> lane) & 1u)) cv_next = __ldg(b.cache + (size_t)lo * rp.n + ray);
#endif
    }
  }

  float T = 1.0f, prefix = 0.f;
  float z = 0.f;
  bool have_z = false;
  DepthMarch dm;
  dm.bm = dm.bc = 0.f;
  for (int i = lo; i <= hi; ++i) {
    // ------------------------------------------------------------------ 1. per-lane sample maths
    bool contributes = false;
    float wc[8];
    int vox[8];
    float draw[3] = {0.f, 0.f, 0.f}, dpre = 0.f;
    float draw0[3] = {0.f, 0.f, 0.f};  // DUAL: d L / d raw_diffuse, lands on the k = 0 coefficients only
    unsigned long long key = 0ull;  // 64-bit: voxel offset (up to 2^31) + the low corner's validity pattern
    unsigned fmask = 0u;
    float4 cv_pre = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (MASK) {
      fmask = fmask_next;
      cv_pre = cv_next;
      if (i < hi) {
        fmask_next = __ldg(b.mask + (size_t)(i + 1) * mask_stride + mask_warp);
#if R3D_BWD_PREFETCH
        // request the next step's per-sample record now: its HBM latency hides behind this step's sweep
        if (alive && ((fmask_next >> lane) & 1u)) cv_next = __ldg(b.cache + (size_t)(i + 1) * rp.n + ray);
#endif
      }
      if (fmask == 0u) {  // no ray of this warp contributed at this step: nothing to do, not even the depths
        have_z = false;   // (the stratum marcher restarts at the next contributing step)
        continue;
      }
    }
    if (alive && i >= s.i_lo && i <= s.i_hi) {
      if (!have_z) dm.start(s.dg, i), z = dm.next(s.dg, i), have_z = true;
      const bool last = (i == c.S - 1);
      const float zn = last ? 0.0f : dm.next(s.dg, i + 1);
      const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
      const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
      const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
      if (MASK ? ((fmask >> lane) & 1u) != 0u : inside_aabb(g, px, py, pz)) {
        Cell cell;
        make_cell_inside(g, px, py, pz, cell);
        float dpost, sigma;
        float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (MASK) {
#if R3D_BWD_PREFETCH
          cv = cv_pre;
#else
          cv = __ldg(b.cache + (size_t)i * rp.n + ray);
#endif
          sigma = cv.w, dpost = 1.0f;  // ReLU with sigma != 0
        } else {
          sigma = density_post(g.post, density_pre_interp(g, cell), dpost);
        }
        if (sigma != 0.0f || dpost != 0.0f) {
          const float delta = last ? __fmul_rn(kInfinity, s.dnorm) : __fmul_rn(__fsub_rn(zn, z), s.dnorm);
          const float alpha = 1.0f - exp_neg(sigma * delta);
          const float w = alpha * T;
          const float Tn = T * (1.0f - alpha);
          float sr, sg, sb;
          if (MASK || (b.cache && sigma != 0.0f)) {  // written by the forward for exactly the sigma != 0 samples
            if constexpr (!MASK) cv = __ldg(b.cache + (size_t)i * rp.n + ray);
            sr = cv.x, sg = cv.y, sb = cv.z;
          } else {
            float rr, rg, rb;
            gather_radiance<DEG, VEC>(g, cell, Y, diffuse, rr, rg, rb);
            sr = sigmoidf_(rr), sg = sigmoidf_(rg), sb = sigmoidf_(rb);
          }
          float q = fmaf(rgd.gc[0], sr, fmaf(rgd.gc[1], sg, fmaf(rgd.gc[2], sb, fmaf(rgd.gd, z, rgd.ga))));
          float dr = 0.f, dg_ = 0.f, db = 0.f;  // band-0 radiance of the sample (DUAL)
          if constexpr (DUAL) {
            if (b.cache_diffuse && sigma != 0.0f) {
              const float4 cd = __ldg(b.cache_diffuse + (size_t)i * rp.n + ray);
              dr = cd.x, dg_ = cd.y, db = cd.z;
            } else {
              float rr, rg, rb;
              gather_radiance<DEG, VEC>(g, cell, Y, true, rr, rg, rb);
              dr = sigmoidf_(rr), dg_ = sigmoidf_(rg), db = sigmoidf_(rb);
            }
            q = fmaf(rgd.gcd[0], dr, fmaf(rgd.gcd[1], dg_, fmaf(rgd.gcd[2], db, q)));
          }
          prefix = fmaf(w, q, prefix);
          float suffix = 0.0f;  // see render_bwd_kernel for the clamp
          if (!last && Tn != 0.0f) {
            const float bound = Tn * qmax;
            suffix = fminf(fmaxf(rgd.total - prefix, -bound), bound);
          }
          dpre = delta * (Tn * q - suffix) * dpost * dmul;
          if (!b.gdens) dpre = 0.f;
          if (b.gfeat) {
            draw[0] = w * rgd.gc[0] * sr * (1.0f - sr);
            draw[1] = w * rgd.gc[1] * sg * (1.0f - sg);
            draw[2] = w * rgd.gc[2] * sb * (1.0f - sb);
            if constexpr (DUAL) {
              draw0[0] = w * rgd.gcd[0] * dr * (1.0f - dr);
              draw0[1] = w * rgd.gcd[1] * dg_ * (1.0f - dg_);
              draw0[2] = w * rgd.gcd[2] * db * (1.0f - db);
            }
          }
          contributes = (dpre != 0.f) || (draw[0] != 0.f) || (draw[1] != 0.f) || (draw[2] != 0.f);
          if constexpr (DUAL) contributes = contributes || (draw0[0] != 0.f) || (draw0[1] != 0.f) || (draw0[2] != 0.f);
          if (contributes) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
              wc[k] = cell.wx[ix] * cell.wy[iy] * cell.wz[iz];
              vox[k] = cell.ox[ix] + cell.oy[iy] + cell.oz[iz];
            }
            // the clamped offsets of the low corner identify the cell unless it is clamped at the border; fold the
            // validity pattern in so that border cells with different zero-padding never merge
            key = (unsigned long long)(unsigned)vox[0] | ((unsigned long long)((cell.wx[0] != 0.f) | ((cell.wy[0] != 0.f) << 1) | ((cell.wz[0] != 0.f) << 2)) << 32);
          }
          T = Tn;
          if (T == 0.0f) alive = false;  // every later weight is exactly 0
        }
      }
      z = zn;
    }
    const unsigned act = __ballot_sync(FULL, contributes);
    if (act == 0u) continue;

    // ------------------------------------------------------------------ 2. publish
    if (contributes) {
      float* Prow = sm.P + lane * S::PROW;
#pragma unroll
      for (int j = 0; j < S::NV; ++j) {
        float q4[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const int e = 4 * j + l;
          const bool on = (e < F) && (!(DEG > 0 && diffuse) || (e % K) == 0);
          q4[l] = on ? draw[e / K] * Y[e % K] : 0.0f;
          if (DUAL && e < F && (e % K) == 0) q4[l] = (draw[e / K] + draw0[e / K]) * Y[0];
        }
        *reinterpret_cast<float4*>(Prow + 4 * j) = make_float4(q4[0], q4[1], q4[2], q4[3]);
      }
      float wt[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) wt[S::wpos(k)] = wc[k];
      float* Wrow = sm.W + lane * S::WROW2;
      *reinterpret_cast<float4*>(Wrow) = make_float4(wt[0], wt[1], wt[2], wt[3]);
      *reinterpret_cast<float4*>(Wrow + 4) = make_float4(wt[4], wt[5], wt[6], wt[7]);
      *reinterpret_cast<float4*>(Wrow + 8) = make_float4(wc[0] * dpre, wc[1] * dpre, wc[2] * dpre, wc[3] * dpre);
      *reinterpret_cast<float4*>(Wrow + 12) = make_float4(wc[4] * dpre, wc[5] * dpre, wc[6] * dpre, wc[7] * dpre);
      int* Vrow = sm.V + lane * S::WROW;
      *reinterpret_cast<int4*>(Vrow) = make_int4(vox[0], vox[1], vox[2], vox[3]);
      *reinterpret_cast<int4*>(Vrow + 4) = make_int4(vox[4], vox[5], vox[6], vox[7]);
    }
    // ------------------------------------------------------------------ 3. group by cell, cooperative reduction
    __syncwarp();  // tables visible to the whole warp
    unsigned peers = 0u;
    if (contributes) peers = __match_any_sync(act, key);
    const bool leader = contributes && ((int)(__ffs(peers) - 1) == lane);
    unsigned leaders = __ballot_sync(FULL, leader);
    while (leaders) {
      const int L = __ffs(leaders) - 1;
      leaders &= leaders - 1;
      const unsigned members = __shfl_sync(FULL, peers, L);
      // One sweep over the cell's member samples accumulates, per lane, its float4 of up to PASSES corner records
      // (the P row is loaded once and reused for every pass) and the density gradient of corner (lane & 7).  Every lane
      // runs the sweep (lanes without a role read in-bounds garbage and never store): no divergence inside the loop.
      float2 a01[S::PASSES], a23[S::PASSES];  // packed pairs: one FFMA2 per two record elements (sm_100)
#pragma unroll
      for (int pass = 0; pass < S::PASSES; ++pass) a01[pass] = a23[pass] = make_float2(0.f, 0.f);
      float ad = 0.f;
      unsigned mm = members;
      while (mm) {
        // members are taken from the top (any order will do): z = leading zeros = 31 - lane, one FLO; every table
        // address is (per-lane base of row 31) - z * (row bytes), one IMAD each.  The mask is warp-uniform, so the bit
        // scan runs on the uniform datapath beside the vector pipes; the classic ffs / mm &= mm - 1 form has 4 fewer
        // instructions per member but keeps them on the vector ALU and measured slower (4.49 vs 4.36 ms at c3).
        const unsigned zc = (unsigned)__clz(mm);
        mm &= ~(0x80000000u >> zc);
        float wm[S::PASSES];
        if constexpr (S::PASSES == 1) {
          wm[0] = lds_f32(w_top - zc * (4u * S::WROW2));
        } else if constexpr (S::PASSES == 2) {
          lds_v2(w_top - zc * (4u * S::WROW2), wm[0], wm[1]);
        } else {
          lds_v4(w_top - zc * (4u * S::WROW2), wm[0], wm[1], wm[2], wm[3]);
        }
        ad += lds_f32(wd_top - zc * (4u * S::WROW2));
        float4 p4;
        lds_v4(p_top - zc * (4u * S::PROW), p4.x, p4.y, p4.z, p4.w);
#pragma unroll
        for (int pass = 0; pass < S::PASSES; ++pass) {
          a01[pass] = ffma2(make_float2(p4.x, p4.y), wm[pass], a01[pass]);
          a23[pass] = ffma2(make_float2(p4.z, p4.w), wm[pass], a23[pass]);
        }
      }
      const int* VL = sm.V + L * S::WROW;
      if (b.gfeat && role_ok && band_ok) {
#pragma unroll
        for (int pass = 0; pass < S::PASSES; ++pass) {
          const float4 v = make_float4(a01[pass].x, a01[pass].y, a23[pass].x, a23[pass].y);
          // 32x32 -> 64-bit unsigned multiply: record offsets exceed 2^31 floats at 512^3 / degree 3
          float* dst = b.gfeat + (size_t)(unsigned)VL[pass * S::CPP + cq] * (size_t)ustride + 4 * cj;
          if constexpr (VEC != 0) {
            red_add_v4(dst, v.x, v.y, v.z, v.w);
          } else {
            if (4 * cj + 0 < F) atomicAdd(dst + 0, v.x);
            if (4 * cj + 1 < F) atomicAdd(dst + 1, v.y);
            if (4 * cj + 2 < F) atomicAdd(dst + 2, v.z);
            if (4 * cj + 3 < F) atomicAdd(dst + 3, v.w);
          }
        }
      }
      if (b.gdens && lane < 8 && ad != 0.f) {
        const unsigned vx_ = (unsigned)VL[lane];
        if (g.pre == R3D_PRE_ABS) {
          const float v = __ldg(g.dens + vx_);
          ad = (v > 0.f) ? ad : ((v < 0.f) ? -ad : 0.0f);  // d|x|/dx = sign(x), 0 at 0 (torch.abs)
        }
        atomicAdd(b.gdens + vx_, ad);
      }
    }
    __syncwarp();  // the tables are rewritten at the next contributing step
  }
}

// =================================================================================================
// measurement helper: mark voxels referenced as interpolation corners by in-volume samples
// =================================================================================================
__global__ void __launch_bounds__(128) mark_touched_kernel(const GridP g, const RaysP rp, const CfgP c, uint8_t* __restrict__ bitmap) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  if (ray < 0) return;
  RayCtx s;
  float vx, vy, vz;
  setup_ray(g, rp, c, ray, s, vx, vy, vz);
  const Ray& r = s.r;
  for (int i = s.i_lo; i <= s.i_hi; ++i) {
    const float z = s.dg.at(i);
    const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
    const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
    const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
    if (!inside_aabb(g, px, py, pz)) continue;
    Cell cell;
    make_cell_inside(g, px, py, pz, cell);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
      if (cell.wx[ix] * cell.wy[iy] * cell.wz[iz] != 0.0f) bitmap[(size_t)(cell.ox[ix] + cell.oy[iy] + cell.oz[iz])] = 1;
    }
  }
}

// =================================================================================================
// measurement helper: sample statistics of a batch (SURVEY.md 8d companion figures)
//   counters[0] samples visited (conservative in-range march)   counters[1] samples strictly inside the AABB
//   counters[2] in-range trilinear corner references of those    counters[3] contributing samples (sigma != 0)
// =================================================================================================
__global__ void __launch_bounds__(128) sample_stats_kernel(const GridP g, const RaysP rp, const CfgP c, unsigned long long* __restrict__ counters) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  unsigned visited = 0u, inside = 0u, refs = 0u, contrib = 0u;
  if (ray >= 0) {
    RayCtx s;
    float vx, vy, vz;
    setup_ray(g, rp, c, ray, s, vx, vy, vz);
    const Ray& r = s.r;
    for (int i = s.i_lo; i <= s.i_hi; ++i) {
      const float z = s.dg.at(i);
      const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
      const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
      const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
      ++visited;
      if (!inside_aabb(g, px, py, pz)) continue;
      ++inside;
      Cell cell;
      make_cell_inside(g, px, py, pz, cell);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
        if (cell.wx[ix] * cell.wy[iy] * cell.wz[iz] != 0.0f) ++refs;
      }
      float dpost;
      if (density_post(g.post, density_pre_interp(g, cell), dpost) != 0.0f) ++contrib;
    }
  }
  unsigned v[4] = {visited, inside, refs, contrib};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    unsigned x = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(counters + q, (unsigned long long)x);
  }
}

// =================================================================================================
// host-side dispatch
// =================================================================================================
#ifdef R3D_AB_VARIANTS
// ---- tensor maps of the TMA forward (density volume [W][D][H] fp32, feature volume [W][D][H][stride] fp32) ----
struct TmaMaps {
  CUtensorMap dens, feat;
  bool ok;
};
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tma_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// The maps depend on the buffers' addresses and shapes only; they are cached (the one piece of state the library keeps,
// behind a mutex) and re-encoded when a grid of another shape or at another address shows up.
static const TmaMaps& tma_maps_for(const GridP& g) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, const void*, int, int, int, int>, TmaMaps> cache;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple((const void*)g.dens, (const void*)g.feat, g.W, g.D, g.H, g.stride);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  if (cache.size() > 64) cache.clear();
  TmaMaps m;
  m.ok = false;
  EncodeTiledFn enc = tma_encoder();
  // strides must be multiples of 16 bytes: H % 4 == 0 for the density volume; the record stride is one already
  if (enc && g.H % 4 == 0 && g.stride % 4 == 0 && (reinterpret_cast<uintptr_t>(g.dens) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.feat) & 15) == 0) {
    const cuuint64_t ddim[3] = {(cuuint64_t)g.H, (cuuint64_t)g.D, (cuuint64_t)g.W};
    const cuuint64_t dstr[2] = {(cuuint64_t)g.H * 4, (cuuint64_t)g.D * g.H * 4};
    const cuuint32_t dbox[3] = {(cuuint32_t)kTmaBoxZ, (cuuint32_t)kTmaBoxY, (cuuint32_t)kTmaBoxX};
    const cuuint32_t one3[3] = {1, 1, 1};
    const CUresult r1 = enc(&m.dens, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g.dens), ddim, dstr, dbox, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const cuuint64_t fdim[4] = {(cuuint64_t)g.stride, (cuuint64_t)g.H, (cuuint64_t)g.D, (cuuint64_t)g.W};
    const cuuint64_t fstr[3] = {(cuuint64_t)g.stride * 4, (cuuint64_t)g.H * g.stride * 4, (cuuint64_t)g.D * g.H * g.stride * 4};
    const cuuint32_t fbox[4] = {(cuuint32_t)g.stride, (cuuint32_t)kTmaBoxZ, (cuuint32_t)kTmaBoxY, (cuuint32_t)kTmaBoxX};
    const cuuint32_t one4[4] = {1, 1, 1, 1};
    const CUresult r2 = g.stride <= 256
                            ? enc(&m.feat, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(g.feat), fdim, fstr, fbox, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                            : CUDA_ERROR_INVALID_VALUE;
    m.ok = (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS);
  }
  return cache.emplace(key, m).first->second;
}

#endif  // R3D_AB_VARIANTS

// the lane-group forward addresses records with 32-bit float4 indices
static bool group_indexable(const GridP& g) {
  return (unsigned long long)g.W * g.D * g.H * (unsigned long long)(g.stride / 4) <= 0xffffffffull;
}
// does r3d_render_fwd dispatch the lane-group kernel (the one that writes OutP::mask) for these arguments?
static bool fwd_uses_group_kernel(const GridP& g, const CfgP& c, int vec, int variant, int sh_degree) {
  const bool diffuse = (c.flags & R3D_FLAG_DIFFUSE) != 0 && sh_degree > 0;
  return vec != 0 && !diffuse && !(variant & (2 | 4 | 8)) && group_indexable(g);  // (the ws / sorted variants write it too)
}
// the backward can march by the forward's contribution ballots when it also has the per-sample records (sigma) and the
// density post-activation is ReLU (see render_bwd_coop_kernel)
static bool mask_usable(const GridP& g, const BwdP& b, int vec) {
  return b.mask != nullptr && b.cache != nullptr && vec != 0 && g.post == R3D_POST_RELU;
}

// warp-specialised forward (r3d_fwd_ws.cuh): 4 producer + 4 consumer warps per CTA, dynamic shared memory
#ifdef R3D_AB_VARIANTS
template <int DEG, bool DUAL, bool SORT, int PF = 0, bool DQ = false>
static void launch_fwd_ws(dim3 grid, cudaStream_t st, const GridP& g, const RaysP& r, const CfgP& c, const OutP& o) {
  // the per-CTA stratum table needs one (near, far) for all rays and has to fit beside the stage rings
  const bool table = r.bounds == nullptr && !(c.flags & R3D_FLAG_OPTIMIZED_SAMPLING) && c.S <= 4096;
  const size_t smem = ws_smem_bytes<DEG, DUAL>(c.S, table);
  static const bool attr = [] {
    cudaFuncSetAttribute(render_fwd_ws_kernel<DEG, DUAL, SORT, PF, DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_smem_bytes<DEG, DUAL>(4096, true));
    return true;
  }();
  (void)attr;
  render_fwd_ws_kernel<DEG, DUAL, SORT, PF, DQ><<<grid, 256, smem, st>>>(g, r, c, o, table ? 1 : 0);
}

#endif

template <int DEG>
static void launch_fwd(int vec, int variant, dim3 grid, cudaStream_t st, const GridP& g, const RaysP& r, const CfgP& c, const OutP& o) {
  // the lane-group gather needs 16-byte aligned records addressable with 32-bit float4 indices; band-0-only (diffuse) renders
  // read 3 floats per record and keep the per-ray gather
  const bool diffuse = (c.flags & R3D_FLAG_DIFFUSE) != 0 && DEG > 0;
#ifdef R3D_AB_VARIANTS
  if (vec != 0 && !diffuse && !(variant & 2)) {
    if (variant & 4) {  // TMA (cp.async.bulk) staging instead of per-lane cp.async: measured, see DESIGN.md
      render_fwd_coop_kernel<DEG, true><<<grid, 128, 0, st>>>(g, r, c, o);
      return;
    }
    if (variant & 8) {  // shared-memory staged gather (the default before the lane-group kernel)
      render_fwd_coop_kernel<DEG, false><<<grid, 128, 0, st>>>(g, r, c, o);
      return;
    }
    if (group_indexable(g) && (variant & 32768) && o.mask && o.cache) {  // two-kernel forward (r3d_fwd_split.cuh)
      static float* wplane = nullptr;
      static size_t wplane_floats = 0;
      const size_t need = (size_t)c.S * (size_t)r.n;
      if (need > wplane_floats) {
        if (wplane) cudaFree(wplane);
        cudaMalloc(&wplane, need * sizeof(float));
        wplane_floats = need;
      }
      if (g.quads)
        render_fwd_probe_kernel<true><<<grid, 128, 0, st>>>(g, r, c, o, wplane);
      else
        render_fwd_probe_kernel<false><<<grid, 128, 0, st>>>(g, r, c, o, wplane);
      static const int mode = [] {
        const char* e = getenv("R3D_SPLIT_MODE");
        return e ? atoi(e) : 0;
      }();
      static const int carve = [] {  // tuning hook: preferred shared-memory carve-out of the gather kernel in percent
        const char* e = getenv("R3D_SPLIT_CARVEOUT");
        const int v = e ? atoi(e) : -1;
        if (v >= 0) {
          cudaFuncSetAttribute(render_fwd_gather_kernel<DEG, 0, false>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
          cudaFuncSetAttribute(render_fwd_gather_kernel<DEG, 1, false>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
        }
        return v;
      }();
      (void)carve;
      switch (mode) {  // $R3D_SPLIT_MODE: prefetch flavour (0..3) + 10 for the run-merged gather
        case 1: render_fwd_gather_kernel<DEG, 1, false><<<grid, 128, 0, st>>>(g, r, c, o, wplane); break;
        case 2: render_fwd_gather_kernel<DEG, 2, false><<<grid, 128, 0, st>>>(g, r, c, o, wplane); break;
        case 3: render_fwd_gather_kernel<DEG, 3, false><<<grid, 128, 0, st>>>(g, r, c, o, wplane); break;
        case 9: render_fwd_gather_kernel<DEG, 9, false><<<grid, 128, 0, st>>>(g, r, c, o, wplane); break;
        case 10: render_fwd_gather_kernel<DEG, 0, true><<<grid, 128, 0, st>>>(g, r, c, o, wplane); break;
        case 11: render_fwd_gather_kernel<DEG, 1, true><<<grid, 128, 0, st>>>(g, r, c, o, wplane); break;
        case 13: render_fwd_gather_kernel<DEG, 3, true><<<grid, 128, 0, st>>>(g, r, c, o, wplane); break;
        default: render_fwd_gather_kernel<DEG, 0, false><<<grid, 128, 0, st>>>(g, r, c, o, wplane); break;
      }
      return;
    }
    if (group_indexable(g) && (variant & 65536) && g.quads) {  // lane-group kernel probing the density quad volume
      render_fwd_group_dq_kernel<DEG, false><<<grid, 128, 0, st>>>(g, r, c, o);
      return;
    }
    if (group_indexable(g) && (variant & (32 | 2048 | 4096 | 8192 | 16384))) {
      if ((variant & 96) == 96 && (variant & 256) && (variant & 1024))
        launch_fwd_ws<DEG, false, true, 1, true>(grid, st, g, r, c, o);
      else if ((variant & 96) == 96 && (variant & 1024))
        launch_fwd_ws<DEG, false, true, 0, true>(grid, st, g, r, c, o);
      else if ((variant & 96) == 96 && (variant & 256))
        launch_fwd_ws<DEG, false, true, 1>(grid, st, g, r, c, o);
      else if ((variant & 96) == 96 && (variant & 512))
        launch_fwd_ws<DEG, false, true, 2>(grid, st, g, r, c, o);
      else if ((variant & 96) == 96)
        launch_fwd_ws<DEG, false, true>(grid, st, g, r, c, o);
      else if (variant & 32)
        launch_fwd_ws<DEG, false, false>(grid, st, g, r, c, o);
      else if (variant & 16384)
        render_fwd_group_la_kernel<DEG, false><<<grid, 128, 0, st>>>(g, r, c, o);
      else if (variant & 8192)
        render_fwd_group_kernel<DEG, false, false, 2><<<grid, 128, 0, st>>>(g, r, c, o);
      else if (variant & 4096)
        render_fwd_group_kernel<DEG, false, false, 1><<<grid, 128, 0, st>>>(g, r, c, o);
      else
        render_fwd_group_kernel<DEG, false, true><<<grid, 128, 0, st>>>(g, r, c, o);
      return;
    }
  }
  const bool per_ray = (variant & 2) != 0;
#else
  (void)variant;
  const bool per_ray = false;
#endif
  if (vec != 0 && !diffuse && !per_ray && group_indexable(g)) {
#ifdef R3D_AB_VARIANTS
    // TMA density bricks need a coherent warp (an image tile per warp): image-shaped batches and in-kernel ray generation.
    // $R3D_FWD_TMA: 0 = off (default: measured slower), 1 = density bricks, 2 = + feature-brick L2 prefetch
    static const int tma_mode = [] {
      const char* e = getenv("R3D_FWD_TMA");
      return e ? atoi(e) : 0;
    }();
    if (tma_mode > 0 && r.tile_w > 0) {
      const TmaMaps& m = tma_maps_for(g);
      if (m.ok) {
        render_fwd_group_tma_kernel<DEG, false><<<grid, 128, 0, st>>>(g, r, c, o, m.dens, m.feat, tma_mode > 1 ? 1 : 0);
        return;
      }
    }
#endif
    // tuning hook: $R3D_FWD_CARVEOUT = preferred shared-memory carve-out in percent (the rest of the 228 KB is L1)
    static const int carve = [] {
      const char* e = getenv("R3D_FWD_CARVEOUT");
      const int v = e ? atoi(e) : -1;
      if (v >= 0) cudaFuncSetAttribute(render_fwd_group_kernel<DEG, false>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
      return v;
    }();
    (void)carve;
    render_fwd_group_kernel<DEG, false><<<grid, 128, 0, st>>>(g, r, c, o);
    return;
  }
#ifdef R3D_AB_VARIANTS
  if (vec == 8) {  // 256-bit record loads (R3D_FEATURE_PAD=8 layouts): measured, no gain
    render_fwd_kernel<DEG, 8><<<grid, 128, 0, st>>>(g, r, c, o);
    return;
  }
#endif
  if (vec != 0)
    render_fwd_kernel<DEG, 4><<<grid, 128, 0, st>>>(g, r, c, o);
  else
    render_fwd_kernel<DEG, 0><<<grid, 128, 0, st>>>(g, r, c, o);
}

#ifdef R3D_AB_VARIANTS
// warp-specialised backward (r3d_bwd_ws.cuh): ReLU field + the forward's sample cache and ballots
template <int DEG, bool DUAL>
static void launch_bwd_ws(dim3 grid, cudaStream_t st, const GridP& g, const RaysP& r, const CfgP& c, const BwdP& b) {
  const bool table = r.bounds == nullptr && !(c.flags & R3D_FLAG_OPTIMIZED_SAMPLING) && c.S <= 4096;
  const size_t smem = wsb_smem_bytes<DEG, DUAL>(c.S, table);
  static const bool attr = [] {
    cudaFuncSetAttribute(render_bwd_ws_kernel<DEG, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsb_smem_bytes<DEG, DUAL>(4096, true));
    return true;
  }();
  (void)attr;
  render_bwd_ws_kernel<DEG, DUAL><<<grid, 256, smem, st>>>(g, r, c, b, table ? 1 : 0);
}

#endif

template <int DEG>
static void launch_bwd(int vec, int variant, dim3 grid, cudaStream_t st, const GridP& g, const RaysP& r, const CfgP& c, const BwdP& b) {
#ifdef R3D_AB_VARIANTS
  if (variant & 1) {  // thread-per-ray scatter (kept for A/B measurement)
    if (vec == 8)
      render_bwd_kernel<DEG, 8><<<grid, 128, 0, st>>>(g, r, c, b);
    else if (vec == 4)
      render_bwd_kernel<DEG, 4><<<grid, 128, 0, st>>>(g, r, c, b);
    else
      render_bwd_kernel<DEG, 0><<<grid, 128, 0, st>>>(g, r, c, b);
    return;
  }
  if (mask_usable(g, b, vec) && (variant & 128)) {
    launch_bwd_ws<DEG, false>(grid, st, g, r, c, b);
    return;
  }
#else
  (void)variant;
#endif
  // 128-bit vector reductions whenever the layout allows (a stride that is a multiple of 8 floats is one of 4 too)
  if (mask_usable(g, b, vec))
    render_bwd_coop_kernel<DEG, 4, false, true><<<grid, 128, 0, st>>>(g, r, c, b);
  else if (vec != 0)
    render_bwd_coop_kernel<DEG, 4, false, false><<<grid, 128, 0, st>>>(g, r, c, b);
  else
    render_bwd_coop_kernel<DEG, 0, false, false><<<grid, 128, 0, st>>>(g, r, c, b);
}

// The single-pass specular + diffuse render exists in the lane-group forward and the cooperative backward only.
static bool dual_supported(const GridP& g, const CfgP& c, int vec) {
  return vec != 0 && !(c.flags & R3D_FLAG_DIFFUSE) && group_indexable(g);
}
template <int DEG>
static void launch_fwd_dual(dim3 grid, cudaStream_t st, const GridP& g, const RaysP& r, const CfgP& c, const OutP& o) {
  render_fwd_group_kernel<DEG, true><<<grid, 128, 0, st>>>(g, r, c, o);
}
template <int DEG>
static void launch_bwd_dual(int vec, dim3 grid, cudaStream_t st, const GridP& g, const RaysP& r, const CfgP& c, const BwdP& b) {
  if (mask_usable(g, b, vec))
    render_bwd_coop_kernel<DEG, 4, true, true><<<grid, 128, 0, st>>>(g, r, c, b);
  else
    render_bwd_coop_kernel<DEG, 4, true, false><<<grid, 128, 0, st>>>(g, r, c, b);
}

// widest vector access the feature layout allows: 8 floats (256-bit loads; records are whole 32-byte sectors),
// 4 floats (128-bit) or scalar (the reference's unpadded layout)
static int vector_width(const GridP& g, const void* grad_features) {
  const auto aligned = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const bool grad16 = grad_features == nullptr || aligned(grad_features, 16);
  if (g.stride % 8 == 0 && aligned(g.feat, 32) && grad16) return 8;
  if (g.stride % 4 == 0 && aligned(g.feat, 16) && grad16) return 4;
  return 0;
}

static int grid_blocks(const RaysP& r, dim3& grid) {
  const long long threads = threads_for_rays(r.n, r.tile_w, r.tile_h);
  const long long blocks = (threads + 127) / 128;
  if (blocks > 0x7fffffffLL) return fail(R3D_ERR_UNSUPPORTED, "too many rays for one launch (%lld)", (long long)r.n);
  grid = dim3((unsigned)blocks);
  return R3D_OK;
}

}  // namespace r3d

using namespace r3d;

extern "C" int64_t r3d_sample_mask_words(const R3dRays* rays) {
  RaysP r;
  if (to_device_params(rays, r)) return -1;
  return (threads_for_rays(r.n, r.tile_w, r.tile_h) + 127) / 128 * 4;
}

extern "C" int r3d_render_fwd(const R3dGrid* grid, const R3dRays* rays, const R3dRenderConfig* cfg, const R3dRenderOut* out,
                              void* cuda_stream) {
  GridP g;
  RaysP r;
  CfgP c;
  int rc;
  if ((rc = to_device_params(grid, g)) || (rc = to_device_params(rays, r)) || (rc = to_device_params(cfg, r, c))) return rc;
  if (r.n == 0) return R3D_OK;
  if (!out || !out->colour || !out->depth || !out->acc) return fail(R3D_ERR_INVALID_ARGUMENT, "render output buffers are NULL");
  OutP o{out->colour, out->depth, out->acc, out->disparity, reinterpret_cast<float4*>(out->sample_cache),
         out->colour_diffuse, reinterpret_cast<float4*>(out->sample_cache_diffuse), out->sample_mask};
  if ((o.cache && !aligned16(o.cache)) || (o.cache_diffuse && !aligned16(o.cache_diffuse)))
    return fail(R3D_ERR_INVALID_ARGUMENT, "sample_cache must be 16-byte aligned");
#ifndef R3D_AB_VARIANTS
  if (cfg->variant != 0) return fail(R3D_ERR_UNSUPPORTED, "kernel variant %d: A/B variants are compiled only with -DR3D_AB_VARIANTS", cfg->variant);
#endif
  dim3 blocks;
  if ((rc = grid_blocks(r, blocks))) return rc;
  const int vec = vector_width(g, nullptr);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (o.mask && (!o.cache || !fwd_uses_group_kernel(g, c, vec, o.colour_diffuse ? 0 : cfg->variant, grid->sh_degree)))
    return fail(R3D_ERR_UNSUPPORTED, "sample_mask needs sample_cache and the default forward kernel (16-byte aligned feature layout with a "
                                     "stride that is a multiple of 4 floats, no diffuse-only render, variant 0)");
  if (o.colour_diffuse) {  // single-pass specular + diffuse render
    if (!dual_supported(g, c, vec))
      return fail(R3D_ERR_UNSUPPORTED, "colour_diffuse (single-pass specular + diffuse render) needs a 16-byte aligned feature layout "
                                       "with a stride that is a multiple of 4 floats, and R3D_FLAG_DIFFUSE clear");
    switch (grid->sh_degree) {
      case 0: launch_fwd_dual<0>(blocks, st, g, r, c, o); break;
      case 1: launch_fwd_dual<1>(blocks, st, g, r, c, o); break;
      case 2: launch_fwd_dual<2>(blocks, st, g, r, c, o); break;
      default: launch_fwd_dual<3>(blocks, st, g, r, c, o); break;
    }
    return check_launch("r3d_render_fwd");
  }
  if (o.cache_diffuse) return fail(R3D_ERR_INVALID_ARGUMENT, "sample_cache_diffuse without colour_diffuse");
  switch (grid->sh_degree) {
    case 0: launch_fwd<0>(vec, cfg->variant, blocks, st, g, r, c, o); break;
    case 1: launch_fwd<1>(vec, cfg->variant, blocks, st, g, r, c, o); break;
    case 2: launch_fwd<2>(vec, cfg->variant, blocks, st, g, r, c, o); break;
    default: launch_fwd<3>(vec, cfg->variant, blocks, st, g, r, c, o); break;
  }
  return check_launch("r3d_render_fwd");
}

extern "C" int r3d_render_bwd(const R3dGrid* grid, const R3dRays* rays, const R3dRenderConfig* cfg, const R3dRenderOut* saved,
                              const R3dRenderOutGrad* grad_out, const R3dGridGrad* grad_grid, void* cuda_stream) {
  GridP g;
  RaysP r;
  CfgP c;
  int rc;
  if ((rc = to_device_params(grid, g)) || (rc = to_device_params(rays, r)) || (rc = to_device_params(cfg, r, c))) return rc;
  if (r.n == 0) return R3D_OK;
  if (!saved || !saved->colour || !saved->depth || !saved->acc)
    return fail(R3D_ERR_INVALID_ARGUMENT, "saved forward outputs (colour, depth, acc) are required by the backward pass");
  if (!grad_out || !grad_grid) return fail(R3D_ERR_INVALID_ARGUMENT, "grad_out / grad_grid is NULL");
  if (!grad_grid->densities && !grad_grid->features) return R3D_OK;
  BwdP b{saved->colour,   saved->depth,        saved->acc,           grad_out->colour,     grad_out->depth,
         grad_out->acc,   grad_out->disparity, grad_grid->densities, grad_grid->features,
         reinterpret_cast<const float4*>(saved->sample_cache),
         saved->colour_diffuse, grad_out->colour_diffuse, reinterpret_cast<const float4*>(saved->sample_cache_diffuse),
         saved->sample_mask};
  if ((b.cache && !aligned16(b.cache)) || (b.cache_diffuse && !aligned16(b.cache_diffuse)))
    return fail(R3D_ERR_INVALID_ARGUMENT, "sample_cache must be 16-byte aligned");
#ifndef R3D_AB_VARIANTS
  if (cfg->variant != 0) return fail(R3D_ERR_UNSUPPORTED, "kernel variant %d: A/B variants are compiled only with -DR3D_AB_VARIANTS", cfg->variant);
#endif
  dim3 blocks;
  if ((rc = grid_blocks(r, blocks))) return rc;
  const int vec = vector_width(g, b.gfeat);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (b.g_colour_diffuse) {  // backward of the single-pass specular + diffuse render
    if (!b.colour_diffuse) return fail(R3D_ERR_INVALID_ARGUMENT, "grad_out.colour_diffuse needs the saved colour_diffuse of the forward call");
    if (!dual_supported(g, c, vec))
      return fail(R3D_ERR_UNSUPPORTED, "colour_diffuse gradients need a 16-byte aligned feature / gradient layout with a stride that is a "
                                       "multiple of 4 floats, and R3D_FLAG_DIFFUSE clear");
    switch (grid->sh_degree) {
      case 0: launch_bwd_dual<0>(vec, blocks, st, g, r, c, b); break;
      case 1: launch_bwd_dual<1>(vec, blocks, st, g, r, c, b); break;
      case 2: launch_bwd_dual<2>(vec, blocks, st, g, r, c, b); break;
      default: launch_bwd_dual<3>(vec, blocks, st, g, r, c, b); break;
    }
    return check_launch("r3d_render_bwd");
  }
  b.colour_diffuse = nullptr, b.cache_diffuse = nullptr;  // no diffuse gradient: the plain kernels
  switch (grid->sh_degree) {
    case 0: launch_bwd<0>(vec, cfg->variant, blocks, st, g, r, c, b); break;
    case 1: launch_bwd<1>(vec, cfg->variant, blocks, st, g, r, c, b); break;
    case 2: launch_bwd<2>(vec, cfg->variant, blocks, st, g, r, c, b); break;
    default: launch_bwd<3>(vec, cfg->variant, blocks, st, g, r, c, b); break;
  }
  return check_launch("r3d_render_bwd");
}

extern "C" int r3d_mark_touched_voxels(const R3dGrid* grid, const R3dRays* rays, const R3dRenderConfig* cfg, uint8_t* bitmap,
                                       void* cuda_stream) {
  GridP g;
  RaysP r;
  CfgP c;
  int rc;
  if ((rc = to_device_params(grid, g)) || (rc = to_device_params(rays, r)) || (rc = to_device_params(cfg, r, c))) return rc;
  if (!bitmap) return fail(R3D_ERR_INVALID_ARGUMENT, "bitmap is NULL");
  if (r.n == 0) return R3D_OK;
  dim3 blocks;
  if ((rc = grid_blocks(r, blocks))) return rc;
  mark_touched_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(g, r, c, bitmap);
  return check_launch("r3d_mark_touched_voxels");
}

extern "C" int r3d_sample_statistics(const R3dGrid* grid, const R3dRays* rays, const R3dRenderConfig* cfg, uint64_t* counters,
                                     void* cuda_stream) {
  GridP g;
  RaysP r;
  CfgP c;
  int rc;
  if ((rc = to_device_params(grid, g)) || (rc = to_device_params(rays, r)) || (rc = to_device_params(cfg, r, c))) return rc;
  if (!counters) return fail(R3D_ERR_INVALID_ARGUMENT, "counters is NULL");
  if (r.n == 0) return R3D_OK;
  dim3 blocks;
  if ((rc = grid_blocks(r, blocks))) return rc;
  sample_stats_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(g, r, c, reinterpret_cast<unsigned long long*>(counters));
  return check_launch("r3d_sample_statistics");
}

extern "C" int r3d_has_ab_variants(void) {
#ifdef R3D_AB_VARIANTS
  return 1;
#else
  return 0;
#endif
}
