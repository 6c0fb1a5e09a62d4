// r3d_device.cuh -- device-side building blocks shared by the sm_100a kernels.
//
// Everything here restates the mathematical contract of the reference hot path (SURVEY.md
// appendix A); reference lines are cited relative to /root/reference/thre3d_atom/.
//
// Numerical policy
//   * Sample POSITIONS (depths z, points p = o + d*z, the normalised coordinate n = p*scale + bias)
//     are computed with explicitly un-fused fp32 operations (__fmul_rn / __fadd_rn) in the
//     reference's op order, because they feed the two discontinuous decisions of the algorithm:
//     the strict inside-AABB test and the floor() that selects the interpolation cell.
//   * Everything downstream (interpolation, SH, compositing) is free to use FMAs; results agree
//     with the fp32 reference to rounding (tolerances in tests/).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "r3d_b200.h"

namespace r3d {

constexpr float kZeroPlus = 1e-10f;  // utils/constants.py:7
constexpr float kInfinity = 1e10f;   // utils/constants.py:8

// ---------------------------------------------------------------------------------------------
// kernel parameter blocks (passed by value; plain data)
// ---------------------------------------------------------------------------------------------
struct GridP {
  const float* __restrict__ dens;
  const float* __restrict__ feat;
  const float4* __restrict__ quads;  // optional density quad volume (see QuadDims), nullptr = probe `dens` directly
  int W, D, H;
  int F, stride, K;  // features per voxel, record stride (floats), SH coeffs per colour channel
  float lo[3], hi[3], ns[3], nb[3];
  float dscale;
  int pre, post;
};

struct RaysP {
  const float* __restrict__ origins;
  const float* __restrict__ directions;
  const float* __restrict__ bounds;
  long long n;
  int tile_w, tile_h;  // image shape hint (0 = flat list)
  int has_camera;
  R3dCamera cam;
};

struct CfgP {
  int S;
  float near, far;
  unsigned flags;
  const float* __restrict__ jitter;
  unsigned seed_lo, seed_hi;
};

// ---------------------------------------------------------------------------------------------
// thread -> ray mapping.  With an image-shape hint a warp owns an 8x4 pixel tile and a 128-thread
// CTA a 16x8 block of pixels, so that the 32 lanes of a gather instruction fall into a handful of
// voxel cells (adjacent pixels are ~0.3 voxel apart at the BASELINE shapes) and share L1 lines.
// Without the hint rays are taken in list order.  The mapping only permutes which thread does
// which ray; per-ray results do not depend on it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long thread_to_ray(const RaysP& r, long long t) {
  if (r.tile_w <= 0) return t < r.n ? t : -1;
  const int ctas_x = (r.tile_w + 15) >> 4;
  const long long cta = t >> 7;
  const int in_cta = (int)(t & 127);
  const int warp = in_cta >> 5, lane = in_cta & 31;
  const int cta_x = (int)(cta % ctas_x);
  const long long cta_y = cta / ctas_x;
  const int x = cta_x * 16 + (warp & 1) * 8 + (lane & 7);
  const long long y = cta_y * 8 + (warp >> 1) * 4 + (lane >> 3);
  if (x >= r.tile_w || y >= r.tile_h) return -1;
  const long long ray = y * r.tile_w + x;
  return ray < r.n ? ray : -1;
}

__host__ __device__ inline long long threads_for_rays(long long n, int tile_w, int tile_h) {
  if (tile_w <= 0) return n;
  const long long ctas_x = (tile_w + 15) >> 4, ctas_y = (tile_h + 7) >> 3;
  return ctas_x * ctas_y * 128;
}

// ---------------------------------------------------------------------------------------------
// rays: either loaded, or generated in-kernel exactly like cast_rays
// (rendering/volumetric/utils/misc.py:27-50): pixel centre (x+0.5, y+0.5),
// dir_cam = ((x+.5 - W/2)/f, -(y+.5 - H/2)/f, -1), d = R * dir_cam, o = t.
// ---------------------------------------------------------------------------------------------
struct Ray {
  float ox, oy, oz, dx, dy, dz;
};

__device__ __forceinline__ void camera_ray(const R3dCamera& c, int x, int y, Ray& r) {
  const float cx = __fdiv_rn(__fsub_rn((float)x + 0.5f, (float)c.width * 0.5f), c.focal);
  const float cy = -__fdiv_rn(__fsub_rn((float)y + 0.5f, (float)c.height * 0.5f), c.focal);
  const float cz = -1.0f;
  r.dx = fmaf(c.rotation[2], cz, fmaf(c.rotation[1], cy, c.rotation[0] * cx));
  r.dy = fmaf(c.rotation[5], cz, fmaf(c.rotation[4], cy, c.rotation[3] * cx));
  r.dz = fmaf(c.rotation[8], cz, fmaf(c.rotation[7], cy, c.rotation[6] * cx));
  r.ox = c.translation[0];
  r.oy = c.translation[1];
  r.oz = c.translation[2];
}

__device__ __forceinline__ Ray load_ray(const RaysP& rp, long long ray) {
  Ray r;
  if (rp.has_camera) {
    camera_ray(rp.cam, (int)(ray % rp.cam.width), (int)(ray / rp.cam.width), r);
  } else {
    const float* o = rp.origins + 3 * ray;
    const float* d = rp.directions + 3 * ray;
    r.ox = __ldg(o), r.oy = __ldg(o + 1), r.oz = __ldg(o + 2);
    r.dx = __ldg(d), r.dy = __ldg(d + 1), r.dz = __ldg(d + 2);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// counter-based stratified-jitter RNG: u(ray, sample) in [0,1).  Stateless, so the backward pass
// regenerates the forward's offsets.  (The reference draws torch.rand[N,S], sample.py:63; a caller
// that needs that exact stream passes it through R3dRenderConfig.jitter instead.)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned mix32(unsigned h) {
  h ^= h >> 16;
  h *= 0x7feb352dU;
  h ^= h >> 15;
  h *= 0x846ca68bU;
  h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ unsigned ray_rng_key(unsigned seed_lo, unsigned seed_hi, long long ray) {
  unsigned k = mix32((unsigned)ray + seed_lo);
  return mix32(k ^ seed_hi ^ (unsigned)((unsigned long long)ray >> 32));
}
__host__ __device__ __forceinline__ float jitter_u(unsigned key, int sample) {
  return (float)(mix32(key + (unsigned)sample * 0x9E3779B9U) >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------------
// per-ray depth generator (rendering/volumetric/sample.py:38-64)
//   t_i = linspace(0,1,S)[i]  (ATen per-element formula, fused multiply-add as on the device)
//   z_i = near*(1-t_i) + far*t_i
//   perturb: z_i <- lower_i + (upper_i - lower_i) * u_i, mid-point strata
// ---------------------------------------------------------------------------------------------
struct DepthGen {
  float near, far, step;
  int S, half;
  bool perturb;
  const float* __restrict__ jit;  // row of the explicit jitter tensor or nullptr
  unsigned key;

  __device__ __forceinline__ float base(int i) const {
    // linspace(0, 1, 1) == [0]: step is 0 for S == 1, which makes the first branch right for it too
    const float t = (i < half || S == 1) ? step * (float)i : fmaf(-step, (float)(S - 1 - i), 1.0f);
    return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
  }
  __device__ __forceinline__ float at(int i) const {
    const float b = base(i);
    if (!perturb) return b;
    const float lower = (i > 0) ? 0.5f * __fadd_rn(b, base(i - 1)) : b;
    const float upper = (i < S - 1) ? 0.5f * __fadd_rn(base(i + 1), b) : b;
    const float u = jit ? __ldg(jit + i) : jitter_u(key, i);
    return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
  }
};

// Depths of CONSECUTIVE samples j, j+1, j+2, ...: same values as DepthGen::at (bit for bit), but the two unjittered
// depths a stratum shares with its predecessor are carried in registers, so a step evaluates one base() instead of three.
struct DepthMarch {
  float bm, bc;  // base(j - 1), base(j) of the next sample j to be produced
  __device__ __forceinline__ void start(const DepthGen& dg, int j) {
    bc = dg.base(j);
    bm = (dg.perturb && j > 0) ? dg.base(j - 1) : bc;
  }
  __device__ __forceinline__ float next(const DepthGen& dg, int j) {
    const float b = bc;
    const float bn = (j < dg.S - 1) ? dg.base(j + 1) : b;
    float z = b;
    if (dg.perturb) {
      const float lower = (j > 0) ? 0.5f * __fadd_rn(b, bm) : b;
      const float upper = (j < dg.S - 1) ? 0.5f * __fadd_rn(bn, b) : b;
      const float u = dg.jit ? __ldg(dg.jit + j) : jitter_u(dg.key, j);
      z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
    }
    bm = b, bc = bn;
    return z;
  }
};

// Per-ray (near, far) of `optimized_sampling` -- the reference's slab test with all its quirks
// (rendering/volumetric/sample.py:71-183): denominators d + 1e-10, the miss test of an axis uses the
// interval accumulated over the previous axes, misses fall back to the camera bounds, clip at 0.
__device__ __forceinline__ void reference_aabb_bounds(const GridP& g, const Ray& r, float cam_near, float cam_far,
                                                      float& near, float& far) {
  const float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
  float lo = 0.f, hi = 0.f;
  bool hit = true;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float den = __fadd_rn(d[ax], kZeroPlus);
    const float t0 = __fdiv_rn(__fsub_rn(g.lo[ax], o[ax]), den);
    const float t1 = __fdiv_rn(__fsub_rn(g.hi[ax], o[ax]), den);
    const float ta = (t0 > t1) ? t1 : t0, tb = (t0 > t1) ? t0 : t1;
    if (ax == 0) {
      lo = ta, hi = tb;
    } else {
      if ((lo > tb) || (ta > hi)) hit = false;
      lo = (ta > lo) ? ta : lo;
      hi = (tb < hi) ? tb : hi;
    }
  }
  if (!hit) lo = cam_near, hi = cam_far;
  near = fmaxf(lo, 0.0f);  // torch.clip(min=0) propagates NaN; NaN bounds poison the ray either way
  far = fmaxf(hi, 0.0f);
  if (lo != lo) near = lo;
  if (hi != hi) far = hi;
}

// Conservative range [i_lo, i_hi] of sample indices that can lie strictly inside the AABB.
// Samples outside the AABB contribute exactly zero (sigma := 0, process.py:80-84), so skipping
// them is exact; the strict per-sample test is still applied inside the range.
__device__ __forceinline__ void sample_range(const GridP& g, const Ray& r, float near, float far, int S, int& i_lo,
                                             int& i_hi) {
  i_lo = 0, i_hi = S - 1;
  const float span = far - near;
  if (!(span > 0.0f) || S < 2) return;  // degenerate / NaN bounds: visit everything
  const float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
  float t0 = -3.0e38f, t1 = 3.0e38f;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    if (d[ax] != 0.0f) {
      const float inv = 1.0f / d[ax];
      const float a = (g.lo[ax] - o[ax]) * inv, b = (g.hi[ax] - o[ax]) * inv;
      t0 = fmaxf(t0, fminf(a, b));
      t1 = fminf(t1, fmaxf(a, b));
    } else if (!(o[ax] > g.lo[ax] && o[ax] < g.hi[ax])) {
      t1 = -3.0e38f;  // parallel and outside the slab: never inside
    }
  }
  if (!(t0 <= t1)) {  // miss (or NaN): nothing to visit
    i_lo = 1, i_hi = 0;
    return;
  }
  const float to_idx = (float)(S - 1) / span;
  const float f_lo = floorf((t0 - near) * to_idx) - 2.0f;  // +-1 for the jitter stratum, +-1 for rounding
  const float f_hi = ceilf((t1 - near) * to_idx) + 2.0f;
  if (f_lo > (float)(S - 1) || f_hi < 0.0f) {
    i_lo = 1, i_hi = 0;
    return;
  }
  i_lo = f_lo > 0.0f ? (int)f_lo : 0;
  i_hi = f_hi < (float)(S - 1) ? (int)f_hi : S - 1;
}

// ---------------------------------------------------------------------------------------------
// trilinear cell of a point: VoxelGrid._normalize_points (voxels.py:214-223) followed by
// grid_sample(align_corners=False, padding zeros): i = ((n+1)*dim - 1)/2, corners floor(i), floor(i)+1,
// out-of-range corners contribute zero.
// ---------------------------------------------------------------------------------------------
struct Cell {
  int ox[2], oy[2], oz[2];  // clamped linear-offset parts: x*D*H, y*H, z
  float wx[2], wy[2], wz[2];  // weights, zeroed for out-of-range corners
};

__device__ __forceinline__ void axis_cell(float p, float ns, float nb, int dim, int mul, int (&off)[2], float (&w)[2]) {
  const float n = __fadd_rn(__fmul_rn(p, ns), nb);
  const float gi = ((n + 1.0f) * (float)dim - 1.0f) * 0.5f;
  // keep the float->int conversion in range for far-away points (point-lookup API); such points
  // have both corners out of range and therefore weight zero.
  const float gc = fminf(fmaxf(gi, -2.0f), (float)dim + 1.0f);
  const float fl = floorf(gc);
  const int i0 = (int)fl;
  const float f1 = gc - fl, f0 = (fl + 1.0f) - gc;
  const bool v0 = (i0 >= 0) && (i0 < dim) && (gi == gc);
  const bool v1 = (i0 + 1 >= 0) && (i0 + 1 < dim) && (gi == gc);
  w[0] = v0 ? f0 : 0.0f;
  w[1] = v1 ? f1 : 0.0f;
  off[0] = min(max(i0, 0), dim - 1) * mul;
  off[1] = min(max(i0 + 1, 0), dim - 1) * mul;
}

__device__ __forceinline__ void make_cell(const GridP& g, float px, float py, float pz, Cell& c) {
  axis_cell(px, g.ns[0], g.nb[0], g.W, g.D * g.H, c.ox, c.wx);
  axis_cell(py, g.ns[1], g.nb[1], g.D, g.H, c.oy, c.wy);
  axis_cell(pz, g.ns[2], g.nb[2], g.H, 1, c.oz, c.wz);
}

// Same cell for a point that passed inside_aabb: n is then within a rounding error of [-1, 1], so the continuous index
// lies in (-0.5 - eps, dim - 0.5 + eps) and the float-side range clamp of axis_cell is the identity (bit-identical
// weights and offsets); the integer clamps stay, they keep every address in range whatever the caller passed.
__device__ __forceinline__ void axis_cell_inside(float p, float ns, float nb, int dim, int mul, int (&off)[2], float (&w)[2]) {
  const float n = __fadd_rn(__fmul_rn(p, ns), nb);
  const float gi = ((n + 1.0f) * (float)dim - 1.0f) * 0.5f;
  const float fl = floorf(gi);
  const int i0 = (int)fl;
  w[0] = ((unsigned)i0 < (unsigned)dim) ? (fl + 1.0f) - gi : 0.0f;
  w[1] = ((unsigned)(i0 + 1) < (unsigned)dim) ? gi - fl : 0.0f;
  off[0] = min(max(i0, 0), dim - 1) * mul;
  off[1] = min(max(i0 + 1, 0), dim - 1) * mul;
}

__device__ __forceinline__ void make_cell_inside(const GridP& g, float px, float py, float pz, Cell& c) {
  axis_cell_inside(px, g.ns[0], g.nb[0], g.W, g.D * g.H, c.ox, c.wx);
  axis_cell_inside(py, g.ns[1], g.nb[1], g.D, g.H, c.oy, c.wy);
  axis_cell_inside(pz, g.ns[2], g.nb[2], g.H, 1, c.oz, c.wz);
}

__device__ __forceinline__ bool inside_aabb(const GridP& g, float px, float py, float pz) {
  // strict inequalities, voxels.py:252-274
  return (px > g.lo[0]) && (px < g.hi[0]) && (py > g.lo[1]) && (py < g.hi[1]) && (pz > g.lo[2]) && (pz < g.hi[2]);
}

// ---------------------------------------------------------------------------------------------
// Density quad volume: a derived, zero-padded copy of the (pre-activated, un-scaled) densities laid out for the
// probe.  Entry (cx, cy, cz), cx in [0, W+2), cy in [0, D+1), cz in [0, H+1), belongs to the interpolation cell with low
// corner (x, y, z) = (cx-1, cy-1, cz-1) and holds the four values of its x-plane:
//     ( v[x][y][z], v[x][y][z+1], v[x][y+1][z], v[x][y+1][z+1] ),   out-of-range voxels = 0
// so a sample's 8 corner densities are TWO 16-byte loads (entries (cx,cy,cz) and (cx+1,cy,cz)) with no clamping and no
// validity logic: the zero padding of grid_sample (voxels.py:296-303, padding_mode zeros) is physically in the table.
// 16 B per entry: 273 MB at 256^3; rebuilt from the parameters by quads_build_kernel (r3d_aux.cu) in ~0.06 ms.
// ---------------------------------------------------------------------------------------------
struct CellQ {
  int ix, iy, iz;              // low corner, each in [-1, dim-1]
  float wx[2], wy[2], wz[2];   // plain trilinear weights (not zeroed for out-of-range corners)
};

__device__ __forceinline__ void axis_cell_q(float p, float ns, float nb, int dim, int& i0, float (&w)[2]) {
  const float n = __fadd_rn(__fmul_rn(p, ns), nb);
  const float gi = ((n + 1.0f) * (float)dim - 1.0f) * 0.5f;
  const float fl = floorf(gi);
  w[0] = (fl + 1.0f) - gi;
  w[1] = gi - fl;
  // a point that passed inside_aabb has floor(gi) in [-1, dim-1] (see axis_cell_inside); the clamp only keeps the
  // address in range whatever the caller passed
  i0 = min(max((int)fl, -1), dim - 1);
}

__device__ __forceinline__ void make_cell_q(const GridP& g, float px, float py, float pz, CellQ& c) {
  axis_cell_q(px, g.ns[0], g.nb[0], g.W, c.ix, c.wx);
  axis_cell_q(py, g.ns[1], g.nb[1], g.D, c.iy, c.wy);
  axis_cell_q(pz, g.ns[2], g.nb[2], g.H, c.iz, c.wz);
}

// the clamped offsets / zeroed weights of the same cell (bit-identical to make_cell_inside), needed only by samples
// that go on to gather feature records
__device__ __forceinline__ void cell_from_q(const GridP& g, const CellQ& q, Cell& c) {
  const int dims[3] = {g.W, g.D, g.H}, mul[3] = {g.D * g.H, g.H, 1}, i0[3] = {q.ix, q.iy, q.iz};
  const float* w[3] = {q.wx, q.wy, q.wz};
  int* off[3] = {c.ox, c.oy, c.oz};
  float* wo[3] = {c.wx, c.wy, c.wz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    wo[a][0] = (i0[a] >= 0) ? w[a][0] : 0.0f;
    wo[a][1] = (i0[a] + 1 < dims[a]) ? w[a][1] : 0.0f;
    off[a][0] = max(i0[a], 0) * mul[a];
    off[a][1] = min(i0[a] + 1, dims[a] - 1) * mul[a];
  }
}

// same value as density_pre_interp (same products, same order of accumulation; an out-of-range corner adds w * 0
// instead of 0 * v)
__device__ __forceinline__ float density_pre_interp_q(const GridP& g, const CellQ& c) {
  const unsigned plane = (unsigned)(g.D + 1) * (unsigned)(g.H + 1);
  const unsigned qi = (unsigned)(c.ix + 1) * plane + (unsigned)(c.iy + 1) * (unsigned)(g.H + 1) + (unsigned)(c.iz + 1);
  const float4 a = __ldg(g.quads + qi), b = __ldg(g.quads + (qi + plane));
  float s = 0.0f;
  float wxy = c.wx[0] * c.wy[0];
  s = fmaf(wxy * c.wz[0], a.x, s), s = fmaf(wxy * c.wz[1], a.y, s);
  wxy = c.wx[0] * c.wy[1];
  s = fmaf(wxy * c.wz[0], a.z, s), s = fmaf(wxy * c.wz[1], a.w, s);
  wxy = c.wx[1] * c.wy[0];
  s = fmaf(wxy * c.wz[0], b.x, s), s = fmaf(wxy * c.wz[1], b.y, s);
  wxy = c.wx[1] * c.wy[1];
  s = fmaf(wxy * c.wz[0], b.z, s), s = fmaf(wxy * c.wz[1], b.w, s);
  return s * (g.pre == R3D_PRE_ABS ? fabsf(g.dscale) : g.dscale);
}

// interpolated, pre-activated, scaled density (voxels.py:292-308) -- before the post-activation.
// U32: address the 8 corners with unsigned 32-bit voxel indices (one IADD3 + one IMAD.WIDE.U32 per address instead of
// sign extension and 64-bit adds: -24 instructions per marching step).  The forward kernels use it; the cooperative
// backward is compiled for 96 registers and the different register allocation made it spill (5.6 -> 6.5 ms), so it
// keeps the 64-bit form.
template <bool U32 = false>
__device__ __forceinline__ float density_pre_interp(const GridP& g, const Cell& c) {
  float s = 0.0f;
#pragma unroll
  for (int ix = 0; ix < 2; ++ix)
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
      const float wxy = c.wx[ix] * c.wy[iy];
      float v0, v1;
      if constexpr (U32) {  // < 2^31 voxels, checked on the host
        const unsigned col = (unsigned)(c.ox[ix] + c.oy[iy]);
        v0 = __ldg(g.dens + (col + (unsigned)c.oz[0])), v1 = __ldg(g.dens + (col + (unsigned)c.oz[1]));
      } else {
        const float* p = g.dens + (size_t)(c.ox[ix] + c.oy[iy]);
        v0 = __ldg(p + c.oz[0]), v1 = __ldg(p + c.oz[1]);
      }
      if (g.pre == R3D_PRE_ABS) v0 = fabsf(v0), v1 = fabsf(v1);
      s = fmaf(wxy * c.wz[0], v0, s);
      s = fmaf(wxy * c.wz[1], v1, s);
    }
  return s * (g.pre == R3D_PRE_ABS ? fabsf(g.dscale) : g.dscale);
}

// density post-activation (voxels.py:309) and its derivative
__device__ __forceinline__ float density_post(int post, float x, float& dpost) {
  if (post == R3D_POST_RELU) {
    dpost = x > 0.0f ? 1.0f : 0.0f;
    return fmaxf(x, 0.0f);
  }
  if (post == R3D_POST_SOFTPLUS) {  // torch.nn.Softplus(beta=1, threshold=20)
    if (x > 20.0f) {
      dpost = 1.0f;
      return x;
    }
    const float e = expf(x);
    dpost = e / (1.0f + e);
    return log1pf(e);
  }
  dpost = 1.0f;
  return x;
}

// ---------------------------------------------------------------------------------------------
// signed real SH basis, so that raw_ch = sum_k Y[k] * coeff[ch][k]
// (rendering/volumetric/utils/spherical_harmonics.py:33-50 constants, :86-116 ladder)
// ---------------------------------------------------------------------------------------------
template <int DEG>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float (&Y)[(DEG + 1) * (DEG + 1)]) {
  Y[0] = 0.28209479177387814f;
  if constexpr (DEG > 0) {
    Y[1] = -0.4886025119029199f * y;
    Y[2] = 0.4886025119029199f * z;
    Y[3] = -0.4886025119029199f * x;
  }
  if constexpr (DEG > 1) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = 1.0925484305920792f * xy;
    Y[5] = -1.0925484305920792f * yz;
    Y[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    Y[7] = -1.0925484305920792f * xz;
    Y[8] = 0.5462742152960396f * (xx - yy);
    if constexpr (DEG > 2) {
      Y[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
      Y[10] = 2.890611442640554f * xy * z;
      Y[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
      Y[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
      Y[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
      Y[14] = 1.445305721320277f * z * (xx - yy);
      Y[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
    }
  }
}

// Transcendentals of the compositing stage.  Default: the SFU forms (ex2.approx behind __expf, rcp.approx behind
// __fdividef) -- 2 + floor(|1.44 x|) ulp on exp, i.e. ~2e-7 relative for the sigma*delta and raw-radiance ranges that
// matter, against a stated colour tolerance of 1e-5; ~6 instructions instead of ~25 per sample.  -DR3D_PRECISE_MATH=1
// selects expf and IEEE division.
#ifndef R3D_PRECISE_MATH
#define R3D_PRECISE_MATH 0
#endif
__device__ __forceinline__ float exp_neg(float x) {  // exp(-x)
#if R3D_PRECISE_MATH
  return expf(-x);
#else
  return __expf(-x);
#endif
}
__device__ __forceinline__ float sigmoidf_(float x) {
#if R3D_PRECISE_MATH
  return 1.0f / (1.0f + expf(-x));
#else
  return __fdividef(1.0f, 1.0f + __expf(-x));  // exp -> inf gives 0, exp -> 0 gives 1, like the IEEE form
#endif
}

// Packed fp32 FMA (new on sm_100: SASS FFMA2 Rd, Ra.F32x2, Rb.F32 (broadcast), Rc.F32x2): a.xy * b + c.xy in ONE issue slot.
// Measured on the B200 (profiles/r02_ubench_ffma2.json): the FMA rate stays 128 lanes/clk/SM, the issue slots halve.
#ifndef R3D_FFMA2
#define R3D_FFMA2 1
#endif
__device__ __forceinline__ float2 ffma2(const float2 a, const float b, const float2 c) {
#if R3D_FFMA2
  // packed fp32 FMA (sm_100: FFMA2 Rd, Ra.F32x2, Rb.F32 (broadcast), Rc.F32x2): two FMAs per issue slot
  const float2 bb = make_float2(b, b);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(rd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&bb)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
#else
  return make_float2(fmaf(a.x, b, c.x), fmaf(a.y, b, c.y));
#endif
}

}  // namespace r3d
