// r3d_aux.cu -- the small kernels around the fused renderer:
//   r3d_cast_rays          cast_rays                      rendering/volumetric/utils/misc.py:12-50
//   r3d_grid_lookup_fwd    VoxelGrid.forward + test_inside_volume   thre3d_reprs/voxels.py:252-331
//   r3d_grid_lookup_bwd    autograd of the lookup into _densities/_features
//   r3d_adam_step          torch.optim.Adam on the dense grid       modules/trainers.py:242-245,341
#include "r3d_host.h"

namespace r3d {

__global__ void __launch_bounds__(256) cast_rays_kernel(const R3dCamera cam, float* __restrict__ origins, float* __restrict__ directions) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)cam.height * cam.width;
  if (i >= n) return;
  Ray r;
  camera_ray(cam, (int)(i % cam.width), (int)(i / cam.width), r);
  origins[3 * i] = r.ox, origins[3 * i + 1] = r.oy, origins[3 * i + 2] = r.oz;
  directions[3 * i] = r.dx, directions[3 * i + 1] = r.dy, directions[3 * i + 2] = r.dz;
}

// One warp per point, lanes over the channels of a voxel record: every corner record is read with
// one coalesced request (F contiguous floats), which is the natural mapping for scattered points
// that share no cells.
__global__ void __launch_bounds__(256) lookup_fwd_kernel(const GridP g, const float* __restrict__ points, long long n,
                                                         float* __restrict__ out, uint8_t* __restrict__ inside) {
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n) return;
  const float px = __ldg(points + 3 * p), py = __ldg(points + 3 * p + 1), pz = __ldg(points + 3 * p + 2);
  Cell c;
  make_cell(g, px, py, pz, c);
  if (lane == 0) {
    float dpost;
    out[p * (g.F + 1) + g.F] = density_post(g.post, density_pre_interp(g, c), dpost);
    if (inside) inside[p] = inside_aabb(g, px, py, pz) ? 1 : 0;
  }
  for (int e = lane; e < g.F; e += 32) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
      const float w = c.wx[ix] * c.wy[iy] * c.wz[iz];
      const size_t vox = (size_t)(c.ox[ix] + c.oy[iy] + c.oz[iz]);
      acc = fmaf(w, __ldg(g.feat + vox * (size_t)g.stride + e), acc);
    }
    out[p * (g.F + 1) + e] = acc;
  }
}

__global__ void __launch_bounds__(256) lookup_bwd_kernel(const GridP g, const float* __restrict__ points, long long n,
                                                         const float* __restrict__ gout, float* __restrict__ gdens,
                                                         float* __restrict__ gfeat) {
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n) return;
  const float px = __ldg(points + 3 * p), py = __ldg(points + 3 * p + 1), pz = __ldg(points + 3 * p + 2);
  Cell c;
  make_cell(g, px, py, pz, c);
  if (lane == 0 && gdens) {
    float dpost;
    density_post(g.post, density_pre_interp(g, c), dpost);
    const float dpre = __ldg(gout + p * (g.F + 1) + g.F) * dpost * (g.pre == R3D_PRE_ABS ? fabsf(g.dscale) : g.dscale);
    if (dpre != 0.f) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
        const float w = c.wx[ix] * c.wy[iy] * c.wz[iz];
        if (w == 0.f) continue;
        const size_t vox = (size_t)(c.ox[ix] + c.oy[iy] + c.oz[iz]);
        float gv = w * dpre;
        if (g.pre == R3D_PRE_ABS) {
          const float v = __ldg(g.dens + vox);
          gv = (v > 0.f) ? gv : ((v < 0.f) ? -gv : 0.f);
        }
        atomicAdd(gdens + vox, gv);
      }
    }
  }
  if (!gfeat) return;
  for (int e = lane; e < g.F; e += 32) {
    const float go = __ldg(gout + p * (g.F + 1) + e);
    if (go == 0.f) continue;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
      const float w = c.wx[ix] * c.wy[iy] * c.wz[iz];
      if (w == 0.f) continue;
      const size_t vox = (size_t)(c.ox[ix] + c.oy[iy] + c.oz[iz]);
      atomicAdd(gfeat + vox * (size_t)g.stride + e, w * go);
    }
  }
}

// Adam, torch.optim.Adam semantics (no weight decay / amsgrad / maximize):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// Pure streaming kernel: 4 reads + 3 writes of 4 B per element.  A 16-byte group whose gradient AND moments are all zero (voxels no
// ray has reached since training began: g = 0, m = v = 0  =>  m' = v' = 0 and p' = p - step * 0 / eps = p exactly) keeps its three
// stores: the update is the identity there, 12 of its 28 bytes of traffic are not spent.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ gr, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2_sqrt, float gscale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  const float step = lr / bc1;
#ifndef R3D_ADAM_UNROLL
#define R3D_ADAM_UNROLL 4  // measured inside the bench step at 256^3 deg 2: 2.26 / 2.10-2.38 / 2.18 ms at 1 / 2 / 4
#endif
  constexpr int U = R3D_ADAM_UNROLL;  // 16-byte groups per thread and iteration (all loads issued before the arithmetic)
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += U * stride) {
    float4 P[U], G[U], M[U], V[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        P[u] = reinterpret_cast<float4*>(p)[i];
        G[u] = __ldg(reinterpret_cast<const float4*>(gr) + i);
        M[u] = reinterpret_cast<float4*>(m)[i], V[u] = reinterpret_cast<float4*>(v)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n4) continue;
      const bool still = G[u].x == 0.0f && G[u].y == 0.0f && G[u].z == 0.0f && G[u].w == 0.0f && M[u].x == 0.0f && M[u].y == 0.0f &&
                         M[u].z == 0.0f && M[u].w == 0.0f && V[u].x == 0.0f && V[u].y == 0.0f && V[u].z == 0.0f && V[u].w == 0.0f;
      if (still) continue;
#define R3D_ADAM1(c)                                                \
  {                                                                 \
    const float gg = G[u].c * gscale;                               \
    M[u].c = fmaf(b1, M[u].c, (1.0f - b1) * gg);                    \
    V[u].c = fmaf(b2, V[u].c, (1.0f - b2) * gg * gg);               \
    P[u].c -= step * (M[u].c / (sqrtf(V[u].c) / bc2_sqrt + eps));   \
  }
      R3D_ADAM1(x) R3D_ADAM1(y) R3D_ADAM1(z) R3D_ADAM1(w)
#undef R3D_ADAM1
      reinterpret_cast<float4*>(p)[i] = P[u];
      reinterpret_cast<float4*>(m)[i] = M[u];
      reinterpret_cast<float4*>(v)[i] = V[u];
    }
  }
  // tail (n % 4 elements)
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gg = gr[i] * gscale;
    const float mm = fmaf(b1, m[i], (1.0f - b1) * gg);
    const float vv = fmaf(b2, v[i], (1.0f - b2) * gg * gg);
    m[i] = mm, v[i] = vv;
    p[i] -= step * (mm / (sqrtf(vv) / bc2_sqrt + eps));
  }
#undef R3D_ADAM1
}

}  // namespace r3d

using namespace r3d;

// ---- training-batch sampler: rays + target pixels of random pixels (or random 8x4 pixel tiles) of a set of posed views ----
// Replaces, per training iteration, cast_rays for every cached view + randperm over all their pixels + three gathers
// (reference modules/trainers.py:281-303, rendering/volumetric/utils/misc.py:117-129): nothing of size V*H*W is touched.
__global__ void __launch_bounds__(256) sample_ray_batch_kernel(const float* __restrict__ rot, const float* __restrict__ trans,
                                                               const float* __restrict__ images, int V, int H, int W, float focal,
                                                               long long batch, int tile_w, int tile_h, unsigned seed_lo, unsigned seed_hi,
                                                               float* __restrict__ origins, float* __restrict__ directions,
                                                               float* __restrict__ pixels, long long* __restrict__ indices) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= batch) return;
  const int per_tile = tile_w * tile_h;
  const long long pick = t / per_tile;  // one random draw per tile (per pixel when the tile is 1x1)
  const int in_tile = (int)(t % per_tile);
  const unsigned key = ray_rng_key(seed_lo, seed_hi, pick);
  const unsigned r0 = mix32(key + 0x9E3779B9U), r1 = mix32(key + 2u * 0x9E3779B9U), r2 = mix32(key + 3u * 0x9E3779B9U);
  const int tiles_x = W / tile_w, tiles_y = H / tile_h;  // whole tiles only (checked on the host)
  const int v = (int)(((unsigned long long)r0 * (unsigned)V) >> 32);
  const int tx = (int)(((unsigned long long)r1 * (unsigned)tiles_x) >> 32), ty = (int)(((unsigned long long)r2 * (unsigned)tiles_y) >> 32);
  const int x = tx * tile_w + in_tile % tile_w, y = ty * tile_h + in_tile / tile_w;
  R3dCamera cam;
  cam.height = H, cam.width = W, cam.focal = focal;
#pragma unroll
  for (int k = 0; k < 9; ++k) cam.rotation[k] = __ldg(rot + 9 * v + k);
#pragma unroll
  for (int k = 0; k < 3; ++k) cam.translation[k] = __ldg(trans + 3 * v + k);
  Ray r;
  camera_ray(cam, x, y, r);  // bit-identical to cast_rays for that pixel
  origins[3 * t] = r.ox, origins[3 * t + 1] = r.oy, origins[3 * t + 2] = r.oz;
  directions[3 * t] = r.dx, directions[3 * t + 1] = r.dy, directions[3 * t + 2] = r.dz;
  const long long flat = ((long long)v * H + y) * W + x;
  if (pixels && images) {
    pixels[3 * t] = __ldg(images + 3 * flat), pixels[3 * t + 1] = __ldg(images + 3 * flat + 1), pixels[3 * t + 2] = __ldg(images + 3 * flat + 2);
  }
  if (indices) indices[t] = flat;
}

extern "C" int r3d_sample_ray_batch(const R3dViewSet* views, int64_t batch, int32_t tile_width, int32_t tile_height, uint64_t seed,
                                    float* origins, float* directions, float* pixels, int64_t* indices, void* cuda_stream) {
  if (!views || !views->rotations || !views->translations) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_sample_ray_batch: views / poses are NULL");
  if (views->num_views < 1 || views->height < 1 || views->width < 1 || !(views->focal > 0.f))
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_sample_ray_batch: bad view set (%d views of %dx%d)", views->num_views, views->height, views->width);
  if (tile_width < 1 || tile_height < 1 || views->width % tile_width != 0 || views->height % tile_height != 0)
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_sample_ray_batch: the image (%dx%d) must be a whole number of %dx%d tiles", views->height,
                views->width, tile_height, tile_width);
  if (batch < 0 || batch % ((int64_t)tile_width * tile_height) != 0)
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_sample_ray_batch: batch (%lld) must be a multiple of the tile size", (long long)batch);
  if (batch == 0) return R3D_OK;
  if (!origins || !directions) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_sample_ray_batch: output buffers are NULL");
  if (pixels && !views->images) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_sample_ray_batch: pixels requested but views->images is NULL");
  const long long blocks = (batch + 255) / 256;
  sample_ray_batch_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      views->rotations, views->translations, views->images, views->num_views, views->height, views->width, views->focal, batch, tile_width,
      tile_height, (unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32), origins, directions, pixels, reinterpret_cast<long long*>(indices));
  return check_launch("r3d_sample_ray_batch");
}

// ---- density quad volume (r3d_device.cuh: CellQ / density_pre_interp_q) ----
__global__ void __launch_bounds__(256) quads_build_kernel(const float* __restrict__ dens, float4* __restrict__ quads, int W, int D, int H,
                                                          int pre, unsigned total) {
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total) return;
  const unsigned HZ = (unsigned)H + 1u, DY = (unsigned)D + 1u;
  const int cz = (int)(idx % HZ), cy = (int)((idx / HZ) % DY), cx = (int)(idx / (HZ * DY));
  const int x = cx - 1, y = cy - 1, z = cz - 1;
  auto val = [&](int xx, int yy, int zz) -> float {
    if ((unsigned)xx >= (unsigned)W || (unsigned)yy >= (unsigned)D || (unsigned)zz >= (unsigned)H) return 0.0f;
    const float v = __ldg(dens + ((size_t)xx * D + yy) * H + zz);
    return pre == R3D_PRE_ABS ? fabsf(v) : v;
  };
  quads[idx] = make_float4(val(x, y, z), val(x, y, z + 1), val(x, y + 1, z), val(x, y + 1, z + 1));
}

extern "C" int64_t r3d_density_quad_floats(const int32_t dims[3]) {
  if (!dims || dims[0] < 1 || dims[1] < 1 || dims[2] < 1) return -1;
  return 4ll * ((int64_t)dims[0] + 2) * ((int64_t)dims[1] + 1) * ((int64_t)dims[2] + 1);
}

extern "C" int r3d_build_density_quads(const R3dGrid* grid, float* quads, void* cuda_stream) {
  GridP g;
  int rc;
  if ((rc = to_device_params(grid, g))) return rc;
  if (!quads || !aligned16(quads)) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_build_density_quads: quads must be a 16-byte aligned buffer");
  const long long total = ((long long)g.W + 2) * (g.D + 1) * (g.H + 1);
  if (total > 0xffffffffLL) return fail(R3D_ERR_UNSUPPORTED, "r3d_build_density_quads: grid too large for 32-bit quad indices");
  const unsigned blocks = (unsigned)((total + 255) / 256);
  quads_build_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(g.dens, reinterpret_cast<float4*>(quads), g.W, g.D, g.H,
                                                                                    g.pre, (unsigned)total);
  return check_launch("r3d_build_density_quads");
}

extern "C" int r3d_cast_rays(const R3dCamera* camera, float* origins, float* directions, void* cuda_stream) {
  if (!camera || !origins || !directions) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_cast_rays: NULL argument");
  if (camera->height < 1 || camera->width < 1 || !(camera->focal > 0.f))
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_cast_rays: bad camera intrinsics");
  const long long n = (long long)camera->height * camera->width;
  cast_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(*camera, origins, directions);
  return check_launch("r3d_cast_rays");
}

extern "C" int r3d_grid_lookup_fwd(const R3dGrid* grid, const float* points, int64_t num_points, float* out, uint8_t* inside,
                                   void* cuda_stream) {
  GridP g;
  int rc;
  if ((rc = to_device_params(grid, g))) return rc;
  if (num_points < 0 || (num_points > 0 && (!points || !out))) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_grid_lookup_fwd: NULL argument");
  if (num_points == 0) return R3D_OK;
  const long long blocks = (num_points * 32 + 255) / 256;
  lookup_fwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(g, points, num_points, out, inside);
  return check_launch("r3d_grid_lookup_fwd");
}

extern "C" int r3d_grid_lookup_bwd(const R3dGrid* grid, const float* points, int64_t num_points, const float* grad_out,
                                   const R3dGridGrad* grad_grid, void* cuda_stream) {
  GridP g;
  int rc;
  if ((rc = to_device_params(grid, g))) return rc;
  if (num_points < 0 || !grad_grid || (num_points > 0 && (!points || !grad_out)))
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_grid_lookup_bwd: NULL argument");
  if (num_points == 0) return R3D_OK;
  const long long blocks = (num_points * 32 + 255) / 256;
  lookup_bwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(g, points, num_points, grad_out,
                                                                                         grad_grid->densities, grad_grid->features);
  return check_launch("r3d_grid_lookup_bwd");
}

extern "C" int r3d_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                             float beta2, float eps, float bias_correction1, float bias_correction2, float grad_scale,
                             void* cuda_stream) {
  if (n < 0 || (n > 0 && (!param || !grad || !exp_avg || !exp_avg_sq))) return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_adam_step: NULL argument");
  if (n == 0) return R3D_OK;
  if (!aligned16(param) || !aligned16(grad) || !aligned16(exp_avg) || !aligned16(exp_avg_sq))
    return fail(R3D_ERR_INVALID_ARGUMENT, "r3d_adam_step: buffers must be 16-byte aligned");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sms * 8;  // persistent grid-stride: a multiple of the SM count
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bias_correction1, sqrtf(bias_correction2), grad_scale);
  return check_launch("r3d_adam_step");
}
