// r3d_api.cu -- ABI bookkeeping: version, error string, argument validation.
#include <cstdio>
#include <cstring>

#include "r3d_host.h"

namespace r3d {

static thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(R3D_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return R3D_OK;
}

int to_device_params(const R3dGrid* grid, GridP& g) {
  if (!grid) return fail(R3D_ERR_INVALID_ARGUMENT, "grid is NULL");
  if (!grid->densities || !grid->features) return fail(R3D_ERR_INVALID_ARGUMENT, "grid.densities / grid.features is NULL");
  for (int a = 0; a < 3; ++a) {
    if (grid->dims[a] < 1) return fail(R3D_ERR_INVALID_ARGUMENT, "grid.dims[%d] = %d must be >= 1", a, grid->dims[a]);
    if (!(grid->aabb_max[a] > grid->aabb_min[a])) return fail(R3D_ERR_INVALID_ARGUMENT, "grid AABB is empty on axis %d", a);
  }
  if ((long long)grid->dims[0] * grid->dims[1] * grid->dims[2] > 0x7fffffffLL)
    return fail(R3D_ERR_UNSUPPORTED, "grids with more than 2^31-1 voxels are not supported");
  // only degrees 0..3 exist in the reference (spherical_harmonics.py:79)
  if (grid->sh_degree < 0 || grid->sh_degree > 3)
    return fail(R3D_ERR_UNSUPPORTED, "only degrees 0, 1, 2, and 3 are supported (got %d)", grid->sh_degree);
  const int K = (grid->sh_degree + 1) * (grid->sh_degree + 1);
  if (grid->num_features != 3 * K)
    return fail(R3D_ERR_INVALID_ARGUMENT, "number of features (%d) does not match 3*(deg+1)^2 = %d for degree %d",
                grid->num_features, 3 * K, grid->sh_degree);
  if (grid->feature_stride < grid->num_features)
    return fail(R3D_ERR_INVALID_ARGUMENT, "feature_stride (%d) < num_features (%d)", grid->feature_stride, grid->num_features);
  if (grid->density_pre != R3D_PRE_IDENTITY && grid->density_pre != R3D_PRE_ABS)
    return fail(R3D_ERR_UNSUPPORTED, "unknown density pre-activation %d", grid->density_pre);
  if (grid->density_post < R3D_POST_IDENTITY || grid->density_post > R3D_POST_SOFTPLUS)
    return fail(R3D_ERR_UNSUPPORTED, "unknown density post-activation %d", grid->density_post);
  g.dens = grid->densities;
  g.feat = grid->features;
  g.quads = reinterpret_cast<const float4*>(grid->density_quads);
  if (g.quads && !aligned16(g.quads)) return fail(R3D_ERR_INVALID_ARGUMENT, "grid.density_quads must be 16-byte aligned");
  g.W = grid->dims[0], g.D = grid->dims[1], g.H = grid->dims[2];
  g.F = grid->num_features, g.stride = grid->feature_stride, g.K = K;
  for (int a = 0; a < 3; ++a) {
    g.lo[a] = grid->aabb_min[a], g.hi[a] = grid->aabb_max[a];
    g.ns[a] = grid->norm_scale[a], g.nb[a] = grid->norm_bias[a];
  }
  g.dscale = grid->density_scale;
  g.pre = grid->density_pre, g.post = grid->density_post;
  return R3D_OK;
}

int to_device_params(const R3dRays* rays, RaysP& r) {
  if (!rays) return fail(R3D_ERR_INVALID_ARGUMENT, "rays is NULL");
  if (rays->num_rays < 0) return fail(R3D_ERR_INVALID_ARGUMENT, "num_rays < 0");
  r.origins = rays->origins, r.directions = rays->directions, r.bounds = rays->bounds;
  r.n = rays->num_rays;
  r.tile_w = rays->tile_width, r.tile_h = rays->tile_height;
  r.has_camera = 0;
  memset(&r.cam, 0, sizeof(r.cam));
  if (rays->camera) {
    const R3dCamera& c = *rays->camera;
    if (c.height < 1 || c.width < 1 || !(c.focal > 0.f)) return fail(R3D_ERR_INVALID_ARGUMENT, "bad camera intrinsics");
    if ((long long)c.height * c.width != rays->num_rays)
      return fail(R3D_ERR_INVALID_ARGUMENT, "camera %dx%d does not match num_rays %lld", c.height, c.width, (long long)rays->num_rays);
    r.cam = c;
    r.has_camera = 1;
    r.tile_w = c.width, r.tile_h = c.height;
  } else if (rays->num_rays > 0 && (!rays->origins || !rays->directions)) {  // empty batches may carry NULL
    return fail(R3D_ERR_INVALID_ARGUMENT, "rays.origins / rays.directions is NULL and no camera was given");
  }
  if (r.tile_w > 0) {
    if (r.tile_h < 1 || (long long)r.tile_w * r.tile_h != r.n)
      return fail(R3D_ERR_INVALID_ARGUMENT, "tile hint %dx%d does not match num_rays %lld", r.tile_w, r.tile_h, (long long)r.n);
  } else {
    r.tile_w = r.tile_h = 0;
  }
  return R3D_OK;
}

int to_device_params(const R3dRenderConfig* cfg, const RaysP& r, CfgP& c) {
  (void)r;
  if (!cfg) return fail(R3D_ERR_INVALID_ARGUMENT, "render config is NULL");
  if (cfg->num_samples < 1) return fail(R3D_ERR_INVALID_ARGUMENT, "num_samples_per_ray must be >= 1 (got %d)", cfg->num_samples);
  c.S = cfg->num_samples;
  c.near = cfg->near, c.far = cfg->far;
  c.flags = cfg->flags;
  c.jitter = (cfg->flags & R3D_FLAG_PERTURB) ? cfg->jitter : nullptr;
  c.seed_lo = (unsigned)(cfg->rng_seed & 0xffffffffu);
  c.seed_hi = (unsigned)(cfg->rng_seed >> 32);
  return R3D_OK;
}

}  // namespace r3d

extern "C" {

int r3d_abi_version(void) { return R3D_ABI_VERSION; }

const char* r3d_last_error(void) { return r3d::g_error; }

}  // extern "C"
