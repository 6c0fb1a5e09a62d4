/*
 * r3d_b200.h -- C ABI of the B200-native SH-voxel-grid volumetric renderer.
 *
 * This is the drop-in boundary for ONE hot path of akanimax/thr3ed_atom: the render procedure
 *     thre3d_atom/thre3d_reprs/renderers.py:48-102   render_sh_voxel_grid(voxel_grid, rays, cfg)
 * (sampler -> point processor -> accumulator, rendering/volumetric/render_interface.py:103-134)
 * and autograd's backward of it into VoxelGrid._densities / ._features.  The reference has no FFI
 * (it is pure Python); these entry points are what a ctypes binding inside the reference's
 * renderers.py would call -- see INTEGRATION.md for that stub.
 *
 * Conventions
 *  - Plain C: raw device pointers + sizes, no torch / C++ types.  All arrays are fp32, contiguous.
 *  - The caller (PyTorch) allocates and owns every buffer, including outputs and gradient buffers.
 *    Gradient buffers are ACCUMULATED into (caller zeroes them), so several ray batches / several
 *    renders can share one buffer exactly like autograd's .grad accumulation.
 *  - Work is enqueued on the given CUDA stream (cudaStream_t passed as void*); no internal
 *    synchronisation, no global mutable state apart from the thread-local error string.
 *  - Every function returns R3D_OK (0) or an error code; r3d_last_error() gives the message.
 *    No C++ exception crosses this boundary.
 */
#ifndef R3D_B200_H_
#define R3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R3D_ABI_VERSION 5

#if defined(__GNUC__)
#define R3D_API __attribute__((visibility("default")))
#else
#define R3D_API
#endif

enum R3dStatus {
  R3D_OK = 0,
  R3D_ERR_INVALID_ARGUMENT = 1,
  R3D_ERR_UNSUPPORTED = 2,
  R3D_ERR_CUDA = 3
};

/* density activations selectable by the reference's train script
 * (thre3d_elements/relu_fields/train_sh_based_voxel_grid_with_posed_images.py:169-192) */
enum R3dDensityPre { R3D_PRE_IDENTITY = 0, R3D_PRE_ABS = 1 };
enum R3dDensityPost { R3D_POST_IDENTITY = 0, R3D_POST_RELU = 1, R3D_POST_SOFTPLUS = 2 };

/* SHVoxGridRenderConfig booleans (thre3d_reprs/renderers.py:28-45) */
enum R3dRenderFlags {
  R3D_FLAG_PERTURB = 1u << 0,            /* perturb_sampled_points */
  R3D_FLAG_WHITE_BKGD = 1u << 1,         /* white_bkgd */
  R3D_FLAG_DIFFUSE = 1u << 2,            /* render_diffuse: SH band 0 only (process.py:59-63) */
  R3D_FLAG_OPTIMIZED_SAMPLING = 1u << 3  /* optimized_sampling: per-ray near/far from the slab test (sample.py:71-202) */
};

/* A dense voxel grid = reference VoxelGrid (thre3d_reprs/voxels.py:46-124).
 * Index order V[ix][iy][iz][channel], channel fastest -- the reference's own layout (voxels.py:70-71).
 * `features` may be padded: `feature_stride` floats per voxel record, of which the first
 * `num_features` = 3*(sh_degree+1)^2 are used (channel-major: coeff[ch][k] = rec[ch*K + k],
 * process.py:61,66).  A stride that is a multiple of 4 with a 16-byte aligned base enables 128-bit
 * vector loads / reductions; the unpadded reference layout (stride == num_features) also works. */
typedef struct R3dGrid {
  const float* densities;  /* [W][D][H]          raw (pre-activation, un-scaled) density */
  const float* features;   /* [W][D][H][feature_stride] */
  int32_t dims[3];         /* W (x), D (y), H (z)  voxels.py:116-121 */
  int32_t sh_degree;       /* 0..3 (spherical_harmonics.py:79) */
  int32_t num_features;    /* 3*(sh_degree+1)^2 */
  int32_t feature_stride;  /* >= num_features */
  float aabb_min[3];       /* voxels.py:187-212, rounded to fp32 (what the fp32 comparisons at :262-272 see) */
  float aabb_max[3];
  float norm_scale[3];     /* n = p*scale + bias maps the AABB to [-1,1]; fp32 values of */
  float norm_bias[3];      /*   utils/imaging_utils.py:58-63 (adjust_dynamic_range, slack=True) */
  float density_scale;     /* expected_density_scale (voxels.py:63,292-294) */
  int32_t density_pre;     /* R3dDensityPre  */
  int32_t density_post;    /* R3dDensityPost */
  /* Optional derived buffer (NULL = not used): the density quad volume, r3d_density_quad_floats(dims) floats, 16-byte
   * aligned, filled from `densities` by r3d_build_density_quads.  The caller owns it and must rebuild it whenever the
   * densities changed.  With it the forward kernel probes the 8 corner densities of a sample with two 16-byte loads. */
  const float* density_quads;
} R3dGrid;

/* Pinhole camera for in-kernel ray generation = cast_rays (rendering/volumetric/utils/misc.py:12-50). */
typedef struct R3dCamera {
  int32_t height, width;
  float focal;
  float rotation[9];     /* row-major 3x3 camera-to-world */
  float translation[3];
} R3dCamera;

/* A flat batch of rays = reference Rays (render_interface.py:14-44), always [N,3] (:127-129). */
typedef struct R3dRays {
  const float* origins;     /* [N][3]; may be NULL when `camera` is given (rays are generated in-kernel) */
  const float* directions;  /* [N][3]; need not be unit length */
  const float* bounds;      /* optional [N][2] per-ray (near, far) (sample.py:42-43); NULL = cfg near/far */
  const R3dCamera* camera;  /* optional HOST pointer; N must equal height*width, ray r = y*width + x */
  int64_t num_rays;
  int32_t tile_width;       /* optional coherence hint: rays are a row-major image of this width (0 = unknown). */
  int32_t tile_height;      /*   Threads are then mapped to 8x4 pixel tiles; results are identical either way. */
} R3dRays;

typedef struct R3dRenderConfig {
  int32_t num_samples;   /* num_samples_per_ray */
  float near, far;       /* camera_bounds */
  uint32_t flags;        /* R3dRenderFlags */
  const float* jitter;   /* optional [N][S] U[0,1) stratified offsets (what the reference draws with torch.rand,
                            sample.py:63).  NULL with R3D_FLAG_PERTURB => counter-based in-kernel RNG below. */
  uint64_t rng_seed;     /* in-kernel jitter = hash(seed, ray, sample); the backward pass re-derives it */
  int32_t variant;       /* kernel variant selector for A/B measurement; 0 = default */
} R3dRenderConfig;

/* = reference RenderOut (render_interface.py:47-83) with extra{disparity, accumulated_weight}. */
typedef struct R3dRenderOut {
  float* colour;     /* [N][3] */
  float* depth;      /* [N]    (== [N][1]) */
  float* acc;        /* [N]    accumulated_weight */
  float* disparity;  /* [N]    may be NULL */
  float* sample_cache; /* optional [S][N][4] fp32, 16-byte aligned.  Forward: when non-NULL, (sigmoid(raw) rgb, sigma) of every
                          sample with sigma != 0 is stored at [sample][ray] (other entries are left untouched).  Backward
                          (`saved`): when non-NULL the per-sample radiance is read back instead of being re-gathered from
                          the grid.  Trades 16*N*S bytes of HBM for the second 8-corner gather. */
  /* Single-pass specular + diffuse render.  The reference trainer renders every batch twice, once with
   * all SH bands and once with band 0 only (modules/trainers.py:306-330, process.py:59-63); the diffuse radiance is the
   * k = 0 slice of the very records the specular render interpolates, so one gather serves both.  When `colour_diffuse`
   * is non-NULL (R3D_FLAG_DIFFUSE must be clear) the forward also writes the band-0 image of the SAME samples there;
   * depth, acc and disparity are common to both (they do not depend on the radiance). */
  float* colour_diffuse;        /* [N][3], optional */
  float* sample_cache_diffuse;  /* optional [S][N][4] like sample_cache: (sigmoid(raw_diffuse) rgb, -) per sample */
  /* Optional [S][r3d_sample_mask_words(rays)] uint32, with sample_cache.  Forward: word [sample][warp] = ballot of the warp's
   * rays whose sample contributed (sigma != 0); written by the default (lane-group) forward kernel only -- r3d_render_fwd
   * fails with R3D_ERR_UNSUPPORTED if another kernel would be dispatched.  Backward (`saved`): with a ReLU density
   * post-activation the march then needs neither the inside test nor the 8-corner density gather (sigma comes from
   * sample_cache), and marching steps in which no ray of a warp contributed are skipped outright.  The words are indexed
   * by the launch's warps: the backward call must describe the rays exactly as the forward call did (same num_rays and
   * tile hint), as it must anyway for the saved outputs to belong to it. */
  uint32_t* sample_mask;
} R3dRenderOut;

/* upstream gradients dL/d(output); any pointer may be NULL (= zero). */
typedef struct R3dRenderOutGrad {
  const float* colour;     /* [N][3] */
  const float* depth;      /* [N] */
  const float* acc;        /* [N] */
  const float* disparity;  /* [N] */
  const float* colour_diffuse;  /* [N][3]: dL/d(colour_diffuse) of a single-pass specular + diffuse render */
} R3dRenderOutGrad;

/* gradient buffers, same layout as R3dGrid.densities / .features; accumulated into. */
typedef struct R3dGridGrad {
  float* densities;  /* [W][D][H]; may be NULL */
  float* features;   /* [W][D][H][feature_stride]; may be NULL */
} R3dGridGrad;

R3D_API int r3d_abi_version(void);
/* 1 when the library was built with -DR3D_AB_VARIANTS (R3dRenderConfig.variant != 0 selects measurement-only kernels),
 * 0 for the product build, which carries only the kernels the dispatch uses and refuses a non-zero variant. */
R3D_API int r3d_has_ab_variants(void);
/* words per sample of R3dRenderOut.sample_mask for this ray batch (= warps of the render launch) */
R3D_API int64_t r3d_sample_mask_words(const R3dRays* rays);
R3D_API const char* r3d_last_error(void);

/* Forward: replaces render_sh_voxel_grid (thre3d_reprs/renderers.py:48-102) =
 *   sample_uniform_points_on_rays / sample_aabb_bound_uniform_points_on_rays (sample.py:15-202)
 *   -> process_points_with_sh_voxel_grid (process.py:20-96) incl. VoxelGrid.forward (voxels.py:276-331),
 *      test_inside_volume (voxels.py:252-274), evaluate_spherical_harmonics (spherical_harmonics.py:64-116)
 *   -> accumulate_radiance_density_on_rays (accumulate.py:31-113)
 * in one fused kernel.  density2occupancy is density2occupancy_pb (accumulate.py:24-28), the tone map is
 * torch.sigmoid, stochastic_density_noise_std is 0 (the defaults of renderers.py:37-39). */
R3D_API int r3d_render_fwd(const R3dGrid* grid, const R3dRays* rays, const R3dRenderConfig* cfg,
                   const R3dRenderOut* out, void* cuda_stream);

/* Backward: replaces autograd's backward of the above into _densities/_features
 * (the graph behind trainers.py:339-341).  `saved` holds the forward outputs of the same call. */
R3D_API int r3d_render_bwd(const R3dGrid* grid, const R3dRays* rays, const R3dRenderConfig* cfg,
                   const R3dRenderOut* saved, const R3dRenderOutGrad* grad_out,
                   const R3dGridGrad* grad_grid, void* cuda_stream);

/* cast_rays (rendering/volumetric/utils/misc.py:12-50): fills origins/directions [H*W][3]. */
R3D_API int r3d_cast_rays(const R3dCamera* camera, float* origins, float* directions, void* cuda_stream);

/* A set of posed pinhole views sharing one intrinsics (what the reference trainer keeps cached on the device,
 * modules/trainers.py:59, data/datasets.py:74-89). */
typedef struct R3dViewSet {
  const float* rotations;     /* [V][9] row-major camera-to-world rotations (device) */
  const float* translations;  /* [V][3] (device) */
  const float* images;        /* optional [V][H][W][3] target pixels (device, channel-last) */
  int32_t num_views, height, width;
  float focal;
} R3dViewSet;

/* Training-batch sampler: `batch` rays (+ their target pixels) drawn uniformly with replacement from all pixels of all
 * views, in units of tile_width x tile_height pixel tiles (1 x 1 = independent pixels; 8 x 4 = one coherent tile per warp of
 * the render kernels).  Replaces cast_rays on every cached view + randperm over all pixels + gathers, every iteration
 * (modules/trainers.py:281-303, rendering/volumetric/utils/misc.py:117-129).  Rays are bit-identical to r3d_cast_rays for
 * the same pixel.  Outputs: origins / directions [batch][3], pixels [batch][3] (optional), indices [batch] = flat pixel
 * index (view*H + y)*W + x (optional).  Counter-based RNG keyed by `seed`: the same seed gives the same batch. */
R3D_API int r3d_sample_ray_batch(const R3dViewSet* views, int64_t batch, int32_t tile_width, int32_t tile_height, uint64_t seed,
                                 float* origins, float* directions, float* pixels, int64_t* indices, void* cuda_stream);

/* VoxelGrid.forward (voxels.py:276-331) on free points: out [P][F+1] = (features..., density);
 * `inside` (optional, [P] bytes) = test_inside_volume (voxels.py:252-274). */
R3D_API int r3d_grid_lookup_fwd(const R3dGrid* grid, const float* points, int64_t num_points, float* out,
                        uint8_t* inside, void* cuda_stream);
R3D_API int r3d_grid_lookup_bwd(const R3dGrid* grid, const float* points, int64_t num_points,
                        const float* grad_out, const R3dGridGrad* grad_grid, void* cuda_stream);

/* Density quad volume (see R3dGrid.density_quads): entry (cx, cy, cz) of a [W+2][D+1][H+1] array of float4 holds the
 * pre-activated, un-scaled densities (v[x][y][z], v[x][y][z+1], v[x][y+1][z], v[x][y+1][z+1]) of x-plane x = cx-1 of the
 * interpolation cell with low corner (cx-1, cy-1, cz-1); out-of-range voxels are 0 (grid_sample's zero padding,
 * voxels.py:296-303).  r3d_build_density_quads fills `quads` (writable alias of grid->density_quads) from grid->densities. */
R3D_API int64_t r3d_density_quad_floats(const int32_t dims[3]);
R3D_API int r3d_build_density_quads(const R3dGrid* grid, float* quads, void* cuda_stream);

/* Measurement helper (not on the product path): marks every voxel that the batch's in-volume
 * samples reference as an interpolation corner in `bitmap` ([W*D*H] bytes, caller-zeroed), so the
 * host can count U = unique voxels touched for the algorithmic-bytes figure. */
R3D_API int r3d_mark_touched_voxels(const R3dGrid* grid, const R3dRays* rays, const R3dRenderConfig* cfg,
                            uint8_t* bitmap, void* cuda_stream);

/* Measurement helper (not on the product path): sample statistics of a batch for the roofline companion figures
 * (SURVEY.md 8d): counters[0] samples visited by the march, [1] samples strictly inside the AABB (voxels.py:252-274),
 * [2] in-range trilinear corner references of those (what the reference's two grid_sample calls request,
 * voxels.py:296-318), [3] contributing samples (sigma != 0).  `counters` = 4 device uint64, caller-zeroed, accumulated. */
R3D_API int r3d_sample_statistics(const R3dGrid* grid, const R3dRays* rays, const R3dRenderConfig* cfg, uint64_t* counters,
                                  void* cuda_stream);

/* Fused dense Adam on a grid tensor (next-row f1; replaces torch.optim.Adam at trainers.py:242-245,341).
 * p, g, m, v: [n] fp32.  Standard Adam (no weight decay, no amsgrad): bias corrections are passed in
 * as 1-beta^t so the kernel stays stateless. */
R3D_API int r3d_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, float bias_correction1,
                  float bias_correction2, float grad_scale, void* cuda_stream);

/* In-switch (NVLS) sum all-reduce of a gradient buffer over NVLink/NVSwitch: the exchange between backward() and
 * optimizer.step() (reference modules/trainers.py:339-341; the reference itself is single-device).  `multicast_ptr` is the
 * multicast address of a buffer allocated symmetrically on all ranks and bound to one multicast object; rank r reduces and
 * re-broadcasts slice r (multimem.ld_reduce / multimem.st).  The caller brackets the call with cross-rank barriers on the
 * same stream.  num_blocks <= 0 picks 2 CTAs per SM. */
R3D_API int r3d_multimem_all_reduce(void* multicast_ptr, int64_t num_floats, int32_t rank, int32_t world_size,
                                    int32_t num_blocks, void* cuda_stream);

/* Fused exchange + optimizer over NVLink/NVSwitch: reduce-scatter -> shard-local Adam -> all-gather in one kernel
 * (replaces the gradient all-reduce followed by optimizer.step() of reference modules/trainers.py:339-341 with the Adam of
 * :242-245).  `grad_multicast_ptr` / `param_multicast_ptr`: multicast addresses of the flat gradient / parameter buffers,
 * each allocated symmetrically on all ranks and bound to one multicast object; `param_local`: this rank's own replica of the
 * parameters.  Rank r owns slice r of r3d_multimem_shard_floats(num_floats, world_size) floats (the last slice may be
 * shorter): it pulls the summed gradient of the slice (multimem.ld_reduce), updates the slice with its shard of the
 * optimizer state (`exp_avg_shard`, `exp_avg_sq_shard`: that many floats, kept by the caller between steps) and broadcasts
 * the new parameters to every replica (multimem.st).  Adam semantics and bias corrections as in r3d_adam_step.  The caller
 * brackets the call with cross-rank barriers on the same stream.  num_blocks <= 0 picks 2 CTAs per SM. */
R3D_API int64_t r3d_multimem_shard_floats(int64_t num_floats, int32_t world_size);
R3D_API int r3d_multimem_adam_step(void* grad_multicast_ptr, void* param_multicast_ptr, const float* param_local, float* exp_avg_shard,
                                   float* exp_avg_sq_shard, int64_t num_floats, int32_t rank, int32_t world_size, float lr,
                                   float beta1, float beta2, float eps, float bias_correction1, float bias_correction2,
                                   float grad_scale, int32_t num_blocks, void* cuda_stream);

/* The same fused exchange + optimizer over peer-to-peer loads / stores, no multicast object needed: `grad_ptrs[r]` /
 * `param_ptrs[r]` (host arrays of world_size device pointers, 1 <= world_size <= 8) are rank r's flat gradient / parameter
 * replica as mapped into THIS process (symmetric-memory peer pointers; entry `rank` is the local buffer).  Rank `rank` sums
 * the gradient of its slice over the replicas in rank order, applies Adam with its state shard and writes the new parameters
 * into every replica.  Moves fewer NVLink bytes than the in-switch version at world_size == 2 (1.0x vs 1.5x the gradient
 * bytes per direction), more from 4 ranks on.  Same slicing, Adam semantics and barrier contract as r3d_multimem_adam_step. */
R3D_API int r3d_peer_adam_step(const void* const* grad_ptrs, void* const* param_ptrs, float* exp_avg_shard, float* exp_avg_sq_shard,
                               int64_t num_floats, int32_t rank, int32_t world_size, float lr, float beta1, float beta2, float eps,
                               float bias_correction1, float bias_correction2, float grad_scale, int32_t num_blocks,
                               void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* R3D_B200_H_ */
