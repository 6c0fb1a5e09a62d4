// r3d_bwd_ws.cuh -- backward render, warp-specialised producer / consumer kernel (sm_100a).  Included by r3d_render.cu.
//
// Covers the default training configuration: ReLU density post-activation with the forward's per-sample records
// (R3dRenderOut.sample_cache) and contribution ballots (sample_mask); every other configuration keeps
// render_bwd_coop_kernel.  Same maths as that kernel; what changes is who does the scatter and how:
//
//   producer warp  (4 per CTA, one thread per ray of an 8x4 pixel tile)
//       marches by the forward's ballots: depth (per-CTA stratum table), position, cell, alpha / T / q / suffix from the
//       cached (sigmoid(raw), sigma) -> dL/dsigma_pre and dL/draw; groups the warp's contributing samples by interpolation
//       cell (__match_any_sync + a warp scan) and publishes them cell by cell into a stage: 8 trilinear weights, 8 corner
//       voxels, (dL/draw rgb, dL/dsigma_pre), owning ray.
//   consumer warp  (4 per CTA, paired 1:1 with a producer)
//       lane groups of LPR lanes (one float4 of a voxel record each) take contiguous chunks of a stage.  A group keeps the
//       gradient of ONE cell -- 8 records, i.e. 8 float4 per lane -- in registers, adds w[corner] * dL/draw[ch] * Y[k] for
//       every sample of the run (packed FFMA2) and flushes the cell with 8 red.global.add.v4.f32 per lane when the cell
//       changes.  The per-sample product row (28 floats) of the cooperative kernel is never materialised: the group
//       forms it from 4 published floats and the ray's SH row.
//
// render_bwd_coop_kernel spends 49 % of the SM's L1/shared-memory data pipe on shared-memory traffic (603 M wavefronts per
// launch at c3: the published product rows and the member sweep that re-reads them for every corner pass) and 34 % of its
// instructions on that sweep; here a contributing sample costs ~2.5 shared-memory wavefronts and ~20 instructions.
#pragma once

#ifndef R3D_WSB_BLOCKS
#define R3D_WSB_BLOCKS 3
#endif

namespace r3d {

constexpr int kWbStages = 3;

template <bool DUAL>
struct alignas(16) WbStage {
  float4 Wlo[32], Whi[32];   // trilinear weights of corners 0..3 / 4..7, per published sample (cell-sorted slot order)
  uint4 Vlo[32], Vhi[32];    // corner VOXEL indices (record offset = voxel * stride floats: 64-bit at 512^3 degree 3)
  float4 G[32];              // (dL/draw r, g, b, dL/dsigma_pre)
  float4 G2[DUAL ? 32 : 1];  // dL/draw_diffuse (r, g, b, -): lands on the k = 0 coefficients only
  unsigned yoff[32];         // byte offset of the owning ray's row in the pair's SH table
  int n, pad0, pad1, pad2;   // samples in this stage (-1: the producer is done)
};

template <int DEG, bool DUAL>
struct alignas(16) WbPair {
  using H = FwdGroupShape<DEG>;
  WbStage<DUAL> st[kWbStages];
  float Y[32 * H::YROW];
  unsigned long long full[kWbStages], done[kWbStages];
};

template <int DEG, bool DUAL>
__host__ __device__ constexpr size_t wsb_smem_bytes(int S, bool table) {
  return 4 * sizeof(WbPair<DEG, DUAL>) + (table ? sizeof(float2) * (size_t)S : 0);
}

template <int DEG, bool DUAL>
__global__ void __launch_bounds__(256, DEG >= 3 ? 2 : R3D_WSB_BLOCKS)
    render_bwd_ws_kernel(const GridP g, const RaysP rp, const CfgP c, const BwdP b, const int use_tab) {
  using H = FwdGroupShape<DEG>;
  using S = CoopShape<DEG>;
  constexpr int K = S::K, F = S::F, NV = S::NV, LPR = H::LPR, MPI = H::MPI, NS = kWbStages;
  constexpr int CPL = LPR >= 8 ? 1 : 8 / LPR;  // density corners per lane of a group
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char ws_smem[];
  WbPair<DEG, DUAL>* pairs = reinterpret_cast<WbPair<DEG, DUAL>*>(ws_smem);
  float2* ztab = reinterpret_cast<float2*>(ws_smem + 4 * sizeof(WbPair<DEG, DUAL>));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WbPair<DEG, DUAL>& sm = pairs[warp & 3];
  const bool producer = warp < 4;

  if (producer && lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&sm.full[s], 1), mbar_init(&sm.done[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (use_tab) {
    DepthGen dg;
    dg.near = c.near, dg.far = c.far, dg.S = c.S, dg.half = c.S / 2;
    dg.step = c.S > 1 ? __fdiv_rn(1.0f, (float)(c.S - 1)) : 0.0f;
    const bool perturb = (c.flags & R3D_FLAG_PERTURB) != 0;
    for (int j = threadIdx.x; j < c.S; j += 256) {
      const float bb = dg.base(j);
      float lower = bb, span = 0.0f;
      if (perturb) {
        lower = (j > 0) ? 0.5f * __fadd_rn(bb, dg.base(j - 1)) : bb;
        const float upper = (j < c.S - 1) ? 0.5f * __fadd_rn(dg.base(j + 1), bb) : bb;
        span = __fsub_rn(upper, lower);
      }
      ztab[j] = make_float2(lower, span);
    }
  }
  __syncthreads();

  if (!producer) {
    // =========================================================================== consumer: run-merged scatter
    const int ms = lane / LPR, cj = lane % LPR;
    const bool role_ok = cj < NV;
    const unsigned y_lane = (unsigned)__cvta_generic_to_shared(sm.Y) + 16u * (unsigned)cj;
    const unsigned ustride = (unsigned)g.stride;
    // channel of this lane's four record elements: (lo, lo|hi, hi, hi) -- at most one channel boundary per float4
    const int e0 = 4 * cj;
    const int ch_lo = min(e0 / K, 2), ch_hi = min((e0 + 3) / K, 2), ch_1 = min((e0 + 1) / K, 2), ch_2 = min((e0 + 2) / K, 2);
    // DUAL: component of this lane's float4 that is a k = 0 coefficient (element ch * K), if any
    int dslot = -1, dcomp = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (DUAL && cj == (ch * K) / 4) dslot = ch, dcomp = (ch * K) % 4;

    float2 a01[8], a23[8];  // gradient of the cell in flight: this lane's float4 of its 8 corner records
    float ad[CPL];          // density gradient of corners cj + t * LPR
    unsigned vox[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a01[k] = a23[k] = make_float2(0.f, 0.f), vox[k] = 0u;
#pragma unroll
    for (int t = 0; t < CPL; ++t) ad[t] = 0.f;
    bool have = false;  // a cell is in flight
    unsigned key0 = 0xffffffffu, key7 = 0xffffffffu;

    auto flush = [&]() {
      if (b.gfeat && role_ok) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float* dst = b.gfeat + (size_t)vox[k] * (size_t)ustride + 4 * cj;
          red_add_v4(dst, a01[k].x, a01[k].y, a23[k].x, a23[k].y);
        }
      }
      if (b.gdens) {
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
          const int corner = cj + t * LPR;
          if (corner < 8 && ad[t] != 0.f) {
            unsigned v = vox[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) v = (corner == k) ? vox[k] : v;
            float gv = ad[t];
            if (g.pre == R3D_PRE_ABS) {
              const float dv = __ldg(g.dens + v);
              gv = (dv > 0.f) ? gv : ((dv < 0.f) ? -gv : 0.0f);  // d|x|/dx = sign(x), 0 at 0 (torch.abs)
            }
            atomicAdd(b.gdens + v, gv);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) a01[k] = a23[k] = make_float2(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < CPL; ++t) ad[t] = 0.f;
    };

    int gs = 0;
    unsigned gpar = 0u;
    while (true) {
      WbStage<DUAL>& st = sm.st[gs];
      mbar_wait_bounded(&sm.full[gs], gpar);
      const int n = st.n;
      if (n < 0) break;
      const int chunk = (n + MPI - 1) / MPI;
      const int m_end = min(n, (ms + 1) * chunk);
      for (int it = 0; it < chunk; ++it) {
        const int m = ms * chunk + it;
        if (m < m_end) {  // whole groups take this branch together
          const uint4 v0 = st.Vlo[m], v1 = st.Vhi[m];
          if (v0.x != key0 || v1.w != key7) {  // corners 0 and 7 identify the cell
            if (have) flush();
            have = true, key0 = v0.x, key7 = v1.w;
            vox[0] = v0.x, vox[1] = v0.y, vox[2] = v0.z, vox[3] = v0.w, vox[4] = v1.x, vox[5] = v1.y, vox[6] = v1.z, vox[7] = v1.w;
          }
          const float4 w0 = st.Wlo[m], w1 = st.Whi[m];
          const float4 gr = st.G[m];
          const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          if (role_ok) {
            float4 y4;
            lds_v4(y_lane + st.yoff[m], y4.x, y4.y, y4.z, y4.w);
            const float d_lo = ch_lo == 0 ? gr.x : (ch_lo == 1 ? gr.y : gr.z), d_hi = ch_hi == 0 ? gr.x : (ch_hi == 1 ? gr.y : gr.z);
            float d_1, d_2;
            if constexpr (DEG == 0) {  // (r, g, b, pad): three channels in one float4
              d_1 = gr.y, d_2 = gr.z;
            } else {  // K >= 4: at most one channel boundary inside a float4
              d_1 = ch_1 == ch_lo ? d_lo : d_hi, d_2 = ch_2 == ch_lo ? d_lo : d_hi;
            }
            float2 p01 = make_float2(d_lo * y4.x, d_1 * y4.y), p23 = make_float2(d_2 * y4.z, d_hi * y4.w);  // the pad element has Y = 0
            if constexpr (DUAL) {
              if (dslot >= 0) {
                const float4 g2 = st.G2[m];
                const float d0 = (dslot == 0 ? g2.x : (dslot == 1 ? g2.y : g2.z)) * 0.28209479177387814f;  // Y[0] = C0 for every ray
                if (dcomp == 0) p01.x += d0;
                if (dcomp == 1) p01.y += d0;
                if (dcomp == 2) p23.x += d0;
                if (dcomp == 3) p23.y += d0;
              }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              a01[k] = ffma2(p01, wk[k], a01[k]);
              a23[k] = ffma2(p23, wk[k], a23[k]);
            }
          }
#pragma unroll
          for (int t = 0; t < CPL; ++t) {
            const int corner = cj + t * LPR;
            float wcn = wk[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) wcn = (corner == k) ? wk[k] : wcn;
            if (corner < 8) ad[t] = fmaf(wcn, gr.w, ad[t]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.done[gs]);
      if (++gs == NS) gs = 0, gpar ^= 1u;
    }
    if (have) flush();
    return;
  }

  // ============================================================================= producer: march by the forward's ballots
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  RayGrad rgd;
  bool alive = (ray >= 0) && load_ray_grad(b, c, ray, rgd);
  RayCtx s;
  float qmax = 0.f;
  s.i_lo = 1, s.i_hi = 0;
  {
    float Y[K];
#pragma unroll
    for (int k = 0; k < K; ++k) Y[k] = 0.f;
    if (alive) {
      float vx, vy, vz;
      setup_ray(g, rp, c, ray, s, vx, vy, vz);
      sh_basis<DEG>(vx, vy, vz, Y);
      qmax = fabsf(rgd.gc[0]) + fabsf(rgd.gc[1]) + fabsf(rgd.gc[2]) + fabsf(rgd.gd) * fmaxf(fabsf(s.dg.near), fabsf(s.dg.far)) + fabsf(rgd.ga);
      if constexpr (DUAL) qmax += fabsf(rgd.gcd[0]) + fabsf(rgd.gcd[1]) + fabsf(rgd.gcd[2]);
      alive = s.i_lo <= s.i_hi;
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
  const float dmul = g.dscale;  // identity pre-activation or |.|: d pre / d raw is applied at the flush (sign), the scale here
  const float dscale_abs = (g.pre == R3D_PRE_ABS) ? fabsf(dmul) : dmul;
  const Ray& r = s.r;
  int lo = alive ? s.i_lo : 0x7fffffff, hi = alive ? s.i_hi : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(FULL, lo, o));
    hi = max(hi, __shfl_xor_sync(FULL, hi, o));
  }
  const DepthTab dt{ztab, s.dg.perturb, s.dg.jit, s.dg.key};
  const unsigned mask_stride = gridDim.x * 4u, mask_warp = blockIdx.x * 4u + warp;
  unsigned fmask_next = 0u;
  float4 cv_next = make_float4(0.f, 0.f, 0.f, 0.f), cd_next = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lo <= hi) {
    fmask_next = __ldg(b.mask + (size_t)lo * mask_stride + mask_warp);
    if (alive && ((fmask_next >> lane) & 1u)) {
      cv_next = __ldg(b.cache + (size_t)lo * rp.n + ray);
      if constexpr (DUAL) cd_next = __ldg(b.cache_diffuse + (size_t)lo * rp.n + ray);
    }
  }

  float T = 1.0f, prefix = 0.f;
  int ps = 0, published = 0;
  unsigned ppar = 0u;  // parity of the done[] phase that frees stage ps (valid once published >= NS)
  DepthMarch dm;
  dm.bm = dm.bc = 0.f;
  for (int i = lo; i <= hi; ++i) {
    const unsigned fmask = fmask_next;
    const float4 cv = cv_next, cd = cd_next;
    if (i < hi) {
      fmask_next = __ldg(b.mask + (size_t)(i + 1) * mask_stride + mask_warp);
      // request the next step's per-sample record now: its HBM latency hides behind this step
      if (alive && ((fmask_next >> lane) & 1u)) {
        cv_next = __ldg(b.cache + (size_t)(i + 1) * rp.n + ray);
        if constexpr (DUAL) cd_next = __ldg(b.cache_diffuse + (size_t)(i + 1) * rp.n + ray);
      }
    }
    if (fmask == 0u) continue;  // no ray of this warp contributed at this step
    bool contributes = false;
    float draw[3] = {0.f, 0.f, 0.f}, draw0[3] = {0.f, 0.f, 0.f}, dpre = 0.f;
    Cell cell;
    if (alive && ((fmask >> lane) & 1u)) {  // the ballot bit implies i in [i_lo, i_hi] and a point strictly inside the AABB
      const bool last = (i == c.S - 1);
      float z, zn;
      if (use_tab) {
        z = dt.at(i);
        zn = last ? 0.0f : dt.at(i + 1);
      } else {
        dm.start(s.dg, i), z = dm.next(s.dg, i);
        zn = last ? 0.0f : dm.next(s.dg, i + 1);
      }
      const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
      const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
      const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
      make_cell_inside(g, px, py, pz, cell);
      const float sigma = cv.w;  // ReLU with sigma != 0: d sigma / d pre = 1
      const float delta = last ? __fmul_rn(kInfinity, s.dnorm) : __fmul_rn(__fsub_rn(zn, z), s.dnorm);
      const float alpha = 1.0f - exp_neg(sigma * delta);
      const float w = alpha * T;
      const float Tn = T * (1.0f - alpha);
      const float sr = cv.x, sg = cv.y, sb = cv.z;
      float q = fmaf(rgd.gc[0], sr, fmaf(rgd.gc[1], sg, fmaf(rgd.gc[2], sb, fmaf(rgd.gd, z, rgd.ga))));
      if constexpr (DUAL) q = fmaf(rgd.gcd[0], cd.x, fmaf(rgd.gcd[1], cd.y, fmaf(rgd.gcd[2], cd.z, q)));
      prefix = fmaf(w, q, prefix);
      float suffix = 0.0f;  // see render_bwd_kernel for the clamp
      if (!last && Tn != 0.0f) {
        const float bound = Tn * qmax;
        suffix = fminf(fmaxf(rgd.total - prefix, -bound), bound);
      }
      dpre = delta * (Tn * q - suffix) * dscale_abs;
      if (!b.gdens) dpre = 0.f;
      if (b.gfeat) {
        draw[0] = w * rgd.gc[0] * sr * (1.0f - sr);
        draw[1] = w * rgd.gc[1] * sg * (1.0f - sg);
        draw[2] = w * rgd.gc[2] * sb * (1.0f - sb);
        if constexpr (DUAL) {
          draw0[0] = w * rgd.gcd[0] * cd.x * (1.0f - cd.x);
          draw0[1] = w * rgd.gcd[1] * cd.y * (1.0f - cd.y);
          draw0[2] = w * rgd.gcd[2] * cd.z * (1.0f - cd.z);
        }
      }
      contributes = (dpre != 0.f) || (draw[0] != 0.f) || (draw[1] != 0.f) || (draw[2] != 0.f);
      if constexpr (DUAL) contributes = contributes || (draw0[0] != 0.f) || (draw0[1] != 0.f) || (draw0[2] != 0.f);
      T = Tn;
      if (T == 0.0f) alive = false;  // every later weight is exactly 0
    }
    const unsigned act = __ballot_sync(FULL, contributes);
    if (act == 0u) continue;

    // ---- slot = cell-sorted position: (samples of cells whose first lane precedes this cell's first lane) + (position in the cell)
    unsigned peers = 0u;
    if (contributes)
      peers = __match_any_sync(act, (unsigned long long)(unsigned)(cell.ox[0] + cell.oy[0] + cell.oz[0]) |
                                        ((unsigned long long)(unsigned)(cell.ox[1] + cell.oy[1] + cell.oz[1]) << 32));
    const int leader = contributes ? (__ffs(peers) - 1) : lane;
    int scan = (contributes && leader == lane) ? __popc(peers) : 0;
    const int own = scan;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(FULL, scan, o);
      if (lane >= o) scan += up;
    }
    const int slot = __shfl_sync(FULL, scan - own, leader) + __popc(peers & ((1u << lane) - 1u));

    if (published >= NS) mbar_wait_bounded(&sm.done[ps], ppar);  // the consumer has finished the stage's previous contents
    WbStage<DUAL>& st = sm.st[ps];
    if (contributes) {
      float wc[8];
      unsigned vx_[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
        wc[k] = cell.wx[ix] * cell.wy[iy] * cell.wz[iz];
        vx_[k] = (unsigned)(cell.ox[ix] + cell.oy[iy] + cell.oz[iz]);
      }
      st.Wlo[slot] = make_float4(wc[0], wc[1], wc[2], wc[3]);
      st.Whi[slot] = make_float4(wc[4], wc[5], wc[6], wc[7]);
      st.Vlo[slot] = make_uint4(vx_[0], vx_[1], vx_[2], vx_[3]);
      st.Vhi[slot] = make_uint4(vx_[4], vx_[5], vx_[6], vx_[7]);
      st.G[slot] = make_float4(draw[0], draw[1], draw[2], dpre);
      if constexpr (DUAL) st.G2[slot] = make_float4(draw0[0], draw0[1], draw0[2], 0.0f);
      st.yoff[slot] = (unsigned)lane * (unsigned)(H::YROW * 4);
    }
    if (lane == 0) st.n = __popc(act);
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.full[ps]);
    ++published;
    if (++ps == NS) {
      ps = 0;
      if (published > NS) ppar ^= 1u;
    }
  }
  // end marker: the stage must be free like any other
  if (published >= NS) mbar_wait_bounded(&sm.done[ps], ppar);
  if (lane == 0) sm.st[ps].n = -1;
  __syncwarp();
  if (lane == 0) mbar_arrive(&sm.full[ps]);
}

}  // namespace r3d
