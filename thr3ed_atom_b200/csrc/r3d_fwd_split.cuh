// r3d_fwd_split.cuh -- the forward as TWO kernels (training renders: sample cache + contribution ballots requested).
//
// Why.  DESIGN.md 4.5: the fused lane-group forward is bound by the dependent chain of a marching step
// (depth -> position -> cell -> density probe -> vote -> publish -> record gather -> reduction -> sigmoid -> compositing),
// ~850 dependent instructions per step at 4 warps per scheduler.  Two observations split that chain without changing a
// single result bit:
//   * transmittance, weight, depth and accumulated weight of a sample depend on sigma and the sample interval only
//     (accumulate.py:43-88) -- never on the radiance; only  colour = sum_i w_i * sigmoid(raw_i)  needs the SH records;
//   * the training forward already writes one 16-byte record per contributing sample (the sample cache) and one ballot per
//     warp and marching step (the backward marches by them).
// So:
//   render_fwd_probe_kernel   marches the ray exactly like the fused kernel (same depths, inside test, cell, density,
//                             alpha / T chain), composites depth and acc, and leaves per contributing sample the continuous
//                             grid coordinates of the sample + sigma in its sample-cache slot and its weight w_i in a
//                             [S][N] float plane; writes the ballots.  No feature traffic, no shared memory: a light kernel.
//   render_fwd_gather_kernel  walks the ballots (empty steps cost one bit test), rebuilds the cell from the stored
//                             coordinates (the weights are (floor + 1) - gi and gi - floor: bit-identical), gathers the 8
//                             corner records with lane groups exactly like the fused kernel, applies the SH basis, replaces
//                             the slot by (sigmoid(raw) rgb, sigma) -- the record the backward wants -- and accumulates
//                             colour += w_i * sigmoid(raw_i) in step order.  Steps are independent of one another here (no
//                             transmittance chain, no probe), so the next step's slot is requested one step ahead and the
//                             register file is free for the gather.
// Same samples, same operations in the same order: colour, depth, acc, cache and ballots are bit-identical to the fused
// kernel.  Extra HBM traffic: the slot is written twice and read once, the weight plane written and read once
// (~2 GB at c3); HBM is ~10 % busy in this path.
#pragma once

namespace r3d {

// grid coordinate of axis_cell_inside (same operations)
__device__ __forceinline__ float axis_gi(float p, float ns, float nb, int dim) {
  const float n = __fadd_rn(__fmul_rn(p, ns), nb);
  return ((n + 1.0f) * (float)dim - 1.0f) * 0.5f;
}
// the rest of axis_cell_inside, from the stored coordinate
__device__ __forceinline__ void axis_cell_from_gi(float gi, int dim, int mul, int (&off)[2], float (&w)[2]) {
  const float fl = floorf(gi);
  const int i0 = (int)fl;
  w[0] = ((unsigned)i0 < (unsigned)dim) ? (fl + 1.0f) - gi : 0.0f;
  w[1] = ((unsigned)(i0 + 1) < (unsigned)dim) ? gi - fl : 0.0f;
  off[0] = min(max(i0, 0), dim - 1) * mul;
  off[1] = min(max(i0 + 1, 0), dim - 1) * mul;
}

#ifndef R3D_SPLIT_PROBE_BLOCKS
#define R3D_SPLIT_PROBE_BLOCKS 8
#endif
#ifndef R3D_SPLIT_GATHER_BLOCKS
#define R3D_SPLIT_GATHER_BLOCKS 6
#endif

// DQ: the 8 corner densities come from the density quad volume (r3d_device.cuh: two 16-byte loads per sample, no clamps);
// the probe is L1-data-pipe bound like everything else here (8 scalar loads of a warp touch ~7 lines each), the quads cut
// its wavefronts ~4x.  Same products in the same order: bit-identical sigma.
template <bool DQ>
__global__ void __launch_bounds__(128, R3D_SPLIT_PROBE_BLOCKS)
    render_fwd_probe_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out, float* __restrict__ wplane) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  const bool alive = ray >= 0;
  RayCtx s;
  s.i_lo = 1, s.i_hi = 0;
  if (alive) {
    float vx, vy, vz;
    setup_ray(g, rp, c, ray, s, vx, vy, vz);
  }
  const Ray& r = s.r;
  bool marching = alive && s.i_lo <= s.i_hi;
  int lo = marching ? s.i_lo : 0x7fffffff, hi = marching ? s.i_hi : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(FULL, lo, o));
    hi = max(hi, __shfl_xor_sync(FULL, hi, o));
  }
  const size_t mask_col = blockIdx.x * 4u + (threadIdx.x >> 5), mask_row = gridDim.x * 4u;
  float T = 1.0f, dep = 0.f, acc = 0.f, z = 0.f;
  bool have_z = false;
  DepthMarch dm;
  dm.bm = dm.bc = 0.f;
  for (int i = lo; i <= hi; ++i) {
    bool contributes = false, last = false;
    float sigma = 0.f, zn = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    const bool mine = marching && i >= s.i_lo && i <= s.i_hi;
    if (mine) {
      if (!have_z) dm.start(s.dg, i), z = dm.next(s.dg, i), have_z = true;
      last = (i == c.S - 1);
      zn = last ? 0.0f : dm.next(s.dg, i + 1);
      const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
      const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
      const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
      if (inside_aabb(g, px, py, pz)) {
        gx = axis_gi(px, g.ns[0], g.nb[0], g.W), gy = axis_gi(py, g.ns[1], g.nb[1], g.D), gz = axis_gi(pz, g.ns[2], g.nb[2], g.H);
        float dpost;
        if constexpr (DQ) {
          CellQ cq;  // axis_cell_q from the coordinate
          const float fx = floorf(gx), fy = floorf(gy), fz = floorf(gz);
          cq.wx[0] = (fx + 1.0f) - gx, cq.wx[1] = gx - fx, cq.ix = min(max((int)fx, -1), g.W - 1);
          cq.wy[0] = (fy + 1.0f) - gy, cq.wy[1] = gy - fy, cq.iy = min(max((int)fy, -1), g.D - 1);
          cq.wz[0] = (fz + 1.0f) - gz, cq.wz[1] = gz - fz, cq.iz = min(max((int)fz, -1), g.H - 1);
          sigma = density_post(g.post, density_pre_interp_q(g, cq), dpost);
        } else {
          Cell cell;
          axis_cell_from_gi(gx, g.W, g.D * g.H, cell.ox, cell.wx);
          axis_cell_from_gi(gy, g.D, g.H, cell.oy, cell.wy);
          axis_cell_from_gi(gz, g.H, 1, cell.oz, cell.wz);
          sigma = density_post(g.post, density_pre_interp<true>(g, cell), dpost);
        }
        contributes = sigma != 0.0f;
      }
    }
    const unsigned act = __ballot_sync(FULL, contributes);
    if (lane == 0) out.mask[(size_t)i * mask_row + mask_col] = act;
    if (contributes) {
      const float delta = last ? __fmul_rn(kInfinity, s.dnorm) : __fmul_rn(__fsub_rn(zn, z), s.dnorm);
      const float alpha = 1.0f - exp_neg(sigma * delta);
      const float w = alpha * T;
      const size_t slot = (size_t)i * rp.n + ray;
      __stcs(out.cache + slot, make_float4(gx, gy, gz, sigma));
      __stcs(wplane + slot, w);
      dep = fmaf(w, z, dep);
      acc += w;
      T *= (1.0f - alpha);
      if (T == 0.0f) marching = false;
    }
    if (mine) z = zn;
  }
  if (!alive) return;
  out.depth[ray] = dep;
  out.acc[ray] = acc;
  if (out.disparity) {
    const float ratio = __fdiv_rn(dep, acc);
    const float m = (ratio != ratio) ? ratio : fmaxf(kZeroPlus, ratio);
    out.disparity[ray] = __fdiv_rn(1.0f, m);
  }
}

// Shared memory of one gather warp.  The kernel lives on L1 hits (69 % of its sectors), and the L1 is what the CTAs' shared
// memory leaves of the SM's 256 KB: the SH table is therefore stored compact (K values per ray, not expanded per record
// element as in the fused kernel: 1.1 KB instead of 3.5 KB per warp at degree 2) -- 6 CTAs/SM then fit the 100 KB carve-out.
template <int DEG>
struct alignas(16) GatherSmem {
  static constexpr int K = (DEG + 1) * (DEG + 1);
  float W[32 * 8];     // rows per contributing sample of the current marching step (rank order): 8 corner weights; once the
                       // sample's group has read them, elements 0..2 of the row take the raw radiance (r, g, b) back to the owner
  unsigned V[32 * 8];  // corner record indices in float4 units
  float Y[32 * K];     // SH basis per lane (= ray)
  unsigned char src[32];  // owning lane of each rank
};

// The gather kernel walks the NON-EMPTY marching steps of its warp (A = gathered now, B = next, C = the one after):
//   C: its sample-cache slots + weights are requested                                        (HBM, two steps ahead)
//   B: with PFK != 0 the (x, y) columns of its cell are requested into L2, a whole gather ahead of their use: the gather of
//      a step waits for its slowest sector, 15 % of the sectors come from DRAM and every iteration had one -- 57 % of all
//      stall samples sat on the first use of a gathered record (profiles/r02_split_forward_ncu.md)
//   A: cell, 8 weights, 8 record indices from the stored coordinates -> published; lane groups gather / reduce; owners apply
//      sigmoid, write the backward's record, accumulate colour.
// PFK: 1 = prefetch.global.L2 at the start / middle / end of each (x, y) column (two z-adjacent records = 8 * stride bytes),
//      2 = one cp.async.bulk.prefetch.L2 per column, 3 = prefetch.global.L2 every 32 bytes of the column.
// RUN: run-merged gather.  The contributing samples of a marching step are handed to the lane groups in contiguous quarters
// (group g takes samples [g * per, (g + 1) * per) of the published order) and a group keeps the 8 corner records in registers
// while consecutive samples of its quarter lie in the same interpolation cell.
template <int DEG, int PFK = 0, bool RUN = false>
__global__ void __launch_bounds__(128, DEG >= 3 ? 4 : R3D_SPLIT_GATHER_BLOCKS)
    render_fwd_gather_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out, const float* __restrict__ wplane) {
  using H = FwdGroupShape<DEG>;
  using S = CoopShape<DEG>;
  constexpr int K = S::K, F = S::F, NV = S::NV, LPR = H::LPR, MPI = H::MPI;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) GatherSmem<DEG> smem_all[4];
  GatherSmem<DEG>& sm = smem_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ray = thread_to_ray(rp, t);
  const bool alive = ray >= 0;
  int my_lo = 1, my_hi = 0;
  {
    float Y[K];
#pragma unroll
    for (int k = 0; k < K; ++k) Y[k] = 0.f;
    if (alive) {
      RayCtx s;
      float vx, vy, vz;
      setup_ray(g, rp, c, ray, s, vx, vy, vz);
      sh_basis<DEG>(vx, vy, vz, Y);
      my_lo = s.i_lo, my_hi = s.i_hi;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) sm.Y[lane * K + k] = Y[k];
  }
  const bool marching = alive && my_lo <= my_hi;
  int lo = marching ? my_lo : 0x7fffffff, hi = marching ? my_hi : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(FULL, lo, o));
    hi = max(hi, __shfl_xor_sync(FULL, hi, o));
  }
  __syncwarp();

  const int ms = lane / LPR, cj = lane % LPR;
  const bool role_ok = cj < NV;
  const unsigned stride4 = (unsigned)g.stride >> 2;
  const unsigned long long feat_base = reinterpret_cast<unsigned long long>(g.feat);
  const unsigned long long feat_lane = feat_base + 16ull * (unsigned)cj;
  int out_slot = -1, trade_lane = lane;
  if constexpr (DEG == 1) out_slot = role_ok ? cj : -1;
  if constexpr (DEG == 3) out_slot = (role_ok && (cj & 3) == 0) ? (cj >> 2) : -1;
  if constexpr (DEG == 2) {
    out_slot = cj == 0 ? 0 : (cj == 3 ? 1 : (cj == 5 ? 2 : -1));
    const int x = (cj == 0 || cj == 2) ? 2 : ((cj == 3 || cj == 4) ? 7 : ((cj == 5 || cj == 6) ? 3 : 0));
    trade_lane = lane ^ x;
  }
  const bool split1 = DEG == 2 && cj == 2, split2 = DEG == 2 && cj == 4, odd = (cj & 1) != 0;
  // record element e = 4 * cj + l belongs to SH coefficient e % K; the pad elements (e >= F) get weight 0
  int yk[4];
  bool ypad[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) yk[l] = (4 * cj + l) % K, ypad[l] = (4 * cj + l) >= F;

  const size_t mask_col = blockIdx.x * 4u + (threadIdx.x >> 5), mask_row = gridDim.x * 4u;
  const size_t n = (size_t)rp.n;
  const unsigned lt = (1u << lane) - 1u;
  float cr = 0.f, cg = 0.f, cb = 0.f;

  // ---- the non-empty marching steps of this warp, in order (ballots are fetched 32 steps at a time, one per lane) ----
  int base = (hi >= lo) ? lo - 32 : 0;
  unsigned mymask = 0u, todo = 0u;
  auto next_step = [&](unsigned& act) -> int {
    while (todo == 0u) {
      base += 32;
      if (base > hi) return -1;
      mymask = (base + lane <= hi) ? __ldcs(out.mask + (size_t)(base + lane) * mask_row + mask_col) : 0u;
      todo = __ballot_sync(FULL, mymask != 0u);
    }
    const int k = __ffs(todo) - 1;
    todo &= todo - 1u;
    act = __shfl_sync(FULL, mymask, k);
    return base + k;
  };
  auto request = [&](int i, unsigned act, float4& rec, float& w) {
    if (i >= 0 && ((act >> lane) & 1u)) {
      const size_t slot = (size_t)i * n + ray;
      rec = __ldcs(out.cache + slot), w = __ldcs(wplane + slot);
    }
  };
  // the (x, y) columns of a sample's cell -> L2 (a hint: same cell arithmetic, only the four column origins are formed)
  auto prefetch_cell = [&](unsigned act, const float4& rec) {
    if ((act >> lane) & 1u) {
      int ox[2], oy[2], oz[2];
      float unused[2];
      axis_cell_from_gi(rec.x, g.W, g.D * g.H, ox, unused);
      axis_cell_from_gi(rec.y, g.D, g.H, oy, unused);
      axis_cell_from_gi(rec.z, g.H, 1, oz, unused);
      const unsigned bytes = (oz[1] != oz[0] ? 8u : 4u) * (unsigned)g.stride;  // the z + 1 record follows the z record in memory
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned long long col = feat_base + 16ull * ((unsigned)(ox[q >> 1] + oy[q & 1] + oz[0]) * stride4);
        if constexpr (PFK == 9) {
        } else if constexpr (PFK == 1) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(col));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(col + (bytes >> 1)));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(col + bytes - 16u));
        } else if constexpr (PFK == 2) {
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(col), "r"(bytes) : "memory");
        } else {
          for (unsigned o = 0; o < bytes; o += 32u) asm volatile("prefetch.global.L2 [%0];" ::"l"(col + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(col + bytes - 16u));
        }
      }
    }
  };
  // cell of a sample from its stored coordinates -> published rows
  auto publish = [&](unsigned act, const float4& rec) {
    if ((act >> lane) & 1u) {
      const int rank = __popc(act & lt);
      Cell cell;
      axis_cell_from_gi(rec.x, g.W, g.D * g.H, cell.ox, cell.wx);
      axis_cell_from_gi(rec.y, g.D, g.H, cell.oy, cell.wy);
      axis_cell_from_gi(rec.z, g.H, 1, cell.oz, cell.wz);
      float wc[8];
      unsigned rec4[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int ix = q >> 2, iy = (q >> 1) & 1, iz = q & 1;
        wc[q] = cell.wx[ix] * cell.wy[iy] * cell.wz[iz];
        rec4[q] = (unsigned)(cell.ox[ix] + cell.oy[iy] + cell.oz[iz]) * stride4;
      }
      float* Wrow = sm.W + rank * 8;
      *reinterpret_cast<float4*>(Wrow) = make_float4(wc[0], wc[1], wc[2], wc[3]);
      *reinterpret_cast<float4*>(Wrow + 4) = make_float4(wc[4], wc[5], wc[6], wc[7]);
      unsigned* Vrow = sm.V + rank * 8;
      *reinterpret_cast<uint4*>(Vrow) = make_uint4(rec4[0], rec4[1], rec4[2], rec4[3]);
      *reinterpret_cast<uint4*>(Vrow + 4) = make_uint4(rec4[4], rec4[5], rec4[6], rec4[7]);
      sm.src[rank] = (unsigned char)lane;
    }
  };

  unsigned act_a = 0u, act_b = 0u, act_c = 0u;
  int i_a = -1, i_b = -1, i_c = -1;
  float w_a = 0.f, w_b = 0.f, w_c = 0.f;
  float4 rec_a = make_float4(0.f, 0.f, 0.f, 0.f), rec_b = rec_a, rec_c = rec_a;
  i_b = next_step(act_b);
  request(i_b, act_b, rec_b, w_b);
  if (i_b >= 0) {
    i_c = next_step(act_c);
    request(i_c, act_c, rec_c, w_c);
  }
  while (i_b >= 0) {
    // rotate: A <- B, B <- C (requested one iteration ago), request the new C, prefetch B's records
    i_a = i_b, act_a = act_b, rec_a = rec_b, w_a = w_b;
    i_b = i_c, act_b = act_c, rec_b = rec_c, w_b = w_c;
    if (i_b >= 0) {
      i_c = next_step(act_c);
      request(i_c, act_c, rec_c, w_c);
      if constexpr (PFK != 0 && PFK != 9) prefetch_cell(act_b, rec_b);
    }
    publish(act_a, rec_a);
    __syncwarp();
    const int total = __popc(act_a);
    const int per = (total + MPI - 1) / MPI;  // RUN: samples per lane group (contiguous quarters of the published order)
    float4 q[8];
    unsigned held0 = 0xffffffffu, held7 = 0xffffffffu;  // RUN: the cell whose records this lane holds (corner 0 / corner 7 record)
    for (int b = 0; b < (RUN ? per : total); b += (RUN ? 1 : MPI)) {
      const int m = RUN ? ms * per + b : b + ms;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < total && role_ok) {
        const uint4 v0 = *reinterpret_cast<const uint4*>(sm.V + m * 8);
        const uint4 v1 = *reinterpret_cast<const uint4*>(sm.V + m * 8 + 4);
        const unsigned vk[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        if (!RUN || v0.x != held0 || v1.w != held7) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            unsigned long long addr;
            asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(addr) : "r"(vk[e]), "l"(feat_lane));
            if constexpr (PFK == 9)  // (the compiler would narrow the load to the one component the measurement variant uses)
              asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q[e].x), "=f"(q[e].y), "=f"(q[e].z), "=f"(q[e].w) : "l"(addr));
            else
              q[e] = __ldg(reinterpret_cast<const float4*>(addr));
          }
          held0 = v0.x, held7 = v1.w;
        }
        const float4 w0 = *reinterpret_cast<const float4*>(sm.W + m * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sm.W + m * 8 + 4);
        const float* Yr = sm.Y + sm.src[m] * K;
        const float y0 = ypad[0] ? 0.0f : Yr[yk[0]], y1 = ypad[1] ? 0.0f : Yr[yk[1]];
        const float y2 = ypad[2] ? 0.0f : Yr[yk[2]], y3 = ypad[3] ? 0.0f : Yr[yk[3]];
        const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        if constexpr (PFK == 9) {  // measurement only: the gather without its arithmetic.  All four components of every load are
          // consumed (ptxas narrows a vector load whose upper components are dead): two LOP3 per record instead of two FFMA2
          unsigned bits = 0u;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            bits ^= __float_as_uint(q[e].x) ^ __float_as_uint(q[e].y);
            bits ^= __float_as_uint(q[e].z) ^ __float_as_uint(q[e].w);
          }
          a.x = __uint_as_float((bits & 0x007fffffu) | 0x3f000000u) * (wk[0] + wk[7]) * (y0 + y1 + y2 + y3);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a.x = fmaf(wk[e], q[e].x, a.x), a.y = fmaf(wk[e], q[e].y, a.y);
            a.z = fmaf(wk[e], q[e].z, a.z), a.w = fmaf(wk[e], q[e].w, a.w);
          }
          a.x *= y0, a.y *= y1, a.z *= y2, a.w *= y3;
        }
      }
      if constexpr (DEG == 0) {
        if (m < total) *reinterpret_cast<float4*>(sm.W + m * 8) = a;
      } else {
        const float u01 = a.x + a.y, u23 = a.z + a.w;
        float v;
        if constexpr (DEG == 2) {
          const float A = split1 ? a.x : (split2 ? u01 : u01 + u23);
          const float B = split1 ? a.y + u23 : (split2 ? u23 : 0.0f);
          const float x = __shfl_xor_sync(FULL, odd ? A : B, 1);
          v = (split1 || split2) ? A : A + x;
          v += __shfl_sync(FULL, v, trade_lane);
        } else {
          v = u01 + u23;
          if constexpr (DEG == 3) {
            v += __shfl_xor_sync(FULL, v, 1);
            v += __shfl_xor_sync(FULL, v, 2);
          }
        }
        if (m < total && out_slot >= 0) sm.W[m * 8 + out_slot] = v;
      }
    }
    __syncwarp();
    if ((act_a >> lane) & 1u) {
      const float4 raw = *reinterpret_cast<const float4*>(sm.W + __popc(act_a & lt) * 8);
      const float sr = sigmoidf_(raw.x), sg = sigmoidf_(raw.y), sb2 = sigmoidf_(raw.z);
      __stcs(out.cache + ((size_t)i_a * n + ray), make_float4(sr, sg, sb2, rec_a.w));
      cr = fmaf(w_a, sr, cr);
      cg = fmaf(w_a, sg, cg);
      cb = fmaf(w_a, sb2, cb);
    }
  }
  if (!alive) return;
  if (c.flags & R3D_FLAG_WHITE_BKGD) {
    const float bg = 1.0f - out.acc[ray];
    cr += bg, cg += bg, cb += bg;
  }
  out.colour[3 * ray] = cr, out.colour[3 * ray + 1] = cg, out.colour[3 * ray + 2] = cb;
}

}  // namespace r3d
