// r3d_fwd_ws.cuh -- forward render, warp-specialised producer / consumer kernel (sm_100a).  Included by r3d_render.cu.
//
// Why.  The lane-group forward (render_fwd_group_kernel) does everything in one warp: march a step (position, inside test,
// density probe), vote, publish the contributing samples, gather their records with lane groups, read the sums back,
// composite -- one long dependent chain with two global-memory round trips per marching step.  Its profile
// (profiles/r01_v6_ncu_full_summary.md) is neither HBM- nor pipe-bound: issue slots 52 % busy, 2.95 warps stalled on the
// long scoreboard per issued instruction, 16 warps/SM at 128 registers.  Here the two halves of a step run in DIFFERENT
// warps of the same CTA, decoupled by a small ring of stages in shared memory and mbarriers:
//
//   producer warp  (4 per CTA, one thread per ray of an 8x4 pixel tile)
//       march: depth (from a per-CTA table of the stratum bounds), position, strict inside test, density probe, ReLU ->
//       ballot of the contributing samples; publish their 8 trilinear weights + 8 record indices + alpha into stage s,
//       arrive on full[s]; then composite the stage published LAG steps earlier (wait on done[], read the raw radiance,
//       sigmoid, front-to-back accumulation, sample-cache record for the backward).
//   consumer warp  (4 per CTA, paired 1:1 with a producer)
//       wait on full[s]; lane groups of LPR lanes gather the 8 corner records of each published sample straight from
//       global memory (coalesced 112-byte reads), apply weights and the ray's SH basis (packed FFMA2), reduce over the
//       group, store the raw radiance into the stage, arrive on done[s].
//
// The producer is always LAG stages ahead, so the consumer's loop never waits for a density probe and the producer's
// probe never waits for a record gather; each stream needs fewer registers than the fused loop (more resident warps).
// Results are bit-identical to render_fwd_group_kernel (same per-sample arithmetic, same order of accumulation).
#pragma once

#ifndef R3D_WS_BLOCKS
#define R3D_WS_BLOCKS 3
#endif
#ifndef R3D_WS_BLOCKS_DQ
#define R3D_WS_BLOCKS_DQ 2
#endif
#ifndef R3D_WS_STAGES
#define R3D_WS_STAGES 4
#endif
#ifndef R3D_WS_LAG
#define R3D_WS_LAG 3
#endif

namespace r3d {

constexpr int kWsStages = R3D_WS_STAGES;  // ring depth per producer/consumer pair
constexpr int kWsLag = R3D_WS_LAG;        // the producer composites stage p - kWsLag after publishing stage p
static_assert(kWsLag < kWsStages, "a stage must have been composited before it is published again");

// Bounded wait: a protocol bug must surface as a CUDA error (trap), never as a hung GPU.  try_wait suspends the warp in
// hardware for up to a system-dependent time per attempt, so the bound is generous (seconds).
__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, unsigned parity) {
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
    if (spin > (1u << 24)) __trap();
  }
}

// one non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <bool DUAL>
struct alignas(16) WsStage {
  // per published sample, in rank order (rank = position of the ray's lane among the contributing lanes).  The 8-float
  // rows are split into two 16-byte-stride arrays: 128-bit stores of consecutive ranks then fall into consecutive banks.
  float4 Wlo[32], Whi[32];      // trilinear weights of corners 0..3 / 4..7
  uint4 Vlo[32], Vhi[32];       // record indices (float4 units) of corners 0..3 / 4..7
  float4 pend[32];              // (sigma, z, alpha, -) kept for the compositing pass
  float4 R[32];                 // raw radiance (r, g, b, -), written by the consumer
  float4 R2[DUAL ? 32 : 1];     // raw band-0 radiance (single-pass specular + diffuse render)
  unsigned yoff[32];            // byte offset of the owning ray's row in the pair's SH table
  int n, step;                  // samples in this stage (-1: the producer is done), marching step
  unsigned act, pad0;           // ballot of the contributing lanes
};

template <int DEG, bool DUAL>
struct alignas(16) WsPair {
  using H = FwdGroupShape<DEG>;
  WsStage<DUAL> st[kWsStages];
  float Y[32 * H::YROW];  // expanded SH row per ray (Y[e % K] for record element e < F, 0 for the pad)
  unsigned long long full[kWsStages], done[kWsStages];
};

template <int DEG, bool DUAL>
__host__ __device__ constexpr size_t ws_smem_bytes(int S, bool table) {
  return 4 * sizeof(WsPair<DEG, DUAL>) + (table ? sizeof(float2) * (size_t)S : 0);
}

// Depth of sample j from the per-CTA table of stratum bounds: bit-identical to DepthGen::at (the table holds `lower` and
// `upper - lower` formed with the same un-fused operations; without jitter it holds base(j) and 0).
struct DepthTab {
  const float2* tab;
  bool perturb;
  const float* __restrict__ jit;
  unsigned key;
  __device__ __forceinline__ float at(int j) const {
    const float2 t = tab[j];
    if (!perturb) return t.x;
    const float u = jit ? __ldg(jit + j) : jitter_u(key, j);
    return __fadd_rn(t.x, __fmul_rn(t.y, u));
  }
};

// SORT: the producer publishes the contributing samples grouped by interpolation cell (__match_any_sync on the cell key +
// a warp scan of the group sizes) instead of in lane order, so that every distinct cell of a marching step is one run.
// PF: the producer prefetches the published samples' corner records (1 = into L2, 2 = into L1) kWsLag stages before the
// consumer gathers them.
// DQ: the consumer gathers for two stages at a time (two record register sets per lane).
template <int DEG, bool DUAL, bool SORT, int PF, bool DQ>
__global__ void __launch_bounds__(256, DEG >= 3 ? 2 : (DQ ? R3D_WS_BLOCKS_DQ : R3D_WS_BLOCKS))
    render_fwd_ws_kernel(const GridP g, const RaysP rp, const CfgP c, const OutP out, const int use_tab) {
  using H = FwdGroupShape<DEG>;
  using S = CoopShape<DEG>;
  constexpr int K = S::K, F = S::F, NV = S::NV, LPR = H::LPR, MPI = H::MPI, NS = kWsStages;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char ws_smem[];
  WsPair<DEG, DUAL>* pairs = reinterpret_cast<WsPair<DEG, DUAL>*>(ws_smem);
  float2* ztab = reinterpret_cast<float2*>(ws_smem + 4 * sizeof(WsPair<DEG, DUAL>));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WsPair<DEG, DUAL>& sm = pairs[warp & 3];
  const bool producer = warp < 4;

  if (producer && lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&sm.full[s], 1), mbar_init(&sm.done[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (use_tab) {  // stratum table: (lower_j, upper_j - lower_j), sample.py:54-64
    DepthGen dg;
    dg.near = c.near, dg.far = c.far, dg.S = c.S, dg.half = c.S / 2;
    dg.step = c.S > 1 ? __fdiv_rn(1.0f, (float)(c.S - 1)) : 0.0f;
    const bool perturb = (c.flags & R3D_FLAG_PERTURB) != 0;
    for (int j = threadIdx.x; j < c.S; j += 256) {
      const float b = dg.base(j);
      float lower = b, span = 0.0f;
      if (perturb) {
        lower = (j > 0) ? 0.5f * __fadd_rn(b, dg.base(j - 1)) : b;
        const float upper = (j < c.S - 1) ? 0.5f * __fadd_rn(dg.base(j + 1), b) : b;
        span = __fsub_rn(upper, lower);
      }
      ztab[j] = make_float2(lower, span);
    }
  }
  __syncthreads();

  if (!producer) {
    // =========================================================================== consumer: record gather
    // Group `ms` (LPR lanes) takes a CONTIGUOUS chunk of the stage's samples.  Ranks follow the lanes of the 8x4 pixel
    // tile, so neighbours in rank are neighbours in the image and often fall into the same interpolation cell: the
    // group keeps the 8 corner records in registers and reloads them only when the cell changes -- the L1 data pipe
    // (the binding unit of this path, profiles/r02_*) then moves a cell's 8 x 112 bytes once per run instead of once
    // per sample.
    const int ms = lane / LPR, cj = lane % LPR;
    const bool role_ok = cj < NV;
    const unsigned long long feat_lane = reinterpret_cast<unsigned long long>(g.feat) + 16ull * (unsigned)cj;
    const unsigned y_lane = (unsigned)__cvta_generic_to_shared(sm.Y) + 16u * (unsigned)cj;
    int out_slot = -1, trade_lane = lane;
    if constexpr (DEG == 1) out_slot = role_ok ? cj : -1;
    if constexpr (DEG == 3) out_slot = (role_ok && (cj & 3) == 0) ? (cj >> 2) : -1;
    if constexpr (DEG == 2) {
      out_slot = cj == 0 ? 0 : (cj == 3 ? 1 : (cj == 5 ? 2 : -1));
      const int x = (cj == 0 || cj == 2) ? 2 : ((cj == 3 || cj == 4) ? 7 : ((cj == 5 || cj == 6) ? 3 : 0));
      trade_lane = lane ^ x;
    }
    const bool split1 = DEG == 2 && cj == 2, split2 = DEG == 2 && cj == 4, odd = (cj & 1) != 0;
    int dslot = -1, dcomp = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (DUAL && DEG > 0 && cj == (ch * K) / 4) dslot = ch, dcomp = (ch * K) % 4;

    // load(): V row of sample m; when it names another cell than the one held in q[], request its 8 corner records
    auto load = [&](WsStage<DUAL>& st, int m, bool on, float4(&q)[8], unsigned& key0, unsigned& key7) {
      if (!on) return;
      const uint4 v0 = st.Vlo[m], v1 = st.Vhi[m];
      if (v0.x != key0 || v1.w != key7) {  // corners 0 and 7 identify the cell (low and high voxel of every axis)
        key0 = v0.x, key7 = v1.w;
        const unsigned vk[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // address = lane base + 16 * record index
          unsigned long long addr;
          asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(addr) : "r"(vk[k]), "l"(feat_lane));
          q[k] = __ldg(reinterpret_cast<const float4*>(addr));
        }
      }
    };
    // maths(): weights, SH row, channel sums of the group -> raw radiance of sample m (executed by every lane: shuffles)
    auto maths = [&](WsStage<DUAL>& st, int m, bool on, const float4(&q)[8]) {
      float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
      if (on) {
        const float4 w0 = st.Wlo[m], w1 = st.Whi[m];
        float4 y4;
        lds_v4(y_lane + st.yoff[m], y4.x, y4.y, y4.z, y4.w);
        const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a01 = ffma2(make_float2(q[k].x, q[k].y), wk[k], a01);
          a23 = ffma2(make_float2(q[k].z, q[k].w), wk[k], a23);
        }
        if constexpr (DUAL && DEG > 0) {  // band-0 radiance: C0 * interpolated coeff[ch][0]
          if (dslot >= 0)
            reinterpret_cast<float*>(&st.R2[m])[dslot] = 0.28209479177387814f * (dcomp == 0 ? a01.x : (dcomp == 1 ? a01.y : (dcomp == 2 ? a23.x : a23.y)));
        }
        a01.x *= y4.x, a01.y *= y4.y, a23.x *= y4.z, a23.y *= y4.w;  // the pad element has Y = 0
      }
      if constexpr (DEG == 0) {
        if (on) st.R[m] = make_float4(a01.x, a01.y, a23.x, a23.y);
      } else {
        const float u01 = a01.x + a01.y, u23 = a23.x + a23.y;
        float v;
        if constexpr (DEG == 2) {  // see render_fwd_group_kernel: float4s 2 and 4 straddle a channel boundary
          const float A = split1 ? a01.x : (split2 ? u01 : u01 + u23);
          const float B = split1 ? a01.y + u23 : (split2 ? u23 : 0.0f);
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
        if (on && out_slot >= 0) reinterpret_cast<float*>(&st.R[m])[out_slot] = v;
      }
    };

    int gs = 0;
    unsigned gpar = 0u;
    float4 qa[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) qa[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned ka0 = 0xffffffffu, ka7 = 0xffffffffu;  // corner-0 / corner-7 record of the cell held in qa[] (the grid is read-only)
    if constexpr (DQ) {
      // Two stages at a time whenever the producer is far enough ahead (it runs kWsLag stages ahead, so almost always): each
      // group walks one chunk of stage A and one of stage B in the same iteration, with separate record registers, so 16
      // independent 128-bit loads are in flight per lane before the first is consumed.  The consumer is latency-bound on
      // exactly those loads (profiles/r02_ws2_*: 40 % of all stall samples sit on their first use).
      float4 qb[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) qb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      unsigned kb0 = 0xffffffffu, kb7 = 0xffffffffu;
      while (true) {
        WsStage<DUAL>& sa = sm.st[gs];
        mbar_wait_bounded(&sm.full[gs], gpar);
        const int na = sa.n;
        if (na < 0) break;
        const int gs2 = (gs + 1 == NS) ? 0 : gs + 1;
        const unsigned gpar2 = (gs + 1 == NS) ? (gpar ^ 1u) : gpar;
        WsStage<DUAL>& sb = sm.st[gs2];
        int nb = 0;
        const bool two = __all_sync(FULL, mbar_test(&sm.full[gs2], gpar2)) && (nb = sb.n) > 0;  // warp-uniform
        const int ca = (na + MPI - 1) / MPI, cb = two ? (nb + MPI - 1) / MPI : 0;
        const int ea = min(na, (ms + 1) * ca), eb = min(nb, (ms + 1) * cb);
        const int iters = max(ca, cb);
        for (int it = 0; it < iters; ++it) {
          const int ma = ms * ca + it, mb = ms * cb + it;
          const bool ona = it < ca && ma < ea && role_ok, onb = it < cb && mb < eb && role_ok;
          load(sa, ma, ona, qa, ka0, ka7);
          load(sb, mb, onb, qb, kb0, kb7);
          maths(sa, ma, ona, qa);
          if (two) maths(sb, mb, onb, qb);
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sm.done[gs]);
          if (two) mbar_arrive(&sm.done[gs2]);
        }
        if (two) gs = gs2, gpar = gpar2;
        if (++gs == NS) gs = 0, gpar ^= 1u;
      }
    } else {
      while (true) {
        WsStage<DUAL>& st = sm.st[gs];
        mbar_wait_bounded(&sm.full[gs], gpar);
        const int n = st.n;
        if (n < 0) break;
        const int chunk = (n + MPI - 1) / MPI;
        const int m_end = min(n, (ms + 1) * chunk);
        for (int it = 0; it < chunk; ++it) {
          const int m = ms * chunk + it;
          const bool on = m < m_end && role_ok;
          load(st, m, on, qa, ka0, ka7);
          maths(st, m, on, qa);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.done[gs]);
        if (++gs == NS) gs = 0, gpar ^= 1u;
      }
    }
    return;
  }

  // ============================================================================= producer: march + composite
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
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
  const unsigned stride4 = (unsigned)g.stride >> 2;
  const DepthTab dt{ztab, s.dg.perturb, s.dg.jit, s.dg.key};
  const bool use_quads = g.quads != nullptr;

  float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, acc = 0.f;
  float cdr = 0.f, cdg = 0.f, cdb = 0.f;
  int ps = 0, cs = 0, pending = 0;  // stage to publish next, stage to composite next, published but not composited
  unsigned cpar = 0u;

  auto composite = [&]() {
    WsStage<DUAL>& st = sm.st[cs];
    mbar_wait_bounded(&sm.done[cs], cpar);
    const unsigned sact = st.act;
    if ((sact >> lane) & 1u) {
      int rank = __popc(sact & ((1u << lane) - 1u));
      const float4 pd = st.pend[SORT ? lane : rank];
      if constexpr (SORT) rank = (int)__float_as_uint(pd.w);
      const float4 raw = st.R[rank];
      const float sigma = pd.x, z = pd.y, alpha = pd.z;
      const float w = alpha * T;
      const float sr = sigmoidf_(raw.x), sg = sigmoidf_(raw.y), sb2 = sigmoidf_(raw.z);
      if (out.cache) out.cache[(size_t)st.step * rp.n + ray] = make_float4(sr, sg, sb2, sigma);
      cr = fmaf(w, sr, cr);
      cg = fmaf(w, sg, cg);
      cb = fmaf(w, sb2, cb);
      if constexpr (DUAL) {
        const float4 raw2 = DEG > 0 ? st.R2[rank] : raw;
        const float dr = sigmoidf_(raw2.x), dg_ = sigmoidf_(raw2.y), db = sigmoidf_(raw2.z);
        if (out.cache_diffuse) out.cache_diffuse[(size_t)st.step * rp.n + ray] = make_float4(dr, dg_, db, 0.0f);
        cdr = fmaf(w, dr, cdr);
        cdg = fmaf(w, dg_, cdg);
        cdb = fmaf(w, db, cdb);
      }
      dep = fmaf(w, z, dep);
      acc += w;
      T *= (1.0f - alpha);
      if (T == 0.0f) marching = false;  // every later weight is alpha * 0 = 0 exactly
    }
    if (++cs == NS) cs = 0, cpar ^= 1u;
    --pending;
  };

  float z = 0.f;
  bool have_z = false;
  DepthMarch dm;
  dm.bm = dm.bc = 0.f;
  for (int i = lo; i <= hi; ++i) {
    bool contributes = false;
    float sigma = 0.f, zn = 0.f;
    bool last = false;
    Cell cell;
    CellQ cq;
    const bool mine = marching && i >= s.i_lo && i <= s.i_hi;
    if (mine) {
      last = (i == c.S - 1);
      if (use_tab) {
        if (!have_z) z = dt.at(i), have_z = true;
        zn = last ? 0.0f : dt.at(i + 1);
      } else {
        if (!have_z) dm.start(s.dg, i), z = dm.next(s.dg, i), have_z = true;
        zn = last ? 0.0f : dm.next(s.dg, i + 1);
      }
      const float px = __fadd_rn(r.ox, __fmul_rn(r.dx, z));
      const float py = __fadd_rn(r.oy, __fmul_rn(r.dy, z));
      const float pz = __fadd_rn(r.oz, __fmul_rn(r.dz, z));
      if (inside_aabb(g, px, py, pz)) {
        float dpost;
        if (use_quads) {
          make_cell_q(g, px, py, pz, cq);
          sigma = density_post(g.post, density_pre_interp_q(g, cq), dpost);
        } else {
          make_cell_inside(g, px, py, pz, cell);
          sigma = density_post(g.post, density_pre_interp<true>(g, cell), dpost);
        }
        contributes = sigma != 0.0f;
      }
    }
    const unsigned act = __ballot_sync(FULL, contributes);
    if (out.mask && lane == 0) out.mask[(size_t)i * (gridDim.x * 4u) + (blockIdx.x * 4u + warp)] = act;
    if (act != 0u) {
      WsStage<DUAL>& st = sm.st[ps];
      int rank = __popc(act & ((1u << lane) - 1u));
      if (contributes && use_quads) cell_from_q(g, cq, cell);
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
        const int base = __shfl_sync(FULL, scan - own, leader);  // exclusive prefix at the leader
        rank = base + __popc(peers & ((1u << lane) - 1u));
      }
      if (contributes) {
        float wc[8];
        unsigned rec4[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ix = k >> 2, iy = (k >> 1) & 1, iz = k & 1;
          wc[k] = cell.wx[ix] * cell.wy[iy] * cell.wz[iz];
          rec4[k] = (unsigned)(cell.ox[ix] + cell.oy[iy] + cell.oz[iz]) * stride4;
        }
        if constexpr (PF != 0) {
          // a record is F floats at a 16-byte aligned offset: its first and its last byte name the (at most two) 128-byte lines
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const char* rec = reinterpret_cast<const char*>(g.feat) + 16ull * rec4[k];
            if constexpr (PF == 1) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rec));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (4 * F - 4)));
            } else {
              asm volatile("prefetch.global.L1 [%0];" ::"l"(rec));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (4 * F - 4)));
            }
          }
        }
        st.Wlo[rank] = make_float4(wc[0], wc[1], wc[2], wc[3]);
        st.Whi[rank] = make_float4(wc[4], wc[5], wc[6], wc[7]);
        st.Vlo[rank] = make_uint4(rec4[0], rec4[1], rec4[2], rec4[3]);
        st.Vhi[rank] = make_uint4(rec4[4], rec4[5], rec4[6], rec4[7]);
        st.yoff[rank] = (unsigned)lane * (unsigned)(H::YROW * 4);
        const float delta = last ? __fmul_rn(kInfinity, s.dnorm) : __fmul_rn(__fsub_rn(zn, z), s.dnorm);
        st.pend[SORT ? lane : rank] = make_float4(sigma, z, 1.0f - exp_neg(sigma * delta), __uint_as_float((unsigned)rank));
      }
      if (lane == 0) st.n = __popc(act), st.step = i, st.act = act;
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full[ps]);
      if (++ps == NS) ps = 0;
      ++pending;
      if (pending > kWsLag) composite();
    }
    if (mine) z = zn;
  }
  while (pending > 0) composite();
  if (lane == 0) sm.st[ps].n = -1;  // every published stage has been composited: this one is free
  __syncwarp();
  if (lane == 0) mbar_arrive(&sm.full[ps]);

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

}  // namespace r3d
