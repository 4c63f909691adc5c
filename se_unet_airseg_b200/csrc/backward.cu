// HBM-bound backward kernels: fused sSE-gate / LeakyReLU / InstanceNorm backward (two passes around the
// per-(n,c) reductions), CAT-block backward with max-pool routing and the analytic injection branch,
// adjoint trilinear up-sampling, head adjoint and the small-parameter gradient assembly.
#include "backward.cuh"
#include <algorithm>
#include <cstring>
#include <cstdint>
#include <cstdlib>

constexpr float kInEps = 1e-5f;

__device__ __forceinline__ float lrelu_grad(float n) { return n > 0.f ? 1.f : 0.01f; }
__device__ __forceinline__ void atomic_max_pos(unsigned int* p, float v) { atomicMax(p, __float_as_uint(v)); }

// =============================================================================================
// SSE block backward, pass A
// =============================================================================================
// RING > 0 (round 2, late): the inputs of a warp's next RING voxel spans travel as bulk copies (cp.async.bulk, one mbarrier
// per warp and ring slot) into a warp-private shared-memory ring instead of as register prefetch.  ncu on the register
// version: 25 % occupancy (128 registers), a third of all stall samples on the first use of the prefetched chunk (long
// scoreboard) - 26 KB in flight per SM cap the read rate near 3 TB/s and a second group in registers spills.  A ring slot
// costs no registers.  One ELECTED lane issues the copies with warp-uniform addresses (the first ring version let 2 LPV + 1
// lanes issue one row each: UBLKCP takes uniform registers, so ptxas serialised the lanes in a waterfall loop, +45 % warp
// instructions, and the kernel got slower although the long-scoreboard stalls were gone); a slot holds a SPAN of 1-2 voxel
// groups so that a row is >= 128 B and there are half as many copies.
template <int C> struct SseRing {
  static constexpr int LPV = C / 8, VPW = 32 / LPV;
  static constexpr int SPAN = C >= 32 ? 2 : 1;                // voxel groups per slot
  static constexpr int SVX = SPAN * VPW;                      // voxels per slot
  static constexpr int GB = 8 * (int)sizeof(grad_t);          // bytes of one gradient chunk
  static constexpr int RS_RAW = SVX * 16 + 128 / LPV;         // row = one chunk plane of the span; the pad spreads the LPV rows over the banks
  static constexpr int RS_DE0 = SVX * GB + 16;
  static constexpr int DE0_OFF = LPV * RS_RAW;
  static constexpr int DT_OFF = DE0_OFF + LPV * RS_DE0;
  static constexpr int STAGE = (DT_OFF + SVX * 4 + 127) / 128 * 128;
};
constexpr int kSseRingDepth = 3;

template <int C, int GATES, int PF, int RING>
__global__ void __launch_bounds__(256, 2) sse_bwd_a_kernel(const __grid_constant__ SseBwdArgs a) {
  constexpr int LPV = C / 8;     // lanes cooperating on one voxel (one 8-channel chunk each)
  constexpr int VPW = 32 / LPV;  // voxels per warp
  extern __shared__ __align__(128) uint8_t s_ring[];          // RING > 0: [8 warps][RING][SseRing<C>::STAGE]
  __shared__ __align__(8) unsigned long long s_bar[8 * (RING > 0 ? RING : 1)];
  if (RING > 0) {
    if (threadIdx.x < 8 * RING) mbar_init(smem_u32(&s_bar[threadIdx.x]), 1);
    fence_mbar_init();        // (the __syncthreads() after the statistics prologue publishes the barriers)
  }
  __shared__ float s_mean[C], s_rstd[C], s_wse[C], s_wse2[C], s_weff[C];
  __shared__ float s_red[8][5][C];
  __shared__ float s_cst[8], s_max[8];
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s = a.stats[((size_t)n * a.stats_c + c) * 2], q = a.stats[((size_t)n * a.stats_c + c) * 2 + 1];
    const double mean = s / (double)a.V;
    double var = q / (double)a.V - mean * mean;
    if (var < 0) var = 0;
    s_mean[c] = (float)mean;
    s_rstd[c] = (float)(1.0 / sqrt(var + (double)kInEps));
    s_wse[c] = a.wse[c];
    s_wse2[c] = GATES == 2 ? a.wse2[c] : 0.f;
    s_weff[c] = a.weff[(size_t)n * 64 + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (bulk-copy addresses live in uniform registers)
  const int k = lane % LPV, vsub = lane / LPV;
  float S1[8], S2[8], Wse[8], Wse2[8], Weff[8], cst = 0.f, mx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) S1[i] = S2[i] = Wse[i] = Wse2[i] = Weff[i] = 0.f;
  auto vreduce = [&](float v) {   // sum over the LPV lanes of one voxel
#pragma unroll
    for (int o = 1; o < LPV; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  // Software pipeline: the loads of the next TWO voxel groups are in flight during the arithmetic of the current one.  The
  // kernel runs at 25 % occupancy (two blocks per SM), so the bytes in flight per warp decide the achieved bandwidth (ncu:
  // stalled on long-scoreboard, 2.9 TB/s without prefetch; with ONE group ahead 512 threads x 52 B = 26 KB per SM are in
  // flight, which caps the read rate near 3 TB/s - the 64-channel instances sat at 2.1 TB/s, the 32-channel ones at 4.4).
  const long long vstep = (long long)gridDim.x * 8 * VPW;
  const act_t* rawp = a.raw + ((size_t)n * a.raw_chunks + k) * a.V * 8;
  const grad_t* de0p = a.dE0 ? a.dE0 + ((size_t)n * a.dE0_chunks + a.dE0_off + k) * a.V * 8 : nullptr;
  const float* dTp = a.dT + (size_t)n * a.V;
  struct Pre { Chunk8 raw; float de0[8]; float dT; };
  auto prefetch = [&](Pre& p, long long v) {
    if (v >= a.V) {   // tail of a level whose voxel count is not a multiple of the warp's group (coarsest level of ragged shapes)
      p.raw.u[0] = p.raw.u[1] = p.raw.u[2] = p.raw.u[3] = 0u;
      p.dT = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) p.de0[i] = 0.f;
      return;
    }
    p.raw = ld_chunk_stream(rawp + (size_t)v * 8);
    p.dT = dTp[v];
    if (de0p) ld_grad8(de0p + (size_t)v * 8, p.de0);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p.de0[i] = 0.f;
    }
  };
  auto process = [&](const Pre& p, long long v, const bool live) {   // live: v < V (the lanes of a dead voxel only feed each other)
    float f[8], nn[8], av[8], e0[8], de0[8];
    chunk_to_floats(p.raw, f);
    const float dT = p.dT;
#pragma unroll
    for (int i = 0; i < 8; ++i) de0[i] = p.de0[i];
    float p1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = k * 8 + i;
      nn[i] = (f[i] - s_mean[c]) * s_rstd[c];
      av[i] = lrelu_(nn[i]);
      p1 = fmaf(s_wse[c], av[i], p1);
    }
    const float g1 = sigmoidf_(vreduce(p1));
    float a1[8], g2 = 1.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a1[i] = av[i] * g1;
    if (GATES == 2) {
      float p2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) p2 = fmaf(s_wse2[k * 8 + i], a1[i], p2);
      g2 = sigmoidf_(vreduce(p2));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) e0[i] = a1[i] * g2;
#pragma unroll
    for (int i = 0; i < 8; ++i) de0[i] = fmaf(s_weff[k * 8 + i], dT, de0[i]);
    float da1[8], k2 = 0.f;
    if (GATES == 2) {
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s2 = fmaf(de0[i], a1[i], s2);
      k2 = vreduce(s2) * g2 * (1.f - g2);
#pragma unroll
      for (int i = 0; i < 8; ++i) da1[i] = fmaf(s_wse2[k * 8 + i], k2, de0[i] * g2);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) da1[i] = de0[i];
    }
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s1 = fmaf(da1[i], av[i], s1);
    const float k1 = vreduce(s1) * g1 * (1.f - g1);
    float dn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = k * 8 + i;
      const float da = fmaf(s_wse[c], k1, da1[i] * g1);
      dn[i] = da * lrelu_grad(nn[i]);
      if (live) {
        S1[i] += dn[i];
        S2[i] = fmaf(dn[i], nn[i], S2[i]);
        Wse[i] = fmaf(k1, av[i], Wse[i]);
        Wse2[i] = fmaf(k2, a1[i], Wse2[i]);
        Weff[i] = fmaf(dT, e0[i], Weff[i]);
        mx = fmaxf(mx, fabsf(dn[i] * s_rstd[c]));
      }
    }
    if (live) {
      if (k == 0) cst += dT;
      st_grad8(a.dn + (((size_t)n * a.dn_chunks + k) * a.V + v) * 8, dn);
    }
  };
  if constexpr (RING > 0) {
    using R = SseRing<C>;
    const uint32_t ring0 = smem_u32(s_ring) + (uint32_t)(warp * RING * R::STAGE);
    const uint32_t bar0 = smem_u32(&s_bar[warp * RING]);
    const long long sstep = (long long)gridDim.x * 8 * R::SVX;
    const long long vs0 = ((long long)blockIdx.x * 8 + warp) * R::SVX;
    const act_t* raw_n = a.raw + (size_t)n * a.raw_chunks * a.V * 8;                                   // chunk plane 0 of the sample
    const grad_t* de0_n = a.dE0 ? a.dE0 + ((size_t)n * a.dE0_chunks + a.dE0_off) * a.V * 8 : nullptr;
    const uint32_t span_bytes = (uint32_t)(LPV * R::SVX * 16 + R::SVX * 4) + (de0_n ? (uint32_t)(LPV * R::SVX * R::GB) : 0u);
    auto fill = [&](int slot, long long vs) {     // whole warp, converged; one elected lane issues the 2 LPV + 1 row copies
      if (elect_one_sync()) {
        const uint32_t bar = bar0 + 8u * (uint32_t)slot;
        const uint32_t dst = ring0 + (uint32_t)(slot * R::STAGE);
        mbar_expect_tx(bar, span_bytes);
        const act_t* rp = raw_n + (size_t)vs * 8;
#pragma unroll
        for (int kk = 0; kk < LPV; ++kk) bulk_g2s(dst + (uint32_t)(kk * R::RS_RAW), rp + (size_t)kk * a.V * 8, (uint32_t)(R::SVX * 16), bar);
        if (de0_n) {
          const grad_t* gp = de0_n + (size_t)vs * 8;
#pragma unroll
          for (int kk = 0; kk < LPV; ++kk)
            bulk_g2s(dst + (uint32_t)(R::DE0_OFF + kk * R::RS_DE0), gp + (size_t)kk * a.V * 8, (uint32_t)(R::SVX * R::GB), bar);
        }
        bulk_g2s(dst + (uint32_t)R::DT_OFF, dTp + vs, (uint32_t)(R::SVX * 4), bar);
      }
      __syncwarp();
    };
#pragma unroll
    for (int j = 0; j < RING; ++j)
      if (vs0 + j * sstep < a.V) fill(j, vs0 + j * sstep);
    int slot = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (long long vs = vs0; vs < a.V; vs += sstep) {
      mbar_wait(bar0 + 8u * (uint32_t)slot, phase);
      const uint32_t src = ring0 + (uint32_t)(slot * R::STAGE);
#pragma unroll
      for (int s = 0; s < R::SPAN; ++s) {
        const int vl = s * VPW + vsub;      // voxel inside the span
        Pre cur;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(cur.raw.u[0]), "=r"(cur.raw.u[1]), "=r"(cur.raw.u[2]), "=r"(cur.raw.u[3])
                     : "r"(src + (uint32_t)(k * R::RS_RAW + vl * 16)));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cur.dT) : "r"(src + (uint32_t)(R::DT_OFF + vl * 4)));
        if (de0_n) {
          const uint32_t ga = src + (uint32_t)(R::DE0_OFF + k * R::RS_DE0 + vl * R::GB);
#ifdef SEUNET_GRAD_BF16
          Chunk8 gc;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(gc.u[0]), "=r"(gc.u[1]), "=r"(gc.u[2]), "=r"(gc.u[3]) : "r"(ga));
          chunk_to_floats_bf16(gc, cur.de0);
#else
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cur.de0[0]), "=f"(cur.de0[1]), "=f"(cur.de0[2]), "=f"(cur.de0[3]) : "r"(ga));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cur.de0[4]), "=f"(cur.de0[5]), "=f"(cur.de0[6]), "=f"(cur.de0[7]) : "r"(ga + 16u));
#endif
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) cur.de0[i] = 0.f;
        }
        process(cur, vs + vl, true);      // the ring path is only launched when V is a multiple of 32 (whole spans)
      }
      // refill the slot only after its values have been consumed (generic-proxy reads before the async-proxy write)
      __syncwarp();
      if (vs + RING * sstep < a.V) fill(slot, vs + RING * sstep);
      if (++slot == RING) { slot = 0; phase ^= 1u; }
    }
  } else {
    const long long vb0 = ((long long)blockIdx.x * 8 + warp) * VPW;
    Pre pre[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j)
      if (vb0 + j * vstep < a.V) prefetch(pre[j], vb0 + j * vstep + vsub);
#pragma unroll 1
    for (long long vb = vb0; vb < a.V; vb += vstep) {
      const Pre cur = pre[0];
#pragma unroll
      for (int j = 0; j + 1 < PF; ++j) pre[j] = pre[j + 1];   // (13 register moves per ~200-instruction body)
      if (vb + PF * vstep < a.V) prefetch(pre[PF - 1], vb + PF * vstep + vsub);
      process(cur, vb + vsub, vb + vsub < a.V);
    }
  }
  // reduce over the voxel sub-lanes of the warp, then over warps, then one atomic per value per block
  auto wreduce = [&](float v) {
#pragma unroll
    for (int o = LPV; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float r0 = wreduce(S1[i]), r1 = wreduce(S2[i]), r2 = wreduce(Wse[i]), r3 = wreduce(Wse2[i]), r4 = wreduce(Weff[i]);
    if (vsub == 0) {
      s_red[warp][0][k * 8 + i] = r0; s_red[warp][1][k * 8 + i] = r1; s_red[warp][2][k * 8 + i] = r2;
      s_red[warp][3][k * 8 + i] = r3; s_red[warp][4][k * 8 + i] = r4;
    }
  }
  cst = warp_sum(cst);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) { s_cst[warp] = cst; s_max[warp] = mx; }
  __syncthreads();
  for (int t = threadIdx.x; t < 5 * C; t += blockDim.x) {
    const int q = t / C, c = t % C;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w][q][c];
    if (q < 2) atomicAdd(a.redS + ((size_t)n * 64 + c) * 2 + q, (double)s);
    else if (q == 2) atomicAdd(a.dwse + c, s);
    else if (q == 3) { if (GATES == 2) atomicAdd(a.dwse2 + c, s); }
    else atomicAdd(a.dweff + (size_t)n * 64 + c, s);
  }
  if (threadIdx.x == 0) {
    float s = 0.f, m = 0.f;
    for (int w = 0; w < 8; ++w) { s += s_cst[w]; m = fmaxf(m, s_max[w]); }
    atomicAdd(a.dcst + n, s);
    atomic_max_pos(a.dymax, m);
  }
}

template <int C> constexpr bool kSseBwdDeep = false;   // measured: depth 2 spills (128 registers) and is 10-30 % slower

template <int C>
static int launch_sse_bwd_a_c(const SseBwdArgs& a, cudaStream_t st) {
  static const bool ring_env = !(getenv("SEUNET_BWDA_RING") && atoi(getenv("SEUNET_BWDA_RING")) == 0);   // A/B knob
  static const int pf_env = getenv("SEUNET_BWDA_PF") ? atoi(getenv("SEUNET_BWDA_PF")) : 0;
  const bool deep = pf_env ? pf_env == 2 : kSseBwdDeep<C>;
  // bulk copies need 16-byte aligned sources (plan buffers always are; dT of level 0 is the caller's dpred tensor)
  const bool aligned = ((((uintptr_t)a.raw) | ((uintptr_t)a.dE0) | ((uintptr_t)a.dT)) & 15u) == 0 && a.V % 32 == 0;
  const bool ring = ring_env && !deep && aligned;
  const int VPB = 8 * (ring ? SseRing<C>::SVX : 32 / (C / 8));   // voxels per block and loop iteration
  const long long need = (a.V + VPB - 1) / VPB;
  // Every block ends with 5*C same-address atomics per sample: at the coarse levels (few voxels) thousands of one-iteration
  // blocks spent their time in that tail and in the per-block statistics prologue.  Give each warp >= 16 voxel groups, but
  // keep >= 4 blocks per SM in flight over the whole batch.
  const long long floor_blocks = (148 * 4 + a.N - 1) / a.N;
  const long long gx = std::min<long long>(need, std::max<long long>(floor_blocks, std::min<long long>(148 * 8, need / 16)));
  dim3 grid((unsigned)gx, a.N);
  if (ring) {
    constexpr int smem = 8 * kSseRingDepth * SseRing<C>::STAGE;
    static bool attr_set[64] = {};   // per-device function attribute (see conv_launch_t)
    int dev = 0;
    SEUNET_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      SEUNET_CUDA_CHECK(cudaFuncSetAttribute(sse_bwd_a_kernel<C, 1, 1, kSseRingDepth>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      SEUNET_CUDA_CHECK(cudaFuncSetAttribute(sse_bwd_a_kernel<C, 2, 1, kSseRingDepth>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (a.wse2) sse_bwd_a_kernel<C, 2, 1, kSseRingDepth><<<grid, 256, smem, st>>>(a);
    else sse_bwd_a_kernel<C, 1, 1, kSseRingDepth><<<grid, 256, smem, st>>>(a);
  } else if (deep) {   // register prefetch two groups ahead (A/B, tools/r02_call50.sh: spills, slower)
    if (a.wse2) sse_bwd_a_kernel<C, 2, 2, 0><<<grid, 256, 0, st>>>(a);
    else sse_bwd_a_kernel<C, 1, 2, 0><<<grid, 256, 0, st>>>(a);
  } else {
    if (a.wse2) sse_bwd_a_kernel<C, 2, 1, 0><<<grid, 256, 0, st>>>(a);
    else sse_bwd_a_kernel<C, 1, 1, 0><<<grid, 256, 0, st>>>(a);
  }
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int launch_sse_bwd_a(int C, const SseBwdArgs& a, cudaStream_t st) {
  switch (C) {
    case 8: return launch_sse_bwd_a_c<8>(a, st);
    case 16: return launch_sse_bwd_a_c<16>(a, st);
    case 32: return launch_sse_bwd_a_c<32>(a, st);
    case 64: return launch_sse_bwd_a_c<64>(a, st);
  }
  seunet_set_error("sse_bwd_a: C=%d unsupported", C);
  return 1;
}

// =============================================================================================
// InstanceNorm backward, pass B (shared by SSE and CAT blocks)
// =============================================================================================
__device__ __forceinline__ float dy_scale_from_max(unsigned int bits) {
  const float mx = __uint_as_float(bits);
  if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
  float e = 8.f - ceilf(log2f(mx));          // largest |dn*rstd| lands in [2^7, 2^8]
  e = fminf(fmaxf(e, -100.f), 100.f);
  return exp2f(e);
}

// VPT voxel groups per block (rolled loop): the fp64 prologue + barrier of a block is as long as one load/store round trip, so
// where enough blocks remain it is paid once per four groups (same finding as apply_sse in pointwise.cu): pass B total
// 4.87 -> 3.80 ms at 8 x 128^3 (tools/r02_call72.sh); eight or sixteen groups per block: no further gain.
template <int VPT>
__global__ void __launch_bounds__(256) norm_bwd_b_kernel(const __grid_constant__ NormBwdArgs a) {
  __shared__ float s_mean[8], s_rstd[8], s_m1[8], s_m2[8];
  __shared__ float s_scale;
  const int n = blockIdx.z, k = blockIdx.y;
  if (threadIdx.x < 8) {
    const int c = k * 8 + threadIdx.x;
    const double s = a.stats[((size_t)n * a.stats_c + c) * 2], q = a.stats[((size_t)n * a.stats_c + c) * 2 + 1];
    const double mean = s / (double)a.V;
    double var = q / (double)a.V - mean * mean;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)kInEps));
    s_m1[threadIdx.x] = (float)(a.redS[((size_t)n * 64 + c) * 2] / (double)a.V);
    s_m2[threadIdx.x] = (float)(a.redS[((size_t)n * 64 + c) * 2 + 1] / (double)a.V);
  }
  if (threadIdx.x == 8) {
    const float sc = dy_scale_from_max(*a.dymax);
    s_scale = sc;
    if (blockIdx.x == 0 && k == 0 && n == 0) { a.scale_out[0] = sc; a.scale_out[1] = 1.f / sc; }
  }
  __syncthreads();
  const float sc = s_scale;
#pragma unroll 1
  for (int u = 0; u < VPT; ++u) {
    const long long v = ((long long)blockIdx.x * VPT + u) * blockDim.x + threadIdx.x;
    if (v >= a.V) return;
    float f[8], dn[8], dy[8];
    chunk_to_floats(ld_chunk_stream(a.raw + (((size_t)n * a.raw_chunks + k) * a.V + v) * 8), f);
    ld_grad8(a.dn + (((size_t)n * a.dn_chunks + k) * a.V + v) * 8, dn);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float nn = (f[i] - s_mean[i]) * s_rstd[i];
      float t = s_rstd[i] * (dn[i] - s_m1[i] - nn * s_m2[i]) * sc;
      dy[i] = fminf(fmaxf(t, -60000.f), 60000.f);
    }
    st_chunk(a.dy + (((size_t)n * a.dy_chunks + k) * a.V + v) * 8, floats_to_chunk(dy));
  }
}

int launch_norm_bwd_b(const NormBwdArgs& a, cudaStream_t st) {
  static const int vpt_env = getenv("SEUNET_BWDB_VPT") ? atoi(getenv("SEUNET_BWDB_VPT")) : 4;   // A/B knob
  const long long groups = (a.V + 255) / 256;
  if (vpt_env >= 4 && groups / 4 * (a.C / 8) * a.N >= 148 * 8) {
    dim3 grid((unsigned)((groups + 3) / 4), a.C / 8, a.N);
    norm_bwd_b_kernel<4><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid((unsigned)groups, a.C / 8, a.N);
    norm_bwd_b_kernel<1><<<grid, 256, 0, st>>>(a);
  }
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// CAT block backward, pass A:  out = lrelu(IN(y)) [+ lrelu(IN(Wx x))], optional 2x2x2 max-pool fan-out
// =============================================================================================
struct __align__(16) CatCoef { float mean[8], rstd[8], mx[8], rx[8], w0[8], w1[8]; };

__device__ __forceinline__ void cat_prologue(const act_t*, const double* stats, int stats_c, long long V, int n, int k, bool hasx,
                                             const float* wx, int in_ch, const double* mom, CatCoef* s) {
  if (threadIdx.x < 8) {
    const int c = k * 8 + threadIdx.x;
    const double su = stats[((size_t)n * stats_c + c) * 2], q = stats[((size_t)n * stats_c + c) * 2 + 1];
    const double mean = su / (double)V;
    double var = q / (double)V - mean * mean;
    if (var < 0) var = 0;
    s->mean[threadIdx.x] = (float)mean;
    s->rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)kInEps));
    if (hasx) {
      const double* m = mom + (size_t)n * kMomStride;
      const double mu0 = m[0] / V, mu1 = m[1] / V;
      const double c00 = m[2] / V - mu0 * mu0, c11 = m[3] / V - mu1 * mu1, c01 = m[4] / V - mu0 * mu1;
      const double w0 = wx[c * in_ch], w1 = in_ch > 1 ? wx[c * in_ch + 1] : 0.0;
      double vx = w0 * w0 * c00 + w1 * w1 * c11 + 2.0 * w0 * w1 * c01;
      if (vx < 0) vx = 0;
      s->mx[threadIdx.x] = (float)(w0 * mu0 + w1 * mu1);
      s->rx[threadIdx.x] = (float)(1.0 / sqrt(vx + (double)kInEps));
      s->w0[threadIdx.x] = (float)w0;
      s->w1[threadIdx.x] = (float)w1;
    }
  }
}

template <bool HASX, bool POOL>
__global__ void __launch_bounds__(256, 2) cat_bwd_a_kernel(const __grid_constant__ CatBwdArgs a) {
  __shared__ CatCoef s;
  __shared__ float s_red[8][32];
  __shared__ float s_max[8];
  const int n = blockIdx.z, k = blockIdx.y;
  const long long V = dims_vox(a.d);
  cat_prologue(a.raw, a.stats, a.stats_c, V, n, k, HASX, a.wx, a.in_ch, a.mom, &s);
  __syncthreads();
  const act_t* rawp = a.raw + ((size_t)n * a.raw_chunks + k) * V * 8;
  const grad_t* gp_ = a.g + ((size_t)n * a.g_chunks + a.g_off + k) * V * 8;
  grad_t* dnp = a.dn + ((size_t)n * a.dn_chunks + k) * V * 8;
  float red[32];   // [0..7] S1, [8..15] S2, [16..23] Sx1, [24..31] Sx2
#pragma unroll
  for (int i = 0; i < 32; ++i) red[i] = 0.f;
  float mxv = 0.f;
  auto xvals = [&](int dz, int hy, int wx, float& x0, float& x1) {
    x0 = 0.f; x1 = 0.f;
    if (HASX) {
      const float* xp = a.x + (a.xo.use ? a.xo.off[n] : n * a.xs[0]) + dz * a.xs[2] + hy * a.xs[3] + wx * a.xs[4];
      x0 = xp[0];
      if (a.in_ch > 1) x1 = xp[a.xs[1]];
    }
  };
  // gradient bookkeeping of one voxel given its total output gradient G[8]
  auto accumulate = [&](long long v, const float* ny, const float* nx, const float* G) {
    float dn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dn[i] = G[i] * lrelu_grad(ny[i]);
      red[i] += dn[i];
      red[8 + i] = fmaf(dn[i], ny[i], red[8 + i]);
      mxv = fmaxf(mxv, fabsf(dn[i] * s.rstd[i]));
      if (HASX) {
        const float dx = G[i] * lrelu_grad(nx[i]);
        red[16 + i] += dx;
        red[24 + i] = fmaf(dx, nx[i], red[24 + i]);
      }
    }
    st_grad8(dnp + (size_t)v * 8, dn);
  };
  auto norms = [&](long long v, int dz, int hy, int wx, float* ny, float* nx) {
    float f[8];
    chunk_to_floats(ld_chunk(rawp + (size_t)v * 8), f);
    float x0, x1;
    xvals(dz, hy, wx, x0, x1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ny[i] = (f[i] - s.mean[i]) * s.rstd[i];
      nx[i] = HASX ? (fmaf(s.w0[i], x0, s.w1[i] * x1) - s.mx[i]) * s.rx[i] : 0.f;
    }
  };
  // blockIdx.x walks d-planes (pooled planes when POOL); a thread owns one (h, w) position of the plane - 32-bit index
  // arithmetic, fully coalesced 16/32-byte accesses (the 64-bit div/mod per voxel of the first version cost more than
  // the memory traffic).
  const int W = a.d.W, H = a.d.H;
  if (!POOL) {
    for (int dz = blockIdx.x; dz < a.d.D; dz += gridDim.x)
      for (int t = threadIdx.x; t < H * W; t += blockDim.x) {
        const int hy = t / W, wx = t - hy * W;
        const long long v = ((long long)dz * H + hy) * W + wx;
        float ny[8], nx[8], G[8];
        norms(v, dz, hy, wx, ny, nx);
        ld_grad8(gp_ + (size_t)v * 8, G);
        accumulate(v, ny, nx, G);
      }
  } else {
    // POOL: a thread owns the 2 (h) voxels of a pooling window at ONE (d, w); the four lanes l, l^1 (w pair; W is even) and
    // l^16 (d pair) settle the window's arg-max with two shuffle rounds.  Scan index inside the window: J = 4*dd + 2*dh + dw
    // (max_pool3d keeps the first maximum in (d,h,w) order).
    // Round 2: every load of the thread is issued up front - the two raw chunks stay packed in registers for both passes, the
    // x values and the pooled gradient ride along, the second output gradient is fetched while the first is processed.  (The
    // first version walked a 2x2 (d,h) window twice with one load in flight per thread: ncu 2.4 TB/s, 58 % of the stall
    // samples on that load; with all four voxels unrolled in one thread the x-branch instance spilled 500 bytes.)
    const int Dp = a.d.D >> 1, Hp = H >> 1, Wp = W >> 1;
    const long long Vp = (long long)Dp * Hp * Wp;
    const grad_t* gpool = a.gp + ((size_t)n * a.gp_chunks + a.gp_off + k) * Vp * 8;
    const int per_plane = Hp * W, groups = (per_plane + 15) >> 4;
    for (int pd = blockIdx.x; pd < Dp; pd += gridDim.x)
      for (int t = threadIdx.x; t < groups * 32; t += blockDim.x) {
        const int l = t & 31, dd = l >> 4, idx = (t >> 5) * 16 + (l & 15);
        const bool live = idx < per_plane;
        const int ii = live ? idx : 0;
        const int ph = ii / W, wx = ii - ph * W, dz = pd * 2 + dd;
        Chunk8 rc[2];
        float xa[2], xb[2];
        long long vj[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int hy = ph * 2 + j;
          vj[j] = ((long long)dz * H + hy) * W + wx;
          rc[j] = ld_chunk(rawp + (size_t)vj[j] * 8);
          xvals(dz, hy, wx, xa[j], xb[j]);
        }
        float GP[8], Gn[8];
        ld_grad8_cached(gpool + (size_t)(((long long)pd * Hp + ph) * Wp + (wx >> 1)) * 8, GP);   // the four lanes of a window read the same chunk
        ld_grad8(gp_ + (size_t)vj[0] * 8, Gn);
        auto norms_r = [&](int j, float* ny, float* nx) {
          float f[8];
          chunk_to_floats(rc[j], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            ny[i] = (f[i] - s.mean[i]) * s.rstd[i];
            nx[i] = HASX ? (fmaf(s.w0[i], xa[j], s.w1[i] * xb[j]) - s.mx[i]) * s.rx[i] : 0.f;
          }
        };
        float best[8];
        unsigned argp = 0;     // 4 bits per channel: scan index J of the running maximum
#pragma unroll
        for (int i = 0; i < 8; ++i) best[i] = -INFINITY;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float ny[8], nx[8];
          norms_r(j, ny, nx);
          const unsigned J = (unsigned)(dd * 4 + j * 2 + (wx & 1));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float o = lrelu_(ny[i]) + (HASX ? lrelu_(nx[i]) : 0.f);
            if (o > best[i]) { best[i] = o; argp = (argp & ~(0xFu << (4 * i))) | (J << (4 * i)); }
          }
        }
#pragma unroll
        for (int m = 1; m <= 16; m <<= 4) {           // partner along w, then partner along d
          const unsigned oap = __shfl_xor_sync(0xffffffffu, argp, m);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ob = __shfl_xor_sync(0xffffffffu, best[i], m);
            const unsigned ma = (argp >> (4 * i)) & 0xFu, oa = (oap >> (4 * i)) & 0xFu;
            if (ob > best[i] || (ob == best[i] && oa < ma)) { best[i] = ob; argp = (argp & ~(0xFu << (4 * i))) | (oa << (4 * i)); }
          }
        }
        if (!live) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const unsigned J = (unsigned)(dd * 4 + j * 2 + (wx & 1));
          float ny[8], nx[8], G[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) G[i] = Gn[i];
          if (j == 0) ld_grad8(gp_ + (size_t)vj[1] * 8, Gn);
          norms_r(j, ny, nx);
#pragma unroll
          for (int i = 0; i < 8; ++i) G[i] += (((argp >> (4 * i)) & 0xFu) == J) ? GP[i] : 0.f;
          accumulate(vj[j], ny, nx, G);
        }
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float tot = warp_xreduce32(red, lane);
  s_red[warp][lane] = tot;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mxv = fmaxf(mxv, __shfl_xor_sync(0xffffffffu, mxv, o));
  if (lane == 0) s_max[warp] = mxv;
  __syncthreads();
  if (threadIdx.x < 32) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_red[w][threadIdx.x];
    const int q = threadIdx.x >> 3, i = threadIdx.x & 7, c = k * 8 + i;
    if (q < 2) atomicAdd(a.redS + ((size_t)n * 64 + c) * 2 + q, (double)sum);
    else if (HASX) atomicAdd(a.redSx + ((size_t)n * 64 + c) * 2 + (q - 2), (double)sum);
  }
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int w = 0; w < 8; ++w) m = fmaxf(m, s_max[w]);
    atomic_max_pos(a.dymax, m);
  }
}

int launch_cat_bwd_a(const CatBwdArgs& a, cudaStream_t st) {
  const long long V = dims_vox(a.d);
  const bool pool = a.gp != nullptr, hasx = a.x != nullptr;
  const int planes = pool ? a.d.D / 2 : a.d.D;
  dim3 grid((unsigned)std::min(planes, 148 * 2), a.C / 8, a.d.N);
  if (hasx && pool) cat_bwd_a_kernel<true, true><<<grid, 256, 0, st>>>(a);
  else if (hasx) cat_bwd_a_kernel<true, false><<<grid, 256, 0, st>>>(a);
  else if (pool) cat_bwd_a_kernel<false, true><<<grid, 256, 0, st>>>(a);
  else cat_bwd_a_kernel<false, false><<<grid, 256, 0, st>>>(a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// x-branch weight gradient (x33/x63/x93.conv1.weight): du = rx * (dnx - mean(dnx) - nx * mean(dnx*nx)), dWx = sum du x^T
__global__ void __launch_bounds__(256) cat_bwd_x_kernel(const __grid_constant__ CatBwdXArgs a) {
  __shared__ CatCoef s;
  __shared__ float s_m1[8], s_m2[8];
  __shared__ float s_red[8][32];
  const int n = blockIdx.z, k = blockIdx.y;
  const long long V = dims_vox(a.d);
  cat_prologue(a.raw, a.stats, a.stats_c, V, n, k, true, a.wx, a.in_ch, a.mom, &s);
  if (threadIdx.x >= 32 && threadIdx.x < 40) {
    const int i = threadIdx.x - 32, c = k * 8 + i;
    s_m1[i] = (float)(a.redSx[((size_t)n * 64 + c) * 2] / (double)V);
    s_m2[i] = (float)(a.redSx[((size_t)n * 64 + c) * 2 + 1] / (double)V);
  }
  __syncthreads();
  float red[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) red[i] = 0.f;
  const act_t* rawp = a.raw + ((size_t)n * a.raw_chunks + k) * V * 8;
  const grad_t* dnp = a.dn + ((size_t)n * a.dn_chunks + k) * V * 8;
  const int W = a.d.W, H = a.d.H;
  for (int dz = blockIdx.x; dz < a.d.D; dz += gridDim.x)
  for (int t = threadIdx.x; t < H * W; t += blockDim.x) {
    const int hy = t / W, wx = t - hy * W;
    const long long v = ((long long)dz * H + hy) * W + wx;
    float f[8], dn[8];
    chunk_to_floats(ld_chunk_stream(rawp + (size_t)v * 8), f);
    ld_grad8(dnp + (size_t)v * 8, dn);
    const float* xp = a.x + (a.xo.use ? a.xo.off[n] : n * a.xs[0]) + dz * a.xs[2] + hy * a.xs[3] + wx * a.xs[4];
    const float x0 = xp[0], x1 = a.in_ch > 1 ? xp[a.xs[1]] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float ny = (f[i] - s.mean[i]) * s.rstd[i];
      const float G = dn[i] * (ny > 0.f ? 1.f : 100.f);          // undo the y-branch LeakyReLU derivative
      const float nx = (fmaf(s.w0[i], x0, s.w1[i] * x1) - s.mx[i]) * s.rx[i];
      const float dx = G * lrelu_grad(nx);
      const float du = s.rx[i] * (dx - s_m1[i] - nx * s_m2[i]);
      red[i] = fmaf(du, x0, red[i]);
      red[8 + i] = fmaf(du, x1, red[8 + i]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float tot = warp_xreduce32(red, lane);
  s_red[warp][lane] = tot;
  __syncthreads();
  if (threadIdx.x < 16) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_red[w][threadIdx.x];
    const int inp = threadIdx.x >> 3, c = k * 8 + (threadIdx.x & 7);
    if (inp < a.in_ch) atomicAdd(a.dwx + c * a.in_ch + inp, sum);
  }
}

int launch_cat_bwd_x(const CatBwdXArgs& a, cudaStream_t st) {
  dim3 grid((unsigned)std::min(a.d.D, 148 * 2), a.C / 8, a.d.N);
  cat_bwd_x_kernel<<<grid, 256, 0, st>>>(a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// adjoint trilinear interpolation (align_corners=True)
// =============================================================================================
struct LerpB { int i0, i1; float l0, l1; };
__device__ __forceinline__ LerpB lerp_ac_b(int dst, int in_size, int out_size) {
  const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float src = scale * (float)dst;
  LerpB r;
  r.i0 = (int)src;
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}
// weight with which output index o reads source index j along one axis
__device__ __forceinline__ float axis_weight(int o, int j, int in_size, int out_size) {
  const LerpB l = lerp_ac_b(o, in_size, out_size);
  return (l.i0 == j ? l.l0 : 0.f) + (l.i1 == j ? l.l1 : 0.f);
}
// candidate output range [lo, hi] whose interpolation may touch source index j
__device__ __forceinline__ void axis_range(int j, int in_size, int out_size, int& lo, int& hi) {
  const float inv = in_size > 1 ? (float)(out_size - 1) / (float)(in_size - 1) : 0.f;
  lo = max(0, (int)floorf((float)(j - 1) * inv) - 1);
  hi = min(out_size - 1, (int)ceilf((float)(j + 1) * inv) + 1);
}

__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const grad_t* __restrict__ gdst, int gdst_chunks, int gdst_off, Dims sd,
                                                            grad_t* __restrict__ gsrc, int C8) {
  const int n = blockIdx.z, k = blockIdx.y;
  const int Do = sd.D * 2, Ho = sd.H * 2, Wo = sd.W * 2;
  const long long Vs = dims_vox(sd), Vo = Vs * 8;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= Vs) return;
  const int jw = (int)(v % sd.W), jh = (int)((v / sd.W) % sd.H), jd = (int)(v / ((long long)sd.W * sd.H));
  int dlo, dhi, hlo, hhi, wlo, whi;
  axis_range(jd, sd.D, Do, dlo, dhi);
  axis_range(jh, sd.H, Ho, hlo, hhi);
  axis_range(jw, sd.W, Wo, wlo, whi);
  const grad_t* gp = gdst + ((size_t)n * gdst_chunks + gdst_off + k) * Vo * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int od = dlo; od <= dhi; ++od) {
    const float wd = axis_weight(od, jd, sd.D, Do);
    if (wd == 0.f) continue;
    for (int oh = hlo; oh <= hhi; ++oh) {
      const float wh = axis_weight(oh, jh, sd.H, Ho) * wd;
      if (wh == 0.f) continue;
      for (int ow = wlo; ow <= whi; ++ow) {
        const float ww = axis_weight(ow, jw, sd.W, Wo) * wh;
        if (ww == 0.f) continue;
        float f[8];
        ld_grad8_cached(gp + (((size_t)od * Ho + oh) * Wo + ow) * 8, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(ww, f[i], acc[i]);
      }
    }
  }
  st_grad8(gsrc + (((size_t)n * C8 + k) * Vs + v) * 8, acc);
}

int launch_upsample2_bwd(const grad_t* gdst, int gdst_chunks, int gdst_off, int C, Dims sd, grad_t* gsrc, cudaStream_t st) {
  const long long Vs = dims_vox(sd);
  dim3 grid((unsigned)((Vs + 255) / 256), C / 8, sd.N);
  upsample2_bwd_kernel<<<grid, 256, 0, st>>>(gdst, gdst_chunks, gdst_off, sd, gsrc, C / 8);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) sum_kernel(const float* __restrict__ src, long long n, float* __restrict__ dst) {
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += src[i];
  __shared__ double sh[8];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(dst, (float)t);
  }
}
int launch_sum(const float* src, long long n, float* dst, cudaStream_t st) {
  SEUNET_CUDA_CHECK(cudaMemsetAsync(dst, 0, sizeof(float), st));
  sum_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 592), 256, 0, st>>>(src, n, dst);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// small parameters: conv2 (weight, bias), conv_se, conv_se2, dc0_0 / dc0_1 weights
//   weff[n][c] = sum_j h_j[n] W2[j][c],  cst[n] = sum_j h_j[n] b2[j],  h_j[n] = hw[2k+j] * drop[n][2k+j]
// =============================================================================================
__global__ void small_grads_kernel(const float* __restrict__ params, const float* __restrict__ drop0,
                                   const float* __restrict__ drop1, const __grid_constant__ SmallGradArgs a,
                                   const float* __restrict__ dweff, const float* __restrict__ dcst,
                                   const float* __restrict__ dwse, const float* __restrict__ dwse2, float* __restrict__ grads) {
  const int b = blockIdx.x, c = threadIdx.x, N = a.N;
  const SmallGradBlock blk = a.blk[b];
  const int hc = blk.head == 0 ? 24 : 12;
  const float* drop = blk.head == 0 ? drop0 : drop1;
  const float* hw = params + a.hw_off[blk.head];
  __shared__ float s_part[2][64];
  float dw2[2] = {0.f, 0.f}, dh[2] = {0.f, 0.f};
  if (c < blk.C) {
    for (int n = 0; n < N; ++n) {
      const float g = dweff[((size_t)b * N + n) * 64 + c];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float dr = drop[n * hc + 2 * blk.k + j];
        dw2[j] = fmaf(hw[2 * blk.k + j] * dr, g, dw2[j]);
        dh[j] = fmaf(dr * params[blk.w2_off + j * blk.C + c], g, dh[j]);
      }
    }
    grads[blk.w2_off + c] = dw2[0];
    grads[blk.w2_off + blk.C + c] = dw2[1];
    grads[blk.wse_off + c] = dwse[b * 64 + c];
    if (blk.wse2_off >= 0) grads[blk.wse2_off + c] = dwse2[b * 64 + c];
  }
  s_part[0][c] = dh[0];
  s_part[1][c] = dh[1];
  __syncthreads();
  if (c < 2) {
    float s = 0.f;
    for (int i = 0; i < 64; ++i) s += s_part[c][i];
    float db2 = 0.f;
    for (int n = 0; n < N; ++n) {
      const float dr = drop[n * hc + 2 * blk.k + c], gc = dcst[(size_t)b * N + n];
      db2 = fmaf(hw[2 * blk.k + c] * dr, gc, db2);
      s = fmaf(dr * params[blk.b2_off + c], gc, s);
    }
    grads[blk.b2_off + c] = db2;
    grads[a.hw_off[blk.head] + 2 * blk.k + c] = s;
  }
}

int launch_small_grads(const float* params, const float* drop0, const float* drop1, const SmallGradArgs& a,
                       const float* dweff, const float* dcst, const float* dwse, const float* dwse2, float* grads,
                       cudaStream_t st) {
  small_grads_kernel<<<18, 64, 0, st>>>(params, drop0, drop1, a, dweff, dcst, dwse, dwse2, grads);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
