// HBM-bound forward kernels: input prep, fused InstanceNorm+LeakyReLU+sSE gate(s)+side-branch fold,
// CAT-block apply (+ detail injection + 2x2x2 max-pool), trilinear x2 upsample, deep-supervision head.
// All activations are chunk planes [n][C/8][D][H][W][8] so every thread moves 16-byte vectors and a
// warp touches 512 contiguous bytes per chunk plane.
#include "pointwise.cuh"
#include <cstring>
#include <cstdlib>
#include <cstdint>

constexpr float kInEps = 1e-5f;   // nn.InstanceNorm3d default eps (SE_UNet.py:17,43,59)

// =============================================================================================
// input prep
// =============================================================================================
// One thread per 4x4x4 block of the full-resolution input: writes the 64 chunk-plane voxels, the
// 8 half-resolution and 1 quarter-resolution max-pooled values (SE_UNet.py:189,198) and contributes
// to the first/second moments of x at each level (used for the analytic InstanceNorm statistics of
// the x33/x63/x93 injection branches).
__global__ void __launch_bounds__(128) input_prep_kernel(const float* __restrict__ x, long long sN, long long sC, long long sD,
                                                         long long sH, long long sW, const __grid_constant__ XOffsets xo,
                                                         int in_ch, Dims d,
                                                         act_t* __restrict__ xb, float* __restrict__ xp1,
                                                         float* __restrict__ xp2, double* __restrict__ mom) {
  const int n = blockIdx.y;
  const long long xbase = xo.use ? xo.off[n] : n * sN;
  const int D4 = d.D >> 2, H4 = d.H >> 2, W4 = d.W >> 2;
  const long long nb = (long long)D4 * H4 * W4;
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // moments: [level][5] = s0, s1, s00, s11, s01
  float m[3][5];
#pragma unroll
  for (int l = 0; l < 3; ++l)
#pragma unroll
    for (int i = 0; i < 5; ++i) m[l][i] = 0.f;
  if (b < nb) {
    const int bw = (int)(b % W4), bh = (int)((b / W4) % H4), bd = (int)(b / ((long long)W4 * H4));
    float p2[kMaxInCh];
#pragma unroll
    for (int c = 0; c < kMaxInCh; ++c) p2[c] = -INFINITY;
    for (int dd2 = 0; dd2 < 2; ++dd2)
      for (int hh2 = 0; hh2 < 2; ++hh2)
        for (int ww2 = 0; ww2 < 2; ++ww2) {
          float p1[kMaxInCh];
#pragma unroll
          for (int c = 0; c < kMaxInCh; ++c) p1[c] = -INFINITY;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int dz = bd * 4 + dd2 * 2 + (k >> 2), hy = bh * 4 + hh2 * 2 + ((k >> 1) & 1), wx = bw * 4 + ww2 * 2 + (k & 1);
            float v[kMaxInCh];
#pragma unroll
            for (int c = 0; c < kMaxInCh; ++c)
              v[c] = c < in_ch ? x[xbase + c * sC + dz * sD + hy * sH + wx * sW] : 0.f;
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < kMaxInCh; ++c) { f[c] = v[c]; p1[c] = fmaxf(p1[c], v[c]); }
            st_chunk(xb + (((size_t)n * d.D + dz) * d.H + hy) * (size_t)d.W * 8 + (size_t)wx * 8, floats_to_chunk(f));
            m[0][0] += v[0]; m[0][1] += v[1]; m[0][2] += v[0] * v[0]; m[0][3] += v[1] * v[1]; m[0][4] += v[0] * v[1];
          }
          const int d1 = bd * 2 + dd2, h1 = bh * 2 + hh2, w1 = bw * 2 + ww2;
#pragma unroll
          for (int c = 0; c < kMaxInCh; ++c) {
            if (c < in_ch) xp1[(((size_t)n * in_ch + c) * (d.D >> 1) + d1) * (size_t)(d.H >> 1) * (d.W >> 1) + (size_t)h1 * (d.W >> 1) + w1] = p1[c];
            else p1[c] = 0.f;
            p2[c] = fmaxf(p2[c], p1[c]);
          }
          m[1][0] += p1[0]; m[1][1] += p1[1]; m[1][2] += p1[0] * p1[0]; m[1][3] += p1[1] * p1[1]; m[1][4] += p1[0] * p1[1];
        }
#pragma unroll
    for (int c = 0; c < kMaxInCh; ++c) {
      if (c < in_ch) xp2[(((size_t)n * in_ch + c) * D4 + bd) * (size_t)H4 * W4 + (size_t)bh * W4 + bw] = p2[c];
      else p2[c] = 0.f;
    }
    m[2][0] = p2[0]; m[2][1] = p2[1]; m[2][2] = p2[0] * p2[0]; m[2][3] = p2[1] * p2[1]; m[2][4] = p2[0] * p2[1];
  }
  // block reduce (fp32 per thread covers <= 64 values; cross-thread accumulation in fp64)
  __shared__ double red[4][15];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int l = 0; l < 3; ++l)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double s = warp_sum_d((double)m[l][i]);
      if (lane == 0) red[warp][l * 5 + i] = s;
    }
  __syncthreads();
  if (threadIdx.x < 15) {
    const double s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    const int l = threadIdx.x / 5, i = threadIdx.x % 5;
    atomicAdd(mom + ((size_t)l * gridDim.y + n) * kMomStride + i, s);
  }
}

// Round 2, contiguous inputs (last-axis stride 1, 16-byte aligned rows - every caller in this repo): a lane owns 4 consecutive
// w voxels of one h row and walks the 4 d-planes of a 4x4x4 block row; a warp is 4 (h) x 8 (w quads), so every load is a
// coalesced 128-bit vector per channel, every lane stores 64 contiguous bytes of chunks, the w pairs of the pooling windows
// stay inside the thread and the h pairs / quads are two shuffles.  (ncu on the one-thread-per-4x4x4-block version above:
// 47 % of the DRAM peak; its scalar loads are 16 B apart between lanes and its chunk stores 64 B apart.)
constexpr int kPrepDG = 4;   // 4x4x4 block rows along d per warp (amortises the moment reduction)
__global__ void __launch_bounds__(128) input_prep_vec_kernel(const float* __restrict__ x, long long sN, long long sC, long long sD,
                                                             long long sH, const __grid_constant__ XOffsets xo, int in_ch, Dims d,
                                                             act_t* __restrict__ xb, float* __restrict__ xp1,
                                                             float* __restrict__ xp2, double* __restrict__ mom) {
  const int n = blockIdx.y;
  const long long xbase = xo.use ? xo.off[n] : n * sN;
  const int D4 = d.D >> 2, H4 = d.H >> 2, W4 = d.W >> 2;
  const int segs = (d.W + 31) >> 5, dgs = (D4 + kPrepDG - 1) / kPrepDG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * 4 + warp;
  float m[3][5];
#pragma unroll
  for (int l = 0; l < 3; ++l)
#pragma unroll
    for (int i = 0; i < 5; ++i) m[l][i] = 0.f;
  if (gw < (long long)segs * H4 * dgs) {
    const int seg = (int)(gw % segs), bh = (int)((gw / segs) % H4), bdg = (int)(gw / ((long long)segs * H4));
    const int hq = lane >> 3, wq = lane & 7;
    const int w0 = seg * 32 + wq * 4, hy = bh * 4 + hq;
    const bool act = w0 < d.W;
    const int H2 = d.H >> 1, W2 = d.W >> 1;
    for (int bd = bdg * kPrepDG; bd < min(D4, bdg * kPrepDG + kPrepDG); ++bd) {
      float p2[kMaxInCh], q01[kMaxInCh], q23[kMaxInCh];
#pragma unroll
      for (int c = 0; c < kMaxInCh; ++c) p2[c] = q01[c] = q23[c] = -INFINITY;
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) {
        const int dz = bd * 4 + dd;
        float4 v[kMaxInCh];
#pragma unroll
        for (int c = 0; c < kMaxInCh; ++c) {
          v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (act && c < in_ch) v[c] = __ldg(reinterpret_cast<const float4*>(x + xbase + c * sC + dz * sD + hy * sH + w0));
        }
        if (act) {
          act_t* op = xb + ((((size_t)n * d.D + dz) * d.H + hy) * (size_t)d.W + w0) * 8;
          const float* v0 = reinterpret_cast<const float*>(&v[0]);
          const float* v1 = reinterpret_cast<const float*>(&v[kMaxInCh - 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            f[0] = v0[j];
            if (kMaxInCh > 1) f[1] = v1[j];
            st_chunk(op + j * 8, floats_to_chunk(f));
            const float a0 = v0[j], a1 = kMaxInCh > 1 ? v1[j] : 0.f;
            m[0][0] += a0; m[0][1] += a1; m[0][2] += a0 * a0; m[0][3] += a1 * a1; m[0][4] += a0 * a1;
          }
        }
#pragma unroll
        for (int c = 0; c < kMaxInCh; ++c) {
          const float a = act ? fmaxf(v[c].x, v[c].y) : -INFINITY, b = act ? fmaxf(v[c].z, v[c].w) : -INFINITY;
          q01[c] = (dd & 1) ? fmaxf(q01[c], a) : a;
          q23[c] = (dd & 1) ? fmaxf(q23[c], b) : b;
        }
        if (dd & 1) {   // a 2x2x2 window is complete along d: finish it along h
          float r01[kMaxInCh], r23[kMaxInCh];
#pragma unroll
          for (int c = 0; c < kMaxInCh; ++c) {
            r01[c] = fmaxf(q01[c], __shfl_xor_sync(0xffffffffu, q01[c], 8));
            r23[c] = fmaxf(q23[c], __shfl_xor_sync(0xffffffffu, q23[c], 8));
            p2[c] = fmaxf(p2[c], fmaxf(r01[c], r23[c]));
          }
          if (act && !(hq & 1)) {
            const int d1 = dz >> 1, h1 = hy >> 1, w1 = w0 >> 1;
            float pa[2] = {0.f, 0.f}, pb[2] = {0.f, 0.f};
#pragma unroll
            for (int c = 0; c < kMaxInCh; ++c)
              if (c < in_ch) {
                pa[c] = r01[c]; pb[c] = r23[c];
                *reinterpret_cast<float2*>(xp1 + (((size_t)n * in_ch + c) * (d.D >> 1) + d1) * (size_t)H2 * W2 + (size_t)h1 * W2 + w1) =
                    make_float2(r01[c], r23[c]);
              }
            m[1][0] += pa[0] + pb[0]; m[1][1] += pa[1] + pb[1]; m[1][2] += pa[0] * pa[0] + pb[0] * pb[0];
            m[1][3] += pa[1] * pa[1] + pb[1] * pb[1]; m[1][4] += pa[0] * pa[1] + pb[0] * pb[1];
          }
        }
      }
      float t2[2] = {0.f, 0.f};
#pragma unroll
      for (int c = 0; c < kMaxInCh; ++c) {
        p2[c] = fmaxf(p2[c], __shfl_xor_sync(0xffffffffu, p2[c], 16));
        if (act && hq == 0 && c < in_ch) {
          t2[c] = p2[c];
          xp2[(((size_t)n * in_ch + c) * D4 + bd) * (size_t)H4 * W4 + (size_t)bh * W4 + (w0 >> 2)] = p2[c];
        }
      }
      if (act && hq == 0) { m[2][0] += t2[0]; m[2][1] += t2[1]; m[2][2] += t2[0] * t2[0]; m[2][3] += t2[1] * t2[1]; m[2][4] += t2[0] * t2[1]; }
    }
  }
  __shared__ double red[4][15];
#pragma unroll
  for (int l = 0; l < 3; ++l)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double s = warp_sum_d((double)m[l][i]);
      if (lane == 0) red[warp][l * 5 + i] = s;
    }
  __syncthreads();
  if (threadIdx.x < 15) {
    const double s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    const int l = threadIdx.x / 5, i = threadIdx.x % 5;
    atomicAdd(mom + ((size_t)l * gridDim.y + n) * kMomStride + i, s);
  }
}

int launch_input_prep(const float* x, const long long* xs, const XOffsets& xo, int in_ch, Dims d, act_t* xb, float* xp1, float* xp2,
                      double* mom, cudaStream_t st, bool inference) {
  if (in_ch < 1 || in_ch > kMaxInCh) { seunet_set_error("in_channel %d unsupported (1..%d)", in_ch, kMaxInCh); return 1; }
  if ((d.D | d.H | d.W) & 7) { seunet_set_error("spatial dims must be multiples of 8"); return 1; }
  SEUNET_CUDA_CHECK(cudaMemsetAsync(mom, 0, sizeof(double) * 3 * d.N * kMomStride, st));
  // vector path: contiguous last axis and 16-byte aligned rows
  bool vec = xs[4] == 1 && ((uintptr_t)x & 15) == 0 && ((xs[0] | xs[1] | xs[2] | xs[3]) & 3) == 0;
  if (xo.use)
    for (int n = 0; n < d.N && n < kMaxWindowBatch; ++n) vec = vec && (xo.off[n] & 3) == 0;
  static const bool allow_vec = !(getenv("SEUNET_PREP_VEC") && atoi(getenv("SEUNET_PREP_VEC")) == 0);
  if (vec && allow_vec && inference) {
    const long long warps = (long long)((d.W + 31) / 32) * (d.H / 4) * ((d.D / 4 + kPrepDG - 1) / kPrepDG);
    dim3 gridv((unsigned)((warps + 3) / 4), d.N);
    input_prep_vec_kernel<<<gridv, 128, 0, st>>>(x, xs[0], xs[1], xs[2], xs[3], xo, in_ch, d, xb, xp1, xp2, mom);
    SEUNET_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  const long long nb = (long long)(d.D / 4) * (d.H / 4) * (d.W / 4);
  dim3 grid((unsigned)((nb + 127) / 128), d.N);
  input_prep_kernel<<<grid, 128, 0, st>>>(x, xs[0], xs[1], xs[2], xs[3], xs[4], xo, in_ch, d, xb, xp1, xp2, mom);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// SSE block apply:  y -> IN -> LeakyReLU -> sSE gate(s) -> e0 (+ folded conv2/head contribution)
// (SE_UNet.py:26-33 / 70-80; fold of conv2 + up_sample + dc0_x per SURVEY App. C)
// =============================================================================================
// (register caps chosen by A/B: C = 64 gains 17 % from three blocks per SM instead of two; the narrower variants need <= 64
// registers for four blocks (C = 32) / 42 for six (C <= 16) - with fewer resident blocks they are 7-10 % slower)
// VPT voxels per thread (round 2): ncu on the 8- and 16-channel instances showed 60-70 % issue-slot utilisation at 46 % of the DRAM
// peak - with one 16/32-byte voxel per thread the per-block prologue (fp64 statistics, barrier) and the address arithmetic
// outweigh the payload; those instances now loop over 4 voxels per thread (rolled: hoisting the loads of several voxels spilled).
constexpr int kSseSplitDefault = 1;
template <int C> struct SseVpt { static constexpr int value = C <= 16 ? 4 : 1; };

// SPLIT = 2 (round 2, the 64-channel instances): the two lanes l and l^16 of a warp share a voxel, 32 channels each - half the
// registers (four resident blocks instead of three, no spill), half the serial arithmetic per thread and twice the threads
// on the 32^3 / 16^3 levels, where the one-thread-per-voxel version was latency-bound (ncu: 19 us for 59 MB, 34 % of the
// warp slots).  The gate sums and the folded side-branch sum are completed with one shuffle each.
template <int C, int GATES, int SPLIT, int VPT>
__global__ void __launch_bounds__(256, SPLIT == 2 ? 4 : (C == 64 ? 3 : (C == 32 ? 4 : 6))) apply_sse_kernel(const __grid_constant__ SseArgs a) {
  constexpr int KPT = C / 8 / SPLIT;       // channel chunks per thread
  constexpr int CT = C / SPLIT;            // channels per thread
  __shared__ __align__(16) float s_mean[C], s_rstd[C], s_wse[C], s_wse2[C], s_weff[C];
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
  // per-channel constants come from shared memory as broadcast 128-bit loads (the kernel is issue-bound otherwise)
  auto ld8 = [](const float* sm, int k, float* r) {
    *reinterpret_cast<float4*>(r) = *reinterpret_cast<const float4*>(sm + k * 8);
    *reinterpret_cast<float4*>(r + 4) = *reinterpret_cast<const float4*>(sm + k * 8 + 4);
  };
  const bool has_t = a.T != nullptr;     // window plans drop head 0: its blocks neither fold nor touch the accumulator
  float* tp0 = a.T + (size_t)n * a.V;
  const float wcst = a.wcst[n];
  // SPLIT == 2: lanes 0-15 hold chunks 0..KPT-1 of 16 consecutive voxels, lanes 16-31 chunks KPT..2*KPT-1 of the same voxels
  const int sub = SPLIT == 2 ? (threadIdx.x >> 4) & 1 : 0;
  const int tvox = SPLIT == 2 ? (threadIdx.x & 15) | ((threadIdx.x >> 5) << 4) : threadIdx.x;
  constexpr int VPB = 256 / SPLIT;
  const int k0 = sub * KPT;
#pragma unroll 1     // (unrolled x2 / x4 for the 16- / 8-channel instances, round 2 late: ec2 24.9 -> 26.1 us, dc6 17.6 -> 18.4, ec1 unchanged)
  for (int u = 0; u < VPT; ++u) {
    const long long v = (blockIdx.x * (long long)VPT + u) * VPB + tvox;
    if (v >= a.V) break;                    // (V is a multiple of 32: both lanes of a pair leave together)
    float* tp = tp0 + v;
    const float t_old = (a.t_init || !has_t || sub) ? 0.f : *tp;   // issued with the raw loads, not after the gate arithmetic
    float e[CT];
    float g1 = 0.f;
    Chunk8 in[KPT];
#pragma unroll
    for (int k = 0; k < KPT; ++k) in[k] = ld_chunk_stream(a.raw + (((size_t)n * a.raw_chunks + k0 + k) * a.V + v) * 8);
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      float f[8], mean[8], rstd[8], wse[8];
      chunk_to_floats(in[k], f);
      ld8(s_mean, k0 + k, mean); ld8(s_rstd, k0 + k, rstd); ld8(s_wse, k0 + k, wse);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = lrelu_((f[i] - mean[i]) * rstd[i]);
        e[k * 8 + i] = t;
        g1 = fmaf(wse[i], t, g1);
      }
    }
    if (SPLIT == 2) g1 += __shfl_xor_sync(0xffffffffu, g1, 16);
    g1 = sigmoidf_(g1);
    if (GATES == 2) {
      float g2 = 0.f;
#pragma unroll
      for (int k = 0; k < KPT; ++k) {
        float wse2[8];
        ld8(s_wse2, k0 + k, wse2);
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[k * 8 + i] *= g1; g2 = fmaf(wse2[i], e[k * 8 + i], g2); }
      }
      if (SPLIT == 2) g2 += __shfl_xor_sync(0xffffffffu, g2, 16);
      g1 = sigmoidf_(g2);   // the second gate multiplies below
    }
    float t = sub ? 0.f : wcst;
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      float weff[8];
      ld8(s_weff, k0 + k, weff);
#pragma unroll
      for (int i = 0; i < 8; ++i) { e[k * 8 + i] *= g1; t = fmaf(weff[i], e[k * 8 + i], t); }
    }
    if (SPLIT == 2) t += __shfl_xor_sync(0xffffffffu, t, 16);
    if (has_t && !sub) *tp = t_old + t;
    if (a.dest) {
#pragma unroll
      for (int k = 0; k < KPT; ++k)
        st_chunk(a.dest + (((size_t)n * a.dest_chunks + a.dest_off + k0 + k) * a.V + v) * 8, floats_to_chunk(e + k * 8));
    }
  }
}

template <int C, int SPLIT, int VPT>
static int launch_apply_sse_v(int N, const SseArgs& a, cudaStream_t st) {
  constexpr int VPB = 256 / SPLIT * VPT;
  dim3 grid((unsigned)((a.V + VPB - 1) / VPB), N);
  if (a.wse2) apply_sse_kernel<C, 2, SPLIT, VPT><<<grid, 256, 0, st>>>(a);
  else apply_sse_kernel<C, 1, SPLIT, VPT><<<grid, 256, 0, st>>>(a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int C>
static int launch_apply_sse_c(int N, const SseArgs& a, cudaStream_t st) {
  static const int split_mask = getenv("SEUNET_SSE_SPLIT") ? atoi(getenv("SEUNET_SSE_SPLIT")) : kSseSplitDefault;   // bit 0: C = 64, bit 1: C = 32
  if constexpr (C >= 32) {
    // Voxels per thread of the wide instances: with one voxel per thread every 128/256 voxels pay the block's fp64 statistics
    // prologue and barrier, which is as long as the payload; looping over 4 voxel groups amortises it (dc5 48.9 -> 44.0 us per
    // window, ec6 16.8 -> 14.6) - but only where enough blocks remain to fill the SMs (the 32^3 / 16^3 levels got slower).
    // Per voxel the arithmetic is unchanged (8 groups per thread: no further gain).  SEUNET_SSE_VPT = 1 | 4: A/B knob.
    static const int vpt_env = getenv("SEUNET_SSE_VPT") ? atoi(getenv("SEUNET_SSE_VPT")) : 4;
    const bool split = a.inference && (split_mask & (C == 64 ? 1 : 2));
    const long long groups = (long long)N * ((a.V + (split ? 127 : 255)) / (split ? 128 : 256));
    const bool many = vpt_env >= 4 && groups / 4 >= 148 * 8;
    if (split) return many ? launch_apply_sse_v<C, 2, 4>(N, a, st) : launch_apply_sse_v<C, 2, 1>(N, a, st);
    return many ? launch_apply_sse_v<C, 1, 4>(N, a, st) : launch_apply_sse_v<C, 1, 1>(N, a, st);
  } else {
    return launch_apply_sse_v<C, 1, SseVpt<C>::value>(N, a, st);   // (8 voxels per thread measured: no further gain)
  }
}
int launch_apply_sse(int C, int N, const SseArgs& a, cudaStream_t st) {
  switch (C) {
    case 8: return launch_apply_sse_c<8>(N, a, st);
    case 16: return launch_apply_sse_c<16>(N, a, st);
    case 32: return launch_apply_sse_c<32>(N, a, st);
    case 64: return launch_apply_sse_c<64>(N, a, st);
  }
  seunet_set_error("apply_sse: C=%d unsupported", C);
  return 1;
}

// =============================================================================================
// CAT block apply:  out = lrelu(IN(y)) [+ lrelu(IN(Wx x))]  -> full-res slot and/or 2x2x2 max-pool
// (SE_UNet.py:45-49, 186-189, 195-198, 204-206, 212, 218, 224)
// =============================================================================================
// One thread owns one (h, w) position of a d-plane (POOL: the 2 x 2 (d, h) voxels of a pooling window at one w) and loops
// over all channel chunks: index arithmetic, the x-branch loads and the per-channel constants are amortised over C
// channels, every load/store is a coalesced 16 B per lane, and the w pair of the pooling window is reduced with a shuffle.
template <int C, bool HASX, bool POOL>
__global__ void __launch_bounds__(256, 3) apply_cat_kernel(const __grid_constant__ CatArgs a) {
  __shared__ __align__(16) float s_mean[C], s_rstd[C], s_ax[C], s_bx[C], s_cx[C];
  const int n = blockIdx.z;
  const int W = a.d.W, H = a.d.H;
  const long long V = dims_vox(a.d);
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const double s = a.stats[((size_t)n * a.stats_c + c) * 2], q = a.stats[((size_t)n * a.stats_c + c) * 2 + 1];
    const double mean = s / (double)V;
    double var = q / (double)V - mean * mean;
    if (var < 0) var = 0;
    s_mean[c] = (float)mean;
    s_rstd[c] = (float)(1.0 / sqrt(var + (double)kInEps));
    if (HASX) {
      // analytic InstanceNorm statistics of the 1x1x1 conv of x: mean' = w.mu, var' = w^T Cov w;
      // (w.x - mean') * rstd' is folded into one affine form ax*x0 + bx*x1 + cx
      const double* m = a.mom + (size_t)n * kMomStride;
      const double mu0 = m[0] / V, mu1 = m[1] / V;
      const double c00 = m[2] / V - mu0 * mu0, c11 = m[3] / V - mu1 * mu1, c01 = m[4] / V - mu0 * mu1;
      const double w0 = a.wx[c * a.in_ch], w1 = a.in_ch > 1 ? a.wx[c * a.in_ch + 1] : 0.0;
      double vx = w0 * w0 * c00 + w1 * w1 * c11 + 2.0 * w0 * w1 * c01;
      if (vx < 0) vx = 0;
      const double rx = 1.0 / sqrt(vx + (double)kInEps);
      s_ax[c] = (float)(w0 * rx); s_bx[c] = (float)(w1 * rx); s_cx[c] = (float)(-(w0 * mu0 + w1 * mu1) * rx);
    }
  }
  __syncthreads();
  constexpr int NV = POOL ? 4 : 1;
  const int rows = POOL ? (H >> 1) : H;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < rows * W;
  const int tt = live ? t : 0;
  const int wx = tt % W, r = tt / W;
  int vox[NV];
  float x0[NV], x1[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int dz = POOL ? (int)blockIdx.y * 2 + (j >> 1) : (int)blockIdx.y;
    const int hy = POOL ? r * 2 + (j & 1) : r;
    vox[j] = (dz * H + hy) * W + wx;
    x0[j] = 0.f; x1[j] = 0.f;
    if (HASX) {
      const float* xp = a.x + (a.xo.use ? a.xo.off[n] : n * a.xs[0]) + dz * a.xs[2] + hy * a.xs[3] + wx * a.xs[4];
      x0[j] = __ldg(xp);
      if (a.in_ch > 1) x1[j] = __ldg(xp + a.xs[1]);
    }
  }
  const int Hp = H >> 1, Wp = W >> 1;
  const long long Vp = V >> 3;
  const int pvox = POOL ? ((int)blockIdx.y * Hp + r) * Wp + (wx >> 1) : 0;
#pragma unroll 1
  for (int k = 0; k < C / 8; ++k) {
    const act_t* rawp = a.raw + ((size_t)n * a.raw_chunks + k) * V * 8;
    float mean[8], rstd[8], ax[8], bx[8], cx[8];
#pragma unroll
    for (int i = 0; i < 8; i += 4) {
      *reinterpret_cast<float4*>(mean + i) = *reinterpret_cast<const float4*>(s_mean + k * 8 + i);
      *reinterpret_cast<float4*>(rstd + i) = *reinterpret_cast<const float4*>(s_rstd + k * 8 + i);
      if (HASX) {
        *reinterpret_cast<float4*>(ax + i) = *reinterpret_cast<const float4*>(s_ax + k * 8 + i);
        *reinterpret_cast<float4*>(bx + i) = *reinterpret_cast<const float4*>(s_bx + k * 8 + i);
        *reinterpret_cast<float4*>(cx + i) = *reinterpret_cast<const float4*>(s_cx + k * 8 + i);
      }
    }
    Chunk8 in[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) in[j] = ld_chunk_stream(rawp + (size_t)vox[j] * 8);
    float mx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float f[8], o[8];
      chunk_to_floats(in[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = lrelu_((f[i] - mean[i]) * rstd[i]);
        if (HASX) v += lrelu_(fmaf(ax[i], x0[j], fmaf(bx[i], x1[j], cx[i])));
        o[i] = v;
        mx[i] = fmaxf(mx[i], v);
      }
      if (a.dest && live) st_chunk(a.dest + (((size_t)n * a.dest_chunks + a.dest_off + k) * V + vox[j]) * 8, floats_to_chunk(o));
    }
    if (POOL) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      if (live && !(wx & 1))
        st_chunk(a.pdest + (((size_t)n * a.pdest_chunks + a.pdest_off + k) * Vp + pvox) * 8, floats_to_chunk(mx));
    }
  }
}

template <int C>
static int launch_apply_cat_c(const CatArgs& a, cudaStream_t st) {
  const bool pool = a.pdest != nullptr, hasx = a.x != nullptr;
  const int rows = pool ? a.d.H / 2 : a.d.H, planes = pool ? a.d.D / 2 : a.d.D;
  dim3 grid((unsigned)((rows * a.d.W + 255) / 256), planes, a.d.N);
  if (hasx && pool) apply_cat_kernel<C, true, true><<<grid, 256, 0, st>>>(a);
  else if (hasx) apply_cat_kernel<C, true, false><<<grid, 256, 0, st>>>(a);
  else if (pool) apply_cat_kernel<C, false, true><<<grid, 256, 0, st>>>(a);
  else apply_cat_kernel<C, false, false><<<grid, 256, 0, st>>>(a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int launch_apply_cat(int C, const CatArgs& a, cudaStream_t st) {
  if (a.x && (a.in_ch < 1 || a.in_ch > kMaxInCh)) { seunet_set_error("apply_cat: in_ch unsupported"); return 1; }
  switch (C) {
    case 16: return launch_apply_cat_c<16>(a, st);
    case 32: return launch_apply_cat_c<32>(a, st);
    case 64: return launch_apply_cat_c<64>(a, st);
  }
  seunet_set_error("apply_cat: C=%d unsupported", C);
  return 1;
}

// =============================================================================================
// folded head weights:  weff[blk][n][c] = sum_j hw[2k+j] * drop[n][2k+j] * W2[j][c]
// (conv2 SE_UNet.py:33/80, DropLayer 89-97, dc0_0/dc0_1 232-233; all linear, so they commute
//  with the trilinear up-sampling of the side branch)
// =============================================================================================
__global__ void headw_kernel(const float* __restrict__ params, const float* __restrict__ drop0,
                             const float* __restrict__ drop1, const __grid_constant__ HeadwArgs a,
                             float* __restrict__ weff, float* __restrict__ wcst) {
  const int b = blockIdx.x, n = blockIdx.y, N = gridDim.y;
  const HeadwBlock blk = a.blk[b];
  const int hc = blk.head == 0 ? 24 : 12;
  const float* drop = blk.head == 0 ? drop0 : drop1;
  const float* hw = params + a.hw_off[blk.head];
  const float h0 = hw[2 * blk.k] * drop[n * hc + 2 * blk.k];
  const float h1 = hw[2 * blk.k + 1] * drop[n * hc + 2 * blk.k + 1];
  const int c = threadIdx.x;
  if (c < 64) {
    float v = 0.f;
    if (c < blk.C) v = h0 * params[blk.w2_off + c] + h1 * params[blk.w2_off + blk.C + c];
    weff[((size_t)b * N + n) * 64 + c] = v;
  }
  if (c == 0) wcst[(size_t)b * N + n] = h0 * params[blk.b2_off] + h1 * params[blk.b2_off + 1];
}

int launch_headw(const float* params, const float* drop0, const float* drop1, int N, const HeadwArgs& a,
                 float* weff, float* wcst, cudaStream_t st) {
  dim3 grid(a.nblk, N);
  headw_kernel<<<grid, 64, 0, st>>>(params, drop0, drop1, a, weff, wcst);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
