// Separable adjoint of the trilinear (align_corners=True) up-sampling folded into the heads: Up = U_d (x) U_h (x) U_w, so
// Up^T is three 1-D adjoint passes (w, then h, then d) over shrinking tensors instead of one (2*2^l+1)^3 gather per voxel.
#include "backward.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ float axis_w1(int o, int j, int in_size, int out_size) {   // weight of source j in output o
  const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float src = scale * (float)o;
  const int i0 = (int)src;
  const int i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  const float l1 = src - (float)i0;
  return (i0 == j ? 1.f - l1 : 0.f) + (i1 == j ? l1 : 0.f);
}

// tensor viewed as [outer][axis][inner]; out[outer][j][inner] = sum_o w(o -> j) * in[outer][o][inner]
// VEC = 4: four consecutive inner elements per thread (the h and d passes), so the candidate scan and its weights are paid once
// per float4; all index arithmetic is 32-bit (ncu on the first version: 80 % issue-slot utilisation at 5 % of the DRAM
// bandwidth - three 64-bit divisions per element cost more than the taps).
template <int VEC>
__global__ void __launch_bounds__(256) adjoint_axis_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned total /*threads' worth*/,
                                                           int out_len /*fine*/, int in_len /*coarse*/, int innerV /*inner / VEC*/) {
  const float inv = in_len > 1 ? (float)(out_len - 1) / (float)(in_len - 1) : 0.f;
  const float scale = out_len > 1 ? (float)(in_len - 1) / (float)(out_len - 1) : 0.f;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const unsigned q = e % (unsigned)innerV, r = e / (unsigned)innerV;
    const int j = (int)(r % (unsigned)in_len);
    const unsigned o_ = r / (unsigned)in_len;
    const int lo = max(0, (int)floorf((float)(j - 1) * inv) - 1);
    const int hi = min(out_len - 1, (int)ceilf((float)(j + 1) * inv) + 1);
    const float* p = in + ((size_t)o_ * out_len * innerV + q) * VEC;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    for (int o = lo; o <= hi; ++o) {
      const float src = scale * (float)o;            // axis_w1(o, j, in_len, out_len) with the loop invariants hoisted
      const int i0 = (int)src;
      const int i1 = i0 + (i0 < in_len - 1 ? 1 : 0);
      const float l1 = src - (float)i0;
      const float w = (i0 == j ? 1.f - l1 : 0.f) + (i1 == j ? l1 : 0.f);
      if (w != 0.f) {
        if (VEC == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)o * innerV * 4));
          acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]); acc[2] = fmaf(w, v.z, acc[2]); acc[VEC - 1] = fmaf(w, v.w, acc[VEC - 1]);
        } else {
          acc[0] = fmaf(w, __ldg(p + (size_t)o * innerV), acc[0]);
        }
      }
    }
    if (VEC == 4) reinterpret_cast<float4*>(out)[e] = make_float4(acc[0], acc[1], acc[2], acc[VEC - 1]);
    else out[e] = acc[0];
  }
}

static int adjoint_axis(const float* in, float* out, long long outer, int fine, int coarse, long long inner, cudaStream_t st) {
  const bool vec = inner % 4 == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
  const long long total = outer * coarse * (vec ? inner / 4 : inner);
  if (total >= (1LL << 31) || inner >= (1LL << 31)) { seunet_set_error("head adjoint: tensor too large for 32-bit indexing"); return 1; }
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  if (vec) adjoint_axis_kernel<4><<<blocks, 256, 0, st>>>(in, out, (unsigned)total, fine, coarse, (int)(inner / 4));
  else adjoint_axis_kernel<1><<<blocks, 256, 0, st>>>(in, out, (unsigned)total, fine, coarse, (int)inner);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// dT[n][Ds][Hs][Ws] = Up_{2^level}^T(dpred[n][D][H][W]); tmp1 >= N*D*H*Ws floats, tmp2 >= N*D*Hs*Ws floats
int launch_head_bwd_level_sep(const float* dpred, Dims full, int level, float* dT, float* tmp1, float* tmp2, cudaStream_t st) {
  const int Ds = full.D >> level, Hs = full.H >> level, Ws = full.W >> level;
  if (adjoint_axis(dpred, tmp1, (long long)full.N * full.D * full.H, full.W, Ws, 1, st)) return 1;          // along w
  if (adjoint_axis(tmp1, tmp2, (long long)full.N * full.D, full.H, Hs, Ws, st)) return 1;                   // along h
  return adjoint_axis(tmp2, dT, (long long)full.N, full.D, Ds, (long long)Hs * Ws, st);                     // along d
}


#ifndef SEUNET_GRAD_BF16
// ---------------------------------------------------------------------------------------------------------------------
// Separable adjoint of the trunk x2 up-sampling (SE_UNet.py:136-138) on fp32 gradient chunk planes [plane][D][H][W][8]:
// three 1-D passes (w, h, d) over shrinking tensors, 16 bytes per thread, instead of a ~4x4x4-tap gather of 32-byte
// chunks per source voxel (277 us per patch for the 32-channel 64^3 -> 128^3 layer).
// ---------------------------------------------------------------------------------------------------------------------
// tensor plane viewed as [A][axis][inner4 float4]; out[plane][a][j][q] = sum_o w(o -> j) * in[plane][a][o][q]
__global__ void __launch_bounds__(256) adjoint_axis_vec_kernel(const float4* __restrict__ in, long long in_sample_stride4, long long in_plane_stride4,
                                                               float4* __restrict__ out, int K, int A, int fine, int coarse, int inner4) {
  const int plane = blockIdx.z, n = plane / K, k = plane - n * K, a = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= coarse * inner4) return;
  const int j = t / inner4, q = t - j * inner4;
  const float inv = coarse > 1 ? (float)(fine - 1) / (float)(coarse - 1) : 0.f;
  // outputs that read source j have their source coordinate in (j-1, j+1): at most first+1 .. first+5
  const int first = max(0, (int)floorf((float)(j - 1) * inv));
  const float4* p = in + n * in_sample_stride4 + k * in_plane_stride4 + ((long long)a * fine + first) * inner4 + q;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const int o = first + c;
    if (o >= fine) break;
    const float w = axis_w1(o, j, coarse, fine);
    if (w != 0.f) {
      const float4 v = __ldg(p + (long long)c * inner4);
      acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
    }
  }
  out[(((long long)plane * A + a) * coarse + j) * inner4 + q] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same adjoint in ONE pass (round 2): the three 1-D passes above move 268 + 134 + 134 + 67 + 67 + 33 MB per 128^3 patch
// for the 32-channel layer; only the first read and the last write are algorithmic.  A block owns TH source rows (all of W)
// of one chunk plane and marches through the fine d-planes: the w pass reads the fine rows straight from global memory
// (neighbouring threads share sectors through L1), the h pass runs out of shared memory, and the d pass is a ROLLING pair
// of register accumulators (fine plane o feeds source planes i0(o) and i0(o)+1; a source plane is written when i0 moves
// past it).  Same taps, weights and summation order as the three passes, so the result is bit-identical.
// ---------------------------------------------------------------------------------------------------------------------
struct AxisTaps { int first; float w[6]; int pad; };   // fine indices first .. first+5 feed source index j with weights w (zeros allowed)
__device__ __forceinline__ AxisTaps axis_taps(int j, int coarse, int fine) {
  const float inv = coarse > 1 ? (float)(fine - 1) / (float)(coarse - 1) : 0.f;
  AxisTaps t;
  t.first = max(0, (int)floorf((float)(j - 1) * inv));
  t.pad = 0;
#pragma unroll
  for (int c = 0; c < 6; ++c) t.w[c] = t.first + c < fine ? axis_w1(t.first + c, j, coarse, fine) : 0.f;
  return t;
}

constexpr int kUbThreads = 512;   // one (source row, w) position per thread: TH * Ws <= 512
constexpr int kUbMaxItems = 1;

__global__ void __launch_bounds__(kUbThreads, 2) upsample2_bwd_fused_kernel(const float* __restrict__ gdst, long long sample_stride, long long plane_stride,
                                                                     Dims sd, float* __restrict__ gsrc, int K, int TH, int dper) {
  extern __shared__ __align__(32) uint8_t s_ub[];
  const int Ws = sd.W, Hs = sd.H, Ds = sd.D, Wo = Ws * 2, Ho = Hs * 2, Do = Ds * 2;
  AxisTaps* tabW = reinterpret_cast<AxisTaps*>(s_ub);           // [Ws]
  AxisTaps* tabH = tabW + Ws;                                   // [TH]
  float* R = reinterpret_cast<float*>(tabH + TH);               // fine rows of the current planes after the w pass (layout below)
  const int Hmax = 2 * TH + 6;
  const int plane = blockIdx.z, n = plane / K, k = plane - n * K;
  const int jh0 = blockIdx.y * TH, the = min(TH, Hs - jh0);
  const int jd0 = blockIdx.x * dper, jd1 = min(Ds, jd0 + dper);
  for (int t = threadIdx.x; t < Ws; t += blockDim.x) tabW[t] = axis_taps(t, Ws, Wo);
  if (threadIdx.x < the) tabH[threadIdx.x] = axis_taps(jh0 + threadIdx.x, Hs, Ho);
  __syncthreads();
  const int ohlo = tabH[0].first, ohhi = min(Ho - 1, tabH[the - 1].first + 5), Hn = ohhi - ohlo + 1;
  const int odlo = axis_taps(jd0, Ds, Do).first, odhi = min(Do - 1, axis_taps(jd1 - 1, Ds, Do).first + 5);
  const float* gp = gdst + n * sample_stride + k * plane_stride;
  float* op = gsrc + (size_t)plane * Ds * Hs * Ws * 8;
  const float dscale = Do > 1 ? (float)(Ds - 1) / (float)(Do - 1) : 0.f;
  // this thread's (source row, w) positions of the h / d passes
  int jh_[kUbMaxItems], jw_[kUbMaxItems];
  bool live[kUbMaxItems];
#pragma unroll
  for (int it = 0; it < kUbMaxItems; ++it) {
    const int item = threadIdx.x + it * kUbThreads;
    live[it] = item < the * Ws;
    jh_[it] = live[it] ? item / Ws : 0;
    jw_[it] = live[it] ? item - jh_[it] * Ws : 0;
  }
  float acc[2][kUbMaxItems][8];
#pragma unroll
  for (int sl = 0; sl < 2; ++sl)
#pragma unroll
    for (int it = 0; it < kUbMaxItems; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[sl][it][i] = 0.f;
  int cur = (int)(dscale * (float)odlo);   // source plane of slot 0 (= i0 of the current fine plane)
  auto flush0 = [&]() {                    // slot 0 is complete: write it (if it is one of this block's planes), shift the pair
    if (cur >= jd0 && cur < jd1) {
#pragma unroll
      for (int it = 0; it < kUbMaxItems; ++it)
        if (live[it]) st_grad8(op + (((size_t)cur * Hs + jh0 + jh_[it]) * Ws + jw_[it]) * 8, acc[0][it]);
    }
#pragma unroll
    for (int it = 0; it < kUbMaxItems; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[0][it][i] = acc[1][it][i]; acc[1][it][i] = 0.f; }
    ++cur;
  };
  // PL fine planes per step: their w passes are independent, so a thread has PL x as many global loads between two barriers
  // (ncu on the one-plane version: 59 % of the stall samples on the load round trip, 6.9 us per plane and block whatever the
  // block size).  R holds the two 16-byte halves of a chunk in separate arrays: 32-byte strides were 2-way bank conflicts.
  constexpr int PL = 2;
  const int rows = Hn * Ws;                      // (row, w) positions of one fine plane after the w pass
  float4* Ra = reinterpret_cast<float4*>(R);     // [PL][rows] channels 0..3
  float4* Rb = Ra + PL * Hmax * Ws;              // [PL][rows] channels 4..7
  for (int od0 = odlo; od0 <= odhi; od0 += PL) {
    const int npl = min(PL, odhi - od0 + 1);
    // ---- w pass: fine rows ohlo..ohhi of planes od0 .. od0+npl-1 -> R
    for (int item = threadIdx.x; item < npl * rows; item += blockDim.x) {
      const int pz = item >= rows ? 1 : 0, ir = item - pz * rows;
      const int r = ir / Ws, jw = ir - r * Ws;
      const AxisTaps t = tabW[jw];
      const float* row = gp + (((size_t)(od0 + pz) * Ho + ohlo + r) * Wo + t.first) * 8;
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 6; ++c)
        if (t.w[c] != 0.f) {
          float v[8];
          ld_grad8_cached(row + c * 8, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = fmaf(t.w[c], v[i], a[i]);
        }
      Ra[pz * Hmax * Ws + ir] = make_float4(a[0], a[1], a[2], a[3]);
      Rb[pz * Hmax * Ws + ir] = make_float4(a[4], a[5], a[6], a[7]);
    }
    __syncthreads();
    // ---- h pass from shared memory, then the rolling d pass in registers, plane by plane
    for (int pz = 0; pz < npl; ++pz) {
      const int od = od0 + pz;
      const int i0 = (int)(dscale * (float)od);
      while (cur < i0) flush0();             // (block-uniform)
      const float wa = axis_w1(od, cur, Ds, Do), wb = cur + 1 < Ds ? axis_w1(od, cur + 1, Ds, Do) : 0.f;
#pragma unroll
      for (int it = 0; it < kUbMaxItems; ++it) {
        if (!live[it]) continue;
        const AxisTaps t = tabH[jh_[it]];
        const int ro = pz * Hmax * Ws + (t.first - ohlo) * Ws + jw_[it];
        float pv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 6; ++c)
          if (t.w[c] != 0.f) {
            const float4 u0 = Ra[ro + c * Ws], u1 = Rb[ro + c * Ws];
            pv[0] = fmaf(t.w[c], u0.x, pv[0]); pv[1] = fmaf(t.w[c], u0.y, pv[1]); pv[2] = fmaf(t.w[c], u0.z, pv[2]); pv[3] = fmaf(t.w[c], u0.w, pv[3]);
            pv[4] = fmaf(t.w[c], u1.x, pv[4]); pv[5] = fmaf(t.w[c], u1.y, pv[5]); pv[6] = fmaf(t.w[c], u1.z, pv[6]); pv[7] = fmaf(t.w[c], u1.w, pv[7]);
          }
        if (wa != 0.f) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[0][it][i] = fmaf(wa, pv[i], acc[0][it][i]);
        }
        if (wb != 0.f) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[1][it][i] = fmaf(wb, pv[i], acc[1][it][i]);
        }
      }
    }
    __syncthreads();                       // R is overwritten by the next planes
  }
  flush0();
  flush0();
}

static int upsample2_bwd_fused(const grad_t* gdst, int gdst_chunks, int gdst_off, int C, Dims sd, grad_t* gsrc, int num_sms, cudaStream_t st) {
  const int K = C / 8, TH = sd.W <= 64 ? 8 : 4;
  const long long Vo = (long long)sd.D * sd.H * sd.W * 8;
  const size_t smem = (size_t)(sd.W + TH) * sizeof(AxisTaps) + (size_t)2 * (2 * TH + 6) * sd.W * 32;   // tables + two fine planes of w-reduced rows
  static bool attr_set[64] = {};
  int dev = 0;
  SEUNET_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { seunet_set_error("upsample2_bwd: device index %d", dev); return 1; }
  if (!attr_set[dev]) {
    SEUNET_CUDA_CHECK(cudaFuncSetAttribute(upsample2_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr_set[dev] = true;
  }
  const int htiles = (sd.H + TH - 1) / TH;
  // Split the d range so that the waves of (two blocks per SM) are full: a split costs ~5 re-read fine planes of halo, a
  // partial wave costs a whole block time.  Pick the split with the smallest (waves x planes per block).
  const long long base = (long long)htiles * sd.N * K, slots = 2LL * num_sms;
  int dsplit = 1;
  long long best = -1;
  for (int s = 1; s <= std::max(1, sd.D / 2); ++s) {
    const int per = (sd.D + s - 1) / s, ns = (sd.D + per - 1) / per;
    const long long cost = ((base * ns + slots - 1) / slots) * (2LL * per + 5);
    if (best < 0 || cost < best) { best = cost; dsplit = ns; }
  }
  const int dper = (sd.D + dsplit - 1) / dsplit;
  dsplit = (sd.D + dper - 1) / dper;
  upsample2_bwd_fused_kernel<<<dim3((unsigned)dsplit, (unsigned)htiles, (unsigned)(sd.N * K)), kUbThreads, smem, st>>>(
      gdst + (size_t)gdst_off * Vo * 8, (long long)gdst_chunks * Vo * 8, Vo * 8, sd, gsrc, K, TH, dper);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// gsrc[n][C/8][sd] = Up2^T(gdst slice); tmp1 >= N*C*(2D*2H*W) floats, tmp2 >= N*C*(2D*H*W) floats (three-pass fallback only)
int launch_upsample2_bwd_sep(const grad_t* gdst, int gdst_chunks, int gdst_off, int C, Dims sd, grad_t* gsrc, float* tmp1, float* tmp2,
                             cudaStream_t st) {
  const int K = C / 8, Do = sd.D * 2, Ho = sd.H * 2, Wo = sd.W * 2;
  const long long Vo = (long long)Do * Ho * Wo;
  if (Do * Ho > 65535 || sd.N * K > 65535) return launch_upsample2_bwd(gdst, gdst_chunks, gdst_off, C, sd, gsrc, st);
  static const bool three_pass = getenv("SEUNET_UPBWD_3PASS") != nullptr && atoi(getenv("SEUNET_UPBWD_3PASS")) != 0;   // A/B switch
  if (!three_pass && sd.W <= 128 && sd.W * (sd.W <= 64 ? 8 : 4) <= kUbThreads * kUbMaxItems && sd.D >= 2 && sd.H >= 2 && sd.W >= 2) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return upsample2_bwd_fused(gdst, gdst_chunks, gdst_off, C, sd, gsrc, sms, st);
  }
  const float4* in0 = reinterpret_cast<const float4*>(gdst + (size_t)gdst_off * Vo * 8);
  auto blocks = [](long long n) { return (unsigned)((n + 255) / 256); };
  // along w: [Do*Ho][Wo -> W][2]   (a line has only 2*W float4 outputs: right-size the block)
  const int tw = std::min(256, (sd.W * 2 + 31) / 32 * 32);
  adjoint_axis_vec_kernel<<<dim3((unsigned)((sd.W * 2 + tw - 1) / tw), Do * Ho, sd.N * K), tw, 0, st>>>(
      in0, (long long)gdst_chunks * Vo * 2, Vo * 2, reinterpret_cast<float4*>(tmp1), K, Do * Ho, Wo, sd.W, 2);
  // along h: [Do][Ho -> H][W*2]
  const long long p1 = (long long)Do * Ho * sd.W * 2;   // float4 per plane of tmp1
  adjoint_axis_vec_kernel<<<dim3(blocks((long long)sd.H * sd.W * 2), Do, sd.N * K), 256, 0, st>>>(
      reinterpret_cast<const float4*>(tmp1), p1 * K, p1, reinterpret_cast<float4*>(tmp2), K, Do, Ho, sd.H, sd.W * 2);
  // along d: [1][Do -> D][H*W*2]
  const long long p2 = (long long)Do * sd.H * sd.W * 2;
  adjoint_axis_vec_kernel<<<dim3(blocks((long long)sd.D * sd.H * sd.W * 2), 1, sd.N * K), 256, 0, st>>>(
      reinterpret_cast<const float4*>(tmp2), p2 * K, p2, reinterpret_cast<float4*>(gsrc), K, 1, Do, sd.D, sd.H * sd.W * 2);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
#else
int launch_upsample2_bwd_sep(const grad_t* gdst, int gdst_chunks, int gdst_off, int C, Dims sd, grad_t* gsrc, float*, float*,
                             cudaStream_t st) {
  return launch_upsample2_bwd(gdst, gdst_chunks, gdst_off, C, sd, gsrc, st);
}
#endif
