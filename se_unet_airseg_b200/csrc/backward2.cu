// Separable adjoint of the trilinear (align_corners=True) up-sampling folded into the heads: Up = U_d (x) U_h (x) U_w, so
// Up^T is three 1-D adjoint passes (w, then h, then d) over shrinking tensors instead of one (2*2^l+1)^3 gather per voxel.
#include "backward.cuh"
#include <algorithm>

__device__ __forceinline__ float axis_w1(int o, int j, int in_size, int out_size) {   // weight of source j in output o
  const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float src = scale * (float)o;
  const int i0 = (int)src;
  const int i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  const float l1 = src - (float)i0;
  return (i0 == j ? 1.f - l1 : 0.f) + (i1 == j ? l1 : 0.f);
}

// tensor viewed as [outer][axis][inner]; out[outer][j][inner] = sum_o w(o -> j) * in[outer][o][inner]
__global__ void __launch_bounds__(256) adjoint_axis_kernel(const float* __restrict__ in, float* __restrict__ out, long long outer,
                                                           int out_len /*fine*/, int in_len /*coarse*/, long long inner) {
  const long long total = outer * in_len * inner;
  const float inv = in_len > 1 ? (float)(out_len - 1) / (float)(in_len - 1) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long q = i % inner;
    const int j = (int)((i / inner) % in_len);
    const long long o_ = i / (inner * in_len);
    const int lo = max(0, (int)floorf((float)(j - 1) * inv) - 1);
    const int hi = min(out_len - 1, (int)ceilf((float)(j + 1) * inv) + 1);
    const float* p = in + (o_ * out_len) * inner + q;
    float acc = 0.f;
    for (int o = lo; o <= hi; ++o) {
      const float w = axis_w1(o, j, in_len, out_len);
      if (w != 0.f) acc = fmaf(w, __ldg(p + (long long)o * inner), acc);
    }
    out[i] = acc;
  }
}

static int adjoint_axis(const float* in, float* out, long long outer, int fine, int coarse, long long inner, cudaStream_t st) {
  const long long total = outer * coarse * inner;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  adjoint_axis_kernel<<<blocks, 256, 0, st>>>(in, out, outer, fine, coarse, inner);
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

// gsrc[n][C/8][sd] = Up2^T(gdst slice); tmp1 >= N*C*(2D*2H*W) floats, tmp2 >= N*C*(2D*H*W) floats
int launch_upsample2_bwd_sep(const grad_t* gdst, int gdst_chunks, int gdst_off, int C, Dims sd, grad_t* gsrc, float* tmp1, float* tmp2,
                             cudaStream_t st) {
  const int K = C / 8, Do = sd.D * 2, Ho = sd.H * 2, Wo = sd.W * 2;
  const long long Vo = (long long)Do * Ho * Wo;
  if (Do * Ho > 65535 || sd.N * K > 65535) return launch_upsample2_bwd(gdst, gdst_chunks, gdst_off, C, sd, gsrc, st);
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
