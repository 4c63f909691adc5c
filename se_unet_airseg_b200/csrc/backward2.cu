// Separable adjoint of the trilinear (align_corners=True) up-sampling folded into the heads: Up = U_d (x) U_h (x) U_w, so
// Up^T is three 1-D adjoint passes (w, then h, then d) over shrinking tensors instead of one (2*2^l+1)^3 gather per voxel.
#include "backward.cuh"

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
