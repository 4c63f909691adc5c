// Fused loss reductions / gradients (train.py:51-76 and the stage combinations train.py:597-599, 432-435, 238-243) and a
// fused AdamW step over the flat parameter buffer (torch.optim.AdamW as used in train.py:188/386/569).
//
// The three losses are ratios of BATCH-GLOBAL sums, so the forward half only produces partial sums
//   per head h in {en (pred0), de (pred1)}:  [I=sum p t, P=sum p, T=sum t, A=sum w (p+1e-4)^0.7 t, Bs=sum w (0.2p+0.8t),
//                                            Ia=sum w p s^2, Ja=sum w (p s + s), pad]
// (p = sigmoid(logit)); data-parallel ranks all-reduce these 16 doubles, then the backward half turns the GLOBAL sums
// into the scalar loss and d loss / d logit.  This is what makes N-rank training equal 1-rank training on the
// concatenated batch (the reference computes the loss on the gathered batch on GPU 0).
#include "../../include/seunet_b200.h"
#include "common.cuh"
#include <algorithm>

constexpr int kLossSums = 8;

__device__ __forceinline__ float sigmoid_acc(float z) { return 1.f / (1.f + expf(-z)); }

__global__ void __launch_bounds__(256) loss_sums_kernel(const float* __restrict__ pen, const float* __restrict__ pde,
                                                        const float* __restrict__ label, const float* __restrict__ weight,
                                                        const float* __restrict__ skel, long long V, int stage,
                                                        double* __restrict__ sums, double* __restrict__ per_sample) {
  const int b = blockIdx.y;
  float acc[2][7];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[h][i] = 0.f;
  const size_t base = (size_t)b * V;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const float t = label[base + v];
    const float w = weight ? weight[base + v] : 1.f;
    const float s = skel ? skel[base + v] : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float p = sigmoid_acc(h == 0 ? pen[base + v] : pde[base + v]);
      acc[h][0] += p * t; acc[h][1] += p; acc[h][2] += t;
      if (stage >= 2) {
        acc[h][3] += w * powf(p + 1e-4f, 0.7f) * t;
        acc[h][4] += w * (0.2f * p + 0.8f * t);
      }
      if (stage == 3) {
        acc[h][5] += w * p * s * s;
        acc[h][6] += w * (p * s + s);
      }
    }
  }
  __shared__ double sh[8][14];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const double r = warp_sum_d((double)acc[h][i]);
      if (lane == 0) sh[warp][h * 7 + i] = r;
    }
  __syncthreads();
  if (threadIdx.x < 14) {
    double r = 0.0;
    for (int w = 0; w < 8; ++w) r += sh[w][threadIdx.x];
    const int h = threadIdx.x / 7, i = threadIdx.x % 7;
    atomicAdd(sums + h * kLossSums + i, r);
    if (per_sample && (i == 3 || i == 4)) atomicAdd(per_sample + ((size_t)b * 2 + h) * 2 + (i - 3), r);
  }
}

extern "C" int seunet_loss_sums(int stage, const float* pred_en, const float* pred_de, const float* label,
                                const float* weight, const float* skel, int batch, int64_t voxels_per_sample, double* sums,
                                double* per_sample, seunet_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (stage < 1 || stage > 3) { seunet_set_error("loss: stage %d unsupported (1..3)", stage); return 1; }
  if (stage >= 2 && !weight) { seunet_set_error("loss: stage %d needs the weight tensor", stage); return 1; }
  if (stage == 3 && !skel) { seunet_set_error("loss: stage 3 needs the skeleton tensor"); return 1; }
  SEUNET_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * kLossSums, st));
  if (per_sample) SEUNET_CUDA_CHECK(cudaMemsetAsync(per_sample, 0, sizeof(double) * batch * 4, st));
  dim3 grid((unsigned)std::min<long long>((voxels_per_sample + 255) / 256, 148 * 2), batch);
  loss_sums_kernel<<<grid, 256, 0, st>>>(pred_en, pred_de, label, weight, skel, voxels_per_sample, stage, sums, per_sample);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// loss terms from the (global) sums
__device__ __forceinline__ double dice_l(const double* s) { return 1.0 - (2.0 * s[0] + 1.0) / (s[1] + s[2] + 1.0); }
__device__ __forceinline__ double gul_l(const double* s) { return 1.0 - (s[3] + 1.0) / (s[4] + 1.0); }
__device__ __forceinline__ double atr_l(const double* s) { return 1.0 - (s[5] + 1.0) / (s[6] + 1.0); }

__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ pen, const float* __restrict__ pde,
                                                        const float* __restrict__ label, const float* __restrict__ weight,
                                                        const float* __restrict__ skel, long long n, int stage,
                                                        const double* __restrict__ sums, float* __restrict__ den,
                                                        float* __restrict__ dde, float* __restrict__ loss_out) {
  __shared__ float c[2][6];   // per head: dice a, dice b, gul a, gul b, atr a, atr b
  if (threadIdx.x < 2) {
    const double* s = sums + threadIdx.x * kLossSums;
    // head weights of the stage combination: stage 1: dice_de + dice_en ; stage >= 2: gul_de + 0.5 gul_en (+ 0.5 atr each)
    const double wd = 1.0, wg = threadIdx.x == 0 ? 0.5 : 1.0, wa = 0.5;
    const double dd = s[1] + s[2] + 1.0;
    c[threadIdx.x][0] = (float)(wd * -2.0 / dd);
    c[threadIdx.x][1] = (float)(wd * (2.0 * s[0] + 1.0) / (dd * dd));
    c[threadIdx.x][2] = (float)(wg * -0.7 / (s[4] + 1.0));
    c[threadIdx.x][3] = (float)(wg * 0.2 * (s[3] + 1.0) / ((s[4] + 1.0) * (s[4] + 1.0)));
    c[threadIdx.x][4] = (float)(wa * -1.0 / (s[6] + 1.0));
    c[threadIdx.x][5] = (float)(wa * (s[5] + 1.0) / ((s[6] + 1.0) * (s[6] + 1.0)));
  }
  if (blockIdx.x == 0 && threadIdx.x == 32 && loss_out) {
    const double* e = sums;
    const double* d = sums + kLossSums;
    double l;
    if (stage == 1) l = dice_l(d) + dice_l(e);
    else {
      l = gul_l(d) + 0.5 * gul_l(e);
      if (stage == 3) l += 0.5 * (atr_l(e) + atr_l(d));
    }
    loss_out[0] = (float)l;
  }
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float t = label[i];
    const float w = weight ? weight[i] : 1.f;
    const float s = skel ? skel[i] : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float p = sigmoid_acc(h == 0 ? pen[i] : pde[i]);
      float g;
      if (stage == 1) g = c[h][0] * t + c[h][1];
      else {
        g = c[h][2] * w * t * powf(p + 1e-4f, -0.3f) + c[h][3] * w;
        if (stage == 3) g += c[h][4] * w * s * s + c[h][5] * w * s;
      }
      g *= p * (1.f - p);
      if (h == 0) den[i] = g; else dde[i] = g;
    }
  }
}

extern "C" int seunet_loss_grad(int stage, const float* pred_en, const float* pred_de, const float* label,
                                const float* weight, const float* skel, int64_t n, const double* sums, float* dpred_en,
                                float* dpred_de, float* loss_out, seunet_stream_t stream) {
  if (stage < 1 || stage > 3) { seunet_set_error("loss: stage %d unsupported (1..3)", stage); return 1; }
  loss_grad_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      pred_en, pred_de, label, weight, skel, n, stage, sums, dpred_en, dpred_de, loss_out);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// torch.optim.AdamW (decoupled weight decay, amsgrad=False) over one flat buffer; entries in [skip_off, skip_off+skip_len)
// are left untouched (dc62.conv1.weight has grad None in the reference, so AdamW never updates or decays it).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2_sqrt, float gscale, long long skip_off,
                                                    long long skip_len) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (i >= skip_off && i < skip_off + skip_len) continue;
    const float gi = g[i] * gscale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

extern "C" int seunet_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                 int64_t skip_off, int64_t skip_len, seunet_stream_t stream) {
  if (step < 1) { seunet_set_error("adamw: step must be >= 1"); return 1; }
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adamw_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale, skip_off, skip_len);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
