// Sliding-window inference support kernels: HU windowing of the CT volume,
// device-side overlap accumulation of window probabilities, mean + threshold.
// Reference: prediction.py:39-49, 69-111.
#include "../../include/seunet_b200.h"
#include "common.cuh"

// prediction.py:69 (img - 1024) and two_channel() (prediction.py:39-49), evaluated in float64 exactly
// like the numpy reference and rounded once to fp32 (x = torch.from_numpy(img.astype(np.float32))).
template <typename T>
__global__ void __launch_bounds__(256) hu_windows_kernel(const T* __restrict__ img, long long n, long long cstride, double offset,
                                                         float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)img[i] + offset;
    const double a = fmin(fmax(v, -1024.0), 1024.0);
    const double b = fmin(fmax(v, -1000.0), 500.0);
    out[i] = (float)((a + 1024.0) / 2048.0);
    out[cstride + i] = (float)((b + 1000.0) / 1500.0);
  }
}

extern "C" int seunet_hu_windows_slab(const void* img, int dtype, int64_t nvox, int64_t channel_stride, double offset, float* out,
                                      seunet_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nvox <= 0) return 0;
  if (channel_stride < nvox) { seunet_set_error("hu_windows: channel stride smaller than the slab"); return 1; }
  const int blocks = (int)((nvox + 255) / 256 < 148 * 16 ? (nvox + 255) / 256 : 148 * 16);
  if (dtype == 0) hu_windows_kernel<short><<<blocks, 256, 0, st>>>((const short*)img, nvox, channel_stride, offset, out);
  else if (dtype == 1) hu_windows_kernel<float><<<blocks, 256, 0, st>>>((const float*)img, nvox, channel_stride, offset, out);
  else { seunet_set_error("hu_windows: dtype %d unsupported (0=int16, 1=float32)", dtype); return 1; }
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int seunet_hu_windows(const void* img, int dtype, int64_t nvox, double offset, float* out,
                                 seunet_stream_t stream) {
  return seunet_hu_windows_slab(img, dtype, nvox, nvox, offset, out, stream);
}

struct WinStarts { int n; int s[32][3]; };

// pred[..window..] += sigmoid(logits)   (prediction.py:103-106; pred_num is analytic, see finalize)
// The sum is kept in FIXED POINT (units of 2^-acc_log2, 32-bit; the caller picks acc_log2 so that the largest window
// overlap count times 2^acc_log2 stays below 2^31 - 26 for the reference's 128/64 grid with <= 27 overlaps): integer
// addition is associative, so the result does not depend on the order in which windows arrive - across CUDA streams,
// across window batches, or across the ranks of a patch-sharded volume whose partial volumes are summed by NCCL.  The
// reference accumulates in float64 (prediction.py:78, 106); the quantisation here (<= 2^-27 per window) is below the
// fp32 resolution of the probabilities themselves near 0.5.
__global__ void __launch_bounds__(256) window_accumulate_kernel(const float* __restrict__ logits, const __grid_constant__ WinStarts ws,
                                                                int cd, int ch, int cw, unsigned int* __restrict__ acc, int X, int Y,
                                                                int Z, int apply_sigmoid, float acc_scale) {
  const int b = blockIdx.y;
  const long long V = (long long)cd * ch * cw;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int w = (int)(v % cw), h = (int)((v / cw) % ch), d = (int)(v / ((long long)cw * ch));
  float p = logits[(size_t)b * V + v];
  if (apply_sigmoid) p = 1.f / (1.f + expf(-p));
  p = fminf(fmaxf(p, 0.f), 1.f);
  atomicAdd(acc + ((size_t)(ws.s[b][0] + d) * Y + (ws.s[b][1] + h)) * Z + ws.s[b][2] + w, __float2uint_rn(p * acc_scale));
}

extern "C" int seunet_window_accumulate(const float* logits, const int* starts /*host [B][3]*/, int B, int cd, int ch,
                                        int cw, uint32_t* acc, int X, int Y, int Z, int apply_sigmoid, int acc_log2,
                                        seunet_stream_t stream) {
  if (B < 1 || B > 32) { seunet_set_error("window_accumulate: batch %d out of range (1..32)", B); return 1; }
  if (acc_log2 < 8 || acc_log2 > 30) { seunet_set_error("window_accumulate: acc_log2 %d out of range (8..30)", acc_log2); return 1; }
  WinStarts ws;
  ws.n = B;
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < 3; ++k) {
      ws.s[b][k] = starts[b * 3 + k];
      const int lim = k == 0 ? X - cd : (k == 1 ? Y - ch : Z - cw);
      if (ws.s[b][k] < 0 || ws.s[b][k] > lim) { seunet_set_error("window_accumulate: window %d out of bounds", b); return 1; }
    }
  const long long V = (long long)cd * ch * cw;
  dim3 grid((unsigned)((V + 255) / 256), B);
  window_accumulate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, ws, cd, ch, cw, acc, X, Y, Z, apply_sigmoid,
                                                                   (float)(1u << acc_log2));
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// mean = acc / count (prediction.py:109) and mask = mean >= threshold.  count[x][y][z] is the product of the
// per-axis window coverage counts (the windows form a full grid, prediction.py:83-100).
__global__ void __launch_bounds__(256) window_finalize_kernel(unsigned int* __restrict__ acc, const int* __restrict__ cx,
                                                              const int* __restrict__ cy, const int* __restrict__ cz, int X,
                                                              int Y, int Z, float thr, unsigned char* __restrict__ mask,
                                                              int write_mean, double acc_inv) {
  const long long V = (long long)X * Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(i % Z), y = (int)((i / Z) % Y), x = (int)(i / ((long long)Z * Y));
    const double cnt = (double)(cx[x] * cy[y] * cz[z]);
    const float m = (float)((double)acc[i] * acc_inv / cnt);
    if (write_mean) acc[i] = __float_as_uint(m);
    if (mask) mask[i] = m >= thr ? 1 : 0;
  }
}

extern "C" int seunet_window_finalize(uint32_t* acc, const int* counts_dev /*device [X+Y+Z]*/, int X, int Y, int Z,
                                      float threshold, unsigned char* mask, int write_mean, int acc_log2,
                                      seunet_stream_t stream) {
  if (acc_log2 < 8 || acc_log2 > 30) { seunet_set_error("window_finalize: acc_log2 %d out of range (8..30)", acc_log2); return 1; }
  window_finalize_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(acc, counts_dev, counts_dev + X, counts_dev + X + Y, X, Y,
                                                                     Z, threshold, mask, write_mean, 1.0 / (double)(1u << acc_log2));
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
