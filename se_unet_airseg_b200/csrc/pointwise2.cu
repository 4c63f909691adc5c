// Line-per-block versions of the two trilinear kernels of the forward pass.  One thread block owns one output (d,h)
// line: the d/h interpolation parameters are block-uniform and the w parameters come from a small shared-memory table,
// so the per-output work is 8 vector loads + the blend, without per-thread index/division arithmetic.
#include "pointwise.cuh"

struct Lerp1 { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp1 lerp1_ac(int dst, int in_size, int out_size) {   // ATen area_pixel_compute_source_index, align_corners=True
  const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float src = scale * (float)dst;
  Lerp1 r;
  r.i0 = (int)src;
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

// trilinear x2 of C channels into a chunk slot (SE_UNet.py:136-138, 214/220/226)
__global__ void __launch_bounds__(256) upsample2_line_kernel(const act_t* __restrict__ src, Dims sd, act_t* __restrict__ dst,
                                                             int dst_chunks, int dst_off, int C8) {
  extern __shared__ int s_tab[];   // [Wo] i0*8 | [Wo] i1*8 | [Wo] l1 (float bits)
  const int Wo = sd.W * 2, Ho = sd.H * 2, Do = sd.D * 2;
  const int oh = blockIdx.x, od = blockIdx.y, n = blockIdx.z;
  for (int w = threadIdx.x; w < Wo; w += blockDim.x) {
    const Lerp1 lw = lerp1_ac(w, sd.W, Wo);
    s_tab[w] = lw.i0 * 8; s_tab[Wo + w] = lw.i1 * 8; s_tab[2 * Wo + w] = __float_as_int(lw.l1);
  }
  const Lerp1 ld = lerp1_ac(od, sd.D, Do), lh = lerp1_ac(oh, sd.H, Ho);
  __syncthreads();
  const size_t Vs = (size_t)sd.D * sd.H * sd.W, Vo = Vs * 8;
  const size_t line00 = ((size_t)ld.i0 * sd.H + lh.i0) * sd.W * 8, line01 = ((size_t)ld.i0 * sd.H + lh.i1) * sd.W * 8;
  const size_t line10 = ((size_t)ld.i1 * sd.H + lh.i0) * sd.W * 8, line11 = ((size_t)ld.i1 * sd.H + lh.i1) * sd.W * 8;
  const float w00 = ld.l0 * lh.l0, w01 = ld.l0 * lh.l1, w10 = ld.l1 * lh.l0, w11 = ld.l1 * lh.l1;
  const size_t oline = ((size_t)od * Ho + oh) * Wo * 8;
  for (int item = threadIdx.x; item < Wo * C8; item += blockDim.x) {
    const int ow = item % Wo, k = item / Wo;
    const int a0 = s_tab[ow], a1 = s_tab[Wo + ow];
    const float l1 = __int_as_float(s_tab[2 * Wo + ow]), l0 = 1.f - l1;
    const act_t* sp = src + ((size_t)n * C8 + k) * Vs * 8;
    float acc[8], f0[8], f1[8];
    chunk_to_floats(ld_chunk(sp + line00 + a0), f0); chunk_to_floats(ld_chunk(sp + line00 + a1), f1);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = w00 * (l0 * f0[i] + l1 * f1[i]);
    chunk_to_floats(ld_chunk(sp + line01 + a0), f0); chunk_to_floats(ld_chunk(sp + line01 + a1), f1);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(w01, l0 * f0[i] + l1 * f1[i], acc[i]);
    chunk_to_floats(ld_chunk(sp + line10 + a0), f0); chunk_to_floats(ld_chunk(sp + line10 + a1), f1);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(w10, l0 * f0[i] + l1 * f1[i], acc[i]);
    chunk_to_floats(ld_chunk(sp + line11 + a0), f0); chunk_to_floats(ld_chunk(sp + line11 + a1), f1);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(w11, l0 * f0[i] + l1 * f1[i], acc[i]);
    st_chunk(dst + ((size_t)n * dst_chunks + dst_off + k) * Vo * 8 + oline + (size_t)ow * 8, floats_to_chunk(acc));
  }
}

int launch_upsample2(const act_t* src, int C, Dims sd, act_t* dst, int dst_chunks, int dst_off, cudaStream_t st) {
  dim3 grid(sd.H * 2, sd.D * 2, sd.N);
  const size_t smem = (size_t)sd.W * 2 * 3 * sizeof(int);
  upsample2_line_kernel<<<grid, 256, smem, st>>>(src, sd, dst, dst_chunks, dst_off, C / 8);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// head: pred = bias + T(S) + Up2(T(S/2)) + Up4(T(S/4)) [+ Up8(T(S/8))]   (SE_UNet.py:232-233 with the side branches folded)
__global__ void __launch_bounds__(128) head_line_kernel(const __grid_constant__ HeadArgs a) {
  extern __shared__ int s_tab[];   // per level l=1..3: [W] i0 | [W] i1 | [W] l1
  const Dims d = a.d;
  const int hy = blockIdx.x, dz = blockIdx.y, n = blockIdx.z;
  const int W = d.W;
  for (int t = threadIdx.x; t < 3 * W; t += blockDim.x) {
    const int l = t / W + 1, w = t % W;
    const Lerp1 lw = lerp1_ac(w, d.W >> l, d.W);
    int* tab = s_tab + (l - 1) * 3 * W;
    tab[w] = lw.i0; tab[W + w] = lw.i1; tab[2 * W + w] = __float_as_int(lw.l1);
  }
  __syncthreads();
  const size_t V = (size_t)d.D * d.H * d.W;
  const size_t line = ((size_t)dz * d.H + hy) * d.W;
  // block-uniform d/h interpolation per level
  size_t off[3][4];
  float wq[3][4];
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    const int Ds = d.D >> l, Hs = d.H >> l, Ws = d.W >> l;
    const Lerp1 ldd = lerp1_ac(dz, Ds, d.D), lhh = lerp1_ac(hy, Hs, d.H);
    off[l - 1][0] = ((size_t)ldd.i0 * Hs + lhh.i0) * Ws; off[l - 1][1] = ((size_t)ldd.i0 * Hs + lhh.i1) * Ws;
    off[l - 1][2] = ((size_t)ldd.i1 * Hs + lhh.i0) * Ws; off[l - 1][3] = ((size_t)ldd.i1 * Hs + lhh.i1) * Ws;
    wq[l - 1][0] = ldd.l0 * lhh.l0; wq[l - 1][1] = ldd.l0 * lhh.l1; wq[l - 1][2] = ldd.l1 * lhh.l0; wq[l - 1][3] = ldd.l1 * lhh.l1;
  }
  const float b0 = a.bias0[0], b1 = a.bias1[0];
  for (int wx = threadIdx.x; wx < W; wx += blockDim.x) {
    float p0 = b0 + a.T0[0][(size_t)n * V + line + wx];
    float p1 = b1 + a.T1[0][(size_t)n * V + line + wx];
#pragma unroll
    for (int l = 1; l < 4; ++l) {
      const int* tab = s_tab + (l - 1) * 3 * W;
      const int i0 = tab[wx], i1 = tab[W + wx];
      const float l1 = __int_as_float(tab[2 * W + wx]), l0 = 1.f - l1;
      const size_t Vs = (size_t)(d.D >> l) * (d.H >> l) * (d.W >> l);
      const float* t0 = a.T0[l] + (size_t)n * Vs;
      float s0 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) s0 = fmaf(wq[l - 1][q], l0 * __ldg(t0 + off[l - 1][q] + i0) + l1 * __ldg(t0 + off[l - 1][q] + i1), s0);
      p0 += s0;
      if (l < 3) {
        const float* t1 = a.T1[l] + (size_t)n * Vs;
        float s1 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) s1 = fmaf(wq[l - 1][q], l0 * __ldg(t1 + off[l - 1][q] + i0) + l1 * __ldg(t1 + off[l - 1][q] + i1), s1);
        p1 += s1;
      }
    }
    a.pred0[(size_t)n * V + line + wx] = p0;
    a.pred1[(size_t)n * V + line + wx] = p1;
  }
}

int launch_head(const HeadArgs& a, cudaStream_t st) {
  dim3 grid(a.d.H, a.d.D, a.d.N);
  const size_t smem = (size_t)a.d.W * 9 * sizeof(int);
  head_line_kernel<<<grid, 128, smem, st>>>(a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
