// Block-tiled versions of the two trilinear kernels of the forward pass.
//  * upsample2: one block owns one output (d,h) line (block-uniform d/h parameters, shared-memory w table, packed blend).
//  * head: one block owns 8 output h-lines; the w interpolation tables of the three coarse levels are built once per
//    block, the d/h parameters are block-uniform per line.
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

constexpr int kUpLinesPerBlock = 8;

// trilinear x2 of C channels into a chunk slot (SE_UNet.py:136-138, 214/220/226).  One block owns 8 output (d,h) lines:
// d/h parameters are block-uniform, the w table lives in shared memory, and the 8-tap blend runs in packed 16-bit FMAs
// (weights are a convex combination, so the packed accumulation adds ~2 storage ulps; a fp32 blend spends 2/3 of the
// kernel's issue slots on 16->32-bit conversions and runs 2.5x slower).
__global__ void __launch_bounds__(256) upsample2_line_kernel(const act_t* __restrict__ src, Dims sd, act_t* __restrict__ dst,
                                                             int dst_chunks, int dst_off, int C8) {
  extern __shared__ int s_tab[];   // [Wo] i0*8 | [Wo] i1*8 | [Wo] l1 (float bits)
  const int Wo = sd.W * 2, Ho = sd.H * 2, Do = sd.D * 2;
  const int od = blockIdx.y, n = blockIdx.z;
  for (int w = threadIdx.x; w < Wo; w += blockDim.x) {
    const Lerp1 lw = lerp1_ac(w, sd.W, Wo);
    s_tab[w] = lw.i0 * 8; s_tab[Wo + w] = lw.i1 * 8; s_tab[2 * Wo + w] = __float_as_int(lw.l1);
  }
  const Lerp1 ld = lerp1_ac(od, sd.D, Do);
  __syncthreads();
  const size_t Vs = (size_t)sd.D * sd.H * sd.W, Vo = Vs * 8;
  for (int ol = 0; ol < kUpLinesPerBlock; ++ol) {   // the w table is amortised over several output lines
    const int oh = blockIdx.x * kUpLinesPerBlock + ol;
    if (oh >= Ho) break;
    const Lerp1 lh = lerp1_ac(oh, sd.H, Ho);
    size_t line[4];
    float wq[4];
    line[0] = ((size_t)ld.i0 * sd.H + lh.i0) * sd.W * 8; line[1] = ((size_t)ld.i0 * sd.H + lh.i1) * sd.W * 8;
    line[2] = ((size_t)ld.i1 * sd.H + lh.i0) * sd.W * 8; line[3] = ((size_t)ld.i1 * sd.H + lh.i1) * sd.W * 8;
    wq[0] = ld.l0 * lh.l0; wq[1] = ld.l0 * lh.l1; wq[2] = ld.l1 * lh.l0; wq[3] = ld.l1 * lh.l1;
    const size_t oline = ((size_t)od * Ho + oh) * Wo * 8;
    for (int item = threadIdx.x; item < Wo * C8; item += blockDim.x) {
      const int ow = item % Wo, k = item / Wo;
      const int a0 = s_tab[ow], a1 = s_tab[Wo + ow];
      const float l1 = __int_as_float(s_tab[2 * Wo + ow]), l0 = 1.f - l1;
      const act_t* sp = src + ((size_t)n * C8 + k) * Vs * 8;
      act2_t acc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 u0 = *reinterpret_cast<const uint4*>(sp + line[q] + a0);
        const uint4 u1 = *reinterpret_cast<const uint4*>(sp + line[q] + a1);
        const uint32_t w0b = pack_act2(wq[q] * l0, wq[q] * l0), w1b = pack_act2(wq[q] * l1, wq[q] * l1);
        const act2_t w0 = *reinterpret_cast<const act2_t*>(&w0b), w1 = *reinterpret_cast<const act2_t*>(&w1b);
        const act2_t* v0 = reinterpret_cast<const act2_t*>(&u0);
        const act2_t* v1 = reinterpret_cast<const act2_t*>(&u1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (q == 0) acc[i] = __hmul2(w0, v0[i]);
          else acc[i] = __hfma2(w0, v0[i], acc[i]);
          acc[i] = __hfma2(w1, v1[i], acc[i]);
        }
      }
      *reinterpret_cast<uint4*>(dst + ((size_t)n * dst_chunks + dst_off + k) * Vo * 8 + oline + (size_t)ow * 8) =
          *reinterpret_cast<const uint4*>(acc);
    }
  }
}

int launch_upsample2(const act_t* src, int C, Dims sd, act_t* dst, int dst_chunks, int dst_off, cudaStream_t st) {
  dim3 grid((sd.H * 2 + kUpLinesPerBlock - 1) / kUpLinesPerBlock, sd.D * 2, sd.N);
  const size_t smem = (size_t)sd.W * 2 * 3 * sizeof(int);
  upsample2_line_kernel<<<grid, 256, smem, st>>>(src, sd, dst, dst_chunks, dst_off, C / 8);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

constexpr int kHeadLines = 8;

// head: pred = bias + T(S) + Up2(T(S/2)) + Up4(T(S/4)) [+ Up8(T(S/8))]   (SE_UNet.py:232-233 with the side branches folded)
__global__ void __launch_bounds__(256) head_tile_kernel(const __grid_constant__ HeadArgs a) {
  extern __shared__ int s_tab[];   // per level l=1..3: [W] i0 | [W] i1 | [W] l1
  const Dims d = a.d;
  const int hy0 = blockIdx.x * kHeadLines, dz = blockIdx.y, n = blockIdx.z;
  const int W = d.W;
  for (int t = threadIdx.x; t < 3 * W; t += blockDim.x) {
    const int l = t / W + 1, w = t % W;
    const Lerp1 lw = lerp1_ac(w, d.W >> l, d.W);
    int* tab = s_tab + (l - 1) * 3 * W;
    tab[w] = lw.i0; tab[W + w] = lw.i1; tab[2 * W + w] = __float_as_int(lw.l1);
  }
  __syncthreads();
  const size_t V = (size_t)d.D * d.H * d.W;
  const float b0 = a.bias0[0], b1 = a.bias1[0];
  for (int item = threadIdx.x; item < kHeadLines * W; item += blockDim.x) {
    const int wx = item % W, hy = hy0 + item / W;   // a warp stays within one line when W is a multiple of 32
    if (hy >= d.H) break;
    const size_t idx = (size_t)n * V + ((size_t)dz * d.H + hy) * d.W + wx;
    float p0 = b0 + a.T0[0][idx];
    float p1 = b1 + a.T1[0][idx];
#pragma unroll
    for (int l = 1; l < 4; ++l) {
      const int Ds = d.D >> l, Hs = d.H >> l, Ws = d.W >> l;
      const Lerp1 ldd = lerp1_ac(dz, Ds, d.D), lhh = lerp1_ac(hy, Hs, d.H);
      const int* tab = s_tab + (l - 1) * 3 * W;
      const int i0 = tab[wx], i1 = tab[W + wx];
      const float l1 = __int_as_float(tab[2 * W + wx]), l0 = 1.f - l1;
      const size_t Vs = (size_t)Ds * Hs * Ws;
      const size_t o00 = ((size_t)ldd.i0 * Hs + lhh.i0) * Ws, o01 = ((size_t)ldd.i0 * Hs + lhh.i1) * Ws;
      const size_t o10 = ((size_t)ldd.i1 * Hs + lhh.i0) * Ws, o11 = ((size_t)ldd.i1 * Hs + lhh.i1) * Ws;
      const float w00 = ldd.l0 * lhh.l0, w01 = ldd.l0 * lhh.l1, w10 = ldd.l1 * lhh.l0, w11 = ldd.l1 * lhh.l1;
      const float* t0 = a.T0[l] + (size_t)n * Vs;
      p0 += w00 * (l0 * __ldg(t0 + o00 + i0) + l1 * __ldg(t0 + o00 + i1)) + w01 * (l0 * __ldg(t0 + o01 + i0) + l1 * __ldg(t0 + o01 + i1)) +
            w10 * (l0 * __ldg(t0 + o10 + i0) + l1 * __ldg(t0 + o10 + i1)) + w11 * (l0 * __ldg(t0 + o11 + i0) + l1 * __ldg(t0 + o11 + i1));
      if (l < 3) {
        const float* t1 = a.T1[l] + (size_t)n * Vs;
        p1 += w00 * (l0 * __ldg(t1 + o00 + i0) + l1 * __ldg(t1 + o00 + i1)) + w01 * (l0 * __ldg(t1 + o01 + i0) + l1 * __ldg(t1 + o01 + i1)) +
              w10 * (l0 * __ldg(t1 + o10 + i0) + l1 * __ldg(t1 + o10 + i1)) + w11 * (l0 * __ldg(t1 + o11 + i0) + l1 * __ldg(t1 + o11 + i1));
      }
    }
    a.pred0[idx] = p0;
    a.pred1[idx] = p1;
  }
}

int launch_head(const HeadArgs& a, cudaStream_t st) {
  dim3 grid((a.d.H + kHeadLines - 1) / kHeadLines, a.d.D, a.d.N);
  const size_t smem = (size_t)a.d.W * 9 * sizeof(int);
  head_tile_kernel<<<grid, 256, smem, st>>>(a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
