// Block-tiled versions of the two trilinear kernels of the forward pass.
//  * upsample2: one block owns one output (d,h) line (block-uniform d/h parameters, shared-memory w table, packed blend).
//  * head: one block owns 8 output h-lines; the w interpolation tables of the three coarse levels are built once per
//    block, the d/h parameters are block-uniform per line.
#include "pointwise.cuh"
#include <cstdlib>

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

#ifndef SEUNET_UP_LINES
#define SEUNET_UP_LINES 8
#endif
constexpr int kUpLinesPerBlock = SEUNET_UP_LINES;

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

// Separable version: phase 1 blends the four (d,h) source lines of every output line into shared memory (block-uniform
// weights, one pass over the coarse line), phase 2 does the w blend from shared memory.  ~2x fewer instructions per
// output voxel than blending all 8 taps per output element.  The kernel is issue-bound, so every interpolation
// parameter comes from a small shared-memory table (computing them per element in registers measured 7 % slower).
__global__ void __launch_bounds__(256) upsample2_sep_kernel(const act_t* __restrict__ src, Dims sd, act_t* __restrict__ dst,
                                                            int dst_chunks, int dst_off, int C8) {
  extern __shared__ __align__(16) uint8_t s_up[];
  const int Ws = sd.W, Wo = sd.W * 2, Ho = sd.H * 2, Do = sd.D * 2;
  uint4* s_line = reinterpret_cast<uint4*>(s_up);                       // [line][C8][Ws] blended coarse lines
  int* s_tab = reinterpret_cast<int*>(s_line + kUpLinesPerBlock * C8 * Ws);   // [Wo] i0 | [Wo] i1 | [Wo] packed l0 | [Wo] packed l1
  __shared__ __align__(16) int s_lo[kUpLinesPerBlock][4];
  __shared__ __align__(16) uint32_t s_wq[kUpLinesPerBlock][4];
  const int od = blockIdx.y, n = blockIdx.z;
  const Lerp1 ld = lerp1_ac(od, sd.D, Do);
  for (int w = threadIdx.x; w < Wo; w += blockDim.x) {
    const Lerp1 lw = lerp1_ac(w, Ws, Wo);
    s_tab[w] = lw.i0; s_tab[Wo + w] = lw.i1;
    s_tab[2 * Wo + w] = (int)pack_act2(lw.l0, lw.l0); s_tab[3 * Wo + w] = (int)pack_act2(lw.l1, lw.l1);
  }
  if (threadIdx.x < kUpLinesPerBlock) {
    const int oh = min(blockIdx.x * kUpLinesPerBlock + (int)threadIdx.x, Ho - 1);
    const Lerp1 lh = lerp1_ac(oh, sd.H, Ho);
    s_lo[threadIdx.x][0] = (ld.i0 * sd.H + lh.i0) * Ws; s_lo[threadIdx.x][1] = (ld.i0 * sd.H + lh.i1) * Ws;
    s_lo[threadIdx.x][2] = (ld.i1 * sd.H + lh.i0) * Ws; s_lo[threadIdx.x][3] = (ld.i1 * sd.H + lh.i1) * Ws;
    s_wq[threadIdx.x][0] = pack_act2(ld.l0 * lh.l0, ld.l0 * lh.l0); s_wq[threadIdx.x][1] = pack_act2(ld.l0 * lh.l1, ld.l0 * lh.l1);
    s_wq[threadIdx.x][2] = pack_act2(ld.l1 * lh.l0, ld.l1 * lh.l0); s_wq[threadIdx.x][3] = pack_act2(ld.l1 * lh.l1, ld.l1 * lh.l1);
  }
  __syncthreads();
  const size_t Vs = (size_t)sd.D * sd.H * Ws, Vo = Vs * 8;
  const int nlines = min(kUpLinesPerBlock, Ho - (int)blockIdx.x * kUpLinesPerBlock);
  // Warp-per-(line, chunk) mapping: no runtime integer divisions in the inner loops (they cost more than the blend).
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // phase 1
  for (int lk = warp; lk < nlines * C8; lk += 8) {
    {
      const int line = lk / C8, k = lk - line * C8;   // one division per (line, chunk) pair, not per element
      const uint4* sp = reinterpret_cast<const uint4*>(src + ((size_t)n * C8 + k) * Vs * 8);
      const int4 lo = *reinterpret_cast<const int4*>(s_lo[line]);
      const uint4 wqv = *reinterpret_cast<const uint4*>(s_wq[line]);
      const int o[4] = {lo.x, lo.y, lo.z, lo.w};
      const uint32_t wq32[4] = {wqv.x, wqv.y, wqv.z, wqv.w};
      for (int w = lane; w < Ws; w += 32) {
        uint4 u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = __ldg(sp + o[q] + w);
        act2_t acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const act2_t wq = *reinterpret_cast<const act2_t*>(&wq32[q]);
          const act2_t* v = reinterpret_cast<const act2_t*>(&u[q]);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = q == 0 ? __hmul2(wq, v[i]) : __hfma2(wq, v[i], acc[i]);
        }
        s_line[lk * Ws + w] = *reinterpret_cast<const uint4*>(acc);
      }
    }
  }
  __syncthreads();
  // phase 2
  for (int lk = warp; lk < nlines * C8; lk += 8) {
    {
      const int line = lk / C8, k = lk - line * C8;
      const int oh = blockIdx.x * kUpLinesPerBlock + line;
      act_t* dp = dst + ((size_t)n * dst_chunks + dst_off + k) * Vo * 8 + ((size_t)od * Ho + oh) * Wo * 8;
      for (int ow = lane; ow < Wo; ow += 32) {
        const uint4 u0 = s_line[lk * Ws + s_tab[ow]], u1 = s_line[lk * Ws + s_tab[Wo + ow]];
        const act2_t w0 = *reinterpret_cast<const act2_t*>(&s_tab[2 * Wo + ow]), w1 = *reinterpret_cast<const act2_t*>(&s_tab[3 * Wo + ow]);
        const act2_t* v0 = reinterpret_cast<const act2_t*>(&u0);
        const act2_t* v1 = reinterpret_cast<const act2_t*>(&u1);
        act2_t acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = __hfma2(w1, v1[i], __hmul2(w0, v0[i]));
        *reinterpret_cast<uint4*>(dp + (size_t)ow * 8) = *reinterpret_cast<const uint4*>(acc);
      }
    }
  }
}

// Round 2: warp-private version.  Warp w of a block owns output line w (one (od, oh) pair) for all channel chunks: the blended
// coarse line goes through a warp-private shared-memory row, so the two phases need __syncwarp only (the sep kernel above
// waits for the slowest warp of the block between its phases), and the four coarse loads of chunk k+1 are in flight while
// chunk k is blended and stored (ncu on the sep kernel: 0.8 IPC per SM and 58 % of the DRAM peak on the 128^3 instance,
// 3 % on the 32^3 one - eight dependent load round trips per warp).  The w table is one 128-bit entry per output voxel.
// Arithmetic (operand order, packed 16-bit FMAs) is the sep kernel's: results are bit-identical.
template <int NW>   // ceil(Ws / 32)
__global__ void __launch_bounds__(256, NW <= 2 ? 4 : 2) upsample2_warp_kernel(const act_t* __restrict__ src, Dims sd, act_t* __restrict__ dst,
                                                             int dst_chunks, int dst_off, int C8) {
  extern __shared__ __align__(16) uint8_t s_up[];
  const int Ws = sd.W, Wo = sd.W * 2, Ho = sd.H * 2, Do = sd.D * 2;
  int4* s_tab = reinterpret_cast<int4*>(s_up);          // [Wo] {i0, i1, packed l0, packed l1}
  uint4* s_row = reinterpret_cast<uint4*>(s_tab + Wo);  // [8 warps][Ws] blended coarse line of the warp's current chunk
  const int od = blockIdx.y, n = blockIdx.z;
  for (int w = threadIdx.x; w < Wo; w += blockDim.x) {
    const Lerp1 lw = lerp1_ac(w, Ws, Wo);
    s_tab[w] = make_int4(lw.i0, lw.i1, (int)pack_act2(lw.l0, lw.l0), (int)pack_act2(lw.l1, lw.l1));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oh = blockIdx.x * 8 + warp;
  if (oh >= Ho) return;
  const Lerp1 ld = lerp1_ac(od, sd.D, Do), lh = lerp1_ac(oh, sd.H, Ho);
  const int o[4] = {(ld.i0 * sd.H + lh.i0) * Ws, (ld.i0 * sd.H + lh.i1) * Ws, (ld.i1 * sd.H + lh.i0) * Ws, (ld.i1 * sd.H + lh.i1) * Ws};
  const uint32_t wq32[4] = {pack_act2(ld.l0 * lh.l0, ld.l0 * lh.l0), pack_act2(ld.l0 * lh.l1, ld.l0 * lh.l1),
                            pack_act2(ld.l1 * lh.l0, ld.l1 * lh.l0), pack_act2(ld.l1 * lh.l1, ld.l1 * lh.l1)};
  const size_t Vs = (size_t)sd.D * sd.H * Ws, Vo = Vs * 8;
  uint4* row = s_row + warp * Ws;
  uint4 nxt[NW][4];
  auto prefetch = [&](int k) {
    const uint4* sp = reinterpret_cast<const uint4*>(src + ((size_t)n * C8 + k) * Vs * 8);
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const int w = lane + 32 * j;
      if (w < Ws) {
#pragma unroll
        for (int q = 0; q < 4; ++q) nxt[j][q] = __ldg(sp + o[q] + w);
      }
    }
  };
  prefetch(0);
  for (int k = 0; k < C8; ++k) {
    // phase 1: blend the four (d, h) source lines of this chunk into the warp's row
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const int w = lane + 32 * j;
      if (w < Ws) {
        act2_t acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const act2_t wq = *reinterpret_cast<const act2_t*>(&wq32[q]);
          const act2_t* v = reinterpret_cast<const act2_t*>(&nxt[j][q]);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = q == 0 ? __hmul2(wq, v[i]) : __hfma2(wq, v[i], acc[i]);
        }
        row[w] = *reinterpret_cast<const uint4*>(acc);
      }
    }
    if (k + 1 < C8) prefetch(k + 1);
    __syncwarp();
    // phase 2: w blend from the row
    act_t* dp = dst + ((size_t)n * dst_chunks + dst_off + k) * Vo * 8 + ((size_t)od * Ho + oh) * Wo * 8;
#pragma unroll
    for (int j = 0; j < 2 * NW; ++j) {
      const int ow = lane + 32 * j;
      if (ow < Wo) {
        const int4 t = s_tab[ow];
        const uint4 u0 = row[t.x], u1 = row[t.y];
        const act2_t w0 = *reinterpret_cast<const act2_t*>(&t.z), w1 = *reinterpret_cast<const act2_t*>(&t.w);
        const act2_t* v0 = reinterpret_cast<const act2_t*>(&u0);
        const act2_t* v1 = reinterpret_cast<const act2_t*>(&u1);
        act2_t acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = __hfma2(w1, v1[i], __hmul2(w0, v0[i]));
        *reinterpret_cast<uint4*>(dp + (size_t)ow * 8) = *reinterpret_cast<const uint4*>(acc);
      }
    }
    __syncwarp();   // the row is overwritten by the next chunk
  }
}

template <int NW>
static void launch_upsample2_warp(const act_t* src, Dims sd, act_t* dst, int dst_chunks, int dst_off, int C8, cudaStream_t st) {
  dim3 grid((sd.H * 2 + 7) / 8, sd.D * 2, sd.N);
  const size_t smem = (size_t)sd.W * 2 * sizeof(int4) + (size_t)8 * sd.W * sizeof(uint4);
  upsample2_warp_kernel<NW><<<grid, 256, smem, st>>>(src, sd, dst, dst_chunks, dst_off, C8);
}

int launch_upsample2(const act_t* src, int C, Dims sd, act_t* dst, int dst_chunks, int dst_off, cudaStream_t st) {
  dim3 grid((sd.H * 2 + kUpLinesPerBlock - 1) / kUpLinesPerBlock, sd.D * 2, sd.N);
  const int C8 = C / 8;
  const size_t smem = (size_t)kUpLinesPerBlock * C8 * sd.W * 16 + (size_t)sd.W * 2 * 4 * sizeof(int);
  static const bool warp_version = !(getenv("SEUNET_UP_WARP") && atoi(getenv("SEUNET_UP_WARP")) == 0);
  const int nw = (sd.W + 31) / 32;
  if (warp_version && nw <= 4) {
    switch (nw) {
      case 1: launch_upsample2_warp<1>(src, sd, dst, dst_chunks, dst_off, C8, st); break;
      case 2: launch_upsample2_warp<2>(src, sd, dst, dst_chunks, dst_off, C8, st); break;
      case 3: launch_upsample2_warp<3>(src, sd, dst, dst_chunks, dst_off, C8, st); break;
      default: launch_upsample2_warp<4>(src, sd, dst, dst_chunks, dst_off, C8, st); break;
    }
  } else if (smem <= 48 * 1024) {
    upsample2_sep_kernel<<<grid, 256, smem, st>>>(src, sd, dst, dst_chunks, dst_off, C8);
  } else {   // very wide volumes: direct 8-tap blend
    upsample2_line_kernel<<<grid, 256, (size_t)sd.W * 2 * 3 * sizeof(int), st>>>(src, sd, dst, dst_chunks, dst_off, C8);
  }
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

constexpr int kHeadLines = 8;

// head: pred = bias + T(S) + Up2(T(S/2)) + Up4(T(S/4)) [+ Up8(T(S/8))]   (SE_UNet.py:232-233 with the side branches folded)
// One block owns 8 output h-lines of one d-plane.  Phase 1 interpolates, for every line, level and head, the coarse line
// along d and h (block-/line-uniform weights, 4 coalesced loads per coarse element) into shared memory; phase 2 is the w
// lerp from shared memory: 2 LDS + 2 FMA per (voxel, level, head) instead of 8 gathers + 7 lerps.  Issue-bound: the
// interpolation parameters come from shared-memory tables (per-element recomputation measured 40 % slower).
// SINK (sliding-window plans): only head 1 (prediction.py:103 drops p0), and instead of storing the logits the block adds
// sigmoid(pred1) in fixed point to the volume accumulator - same arithmetic as window_accumulate_kernel (window.cu), so the
// fused and the two-kernel paths agree bit for bit.
template <bool SINK>
__global__ void __launch_bounds__(256) head_tile_kernel(const __grid_constant__ HeadArgs a) {
  extern __shared__ float s_head[];   // lines: [level 1..3][head][line][W >> l]; then w tables [level][W]{i0 | i1<<16, l1}
  __shared__ __align__(16) int s_hoff[3][kHeadLines][4];
  __shared__ __align__(16) float s_hw[3][kHeadLines][4];
  const Dims d = a.d;
  const int hy0 = blockIdx.x * kHeadLines, dz = blockIdx.y, n = blockIdx.z;
  const int W = d.W;
  int base[4][2];
  int off = 0;
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    base[l][0] = off; off += kHeadLines * (W >> l);
    base[l][1] = off; if (l < 3) off += kHeadLines * (W >> l);
  }
  int2* s_wtab = reinterpret_cast<int2*>(s_head + off);
  if (threadIdx.x < 3 * kHeadLines) {
    const int l = threadIdx.x / kHeadLines + 1, line = threadIdx.x % kHeadLines;
    const int Ds = d.D >> l, Hs = d.H >> l, Ws = W >> l;
    const Lerp1 ld = lerp1_ac(dz, Ds, d.D), lh = lerp1_ac(min(hy0 + line, d.H - 1), Hs, d.H);
    s_hoff[l - 1][line][0] = (ld.i0 * Hs + lh.i0) * Ws; s_hoff[l - 1][line][1] = (ld.i0 * Hs + lh.i1) * Ws;
    s_hoff[l - 1][line][2] = (ld.i1 * Hs + lh.i0) * Ws; s_hoff[l - 1][line][3] = (ld.i1 * Hs + lh.i1) * Ws;
    s_hw[l - 1][line][0] = ld.l0 * lh.l0; s_hw[l - 1][line][1] = ld.l0 * lh.l1;
    s_hw[l - 1][line][2] = ld.l1 * lh.l0; s_hw[l - 1][line][3] = ld.l1 * lh.l1;
  }
  for (int t = threadIdx.x; t < 3 * W; t += blockDim.x) {
    const int l = t / W + 1, w = t - (l - 1) * W;
    const Lerp1 lw = lerp1_ac(w, W >> l, W);
    s_wtab[t] = make_int2(lw.i0 | (lw.i1 << 16), __float_as_int(lw.l1));
  }
  __syncthreads();
  // Warp w owns output line w of the block in both phases: no runtime integer divisions in the inner loops.
  const int line = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // phase 1: (d,h)-interpolated coarse lines
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    const int Ws = W >> l;
    const size_t Vs = (size_t)(d.D >> l) * (d.H >> l) * Ws;
    if (SINK && l == 3) break;
    const float* t0 = SINK ? nullptr : a.T0[l] + (size_t)n * Vs;
    const float* t1 = l < 3 ? a.T1[l] + (size_t)n * Vs : nullptr;
    const int4 o = *reinterpret_cast<const int4*>(s_hoff[l - 1][line]);
    const float4 q = *reinterpret_cast<const float4*>(s_hw[l - 1][line]);
    for (int w = lane; w < Ws; w += 32) {
      if (!SINK)
        s_head[base[l][0] + line * Ws + w] =
            q.x * __ldg(t0 + o.x + w) + q.y * __ldg(t0 + o.y + w) + q.z * __ldg(t0 + o.z + w) + q.w * __ldg(t0 + o.w + w);
      if (l < 3)
        s_head[base[l][1] + line * Ws + w] =
            q.x * __ldg(t1 + o.x + w) + q.y * __ldg(t1 + o.y + w) + q.z * __ldg(t1 + o.z + w) + q.w * __ldg(t1 + o.w + w);
    }
  }
  __syncwarp();   // a line is produced and consumed by the same warp
  // phase 2: w lerp
  const size_t V = (size_t)d.D * d.H * d.W;
  const float b0 = a.bias0[0], b1 = a.bias1[0];
  const int hy = hy0 + line;
  if (hy >= d.H) return;
  const size_t row = (size_t)n * V + ((size_t)dz * d.H + hy) * d.W;
  unsigned int* accp = SINK ? a.acc + ((size_t)(a.s[n][0] + dz) * a.Y + (a.s[n][1] + hy)) * a.Z + a.s[n][2] : nullptr;
  for (int wx = lane; wx < W; wx += 32) {
    float p0 = SINK ? 0.f : b0 + __ldg(a.T0[0] + row + wx);
    float p1 = b1 + __ldg(a.T1[0] + row + wx);
#pragma unroll
    for (int l = 1; l < 4; ++l) {
      if (SINK && l == 3) break;
      const int Ws = W >> l;
      const int2 tw = s_wtab[(l - 1) * W + wx];
      const int i0 = tw.x & 0xffff, i1 = tw.x >> 16;
      const float l1 = __int_as_float(tw.y);
      if (!SINK) {
        const float* s0 = s_head + base[l][0] + line * Ws;
        const float v0 = s0[i0];
        p0 += fmaf(l1, s0[i1] - v0, v0);
      }
      if (l < 3) {
        const float* s1 = s_head + base[l][1] + line * Ws;
        const float u0 = s1[i0];
        p1 += fmaf(l1, s1[i1] - u0, u0);
      }
    }
    if (SINK) {
      float pr = 1.f / (1.f + expf(-p1));
      pr = fminf(fmaxf(pr, 0.f), 1.f);
      atomicAdd(accp + wx, __float2uint_rn(pr * a.acc_scale));
    } else {
      a.pred0[row + wx] = p0;
      a.pred1[row + wx] = p1;
    }
  }
}

int launch_head(const HeadArgs& a, cudaStream_t st) {
  if (a.d.W >= 65536) { seunet_set_error("head: W too large"); return 1; }
  dim3 grid((a.d.H + kHeadLines - 1) / kHeadLines, a.d.D, a.d.N);
  size_t floats = 0;
  for (int l = 1; l < 4; ++l) floats += (size_t)kHeadLines * (a.d.W >> l) * (l < 3 ? 2 : 1);
  const size_t smem = floats * sizeof(float) + (size_t)3 * a.d.W * sizeof(int2);
  if (a.acc) head_tile_kernel<true><<<grid, 256, smem, st>>>(a);
  else head_tile_kernel<false><<<grid, 256, smem, st>>>(a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
