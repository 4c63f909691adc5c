// Fused "SSE apply + CAT 1x1x1 conv" pass (inference plans only).
//
// The last SSE block of every encoder level (ec3 / ec6 / ec9 / ec12) and decoder level (dc2 / dc4) feeds exactly one consumer:
// the CATConv 1x1x1 conv over the concat [that block's output | earlier blocks of the level] (SE_UNet.py:186, 195, 204,
// 212, 218, 224).  Unfused, the apply pass writes the gated activations to the concat buffer and the tcgen05 conv reads
// the whole concat back - at full resolution (ec33) that is 134 MB written and 268 MB read per 128^3 window for a
// 7.5 GFLOP GEMV-like contraction that runs at 5-12 % tensor-pipe utilisation and HBM speed.  Here the apply pass keeps
// its own output channels in registers/shared memory, reads only the OTHER concat slices, and does the 1x1x1 contraction
// itself with warp-level mma.sync (m16n8k16, fp16 x fp16 -> fp32: the work is HBM-bound, the legacy tensor path is ample),
// writing the raw CAT-conv output and its InstanceNorm statistics exactly as the tcgen05 conv would have.
// In training plans the unfused path stays: the backward pass needs the block's output in the concat buffer.
#include "pointwise.cuh"
#include <cstring>

constexpr float kInEps3 = 1e-5f;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_m16n8k16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
#ifdef SEUNET_ACT_BF16
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#endif
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// C: channels of the SSE block, GATES: 1|2, KCAT: padded input channels of the CAT conv (multiple of 16), NOUT: its outputs.
template <int C, int GATES, int KCAT, int NOUT>
__global__ void __launch_bounds__(256, C == 64 ? 1 : 2) apply_sse_cat_kernel(const __grid_constant__ SseArgs a, const __grid_constant__ CatFuseArgs f) {
  constexpr int ROW = KCAT * 2 + 16;     // bytes per shared-memory row (+16: ldmatrix rows land in different bank groups)
  __shared__ __align__(16) float s_mean[C], s_rstd[C], s_wse[C], s_wse2[C], s_weff[C];
  extern __shared__ __align__(16) uint8_t dsm[];
  uint8_t* sA = dsm;                     // [256 voxels][KCAT] storage type, row-major (mma A operand)
  uint8_t* sW = dsm + 256 * ROW;         // [NOUT][KCAT]  (mma B operand, "col-major")
  float* s_red = reinterpret_cast<float*>(sW + NOUT * ROW);   // [8 warps][2 * NOUT] statistics partials
  const int n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < C; c += blockDim.x) {
    const double s = a.stats[((size_t)n * a.stats_c + c) * 2], q = a.stats[((size_t)n * a.stats_c + c) * 2 + 1];
    const double mean = s / (double)a.V;
    double var = q / (double)a.V - mean * mean;
    if (var < 0) var = 0;
    s_mean[c] = (float)mean;
    s_rstd[c] = (float)(1.0 / sqrt(var + (double)kInEps3));
    s_wse[c] = a.wse[c];
    s_wse2[c] = GATES == 2 ? a.wse2[c] : 0.f;
    s_weff[c] = a.weff[(size_t)n * 64 + c];
  }
  // CAT weights (fp32 [NOUT][cin_real], concat channel order == chunk order of the concat buffer) -> storage type [NOUT][KCAT]
  for (int i = tid; i < NOUT * KCAT; i += blockDim.x) {
    const int o = i / KCAT, k = i % KCAT;
    const float w = k < f.cin_real ? f.w[(size_t)o * f.cin_real + k] : 0.f;
    *reinterpret_cast<act_t*>(sW + o * ROW + k * 2) = f2act(w);
  }
  __syncthreads();
  auto ld8 = [](const float* sm, int k, float* r) {
    *reinterpret_cast<float4*>(r) = *reinterpret_cast<const float4*>(sm + k * 8);
    *reinterpret_cast<float4*>(r + 4) = *reinterpret_cast<const float4*>(sm + k * 8 + 4);
  };
  // statistics of the CAT-conv output: this thread's accumulator columns are channels nt*8 + (lane%4)*2 + {0,1}
  float ssum[NOUT / 8][2], ssq[NOUT / 8][2];
#pragma unroll
  for (int nt = 0; nt < NOUT / 8; ++nt) { ssum[nt][0] = ssum[nt][1] = ssq[nt][0] = ssq[nt][1] = 0.f; }
  const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
  const long long ntiles = (a.V + 255) / 256;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long v = tile * 256 + tid;
    const bool live = v < a.V;
    uint8_t* rowp = sA + tid * ROW;
    if (live) {
      float* tp = a.T + (size_t)n * a.V + v;
      const float t_old = a.t_init ? 0.f : *tp;
      Chunk8 in[C / 8];
#pragma unroll
      for (int k = 0; k < C / 8; ++k) in[k] = ld_chunk_stream(a.raw + (((size_t)n * a.raw_chunks + k) * a.V + v) * 8);
      // the other slices of the concat (written by earlier blocks of the level) go straight to the operand tile
      Chunk8 oth[(KCAT - C) / 8];
#pragma unroll
      for (int k = 0; k < (KCAT - C) / 8; ++k)
        oth[k] = (C / 8 + k) < f.cat_real_chunks
                     ? ld_chunk_stream(f.cat + (((size_t)n * f.cat_chunks + C / 8 + k) * a.V + v) * 8)
                     : Chunk8{{0u, 0u, 0u, 0u}};
      float e[C];
      float g1 = 0.f;
#pragma unroll
      for (int k = 0; k < C / 8; ++k) {
        float fv[8], mean[8], rstd[8], wse[8];
        chunk_to_floats(in[k], fv);
        ld8(s_mean, k, mean); ld8(s_rstd, k, rstd); ld8(s_wse, k, wse);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float t = lrelu_((fv[i] - mean[i]) * rstd[i]);
          e[k * 8 + i] = t;
          g1 = fmaf(wse[i], t, g1);
        }
      }
      g1 = sigmoidf_(g1);
      if (GATES == 2) {
        float g2 = 0.f;
#pragma unroll
        for (int k = 0; k < C / 8; ++k) {
          float wse2[8];
          ld8(s_wse2, k, wse2);
#pragma unroll
          for (int i = 0; i < 8; ++i) { e[k * 8 + i] *= g1; g2 = fmaf(wse2[i], e[k * 8 + i], g2); }
        }
        g1 = sigmoidf_(g2);
      }
      float t = a.wcst[n];
#pragma unroll
      for (int k = 0; k < C / 8; ++k) {
        float weff[8];
        ld8(s_weff, k, weff);
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[k * 8 + i] *= g1; t = fmaf(weff[i], e[k * 8 + i], t); }
      }
      *tp = t_old + t;
#pragma unroll
      for (int k = 0; k < C / 8; ++k) st_chunk(rowp + k * 16, floats_to_chunk(e + k * 8));   // same rounding as the stored concat slice
#pragma unroll
      for (int k = 0; k < (KCAT - C) / 8; ++k) st_chunk(rowp + (C / 8 + k) * 16, oth[k]);
    } else {
#pragma unroll
      for (int k = 0; k < KCAT / 8; ++k) st_chunk(rowp + k * 16, Chunk8{{0u, 0u, 0u, 0u}});
    }
    __syncwarp();
    // ---- this warp's 32 voxels x KCAT  times  KCAT x NOUT, fp32 accumulate
    float acc[2][NOUT / 8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NOUT / 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KCAT / 16; ++ks) {
      uint32_t af[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = warp * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(sA_u + r * ROW + (ks * 16 + (lane >> 4) * 8) * 2, af[mt][0], af[mt][1], af[mt][2], af[mt][3]);
      }
#pragma unroll
      for (int np = 0; np < NOUT / 16; ++np) {   // two 8-wide output tiles per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        const int o = np * 16 + (lane & 7) + (lane >> 4) * 8;
        ldmatrix_x4(sW_u + o * ROW + (ks * 16 + ((lane >> 3) & 1) * 8) * 2, b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_m16n8k16(acc[mt][np * 2], af[mt][0], af[mt][1], af[mt][2], af[mt][3], b0, b1);
          mma_m16n8k16(acc[mt][np * 2 + 1], af[mt][0], af[mt][1], af[mt][2], af[mt][3], b2, b3);
        }
      }
    }
    __syncwarp();   // all lanes done with the operand rows before the next tile overwrites them
    // ---- raw output (chunk planes, storage type) + statistics from the fp32 accumulators
    const long long vbase = tile * 256 + warp * 32;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const long long vv = vbase + mt * 16 + half * 8 + (lane >> 2);
        if (vv < a.V) {
#pragma unroll
          for (int nt = 0; nt < NOUT / 8; ++nt) {
            const float x0 = acc[mt][nt][half * 2], x1 = acc[mt][nt][half * 2 + 1];
            ssum[nt][0] += x0; ssum[nt][1] += x1; ssq[nt][0] += x0 * x0; ssq[nt][1] += x1 * x1;
            uint32_t* dst = reinterpret_cast<uint32_t*>(f.out + (((size_t)n * f.out_chunks + nt) * a.V + vv) * 8) + (lane & 3);
            *dst = pack_act2(x0, x1);
          }
        }
      }
  }
  // ---- statistics: reduce over the 8 row lanes of a warp, then over the 8 warps, one fp64 atomic per (channel, moment) and block
#pragma unroll
  for (int nt = 0; nt < NOUT / 8; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float s = ssum[nt][j], q = ssq[nt][j];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if (lane < 4) {
        const int ch = nt * 8 + lane * 2 + j;
        s_red[warp * 2 * NOUT + ch] = s;
        s_red[warp * 2 * NOUT + NOUT + ch] = q;
      }
    }
  __syncthreads();
  if (tid < 2 * NOUT) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += (double)s_red[w * 2 * NOUT + tid];
    const int ch = tid % NOUT, mom = tid / NOUT;
    atomicAdd(f.out_stats + ((size_t)n * f.out_stats_c + ch) * 2 + mom, tot);
  }
}

template <int C, int GATES, int KCAT, int NOUT>
static int launch_t(int N, const SseArgs& a, const CatFuseArgs& f, int num_sms, cudaStream_t st) {
  constexpr int ROW = KCAT * 2 + 16;
  const int smem = 256 * ROW + NOUT * ROW + 8 * 2 * NOUT * 4;
  static bool attr_set[64] = {};
  int dev = 0;
  SEUNET_CUDA_CHECK(cudaGetDevice(&dev));
  if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || !attr_set[dev])) {
    SEUNET_CUDA_CHECK(cudaFuncSetAttribute(apply_sse_cat_kernel<C, GATES, KCAT, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const long long ntiles = (a.V + 255) / 256;
  // persistent-ish: a few blocks per SM loop over the voxel tiles so that the statistics cost one set of atomics per block
  const int per_sample = (int)std::min<long long>(ntiles, std::max(1, (num_sms * 2 + N - 1) / N));
  dim3 grid((unsigned)per_sample, N);
  apply_sse_cat_kernel<C, GATES, KCAT, NOUT><<<grid, 256, smem, st>>>(a, f);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_apply_sse_cat(int C, int N, const SseArgs& a, const CatFuseArgs& f, int num_sms, cudaStream_t st) {
  if (a.dest != nullptr) { seunet_set_error("apply_sse_cat: the fused pass does not store the block output"); return 1; }
  const int gates = a.wse2 ? 2 : 1;
  if (C == 32 && gates == 1 && f.kcat == 64 && f.nout == 32) return launch_t<32, 1, 64, 32>(N, a, f, num_sms, st);     // ec3 -> ec33
  if (C == 64 && gates == 2 && f.kcat == 128 && f.nout == 64) return launch_t<64, 2, 128, 64>(N, a, f, num_sms, st);   // ec6 -> ec63, dc2 -> dc22
  if (C == 64 && gates == 2 && f.kcat == 192 && f.nout == 64) return launch_t<64, 2, 192, 64>(N, a, f, num_sms, st);   // ec9 -> ec93, ec12 -> ec123
  if (C == 32 && gates == 2 && f.kcat == 96 && f.nout == 32) return launch_t<32, 2, 96, 32>(N, a, f, num_sms, st);     // dc4 -> dc42
  seunet_set_error("apply_sse_cat: no instance for C=%d gates=%d K=%d N=%d", C, gates, f.kcat, f.nout);
  return 1;
}
