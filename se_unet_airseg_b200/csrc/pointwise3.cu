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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ Chunk8 lds_chunk(uint32_t addr) {
  Chunk8 c;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.u[0]), "=r"(c.u[1]), "=r"(c.u[2]), "=r"(c.u[3]) : "r"(addr));
  return c;
}

constexpr int kCatTileGroup = 4;   // tiles per statistics group (see below)
// cp.async ring depth per variant: as deep as two resident blocks per SM allow (227 KB)
__host__ __device__ constexpr int cat_stages(int KCAT) { return KCAT <= 64 ? 2 : 1; }

// C: channels of the SSE block, GATES: 1|2, KCAT: padded input channels of the CAT conv (multiple of 16), NOUT: its outputs.
//
// A warp owns 32 voxels of a 256-voxel tile from the loads to the stores (warp-level synchronisation only), so the 16
// warps of an SM drift apart and overlap each other's load / arithmetic / mma phases.  With 128 registers per thread only
// two blocks fit an SM, so the memory latency is covered by a per-thread cp.async ring instead of by occupancy (ncu on the
// first version: 49 % of DRAM peak, half of all stall samples on the first use of a loaded value): every thread keeps the
// raw values, the OTHER concat slices and the head accumulator of its voxel of the next STAGES tiles in flight in a private
// shared-memory slot (no synchronisation needed: a thread reads back only what it copied itself).  Per tile and warp:
//   1. wait for the oldest ring stage; copy the other concat slices from it into the mma operand rows;
//   2. InstanceNorm + LeakyReLU + gate(s) + head fold on the block's own channels, rounded to the storage type into the
//      operand rows;   3. refill the stage with the tile STAGES ahead;   4. 32 x KCAT x NOUT mma.sync;
//   5. store the raw CAT-conv output + statistics.
// Statistics: fp32 per thread over a GROUP of kCatTileGroup consecutive tiles (a fixed function of the group), then fp64 -
// blocks walk whole groups, so the sums do not depend on the grid (batch size) beyond fp64 rounding, as in conv_tc.cu.
template <int C, int GATES, int KCAT, int NOUT>
__global__ void __launch_bounds__(256, C == 64 ? 1 : 2) apply_sse_cat_kernel(const __grid_constant__ SseArgs a, const __grid_constant__ CatFuseArgs f) {
  constexpr int ROW = KCAT * 2 + 16;     // bytes per shared-memory row (+16: ldmatrix rows land in different bank groups)
  constexpr int NT = NOUT / 8;
  constexpr int NCH = KCAT / 8;          // chunks per voxel in the ring
  constexpr int STAGES = cat_stages(KCAT);
  constexpr int STAGE_BYTES = NCH * 512 + 128;   // per warp: [chunk][lane] 16 B + [lane] 4 B head accumulator
  __shared__ __align__(16) float s_mean[C], s_rstd[C], s_wse[C], s_wse2[C], s_weff[C];
  extern __shared__ __align__(16) uint8_t dsm[];
  uint8_t* sA = dsm;                     // [256 voxels][KCAT] storage type, row-major (mma A operand)
  uint8_t* sW = dsm + 256 * ROW;         // [NOUT][KCAT]  (mma B operand, "col-major")
  const int n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ring_u = smem_u32(dsm + 256 * ROW + NOUT * ROW) + (uint32_t)(warp * STAGES * STAGE_BYTES + lane * 16);
  for (int c = tid; c < C; c += blockDim.x) {
    const double s = a.stats[((size_t)n * a.stats_c + c) * 2], q = a.stats[((size_t)n * a.stats_c + c) * 2 + 1];
    const double mean = s / (double)a.V;
    double var = q / (double)a.V - mean * mean;
    if (var < 0) var = 0;
    const double rstd = 1.0 / sqrt(var + (double)kInEps3);
    s_mean[c] = (float)(-mean * rstd);     // (y - mean) * rstd as one fma: y * rstd + (-mean * rstd)
    s_rstd[c] = (float)rstd;
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
  uint8_t* rowp = sA + tid * ROW;
  // padding chunks of the concat (56 -> 64 channels) are never loaded: zero them once
#pragma unroll
  for (int k = C / 8; k < NCH; ++k)
    if (k >= f.cat_real_chunks) st_chunk(rowp + k * 16, Chunk8{{0u, 0u, 0u, 0u}});
  __syncthreads();
  auto ld8 = [](const float* sm, int k, float* r) {
    *reinterpret_cast<float4*>(r) = *reinterpret_cast<const float4*>(sm + k * 8);
    *reinterpret_cast<float4*>(r + 4) = *reinterpret_cast<const float4*>(sm + k * 8 + 4);
  };
  // statistics of the CAT-conv output: this thread's accumulator columns are channels nt*8 + (lane%4)*2 + {0,1}
  float ssum[NT][2], ssq[NT][2];
  double dsum[NT][2], dsq[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) { ssum[nt][j] = ssq[nt][j] = 0.f; dsum[nt][j] = dsq[nt][j] = 0.0; }
  const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
  const long long ntiles = (a.V + 255) / 256;
  const long long ngroups = (ntiles + kCatTileGroup - 1) / kCatTileGroup;
  // per-thread byte pointers at tile 0 (a tile is 256 voxels = 4096 bytes of every chunk plane); planes are V * 16 bytes apart.
  // (ncu, first version: a quarter of the executed instructions were 64-bit index arithmetic recomputed per load / store)
  const size_t plane_bytes = (size_t)a.V * 16;
  const uint8_t* raw_p = reinterpret_cast<const uint8_t*>(a.raw) + (size_t)n * a.raw_chunks * plane_bytes + tid * 16;
  const uint8_t* cat_p = reinterpret_cast<const uint8_t*>(f.cat) + ((size_t)n * f.cat_chunks + C / 8) * plane_bytes + tid * 16;
  uint8_t* out_p = reinterpret_cast<uint8_t*>(f.out) + (size_t)n * f.out_chunks * plane_bytes + (warp * 32 + (lane >> 2)) * 16 + (lane & 3) * 4;
  float* Tn = a.T + (size_t)n * a.V + tid;
  const float wcst = a.wcst[n];
  const bool has_t = a.T != nullptr;               // window plans drop head 0 (ec3): no accumulator traffic at all
  const bool t_init = a.t_init || !has_t;
  // the block's tile sequence: whole groups g = blockIdx.x, + gridDim.x, ...; `cur` is consumed, `pf` runs STAGES tiles ahead
  struct TileIt { long long g; int ts; };
  auto tile_of = [&](const TileIt& it) { return it.g * kCatTileGroup + it.ts; };
  auto valid = [&](const TileIt& it) { return it.g < ngroups; };
  auto advance = [&](TileIt& it) {
    if (++it.ts == kCatTileGroup || tile_of(it) >= ntiles) { it.ts = 0; it.g += gridDim.x; }
  };
  auto issue_loads = [&](const TileIt& it, int stage) {   // this thread's voxel of tile `it` -> ring stage (one commit group, maybe empty)
    if (valid(it)) {
      const long long t256 = tile_of(it) * 256;
      if (t256 + tid < a.V) {
        const uint32_t dst = ring_u + (uint32_t)(stage * STAGE_BYTES);
        const uint8_t* rp = raw_p + t256 * 16;
#pragma unroll
        for (int k = 0; k < C / 8; ++k) { cp_async16(dst + k * 512, rp); rp += plane_bytes; }
        const uint8_t* cp = cat_p + t256 * 16;
#pragma unroll
        for (int k = C / 8; k < NCH; ++k) {
          if (k < f.cat_real_chunks) cp_async16(dst + k * 512, cp);
          cp += plane_bytes;
        }
        if (!t_init) cp_async4(dst + NCH * 512 - lane * 12, Tn + t256);   // (lane * 16 is in ring_u: head slot = lane * 4)
      }
    }
    cp_async_commit();
  };
  TileIt cur{(long long)blockIdx.x, 0}, pf = cur;
#pragma unroll
  for (int s = 0; s < STAGES; ++s) { issue_loads(pf, s); advance(pf); }
  int stage = 0;
  while (valid(cur)) {
    const long long t256 = tile_of(cur) * 256;
    const int lim = (int)min((long long)256, a.V - t256);   // voxels of this tile (only the last tile of a small volume is ragged)
    const bool live = tid < lim;
    cp_async_wait<STAGES - 1>();
    const uint32_t src = ring_u + (uint32_t)(stage * STAGE_BYTES);
    if (live) {
      // ---- 1. staged values of this thread's voxel
      Chunk8 in[C / 8];
#pragma unroll
      for (int k = 0; k < C / 8; ++k) in[k] = lds_chunk(src + k * 512);
#pragma unroll
      for (int k = C / 8; k < NCH; ++k)
        if (k < f.cat_real_chunks) st_chunk(rowp + k * 16, lds_chunk(src + k * 512));
      float t_old = 0.f;
      if (!t_init) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t_old) : "r"(src + NCH * 512 - lane * 12));
      // ---- 2. the block's own channels
      float e[C];
      float g1 = 0.f;
#pragma unroll
      for (int k = 0; k < C / 8; ++k) {
        float fv[8], mean[8], rstd[8], wse[8];
        chunk_to_floats(in[k], fv);
        ld8(s_mean, k, mean); ld8(s_rstd, k, rstd); ld8(s_wse, k, wse);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float t = lrelu_(fmaf(fv[i], rstd[i], mean[i]));
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
      float t = wcst;
#pragma unroll
      for (int k = 0; k < C / 8; ++k) {
        float weff[8];
        ld8(s_weff, k, weff);
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[k * 8 + i] *= g1; t = fmaf(weff[i], e[k * 8 + i], t); }
        st_chunk(rowp + k * 16, floats_to_chunk(e + k * 8));   // same rounding as the stored concat slice
      }
      if (has_t) Tn[t256] = t_old + t;
    } else {
#pragma unroll
      for (int k = 0; k < NCH; ++k) st_chunk(rowp + k * 16, Chunk8{{0u, 0u, 0u, 0u}});   // zero rows: zero outputs, nothing added to the statistics
    }
    // ---- 3. refill the stage (its values have been consumed above) with the tile STAGES ahead
    issue_loads(pf, stage);
    advance(pf);
    if (++stage == STAGES) stage = 0;
    __syncwarp();
    // ---- 4. this warp's 32 voxels x KCAT  times  KCAT x NOUT, fp32 accumulate
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
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
    // ---- 5. raw output (chunk planes, storage type) + statistics from the fp32 accumulators
    const int r0 = warp * 32 + (lane >> 2);      // this thread's first row of the tile; the others are + 8, + 16, + 24
    uint8_t* op = out_p + t256 * 16;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float x0 = acc[mt][nt][half * 2], x1 = acc[mt][nt][half * 2 + 1];
          ssum[nt][0] += x0; ssum[nt][1] += x1; ssq[nt][0] = fmaf(x0, x0, ssq[nt][0]); ssq[nt][1] = fmaf(x1, x1, ssq[nt][1]);
          if (r0 + mt * 16 + half * 8 < lim) *reinterpret_cast<uint32_t*>(op + (mt * 16 + half * 8) * 16) = pack_act2(x0, x1);
        }
      op += plane_bytes;
    }
    advance(cur);
    if (cur.ts == 0) {
      // group done: fold the fp32 partials (16 voxels per thread) into the fp64 running sums
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dsum[nt][j] += (double)ssum[nt][j]; dsq[nt][j] += (double)ssq[nt][j];
          ssum[nt][j] = 0.f; ssq[nt][j] = 0.f;
        }
    }
  }
  cp_async_wait<0>();
  // ---- statistics: reduce over the 8 row lanes of a warp, then over the 8 warps, one fp64 atomic per (channel, moment) and block
  __syncthreads();                                       // the operand tile is free: reuse it for the cross-warp reduction
  double* s_red = reinterpret_cast<double*>(sA);         // [8 warps][2 * NOUT]
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double s = dsum[nt][j], q = dsq[nt][j];
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
    for (int w = 0; w < 8; ++w) tot += s_red[w * 2 * NOUT + tid];
    const int ch = tid % NOUT, mom = tid / NOUT;
    atomicAdd(f.out_stats + ((size_t)n * f.out_stats_c + ch) * 2 + mom, tot);
  }
}

template <int C, int GATES, int KCAT, int NOUT>
static int launch_t(int N, const SseArgs& a, const CatFuseArgs& f, int num_sms, cudaStream_t st) {
  constexpr int ROW = KCAT * 2 + 16;
  const int smem = 256 * ROW + NOUT * ROW + 8 * cat_stages(KCAT) * ((KCAT / 8) * 512 + 128);   // operand tile, weights, cp.async rings
  static bool attr_set[64] = {};
  static int blocks_per_sm[64] = {};
  int dev = 0;
  SEUNET_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { seunet_set_error("apply_sse_cat: device index %d", dev); return 1; }
  if (!attr_set[dev]) {
    if (smem > 48 * 1024)
      SEUNET_CUDA_CHECK(cudaFuncSetAttribute(apply_sse_cat_kernel<C, GATES, KCAT, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int b = 0;
    SEUNET_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, apply_sse_cat_kernel<C, GATES, KCAT, NOUT>, 256, smem));
    blocks_per_sm[dev] = std::max(1, b);
    attr_set[dev] = true;
  }
  const long long ngroups = ((a.V + 255) / 256 + kCatTileGroup - 1) / kCatTileGroup;
  // ONE wave: the blocks of all samples together never exceed the resident slots (a 2-block second wave would double the time)
  const long long slots = (long long)num_sms * blocks_per_sm[dev];
  const int per_sample = (int)std::min<long long>(ngroups, std::max<long long>(1, slots / N));
  dim3 grid((unsigned)per_sample, N);
  apply_sse_cat_kernel<C, GATES, KCAT, NOUT><<<grid, 256, smem, st>>>(a, f);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_apply_sse_cat(int C, int N, const SseArgs& a, const CatFuseArgs& f, int num_sms, cudaStream_t st) {
  if (a.dest != nullptr) { seunet_set_error("apply_sse_cat: the fused pass does not store the block output"); return 1; }
  const int gates = a.wse2 ? 2 : 1;
  // (the C = 64 blocks - ec6, ec9, ec12, dc2 - were measured too: 254 registers, one block per SM, no faster than the two passes)
  if (C == 32 && gates == 1 && f.kcat == 64 && f.nout == 32) return launch_t<32, 1, 64, 32>(N, a, f, num_sms, st);     // ec3 -> ec33
  if (C == 32 && gates == 2 && f.kcat == 96 && f.nout == 32) return launch_t<32, 2, 96, 32>(N, a, f, num_sms, st);     // dc4 -> dc42
  seunet_set_error("apply_sse_cat: no instance for C=%d gates=%d K=%d N=%d", C, gates, f.kcat, f.nout);
  return 1;
}
