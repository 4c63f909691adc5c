// tcgen05 weight-gradient kernel: dW[co][ci][tap] = sum_{n,voxel} X[n,ci,voxel+shift(tap)] * DY[n,co,voxel]
// (autograd of the nn.Conv3d call sites SE_UNet.py:15/42/57, reached from loss.backward(), train.py:246/439/602).
//
// GEMM view per tap: D[ci, co] += A[ci, k] * B[co, k] with k = voxels.  Both operands are "MN-major" in UMMA terms and
// the chunk-plane activation layout [C/8][h][w][8] IS the no-swizzle MN-major canonical layout
// ((8,1,m),(8,k)):((1,8,SBO),(8,LBO)): 8 channels contiguous (16 B), 8 consecutive w voxels 16 B apart form a core
// matrix, the next 8 voxels (next h line) are LBO away, the next 8 channels (next chunk plane) SBO away.  So, exactly
// as in the forward kernel, one TMA box per input plane feeds all kw taps by start-address offsets.
//   * UMMA M = 128 input channels (16 chunk planes), or M = 64 when Cin <= 64 (8 planes: half the shared-memory
//     traffic; its accumulator rows live in TMEM lanes (r/16)*32 + r%16, measured); planes beyond Cin read junk whose
//     rows are never used,
//     N = Cout (16/32/64), K = 16 voxels (two h lines of the 16x8 tile) -> 8 MMAs per (tile plane, tap).
//   * The three kw taps are stacked along N (three w-shifted dY boxes), and as many kh taps as fit are stacked along M
//     (round 2): the A tile holds nkh h-shifted copies of the X box back to back - the chunk-plane stride stays uniform
//     across the copies - so ONE M = 64/128 instruction covers nkh * 3 taps: nkh = 3 for Cin <= 40, 2 for Cin <= 64,
//     1 above.  The kernel is bound by the NUMBER of MMAs (K is only 16 voxels; an M=128 instruction costs >= ~48 cycles
//     whatever M and N are), so the full-resolution layers (Cin <= 32) need 3x fewer of them.
//   * A pass fixes kd and a group of nkh kh taps - 3, 6 or 9 passes for 3x3x3 - and keeps its accumulators
//     [nkh*Cin x 3*Cout] in TMEM for the whole kernel: the accumulation over ALL voxels of the CTA's tile planes happens
//     inside TMEM (fp32).  1x1x1 convs use one pass per block of 128 input channels.
//   * grid = npass * ctas_per_pass persistent CTAs; each writes its partial [taps][128][Cout] once; a small
//     reduction kernel sums the partials in a fixed order (deterministic) into the flat fp32 gradient.
//   * both operands use the activation storage type (tcgen05 kind::f16 faults on mixed f16 x bf16 operands, measured);
//     dY is therefore stored in that type too, pre-scaled by a per-layer power of two (backward.cu) that the
//     reduction kernel divides out again.
#include "wgrad_tc.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

constexpr int kWgThreads = 192;

template <int COUT>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                const __grid_constant__ WgradKArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t x_addr = smem_base;
  const uint32_t dy_addr = smem_base + a.nstages * a.x_stage_bytes;
  const uint32_t bar_addr = smem_base + a.bar_off;
  auto full_bar = [&](int i) { return bar_addr + 8u * i; };
  auto empty_bar = [&](int i) { return bar_addr + 8u * (8 + i); };
  const uint32_t done_bar = bar_addr + 8u * 16;
  const uint32_t tmem_slot_addr = bar_addr + 8u * 17;
  const uint32_t count_addr = bar_addr + 8u * 18;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (see conv_tc.cu)
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstages; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
  }
  if (warp == 1) { tmem_alloc(tmem_slot_addr, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot_addr));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const int pass = blockIdx.x / a.ctas_per_pass;
  const int rank = blockIdx.x % a.ctas_per_pass;
  // 3x3x3: pass = kd * npkh + (group of nkh kh taps); 1x1x1: pass = block of 128 input channels
  const int kd = a.ksize == 3 ? pass / a.npkh : 1;
  const int kh0 = a.ksize == 3 ? (pass % a.npkh) * a.nkh : 1;
  const int nkh = a.ksize == 3 ? min(a.nkh, 3 - kh0) : 1;   // (the last group of a 2 + 1 split holds one tap: its second M half is junk)
  const int ntap = a.ksize == 3 ? 3 : 1;
  const int x_chunk = a.x_chunk_off + (a.ksize == 3 ? 0 : pass * 16);
  const int dshift = (kd - 1) * a.dil;

  // Each CTA of a pass walks a CONTIGUOUS range of tile planes; the (w, h, plane, sample) coordinates are advanced with
  // carries - the single-thread producer and issuer loops are issue-bound, and four runtime integer divisions per tile
  // cost more than the eight MMAs they feed.
  const int tp_begin = (int)((long long)rank * a.numTilePlanes / a.ctas_per_pass);
  const int tp_end = (int)((long long)(rank + 1) * a.numTilePlanes / a.ctas_per_pass);
  struct Walk { int tw, th, p, n; };
  auto walk_init = [&](int tp) {
    Walk w;
    w.tw = tp % a.tilesW; tp /= a.tilesW;
    w.th = tp % a.tilesH; tp /= a.tilesH;
    w.p = tp % a.D; w.n = tp / a.D;
    return w;
  };
  auto walk_next = [&](Walk& w) {
    if (++w.tw == a.tilesW) { w.tw = 0; if (++w.th == a.tilesH) { w.th = 0; if (++w.p == a.D) { w.p = 0; ++w.n; } } }
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      Walk w = walk_init(tp_begin);
      const uint32_t tx_bytes = nkh * a.x_box_bytes + ntap * a.dy_box_bytes;
      const int kw0 = ntap == 3 ? 1 : 0;
      for (int tp = tp_begin; tp < tp_end; ++tp, walk_next(w)) {
        // tile plane `p` indexes the INPUT plane; it pairs with output-gradient plane p - dshift
        const int po = w.p - dshift;
        if (po < 0 || po >= a.D) continue;
        const int w0 = w.tw * a.tw, h0 = w.th * 16;
        mbar_wait(empty_bar(st), ph ^ 1u);
        mbar_expect_tx(full_bar(st), tx_bytes);
        // dW[tap] = sum_v X[v + shift(tap)] dY[v]: the kh shift is applied to the X copies (stacked along M), the kw shift
        // to the dY copies (stacked along N), the kd shift to the plane pairing
        for (int k = 0; k < nkh; ++k)
          tma_load_4d(x_addr + st * a.x_stage_bytes + k * a.x_box_bytes, &tmap_x, full_bar(st), 8 * w0,
                      h0 + (a.ksize == 3 ? (kh0 + k - 1) * a.dil : 0), w.p, w.n * a.x_chunks_total + x_chunk);
        for (int kw = 0; kw < ntap; ++kw)
          tma_load_4d(dy_addr + st * a.dy_stage_bytes + kw * a.dy_box_bytes, &tmap_dy, full_bar(st),
                      8 * (w0 - (kw - kw0) * a.dil), h0, po, w.n * a.dy_chunks_total + a.dy_chunk_off);
        if (++st == (uint32_t)a.nstages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // One elected thread runs the whole issue loop (as in conv_tc.cu: re-converging the warp after every elected
    // tcgen05.mma drains the tensor pipe's instruction queue; ptxas keeps the elect region on the uniform datapath).
    if (elect_one_sync()) {
    uint32_t st = 0, ph = 0, any = 0;
    const uint32_t idesc = umma_idesc2(a.fmt_a, a.fmt_b, 1u, 1u, a.m64 ? 64 : 128, ntap * COUT);
    // A (X tile, 16x8 voxels, no halo) and B (ntap shifted dY tiles stacked along N): MN-major, m/n-groups = chunk planes
    // 2048 B apart (SBO), k-groups = h lines 128 B apart (LBO).  Only the low descriptor word (address) changes.
    const uint32_t desc_hi = ((a.x_plane_bytes >> 4) & 0x3FFFu) | (1u << 14);   // SBO = chunk-plane stride (16 x tw voxels)
    const int ksteps = a.tw;   // K=16 steps per tile plane
    const uint32_t lbo_bits = ((128u >> 4) & 0x3FFFu) << 16;
    const uint32_t a_lo_first = lbo_bits | ((x_addr & 0x3FFFFu) >> 4), b_lo_first = lbo_bits | ((dy_addr & 0x3FFFFu) >> 4);
    const uint32_t a_stage16 = a.x_stage_bytes >> 4, b_stage16 = a.dy_stage_bytes >> 4;
    uint32_t a_lo_st = a_lo_first, b_lo_st = b_lo_first, fbar = full_bar(0);
    int p = walk_init(tp_begin).p, tw = walk_init(tp_begin).tw, th = walk_init(tp_begin).th;
    for (int tp = tp_begin; tp < tp_end; ++tp) {
      const int po = p - dshift;
      if (++tw == a.tilesW) { tw = 0; if (++th == a.tilesH) { th = 0; if (++p == a.D) p = 0; } }
      if (po < 0 || po >= a.D) continue;
      mbar_wait(fbar, ph);      // TMA -> mbarrier -> MMA: ordered by the mbarrier itself
      if (!any) {   // first tile plane of this CTA: the first MMA overwrites the accumulators
        {
          const uint64_t ad = ((uint64_t)desc_hi << 32) | a_lo_st, bd = ((uint64_t)desc_hi << 32) | b_lo_st;
          umma_f16(tmem_base, ad, bd, idesc, 0u);
        }
        any = 1;
#pragma unroll 1
        for (int j = 1; j < ksteps; ++j)
          umma_f16_lohi(tmem_base, a_lo_st + 16u * j, desc_hi, b_lo_st + 16u * j, desc_hi, idesc);
      } else {
#pragma unroll 1
        for (int j8 = 0; j8 < ksteps; j8 += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j)   // two h-line segments of 8 voxels = 256 B per K=16 step
            umma_f16_lohi(tmem_base, a_lo_st + 16u * (j8 + j), desc_hi, b_lo_st + 16u * (j8 + j), desc_hi, idesc);
        }
      }
      umma_commit(fbar + 64u);   // empty barrier of this stage
      a_lo_st += a_stage16; b_lo_st += b_stage16; fbar += 8u;
      if (++st == (uint32_t)a.nstages) { st = 0; ph ^= 1u; a_lo_st = a_lo_first; b_lo_st = b_lo_first; fbar = full_bar(0); }
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(count_addr), "r"(any) : "memory");
    umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    // epilogue: after ALL MMAs of this CTA, dump [ntap][128][COUT] fp32 to the partial buffer
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    uint32_t any;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(any) : "r"(count_addr));
    float* dst = a.partial + ((size_t)blockIdx.x * ntap * 128 + row) * COUT;
    for (int kw = 0; kw < ntap; ++kw) {
#pragma unroll 1
      for (int cg = 0; cg < COUT / 16; ++cg) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + kw * COUT + cg * 16, v);
        tmem_ld_wait();
        float4* o = reinterpret_cast<float4*>(dst + (size_t)kw * 128 * COUT + cg * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 f;
          f.x = any ? __uint_as_float(v[4 * i]) : 0.f; f.y = any ? __uint_as_float(v[4 * i + 1]) : 0.f;
          f.z = any ? __uint_as_float(v[4 * i + 2]) : 0.f; f.w = any ? __uint_as_float(v[4 * i + 3]) : 0.f;
          o[i] = f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}


// ---------------------------------------------------------------------------------------------
// Version 3 (round 2), 3x3x3 layers with Cout <= 32: the three kd taps share ONE walk over the volume.
// ncu on the pass-per-(kd, kh-group) kernel above (profiles/r02_wgrad_ncu_b8.txt): dc5 moves 45 GB from L2 to shared memory
// per 8-patch step (13 TB/s) and 12.9 GB from DRAM for 3.2 GB of operands - every pass re-reads X and dY.  Here a CTA walks
// whole (h, w) tile COLUMNS along d and keeps the previous plane's X and dY stage alive, so plane p is loaded once and used
// three times:   kd = 1: X_p x dY_p      kd = 2: X_p x dY_(p-dil)      kd = 0: X_(p-dil) x dY_p
// into three accumulator sets [nkh*Cin x 3*Cout] (3 * 96 <= 512 TMEM columns - why Cout = 64 stays on the kernel above).
// Dilation-2 layers walk the two plane parity classes as separate columns (inside a class the kd taps are neighbours).
// A stage is released after the NEXT step's MMAs (or at the end of its column).  L2->SMEM traffic drops 3x; the MMA count is
// unchanged, so the full-resolution layers end up MMA-count-bound (K = 16 voxels per instruction).
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                 const __grid_constant__ WgradKArgs a) {
  constexpr int NCOLS = 3 * COUT;            // accumulator columns per kd set (three kw taps)
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t x_addr = smem_base;
  const uint32_t dy_addr = smem_base + a.nstages * a.x_stage_bytes;
  const uint32_t bar_addr = smem_base + a.bar_off;
  auto full_bar = [&](int i) { return bar_addr + 8u * i; };
  auto empty_bar = [&](int i) { return bar_addr + 8u * (8 + i); };
  const uint32_t done_bar = bar_addr + 8u * 16;
  const uint32_t tmem_slot_addr = bar_addr + 8u * 17;
  const uint32_t count_addr = bar_addr + 8u * 18;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstages; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
  }
  if (warp == 1) { tmem_alloc(tmem_slot_addr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot_addr));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const int pass = blockIdx.x / a.ctas_per_pass;       // group of nkh kh taps
  const int rank = blockIdx.x % a.ctas_per_pass;
  const int kh0 = pass * a.nkh;
  const int nkh = min(a.nkh, 3 - kh0);
  const int dstep = a.dil;                             // plane distance of the kd taps = stride of a column's plane walk
  // Work units = column segments: a column is cut into a.nseg runs of planes when there are too few columns to fill the
  // CTAs (small batches).  A segment that does not start at the first plane of its class first loads the plane before it as
  // a warm-up stage (no MMAs of its own, it only serves the two cross terms of the segment's first plane).
  const int numUnits = a.numCols * a.nseg;
  const int c_begin = (int)((long long)rank * numUnits / a.ctas_per_pass);
  const int c_end = (int)((long long)(rank + 1) * numUnits / a.ctas_per_pass);
  struct Col { int n, h0, w0, par, i0, i1, warm; };   // planes par + dstep * i, i in [i0, i1); warm: also load i0 - 1 first
  auto col_decode = [&](int u) {
    Col q;
    const int seg = u % a.nseg;
    int c = u / a.nseg;
    q.par = c % dstep; c /= dstep;
    q.w0 = (c % a.tilesW) * a.tw; c /= a.tilesW;
    q.h0 = (c % a.tilesH) * 16; q.n = c / a.tilesH;
    const int P = (a.D - q.par + dstep - 1) / dstep;
    q.i0 = (int)((long long)seg * P / a.nseg); q.i1 = (int)((long long)(seg + 1) * P / a.nseg);
    q.warm = (q.i0 > 0 && q.i1 > q.i0) ? 1 : 0;
    return q;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      const uint32_t tx_bytes = nkh * a.x_box_bytes + 3 * a.dy_box_bytes;
      for (int c = c_begin; c < c_end; ++c) {
        const Col q = col_decode(c);
        for (int i = q.i0 - q.warm; i < q.i1; ++i) {
          const int p = q.par + dstep * i;
          mbar_wait(empty_bar(st), ph ^ 1u);
          mbar_expect_tx(full_bar(st), tx_bytes);
          for (int k = 0; k < nkh; ++k)
            tma_load_4d(x_addr + st * a.x_stage_bytes + k * a.x_box_bytes, &tmap_x, full_bar(st), 8 * q.w0,
                        q.h0 + (kh0 + k - 1) * a.dil, p, q.n * a.x_chunks_total + a.x_chunk_off);
          for (int kw = 0; kw < 3; ++kw)
            tma_load_4d(dy_addr + st * a.dy_stage_bytes + kw * a.dy_box_bytes, &tmap_dy, full_bar(st),
                        8 * (q.w0 - (kw - 1) * a.dil), q.h0, p, q.n * a.dy_chunks_total + a.dy_chunk_off);
          if (++st == (uint32_t)a.nstages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      uint32_t st = 0, ph = 0, any = 0;     // any: bit kd set once accumulator set kd has been written
      const uint32_t idesc = umma_idesc2(a.fmt_a, a.fmt_b, 1u, 1u, a.m64 ? 64 : 128, NCOLS);
      const uint32_t desc_hi = ((a.x_plane_bytes >> 4) & 0x3FFFu) | (1u << 14);   // SBO = chunk-plane stride
      const int ksteps = a.tw;
      const uint32_t lbo_bits = ((128u >> 4) & 0x3FFFu) << 16;
      const uint32_t a_lo0 = lbo_bits | ((x_addr & 0x3FFFFu) >> 4), b_lo0 = lbo_bits | ((dy_addr & 0x3FFFFu) >> 4);
      const uint32_t a_stage16 = a.x_stage_bytes >> 4, b_stage16 = a.dy_stage_bytes >> 4;
      // one kd group: ksteps MMAs of A (X stage sa) x B (dY stage sb) into accumulator set kd
      auto group = [&](uint32_t sa, uint32_t sb, int kd) {
        const uint32_t a_lo = a_lo0 + sa * a_stage16, b_lo = b_lo0 + sb * b_stage16;
        const uint32_t d = tmem_base + (uint32_t)(kd * NCOLS);
        {
          const uint64_t ad = ((uint64_t)desc_hi << 32) | a_lo, bd = ((uint64_t)desc_hi << 32) | b_lo;
          umma_f16(d, ad, bd, idesc, (any >> kd) & 1u);   // the very first MMA of a set overwrites it
        }
        any |= 1u << kd;
#pragma unroll 1
        for (int j = 1; j < ksteps; ++j) umma_f16_lohi(d, a_lo + 16u * j, desc_hi, b_lo + 16u * j, desc_hi, idesc);
      };
      for (int c = c_begin; c < c_end; ++c) {
        const Col q = col_decode(c);
        uint32_t prev = 0;
        for (int i = q.i0 - q.warm; i < q.i1; ++i) {
          mbar_wait(full_bar(st), ph);
          if (i >= q.i0) {             // (the warm-up plane belongs to the previous segment: it only becomes `prev`)
            group(st, st, 1);
            if (i > 0) {
              group(st, prev, 2);      // X_p x dY_(p - dil)
              group(prev, st, 0);      // X_(p - dil) x dY_p
              umma_commit(empty_bar(prev));
            }
            if (i == q.i1 - 1) umma_commit(empty_bar(st));
          }
          prev = st;
          if (++st == (uint32_t)a.nstages) { st = 0; ph ^= 1u; }
        }
      }
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(count_addr), "r"(any) : "memory");
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    // epilogue: after ALL MMAs of this CTA, dump [kd][kw][128][COUT] fp32 to the partial buffer
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    uint32_t any;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(any) : "r"(count_addr));
    float* dst = a.partial + ((size_t)blockIdx.x * 9 * 128 + row) * COUT;
#pragma unroll 1
    for (int t = 0; t < 9; ++t) {          // t = kd * 3 + kw
      const bool live = (any >> (t / 3)) & 1u;
#pragma unroll 1
      for (int cg = 0; cg < COUT / 16; ++cg) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + t * COUT + cg * 16, v);
        tmem_ld_wait();
        float4* o = reinterpret_cast<float4*>(dst + (size_t)t * 128 * COUT + cg * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 f;
          f.x = live ? __uint_as_float(v[4 * i]) : 0.f; f.y = live ? __uint_as_float(v[4 * i + 1]) : 0.f;
          f.z = live ? __uint_as_float(v[4 * i + 2]) : 0.f; f.w = live ? __uint_as_float(v[4 * i + 3]) : 0.f;
          o[i] = f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// v3 partial layout: [pass (kh group)][cta][kd][kw][128 rows = (kh in group, ci)][COUT]
__global__ void wgrad3_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Cin, int Cout, int COUT,
                                     int ctas_per_pass, const float* __restrict__ inv_scale, int m64, int nkh, int x_rows) {
  const float mul = inv_scale ? inv_scale[0] : 1.f;
  const int total = Cout * Cin * 27;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 27, ci = (i / 27) % Cin, co = i / (27 * Cin);
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    const int pass = kh / nkh;
    int row = (kh % nkh) * x_rows + ci;
    if (m64 == 1) row = (row >> 4) * 32 + (row & 15);
    float s = 0.f;
    for (int r = 0; r < ctas_per_pass; ++r)
      s += partial[(((size_t)(pass * ctas_per_pass + r) * 9 + kd * 3 + kw) * 128 + row) * COUT + co];
    dw[i] = s * mul;
  }
}

// dW[co][ci][kd][kh][kw] = sum_r partial[pass(kd, kh group) or ci-block][r][kw][(kh in group,) ci][co]   (fixed summation order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Cin, int Cout, int COUT,
                                    int ksize, int ctas_per_pass, const float* __restrict__ inv_scale, int m64, int nkh, int npkh,
                                    int x_rows) {
  const float mul = inv_scale ? inv_scale[0] : 1.f;
  const int K3 = ksize * ksize * ksize;
  const int total = Cout * Cin * K3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % K3, ci = (i / K3) % Cin, co = i / (K3 * Cin);
    int pass, kw, row;
    if (ksize == 3) {
      const int kd = tap / 9, kh = (tap / 3) % 3;
      pass = kd * npkh + kh / nkh; kw = tap % 3; row = (kh % nkh) * x_rows + ci;
    }
    else { pass = ci / 128; kw = 0; row = ci % 128; }
    // UMMA M=64 accumulators use 16 TMEM lanes of each of the four 32-lane sub-partitions
    if (m64 == 1) row = (row >> 4) * 32 + (row & 15);
    const int ntap = ksize == 3 ? 3 : 1;
    float s = 0.f;
    for (int r = 0; r < ctas_per_pass; ++r)
      s += partial[(((size_t)(pass * ctas_per_pass + r) * ntap + kw) * 128 + row) * COUT + co];
    dw[i] = s * mul;
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled wg_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (PFN_encodeTiled)p;
  return fn;
}

size_t wgrad_partial_bytes(int Cin, int Cout, int ksize, int num_sms) {
  WgradLaunch L;
  if (wgrad_launch_init(&L, 1, 16, 16, 8, Cin, Cout, ksize, 1, nullptr, (Cin + 7) / 8, 0, 0, nullptr, 8, 0, nullptr, num_sms, true))
    return 0;
  return (size_t)L.grid * (L.a.v3 ? 9 : (ksize == 3 ? 3 : 1)) * 128 * L.COUT * sizeof(float);
}

int wgrad_launch_init(WgradLaunch* L, int N, int D, int H, int W, int Cin, int Cout, int ksize, int dil,
                      const void* x, int x_chunks_total, int x_chunk_off, int x_bf16,
                      const void* dy, int dy_chunks_total, int dy_chunk_off, float* partial, int num_sms, bool geometry_only) {
  memset(L, 0, sizeof(*L));
  WgradKArgs& a = L->a;
  L->Cin = Cin; L->Cout = Cout; L->ksize = ksize;
  L->COUT = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);
  if (Cout > 64 || (ksize != 1 && ksize != 3)) { seunet_set_error("wgrad: unsupported shape"); return 1; }
  if (ksize == 3 && Cin > 128) { seunet_set_error("wgrad: 3x3x3 with Cin > 128 unsupported"); return 1; }
  const int halo = ksize == 3 ? dil : 0;
  // kh taps stacked along M: as many h-shifted copies of the X box as fit 128 accumulator rows
  const int xpl_all = std::min(16, (Cin + 7) / 8);
  a.nkh = ksize == 3 ? (3 * xpl_all <= 16 ? 3 : (2 * xpl_all <= 16 ? 2 : 1)) : 1;
  if (getenv("SEUNET_WG_NKH")) a.nkh = std::max(1, std::min(a.nkh, atoi(getenv("SEUNET_WG_NKH"))));   // developer A/B knob
  // version 3 (kd taps merged into one walk, see wgrad3_tc_kernel): 3x3x3 layers whose nine accumulator blocks fit TMEM
  static const bool v3_enabled = !(getenv("SEUNET_WG_V3") && atoi(getenv("SEUNET_WG_V3")) == 0);   // developer A/B knob
  a.v3 = (ksize == 3 && L->COUT <= 32 && v3_enabled) ? 1 : 0;
  if (a.v3 && a.nkh == 2) a.nkh = 1;   // Cin = 64: one kh tap per pass keeps the stage at 40 KB (5 stages; 56 KB stages would leave one in flight)
  a.npkh = (3 + a.nkh - 1) / a.nkh;
  const int npass = ksize == 3 ? (a.v3 ? a.npkh : 3 * a.npkh) : (Cin + 127) / 128;
  a.N = N; a.D = D; a.H = H; a.W = W;
  // Tile width: the per-stage costs (4 TMA instructions, barrier round trip, loop bookkeeping of the single-thread
  // producer/issuer) dominate narrow layers, so a stage covers as many voxels as ~100 KB of shared memory allow
  // (two stages of 80 KB measured faster than five of 40 KB).
  {
    const int coutp = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);
    const int xpl = std::min(16, (Cin + 7) / 8);
    const uint32_t per8 = 2048u * xpl * a.nkh + (ksize == 3 ? 3u : 1u) * 2048u * (coutp / 8);   // stage bytes at tw = 8
    int tw = 8;
    static const uint32_t cap_kb = getenv("SEUNET_WG_STAGE_KB") ? (uint32_t)atoi(getenv("SEUNET_WG_STAGE_KB")) : 100u;
    const uint32_t cap = a.v3 ? 40u * 1024u : cap_kb * 1024u;   // v3 keeps one extra stage alive: smaller stages, deeper ring
    while (tw < 32 && per8 * (tw * 2 / 8) <= cap && W >= tw * 2) tw *= 2;
    a.tw = tw;
  }
  a.tilesW = (W + a.tw - 1) / a.tw; a.tilesH = (H + 15) / 16;
  a.numTilePlanes = N * D * a.tilesH * a.tilesW;
  a.numCols = N * a.tilesH * a.tilesW * (halo == 2 ? 2 : 1);   // v3: (sample, h tile, w tile, plane parity class)
  a.nseg = 1;                                                    // (set below once the pass count is known)
  a.dil = halo; a.ksize = ksize;
  a.lineW = 8 + 2 * halo;
  a.ctas_per_pass = std::max(1, num_sms / npass);
  L->grid = npass * a.ctas_per_pass;
  if (a.v3) {
    // >= 2 work units per CTA, segments of >= 8 planes (every extra segment costs one extra stage load)
    // pick the segment count with the best (load balance) x (extra warm-up loads) product
    const int P = std::max(1, D / std::max(1, halo));
    double best = 1e30;
    for (int ns = 1; ns <= std::max(1, std::min(16, P / 8)); ++ns) {
      const double units = (double)a.numCols * ns;
      const double per_cta = std::ceil(units / a.ctas_per_pass) / (units / a.ctas_per_pass);   // slowest CTA / average
      const double cost = per_cta * (1.0 + (double)(ns - 1) / P);
      if (cost < best - 1e-9) { best = cost; a.nseg = ns; }
    }
  }
  // planes per X box: the real channel planes of this pass (<= 16)
  const int cin_planes_total = (Cin + 7) / 8;
  const int xplanes = std::min(16, cin_planes_total);
  a.x_chunks_total = x_chunks_total; a.x_chunk_off = x_chunk_off;
  a.dy_chunks_total = dy_chunks_total; a.dy_chunk_off = dy_chunk_off;
  a.x_plane_bytes = 256u * a.tw;   // 16 x tw voxels x 16 B, no halo: the tap shifts are applied to dY
  a.x_box_bytes = a.x_plane_bytes * xplanes;
  a.x_stage_bytes = a.nkh * a.x_box_bytes;   // nkh copies back to back (a multiple of 2 KB)
  a.x_rows = xplanes * 8;
  a.dy_box_bytes = a.x_plane_bytes * (L->COUT / 8);
  a.dy_stage_bytes = (ksize == 3 ? 3u : 1u) * a.dy_box_bytes;
  int nst = (int)(((a.v3 ? 216u : 200u) * 1024u) / (a.x_stage_bytes + a.dy_stage_bytes));
  nst = std::min(nst, a.v3 ? 8 : 6);
  if (nst < (a.v3 ? 3 : 2)) { seunet_set_error("wgrad: shared memory budget exceeded"); return 1; }
  a.nstages = nst;
  uint32_t natural = nst * (a.x_stage_bytes + a.dy_stage_bytes);
  // junk rows: an M=128 A operand spans 16 chunk planes from the start of the LAST X stage
  const uint32_t junk_end = (nst - 1) * a.x_stage_bytes + 16u * a.x_plane_bytes;
  natural = std::max(natural, junk_end);
  a.bar_off = (natural + 127u) & ~127u;
  L->smem_bytes = a.bar_off + 256u + 128u;
  if (L->smem_bytes > 224u * 1024u) { seunet_set_error("wgrad: shared memory budget exceeded (%u)", L->smem_bytes); return 1; }
  a.partial = partial;
  a.fmt_a = x_bf16 ? 1u : (uint32_t)SEUNET_UMMA_FMT;
  a.m64 = a.nkh * xplanes <= 8 ? 1 : 0;   // UMMA M=64 when 8 chunk planes cover all (stacked) input rows
  a.fmt_b = a.fmt_a;  // tcgen05 kind::f16 requires A and B in the SAME format (mixed f16 x bf16 is an illegal instruction)
  if (geometry_only) return 0;
  if (ksize == 1 && x_chunk_off + npass * 16 > x_chunks_total + 15) { seunet_set_error("wgrad: bad 1x1 channel blocks"); return 1; }
  PFN_encodeTiled enc = wg_encode_fn();
  if (!enc) { seunet_set_error("cuTensorMapEncodeTiled not available"); return 1; }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  {
    cuuint64_t gdim[4] = {(cuuint64_t)8 * W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N * x_chunks_total};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
    cuuint32_t box[4] = {(cuuint32_t)(8 * a.tw), 16u, 1u, (cuuint32_t)xplanes};
    CUresult r = enc(&L->tmap_x, x_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                     const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seunet_set_error("wgrad: cuTensorMapEncodeTiled(x) failed: %d", (int)r); return 1; }
  }
  {
    cuuint64_t gdim[4] = {(cuuint64_t)8 * W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N * dy_chunks_total};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
    cuuint32_t box[4] = {(cuuint32_t)(8 * a.tw), 16u, 1u, (cuuint32_t)(L->COUT / 8)};
    CUresult r = enc(&L->tmap_dy, x_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(dy), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seunet_set_error("wgrad: cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return 1; }
  }
  return 0;
}

template <int COUT>
static int wgrad_launch_t(const WgradLaunch& L, cudaStream_t st) {
  if (L.a.v3) {
    SEUNET_CUDA_CHECK(cudaFuncSetAttribute(wgrad3_tc_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    wgrad3_tc_kernel<COUT><<<L.grid, kWgThreads, L.smem_bytes, st>>>(L.tmap_x, L.tmap_dy, L.a);
  } else {
    SEUNET_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    wgrad_tc_kernel<COUT><<<L.grid, kWgThreads, L.smem_bytes, st>>>(L.tmap_x, L.tmap_dy, L.a);
  }
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int wgrad_launch_run(const WgradLaunch& L, float* dw, const float* inv_scale, cudaStream_t st) {
  int rc;
  switch (L.COUT) {
    case 16: rc = wgrad_launch_t<16>(L, st); break;
    case 32: rc = wgrad_launch_t<32>(L, st); break;
    case 64: rc = wgrad_launch_t<64>(L, st); break;
    default: seunet_set_error("wgrad: bad COUT"); return 1;
  }
  if (rc) return rc;
  const int total = L.Cout * L.Cin * L.ksize * L.ksize * L.ksize;
  if (L.a.v3) {
    wgrad3_reduce_kernel<<<std::min((total + 255) / 256, 592), 256, 0, st>>>(L.a.partial, dw, L.Cin, L.Cout, L.COUT, L.a.ctas_per_pass,
                                                                              inv_scale, L.a.m64, L.a.nkh, L.a.x_rows);
    SEUNET_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  wgrad_reduce_kernel<<<std::min((total + 255) / 256, 592), 256, 0, st>>>(L.a.partial, dw, L.Cin, L.Cout, L.COUT, L.ksize,
                                                                           L.a.ctas_per_pass, inv_scale, L.a.m64, L.a.nkh, L.a.npkh,
                                                                           L.a.x_rows);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
