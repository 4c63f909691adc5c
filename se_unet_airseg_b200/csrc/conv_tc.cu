// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Replaces the reference's nn.Conv3d call sites SE_UNet.py:15/57 (3x3x3, dilation 1|2, padding=dilation)
// and SE_UNet.py:42 (CATConv 1x1x1), forward (and data-gradient, with mirrored/transposed weights).
//
// GEMM view:  out[voxel, cout] = sum_{tap, cin} in[voxel + shift(tap), cin] * W[tap, cin, cout].
//   * UMMA M = 128 voxels  = a 16(h) x 8(w) patch of ONE d-plane.
//   * UMMA K = 16 channels of one (kh,kw) tap.  Activations live in HBM as [n][C/8][D][H][W][8]
//     ("chunk planes"), and one TMA box brings a halo'd (16+2h)x(8+2h) patch of KC channels of one
//     input plane into shared memory as [KC/8][hh][ww][8].  In the no-swizzle K-major canonical UMMA
//     layout ((8,m),2):((16B,SBO),LBO) this means: the 8 rows of a core matrix are 8 consecutive w
//     voxels, SBO = one halo line, LBO = one chunk plane - and EVERY (kh,kw) tap is just a different
//     start address into the same halo tile.  Activations are therefore read from L2 once per
//     input plane (x1.4-1.9 halo) instead of 27 times.
//   * UMMA N = 3*Cout: the three kd taps that consume the same input plane q are stacked along N.
//     They accumulate into the output planes q-dil, q, q+dil, whose TMEM accumulators are laid out
//     adjacently (for dil=2 the even planes first, then the odd ones), so one instruction with
//     N = 3*Cout feeds three accumulators.  This triples the reuse of the A tile read from shared
//     memory, which is what limits small-N UMMA shapes.
//   * A CTA tile is DT = 512/Cout output planes of one 16x8 patch: the whole TMEM is one accumulator stage whose
//     plane slots are handed back and forth individually (per-slot mbarriers), so the epilogue of tile i still
//     overlaps the MMAs of tile i+1 (see the comment above the kernel).
//   * Weights are pre-packed into the exact shared-memory image ([chunk][step][khalf][3*Cout][8]) and
//     either stay resident for the whole (persistent) CTA or stream through a 2-slot ring when
//     27*Cin*Cout*2 B does not fit (Cin*Cout >= 64*64).
//   * Epilogue: TMEM -> registers -> fp16/bf16 raw conv output (chunk planes) + per-(n,c) sum / sum
//     of squares for InstanceNorm (fp32 partials from the fp32 accumulators, fp64 atomics).
#include "conv_tc.cuh"
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <algorithm>

// ---------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------
struct TileCoord { int n, d0, h0, w0, par, Dp; };   // d0 / Dp count planes of the tile's parity class (see ConvKArgs::dstep)

template <int DT>
__device__ __forceinline__ TileCoord decode_tile(const ConvKArgs& a, int tile) {
  TileCoord t;
  const int tw = tile % a.tilesW; tile /= a.tilesW;
  const int th = tile % a.tilesH; tile /= a.tilesH;
  int td = tile % a.tilesD; tile /= a.tilesD;
  // dstep == 2 (dilation 2): a tile holds DT planes of ONE parity, d = par + 2 * (d0 + i); the kd taps then reach the
  // neighbouring planes of the same parity class, so a tile needs DT + 2 input planes instead of DT + 4.
  t.par = 0;
  if (a.dstep == 2) { const int half = a.tilesD >> 1; t.par = td >= half ? 1 : 0; td -= t.par * half; }
  t.Dp = a.dstep == 2 ? ((a.D - t.par + 1) >> 1) : a.D;
  t.n = tile; t.d0 = td * DT; t.h0 = th * kConvTileH; t.w0 = tw * kConvTileW;
  return t;
}

// The 9 in-plane taps x J 16-channel blocks of one input plane, fully unrolled (J = 0: runtime block count).  Runs on
// the whole issuer warp with uniform operands; only the tcgen05.mma itself is issued by the elected lane.
template <int J>
__device__ __forceinline__ void issue_taps3(uint32_t dcol, uint32_t a_lo_st, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t kh_step, uint32_t kw_step, uint32_t j_step,
                                            uint32_t b_step, int jsteps) {
  // Only a few tcgen05.mma per basic block: every in-flight instruction pins its own uniform-register operands, and a
  // fully unrolled plane (36 of them) pushes the loop-carried state out of the uniform register file.
  uint32_t a_row = a_lo_st;
#pragma unroll 1
  for (int kh = 0; kh < 3; ++kh) {
    uint32_t a_tap = a_row;
#pragma unroll 1
    for (int kw = 0; kw < 3; ++kw) {
      uint32_t a_lo = a_tap;
      if (J > 0) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
          if (elect_one_sync()) umma_f16_lohi(dcol, a_lo, a_hi, b_lo, b_hi, idesc);
          a_lo += j_step; b_lo += b_step;
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < jsteps; ++j) {
          if (elect_one_sync()) umma_f16_lohi(dcol, a_lo, a_hi, b_lo, b_hi, idesc);
          a_lo += j_step; b_lo += b_step;
        }
      }
      a_tap += kw_step;
    }
    a_row += kh_step;
  }
}

// Accumulator organisation (round 2).  The whole TMEM (512 columns) is ONE stage of DT = 512/COUT output-plane slots, and
// every slot has its own full/empty mbarrier pair:
//   * the issuer acquires slot p (waits for "empty": drained AND zeroed by the epilogue) right before the first input
//     plane that touches output plane p, accumulates unconditionally, and commits "full" for p right after the last input
//     plane that touches it (input p + dil, last channel chunk);
//   * the epilogue drains the slots in plane order as they complete: tcgen05.ld, then tcgen05.st of zeros (off the tensor
//     pipe), arrive "empty", and only then the conversions / global stores / statistics.
// The next tile starts on slot 0 while the last slots of the previous tile are still being drained, so the epilogue is
// hidden exactly as with two half-size stages, but a tile is twice as deep: (DT + 2)/DT input planes per output plane
// instead of (DT/2 + 2)/(DT/2) - 1.25 instead of 1.5 for COUT = 64, 1.125 instead of 1.25 for COUT = 32.
template <int COUT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvKArgs a) {
  constexpr int DT = kConvAccCols / COUT;
  constexpr int NCG = COUT / 16;
  // per-epilogue-warp running InstanceNorm partial sums.  fp64: the per-tile fp32 partials are a fixed function of the
  // tile, so the statistics do not depend on which tiles a CTA happens to process (batch size, grid) beyond fp64 rounding.
  __shared__ double s_run[4][NCG][32];
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16-byte aligned: align to 128 by hand
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t w_region = (a.wslots * a.wchunk_bytes + 127u) & ~127u;
  const uint32_t w_addr = smem_base;
  const uint32_t s_addr = smem_base + w_region;
  const uint32_t bar_addr = s_addr + a.nstages * a.stage_bytes;
  // barrier slots (8 bytes each)
  auto full_bar = [&](int i) { return bar_addr + 8u * i; };
  auto empty_bar = [&](int i) { return bar_addr + 8u * (kConvMaxStages + i); };
  auto wfull_bar = [&](int i) { return bar_addr + 8u * (2 * kConvMaxStages + i); };
  auto wempty_bar = [&](int i) { return bar_addr + 8u * (2 * kConvMaxStages + 8 + i); };
  auto tfull_bar = [&](int i) { return bar_addr + 8u * (2 * kConvMaxStages + 16 + i); };
  auto tempty_bar = [&](int i) { return bar_addr + 8u * (2 * kConvMaxStages + 48 + i); };
  const uint32_t tmem_slot_addr = bar_addr + 8u * (2 * kConvMaxStages + 80);

  // warp index through a shuffle so that the compiler KNOWS it is warp-uniform: the role branches and everything
  // inside them (loop counters, descriptors) can then live in uniform registers, which UTCHMMA needs anyway.
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstages; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    for (int i = 0; i < a.wslots; ++i) { mbar_init(wfull_bar(i), 1); mbar_init(wempty_bar(i), 1); }
    for (int i = 0; i < DT; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
    fence_mbar_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot_addr));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const bool resident = a.nchunks <= a.wslots;

  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (lane == 0) {
      uint32_t st = 0, ph = 0, wcount = 0;
      auto load_weights = [&](int chunk, int slot) {
        mbar_expect_tx(wfull_bar(slot), a.wchunk_bytes);
        const uint8_t* src = a.wimg + (size_t)chunk * a.wchunk_bytes;
        const uint32_t dst = w_addr + slot * a.wchunk_bytes;
        for (uint32_t off = 0; off < a.wchunk_bytes; off += 16384u) {
          const uint32_t n = min(16384u, a.wchunk_bytes - off);
          bulk_g2s(dst + off, src + off, n, wfull_bar(slot));
        }
      };
      if (resident)
        for (int c = 0; c < a.nchunks; ++c) load_weights(c, c);
      for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<DT>(a, tile);
        const int dteff = min(DT, t.Dp - t.d0);
        const int q_begin = max(-a.dil, -t.d0), q_end = min(dteff + a.dil, t.Dp - t.d0);   // same range as the issuer
        for (int c = 0; c < a.nchunks; ++c) {
          if (!resident) {
            const int slot = wcount % a.wslots;
            const uint32_t par = (wcount / a.wslots) & 1u;
            mbar_wait(wempty_bar(slot), par ^ 1u);
            load_weights(c, slot);
            ++wcount;
          }
          for (int q_rel = q_begin; q_rel < q_end; ++q_rel) {
            mbar_wait(empty_bar(st), ph ^ 1u);
            mbar_expect_tx(full_bar(st), a.box_bytes);
            tma_load_4d(s_addr + st * a.stage_bytes, &tmap, full_bar(st),
                        8 * (t.w0 - a.halo), t.h0 - a.halo, t.par + a.dstep * (t.d0 + q_rel),
                        t.n * a.in_chunks_total + a.in_chunk_off + c * a.kc8);
            if (++st == (uint32_t)a.nstages) { st = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =================================== UMMA issuer ===================================
    // The whole warp runs this loop with warp-uniform control flow so that descriptors live in
    // uniform registers; only the tcgen05.mma / tcgen05.commit themselves are issued by one elected lane.
    {
      uint32_t st = 0, ph = 0, wcount = 0, wready = 0;
      uint32_t use_bits = 0;                  // bit p: parity of the number of completed uses of accumulator slot p
      const uint32_t idesc1 = umma_idesc(a.fmt, 128, COUT);
      constexpr uint32_t kIdescNStep = (uint32_t)(COUT >> 3) << 17;   // one more kd block along N
      const int dil = a.dil, nkd = a.nkd;
      const uint32_t a_hi = a.a_hi, b_hi = a.b_hi, stage16 = a.stage16;
      const uint32_t a_lo_first = a.a_lo0 | ((s_addr & 0x3FFFFu) >> 4);
      const uint32_t b_lo_lbo = ((a.b_lbo >> 4) & 0x3FFFu) << 16;
      uint32_t a_lo_st = a_lo_first;          // A descriptor low word of ring stage st
      uint32_t fbar = full_bar(0);            // full barrier of ring stage st (empty barrier = + 8 * kConvMaxStages)
      const int regular = a.regular, ntap = a.ntap, jsteps = a.jsteps;
      // bit mask instead of the count, so that the dispatch below compiles to uniform branches and not to a jump table
      const int jmode = jsteps == 1 ? 1 : (jsteps == 2 ? 2 : (jsteps == 4 ? 4 : 8));
      const uint32_t kh_step = a.kh_step, kw_step = a.kw_step, j_step = a.j_step, b_step = a.b_step;
      const int last_chunk = a.nchunks - 1;
      for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<DT>(a, tile);
        const int D = t.Dp;
        const int dteff = min(DT, D - t.d0);
        // input planes q_rel in [-dil, dteff + dil) that exist in the volume
        const int q_begin = max(-dil, -t.d0), q_end = min(dteff + dil, D - t.d0);
        int next_fresh = 0, next_done = 0;    // output planes not yet acquired / not yet committed
        for (int c = 0; c < a.nchunks; ++c) {
          int slot;
          if (resident) {
            slot = c;
            if (!((wready >> c) & 1u)) { mbar_wait(wfull_bar(c), 0); wready |= 1u << c; }
          } else {
            slot = wcount % a.wslots;
            mbar_wait(wfull_bar(slot), (wcount / a.wslots) & 1u);
          }
          const uint32_t b_lo_slot = b_lo_lbo | (((w_addr + slot * a.wchunk_bytes) & 0x3FFFFu) >> 4);
          for (int q_rel = q_begin; q_rel < q_end; ++q_rel) {
            // kd block j (j = 0,1,2) of this input plane feeds output plane q_rel + (j-1)*dil; the valid ones are contiguous
            int jlo = 0, nj = 1, p_lo = q_rel;
            if (nkd == 3) {
              const int p0 = q_rel - dil, p2 = q_rel + dil;
              jlo = (p0 < 0) + (q_rel < 0);                                   // p2 >= 0 always holds here
              nj = (p0 < dteff) + (q_rel < dteff) + (p2 < dteff) - jlo;
              p_lo = q_rel + (jlo - 1) * dil;
            }
            if (c == 0) {
              // first touch of output planes <= p_hi: their slots must have been drained and zeroed by the epilogue
              const int p_hi = p_lo + nj - 1;
              while (next_fresh <= p_hi) {
                mbar_wait(tempty_bar(next_fresh), (use_bits >> next_fresh) & 1u);
                ++next_fresh;
              }
            }
            const uint32_t dcol = tmem_base + p_lo * COUT;
            const uint32_t idesc = idesc1 + (uint32_t)(nj - 1) * kIdescNStep;
            mbar_wait(fbar, ph);
            tc_fence_after();
            uint32_t b_lo = b_lo_slot + (uint32_t)(jlo * COUT);
            if (regular) {
              if (ntap == 3) {
                // (an if-chain, not a switch: a jump table would leave the uniform datapath)
                if (jmode == 1) issue_taps3<1>(dcol, a_lo_st, a_hi, b_lo, b_hi, idesc, kh_step, kw_step, j_step, b_step, 1);
                else if (jmode & 2) issue_taps3<2>(dcol, a_lo_st, a_hi, b_lo, b_hi, idesc, kh_step, kw_step, j_step, b_step, 2);
                else if (jmode & 4) issue_taps3<4>(dcol, a_lo_st, a_hi, b_lo, b_hi, idesc, kh_step, kw_step, j_step, b_step, 4);
                else issue_taps3<0>(dcol, a_lo_st, a_hi, b_lo, b_hi, idesc, kh_step, kw_step, j_step, b_step, jsteps);
              } else {
                uint32_t a_lo = a_lo_st;
#pragma unroll 1
                for (int j = 0; j < jsteps; ++j) {
                  if (elect_one_sync()) umma_f16_lohi(dcol, a_lo, a_hi, b_lo, b_hi, idesc);
                  a_lo += j_step; b_lo += b_step;
                }
              }
            } else {
#pragma unroll
              for (int s = 0; s < 5; ++s) {   // paired taps (Cin = 8): always 5 steps
                const uint2 dl = a.dlt[s];
                if (elect_one_sync()) umma_f16_lohi(dcol, a_lo_st + dl.x, a_hi, b_lo + dl.y, b_hi, idesc);
              }
            }
            if (elect_one_sync()) umma_commit(fbar + 8u * kConvMaxStages);  // frees the activation stage when these MMAs retire
            if (c == last_chunk) {
              // output planes whose last contribution was just issued: q_rel - dil, and everything left after the last input plane
              const int p_dn = (q_rel == q_end - 1) ? dteff - 1 : q_rel - dil;
              while (next_done <= p_dn) {
                if (elect_one_sync()) umma_commit(tfull_bar(next_done));
                ++next_done;
              }
            }
            a_lo_st += stage16; fbar += 8u;
            if (++st == (uint32_t)a.nstages) { st = 0; ph ^= 1u; a_lo_st = a_lo_first; fbar = full_bar(0); }
          }
          if (!resident) {
            if (elect_one_sync()) umma_commit(wempty_bar(slot));
            ++wcount;
          }
        }
        use_bits ^= (dteff >= 32 ? 0xffffffffu : ((1u << dteff) - 1u));
      }
      __syncwarp();
    }
  } else {
    // =================================== epilogue (4 warps) ===================================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp may read
    const int row = quarter * 32 + lane;
    const int hh = row >> 3, ww = row & 7;
    uint32_t use_bits = 0;
    int run_n = -1;
    const float oscale = a.out_scale ? __ldg(a.out_scale) : 1.f;
    const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16);
    auto flush_stats = [&]() {
      if (run_n < 0 || a.stats == nullptr) return;
#pragma unroll
      for (int cg = 0; cg < NCG; ++cg) {
        double* sp = a.stats + ((size_t)run_n * COUT + cg * 16 + (lane & 15)) * 2 + (lane >> 4);
        atomicAdd(sp, s_run[quarter][cg][lane]);
        s_run[quarter][cg][lane] = 0.0;
      }
    };
#pragma unroll
    for (int cg = 0; cg < NCG; ++cg) s_run[quarter][cg][lane] = 0.0;
    // hand every slot to the issuer zeroed (phase 0 of the "empty" barriers)
#pragma unroll 1
    for (int s = 0; s < 512 / 16; ++s) tmem_st16_zero(tacc + s * 16);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < DT; ++s) mbar_arrive(tempty_bar(s));
    const size_t plane_elems = (size_t)a.H * a.W * 8;
    for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
      const TileCoord t = decode_tile<DT>(a, tile);
      if (t.n != run_n) { flush_stats(); run_n = t.n; }
      const int h = t.h0 + hh, w = t.w0 + ww;
      const bool inb = (h < a.H) && (w < a.W);
      const int dteff = min(DT, t.Dp - t.d0);
      // per-lane partial sums of this tile: red[cg*32 + i] = sum of channel cg*16+i, red[cg*32 + 16 + i] = sum of squares
      float red[2 * COUT];
#pragma unroll
      for (int i = 0; i < 2 * COUT; ++i) red[i] = 0.f;
#pragma unroll 1
      for (int p = 0; p < dteff; ++p) {
        const int d = t.par + a.dstep * (t.d0 + p);
        mbar_wait(tfull_bar(p), (use_bits >> p) & 1u);
        tc_fence_after();
        uint32_t v[COUT];
#pragma unroll
        for (int cg = 0; cg < NCG; ++cg) tmem_ld16(tacc + p * COUT + cg * 16, v + cg * 16);
        tmem_ld_wait();
        // give the slot back (zeroed) before the slow part: conversions, global stores, statistics
#pragma unroll
        for (int cg = 0; cg < NCG; ++cg) tmem_st16_zero(tacc + p * COUT + cg * 16);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(p));
        if (inb) {
#pragma unroll
          for (int cg = 0; cg < NCG; ++cg) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              f[i] = __uint_as_float(v[cg * 16 + i]) * oscale;
              red[cg * 32 + i] += f[i];
              red[cg * 32 + 16 + i] += f[i] * f[i];
            }
            const size_t eo = ((size_t)(t.n * a.out_chunks_total + a.out_chunk_off + cg * 2) * a.D + d) * plane_elems +
                              ((size_t)h * a.W + w) * 8;   // element offset
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (cg * 2 + k >= a.out_real_chunks) break;
              const size_t ek = eo + (size_t)k * a.D * plane_elems;
              if (a.out_bf16) {   // gradient-format output (grad_t), optionally accumulating
                grad_t* ok = reinterpret_cast<grad_t*>(a.out) + ek;
#ifdef SEUNET_HAVE_RED_GRAD8
                // gradient accumulation of fan-out nodes: vector reduction in L2 instead of load + add + store (the read
                // latency, once per plane and chunk, made the epilogue the bottleneck of the accumulating dgrads)
                if (a.accum_out) red_grad8(ok, f + 8 * k);
                else st_grad8(ok, f + 8 * k);
#else
                if (a.accum_out) {
                  float old[8];
                  ld_grad8_cached(ok, old);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[8 * k + i] += old[i];
                }
                st_grad8(ok, f + 8 * k);
#endif
              } else {
                st_chunk(reinterpret_cast<uint16_t*>(a.out) + ek, floats_to_chunk(f + 8 * k));
              }
            }
          }
        }
      }
      use_bits ^= (dteff >= 32 ? 0xffffffffu : ((1u << dteff) - 1u));
      if (a.stats != nullptr) {
        // lane l < 16: sum of channel cg*16+l ; lane l >= 16: sum of squares of channel cg*16+l-16.
        // Totals are kept per warp across the tiles of this persistent CTA and flushed with one fp64 atomic
        // per (channel, moment) when the sample changes / at the end: same-address atomics serialise in L2.
#pragma unroll
        for (int cg = 0; cg < NCG; ++cg) s_run[quarter][cg][lane] += (double)warp_xreduce32(red + cg * 32, lane);
      }
    }
    flush_stats();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 (Cout, Cin, k,k,k) -> UMMA image [chunk][step][khalf][nkd*COUT][8]
// ---------------------------------------------------------------------------------------------
struct PackArgs {
  int Cin_real, Cout_real, COUT, ksize, nkd, KC, nchunks, nsteps, transpose_flip, bf16, co_off, co_total;
  PackStep steps[kConvMaxSteps];
};

__global__ void conv_pack_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, const __grid_constant__ PackArgs p) {
  const int rows = p.nkd * p.COUT;
  const size_t total = (size_t)p.nchunks * p.nsteps * 2 * rows * 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = r % 8; r /= 8;
    const int row = r % rows; r /= rows;
    const int khalf = r % 2; r /= 2;
    const int s = r % p.nsteps; r /= p.nsteps;
    const int c = (int)r;
    const PackStep ps = p.steps[s];
    const int tap = khalf ? ps.tap_b : ps.tap_a;
    const int ci = c * p.KC + (khalf ? ps.cbase_b : ps.cbase_a) + e;
    const int jkd = row / p.COUT, co = row % p.COUT;
    float val = 0.f;
    if (tap >= 0) {
      int kd = (p.nkd == 3) ? (2 - jkd) : 0;
      int kh = tap / 3, kw = tap % 3;
      const int K = p.ksize, K3 = K * K * K;
      if (!p.transpose_flip) {
        if (co < p.Cout_real && ci < p.Cin_real)
          val = w[((size_t)co * p.Cin_real + ci) * K3 + (kd * K + kh) * K + kw];
      } else {
        // data gradient: "input" channels are the forward Cout, "output" channels the forward Cin,
        // taps mirrored.  Cin_real/Cout_real are given in the gradient operator's own roles.
        if (co < p.Cout_real && ci < p.Cin_real) {
          if (K == 3) { kd = 2 - kd; kh = 2 - kh; kw = 2 - kw; }
          val = w[((size_t)ci * p.co_total + p.co_off + co) * K3 + (kd * K + kh) * K + kw];
        }
      }
    }
    if (p.bf16) { __nv_bfloat16 b = __float2bfloat16_rn(val); img[i] = *reinterpret_cast<uint16_t*>(&b); }
    else { act_t h = f2act(val); img[i] = *reinterpret_cast<uint16_t*>(&h); }
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static constexpr uint32_t kSmemBudgetMax = 220u * 1024u;  // + 4 KB static smem (s_run, fp64) + alignment slack <= 227 KB
static uint32_t smem_budget() {   // SEUNET_CONV_SMEM_KB: experiment knob (leave room for co-resident streaming kernels)
  static const uint32_t v = [] {
    const char* e = getenv("SEUNET_CONV_SMEM_KB");
    uint32_t kb = e ? (uint32_t)atoi(e) : 220u;
    return std::min(kSmemBudgetMax, std::max(64u, kb) * 1024u);
  }();
  return v;
}
#define kSmemBudget smem_budget()
static constexpr uint32_t kBarBytes = 8u * (2 * kConvMaxStages + 82) + 64u;   // stage / weight / per-slot accumulator barriers + TMEM address

int conv_geom_init(ConvGeom* g, int Cin_real, int Cout_real, int ksize, int dil, int bf16) {
  memset(g, 0, sizeof(*g));
#ifdef SEUNET_ACT_BF16
  g->bf16 = 1;
#else
  g->bf16 = bf16 < 0 ? 0 : bf16;
#endif
  g->Cin_real = Cin_real; g->Cout_real = Cout_real; g->ksize = ksize; g->dil = (ksize == 3) ? dil : 0;
  if (ksize != 1 && ksize != 3) { seunet_set_error("conv: kernel size %d unsupported", ksize); return 1; }
  if (ksize == 3 && dil != 1 && dil != 2) { seunet_set_error("conv: dilation %d unsupported", dil); return 1; }
  g->COUT = Cout_real <= 16 ? 16 : (Cout_real <= 32 ? 32 : 64);
  if (Cout_real > 64) { seunet_set_error("conv: Cout %d > 64 unsupported", Cout_real); return 1; }
  const int nkd = ksize == 3 ? 3 : 1;
  const int ntaps = ksize == 3 ? 9 : 1;
  const int halo = g->dil;
  const int HV = (kConvTileH + 2 * halo) * (kConvTileW + 2 * halo);
  int Cin = Cin_real <= 8 ? 8 : ((Cin_real + 15) / 16) * 16;
  if (ksize == 1 && Cin == 8) Cin = 16;
  g->Cin = Cin;
  g->paired = (Cin == 8);
  if (g->paired) {
    // 9 (kh,kw) taps -> 5 K=16 steps.  The unpaired tap must be tap 0: its dummy second K half (zero weights)
    // then reads the next voxel, which is still inside the halo tile - a dummy half after the LAST tap would read
    // past the TMA box, and 0 * (stale shared memory that happens to be Inf/NaN) poisons the accumulator.
    g->KC = 8; g->nchunks = 1; g->nsteps = 5;
    for (int s = 0; s < 5; ++s) {
      g->psteps[s].tap_a = (int8_t)(s == 0 ? 0 : 2 * s - 1);
      g->psteps[s].tap_b = (int8_t)(s == 0 ? -1 : 2 * s);
      g->psteps[s].cbase_a = 0; g->psteps[s].cbase_b = 0;
    }
  } else {
    const size_t total_w = (size_t)Cin * ntaps * nkd * g->COUT * 2;
    int KC;
    if (total_w <= 112u * 1024u) {
      KC = 64;
      while (Cin % KC) KC /= 2;
      g->wslots = Cin / KC;
    } else {
      KC = 16;
      g->wslots = 2;
    }
    g->KC = KC; g->nchunks = Cin / KC;
    g->nsteps = ntaps * (KC / 16);
    for (int t = 0; t < ntaps; ++t)
      for (int j = 0; j < KC / 16; ++j) {
        PackStep& ps = g->psteps[t * (KC / 16) + j];
        ps.tap_a = ps.tap_b = (int8_t)t;
        ps.cbase_a = (int16_t)(j * 16); ps.cbase_b = (int16_t)(j * 16 + 8);
      }
  }
  if (g->paired) g->wslots = 1;
  if (g->nsteps > kConvMaxSteps) { seunet_set_error("conv: too many steps"); return 1; }
  g->wchunk_bytes = (uint32_t)g->nsteps * 2u * nkd * g->COUT * 16u;
  g->box_bytes = (uint32_t)HV * g->KC * 2u;
  g->stage_bytes = (g->box_bytes + 127u) & ~127u;
  const uint32_t wregion = ((uint32_t)g->wslots * g->wchunk_bytes + 127u) & ~127u;
  int nst = (int)((kSmemBudget - wregion - kBarBytes - 128u) / g->stage_bytes);
  // TMA is latency-bound: keep >= ~96 KB in flight per SM when the stages are small (ec1/ec2: 2.9 KB per plane)
  nst = std::min(nst, std::max(4, std::min(kConvMaxStages, (int)(98304u / g->stage_bytes))));
  if (nst < 2) { seunet_set_error("conv: shared memory budget exceeded"); return 1; }
  g->nstages = nst;
  g->smem_bytes = wregion + nst * g->stage_bytes + kBarBytes + 128u;
  return 0;
}

int conv_pack_weights(const ConvGeom& g, const float* w_fp32, void* wimg, int transpose_flip, cudaStream_t st, int co_off,
                      int co_total) {
  PackArgs p;
  memset(&p, 0, sizeof(p));
  p.Cin_real = g.Cin_real; p.Cout_real = g.Cout_real; p.COUT = g.COUT; p.ksize = g.ksize;
  p.nkd = g.ksize == 3 ? 3 : 1; p.KC = g.KC; p.nchunks = g.nchunks; p.nsteps = g.nsteps;
  p.transpose_flip = transpose_flip;
  p.bf16 = g.bf16;
  p.co_off = co_off;
  p.co_total = co_total < 0 ? g.Cout_real : co_total;
  memcpy(p.steps, g.psteps, sizeof(p.steps));
  const size_t total = g.wimg_bytes() / 2;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 1024);
  conv_pack_kernel<<<blocks, 256, 0, st>>>(w_fp32, (uint16_t*)wimg, p);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
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

int conv_launch_init(ConvLaunch* L, const ConvGeom& g, int N, int D, int H, int W,
                     const void* in, int in_chunks_total, int in_chunk_off,
                     void* out, int out_chunks_total, int out_chunk_off,
                     double* stats, const void* wimg, int num_sms, int accum_out, int out_real_chunks, int grad_out,
                     const float* out_scale) {
  L->g = g;
  ConvKArgs& a = L->a;
  memset(&a, 0, sizeof(a));
  const int nkd = g.ksize == 3 ? 3 : 1;
  const int halo = g.dil;
  const int DT = kConvAccCols / g.COUT;
  const int lineW = kConvTileW + 2 * halo;
  const int HV = (kConvTileH + 2 * halo) * lineW;
  a.N = N; a.D = D; a.H = H; a.W = W;
  a.tilesW = (W + kConvTileW - 1) / kConvTileW;
  a.tilesH = (H + kConvTileH - 1) / kConvTileH;
  // dilation 2: split the planes into the two parity classes (plane distance of the kd taps becomes 1 inside a class)
  a.dstep = (nkd == 3 && g.dil == 2) ? 2 : 1;
  a.tilesD = a.dstep == 2 ? 2 * (((D + 1) / 2 + DT - 1) / DT) : (D + DT - 1) / DT;
  a.numTiles = a.tilesW * a.tilesH * a.tilesD * N;
  a.dil = nkd == 3 ? (a.dstep == 2 ? 1 : g.dil) : 0;
  a.nkd = nkd; a.halo = halo;
  a.nchunks = g.nchunks; a.kc8 = g.KC / 8;
  a.in_chunks_total = in_chunks_total; a.in_chunk_off = in_chunk_off;
  a.out_chunks_total = out_chunks_total; a.out_chunk_off = out_chunk_off;
  a.nstages = g.nstages; a.wslots = g.wslots; a.nsteps = g.nsteps;
  a.stage_bytes = g.stage_bytes; a.box_bytes = g.box_bytes; a.wchunk_bytes = g.wchunk_bytes;
  a.a_sbo = (uint32_t)lineW * 16u;
  a.b_lbo = (uint32_t)nkd * g.COUT * 16u;
  a.wimg = (const uint8_t*)wimg;
  a.out = out;
  a.stats = stats;
  a.fmt = g.bf16 ? 1u : 0u;
  a.out_bf16 = grad_out;
  a.out_scale = out_scale;
  a.accum_out = accum_out;
  a.out_real_chunks = out_real_chunks < 0 ? g.COUT / 8 : out_real_chunks;
  const uint32_t a_lbo = (uint32_t)HV * 16u;
  auto tapoff = [&](int t) { return (uint32_t)(((t / 3) * g.dil * lineW + (t % 3) * g.dil) * 16); };
  a.a_hi = (((a.a_sbo >> 4) & 0x3FFFu)) | (1u << 14);   // SBO | descriptor version (bit 46)
  a.b_hi = ((128u >> 4) & 0x3FFFu) | (1u << 14);
  a.stage16 = a.stage_bytes >> 4;
  a.b_step = (uint32_t)(2 * nkd * g.COUT);   // one step of the packed weight image = 2 K halves x nkd x COUT rows of 16 B
  if (g.paired) {
    if (g.nsteps != 5 || g.nsteps > kConvTableSteps) { seunet_set_error("conv: paired schedule must have 5 steps"); return 1; }
    a.regular = 0; a.a_lo0 = 0;
    for (int s = 0; s < g.nsteps; ++s) {
      const PackStep& ps = g.psteps[s];
      const uint32_t off = tapoff(ps.tap_a);
      const uint32_t lbo = ps.tap_b >= 0 ? tapoff(ps.tap_b) - tapoff(ps.tap_a) : 16u;
      a.dlt[s] = make_uint2((off >> 4) | ((lbo >> 4) << 16), (uint32_t)s * a.b_step);
    }
  } else {
    // steps are ordered tap-major, 16-channel block minor (conv_geom_init), weights packed in the same order
    a.regular = 1; a.a_lo0 = ((a_lbo >> 4) & 0x3FFFu) << 16;
    a.ntap = g.ksize == 3 ? 3 : 1;
    a.jsteps = g.KC / 16;
    a.kh_step = (uint32_t)(g.dil * lineW); a.kw_step = (uint32_t)g.dil;
    a.j_step = 2u * (a_lbo >> 4);
    if (g.nsteps != (g.ksize == 3 ? 9 : 1) * a.jsteps) { seunet_set_error("conv: step schedule mismatch"); return 1; }
  }
  if (in_chunk_off + g.Cin / 8 > in_chunks_total) { seunet_set_error("conv: input slice exceeds buffer"); return 1; }
  if (out_chunk_off + a.out_real_chunks > out_chunks_total) { seunet_set_error("conv: output slice exceeds buffer"); return 1; }

  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { seunet_set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)"); return 1; }
  cuuint64_t gdim[4] = {(cuuint64_t)8 * W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N * in_chunks_total};
  cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)(8 * lineW), (cuuint32_t)(kConvTileH + 2 * halo), 1u, (cuuint32_t)(g.KC / 8)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = g.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(&L->tmap, dt, 4, const_cast<void*>(in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { seunet_set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return 1; }
  L->grid = std::min(a.numTiles, num_sms);
  return 0;
}

template <int COUT>
static int conv_launch_t(const ConvLaunch& L, cudaStream_t st) {
  // the dynamic shared-memory limit is a PER-DEVICE function attribute (nn.DataParallel replicas launch on several devices
  // of one process, one host thread each): remember it per device; a racing duplicate call is harmless
  static bool attr_set[64] = {};
  int dev = 0;
  SEUNET_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    SEUNET_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  conv_tc_kernel<COUT><<<L.grid, kConvThreads, L.g.smem_bytes, st>>>(L.tmap, L.a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int conv_launch_run(const ConvLaunch& L, cudaStream_t st) {
  switch (L.g.COUT) {
    case 16: return conv_launch_t<16>(L, st);
    case 32: return conv_launch_t<32>(L, st);
    case 64: return conv_launch_t<64>(L, st);
  }
  seunet_set_error("conv: bad COUT %d", L.g.COUT);
  return 1;
}
