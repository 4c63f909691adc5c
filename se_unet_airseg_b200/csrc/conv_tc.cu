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
  t.n = tile; t.d0 = td * a.dt_use; t.h0 = th * kConvTileH; t.w0 = tw * kConvTileW;
  return t;
}

// One input plane of a 3x3x3 layer: 9 in-plane taps x J 16-channel blocks, fully unrolled.  Tap offsets into the halo tile
// and weight-image offsets are IMMEDIATES (functions of DIL, J, COUT only); the only register operands are the plane's
// base descriptors.  Called by the single issuing thread.
template <int COUT, int DIL, int J>
__device__ __forceinline__ void issue_plane3(uint32_t dcol, uint32_t a_plane, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t j_step) {
  constexpr uint32_t LW = kConvTileW + 2 * DIL;     // halo line in voxels (= 16-byte units)
  constexpr uint32_t BST = 2u * 3u * COUT;          // one K=16 step of the weight image: 2 K halves x 3 kd blocks x COUT rows
  uint32_t aj[J];
#pragma unroll
  for (int j = 0; j < J; ++j) aj[j] = a_plane + (uint32_t)j * j_step;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int j = 0; j < J; ++j)
        umma_f16_lohi(dcol, aj[j] + (uint32_t)(kh * DIL * LW + kw * DIL), a_hi, b_lo + (uint32_t)(((kh * 3 + kw) * J + j) * BST), b_hi, idesc);
}

// Accumulator organisation (round 2).  The whole TMEM (512 columns) is ONE stage of DT = 512/COUT output-plane slots in
// kConvGroups groups with a full/empty mbarrier pair each:
//   * the issuer acquires a group (waits for "empty": drained by the epilogue) right before the first input plane that
//     touches it, clears it with one zero-operand UMMA, accumulates unconditionally, and commits "full" right after the
//     last input plane that touches it (last channel chunk);
//   * the epilogue drains the groups in order as they complete: tcgen05.ld, conversions / global stores / statistics,
//     arrive "empty".
// The next tile starts on group 0 while the last group of the previous tile is still being drained, so the epilogue is
// hidden as with two half-size stages, but a tile is twice as deep: (DT + 2)/DT input planes per output plane instead of
// (DT/2 + 2)/(DT/2) - 1.25 instead of 1.5 for COUT = 64, 1.125 instead of 1.25 for COUT = 32.
// Every tile runs the SAME schedule: DT + 2 input planes (TMA zero-fills planes outside the volume, their MMAs are
// skipped) in boxes of `pb` planes, so the issuer pays one mbarrier wait and one commit per BOX, not per plane -
// measured (tools/umma_bench.cu, groups sync/issue): a try_wait on a completed barrier costs the issuing thread ~120
// cycles and a tcgen05.commit ~125, neither overlaps with issuing, and an M=128 tcgen05.mma cannot be issued faster than
// one per ~45 cycles whatever N is, so for N <= 96 every cycle of issuer overhead is a cycle of idle tensor pipe.
// Issue modes of the per-tile plane loop (one template instance each, selected ONCE per tile and channel chunk: inside the
// loop there is no layer-dependent branching and no kernel-parameter reload - ncu's source view showed the issuer stalled
// on the latency of exactly those uniform-datapath compare/branch chains, ~500 cycles per plane).
enum { kModeTaps3 = 0 /* + (dil-1)*3 + log2(J): six 3x3x3 variants */, kModePaired = 6, kModePointwise = 7 };

template <int COUT, int MODE>
__device__ __forceinline__ void issue_mode(uint32_t dcol, uint32_t a_plane, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t j_step, uint32_t b_step, int jsteps, const uint2* dlt) {
  if (MODE < 6) {
    issue_plane3<COUT, 1 + MODE / 3, 1 << (MODE % 3)>(dcol, a_plane, a_hi, b_lo, b_hi, idesc, j_step);
  } else if (MODE == kModePaired) {
#pragma unroll
    for (int s = 0; s < 5; ++s) umma_f16_lohi(dcol, a_plane + dlt[s].x, a_hi, b_lo + dlt[s].y, b_hi, idesc);   // paired taps (Cin = 8)
  } else {
#pragma unroll 1
    for (int j = 0; j < jsteps; ++j) {
      umma_f16_lohi(dcol, a_plane, a_hi, b_lo, b_hi, idesc);
      a_plane += j_step; b_lo += b_step;
    }
  }
}

// Loop-carried state of the issuer across boxes / tiles (activation ring position).
struct IssueRing { uint32_t st, ph, a_lo_st, fbar; };

// All planes k = 0 .. KMAX of one (tile, channel chunk).  Plane k first touches output plane k (3x3x3: through its kd = 0
// tap) and gives output plane k - 2*DIL its last contribution, so the accumulator-group events sit at fixed k:
//   wait "empty" + clear of group k/G          when k % G == 0, k < DT              (first channel chunk only)
//   commit "full" of group (k - 2*DIL)/G       when (k - 2*DIL) % G == G - 1        (last channel chunk only)
template <int COUT, int MODE>
__device__ __forceinline__ void issue_tile(IssueRing& r, const uint32_t a_lo_first, const uint32_t full0, const int nstages,
                                           const uint32_t stage16, const uint32_t plane16, const int pb, const int kv0, const int kv1,
                                           const bool first_chunk, const bool last_chunk, const uint32_t tphase, const uint32_t tmem_base,
                                           const uint32_t tempty0, const uint32_t tfull0, const uint32_t idesc1, const uint32_t idesc_clear,
                                           const uint64_t zdesc, const uint32_t a_hi, const uint32_t b_hi, const uint32_t b_lo_slot,
                                           const uint32_t j_step, const uint32_t b_step, const int jsteps, const uint2* dlt) {
  constexpr int DT = kConvAccCols / COUT;
  constexpr int G = DT / kConvGroups;
  constexpr int DIL = (MODE == kModePointwise) ? 0 : 1;     // plane distance of the kd taps (parity classes make it 1)
  constexpr int KMAX = DT - 1 + 2 * DIL;
  constexpr uint32_t kIdescNStep = (uint32_t)(COUT >> 3) << 17;   // one more kd block along N
  uint32_t dcol = tmem_base, idesc = idesc1, b_lo = b_lo_slot + (uint32_t)(2 * DIL * COUT);
  uint32_t a_plane = r.a_lo_st;
  int box_left = 0;
  const uint32_t kvspan = (uint32_t)(kv1 - kv0);
  // One plane.  Everything that depends on the plane's position inside its group is a compile-time constant (the group loop
  // below is unrolled), so the only run-time work besides the MMAs is the box bookkeeping and the in-volume test.
  auto plane = [&](const int k, const bool grow, const bool shrink_after, const int commit_group) {
    if (box_left == 0) {
      mbar_wait(r.fbar, r.ph);      // TMA -> mbarrier -> MMA: ordered by the mbarrier itself
      a_plane = r.a_lo_st;
      box_left = pb;
    }
    if ((uint32_t)(k - kv0) <= kvspan)       // planes outside the volume arrive as zeros: nothing to add
      issue_mode<COUT, MODE>(dcol, a_plane, a_hi, b_lo, b_hi, idesc, j_step, b_step, jsteps, dlt);
    if (commit_group >= 0 && last_chunk) umma_commit(tfull0 + 8u * (uint32_t)commit_group);
    // next plane: stacked kd blocks grow 1 -> 2 -> 3 at the start of the tile, shrink 3 -> 2 -> 1 at its end
    a_plane += plane16;
    if (grow) { idesc += kIdescNStep; b_lo -= COUT; }
    else { dcol += COUT; if (shrink_after) idesc -= kIdescNStep; }
    if (--box_left == 0 || k == KMAX) {
      umma_commit(r.fbar + 8u * kConvMaxStages);  // frees the activation stage when these MMAs retire
      box_left = 0;
      r.a_lo_st += stage16; r.fbar += 8u;
      if (++r.st == (uint32_t)nstages) { r.st = 0; r.ph ^= 1u; r.a_lo_st = a_lo_first; r.fbar = full0; }
    }
  };
#pragma unroll 1
  for (int g = 0; g < kConvGroups; ++g) {
    if (first_chunk) {
      // acquire the group (drained by the epilogue) and clear it with one zero-operand UMMA
      mbar_wait(tempty0 + 8u * (uint32_t)g, tphase);
      tc_fence_after();
      umma_f16(tmem_base + g * G * COUT, zdesc, zdesc, idesc_clear, 0u);
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int k = g * G + i;
      if (DIL == 1)   // output plane k-2 is complete after plane k: group g-1 after the plane with i == 1
        plane(k, g == 0 && i < 2, g == kConvGroups - 1 && i == G - 1, (i == 1 && g > 0) ? g - 1 : -1);
      else            // pointwise: plane k only touches output plane k
        plane(k, false, false, i == G - 1 ? g : -1);
    }
  }
  if (DIL == 1) {     // the two halo planes behind the last output plane
    plane(DT, false, true, -1);
    plane(DT + 1, false, false, kConvGroups - 1);
  }
}

#ifdef SEUNET_CONV_PROFILE
// developer build (tools/conv_profile.sh): per-CTA cycle counters of the three roles, printed by the host after each launch
__device__ long long g_conv_prof[148 * 8];
#define PROF_T0(v) const long long v = clock64()
#define PROF_ADD(acc, v) acc += clock64() - v
#else
#define PROF_T0(v)
#define PROF_ADD(acc, v)
#endif

template <int COUT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvKArgs a) {
  constexpr int DT = kConvAccCols / COUT;
  constexpr int G = DT / kConvGroups;   // output planes per accumulator group
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
  auto tempty_bar = [&](int i) { return bar_addr + 8u * (2 * kConvMaxStages + 24 + i); };
  const uint32_t tmem_slot_addr = bar_addr + 8u * (2 * kConvMaxStages + 32);
  const uint32_t zero_addr = bar_addr + 8u * (2 * kConvMaxStages + 48);   // 128 zero bytes (128-byte aligned): operands of the accumulator-clearing UMMA

  // warp index through a shuffle so that the compiler KNOWS it is warp-uniform: the role branches and everything
  // inside them (loop counters, descriptors) can then live in uniform registers, which UTCHMMA needs anyway.
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstages; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    for (int i = 0; i < a.wslots; ++i) { mbar_init(wfull_bar(i), 1); mbar_init(wempty_bar(i), 1); }
    for (int i = 0; i < kConvGroups; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
    fence_mbar_init();
    tma_prefetch_desc(&tmap);
  }
  if (threadIdx.x < 32) asm volatile("st.shared.u32 [%0], %1;" ::"r"(zero_addr + 4u * threadIdx.x), "r"(0u) : "memory");
  fence_proxy_async();   // generic-proxy zeros must be visible to the tensor core (async proxy)
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
  const int q0 = -a.dil;                 // first input plane of a tile (tile-relative); dil is 1 (3x3x3) or 0 (1x1x1) here

  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (lane == 0) {
      uint32_t st = 0, ph = 0, wcount = 0;
#ifdef SEUNET_CONV_PROFILE
      long long p_wait = 0; PROF_T0(p_t0);
#endif
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
        for (int c = 0; c < a.nchunks; ++c) {
          if (!resident) {
            const int slot = wcount % a.wslots;
            const uint32_t par = (wcount / a.wslots) & 1u;
            mbar_wait(wempty_bar(slot), par ^ 1u);
            load_weights(c, slot);
            ++wcount;
          }
          for (int bx = 0; bx < a.nboxes; ++bx) {
            // one box = pb consecutive planes of the tile's parity class; planes outside the volume arrive as zeros
            { PROF_T0(tw); mbar_wait(empty_bar(st), ph ^ 1u); PROF_ADD(p_wait, tw); }
            mbar_expect_tx(full_bar(st), a.box_bytes);
            tma_load_4d(s_addr + st * a.stage_bytes, &tmap, full_bar(st),
                        8 * (t.w0 - a.halo), t.h0 - a.halo, t.par + a.dstep * (t.d0 + q0 + bx * a.pb),
                        t.n * a.in_chunks_total + a.in_chunk_off + c * a.kc8);
            if (++st == (uint32_t)a.nstages) { st = 0; ph ^= 1u; }
          }
        }
      }
#ifdef SEUNET_CONV_PROFILE
      if (blockIdx.x < 148) { g_conv_prof[blockIdx.x * 8 + 0] = clock64() - p_t0; g_conv_prof[blockIdx.x * 8 + 1] = p_wait; }
#endif
    }
  } else if (warp == 1) {
    // =================================== UMMA issuer ===================================
    // ONE thread runs the whole issue loop.  (Round 1 ran it warp-converged with an elected lane per instruction; ncu's
    // source view shows what that costs: the BSYNC that re-converges the warp after every `if (elect_one_sync())` region
    // waits on the scoreboard of the tcgen05.mma instructions inside it, i.e. drains the tensor pipe's instruction queue -
    // once per MMA in round 1 (66-68 cycles per MMA instead of 56), once per plane with an unrolled plane.  A single
    // active thread never re-converges, so the queue stays full across planes, boxes and tiles.)
    // (`elect_one_sync()` rather than `lane == 0`: ptxas recognises the elect pattern as "exactly one thread active" and keeps
    // the descriptor arithmetic on the uniform datapath; with a lane test every tcgen05.mma is wrapped in a vote loop.)
    if (elect_one_sync()) {
      uint32_t wcount = 0, wready = 0, tphase = 0;
      const uint32_t idesc1 = umma_idesc(a.fmt, 128, COUT);
      // One UMMA with all-zero operands and accumulate OFF clears a whole accumulator group (G * COUT = 128 columns, 64
      // cycles) right after it is acquired: 4x cheaper than tcgen05.st from the epilogue warps, whose TMEM traffic stalls
      // the tensor pipe just like their tcgen05.ld does (measured: ~30 cycles of lost MMA time per x16 access).
      const uint32_t idesc_clear = umma_idesc(a.fmt, 128, G * COUT);
      const uint64_t zdesc = umma_desc(zero_addr, 0, 0);   // every core matrix reads the same 128 zero bytes
      const int pb = a.pb;
      const uint32_t a_hi = a.a_hi, b_hi = a.b_hi, stage16 = a.stage16, plane16 = a.plane16;
      const uint32_t a_lo_first = a.a_lo0 | ((s_addr & 0x3FFFFu) >> 4);
      const uint32_t b_lo_lbo = ((a.b_lbo >> 4) & 0x3FFFu) << 16;
      // activation ring position: stage, phase, A descriptor low word and full barrier of the stage (empty barrier = + 8 * kConvMaxStages)
      IssueRing ring{0u, 0u, a_lo_first, full_bar(0)};
      const int jsteps = a.jsteps;
      const int mode = !a.regular ? kModePaired : (a.ntap == 3 ? (a.cdil - 1) * 3 + (jsteps == 1 ? 0 : (jsteps == 2 ? 1 : 2)) : kModePointwise);
      const uint32_t j_step = a.j_step, b_step = a.b_step;
      const int last_chunk = a.nchunks - 1;
#ifdef SEUNET_CONV_PROFILE
      long long i_full = 0, i_tempty = 0, i_w = 0; PROF_T0(i_t0);
#endif
      for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<DT>(a, tile);
        const int D = t.Dp;
        for (int c = 0; c < a.nchunks; ++c) {
          int slot;
          if (resident) {
            slot = c;
            if (!((wready >> c) & 1u)) { PROF_T0(tw); mbar_wait(wfull_bar(c), 0); wready |= 1u << c; PROF_ADD(i_w, tw); }
          } else {
            slot = wcount % a.wslots;
            PROF_T0(tw); mbar_wait(wfull_bar(slot), (wcount / a.wslots) & 1u); PROF_ADD(i_w, tw);
          }
          const uint32_t b_lo_slot = b_lo_lbo | (((w_addr + slot * a.wchunk_bytes) & 0x3FFFFu) >> 4);
          // planes inside the volume (k = q - q0) that feed the dt_use output planes this tile owns (shallow tiles, see
          // conv_launch_init: the planes behind them are loaded but not multiplied)
          const int kv0 = max(0, -q0 - t.d0), kv1 = min(D - 1 - t.d0 - q0, a.dt_use - 1 - 2 * q0);
#define SEUNET_ISSUE_TILE(M) issue_tile<COUT, M>(ring, a_lo_first, full_bar(0), a.nstages, stage16, plane16, pb, kv0, kv1, c == 0, \
                                                 c == last_chunk, tphase, tmem_base, tempty_bar(0), tfull_bar(0), idesc1, idesc_clear, \
                                                 zdesc, a_hi, b_hi, b_lo_slot, j_step, b_step, jsteps, a.dlt)
          // (an if-chain, not a switch: a jump table would leave the uniform datapath)
          if (mode == 0) SEUNET_ISSUE_TILE(0);
          else if (mode == 1) SEUNET_ISSUE_TILE(1);
          else if (mode == 2) SEUNET_ISSUE_TILE(2);
          else if (mode == 3) SEUNET_ISSUE_TILE(3);
          else if (mode == 4) SEUNET_ISSUE_TILE(4);
          else if (mode == 5) SEUNET_ISSUE_TILE(5);
          else if (mode == kModePaired) SEUNET_ISSUE_TILE(kModePaired);
          else SEUNET_ISSUE_TILE(kModePointwise);
#undef SEUNET_ISSUE_TILE
          if (!resident) {
            umma_commit(wempty_bar(slot));
            ++wcount;
          }
        }
        tphase ^= 1u;
      }
#ifdef SEUNET_CONV_PROFILE
      if (blockIdx.x < 148) {
        g_conv_prof[blockIdx.x * 8 + 2] = clock64() - i_t0; g_conv_prof[blockIdx.x * 8 + 3] = i_full;
        g_conv_prof[blockIdx.x * 8 + 4] = i_tempty; g_conv_prof[blockIdx.x * 8 + 5] = i_w;
      }
#endif
    }
    __syncwarp();
  } else {
    // =================================== epilogue (4 warps) ===================================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp may read
    const int row = quarter * 32 + lane;
    const int hh = row >> 3, ww = row & 7;
    uint32_t tphase = 0;
    int run_n = -1;
    const bool has_scale = a.out_scale != nullptr;
    const float oscale = has_scale ? __ldg(a.out_scale) : 1.f;
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
    // all groups start out free (phase 0 of the "empty" barriers); the issuer clears a group when it acquires it
    if (lane == 0)
      for (int s = 0; s < kConvGroups; ++s) mbar_arrive(tempty_bar(s));
    const size_t plane_elems = (size_t)a.H * a.W * 8;
#ifdef SEUNET_CONV_PROFILE
    long long e_wait = 0; PROF_T0(e_t0);
#endif
    for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
      const TileCoord t = decode_tile<DT>(a, tile);
      if (t.n != run_n) { flush_stats(); run_n = t.n; }
      const int h = t.h0 + hh, w = t.w0 + ww;
      const bool inb = (h < a.H) && (w < a.W);
      const int dteff = min(a.dt_use, t.Dp - t.d0);
      // per-lane partial sums of this tile: red[cg*32 + i] = sum of channel cg*16+i, red[cg*32 + 16 + i] = sum of squares
      float red[2 * COUT];
#pragma unroll
      for (int i = 0; i < 2 * COUT; ++i) red[i] = 0.f;
#pragma unroll 1
      for (int g = 0; g < kConvGroups; ++g) {
        { PROF_T0(tw); mbar_wait(tfull_bar(g), tphase); PROF_ADD(e_wait, tw); }
        tc_fence_after();
#pragma unroll 1
        for (int p = g * G; p < (g + 1) * G; ++p) {
          if (p < dteff) {   // (warp-uniform) planes past the end of the volume hold nothing
            const int d = t.par + a.dstep * (t.d0 + p);
            uint32_t v[COUT];
#ifdef SEUNET_CONV_PROFILE
            if (a.dbg & 1) {   // experiment: no TMEM reads at all (results are garbage)
#pragma unroll
              for (int i = 0; i < COUT; ++i) v[i] = 0x3f800000u;
            } else
#endif
            {
#pragma unroll
              for (int cg = 0; cg < NCG; ++cg) tmem_ld16(tacc + p * COUT + cg * 16, v + cg * 16);
              tmem_ld_wait();
            }
            if (inb) {
#pragma unroll
              for (int cg = 0; cg < NCG; ++cg) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  f[i] = __uint_as_float(v[cg * 16 + i]);
                  if (has_scale) f[i] *= oscale;
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
        }
        // give the group back (the issuer clears it with one UMMA when it re-acquires it)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(g));
      }
      tphase ^= 1u;
      if (a.stats != nullptr) {
        // lane l < 16: sum of channel cg*16+l ; lane l >= 16: sum of squares of channel cg*16+l-16.
        // Totals are kept per warp across the tiles of this persistent CTA and flushed with one fp64 atomic
        // per (channel, moment) when the sample changes / at the end: same-address atomics serialise in L2.
#pragma unroll
        for (int cg = 0; cg < NCG; ++cg) s_run[quarter][cg][lane] += (double)warp_xreduce32(red + cg * 32, lane);
      }
    }
    flush_stats();
#ifdef SEUNET_CONV_PROFILE
    if (warp == 2 && lane == 0 && blockIdx.x < 148) { g_conv_prof[blockIdx.x * 8 + 6] = clock64() - e_t0; g_conv_prof[blockIdx.x * 8 + 7] = e_wait; }
#endif
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

// One element of the packed image: image index i of a layer described by `p`, whose K=16 step s maps to taps / channel
// bases `ps`.  Shared by the single-layer kernel (stand-alone conv API, tests) and the batched one (plans).
template <class P>
__device__ __forceinline__ uint16_t pack_element(const float* __restrict__ w, const P& p, const PackStep ps, int c, int khalf, int row, int e) {
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
  if (p.bf16) { __nv_bfloat16 b = __float2bfloat16_rn(val); return *reinterpret_cast<uint16_t*>(&b); }
  act_t h = f2act(val);
  return *reinterpret_cast<uint16_t*>(&h);
}

__global__ void conv_pack_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, const __grid_constant__ PackArgs p) {
  const int rows = p.nkd * p.COUT;
  const size_t total = (size_t)p.nchunks * p.nsteps * 2 * rows * 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = r % 8; r /= 8;
    const int row = r % rows; r /= rows;
    const int khalf = r % 2; r /= 2;
    const int s = r % p.nsteps; r /= p.nsteps;
    img[i] = pack_element(w, p, p.steps[s], (int)r, khalf, row, e);
  }
}

// All layers of a plan in one launch (blockIdx.y = layer): a training step re-packs ~70 images (forward convs + the
// mirrored data-gradient pieces) after every optimizer step, and at one patch per rank 70 launches of a few microseconds
// each were 3 % of the step.  The step table is a function of (paired, KC) only (conv_geom_init), so a job is 72 bytes and
// 48 of them travel as kernel parameters.
struct PackBatch { ConvPackJob job[kConvPackBatch]; };

__global__ void __launch_bounds__(256) conv_pack_batch_kernel(const float* __restrict__ params, uint8_t* __restrict__ wimg,
                                                              const __grid_constant__ PackBatch b) {
  const ConvPackJob& p = b.job[blockIdx.y];
  const float* w = params + p.src_off;
  uint16_t* img = reinterpret_cast<uint16_t*>(wimg + p.dst_off);
  const unsigned rows = (unsigned)(p.nkd * p.COUT), nsteps = (unsigned)p.nsteps, jsteps = (unsigned)(p.KC / 16);
  const unsigned total = (unsigned)p.nchunks * nsteps * 2u * rows * 8u;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r = i;
    const int e = (int)(r % 8u); r /= 8u;
    const int row = (int)(r % rows); r /= rows;
    const int khalf = (int)(r % 2u); r /= 2u;
    const int s = (int)(r % nsteps); r /= nsteps;
    PackStep ps;   // the same table conv_geom_init writes into ConvGeom::psteps
    if (p.paired) {
      ps.tap_a = (int8_t)(s == 0 ? 0 : 2 * s - 1); ps.tap_b = (int8_t)(s == 0 ? -1 : 2 * s);
      ps.cbase_a = 0; ps.cbase_b = 0;
    } else {
      const int t = s / (int)jsteps, j = s % (int)jsteps;
      ps.tap_a = ps.tap_b = (int8_t)t;
      ps.cbase_a = (int16_t)(j * 16); ps.cbase_b = (int16_t)(j * 16 + 8);
    }
    img[i] = pack_element(w, p, ps, (int)r, khalf, row, e);
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
static constexpr uint32_t kBarBytes = 8u * (2 * kConvMaxStages + 48) + 128u + 64u;   // stage / weight / accumulator-group barriers, TMEM address, zero block

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
  const uint32_t wregion = ((uint32_t)g->wslots * g->wchunk_bytes + 127u) & ~127u;
  const uint32_t avail = kSmemBudget - wregion - kBarBytes - 128u;
  // Planes per TMA box / ring stage.  A tile always consumes DTIN = DT (+2 for 3x3x3) input planes; the issuer pays ~300
  // cycles of synchronisation per box (see the kernel comment), an unused plane at the end of the last box only costs
  // L2->SMEM traffic.  At least two stages must fit.
  const uint32_t plane_bytes = (uint32_t)HV * g->KC * 2u;
  const int DTIN = kConvAccCols / g->COUT + (nkd == 3 ? 2 : 0);
  const uint32_t cap = std::min<uint32_t>(56u * 1024u, avail / 2u);
  int best_pb = 1; long best_cost = -1;
  for (int pb = 1; pb <= std::min(DTIN, 17); ++pb) {
    if (pb > 1 && (uint32_t)pb * plane_bytes > cap) break;
    const int nb = (DTIN + pb - 1) / pb;
    const long cost = (long)nb * 300 + (long)(nb * pb - DTIN) * 40;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_pb = pb; }
  }
  g->pb = best_pb; g->nboxes = (DTIN + best_pb - 1) / best_pb;
  g->box_bytes = (uint32_t)best_pb * plane_bytes;
  g->stage_bytes = (g->box_bytes + 127u) & ~127u;
  int nst = (int)(avail / g->stage_bytes);
  // TMA is latency-bound: keep >= ~96 KB in flight per SM when the stages are small
  nst = std::min(nst, std::max(2, std::min(kConvMaxStages, (int)(98304u / g->stage_bytes))));
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

ConvPackJob conv_pack_job(const ConvGeom& g, long long src_off, long long dst_off, int transpose_flip, int co_off, int co_total) {
  ConvPackJob j;
  memset(&j, 0, sizeof(j));
  j.src_off = src_off; j.dst_off = dst_off;
  j.Cin_real = g.Cin_real; j.Cout_real = g.Cout_real; j.COUT = g.COUT; j.ksize = g.ksize;
  j.nkd = g.ksize == 3 ? 3 : 1; j.KC = g.KC; j.nchunks = g.nchunks; j.nsteps = g.nsteps;
  j.transpose_flip = transpose_flip; j.bf16 = g.bf16;
  j.co_off = co_off; j.co_total = co_total < 0 ? g.Cout_real : co_total;
  j.paired = g.paired ? 1 : 0;
  return j;
}

int conv_pack_weights_batch(const ConvPackJob* jobs, int njobs, const float* params, void* wimg, cudaStream_t st) {
  for (int j0 = 0; j0 < njobs; j0 += kConvPackBatch) {
    PackBatch b;
    const int nb = std::min(kConvPackBatch, njobs - j0);
    memset(&b, 0, sizeof(b));
    memcpy(b.job, jobs + j0, sizeof(ConvPackJob) * nb);
    conv_pack_batch_kernel<<<dim3(48, nb), 256, 0, st>>>(params, (uint8_t*)wimg, b);
    SEUNET_CUDA_CHECK(cudaGetLastError());
  }
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
                     const float* out_scale, int shallow_ok) {
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
  auto tiles_d = [&](int dt) { return a.dstep == 2 ? 2 * (((D + 1) / 2 + dt - 1) / dt) : (D + dt - 1) / dt; };
  // Shallow tiles: a CTA tile normally owns DT output planes (the whole TMEM), which leaves most SMs idle when the layer has
  // fewer tiles than SMs - a 16^3 level of one patch is FOUR tiles of 10 sequential input planes each (30 us whatever the
  // batch size up to 8).  With dt_use < DT a tile owns only its first dt_use planes: the schedule (boxes, accumulator groups)
  // is unchanged, the MMAs of the input planes behind plane dt_use + 1 are skipped and their accumulators never stored, so
  // a tile costs ~(dt_use + 2) planes of MMAs and there are DT/dt_use times as many tiles.  Per output voxel the accumulation
  // order is the same, so the result is bit-identical; only launches WITHOUT InstanceNorm statistics use it (data gradients):
  // the per-tile fp32 statistics partials of the forward convs would otherwise depend on the batch size.
  int dtu = DT;
  static const bool shallow_env = !(getenv("SEUNET_CONV_SHALLOW") && atoi(getenv("SEUNET_CONV_SHALLOW")) == 0);   // A/B knob
  if (shallow_ok && shallow_env && stats == nullptr) {
    const int extra = (nkd == 3 ? 2 : 0) + 3;   // halo planes + per-tile fixed cost (pipeline fill, weight ring) in plane units
    long best = -1;
    for (int dt = DT; dt >= 2; --dt) {
      const long tiles = (long)a.tilesW * a.tilesH * tiles_d(dt) * N;
      const long cost = ((tiles + num_sms - 1) / num_sms) * (dt + extra);
      if (best < 0 || cost < best) { best = cost; dtu = dt; }
    }
  }
  a.dt_use = dtu;
  a.tilesD = tiles_d(dtu);
  a.numTiles = a.tilesW * a.tilesH * a.tilesD * N;
  a.dil = nkd == 3 ? (a.dstep == 2 ? 1 : g.dil) : 0;
  a.nkd = nkd; a.halo = halo;
  a.nchunks = g.nchunks; a.kc8 = g.KC / 8;
  a.in_chunks_total = in_chunks_total; a.in_chunk_off = in_chunk_off;
  a.out_chunks_total = out_chunks_total; a.out_chunk_off = out_chunk_off;
  a.nstages = g.nstages; a.wslots = g.wslots; a.nsteps = g.nsteps;
  a.pb = g.pb; a.nboxes = g.nboxes; a.plane16 = (uint32_t)HV; a.cdil = g.dil;
  a.dbg = getenv("SEUNET_CONV_DBG") ? atoi(getenv("SEUNET_CONV_DBG")) : 0;   // developer experiments (profile build only)
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
  const uint32_t a_lbo = (uint32_t)g.pb * HV * 16u;   // the two K halves (8-channel planes) of a box are pb halo planes apart
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
  // a box spans pb planes of one parity class: extent pb * dstep along d, traversed with stride dstep
  cuuint32_t box[4] = {(cuuint32_t)(8 * lineW), (cuuint32_t)(kConvTileH + 2 * halo), (cuuint32_t)(g.pb * a.dstep), (cuuint32_t)(g.KC / 8)};
  cuuint32_t estr[4] = {1, 1, (cuuint32_t)a.dstep, 1};
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
    SEUNET_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));   // + <= 4 KB static (s_run) <= 227 KB
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  conv_tc_kernel<COUT><<<L.grid, kConvThreads, L.g.smem_bytes, st>>>(L.tmap, L.a);
  SEUNET_CUDA_CHECK(cudaGetLastError());
#ifdef SEUNET_CONV_PROFILE
  {
    static long long h[148 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_conv_prof, sizeof(h));
    double s[8] = {0};
    const int n = std::min(L.grid, 148);
    for (int b = 0; b < n; ++b) for (int k = 0; k < 8; ++k) s[k] += (double)h[b * 8 + k] / n;
    const double mmas = (double)L.a.numTiles / L.grid * L.a.nchunks * (kConvAccCols / COUT + 2 * L.a.dil) * L.a.nsteps;
    fprintf(stderr, "[conv prof] COUT %d Cin %d k%d pb %d nboxes %d stages %d tiles %d: producer %.0f kclk (wait empty %.0f) | issuer %.0f kclk "
            "(wait full %.0f, wait tmem %.0f, wait weights %.0f; %.1f clk per issued-plane MMA) | epilogue %.0f kclk (wait full %.0f)\n",
            COUT, L.g.Cin, L.g.ksize, L.g.pb, L.g.nboxes, L.g.nstages, L.a.numTiles, s[0] / 1e3, s[1] / 1e3, s[2] / 1e3, s[3] / 1e3,
            s[4] / 1e3, s[5] / 1e3, s[2] / std::max(1.0, mmas), s[6] / 1e3, s[7] / 1e3);
  }
#endif
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
