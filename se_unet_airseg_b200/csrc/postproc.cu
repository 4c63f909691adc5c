// GPU post-processing of the sliding-window probability volume:
//   * double_threshold_iteration (prediction.py:13-37): hysteresis thresholding.  The reference's `while` runs exactly ONE
//     in-place raster sweep (gbin_pre aliases gbin), so the result depends on the (i, j, k) visiting order: a weak voxel is
//     set iff one of its 13 raster-PRECEDING 26-neighbours is set (already updated) or one of its 13 FOLLOWING neighbours
//     is strong (not yet updated).  That recurrence is evaluated exactly: rows (i, j) with equal 2i + j are independent
//     (wavefront), and inside a row the k-1 dependence is a bit-parallel prefix fill over the weak bits.
//   * border zeroing (prediction.py:112-115),
//   * maximum_3d (util.py:58-75): largest 26-connected component (second largest if the largest misses the three probe
//     slices), then scipy binary_fill_holes (6-connected background not reachable from the volume border).
//     Components are found by union-find over ROW RUNS (maximal runs of set bits along the last axis), not voxels: an
//     airway mask has a few 1e5 runs against 1e8 voxels.
// Volumes are bit-packed along the last axis: row r = i*H + j, word w holds voxels k = 32w .. 32w+31 (bit b = k - 32w).
#include "../../include/seunet_b200.h"
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int kInfoError = 0, kInfoRunsFg = 1, kInfoBest = 2, kInfoSecond = 3, kInfoUsedSecond = 4, kInfoRunsBg = 5,
              kInfoBestId = 6, kInfoSecondId = 7, kInfoFlagBest = 8, kInfoFlagSecond = 9, kInfoChosen = 10, kInfoTotal = 11;
constexpr int kInfoInts = 16;

struct PpLayout {
  size_t strong, weak, setm, outm, offsets, runs, parent, size, info, total;
};
PpLayout pp_layout(int D, int H, int W, int64_t max_runs) {
  const size_t R = (size_t)D * H, WW = (size_t)(W + 31) / 32;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  PpLayout L;
  size_t off = 0;
  L.strong = off; off += al(R * WW * 4);
  L.weak = off; off += al(R * WW * 4);
  L.setm = off; off += al(R * WW * 4);
  L.outm = off; off += al(R * WW * 4);
  L.offsets = off; off += al((R + 1) * 4);
  L.runs = off; off += al((size_t)max_runs * 8);
  L.parent = off; off += al(((size_t)max_runs + 1) * 4);
  L.size = off; off += al(((size_t)max_runs + 1) * 4);
  L.info = off; off += al(kInfoInts * 4);
  L.total = off;
  return L;
}

// ---------------------------------------------------------------------------------------------------------------------
// classification into strong / weak bit planes: one warp per row
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pp_classify_kernel(const float* __restrict__ prob, int R, int W, int WW, double h255, double l255,
                                                          uint32_t* __restrict__ strong, uint32_t* __restrict__ weak) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += gridDim.x * 8) {
    for (int w = 0; w < WW; ++w) {
      const int k = w * 32 + lane;
      bool s = false, wk = false;
      if (k < W) {
        const double v = (double)prob[(size_t)r * W + k] * 255.0;   // pred = np.array(pred*255, float64)
        s = v >= h255;
        wk = !s && v >= l255;
      }
      const uint32_t sb = __ballot_sync(0xffffffffu, s), wb = __ballot_sync(0xffffffffu, wk);
      if (lane == 0) { strong[(size_t)r * WW + w] = sb; weak[(size_t)r * WW + w] = wb; }
    }
  }
}

__global__ void __launch_bounds__(256) pp_pack_kernel(const unsigned char* __restrict__ mask, int R, int W, int WW, uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += gridDim.x * 8)
    for (int w = 0; w < WW; ++w) {
      const int k = w * 32 + lane;
      const uint32_t b = __ballot_sync(0xffffffffu, k < W && mask[(size_t)r * W + k] != 0);
      if (lane == 0) bits[(size_t)r * WW + w] = b;
    }
}

__global__ void __launch_bounds__(256) pp_unpack_kernel(const uint32_t* __restrict__ bits, int R, int W, int WW, unsigned char* __restrict__ mask) {
  const size_t n = (size_t)R * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / W;
    const int k = (int)(i - r * W);
    mask[i] = (bits[r * WW + (k >> 5)] >> (k & 31)) & 1u;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-word shifts inside a warp: lane = word index of the row (lanes >= WW hold zeros)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t row_shl(uint32_t x, int s, int lane) {   // towards higher k
  if (s < 32) {
    uint32_t lo = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) lo = 0;
    return (x << s) | (lo >> (32 - s));
  }
  const int d = s >> 5;
  uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
  if (lane < d) y = 0;
  return y;
}
__device__ __forceinline__ uint32_t row_shr1(uint32_t x, int lane) {        // towards lower k
  uint32_t hi = __shfl_down_sync(0xffffffffu, x, 1);
  if (lane == 31) hi = 0;
  return (x >> 1) | (hi << 31);
}

// one wavefront t = 2i + j of the in-place raster sweep; one warp per row
__global__ void __launch_bounds__(128) pp_dti_wavefront_kernel(const uint32_t* __restrict__ strong, const uint32_t* __restrict__ weak,
                                                               uint32_t* __restrict__ setm, int t, int i_lo, int n_rows, int D, int H,
                                                               int WW) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= n_rows) return;
  const int i = i_lo + q, j = t - 2 * i;
  auto ld = [&](const uint32_t* p, int ii, int jj) -> uint32_t {
    if (ii < 0 || ii >= D || jj < 0 || jj >= H || lane >= WW) return 0u;
    return p[((size_t)ii * H + jj) * WW + lane];
  };
  const uint32_t S = ld(strong, i, j), Wk = ld(weak, i, j);
  // preceding neighbour rows: already final; following neighbour rows: still the initial (strong) state
  const uint32_t nb = ld(setm, i - 1, j - 1) | ld(setm, i - 1, j) | ld(setm, i - 1, j + 1) | ld(setm, i, j - 1) |
                      ld(strong, i, j + 1) | ld(strong, i + 1, j - 1) | ld(strong, i + 1, j) | ld(strong, i + 1, j + 1);
  const uint32_t A = nb | row_shl(nb, 1, lane) | row_shr1(nb, lane) | row_shr1(S, lane);   // + the k+1 voxel of this row (strong only)
  uint32_t x = S | (Wk & A), p = Wk;
  // set(k) |= weak(k) & set(k-1): prefix fill through runs of weak bits (Kogge-Stone over up to 1024 bits)
#pragma unroll 1
  for (int s = 1; s < WW * 32; s <<= 1) {
    x |= p & row_shl(x, s, lane);
    p &= row_shl(p, s, lane);
  }
  if (lane < WW) setm[((size_t)i * H + j) * WW + lane] = x;
}

__global__ void __launch_bounds__(256) pp_border_kernel(uint32_t* __restrict__ bits, int D, int H, int WW, int dlo, int dhi, int hlo, int hhi) {
  const size_t n = (size_t)D * H * WW;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / WW;
    const int i = (int)(r / H), j = (int)(r - (size_t)i * H);
    if (i < dlo || i >= dhi || j < hlo || j >= hhi) bits[idx] = 0u;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// row runs
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t row_word(const uint32_t* bits, size_t r, int w, int W, int WW, int invert) {
  uint32_t x = bits[r * WW + w];
  if (invert) {
    x = ~x;
    if (w == WW - 1 && (W & 31)) x &= (1u << (W & 31)) - 1u;
  }
  return x;
}

__global__ void __launch_bounds__(256) pp_count_runs_kernel(const uint32_t* __restrict__ bits, int R, int W, int WW, int invert,
                                                            int* __restrict__ counts) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    int c = 0;
    uint32_t carry = 0;
    for (int w = 0; w < WW; ++w) {
      const uint32_t x = row_word(bits, r, w, W, WW, invert);
      c += __popc(x & ~((x << 1) | carry));
      carry = x >> 31;
    }
    counts[r] = c;
  }
}

// exclusive scan of counts[0..R) in place (-> offsets[0..R]), single block; total also goes to info[slot]
__global__ void __launch_bounds__(1024) pp_scan_kernel(int* __restrict__ a, int R, int* __restrict__ info, int slot, long long max_runs) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int start = 0; start < R; start += 1024) {
    const int idx = start + threadIdx.x;
    const int v = idx < R ? a[idx] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int wv = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, wv, o); if (lane >= o) wv += y; }
      s_warp[lane] = wv;
    }
    __syncthreads();
    const int base = s_base + (warp ? s_warp[warp - 1] : 0);
    if (idx < R) a[idx] = base + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_base = base + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a[R] = s_base;
    info[slot] = s_base;
    if ((long long)s_base > max_runs) info[kInfoError] = 1;
  }
}

__global__ void __launch_bounds__(256) pp_fill_runs_kernel(const uint32_t* __restrict__ bits, int R, int W, int WW, int invert,
                                                           const int* __restrict__ offsets, int2* __restrict__ runs,
                                                           const int* __restrict__ info) {
  if (info[kInfoError]) return;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    int o = offsets[r];
    int start = -1;
    for (int w = 0; w < WW; ++w) {
      uint32_t x = row_word(bits, r, w, W, WW, invert);
      int b = 0;   // bits below b are consumed
      while (b < 32) {
        const uint32_t rest = b ? (x >> b) : x;
        if (start < 0) {
          if (!rest) break;
          b += __ffs(rest) - 1;
          start = w * 32 + b;
        } else {
          const uint32_t inv = ~rest;   // (the shifted-in zeros read as ones here, handled by the b + z >= 32 test)
          if (!inv) break;              // b == 0 and the whole word is set
          const int z = __ffs(inv) - 1;
          if (b + z >= 32) break;       // run continues into the next word
          b += z;
          runs[o++] = make_int2(start, w * 32 + b - 1);
          start = -1;
        }
      }
    }
    if (start >= 0) runs[o++] = make_int2(start, W - 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// union-find over runs; node id = run index + 1, node 0 = "outside the volume" (background pass only)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* parent, int x) {
  while (true) {
    const int y = ((volatile int*)parent)[x];
    if (y == x) return x;
    x = y;
  }
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }   // the smaller id becomes the root
    if (atomicCAS(parent + a, a, b) == a) return;
  }
}

__global__ void __launch_bounds__(256) pp_init_kernel(int* __restrict__ parent, unsigned int* __restrict__ size, const int* __restrict__ info,
                                                      int slot) {
  if (info[kInfoError]) return;
  const int n = info[slot] + 1;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) { parent[x] = x; size[x] = 0u; }
}

// conn26 = 1: rows (i-1, j-1..j+1) and (i, j-1), runs overlap when dilated by one voxel; conn26 = 0: rows (i-1, j), (i, j-1),
// plain overlap, and runs touching the volume border are united with node 0.
__global__ void __launch_bounds__(128) pp_union_kernel(const int* __restrict__ offsets, const int2* __restrict__ runs, int* __restrict__ parent,
                                                       int D, int H, int W, int conn26, const int* __restrict__ info) {
  if (info[kInfoError]) return;
  const int R = D * H;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    const int o0 = offsets[r], o1 = offsets[r + 1];
    if (o0 == o1) continue;
    const int i = r / H, j = r - i * H;
    if (!conn26) {
      const bool row_border = i == 0 || i == D - 1 || j == 0 || j == H - 1;
      for (int a = o0; a < o1; ++a)
        if (row_border || runs[a].x == 0 || runs[a].y == W - 1) uf_union(parent, a + 1, 0);
    }
    const int dil = conn26 ? 1 : 0;
    const int nn = conn26 ? 4 : 2;
    for (int q = 0; q < nn; ++q) {
      int ii, jj;
      if (conn26) { ii = q < 3 ? i - 1 : i; jj = q < 3 ? j - 1 + q : j - 1; }
      else { ii = q == 0 ? i - 1 : i; jj = q == 0 ? j : j - 1; }
      if (ii < 0 || jj < 0 || jj >= H) continue;
      const int r2 = ii * H + jj;
      int b = offsets[r2];
      const int b1 = offsets[r2 + 1];
      for (int a = o0; a < o1 && b < b1; ++a) {
        const int2 ra = runs[a];
        while (b < b1 && runs[b].y < ra.x - dil) ++b;            // runs entirely before ra
        for (int c = b; c < b1 && runs[c].x <= ra.y + dil; ++c)   // overlapping (dilated) runs
          uf_union(parent, a + 1, c + 1);
      }
    }
  }
}

__global__ void __launch_bounds__(256) pp_compress_kernel(int* __restrict__ parent, const int* __restrict__ info, int slot) {
  if (info[kInfoError]) return;
  const int n = info[slot] + 1;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) parent[x] = uf_find(parent, x);
}

__global__ void __launch_bounds__(256) pp_sizes_kernel(const int2* __restrict__ runs, const int* __restrict__ parent, unsigned int* __restrict__ size,
                                                       const int* __restrict__ info, int slot) {
  if (info[kInfoError]) return;
  const int n = info[slot];
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x)
    atomicAdd(size + parent[a + 1], (unsigned int)(runs[a].y - runs[a].x + 1));
}

// largest and second largest component; ties go to the component that starts LATER in raster order (util.py:62-63: stable
// ascending sort by area, reversed - with labels numbered in raster order of their first voxel)
__global__ void __launch_bounds__(1024) pp_top2_kernel(const int* __restrict__ parent, const unsigned int* __restrict__ size, int* __restrict__ info) {
  if (info[kInfoError]) return;
  __shared__ unsigned long long s_key[1024];
  const int n = info[kInfoRunsFg];
  unsigned long long best_key = 0ull;
  for (int pass = 0; pass < 2; ++pass) {
    unsigned long long k = 0ull;
    for (int x = 1 + threadIdx.x; x <= n; x += blockDim.x)
      if (parent[x] == x) {
        const unsigned long long key = ((unsigned long long)size[x] << 32) | (unsigned int)x;
        if (pass == 1 && key == best_key) continue;
        k = key > k ? key : k;
      }
    s_key[threadIdx.x] = k;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
      if (threadIdx.x < o && s_key[threadIdx.x + o] > s_key[threadIdx.x]) s_key[threadIdx.x] = s_key[threadIdx.x + o];
      __syncthreads();
    }
    const unsigned long long top = s_key[0];
    __syncthreads();
    if (pass == 0) {
      best_key = top;
      if (threadIdx.x == 0) { info[kInfoBest] = (int)(top >> 32); info[kInfoBestId] = (int)(top & 0xffffffffu); }
    } else if (threadIdx.x == 0) {
      info[kInfoSecond] = (int)(top >> 32); info[kInfoSecondId] = (int)(top & 0xffffffffu);
    }
  }
}

// does the component touch the probe slices k = W/2, W/3, (W/3)*2 (util.py:65-70)?
__global__ void __launch_bounds__(256) pp_probe_kernel(const int2* __restrict__ runs, const int* __restrict__ parent, int W, int* __restrict__ info) {
  if (info[kInfoError]) return;
  const int n = info[kInfoRunsFg], best = info[kInfoBestId], second = info[kInfoSecondId];
  const int k0 = W / 2, k1 = W / 3, k2 = W / 3 * 2;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
    const int root = parent[a + 1];
    if (root != best && root != second) continue;
    const int2 r = runs[a];
    if ((r.x <= k0 && k0 <= r.y) || (r.x <= k1 && k1 <= r.y) || (r.x <= k2 && k2 <= r.y))
      atomicOr(info + (root == best ? kInfoFlagBest : kInfoFlagSecond), 1);
  }
}

__global__ void pp_choose_kernel(int* __restrict__ info) {
  if (info[kInfoError]) return;
  const bool use_second = !info[kInfoFlagBest] && info[kInfoSecondId] > 0;
  info[kInfoUsedSecond] = use_second ? 1 : 0;
  info[kInfoChosen] = use_second ? info[kInfoSecondId] : info[kInfoBestId];
}

// mode 0: out row = bits of the runs whose root is the chosen component; mode 1 (background runs): OR in the runs that are
// NOT connected to the outside node (holes)
__global__ void __launch_bounds__(256) pp_paint_kernel(const int* __restrict__ offsets, const int2* __restrict__ runs, const int* __restrict__ parent,
                                                       int R, int WW, int mode, uint32_t* __restrict__ out, const int* __restrict__ info) {
  if (info[kInfoError]) return;
  const int chosen = info[kInfoChosen];
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    uint32_t* row = out + (size_t)r * WW;
    if (mode == 0) for (int w = 0; w < WW; ++w) row[w] = 0u;
    for (int a = offsets[r]; a < offsets[r + 1]; ++a) {
      const int root = parent[a + 1];
      if (mode == 0 ? (root != chosen || chosen == 0) : (root == 0)) continue;
      const int2 run = runs[a];
      for (int w = run.x >> 5; w <= (run.y >> 5); ++w) {
        const int lo = max(run.x - w * 32, 0), hi = min(run.y - w * 32, 31);
        const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        row[w] |= m;
      }
    }
  }
}

int check_dims(int D, int H, int W, int64_t max_runs) {
  if (D < 1 || H < 1 || W < 1 || W > 1024) { seunet_set_error("postproc: dims %dx%dx%d unsupported (W <= 1024)", D, H, W); return 1; }
  if ((int64_t)D * H >= (1ll << 31) - 2 || max_runs < 1 || max_runs >= (1ll << 31) - 2) { seunet_set_error("postproc: volume too large"); return 1; }
  return 0;
}

}  // namespace

extern "C" size_t seunet_postproc_scratch_bytes(int D, int H, int W, int64_t max_runs) {
  if (check_dims(D, H, W, max_runs)) return 0;
  return pp_layout(D, H, W, max_runs).total;
}

extern "C" int seunet_postproc_dti(const float* prob, int D, int H, int W, double h_thresh, double l_thresh, double border_frac,
                                   unsigned char* mask_out, void* scratch, int64_t max_runs, seunet_stream_t stream) {
  if (check_dims(D, H, W, max_runs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const PpLayout L = pp_layout(D, H, W, max_runs);
  uint8_t* base = (uint8_t*)scratch;
  uint32_t* strong = (uint32_t*)(base + L.strong);
  uint32_t* weak = (uint32_t*)(base + L.weak);
  uint32_t* setm = (uint32_t*)(base + L.setm);
  const int R = D * H, WW = (W + 31) / 32;
  const int rb = std::min((R + 7) / 8, 148 * 8);
  pp_classify_kernel<<<rb, 256, 0, st>>>(prob, R, W, WW, h_thresh * 255.0, l_thresh * 255.0, strong, weak);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  // wavefronts t = 2i + j
  for (int t = 0; t <= 2 * (D - 1) + (H - 1); ++t) {
    const int i_lo = std::max(0, (t - (H - 1) + 1) / 2), i_hi = std::min(D - 1, t / 2);
    if (i_hi < i_lo) continue;
    const int n_rows = i_hi - i_lo + 1;
    pp_dti_wavefront_kernel<<<(n_rows + 3) / 4, 128, 0, st>>>(strong, weak, setm, t, i_lo, n_rows, D, H, WW);
  }
  SEUNET_CUDA_CHECK(cudaGetLastError());
  if (border_frac >= 0.0) {   // prediction.py:112-115: int(0.15 * n), int(0.85 * n) in double arithmetic
    const int dlo = (int)(border_frac * D), dhi = (int)((1.0 - border_frac) * D);
    const int hlo = (int)(border_frac * H), hhi = (int)((1.0 - border_frac) * H);
    pp_border_kernel<<<148 * 4, 256, 0, st>>>(setm, D, H, WW, dlo, dhi, hlo, hhi);
    SEUNET_CUDA_CHECK(cudaGetLastError());
  }
  if (mask_out) {
    pp_unpack_kernel<<<148 * 8, 256, 0, st>>>(setm, R, W, WW, mask_out);
    SEUNET_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

extern "C" int seunet_postproc_largest_component(const unsigned char* mask_in, int D, int H, int W, int fill_holes,
                                                 unsigned char* mask_out, int* info_out, void* scratch, int64_t max_runs,
                                                 seunet_stream_t stream) {
  if (check_dims(D, H, W, max_runs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const PpLayout L = pp_layout(D, H, W, max_runs);
  uint8_t* base = (uint8_t*)scratch;
  uint32_t* setm = (uint32_t*)(base + L.setm);
  uint32_t* outm = (uint32_t*)(base + L.outm);
  int* offsets = (int*)(base + L.offsets);
  int2* runs = (int2*)(base + L.runs);
  int* parent = (int*)(base + L.parent);
  unsigned int* size = (unsigned int*)(base + L.size);
  int* info = (int*)(base + L.info);
  const int R = D * H, WW = (W + 31) / 32;
  const int rowb = std::min((R + 255) / 256, 148 * 8);
  SEUNET_CUDA_CHECK(cudaMemsetAsync(info, 0, kInfoInts * sizeof(int), st));
  if (mask_in) {
    pp_pack_kernel<<<std::min((R + 7) / 8, 148 * 8), 256, 0, st>>>(mask_in, R, W, WW, setm);
    SEUNET_CUDA_CHECK(cudaGetLastError());
  }
  // foreground components, 26-connectivity (cc3d.connected_components(..., connectivity=26), util.py:59)
  pp_count_runs_kernel<<<rowb, 256, 0, st>>>(setm, R, W, WW, 0, offsets);
  pp_scan_kernel<<<1, 1024, 0, st>>>(offsets, R, info, kInfoRunsFg, max_runs);
  pp_fill_runs_kernel<<<rowb, 256, 0, st>>>(setm, R, W, WW, 0, offsets, runs, info);
  pp_init_kernel<<<148 * 4, 256, 0, st>>>(parent, size, info, kInfoRunsFg);
  pp_union_kernel<<<std::min((R + 127) / 128, 148 * 16), 128, 0, st>>>(offsets, runs, parent, D, H, W, 1, info);
  pp_compress_kernel<<<148 * 4, 256, 0, st>>>(parent, info, kInfoRunsFg);
  pp_sizes_kernel<<<148 * 4, 256, 0, st>>>(runs, parent, size, info, kInfoRunsFg);
  pp_top2_kernel<<<1, 1024, 0, st>>>(parent, size, info);
  pp_probe_kernel<<<148 * 4, 256, 0, st>>>(runs, parent, W, info);
  pp_choose_kernel<<<1, 1, 0, st>>>(info);
  pp_paint_kernel<<<rowb, 256, 0, st>>>(offsets, runs, parent, R, WW, 0, outm, info);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  if (fill_holes) {
    // scipy.ndimage.binary_fill_holes (util.py:73): background not 6-connected to the outside of the volume
    pp_count_runs_kernel<<<rowb, 256, 0, st>>>(outm, R, W, WW, 1, offsets);
    pp_scan_kernel<<<1, 1024, 0, st>>>(offsets, R, info, kInfoRunsBg, max_runs);
    pp_fill_runs_kernel<<<rowb, 256, 0, st>>>(outm, R, W, WW, 1, offsets, runs, info);
    pp_init_kernel<<<148 * 4, 256, 0, st>>>(parent, size, info, kInfoRunsBg);
    pp_union_kernel<<<std::min((R + 127) / 128, 148 * 16), 128, 0, st>>>(offsets, runs, parent, D, H, W, 0, info);
    pp_compress_kernel<<<148 * 4, 256, 0, st>>>(parent, info, kInfoRunsBg);
    pp_paint_kernel<<<rowb, 256, 0, st>>>(offsets, runs, parent, R, WW, 1, outm, info);
    SEUNET_CUDA_CHECK(cudaGetLastError());
  }
  pp_unpack_kernel<<<148 * 8, 256, 0, st>>>(outm, R, W, WW, mask_out);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  if (info_out) SEUNET_CUDA_CHECK(cudaMemcpyAsync(info_out, info, kInfoInts * sizeof(int), cudaMemcpyDeviceToDevice, st));
  return 0;
}
