// tcgen05 weight-gradient launcher (see wgrad_tc.cu).
#pragma once
#include <cuda.h>
#include "common.cuh"

struct WgradKArgs {
  int N, D, H, W;
  int tilesW, tilesH, numTilePlanes;
  int dil, ksize, lineW;
  int tw;              // tile width in voxels (8, 16 or 32): one stage = 16 x tw voxels = tw K=16 MMA steps
  int ctas_per_pass, nstages;
  int x_chunks_total, x_chunk_off, dy_chunks_total, dy_chunk_off;
  uint32_t x_plane_bytes, x_box_bytes, x_stage_bytes, dy_box_bytes, dy_stage_bytes, bar_off;
  uint32_t fmt_a, fmt_b;
  int m64;             // UMMA M=64 (stacked input rows <= 64): halves the shared-memory traffic of the A operand
  int nkh, npkh;       // kh taps stacked along M per pass (h-shifted copies of the X box), passes per kd = ceil(3 / nkh)
  int x_rows;          // accumulator rows per kh copy (chunk planes of the X box * 8)
  int nseg;            // v3: segments (runs of planes) per column
  int v3, numCols;     // version 3 (kd taps merged into one walk along d): tile columns (sample, h tile, w tile, parity class)
  float* partial;
};

struct WgradLaunch {
  CUtensorMap tmap_x, tmap_dy;
  WgradKArgs a;
  int grid, Cin, Cout, COUT, ksize;
  uint32_t smem_bytes;
};

size_t wgrad_partial_bytes(int Cin, int Cout, int ksize, int num_sms);
// x: activations [N][x_chunks_total][D][H][W][8] (slice at x_chunk_off, Cin channels), dy: same 16-bit type [N][dy_chunks_total][..][8]
int wgrad_launch_init(WgradLaunch* L, int N, int D, int H, int W, int Cin, int Cout, int ksize, int dil,
                      const void* x, int x_chunks_total, int x_chunk_off, int x_bf16,
                      const void* dy, int dy_chunks_total, int dy_chunk_off, float* partial, int num_sms,
                      bool geometry_only = false);
// runs the kernel and reduces the partials into dw (fp32, (Cout, Cin, k, k, k) layout, overwritten)
// inv_scale: optional DEVICE scalar multiplied into dw (undoes the power-of-two pre-scaling of dY)
int wgrad_launch_run(const WgradLaunch& L, float* dw, const float* inv_scale, cudaStream_t st);
