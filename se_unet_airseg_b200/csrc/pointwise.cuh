// Launchers for the HBM-bound kernels of the SE-UNet forward pass (all chunk-plane layout).
#pragma once
#include "common.cuh"
#include <algorithm>

constexpr int kMaxInCh = 2;     // reference callers use in_channel=2 (default ctor: 1)
constexpr int kMomStride = 8;   // doubles per (level, sample): sum x_i (2), sum x_i x_j (3), pad

struct Dims { int N, D, H, W; };
// Per-sample element offsets of the network input inside a larger tensor (sliding windows of one CT
// volume, prediction.py:102).  use == 0: sample n starts at n * (sample stride).
constexpr int kMaxWindowBatch = 32;
struct XOffsets { int use; long long off[kMaxWindowBatch]; };
__host__ __device__ inline long long dims_vox(const Dims& d) { return (long long)d.D * d.H * d.W; }

// x (fp32, arbitrary strides) -> XB (storage type, 1 chunk plane, channels >= in_ch zero),
// max-pooled fp32 copies at 1/2 and 1/4 resolution, and first/second moments at all three levels.
int launch_input_prep(const float* x, const long long* xstride /*n,c,d,h,w in elements*/, const XOffsets& xo, int in_ch, Dims d,
                      act_t* xb, float* xp1, float* xp2, double* mom /*[3][N][kMomStride]*/, cudaStream_t st,
                      bool inference = false /* inference plans may use the vector kernel (different moment summation order) */);

struct SseArgs {
  const act_t* raw; int raw_chunks;        // conv output [n][raw_chunks][V][8]
  const double* stats; int stats_c;        // [n][stats_c][2]
  long long V;
  const float* wse; const float* wse2;     // gate weights (C floats each), wse2 may be null (1 gate)
  const float* weff; const float* wcst;    // folded side-branch/head weights [n][C], [n]
  float* T; int t_init;                    // head accumulator [n][V]; null = this block's head is not wanted (window plans: head 0)
  act_t* dest; int dest_chunks; int dest_off;  // gated activations, may be null
  int inference;                           // inference plan: the two-lanes-per-voxel variant may be used (training plans keep the
                                           // summation order the gradient parity bounds were measured with)
};
int launch_apply_sse(int C, int N, const SseArgs& a, cudaStream_t st);

// Fused SSE apply + CAT 1x1x1 conv (pointwise3.cu; inference plans): the block's own C output channels are concat chunks
// [0, C/8); chunks [C/8, cat_real_chunks) are read from the concat buffer; the conv output goes to `out` with statistics.
struct CatFuseArgs {
  const act_t* cat; int cat_chunks; int cat_real_chunks;   // concat buffer [n][cat_chunks][V][8]; chunks beyond cat_real_chunks are padding
  const float* w; int cin_real;                            // CATConv weight fp32 [nout][cin_real]
  int kcat, nout;                                          // padded K (multiple of 16) and output channels
  act_t* out; int out_chunks;                              // raw conv output [n][out_chunks][V][8]
  double* out_stats; int out_stats_c;                      // [n][out_stats_c][2], accumulated with atomics (zeroed by the caller)
};
int launch_apply_sse_cat(int C, int N, const SseArgs& a, const CatFuseArgs& f, int num_sms, cudaStream_t st);

struct CatArgs {
  const act_t* raw; int raw_chunks;
  const double* stats; int stats_c;
  Dims d;                                   // resolution of this block
  // optional injection branch lrelu(IN(Wx * x)) (SE_UNet.py:187,196,205)
  const float* x; long long xs[5]; XOffsets xo; int in_ch; const float* wx; const double* mom;
  act_t* dest; int dest_chunks; int dest_off;       // full-resolution destination (may be null)
  act_t* pdest; int pdest_chunks; int pdest_off;    // 2x2x2 max-pooled destination (may be null)
};
int launch_apply_cat(int C, const CatArgs& a, cudaStream_t st);

// trilinear x2, align_corners=True (SE_UNet.py:136-138), C channels src -> chunk slot of dst
int launch_upsample2(const act_t* src, int C, Dims sd, act_t* dst, int dst_chunks, int dst_off, cudaStream_t st);

struct HeadwBlock { int w2_off, b2_off, C, head, k; };  // offsets into the flat fp32 parameter buffer
struct HeadwArgs {
  HeadwBlock blk[18];
  int hw_off[2];      // dc0_0.weight / dc0_1.weight offsets
  int nblk;
};
// weff[blk][n][64], wcst[blk][n]
int launch_headw(const float* params, const float* drop0, const float* drop1, int N, const HeadwArgs& a,
                 float* weff, float* wcst, cudaStream_t st);

struct HeadArgs {
  Dims d;                         // full resolution
  const float* T0[4];             // head 0 accumulators at S, S/2, S/4, S/8
  const float* T1[3];             // head 1 accumulators at S, S/2, S/4
  const float* bias0; const float* bias1;
  float* pred0; float* pred1;     // [n][1][D][H][W] contiguous fp32
  // window sink (seunet_forward_window): head 0 is skipped and sigmoid(pred1) of sample n is added, in fixed point, to the
  // (X,Y,Z) accumulator volume at origin s[n] (prediction.py:103-106) instead of being stored
  unsigned int* acc; int X, Y, Z; float acc_scale; int s[kMaxWindowBatch][3];
};
int launch_head(const HeadArgs& a, cudaStream_t st);
