// Launchers for the HBM-bound kernels of the SE-UNet backward pass (autograd of SE_UNet.forward, reached from
// loss.backward() in train.py:246/300/439/490/602).  Gradient tensors w.r.t. activations are bf16 chunk planes
// (range of fp32, no loss scaling); dY - the operand of the tensor-core dgrad/wgrad - is written in the activation
// storage type, pre-scaled by a per-layer power of two.
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

// per-layer reduction block (doubles / floats zeroed at the start of every backward)
constexpr int kRedS = 0;          // double [N][64][2]  sum dn, sum dn*n          (y branch)
// float offsets inside the per-layer float block
constexpr int kRedMaxN = 64;      // max batch for the per-layer float block layout below
struct RedLayout {
  // doubles
  static constexpr int kS = 0;                     // [N][64][2]
  static constexpr int kSx = 1;                    // [N][64][2] (x branch of CAT blocks)
};

struct SseBwdArgs {
  // forward tensors
  const act_t* raw; int raw_chunks;
  const double* stats; int stats_c;
  long long V; int N;
  const float* wse; const float* wse2;
  const float* weff; const float* wcst;
  // incoming gradients
  const grad_t* dE0; int dE0_chunks; int dE0_off;   // may be null (dc6: e0 has no consumer)
  const float* dT;                                  // [n][V] gradient of the head accumulator of this level
  // outputs
  grad_t* dn; int dn_chunks;                        // [n][dn_chunks][V][8]
  double* redS;                                     // [N][64][2]
  float* dwse; float* dwse2;                        // [64] each (summed over n)
  float* dweff; float* dcst;                        // [N][64], [N]
  unsigned int* dymax;                              // max |dn * rstd| as float bits
};
int launch_sse_bwd_a(int C, const SseBwdArgs& a, cudaStream_t st);

struct NormBwdArgs {          // pass B: dy = rstd * (dn - mean(dn) - n * mean(dn*n)) * 2^s
  const act_t* raw; int raw_chunks;
  const double* stats; int stats_c;
  long long V; int N; int C;
  const grad_t* dn; int dn_chunks;
  const double* redS;
  const unsigned int* dymax;
  act_t* dy; int dy_chunks;
  float* scale_out;           // [2]: scale, 1/scale (written by block 0)
};
int launch_norm_bwd_b(const NormBwdArgs& a, cudaStream_t st);

struct CatBwdArgs {
  const act_t* raw; int raw_chunks;
  const double* stats; int stats_c;
  Dims d; int C;
  const float* x; long long xs[5]; XOffsets xo; int in_ch; const float* wx; const double* mom;   // injection branch or null
  const grad_t* g; int g_chunks; int g_off;          // gradient at the full-resolution destination slot
  const grad_t* gp; int gp_chunks; int gp_off;       // gradient of the 2x2x2 max-pooled copy (or null)
  grad_t* dn; int dn_chunks;
  double* redS; double* redSx;                       // [N][64][2]
  unsigned int* dymax;
};
int launch_cat_bwd_a(const CatBwdArgs& a, cudaStream_t st);

struct CatBwdXArgs {          // x-branch part of pass B: dWx[c][i] += sum du_c * x_i
  const act_t* raw; int raw_chunks;
  const double* stats; int stats_c;
  Dims d; int C;
  const float* x; long long xs[5]; XOffsets xo; int in_ch; const float* wx; const double* mom;
  const grad_t* dn; int dn_chunks;
  const double* redSx;
  float* dwx;                 // [C][in_ch] in the flat gradient buffer (zeroed before)
};
int launch_cat_bwd_x(const CatBwdXArgs& a, cudaStream_t st);

// adjoint of launch_upsample2: gsrc (C ch at sd) = Up2^T(gdst slice)
int launch_upsample2_bwd(const grad_t* gdst, int gdst_chunks, int gdst_off, int C, Dims sd, grad_t* gsrc, cudaStream_t st);

// separable version on fp32 gradient planes (three 1-D passes); tmp1 >= N*C*(2D*2H*W) floats, tmp2 >= N*C*(2D*H*W) floats
int launch_upsample2_bwd_sep(const grad_t* gdst, int gdst_chunks, int gdst_off, int C, Dims sd, grad_t* gsrc, float* tmp1, float* tmp2,
                             cudaStream_t st);

// adjoint of the head: dT_l = Up_{2^l}^T(dpred) for one level, and sum(dpred) for the bias
// three 1-D adjoint passes; tmp1 >= N*D*H*(W>>level) floats, tmp2 >= N*D*(H>>level)*(W>>level) floats
int launch_head_bwd_level_sep(const float* dpred, Dims full, int level, float* dT, float* tmp1, float* tmp2, cudaStream_t st);
int launch_sum(const float* src, long long n, float* dst /* single float, overwritten */, cudaStream_t st);

struct SmallGradBlock { int w2_off, b2_off, wse_off, wse2_off, C, head, k; };
struct SmallGradArgs {
  SmallGradBlock blk[18];
  int hw_off[2];
  int N;
};
// conv2 / conv_se / conv_se2 / dc0_x.weight gradients from the per-layer reductions
int launch_small_grads(const float* params, const float* drop0, const float* drop1, const SmallGradArgs& a,
                       const float* dweff /*[18][N][64]*/, const float* dcst /*[18][N]*/, const float* dwse /*[18][64]*/,
                       const float* dwse2 /*[18][64]*/, float* grads, cudaStream_t st);
