// Host-side description of one tcgen05 implicit-GEMM convolution launch (3x3x3 dilated or 1x1x1).
#pragma once
#include <cuda.h>
#include "common.cuh"

constexpr int kConvThreads = 192;   // warp 0: TMA producer, warp 1: UMMA issuer, warps 2-5: epilogue
constexpr int kConvMaxStages = 32;  // activation stage ring (mbarrier slots)
constexpr int kConvMaxSteps = 40;   // UMMA K=16 steps per (input plane, channel chunk)
constexpr int kConvTableSteps = 8;  // steps of the table-driven issue mode (Cin = 8 layers: 5)
constexpr int kConvTileH = 16;      // one UMMA M=128 block = 16 (h) x 8 (w) voxels of one d-plane
constexpr int kConvTileW = 8;
constexpr int kConvAccCols = 512;   // TMEM columns of the (single) accumulator stage: DT = 512/COUT output-plane slots
constexpr int kConvGroups = 4;      // the slots are handed between issuer and epilogue in this many groups

// Kernel arguments (passed by value as a __grid_constant__).
struct ConvKArgs {
  int N, D, H, W;
  int tilesW, tilesH, tilesD, numTiles;
  int dil;            // distance of the kd taps in planes of the tile's parity class (0 when nkd == 1)
  int dstep;          // 2 for dilation-2 layers: tiles hold planes of one parity (d = par + 2*i); else 1
  int nkd;            // 3: kd taps stacked along UMMA N; 1: pointwise conv
  int halo;           // in-plane halo (dil for 3x3x3, 0 for 1x1x1)
  int nchunks, kc8;   // channel chunks per tile, 8-channel planes per chunk
  int in_chunks_total, in_chunk_off;
  int out_chunks_total, out_chunk_off;
  int nstages, wslots, nsteps;
  int pb, nboxes;        // planes per TMA box (= ring stage) and boxes per (tile, channel chunk)
  int dt_use;            // output planes a tile owns (<= DT = 512/COUT; < DT: shallow tiles of small layers)
  int dbg;               // developer experiment flags (only read by -DSEUNET_CONV_PROFILE builds)
  int cdil;              // the conv's own dilation (1|2; tap offsets inside the halo tile), 0 for 1x1x1
  uint32_t plane16;      // one halo plane inside a box, in 16-byte units
  uint32_t stage_bytes, box_bytes, wchunk_bytes;
  uint32_t a_sbo, b_lbo;
  uint32_t fmt;          // UMMA operand format of activations AND weights: 0 = f16, 1 = bf16
  int out_bf16;          // epilogue writes the gradient format grad_t (fp32 by default) instead of the storage type
  int accum_out;         // epilogue adds to the existing output (gradient accumulation of fan-out nodes)
  int out_real_chunks;   // only the first out_real_chunks 8-channel planes are written (COUT padding)
  const float* out_scale; // optional DEVICE scalar multiplied into the result (undoes the dY pre-scaling in dgrad)
  const uint8_t* wimg;   // packed weights: [chunk][step][khalf][nkd*COUT rows][8]
  void* out;             // conv output, chunk-plane layout (16-bit elements)
  double* stats;         // [N][COUT][2] running (sum, sum of squares), fp64 atomics
  // UMMA issue schedule.  Descriptor low words are (start address >> 4) | (LBO >> 4) << 16; the high words are constant.
  uint32_t a_lo0;        // LBO field of the A descriptor (address part added per stage)
  uint32_t a_hi, b_hi;   // SBO | version
  uint32_t stage16;      // stage_bytes >> 4
  int regular;           // 1: steps follow the (kh, kw, 16-channel block) nest below; 0: use the dlt table
  int ntap;              // regular: 3 (3x3 in-plane taps) or 1 (pointwise)
  int jsteps;            // regular: K=16 steps per tap (KC / 16)
  uint32_t kh_step, kw_step, j_step;   // regular: A start-address increments (16-byte units)
  uint32_t b_step;       // B start-address increment per step (16-byte units)
  uint2 dlt[kConvTableSteps];   // table mode (paired taps): per-step {A low-word delta, B low-word delta}
};

// How the fp32 reference weights map into one UMMA K=16 step of the packed image.
struct PackStep {
  int8_t tap_a, tap_b;      // (kh*3+kw) of K half 0 / 1, -1 = zeros
  int16_t cbase_a, cbase_b; // channel offset inside the chunk of K half 0 / 1
};

struct ConvGeom {
  int Cin_real, Cout_real;  // reference tensor sizes
  int Cin, COUT;            // padded: Cin in {8, 16k}, COUT in {16,32,64}
  int ksize;                // 3 or 1
  int dil;                  // conv dilation (3x3x3 only)
  int KC, nchunks, nsteps, wslots, nstages;
  int pb, nboxes;           // planes per TMA box / boxes per tile
  bool paired;              // Cin == 8: two taps share one K=16 step
  int bf16;                 // operands (activations + packed weights) are bf16 (gradient operators)
  uint32_t stage_bytes, box_bytes, wchunk_bytes, smem_bytes;
  PackStep psteps[kConvMaxSteps];
  size_t wimg_bytes() const { return (size_t)nchunks * wchunk_bytes; }
};

// Fill geometry for a layer. Returns 0 on success.
int conv_geom_init(ConvGeom* g, int Cin_real, int Cout_real, int ksize, int dil, int bf16 = -1 /* -1: storage type */);

// Pack fp32 weights (Cout_real, Cin_real, k, k, k) into the UMMA image.  transpose_flip=1 builds
// the data-gradient operator (roles of Cin/Cout swapped, taps mirrored).
// co_off/co_total (transpose_flip only): the operator produces output channels [co_off, co_off+Cout_real) of a
// forward weight tensor with co_total input channels (dgrad of wide layers is split into <= 64-channel pieces).
int conv_pack_weights(const ConvGeom& g, const float* w_fp32, void* wimg, int transpose_flip, cudaStream_t st,
                      int co_off = 0, int co_total = -1);

// Batched packing (plans): every job describes one layer image; sources are offsets into ONE flat fp32 parameter buffer,
// destinations byte offsets into ONE packed-image buffer.  kConvPackBatch jobs per launch (kernel-parameter space).
constexpr int kConvPackBatch = 48;
struct ConvPackJob {
  long long src_off, dst_off;
  int Cin_real, Cout_real, COUT, ksize, nkd, KC, nchunks, nsteps, transpose_flip, bf16, co_off, co_total, paired, pad_;
};
ConvPackJob conv_pack_job(const ConvGeom& g, long long src_off, long long dst_off, int transpose_flip, int co_off = 0, int co_total = -1);
int conv_pack_weights_batch(const ConvPackJob* jobs, int njobs, const float* params, void* wimg, cudaStream_t st);

struct ConvLaunch {
  ConvGeom g;
  CUtensorMap tmap;
  ConvKArgs a;
  int grid;
};

// Bind a layer to concrete buffers: input chunk-plane buffer (in_chunks_total planes per sample,
// slice starting at in_chunk_off), output raw buffer, stats.
int conv_launch_init(ConvLaunch* L, const ConvGeom& g, int N, int D, int H, int W,
                     const void* in, int in_chunks_total, int in_chunk_off,
                     void* out, int out_chunks_total, int out_chunk_off,
                     double* stats, const void* wimg, int num_sms, int accum_out = 0, int out_real_chunks = -1,
                     int grad_out = 0 /* 1: output is grad_t */, const float* out_scale = nullptr,
                     int shallow_ok = 0 /* 1: tiles may own fewer than DT planes when the layer cannot fill the SMs */);
int conv_launch_run(const ConvLaunch& L, cudaStream_t st);
