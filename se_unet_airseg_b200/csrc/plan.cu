// Network plan + C ABI: the SE_UNet graph (SE_UNet.py:100-153, 181-238) as a fixed schedule of
// sm_100a kernel launches over caller-owned memory.
#include "../../include/seunet_b200.h"
#include "conv_tc.cuh"
#include "pointwise.cuh"
#include "wgrad_tc.cuh"
#include "backward.cuh"
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void seunet_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// network description (channel counts / order from SE_UNet.py:108-151)
// ---------------------------------------------------------------------------------------------
enum BufId {
  B_XB, B_CAT1, B_DC5IN, B_D2,                 // level 0
  B_P1, B_CAT2, B_DC3IN, B_DC42IN, B_D1F,      // level 1
  B_P2, B_CAT3, B_DC1IN, B_DC22IN, B_D0F,      // level 2
  B_P3, B_CAT4, B_E7F,                         // level 3
  B_COUNT, B_NONE = -1
};
struct BufDesc { int level, chunks; };
static const BufDesc kBufs[B_COUNT] = {
    {0, 1}, {0, 8}, {0, 8}, {0, 4},
    {1, 4}, {1, 16}, {1, 16}, {1, 12}, {1, 4},
    {2, 8}, {2, 24}, {2, 16}, {2, 16}, {2, 8},
    {3, 8}, {3, 24}, {3, 8}};

struct SseDesc {
  const char* name; int cin, cout, dil, gates, level;
  int in_buf, in_off;      // input slice (chunks)
  int out_buf, out_off;    // destination of the gated activations (chunks), B_NONE = not needed
  int head, k;             // deep-supervision head (0: dc0_0, 1: dc0_1) and index of the 2-ch pair
};
// in_ch-dependent entries (ec1 cin) are patched at plan creation.
static const SseDesc kSse[18] = {
    {"ec1", 0, 8, 1, 1, 0, B_XB, 0, B_CAT1, 4, 0, 0},
    {"ec2", 8, 16, 1, 1, 0, B_CAT1, 4, B_CAT1, 5, 0, 1},
    {"ec3", 16, 32, 2, 1, 0, B_CAT1, 5, B_CAT1, 0, 0, 2},
    {"ec4", 32, 32, 1, 2, 1, B_P1, 0, B_CAT2, 8, 0, 3},
    {"ec5", 32, 32, 2, 2, 1, B_CAT2, 8, B_CAT2, 12, 0, 4},
    {"ec6", 32, 64, 2, 2, 1, B_CAT2, 12, B_CAT2, 0, 0, 5},
    {"ec7", 64, 64, 1, 2, 2, B_P2, 0, B_CAT3, 8, 0, 6},
    {"ec8", 64, 64, 2, 2, 2, B_CAT3, 8, B_CAT3, 16, 0, 7},
    {"ec9", 64, 64, 2, 2, 2, B_CAT3, 16, B_CAT3, 0, 0, 8},
    {"ec10", 64, 64, 1, 2, 3, B_P3, 0, B_CAT4, 8, 0, 9},
    {"ec11", 64, 64, 1, 2, 3, B_CAT4, 8, B_CAT4, 16, 0, 10},
    {"ec12", 64, 64, 1, 2, 3, B_CAT4, 16, B_CAT4, 0, 0, 11},
    {"dc1", 128, 64, 1, 2, 2, B_DC1IN, 0, B_DC22IN, 8, 1, 0},
    {"dc2", 64, 64, 1, 2, 2, B_DC22IN, 8, B_DC22IN, 0, 1, 1},
    {"dc3", 128, 64, 1, 2, 1, B_DC3IN, 0, B_DC42IN, 4, 1, 2},
    {"dc4", 64, 32, 1, 2, 1, B_DC42IN, 4, B_DC42IN, 0, 1, 3},
    {"dc5", 64, 32, 1, 1, 0, B_DC5IN, 0, B_D2, 0, 1, 4},
    {"dc6", 32, 16, 1, 1, 0, B_D2, 0, B_NONE, 0, 1, 5},
};
enum SseId { S_EC1, S_EC2, S_EC3, S_EC4, S_EC5, S_EC6, S_EC7, S_EC8, S_EC9, S_EC10, S_EC11, S_EC12,
             S_DC1, S_DC2, S_DC3, S_DC4, S_DC5, S_DC6 };

struct CatDesc {
  const char* name; int cin, cout, level;
  int in_buf;               // whole concat buffer
  const char* xname;        // injection branch (x33/x63/x93) or null
  int out_buf, out_off;     // full-resolution destination
  int pool_buf;             // pooled destination or B_NONE
};
static const CatDesc kCat[6] = {
    {"ec33", 56, 32, 0, B_CAT1, "x33", B_DC5IN, 4, B_P1},
    {"ec63", 128, 64, 1, B_CAT2, "x63", B_DC3IN, 8, B_P2},
    {"ec93", 192, 64, 2, B_CAT3, "x93", B_DC1IN, 8, B_P3},
    {"ec123", 192, 64, 3, B_CAT4, nullptr, B_E7F, 0, B_NONE},
    {"dc22", 128, 64, 2, B_DC22IN, nullptr, B_D0F, 0, B_NONE},
    {"dc42", 96, 32, 1, B_DC42IN, nullptr, B_D1F, 0, B_NONE},
};
enum CatId { C_EC33, C_EC63, C_EC93, C_EC123, C_DC22, C_DC42 };
// Inference plans CAN fuse every CAT 1x1x1 conv into the apply pass of the LAST block that writes its concat (pointwise3.cu):
// kSseFuse[i] = CAT block whose concat starts with SSE block i's output (chunk offset 0 of kCat[..].in_buf), or -1.
// Only the 32-channel blocks (ec3 at full, dc4 at half resolution) are fused: the 64-channel variants of the pass need the whole
// register file of an SM for one block and measured no faster than apply + tcgen05 conv (ec6 33.4 vs 15.1 + 18.0 us per window).
static const int kSseFuse[18] = {-1, -1, C_EC33, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, C_DC42, -1, -1};

// ---------------------------------------------------------------------------------------------
// flat parameter table (state_dict order)
// ---------------------------------------------------------------------------------------------
struct ParamEntry { std::string name; int64_t off, numel; };
struct ParamTable {
  std::vector<ParamEntry> e;
  int64_t total = 0;
  void add(const std::string& n, int64_t numel) { e.push_back({n, total, numel}); total += numel; }
  int64_t off(const std::string& n) const {
    for (auto& p : e) if (p.name == n) return p.off;
    return -1;
  }
};
static void add_sse(ParamTable& t, const char* n, int cin, int cout, int gates) {
  std::string s(n);
  t.add(s + ".conv1.weight", (int64_t)cout * cin * 27);
  t.add(s + ".conv1.bias", cout);
  t.add(s + ".conv2.weight", 2 * cout);
  t.add(s + ".conv2.bias", 2);
  t.add(s + ".conv_se.weight", cout);
  if (gates == 2) t.add(s + ".conv_se2.weight", cout);
}
static void add_cat(ParamTable& t, const char* n, int cin, int cout) { t.add(std::string(n) + ".conv1.weight", (int64_t)cout * cin); }
static ParamTable build_params(int ic, int ncls) {
  ParamTable t;
  add_sse(t, "ec1", ic, 8, 1); add_sse(t, "ec2", 8, 16, 1); add_sse(t, "ec3", 16, 32, 1);
  add_cat(t, "ec33", 56, 32); add_cat(t, "x33", ic, 32);
  add_sse(t, "ec4", 32, 32, 2); add_sse(t, "ec5", 32, 32, 2); add_sse(t, "ec6", 32, 64, 2);
  add_cat(t, "ec63", 128, 64); add_cat(t, "x63", ic, 64);
  add_sse(t, "ec7", 64, 64, 2); add_sse(t, "ec8", 64, 64, 2); add_sse(t, "ec9", 64, 64, 2);
  add_cat(t, "ec93", 192, 64); add_cat(t, "x93", ic, 64);
  add_sse(t, "ec10", 64, 64, 2); add_sse(t, "ec11", 64, 64, 2); add_sse(t, "ec12", 64, 64, 2);
  add_cat(t, "ec123", 192, 64);
  add_sse(t, "dc1", 128, 64, 2); add_sse(t, "dc2", 64, 64, 2); add_cat(t, "dc22", 128, 64);
  add_sse(t, "dc3", 128, 64, 2); add_sse(t, "dc4", 64, 32, 2); add_cat(t, "dc42", 96, 32);
  add_sse(t, "dc5", 64, 32, 1); add_sse(t, "dc6", 32, 16, 1); add_cat(t, "dc62", 48, 16);
  t.add("dc0_0.weight", 24 * ncls); t.add("dc0_0.bias", ncls);
  t.add("dc0_1.weight", 12 * ncls); t.add("dc0_1.bias", ncls);
  return t;
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct ConvSlot {
  ConvGeom g;
  ConvLaunch L;
  int64_t w_off;        // fp32 weights in the flat parameter buffer
  size_t wimg_off;      // packed image offset
  size_t raw_off;       // raw conv output in the workspace
  size_t stats_off;     // fp64 stats in the workspace
};

struct seunet_plan {
  int N, D, H, W, in_ch, ncls, mode, device, num_sms;
  // inference plans: CAT 1x1x1 convs fused into the producer's apply pass (pointwise3.cu); SEUNET_CAT_FUSION=0 disables it
  bool fuse_cat = true;
  bool skip_head0 = false;          // set for the duration of a seunet_forward_window call
  ParamTable pt;
  ConvSlot sse_conv[18], cat_conv[6];
  size_t buf_off[B_COUNT];
  size_t T0_off[4], T1_off[3];
  size_t xp1_off, xp2_off, mom_off, weff_off, wcst_off, stats_off, stats_bytes;
  size_t ws_bytes, wimg_bytes;
  HeadwArgs headw;
  // ---- backward (training-mode plans only; see plan_bwd.inc)
  struct DgradPiece { ConvGeom g; ConvLaunch L; size_t wimg_off; int co_off, co_total, out_buf, out_chunk, accum; };
  std::vector<DgradPiece> sse_dgrad[18], cat_dgrad[6];
  WgradLaunch sse_wgrad[18], cat_wgrad[6];
  SmallGradArgs smallg;
  size_t gbuf_off[B_COUNT];          // bf16 gradients w.r.t. the activation buffers
  size_t dn_off[4], dy_off[4];       // per-level dn (bf16) / dY (storage type) scratch
  size_t dT0_off[4], dT1_off[3];     // head-accumulator gradients (level 0 aliases dpred)
  size_t bwd_red_off, bwd_red_bytes; // zeroed at the start of every backward
  size_t redS_off, redSx_off, dwse_off, dwse2_off, dweff_off, dcst_off, dymax_off, scale_off, partial_off, htmp1_off, htmp2_off, utmp1_off, utmp2_off;
  uint8_t* ws = nullptr;
  uint8_t* wimg = nullptr;
  XOffsets xo;                      // per-sample input offsets of the current forward
  // backward concurrency (plan_bwd.inc): plan-owned side streams forked from / joined to the caller's stream with events.
  // side[0] carries the weight gradients, side[1..3] the extra dgrad pieces of layers wider than 64 channels, side[4] the
  // adjoint of the heads' folded side-branch up-sampling (needed only when the backward reaches level 1).
  static constexpr int kSideStreams = 5;
  cudaStream_t side[kSideStreams] = {};
  cudaEvent_t ev_fork = nullptr, ev_side[kSideStreams] = {};
  bool wg_pending = false;          // a wgrad on side[0] has not been joined to the caller's stream yet
  long long conc_vox = 0;           // layers with batch * voxels <= conc_vox use the side streams (0: never)
  // optional per-launch CUDA-event timing (bench roofline); events live on the caller's stream
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> ev_label;
  std::vector<double> ev_flops;
  size_t ev_n = 0;
  void mark(const char* label, cudaStream_t st, double flops = 0.0) {
    if (!timing) return;
    if (ev_n >= ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); ev_label.emplace_back(); ev_flops.push_back(0.0); }
    ev_label[ev_n] = label; ev_flops[ev_n] = flops;
    cudaEventRecord(ev[ev_n++], st);
  }
  Dims dims(int level) const { return Dims{N, D >> level, H >> level, W >> level}; }
  long long vox(int level) const { return (long long)(D >> level) * (H >> level) * (W >> level); }
};

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
// backward-pass planning hooks (plan_bwd.inc)
static int bwd_plan_create(seunet_plan* p, size_t& wimg, size_t& off);
static int bwd_plan_bind(seunet_plan* p);
static void bwd_pack_jobs(seunet_plan* p, std::vector<ConvPackJob>& jobs);

extern "C" int seunet_version(void) { return 1; }
extern "C" int seunet_act_dtype(void) {
#ifdef SEUNET_ACT_BF16
  return 1;
#else
  return 0;
#endif
}
extern "C" const char* seunet_last_error(void) { return g_err; }

static bool check_model(int ic, int ncls) {
  if (ic < 1 || ic > kMaxInCh) { seunet_set_error("in_channel=%d unsupported (1..%d)", ic, kMaxInCh); return false; }
  if (ncls != 1) { seunet_set_error("n_classes=%d unsupported (only 1)", ncls); return false; }
  return true;
}
extern "C" int64_t seunet_param_count(int ic, int ncls) { return check_model(ic, ncls) ? build_params(ic, ncls).total : -1; }
extern "C" int64_t seunet_param_offset(int ic, int ncls, const char* name) {
  return check_model(ic, ncls) ? build_params(ic, ncls).off(name) : -1;
}
extern "C" int seunet_param_tensors(int ic, int ncls) { return check_model(ic, ncls) ? (int)build_params(ic, ncls).e.size() : -1; }
extern "C" const char* seunet_param_name(int ic, int ncls, int i) {
  static thread_local std::string s;
  if (!check_model(ic, ncls)) return nullptr;
  ParamTable t = build_params(ic, ncls);
  if (i < 0 || i >= (int)t.e.size()) return nullptr;
  s = t.e[i].name;
  return s.c_str();
}
extern "C" int64_t seunet_param_numel(int ic, int ncls, int i) {
  if (!check_model(ic, ncls)) return -1;
  ParamTable t = build_params(ic, ncls);
  if (i < 0 || i >= (int)t.e.size()) return -1;
  return t.e[i].numel;
}

extern "C" int seunet_plan_create(seunet_plan_t** out, int batch, int D, int H, int W, int in_ch, int ncls, int mode,
                                  int device) {
  if (!out) { seunet_set_error("null plan pointer"); return 1; }
  if (!check_model(in_ch, ncls)) return 1;
  if (batch < 1 || D < 8 || H < 8 || W < 8 || ((D | H | W) & 7)) {
    seunet_set_error("bad shape: batch=%d D=%d H=%d W=%d (dims must be positive multiples of 8)", batch, D, H, W);
    return 1;
  }
  seunet_plan* p = new seunet_plan();
  p->N = batch; p->D = D; p->H = H; p->W = W; p->in_ch = in_ch; p->ncls = ncls; p->mode = mode; p->device = device;
  p->pt = build_params(in_ch, ncls);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    seunet_set_error("cudaGetDeviceProperties(%d) failed - no CUDA device", device);
    delete p; return 1;
  }
  if (prop.major != 10) {
    seunet_set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    delete p; return 1;
  }
  p->num_sms = prop.multiProcessorCount;
  p->fuse_cat = getenv("SEUNET_CAT_FUSION") == nullptr || atoi(getenv("SEUNET_CAT_FUSION")) != 0;

  // --- conv geometry + packed weight image layout
  size_t wimg = 0;
  for (int i = 0; i < 18; ++i) {
    const SseDesc& s = kSse[i];
    const int cin = (i == S_EC1) ? in_ch : s.cin;
    if (conv_geom_init(&p->sse_conv[i].g, cin, s.cout, 3, s.dil)) { delete p; return 1; }
    p->sse_conv[i].w_off = p->pt.off(std::string(s.name) + ".conv1.weight");
    p->sse_conv[i].wimg_off = wimg;
    wimg += align_up(p->sse_conv[i].g.wimg_bytes());
  }
  for (int i = 0; i < 6; ++i) {
    const CatDesc& c = kCat[i];
    if (conv_geom_init(&p->cat_conv[i].g, c.cin, c.cout, 1, 0)) { delete p; return 1; }
    p->cat_conv[i].w_off = p->pt.off(std::string(c.name) + ".conv1.weight");
    p->cat_conv[i].wimg_off = wimg;
    wimg += align_up(p->cat_conv[i].g.wimg_bytes());
  }
  // --- workspace layout
  size_t off = 0;
  for (int b = 0; b < B_COUNT; ++b) {
    p->buf_off[b] = off;
    off += align_up((size_t)batch * kBufs[b].chunks * p->vox(kBufs[b].level) * 16);
  }
  // raw conv outputs: per layer in training mode, one scratch per level in inference mode
  size_t raw_level_off[4];
  for (int l = 0; l < 4; ++l) {
    raw_level_off[l] = off;
    if (mode == 0) off += align_up((size_t)batch * (l == 0 ? 4 : 8) * p->vox(l) * 16);
  }
  auto raw_alloc = [&](ConvSlot& cs, int level) {
    if (mode == 0) { cs.raw_off = raw_level_off[level]; return; }
    cs.raw_off = off;
    off += align_up((size_t)batch * (cs.g.COUT / 8) * p->vox(level) * 16);
  };
  for (int i = 0; i < 18; ++i) raw_alloc(p->sse_conv[i], kSse[i].level);
  for (int i = 0; i < 6; ++i) raw_alloc(p->cat_conv[i], kCat[i].level);
  p->stats_off = off;
  for (int i = 0; i < 18; ++i) { p->sse_conv[i].stats_off = off; off += (size_t)batch * 64 * 2 * sizeof(double); }
  for (int i = 0; i < 6; ++i) { p->cat_conv[i].stats_off = off; off += (size_t)batch * 64 * 2 * sizeof(double); }
  p->stats_bytes = off - p->stats_off;
  off = align_up(off);
  for (int l = 0; l < 4; ++l) { p->T0_off[l] = off; off += align_up((size_t)batch * p->vox(l) * 4); }
  for (int l = 0; l < 3; ++l) { p->T1_off[l] = off; off += align_up((size_t)batch * p->vox(l) * 4); }
  p->xp1_off = off; off += align_up((size_t)batch * in_ch * p->vox(1) * 4);
  p->xp2_off = off; off += align_up((size_t)batch * in_ch * p->vox(2) * 4);
  p->mom_off = off; off += align_up((size_t)3 * batch * kMomStride * sizeof(double));
  p->weff_off = off; off += align_up((size_t)18 * batch * 64 * 4);
  p->wcst_off = off; off += align_up((size_t)18 * batch * 4);
  if (mode == 1 && bwd_plan_create(p, wimg, off)) { delete p; return 1; }
  p->wimg_bytes = wimg;
  p->ws_bytes = off;

  // --- folded head weights table
  memset(&p->headw, 0, sizeof(p->headw));
  p->headw.nblk = 18;
  for (int i = 0; i < 18; ++i) {
    const SseDesc& s = kSse[i];
    HeadwBlock& hb = p->headw.blk[i];
    hb.w2_off = (int)p->pt.off(std::string(s.name) + ".conv2.weight");
    hb.b2_off = (int)p->pt.off(std::string(s.name) + ".conv2.bias");
    hb.C = s.cout; hb.head = s.head; hb.k = s.k;
  }
  p->headw.hw_off[0] = (int)p->pt.off("dc0_0.weight");
  p->headw.hw_off[1] = (int)p->pt.off("dc0_1.weight");
  *out = p;
  return 0;
}

extern "C" void seunet_plan_destroy(seunet_plan_t* p) {
  if (!p) return;
  for (auto e : p->ev) cudaEventDestroy(e);
  for (int i = 0; i < seunet_plan::kSideStreams; ++i) {
    if (p->ev_side[i]) cudaEventDestroy(p->ev_side[i]);
    if (p->side[i]) cudaStreamDestroy(p->side[i]);
  }
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  delete p;
}
extern "C" int seunet_plan_set_timing(seunet_plan_t* p, int on) { if (!p) return 1; p->timing = on != 0; p->ev_n = 0; return 0; }
extern "C" int seunet_plan_timing_count(const seunet_plan_t* p) { return p && p->ev_n ? (int)p->ev_n - 1 : 0; }
// Interval i = time between mark i and mark i+1 (the launches issued after mark i), label/flops of mark i+1.
extern "C" int seunet_plan_timing_get(const seunet_plan_t* p, int i, const char** label, float* ms, double* flops) {
  if (!p || i < 0 || i + 1 >= (int)p->ev_n) { seunet_set_error("timing_get: index out of range"); return 1; }
  SEUNET_CUDA_CHECK(cudaEventElapsedTime(ms, p->ev[i], p->ev[i + 1]));
  *label = p->ev_label[i + 1].c_str();
  *flops = p->ev_flops[i + 1];
  return 0;
}
extern "C" size_t seunet_plan_workspace_bytes(const seunet_plan_t* p) { return p ? p->ws_bytes : 0; }

// Test hook: locate an intermediate tensor inside the bound workspace.  Names: the activation buffers
// "XB","CAT1","DC5IN","D2","P1","CAT2","DC3IN","DC42IN","D1F","P2","CAT3","DC1IN","DC22IN","D0F","P3","CAT4","E7F"
// (chunk planes), "raw:<layer>" (raw conv output, chunk planes; training-mode plans keep one per layer),
// "T0:<level>" / "T1:<level>" (fp32 head accumulators, chunks = 0).
extern "C" int seunet_plan_debug_buffer(const seunet_plan_t* p, const char* name, void** ptr, int* chunks, int* level) {
  if (!p || !p->ws) { seunet_set_error("debug_buffer: plan not bound"); return 1; }
  static const char* bnames[B_COUNT] = {"XB", "CAT1", "DC5IN", "D2", "P1", "CAT2", "DC3IN", "DC42IN", "D1F",
                                        "P2", "CAT3", "DC1IN", "DC22IN", "D0F", "P3", "CAT4", "E7F"};
  const std::string n(name);
  for (int b = 0; b < B_COUNT; ++b)
    if (n == bnames[b]) { *ptr = p->ws + p->buf_off[b]; *chunks = kBufs[b].chunks; *level = kBufs[b].level; return 0; }
  if (n.rfind("raw:", 0) == 0) {
    for (int i = 0; i < 18; ++i)
      if (n.substr(4) == kSse[i].name) { *ptr = p->ws + p->sse_conv[i].raw_off; *chunks = p->sse_conv[i].g.COUT / 8; *level = kSse[i].level; return 0; }
    for (int i = 0; i < 6; ++i)
      if (n.substr(4) == kCat[i].name) { *ptr = p->ws + p->cat_conv[i].raw_off; *chunks = p->cat_conv[i].g.COUT / 8; *level = kCat[i].level; return 0; }
  }
  if ((n.rfind("T0:", 0) == 0 || n.rfind("T1:", 0) == 0) && n.size() == 4) {
    const int l = n[3] - '0';
    if (l >= 0 && l < (n[1] == '0' ? 4 : 3)) { *ptr = p->ws + (n[1] == '0' ? p->T0_off[l] : p->T1_off[l]); *chunks = 0; *level = l; return 0; }
  }
  if (n.rfind("stats:", 0) == 0) {   // fp64 [N][COUT][2] (sum, sum of squares); *chunks = COUT
    for (int i = 0; i < 18; ++i)
      if (n.substr(6) == kSse[i].name) { *ptr = p->ws + p->sse_conv[i].stats_off; *chunks = p->sse_conv[i].g.COUT; *level = kSse[i].level; return 0; }
    for (int i = 0; i < 6; ++i)
      if (n.substr(6) == kCat[i].name) { *ptr = p->ws + p->cat_conv[i].stats_off; *chunks = p->cat_conv[i].g.COUT; *level = kCat[i].level; return 0; }
  }
  if (p->mode == 1 && (n.rfind("dy:", 0) == 0 || n.rfind("dn:", 0) == 0) && n.size() == 4) {
    const int l = n[3] - '0';
    if (l >= 0 && l < 4) { *ptr = p->ws + (n[1] == 'y' ? p->dy_off[l] : p->dn_off[l]); *chunks = l == 0 ? 4 : 8; *level = l; return 0; }
  }
  if (p->mode == 1 && n.rfind("scale:", 0) == 0) {
    *ptr = p->ws + p->scale_off + 8 * atoi(n.c_str() + 6); *chunks = 0; *level = 0; return 0;
  }
  if (p->mode == 1 && n.rfind("g:", 0) == 0)
    for (int b = 1; b < B_COUNT; ++b)
      if (n.substr(2) == bnames[b]) { *ptr = p->ws + p->gbuf_off[b]; *chunks = kBufs[b].chunks; *level = kBufs[b].level; return 0; }
  seunet_set_error("debug_buffer: unknown buffer '%s'", name);
  return 1;
}
extern "C" size_t seunet_plan_wimg_bytes(const seunet_plan_t* p) { return p ? p->wimg_bytes : 0; }

extern "C" int seunet_plan_bind(seunet_plan_t* p, void* workspace, void* wimg, seunet_stream_t stream) {
  if (!p || !workspace || !wimg) { seunet_set_error("plan_bind: null argument"); return 1; }
  if (((uintptr_t)workspace | (uintptr_t)wimg) & 255) { seunet_set_error("plan_bind: buffers must be 256-byte aligned"); return 1; }
  cudaStream_t st = (cudaStream_t)stream;
  p->ws = (uint8_t*)workspace;
  p->wimg = (uint8_t*)wimg;
  // padding chunk of the 56-channel concat (SE_UNet.py:186) must read as zeros
  for (int n = 0; n < p->N; ++n) {
    uint8_t* pad = p->ws + p->buf_off[B_CAT1] + ((size_t)n * 8 + 7) * p->vox(0) * 16;
    SEUNET_CUDA_CHECK(cudaMemsetAsync(pad, 0, (size_t)p->vox(0) * 16, st));
  }
  for (int i = 0; i < 18; ++i) {
    const SseDesc& s = kSse[i];
    ConvSlot& cs = p->sse_conv[i];
    const Dims d = p->dims(s.level);
    // inference plans do not store the zero padding planes of the raw output (ec1: 8 real channels in a COUT = 16 tile; ncu showed
    // the conv writing as many DRAM bytes as ec2); training plans keep them defined for the debug/test accessors
    if (conv_launch_init(&cs.L, cs.g, d.N, d.D, d.H, d.W, p->ws + p->buf_off[s.in_buf], kBufs[s.in_buf].chunks, s.in_off,
                         p->ws + cs.raw_off, cs.g.COUT / 8, 0, (double*)(p->ws + cs.stats_off), p->wimg + cs.wimg_off,
                         p->num_sms, 0, p->mode == 0 ? (cs.g.Cout_real + 7) / 8 : -1))
      return 1;
  }
  for (int i = 0; i < 6; ++i) {
    const CatDesc& c = kCat[i];
    ConvSlot& cs = p->cat_conv[i];
    const Dims d = p->dims(c.level);
    if (conv_launch_init(&cs.L, cs.g, d.N, d.D, d.H, d.W, p->ws + p->buf_off[c.in_buf], kBufs[c.in_buf].chunks, 0,
                         p->ws + cs.raw_off, cs.g.COUT / 8, 0, (double*)(p->ws + cs.stats_off), p->wimg + cs.wimg_off,
                         p->num_sms))
      return 1;
  }
  if (p->mode == 1 && bwd_plan_bind(p)) return 1;
  return 0;
}

extern "C" int seunet_pack_weights(seunet_plan_t* p, const float* params, seunet_stream_t stream) {
  if (!p || !p->wimg) { seunet_set_error("pack_weights: plan not bound"); return 1; }
  cudaStream_t st = (cudaStream_t)stream;
  // every layer image of the plan (forward convs, and in training plans the mirrored data-gradient pieces) in one batched
  // launch per kConvPackBatch layers
  std::vector<ConvPackJob> jobs;
  for (int i = 0; i < 18; ++i) jobs.push_back(conv_pack_job(p->sse_conv[i].g, p->sse_conv[i].w_off, (long long)p->sse_conv[i].wimg_off, 0));
  for (int i = 0; i < 6; ++i) jobs.push_back(conv_pack_job(p->cat_conv[i].g, p->cat_conv[i].w_off, (long long)p->cat_conv[i].wimg_off, 0));
  if (p->mode == 1) bwd_pack_jobs(p, jobs);
  return conv_pack_weights_batch(jobs.data(), (int)jobs.size(), params, p->wimg, st);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
static bool fuse_cat(const seunet_plan* p, int cat) {
  if (p->mode != 0 || cat < 0 || !p->fuse_cat) return false;
  for (int i = 0; i < 18; ++i) if (kSseFuse[i] == cat) return true;   // some block's apply pass computes this CAT conv
  return false;
}

static int run_sse(seunet_plan* p, int i, const float* params, cudaStream_t st) {
  const SseDesc& s = kSse[i];
  ConvSlot& cs = p->sse_conv[i];
  if (conv_launch_run(cs.L, st)) return 1;
  p->mark((std::string("conv:") + s.name).c_str(), st, 2.0 * p->N * p->vox(s.level) * cs.g.Cin_real * cs.g.Cout_real * 27.0);
  SseArgs a;
  memset(&a, 0, sizeof(a));
  a.raw = (const act_t*)(p->ws + cs.raw_off); a.raw_chunks = cs.g.COUT / 8;
  a.stats = (const double*)(p->ws + cs.stats_off); a.stats_c = cs.g.COUT;
  a.V = p->vox(s.level);
  const std::string nm(s.name);
  a.wse = params + p->pt.off(nm + ".conv_se.weight");
  a.wse2 = s.gates == 2 ? params + p->pt.off(nm + ".conv_se2.weight") : nullptr;
  a.weff = (const float*)(p->ws + p->weff_off) + (size_t)i * p->N * 64;
  a.wcst = (const float*)(p->ws + p->wcst_off) + (size_t)i * p->N;
  a.T = (s.head == 0 && p->skip_head0) ? nullptr : (float*)(p->ws + (s.head == 0 ? p->T0_off[s.level] : p->T1_off[s.level]));
  a.t_init = (s.k % 3 == 0 && s.head == 0) || (s.head == 1 && s.k % 2 == 0);
  a.inference = p->mode == 0;
  if (fuse_cat(p, kSseFuse[i])) {
    // the block output never leaves the SM: apply + CAT 1x1x1 conv in one pass (the CAT conv launch is skipped in run_cat)
    const CatDesc& c = kCat[kSseFuse[i]];
    ConvSlot& cc = p->cat_conv[kSseFuse[i]];
    CatFuseArgs f;
    memset(&f, 0, sizeof(f));
    f.cat = (const act_t*)(p->ws + p->buf_off[c.in_buf]); f.cat_chunks = kBufs[c.in_buf].chunks; f.cat_real_chunks = (c.cin + 7) / 8;
    f.w = params + cc.w_off; f.cin_real = c.cin;
    f.kcat = ((c.cin + 15) / 16) * 16; f.nout = c.cout;
    f.out = (act_t*)(p->ws + cc.raw_off); f.out_chunks = cc.g.COUT / 8;
    f.out_stats = (double*)(p->ws + cc.stats_off); f.out_stats_c = cc.g.COUT;
    if (launch_apply_sse_cat(s.cout, p->N, a, f, p->num_sms, st)) return 1;
    p->mark((std::string("apply:") + s.name).c_str(), st);
    return 0;
  }
  if (s.out_buf != B_NONE) {
    a.dest = (act_t*)(p->ws + p->buf_off[s.out_buf]); a.dest_chunks = kBufs[s.out_buf].chunks; a.dest_off = s.out_off;
  }
  if (launch_apply_sse(s.cout, p->N, a, st)) return 1;
  p->mark((std::string("apply:") + s.name).c_str(), st);
  return 0;
}

static int run_cat(seunet_plan* p, int i, const float* params, const float* x, const int64_t* xs, cudaStream_t st) {
  const CatDesc& c = kCat[i];
  ConvSlot& cs = p->cat_conv[i];
  if (!fuse_cat(p, i)) {   // fused plans: the contraction was done inside the last producer's apply pass (no launch, no timing entry)
    if (conv_launch_run(cs.L, st)) return 1;
    p->mark((std::string("conv:") + c.name).c_str(), st, 2.0 * p->N * p->vox(c.level) * cs.g.Cin_real * cs.g.Cout_real);
  }
  CatArgs a;
  memset(&a, 0, sizeof(a));
  a.raw = (const act_t*)(p->ws + cs.raw_off); a.raw_chunks = cs.g.COUT / 8;
  a.stats = (const double*)(p->ws + cs.stats_off); a.stats_c = cs.g.COUT;
  a.d = p->dims(c.level);
  if (c.xname) {
    a.in_ch = p->in_ch;
    a.wx = params + p->pt.off(std::string(c.xname) + ".conv1.weight");
    a.mom = (const double*)(p->ws + p->mom_off) + (size_t)c.level * p->N * kMomStride;
    if (c.level == 0) {
      a.x = x;
      for (int k = 0; k < 5; ++k) a.xs[k] = xs[k];
      a.xo = p->xo;
    } else {
      a.x = (const float*)(p->ws + (c.level == 1 ? p->xp1_off : p->xp2_off));
      const long long V = p->vox(c.level);
      a.xs[0] = (long long)p->in_ch * V; a.xs[1] = V; a.xs[2] = (long long)a.d.H * a.d.W; a.xs[3] = a.d.W; a.xs[4] = 1;
    }
  }
  a.dest = (act_t*)(p->ws + p->buf_off[c.out_buf]); a.dest_chunks = kBufs[c.out_buf].chunks; a.dest_off = c.out_off;
  if (c.pool_buf != B_NONE) {
    a.pdest = (act_t*)(p->ws + p->buf_off[c.pool_buf]); a.pdest_chunks = kBufs[c.pool_buf].chunks; a.pdest_off = 0;
  }
  if (launch_apply_cat(c.cout, a, st)) return 1;
  p->mark((std::string("cat:") + c.name).c_str(), st);
  return 0;
}

struct WindowSink { uint32_t* acc; int X, Y, Z, acc_log2; const int* starts; };

static int forward_impl(seunet_plan_t* p, const float* x, const int64_t* xs, const int64_t* x_offsets, const float* params,
                        const float* drop0, const float* drop1, float* pred0, float* pred1, const WindowSink* sink,
                        cudaStream_t st) {
  auto act = [&](int b) { return (act_t*)(p->ws + p->buf_off[b]); };
  memset(&p->xo, 0, sizeof(p->xo));
  if (x_offsets) {
    if (p->N > kMaxWindowBatch) { seunet_set_error("forward: x_offsets supports at most %d samples", kMaxWindowBatch); return 1; }
    p->xo.use = 1;
    for (int n = 0; n < p->N; ++n) p->xo.off[n] = x_offsets[n];
  }
  p->ev_n = 0;
  p->mark("start", st);
  SEUNET_CUDA_CHECK(cudaMemsetAsync(p->ws + p->stats_off, 0, p->stats_bytes, st));
  long long xs_ll[5];
  for (int k = 0; k < 5; ++k) xs_ll[k] = xs[k];
  if (launch_input_prep(x, xs_ll, p->xo, p->in_ch, p->dims(0), act(B_XB), (float*)(p->ws + p->xp1_off),
                        (float*)(p->ws + p->xp2_off), (double*)(p->ws + p->mom_off), st, p->mode == 0)) return 1;
  if (launch_headw(params, drop0, drop1, p->N, p->headw, (float*)(p->ws + p->weff_off), (float*)(p->ws + p->wcst_off), st)) return 1;
  p->mark("prep", st);
  // encoder, level 0 (SE_UNet.py:183-189)
  if (run_sse(p, S_EC1, params, st) || run_sse(p, S_EC2, params, st) || run_sse(p, S_EC3, params, st)) return 1;
  if (run_cat(p, C_EC33, params, x, xs, st)) return 1;
  // level 1 (192-198)
  if (run_sse(p, S_EC4, params, st) || run_sse(p, S_EC5, params, st) || run_sse(p, S_EC6, params, st)) return 1;
  if (run_cat(p, C_EC63, params, x, xs, st)) return 1;
  // level 2 (201-206)
  if (run_sse(p, S_EC7, params, st) || run_sse(p, S_EC8, params, st) || run_sse(p, S_EC9, params, st)) return 1;
  if (run_cat(p, C_EC93, params, x, xs, st)) return 1;
  // level 3 (209-212)
  if (run_sse(p, S_EC10, params, st) || run_sse(p, S_EC11, params, st) || run_sse(p, S_EC12, params, st)) return 1;
  if (run_cat(p, C_EC123, params, x, xs, st)) return 1;
  // decoder (214-229)
  if (launch_upsample2(act(B_E7F), 64, p->dims(3), act(B_DC1IN), kBufs[B_DC1IN].chunks, 0, st)) return 1;
  p->mark("up:e7", st);
  if (run_sse(p, S_DC1, params, st) || run_sse(p, S_DC2, params, st)) return 1;
  if (run_cat(p, C_DC22, params, x, xs, st)) return 1;
  if (launch_upsample2(act(B_D0F), 64, p->dims(2), act(B_DC3IN), kBufs[B_DC3IN].chunks, 0, st)) return 1;
  p->mark("up:d0", st);
  if (run_sse(p, S_DC3, params, st) || run_sse(p, S_DC4, params, st)) return 1;
  if (run_cat(p, C_DC42, params, x, xs, st)) return 1;
  if (launch_upsample2(act(B_D1F), 32, p->dims(1), act(B_DC5IN), kBufs[B_DC5IN].chunks, 0, st)) return 1;
  p->mark("up:d1", st);
  if (run_sse(p, S_DC5, params, st) || run_sse(p, S_DC6, params, st)) return 1;
  // dc62 (SE_UNet.py:230) is dead code in the reference: its result is never used.
  // heads (232-233)
  HeadArgs h;
  memset(&h, 0, sizeof(h));
  h.d = p->dims(0);
  for (int l = 0; l < 4; ++l) h.T0[l] = (const float*)(p->ws + p->T0_off[l]);
  for (int l = 0; l < 3; ++l) h.T1[l] = (const float*)(p->ws + p->T1_off[l]);
  h.bias0 = params + p->pt.off("dc0_0.bias");
  h.bias1 = params + p->pt.off("dc0_1.bias");
  h.pred0 = pred0; h.pred1 = pred1;
  if (sink) {
    h.acc = sink->acc; h.X = sink->X; h.Y = sink->Y; h.Z = sink->Z; h.acc_scale = (float)(1u << sink->acc_log2);
    for (int n = 0; n < p->N; ++n)
      for (int k = 0; k < 3; ++k) h.s[n][k] = sink->starts[n * 3 + k];
  }
  if (launch_head(h, st)) return 1;
  p->mark("head", st);
  return 0;
}

extern "C" int seunet_forward(seunet_plan_t* p, const float* x, const int64_t* xs, const int64_t* x_offsets,
                              const float* params, const float* drop0, const float* drop1, float* pred0, float* pred1,
                              seunet_stream_t stream) {
  if (!p || !p->ws) { seunet_set_error("forward: plan not bound"); return 1; }
  if (!x || !xs || !params || !drop0 || !drop1 || !pred0 || !pred1) { seunet_set_error("forward: null argument"); return 1; }
  p->skip_head0 = false;
  return forward_impl(p, x, xs, x_offsets, params, drop0, drop1, pred0, pred1, nullptr, (cudaStream_t)stream);
}

// One sliding-window step of prediction.py:102-106 for a batch of windows: forward, sigmoid of the second output, += into the
// volume accumulator.  prediction.py drops the first output (p0), so the whole deep-supervision head 0 - the folded side
// branches of ec1..ec12 and their accumulators - is not computed, and the head-1 logits never reach memory: the head kernel
// adds sigmoid(pred1) to the fixed-point accumulator directly (same arithmetic as seunet_window_accumulate).
extern "C" int seunet_forward_window(seunet_plan_t* p, const float* x, const int64_t* xs, const int64_t* x_offsets,
                                     const float* params, const float* drop0, const float* drop1, const int* starts,
                                     uint32_t* acc, int X, int Y, int Z, int acc_log2, seunet_stream_t stream) {
  if (!p || !p->ws) { seunet_set_error("forward_window: plan not bound"); return 1; }
  if (!x || !xs || !params || !drop0 || !drop1 || !starts || !acc) { seunet_set_error("forward_window: null argument"); return 1; }
  if (p->mode != 0) { seunet_set_error("forward_window: needs an inference plan (mode 0)"); return 1; }
  if (p->N > kMaxWindowBatch) { seunet_set_error("forward_window: at most %d windows per call", kMaxWindowBatch); return 1; }
  if (acc_log2 < 8 || acc_log2 > 30) { seunet_set_error("forward_window: acc_log2 %d out of range (8..30)", acc_log2); return 1; }
  for (int n = 0; n < p->N; ++n)
    for (int k = 0; k < 3; ++k) {
      const int lim = k == 0 ? X - p->D : (k == 1 ? Y - p->H : Z - p->W);
      if (starts[n * 3 + k] < 0 || starts[n * 3 + k] > lim) { seunet_set_error("forward_window: window %d out of bounds", n); return 1; }
    }
  WindowSink sink{acc, X, Y, Z, acc_log2, starts};
  p->skip_head0 = true;
  const int rc = forward_impl(p, x, xs, x_offsets, params, drop0, drop1, nullptr, nullptr, &sink, (cudaStream_t)stream);
  p->skip_head0 = false;
  return rc;
}

// ---------------------------------------------------------------------------------------------
// single-op entry points
// ---------------------------------------------------------------------------------------------
__global__ void to_chunks_kernel(const float* __restrict__ src, int C, long long V, act_t* __restrict__ dst, int dst_chunks,
                                 int dst_off) {
  const int n = blockIdx.z, k = blockIdx.y;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= V) return;
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = k * 8 + i;
    f[i] = c < C ? src[((size_t)n * C + c) * V + v] : 0.f;
  }
  st_chunk(dst + (((size_t)n * dst_chunks + dst_off + k) * V + v) * 8, floats_to_chunk(f));
}
__global__ void from_chunks_kernel(const act_t* __restrict__ src, int src_chunks, int src_off, int C, long long V,
                                   float* __restrict__ dst) {
  const int n = blockIdx.z, k = blockIdx.y;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= V) return;
  float f[8];
  chunk_to_floats(ld_chunk(src + (((size_t)n * src_chunks + src_off + k) * V + v) * 8), f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = k * 8 + i;
    if (c < C) dst[((size_t)n * C + c) * V + v] = f[i];
  }
}
extern "C" int seunet_to_chunks(const float* src, int N, int C, int D, int H, int W, void* dst, int dst_chunks, int dst_off,
                                seunet_stream_t stream) {
  const long long V = (long long)D * H * W;
  dim3 grid((unsigned)((V + 255) / 256), (C + 7) / 8, N);
  to_chunks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, C, V, (act_t*)dst, dst_chunks, dst_off);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int seunet_from_chunks(const void* src, int src_chunks, int src_off, int N, int C, int D, int H, int W,
                                  float* dst, seunet_stream_t stream) {
  const long long V = (long long)D * H * W;
  dim3 grid((unsigned)((V + 255) / 256), (C + 7) / 8, N);
  from_chunks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const act_t*)src, src_chunks, src_off, C, V, dst);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Test hook: fill the shared memory of every SM with 0xFF (NaN in every 16/32-bit float format) so that a kernel
// relying on stale shared memory (e.g. 0-weight * garbage in a padded UMMA K half) is caught by the parity tests.
__global__ void poison_smem_kernel(int bytes) {
  extern __shared__ uint4 sm[];
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) sm[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  __syncthreads();
  if (sm[(threadIdx.x * 7) % (bytes / 16)].x == 1u) printf("unreachable\n");
}
extern "C" int seunet_debug_poison_smem(seunet_stream_t stream) {
  const int bytes = 200 * 1024;
  SEUNET_CUDA_CHECK(cudaFuncSetAttribute(poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  int dev = 0, sms = 0;
  SEUNET_CUDA_CHECK(cudaGetDevice(&dev));
  SEUNET_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  poison_smem_kernel<<<sms * 2, 256, bytes, (cudaStream_t)stream>>>(bytes);
  SEUNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" size_t seunet_conv_scratch_bytes(int Cin, int Cout, int ksize, int dil) {
  ConvGeom g;
  if (conv_geom_init(&g, Cin, Cout, ksize, dil)) return 0;
  return align_up(g.wimg_bytes());
}
extern "C" int seunet_conv_fprop(const void* in, int in_chunks, int in_chunk_off, const float* w, int N, int D, int H,
                                 int W, int Cin, int Cout, int ksize, int dil, void* out, double* stats, void* scratch,
                                 int transpose_flip, int bf16, int grad_out, int accum_out, seunet_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ConvGeom g;
  if (conv_geom_init(&g, Cin, Cout, ksize, dil, bf16)) return 1;
  int dev = 0, sms = 0;
  SEUNET_CUDA_CHECK(cudaGetDevice(&dev));
  SEUNET_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (conv_pack_weights(g, w, scratch, transpose_flip, st)) return 1;
  if (stats) SEUNET_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * N * g.COUT * 2, st));
  ConvLaunch L;
  if (accum_out && !grad_out) { seunet_set_error("conv_fprop: accum_out needs grad_out"); return 1; }
  if (conv_launch_init(&L, g, N, D, H, W, in, in_chunks, in_chunk_off, out, (Cout + 7) / 8, 0, stats, scratch, sms, accum_out,
                       (Cout + 7) / 8, grad_out, nullptr, /*shallow_ok=*/grad_out))   // gradient-format launches tile like the plans' dgrads
    return 1;
  return conv_launch_run(L, st);
}

extern "C" size_t seunet_wgrad_scratch_bytes(int Cin, int Cout, int ksize) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  return wgrad_partial_bytes(Cin, Cout, ksize, sms);
}
extern "C" int seunet_conv_wgrad(const void* x, int x_chunks, int x_chunk_off, const void* dy, int dy_chunks, int dy_chunk_off,
                                 int N, int D, int H, int W, int Cin, int Cout, int ksize, int dil, void* scratch, float* dw,
                                 seunet_stream_t stream) {
  int dev = 0, sms = 0;
  SEUNET_CUDA_CHECK(cudaGetDevice(&dev));
  SEUNET_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  WgradLaunch L;
  if (wgrad_launch_init(&L, N, D, H, W, Cin, Cout, ksize, dil, x, x_chunks, x_chunk_off, 0, dy, dy_chunks, dy_chunk_off,
                        (float*)scratch, sms))
    return 1;
  return wgrad_launch_run(L, dw, nullptr, (cudaStream_t)stream);
}

#include "plan_bwd.inc"
