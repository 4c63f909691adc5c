// Shared device helpers for the SE-UNet sm_100a kernels: storage type, PTX wrappers for
// mbarrier / TMA / tcgen05 (TMEM alloc, UMMA issue, commit, TMEM loads), small math helpers.
// Everything here is inline PTX for sm_100a; there is no fallback path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>

// ---------------------------------------------------------------------------------------------
// Activation / weight storage type for the tensor-core path.
// Default is IEEE fp16 (same tcgen05 kind::f16 pipe and rate as bf16, 3 more mantissa bits):
// with bf16 storage the measured logit error of the 24-conv-deep network exceeds the 2e-2
// parity tolerance (see DESIGN.md "Numerics").  -DSEUNET_ACT_BF16 switches the whole path to bf16.
// ---------------------------------------------------------------------------------------------
#ifdef SEUNET_ACT_BF16
typedef __nv_bfloat16 act_t;
typedef __nv_bfloat162 act2_t;
#define SEUNET_UMMA_FMT 1u  // BF16
__device__ __forceinline__ float act2f(act_t v) { return __bfloat162float(v); }
__device__ __forceinline__ act_t f2act(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ uint32_t pack_act2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_act2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}
#else
typedef __half act_t;
typedef __half2 act2_t;
#define SEUNET_UMMA_FMT 0u  // F16
__device__ __forceinline__ float act2f(act_t v) { return __half2float(v); }
__device__ __forceinline__ act_t f2act(float v) { return __float2half_rn(v); }
__device__ __forceinline__ uint32_t pack_act2(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_act2(uint32_t u) {
  __half2 t = *reinterpret_cast<__half2*>(&u);
  return __half22float2(t);
}
#endif

// One "chunk" = 8 channels of one voxel = 16 bytes.  Activations live in HBM as
// [n][C/8][D][H][W][8] ("channel-chunk planes"): a voxel shift of one along W is a 16-byte
// address shift, which is what lets every 3x3x3 tap be a plain start-address offset of a UMMA
// shared-memory descriptor (no-swizzle K-major canonical layout).
struct __align__(16) Chunk8 { uint32_t u[4]; };

__device__ __forceinline__ void chunk_to_floats(const Chunk8& c, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = unpack_act2(c.u[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ Chunk8 floats_to_chunk(const float* f) {
  Chunk8 c;
#pragma unroll
  for (int i = 0; i < 4; ++i) c.u[i] = pack_act2(f[2 * i], f[2 * i + 1]);
  return c;
}

// Gradient tensors are always bf16 (range of fp32: no loss scaling needed in the backward pass).
__device__ __forceinline__ void chunk_to_floats_bf16(const Chunk8& c, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&c.u[i]);
    float2 v = __bfloat1622float2(t);
    f[2 * i] = v.x; f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ Chunk8 floats_to_chunk_bf16(const float* f) {
  Chunk8 c;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    c.u[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  return c;
}
__device__ __forceinline__ Chunk8 ld_chunk(const void* p) {
  Chunk8 c;
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  c.u[0] = v.x; c.u[1] = v.y; c.u[2] = v.z; c.u[3] = v.w;
  return c;
}
// streaming variants: data touched once per kernel should not pollute L1
__device__ __forceinline__ Chunk8 ld_chunk_stream(const void* p) {
  Chunk8 c;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(c.u[0]), "=r"(c.u[1]), "=r"(c.u[2]), "=r"(c.u[3]) : "l"(p));
  return c;
}
__device__ __forceinline__ void st_chunk(void* p, const Chunk8& c) {
  *reinterpret_cast<uint4*>(p) = make_uint4(c.u[0], c.u[1], c.u[2], c.u[3]);
}

// Gradients w.r.t. activations (dn scratch and the per-buffer gradient accumulators) are fp32 by default: the weight
// gradient is a heavily cancelling sum, and bf16 storage of its inputs costs ~10x the 1e-2 tolerance at small volumes.
// -DSEUNET_GRAD_BF16 halves that traffic at the price of accuracy.
#ifdef SEUNET_GRAD_BF16
typedef __nv_bfloat16 grad_t;
__device__ __forceinline__ void ld_grad8(const grad_t* p, float* f) { chunk_to_floats_bf16(ld_chunk_stream(p), f); }
__device__ __forceinline__ void ld_grad8_cached(const grad_t* p, float* f) {
  Chunk8 c; const uint4 v = *reinterpret_cast<const uint4*>(p); c.u[0] = v.x; c.u[1] = v.y; c.u[2] = v.z; c.u[3] = v.w;
  chunk_to_floats_bf16(c, f);
}
__device__ __forceinline__ void st_grad8(grad_t* p, const float* f) {
  const Chunk8 c = floats_to_chunk_bf16(f);
  *reinterpret_cast<uint4*>(p) = make_uint4(c.u[0], c.u[1], c.u[2], c.u[3]);
}
#else
typedef float grad_t;
// One fp32 gradient chunk (8 channels of a voxel) is 32 bytes = one DRAM sector: sm_100 has 256-bit global loads/stores, so a
// chunk is ONE instruction.  (ncu on the two-float4 version: 44 % "excessive sectors" in sse_bwd_a - each 128-bit half request
// of a warp touches only half of every sector it fetches - and the L1 data pipeline, not DRAM, was the busy unit.)
__device__ __forceinline__ void ld_grad8(const grad_t* p, float* f) {
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "l"(p));
}
__device__ __forceinline__ void ld_grad8_cached(const grad_t* p, float* f) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "l"(p));
}
__device__ __forceinline__ void st_grad8(grad_t* p, const float* f) {
  asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]),
               "f"(f[6]), "f"(f[7]), "l"(p) : "memory");
}
// p[0..7] += f[0..7] as two fire-and-forget vector reductions executed in L2: no read round trip through the SM
#define SEUNET_HAVE_RED_GRAD8 1
__device__ __forceinline__ void red_grad8(grad_t* p, const float* f) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]) : "memory");
}
#endif

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float lrelu_(float x) { return x > 0.f ? x : 0.01f * x; }

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk[.tensor])
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, UMMA, commit, loads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, no-swizzle ("interleave") canonical layout, sm_100 version 1.
// In 16-byte units the K-major canonical form is ((8,n),2):((1,SBO),LBO): 8 rows 16 B apart form a
// core matrix, 8-row groups are SBO apart, the two 8-element K halves of one K=16 MMA are LBO apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// Instruction descriptor for kind::f16, fp32 accumulate, both operands K-major.
__host__ __device__ __forceinline__ uint32_t umma_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// general form: separate A/B formats (0 = f16, 1 = bf16) and majors (0 = K-major, 1 = MN-major)
__host__ __device__ __forceinline__ uint32_t umma_idesc2(uint32_t fmt_a, uint32_t fmt_b, uint32_t a_mn, uint32_t b_mn,
                                                         uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Same instruction with the two shared-memory descriptors given as (low, high) 32-bit words: only the low word (start
// address | LBO) changes from step to step, so the issuing warp does 32-bit uniform adds instead of 64-bit ones.
__device__ __forceinline__ void umma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// zero 16 consecutive 32-bit columns of this warp's 32 TMEM lanes (accumulator clearing off the tensor pipe)
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
      ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Warp "transpose-reduce": every lane holds 32 partial values v[0..31]; on return lane l holds the
// warp-wide total of v[l] (31 shuffles instead of 160).
// ---------------------------------------------------------------------------------------------
template <int HALF>
__device__ __forceinline__ void warp_xreduce_step(float* v, int lane) {
  const bool hi = (lane & HALF) != 0;
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    const float send = hi ? v[i] : v[i + HALF];
    const float keep = hi ? v[i + HALF] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
  }
}
__device__ __forceinline__ float warp_xreduce32(float* v, int lane) {
  warp_xreduce_step<16>(v, lane);
  warp_xreduce_step<8>(v, lane);
  warp_xreduce_step<4>(v, lane);
  warp_xreduce_step<2>(v, lane);
  warp_xreduce_step<1>(v, lane);
  return v[0];
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#define SEUNET_CUDA_CHECK(expr)                                                         \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) { seunet_set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); return 1; } \
  } while (0)

void seunet_set_error(const char* fmt, ...);
