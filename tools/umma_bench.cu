// Micro-benchmark: issue cost of tcgen05.mma (kind::f16, K = 16) on sm_100a as a function of N, operand source
// (SS: A and B in shared memory, TS: A in TMEM), shared-memory layout (no swizzle / 128-byte swizzle), and CTA group.
// It answers one question for the implicit-GEMM conv (csrc/conv_tc.cu): is the ~64-cycle floor of an M=128 SS-mode
// instruction a property of the no-swizzle halo-tile layout, of SS mode, or of the instruction itself?
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/umma_bench tools/umma_bench.cu
// run:   tools/umma_bench <group>      group in {ss, layout, sw128, ts, m64, cg2, cp, mix, sync, issue, dswitch}
//
// Every CTA (one per SM, 148) runs the same instruction stream; the elected thread issues `iters` MMAs back to back,
// commits, waits for the commit, and reports clock64() deltas.  cycles/MMA = (T(iters=2052) - T(iters=540)) / 1512,
// the median over CTAs.  Operand data are zeros: timing of the tensor pipe does not depend on values.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); fflush(stdout); exit(2); } } while (0)

struct Args {
  uint32_t idesc;
  uint32_t a_lo, a_hi, b_lo, b_hi;     // descriptor words relative to the dynamic smem base (address part added in-kernel)
  uint32_t a_delta[16]; int a_period;  // cyclic start-address offsets of A (16-byte units)
  uint32_t b_delta[16]; int b_period;
  uint32_t d_stride; int d_period;     // accumulator rotation (columns)
  // expanded by the host to one unrolled block of kBlock instructions (periods 1, 4, 9 all divide 36)
  uint32_t a_tab[36], b_tab[36], d_tab[36];
  int iters;                           // multiple of 36
  int plane;                           // "sync" group: MMAs per emulated input plane (divides 36), 0 = off
  int nwait, ncommit;                  // per plane: try_wait on already-complete barriers / tcgen05.commit to a dummy barrier
  int ts;                              // A operand from TMEM
  int cp;                              // benchmark tcgen05.cp instead of mma (128x256b, 4 KB per instruction)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile("{\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, 1;\n\t}" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint32_t lo, uint32_t hi) {
  asm volatile("{\n\t.reg .b64 ds;\n\tmov.b64 ds, {%1, %2};\n\ttcgen05.cp.cta_group::1.128x256b [%0], ds;\n\t}" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
  if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int CG>
__global__ void __launch_bounds__(128, 1) bench_kernel(const __grid_constant__ Args a, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_store;
  __shared__ __align__(8) uint64_t dummy_bar[2];   // [0]: never completes (commit sink), [1]: fresh (waits on parity 1 return at once)
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t bar = smem_u32(&bar_store);
  // zero the operand area (generic proxy), make it visible to the async proxy
  for (uint32_t i = threadIdx.x; i < (200u * 1024u) / 16u; i += blockDim.x)
    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(base + 16u * i), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); mbar_init(smem_u32(&dummy_bar[0]), 1000000u); mbar_init(smem_u32(&dummy_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t a_lo0 = a.a_lo + (base >> 4), b_lo0 = a.b_lo + ((base + 100u * 1024u) >> 4);
    const uint32_t a_tmem = tmem + 448u;   // TS mode: A lives in the last TMEM columns (M=128 x K=16 f16 = 8 columns)
    const uint32_t a_hi = a.a_hi, b_hi = a.b_hi, idesc = a.idesc;
    const int mode = a.cp ? 2 : (a.ts ? 1 : 0);
    long long t0 = clock64();
    // one unrolled block of 36 instructions per trip: per instruction only three adds with constant-bank operands
    if (mode == 0 && a.plane > 0) {
      // emulated conv issuer: per input plane `plane` MMAs, then the per-plane synchronisation the real kernel does
      const uint32_t sink = smem_u32(&dummy_bar[0]), fresh = smem_u32(&dummy_bar[1]);
      const int plane = a.plane, nwait = a.nwait, ncommit = a.ncommit;
#pragma unroll 1
      for (int i = 0; i < a.iters; i += 36) {
#pragma unroll 1
        for (int k0 = 0; k0 < 36; k0 += plane) {
          for (int w = 0; w < nwait; ++w) mbar_wait(fresh, 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (plane == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss<CG>(tmem + a.d_tab[k], a_lo0 + a.a_tab[k], a_hi, b_lo0 + a.b_tab[k], b_hi, idesc);
          } else if (plane == 9) {
#pragma unroll
            for (int k = 0; k < 9; ++k) mma_ss<CG>(tmem + a.d_tab[k], a_lo0 + a.a_tab[k], a_hi, b_lo0 + a.b_tab[k], b_hi, idesc);
          } else if (plane == 18) {
#pragma unroll
            for (int k = 0; k < 18; ++k) mma_ss<CG>(tmem + a.d_tab[k], a_lo0 + a.a_tab[k], a_hi, b_lo0 + a.b_tab[k], b_hi, idesc);
          } else {
#pragma unroll
            for (int k = 0; k < 36; ++k) mma_ss<CG>(tmem + a.d_tab[k], a_lo0 + a.a_tab[k], a_hi, b_lo0 + a.b_tab[k], b_hi, idesc);
          }
          for (int c = 0; c < ncommit; ++c) commit<CG>(sink);
        }
      }
    } else if (mode == 0) {
#pragma unroll 1
      for (int i = 0; i < a.iters; i += 36) {
#pragma unroll
        for (int k = 0; k < 36; ++k) mma_ss<CG>(tmem + a.d_tab[k], a_lo0 + a.a_tab[k], a_hi, b_lo0 + a.b_tab[k], b_hi, idesc);
      }
    } else if (mode == 1) {
#pragma unroll 1
      for (int i = 0; i < a.iters; i += 36) {
#pragma unroll
        for (int k = 0; k < 36; ++k) mma_ts(tmem + a.d_tab[k], a_tmem, b_lo0 + a.b_tab[k], b_hi, idesc);
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < a.iters; i += 36) {
#pragma unroll
        for (int k = 0; k < 36; ++k) cp_128x256b(tmem + (uint32_t)((k & 7) * 8), a_lo0 + a.a_tab[k], a_hi);
      }
    }
    commit<CG>(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[blockIdx.x / CG] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (threadIdx.x < 32) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// "issue" group: how fast can ONE warp feed the tensor pipe?  Mirrors the conv kernel's issuer: the whole warp runs the loop
// converged (warp-uniform control flow, descriptors in uniform registers) and one elected lane issues.  Variants:
//   style 0: rolled (kh, kw) tap nest with register increments, J k-blocks unrolled (round-1 issue_taps3)
//   style 1: one input plane (9 taps x J blocks) fully unrolled with IMMEDIATE descriptor offsets
// issuer warp id 1 or 5 of 6 warps; the other warps optionally run an FMA/ST loop ("noise", like the epilogue warps).
// The warp scheduler prefers the highest warp id of an SM sub-partition (warp % 4), so an issuer at warp 1 competes with warp 5.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void mma_lohi(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
struct IssueArgs { uint32_t idesc, a_lo, a_hi, b_lo, b_hi; uint32_t kh_step, kw_step, j_step, b_step; int planes, style, J, issuer_warp, noise, nwait, ncommit; };

template <int J>
__device__ __forceinline__ void plane_rolled(uint32_t dcol, uint32_t a0, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t kh_step, uint32_t kw_step, uint32_t j_step, uint32_t b_step) {
  uint32_t a_row = a0;
#pragma unroll 1
  for (int kh = 0; kh < 3; ++kh) {
    uint32_t a_tap = a_row;
#pragma unroll 1
    for (int kw = 0; kw < 3; ++kw) {
      uint32_t a_lo = a_tap;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (elect_one()) mma_lohi(dcol, a_lo, a_hi, b_lo, b_hi, idesc);
        a_lo += j_step; b_lo += b_step;
      }
      a_tap += kw_step;
    }
    a_row += kh_step;
  }
}
// immediates: dil 1 halo tile (lineW = 10, HV = 180), B step = 2 * 3 * COUT rows with COUT = N / 3
template <int J, int N>
__device__ __forceinline__ void plane_unrolled(uint32_t dcol, uint32_t a0, uint32_t a_hi, uint32_t b0, uint32_t b_hi, uint32_t idesc) {
  const bool lead = elect_one();
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t ao = (uint32_t)(kh * 10 + kw + j * 2 * 180), bo = (uint32_t)(((kh * 3 + kw) * J + j) * 2 * N);
        if (lead) mma_lohi(dcol, a0 + ao, a_hi, b0 + bo, b_hi, idesc);
      }
}

__global__ void __launch_bounds__(192, 1) issue_kernel(const __grid_constant__ IssueArgs a, long long* __restrict__ out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (uint32_t i = threadIdx.x; i < (200u * 1024u) / 16u; i += blockDim.x)
    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(base + 16u * i), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1000000u); mbar_init(smem_u32(&bars[2]), 1);
    stop_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == a.issuer_warp) {
    const uint32_t a0 = a.a_lo + (base >> 4), b0 = a.b_lo + ((base + 100u * 1024u) >> 4);
    const uint32_t a_hi = a.a_hi, b_hi = a.b_hi, idesc = a.idesc;
    const uint32_t kh_step = a.kh_step, kw_step = a.kw_step, j_step = a.j_step, b_step = a.b_step;
    const uint32_t sinkb = smem_u32(&bars[1]), fresh = smem_u32(&bars[2]);
    const int style = a.style, J = a.J, nwait = a.nwait, ncommit = a.ncommit;
    long long t0 = clock64();
#pragma unroll 1
    for (int p = 0; p < a.planes; ++p) {
      for (int w = 0; w < nwait; ++w) mbar_wait(fresh, 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t dcol = tmem + (uint32_t)((p & 3) * 96);
      if (style == 0) {
        if (J == 1) plane_rolled<1>(dcol, a0, a_hi, b0, b_hi, idesc, kh_step, kw_step, j_step, b_step);
        else if (J == 2) plane_rolled<2>(dcol, a0, a_hi, b0, b_hi, idesc, kh_step, kw_step, j_step, b_step);
        else plane_rolled<4>(dcol, a0, a_hi, b0, b_hi, idesc, kh_step, kw_step, j_step, b_step);
      } else {
        if (J == 1) plane_unrolled<1, 96>(dcol, a0, a_hi, b0, b_hi, idesc);
        else if (J == 2) plane_unrolled<2, 96>(dcol, a0, a_hi, b0, b_hi, idesc);
        else plane_unrolled<4, 96>(dcol, a0, a_hi, b0, b_hi, idesc);
      }
      for (int c = 0; c < ncommit; ++c)
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sinkb) : "memory");
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
    mbar_wait(smem_u32(&bars[0]), 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[blockIdx.x] = t1 - t0; stop_flag = 1; }
  } else if (a.noise) {
    // epilogue-like noise: dependent FMAs + a store now and then, until the issuer is done
    float x = (float)threadIdx.x, y = 1.0001f;
    int it = 0;
    while (!stop_flag) {
#pragma unroll
      for (int k = 0; k < 32; ++k) x = fmaf(x, y, 0.5f);
      if ((++it & 15) == 0) sink[blockIdx.x * 192 + threadIdx.x] = x;
    }
    sink[blockIdx.x * 192 + threadIdx.x] = x;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
// descriptor words: lo = (addr>>4) | (LBO>>4)<<16 ; hi = (SBO>>4) | version(1<<14) | layout_type << 29
static void desc_words(uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t* lo, uint32_t* hi) {
  *lo = ((lbo >> 4) & 0x3FFFu) << 16;
  *hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}

static long long* d_out = nullptr;
static int g_dswitch_P = 0, g_dswitch_S = 0;
template <int CG>
static double run(Args a, int ctas) {
  double med[2];
  const int its[2] = {540, 2052};
  for (int k = 0; k < 36; ++k) {
    a.a_tab[k] = a.a_delta[k % a.a_period];
    a.b_tab[k] = a.b_delta[k % a.b_period];
    a.d_tab[k] = (uint32_t)(k % a.d_period) * a.d_stride;
    if (g_dswitch_P > 0) a.d_tab[k] = (uint32_t)(((k / g_dswitch_P) * g_dswitch_S) % 256);
  }
  std::vector<long long> h(ctas);
  for (int r = 0; r < 2; ++r) {
    a.iters = its[r];
    std::vector<double> best;
    for (int rep = 0; rep < 3; ++rep) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ctas * CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 202 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, bench_kernel<CG>, a, d_out));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
      std::vector<long long> s(h); std::sort(s.begin(), s.end());
      best.push_back((double)s[s.size() / 2]);
    }
    med[r] = *std::min_element(best.begin(), best.end());
  }
  return (med[1] - med[0]) / (its[1] - its[0]);
}

struct Row { std::string name; double cyc; int M, N, cg; };

int main(int argc, char** argv) {
  const std::string group = argc > 1 ? argv[1] : "ss";
  const int ctas_arg = argc > 2 ? atoi(argv[2]) : 148;
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&d_out, sizeof(long long) * 256));
  CK(cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024));
  CK(cudaFuncSetAttribute(bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024));
  CK(cudaFuncSetAttribute(bench_kernel<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int Ns[] = {16, 32, 48, 64, 96, 128, 144, 192, 256};
  auto report = [&](const char* name, int M, int N, int cg, double cyc) {
    const double ideal = (double)std::max(M / cg, 128) * N / 256.0 / 1.0;   // per-SM math floor: 128 x N x 16 in N/2 cycles
    printf("%-44s M=%3d N=%3d cg=%d  %7.1f cyc/MMA   math floor %5.1f   pipe %5.1f%%\n", name, M, N, cg, cyc, ideal, 100.0 * ideal / cyc);
    fflush(stdout);
  };
  Args a;
  auto base_args = [&](int M, int N) {
    memset(&a, 0, sizeof(a));
    a.idesc = idesc_f16(M, N);
    a.a_period = a.b_period = a.d_period = 1;
  };
  // conv-like tap walk over a halo tile [chunk][18][10][16 B]: 9 taps x 2 K-blocks
  auto conv_walk = [&](int dil) {
    const int lineW = 8 + 2 * dil, HV = (16 + 2 * dil) * lineW;
    int n = 0;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) a.a_delta[n++] = (uint32_t)(kh * dil * lineW + kw * dil);
    a.a_period = 9;
    desc_words((uint32_t)HV * 16u, (uint32_t)lineW * 16u, 0, &a.a_lo, &a.a_hi);
  };
  if (group == "ss") {
    // (1) the conv's own operand layouts: A = halo tile (SBO = 160 B, LBO = one chunk plane, taps = start offsets),
    //     B = packed weights [khalf][N rows][16 B] (SBO = 128 B, LBO = N * 16 B), same accumulator for 9 taps
    for (int N : Ns) {
      base_args(128, N);
      conv_walk(1);
      desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
      for (int i = 0; i < 9; ++i) a.b_delta[i] = (uint32_t)(i * 2 * N);
      a.b_period = 9;
      report("SS no-swizzle, conv halo-tile walk (dil 1)", 128, N, 1, run<1>(a, ctas_arg));
    }
    for (int N : {48, 96, 192}) {
      base_args(128, N);
      conv_walk(2);
      desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
      report("SS no-swizzle, conv halo-tile walk (dil 2)", 128, N, 1, run<1>(a, ctas_arg));
    }
  } else if (group == "layout") {
    // (2) same instruction, dense canonical no-swizzle A ([khalf][128 rows][16 B]: SBO = 128 B, LBO = 2048 B), fixed address
    for (int N : Ns) {
      base_args(128, N);
      desc_words(2048u, 128u, 0, &a.a_lo, &a.a_hi);
      desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
      report("SS no-swizzle, dense A, fixed address", 128, N, 1, run<1>(a, ctas_arg));
    }
    // dense A but start address shifted by 16 B (what a kw tap does): alignment effect in isolation
    for (int N : {48, 96}) {
      base_args(128, N);
      desc_words(2048u + 128u, 128u, 0, &a.a_lo, &a.a_hi);
      a.a_delta[0] = 1; a.a_period = 1;
      desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
      report("SS no-swizzle, dense A, start + 16 B", 128, N, 1, run<1>(a, ctas_arg));
    }
    // rotating accumulators (4 slots) to rule out an accumulator dependency
    for (int N : {48, 96}) {
      base_args(128, N);
      desc_words(2048u, 128u, 0, &a.a_lo, &a.a_hi);
      desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
      a.d_stride = 96; a.d_period = 4;
      report("SS no-swizzle, dense A, 4 accumulators", 128, N, 1, run<1>(a, ctas_arg));
    }
  } else if (group == "sw128") {
    // (3) 128-byte swizzle, K-major: A tile = 128 rows x 128 B (64 K elements), SBO = 1024 B; K blocks advance by 32 B
    for (int N : Ns) {
      base_args(128, N);
      desc_words(16u, 1024u, 2, &a.a_lo, &a.a_hi);
      desc_words(16u, 1024u, 2, &a.b_lo, &a.b_hi);
      for (int i = 0; i < 4; ++i) { a.a_delta[i] = (uint32_t)(2 * i); a.b_delta[i] = (uint32_t)(2 * i); }
      a.a_period = a.b_period = 4;
      report("SS swizzle-128B K-major, 4 K blocks", 128, N, 1, run<1>(a, ctas_arg));
    }
  } else if (group == "ts") {
    // (4) A from TMEM, B no-swizzle
    for (int N : Ns) {
      if (N > 256) continue;
      base_args(128, N);
      a.ts = 1;
      desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
      for (int i = 0; i < 9; ++i) a.b_delta[i] = (uint32_t)(i * 2 * N);
      a.b_period = 9;
      report("TS (A in TMEM), B no-swizzle", 128, N, 1, run<1>(a, ctas_arg));
    }
  } else if (group == "m64") {
    for (int N : Ns) {
      base_args(64, N);
      conv_walk(1);
      desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
      report("SS no-swizzle, M=64, conv walk", 64, N, 1, run<1>(a, ctas_arg));
    }
  } else if (group == "cg2") {
    // (5) CTA pair: M = 256 (128 rows of A per CTA), each CTA holds N/2 rows of B
    for (int N : {32, 64, 96, 128, 192, 256}) {
      base_args(256, N);
      conv_walk(1);
      desc_words((uint32_t)(N / 2) * 16u, 128u, 0, &a.b_lo, &a.b_hi);
      report("SS no-swizzle, cta_group::2, conv walk", 256, N, 2, run<2>(a, ctas_arg / 2));
    }
    for (int N : {32, 64, 96, 128, 192, 256}) {
      base_args(256, N);
      desc_words(16u, 1024u, 2, &a.a_lo, &a.a_hi);
      desc_words(16u, 1024u, 2, &a.b_lo, &a.b_hi);
      for (int i = 0; i < 4; ++i) { a.a_delta[i] = (uint32_t)(2 * i); a.b_delta[i] = (uint32_t)(2 * i); }
      a.a_period = a.b_period = 4;
      report("SS swizzle-128B, cta_group::2", 256, N, 2, run<2>(a, ctas_arg / 2));
    }
  } else if (group == "cp") {
    // (6) tcgen05.cp 128x256b (4 KB per instruction) from the halo tile: the cost of staging A in TMEM
    base_args(128, 64);
    a.cp = 1;
    desc_words(2048u, 128u, 0, &a.a_lo, &a.a_hi);
    printf("tcgen05.cp.128x256b (4 KB): %7.1f cyc/instr\n", run<1>(a, ctas_arg));
  } else if (group == "mix") {
    // (7) one SM alone vs all 148 (shared-resource / power effects on the cycle count)
    for (int ctas : {1, 148}) {
      for (int N : {96, 192}) {
        base_args(128, N);
        conv_walk(1);
        desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
        char nm[64]; snprintf(nm, sizeof nm, "SS conv walk, %d CTA(s)", ctas);
        report(nm, 128, N, 1, run<1>(a, ctas));
      }
    }
  } else if (group == "dswitch") {
    // (9) does moving to another accumulator cost anything?  36 MMAs in "planes" of P; each plane accumulates into columns
    //     advanced by S from the previous plane (S = 0: same accumulator, S < N: OVERLAPPING ranges as in the kd-stacked conv
    //     where plane k writes output slots k-2..k and plane k+1 writes k-1..k+1, S >= N: disjoint)
    for (int N : {48, 96, 192}) {
      for (int P : {4, 9, 18, 36}) {
        for (int S : {0, N / 3, N, 128}) {
          if (S == 128 && N > 128) continue;
          base_args(128, N);
          conv_walk(1);
          desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
          a.a_period = 9;
          // d_tab is filled below through d_stride/d_period only for uniform rotation; build it by hand here
          a.d_period = 36; a.d_stride = 0;
          char nm[96]; snprintf(nm, sizeof nm, "SS conv walk, %2d MMAs per accumulator, next accumulator +%3d cols", P, S);
          // encode: run<> expands tables from periods; override afterwards via a custom hook
          g_dswitch_P = P; g_dswitch_S = S;
          report(nm, 128, N, 1, run<1>(a, ctas_arg));
          g_dswitch_P = 0;
        }
      }
    }
  } else if (group == "issue") {
    CK(cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024));
    float* d_sink; CK(cudaMalloc(&d_sink, sizeof(float) * 192 * 256));
    std::vector<long long> h(ctas_arg);
    auto run_issue = [&](IssueArgs ia) {
      double med[2]; const int pl[2] = {40, 160};
      for (int r = 0; r < 2; ++r) {
        ia.planes = pl[r];
        double best = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
          issue_kernel<<<ctas_arg, 192, 202 * 1024>>>(ia, d_out, d_sink);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * ctas_arg, cudaMemcpyDeviceToHost));
          std::vector<long long> sd(h); std::sort(sd.begin(), sd.end());
          best = std::min(best, (double)sd[sd.size() / 2]);
        }
        med[r] = best;
      }
      return (med[1] - med[0]) / ((pl[1] - pl[0]) * 9.0 * ia.J);
    };
    for (int N : {48, 96}) {
      for (int J : {1, 2}) {   // (J = 4 would need a weight image beyond the 200 KB operand area of this bench)
        for (int style : {0, 1}) {
          for (int iw : {1, 5}) {
            for (int noise : {0, 1}) {
              for (int sync : {0, 1}) {
                IssueArgs ia; memset(&ia, 0, sizeof ia);
                ia.idesc = idesc_f16(128, N);
                desc_words(180u * 16u, 160u, 0, &ia.a_lo, &ia.a_hi);
                desc_words(96u * 16u, 128u, 0, &ia.b_lo, &ia.b_hi);      // weight image rows as for COUT = 32 (N = 96 columns)
                ia.kh_step = 10; ia.kw_step = 1; ia.j_step = 2 * 180; ia.b_step = 2 * 96;
                ia.style = style; ia.J = J; ia.issuer_warp = iw; ia.noise = noise; ia.nwait = sync; ia.ncommit = sync;
                const double cyc = run_issue(ia);
                printf("N=%3d J=%d %-9s issuer warp %d noise %d wait+commit/plane %d : %6.1f cyc/MMA (math floor %4.1f)\n", N, J,
                       style ? "unrolled" : "rolled", iw, noise, sync, cyc, N / 2.0);
                fflush(stdout);
              }
            }
          }
        }
      }
    }
  } else if (group == "sync") {
    // (8) how much of the issuer's per-plane synchronisation (mbarrier try_wait on a complete barrier, tcgen05.commit) is
    //     hidden behind the tensor pipe's instruction queue: MMAs per plane x {no sync, commit, wait+commit, 2 x (wait+commit)}
    for (int N : {48, 96, 192}) {
      for (int plane : {4, 9, 18, 36}) {
        for (int lvl = 0; lvl < 4; ++lvl) {
          base_args(128, N);
          conv_walk(1);
          desc_words((uint32_t)N * 16u, 128u, 0, &a.b_lo, &a.b_hi);
          a.plane = plane; a.nwait = lvl >= 2 ? lvl - 1 : 0; a.ncommit = lvl >= 1 ? (lvl == 3 ? 2 : 1) : 0;
          char nm[96]; snprintf(nm, sizeof nm, "SS conv walk, %2d MMAs/plane, %d wait + %d commit", plane, a.nwait, a.ncommit);
          report(nm, 128, N, 1, run<1>(a, ctas_arg));
        }
      }
    }
  } else {
    printf("unknown group %s\n", group.c_str());
    return 1;
  }
  return 0;
}
