// xu_probe.cu — how much of the MUFU (exp2) rate can 16 warps per SM sustain with the attention kernel's per-element
// instruction mix, and what do the per-tile synchronisation steps cost?  Standalone micro-benchmark (no library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/xu_probe tools/xu_probe.cu && /tmp/xu_probe
// Each softmax-like warp processes "tiles" of 32 values per thread: FFMA2 scale, ex2, pack to half2, packed max,
// 4 x st.shared.v4 — optionally followed by the proxy fence, a 64-thread named-barrier vote and an mbarrier arrive.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t pack_f2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void ffma2(float& x, float& y, uint32_t a, uint32_t b, uint64_t s, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pack_f2(a, b)), "l"(s), "l"(c));
  uint32_t lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(d));
  x = __uint_as_float(lo);
  y = __uint_as_float(hi);
}
__device__ __forceinline__ uint32_t max3(uint32_t m, uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(a));
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(b));
  return r;
}

// 0: math only  1: + st.shared  2: + fence.proxy.async  3: + pair vote (bar.red)  4: + mbarrier arrive
// 5: math + tcgen05.ld x32 per tile   6: math + tcgen05.ld x32 + tcgen05.st x16 (P to TMEM instead of smem) + wait::st
// 7: mode 6 + per-warp mbarrier arrive (the whole per-tile protocol of a "P in TMEM, no vote" softmax warp)
template <int kMode>
__global__ void __launch_bounds__(320, 2) probe(int tiles, float scale, float* sink) {
  extern __shared__ uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(8));
  if (kMode >= 5 && warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = kMode >= 5 ? tmem_slot : 0u;  // (modes >= 5 allocate TMEM)
  if (warp < 2) {
    if (kMode >= 5) {
      // keep the allocation alive until the math warps are done: they signal through a named barrier at the end
      asm volatile("bar.sync 9, 320;" ::: "memory");
      if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
    }
    return;
  }
  const int quarter = warp & 3;
  const int r = quarter * 32 + lane;
  const uint32_t p_row = (uint32_t)__cvta_generic_to_shared(smem) + r * 128u + ((warp - 2) >> 2) * 64u;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-0.01f * (i + lane));
  const uint64_t s2 = pack_f2(__float_as_uint(scale), __float_as_uint(scale));
  uint64_t m2 = pack_f2(__float_as_uint(-1.0f), __float_as_uint(-1.0f));
  uint32_t pm = 0;
  const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16) + ((warp - 2) >> 2) * 32u;
  for (int t = 0; t < tiles; ++t) {
    uint32_t pk[16];
    if (kMode == 9 || kMode == 10) {  // half the returned registers: x16, or 32 columns packed 2 x 16 bit per register
      uint32_t w[16];
      if (kMode == 9)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
              "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
            : "r"(t_row + 64u * (t & 1)));
      else
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.pack::16b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
              "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
            : "r"(t_row + 64u * (t & 1)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[2 * i] = (w[i] & 0x007fffffu) | 0xbd000000u;
        v[2 * i + 1] = ((w[i] >> 3) & 0x007fffffu) | 0xbd000000u;
      }
    } else if (kMode >= 5) {
      uint32_t w[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
            "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]),
            "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]),
            "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
          : "r"(t_row + 64u * (t & 1)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (w[i] & 0x007fffffu) | 0xbd000000u;  // small negative floats, data dependent
    }
    if (kMode == 8) {  // TMEM read bandwidth alone
#pragma unroll
      for (int i = 0; i < 32; ++i) pm ^= v[i];
      continue;
    }
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float a, b;
      ffma2(a, b, v[i], v[i + 1], s2, m2);
      __half2 h = __floats2half2_rn(ex2(a), ex2(b));
      pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) pm = max3(pm, pk[i], pk[i + 1]);
    m2 ^= static_cast<uint64_t>(pm & 1u);  // loop-carried: nothing can be hoisted, no dynamic register indexing
    if (kMode >= 6) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
          ::"r"(t_row + 128u + 16u * (t & 1)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]),
          "r"(pk[6]), "r"(pk[7]), "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]),
          "r"(pk[15])
          : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (kMode >= 7) {
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
      }
    }
    if (kMode >= 1 && kMode <= 4) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + ((c ^ (r & 3)) << 4)), "r"(pk[4 * c]),
                     "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                     : "memory");
    }
    if (kMode == 3 || kMode == 4) {
      uint32_t out;
      asm volatile(
          "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, 64, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(out)
          : "r"((uint32_t)(pm == 0x7c007c00u)), "r"(quarter + 1)
          : "memory");
      if (out) m2 = pack_f2(__float_as_uint(-2.0f), __float_as_uint(-2.0f));
    }
    if (kMode >= 2 && kMode <= 4) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (kMode == 4) {
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    }
  }
  if (pm == 0x12345678u) sink[threadIdx.x] = __uint_as_float(v[3]);
  if (kMode >= 5) asm volatile("bar.sync 9, 320;" ::: "memory");
}

template <int kMode>
static void run(const char* what, float* sink) {
  const int tiles = 4096, grid = 148 * 2;
  cudaFuncSetAttribute(probe<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<kMode><<<grid, 320, 64 * 1024>>>(64, 0.1f, sink);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<kMode><<<grid, 320, 64 * 1024>>>(tiles, 0.1f, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double exps = double(grid) * 256 * 32 * tiles;
  printf("%-44s %8.3f ms  %6.2f Texp/s  (%s)\n", what, ms, exps / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* sink;
  cudaMalloc(&sink, 4096);
  run<0>("math only (FFMA2, ex2, pack, max)", sink);
  run<1>("+ 4 x st.shared.v4", sink);
  run<2>("+ fence.proxy.async", sink);
  run<3>("+ 64-thread bar.red vote", sink);
  run<4>("+ per-warp mbarrier arrive", sink);
  run<5>("math + tcgen05.ld x32 / tile", sink);
  run<6>("math + ld x32 + tcgen05.st x16 (P->TMEM)", sink);
  run<7>("mode 6 + per-warp mbarrier arrive", sink);
  run<8>("tcgen05.ld x32 only (4 B/'exp')", sink);
  run<9>("math + tcgen05.ld x16 / tile (half the regs)", sink);
  run<10>("math + tcgen05.ld x32.pack::16b / tile", sink);
  return 0;
}
