// Microbenchmark (round 2): TMEM read / write bandwidth per SM and the MUFU ex2 rate next to it - the three resources the
// spatial-attention softmax warps contend for.  One CTA, W warps (4 or 8), each warp loops over tcgen05.ld 32x32b.x32 of its
// own 32-lane quarter.  Prints cycles per warp-level instruction and aggregate bytes / clk / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dino-video-summarization-transformer_b200/csrc/ptx.cuh"
using namespace dvst;

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ float ex2a(float x) {
  float r;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// mode 0: ld.x32 + wait each; 1: two ld.x32 in flight; 2: st.x16 + wait::st; 3: 32 ex2 per iteration; 4: ld.x32 + 32 ex2
// (the softmax inner loop); 5: ld + 32 ex2 + 16 cvt-pack + st.x16 (the full exp pass)
__global__ void __launch_bounds__(256, 1) bench(int mode, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<1>(smem_u32(&slot), 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t v[32], v2[32], pk[16];
  float acc = threadIdx.x * 1e-3f;
  for (int j = 0; j < 16; ++j) pk[j] = threadIdx.x + j;
  for (int j = 0; j < 32; ++j) v[j] = v2[j] = 0;
  __syncthreads();
  const long long t0 = clock64();
  float xs[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) xs[j] = -1.0f - j * 0.01f - threadIdx.x * 1e-4f;
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (i & 3) * 32;
    if (mode == 0) {
      tmem_ld_32x32(base + col, v);
      tmem_ld_wait();
      acc += __uint_as_float(v[0] ^ v[13] ^ v[31]);
    } else if (mode == 1) {
      tmem_ld_32x32(base + col, v);
      tmem_ld_32x32(base + col + 128, v2);
      tmem_ld_wait();
      acc += __uint_as_float(v[0] ^ v[31] ^ v2[0] ^ v2[31]);
    } else if (mode == 2) {
      tmem_st_32x16(base + col, pk);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (mode == 3) {
#pragma unroll
      for (int j = 0; j < 32; ++j) xs[j] = ex2a(xs[j]) - 1.5f;   // 32 independent chains: MUFU throughput
    } else if (mode == 4) {
      tmem_ld_32x32(base + col, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) xs[j] += ex2a(__uint_as_float(v[j]) * 1e-30f);
    } else if (mode == 5) {
      tmem_ld_32x32(base + col, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a = ex2a(__uint_as_float(v[2 * j]) * 1e-30f), b = ex2a(__uint_as_float(v[2 * j + 1]) * 1e-30f);
        xs[j] += a + b;
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st_32x16(base + (i & 7) * 16, pk);
    } else if (mode == 6) {
      // software-pipelined: next chunk's load in flight while this chunk's 32 ex2 run (the kernel's exp pass)
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v2[j] = v[j];
      tmem_ld_32x32(base + col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) xs[j] += ex2a(__uint_as_float(v2[j]) * 1e-30f);
    } else if (mode == 8) {
      // packed half-precision exponentials: 16 x ex2.approx.f16x2 = 32 exponentials per iteration
      uint32_t* hx = reinterpret_cast<uint32_t*>(xs);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint32_t r;
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(hx[j]));
        hx[j] = r ^ 0x80008000u;   // keep the values bounded: alternate the sign
      }
    } else if (mode == 9) {
      uint32_t* hx = reinterpret_cast<uint32_t*>(xs);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint32_t r;
        asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(hx[j]));
        hx[j] = r ^ 0x80008000u;
      }
    } else if (mode == 10) {
      // FMA-pipe exp2 (Cody-Waite + cubic, packed fp32x2): 32 exponentials per iteration
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint64_t x2, r2;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(xs[2 * j]), "f"(xs[2 * j + 1]));
        // n = round(x) via the magic constant, f = x - n in [-0.5, 0.5]
        const uint64_t magic = 0x4B4000004B400000ull;   // 12582912.0f twice
        uint64_t t2, f2;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t2) : "l"(x2), "l"(magic));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(f2) : "l"(t2), "l"(magic));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(f2) : "l"(x2), "l"(f2));
        const uint64_t c3 = 0x3D6359A43D6359A4ull, c2 = 0x3E75FDF03E75FDF0ull, c1 = 0x3F3172183F317218ull, c0 = 0x3F8000003F800000ull;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r2) : "l"(c3), "l"(f2), "l"(c2));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r2) : "l"(r2), "l"(f2), "l"(c1));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r2) : "l"(r2), "l"(f2), "l"(c0));
        uint32_t lo, hi, tl, th;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(r2));
        asm("mov.b64 {%0, %1}, %2;" : "=r"(tl), "=r"(th) : "l"(t2));
        lo += tl << 23; hi += th << 23;
        xs[2 * j] = __uint_as_float(lo) * 1e-3f - 1.0f;
        xs[2 * j + 1] = __uint_as_float(hi) * 1e-3f - 1.0f;
      }
    } else {
      // x16 loads
      uint32_t w[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
            "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "r"(base + col) : "memory");
      tmem_ld_wait();
      acc += __uint_as_float(w[0] ^ w[15]);
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) acc += xs[j];
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(slot, 512); }
}

int main() {
  long long* d_c; float* d_s;
  cudaMalloc(&d_c, 8); cudaMalloc(&d_s, 4096);
  const char* names[] = {"ld.x32 + wait", "2 x ld.x32 in flight", "st.x16 + wait", "32 x ex2", "ld.x32 + 32 ex2",
                         "ld.x32 + 32 ex2 + pack + st.x16", "pipelined ld.x32 | 32 ex2", "ld.x16 + wait", "16 x ex2.f16x2 (32 exp)", "16 x ex2.bf16x2 (32 exp)",
                         "32 x FMA-pipe exp2 (packed)"};
  const int iters = 4096;
  for (int warps = 1; warps <= 8; warps *= 2) {
    for (int mode = 0; mode < 11; ++mode) {
      bench<<<1, warps * 32, 0>>>(mode, iters, d_c, d_s);  // warm
      bench<<<1, warps * 32, 0>>>(mode, iters, d_c, d_s);
      long long c = 0;
      cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      const double per = static_cast<double>(c) / iters;
      double bytes = 0;
      if (mode == 0 || mode == 4 || mode == 6) bytes = 32.0 * 32 * 4;
      if (mode == 7) bytes = 32.0 * 16 * 4;
      if (mode == 1) bytes = 2 * 32.0 * 32 * 4;
      if (mode == 2) bytes = 32.0 * 16 * 4;
      if (mode == 5) bytes = 32.0 * 32 * 4;
      printf("warps %d  %-34s %8.1f clk/iter/warp   TMEM %7.1f B/clk/SM   ex2 %6.2f /clk/SM  (%s)\n", warps, names[mode], per,
             bytes * warps / per, ((mode >= 3 && mode <= 6) || mode >= 8 ? 32.0 * 32 * warps / per : 0.0), cudaGetErrorString(e));
    }
  }
  return 0;
}
