// Device helpers shared by the tcgen05 attention kernels (attention_sm100.cu, attention_split_sm100.cu): packed fp32x2
// arithmetic, the MN-major V descriptor, the A-from-TMEM MMA form, TMEM stores, the 3-D TMA store and the launch parameters
// of the spatial kernels.
#pragma once
#include "dvst_internal.h"
#include "ptx.cuh"

namespace dvst {
namespace attn {

constexpr int HD = 64;
constexpr int DMODEL = 768;
constexpr int TILE_BYTES = 128 * 128;   // one 128-row tile of 128-byte rows
constexpr int O_COL = 128;              // TMEM column of the O accumulator inside a tile's 256-column region
constexpr int NUM_THREADS = 384;

struct SpParams {
  const __nv_bfloat16* qkv;
  long long ld;
  __nv_bfloat16* out;   // space-only: CLS rows are ordinary bf16 output rows
  long long ldo;
  float* cls_out;       // divided: [B*T, 768] fp32 CLS attention outputs (averaged over frames by the caller)
  int mode;             // 1 divided-spatial, 3 space-only
  int S, N, NPAD, T;
  int n_prob, heads;
  long long cls_base;
  int rows_a, rows_b;   // patch rows of tile A / tile B (what the TMA stores cover)
  int* fix_count;       // one-pass softmax: number of rows to recompute ...
  int2* fix_list;       // ... and their (problem, row) pairs (capacity n_prob * S)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// sm_100 packed / 3-input fp32 forms: half the issue slots for the softmax's scaling, row-sum and row-max work
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// V as the B operand of O = P V: B is [N = 64 head dims] x [K = 16 keys]; in shared memory the keys are rows of 128
// bytes (64 dims) with the TMA 128-byte swizzle, i.e. the canonical MN-major SWIZZLE_128B layout
// ((8 chunks of 16 B along MN), (8 keys, k)) : ((1, LBO), (128 B, SBO)) with one MN atom (LBO unused) and SBO = 1024 B
// between 8-key groups.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // leading byte offset (single MN atom: unused)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // stride byte offset: next group of 8 keys
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version 1 (sm_100)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

}  // namespace attn
}  // namespace dvst
