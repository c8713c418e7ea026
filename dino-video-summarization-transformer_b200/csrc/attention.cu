// Fused softmax(q k^T / 8) v for the TimeSformer token layout (models/timesformer.py:68-87 as called from
// Block.forward :124-185), reading q/k/v straight out of the packed qkv GEMM output [rows, 2304] bf16 in the resident
// (b, n, t) row order - none of the reference's rearrange / permute / cat copies are materialised:
//
//   TEMPORAL : one sequence per (b, n)  = the T consecutive rows  (b*N + n)*T + t            (:130-131)
//   SPATIAL  : one sequence per (b, t)  = CLS row (cls_base + b) followed by the N rows (b*N + n)*T + t, i.e. a
//              stride-T gather; the CLS query's output goes to cls_out[b, t, :] (fp32) for the mean over frames (:140-164)
//   JOINT    : one sequence per b       = CLS row followed by the N*T rows b*N*T + i          (:120-123); sequences
//              that do not fit shared memory use the K/V-streaming kernel
//
// Flash-style: K/V of one (sequence, head) are staged once in shared memory (XOR-swizzled 16-byte chunks), each warp
// owns 16-query tiles, scores never leave registers (online softmax over 64-key blocks, exp2 with folded scale).
// Math runs on mma.sync m16n8k16 bf16 (fp32 accumulate); attention is ~3.4 % of the path's FLOPs.
#include "dvst_internal.h"
#include "ptx.cuh"
#include "mma_sync.cuh"

namespace dvst {

namespace {

using namespace mma;

constexpr int HD = 64;     // head dim
constexpr int DMODEL = 768;

struct AttnParams {
  const __nv_bfloat16* qkv;  // [rows, ld]
  long long ld;              // row stride of qkv (2304, or 4608 when hi/lo pairs are stored)
  __nv_bfloat16* out;        // [rows, ldo]
  long long ldo;
  float* cls_out;            // SPATIAL: [B, T, 768] fp32
  int mode;                  // 0 temporal, 1 spatial, 2 joint, 3 space-only
  int S;                     // sequence length
  int n_seq;                 // number of sequences
  int T, N;
  long long cls_base;        // row index of the first CLS row
  int heads;
  int split;                 // also write the bf16 rounding residual (lo) at column + lo_off
  long long lo_off;
};

__device__ __forceinline__ long long row_of(const AttnParams& p, int seq, int i) {
  if (p.mode == 0) return static_cast<long long>(seq) * p.T + i;
  if (p.mode == 1) {
    const int b = seq / p.T, t = seq - b * p.T;
    return i == 0 ? p.cls_base + b : (static_cast<long long>(b) * p.N + (i - 1)) * p.T + t;
  }
  if (p.mode == 2) return i == 0 ? p.cls_base + seq : static_cast<long long>(seq) * p.N * p.T + (i - 1);
  {  // mode 3: space-only, sequence (b, t) with its own CLS row cls_base + b*T + t
    const int b = seq / p.T, t = seq - b * p.T;
    return i == 0 ? p.cls_base + seq : (static_cast<long long>(b) * p.N + (i - 1)) * p.T + t;
  }
}

// bf16x2 rounding residual of (a, b) given their packed bf16 rounding `hi`
__device__ __forceinline__ uint32_t pack_lo2(float a, float b, uint32_t hi) {
  const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&hi);
  return pack2(a - __low2float(hb), b - __high2float(hb));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// One 64-key block of one 16-query tile.  K / V rows of the block live at `sKb + row * pitch` (row = key - key0, 128
// bytes per row, 16-byte chunks XOR-swizzled with row & 7).  qa = Q A-fragments (4 k-steps).  Carries the online-softmax
// state (m, l un-reduced over the quad, un-normalised o).
// SPLIT: q/k/v come as bf16 hi/lo pairs (qb = Q lo fragments, K / V lo rows at + lo_delta); every product is evaluated
// as hi*hi + hi*lo + lo*hi and P is split the same way, ~16 mantissa bits end to end.
// MAXP: compile-time bound on the 16-key tile pairs of a block (4 = the general 64-key block; 1 / 2 for sequences of at
// most 16 / 32 keys - temporal attention - where the general form would run its softmax over 48 / 32 padded keys).
template <bool SPLIT, int MAXP = 4>
__device__ __forceinline__ void attn_block(const uint32_t (&qa)[4][4], const uint32_t (&qb)[4][4], uint32_t sKb,
                                           uint32_t sVb, uint32_t lo_delta, int pitch, int key0, int S,
                                           int ntile_pairs, bool first, int lane, float (&o)[8][4], float& m0,
                                           float& m1, float& l0, float& l1) {
  const int tig = lane & 3;
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  const int mat = lane >> 3, r = lane & 7;
  constexpr int NT = 2 * MAXP;
  float s[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
  // ---- S = Q K^T : k-step outer so the accumulator chains are independent ----
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < MAXP; ++np) {
      if (np < ntile_pairs) {
        const int row = np * 16 + (mat >> 1) * 8 + r;
        const int ch = kk * 2 + (mat & 1);
        uint32_t b0, b1, b2, b3;
        const uint32_t kaddr = sKb + row * pitch + ((ch ^ (row & 7)) << 4);
        ldsm_x4(kaddr, b0, b1, b2, b3);
        mma_bf16(s[2 * np], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b0, b1);
        mma_bf16(s[2 * np + 1], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b2, b3);
        if (SPLIT) {
          mma_bf16(s[2 * np], qb[kk][0], qb[kk][1], qb[kk][2], qb[kk][3], b0, b1);      // q_lo * k_hi
          mma_bf16(s[2 * np + 1], qb[kk][0], qb[kk][1], qb[kk][2], qb[kk][3], b2, b3);
          ldsm_x4(kaddr + lo_delta, b0, b1, b2, b3);
          mma_bf16(s[2 * np], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b0, b1);      // q_hi * k_lo
          mma_bf16(s[2 * np + 1], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b2, b3);
        }
      }
    }
  }
  // ---- mask + online softmax ----
  float mx0 = -INFINITY, mx1 = -INFINITY;
  if (key0 + 16 * MAXP > S) {  // only the last key block can hold padded keys
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c = key0 + nt * 8 + tig * 2;
      if (c >= S) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (c + 1 >= S) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every block holds >= 1 valid key
  const float c0 = ex2_approx((m0 - mn0) * scale_log2), c1 = ex2_approx((m1 - mn1) * scale_log2);
  m0 = mn0; m1 = mn1;
  const float off0 = mn0 * scale_log2, off1 = mn1 * scale_log2;
  float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = ex2_approx(fmaf(s[nt][0], scale_log2, -off0));
    s[nt][1] = ex2_approx(fmaf(s[nt][1], scale_log2, -off0));
    s[nt][2] = ex2_approx(fmaf(s[nt][2], scale_log2, -off1));
    s[nt][3] = ex2_approx(fmaf(s[nt][3], scale_log2, -off1));
    rs0 += s[nt][0] + s[nt][1];
    rs1 += s[nt][2] + s[nt][3];
  }
  l0 = l0 * c0 + rs0;
  l1 = l1 * c1 + rs1;
  if (!first) {
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
  }
  // ---- O += P V ----
#pragma unroll
  for (int kk = 0; kk < MAXP; ++kk) {
    if (kk < ntile_pairs) {
      const uint32_t a0 = pack2(s[2 * kk][0], s[2 * kk][1]), a1 = pack2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
      if (SPLIT) {
        e0 = pack_lo2(s[2 * kk][0], s[2 * kk][1], a0); e1 = pack_lo2(s[2 * kk][2], s[2 * kk][3], a1);
        e2 = pack_lo2(s[2 * kk + 1][0], s[2 * kk + 1][1], a2); e3 = pack_lo2(s[2 * kk + 1][2], s[2 * kk + 1][3], a3);
      }
      const int row = kk * 16 + (mat & 1) * 8 + r;
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        const int ch = dp * 2 + (mat >> 1);
        uint32_t b0, b1, b2, b3;
        const uint32_t vaddr = sVb + row * pitch + ((ch ^ (row & 7)) << 4);
        ldsm_x4_t(vaddr, b0, b1, b2, b3);
        mma_bf16(o[2 * dp], a0, a1, a2, a3, b0, b1);
        mma_bf16(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
        if (SPLIT) {
          mma_bf16(o[2 * dp], e0, e1, e2, e3, b0, b1);          // p_lo * v_hi
          mma_bf16(o[2 * dp + 1], e0, e1, e2, e3, b2, b3);
          ldsm_x4_t(vaddr + lo_delta, b0, b1, b2, b3);
          mma_bf16(o[2 * dp], a0, a1, a2, a3, b0, b1);          // p_hi * v_lo
          mma_bf16(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
        }
      }
    }
  }
}

// One 16-query tile against all S keys staged in shared memory (rows at sK + key * pitch).  Leaves the UN-normalised
// output in o[8][4] and the quad-reduced row sums in l0 / l1.
template <bool SPLIT, int MAXP = 4>
__device__ __forceinline__ void attn_tile(const uint32_t (&qa)[4][4], const uint32_t (&qb)[4][4], uint32_t sK,
                                          uint32_t sV, uint32_t lo_delta, int pitch, int S, int S_pad, int lane,
                                          float (&o)[8][4], float& l0, float& l1) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY;
  l0 = 0.f; l1 = 0.f;
  const int n_kblocks = (S + 63) >> 6;
  for (int kb = 0; kb < n_kblocks; ++kb) {
    const int key0 = kb * 64;
    attn_block<SPLIT, MAXP>(qa, qb, sK + key0 * pitch, sV + key0 * pitch, lo_delta, pitch, key0, S,
                            min(MAXP, (S_pad - key0) >> 4), kb == 0, lane, o, m0, m1, l0, l1);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
}

// Normalise and store one 16-row output tile (rows r0 = tile*16 + lane/4 and r0 + 8 of sequence `seq`).
__device__ __forceinline__ void store_tile(const AttnParams& p, int seq, int head, int r0, const float (&o)[8][4],
                                           float l0, float l1, int lane) {
  const int tig = lane & 3;
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = r0 + half * 8;
    if (r >= p.S) continue;
    const float inv = half == 0 ? inv0 : inv1;
    if (p.mode == 1 && r == 0) {
      const int b = seq / p.T, t = seq - b * p.T;
      float* dst = p.cls_out + (static_cast<long long>(b) * p.T + t) * DMODEL + head * HD;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt)
        *reinterpret_cast<float2*>(dst + dt * 8 + tig * 2) = make_float2(o[dt][2 * half] * inv, o[dt][2 * half + 1] * inv);
    } else {
      __nv_bfloat16* dst = p.out + row_of(p, seq, r) * p.ldo + head * HD;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const float v0 = o[dt][2 * half] * inv, v1 = o[dt][2 * half + 1] * inv;
        const uint32_t hi = pack2(v0, v1);
        *reinterpret_cast<uint32_t*>(dst + dt * 8 + tig * 2) = hi;
        if (p.split) *reinterpret_cast<uint32_t*>(dst + p.lo_off + dt * 8 + tig * 2) = pack_lo2(v0, v1, hi);
      }
    }
  }
}

// Q A-fragments of the 16-row tile starting at row r_base, straight from global (each row's 64-wide head slice is 128
// contiguous bytes).
template <bool SPLIT>
__device__ __forceinline__ void load_q_frags(const AttnParams& p, int seq, int head, int r_base, int lane,
                                             uint32_t (&qh)[4][4], uint32_t (&ql)[4][4]) {
  const int g = lane >> 2, tig = lane & 3;
  const int r0 = r_base + g, r1 = r0 + 8, S = p.S;
  const __nv_bfloat16* q0 = p.qkv + row_of(p, seq, r0 < S ? r0 : 0) * p.ld + head * HD;
  const __nv_bfloat16* q1 = p.qkv + row_of(p, seq, r1 < S ? r1 : 0) * p.ld + head * HD;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int c = kk * 16 + tig * 2;
    qh[kk][0] = r0 < S ? __ldg(reinterpret_cast<const uint32_t*>(q0 + c)) : 0u;
    qh[kk][1] = r1 < S ? __ldg(reinterpret_cast<const uint32_t*>(q1 + c)) : 0u;
    qh[kk][2] = r0 < S ? __ldg(reinterpret_cast<const uint32_t*>(q0 + c + 8)) : 0u;
    qh[kk][3] = r1 < S ? __ldg(reinterpret_cast<const uint32_t*>(q1 + c + 8)) : 0u;
    if (SPLIT) {
      ql[kk][0] = r0 < S ? __ldg(reinterpret_cast<const uint32_t*>(q0 + 3 * DMODEL + c)) : 0u;
      ql[kk][1] = r1 < S ? __ldg(reinterpret_cast<const uint32_t*>(q1 + 3 * DMODEL + c)) : 0u;
      ql[kk][2] = r0 < S ? __ldg(reinterpret_cast<const uint32_t*>(q0 + 3 * DMODEL + c + 8)) : 0u;
      ql[kk][3] = r1 < S ? __ldg(reinterpret_cast<const uint32_t*>(q1 + 3 * DMODEL + c + 8)) : 0u;
    }
  }
}

// Generic kernel.  WPP = warps per problem (4: spatial / joint / space-only, 1: long temporal sequences).  CTA = 4 warps.
// K/V of one (sequence, head) are gathered into shared memory with cp.async (all chunks in flight at once).
template <int WPP, bool SPLIT>
__global__ void __launch_bounds__(128, SPLIT ? 2 : (WPP == 4 ? 4 : 3)) attention_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  constexpr int PPC = 4 / WPP;  // problems per CTA
  const int S = p.S;
  const int S_pad = (S + 15) & ~15;
  const long long prob = static_cast<long long>(blockIdx.x) * PPC + warp / WPP;
  const long long n_prob = static_cast<long long>(p.n_seq) * p.heads;
  const bool active = prob < n_prob;
  const int seq = active ? static_cast<int>(prob / p.heads) : 0;
  const int head = active ? static_cast<int>(prob % p.heads) : 0;

  constexpr int NBUF = SPLIT ? 4 : 2;  // K, V (and their lo parts)
  uint8_t* sK = smem + static_cast<size_t>(warp / WPP) * (NBUF * S_pad * 128);
  uint8_t* sV = sK + S_pad * 128;
  const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  const uint32_t lo_delta = 2 * S_pad * 128;  // K_lo = K + lo_delta, V_lo = V + lo_delta

  // ---- stage K and V (16-byte chunks, chunk index XOR (key & 7)) ----
  {
    const int nthr = WPP * 32;
    const int tid = (warp % WPP) * 32 + lane;
    const int total = S_pad * 8;
    for (int idx = tid; idx < total; idx += nthr) {
      const int key = idx >> 3, ch = idx & 7;
      const int off = key * 128 + ((ch ^ (key & 7)) << 4);
      if (active && key < S) {
        const __nv_bfloat16* src = p.qkv + row_of(p, seq, key) * p.ld + head * HD + ch * 8;
        cp_async16(sK_u + off, src + DMODEL);
        cp_async16(sV_u + off, src + 2 * DMODEL);
        if (SPLIT) {
          cp_async16(sK_u + lo_delta + off, src + 3 * DMODEL + DMODEL);
          cp_async16(sV_u + lo_delta + off, src + 3 * DMODEL + 2 * DMODEL);
        }
      } else {
        *reinterpret_cast<uint4*>(sK + off) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sV + off) = make_uint4(0, 0, 0, 0);
        if (SPLIT) {
          *reinterpret_cast<uint4*>(sK + lo_delta + off) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sV + lo_delta + off) = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  const int n_qtiles = S_pad >> 4;
  // the first tile's Q loads are issued before waiting for K/V, later tiles are prefetched one tile ahead
  uint32_t qn[4][4], ql[4][4];
  if (active && (warp % WPP) < n_qtiles) load_q_frags<SPLIT>(p, seq, head, (warp % WPP) * 16, lane, qn, ql);
  cp_async_wait_all();
  if (WPP == 1) __syncwarp(); else __syncthreads();
  if (!active) return;

  for (int qt = warp % WPP; qt < n_qtiles; qt += WPP) {
    uint32_t qa[4][4], qb[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      qa[kk][0] = qn[kk][0]; qa[kk][1] = qn[kk][1]; qa[kk][2] = qn[kk][2]; qa[kk][3] = qn[kk][3];
      if (SPLIT) { qb[kk][0] = ql[kk][0]; qb[kk][1] = ql[kk][1]; qb[kk][2] = ql[kk][2]; qb[kk][3] = ql[kk][3]; }
    }
    if (qt + WPP < n_qtiles) load_q_frags<SPLIT>(p, seq, head, (qt + WPP) * 16, lane, qn, ql);
    float o[8][4], l0, l1;
    attn_tile<SPLIT>(qa, qb, sK_u, sV_u, lo_delta, 128, S, S_pad, lane, o, l0, l1);
    store_tile(p, seq, head, qt * 16 + g, o, l0, l1, lane);
  }
}

// Long sequences (joint space-time attention, 1 + N*T tokens - 6 273 at T = 32): one CTA per (sequence, head, block of
// 64 queries); K / V are streamed through a double-buffered cp.async ring in 64-key chunks (re-read from L2 by every
// query block), the online-softmax state stays in registers.
template <bool SPLIT>
__global__ void __launch_bounds__(128, SPLIT ? 2 : 3) attention_stream_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2;
  const int S = p.S;
  const int seq = blockIdx.y / p.heads, head = blockIdx.y % p.heads;
  const int q_base = blockIdx.x * 64 + warp * 16;
  constexpr int NBUF = SPLIT ? 4 : 2;
  constexpr int STAGE = NBUF * 64 * 128;          // K | V (| K_lo | V_lo), 64 rows of 128 B each
  constexpr uint32_t lo_delta = 2 * 64 * 128;
  const uint32_t sm = smem_u32(smem);
  const int n_kb = (S + 63) >> 6;

  auto prefetch = [&](int kb) {
    const uint32_t base = sm + (kb & 1) * STAGE;
    uint8_t* basep = smem + (kb & 1) * STAGE;
    for (int idx = threadIdx.x; idx < 64 * 8; idx += 128) {
      const int row = idx >> 3, ch = idx & 7;
      const int key = kb * 64 + row;
      const int off = row * 128 + ((ch ^ (row & 7)) << 4);
      if (key < S) {
        const __nv_bfloat16* src = p.qkv + row_of(p, seq, key) * p.ld + head * HD + ch * 8;
        cp_async16(base + off, src + DMODEL);
        cp_async16(base + 64 * 128 + off, src + 2 * DMODEL);
        if (SPLIT) {
          cp_async16(base + lo_delta + off, src + 3 * DMODEL + DMODEL);
          cp_async16(base + lo_delta + 64 * 128 + off, src + 3 * DMODEL + 2 * DMODEL);
        }
      } else {
        *reinterpret_cast<uint4*>(basep + off) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(basep + 64 * 128 + off) = make_uint4(0, 0, 0, 0);
        if (SPLIT) {
          *reinterpret_cast<uint4*>(basep + lo_delta + off) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(basep + lo_delta + 64 * 128 + off) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  prefetch(0);
  uint32_t qa[4][4], qb[4][4];
  load_q_frags<SPLIT>(p, seq, head, q_base, lane, qa, qb);
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int S_pad = (S + 15) & ~15;
  for (int kb = 0; kb < n_kb; ++kb) {
    if (kb + 1 < n_kb) {
      prefetch(kb + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (q_base < S) {
      const uint32_t base = sm + (kb & 1) * STAGE;
      attn_block<SPLIT>(qa, qb, base, base + 64 * 128, lo_delta, 128, kb * 64, S, min(4, (S_pad - kb * 64) >> 4), kb == 0,
                        lane, o, m0, m1, l0, l1);
    }
    __syncthreads();  // everyone is done with this stage before it is refilled two iterations later
  }
  if (q_base >= S) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  store_tile(p, seq, head, q_base + g, o, l0, l1, lane);
}

// Temporal attention for T <= 48: one CTA per (b, n) sequence.  Its T rows of the packed qkv matrix are CONSECUTIVE
// rows, so the whole [T, 2304] block (all 12 heads, q|k|v) is pulled into shared memory with cp.async in one go
// (T * 4.5 KiB in flight per CTA, 3 CTAs / SM - enough to run at HBM speed), each of the 4 warps owns 3 heads (Q via
// ldmatrix), the output tile is
// written back over the head's own Q slice and the CTA stores the [T, 768] result as fully coalesced 16-byte chunks.
constexpr int TROW = 3 * DMODEL * 2;  // 4608 bytes per staged row
__global__ void __launch_bounds__(128, 3) temporal_attention_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int T = p.S;
  const int T_pad = (T + 15) & ~15;
  const int seq = blockIdx.x;
  const uint32_t sm = smem_u32(smem);
  const long long row0 = static_cast<long long>(seq) * T;

  // ---- stage the sequence: 288 16-byte chunks per row, chunk c -> segment c/8 (q h0..11 | k h0..11 | v h0..11) ----
  for (int idx = threadIdx.x; idx < T_pad * 288; idx += 128) {
    const int row = idx / 288, c = idx - row * 288;
    const uint32_t off = row * TROW + (c >> 3) * 128 + (((c & 7) ^ (row & 7)) << 4);
    if (row < T) cp_async16(sm + off, p.qkv + (row0 + row) * p.ld + c * 8);
    else *reinterpret_cast<uint4*>(smem + off) = make_uint4(0, 0, 0, 0);
  }
  cp_async_wait_all();
  __syncthreads();

  const int mat = lane >> 3, r = lane & 7;
  for (int head = warp; head < 12; head += 4) {  // 4 warps x 3 heads
  const uint32_t sQ = sm + head * 128, sK = sQ + DMODEL * 2, sV = sQ + 2 * DMODEL * 2;
  for (int qt = 0; qt < (T_pad >> 4); ++qt) {
    uint32_t qa[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int row = qt * 16 + (mat & 1) * 8 + r;
      const int ch = kk * 2 + (mat >> 1);
      ldsm_x4(sQ + row * TROW + ((ch ^ (row & 7)) << 4), qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3]);
    }
    float o[8][4], l0, l1;
    if (T_pad == 16) attn_tile<false, 1>(qa, qa, sK, sV, 0, TROW, T, T_pad, lane, o, l0, l1);
    else if (T_pad == 32) attn_tile<false, 2>(qa, qa, sK, sV, 0, TROW, T, T_pad, lane, o, l0, l1);
    else attn_tile<false, 4>(qa, qa, sK, sV, 0, TROW, T, T_pad, lane, o, l0, l1);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    __syncwarp();  // every lane has its Q fragments of this tile: the Q rows may now be overwritten
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = qt * 16 + g + half * 8;
      const float inv = half == 0 ? inv0 : inv1;
      uint8_t* qrow = smem + row * TROW + head * 128;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const float v0 = o[dt][2 * half] * inv, v1 = o[dt][2 * half + 1] * inv;
        const uint32_t hi = pack2(v0, v1);
        const int off = ((dt ^ (row & 7)) << 4) + tig * 4;
        *reinterpret_cast<uint32_t*>(qrow + off) = hi;
      }
    }
  }
  }
  __syncthreads();
  // ---- coalesced store of the [T, 768] (and optional lo) block ----
  for (int idx = threadIdx.x; idx < T * 96; idx += 128) {
    const int row = idx / 96, c = idx - row * 96;
    const uint4 v = *reinterpret_cast<const uint4*>(smem + row * TROW + (c >> 3) * 128 + (((c & 7) ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(p.out + (row0 + row) * p.ldo + c * 8) = v;
  }
}

// CLS-query-only attention (one warp per (sequence, head)), fp32 CUDA-core math.  Used for the LAST block when only the
// CLS feature is returned (timesformer.py:342-345 keeps x[:, 0]): the patch-token queries of that block feed nothing, so
// only the CLS row attends - over the same keys / values as the full kernels.  Each lane walks keys lane, lane + 32, ...
// with an online softmax (own max / sum / 64 accumulators), lanes are merged at the end with a transposing shuffle
// reduction that leaves outputs 2*lane, 2*lane + 1 in each lane.  hi/lo inputs (split precision) are summed on load.
__global__ void __launch_bounds__(128) cls_attention_kernel(const AttnParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long prob = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (prob >= static_cast<long long>(p.n_seq) * p.heads) return;
  const int seq = static_cast<int>(prob / p.heads), head = static_cast<int>(prob % p.heads);
  const float scale_log2 = 0.125f * 1.4426950408889634f;

  auto load64 = [&](const __nv_bfloat16* src, float (&dst)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        dst[c * 8 + 2 * j] = __low2float(h);
        dst[c * 8 + 2 * j + 1] = __high2float(h);
      }
    }
  };
  auto load_row = [&](long long row, int which, float (&dst)[64]) {
    const __nv_bfloat16* src = p.qkv + row * p.ld + which * DMODEL + head * HD;
    load64(src, dst);
    if (p.split) {
      float lo[64];
      load64(src + 3 * DMODEL, lo);
#pragma unroll
      for (int d = 0; d < 64; ++d) dst[d] += lo[d];
    }
  };

  float q[64];
  load_row(row_of(p, seq, 0), 0, q);
#pragma unroll
  for (int d = 0; d < 64; ++d) q[d] *= scale_log2;
  float acc[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int key = lane; key < p.S; key += 32) {
    const long long row = row_of(p, seq, key);
    float kv[64];
    load_row(row, 1, kv);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int d = 0; d < 64; d += 4) {
      s0 = fmaf(q[d], kv[d], s0); s1 = fmaf(q[d + 1], kv[d + 1], s1);
      s2 = fmaf(q[d + 2], kv[d + 2], s2); s3 = fmaf(q[d + 3], kv[d + 3], s3);
    }
    const float sc = (s0 + s1) + (s2 + s3);
    const float mn = fmaxf(m, sc);
    const float corr = exp2f(m - mn), pe = exp2f(sc - mn);  // m = -inf on the first key: corr = 0
    m = mn;
    l = l * corr + pe;
    load_row(row, 2, kv);
#pragma unroll
    for (int d = 0; d < 64; ++d) acc[d] = fmaf(pe, kv[d], acc[d] * corr);
  }
  // merge the lanes
  float mw = m;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, off));
  const float sc = m == -INFINITY ? 0.f : exp2f(m - mw);  // lanes without keys contribute nothing
  l *= sc;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] *= sc;
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = upper ? acc[i] : acc[i + n];
      const float keep = upper ? acc[i + n] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  const float inv = 1.0f / l;
  const float v0 = acc[0] * inv, v1 = acc[1] * inv;  // outputs 2*lane, 2*lane + 1
  if (p.mode == 1) {
    float* dst = p.cls_out + static_cast<long long>(seq) * DMODEL + head * HD;
    *reinterpret_cast<float2*>(dst + 2 * lane) = make_float2(v0, v1);
  } else {
    __nv_bfloat16* dst = p.out + row_of(p, seq, 0) * p.ldo + head * HD;
    const uint32_t hi = pack2(v0, v1);
    *reinterpret_cast<uint32_t*>(dst + 2 * lane) = hi;
    if (p.split) *reinterpret_cast<uint32_t*>(dst + p.lo_off + 2 * lane) = pack_lo2(v0, v1, hi);
  }
}

// Attention MAP of one (sequence, head): softmax(q k^T / 8) written out as fp32 [S, S] in the reference's token order
// (CLS first) - forward_features(get_attn=True) returns the last block's spatial map for visualisation
// (models/timesformer.py:155-157,325-331).  Not a hot path: fp32 CUDA-core math, K staged in shared memory as fp32.
__global__ void __launch_bounds__(256) attention_probs_kernel(const AttnParams p, float* __restrict__ probs) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* sK = reinterpret_cast<float*>(smem);            // [S][65]
  float* sQ = sK + static_cast<size_t>(p.S) * 65;         // [8 warps][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.x / p.heads, head = blockIdx.x % p.heads;
  const int S = p.S;
  auto load2 = [&](long long row, int which, int d) {  // elements d, d + 1 of q (0) / k (1) of token `row`
    const __nv_bfloat16* src = p.qkv + row * p.ld + which * DMODEL + head * HD + d;
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(src);
    float2 f = __bfloat1622float2(h);
    if (p.split) {
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + 3 * DMODEL));
      f.x += g.x; f.y += g.y;
    }
    return f;
  };
  for (int idx = threadIdx.x; idx < S * 32; idx += 256) {
    const int key = idx >> 5, d = (idx & 31) * 2;
    const float2 f = load2(row_of(p, seq, key), 1, d);
    sK[key * 65 + d] = f.x;
    sK[key * 65 + d + 1] = f.y;
  }
  __syncthreads();
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  float* out = probs + (static_cast<size_t>(seq) * p.heads + head) * S * S;
  for (int i = warp; i < S; i += 8) {
    const float2 qf = load2(row_of(p, seq, i), 0, lane * 2);
    sQ[warp * 64 + lane * 2] = qf.x * scale_log2;
    sQ[warp * 64 + lane * 2 + 1] = qf.y * scale_log2;
    __syncwarp();
    float sc[8];  // S <= 256
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = c * 32 + lane;
      float a = -INFINITY;
      if (j < S) {
        a = 0.f;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) a = fmaf(sQ[warp * 64 + d], sK[j * 65 + d], a);
      }
      sc[c] = a;
      mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      sc[c] = c * 32 + lane < S ? exp2f(sc[c] - mx) : 0.f;
      sum += sc[c];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c * 32 + lane < S) out[static_cast<size_t>(i) * S + c * 32 + lane] = sc[c] * inv;
    __syncwarp();
  }
}

}  // namespace

int launch_attention_probs(const void* qkv, long long ld, float* probs, int B, int T, int N, long long cls_base,
                           int split, cudaStream_t st) {
  AttnParams p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.ld = ld;
  p.out = nullptr; p.ldo = 0; p.cls_out = nullptr;
  p.mode = 1;
  p.T = T; p.N = N; p.cls_base = cls_base; p.heads = 12;
  p.split = split; p.lo_off = DMODEL;
  p.S = N + 1; p.n_seq = B * T;
  if (p.n_seq <= 0) return 0;
  if (p.S > 256) return set_error("attention map: %d tokens per frame exceed the supported 256", p.S);
  const size_t smem = (static_cast<size_t>(p.S) * 65 + 8 * 64) * sizeof(float);
  static PerDevice<size_t> cur_cache;
  size_t& cur = cur_cache.get();
  if (smem > cur) {
    cudaError_t ee = cudaFuncSetAttribute(attention_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem));
    if (ee != cudaSuccess) return set_error("attention map attr: %s", cudaGetErrorString(ee));
    cur = smem;
  }
  ProfScope prof(PROF_ATTENTION, 2.0 * p.n_seq * p.heads * static_cast<double>(p.S) * p.S * HD, st);
  attention_probs_kernel<<<static_cast<unsigned>(p.n_seq * p.heads), 256, smem, st>>>(p, probs);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("attention map launch: %s", cudaGetErrorString(e));
}

int launch_cls_attention(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                         int N, long long cls_base, int split, cudaStream_t st) {
  AttnParams p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.ld = ld;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.cls_out = cls_out;
  p.mode = mode;
  p.T = T; p.N = N; p.cls_base = cls_base; p.heads = 12;
  p.split = split; p.lo_off = DMODEL;
  if (mode == 1 || mode == 3) { p.S = N + 1; p.n_seq = B * T; }
  else if (mode == 2) { p.S = N * T + 1; p.n_seq = B; }
  else return set_error("cls attention: bad mode %d", mode);
  if (p.n_seq <= 0) return 0;
  const long long n_prob = static_cast<long long>(p.n_seq) * p.heads;
  ProfScope prof(PROF_ATTENTION, 4.0 * n_prob * static_cast<double>(p.S) * HD, st);
  cls_attention_kernel<<<static_cast<unsigned>((n_prob + 3) / 4), 128, 0, st>>>(p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("cls attention launch: %s", cudaGetErrorString(e));
}

int launch_attention(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                     int N, long long cls_base, int split, cudaStream_t st, void* scratch, long long scratch_bytes) {
  AttnParams p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.ld = ld;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.cls_out = cls_out;
  p.mode = mode;
  p.T = T; p.N = N; p.cls_base = cls_base; p.heads = 12;
  p.split = split; p.lo_off = DMODEL;
  if (mode == 0) { p.S = T; p.n_seq = B * N; }
  else if (mode == 1) { p.S = N + 1; p.n_seq = B * T; }
  else if (mode == 2) { p.S = N * T + 1; p.n_seq = B; }
  else if (mode == 3) { p.S = N + 1; p.n_seq = B * T; }
  else return set_error("attention: bad mode %d", mode);
  if (p.n_seq <= 0) return 0;
  const int S_pad = (p.S + 15) & ~15;
  const long long n_prob = static_cast<long long>(p.n_seq) * p.heads;
  cudaError_t e;
  ProfScope prof(PROF_ATTENTION, 4.0 * p.n_seq * p.heads * static_cast<double>(p.S) * p.S * HD, st);
  auto launch_generic = [&](auto kern, int ppc, size_t smem, size_t* cur) -> int {
    if (smem > 220 * 1024)
      return set_error("attention: sequence of %d tokens does not fit the shared-memory K/V staging", p.S);
    if (smem > *cur) {
      cudaError_t ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (ee != cudaSuccess) return set_error("attention attr: %s", cudaGetErrorString(ee));
      *cur = smem;
    }
    kern<<<static_cast<unsigned>((n_prob + ppc - 1) / ppc), 128, smem, st>>>(p);
    return 0;
  };
  struct Cur { size_t t, k1, k1s, k4, k4s; };
  static PerDevice<Cur> cur_cache;
  Cur& cc = cur_cache.get();
  size_t &cur_t = cc.t, &cur1 = cc.k1, &cur1s = cc.k1s, &cur4 = cc.k4, &cur4s = cc.k4s;
  const size_t per = static_cast<size_t>(S_pad) * 128;
  int rc = 0;
  if (!split && (mode == 1 || mode == 3)) {
    rc = launch_spatial_attention_sm100(qkv, ld, out, ldo, cls_out, mode, B, T, N, cls_base, st, scratch, scratch_bytes);
    if (rc == 0) { count_launch(); return 0; }
    if (rc != 1) return rc;
    rc = 0;
  }
  if (split && (mode == 1 || mode == 3) && attention_impl() == 1) {
    // precision = "high": hi / lo operands on tcgen05 (attention_hp_sm100.cu)
    rc = launch_spatial_attention_hp_sm100(qkv, ld, out, ldo, cls_out, mode, B, T, N, cls_base, st, scratch, scratch_bytes);
    if (rc == 0) { count_launch(); return 0; }
    if (rc != 1) return rc;
    rc = 0;
  }
  if (!split && mode == 2) {
    // joint space-time attention: flash-style tcgen05 kernel (attention_sm100.cu)
    rc = launch_joint_attention_sm100(qkv, ld, out, ldo, B, N * T, cls_base, st);
    if (rc == 0) { count_launch(); return 0; }
    if (rc != 1) return rc;
    rc = 0;
  }
  if (mode == 0 && S_pad <= 48 && !split) {
    // contiguous-sequence kernel
    const size_t smem = static_cast<size_t>(S_pad) * TROW;
    if (smem > cur_t) {
      e = cudaFuncSetAttribute(temporal_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
      if (e != cudaSuccess) return set_error("attention attr: %s", cudaGetErrorString(e));
      cur_t = smem;
    }
    temporal_attention_kernel<<<static_cast<unsigned>(p.n_seq), 128, smem, st>>>(p);
  } else if (mode == 0) {
    rc = split ? launch_generic(attention_kernel<1, true>, 4, 4 * 4 * per, &cur1s)
               : launch_generic(attention_kernel<1, false>, 4, 4 * 2 * per, &cur1);
  } else if ((split ? 4 : 2) * per > 200 * 1024) {
    // long sequence: K/V streamed in 64-key chunks, one CTA per 64-query block
    dim3 grid(static_cast<unsigned>((p.S + 63) / 64), static_cast<unsigned>(n_prob));
    if (n_prob > 65535) return set_error("attention: too many (sequence, head) problems for the streaming kernel");
    if (split) {
      static PerDevice<bool> set_cache;
      bool& set = set_cache.get();
      if (!set) { cudaFuncSetAttribute(attention_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4 * 64 * 128); set = true; }
      attention_stream_kernel<true><<<grid, 128, 2 * 4 * 64 * 128, st>>>(p);
    } else {
      attention_stream_kernel<false><<<grid, 128, 2 * 2 * 64 * 128, st>>>(p);
    }
  } else {
    rc = split ? launch_generic(attention_kernel<4, true>, 1, 4 * per, &cur4s)
               : launch_generic(attention_kernel<4, false>, 1, 2 * per, &cur4);
  }
  if (rc != 0) return rc;
  count_launch();
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("attention launch: %s", cudaGetErrorString(e));
}

}  // namespace dvst
