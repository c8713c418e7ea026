// Spatial attention on tcgen05 / TMEM, second generation ("split" kernel): softmax(q k^T / 8) v over the 1 + N tokens of one
// frame (models/timesformer.py:140-164 divided space-time, :120-123 space_only), restructured around what the round-2
// microbenchmark (tools/tmem_bench.cu) measured on a B200 SM: the exponentials are the bound (MUFU ex2: ~14 / clk / SM with
// two warps per scheduler, ~8.5 with one), a tcgen05.ld round trip costs ~150 clk of latency but TMEM bandwidth is ample,
// and the first-generation kernel ran its two 128-row tiles in lock step, so every MMA round trip (S, P V), the O read-out
// and the store staging were exposed between two half-rate exp phases (7 600 clk per (frame, head) problem for ~4 000 clk of
// exponentials).  Here:
//
//   * EIGHT softmax warps work on ONE row tile at a time: the two threads (warp w, warp w + 4: same TMEM lane quarter) that
//     share a query row each take one half of the keys - keys [0, K0) and [K0, NPAD) - so every scheduler always has two
//     warps feeding the MUFU and a tile's softmax takes half as long.  Both use the same reference point (max of the first
//     32 keys of the second half and the CLS key: the one-pass softmax of the first-generation kernel, rows that overflow go
//     to the fp32 fix-up kernel), so one P V MMA chain serves the row and the two partial row sums just add.
//   * P of the first key half is written in place over the consumed S columns, P of the second half goes to its own TMEM
//     columns [max(NPAD, 192), ..) - a thread never overwrites scores its partner still has to read; O accumulates in two
//     independent chains (even key steps -> columns [128, 192), odd -> [64, 128)): a tcgen05.mma that accumulates onto the
//     previous one's tile waits ~135 clk for it, so one chain of 13 small P V steps took 1 800 clk, two take half.
//   * FOUR dedicated epilogue warps read O out of TMEM, scale by 1 / (sum0 + sum1), stage and TMA-store it, so the softmax
//     warps go straight from one tile to the next: the units (problem, row tile) alternate between the two 256-column TMEM
//     regions, and while the softmax warps work in one region the tensor core runs P V, and then S of the unit after next,
//     in the other.  The MMA round trips, the O read-out and the stores are hidden behind the exponentials.
//   * Q | K and V are separate TMA transactions with their own barriers: Q | K of problem i + 2 are requested as soon as the
//     S MMAs of problem i have read them (more than a full problem ahead), V after its P V MMAs.
//
// Warp roles (512 threads): 0 TMA producer + CLS-row copy, 1 MMA issuer, 2 TMEM allocator, 4-11 softmax, 12-15 epilogue.
// Covers 32 <= NPAD <= 208 key columns (1 + N rounded up to 16); other shapes stay on attention_sm100.cu.
#include <type_traits>

#include "attn_sm100_common.cuh"

namespace dvst {

using namespace attn;

namespace {

constexpr int SP2_THREADS = 512;
constexpr int O2_COL = 64;   // second O accumulator (odd key steps): TMEM columns [64, 128) of the region, free once S is consumed

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[32]) {  // fills v[0..15]
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[16]) {  // stores v[0..7]
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

#ifdef DVST_SP2_TIMING
// debug build only (tools/attn_prof.py with DVST_LIB=libdvst_dbg.so): per-role cycle counters of block 0
__device__ long long dvst_sp2_dbg[32];
__device__ long long dvst_sp2_trace[4 * 8 * 4];   // [role: mma, softmax w4, softmax w8, epilogue w12][unit 40..47][event]
#define SP2_TR(role, ev) do { if (blockIdx.x == 0 && lane == 0 && u >= 40 && u < 48) dvst_sp2_trace[((role) * 8 + (u - 40)) * 4 + (ev)] = clock64(); } while (0)
#define SP2_T0() long long _t0 = clock64()
#define SP2_ACC(slot) do { const long long _t1 = clock64(); if (blockIdx.x == 0) dbg_acc[slot] += _t1 - _t0; _t0 = _t1; } while (0)
#else
#define SP2_T0() do {} while (0)
#define SP2_ACC(slot) do {} while (0)
#define SP2_TR(role, ev) do {} while (0)
#endif

// exp2 on the FMA pipe for a fixed subset of the score pairs (the FlashAttention-4 arrangement): the softmax is bound by
// the MUFU (16 ex2 / clk / SM, ~14 reachable - tools/tmem_bench.cu), while the FMA pipe idles.  Cody-Waite: n = round(x)
// through the 1.5 * 2^23 magic constant, f = x - n in [-0.5, 0.5], 2^f by a cubic (max relative error 1.4e-4 in fp32,
// 30x below the bf16 rounding of P), result = 2^f with n added into the exponent field.  x is clamped to [-126, 127]: below
// the result would leave the normal range, above it must stay huge so that the row still reaches the fix-up kernel.
// Measured (cfg2 launch, same box): 6 of 16 pairs on the FMA pipe (mask 0x5252) 0.2415 -> 0.2518 ms - the two clamps and the
// exponent insertion are ALU-pipe instructions, and the ALU pipe (bf16 packing, masks) is already ~48 % busy, so the
// exponentials saved on the MUFU are paid back there.  Left switched off (mask 0); packed ex2.approx.f16x2 / bf16x2 are no way
// out either: they compile to TWO MUFU.EX2 (one per half) and run at the same 13.5 exp / clk / SM.
constexpr uint32_t SP2_POLY_PAIRS = 0x0000u;   // bit j: pair j (of 16 per 32-column chunk) is computed on the FMA pipe
__device__ __forceinline__ void exp2_poly2(float& a0, float& a1) {
  a0 = fminf(fmaxf(a0, -126.0f), 127.0f);
  a1 = fminf(fmaxf(a1, -126.0f), 127.0f);
  const uint64_t x2 = pk2(a0, a1);
  const uint64_t magic = 0x4B4000004B400000ull, nmagic = 0xCB400000CB400000ull, mone = 0xBF800000BF800000ull;
  const uint64_t t2 = add2(x2, magic);
  const uint64_t f2 = fma2(add2(t2, nmagic), mone, x2);
  uint64_t p2 = fma2(0x3D61FB263D61FB26ull, f2, 0x3E786ECB3E786ECBull);
  p2 = fma2(p2, f2, 0x3F31798D3F31798Dull);
  p2 = fma2(p2, f2, 0x3F7FFB4A3F7FFB4Aull);
  float p0, p1, t0, t1;
  unpk2(p2, p0, p1);
  unpk2(t2, t0, t1);
  a0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  a1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

struct Sp2Geom {
  int K0;      // keys of the first half (multiple of 16); the second half holds NPAD - K0
  int p1_col;  // TMEM column (inside the region) of the second half's P
};

// kFixed: the ViT-B/16 224^2 frame (N = 196 patches: S = 197, NPAD = 208, K0 = 112, P1 at column 208) with every geometry
// value a compile-time constant - the shape of the whole benchmark; the run-time version covers the other crops.
template <bool kFixed>
__global__ void __launch_bounds__(SP2_THREADS, 1)
spatial_attention_split_sm100_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOutA,
                                     const __grid_constant__ CUtensorMap tmOutB, const SpParams p, const Sp2Geom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int S = kFixed ? 197 : p.S, N = kFixed ? 196 : p.N, NPAD = kFixed ? 208 : p.NPAD;
  const int K0 = kFixed ? 112 : g.K0;
  const int P1 = kFixed ? 208 : g.p1_col;
  const int n_tiles = S > 128 ? 2 : 1;
  const uint32_t q_bytes = n_tiles * TILE_BYTES;
  const uint32_t kv_bytes = (static_cast<uint32_t>(NPAD) * 128u + 1023u) & ~1023u;
  const uint32_t stage_bytes = q_bytes + 2 * kv_bytes;
  const uint32_t sO = base + 2 * stage_bytes;        // output staging: one 128-row tile per TMEM region
  const uint32_t sums_off = 2 * stage_bytes + 2 * TILE_BYTES;   // fp32 [2 regions][2 key halves][128 rows]
  const uint32_t bars = base + sums_off + 2048;
  const uint32_t qk_full = bars, qk_empty = bars + 16, v_full = bars + 32, v_empty = bars + 48;   // [2 stages] each
  const uint32_t s_full = bars + 64, p_full = bars + 80, o_full = bars + 96, o_empty = bars + 112;  // [2 regions] each
  const uint32_t tmem_slot = bars + 128;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));
  float* sums = reinterpret_cast<float*>(smem + sums_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // lane quarters (32 rows each) of a row tile that hold at least one real query row
  const int act_a = min(4, (min(S, 128) + 31) >> 5);
  const int act_b = n_tiles == 2 ? (S - 128 + 31) >> 5 : 0;
  const int act_r0 = act_a, act_r1 = n_tiles == 2 ? act_b : act_a;   // by TMEM region (units alternate regions)
  const int n_my = static_cast<int>(blockIdx.x) < p.n_prob
                       ? (p.n_prob - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  const int n_units = n_my * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmOutA);
    tma_prefetch_desc(&tmOutB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(qk_full + 8 * s, 1);
      mbar_init(qk_empty + 8 * s, 1);
      mbar_init(v_full + 8 * s, 1);
      mbar_init(v_empty + 8 * s, 1);
      mbar_init(s_full + 8 * s, 1);
      mbar_init(o_full + 8 * s, 1);
    }
    mbar_init(p_full, 2 * act_r0);
    mbar_init(p_full + 8, 2 * act_r1);
    mbar_init(o_empty, act_r0);
    mbar_init(o_empty + 8, act_r1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  // K / V rows [S, NPAD) are never written by the loads: zero them once (P is 0 there, and 0 * garbage must not be NaN)
  for (int s = 0; s < 2; ++s) {
    const int pad_chunks = (NPAD - S) * 8;
    for (int i = threadIdx.x; i < 2 * pad_chunks; i += SP2_THREADS) {
      const int which = i / pad_chunks, j = i - which * pad_chunks;
      uint8_t* dst = smem + s * stage_bytes + q_bytes + which * kv_bytes + (S * 128 + j * 16);
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // the prologue above touched shared memory and TMEM only
  pdl_launch_dependents();
#ifdef DVST_SP2_TIMING
  long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif

  if (warp == 0) {
    // ===== producer: Q | K (one barrier) and V (its own barrier) of every problem, plus the CLS row of each =====
    for (int it = 0; it < n_my; ++it) {
      const int prob = blockIdx.x + it * gridDim.x;
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int seq = prob / p.heads, head = prob - seq * p.heads;
      const int b = seq / p.T, t = seq - b * p.T;
      const long long cls_row = p.cls_base + (p.mode == 1 ? b : seq);
      const uint32_t st = base + s * stage_bytes;
      if (lane == 0) mbar_wait(qk_empty + 8 * s, ph ^ 1);
      __syncwarp();
      if (lane == 0) {
        tma_load_3d(st, &tmIn, qk_full + 8 * s, head * HD, t, b * N);
        tma_load_3d(st + q_bytes, &tmIn, qk_full + 8 * s, DMODEL + head * HD, t, b * N);
      }
      if (lane < 16) {
        const int which = lane >> 3, ch = lane & 7;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + cls_row * p.ld + which * DMODEL + head * HD) + ch);
        const uint32_t off = (which == 0 ? 0u : q_bytes) + N * 128 + ((ch ^ (N & 7)) << 4);
        *reinterpret_cast<uint4*>(smem + s * stage_bytes + off) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(qk_full + 8 * s, 2u * N * 128u);
        mbar_wait(v_empty + 8 * s, ph ^ 1);
      }
      __syncwarp();
      if (lane == 0) tma_load_3d(st + q_bytes + kv_bytes, &tmIn, v_full + 8 * s, 2 * DMODEL + head * HD, t, b * N);
      if (lane < 8) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + cls_row * p.ld + 2 * DMODEL + head * HD) + lane);
        const uint32_t off = q_bytes + kv_bytes + N * 128 + ((lane ^ (N & 7)) << 4);
        *reinterpret_cast<uint4*>(smem + s * stage_bytes + off) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(v_full + 8 * s, 1u * N * 128u);
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer (the whole warp runs the control flow, one elected lane issues).  Unit u = (problem u / n_tiles, row tile u % n_tiles) lives in TMEM region u & 1. =====
      const uint32_t idesc_s = make_idesc(128, NPAD, 0);
      constexpr uint32_t idesc_pv = make_idesc(128, HD, 1);
      const int ksteps = NPAD >> 4;
      auto issue_s = [&](int u) {
        const int it = n_tiles == 2 ? (u >> 1) : u, tile = n_tiles == 2 ? (u & 1) : 0, s = it & 1;
        const uint32_t st = base + s * stage_bytes;
        if (tile == 0) {
#ifdef DVST_SP2_TIMING
          const long long w0 = clock64();
#endif
          mbar_wait(qk_full + 8 * s, (it >> 1) & 1);
          tc_fence_after();
#ifdef DVST_SP2_TIMING
          if (blockIdx.x == 0) dbg_acc[4] += clock64() - w0;
#endif
        }
        const uint32_t d = tmem_base + (u & 1) * 256;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16<1>(d, make_kmajor_sw128_desc(st + tile * TILE_BYTES + k * 32),
                         make_kmajor_sw128_desc(st + q_bytes + k * 32), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(s_full + 8 * (u & 1));
          if (tile == n_tiles - 1) umma_commit(qk_empty + 8 * s);   // Q | K of this problem have been read
        }
        __syncwarp();
      };
      auto issue_pv = [&](int u) {
        const int it = n_tiles == 2 ? (u >> 1) : u, tile = n_tiles == 2 ? (u & 1) : 0, s = it & 1;
        const uint32_t st = base + s * stage_bytes;
        if (tile == 0) {
#ifdef DVST_SP2_TIMING
          const long long w0 = clock64();
#endif
          mbar_wait(v_full + 8 * s, (it >> 1) & 1);
          tc_fence_after();
#ifdef DVST_SP2_TIMING
          if (blockIdx.x == 0) dbg_acc[5] += clock64() - w0;
#endif
        }
        const uint32_t region = tmem_base + (u & 1) * 256;
        // Fully unrolled with incrementally formed operands: a single thread issues these, and a rolled loop that rebuilt
        // the 64-bit descriptor per step spent ~135 clk of dependent scalar latency per MMA (1 800 clk per 13-step P V).
        const uint64_t vdesc0 = make_mnmajor_sw128_desc(st + q_bytes + kv_bytes);
        const uint32_t a_lo = region, a_hi = region + static_cast<uint32_t>(P1 - (K0 >> 1));
        if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 13; ++j) {
          if (j < ksteps) {
            // P of key step j: TMEM column key / 2 of its half's P block.  Two independent accumulation chains (even /
            // odd key steps -> O_COL / O2_COL).
            const uint32_t a_addr = ((j << 4) < K0 ? a_lo : a_hi) + j * 8;
            umma_ts_bf16(region + ((j & 1) ? O2_COL : O_COL), a_addr, vdesc0 + static_cast<uint64_t>(j * 128), idesc_pv,
                         j >= 2 ? 1u : 0u);
          }
        }
        umma_commit(o_full + 8 * (u & 1));
        if (tile == n_tiles - 1) umma_commit(v_empty + 8 * s);    // V of this problem has been read
        }
        __syncwarp();
      };
      if (n_units > 0) issue_s(0);
      if (n_units > 1) issue_s(1);
      SP2_T0();
      for (int u = 0; u < n_units; ++u) {
        const int r = u & 1;
        const uint32_t par = (u >> 1) & 1;
        mbar_wait(p_full + 8 * r, par);
        tc_fence_after();
        SP2_ACC(0);
        SP2_TR(0, 0);
        issue_pv(u);
        SP2_ACC(1);
        SP2_TR(0, 1);
        if (u + 2 < n_units) {
          mbar_wait(o_empty + 8 * r, par);   // O of unit u has left the region: S of unit u + 2 may overwrite it
          tc_fence_after();
          SP2_ACC(2);
          SP2_TR(0, 2);
          issue_s(u + 2);
          SP2_ACC(3);
          SP2_TR(0, 3);
        }
      }
#ifdef DVST_SP2_TIMING
      if (blockIdx.x == 0 && lane == 0) for (int i = 0; i < 6; ++i) dvst_sp2_dbg[i] = dbg_acc[i];
#endif
    }
  } else if (warp >= 4 && warp < 12) {
    // ===== softmax: two threads per query row, one key half each =====
    const int q = warp & 3;
    auto softmax_units = [&](auto hc) {
      constexpr int h = decltype(hc)::value;
      const float scale_log2 = 0.125f * 1.4426950408889634f;
      const int c_begin = h ? K0 : 0;
      const int len = h ? NPAD - K0 : K0;
      const int nfull = len >> 5;
      const bool tail = (len & 16) != 0;
      const int ref_n = min(32, S - K0);            // real keys among the reference columns [K0, K0 + 32)
      SP2_T0();
      for (int u = 0; u < n_units; ++u) {
        const int tile = n_tiles == 2 ? (u & 1) : 0;
        const int r = u & 1;
        const uint32_t par = (u >> 1) & 1;
        if (q >= (tile == 0 ? act_a : act_b)) continue;
        const uint32_t treg = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + r * 256;
        const uint32_t pdst = treg + (h ? P1 : 0);
        mbar_wait(s_full + 8 * r, par);
        tc_fence_after();
        SP2_ACC(tile);
        SP2_TR(1 + h, 0);
        // va / vb ping-pong.  The reference columns [K0, K0 + 32) are the second half's first chunk; the first half loads its
        // own first chunk in the same round trip.  Chunk c of this thread's half sits in (c even ? first : second).
        uint32_t va[32], vb[32];
        uint32_t (&first)[32] = h ? va : vb;
        uint32_t (&second)[32] = h ? vb : va;
        float mx;
        {
          uint32_t sc;
          tmem_ld_32x32(treg + K0, va);
          if (h == 0) {
            if (nfull > 0) tmem_ld_32x32(treg, vb); else tmem_ld_32x16(treg, vb);
          } else {
            if (nfull > 1) tmem_ld_32x32(treg + K0 + 32, vb);
            else if (nfull == 1 && tail) tmem_ld_32x16(treg + K0 + 32, vb);
          }
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(sc) : "r"(treg + N) : "memory");
          tmem_ld_wait();
          // ---- reference point: max of the first keys of the second half and of the CLS key (same for both threads of a row)
          float m0 = __uint_as_float(sc), m1 = -INFINITY;
          if (ref_n == 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              m0 = max3(m0, __uint_as_float(va[4 * j]), __uint_as_float(va[4 * j + 1]));
              m1 = max3(m1, __uint_as_float(va[4 * j + 2]), __uint_as_float(va[4 * j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ref_n) m1 = fmaxf(m1, __uint_as_float(va[j]));
          }
          mx = fmaxf(m0, m1);
        }
        SP2_ACC(2);
        SP2_TR(1 + h, 1);
        const float moff = mx * scale_log2;
        const uint64_t scale2 = pk2(scale_log2, scale_log2), nmoff2 = pk2(-moff, -moff);
        uint64_t sum2a = 0ull, sum2b = 0ull;
        // p = exp2(s * scale - moff) for `ncols` (32 or 16) score columns starting at key `col`; bf16 pairs -> TMEM at pcol
        auto exp_chunk = [&](const uint32_t (&cur)[32], int col, int ncols, uint32_t pcol) {
          uint32_t pk[16];
          const bool masked = col + ncols > S;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (2 * j < ncols) {
              float a0, a1;
              unpk2(fma2(pk2(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), scale2, nmoff2), a0, a1);
              float e0, e1;
              if ((SP2_POLY_PAIRS >> j) & 1u) {
                e0 = a0; e1 = a1;
                exp2_poly2(e0, e1);
              } else {
                e0 = ex2_approx(a0); e1 = ex2_approx(a1);
              }
              if (masked) {
                if (col + 2 * j >= S) e0 = 0.f;
                if (col + 2 * j + 1 >= S) e1 = 0.f;
              }
              if (j & 1) sum2b = add2(sum2b, pk2(e0, e1)); else sum2a = add2(sum2a, pk2(e0, e1));
              pk[j] = pack_bf16x2(e0, e1);
            }
          }
          if (ncols == 32) tmem_st_32x16(pcol, pk); else tmem_st_32x8(pcol, pk);
        };
        // ---- exp pass over this thread's key half, the next chunk's tcgen05.ld in flight behind the current chunk's math ----
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nfull) {
            uint32_t (&cur)[32] = (c & 1) ? second : first;
            uint32_t (&nxt)[32] = (c & 1) ? first : second;
            if (c > 0) tmem_ld_wait();
            if (!(h == 1 && c == 0)) {       // (the second half already has chunk 1 in flight / landed)
              if (c + 1 < nfull) tmem_ld_32x32(treg + c_begin + (c + 1) * 32, nxt);
              else if (tail) tmem_ld_32x16(treg + c_begin + (c + 1) * 32, nxt);
            }
            exp_chunk(cur, c_begin + c * 32, 32, pdst + c * 16);
          }
        }
        if (tail) {
          if (nfull > 0) tmem_ld_wait();
          if (nfull & 1) exp_chunk(second, c_begin + nfull * 32, 16, pdst + nfull * 16);
          else exp_chunk(first, c_begin + nfull * 32, 16, pdst + nfull * 16);
        }
        SP2_ACC(3 + tile);
        SP2_TR(1 + h, 2);
        float sa, sb;
        unpk2(add2(sum2a, sum2b), sa, sb);
        sums[(r * 2 + h) * 128 + q * 32 + lane] = sa + sb;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + 8 * r);
        SP2_ACC(5);
        if (q == 0) SP2_TR(1 + h, 3);
      }
#ifdef DVST_SP2_TIMING
      if (blockIdx.x == 0 && lane == 0 && q == 0) for (int i = 0; i < 6; ++i) dvst_sp2_dbg[8 + h * 6 + i] = dbg_acc[i];
#endif
    };
    if (warp < 8) softmax_units(std::integral_constant<int, 0>{}); else softmax_units(std::integral_constant<int, 1>{});
  } else if (warp >= 12) {
    // ===== epilogue: O out of TMEM, 1 / row sum, bf16, staging tile, TMA store =====
    const int q = warp & 3;
    const bool store_leader = q == 0 && lane == 0;
    SP2_T0();
    for (int u = 0; u < n_units; ++u) {
      const int it = n_tiles == 2 ? (u >> 1) : u, tile = n_tiles == 2 ? (u & 1) : 0;
      const int r = u & 1;
      const uint32_t par = (u >> 1) & 1;
      const int n_act = tile == 0 ? act_a : act_b;
      if (q >= n_act) continue;
      const int prob = blockIdx.x + it * gridDim.x;
      const int seq = prob / p.heads, head = prob - seq * p.heads;
      const int b = seq / p.T, t = seq - b * p.T;
      const int rloc = q * 32 + lane;
      const int row = tile * 128 + rloc;
      const int rows_tile = tile == 0 ? p.rows_a : p.rows_b;
      const uint32_t treg = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + r * 256;
      const uint32_t so_tile = sO + r * TILE_BYTES;
      mbar_wait(p_full + 8 * r, par);     // the softmax warps' row sums are visible
      SP2_ACC(0);
      if (q == 0) SP2_TR(3, 0);
      mbar_wait(o_full + 8 * r, par);
      tc_fence_after();
      SP2_ACC(1);
      if (q == 0) SP2_TR(3, 1);
      uint32_t o0[32], o1[32];
      {
        // O = even-step accumulator + odd-step accumulator (NPAD >= 32: both chains have run)
        uint32_t t0[32];
        tmem_ld_32x32(treg + O_COL, o0);
        tmem_ld_32x32(treg + O2_COL, t0);
        tmem_ld_wait();
        tmem_ld_32x32(treg + O_COL + 32, o1);
#pragma unroll
        for (int j = 0; j < 32; ++j) o0[j] = __float_as_uint(__uint_as_float(o0[j]) + __uint_as_float(t0[j]));
        tmem_ld_32x32(treg + O2_COL + 32, t0);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o1[j] = __float_as_uint(__uint_as_float(o1[j]) + __uint_as_float(t0[j]));
      }
      const float sum = sums[(r * 2) * 128 + rloc] + sums[(r * 2 + 1) * 128 + rloc];
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty + 8 * r);
      SP2_ACC(2);
      if (q == 0) SP2_TR(3, 2);
      const float inv = 1.0f / sum;
      // exponentials past 2^64 (or NaN): this row is recomputed exactly by attention_fixup_kernel
      if (row < S && !(sum < 1.8e19f)) p.fix_list[atomicAdd(p.fix_count, 1)] = make_int2(prob, row);
      // the store issued two units ago read this staging tile: it must have finished
      if (store_leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      named_bar_sync(1 + tile, n_act * 32);
      if (row < N) {
        const uint32_t dst = so_tile + rloc * 128;
        const int sw = rloc & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t (&o)[32] = j < 4 ? o0 : o1;
          const int j8 = (j & 3) * 8;
          const uint32_t w0 = pack_bf16x2(__uint_as_float(o[j8]) * inv, __uint_as_float(o[j8 + 1]) * inv);
          const uint32_t w1 = pack_bf16x2(__uint_as_float(o[j8 + 2]) * inv, __uint_as_float(o[j8 + 3]) * inv);
          const uint32_t w2 = pack_bf16x2(__uint_as_float(o[j8 + 4]) * inv, __uint_as_float(o[j8 + 5]) * inv);
          const uint32_t w3 = pack_bf16x2(__uint_as_float(o[j8 + 6]) * inv, __uint_as_float(o[j8 + 7]) * inv);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((j ^ sw) << 4)), "r"(w0), "r"(w1),
                       "r"(w2), "r"(w3) : "memory");
        }
      } else if (row == N) {
        if (p.mode == 1) {
          float* dst = p.cls_out + static_cast<long long>(seq) * DMODEL + head * HD;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t (&o)[32] = j < 8 ? o0 : o1;
            const int j4 = (j & 7) * 4;
            *reinterpret_cast<float4*>(dst + j * 4) =
                make_float4(__uint_as_float(o[j4]) * inv, __uint_as_float(o[j4 + 1]) * inv,
                            __uint_as_float(o[j4 + 2]) * inv, __uint_as_float(o[j4 + 3]) * inv);
          }
        } else {
          __nv_bfloat16* dst = p.out + (p.cls_base + seq) * p.ldo + head * HD;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t (&o)[32] = j < 4 ? o0 : o1;
            const int j8 = (j & 3) * 8;
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[j8]) * inv, __uint_as_float(o[j8 + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[j8 + 2]) * inv, __uint_as_float(o[j8 + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[j8 + 4]) * inv, __uint_as_float(o[j8 + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[j8 + 6]) * inv, __uint_as_float(o[j8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + j * 8) = w;
          }
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + tile, n_act * 32);
      if (store_leader && rows_tile > 0) {
        tma_store_3d(tile == 0 ? &tmOutA : &tmOutB, so_tile, head * HD, t, b * N + tile * 128);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      SP2_ACC(3);
      if (q == 0) SP2_TR(3, 3);
    }
#ifdef DVST_SP2_TIMING
    if (blockIdx.x == 0 && warp == 12 && lane == 0) for (int i = 0; i < 4; ++i) dvst_sp2_dbg[24 + i] = dbg_acc[i];
#endif
    if (store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace

#ifdef DVST_SP2_TIMING
extern "C" int dvst_debug_sp2_read(long long* out32) {
  return cudaMemcpyFromSymbol(out32, dvst_sp2_dbg, sizeof(long long) * 32) == cudaSuccess ? 0 : 1;
}
extern "C" int dvst_debug_sp2_trace(long long* out128) {
  return cudaMemcpyFromSymbol(out128, dvst_sp2_trace, sizeof(long long) * 128) == cudaSuccess ? 0 : 1;
}
#endif

// Returns 1 when the shape is not covered (the caller then uses the first-generation kernel), 0 when launched.
int launch_spatial_attention_split_sm100(const SpParams& p, const CUtensorMap& tmIn, const CUtensorMap& tmOutA,
                                         const CUtensorMap& tmOutB, int grid, cudaStream_t st) {
  if (p.NPAD < 32 || p.NPAD > 208) return 1;
  Sp2Geom g;
  g.K0 = ((p.NPAD / 2) + 15) & ~15;
  g.p1_col = p.NPAD > 192 ? p.NPAD : 192;
  if (g.p1_col + (p.NPAD - g.K0) / 2 > 256 || g.K0 >= p.S) return 1;
  const int n_tiles = p.S > 128 ? 2 : 1;
  const size_t kv_bytes = (static_cast<size_t>(p.NPAD) * 128 + 1023) & ~static_cast<size_t>(1023);
  const size_t smem = 1024 + 2 * (n_tiles * TILE_BYTES + 2 * kv_bytes) + 2 * TILE_BYTES + 2048 + 256;
  static PerDevice<size_t> cur_cache;
  size_t& cur = cur_cache.get();
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(spatial_attention_split_sm100_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(spatial_attention_split_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (e != cudaSuccess) return set_error("attention(split sm100) attr: %s", cudaGetErrorString(e));
    cur = smem;
  }
  cudaError_t le;
  if (p.N == 196 && g.K0 == 112 && g.p1_col == 208)
    le = launch_kernel_pdl(spatial_attention_split_sm100_kernel<true>, dim3(grid), dim3(SP2_THREADS), smem, st, tmIn, tmOutA,
                           tmOutB, p, g);
  else
    le = launch_kernel_pdl(spatial_attention_split_sm100_kernel<false>, dim3(grid), dim3(SP2_THREADS), smem, st, tmIn, tmOutA,
                           tmOutB, p, g);
  if (le != cudaSuccess) return set_error("attention(split sm100) launch: %s", cudaGetErrorString(le));
  return 0;
}

}  // namespace dvst
