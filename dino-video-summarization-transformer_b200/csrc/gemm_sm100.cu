// Persistent, warp-specialised tcgen05/TMEM GEMM for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] (bf16, row-major)  x  W[N,K]^T (bf16, row-major = nn.Linear layout) )
//
// Replaces every nn.Linear on the reference hot path (models/timesformer.py:41-52,63-64,71,83,108,133 and
// vision_transformer.py:254-266): both operands are K-major, so they are staged by TMA (SWIZZLE_128B) and fed
// to tcgen05.mma without any transpose.  Accumulators live in TMEM (2 x 256 fp32 columns, double buffered so the
// epilogue of tile i overlaps the MMAs of tile i+1).
//
// Two variants of one template: 1-CTA (128x256 tiles) and 2-CTA cta_group::2 pairs (256x256 tiles, W tile split
// between the two CTAs' shared memories).  Warp roles (384 threads, 1 CTA / SM, grid = #SMs, static round-robin tiles):
//   warp 0      : TMA producer (one lane)       - ring of {A 128x64, W 256x64 (or its 128-row half)} bf16 tiles
//   warp 1      : MMA issuer (one lane, leader) - 4 x tcgen05.mma (128|256)x256x16 per stage, tcgen05.commit -> mbarriers
//   warp 2      : TMEM allocator / deallocator
//   warps 4..11 : epilogue                      - tcgen05.ld 32x32b.x32 -> bias / GELU / dtype -> swizzled smem box ->
//                                                 TMA store (bf16, fp32, hi/lo pair) or TMA reduce-add (fused residual)
//
// Deferred LayerNorm: a residual GEMM (kOut = 3) reads the fp32 residual tile through TMA, adds,
// stores the new fp32 stream AND its bf16 copy `xb`, and leaves per-row partial (sum, sum of squares) of the new row in
// `stats[row][128-column block]`; the GEMM that consumes LayerNorm(x) (kLN = 1) then runs on `xb` with gamma folded into W:
//   LN(x) W^T + b = rstd_r (xb W'^T)[r,c] - rstd_r mean_r colsum(W')[c] + (b + W beta)[c],   W' = bf16(gamma * W)
// so the stand-alone LayerNorm pass (read 4 B + write 2 B per element) disappears from the block.
// Row offsets: xb is stored as bf16(x - m_r) with m_r a recent mean of row r (`xoff`, refreshed once per block from the row
// statistics), and the statistics are those of the centred values x - m_r; LayerNorm is shift invariant, so the consumer
// needs no change.  bf16 rounding and the E[x^2] - mean^2 variance then act on centred values: rows whose mean is large
// next to their spread lose nothing (with m_r = 0 the operand error would grow like |mean| / std).
//
// `split` mode: A and W hold [hi | lo] bf16 halves along K (lo at column offset K); the K loop then runs the three
// products hi*hi, hi*lo, lo*hi into the same accumulator (~16 mantissa bits) - the "precision=high" path that
// reference-exact scoring (dino_similarity.py:59-63, tau_t = 0.02) needs.  The deferred LayerNorm runs in split mode too:
// producers (kOut = 4) store xb as the [hi | lo] pair of x - m_r, consumers (kLN = 1, kOut = 2) fold the normalisation
// around the three-product accumulator with colsum = row sums of W'_hi + W'_lo.
#include "dvst_internal.h"
#include "ptx.cuh"

namespace dvst {

namespace {

constexpr int BM = 128;   // rows per CTA tile (a CTA pair covers 256 rows in 2-CTA mode)
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int ACC_STAGES = 2;
constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;
constexpr int NUM_THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;
constexpr int A_BYTES = BM * BK * 2;
constexpr int STG_BUF_BYTES = 32 * 128;  // one TMA-store box: 32 rows x 128 B (32 fp32 or 64 bf16 columns), SWIZZLE_128B
constexpr int STG_BUFS = 2;              // double buffered per epilogue warp (single-buffering for a 6th operand stage
                                         // measured slower: fc2 0.72 -> 0.78 ms)
constexpr int STG_BYTES = STG_BUFS * STG_BUF_BYTES;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512

// kCG = 1: one CTA per 128x256 tile, each CTA stages the whole 256-row W tile (48 KiB / stage, 4 stages).
// kCG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256x256 tile; each CTA stages its 128 A rows and
//          HALF of the W tile (32 KiB / stage, 6 stages) - the tensor core reads the other half from the peer's shared
//          memory, which halves the L2->SM operand traffic per FLOP (the 1-CTA kernel saturates L2 bandwidth).
// kRing: residual boxes per epilogue warp of the deferred-LayerNorm producers (kOut = 3).  3 = the default (loads two chunks
// ahead); 5 = the short-K producers (K = 768: patch-embed, temporal proj*fc, proj), which are bound by their fp32 residual
// stream, not by the tensor core: one operand stage is traded for two more 4 KiB boxes per warp so that the residual loads
// run four chunks (64 KiB per CTA) ahead (ncu: 40 % of the epilogue warps' time was the wait for a residual box).
// Split-precision producers (kOut = 4) also stage the lo half of the bf16 copy: their two 2 KiB bf16 boxes are [hi | lo] of
// ONE chunk (single buffered - the warp waits for the previous chunk's stores before it refills them) instead of the hi boxes
// of two chunks, so the shared-memory budget and the five operand stages stay (with four stages every three-product producer
// launch was 5 % slower: 514 -> 610 us at K = 768, 1 892 -> 1 990 us at K = 3 072).
template <int kCG, int kRing = 3> struct Cfg {
  static constexpr bool DEEP = kRing > 3;
  static constexpr int B_ROWS = BN / kCG;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = kCG == 1 ? 3 : (DEEP ? 4 : 5);
  static constexpr int PRODUCER_WARP_BYTES = kRing * STG_BUF_BYTES + STG_BUF_BYTES;  // ring + two 2 KiB bf16 boxes
  static constexpr int EPI_BYTES = DEEP ? 4 * PRODUCER_WARP_BYTES : EPI_WARPS * STG_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct GemmParams {
  int M, N, K;
  int split;
  const float* bias;
  int accumulate;  // fp32 output only: out += tile (TMA reduce-add in L2) - the fused residual add
  int out_kind;    // 0 = bf16, 1 = fp32, 2 = bf16 hi/lo pair (lo at column + lo_off)
  int lo_off;
  int act;  // 0 = none, 1 = exact-erf GELU
  int num_m_blocks, num_n_blocks;  // in units of (kCG * 128) x 256 tiles
  const float* ln_stats;   // kLN: fp32 [M, 6, 2] per-128-column (sum, sumsq) of the 768-wide fp32 row A was cast from
  const float* ln_colsum;  // kLN: fp32 [N] row sums of the (gamma-folded, bf16-rounded) weight
  float ln_eps;
  float* stats_out;        // producers: fp32 [M, N / 128, 2]
  const float* xoff;       // producers: per-row offset m_r subtracted before the bf16 rounding and the statistics, or null
};

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) (nn.GELU default, timesformer.py:37) = max(x, 0) - |x| h(|x|) with the Gaussian tail
// h(a) = 1 - Phi(a) = 0.5 erfc(a / sqrt 2).  log2 h is smooth (about -0.72 a^2 - 1.15 a - 1), so h = exp2(Q(min(a, 6))) with a
// degree-6 polynomial Q (weighted least-squares fit on Chebyshev nodes of [0, 6], tools/gelu_fit.py): |GELU error| <= 1.8e-7
// in fp32 arithmetic - below the Abramowitz-Stegun 7.1.26 form used before (3.3e-7 in fp32: 1 MUFU.RCP + 1 MUFU.EX2 + 9 FP32
// ops) at 1 MUFU.EX2 + 7 FMA + 3 min / max per element; after bf16 rounding 9e-5 of the outputs on [-3, 3] differ from the
// rounding of the exact value (A-S: 3.8e-4).  Beyond |x| = 6 the tail (< 1e-9) is held constant.
// Sustained (power-capped) fc1 + deferred LayerNorm + GELU: 0.922 -> 0.876 ms per cfg2 launch (0.767 without GELU,
// tools/sustained_gemm.py).  Measured and not kept: L1 prefetch of the tile's bias / column sums (no change), L1 evict_last /
// no_allocate policies on the epilogue loads (slower); with the MUFU op AND the polynomial compiled out the launch still takes
// 0.83 ms - most of what is left of the GELU cost is not its arithmetic.
constexpr float GELU_C0 = -0.9999954104423523f, GELU_C1 = -1.1511955261230469f, GELU_C2 = -0.45875775814056396f,
                GELU_C3 = -0.05345946177840233f, GELU_C4 = 0.008125263266265392f, GELU_C5 = -0.0007859655888751149f,
                GELU_C6 = 3.5286495403852314e-05f;
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fminf(fabsf(x), 6.0f);
  float q = fmaf(GELU_C6, a, GELU_C5);
  q = fmaf(q, a, GELU_C4);
  q = fmaf(q, a, GELU_C3);
  q = fmaf(q, a, GELU_C2);
  q = fmaf(q, a, GELU_C1);
  q = fmaf(q, a, GELU_C0);
  return fmaf(-fabsf(x), ex2_approx(q), fmaxf(x, 0.0f));
}

// Packed fp32x2 arithmetic (sm_100 FFMA2): two IEEE fp32 FMAs per issued instruction.  The GELU epilogue of fc1 is
// issue-bound next to the 256x256x768 MMA, so the polynomial runs on register pairs; the MUFU op stays scalar.  Same
// operations in the same order as gelu_erf(), i.e. bit-identical results.
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
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t splat2(float v) { return pk2(v, v); }
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x2) {
  float x0, x1;
  unpk2(x2, x0, x1);
  const uint64_t a2 = pk2(fminf(fabsf(x0), 6.0f), fminf(fabsf(x1), 6.0f));
  uint64_t q2 = fma2(splat2(GELU_C6), a2, splat2(GELU_C5));
  q2 = fma2(q2, a2, splat2(GELU_C4));
  q2 = fma2(q2, a2, splat2(GELU_C3));
  q2 = fma2(q2, a2, splat2(GELU_C2));
  q2 = fma2(q2, a2, splat2(GELU_C1));
  q2 = fma2(q2, a2, splat2(GELU_C0));
  float q0, q1;
  unpk2(q2, q0, q1);
  const uint64_t nax2 = x2 | 0x8000000080000000ull;  // -|x|
  return fma2(nax2, pk2(ex2_approx(q0), ex2_approx(q1)), pk2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// kOut: 0 = bf16, 1 = fp32 (plain or reduce-add), 2 = bf16 hi/lo pair.  kAct: 0 = none, 1 = exact-erf GELU.  Compile-time
// so the epilogue (issue-bound next to the 256x256x768 MMA of fc1) carries no predicated code of the other modes.
// kOut = 3: fp32 in-place residual (TMA load + add + TMA store) + bf16 copy + row statistics (deferred LayerNorm producer);
// kOut = 4: the same with the copy as a bf16 [hi | lo] pair (split precision; lo at column + lo_off of xb).
// kLN = 1: deferred LayerNorm consumer (bf16 output, or the hi/lo pair in split precision).
template <int kCG, int kOut, int kAct, int kLN, int kRing = 3>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmXb,
                    const GemmParams p) {
  constexpr bool kProducer = kOut == 3 || kOut == 4;
  constexpr bool kXbLo = kOut == 4;
  constexpr int EW = kProducer ? 4 : EPI_WARPS;  // epilogue warps that take part
  using C = Cfg<kCG, kRing>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * A_BYTES;
  const uint32_t stg_all = base + STAGES * C::STAGE_BYTES;  // 1024-byte aligned
  const uint32_t bars = base + STAGES * C::STAGE_BYTES + C::EPI_BYTES;
  const uint32_t full_bar = bars;                       // [STAGES]   (2-CTA: only the leader's copy is used)
  const uint32_t empty_bar = bars + 8 * STAGES;         // [STAGES]
  const uint32_t tfull_bar = bars + 16 * STAGES;        // [ACC_STAGES]
  const uint32_t tempty_bar = tfull_bar + 8 * ACC_STAGES;  // (2-CTA: only the leader's copy is used)
  const uint32_t tmem_slot = tempty_bar + 8 * ACC_STAGES;
  const uint32_t res_bar = tmem_slot + 8;                // producers: [EW warps][kRing ring slots]
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kCG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if constexpr (kProducer) tma_prefetch_desc(&tmXb);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, EW * kCG);
    }
    if constexpr (kProducer) {
      for (int i = 0; i < kRing * EW; ++i) mbar_init(res_bar + 8 * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<kCG>(tmem_slot, TMEM_COLS);
    tmem_relinquish<kCG>();
  }
  tc_fence_before();
  if (kCG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // nothing above touched global memory: the prologue may overlap the previous kernel's tail
  pdl_launch_dependents();

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  const int tile0 = kCG == 2 ? blockIdx.x / 2 : blockIdx.x;
  const int tile_step = kCG == 2 ? gridDim.x / 2 : gridDim.x;
  const int nk = p.K / BK;
  const int num_iters = p.split ? 3 * nk : nk;

  if (warp == 0) {
    {
      // (the whole warp walks the loop; one elected lane issues - see elect_one() in ptx.cuh)
      int s = 0;
      uint32_t ph = 0;
      // TMA completion bytes of BOTH CTAs are signalled on the leader's `full` barrier
      const uint32_t full_leader = kCG == 2 ? mapa_u32(full_bar, 0) : full_bar;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_blk = tile / p.num_n_blocks;
        const int n_blk = tile % p.num_n_blocks;
        const int row_a = (m_blk * kCG + cta_rank) * BM;
        const int row_b = n_blk * BN + cta_rank * C::B_ROWS;
        for (int it = 0; it < num_iters; ++it) {
          int ka, kb;
          if (p.split) {
            const int kblk = it / 3, term = it - 3 * kblk;  // 0: hi*hi, 1: hi*lo, 2: lo*hi
            ka = kblk * BK + (term == 2 ? p.K : 0);
            kb = kblk * BK + (term == 1 ? p.K : 0);
          } else {
            ka = kb = it * BK;
          }
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          if (elect_one()) {
            if (kCG == 1) {
              mbar_arrive_expect_tx(full_bar + 8 * s, C::STAGE_BYTES);
              tma_load_2d(sA + s * A_BYTES, &tmA, full_bar + 8 * s, ka, row_a);
              tma_load_2d(sB + s * C::B_BYTES, &tmB, full_bar + 8 * s, kb, row_b);
            } else {
              if (leader) mbar_arrive_expect_tx(full_bar + 8 * s, 2 * C::STAGE_BYTES);
              tma_load_2d_cg2(sA + s * A_BYTES, &tmA, full_leader + 8 * s, ka, row_a);
              tma_load_2d_cg2(sB + s * C::B_BYTES, &tmB, full_leader + 8 * s, kb, row_b);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM * kCG, BN);
      int s = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar + 8 * a, aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + a * BN;
        for (int it = 0; it < num_iters; ++it) {
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_addr = sA + s * A_BYTES;
          const uint32_t b_addr = sB + s * C::B_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_bf16<kCG>(tmem_d, make_kmajor_sw128_desc(a_addr + k * 32), make_kmajor_sw128_desc(b_addr + k * 32),
                             idesc, (it | k) != 0 ? 1u : 0u);
            }
            if (kCG == 1) umma_commit(empty_bar + 8 * s); else umma_commit_cg2(empty_bar + 8 * s, 3);
            if (it == num_iters - 1) {
              if (kCG == 1) umma_commit(tfull_bar + 8 * a); else umma_commit_cg2(tfull_bar + 8 * a, 3);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++a == ACC_STAGES) { a = 0; aph ^= 1; }
      }
    }
  } else if (warp >= FIRST_EPI_WARP) {
    // TMEM -> registers -> (bias, GELU, dtype) -> swizzled smem box -> TMA store (or TMA reduce-add for the residual):
    // each output byte crosses the LSU exactly once (one 128-bit STS); the global write is done by the async proxy.
    const int e = warp - FIRST_EPI_WARP;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int half = e >> 2;  // which 128-column half of the accumulator
    const uint32_t stg = stg_all + e * STG_BYTES;
    const uint32_t row_off = lane * 128;
    const int sw = lane & 7;
    int buf = 0;
    int a = 0;
    uint32_t aph = 0;

    auto stage_and_store = [&](const uint32_t (&w)[32], int gcol, int grow) {
      // the TMA store issued STG_BUFS boxes ago must have finished reading this buffer
      if (elect_one()) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STG_BUFS - 1) : "memory");
      __syncwarp();
      const uint32_t dst = stg + buf * STG_BUF_BYTES + row_off;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((j ^ sw) << 4)), "r"(w[4 * j]),
                     "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3]) : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        const uint32_t src = stg + buf * STG_BUF_BYTES;
        if (p.accumulate)
          asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(src), "r"(gcol), "r"(grow) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(src), "r"(gcol), "r"(grow) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      buf = (buf + 1) % STG_BUFS;
    };

    if constexpr (kProducer) {
      // ---- deferred-LayerNorm producer: x <- x + acc + bias (fp32, in place), xb <- bf16(x), stats <- row partials.
      // 4 warps; warp e owns rows q*32.. of the tile and walks its 256 columns in 8 chunks of 32.  Per warp: a ring of
      // three 4 KiB boxes (TMA load of the residual chunk -> add in place -> TMA store from the same box) and two 2 KiB
      // bf16 boxes (SWIZZLE_64B).  Residual loads run two chunks ahead, across tile boundaries.  One (sum, sumsq)
      // partial per 128 columns.  (Measured and dropped: 8 warps x 4 chunks with 3 operand stages for the K = 768
      // producers - twice the residual bytes in flight - gave the same step time, 54.14 vs 54.16 ms.)
      constexpr int NCH = 8;
      constexpr int col_base = 0;
      if (e < EW) {
        const uint32_t ring = stg_all + e * C::PRODUCER_WARP_BYTES;
        const uint32_t bst = ring + kRing * STG_BUF_BYTES;
        const uint32_t rbar = res_bar + e * (8 * kRing);
        const uint32_t brow_off = lane * 64;
        const int bsw = (lane >> 1) & 3;
        int l_tile = tile0, l_c = 0;
        uint32_t l_g = 0;
        auto issue_load = [&]() {  // lane 0 only
          if (l_tile < num_tiles) {
            const int lm = l_tile / p.num_n_blocks, ln = l_tile % p.num_n_blocks;
            const uint32_t sl = l_g % kRing;
            mbar_arrive_expect_tx(rbar + 8 * sl, STG_BUF_BYTES);
            tma_load_2d(ring + sl * STG_BUF_BYTES, &tmO, rbar + 8 * sl, ln * BN + col_base + l_c * 32,
                        (lm * kCG + static_cast<int>(cta_rank)) * BM + q * 32);
            ++l_g;
            if (++l_c == NCH) { l_c = 0; l_tile += tile_step; }
          }
        };
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < kRing - 1; ++i) issue_load();
        }
        uint32_t g = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
          const int m_blk = tile / p.num_n_blocks;
          const int n_blk = tile % p.num_n_blocks;
          const int grow = (m_blk * kCG + cta_rank) * BM + q * 32;
          mbar_wait(tfull_bar + 8 * a, aph);
          tc_fence_after();
          const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN + col_base;
          uint32_t v[32];
          tmem_ld_32x32(trow, v);
          uint64_t s1 = 0ull, s2 = 0ull;  // packed (even, odd column) partial sums
          const float moff = (p.xoff != nullptr && grow + lane < p.M) ? __ldg(p.xoff + grow + lane) : 0.f;
          const uint64_t nmoff2 = splat2(-moff);
          float2* stat_row = reinterpret_cast<float2*>(p.stats_out) +
                             (static_cast<long long>(grow + lane) * (2 * p.num_n_blocks) + 2 * n_blk + (col_base >> 7));
#pragma unroll 1
          for (int c = 0; c < NCH; ++c, ++g) {
            const int gcol0 = n_blk * BN + col_base + c * 32;
            const uint32_t sl = g % kRing;
            mbar_wait(rbar + 8 * sl, (g / kRing) & 1);
            tmem_ld_wait();
            float w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __uint_as_float(v[j]);
            if (c < NCH - 1) tmem_ld_32x32(trow + (c + 1) * 32, v);
            else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (kCG == 1 || leader) mbar_arrive(tempty_bar + 8 * a);
                else mbar_arrive_cluster(tempty_bar + 8 * a, 0);
              }
            }
            const uint32_t box = ring + sl * STG_BUF_BYTES + row_off;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t addr = box + ((j ^ sw) << 4);
              float4 r4;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r4.x), "=f"(r4.y), "=f"(r4.z), "=f"(r4.w)
                           : "r"(addr) : "memory");
              const float4 b4 = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + gcol0) + j)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
              const uint64_t wa = add2(add2(pk2(w[4 * j], w[4 * j + 1]), pk2(b4.x, b4.y)), pk2(r4.x, r4.y));
              const uint64_t wb = add2(add2(pk2(w[4 * j + 2], w[4 * j + 3]), pk2(b4.z, b4.w)), pk2(r4.z, r4.w));
              const uint64_t ca = add2(wa, nmoff2), cb = add2(wb, nmoff2);  // centred: bf16 copy and statistics
              s1 = add2(add2(s1, ca), cb);
              s2 = fma2(cb, cb, fma2(ca, ca, s2));
              unpk2(wa, w[4 * j], w[4 * j + 1]);
              unpk2(wb, w[4 * j + 2], w[4 * j + 3]);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(w[4 * j]), "f"(w[4 * j + 1]),
                           "f"(w[4 * j + 2]), "f"(w[4 * j + 3]) : "memory");
              unpk2(ca, w[4 * j], w[4 * j + 1]);
              unpk2(cb, w[4 * j + 2], w[4 * j + 3]);
            }
            const uint32_t bsrc = bst + (kXbLo ? 0u : (g & 1) * (STG_BUF_BYTES / 2));  // TMA-store source of the hi box
            const uint32_t bbox = bsrc + brow_off;
            if constexpr (kXbLo) {  // single-buffered [hi | lo] boxes: the previous chunk's stores must have read them
              if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t h4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) h4[i] = pack_bf16x2(w[8 * j + 2 * i], w[8 * j + 2 * i + 1]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bbox + ((j ^ bsw) << 4)), "r"(h4[0]), "r"(h4[1]),
                           "r"(h4[2]), "r"(h4[3]) : "memory");
              if constexpr (kXbLo) {  // lo = bf16(c - float(hi)), second box
                uint32_t l4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  l4[i] = pack_bf16x2(w[8 * j + 2 * i] - __uint_as_float(h4[i] << 16),
                                      w[8 * j + 2 * i + 1] - __uint_as_float(h4[i] & 0xFFFF0000u));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bbox + STG_BUF_BYTES / 2 + ((j ^ bsw) << 4)),
                             "r"(l4[0]), "r"(l4[1]), "r"(l4[2]), "r"(l4[3]) : "memory");
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(ring + sl * STG_BUF_BYTES), "r"(gcol0), "r"(grow)
                           : "memory");
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmXb)), "r"(bsrc), "r"(gcol0), "r"(grow) : "memory");
              if constexpr (kXbLo)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmXb)),
                               "r"(bsrc + STG_BUF_BYTES / 2), "r"(gcol0 + p.lo_off), "r"(grow) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              // the stores of the previous chunk have finished reading their boxes: its ring slot takes the load that
              // runs kRing - 1 chunks ahead, its bf16 box is the next chunk's
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              issue_load();
            }
            __syncwarp();
            if ((c & 3) == 3) {  // 128 columns done: one partial
              float s1a, s1b, s2a, s2b;
              unpk2(s1, s1a, s1b);
              unpk2(s2, s2a, s2b);
              if (grow + lane < p.M) stat_row[c >> 2] = make_float2(s1a + s1b, s2a + s2b);
              s1 = 0ull;
              s2 = 0ull;
            }
          }
          if (++a == ACC_STAGES) { a = 0; aph ^= 1; }
        }
        if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
    } else {
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = tile / p.num_n_blocks;
      const int n_blk = tile % p.num_n_blocks;
      const int grow = (m_blk * kCG + cta_rank) * BM + q * 32;
      float ln_rstd = 0.f, ln_nmr = 0.f;  // deferred LayerNorm: rstd and -rstd * mean of this thread's row
      if constexpr (kLN == 1) {
        if (grow + lane < p.M) {
          const float4* sp = reinterpret_cast<const float4*>(p.ln_stats) + static_cast<long long>(grow + lane) * 3;
          const float4 a0 = __ldg(sp), a1 = __ldg(sp + 1), a2 = __ldg(sp + 2);  // 6 x (sum, sumsq) per 128 columns
          const float mean = (((a0.x + a0.z) + (a1.x + a1.z)) + (a2.x + a2.z)) * (1.0f / 768.0f);
          const float var = (((a0.y + a0.w) + (a1.y + a1.w)) + (a2.y + a2.w)) * (1.0f / 768.0f) - mean * mean;
          ln_rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
          ln_nmr = -ln_rstd * mean;  // (mean and variance of the centred row A was cast from)
        }
      }
      mbar_wait(tfull_bar + 8 * a, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN + half * 128;
      const int gcol_base = n_blk * BN + half * 128;
      uint32_t v[32], v2[32];
      tmem_ld_32x32(trow, v);
      if constexpr (kOut == 1) {
        // ---- fp32 output: 4 boxes of 32 columns ----
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int gcol0 = gcol_base + c * 32;
          tmem_ld_wait();
          uint32_t w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] = v[j];
          if (c < 3) tmem_ld_32x32(trow + (c + 1) * 32, v);
          else {  // accumulator fully read: hand the TMEM stage back to the MMA warp before the stores drain
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (kCG == 1 || leader) mbar_arrive(tempty_bar + 8 * a);
              else mbar_arrive_cluster(tempty_bar + 8 * a, 0);
            }
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol0) + j4);
              w[4 * j4] = __float_as_uint(__uint_as_float(w[4 * j4]) + b4.x);
              w[4 * j4 + 1] = __float_as_uint(__uint_as_float(w[4 * j4 + 1]) + b4.y);
              w[4 * j4 + 2] = __float_as_uint(__uint_as_float(w[4 * j4 + 2]) + b4.z);
              w[4 * j4 + 3] = __float_as_uint(__uint_as_float(w[4 * j4 + 3]) + b4.w);
            }
          }
          if constexpr (kAct == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(gelu_erf(__uint_as_float(w[j])));
          }
          stage_and_store(w, gcol0, grow);
        }
      } else {
        // ---- bf16 output (optionally hi/lo pair): 2 boxes of 64 columns; the two 32-column register sets ping-pong so
        // the next tcgen05.ld is in flight while the current set goes through bias / GELU / pack ----
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int gcol0 = gcol_base + c * 64;
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t (&cur)[32] = hh == 0 ? v : v2;
            uint32_t (&nxt)[32] = hh == 0 ? v2 : v;
            tmem_ld_wait();
            if (c * 2 + hh < 3) tmem_ld_32x32(trow + (c * 2 + hh + 1) * 32, nxt);
            else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (kCG == 1 || leader) mbar_arrive(tempty_bar + 8 * a);
                else mbar_arrive_cluster(tempty_bar + 8 * a, 0);
              }
            }
            uint64_t f2[16];  // register pairs (columns 2j, 2j + 1)
#pragma unroll
            for (int j = 0; j < 16; ++j) f2[j] = pk2(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1]));
            const uint64_t rstd2 = splat2(ln_rstd), nmr2 = splat2(ln_nmr);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {  // branch-free: a null bias adds zeros
              const float4 b4 = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + gcol0 + hh * 32) + j4)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
              if constexpr (kLN == 1) {
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + gcol0 + hh * 32) + j4);
                f2[2 * j4] = fma2(rstd2, f2[2 * j4], fma2(nmr2, pk2(s4.x, s4.y), pk2(b4.x, b4.y)));
                f2[2 * j4 + 1] = fma2(rstd2, f2[2 * j4 + 1], fma2(nmr2, pk2(s4.z, s4.w), pk2(b4.z, b4.w)));
              } else {
                f2[2 * j4] = add2(f2[2 * j4], pk2(b4.x, b4.y));
                f2[2 * j4 + 1] = add2(f2[2 * j4 + 1], pk2(b4.z, b4.w));
              }
            }
            if constexpr (kAct == 1) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f2[j] = gelu_erf2(f2[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float fa, fb;
              unpk2(f2[j], fa, fb);
              const uint32_t hw = pack_bf16x2(fa, fb);
              hi[hh * 16 + j] = hw;
              if constexpr (kOut == 2) {
                const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&hw);
                lo[hh * 16 + j] = pack_bf16x2(fa - __low2float(hb), fb - __high2float(hb));
              }
            }
          }
          stage_and_store(hi, gcol0, grow);
          if constexpr (kOut == 2) stage_and_store(lo, gcol0 + p.lo_off, grow);
        }
      }
      if (++a == ACC_STAGES) { a = 0; aph ^= 1; }
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    }  // not a producer
  }

  tc_fence_before();
  if (kCG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kCG>(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols, bool swizzle64 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error("cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (code %d)", static_cast<int>(r));
  return 0;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(map, ptr, 2, rows, cols, ld_elems, box_rows, box_cols);
}

int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error("cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (3-D) failed (code %d)", static_cast<int>(r));
  return 0;
}

int device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

int num_sms() {
  static PerDevice<int> cache;
  int& n = cache.get();
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static int g_producer_ring = 5;       // 3: never the deep ring, 5: for K <= 1024, 6: for every K (experiment)
void set_producer_ring(int r) { g_producer_ring = r <= 3 ? 3 : (r >= 6 ? 6 : 5); }
static int g_gemm_cta_group = 2;
void set_gemm_cta_group(int cg) { g_gemm_cta_group = cg == 1 ? 1 : 2; }

template <int kCG, int kOut, int kAct, int kLN, int kRing = 3>
static int launch_gemm_impl(const GemmArgs& g, cudaStream_t stream) {
  using C = Cfg<kCG, kRing>;
  static PerDevice<bool> attr_cache;
  bool& attr_set = attr_cache.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<kCG, kOut, kAct, kLN, kRing>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return set_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const uint64_t kcols = static_cast<uint64_t>(g.K) * (g.split ? 2 : 1);
  CUtensorMap tmA, tmB, tmO;
  if (int rc = make_tmap_bf16_2d(&tmA, g.A, g.M, kcols, g.lda, BM, BK)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmB, g.W, g.N, kcols, g.ldw, C::B_ROWS, BK)) return rc;
  const uint64_t out_cols = g.out_kind == 2 ? static_cast<uint64_t>(g.lo_off) + g.N : g.N;
  const bool f32out = g.out_kind == 1 || g.out_kind == 3;
  if (int rc = make_tmap_2d(&tmO, g.out, f32out ? 4 : 2, g.M, out_cols, g.ldo, 32, f32out ? 32 : 64)) return rc;
  CUtensorMap tmXb = tmO;
  if (kOut == 3 || kOut == 4)
    if (int rc = make_tmap_2d(&tmXb, g.xb, 2, g.M, kOut == 4 ? static_cast<uint64_t>(g.lo_off) + g.N : g.N, g.ldxb, 32, 32,
                              true))
      return rc;

  GemmParams p;
  p.M = g.M; p.N = g.N; p.K = g.K; p.split = g.split;
  p.bias = g.bias; p.accumulate = g.residual != nullptr ? 1 : 0;
  p.out_kind = g.out_kind; p.lo_off = static_cast<int>(g.lo_off); p.act = g.act;
  p.num_m_blocks = (g.M + BM * kCG - 1) / (BM * kCG);
  p.num_n_blocks = g.N / BN;
  p.ln_stats = g.ln_stats; p.ln_colsum = g.ln_colsum; p.ln_eps = g.ln_eps; p.stats_out = g.stats_out;
  p.xoff = g.xoff;
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int max_units = num_sms() / kCG;
  const int units = tiles < max_units ? tiles : max_units;

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * kCG);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  const bool hbm_bound = kOut == 3 && g.K <= 1024;   // (the three-product split producers, kOut = 4, are tensor bound)
  const double hbm_bytes = static_cast<double>(g.M) * (2.0 * g.K + 10.0 * g.N) + 2.0 * g.N * g.K;  // A, x in / out, xb, W
  ProfScope prof(hbm_bound ? PROF_GEMM_HBM : PROF_GEMM,
                 hbm_bound ? hbm_bytes : 2.0 * g.M * g.N * g.K * (g.split ? 3.0 : 1.0), stream);
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<kCG, kOut, kAct, kLN, kRing>, tmA, tmB, tmO, tmXb, p);
  if (e != cudaSuccess) return set_error("gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0) return 0;
  if (g.N % BN != 0) return set_error("gemm: N=%d must be a multiple of %d", g.N, BN);
  if (g.K % BK != 0) return set_error("gemm: K=%d must be a multiple of %d", g.K, BK);
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15) || (g.lda % 8) ||
      (g.ldw % 8))
    return set_error("gemm: operands must be 16-byte aligned with leading dimensions multiple of 8");
  if (g.out_kind < 0 || g.out_kind > 3) return set_error("gemm: bad out_kind %d", g.out_kind);
  const bool f32out = g.out_kind == 1 || g.out_kind == 3;
  if ((reinterpret_cast<uintptr_t>(g.out) & 15) || (g.ldo % (f32out ? 4 : 8)) || (g.lo_off % 8))
    return set_error("gemm: output must be 16-byte aligned with a 16-byte multiple row pitch");
  if (g.residual != nullptr && (g.residual != g.out || !f32out || g.ldr != g.ldo))
    return set_error("gemm: the fused residual add is in place (residual must alias the fp32 output)");
  if (g.out_kind == 3) {
    // deferred-LayerNorm producer: x += A W^T + b in place, xb = bf16(x), stats[row][n_block] = (sum, sumsq)
    if (g.residual == nullptr || g.xb == nullptr || g.stats_out == nullptr || g.act != 0)
      return set_error("gemm: out_kind 3 needs residual (in place), xb, stats_out, no activation");
    if (g.split != 0 && (g.lo_off < g.N || g.lo_off + g.N > g.ldxb))
      return set_error("gemm: split-precision out_kind 3 writes xb as [hi | lo]: N <= lo_off <= ldxb - N");
    if ((reinterpret_cast<uintptr_t>(g.xb) & 15) || (g.ldxb % 8) || (reinterpret_cast<uintptr_t>(g.stats_out) & 15))
      return set_error("gemm: xb and stats must be 16-byte aligned, ldxb a multiple of 8");
  }
  if (g.ln_stats != nullptr) {
    if (g.out_kind != (g.split ? 2 : 0) || g.K != 768 || g.ln_colsum == nullptr ||
        (reinterpret_cast<uintptr_t>(g.ln_stats) & 15) || (reinterpret_cast<uintptr_t>(g.ln_colsum) & 15))
      return set_error("gemm: deferred LayerNorm needs K = 768, a column-sum vector and bf16 output (the hi/lo pair in "
                       "split precision)");
  }
  // small problems (fewer 256-row tiles than SM pairs would leave half the chip idle) stay on the 1-CTA kernel
  const long long tiles2 = static_cast<long long>((g.M + 2 * BM - 1) / (2 * BM)) * (g.N / BN);
  if (g.act < 0 || g.act > 1) return set_error("gemm: bad act %d", g.act);
  const bool cg2 = g_gemm_cta_group == 2 && tiles2 >= num_sms() / 2;
  if (g.ln_stats != nullptr) {
    if (g.split) {
      if (g.act == 0) return cg2 ? launch_gemm_impl<2, 2, 0, 1>(g, stream) : launch_gemm_impl<1, 2, 0, 1>(g, stream);
      return cg2 ? launch_gemm_impl<2, 2, 1, 1>(g, stream) : launch_gemm_impl<1, 2, 1, 1>(g, stream);
    }
    if (g.act == 0) return cg2 ? launch_gemm_impl<2, 0, 0, 1>(g, stream) : launch_gemm_impl<1, 0, 0, 1>(g, stream);
    return cg2 ? launch_gemm_impl<2, 0, 1, 1>(g, stream) : launch_gemm_impl<1, 0, 1, 1>(g, stream);
  }
  if (g.out_kind == 3) {
    if (g.split) return cg2 ? launch_gemm_impl<2, 4, 0, 0>(g, stream) : launch_gemm_impl<1, 4, 0, 0>(g, stream);
    if (!cg2) return launch_gemm_impl<1, 3, 0, 0>(g, stream);
    // short K: bound by the residual stream - deeper residual prefetch, one operand stage fewer
    return (g.K <= 1024 && g_producer_ring > 3) || g_producer_ring == 6 ? launch_gemm_impl<2, 3, 0, 0, 5>(g, stream)
                                                                         : launch_gemm_impl<2, 3, 0, 0>(g, stream);
  }
#define DVST_GEMM_CASE(OUT, ACT)                                                                   \
  if (g.out_kind == OUT && g.act == ACT)                                                           \
    return cg2 ? launch_gemm_impl<2, OUT, ACT, 0>(g, stream) : launch_gemm_impl<1, OUT, ACT, 0>(g, stream);
  DVST_GEMM_CASE(0, 0)
  DVST_GEMM_CASE(0, 1)
  DVST_GEMM_CASE(1, 0)
  DVST_GEMM_CASE(1, 1)
  DVST_GEMM_CASE(2, 0)
  DVST_GEMM_CASE(2, 1)
#undef DVST_GEMM_CASE
  return set_error("gemm: unsupported epilogue");
}

}  // namespace dvst
