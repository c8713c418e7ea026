// Temporal half of a divided space-time block in ONE kernel (bf16 precision, deferred LayerNorm):
//
//   a[rows of (b, n), head h] = softmax(q k^T / 8) v     with   [q | k | v] = LayerNorm(x) W_qkv^T + b
//
// i.e. temporal_norm1 -> temporal_attn.qkv -> the attention of temporal_attn (models/timesformer.py:127-131 with
// Attention.forward :68-81), up to (not including) temporal_attn.proj.  The reference materialises the [B*N, T, 2304] qkv
// tensor, rearranges it to heads and runs two batched matmuls; here the packed qkv tensor never exists:
//
//   * the weight rows are permuted at pack time so that one 192-column N tile is exactly [q_h | k_h | v_h] of ONE head;
//   * an M tile of the resident (b, n, t) row order is R = floor(128 / T) * T consecutive token rows = a whole number of
//     (b, n) sequences (T = 16: 8 sequences, T = 30: 4, T = 3: 42; tattn_geometry below), optionally tiled clip by clip, so a
//     (128-row, 192-column) accumulator tile holds everything the temporal attention of those sequences and that head needs;
//   * drain warps apply the folded LayerNorm (gemm_sm100.cu) and round q / k / v to bf16 into a swizzled staging tile in
//     SHARED MEMORY; attention warps run the attention of 16 query rows each with mma.sync m16n8k16 straight out of that
//     tile (exactly the arithmetic of temporal_attention_kernel, attention.cu - for power-of-two T the results are
//     bit-identical to the two-kernel form) and TMA-store the 16 x 64 output box into the A operand of the temporal
//     proj*fc GEMM.
//
// Per token and block this removes the 4 608-byte qkv write and read (the qkv GEMM is tensor bound, the attention kernel
// HBM bound: 1.85 GB per cfg2 launch pair) and one launch; the main loop is the 2-CTA tcgen05 pipeline of gemm_sm100.cu
// with 256 x 192 tiles (TMEM: 2 x 192 of 512 columns).
//
// Warp roles (640 threads, CTA pairs, persistent): warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM
// allocator, warps 4..11 drain, warps 12..19 attention.  Per tile and CTA (its own 128 accumulator rows):
//   drain      warp e: TMEM lane quarter e & 3, columns [96 (e >> 2), +96) -> LayerNorm fold -> bf16 -> staging buffer
//              (tile parity) [128 rows][q 128 B | k 128 B | v 128 B], 16-byte chunks XOR-swizzled with (row & 7); the TMEM
//              stage goes back to the MMA warp, the staging buffer to the attention warps (mbarrier)
//   attention  warp w: query rows [16 w, 16 w + 16) as one mma.sync tile against the key rows of the sequences those rows
//              belong to (a 16-, 32- or 64-key block; keys of other sequences masked), output -> swizzled 16 x 128 B box ->
//              TMA store; staging buffer handed back
// The two roles overlap (double-buffered staging): the legacy HMMA instructions of the attention only issue in the gaps
// between the main loop's tcgen05.mma instructions (ncu of the first version, where the drain warps also ran the attention:
// 42 % of their time waiting at HMMA, epilogue-bound at 0.73 ms per cfg2 launch) - dedicated warps take that wait off the
// accumulator-drain path (4 warps: 0.63 ms, 8 warps: 0.60 ms; 0.48 ms with the attention compiled out,
// -DDVST_TATTN_EXP_NO_ATTENTION).
#include "attn_sm100_common.cuh"
#include "dvst_internal.h"
#include "mma_sync.cuh"
#include "ptx.cuh"

namespace dvst {

namespace {

using namespace mma;
using attn::add2;
using attn::fma2;
using attn::named_bar_sync;
using attn::pk2;
using attn::unpk2;

constexpr int BM = 128;        // accumulator rows per CTA
constexpr int BK = 64;
constexpr int HN = 192;        // q | k | v columns of one head
constexpr int KD = 768;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int ACC_STRIDE = 256;                 // TMEM columns between the two accumulator stages
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 8;       // drain warps
constexpr int FIRST_EPI_WARP = 4;
constexpr int ATT_WARPS = 8;
constexpr int FIRST_ATT_WARP = FIRST_EPI_WARP + EPI_WARPS;
constexpr int NUM_THREADS = (FIRST_ATT_WARP + ATT_WARPS) * 32;
constexpr int A_BYTES = BM * BK * 2;            // 16 KiB
constexpr int B_BYTES = (HN / 2) * BK * 2;      // 12 KiB: this CTA's half of the W tile
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int QKV_PITCH = 3 * 128;              // staging row: q | k | v, 128 bytes each
constexpr int QKV_BYTES = BM * QKV_PITCH;       // 48 KiB
constexpr int OBOX_BYTES = 16 * 128;            // one warp's output box
constexpr int NUM_OBOX = ATT_WARPS;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + ACC_STAGES * QKV_BYTES + NUM_OBOX * OBOX_BYTES + 256;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

struct TAttnParams {
  int M;             // patch-token rows
  int T, R;          // frames per sequence, rows per CTA tile (whole sequences)
  int num_m_blocks;  // CTA-pair tiles (two CTA tiles each)
  int clip_rows;     // rows per clip (tiles never straddle a clip: every clip is tiled the same way), a multiple of T
  int tiles_per_clip;  // ceil(clip_rows / R) CTA tiles
  const float* bias;       // fp32 [2304], head-permuted: b + W beta
  const float* ln_colsum;  // fp32 [2304], head-permuted row sums of the folded weight
  const float* ln_stats;   // fp32 [M, 6, 2]
  float ln_eps;
  __nv_bfloat16* out;      // [M, ldo] (plain stores of partially valid boxes)
  long long ldo;
};

// First row and number of rows of CTA tile `ct`: clip ct / tiles_per_clip, R rows per tile, the clip's last tile shorter.
__device__ __forceinline__ void cta_tile_rows(const TAttnParams& p, int ct, int& row0, int& rows) {
  const int clip = ct / p.tiles_per_clip, j = ct - clip * p.tiles_per_clip;
  row0 = clip * p.clip_rows + j * p.R;
  rows = p.clip_rows - j * p.R;
  if (rows > p.R) rows = p.R;
}

// MAXP: 16-key pairs of n-tiles in the key block (key block = 16 MAXP keys).
template <int MAXP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_ln_tattn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const TAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * A_BYTES;
  const uint32_t sQKV = base + STAGES * STAGE_BYTES;
  const uint32_t sO = sQKV + ACC_STAGES * QKV_BYTES;
  const uint32_t bars = sO + NUM_OBOX * OBOX_BYTES;
  const uint32_t full_bar = bars;                          // [STAGES]   (only the leader's copy is used)
  const uint32_t empty_bar = bars + 8 * STAGES;            // [STAGES]
  const uint32_t tfull_bar = bars + 16 * STAGES;           // [ACC_STAGES]
  const uint32_t tempty_bar = tfull_bar + 8 * ACC_STAGES;  // (only the leader's copy is used)
  const uint32_t sfull_bar = tempty_bar + 8 * ACC_STAGES;   // [ACC_STAGES] staging buffer written (drain -> attention)
  const uint32_t sempty_bar = sfull_bar + 8 * ACC_STAGES;   // [ACC_STAGES] staging buffer read (attention -> drain)
  const uint32_t tmem_slot = sempty_bar + 8 * ACC_STAGES;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, EPI_WARPS * 2);
      mbar_init(sfull_bar + 8 * a, EPI_WARPS);
      mbar_init(sempty_bar + 8 * a, ATT_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<2>(tmem_slot, TMEM_COLS);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // nothing above touched global memory: the prologue may overlap the previous kernel's tail
  pdl_launch_dependents();

  constexpr int NUM_HEADS = 12;
  const int num_tiles = p.num_m_blocks * NUM_HEADS;
  const int tile0 = blockIdx.x / 2;
  const int tile_step = gridDim.x / 2;
  constexpr int NK = KD / BK;

  // Register budget per role (640 threads -> 96 by default, i.e. a pool of 61 440 for the CTA): the producer / MMA /
  // allocator warpgroup gives registers back (40), the two drain warpgroups (a 32-column accumulator slice in flight next
  // to the one being folded and packed) take them (120), the attention warpgroups keep 96: 128 x 40 + 256 x 120 + 256 x 96
  // = 60 416 <= 61 440 (a request beyond the pool blocks for ever).
  if (warp < FIRST_EPI_WARP) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    const uint32_t full_leader = mapa_u32(full_bar, 0);  // TMA bytes of BOTH CTAs are signalled on the leader's barrier
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = tile / NUM_HEADS;
      const int head = tile - m_blk * NUM_HEADS;
      int row_a, rows_a;
      cta_tile_rows(p, m_blk * 2 + static_cast<int>(cta_rank), row_a, rows_a);
      const int row_b = head * HN + static_cast<int>(cta_rank) * (HN / 2);
      for (int it = 0; it < NK; ++it) {
        mbar_wait(empty_bar + 8 * s, ph ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(full_bar + 8 * s, 2 * STAGE_BYTES);
          tma_load_2d_cg2(sA + s * A_BYTES, &tmA, full_leader + 8 * s, it * BK, row_a);
          tma_load_2d_cg2(sB + s * B_BYTES, &tmB, full_leader + 8 * s, it * BK, row_b);
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * BM, HN);
      int s = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar + 8 * a, aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + a * ACC_STRIDE;
        for (int it = 0; it < NK; ++it) {
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_addr = sA + s * A_BYTES;
          const uint32_t b_addr = sB + s * B_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16<2>(tmem_d, make_kmajor_sw128_desc(a_addr + k * 32), make_kmajor_sw128_desc(b_addr + k * 32),
                           idesc, (it | k) != 0 ? 1u : 0u);
            umma_commit_cg2(empty_bar + 8 * s, 3);
            if (it == NK - 1) umma_commit_cg2(tfull_bar + 8 * a, 3);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++a == ACC_STAGES) { a = 0; aph ^= 1; }
      }
    }
  }
  } else if (warp < FIRST_ATT_WARP) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;" ::: "memory");
    // ---- drain: accumulators -> folded LayerNorm -> bf16 q | k | v of this head in shared memory ----
    const int e = warp - FIRST_EPI_WARP;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int half = e >> 2;  // columns [96 half, 96 half + 96) of the tile
    const int lrow = q * 32 + lane;
    const int sw = lrow & 7;
    int a = 0;
    uint32_t aph = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = tile / NUM_HEADS;
      const int head = tile - m_blk * NUM_HEADS;
      int tile_row0, tile_rows;
      cta_tile_rows(p, m_blk * 2 + static_cast<int>(cta_rank), tile_row0, tile_rows);
      const long long grow = static_cast<long long>(tile_row0) + lrow;
      float ln_rstd = 0.f, ln_nmr = 0.f;
      if (grow < p.M) {
        const float4* sp = reinterpret_cast<const float4*>(p.ln_stats) + grow * 3;
        // 6 x (sum, sumsq) per 128 columns; streamed once per tile: kept out of the (28 KiB) L1 so that the bias / column-sum
        // vectors of the 12 heads (18 KiB, re-read by every tile) stay resident
        const float4 a0 = ldg_f4_stream(sp), a1 = ldg_f4_stream(sp + 1), a2 = ldg_f4_stream(sp + 2);
        const float mean = (((a0.x + a0.z) + (a1.x + a1.z)) + (a2.x + a2.z)) * (1.0f / 768.0f);
        const float var = (((a0.y + a0.w) + (a1.y + a1.w)) + (a2.y + a2.w)) * (1.0f / 768.0f) - mean * mean;
        ln_rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
        ln_nmr = -ln_rstd * mean;
      }
      const uint64_t rstd2 = pk2(ln_rstd, ln_rstd), nmr2 = pk2(ln_nmr, ln_nmr);
      mbar_wait(tfull_bar + 8 * a, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * ACC_STRIDE + half * 96;
      const uint32_t srow = sQKV + a * QKV_BYTES + lrow * QKV_PITCH;
      uint32_t v[32];
      tmem_ld_32x32(trow, v);
      mbar_wait(sempty_bar + 8 * a, aph ^ 1);  // the attention warps are done with this buffer's previous tile
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        tmem_ld_wait();
        uint64_t f2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f2[j] = pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        if (c < 2) tmem_ld_32x32(trow + (c + 1) * 32, v);
        else {  // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar + 8 * a);
            else mbar_arrive_cluster(tempty_bar + 8 * a, 0);
          }
        }
        const int col = half * 96 + c * 32;  // first of these 32 tile columns
        const int gcol = head * HN + col;
        uint32_t w[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
#ifdef DVST_TATTN_EXP_NO_LDG
          const float4 b4 = make_float4(0.1f, 0.2f, 0.3f, 0.4f), s4 = make_float4(1.f, 2.f, 3.f, 4.f);
#else
          const float4 b4 = ldg_f4_keep(reinterpret_cast<const float4*>(p.bias + gcol) + j4);
          const float4 s4 = ldg_f4_keep(reinterpret_cast<const float4*>(p.ln_colsum + gcol) + j4);
#endif
          float fa, fb;
          unpk2(fma2(rstd2, f2[2 * j4], fma2(nmr2, pk2(s4.x, s4.y), pk2(b4.x, b4.y))), fa, fb);
          w[2 * j4] = pack2(fa, fb);
          unpk2(fma2(rstd2, f2[2 * j4 + 1], fma2(nmr2, pk2(s4.z, s4.w), pk2(b4.z, b4.w))), fa, fb);
          w[2 * j4 + 1] = pack2(fa, fb);
        }
        const uint32_t seg = srow + (col >> 6) * 128;  // q, k or v segment of the row
        const int ch0 = (col & 63) >> 3;               // first 16-byte chunk inside the segment (0 or 4)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(seg + (((ch0 + j) ^ sw) << 4)), "r"(w[4 * j]),
                       "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3]) : "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sfull_bar + 8 * a);
      if (++a == ACC_STAGES) { a = 0; aph ^= 1; }
    }
  } else {
    // ---- temporal attention of query rows [16 w, 16 w + 16) of the CTA tile, one head per tile ----
    const int wq = warp - FIRST_ATT_WARP;
    const int g = lane >> 2, tig = lane & 3;
    const int mat = lane >> 3, r8 = lane & 7;
    const int T = p.T, R = p.R;
    // geometry of this warp's 16-row query tile (the same for every tile of the launch)
    const int r0 = 16 * wq;
    const int n_valid = R - r0 < 0 ? 0 : (R - r0 > 16 ? 16 : R - r0);
    int k0 = 0;
    uint32_t m0 = 0, m1 = 0;  // bit (2 nt + j): key k0 + 8 nt + 2 tig + j belongs to the sequence of row r0 + g / + 8
    constexpr uint32_t ALL_KEYS = MAXP == 4 ? 0xFFFFu : ((1u << (4 * MAXP)) - 1u);
    if (n_valid > 0) {
      k0 = (r0 / T) * T;
      if (k0 > BM - 16 * MAXP) k0 = BM - 16 * MAXP;
      const int ra = r0 + g < R ? r0 + g : r0, rb = r0 + g + 8 < R ? r0 + g + 8 : r0;
      const int sa = ra / T, sb = rb / T;
#pragma unroll
      for (int nt = 0; nt < 2 * MAXP; ++nt) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = k0 + nt * 8 + tig * 2 + j;
          const int ks = k < R ? k / T : -1;
          if (ks == sa) m0 |= 1u << (2 * nt + j);
          if (ks == sb) m1 |= 1u << (2 * nt + j);
        }
      }
    }
    const uint32_t obox = sO + wq * OBOX_BYTES;
    const float scale_log2 = 0.125f * 1.4426950408889634f;
    int a = 0;
    uint32_t aph = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = tile / NUM_HEADS;
      const int head = tile - m_blk * NUM_HEADS;
      int tile_row0, tile_rows;
      cta_tile_rows(p, m_blk * 2 + static_cast<int>(cta_rank), tile_row0, tile_rows);
      const int nv = tile_rows - r0 < n_valid ? tile_rows - r0 : n_valid;  // valid query rows of this warp in this tile
      const uint32_t sT = sQKV + a * QKV_BYTES;
      mbar_wait(sfull_bar + 8 * a, aph);
#ifndef DVST_TATTN_EXP_NO_ATTENTION
      if (nv > 0) {
        uint32_t qa[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int row = r0 + (mat & 1) * 8 + r8;
          const int ch = kk * 2 + (mat >> 1);
          ldsm_x4(sT + row * QKV_PITCH + ((ch ^ (row & 7)) << 4), qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3]);
        }
        constexpr int NT = 2 * MAXP;
        float s[NT][4];
#pragma unroll
        for (int i = 0; i < NT; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
          for (int np = 0; np < MAXP; ++np) {
            const int row = k0 + np * 16 + (mat >> 1) * 8 + r8;
            const int ch = kk * 2 + (mat & 1);
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sT + 128 + row * QKV_PITCH + ((ch ^ (row & 7)) << 4), b0, b1, b2, b3);
            mma_bf16(s[2 * np], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b0, b1);
            mma_bf16(s[2 * np + 1], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b2, b3);
          }
        }
        if ((m0 & m1) != ALL_KEYS) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            if (!((m0 >> (2 * nt)) & 1u)) s[nt][0] = -INFINITY;
            if (!((m0 >> (2 * nt + 1)) & 1u)) s[nt][1] = -INFINITY;
            if (!((m1 >> (2 * nt)) & 1u)) s[nt][2] = -INFINITY;
            if (!((m1 >> (2 * nt + 1)) & 1u)) s[nt][3] = -INFINITY;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float off0 = mx0 * scale_log2, off1 = mx1 * scale_log2;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          s[nt][0] = ex2_approx(fmaf(s[nt][0], scale_log2, -off0));
          s[nt][1] = ex2_approx(fmaf(s[nt][1], scale_log2, -off0));
          s[nt][2] = ex2_approx(fmaf(s[nt][2], scale_log2, -off1));
          s[nt][3] = ex2_approx(fmaf(s[nt][3], scale_log2, -off1));
          l0 += s[nt][0] + s[nt][1];
          l1 += s[nt][2] + s[nt][3];
        }
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < MAXP; ++kk) {
          const uint32_t a0 = pack2(s[2 * kk][0], s[2 * kk][1]), a1 = pack2(s[2 * kk][2], s[2 * kk][3]);
          const uint32_t a2 = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
          const int row = k0 + kk * 16 + (mat & 1) * 8 + r8;
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            const int ch = dp * 2 + (mat >> 1);
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(sT + 256 + row * QKV_PITCH + ((ch ^ (row & 7)) << 4), b0, b1, b2, b3);
            mma_bf16(o[2 * dp], a0, a1, a2, a3, b0, b1);
            mma_bf16(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
          }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        // the TMA store of this box's previous tile must have finished reading it
        if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int lr = g + hh * 8;
          const float inv = hh == 0 ? inv0 : inv1;
          const uint32_t orow = obox + lr * 128 + tig * 4;
#pragma unroll
          for (int dt = 0; dt < 8; ++dt) {
            const uint32_t hv = pack2(o[dt][2 * hh] * inv, o[dt][2 * hh + 1] * inv);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(orow + ((dt ^ (lr & 7)) << 4)), "r"(hv) : "memory");
          }
        }
        if (nv == 16) {
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(obox), "r"(head * 64), "r"(tile_row0 + r0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          // the rows past R belong to the next tile: plain 16-byte stores of the valid rows only
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i;
            const int lr = idx >> 3, ch = idx & 7;
            const long long gr = static_cast<long long>(tile_row0) + r0 + lr;
            if (lr < nv && gr < p.M) {
              uint4 val;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                           : "r"(obox + lr * 128 + ((ch ^ (lr & 7)) << 4)) : "memory");
              *reinterpret_cast<uint4*>(p.out + gr * p.ldo + head * 64 + ch * 8) = val;
            }
          }
        }
      }
#endif
      __syncwarp();  // every lane has read its q / k / v fragments of this buffer
      if (lane == 0) mbar_arrive(sempty_bar + 8 * a);
      if (++a == ACC_STAGES) { a = 0; aph ^= 1; }
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, TMEM_COLS);
  }
}

}  // namespace

int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols, bool swizzle64);

static int g_tattn_fusion = 1;
void set_temporal_fusion(int on) { g_tattn_fusion = on != 0 ? 1 : 0; }
int temporal_fusion() { return g_tattn_fusion; }

// Tile geometry for sequences of T rows: rows per CTA tile and the key-block size (in 16-key units: 1, 2 or 4) that covers
// the sequences any 16-row query tile touches.  Returns false when T is not covered (T > 64).
static bool tattn_geometry(int T, int* R_out, int* maxp_out) {
  if (T < 1 || T > 64) return false;
  // as many whole sequences as fit 128 rows, unless a 16-row query tile could then touch more than 64 rows of sequences
  // (33 <= T <= 63): one sequence per tile there
  for (int R = (BM / T) * T; R >= T; R = (R == T ? 0 : T)) {
    int span = 0;
    for (int r0 = 0; r0 < R; r0 += 16) {
      const int key0 = (r0 / T) * T;
      const int last = r0 + 15 < R ? r0 + 15 : R - 1;
      const int kend = (last / T + 1) * T;
      if (kend - key0 > span) span = kend - key0;
    }
    if (span <= 64) {
      *R_out = R;
      *maxp_out = span <= 16 ? 1 : (span <= 32 ? 2 : 4);
      return true;
    }
  }
  return false;
}

// exported for the host-logic tests (no GPU needed): geometry of the fused kernel for sequences of T rows
int temporal_tile_geometry(int T, int* rows_per_tile, int* key_block) {
  int R = 0, maxp = 0;
  if (!tattn_geometry(T, &R, &maxp)) return -1;
  *rows_per_tile = R;
  *key_block = 16 * maxp;
  return 0;
}

template <int MAXP>
static int launch_tattn_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const TAttnParams& p,
                             double flops, cudaStream_t stream) {
  static PerDevice<bool> attr_cache;
  bool& attr_set = attr_cache.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_tattn_kernel<MAXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_error("gemm_ln_tattn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * 12;
  const int max_units = num_sms() / 2;
  const int units = tiles < max_units ? tiles : max_units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * 2);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  ProfScope prof(PROF_GEMM_TATTN, flops, stream);
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_ln_tattn_kernel<MAXP>, tmA, tmB, tmO, p);
  if (e != cudaSuccess) return set_error("gemm_ln_tattn launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

// 0 = ran, -1 = shape not covered (the caller runs the qkv GEMM and the temporal attention kernel), DVST_ERR = error
int launch_gemm_ln_tattn(const void* xb, long long ldxb, const void* w_perm, const float* bias_perm,
                         const float* colsum_perm, const float* ln_stats, float ln_eps, long long M, int T,
                         long long clip_rows, void* out, long long ldo, cudaStream_t stream) {
  int R = 0, maxp = 0;
  if (!tattn_geometry(T, &R, &maxp)) return -1;
  if (M <= 0) return 0;
  if (M % T != 0) return set_error("gemm_ln_tattn: %lld rows are not whole sequences of %d", M, T);
  if (clip_rows <= 0) clip_rows = M;
  if (clip_rows % T != 0 || M % clip_rows != 0)
    return set_error("gemm_ln_tattn: clip_rows = %lld must be a multiple of T = %d and divide M = %lld", clip_rows, T, M);
  if (M > 2000000000LL) return set_error("gemm_ln_tattn: too many rows");
  if ((reinterpret_cast<uintptr_t>(xb) & 15) || (reinterpret_cast<uintptr_t>(w_perm) & 15) || (ldxb % 8) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (ldo % 8) || (reinterpret_cast<uintptr_t>(ln_stats) & 15) ||
      (reinterpret_cast<uintptr_t>(bias_perm) & 15) || (reinterpret_cast<uintptr_t>(colsum_perm) & 15))
    return set_error("gemm_ln_tattn: operands must be 16-byte aligned with leading dimensions multiple of 8");
  CUtensorMap tmA, tmB, tmO;
  if (int rc = make_tmap_2d(&tmA, xb, 2, static_cast<uint64_t>(M), KD, ldxb, BM, BK, false)) return rc;
  if (int rc = make_tmap_2d(&tmB, w_perm, 2, 3 * KD, KD, KD, HN / 2, BK, false)) return rc;
  if (int rc = make_tmap_2d(&tmO, out, 2, static_cast<uint64_t>(M), KD, ldo, 16, 64, false)) return rc;
  TAttnParams p;
  p.M = static_cast<int>(M); p.T = T; p.R = R;
  p.clip_rows = static_cast<int>(clip_rows);
  p.tiles_per_clip = static_cast<int>((clip_rows + R - 1) / R);
  const long long cta_tiles = (M / clip_rows) * p.tiles_per_clip;
  p.num_m_blocks = static_cast<int>((cta_tiles + 1) / 2);
  p.bias = bias_perm; p.ln_colsum = colsum_perm; p.ln_stats = ln_stats; p.ln_eps = ln_eps;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  // executed work: the qkv GEMM as launched plus 4 T^2 64 FLOPs per (sequence, head)
  const double flops = 2.0 * M * (3.0 * KD) * KD + 4.0 * (static_cast<double>(M) / T) * 12.0 * T * T * 64.0;
  if (maxp == 1) return launch_tattn_impl<1>(tmA, tmB, tmO, p, flops, stream);
  if (maxp == 2) return launch_tattn_impl<2>(tmA, tmB, tmO, p, flops, stream);
  return launch_tattn_impl<4>(tmA, tmB, tmO, p, flops, stream);
}

}  // namespace dvst
