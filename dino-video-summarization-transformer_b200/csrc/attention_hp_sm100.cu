// High-precision spatial attention on tcgen05 / TMEM for precision = "high" (the 2-term bf16 split that reference-exact
// mode-R scoring needs, dino_similarity.py:59-63 with tau_t = 0.02): softmax(q k^T / 8) v over the 1 + N tokens of one frame
// (models/timesformer.py:140-164, :120-123) with every operand a bf16 [hi | lo] pair and every product evaluated as
// hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator (~16 mantissa bits end to end):
//
//   S = Qhi Khi^T + Qhi Klo^T + Qlo Khi^T                     12 tcgen05.mma (M = 128, N = NPAD, K = 16)
//   p = exp2(s * scale - ref)  (fp32),  Phi = bf16(p),  Plo = bf16(p - Phi)   both written to TMEM
//   O = Phi Vhi + Phi Vlo + Plo Vhi                           3 * NPAD / 16 tcgen05.mma (A from TMEM, V MN-major from smem)
//   out = O / sum(p)  ->  bf16 [hi | lo] pair, two TMA stores
//
// Same arrangement as the split kernel (attention_split_sm100.cu): eight softmax warps on one 128-row tile, the two
// threads of a query row take one half of the keys each and share the reference point (one-pass softmax; rows that overflow
// go to a hi/lo-aware fp32 fix-up kernel), dedicated epilogue warps, elect.sync issue.  Differences: P hi and P lo need
// 2 x NPAD / 2 TMEM columns, so ONE unit is in flight per SM (S [0, 208) | Phi second half [208, 256) | Plo [256, 376) |
// O [384, 512)), and Q / K / V hi + lo fill the shared memory with a single stage (Q | K of the next problem are requested as
// soon as the S MMAs of the current one have read them, V after its P V MMAs).  Replaces the mma.sync split-precision kernel
// (attention.cu attention_kernel<4, true>) for 32 <= NPAD <= 208.
#include <type_traits>

#include "attn_sm100_common.cuh"

namespace dvst {

using namespace attn;

namespace {

constexpr int HP_THREADS = 512;
constexpr int HP_P1 = 208;     // Phi of the second key half
constexpr int HP_PLO0 = 256;   // Plo of the first key half
constexpr int HP_PLO1 = 320;   // Plo of the second key half
constexpr int HP_O = 384;      // O accumulator of the even key steps
constexpr int HP_O2 = 448;     // ... of the odd key steps

struct HpParams {
  SpParams sp;
  long long lo_in;    // element offset of the lo halves inside a qkv row (3 * 768)
  long long lo_out;   // ... inside an output row (768)
  int K0;             // keys of the first half
};

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
// bf16 pair -> (hi pair, lo pair) with lo = bf16(x - float(hi))
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
  lo = pack_bf16x2(a - ha, b - hb);
}

__global__ void __launch_bounds__(HP_THREADS, 1)
spatial_attention_hp_sm100_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOutA,
                                  const __grid_constant__ CUtensorMap tmOutB, const HpParams hp) {
  extern __shared__ uint8_t smem_raw[];
  const SpParams& p = hp.sp;
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int S = p.S, N = p.N, NPAD = p.NPAD, K0 = hp.K0;
  const int n_tiles = S > 128 ? 2 : 1;
  const uint32_t q_bytes = n_tiles * TILE_BYTES;
  const uint32_t kv_bytes = (static_cast<uint32_t>(NPAD) * 128u + 1023u) & ~1023u;
  // shared memory: Qhi | Qlo | Khi | Klo | Vhi | Vlo | O staging hi | O staging lo | row sums | barriers
  const uint32_t oQ = 0, oK = 2 * q_bytes, oV = oK + 2 * kv_bytes, oO = oV + 2 * kv_bytes, oSums = oO + 2 * TILE_BYTES;
  const uint32_t bars = base + oSums + 1024;
  // s_full / p_full / o_full / o_empty exist once per row-tile type: tile B has fewer active lane quarters (different arrival
  // counts), and a warp that sits out the tile-B units must still see EVERY phase of the barriers it waits on - a parity wait
  // cannot tell phase k from phase k + 2, so a shared barrier would let it run a unit ahead of the MMAs.
  const uint32_t qk_full = bars, qk_empty = bars + 8, v_full = bars + 16, v_empty = bars + 24,
                 s_full = bars + 32 /* [2] */, o_full = bars + 48 /* [2] */, p_full = bars + 64 /* [2] */,
                 o_empty = bars + 80 /* [2] */, tmem_slot = bars + 96;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));
  float* sums = reinterpret_cast<float*>(smem + oSums);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int act_a = min(4, (min(S, 128) + 31) >> 5);
  const int act_b = n_tiles == 2 ? (S - 128 + 31) >> 5 : 0;
  const int n_my = static_cast<int>(blockIdx.x) < p.n_prob
                       ? (p.n_prob - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  const int n_units = n_my * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmOutA);
    tma_prefetch_desc(&tmOutB);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_full + 8, 1);
    mbar_init(o_full, 1);
    mbar_init(o_full + 8, 1);
    mbar_init(p_full, 2 * act_a);
    mbar_init(p_full + 8, act_b > 0 ? 2 * act_b : 1);
    mbar_init(o_empty, act_a);
    mbar_init(o_empty + 8, act_b > 0 ? act_b : 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  // K / V rows [S, NPAD) (hi and lo) are never written by the loads: zero them once
  {
    const int pad_chunks = (NPAD - S) * 8;
    for (int i = threadIdx.x; i < 4 * pad_chunks; i += HP_THREADS) {
      const int which = i / pad_chunks, j = i - which * pad_chunks;   // Khi, Klo, Vhi, Vlo
      uint8_t* dst = smem + oK + which * kv_bytes + (S * 128 + j * 16);
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== producer =====
    for (int it = 0; it < n_my; ++it) {
      const int prob = blockIdx.x + it * gridDim.x;
      const int seq = prob / p.heads, head = prob - seq * p.heads;
      const int b = seq / p.T, t = seq - b * p.T;
      const long long cls_row = p.cls_base + (p.mode == 1 ? b : seq);
      const int lo = static_cast<int>(hp.lo_in);
      if (it > 0) mbar_wait(qk_empty, (it - 1) & 1);
      if (elect_one()) {
        tma_load_3d(base + oQ, &tmIn, qk_full, head * HD, t, b * N);
        tma_load_3d(base + oQ + q_bytes, &tmIn, qk_full, lo + head * HD, t, b * N);
        tma_load_3d(base + oK, &tmIn, qk_full, DMODEL + head * HD, t, b * N);
        tma_load_3d(base + oK + kv_bytes, &tmIn, qk_full, lo + DMODEL + head * HD, t, b * N);
      }
      {
        const int which = lane >> 3, ch = lane & 7;   // Qhi, Qlo, Khi, Klo
        const long long col = (which & 1 ? hp.lo_in : 0) + (which >> 1) * DMODEL + head * HD;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + cls_row * p.ld + col) + ch);
        const uint32_t off = (which < 2 ? oQ + (which & 1) * q_bytes : oK + (which & 1) * kv_bytes) + N * 128 +
                             ((ch ^ (N & 7)) << 4);
        *reinterpret_cast<uint4*>(smem + off) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) mbar_arrive_expect_tx(qk_full, 4u * N * 128u);
      if (it > 0) mbar_wait(v_empty, (it - 1) & 1);
      if (elect_one()) {
        tma_load_3d(base + oV, &tmIn, v_full, 2 * DMODEL + head * HD, t, b * N);
        tma_load_3d(base + oV + kv_bytes, &tmIn, v_full, lo + 2 * DMODEL + head * HD, t, b * N);
      }
      if (lane < 16) {
        const int which = lane >> 3, ch = lane & 7;   // Vhi, Vlo
        const long long col = (which ? hp.lo_in : 0) + 2 * DMODEL + head * HD;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + cls_row * p.ld + col) + ch);
        *reinterpret_cast<uint4*>(smem + oV + which * kv_bytes + N * 128 + ((ch ^ (N & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) mbar_arrive_expect_tx(v_full, 2u * N * 128u);
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp in the loop, one elected lane issues) =====
    const uint32_t idesc_s = make_idesc(128, NPAD, 0);
    constexpr uint32_t idesc_pv = make_idesc(128, HD, 1);
    const int ksteps = NPAD >> 4;
    for (int u = 0; u < n_units; ++u) {
      const int it = n_tiles == 2 ? (u >> 1) : u, tile = n_tiles == 2 ? (u & 1) : 0;
      if (tile == 0) {
        mbar_wait(qk_full, it & 1);
        tc_fence_after();
      }
      if (u > 0) {
        // O of the previous unit has left TMEM (its tile type and use count select the barrier and the parity)
        const int pu = u - 1, pt = n_tiles == 2 ? (pu & 1) : 0;
        mbar_wait(o_empty + 8 * pt, n_tiles == 2 ? (pu >> 1) & 1 : pu & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        // S = Qhi Khi^T + Qhi Klo^T + Qlo Khi^T
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t qa = base + oQ + (term == 2 ? q_bytes : 0) + tile * TILE_BYTES;
          const uint32_t kb = base + oK + (term == 1 ? kv_bytes : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16<1>(tmem_base, make_kmajor_sw128_desc(qa + k * 32), make_kmajor_sw128_desc(kb + k * 32), idesc_s,
                         (term | k) != 0 ? 1u : 0u);
        }
        umma_commit(s_full + 8 * tile);
        if (tile == n_tiles - 1) umma_commit(qk_empty);
      }
      __syncwarp();
      mbar_wait(p_full + 8 * tile, n_tiles == 2 ? (u >> 1) & 1 : u & 1);
      tc_fence_after();
      if (tile == 0) {
        mbar_wait(v_full, it & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        // O = Phi Vhi + Phi Vlo + Plo Vhi, even / odd key steps into two accumulators
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint64_t vdesc0 = make_mnmajor_sw128_desc(base + oV + (term == 1 ? kv_bytes : 0));
          const uint32_t a_lo = tmem_base + (term == 2 ? HP_PLO0 : 0);
          const uint32_t a_hi = tmem_base + (term == 2 ? HP_PLO1 : HP_P1) - static_cast<uint32_t>(K0 >> 1);
#pragma unroll
          for (int j = 0; j < 13; ++j) {
            if (j < ksteps) {
              const uint32_t a_addr = ((j << 4) < K0 ? a_lo : a_hi) + j * 8;
              umma_ts_bf16(tmem_base + ((j & 1) ? HP_O2 : HP_O), a_addr, vdesc0 + static_cast<uint64_t>(j * 128), idesc_pv,
                           (term != 0 || j >= 2) ? 1u : 0u);
            }
          }
        }
        umma_commit(o_full + 8 * tile);
        if (tile == n_tiles - 1) umma_commit(v_empty);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 12) {
    // ===== softmax: two threads per query row, one key half each; P written as a bf16 hi / lo pair =====
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    const float scale_log2 = 0.125f * 1.4426950408889634f;
    const int c_begin = h ? K0 : 0;
    const int len = h ? NPAD - K0 : K0;
    const int nfull = len >> 5;
    const bool tail = (len & 16) != 0;
    const int ref_n = min(32, S - K0);
    for (int u = 0; u < n_units; ++u) {
      const int tile = n_tiles == 2 ? (u & 1) : 0;
      const int n_act = tile == 0 ? act_a : act_b;
      if (q >= n_act) continue;
      const uint32_t treg = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t pdst_hi = treg + (h ? HP_P1 : 0), pdst_lo = treg + (h ? HP_PLO1 : HP_PLO0);
      mbar_wait(s_full + 8 * tile, n_tiles == 2 ? (u >> 1) & 1 : u & 1);
      tc_fence_after();
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
        float m0 = __uint_as_float(sc), m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ref_n) { if (j & 1) m1 = fmaxf(m1, __uint_as_float(va[j])); else m0 = fmaxf(m0, __uint_as_float(va[j])); }
        mx = fmaxf(m0, m1);
      }
      const float moff = mx * scale_log2;
      const uint64_t scale2 = pk2(scale_log2, scale_log2), nmoff2 = pk2(-moff, -moff);
      uint64_t sum2a = 0ull, sum2b = 0ull;
      auto exp_chunk = [&](const uint32_t (&cur)[32], int col, int ncols, int pc) {
        uint32_t ph[16], pl[16];
        const bool masked = col + ncols > S;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (2 * j < ncols) {
            float a0, a1;
            unpk2(fma2(pk2(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), scale2, nmoff2), a0, a1);
            float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
            if (masked) {
              if (col + 2 * j >= S) e0 = 0.f;
              if (col + 2 * j + 1 >= S) e1 = 0.f;
            }
            if (j & 1) sum2b = add2(sum2b, pk2(e0, e1)); else sum2a = add2(sum2a, pk2(e0, e1));
            split_pack(e0, e1, ph[j], pl[j]);
          }
        }
        if (ncols == 32) { tmem_st_32x16(pdst_hi + pc, ph); tmem_st_32x16(pdst_lo + pc, pl); }
        else { tmem_st_32x8(pdst_hi + pc, ph); tmem_st_32x8(pdst_lo + pc, pl); }
      };
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nfull) {
          uint32_t (&cur)[32] = (c & 1) ? second : first;
          uint32_t (&nxt)[32] = (c & 1) ? first : second;
          if (c > 0) tmem_ld_wait();
          if (!(h == 1 && c == 0)) {
            if (c + 1 < nfull) tmem_ld_32x32(treg + c_begin + (c + 1) * 32, nxt);
            else if (tail) tmem_ld_32x16(treg + c_begin + (c + 1) * 32, nxt);
          }
          exp_chunk(cur, c_begin + c * 32, 32, c * 16);
        }
      }
      if (tail) {
        if (nfull > 0) tmem_ld_wait();
        if (nfull & 1) exp_chunk(second, c_begin + nfull * 32, 16, nfull * 16);
        else exp_chunk(first, c_begin + nfull * 32, 16, nfull * 16);
      }
      float sa, sb;
      unpk2(add2(sum2a, sum2b), sa, sb);
      sums[h * 128 + q * 32 + lane] = sa + sb;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + 8 * tile);
    }
  } else if (warp >= 12) {
    // ===== epilogue: O out of TMEM, 1 / row sum, bf16 hi / lo pair, staging tiles, two TMA stores =====
    const int q = warp & 3;
    const bool store_leader = q == 0 && lane == 0;
    const uint32_t so_hi = base + oO, so_lo = base + oO + TILE_BYTES;
    for (int u = 0; u < n_units; ++u) {
      const int it = n_tiles == 2 ? (u >> 1) : u, tile = n_tiles == 2 ? (u & 1) : 0;
      const int n_act = tile == 0 ? act_a : act_b;
      if (q >= n_act) continue;
      const int prob = blockIdx.x + it * gridDim.x;
      const int seq = prob / p.heads, head = prob - seq * p.heads;
      const int b = seq / p.T, t = seq - b * p.T;
      const int rloc = q * 32 + lane;
      const int row = tile * 128 + rloc;
      const int rows_tile = tile == 0 ? p.rows_a : p.rows_b;
      const uint32_t treg = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      mbar_wait(p_full + 8 * tile, n_tiles == 2 ? (u >> 1) & 1 : u & 1);   // the row sums are visible
      mbar_wait(o_full + 8 * tile, n_tiles == 2 ? (u >> 1) & 1 : u & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      {
        uint32_t t0[32];
        tmem_ld_32x32(treg + HP_O, o0);
        tmem_ld_32x32(treg + HP_O2, t0);
        tmem_ld_wait();
        tmem_ld_32x32(treg + HP_O + 32, o1);
#pragma unroll
        for (int j = 0; j < 32; ++j) o0[j] = __float_as_uint(__uint_as_float(o0[j]) + __uint_as_float(t0[j]));
        tmem_ld_32x32(treg + HP_O2 + 32, t0);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o1[j] = __float_as_uint(__uint_as_float(o1[j]) + __uint_as_float(t0[j]));
      }
      const float sum = sums[rloc] + sums[128 + rloc];
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty + 8 * tile);
      const float inv = 1.0f / sum;
      if (row < S && !(sum < 1.8e19f)) p.fix_list[atomicAdd(p.fix_count, 1)] = make_int2(prob, row);
      if (store_leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      named_bar_sync(1 + tile, n_act * 32);
      if (row < N) {
        const uint32_t dh = so_hi + rloc * 128, dl = so_lo + rloc * 128;
        const int sw = rloc & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t (&o)[32] = j < 4 ? o0 : o1;
          const int j8 = (j & 3) * 8;
          uint32_t wh[4], wl[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            split_pack(__uint_as_float(o[j8 + 2 * k]) * inv, __uint_as_float(o[j8 + 2 * k + 1]) * inv, wh[k], wl[k]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dh + ((j ^ sw) << 4)), "r"(wh[0]), "r"(wh[1]),
                       "r"(wh[2]), "r"(wh[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dl + ((j ^ sw) << 4)), "r"(wl[0]), "r"(wl[1]),
                       "r"(wl[2]), "r"(wl[3]) : "memory");
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
            uint4 wh, wl;
            split_pack(__uint_as_float(o[j8]) * inv, __uint_as_float(o[j8 + 1]) * inv, wh.x, wl.x);
            split_pack(__uint_as_float(o[j8 + 2]) * inv, __uint_as_float(o[j8 + 3]) * inv, wh.y, wl.y);
            split_pack(__uint_as_float(o[j8 + 4]) * inv, __uint_as_float(o[j8 + 5]) * inv, wh.z, wl.z);
            split_pack(__uint_as_float(o[j8 + 6]) * inv, __uint_as_float(o[j8 + 7]) * inv, wh.w, wl.w);
            *reinterpret_cast<uint4*>(dst + j * 8) = wh;
            *reinterpret_cast<uint4*>(dst + hp.lo_out + j * 8) = wl;
          }
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + tile, n_act * 32);
      if (store_leader && rows_tile > 0) {
        const CUtensorMap* tm = tile == 0 ? &tmOutA : &tmOutB;
        tma_store_3d(tm, so_hi, head * HD, t, b * N + tile * 128);
        tma_store_3d(tm, so_lo, static_cast<int>(hp.lo_out) + head * HD, t, b * N + tile * 128);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// Exact fp32 recomputation (hi + lo operands) of the rows the one-pass softmax could not normalise: one warp per row.
__global__ void __launch_bounds__(128) attention_fixup_hp_kernel(const HpParams hp) {
  const SpParams& p = hp.sp;
  const int n_fix = *p.fix_count;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  auto ld2 = [&](const __nv_bfloat16* ptr, int d) {
    return __bfloat162float(ptr[d]) + __bfloat162float(ptr[hp.lo_in + d]);
  };
  for (int i = warp; i < n_fix; i += n_warps) {
    const int2 e = p.fix_list[i];
    const int prob = e.x, row = e.y;
    const int seq = prob / p.heads, head = prob - seq * p.heads;
    const int b = seq / p.T, t = seq - b * p.T;
    const long long cls_row = p.cls_base + (p.mode == 1 ? b : seq);
    auto token_row = [&](int j) { return j < p.N ? (static_cast<long long>(b) * p.N + j) * p.T + t : cls_row; };
    const __nv_bfloat16* qp = p.qkv + token_row(row) * p.ld + head * HD;
    float qv[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) qv[d] = ld2(qp, d);
    float sc[8];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = lane + 32 * u;
      sc[u] = -INFINITY;
      if (j < p.S) {
        const __nv_bfloat16* kp = p.qkv + token_row(j) * p.ld + DMODEL + head * HD;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc = fmaf(qv[d], ld2(kp, d), acc);
        sc[u] = acc * 0.125f;
        mx = fmaxf(mx, sc[u]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    float sum = 0.f;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = lane + 32 * u;
      if (j < p.S) {
        const float w = expf(sc[u] - mx);
        sum += w;
        const __nv_bfloat16* vp = p.qkv + token_row(j) * p.ld + 2 * DMODEL + head * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = fmaf(w, ld2(vp, d), acc[d]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      float v = acc[d];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
      if ((d & 31) == lane) { if (d < 32) o0 = v; else o1 = v; }
    }
    o0 *= inv;
    o1 *= inv;
    auto put = [&](__nv_bfloat16* dst, int d, float v) {
      const __nv_bfloat16 hi = __float2bfloat16(v);
      dst[d] = hi;
      dst[hp.lo_out + d] = __float2bfloat16(v - __bfloat162float(hi));
    };
    if (row < p.N) {
      __nv_bfloat16* dst = p.out + token_row(row) * p.ldo + head * HD;
      put(dst, lane, o0);
      put(dst, lane + 32, o1);
    } else if (p.mode == 1) {
      float* dst = p.cls_out + static_cast<long long>(seq) * DMODEL + head * HD;
      dst[lane] = o0;
      dst[lane + 32] = o1;
    } else {
      __nv_bfloat16* dst = p.out + (p.cls_base + seq) * p.ldo + head * HD;
      put(dst, lane, o0);
      put(dst, lane + 32, o1);
    }
  }
}

}  // namespace

// Split-precision spatial attention (modes 1 / 3) on tcgen05.  Returns 0 when launched, 1 when the shape is not covered
// (the caller then uses the mma.sync split kernels), anything else is an error.  qkv rows: [q k v hi (2304) | q k v lo (2304)],
// out rows: [hi (768) | lo (768)].
int launch_spatial_attention_hp_sm100(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode,
                                      int B, int T, int N, long long cls_base, cudaStream_t st, void* scratch,
                                      long long scratch_bytes) {
  const int S = N + 1;
  const int NPAD = (S + 15) & ~15;
  if ((mode != 1 && mode != 3) || NPAD < 32 || NPAD > 208 || N < 1) return 1;
  if (scratch == nullptr || scratch_bytes < spatial_attention_scratch_bytes(B, T, N) ||
      (reinterpret_cast<uintptr_t>(scratch) & 15))
    return 1;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (ld % 8) || (ldo % 8) ||
      ld < 6 * DMODEL || ldo < 2 * DMODEL)
    return 1;
  HpParams hp;
  SpParams& p = hp.sp;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.ld = ld;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.cls_out = cls_out;
  p.mode = mode;
  p.S = S; p.N = N; p.NPAD = NPAD; p.T = T;
  p.heads = 12;
  p.n_prob = B * T * p.heads;
  p.cls_base = cls_base;
  p.rows_a = N < 128 ? N : 128;
  p.rows_b = N > 128 ? N - 128 : 0;
  p.fix_count = static_cast<int*>(scratch);
  p.fix_list = reinterpret_cast<int2*>(static_cast<uint8_t*>(scratch) + 256);
  hp.lo_in = 3 * DMODEL;
  hp.lo_out = DMODEL;
  hp.K0 = ((NPAD / 2) + 15) & ~15;
  if (hp.K0 >= S) return 1;
  if (p.n_prob <= 0) return 0;
  const int n_tiles = S > 128 ? 2 : 1;
  const size_t kv_bytes = (static_cast<size_t>(NPAD) * 128 + 1023) & ~static_cast<size_t>(1023);
  const size_t smem = 1024 + 2 * n_tiles * TILE_BYTES + 4 * kv_bytes + 2 * TILE_BYTES + 1024 + 256;
  if (smem > 227 * 1024) return 1;
  static PerDevice<size_t> cur_cache;
  size_t& cur = cur_cache.get();
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(spatial_attention_hp_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return set_error("attention(hp sm100) attr: %s", cudaGetErrorString(e));
    cur = smem;
  }
  const uint64_t BN = static_cast<uint64_t>(B) * N;
  CUtensorMap tmIn, tmOutA, tmOutB;
  if (int rc = make_tmap_bf16_3d(&tmIn, qkv, 6 * DMODEL, T, BN, ld * 2, static_cast<uint64_t>(T) * ld * 2, HD, 1, N))
    return rc;
  if (int rc = make_tmap_bf16_3d(&tmOutA, out, 2 * DMODEL, T, BN, ldo * 2, static_cast<uint64_t>(T) * ldo * 2, HD, 1,
                                 p.rows_a))
    return rc;
  if (int rc = make_tmap_bf16_3d(&tmOutB, out, 2 * DMODEL, T, BN, ldo * 2, static_cast<uint64_t>(T) * ldo * 2, HD, 1,
                                 p.rows_b > 0 ? p.rows_b : 1))
    return rc;
  const int grid = p.n_prob < num_sms() ? p.n_prob : num_sms();
  cudaError_t e = cudaMemsetAsync(p.fix_count, 0, sizeof(int), st);
  if (e != cudaSuccess) return set_error("attention(hp sm100) memset: %s", cudaGetErrorString(e));
  spatial_attention_hp_sm100_kernel<<<grid, HP_THREADS, smem, st>>>(tmIn, tmOutA, tmOutB, hp);
  attention_fixup_hp_kernel<<<grid < 64 ? grid : 64, 128, 0, st>>>(hp);
  count_launch();
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("attention(hp sm100) launch: %s", cudaGetErrorString(e));
}

}  // namespace dvst
