// tcgen05 / TMEM spatial attention for sm_100a: softmax(q k^T / 8) v over the 1 + N tokens of one frame (divided
// space-time attention's spatial half, models/timesformer.py:140-164, and space_only, :120-123), one (frame, head)
// problem at a time on a persistent CTA per SM.
//
// Token layout: the N patch rows of sequence (b, t) are rows (b*N + n)*T + t of the packed qkv matrix - a stride-T row
// gather that one 3-D TMA box (64 columns x 1 frame x N patches, SWIZZLE_128B) pulls straight into shared memory; the
// CLS row is copied behind them (attention is invariant to the key order, so CLS is token N here, not token 0).
//
//   S = Q K^T   : tcgen05.mma  M = 128 (x2 row tiles), N = NPAD (= 1 + N rounded up to 16), K = 64   -> TMEM fp32
//   softmax     : one thread per query row (tcgen05.ld 32x32b), row max, exp2, row sum; P (bf16) is written back
//                 over S in TMEM (tcgen05.st) - scores and probabilities never touch shared or global memory
//   O = P V     : tcgen05.mma with A = P from TMEM, B = V from shared memory as an MN-major operand (V rows are keys,
//                 i.e. the K dimension, so no transpose copy), M = 128, N = 64, K = NPAD               -> TMEM fp32
//   epilogue    : tcgen05.ld O, scale by 1 / row sum, bf16, swizzled staging tile, 3-D TMA store back to the token rows
//
// One-pass softmax (kOnePass, used when the caller provides scratch): softmax is invariant to the reference point, and any
// reference within ~44 nats of the row maximum keeps every exp2 and the fp32 / bf16 sums in range, so the reference is
// the maximum of the FIRST 32 keys and the CLS key only (one TMEM round trip instead of a full pass over S).  A row whose exponentials
// still grow past 2^64 (its maximum lies > 44 nats above those 32 keys) is appended to a fix-up list, and
// attention_fixup_kernel recomputes exactly those rows in fp32 from the packed qkv afterwards - results stay exact for
// every input, the common case pays nothing.
//
// Warp roles (384 threads): warp 0 TMA producer + CLS-row copy, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7
// softmax / epilogue of row tile A (rows 0-127), warps 8-11 of row tile B (rows 128-255).  TMEM: 2 x 256 columns; per
// tile S occupies [0, NPAD), P overwrites [0, NPAD/2), O lands in [128, 192) once S has been consumed.
// Shared memory: 2 stages of {Q, K, V} so the loads of problem i+1 overlap the math of problem i.
#include "attn_sm100_common.cuh"

namespace dvst {

using namespace attn;

namespace {

template <bool kOnePass>
__global__ void __launch_bounds__(NUM_THREADS, 1)
spatial_attention_sm100_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOutA,
                               const __grid_constant__ CUtensorMap tmOutB, const SpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int n_tiles = p.S > 128 ? 2 : 1;
  const uint32_t q_bytes = n_tiles * TILE_BYTES;
  const uint32_t kv_bytes = (static_cast<uint32_t>(p.NPAD) * 128u + 1023u) & ~1023u;
  const uint32_t stage_bytes = q_bytes + 2 * kv_bytes;
  const uint32_t sO = base + 2 * stage_bytes;          // output staging, one 128-row tile per row tile
  const uint32_t bars = sO + n_tiles * TILE_BYTES;
  const uint32_t full_bar = bars;            // [2]  TMA + CLS copy -> MMA
  const uint32_t empty_bar = bars + 16;      // [2]  MMA (all reads of the stage done) -> producer
  const uint32_t sfull_bar = bars + 32;      // [2 tiles]  S ready
  const uint32_t pfull_bar = bars + 48;      // [2 tiles]  P written
  const uint32_t ofull_bar = bars + 64;      // [2 tiles]  O ready
  const uint32_t oempty_bar = bars + 80;     // [2 tiles]  O read out: the tile's TMEM region may be overwritten
  const uint32_t tmem_slot = bars + 96;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.S, N = p.N, NPAD = p.NPAD;
  // warps of a row tile that own at least one real query row
  const int act_a = min(4, (min(S, 128) + 31) >> 5);
  const int act_b = n_tiles == 2 ? (S - 128 + 31) >> 5 : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmOutA);
    tma_prefetch_desc(&tmOutB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(sfull_bar, 1);
    mbar_init(sfull_bar + 8, 1);
    mbar_init(pfull_bar, act_a);
    mbar_init(pfull_bar + 8, act_b > 0 ? act_b : 1);
    mbar_init(ofull_bar, 1);
    mbar_init(ofull_bar + 8, 1);
    mbar_init(oempty_bar, act_a);
    mbar_init(oempty_bar + 8, act_b > 0 ? act_b : 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  // K / V rows [S, NPAD) are never written by the loads: zero them once (P is 0 there, and 0 * garbage must not be NaN)
  for (int s = 0; s < 2; ++s) {
    const int pad_chunks = (NPAD - S) * 8;
    for (int i = threadIdx.x; i < 2 * pad_chunks; i += NUM_THREADS) {
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

  if (warp == 0) {
    // ===== producer: 3 TMA boxes (Q, K, V patch rows) + the CLS row of each =====
    int it = 0;
    for (int prob = blockIdx.x; prob < p.n_prob; prob += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int seq = prob / p.heads, head = prob - seq * p.heads;
      const int b = seq / p.T, t = seq - b * p.T;
      if (lane == 0) mbar_wait(empty_bar + 8 * s, ph ^ 1);
      __syncwarp();
      const uint32_t st = base + s * stage_bytes;
      if (elect_one()) {
        tma_load_3d(st, &tmIn, full_bar + 8 * s, head * HD, t, b * N);
        tma_load_3d(st + q_bytes, &tmIn, full_bar + 8 * s, DMODEL + head * HD, t, b * N);
        tma_load_3d(st + q_bytes + kv_bytes, &tmIn, full_bar + 8 * s, 2 * DMODEL + head * HD, t, b * N);
      }
      if (lane < 24) {
        const int which = lane >> 3, ch = lane & 7;
        const long long cls_row = p.cls_base + (p.mode == 1 ? b : seq);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + cls_row * p.ld + which * DMODEL + head * HD) + ch);
        const uint32_t off = (which == 0 ? 0u : q_bytes + (which - 1) * kv_bytes) + N * 128 + ((ch ^ (N & 7)) << 4);
        *reinterpret_cast<uint4*>(smem + s * stage_bytes + off) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(full_bar + 8 * s, 3u * N * 128u);
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer (whole warp in the loop, one elected lane issues - elect_one() in ptx.cuh) =====
      const uint32_t idesc_s = make_idesc(128, NPAD, 0);
      constexpr uint32_t idesc_pv = make_idesc(128, HD, 1);
      const int ksteps = NPAD >> 4;
      auto issue_s = [&](int tile, uint32_t st) {
        const uint32_t d = tmem_base + tile * 256;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16<1>(d, make_kmajor_sw128_desc(st + tile * TILE_BYTES + k * 32),
                         make_kmajor_sw128_desc(st + q_bytes + k * 32), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(sfull_bar + 8 * tile);
        }
      };
      // Software-pipelined issue order: S of problem i+1 for a row tile goes out as soon as that tile's O of problem i
      // has been read back, i.e. BEFORE waiting for the other tile's P - the two tiles' chains
      // (S -> softmax -> P V -> read-out -> next S) run decoupled and keep the MUFU busy from either one.
      int it = 0;
      if (static_cast<int>(blockIdx.x) < p.n_prob) {
        mbar_wait(full_bar, 0);
        tc_fence_after();
        for (int tile = 0; tile < n_tiles; ++tile) issue_s(tile, base);
      }
      for (int prob = blockIdx.x; prob < p.n_prob; prob += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t tp = it & 1;
        const uint32_t st = base + s * stage_bytes;
        const bool has_next = prob + static_cast<int>(gridDim.x) < p.n_prob;
        const int s_next = (it + 1) & 1;
        const uint32_t ph_next = ((it + 1) >> 1) & 1;
        for (int tile = 0; tile < n_tiles; ++tile) {
          mbar_wait(pfull_bar + 8 * tile, tp);
          tc_fence_after();
          const uint32_t region = tmem_base + tile * 256;
          if (elect_one()) {
            const uint64_t vdesc0 = make_mnmajor_sw128_desc(st + q_bytes + kv_bytes);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < ksteps)
                umma_ts_bf16(region + O_COL, region + j * 8, vdesc0 + static_cast<uint64_t>(j * 128), idesc_pv,
                             j != 0 ? 1u : 0u);
            umma_commit(ofull_bar + 8 * tile);
            if (tile == n_tiles - 1) umma_commit(empty_bar + 8 * s);  // every MMA reading this stage has been issued
          }
          if (has_next) {
            if (tile == 0) {
              mbar_wait(full_bar + 8 * s_next, ph_next);
              tc_fence_after();
            }
            mbar_wait(oempty_bar + 8 * tile, tp);  // this tile's O (problem i) has been read out of TMEM
            tc_fence_after();
            issue_s(tile, base + s_next * stage_bytes);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== softmax + epilogue: one thread per query row =====
    const int tile = (warp - 4) >> 2;
    const int q = warp & 3;
    const int n_act = tile == 0 ? act_a : act_b;
    if (tile < n_tiles && q < n_act) {
      const int rloc = q * 32 + lane;          // row inside the tile
      const int row = tile * 128 + rloc;       // token index in the sequence (N = CLS)
      const uint32_t treg = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tile * 256;
      const uint32_t so_tile = sO + tile * TILE_BYTES;
      const bool store_leader = q == 0 && lane == 0;
      const int rows_tile = tile == 0 ? p.rows_a : p.rows_b;
      const float scale_log2 = 0.125f * 1.4426950408889634f;
      const int nch = (NPAD + 31) >> 5;
      int it = 0;
      for (int prob = blockIdx.x; prob < p.n_prob; prob += gridDim.x, ++it) {
        const uint32_t tp = it & 1;
        const int seq = prob / p.heads, head = prob - seq * p.heads;
        const int b = seq / p.T, t = seq - b * p.T;
        mbar_wait(sfull_bar + 8 * tile, tp);
        tc_fence_after();
        // ---- pass 1: row max (one-pass softmax: of the first 32 keys and the CLS key only) ----
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        {
          uint32_t v[32], v2[32];
          tmem_ld_32x32(treg, v);
          if constexpr (kOnePass) {
            // ... plus the CLS key (token N): heads that use CLS as an attention sink put their row maximum there, and
            // the single extra column rides in the same TMEM round trip
            uint32_t sc;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(sc) : "r"(treg + N) : "memory");
            tmem_ld_wait();
            mx4[3] = __uint_as_float(sc);
          }
#pragma unroll
          for (int c = 0; c < (kOnePass ? 1 : 8); ++c) {
            if (c < nch) {
              uint32_t (&cur)[32] = (c & 1) ? v2 : v;
              uint32_t (&nxt)[32] = (c & 1) ? v : v2;
              tmem_ld_wait();
              if (!kOnePass && c + 1 < nch) tmem_ld_32x32(treg + (c + 1) * 32, nxt);
              if (c * 32 + 32 <= S) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  mx4[j & 3] = max3(mx4[j & 3], __uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (c * 32 + j < S) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(cur[j]));
              }
            }
          }
        }
        // ---- pass 2: p = exp2((s - max) * scale), row sum, P -> TMEM (bf16 pairs over the consumed S columns) ----
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const float moff = mx * scale_log2;
        uint64_t sum2[2] = {0ull, 0ull};
        const uint64_t scale2 = pk2(scale_log2, scale_log2), nmoff2 = pk2(-moff, -moff);
        {
          uint32_t v[32], v2[32];
          tmem_ld_32x32(treg, v);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c < nch) {
              uint32_t (&cur)[32] = (c & 1) ? v2 : v;
              uint32_t (&nxt)[32] = (c & 1) ? v : v2;
              tmem_ld_wait();
              if (c + 1 < nch) tmem_ld_32x32(treg + (c + 1) * 32, nxt);
              float e[32];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float a0, a1;
                unpk2(fma2(pk2(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), scale2, nmoff2), a0, a1);
                e[2 * j] = ex2_approx(a0);
                e[2 * j + 1] = ex2_approx(a1);
              }
              if (c * 32 + 32 > S) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (c * 32 + j >= S) e[j] = 0.f;
              }
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                sum2[j & 1] = add2(sum2[j & 1], pk2(e[2 * j], e[2 * j + 1]));
                pk[j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
              }
              tmem_st_32x16(treg + c * 16, pk);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pfull_bar + 8 * tile);
        // ---- O ----
        mbar_wait(ofull_bar + 8 * tile, tp);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(treg + O_COL, o0);
        tmem_ld_32x32(treg + O_COL + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(oempty_bar + 8 * tile);
        float sa, sb;
        unpk2(add2(sum2[0], sum2[1]), sa, sb);
        const float inv = 1.0f / (sa + sb);
        if constexpr (kOnePass) {
          // exponentials past 2^64 (or NaN): this row is recomputed exactly by attention_fixup_kernel
          if (row < S && !(sa + sb < 1.8e19f)) p.fix_list[atomicAdd(p.fix_count, 1)] = make_int2(prob, row);
        }
        // the previous problem's TMA store must have finished reading the staging tile
        if (store_leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
      }
      if (store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// Exact fp32 recomputation of the rows the one-pass softmax could not normalise: one warp per (problem, row).
__global__ void __launch_bounds__(128) attention_fixup_kernel(const SpParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const int n_fix = *p.fix_count;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n_fix; i += n_warps) {
    const int2 e = p.fix_list[i];
    const int prob = e.x, row = e.y;
    const int seq = prob / p.heads, head = prob - seq * p.heads;
    const int b = seq / p.T, t = seq - b * p.T;
    const long long cls_row = p.cls_base + (p.mode == 1 ? b : seq);
    auto token_row = [&](int j) { return j < p.N ? (static_cast<long long>(b) * p.N + j) * p.T + t : cls_row; };
    const __nv_bfloat16* qp = p.qkv + token_row(row) * p.ld + head * HD;
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = __bfloat162float(qp[d]);
    // scores of this lane's keys (j = lane, lane + 32, ...): S <= 256 -> at most 8 per lane
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
        for (int d = 0; d < HD; ++d) acc = fmaf(q[d], __bfloat162float(kp[d]), acc);
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
        for (int d = 0; d < HD; ++d) acc[d] = fmaf(w, __bfloat162float(vp[d]), acc[d]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    const float inv = 1.0f / sum;
    // reduce the 64 output dims over the warp; lane d (and d + 32) keeps dims d, d + 32
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
    if (row < p.N) {
      __nv_bfloat16* dst = p.out + token_row(row) * p.ldo + head * HD;
      dst[lane] = __float2bfloat16(o0);
      dst[lane + 32] = __float2bfloat16(o1);
    } else if (p.mode == 1) {
      float* dst = p.cls_out + static_cast<long long>(seq) * DMODEL + head * HD;
      dst[lane] = o0;
      dst[lane + 32] = o1;
    } else {
      __nv_bfloat16* dst = p.out + (p.cls_base + seq) * p.ldo + head * HD;
      dst[lane] = __float2bfloat16(o0);
      dst[lane + 32] = __float2bfloat16(o1);
    }
  }
}

}  // namespace

// 1 = tcgen05 kernel where it applies (one-pass softmax when scratch is given), 2 = tcgen05 with the two-pass softmax,
// 0 = mma.sync kernels only
// 3 = tcgen05 first-generation kernel with the one-pass softmax (A/B runs against the split kernel, which is what 1 picks
// for 32 <= NPAD <= 208)
static int g_attention_impl = 1;
void set_attention_impl(int impl) { g_attention_impl = impl == 0 ? 0 : (impl == 2 ? 2 : (impl == 3 ? 3 : 1)); }
int attention_impl() { return g_attention_impl; }
int launch_spatial_attention_split_sm100(const SpParams& p, const CUtensorMap& tmIn, const CUtensorMap& tmOutA,
                                         const CUtensorMap& tmOutB, int grid, cudaStream_t st);

// Returns 1 if the shape is not handled here (the caller falls back to the mma.sync kernels), 0 on success, < 0 error.
long long spatial_attention_scratch_bytes(int B, int T, int N) {
  return 256 + static_cast<long long>(B) * T * 12 * (N + 1) * static_cast<long long>(sizeof(int2));
}

int launch_spatial_attention_sm100(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode,
                                   int B, int T, int N, long long cls_base, cudaStream_t st, void* scratch,
                                   long long scratch_bytes) {
  if (!g_attention_impl) return 1;
  const int S = N + 1;
  if ((mode != 1 && mode != 3) || S > 256 || N < 1) return 1;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (ld % 8) || (ldo % 8))
    return 1;
  SpParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.ld = ld;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.cls_out = cls_out;
  p.mode = mode;
  p.S = S; p.N = N; p.NPAD = (S + 15) & ~15; p.T = T;
  p.heads = 12;
  p.n_prob = B * T * p.heads;
  p.cls_base = cls_base;
  p.rows_a = N < 128 ? N : 128;
  p.rows_b = N > 128 ? N - 128 : 0;
  if (p.n_prob <= 0) return 0;
  const int n_tiles = S > 128 ? 2 : 1;
  const size_t kv_bytes = (static_cast<size_t>(p.NPAD) * 128 + 1023) & ~static_cast<size_t>(1023);
  const size_t smem = 1024 + 2 * (n_tiles * TILE_BYTES + 2 * kv_bytes) + n_tiles * TILE_BYTES + 128;
  static PerDevice<size_t> cur_cache;
  size_t& cur = cur_cache.get();
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(spatial_attention_sm100_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(spatial_attention_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (e != cudaSuccess) return set_error("attention(sm100) attr: %s", cudaGetErrorString(e));
    cur = smem;
  }
  // one-pass softmax needs the fix-up list: counter at scratch[0], entries from scratch + 256
  // (the choice must not depend on the batch size: one- and two-pass results differ in the last bits, and a clip's features
  // are promised to be independent of what else is in its batch)
  const bool one_pass = (g_attention_impl == 1 || g_attention_impl == 3) && scratch != nullptr && scratch_bytes >= spatial_attention_scratch_bytes(B, T, N) &&
                        (reinterpret_cast<uintptr_t>(scratch) & 15) == 0;
  p.fix_count = static_cast<int*>(scratch);
  p.fix_list = reinterpret_cast<int2*>(static_cast<uint8_t*>(scratch) + 256);
  const uint64_t BN = static_cast<uint64_t>(B) * N;
  CUtensorMap tmIn, tmOutA, tmOutB;
  if (int rc = make_tmap_bf16_3d(&tmIn, qkv, 3 * DMODEL, T, BN, ld * 2, static_cast<uint64_t>(T) * ld * 2, HD, 1, N))
    return rc;
  if (int rc = make_tmap_bf16_3d(&tmOutA, out, DMODEL, T, BN, ldo * 2, static_cast<uint64_t>(T) * ldo * 2, HD, 1,
                                 p.rows_a))
    return rc;
  if (int rc = make_tmap_bf16_3d(&tmOutB, out, DMODEL, T, BN, ldo * 2, static_cast<uint64_t>(T) * ldo * 2, HD, 1,
                                 p.rows_b > 0 ? p.rows_b : 1))
    return rc;
  const int grid = p.n_prob < num_sms() ? p.n_prob : num_sms();
  if (one_pass) {
    cudaError_t e = cudaMemsetAsync(p.fix_count, 0, sizeof(int), st);
    if (e != cudaSuccess) return set_error("attention(sm100) memset: %s", cudaGetErrorString(e));
    int rc2 = g_attention_impl == 1 ? launch_spatial_attention_split_sm100(p, tmIn, tmOutA, tmOutB, grid, st) : 1;
    if (rc2 != 0 && rc2 != 1) return rc2;
    if (rc2 == 1) spatial_attention_sm100_kernel<true><<<grid, NUM_THREADS, smem, st>>>(tmIn, tmOutA, tmOutB, p);
    launch_kernel_pdl(attention_fixup_kernel, dim3(grid < 64 ? grid : 64), dim3(128), 0, st, p);
    count_launch();
  } else {
    spatial_attention_sm100_kernel<false><<<grid, NUM_THREADS, smem, st>>>(tmIn, tmOutA, tmOutB, p);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("attention(sm100) launch: %s", cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------------------------------
// Long sequences - joint space-time attention (models/timesformer.py:120-123 with attention_type = 'joint_space_time'):
// 1 + N*T tokens per clip (6 273 at T = 32), 57 % of the model's FLOPs.  Flash-style on tcgen05: a persistent CTA takes a
// (clip, head, block of 256 queries) problem, keeps the two 128-row Q tiles in shared memory and streams the clip's K / V
// in 128-key tiles through a 4-stage TMA ring.  Per key tile and Q tile:
//   S = Q K^T (128 x 128 x 64, TMEM) -> online softmax, one thread per query row (running max m, running sum l; when a
//   row's max grows, its O accumulator is rescaled in TMEM with tcgen05.ld / tcgen05.st - the warp does that only if one of
//   its rows needs it) -> P (bf16) written over S -> O += P V (A = P from TMEM, V MN-major from shared memory).
// The tokens of a clip are contiguous rows (patch tokens b*NT .. b*NT + NT - 1) plus the clip's CLS row, kept as the LAST
// token (attention does not depend on the key order): every tile is one 2-D TMA box, except the last one, which is a
// shorter box (NT % 128 rows) plus a 128-byte copy of the CLS row; its padding keys are masked.
// S of the next key tile is issued BEFORE P V of the current one (P has its own TMEM columns), so the softmax warps wait
// for one MMA round trip per key tile; "P V complete" is its own barrier, waited before P is overwritten / O is rescaled.
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int JT_STAGES = 4;
constexpr int JT_TILE = 128 * 128;  // bytes of one 128-row x 64-column bf16 tile
// TMEM columns of one Q tile's 256-column region: S fp32 [0, 128), P bf16 [128, 192), O fp32 [192, 256).  P has its own
// columns so that S of the NEXT key tile can be issued before P V of the current one: the softmax warps then wait for one
// MMA round trip per key tile instead of two.
constexpr int JT_P_COL = 128;
constexpr int JT_O_COL = 192;

struct JtParams {
  const __nv_bfloat16* qkv;
  long long ld;
  __nv_bfloat16* out;
  long long ldo;
  int NT, S;        // patch tokens per clip; sequence length NT + 1 (CLS = token NT)
  int rem;          // NT % 128: patch rows of the last tile (the CLS row follows them)
  int n_tiles;      // 128-token tiles per clip = NT / 128 + 1
  int n_qblk;       // query blocks (two tiles) per (clip, head)
  int n_prob, heads;
  long long cls_base;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
joint_attention_sm100_kernel(const __grid_constant__ CUtensorMap tmFull, const __grid_constant__ CUtensorMap tmRem,
                             const __grid_constant__ CUtensorMap tmOutFull, const __grid_constant__ CUtensorMap tmOutRem,
                             const JtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t sQ = base;                                   // 2 tiles
  const uint32_t sKV = base + 2 * JT_TILE;                    // JT_STAGES x {K tile, V tile}
  const uint32_t sO = sKV + JT_STAGES * 2 * JT_TILE;          // 2 output staging tiles
  const uint32_t bars = sO + 2 * JT_TILE;
  const uint32_t q_full = bars, q_empty = bars + 8;
  const uint32_t kv_full = bars + 16;                         // [JT_STAGES]
  const uint32_t kv_empty = kv_full + 8 * JT_STAGES;          // [JT_STAGES]
  const uint32_t s_full = kv_empty + 8 * JT_STAGES;           // [2]
  const uint32_t p_full = s_full + 16;                        // [2]
  const uint32_t o_full = p_full + 16;                        // [2]  P V of a key tile complete (one phase per key tile)
  const uint32_t o_empty = o_full + 16;                       // [2]
  const uint32_t tmem_slot = o_empty + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmFull);
    tma_prefetch_desc(&tmRem);
    tma_prefetch_desc(&tmOutFull);
    tma_prefetch_desc(&tmOutRem);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < JT_STAGES; ++s) {
      mbar_init(kv_full + 8 * s, 1);
      mbar_init(kv_empty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full + 8 * t, 1);
      mbar_init(p_full + 8 * t, 4);
      mbar_init(o_full + 8 * t, 1);
      mbar_init(o_empty + 8 * t, 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  // stale rows of partial tiles are multiplied by P = 0 (keys) or never stored (queries): they must be finite
  for (uint32_t i = threadIdx.x; i < (2 + 2 * JT_STAGES) * JT_TILE / 16; i += NUM_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_tiles = p.n_tiles;
  auto decode = [&](int prob, int& b, int& head, int& qb) {
    qb = prob % p.n_qblk;
    const int bh = prob / p.n_qblk;
    head = bh % p.heads;
    b = bh / p.heads;
  };

  if (warp == 0) {
    // ===== producer =====
    // one token tile (tile index t of clip b, column offset col) -> dst; returns the TMA bytes to expect
    auto load_tile = [&](uint32_t dst, uint32_t bar, int b, int t, int col) -> uint32_t {
      if (t >= n_tiles) return 0u;                       // inactive second Q tile of the last block
      const int row0 = b * p.NT + t * 128;
      if (t < n_tiles - 1) {
        if (elect_one()) tma_load_2d(dst, &tmFull, bar, col, row0);
        return JT_TILE;
      }
      if (p.rem > 0 && lane == 0) tma_load_2d(dst, &tmRem, bar, col, row0);
      if (lane < 8) {                                    // the CLS row behind the last patch rows
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + (p.cls_base + b) * p.ld + col) + lane);
        *reinterpret_cast<uint4*>(smem + (dst - base) + p.rem * 128 + ((lane ^ (p.rem & 7)) << 4)) = v;
      }
      return static_cast<uint32_t>(p.rem) * 128u;
    };
    uint32_t it_kv = 0;
    int pi = 0;
    for (int prob = blockIdx.x; prob < p.n_prob; prob += gridDim.x, ++pi) {
      int b, head, qb;
      decode(prob, b, head, qb);
      if (lane == 0) mbar_wait(q_empty, (pi & 1) ^ 1);
      __syncwarp();
      uint32_t bytes = load_tile(sQ, q_full, b, 2 * qb, head * HD);
      bytes += load_tile(sQ + JT_TILE, q_full, b, 2 * qb + 1, head * HD);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(q_full, bytes);
      for (int j = 0; j < n_tiles; ++j, ++it_kv) {
        const int s = it_kv % JT_STAGES;
        if (lane == 0) mbar_wait(kv_empty + 8 * s, ((it_kv / JT_STAGES) & 1) ^ 1);
        __syncwarp();
        const uint32_t dK = sKV + s * 2 * JT_TILE, dV = dK + JT_TILE;
        uint32_t kb = load_tile(dK, kv_full + 8 * s, b, j, DMODEL + head * HD);
        kb += load_tile(dV, kv_full + 8 * s, b, j, 2 * DMODEL + head * HD);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(kv_full + 8 * s, kb);
      }
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer: the whole warp runs the state machine (every poll result is lane 0's, so the control flow stays
      // warp-uniform), one elected lane issues - no per-instruction waterfall loops (elect_one() in ptx.cuh) =====
      auto try_wait_u = [&](uint32_t bar, uint32_t parity) -> bool {
        return __shfl_sync(0xFFFFFFFFu, mbar_try_wait(bar, parity) ? 1 : 0, 0) != 0;
      };
      constexpr uint32_t idesc_s = make_idesc(128, 128, 0);
      constexpr uint32_t idesc_pv = make_idesc(128, HD, 1);
      // The two Q tiles run decoupled: whenever a tile's P of key tile j is ready, its P V (j) and right behind it its
      // S (j + 1) are issued - so while one tile sits in the softmax (MUFU bound) the tensor pipe serves the other one,
      // and the max pass of a tile (TMEM latency bound) hides behind the other tile's exponentials.
      uint32_t kv_base = 0;   // key-tile iterations issued by the producer before this problem
      uint32_t bcs = 0;       // blocks completed per tile before this problem (parity of s_full / p_full)
      int pi = 0;
      auto issue_s = [&](int tile, uint32_t it) {
        const int s = it % JT_STAGES;
        tc_fence_after();
        const uint32_t aK = sKV + s * 2 * JT_TILE;
        const uint32_t d = tmem_base + tile * 256;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16<1>(d, make_kmajor_sw128_desc(sQ + tile * JT_TILE + k * 32), make_kmajor_sw128_desc(aK + k * 32),
                         idesc_s, k != 0 ? 1u : 0u);
          umma_commit(s_full + 8 * tile);
        }
      };
      for (int prob = blockIdx.x; prob < p.n_prob; prob += gridDim.x, ++pi) {
        mbar_wait(q_full, pi & 1);
        tc_fence_after();
        int jt[2] = {0, 0};                 // current key tile per Q tile
        bool need_s[2] = {true, true};      // true: S (jt) not issued yet; false: waiting for P (jt)
        // every wait in this loop is a poll: a blocking wait for one tile (e.g. for a K / V stage that the OTHER tile has
        // not released yet) would stop this thread from serving the other tile and deadlock the pair
        while (jt[0] < n_tiles || jt[1] < n_tiles) {
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            const int j = jt[tile];
            if (j >= n_tiles) continue;
            const uint32_t it = kv_base + j;
            const int s = it % JT_STAGES;
            if (need_s[tile]) {
              if (!try_wait_u(kv_full + 8 * s, (it / JT_STAGES) & 1)) continue;
              issue_s(tile, it);
              need_s[tile] = false;
              continue;
            }
            if (!try_wait_u(p_full + 8 * tile, (bcs + j) & 1)) continue;
            // (the softmax warps release O of the previous problem on their own: this wait cannot involve the other tile)
            if (j == 0 && pi > 0) mbar_wait(o_empty + 8 * tile, (pi - 1) & 1);
            tc_fence_after();
            // S of the next key tile first (the softmax warps have finished reading S of this one), if its K is there
            bool next_s_issued = false;
            if (j + 1 < n_tiles) {
              const uint32_t itn = it + 1;
              if (try_wait_u(kv_full + 8 * (itn % JT_STAGES), (itn / JT_STAGES) & 1)) {
                issue_s(tile, itn);
                next_s_issued = true;
              }
            }
            const uint32_t aV = sKV + s * 2 * JT_TILE + JT_TILE;
            const uint32_t region = tmem_base + tile * 256;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_ts_bf16(region + JT_O_COL, region + JT_P_COL + k * 8, make_mnmajor_sw128_desc(aV + k * 2048), idesc_pv,
                             (j | k) != 0 ? 1u : 0u);
              umma_commit(o_full + 8 * tile);
              // the stage is free once BOTH tiles' MMAs on it have completed: the later of the two commits it
              if (jt[tile ^ 1] > j) umma_commit(kv_empty + 8 * s);
            }
            jt[tile] = j + 1;
            need_s[tile] = !next_s_issued;
          }
        }
        if (elect_one()) umma_commit(q_empty);
        kv_base += n_tiles;
        bcs += n_tiles;
      }
    }
  } else if (warp >= 4) {
    // ===== online softmax + epilogue: one thread per query row =====
    const int tile = (warp - 4) >> 2;
    const int q = warp & 3;
    const int rloc = q * 32 + lane;
    const uint32_t treg = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tile * 256;
    const uint32_t so_tile = sO + tile * JT_TILE;
    const bool store_leader = q == 0 && lane == 0;
    const float scale_log2 = 0.125f * 1.4426950408889634f;
    const uint64_t scale2 = pk2(scale_log2, scale_log2);
    uint32_t bc = 0;
    int pi = 0;
    for (int prob = blockIdx.x; prob < p.n_prob; prob += gridDim.x, ++pi) {
      int b, head, qb;
      decode(prob, b, head, qb);
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++bc) {
        const int valid = j == n_tiles - 1 ? p.rem + 1 : 128;   // keys of this tile that exist
        mbar_wait(s_full + 8 * tile, bc & 1);
        tc_fence_after();
        // ---- pass 1: maximum of the tile (two 64-column TMEM loads: the pass is latency bound) ----
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[64];
          tmem_ld_32x64(treg + h * 64, v);
          tmem_ld_wait();
          if (h * 64 + 64 <= valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              mx4[i & 3] = max3(mx4[i & 3], __uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (h * 64 + i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
          }
        }
        const float bmax = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // ---- lazy reference: m only moves when a row's maximum exceeds it by more than 2^8 (44 raw score units), so the
        // probabilities stay <= 256 and the O / l rescale (tcgen05.ld + st of 64 columns, decided per warp) becomes rare
        // after the first key tiles instead of running whenever any of a warp's 32 rows sees a new maximum ----
        const bool grow = j == 0 || bmax > m + 44.0f;
        // P V of the previous key tile must have completed before its P columns are overwritten and before O is touched
        if (j > 0) {
          mbar_wait(o_full + 8 * tile, (bc - 1) & 1);
          tc_fence_after();
        }
        if (j > 0 && __any_sync(0xFFFFFFFFu, grow)) {
          const float alpha = grow ? ex2_approx((m - bmax) * scale_log2) : 1.0f;
          l *= alpha;
          const uint64_t a2 = pk2(alpha, alpha);
          uint32_t o[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld_32x32(treg + JT_O_COL + h * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float x0, x1;
              unpk2(fma2(pk2(__uint_as_float(o[2 * i]), __uint_as_float(o[2 * i + 1])), a2, 0ull), x0, x1);
              o[2 * i] = __float_as_uint(x0);
              o[2 * i + 1] = __float_as_uint(x1);
            }
            tmem_st_32x32(treg + JT_O_COL + h * 32, o);
          }
          tmem_st_wait();
        }
        if (grow) m = bmax;
        // ---- pass 2: p = exp2((s - m) * scale), row sum, P (bf16) over the consumed S columns ----
        const float moff = m * scale_log2;
        const uint64_t nmoff2 = pk2(-moff, -moff);
        uint64_t sum2[2] = {0ull, 0ull};
        {
          uint32_t v[32], v2[32];
          tmem_ld_32x32(treg, v);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t (&cur)[32] = (c & 1) ? v2 : v;
            uint32_t (&nxt)[32] = (c & 1) ? v : v2;
            tmem_ld_wait();
            if (c < 3) tmem_ld_32x32(treg + (c + 1) * 32, nxt);
            float e[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float a0, a1;
              unpk2(fma2(pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), scale2, nmoff2), a0, a1);
              e[2 * i] = ex2_approx(a0);
              e[2 * i + 1] = ex2_approx(a1);
            }
            if (c * 32 + 32 > valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i >= valid) e[i] = 0.f;
            }
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              sum2[i & 1] = add2(sum2[i & 1], pk2(e[2 * i], e[2 * i + 1]));
              pk[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
            }
            tmem_st_32x16(treg + JT_P_COL + c * 16, pk);
          }
        }
        {
          float sa, sb;
          unpk2(add2(sum2[0], sum2[1]), sa, sb);
          l += sa + sb;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + 8 * tile);
      }
      // ---- O of this problem (P V of the last key tile complete) ----
      mbar_wait(o_full + 8 * tile, (bc - 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(treg + JT_O_COL, o0);
      tmem_ld_32x32(treg + JT_O_COL + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty + 8 * tile);
      const float inv = 1.0f / l;
      const int t_idx = 2 * qb + tile;                 // token tile of this clip
      const int token = t_idx * 128 + rloc;
      if (store_leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      named_bar_sync(1 + tile, 128);
      if (token < p.NT) {
        const uint32_t dst = so_tile + rloc * 128;
        const int sw = rloc & 7;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint32_t (&o)[32] = jj < 4 ? o0 : o1;
          const int j8 = (jj & 3) * 8;
          const uint32_t w0 = pack_bf16x2(__uint_as_float(o[j8]) * inv, __uint_as_float(o[j8 + 1]) * inv);
          const uint32_t w1 = pack_bf16x2(__uint_as_float(o[j8 + 2]) * inv, __uint_as_float(o[j8 + 3]) * inv);
          const uint32_t w2 = pack_bf16x2(__uint_as_float(o[j8 + 4]) * inv, __uint_as_float(o[j8 + 5]) * inv);
          const uint32_t w3 = pack_bf16x2(__uint_as_float(o[j8 + 6]) * inv, __uint_as_float(o[j8 + 7]) * inv);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((jj ^ sw) << 4)), "r"(w0), "r"(w1),
                       "r"(w2), "r"(w3) : "memory");
        }
      } else if (token == p.NT) {
        __nv_bfloat16* dst = p.out + (p.cls_base + b) * p.ldo + head * HD;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint32_t (&o)[32] = jj < 4 ? o0 : o1;
          const int j8 = (jj & 3) * 8;
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[j8]) * inv, __uint_as_float(o[j8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[j8 + 2]) * inv, __uint_as_float(o[j8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[j8 + 4]) * inv, __uint_as_float(o[j8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[j8 + 6]) * inv, __uint_as_float(o[j8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + jj * 8) = w;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + tile, 128);
      if (store_leader && t_idx < n_tiles) {
        const int row0 = b * p.NT + t_idx * 128;
        if (t_idx < n_tiles - 1) tma_store_2d(&tmOutFull, so_tile, head * HD, row0);
        else if (p.rem > 0) tma_store_2d(&tmOutRem, so_tile, head * HD, row0);
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

}  // namespace

// Returns 1 if the shape is not handled here (the caller falls back to the mma.sync streaming kernel), 0 on success.
int launch_joint_attention_sm100(const void* qkv, long long ld, void* out, long long ldo, int B, int NT,
                                 long long cls_base, cudaStream_t st) {
  if (!g_attention_impl) return 1;
  if (NT < 1 || B < 1) return 1;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (ld % 8) || (ldo % 8)) return 1;
  JtParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.ld = ld;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.NT = NT; p.S = NT + 1;
  p.rem = NT % 128;
  p.n_tiles = NT / 128 + 1;
  p.n_qblk = (p.n_tiles + 1) / 2;
  p.heads = 12;
  const long long n_prob = static_cast<long long>(B) * p.heads * p.n_qblk;
  if (n_prob > 0x7FFFFFFF || static_cast<long long>(B) * NT > 0x7FFFFFFF) return 1;
  p.n_prob = static_cast<int>(n_prob);
  p.cls_base = cls_base;
  const size_t smem = 1024 + (2 + 2 * JT_STAGES + 2) * static_cast<size_t>(JT_TILE) + 256;
  static PerDevice<bool> attr_cache;
  bool& attr_set = attr_cache.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(joint_attention_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return set_error("attention(joint sm100) attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const uint64_t rows = static_cast<uint64_t>(B) * NT;   // patch rows: the boxes never read the CLS rows behind them
  CUtensorMap tmFull, tmRem, tmOutFull, tmOutRem;
  if (int rc = make_tmap_bf16_2d(&tmFull, qkv, rows, 3 * DMODEL, ld, 128, HD)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmRem, qkv, rows, 3 * DMODEL, ld, p.rem > 0 ? p.rem : 1, HD)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmOutFull, out, rows, DMODEL, ldo, 128, HD)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmOutRem, out, rows, DMODEL, ldo, p.rem > 0 ? p.rem : 1, HD)) return rc;
  const int grid = p.n_prob < num_sms() ? p.n_prob : num_sms();
  joint_attention_sm100_kernel<<<grid, NUM_THREADS, smem, st>>>(tmFull, tmRem, tmOutFull, tmOutRem, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("attention(joint sm100) launch: %s", cudaGetErrorString(e));
}

}  // namespace dvst
