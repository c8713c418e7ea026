// Fused DINO head last layer + centred / sharpened softmax + cross-entropy ("mode H" scoring).
//
// For P (student, teacher) pairs of L2-normalised bottleneck vectors zs, zt in R^256 and the weight-normalised
// last-layer matrix W [out_dim, 256] (vision_transformer.py:266-282), computes
//     loss[p] = logsumexp(ls / ts) - sum_j softmax((lt - center) / tt)_j * (ls_j / ts),   ls = W zs[p], lt = W zt[p]
// (dino_similarity.py:131-133 / train_ssl.py:653-668 per pair) WITHOUT ever writing the out_dim-wide logits:
//
//   * each CTA keeps the student and teacher A tiles (128 pairs x 256, bf16, 2 x 64 KiB) resident in shared memory,
//   * streams its share of W in 64-row tiles through a 3-stage TMA ring (the kernel is bound by this 33.5 MB stream),
//   * issues two tcgen05.mma chains per tile that SHARE the W tile and write adjacent TMEM column ranges
//     (student logits in columns [0,64), teacher logits in [64,128)), so the epilogue thread that owns pair p reads
//     BOTH logits of every column from its own TMEM lane,
//   * folds the tile into per-pair running (max, sum-exp) for student and teacher and the cross term
//     sum_j exp(t'_j - max_t) * s'_j, rescaled online,
//   * writes 5 floats per (CTA, pair); a tiny second kernel merges the per-CTA partials into the loss.
#include "dvst_internal.h"
#include "ptx.cuh"

namespace dvst {

namespace {

constexpr int HM = 128;        // pairs per M tile
constexpr int HN = 64;         // W rows (logit columns) per tile
constexpr int HK = 256;        // bottleneck dim
constexpr int HKB = HK / 64;   // 64-wide k-blocks
constexpr int HSTAGES = 3;
constexpr int A_TILE_BYTES = HM * 64 * 2;       // 16 KiB per k-block
constexpr int A_BYTES_ALL = HKB * A_TILE_BYTES;  // 64 KiB per operand
constexpr int W_TILE_BYTES = HN * 64 * 2;        // 8 KiB per k-block
constexpr int W_STAGE_BYTES = HKB * W_TILE_BYTES;  // 32 KiB
constexpr int H_THREADS = 256;
constexpr int H_TMEM_COLS = 256;  // 2 accumulator stages x (64 student + 64 teacher)
constexpr int H_SMEM = 1024 + 2 * A_BYTES_ALL + HSTAGES * W_STAGE_BYTES + 256;

struct HeadLossParams {
  int P, out_dim;
  const float* center;  // [out_dim] or null
  float inv_tt, inv_ts;
  float* part;          // [5][m_tiles][gridDim.x][128]
  int n_tiles;          // out_dim / 64
};

__global__ void __launch_bounds__(H_THREADS, 1)
head_loss_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmT,
                 const __grid_constant__ CUtensorMap tmW, const HeadLossParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t sAs = base, sAt = base + A_BYTES_ALL, sW = base + 2 * A_BYTES_ALL;
  const uint32_t bars = sW + HSTAGES * W_STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * HSTAGES, a_bar = bars + 16 * HSTAGES;
  const uint32_t tfull_bar = a_bar + 8, tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmT); tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HSTAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    mbar_init(a_bar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar + 8 * a, 1); mbar_init(tempty_bar + 8 * a, 4); }
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc<1>(tmem_slot, H_TMEM_COLS); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    {
      // (whole warp in the loop, one elected lane issues: no per-instruction waterfall loops - elect_one() in ptx.cuh)
      if (elect_one()) {
        mbar_arrive_expect_tx(a_bar, 2 * A_BYTES_ALL);
#pragma unroll
        for (int kb = 0; kb < HKB; ++kb) {
          tma_load_2d(sAs + kb * A_TILE_BYTES, &tmS, a_bar, kb * 64, m_tile * HM);
          tma_load_2d(sAt + kb * A_TILE_BYTES, &tmT, a_bar, kb * 64, m_tile * HM);
        }
      }
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        mbar_wait(empty_bar + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar + 8 * s, W_STAGE_BYTES);
#pragma unroll
          for (int kb = 0; kb < HKB; ++kb)
            tma_load_2d(sW + s * W_STAGE_BYTES + kb * W_TILE_BYTES, &tmW, full_bar + 8 * s, kb * 64, tile * HN);
        }
        if (++s == HSTAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = make_idesc_bf16_f32(HM, HN);
      mbar_wait(a_bar, 0);
      tc_fence_after();
      int s = 0, a = 0; uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar + 8 * a, aph ^ 1);
        mbar_wait(full_bar + 8 * s, ph);
        tc_fence_after();
        const uint32_t d_s = tmem_base + a * 128, d_t = d_s + 64;
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < HKB; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t wdesc = make_kmajor_sw128_desc(sW + s * W_STAGE_BYTES + kb * W_TILE_BYTES + k * 32);
              umma_bf16<1>(d_s, make_kmajor_sw128_desc(sAs + kb * A_TILE_BYTES + k * 32), wdesc, idesc, (kb | k) != 0);
              umma_bf16<1>(d_t, make_kmajor_sw128_desc(sAt + kb * A_TILE_BYTES + k * 32), wdesc, idesc, (kb | k) != 0);
            }
          }
          umma_commit(empty_bar + 8 * s);
          umma_commit(tfull_bar + 8 * a);
        }
        if (++s == HSTAGES) { s = 0; ph ^= 1; }
        if (++a == 2) { a = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    float ms = -INFINITY, zs = 0.f, mt = -INFINITY, zt = 0.f, cr = 0.f;
    int a = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(tfull_bar + 8 * a, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t vs[32], vt[32];
        tmem_ld_32x32(trow + c * 32, vs);
        tmem_ld_32x32(trow + 64 + c * 32, vt);
        tmem_ld_wait();
        const int col0 = tile * HN + c * 32;
        float sp[32], tp[32];
        float cms = -INFINITY, cmt = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float cj = p.center ? __ldg(p.center + col0 + j) : 0.f;
          sp[j] = __uint_as_float(vs[j]) * p.inv_ts;
          tp[j] = (__uint_as_float(vt[j]) - cj) * p.inv_tt;
          cms = fmaxf(cms, sp[j]);
          cmt = fmaxf(cmt, tp[j]);
        }
        const float nms = fmaxf(ms, cms), nmt = fmaxf(mt, cmt);
        const float scs = __expf(ms - nms), sct = __expf(mt - nmt);
        float as = 0.f, at = 0.f, ac = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          as += __expf(sp[j] - nms);
          const float e = __expf(tp[j] - nmt);
          at += e;
          ac = fmaf(e, sp[j], ac);
        }
        zs = zs * scs + as; ms = nms;
        zt = zt * sct + at; cr = cr * sct + ac; mt = nmt;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * a);
      if (++a == 2) { a = 0; aph ^= 1; }
    }
    const int row = q * 32 + lane;
    const long long stride = static_cast<long long>(gridDim.y) * gridDim.x * HM;
    float* dst = p.part + (static_cast<long long>(m_tile) * gridDim.x + blockIdx.x) * HM + row;
    dst[0] = ms; dst[stride] = zs; dst[2 * stride] = mt; dst[3 * stride] = zt; dst[4 * stride] = cr;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem_base, H_TMEM_COLS); }
}

// One warp per pair: merge the per-CTA partials.
__global__ void __launch_bounds__(128) head_loss_merge_kernel(const float* __restrict__ part, int n_cta, int m_tiles,
                                                              int P, float* __restrict__ loss) {
  const int pidx = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (pidx >= P) return;
  const int lane = threadIdx.x & 31;
  const int m_tile = pidx / HM, row = pidx % HM;
  const long long stride = static_cast<long long>(m_tiles) * n_cta * HM;
  const float* b = part + static_cast<long long>(m_tile) * n_cta * HM + row;
  float Ms = -INFINITY, Mt = -INFINITY;
  for (int c = lane; c < n_cta; c += 32) {
    Ms = fmaxf(Ms, b[static_cast<long long>(c) * HM]);
    Mt = fmaxf(Mt, b[2 * stride + static_cast<long long>(c) * HM]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Ms = fmaxf(Ms, __shfl_xor_sync(0xffffffffu, Ms, o));
    Mt = fmaxf(Mt, __shfl_xor_sync(0xffffffffu, Mt, o));
  }
  float Zs = 0.f, Zt = 0.f, C = 0.f;
  for (int c = lane; c < n_cta; c += 32) {
    const long long o = static_cast<long long>(c) * HM;
    const float ws = expf(b[o] - Ms), wt = expf(b[2 * stride + o] - Mt);
    Zs += b[stride + o] * ws;
    Zt += b[3 * stride + o] * wt;
    C += b[4 * stride + o] * wt;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Zs += __shfl_xor_sync(0xffffffffu, Zs, o);
    Zt += __shfl_xor_sync(0xffffffffu, Zt, o);
    C += __shfl_xor_sync(0xffffffffu, C, o);
  }
  if (lane == 0) loss[pidx] = (Ms + logf(Zs)) - C / Zt;
}

}  // namespace

long long head_loss_scratch_bytes(int P) {
  const int m_tiles = (P + HM - 1) / HM;
  return static_cast<long long>(5) * m_tiles * num_sms() * HM * 4;
}

int launch_head_loss(const void* zs, const void* zt, long long ldz, const void* w, const float* center, float tt,
                     float ts, int P, int out_dim, float* loss, void* scratch, long long scratch_bytes,
                     cudaStream_t st) {
  if (P <= 0) return 0;
  if (out_dim % HN != 0 || out_dim <= 0) return set_error("head_loss: out_dim=%d must be a multiple of %d", out_dim, HN);
  if (scratch_bytes < head_loss_scratch_bytes(P)) return set_error("head_loss: scratch too small");
  static PerDevice<bool> attr_cache;
  bool& attr_set = attr_cache.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(head_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM);
    if (e != cudaSuccess) return set_error("head_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tmS, tmT, tmW;
  if (int rc = make_tmap_bf16_2d(&tmS, zs, P, HK, ldz, HM, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmT, zt, P, HK, ldz, HM, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmW, w, out_dim, HK, HK, HN, 64)) return rc;
  const int m_tiles = (P + HM - 1) / HM;
  HeadLossParams p;
  p.P = P; p.out_dim = out_dim; p.center = center; p.inv_tt = 1.0f / tt; p.inv_ts = 1.0f / ts;
  p.part = static_cast<float*>(scratch);
  p.n_tiles = out_dim / HN;
  int gx = num_sms() / m_tiles;
  if (gx < 1) gx = 1;
  if (gx > p.n_tiles) gx = p.n_tiles;
  dim3 grid(gx, m_tiles);
  ProfScope prof(PROF_HEAD_LOSS, static_cast<double>(out_dim) * HK * 2.0 * m_tiles + 2.0 * P * HK * 2.0, st);
  head_loss_kernel<<<grid, H_THREADS, H_SMEM, st>>>(tmS, tmT, tmW, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("head_loss launch: %s", cudaGetErrorString(e));
  count_launch();
  head_loss_merge_kernel<<<(P + 3) / 4, 128, 0, st>>>(p.part, gx, m_tiles, P, loss);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("head_loss merge launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

}  // namespace dvst
