// HBM-bound kernels of the path: LayerNorm (fp32 residual stream -> bf16 GEMM operand), im2col for the patch
// embedding, CLS-row handling, and the mode-R DINO loss.  All use one warp per 768-wide row, 128-bit coalesced
// accesses and warp-shuffle reductions.
#include "dvst_internal.h"
#include "ptx.cuh"

namespace dvst {

namespace {

constexpr int D = 768;
constexpr int VEC_PER_LANE = D / (32 * 4);  // 6 float4 per lane

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

// LayerNorm over rows of 768 fp32 (eps inside the sqrt, biased variance: torch.nn.LayerNorm, timesformer.py:596).
// out_kind 0: bf16 [M, ldo]; 1: fp32 [M, ldo]; 2: bf16 hi at col, lo at col + lo_off.
__global__ void __launch_bounds__(256) layernorm768_kernel(const float* __restrict__ x, long long ldx,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, void* __restrict__ out,
                                                           long long ldo, int out_kind, long long lo_off,
                                                           long long M, float eps, long long rm_group,
                                                           long long rm_stride, long long rm_offset) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  // optional output row remap: orow = (row / group) * stride + offset + row % group
  const long long orow = rm_group > 0 ? (row / rm_group) * rm_stride + rm_offset + row % rm_group : row;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[VEC_PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    float4 y;
    y.x = v[i].x * rstd * g.x + b.x;
    y.y = v[i].y * rstd * g.y + b.y;
    y.z = v[i].z * rstd * g.z + b.z;
    y.w = v[i].w * rstd * g.w + b.w;
    const int col = (lane + 32 * i) * 4;
    if (out_kind == 1) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + orow * ldo + col) = y;
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + orow * ldo + col;
      *reinterpret_cast<uint2*>(o) = make_uint2(pack2(y.x, y.y), pack2(y.z, y.w));
      if (out_kind == 2) {
        *reinterpret_cast<uint2*>(o + lo_off) =
            make_uint2(pack2(y.x - bf16_round(y.x), y.y - bf16_round(y.y)),
                       pack2(y.z - bf16_round(y.z), y.w - bf16_round(y.w)));
      }
    }
  }
}

// Deferred-LayerNorm inputs for rows not produced by a residual GEMM: xb = bf16(x), per-128-column (sum, sumsq).
__global__ void __launch_bounds__(256) rowstats768_kernel(const float* __restrict__ x, long long ldx,
                                                          __nv_bfloat16* __restrict__ xb, long long ldxb,
                                                          float* __restrict__ stats, long long M,
                                                          float* __restrict__ xoff, long long lo_off) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[6];
  float tot = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    v[j] = xr[j * 32 + lane];
    tot += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float m = xoff != nullptr ? warp_sum(tot) * (1.0f / D) : 0.f;
  if (xoff != nullptr && lane == 0) xoff[row] = m;
#pragma unroll
  for (int j = 0; j < 6; ++j) {  // one partial per 128 columns = one float4 per lane, of the centred values
    const float4 c = make_float4(v[j].x - m, v[j].y - m, v[j].z - m, v[j].w - m);
    float s1 = (c.x + c.y) + (c.z + c.w);
    float s2 = (c.x * c.x + c.y * c.y) + (c.z * c.z + c.w * c.w);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) *reinterpret_cast<float2*>(stats + (row * 6 + j) * 2) = make_float2(s1, s2);
    *reinterpret_cast<uint2*>(xb + row * ldxb + (j * 32 + lane) * 4) = make_uint2(pack2(c.x, c.y), pack2(c.z, c.w));
    if (lo_off > 0)  // split precision: xb = [hi | lo]
      *reinterpret_cast<uint2*>(xb + row * ldxb + lo_off + (j * 32 + lane) * 4) =
          make_uint2(pack2(c.x - bf16_round(c.x), c.y - bf16_round(c.y)), pack2(c.z - bf16_round(c.z), c.w - bf16_round(c.w)));
  }
}

__global__ void __launch_bounds__(256) row_offsets768_kernel(const float* __restrict__ stats, float* __restrict__ xoff,
                                                             long long M) {
  pdl_wait();
  pdl_launch_dependents();
  const long long row = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (row >= M) return;
  const float4* sp = reinterpret_cast<const float4*>(stats) + row * 3;
  const float4 a0 = __ldg(sp), a1 = __ldg(sp + 1), a2 = __ldg(sp + 2);
  xoff[row] += (((a0.x + a0.z) + (a1.x + a1.z)) + (a2.x + a2.z)) * (1.0f / D);  // old offset + mean of the centred row
}

// im2col for Conv2d(k=16, s=16) (PatchEmbed, timesformer.py:201-208): x [B,3,T,H,W] fp32 -> rows (b, n, t) in the
// resident token order, 768 columns in (c, ky, kx) order = the flattened conv weight's K order.
// One thread moves 8 consecutive kx (two float4 loads, one 16-byte bf16 store).
__global__ void __launch_bounds__(256) im2col16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                       long long ldo, int split, long long lo_off, int B, int T,
                                                       int H, int W, int gh, int gw,
                                                       const int* __restrict__ starts) {
  const long long total = static_cast<long long>(B) * gh * gw * T * 96;  // 96 chunks of 8 per row
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % 96);
  const long long row = idx / 96;
  const int t = static_cast<int>(row % T);
  const long long bn = row / T;
  const int n = static_cast<int>(bn % (gh * gw));
  const int b = static_cast<int>(bn / (gh * gw));
  const int c = chunk / 32, rem = chunk % 32, ky = rem / 2, kx0 = (rem % 2) * 8;
  const int py = n / gw, px = n % gw;
  // starts == null: x is a clip batch [B,3,T,H,W]; else x is a frame-major video [F,3,H,W] and clip b is the
  // window of T consecutive frames beginning at starts[b] (device-side view construction, dino_loss_loader.py:126-171)
  const float* src = starts == nullptr
      ? x + (((static_cast<long long>(b) * 3 + c) * T + t) * H + (py * 16 + ky)) * W + px * 16 + kx0
      : x + ((static_cast<long long>(__ldg(starts + b) + t) * 3 + c) * H + (py * 16 + ky)) * W + px * 16 + kx0;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(src));
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
  __nv_bfloat16* o = out + row * ldo + chunk * 8;
  *reinterpret_cast<uint4*>(o) = make_uint4(pack2(a0.x, a0.y), pack2(a0.z, a0.w), pack2(a1.x, a1.y), pack2(a1.z, a1.w));
  if (split) {
    *reinterpret_cast<uint4*>(o + lo_off) =
        make_uint4(pack2(a0.x - bf16_round(a0.x), a0.y - bf16_round(a0.y)),
                   pack2(a0.z - bf16_round(a0.z), a0.w - bf16_round(a0.w)),
                   pack2(a1.x - bf16_round(a1.x), a1.y - bf16_round(a1.y)),
                   pack2(a1.z - bf16_round(a1.z), a1.w - bf16_round(a1.w)));
  }
}

// The same patch rows straight from DECODED frames: video uint8 [n_frames, Hf, Wf, 3]; clip b = the T sampled frames
// (every rate-th decoded frame) starting at sampled index starts[b]; pixel -> / 255, - mean, / std with the loader's three
// separately rounded fp32 operations (data_utils.py:308-325, = preprocess_u8_kernel), centre crop at (y0, x0)
// (transform.py:228-229).  Replaces preprocess_u8_kernel + im2col16_kernel when every frame is used once (clips cut from a
// stream of frames): the normalised fp32 frames (4 x the uint8 bytes, written and read back) never exist.
struct U8Norm { float m[3], s[3]; };
__global__ void __launch_bounds__(256) im2col16_u8_kernel(const uint8_t* __restrict__ video, __nv_bfloat16* __restrict__ out,
                                                          long long ldo, int split, long long lo_off, int B, int T, int Hf,
                                                          int Wf, int rate, int y0, int x0, int gh, int gw,
                                                          const int* __restrict__ starts, const U8Norm nrm) {
  const long long total = static_cast<long long>(B) * gh * gw * T * 96;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % 96);
  const long long row = idx / 96;
  const int t = static_cast<int>(row % T);
  const long long bn = row / T;
  const int n = static_cast<int>(bn % (gh * gw));
  const int b = static_cast<int>(bn / (gh * gw));
  const int c = chunk / 32, rem = chunk % 32, ky = rem / 2, kx0 = (rem % 2) * 8;
  const int py = n / gw, px = n % gw;
  const long long f = static_cast<long long>(__ldg(starts + b) + t) * rate;
  const uint8_t* src = video + ((f * Hf + (y0 + py * 16 + ky)) * Wf + (x0 + px * 16 + kx0)) * 3 + c;
  const float mean = nrm.m[c], sd = nrm.s[c];
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    v[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(__ldg(src + 3 * i)), 255.0f), mean), sd);
  __nv_bfloat16* o = out + row * ldo + chunk * 8;
  *reinterpret_cast<uint4*>(o) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
  if (split) {
    *reinterpret_cast<uint4*>(o + lo_off) =
        make_uint4(pack2(v[0] - bf16_round(v[0]), v[1] - bf16_round(v[1])), pack2(v[2] - bf16_round(v[2]), v[3] - bf16_round(v[3])),
                   pack2(v[4] - bf16_round(v[4]), v[5] - bf16_round(v[5])), pack2(v[6] - bf16_round(v[6]), v[7] - bf16_round(v[7])));
  }
}

// emb[n*T + t, :] = pos_embed[1 + pi(n)] + time_embed[tau(t)]: the additive embedding of every patch token
// (timesformer.py:288-323), with the reference's nearest-neighbour resizes (F.interpolate(mode='nearest'), :292-303
// and :314-318) evaluated exactly like ATen: src = min(floor(dst * float(src_len)/float(dst_len)), src_len - 1).
__global__ void embed_table_kernel(const float* __restrict__ pos, int P, const float* __restrict__ time_embed, int T0,
                                   float* __restrict__ out, int gh_eff, int gw, int N, int T) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(N) * T * (D / 4)) return;
  const int c = static_cast<int>(idx % (D / 4));
  const long long row = idx / (D / 4);
  const int t = static_cast<int>(row % T), n = static_cast<int>(row / T);
  int src_n = n;
  if (N != P * P) {
    const int py = n / gw, px = n % gw;
    const float sy = static_cast<float>(P) / static_cast<float>(gh_eff), sx = static_cast<float>(P) / static_cast<float>(gw);
    const int iy = min(static_cast<int>(floorf(py * sy)), P - 1), ix = min(static_cast<int>(floorf(px * sx)), P - 1);
    src_n = iy * P + ix;
  }
  float4 v = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(1 + src_n) * D) + c);
  if (time_embed != nullptr) {
    int st = t;
    if (T != T0) st = min(static_cast<int>(floorf(t * (static_cast<float>(T0) / static_cast<float>(T)))), T0 - 1);
    const float4 w = __ldg(reinterpret_cast<const float4*>(time_embed + static_cast<long long>(st) * D) + c);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  reinterpret_cast<float4*>(out)[idx] = v;
}

// dst[r, :] = table[r % period, :]: replicate the [N*T, 768] pos+time table over the batch (the patch-embed GEMM then
// accumulates onto it with its TMA reduce-add epilogue).
__global__ void tile_rows768_kernel(const float* __restrict__ table, long long period, float* __restrict__ dst,
                                    long long rows) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * (D / 4)) return;
  const long long r = idx / (D / 4);
  const int c = static_cast<int>(idx % (D / 4));
  reinterpret_cast<float4*>(dst)[idx] = __ldg(reinterpret_cast<const float4*>(table + (r % period) * D) + c);
}

// rows[r, :] = src[:] for r in [0, nrows): broadcast of cls_token + pos_embed[0] into the CLS rows (timesformer.py:310).
__global__ void broadcast_row768_kernel(const float* __restrict__ src, float* __restrict__ dst, long long ldd,
                                        int nrows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrows * (D / 4)) return;
  const int r = idx / (D / 4), c = idx % (D / 4);
  reinterpret_cast<float4*>(dst + r * ldd)[c] = __ldg(reinterpret_cast<const float4*>(src) + c);
}

// out[b, :] = mean_t in[b, t, :]  (CLS residual averaged over frames, timesformer.py:162-164; the projection is
// linear so the mean is taken BEFORE attn.proj).  Output bf16 (+ lo).
__global__ void cls_mean_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long ldo, int split,
                                long long lo_off, long long B, int T, float* __restrict__ out_f32) {
  pdl_wait();
  pdl_launch_dependents();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const long long b = idx / D;
  const int c = static_cast<int>(idx % D);
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += in[(b * T + t) * D + c];
  s /= static_cast<float>(T);
  if (out_f32 != nullptr) { out_f32[b * D + c] = s; return; }
  const __nv_bfloat16 h = __float2bfloat16_rn(s);
  out[b * ldo + c] = h;
  if (split) out[b * ldo + lo_off + c] = __float2bfloat16_rn(s - __bfloat162float(h));
}

// Scoring DINOLoss (dino_similarity.py:129-135) for `pairs` (student, teacher) rows of `dim` fp32 values:
//   loss = logsumexp(s/ts) - sum_j softmax((t - c)/tt)_j * s_j/ts.        One warp per pair.
__global__ void __launch_bounds__(128) pair_loss_kernel(const float* __restrict__ s, long long lds,
                                                        const float* __restrict__ t, long long ldt,
                                                        const float* __restrict__ center, float inv_tt, float inv_ts,
                                                        int dim, int pairs, float* __restrict__ loss,
                                                        const int* __restrict__ s_idx,
                                                        const int* __restrict__ t_idx) {
  const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (p >= pairs) return;
  const int lane = threadIdx.x & 31;
  const float* sr = s + static_cast<long long>(s_idx ? s_idx[p] : p) * lds;
  const float* tr = t + static_cast<long long>(t_idx ? t_idx[p] : p) * ldt;
  float ms = -INFINITY, mt = -INFINITY;
  for (int j = lane; j < dim; j += 32) {
    ms = fmaxf(ms, sr[j] * inv_ts);
    mt = fmaxf(mt, (tr[j] - (center ? center[j] : 0.f)) * inv_tt);
  }
  ms = warp_max(ms);
  mt = warp_max(mt);
  float zs = 0.f, zt = 0.f, cross = 0.f;
  for (int j = lane; j < dim; j += 32) {
    const float sv = sr[j] * inv_ts;
    const float e = expf((tr[j] - (center ? center[j] : 0.f)) * inv_tt - mt);
    zs += expf(sv - ms);
    zt += e;
    cross += e * sv;
  }
  zs = warp_sum(zs);
  zt = warp_sum(zt);
  cross = warp_sum(cross);
  if (lane == 0) loss[p] = (ms + logf(zs)) - cross / zt;
}

// fp32 [M, cols] -> bf16 (+ lo)
__global__ void cast_bf16_kernel(const float* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out,
                                 long long ldo, int split, long long lo_off, long long M, int cols) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * cols) return;
  const long long r = idx / cols;
  const int c = static_cast<int>(idx % cols);
  const float v = in[r * ldi + c];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  out[r * ldo + c] = h;
  if (split) out[r * ldo + lo_off + c] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// F.normalize(x, dim=-1, p=2) (eps 1e-12) over rows of `cols` fp32 (vision_transformer.py:280), bf16 (+ lo) output;
// optionally also the fp32 normalised rows.  One warp per row.
__global__ void __launch_bounds__(128) l2norm_rows_kernel(const float* __restrict__ in, long long ldi,
                                                          __nv_bfloat16* __restrict__ out, long long ldo, int split,
                                                          long long lo_off, float* __restrict__ out_f32, long long M,
                                                          int cols) {
  const long long r = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (r >= M) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.f;
  for (int c = lane; c < cols; c += 32) { const float v = in[r * ldi + c]; ss += v * v; }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  for (int c = lane; c < cols; c += 32) {
    const float v = in[r * ldi + c] * inv;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[r * ldo + c] = h;
    if (split) out[r * ldo + lo_off + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    if (out_f32) out_f32[r * cols + c] = v;
  }
}

// out[c] = sum_r in[r, c]   (batch_center = torch.sum(teacher_output, dim=0), train_ssl.py:686)
__global__ void colsum_kernel(const float* __restrict__ in, long long ldi, float* __restrict__ out, int rows, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += in[static_cast<long long>(r) * ldi + c];
  out[c] = s;
}

// center = center * momentum + (batch_sum / denom) * (1 - momentum)   (train_ssl.py:688-691)
__global__ void center_ema_kernel(float* __restrict__ center, const float* __restrict__ batch_sum, float inv_denom,
                                  float momentum, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  center[c] = center[c] * momentum + batch_sum[c] * inv_denom * (1.0f - momentum);
}

}  // namespace

int launch_cast_bf16(const float* in, long long ldi, void* out, long long ldo, int split, long long lo_off, long long M,
                     int cols, cudaStream_t st) {
  if (M <= 0) return 0;
  const long long total = M * cols;
  cast_bf16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      in, ldi, reinterpret_cast<__nv_bfloat16*>(out), ldo, split, lo_off, M, cols);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("cast launch failed");
}

int launch_l2norm_rows(const float* in, long long ldi, void* out, long long ldo, int split, long long lo_off,
                       float* out_f32, long long M, int cols, cudaStream_t st) {
  if (M <= 0) return 0;
  l2norm_rows_kernel<<<static_cast<unsigned>((M + 3) / 4), 128, 0, st>>>(
      in, ldi, reinterpret_cast<__nv_bfloat16*>(out), ldo, split, lo_off, out_f32, M, cols);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("l2norm launch failed");
}

int launch_colsum(const float* in, long long ldi, float* out, int rows, int cols, cudaStream_t st) {
  if (cols <= 0) return 0;
  colsum_kernel<<<(cols + 255) / 256, 256, 0, st>>>(in, ldi, out, rows, cols);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("colsum launch failed");
}

int launch_center_ema(float* center, const float* batch_sum, float denom, float momentum, int cols, cudaStream_t st) {
  if (cols <= 0) return 0;
  center_ema_kernel<<<(cols + 255) / 256, 256, 0, st>>>(center, batch_sum, 1.0f / denom, momentum, cols);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("center_ema launch failed");
}

int launch_layernorm768(const float* x, long long ldx, const float* gamma, const float* beta, void* out,
                        long long ldo, int out_kind, long long lo_off, long long M, float eps, cudaStream_t st,
                        long long rm_group, long long rm_stride, long long rm_offset) {
  if (M <= 0) return 0;
  const unsigned grid = static_cast<unsigned>((M + 7) / 8);
  ProfScope prof(PROF_LAYERNORM, static_cast<double>(M) * D * (4.0 + (out_kind == 1 ? 4.0 : out_kind == 2 ? 4.0 : 2.0)), st);
  layernorm768_kernel<<<grid, 256, 0, st>>>(x, ldx, gamma, beta, out, ldo, out_kind, lo_off, M, eps, rm_group,
                                            rm_stride, rm_offset);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("layernorm launch failed");
}

int launch_rowstats768(const float* x, long long ldx, void* xb, long long ldxb, float* stats, long long M,
                       cudaStream_t st, float* xoff, long long lo_off) {
  if (M <= 0) return 0;
  ProfScope prof(PROF_LAYERNORM, static_cast<double>(M) * D * 6.0, st);
  rowstats768_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, st>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(xb), ldxb,
                                                                        stats, M, xoff, lo_off);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("rowstats launch failed");
}

int launch_row_offsets768(const float* stats, float* xoff, long long M, cudaStream_t st) {
  if (M <= 0) return 0;
  ProfScope prof(PROF_LAYERNORM, static_cast<double>(M) * 52.0, st);
  launch_kernel_pdl(row_offsets768_kernel, dim3(static_cast<unsigned>((M + 255) / 256)), dim3(256), 0, st, stats, xoff, M);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("row_offsets launch failed");
}

int launch_im2col16(const float* x, void* out, long long ldo, int split, long long lo_off, int B, int T, int H, int W,
                    const int* starts, cudaStream_t st) {
  if (H % 16 || W % 16) return set_error("im2col: H, W must be multiples of 16 (got %d x %d)", H, W);
  if (W % 4) return set_error("im2col: W must be a multiple of 4");
  const int gh = H / 16, gw = W / 16;
  const long long total = static_cast<long long>(B) * gh * gw * T * 96;
  if (total == 0) return 0;
  ProfScope prof(PROF_ELEMENTWISE, static_cast<double>(total) * 8 * (4.0 + (split ? 4.0 : 2.0)), st);
  im2col16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      x, reinterpret_cast<__nv_bfloat16*>(out), ldo, split, lo_off, B, T, H, W, gh, gw, starts);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("im2col launch failed");
}

int launch_im2col16_u8(const uint8_t* video, int Hf, int Wf, int rate, const float* mean3, const float* std3, int crop,
                       void* out, long long ldo, int split, long long lo_off, int B, int T, const int* starts,
                       cudaStream_t st) {
  if (crop % 16 || crop <= 0 || Hf < crop || Wf < crop || rate <= 0)
    return set_error("im2col_u8: bad crop %d / rate %d for %dx%d frames", crop, rate, Hf, Wf);
  const int gh = crop / 16, gw = crop / 16;
  const long long total = static_cast<long long>(B) * gh * gw * T * 96;
  if (total == 0) return 0;
  U8Norm nrm;
  for (int i = 0; i < 3; ++i) { nrm.m[i] = mean3[i]; nrm.s[i] = std3[i]; }
  // ceil((H - crop) / 2) like transform.py:228-229 (math.ceil of a float division), as in launch_preprocess_u8
  const int y0 = (Hf - crop + 1) / 2, x0 = (Wf - crop + 1) / 2;
  ProfScope prof(PROF_ELEMENTWISE, static_cast<double>(total) * 8 * (1.0 + (split ? 4.0 : 2.0)), st);
  im2col16_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      video, reinterpret_cast<__nv_bfloat16*>(out), ldo, split, lo_off, B, T, Hf, Wf, rate, y0, x0, gh, gw, starts, nrm);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("im2col_u8 launch failed");
}

int launch_embed_table(const float* pos, int P, const float* time_embed, int T0, float* out, int gh, int gw, int T,
                       cudaStream_t st) {
  const int N = gh * gw;
  const int gh_eff = (N + 1) / gw;  // the reference derives the target grid height as x.size(1) // W (:297)
  const long long total = static_cast<long long>(N) * T * (D / 4);
  if (total <= 0) return 0;
  embed_table_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(pos, P, time_embed, T0, out, gh_eff, gw,
                                                                               N, T);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("embed_table launch failed");
}

int launch_tile_rows768(const float* table, long long period, float* dst, long long rows, cudaStream_t st) {
  if (rows <= 0) return 0;
  const long long total = rows * (D / 4);
  ProfScope prof(PROF_ELEMENTWISE, static_cast<double>(rows) * D * 4.0, st);
  tile_rows768_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(table, period, dst, rows);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("tile_rows launch failed");
}

int launch_broadcast_row768(const float* src, float* dst, long long ldd, int nrows, cudaStream_t st) {
  if (nrows <= 0) return 0;
  const int total = nrows * (D / 4);
  broadcast_row768_kernel<<<(total + 255) / 256, 256, 0, st>>>(src, dst, ldd, nrows);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("broadcast_row launch failed");
}

int launch_cls_mean(const float* in, void* out, long long ldo, int split, long long lo_off, long long B, int T,
                    float* out_f32, cudaStream_t st) {
  if (B <= 0) return 0;
  launch_kernel_pdl(cls_mean_kernel, dim3(static_cast<unsigned>((B * D + 255) / 256)), dim3(256), 0, st, in,
                    reinterpret_cast<__nv_bfloat16*>(out), ldo, split, lo_off, B, T, out_f32);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("cls_mean launch failed");
}

int launch_pair_loss(const float* s, long long lds, const float* t, long long ldt, const float* center, float tt,
                     float ts, int dim, int pairs, float* loss, const int* s_idx, const int* t_idx, cudaStream_t st) {
  if (pairs <= 0) return 0;
  pair_loss_kernel<<<(pairs + 3) / 4, 128, 0, st>>>(s, lds, t, ldt, center, 1.0f / tt, 1.0f / ts, dim, pairs, loss,
                                                    s_idx, t_idx);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : set_error("pair_loss launch failed");
}

}  // namespace dvst
