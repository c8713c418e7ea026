// Forward orchestration of the TimeSformer backbone on the resident token layout.
//
// Layout in HBM (all inside the caller-provided workspace):
//   x    fp32 [M, 768]   residual stream.  Rows 0..M_t-1 are the patch tokens in (b, n, t) order (n-major, t-minor -
//                        the reference's `b (n t) m`, models/timesformer.py:322) and rows M_t.. are the CLS rows
//                        (one per clip; one per (clip, frame) for space_only), so every row-wise op (LayerNorm, the
//                        qkv / proj / MLP GEMMs) covers patch and CLS tokens in ONE launch and no cat / repeat /
//                        rearrange copy of the reference (:127-181) is ever materialised.
//   a    bf16 [M, 768*k] GEMM A operand (LayerNorm output, im2col rows, attention output); k = 2 in split precision.
//   qkv  bf16 [M, 2304*k]
//   hid  bf16 [M, 3072*k] MLP hidden
//   cls_acc fp32 [B, T, 768] per-frame CLS attention outputs (averaged over T before attn.proj, which is linear)
//   emb  fp32 [N*T, 768] additive pos+time embedding table (periodic residual of the patch-embed GEMM)
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dvst.h"
#include "dvst_internal.h"

namespace dvst {

namespace {
constexpr int D = 768;
constexpr float LN_EPS = 1e-6f;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

struct BlockW {
  const float *tnorm_w = nullptr, *tnorm_b = nullptr, *norm1_w = nullptr, *norm1_b = nullptr, *norm2_w = nullptr,
              *norm2_b = nullptr;
  const void *tqkv_w = nullptr, *tproj_w = nullptr, *qkv_w = nullptr, *proj_w = nullptr, *fc1_w = nullptr,
             *fc2_w = nullptr;
  const float *tqkv_b = nullptr, *tproj_b = nullptr, *qkv_b = nullptr, *proj_b = nullptr, *fc1_b = nullptr,
              *fc2_b = nullptr;
  // deferred LayerNorm (optional): W' = bf16(gamma * W) (split precision: the [hi | lo] pair of gamma * W), colsum(W'),
  // b + W beta for the three Linears that consume a LayerNorm output (temporal qkv <- temporal_norm1, qkv <- norm1,
  // fc1 <- norm2)
  const void *tqkv_wf = nullptr, *qkv_wf = nullptr, *fc1_wf = nullptr;
  const float *tqkv_sf = nullptr, *tqkv_bf = nullptr, *qkv_sf = nullptr, *qkv_bf = nullptr, *fc1_sf = nullptr,
              *fc1_bf = nullptr;
  // head-permuted copies of the folded temporal qkv tensors (optional): the fused qkv GEMM + temporal attention kernel
  const void* tqkv_wfp = nullptr;
  const float *tqkv_sfp = nullptr, *tqkv_bfp = nullptr;
};

struct Model {
  dvst_config cfg;
  std::unordered_map<std::string, const void*> named;
  bool finalized = false;
  bool fused_ln = false;  // every block carries the folded tensors
  bool fused_tattn = false;  // ... and the head-permuted temporal qkv tensors (gemm_tattn_sm100.cu)
  const void* pe_w = nullptr;
  const float *pe_b = nullptr, *cls_pos0 = nullptr, *pos = nullptr, *time_embed = nullptr, *norm_w = nullptr,
              *norm_b = nullptr;
  std::vector<BlockW> blocks;
};

struct Workspace {
  float* x;
  void* a;
  void* qkv;
  void* hid;
  float* cls_acc;
  float* emb;
  float* tmp;  // [rows, 768] fp32 scratch for space_only's frame mean
  void* xb;     // bf16 [M, 768 k] copy of x (deferred LayerNorm: the A operand of the consuming GEMM; k = 2: [hi | lo])
  float* stats; // fp32 [M, 6, 2] per-128-column (sum, sumsq) of every row of x
  float* xoff;  // fp32 [M] per-row offsets of xb (xb = bf16(x - xoff)), refreshed once per block
  void* attn_scratch;  // fix-up list of the one-pass spatial softmax
  size_t attn_scratch_bytes;
  size_t total;
};

static Workspace carve(const Model& m, void* base, long long B, int T, int gh, int gw) {
  const long long N = static_cast<long long>(gh) * gw;
  const long long n_cls = m.cfg.attention_type == DVST_ATTN_SPACE_ONLY ? B * T : B;
  const long long M = B * N * T + n_cls;
  const size_t k = m.cfg.precision == DVST_PRECISION_SPLIT ? 2 : 1;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  Workspace w;
  w.x = static_cast<float*>(take(static_cast<size_t>(M) * D * 4));
  w.a = take(static_cast<size_t>(M) * D * 2 * k);
  w.qkv = take(static_cast<size_t>(M) * 3 * D * 2 * k);
  w.hid = take(static_cast<size_t>(M) * 4 * D * 2 * k);
  w.cls_acc = static_cast<float*>(take(static_cast<size_t>(B) * T * D * 4));
  w.emb = static_cast<float*>(take(static_cast<size_t>(N) * T * D * 4));
  w.tmp = static_cast<float*>(take(m.cfg.attention_type == DVST_ATTN_SPACE_ONLY
                                       ? static_cast<size_t>(B) * (N + 1) * D * 4 : 1024));
  const bool fl = m.fused_ln;
  w.xb = take(fl ? static_cast<size_t>(M) * D * 2 * k : 1024);
  w.stats = static_cast<float*>(take(fl ? static_cast<size_t>(M) * 12 * 4 : 1024));
  w.xoff = static_cast<float*>(take(fl ? static_cast<size_t>(M) * 4 : 1024));
  w.attn_scratch_bytes = m.cfg.attention_type == DVST_ATTN_JOINT
                             ? 0 : static_cast<size_t>(spatial_attention_scratch_bytes(static_cast<int>(B), T, static_cast<int>(N)));
  w.attn_scratch = take(w.attn_scratch_bytes ? w.attn_scratch_bytes : 1024);
  w.total = off;
  return w;
}

}  // namespace dvst

using namespace dvst;

extern "C" {

int dvst_create(const dvst_config* cfg, dvst_model** out) {
  if (cfg == nullptr || out == nullptr) return set_error("dvst_create: null argument");
  if (cfg->depth <= 0 || cfg->depth > 64) return set_error("dvst_create: bad depth %d", cfg->depth);
  if (cfg->attention_type < 0 || cfg->attention_type > 2) return set_error("dvst_create: bad attention_type");
  if (cfg->precision < 0 || cfg->precision > 1) return set_error("dvst_create: bad precision");
  if (cfg->pos_grid <= 0 || cfg->num_frames <= 0) return set_error("dvst_create: bad pos_grid / num_frames");
  Model* m = new Model();
  m->cfg = *cfg;
  *out = reinterpret_cast<dvst_model*>(m);
  return 0;
}

void dvst_destroy(dvst_model* h) { delete reinterpret_cast<Model*>(h); }

int dvst_set_weight(dvst_model* h, const char* name, const void* dev_ptr) {
  if (h == nullptr || name == nullptr || dev_ptr == nullptr) return set_error("dvst_set_weight: null argument");
  if (reinterpret_cast<uintptr_t>(dev_ptr) & 15) return set_error("dvst_set_weight(%s): pointer not 16-byte aligned", name);
  Model* m = reinterpret_cast<Model*>(h);
  m->named[name] = dev_ptr;
  m->finalized = false;
  return 0;
}

int dvst_finalize_weights(dvst_model* h) {
  if (h == nullptr) return set_error("dvst_finalize_weights: null handle");
  Model* m = reinterpret_cast<Model*>(h);
  std::string missing;
  auto get = [&](const std::string& k) -> const void* {
    auto it = m->named.find(k);
    if (it == m->named.end()) {
      if (missing.size() < 300) missing += k + " ";
      return nullptr;
    }
    return it->second;
  };
  auto getf = [&](const std::string& k) { return static_cast<const float*>(get(k)); };
  const bool divided = m->cfg.attention_type == DVST_ATTN_DIVIDED;
  m->pe_w = get("patch_embed.w");
  m->pe_b = getf("patch_embed.b");
  m->cls_pos0 = getf("cls_pos0");
  m->pos = getf("pos_embed");
  m->time_embed = m->cfg.attention_type == DVST_ATTN_SPACE_ONLY ? nullptr : getf("time_embed");
  m->norm_w = getf("norm.w");
  m->norm_b = getf("norm.b");
  m->blocks.assign(m->cfg.depth, BlockW());
  for (int i = 0; i < m->cfg.depth; ++i) {
    const std::string p = "blocks." + std::to_string(i) + ".";
    BlockW& b = m->blocks[i];
    if (divided) {
      b.tnorm_w = getf(p + "temporal_norm1.w"); b.tnorm_b = getf(p + "temporal_norm1.b");
      b.tqkv_w = get(p + "temporal_attn.qkv.w"); b.tqkv_b = getf(p + "temporal_attn.qkv.b");
      b.tproj_w = get(p + "temporal_projfc.w"); b.tproj_b = getf(p + "temporal_projfc.b");
    }
    b.norm1_w = getf(p + "norm1.w"); b.norm1_b = getf(p + "norm1.b");
    b.qkv_w = get(p + "attn.qkv.w"); b.qkv_b = getf(p + "attn.qkv.b");
    b.proj_w = get(p + "attn.proj.w"); b.proj_b = getf(p + "attn.proj.b");
    b.norm2_w = getf(p + "norm2.w"); b.norm2_b = getf(p + "norm2.b");
    b.fc1_w = get(p + "mlp.fc1.w"); b.fc1_b = getf(p + "mlp.fc1.b");
    b.fc2_w = get(p + "mlp.fc2.w"); b.fc2_b = getf(p + "mlp.fc2.b");
  }
  if (!missing.empty()) return set_error("dvst_finalize_weights: missing tensors: %s", missing.c_str());
  // optional folded tensors of the deferred LayerNorm
  auto opt = [&](const std::string& k) -> const void* {
    auto it = m->named.find(k);
    return it == m->named.end() ? nullptr : it->second;
  };
  bool all = true;
  bool all_perm = divided && m->cfg.precision == DVST_PRECISION_BF16;  // the fused temporal kernel is bf16 only
  for (int i = 0; i < m->cfg.depth; ++i) {
    const std::string p = "blocks." + std::to_string(i) + ".";
    BlockW& b = m->blocks[i];
    if (divided) {
      b.tqkv_wf = opt(p + "temporal_attn.qkv.wf");
      b.tqkv_sf = static_cast<const float*>(opt(p + "temporal_attn.qkv.sf"));
      b.tqkv_bf = static_cast<const float*>(opt(p + "temporal_attn.qkv.bf"));
      all = all && b.tqkv_wf && b.tqkv_sf && b.tqkv_bf;
      b.tqkv_wfp = opt(p + "temporal_attn.qkv.wfp");
      b.tqkv_sfp = static_cast<const float*>(opt(p + "temporal_attn.qkv.sfp"));
      b.tqkv_bfp = static_cast<const float*>(opt(p + "temporal_attn.qkv.bfp"));
      all_perm = all_perm && b.tqkv_wfp && b.tqkv_sfp && b.tqkv_bfp;
    }
    b.qkv_wf = opt(p + "attn.qkv.wf");
    b.qkv_sf = static_cast<const float*>(opt(p + "attn.qkv.sf"));
    b.qkv_bf = static_cast<const float*>(opt(p + "attn.qkv.bf"));
    b.fc1_wf = opt(p + "mlp.fc1.wf");
    b.fc1_sf = static_cast<const float*>(opt(p + "mlp.fc1.sf"));
    b.fc1_bf = static_cast<const float*>(opt(p + "mlp.fc1.bf"));
    all = all && b.qkv_wf && b.qkv_sf && b.qkv_bf && b.fc1_wf && b.fc1_sf && b.fc1_bf;
  }
  m->fused_ln = all;
  m->fused_tattn = all && all_perm;
  m->finalized = true;
  return 0;
}

// Clips per pass: the batch is walked in chunks of about g_chunk_rows token rows so that the activations of one chunk
// (x, a, qkv, hid) stay resident in the 126 MB L2 between producer and consumer kernels; 0 = whole batch in one pass.
static int g_ln_fusion = 1;
void dvst_set_ln_fusion(int on) { g_ln_fusion = on != 0 ? 1 : 0; }
void dvst_set_temporal_fusion(int on) { set_temporal_fusion(on); }
static long long g_chunk_rows = 0;
void dvst_set_chunk_rows(long long rows) { g_chunk_rows = rows < 0 ? 0 : rows; }
static int chunk_clips(int B, long long rows_per_clip) {
  if (g_chunk_rows <= 0) return B;
  long long c = g_chunk_rows / rows_per_clip;
  if (c < 1) c = 1;
  return c < B ? static_cast<int>(c) : B;
}

long long dvst_workspace_bytes(const dvst_model* h, int B, int T, int H, int W) {
  if (h == nullptr || B <= 0 || T <= 0 || H < 16 || W < 16) return -1;
  const Model* m = reinterpret_cast<const Model*>(h);
  const int cb = chunk_clips(B, static_cast<long long>(H / 16) * (W / 16) * T);
  return static_cast<long long>(carve(*m, nullptr, cb, T, H / 16, W / 16).total);
}

// Decoded uint8 frames as the input of the patch embedding (dvst_forward_windows_u8): x_in is then null
struct U8Source {
  const uint8_t* video = nullptr;
  int Hf = 0, Wf = 0, rate = 1;
  float mean[3] = {0, 0, 0}, std[3] = {1, 1, 1};
};

// get_all: 0 = CLS features, 1 = every token, 2 = attention map of the last block's spatial attention (get_attn)
static int forward_impl(dvst_model* h, const float* x_in, const int* starts, int B, int T, int H, int W, float* out,
                        int get_all, void* workspace, long long workspace_bytes, void* stream,
                        const U8Source* u8 = nullptr);

int dvst_forward_features(dvst_model* h, const float* x_in, int B, int T, int H, int W, float* out, int get_all,
                          void* workspace, long long workspace_bytes, void* stream) {
  return forward_impl(h, x_in, nullptr, B, T, H, W, out, get_all, workspace, workspace_bytes, stream);
}

int dvst_forward_attention_map(dvst_model* h, const float* x_in, int B, int T, int H, int W, float* attn,
                               void* workspace, long long workspace_bytes, void* stream) {
  if (h != nullptr && reinterpret_cast<Model*>(h)->cfg.attention_type != DVST_ATTN_DIVIDED)
    return set_error("dvst_forward_attention_map: only divided_space_time blocks return an attention map "
                     "(models/timesformer.py:120-123 ignores return_attn for the other types)");
  return forward_impl(h, x_in, nullptr, B, T, H, W, attn, 2, workspace, workspace_bytes, stream);
}

int dvst_forward_windows(dvst_model* h, const float* frames, int F, const int* starts, int n_clips, int T, int H,
                         int W, float* out, void* workspace, long long workspace_bytes, void* stream) {
  if (starts == nullptr) return set_error("dvst_forward_windows: null starts");
  if (F < T) return set_error("dvst_forward_windows: video has %d frames, window needs %d", F, T);
  return forward_impl(h, frames, starts, n_clips, T, H, W, out, 0, workspace, workspace_bytes, stream);
}

int dvst_forward_windows_u8(dvst_model* h, const unsigned char* video, int n_frames, int Hf, int Wf, int sampling_rate,
                            const float* mean3, const float* std3, int crop, const int* starts, int n_clips, int T,
                            float* out, void* workspace, long long workspace_bytes, void* stream) {
  if (video == nullptr || mean3 == nullptr || std3 == nullptr || starts == nullptr)
    return set_error("dvst_forward_windows_u8: null argument");
  if (sampling_rate <= 0 || crop <= 0 || crop % 16 || Hf < crop || Wf < crop)
    return set_error("dvst_forward_windows_u8: bad sampling rate %d / crop %d for %dx%d frames", sampling_rate, crop, Hf, Wf);
  const int F = (n_frames + sampling_rate - 1) / sampling_rate;
  if (F < T) return set_error("dvst_forward_windows_u8: video has %d sampled frames, window needs %d", F, T);
  U8Source u8;
  u8.video = video; u8.Hf = Hf; u8.Wf = Wf; u8.rate = sampling_rate;
  for (int i = 0; i < 3; ++i) { u8.mean[i] = mean3[i]; u8.std[i] = std3[i]; }
  return forward_impl(h, nullptr, starts, n_clips, T, crop, crop, out, 0, workspace, workspace_bytes, stream, &u8);
}

static int forward_chunk(dvst_model* h, const float* x_in, const int* starts, int B, int T, int H, int W, float* out,
                         int get_all, void* workspace, long long workspace_bytes, void* stream, const U8Source* u8);

static int forward_impl(dvst_model* h, const float* x_in, const int* starts, int B, int T, int H, int W, float* out,
                        int get_all, void* workspace, long long workspace_bytes, void* stream, const U8Source* u8) {
  if (h == nullptr) return set_error("dvst_forward_features: null handle");
  if (B <= 0 || T <= 0 || H < 16 || W < 16) return set_error("dvst_forward_features: bad shape");
  const Model* m = reinterpret_cast<const Model*>(h);
  const long long N = static_cast<long long>(H / 16) * (W / 16);
  const int cb = chunk_clips(B, N * T);
  const long long out_rows = get_all == 1 ? 1 + (m->cfg.attention_type == DVST_ATTN_SPACE_ONLY ? N : N * T) : 1;
  for (int b0 = 0; b0 < B; b0 += cb) {
    const int nb = B - b0 < cb ? B - b0 : cb;
    const float* xin = (starts || u8) ? x_in : x_in + static_cast<long long>(b0) * 3 * T * H * W;
    float* o = get_all == 2 ? out + static_cast<long long>(b0) * T * 12 * (N + 1) * (N + 1) : out + b0 * out_rows * D;
    const int rc = forward_chunk(h, xin, starts ? starts + b0 : nullptr, nb, T, H, W, o, get_all, workspace,
                                 workspace_bytes, stream, u8);
    if (rc != 0) return rc;
  }
  return 0;
}

static int forward_chunk(dvst_model* h, const float* x_in, const int* starts, int B, int T, int H, int W, float* out,
                         int get_all, void* workspace, long long workspace_bytes, void* stream, const U8Source* u8) {
  if (h == nullptr) return set_error("dvst_forward_features: null handle");
  Model* m = reinterpret_cast<Model*>(h);
  if (!m->finalized) return set_error("dvst_forward_features: weights not finalized");
  if (B <= 0 || T <= 0) return set_error("dvst_forward_features: bad batch %d / frames %d", B, T);
  if (H % 16 || W % 16 || H < 16 || W < 16)
    return set_error("dvst_forward_features: H, W must be multiples of the patch size 16 (got %dx%d)", H, W);
  if ((u8 == nullptr && (reinterpret_cast<uintptr_t>(x_in) & 15)) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 1023))
    return set_error("dvst_forward_features: x/out must be 16-byte and workspace 1024-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int gh = H / 16, gw = W / 16;
  const long long N = static_cast<long long>(gh) * gw;
  const int att = m->cfg.attention_type;
  const long long n_cls = att == DVST_ATTN_SPACE_ONLY ? static_cast<long long>(B) * T : B;
  const long long Mt = static_cast<long long>(B) * N * T;
  const long long M = Mt + n_cls;
  if (M > 2000000000LL / 2) return set_error("dvst_forward_features: too many rows (%lld)", M);
  Workspace w = carve(*m, workspace, B, T, gh, gw);
  if (static_cast<long long>(w.total) > workspace_bytes)
    return set_error("dvst_forward_features: workspace too small (%lld < %lld)", workspace_bytes,
                     static_cast<long long>(w.total));
  const int split = m->cfg.precision == DVST_PRECISION_SPLIT ? 1 : 0;
  const int act_kind = split ? 2 : 0;            // bf16 or bf16 hi/lo
  const long long lda = split ? 2 * D : D;       // row stride of `a`
  const long long ldq = split ? 6 * D : 3 * D;   // row stride of `qkv`
  const long long ldh = split ? 8 * D : 4 * D;   // row stride of `hid`
  int rc;
#define DVST_TRY(expr) do { if ((rc = (expr)) != 0) return rc; } while (0)

  auto gemm = [&](const void* A, long long lda_, const void* Wt, const float* bias, long long rows, int Nn, int K,
                  const float* residual, void* o, long long ldo, int out_kind, int act) {
    GemmArgs g;
    g.A = A; g.lda = lda_; g.W = Wt; g.ldw = split ? 2LL * K : K; g.bias = bias;
    g.M = static_cast<int>(rows); g.N = Nn; g.K = K; g.split = split;
    g.residual = residual; g.ldr = D;
    g.out = o; g.ldo = ldo; g.out_kind = out_kind; g.lo_off = ldo / 2; g.act = act;
    return launch_gemm(g, st);
  };
  auto ln = [&](const float* src, const float* gw_, const float* gb_, long long rows) {
    return launch_layernorm768(src, D, gw_, gb_, w.a, lda, act_kind, D, rows, LN_EPS, st);
  };
  // Deferred LayerNorm (folded tensors present): residual GEMMs also emit xb = bf16(x) (split precision: its [hi | lo]
  // pair) and the row statistics, the GEMMs that consume LayerNorm(x) run on xb with gamma folded into their weights
  // (gemm_sm100.cu).
  const bool fl = m->fused_ln && g_ln_fusion;
  __nv_bfloat16* xb = static_cast<__nv_bfloat16*>(w.xb);
  const long long ldxb = split ? 2 * D : D;
  // x[row0 .. row0+rows) += A W^T + bias, plus xb / stats of those rows
  auto gemm_res = [&](const void* A, const void* Wt, const float* bias, long long row0, long long rows, int K) {
    GemmArgs g;
    g.A = A; g.lda = K == D ? lda : ldh; g.W = Wt; g.ldw = split ? 2LL * K : K; g.bias = bias;
    g.M = static_cast<int>(rows); g.N = D; g.K = K; g.split = split;
    g.residual = w.x + row0 * D; g.ldr = D; g.out = w.x + row0 * D; g.ldo = D; g.out_kind = 3;
    g.xb = xb + row0 * ldxb; g.ldxb = ldxb; g.lo_off = D; g.stats_out = w.stats + row0 * 12; g.xoff = w.xoff + row0;
    return launch_gemm(g, st);
  };
  // out[row0 .. row0+rows) = act(LayerNorm(x) W^T + b) through the folded form
  auto gemm_ln = [&](const void* Wf, const float* sf, const float* bf, long long row0, long long rows, int Nn, void* o,
                     long long ldo, int act) {
    GemmArgs g;
    g.A = xb + row0 * ldxb; g.lda = ldxb; g.W = Wf; g.ldw = ldxb; g.bias = bf;
    g.M = static_cast<int>(rows); g.N = Nn; g.K = D; g.split = split;
    g.out = o; g.ldo = ldo; g.out_kind = act_kind; g.lo_off = ldo / 2; g.act = act;
    g.ln_stats = w.stats + row0 * 12; g.ln_colsum = sf; g.ln_eps = LN_EPS;
    return launch_gemm(g, st);
  };

  // ---- patch embedding + positional / temporal embedding (timesformer.py:203-209,287-323) ----
  DVST_TRY(launch_embed_table(m->pos, m->cfg.pos_grid, m->time_embed, m->cfg.num_frames, w.emb, gh, gw, T, st));
  if (u8 != nullptr)
    DVST_TRY(launch_im2col16_u8(u8->video, u8->Hf, u8->Wf, u8->rate, u8->mean, u8->std, H, w.a, lda, split, D, B, T, starts, st));
  else
    DVST_TRY(launch_im2col16(x_in, w.a, lda, split, D, B, T, H, W, starts, st));
  DVST_TRY(launch_tile_rows768(w.emb, N * T, w.x, Mt, st));  // x[b, n, t] = pos + time; the GEMM accumulates onto it
  if (fl) {
    // offsets start at 0 for the patch rows (the embedded tokens are centred well enough); the CLS rows get their exact mean
    if (cudaMemsetAsync(w.xoff, 0, static_cast<size_t>(Mt) * sizeof(float), st) != cudaSuccess)
      return set_error("dvst_forward_features: cudaMemsetAsync failed");
    DVST_TRY(gemm_res(w.a, m->pe_w, m->pe_b, 0, Mt, D));
  } else DVST_TRY(gemm(w.a, lda, m->pe_w, m->pe_b, Mt, D, D, w.x, w.x, D, 1, 0));
  DVST_TRY(launch_broadcast_row768(m->cls_pos0, w.x + Mt * D, D, static_cast<int>(n_cls), st));
  if (fl)
    DVST_TRY(launch_rowstats768(w.x + Mt * D, D, xb + Mt * ldxb, ldxb, w.stats + Mt * 12, n_cls, st, w.xoff + Mt,
                                split ? D : 0));

  for (int i = 0; i < m->cfg.depth; ++i) {
    const BlockW& b = m->blocks[i];
    if (att == DVST_ATTN_DIVIDED) {
      // temporal attention over the T tokens of every (b, n); CLS rows excluded (:127-135)
      // bf16 + deferred LayerNorm, T <= 64: qkv GEMM and temporal attention are ONE kernel, the qkv tensor stays in shared
      // memory (gemm_tattn_sm100.cu).  The choice depends on the model and the clip shape only, never on the batch.
      // T in {1, 2, 4, 8, 16, 32, 64} with clips that start on a 16-row boundary: rows tiled as one matrix - same bits as the
      // two-kernel form.  Any other T (the 4_3_30 windows T = 3, 30): rows tiled clip by clip, so every sequence sits at a
      // place of its MMA tile that depends on its patch index only and features stay bit-identical between batched and
      // one-at-a-time scoring (they differ from the two-kernel form by the summation order inside the tensor core).
      int fused = -1;
      if (fl && m->fused_tattn && temporal_fusion() && T <= 64) {
        const bool t_pow2 = (T & (T - 1)) == 0;
        const long long clip_rows = t_pow2 && (T >= 16 || (N * T) % 16 == 0) ? 0 : N * T;
        fused = launch_gemm_ln_tattn(xb, D, b.tqkv_wfp, b.tqkv_bfp, b.tqkv_sfp, w.stats, LN_EPS, Mt, T, clip_rows, w.a, lda, st);
        if (fused > 0) return fused;
      }
      if (fused != 0) {
        if (fl) {
          DVST_TRY(gemm_ln(b.tqkv_wf, b.tqkv_sf, b.tqkv_bf, 0, Mt, 3 * D, w.qkv, ldq, 0));
        } else {
          DVST_TRY(ln(w.x, b.tnorm_w, b.tnorm_b, Mt));
          DVST_TRY(gemm(w.a, lda, b.tqkv_w, b.tqkv_b, Mt, 3 * D, D, nullptr, w.qkv, ldq, act_kind, 0));
        }
        DVST_TRY(launch_attention(w.qkv, ldq, w.a, lda, nullptr, 0, B, T, static_cast<int>(N), Mt, split, st));
      }
      // temporal_fc(proj(.)) folded into one 768x768 GEMM at pack time; residual add fused (:131-135)
      if (fl) DVST_TRY(gemm_res(w.a, b.tproj_w, b.tproj_b, 0, Mt, D));
      else DVST_TRY(gemm(w.a, lda, b.tproj_w, b.tproj_b, Mt, D, D, w.x, w.x, D, 1, 0));
    }
    // spatial (or joint / per-frame) attention (:120-123,140-181)
    const int mode = att == DVST_ATTN_DIVIDED ? 1 : (att == DVST_ATTN_JOINT ? 2 : 3);
    if (!fl) DVST_TRY(ln(w.x, b.norm1_w, b.norm1_b, M));
    if (i + 1 == m->cfg.depth && get_all == 2) {
      // get_attn: the last block stops after the spatial q / k and returns softmax(q k^T / 8) (:155-157,325-331)
      if (fl) DVST_TRY(gemm_ln(b.qkv_wf, b.qkv_sf, b.qkv_bf, 0, M, 3 * D, w.qkv, ldq, 0));
      else DVST_TRY(gemm(w.a, lda, b.qkv_w, b.qkv_b, M, 3 * D, D, nullptr, w.qkv, ldq, act_kind, 0));
      return launch_attention_probs(w.qkv, ldq, out, B, T, static_cast<int>(N), Mt, split, st);
    }
    if (i + 1 == m->cfg.depth && !get_all) {
      // Last block, CLS features only (forward() keeps norm(x)[:, 0], :342-345): nothing downstream reads the patch
      // tokens' spatial-attention output, attn.proj, norm2 or MLP of this block.
      // Patch rows still provide keys / values: their K | V thirds of the qkv GEMM (N = 1536) plus the full qkv of the
      // CLS rows, then the CLS-query attention, and proj / LayerNorm / MLP on the n_cls CLS rows alone.
      const long long wrow = split ? 2LL * D : D;  // elements per weight row
      __nv_bfloat16* qkv_bf = static_cast<__nv_bfloat16*>(w.qkv);
      __nv_bfloat16* a_bf = static_cast<__nv_bfloat16*>(w.a);
      if (fl) {
        DVST_TRY(gemm_ln(static_cast<const __nv_bfloat16*>(b.qkv_wf) + D * wrow, b.qkv_sf + D, b.qkv_bf + D, 0, Mt, 2 * D,
                         qkv_bf + D, ldq, 0));
        DVST_TRY(gemm_ln(b.qkv_wf, b.qkv_sf, b.qkv_bf, Mt, n_cls, 3 * D, qkv_bf + Mt * ldq, ldq, 0));
      } else {
        DVST_TRY(gemm(w.a, lda, static_cast<const __nv_bfloat16*>(b.qkv_w) + D * wrow, b.qkv_b + D, Mt, 2 * D, D, nullptr,
                      qkv_bf + D, ldq, act_kind, 0));
        DVST_TRY(gemm(a_bf + Mt * lda, lda, b.qkv_w, b.qkv_b, n_cls, 3 * D, D, nullptr, qkv_bf + Mt * ldq, ldq, act_kind, 0));
      }
      DVST_TRY(launch_cls_attention(w.qkv, ldq, w.a, lda, w.cls_acc, mode, B, T, static_cast<int>(N), Mt, split, st));
      if (att == DVST_ATTN_DIVIDED)
        DVST_TRY(launch_cls_mean(w.cls_acc, a_bf + Mt * lda, lda, split, D, B, T, nullptr, st));
      float* x_cls = w.x + Mt * D;
      DVST_TRY(gemm(a_bf + Mt * lda, lda, b.proj_w, b.proj_b, n_cls, D, D, x_cls, x_cls, D, 1, 0));
      DVST_TRY(launch_layernorm768(x_cls, D, b.norm2_w, b.norm2_b, a_bf + Mt * lda, lda, act_kind, D, n_cls, LN_EPS, st));
      DVST_TRY(gemm(a_bf + Mt * lda, lda, b.fc1_w, b.fc1_b, n_cls, 4 * D, D, nullptr, w.hid, ldh, act_kind, 1));
      DVST_TRY(gemm(w.hid, ldh, b.fc2_w, b.fc2_b, n_cls, D, 4 * D, x_cls, x_cls, D, 1, 0));
      continue;
    }
    if (fl) DVST_TRY(gemm_ln(b.qkv_wf, b.qkv_sf, b.qkv_bf, 0, M, 3 * D, w.qkv, ldq, 0));
    else DVST_TRY(gemm(w.a, lda, b.qkv_w, b.qkv_b, M, 3 * D, D, nullptr, w.qkv, ldq, act_kind, 0));
    DVST_TRY(launch_attention(w.qkv, ldq, w.a, lda, w.cls_acc, mode, B, T, static_cast<int>(N), Mt, split, st,
                              w.attn_scratch_bytes ? w.attn_scratch : nullptr, static_cast<long long>(w.attn_scratch_bytes)));
    if (att == DVST_ATTN_DIVIDED)
      DVST_TRY(launch_cls_mean(w.cls_acc, static_cast<__nv_bfloat16*>(w.a) + Mt * lda, lda, split, D, B, T, nullptr, st));
    if (fl) {
      // once per block, right before a producer that rewrites xb and the statistics of every row: offsets <- row means
      DVST_TRY(launch_row_offsets768(w.stats, w.xoff, M, st));
      DVST_TRY(gemm_res(w.a, b.proj_w, b.proj_b, 0, M, D));
      // MLP (:184)
      DVST_TRY(gemm_ln(b.fc1_wf, b.fc1_sf, b.fc1_bf, 0, M, 4 * D, w.hid, ldh, 1));
      DVST_TRY(gemm_res(w.hid, b.fc2_w, b.fc2_b, 0, M, 4 * D));
    } else {
      DVST_TRY(gemm(w.a, lda, b.proj_w, b.proj_b, M, D, D, w.x, w.x, D, 1, 0));
      DVST_TRY(ln(w.x, b.norm2_w, b.norm2_b, M));
      DVST_TRY(gemm(w.a, lda, b.fc1_w, b.fc1_b, M, 4 * D, D, nullptr, w.hid, ldh, act_kind, 1));
      DVST_TRY(gemm(w.hid, ldh, b.fc2_w, b.fc2_b, M, D, 4 * D, w.x, w.x, D, 1, 0));
    }
  }

  // ---- final norm (:338-345) ----
  if (att == DVST_ATTN_SPACE_ONLY) {
    // frame average of every token, then LayerNorm
    if (get_all) {
      DVST_TRY(launch_cls_mean(w.x, nullptr, 0, 0, 0, static_cast<long long>(B) * N, T, w.tmp + static_cast<long long>(B) * D, st));
      DVST_TRY(launch_cls_mean(w.x + Mt * D, nullptr, 0, 0, 0, B, T, w.tmp, st));
      DVST_TRY(launch_layernorm768(w.tmp, D, m->norm_w, m->norm_b, out, D, 1, 0, B, LN_EPS, st, 1, N + 1, 0));
      DVST_TRY(launch_layernorm768(w.tmp + static_cast<long long>(B) * D, D, m->norm_w, m->norm_b, out, D, 1, 0,
                                   static_cast<long long>(B) * N, LN_EPS, st, N, N + 1, 1));
    } else {
      DVST_TRY(launch_cls_mean(w.x + Mt * D, nullptr, 0, 0, 0, B, T, w.tmp, st));
      DVST_TRY(launch_layernorm768(w.tmp, D, m->norm_w, m->norm_b, out, D, 1, 0, B, LN_EPS, st));
    }
  } else if (get_all) {
    // reference order: [B, 1 + N*T, 768], CLS first
    DVST_TRY(launch_layernorm768(w.x + Mt * D, D, m->norm_w, m->norm_b, out, D, 1, 0, B, LN_EPS, st, 1, N * T + 1, 0));
    DVST_TRY(launch_layernorm768(w.x, D, m->norm_w, m->norm_b, out, D, 1, 0, Mt, LN_EPS, st, N * T, N * T + 1, 1));
  } else {
    // only the CLS rows are needed: LayerNorm is row-wise, so this equals norm(x)[:, 0] (:342-345)
    DVST_TRY(launch_layernorm768(w.x + Mt * D, D, m->norm_w, m->norm_b, out, D, 1, 0, B, LN_EPS, st));
  }
#undef DVST_TRY
  return 0;
}

}  // extern "C"
