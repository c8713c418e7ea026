// C-ABI layer of libdvst.so. Declarations (with the reference interfaces they replace) are in include/dvst.h.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/dvst.h"
#include "dvst_internal.h"

namespace dvst {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return DVST_ERR;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static int g_pdl = 0;   // measured: neutral on the cfg2 step, -4 % on two-stream whole-video scoring, +11 % single-stream
int pdl_enabled() { return g_pdl; }
void set_pdl(int on) { g_pdl = on != 0 ? 1 : 0; }

struct ProfRec { int kind; double work; cudaEvent_t e0, e1; };
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_event_pool;
static std::mutex g_prof_mu;

static cudaEvent_t get_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
ProfScope::ProfScope(int kind, double work, cudaStream_t stream) : idx(-1), st(stream) {
  if (!g_prof) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{kind, work, get_event(), get_event()};
  cudaEventRecord(r.e0, st);
  g_recs.push_back(r);
  idx = static_cast<int>(g_recs.size()) - 1;
}
ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_recs[idx].e1, st);
}

}  // namespace dvst

using namespace dvst;

extern "C" {

const char* dvst_last_error(void) { return g_err; }
long long dvst_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int dvst_abi_version(void) { return DVST_ABI_VERSION; }
void dvst_set_gemm_cta_group(int cg) { set_gemm_cta_group(cg); }
void dvst_set_pdl(int on) { set_pdl(on); }
void dvst_set_producer_ring(int r) { set_producer_ring(r); }
void dvst_set_attention_impl(int impl) { set_attention_impl(impl); }

void dvst_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof = on != 0;
}

int dvst_profile_collect(double* ms, double* work, long long* count, int nkinds) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int k = 0; k < nkinds; ++k) { ms[k] = 0; work[k] = 0; count[k] = 0; }
  for (auto& r : g_recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess)
      return set_error("dvst_profile_collect: event query failed");
    if (r.kind < nkinds) { ms[r.kind] += t; work[r.kind] += r.work; count[r.kind] += 1; }
    g_event_pool.push_back(r.e0);
    g_event_pool.push_back(r.e1);
  }
  g_recs.clear();
  return 0;
}

int dvst_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias, int M, int N, int K,
                   int split, const float* residual, long long ldr, void* out, long long ldo, int out_kind,
                   long long lo_off, int act, void* stream) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias;
  g.M = M; g.N = N; g.K = K; g.split = split;
  g.residual = residual; g.ldr = ldr; g.out = out; g.ldo = ldo; g.out_kind = out_kind; g.lo_off = lo_off; g.act = act;
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

int dvst_gemm_ln_temporal_attention(const void* xb, long long ldxb, const void* w_perm, const float* bias_perm,
                                    const float* colsum_perm, const float* ln_stats, float ln_eps, long long M, int T,
                                    long long clip_rows, void* out, long long ldo, void* stream) {
  const int rc = launch_gemm_ln_tattn(xb, ldxb, w_perm, bias_perm, colsum_perm, ln_stats, ln_eps, M, T, clip_rows, out, ldo,
                                      static_cast<cudaStream_t>(stream));
  if (rc < 0) return set_error("dvst_gemm_ln_temporal_attention: T = %d is not covered (1 <= T <= 64)", T);
  return rc;
}

int dvst_temporal_tile_geometry(int T, int* rows_per_tile, int* key_block) {
  if (rows_per_tile == nullptr || key_block == nullptr) return set_error("dvst_temporal_tile_geometry: null output");
  if (temporal_tile_geometry(T, rows_per_tile, key_block) != 0)
    return set_error("dvst_temporal_tile_geometry: T = %d is not covered (1 <= T <= 64)", T);
  return 0;
}

int dvst_pair_loss(const float* student, long long lds, const float* teacher, long long ldt, const float* center,
                   float teacher_temp, float student_temp, int dim, int pairs, float* loss, void* stream) {
  if (dim <= 0 || pairs < 0) return set_error("dvst_pair_loss: bad dim %d / pairs %d", dim, pairs);
  if (teacher_temp <= 0.f || student_temp <= 0.f) return set_error("dvst_pair_loss: temperatures must be > 0");
  return launch_pair_loss(student, lds, teacher, ldt, center, teacher_temp, student_temp, dim, pairs, loss, nullptr,
                          nullptr, static_cast<cudaStream_t>(stream));
}

int dvst_pair_loss_indexed(const float* student, long long lds, const int* s_idx, const float* teacher, long long ldt,
                           const int* t_idx, const float* center, float teacher_temp, float student_temp, int dim,
                           int pairs, float* loss, void* stream) {
  if (dim <= 0 || pairs < 0) return set_error("dvst_pair_loss_indexed: bad dim %d / pairs %d", dim, pairs);
  if (teacher_temp <= 0.f || student_temp <= 0.f) return set_error("dvst_pair_loss_indexed: temperatures must be > 0");
  return launch_pair_loss(student, lds, teacher, ldt, center, teacher_temp, student_temp, dim, pairs, loss, s_idx,
                          t_idx, static_cast<cudaStream_t>(stream));
}

long long dvst_head_loss_scratch_bytes(int pairs) { return pairs > 0 ? head_loss_scratch_bytes(pairs) : 0; }

int dvst_head_loss(const void* zs, const void* zt, long long ldz, const void* w, const float* center,
                   float teacher_temp, float student_temp, int pairs, int out_dim, float* loss, void* scratch,
                   long long scratch_bytes, void* stream) {
  if (teacher_temp <= 0.f || student_temp <= 0.f) return set_error("dvst_head_loss: temperatures must be > 0");
  if ((reinterpret_cast<uintptr_t>(zs) & 15) || (reinterpret_cast<uintptr_t>(zt) & 15) ||
      (reinterpret_cast<uintptr_t>(w) & 15) || (ldz % 8))
    return set_error("dvst_head_loss: operands must be 16-byte aligned, ldz a multiple of 8");
  return launch_head_loss(zs, zt, ldz, w, center, teacher_temp, student_temp, pairs, out_dim, loss, scratch,
                          scratch_bytes, static_cast<cudaStream_t>(stream));
}

int dvst_cast_bf16(const float* in, long long ldi, void* out, long long ldo, int split, long long lo_off, long long M,
                   int cols, void* stream) {
  return launch_cast_bf16(in, ldi, out, ldo, split, lo_off, M, cols, static_cast<cudaStream_t>(stream));
}

int dvst_l2norm_rows(const float* in, long long ldi, void* out, long long ldo, int split, long long lo_off,
                     float* out_f32, long long M, int cols, void* stream) {
  return launch_l2norm_rows(in, ldi, out, ldo, split, lo_off, out_f32, M, cols, static_cast<cudaStream_t>(stream));
}

int dvst_colsum(const float* in, long long ldi, float* out, int rows, int cols, void* stream) {
  return launch_colsum(in, ldi, out, rows, cols, static_cast<cudaStream_t>(stream));
}

int dvst_center_ema(float* center, const float* batch_sum, float denom, float momentum, int cols, void* stream) {
  if (denom <= 0.f) return set_error("dvst_center_ema: denom must be > 0");
  return launch_center_ema(center, batch_sum, denom, momentum, cols, static_cast<cudaStream_t>(stream));
}

int dvst_layernorm768(const float* x, long long ldx, const float* gamma, const float* beta, void* out, long long ldo,
                      int out_kind, long long lo_off, long long M, float eps, void* stream) {
  return launch_layernorm768(x, ldx, gamma, beta, out, ldo, out_kind, lo_off, M, eps,
                             static_cast<cudaStream_t>(stream));
}

int dvst_attention(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                   int N, long long cls_base, void* stream) {
  return launch_attention(qkv, ld, out, ldo, cls_out, mode, B, T, N, cls_base, 0, static_cast<cudaStream_t>(stream));
}

int dvst_attention_hp(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                      int N, long long cls_base, void* scratch, long long scratch_bytes, void* stream) {
  if (ld < 6 * 768 || ldo < 2 * 768) return set_error("dvst_attention_hp: rows must hold the [hi | lo] pairs (ld >= 4608, ldo >= 1536)");
  return launch_attention(qkv, ld, out, ldo, cls_out, mode, B, T, N, cls_base, 1, static_cast<cudaStream_t>(stream),
                          scratch, scratch_bytes);
}

long long dvst_attention_scratch_bytes(int B, int T, int N) {
  return (B > 0 && T > 0 && N > 0) ? spatial_attention_scratch_bytes(B, T, N) : 0;
}

int dvst_attention_ws(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                      int N, long long cls_base, void* scratch, long long scratch_bytes, void* stream) {
  return launch_attention(qkv, ld, out, ldo, cls_out, mode, B, T, N, cls_base, 0, static_cast<cudaStream_t>(stream),
                          scratch, scratch_bytes);
}

int dvst_select_frames(const double* loss, long long ld, const int* n_loss, const int* n_unsampled, int n_videos,
                       int pre_sampling_rate, int num_frames, int* indices, void* stream) {
  if (loss == nullptr || n_loss == nullptr || n_unsampled == nullptr || indices == nullptr)
    return set_error("dvst_select_frames: null argument");
  return launch_select_frames(loss, ld, n_loss, n_unsampled, n_videos, pre_sampling_rate, num_frames, indices,
                              static_cast<cudaStream_t>(stream));
}

int dvst_preprocess_u8(const unsigned char* video, int n_frames, int H, int W, int sampling_rate, const float* mean3,
                       const float* std3, int crop, float* frames, void* stream) {
  if (video == nullptr || mean3 == nullptr || std3 == nullptr || frames == nullptr)
    return set_error("dvst_preprocess_u8: null argument");
  return launch_preprocess_u8(video, n_frames, H, W, sampling_rate, mean3, std3, crop, frames,
                              static_cast<cudaStream_t>(stream));
}

int dvst_row_offsets768(const float* stats, float* xoff, long long M, void* stream) {
  if (stats == nullptr || xoff == nullptr || (reinterpret_cast<uintptr_t>(stats) & 15))
    return set_error("dvst_row_offsets768: null or misaligned argument");
  return launch_row_offsets768(stats, xoff, M, static_cast<cudaStream_t>(stream));
}

int dvst_rowstats768(const float* x, long long ldx, void* xb, long long ldxb, float* stats, float* xoff, long long M,
                     void* stream) {
  if (x == nullptr || xb == nullptr || stats == nullptr) return set_error("dvst_rowstats768: null argument");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (ldx % 4) || (reinterpret_cast<uintptr_t>(xb) & 7) || (ldxb % 4) ||
      (reinterpret_cast<uintptr_t>(stats) & 7))
    return set_error("dvst_rowstats768: misaligned argument");
  return launch_rowstats768(x, ldx, xb, ldxb, stats, M, static_cast<cudaStream_t>(stream), xoff);
}

int dvst_gemm_bf16_ln(const void* A, long long lda, const void* W, long long ldw, const float* bias, int M, int N, int K,
                      const float* ln_stats, const float* ln_colsum, float ln_eps, float* x_inout, long long ldx,
                      void* xb, long long ldxb, float* stats_out, const float* xoff, void* out, long long ldo, int act,
                      void* stream) {
  GemmArgs g;
  g.xoff = xoff;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.M = M; g.N = N; g.K = K;
  g.act = act;
  if (x_inout != nullptr) {
    if (ln_stats != nullptr) return set_error("dvst_gemm_bf16_ln: choose the consumer (ln_stats) or the producer (x_inout) form");
    if (N != 768) return set_error("dvst_gemm_bf16_ln: the producer form is defined for N = 768 (6 statistic blocks per row), got %d", N);
    g.residual = x_inout; g.ldr = ldx; g.out = x_inout; g.ldo = ldx; g.out_kind = 3;
    g.xb = xb; g.ldxb = ldxb; g.stats_out = stats_out;
  } else {
    if (ln_stats == nullptr) return set_error("dvst_gemm_bf16_ln: neither ln_stats nor x_inout given");
    g.out = out; g.ldo = ldo; g.out_kind = 0;
    g.ln_stats = ln_stats; g.ln_colsum = ln_colsum; g.ln_eps = ln_eps;
  }
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

int dvst_plan_windows(const int* lengths, int n_videos, int local_clip_size, int global_clip_size, int* s_starts,
                      int* t_starts, int* s_idx, int* t_idx, int* n_s, int* n_t, int* n_groups, int* groups, void* stream) {
  if (lengths == nullptr || n_s == nullptr || n_t == nullptr || n_groups == nullptr || groups == nullptr)
    return set_error("dvst_plan_windows: null host argument");
  if (n_videos > 0 && (s_starts == nullptr || t_starts == nullptr || s_idx == nullptr || t_idx == nullptr))
    return set_error("dvst_plan_windows: null device argument");
  return launch_plan_windows(lengths, n_videos, local_clip_size, global_clip_size, s_starts, t_starts, s_idx, t_idx, n_s,
                             n_t, n_groups, groups, static_cast<cudaStream_t>(stream));
}

int dvst_knn_similarity(const float* test, long long ldt, const float* train, long long ldr, int n_test, int n_train,
                        int dim, float* sim, long long lds, void* stream) {
  if (test == nullptr || train == nullptr || sim == nullptr) return set_error("dvst_knn_similarity: null argument");
  return launch_knn_similarity(test, ldt, train, ldr, n_test, n_train, dim, sim, lds, static_cast<cudaStream_t>(stream));
}

int dvst_knn_topk(const float* sim, long long lds, int n_test, int n_train, int k, float* dist, int* idx, int ldk,
                  void* stream) {
  if (sim == nullptr || dist == nullptr || idx == nullptr) return set_error("dvst_knn_topk: null argument");
  return launch_knn_topk(sim, lds, n_test, n_train, k, dist, idx, ldk, static_cast<cudaStream_t>(stream));
}

int dvst_knn_vote(const float* dist, const int* idx, int ldk, const long long* train_labels,
                  const long long* test_labels, int n_test, int k, float temperature, int num_classes, int* rank,
                  int* pred, void* stream) {
  if (dist == nullptr || idx == nullptr || train_labels == nullptr || test_labels == nullptr || rank == nullptr)
    return set_error("dvst_knn_vote: null argument");
  return launch_knn_vote(dist, idx, ldk, train_labels, test_labels, n_test, k, temperature, num_classes, rank, pred,
                         static_cast<cudaStream_t>(stream));
}

}  // extern "C"
