// Internal (C++) interface between the kernel translation units and the C-ABI layer (api.cu).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace dvst {

// Records a printf-style message retrievable through dvst_last_error(); returns a non-zero error code.
int set_error(const char* fmt, ...);
// Every kernel launch made by the library bumps this counter (read through dvst_launch_count()).
void count_launch();
int num_sms();   // of the CURRENT device
// Per-device caches (kernel attributes such as the dynamic shared-memory opt-in are per device, and so is the SM count):
// indexed by cudaGetDevice(), so a process that drives several GPUs (or moves a model to another device) stays correct.
constexpr int kMaxDevices = 64;
int device_slot();
template <class T> struct PerDevice {
  T v[kMaxDevices] = {};
  T& get() { return v[device_slot()]; }
};

// Programmatic dependent launch for the kernels of the forward's hot path (GEMMs, fused temporal kernel, spatial attention +
// fix-up, cls_mean, row_offsets): 1 launches them with cudaLaunchAttributeProgrammaticStreamSerialization, so a
// kernel's prologue overlaps the tail of its predecessor; each of those kernels executes griddepcontrol.wait before it touches
// global memory.  0 (default) = ordinary launches.
int pdl_enabled();
void set_pdl(int on);
// kernel<<<grid, block, smem, st>>>(args...) with the attribute when enabled
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Optional live per-kernel timing (CUDA events on the launching stream), used by bench.py for the roofline figures.
// PROF_GEMM_HBM: short-K (K <= 1024) deferred-LayerNorm producer GEMMs - bounded by their residual / output streams (work is
// counted in bytes), unlike every other GEMM launch (tensor bound, work in FLOPs).
// PROF_GEMM_TATTN: the fused qkv-GEMM + temporal-attention kernel (gemm_tattn_sm100.cu), work in FLOPs (GEMM + attention).
enum ProfKind { PROF_GEMM = 0, PROF_ATTENTION = 1, PROF_LAYERNORM = 2, PROF_ELEMENTWISE = 3, PROF_HEAD_LOSS = 4,
                PROF_GEMM_HBM = 5, PROF_GEMM_TATTN = 6, PROF_NKINDS = 7 };
struct ProfScope {
  int idx;
  cudaStream_t st;
  ProfScope(int kind, double work, cudaStream_t stream);  // work: FLOPs (GEMM / attention) or bytes (others)
  ~ProfScope();
};

struct GemmArgs {
  const void* A = nullptr;  // bf16 [M, lda]  (split: [hi | lo], lo at column K)
  long long lda = 0;
  const void* W = nullptr;  // bf16 [N, ldw]  (nn.Linear layout; split: [hi | lo])
  long long ldw = 0;
  const float* bias = nullptr;      // [N] or null
  const float* residual = nullptr;  // null, or == out (fp32): out += result (in-place residual add)
  long long ldr = 0;
  void* out = nullptr;
  long long ldo = 0;
  int out_kind = 0;  // 0 bf16, 1 fp32, 2 bf16 hi/lo pair, 3 fp32 in-place residual + bf16 copy `xb` + row stats
  long long lo_off = 0;  // out_kind 2: column offset of the lo half of `out`; out_kind 3 with split: ... of `xb`
  // deferred LayerNorm (gemm_sm100.cu): consumer side (K = 768; out_kind 0, or 2 in split precision) ...
  const float* ln_stats = nullptr;   // fp32 [M, 6, 2] per-128-column (sum, sumsq) of the row A was cast from
  const float* ln_colsum = nullptr;  // fp32 [N] row sums of the gamma-folded bf16 weight
  float ln_eps = 1e-6f;
  // ... and producer side (out_kind 3, N = 768)
  void* xb = nullptr;                // bf16 [M, ldxb] (split: [hi | lo], lo at column lo_off)
  long long ldxb = 0;
  float* stats_out = nullptr;        // fp32 [M, N / 128, 2]
  // producer: per-row offset m_r; xb = bf16(x - m_r) and the statistics are those of x - m_r (null = 0)
  const float* xoff = nullptr;
  int act = 0;  // 0 none, 1 GELU(erf)
  int M = 0, N = 0, K = 0;
  int split = 0;
};
int launch_gemm(const GemmArgs& g, cudaStream_t stream);
void set_gemm_cta_group(int cg);  // 1 or 2 (default 2); tuning / test hook
void set_producer_ring(int r);    // 5 (default) or 3 residual boxes per epilogue warp of the short-K producers; tuning hook

// Temporal half of a divided block in one kernel (gemm_tattn_sm100.cu): deferred-LayerNorm qkv GEMM whose epilogue runs the
// temporal attention out of shared memory; the packed qkv tensor is never written.  w / bias / colsum are HEAD-PERMUTED
// (row 192 h + 64 part + d = original row 768 part + 64 h + d).  clip_rows (0 = M): the rows are tiled clip by clip (every clip
// the same way, tiles never straddle a clip), so a sequence's place inside its MMA tile does not depend on the batch.  Returns 0 when it ran, -1 when T is not covered (T > 64:
// callers run the qkv GEMM + temporal attention kernels), DVST_ERR (1) on error.
int launch_gemm_ln_tattn(const void* xb, long long ldxb, const void* w_perm, const float* bias_perm,
                         const float* colsum_perm, const float* ln_stats, float ln_eps, long long M, int T,
                         long long clip_rows, void* out, long long ldo, cudaStream_t stream);
int temporal_tile_geometry(int T, int* rows_per_tile, int* key_block);  // -1 when T is not covered
void set_temporal_fusion(int on);  // 1 (default) use the fused kernel where the permuted tensors are registered
int temporal_fusion();

int launch_layernorm768(const float* x, long long ldx, const float* gamma, const float* beta, void* out,
                        long long ldo, int out_kind, long long lo_off, long long M, float eps, cudaStream_t st,
                        long long rm_group = 0, long long rm_stride = 0, long long rm_offset = 0);
// xb = bf16(x), stats[row][j] = (sum, sumsq) of columns [128 j, 128 j + 128): the deferred-LayerNorm inputs for rows that
// no residual GEMM has written yet (CLS rows after embedding).
// xoff (optional): receives the exact row mean; xb and the statistics are then those of x - mean
int launch_rowstats768(const float* x, long long ldx, void* xb, long long ldxb, float* stats, long long M,
                       cudaStream_t st, float* xoff = nullptr, long long lo_off = 0);  // lo_off > 0: xb = [hi | lo]
// xoff[r] += mean of the centred row (sum of the six per-128-column sums / 768): moves the offsets to the current row means
int launch_row_offsets768(const float* stats, float* xoff, long long M, cudaStream_t st);
int launch_im2col16(const float* x, void* out, long long ldo, int split, long long lo_off, int B, int T, int H, int W,
                    const int* starts, cudaStream_t st);
// patch rows straight from decoded uint8 frames [n, Hf, Wf, 3] (normalise + centre crop + window gather + im2col in one pass)
int launch_im2col16_u8(const uint8_t* video, int Hf, int Wf, int rate, const float* mean3, const float* std3, int crop,
                       void* out, long long ldo, int split, long long lo_off, int B, int T, const int* starts,
                       cudaStream_t st);
int launch_cast_bf16(const float* in, long long ldi, void* out, long long ldo, int split, long long lo_off, long long M,
                     int cols, cudaStream_t st);
int launch_l2norm_rows(const float* in, long long ldi, void* out, long long ldo, int split, long long lo_off,
                       float* out_f32, long long M, int cols, cudaStream_t st);
int launch_colsum(const float* in, long long ldi, float* out, int rows, int cols, cudaStream_t st);
int launch_center_ema(float* center, const float* batch_sum, float denom, float momentum, int cols, cudaStream_t st);
int launch_tile_rows768(const float* table, long long period, float* dst, long long rows, cudaStream_t st);
int launch_broadcast_row768(const float* src, float* dst, long long ldd, int nrows, cudaStream_t st);
int launch_cls_mean(const float* in, void* out, long long ldo, int split, long long lo_off, long long B, int T,
                    float* out_f32, cudaStream_t st);
int launch_pair_loss(const float* s, long long lds, const float* t, long long ldt, const float* center, float tt,
                     float ts, int dim, int pairs, float* loss, const int* s_idx, const int* t_idx, cudaStream_t st);
int launch_embed_table(const float* pos, int P, const float* time_embed, int T0, float* out, int gh, int gw, int T,
                       cudaStream_t st);
// mode: 0 temporal, 1 spatial (shared CLS per clip), 2 joint, 3 space-only (one CLS row per (b, t))
// Kernel choice (all sm_100a device kernels - there is no host / library fallback): bf16 modes 1 / 3 with <= 256 tokens per
// frame and mode 2 run on the tcgen05 kernels of attention_sm100.cu; shapes those do not cover (more than 256 tokens per
// frame, rows not 16-byte aligned), the temporal sequences and split precision run on the mma.sync kernels of attention.cu.
// The dispatch depends on the shape only, never on the batch size.
// scratch (optional, modes 1 / 3): >= spatial_attention_scratch_bytes(B, T, N) bytes enables the one-pass softmax of the
// tcgen05 kernel (fix-up list of rows to recompute exactly)
int launch_attention(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                     int N, long long cls_base, int split, cudaStream_t st, void* scratch = nullptr,
                     long long scratch_bytes = 0);
long long spatial_attention_scratch_bytes(int B, int T, int N);

// Same modes 1..3, but only the CLS query of every sequence attends (last block when only CLS features are returned):
// mode 1 writes cls_out [B*T, 768] fp32, modes 2 / 3 write the CLS rows of `out`.
int launch_cls_attention(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode, int B, int T,
                         int N, long long cls_base, int split, cudaStream_t st);
// Spatial attention map softmax(q k^T / 8) of every (frame, head), fp32 [B*T, 12, 1+N, 1+N], CLS token first.
int launch_attention_probs(const void* qkv, long long ld, float* probs, int B, int T, int N, long long cls_base,
                           int split, cudaStream_t st);
// tcgen05 / TMEM kernel for modes 1 and 3 (bf16): returns 0 when it ran, 1 when the shape is not covered (callers then
// use the mma.sync kernels), anything else is an error code.
int launch_spatial_attention_sm100(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode,
                                   int B, int T, int N, long long cls_base, cudaStream_t st, void* scratch = nullptr,
                                   long long scratch_bytes = 0);
// flash-style tcgen05 kernel for joint space-time attention (mode 2, bf16): NT = patch tokens per clip; same return codes
int launch_joint_attention_sm100(const void* qkv, long long ld, void* out, long long ldo, int B, int NT,
                                 long long cls_base, cudaStream_t st);
// 1 (default) tcgen05: split kernel (attention_split_sm100.cu) where it applies, else first generation, one-pass softmax
// when scratch is given; 3 first generation one-pass; 2 first generation two-pass; 0 mma.sync
void set_attention_impl(int impl);
int attention_impl();
// split-precision (bf16 hi / lo operands) spatial attention on tcgen05 (attention_hp_sm100.cu): same return codes; needs the
// scratch of the one-pass softmax
int launch_spatial_attention_hp_sm100(const void* qkv, long long ld, void* out, long long ldo, float* cls_out, int mode,
                                      int B, int T, int N, long long cls_base, cudaStream_t st, void* scratch,
                                      long long scratch_bytes);

int launch_select_frames(const double* loss, long long ld, const int* n_loss, const int* n_unsampled, int V, int rate,
                         int N, int* out, cudaStream_t st);
int launch_plan_windows(const int* lengths, int n_videos, int local, int global, int* s_starts, int* t_starts,
                        int* s_idx, int* t_idx, int* n_s_out, int* n_t_out, int* n_groups_out, int* groups_out,
                        cudaStream_t st);
int launch_preprocess_u8(const uint8_t* video, int n, int H, int W, int rate, const float* mean3, const float* std3,
                         int crop, float* out, cudaStream_t st);

// weighted k-NN evaluation (knn.cu)
int launch_knn_similarity(const float* test, long long ldt, const float* train, long long ldr, int n_test, int n_train,
                          int dim, float* sim, long long lds, cudaStream_t st);
int launch_knn_topk(const float* sim, long long lds, int n_test, int n_train, int k, float* dist, int* idx, int ldk,
                    cudaStream_t st);
int launch_knn_vote(const float* dist, const int* idx, int ldk, const long long* train_labels,
                    const long long* test_labels, int n_test, int k, float T, int num_classes, int* rank, int* pred,
                    cudaStream_t st);

long long head_loss_scratch_bytes(int P);
int launch_head_loss(const void* zs, const void* zt, long long ldz, const void* w, const float* center, float tt,
                     float ts, int P, int out_dim, float* loss, void* scratch, long long scratch_bytes,
                     cudaStream_t st);

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols);
int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);

}  // namespace dvst
