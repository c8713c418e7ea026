// Weighted k-NN evaluation of frozen features (eval_knn.py:138-178, `knn_classifier`) on the device:
//
//   similarity = test_features @ train_features.T         -> knn_similarity_kernel (fp32 FMA, k ascending per element)
//   distances, indices = similarity.topk(k, sorted=True)  -> knn_topk_kernel       (radix select + bitonic sort per row)
//   probs[c] = sum_j [label_j == c] * exp(distance_j / T) -> knn_vote_kernel       (sequential over j like torch.sum,
//   predictions = probs.sort(descending)                      then the rank of the target class instead of a full sort)
//
// The similarity stays in fp32 CUDA-core arithmetic on purpose: neighbour ORDER is the result here, and bf16 / tf32
// tensor-core products would reorder near-ties that the reference's fp32 `torch.mm` resolves.  The problem is small
// (UCF-101: 3 783 x 9 537 x 768 = 55 GFLOP) next to the backbone forward that produced the features.
#include "dvst_internal.h"

namespace dvst {

namespace {

// ---- C[M,N] = A[M,K] * B[N,K]^T (both row-major, K contiguous), fp32 ----------------------------------------------
constexpr int SB = 128;  // tile rows (test) and columns (train)
constexpr int SK = 16;   // k-slab

__global__ void __launch_bounds__(256)
knn_similarity_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb, int M, int N,
                      int K, float* __restrict__ C, long long ldc) {
  __shared__ __align__(16) float As[SK][SB + 4];
  __shared__ __align__(16) float Bs[SK][SB + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 8 x 8 outputs each (rows ty*4 + {0..3, 64..67}: LDS.128)
  const int m0 = blockIdx.y * SB, n0 = blockIdx.x * SB;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SK) {
    // 128 rows x 16 k = 512 float4 per operand, 2 per thread; stored k-major so the inner loop reads rows contiguously
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int f = tid + r * 256;
      const int row = f >> 2, k4 = (f & 3) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (k0 + k4 < K) {
        if (m0 + row < M) a = __ldg(reinterpret_cast<const float4*>(A + (m0 + row) * lda + k0 + k4));
        if (n0 + row < N) b = __ldg(reinterpret_cast<const float4*>(B + (n0 + row) * ldb + k0 + k4));
      }
      As[k4][row] = a.x; As[k4 + 1][row] = a.y; As[k4 + 2][row] = a.z; As[k4 + 3][row] = a.w;
      Bs[k4][row] = b.x; Bs[k4 + 1][row] = b.y; Bs[k4 + 2][row] = b.z; Bs[k4 + 3][row] = b.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n < N) C[m * ldc + n] = acc[i][j];
    }
  }
}

// ---- per-row top-k: values descending, ties by ascending index ----------------------------------------------------
constexpr int TOPK_THREADS = 256;
constexpr int TOPK_CAP = 1024;  // candidate capacity (k <= 512 plus ties at the threshold)

// order-preserving map fp32 -> uint32 (larger float = larger key); NaN sorts above +inf like torch.topk treats it
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(TOPK_THREADS)
knn_topk_kernel(const float* __restrict__ sim, long long lds, int n_train, int k, float* __restrict__ dist,
                int* __restrict__ idx, int ldk) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_remaining;
  __shared__ int s_count;
  __shared__ unsigned long long cand[TOPK_CAP];  // (key << 32) | ~index : one descending sort orders value desc, index asc
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const float* v = sim + row * lds;

  // 4-pass MSB radix select of the k-th largest key
  if (tid == 0) { s_prefix = 0; s_remaining = static_cast<uint32_t>(k); }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < n_train; i += TOPK_THREADS) {
      const uint32_t key = f2key(v[i]);
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t rem = s_remaining;
      int b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= rem) break;
        rem -= hist[b];
      }
      s_prefix = prefix | (static_cast<uint32_t>(b) << shift);
      s_remaining = rem;
    }
    __syncthreads();
  }
  const uint32_t thr = s_prefix;           // key of the k-th largest element
  const uint32_t need_eq = s_remaining;    // how many elements equal to it belong to the top k (lowest indices first)

  if (tid == 0) s_count = 0;
  __syncthreads();
  // everything strictly above the threshold (k - need_eq elements), order fixed later by the sort
  for (int i = tid; i < n_train; i += TOPK_THREADS) {
    const uint32_t key = f2key(v[i]);
    if (key > thr) {
      const int p = atomicAdd(&s_count, 1);
      cand[p] = (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(i));
    }
  }
  __syncthreads();
  // the `need_eq` lowest-index elements equal to the threshold: ordered scan, 256 elements per step
  {
    __shared__ int s_taken;
    __shared__ uint32_t wsum[TOPK_THREADS / 32];
    if (tid == 0) s_taken = 0;
    __syncthreads();
    for (int base = 0; base < n_train; base += TOPK_THREADS) {
      if (s_taken >= static_cast<int>(need_eq)) break;  // uniform: s_taken only changes between barriers
      const int i = base + tid;
      const bool eq = i < n_train && f2key(v[i]) == thr;
      const uint32_t ball = __ballot_sync(0xFFFFFFFFu, eq);
      const int lane = tid & 31, w = tid >> 5;
      if (lane == 0) wsum[w] = __popc(ball);
      __syncthreads();
      int before = s_taken;
      for (int j = 0; j < w; ++j) before += wsum[j];
      before += __popc(ball & ((1u << lane) - 1u));
      if (eq && before < static_cast<int>(need_eq)) {
        const int p = atomicAdd(&s_count, 1);
        cand[p] = (static_cast<unsigned long long>(thr) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(i));
      }
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int j = 0; j < TOPK_THREADS / 32; ++j) tot += wsum[j];
        s_taken += tot;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // exactly k candidates now; pad to a power of two and bitonic-sort descending
  int n2 = 1;
  while (n2 < k) n2 <<= 1;
  for (int i = k + tid; i < n2; i += TOPK_THREADS) cand[i] = 0ull;
  __syncthreads();
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < n2 / 2; t += TOPK_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = cand[lo], b = cand[hi];
        if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += TOPK_THREADS) {
    const int i = static_cast<int>(~static_cast<uint32_t>(cand[j] & 0xFFFFFFFFull));
    idx[row * static_cast<long long>(ldk) + j] = i;
    dist[row * static_cast<long long>(ldk) + j] = v[i];
  }
}

// ---- vote: rank of the target class in probs.sort(descending) (stable: ties by ascending class index) -------------
constexpr int VOTE_THREADS = 128;

__global__ void __launch_bounds__(VOTE_THREADS)
knn_vote_kernel(const float* __restrict__ dist, const int* __restrict__ idx, int ldk, const long long* __restrict__ train_labels,
                const long long* __restrict__ test_labels, int k, float T, int num_classes, int* __restrict__ rank_out,
                int* __restrict__ pred_out) {
  extern __shared__ float probs[];  // [num_classes]
  __shared__ int s_cnt;
  const int row = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < num_classes; c += VOTE_THREADS) probs[c] = 0.f;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (tid == 0) {
    // torch.sum over the neighbour axis adds the k one-hot rows in order j = 0..k-1 (eval_knn.py:160-166)
    for (int j = 0; j < k; ++j) {
      const long long p = row * static_cast<long long>(ldk) + j;
      const long long lab = train_labels[idx[p]];
      if (lab >= 0 && lab < num_classes) probs[lab] += expf(dist[p] / T);
    }
  }
  __syncthreads();
  const long long target = test_labels[row];
  const float pt = (target >= 0 && target < num_classes) ? probs[target] : -1.f;
  int cnt = 0;
  // argmax with the same tie rule (prediction = first class of the stable descending sort)
  float best = -1.f;
  int best_c = 0;
  for (int c = tid; c < num_classes; c += VOTE_THREADS) {
    const float p = probs[c];
    if (p > pt || (p == pt && c < target)) ++cnt;
    if (p > best) { best = p; best_c = c; }
  }
  atomicAdd(&s_cnt, cnt);
  __shared__ float s_best[VOTE_THREADS];
  __shared__ int s_best_c[VOTE_THREADS];
  s_best[tid] = best;
  s_best_c[tid] = best_c;
  __syncthreads();
  if (tid == 0) {
    for (int t = 1; t < VOTE_THREADS; ++t)
      if (s_best[t] > best || (s_best[t] == best && s_best_c[t] < best_c)) { best = s_best[t]; best_c = s_best_c[t]; }
    rank_out[row] = (target >= 0 && target < num_classes) ? s_cnt : num_classes;
    if (pred_out != nullptr) pred_out[row] = best_c;
  }
}

}  // namespace

int launch_knn_similarity(const float* test, long long ldt, const float* train, long long ldr, int n_test, int n_train,
                          int dim, float* sim, long long lds, cudaStream_t st) {
  if (n_test <= 0 || n_train <= 0) return 0;
  if (dim <= 0 || dim % 4 || ldt % 4 || ldr % 4 || (reinterpret_cast<uintptr_t>(test) & 15) ||
      (reinterpret_cast<uintptr_t>(train) & 15))
    return set_error("knn_similarity: dim and row strides must be multiples of 4 floats, pointers 16-byte aligned");
  dim3 grid((n_train + SB - 1) / SB, (n_test + SB - 1) / SB);
  if (grid.y > 65535) return set_error("knn_similarity: more than %d test rows per call", 65535 * SB);
  ProfScope prof(PROF_ELEMENTWISE, 4.0 * (static_cast<double>(n_test) * dim + static_cast<double>(n_train) * dim +
                                          static_cast<double>(n_test) * n_train), st);
  knn_similarity_kernel<<<grid, 256, 0, st>>>(test, ldt, train, ldr, n_test, n_train, dim, sim, lds);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("knn_similarity launch: %s", cudaGetErrorString(e));
}

int launch_knn_topk(const float* sim, long long lds, int n_test, int n_train, int k, float* dist, int* idx, int ldk,
                    cudaStream_t st) {
  if (n_test <= 0) return 0;
  if (k <= 0 || k > n_train) return set_error("knn_topk: k=%d must be in [1, n_train=%d]", k, n_train);
  if (k > TOPK_CAP / 2) return set_error("knn_topk: k=%d exceeds the supported maximum %d", k, TOPK_CAP / 2);
  if (ldk < k) return set_error("knn_topk: ldk < k");
  ProfScope prof(PROF_ELEMENTWISE, 4.0 * 6.0 * static_cast<double>(n_test) * n_train, st);
  knn_topk_kernel<<<n_test, TOPK_THREADS, 0, st>>>(sim, lds, n_train, k, dist, idx, ldk);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("knn_topk launch: %s", cudaGetErrorString(e));
}

int launch_knn_vote(const float* dist, const int* idx, int ldk, const long long* train_labels,
                    const long long* test_labels, int n_test, int k, float T, int num_classes, int* rank, int* pred,
                    cudaStream_t st) {
  if (n_test <= 0) return 0;
  if (k <= 0 || k > ldk) return set_error("knn_vote: k=%d must be in [1, ldk=%d]", k, ldk);
  if (!(T > 0.f)) return set_error("knn_vote: temperature must be > 0");
  if (num_classes <= 0 || num_classes > 12000) return set_error("knn_vote: num_classes=%d out of range", num_classes);
  ProfScope prof(PROF_ELEMENTWISE, 8.0 * static_cast<double>(n_test) * k, st);
  knn_vote_kernel<<<n_test, VOTE_THREADS, num_classes * sizeof(float), st>>>(dist, idx, ldk, train_labels, test_labels, k,
                                                                            T, num_classes, rank, pred);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("knn_vote launch: %s", cudaGetErrorString(e));
}

}  // namespace dvst
