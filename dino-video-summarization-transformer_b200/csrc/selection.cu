// Host-loop replacements around the hot path:
//   * select_frames_kernel   - adaptive frame selection (datasets_custom/frame_selection_loader.py:115-172) for a whole
//                              batch of videos, one thread per video, in the reference's fp64 arithmetic (numpy pairwise
//                              np.sum, sequential np.cumsum, first-minimum argmin) so the index lists are identical;
//   * plan_windows_kernel    - the loader's window rule (dino_loss_loader.py:126-171) for a batch of videos as index
//                              arithmetic on the device: distinct (local, global) window starts and the per-frame rows
//                              into them - the video lengths travel as kernel parameters, so planning a batch needs no
//                              host-to-device copy and no synchronisation (and is CUDA-graph capturable);
//   * preprocess_u8_kernel   - loader front half (dino_loss_loader.py:80-92): every rate-th frame of the decoded uint8
//                              [n, H, W, 3] video -> / 255, - mean, / std (data_utils.py:308-325, same three fp32
//                              operations), T H W C -> T C H W, centre crop (transform.py:228-229): one pass, fp32 out.
#include <algorithm>
#include <vector>

#include "dvst_internal.h"

namespace dvst {

namespace {

constexpr int MAX_SELECT = 64;

// numpy's pairwise summation for contiguous float64 (umath loops: pairwise_sum_DOUBLE): < 8 sequential, <= 128 eight
// strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail, else split at
// (n/2 rounded down to a multiple of 8).  Written iteratively over the recursion's leaf blocks with an explicit stack.
__device__ double np_block_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}
__device__ double np_pairwise_sum(const double* a, int n) {
  // post-order evaluation of  sum(a, n) = n <= 128 ? block(a, n) : sum(a, n2) + sum(a + n2, n - n2)
  struct Frame { const double* a; int n; int state; double left; };
  Frame st[24];
  int sp = 0;
  st[0] = {a, n, 0, 0.0};
  double ret = 0.0;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.n <= 128) { ret = np_block_sum(f.a, f.n); --sp; continue; }
    int n2 = f.n / 2;
    n2 -= n2 % 8;
    if (f.state == 0) { f.state = 1; st[++sp] = {f.a, n2, 0, 0.0}; }
    else if (f.state == 1) { f.left = ret; f.state = 2; st[++sp] = {f.a + n2, f.n - n2, 0, 0.0}; }
    else { ret = __dadd_rn(f.left, ret); --sp; }
  }
  return ret;
}

__global__ void select_frames_kernel(const double* __restrict__ loss, long long ld, const int* __restrict__ n_loss,
                                     const int* __restrict__ n_unsampled, int V, int rate, int N,
                                     int* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const double* a = loss + static_cast<long long>(v) * ld;
  const int F = n_loss[v], nu = n_unsampled[v];
  int* o = out + static_cast<long long>(v) * N;
  const double total = np_pairwise_sum(a, F);
  int used[MAX_SELECT];
  int n_used = 0;
  for (int i = 0; i < N; ++i) {
    const double q = static_cast<double>(i) / static_cast<double>(N);
    // idx = argmin_k |cumsum(loss / total)_k - q|, first minimum (np.argmin)
    double c = 0.0, best = 0.0;
    int idx = 0;
    for (int k = 0; k < F; ++k) {
      c = __dadd_rn(c, __ddiv_rn(a[k], total));
      const double d = fabs(__dsub_rn(c, q));
      if (k == 0 || d < best) { best = d; idx = k; }
    }
    auto is_used = [&](int s) {
      for (int u = 0; u < n_used; ++u)
        if (used[u] == s) return true;
      return false;
    };
    int s = idx * rate;
    if (is_used(s)) {
      // walk forward one pre-sampling step at a time (:153-168)
      int t = s + rate;
      while (true) {
        if (!is_used(t) && t < nu) { s = t; break; }
        if (t >= nu) { s = t - rate; break; }
        t += rate;
      }
    }
    used[n_used++] = s;
    o[i] = s;
  }
}

// One block per video.  Exactly the reference's clamp code (dino_loss_loader.py:137-156): j = i - loc, k = i + loc + 1,
// l = i - glob, m = i + glob; j < 0 -> (0, local); k >= F -> (F - local, F); l < 0 -> (0, G); m >= F -> (F - G, F).
// With F >= local the distinct local starts are exactly 0 .. F - local and the distinct global starts 0 .. F - G, so the
// row of a window inside its list is its start.
struct PlanVideo { int off, F, G, base_s, base_t; };
constexpr int PLAN_CHUNK = 128;
struct PlanBatch { PlanVideo v[PLAN_CHUNK]; };

__global__ void plan_windows_kernel(const PlanBatch pb, int local, int global, int* __restrict__ s_starts,
                                    int* __restrict__ t_starts, int* __restrict__ s_idx, int* __restrict__ t_idx) {
  const PlanVideo v = pb.v[blockIdx.x];
  const int loc = local / 2, glob = global / 2;
  for (int i = threadIdx.x; i < v.F; i += blockDim.x) {
    int j = i - loc, k = i + loc + 1, l = i - glob, m = i + glob;
    if (j < 0) { j = 0; k = local; }
    if (k >= v.F) { k = v.F; j = v.F - local; }
    if (l < 0) { l = 0; m = v.G; }
    if (m >= v.F) { m = v.F; l = v.F - v.G; }
    s_idx[v.off + i] = v.base_s + j;
    t_idx[v.off + i] = v.base_t + l;
  }
  for (int j = threadIdx.x; j <= v.F - local; j += blockDim.x) s_starts[v.base_s + j] = v.off + j;
  for (int l = threadIdx.x; l <= v.F - v.G; l += blockDim.x) t_starts[v.base_t + l] = v.off + l;
}

__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ video, int H, int W, int rate, int F, int crop, int y0,
                                     int x0, float m0, float m1, float m2, float s0, float s1, float s2,
                                     float* __restrict__ out) {
  // one thread per output pixel (f, y, x): reads its 3 interleaved channel bytes, writes 3 planes
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(F) * crop * crop;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % crop);
  const int y = static_cast<int>((idx / crop) % crop);
  const int f = static_cast<int>(idx / (static_cast<long long>(crop) * crop));
  const uint8_t* src = video + ((static_cast<long long>(f) * rate * H + (y0 + y)) * W + (x0 + x)) * 3;
  const float r = static_cast<float>(src[0]), g = static_cast<float>(src[1]), b = static_cast<float>(src[2]);
  const long long plane = static_cast<long long>(crop) * crop;
  float* dst = out + static_cast<long long>(f) * 3 * plane + static_cast<long long>(y) * crop + x;
  // tensor / 255.0, tensor - mean, tensor / std: three separately rounded fp32 operations, as in the reference
  dst[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(r, 255.0f), m0), s0);
  dst[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(g, 255.0f), m1), s1);
  dst[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(b, 255.0f), m2), s2);
}

}  // namespace

int launch_select_frames(const double* loss, long long ld, const int* n_loss, const int* n_unsampled, int V, int rate,
                         int N, int* out, cudaStream_t st) {
  if (V <= 0) return 0;
  if (N <= 0 || N > MAX_SELECT) return set_error("select_frames: num_frames must be in 1..%d (got %d)", MAX_SELECT, N);
  if (rate <= 0) return set_error("select_frames: bad pre-sampling rate %d", rate);
  select_frames_kernel<<<(V + 63) / 64, 64, 0, st>>>(loss, ld, n_loss, n_unsampled, V, rate, N, out);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("select_frames launch: %s", cudaGetErrorString(e));
}

int launch_plan_windows(const int* lengths, int n_videos, int local, int global, int* s_starts, int* t_starts,
                        int* s_idx, int* t_idx, int* n_s_out, int* n_t_out, int* n_groups_out, int* groups_out,
                        cudaStream_t st) {
  if (n_videos <= 0) { *n_s_out = *n_t_out = *n_groups_out = 0; return 0; }
  if (local <= 0 || global <= 0) return set_error("plan_windows: bad window sizes %d / %d", local, global);
  // groups: distinct teacher window lengths G = min(global, F), ascending; each group's windows are contiguous rows
  std::vector<int> Gs;
  for (int v = 0; v < n_videos; ++v) {
    if (lengths[v] < local) return set_error("plan_windows: video %d has %d frames, fewer than the local window %d", v, lengths[v], local);
    const int G = lengths[v] < global ? lengths[v] : global;
    if (std::find(Gs.begin(), Gs.end(), G) == Gs.end()) Gs.push_back(G);
  }
  std::sort(Gs.begin(), Gs.end());
  std::vector<int> gcount(Gs.size(), 0), gbase(Gs.size(), 0);
  for (int v = 0; v < n_videos; ++v) {
    const int F = lengths[v], G = F < global ? F : global;
    gcount[std::lower_bound(Gs.begin(), Gs.end(), G) - Gs.begin()] += F - G + 1;
  }
  int n_t = 0;
  for (size_t g = 0; g < Gs.size(); ++g) { gbase[g] = n_t; n_t += gcount[g]; }
  std::vector<int> gfill(gbase);
  long long off = 0, n_s = 0;
  for (int v0 = 0; v0 < n_videos; v0 += PLAN_CHUNK) {
    PlanBatch pb;
    const int nv = n_videos - v0 < PLAN_CHUNK ? n_videos - v0 : PLAN_CHUNK;
    for (int u = 0; u < nv; ++u) {
      const int F = lengths[v0 + u], G = F < global ? F : global;
      const size_t g = std::lower_bound(Gs.begin(), Gs.end(), G) - Gs.begin();
      if (off + F > 0x7FFFFFFFLL) return set_error("plan_windows: too many frames");
      pb.v[u] = PlanVideo{static_cast<int>(off), F, G, static_cast<int>(n_s), gfill[g]};
      gfill[g] += F - G + 1;
      n_s += F - local + 1;
      off += F;
    }
    plan_windows_kernel<<<nv, 128, 0, st>>>(pb, local, global, s_starts, t_starts, s_idx, t_idx);
    count_launch();
  }
  *n_s_out = static_cast<int>(n_s);
  *n_t_out = n_t;
  *n_groups_out = static_cast<int>(Gs.size());
  for (size_t g = 0; g < Gs.size(); ++g) {
    groups_out[3 * g] = Gs[g];
    groups_out[3 * g + 1] = gcount[g];
    groups_out[3 * g + 2] = gbase[g];
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("plan_windows launch: %s", cudaGetErrorString(e));
}

int launch_preprocess_u8(const uint8_t* video, int n, int H, int W, int rate, const float* mean3, const float* std3,
                         int crop, float* out, cudaStream_t st) {
  if (n <= 0) return 0;
  if (rate <= 0 || crop <= 0 || H < crop || W < crop)
    return set_error("preprocess: bad rate %d / crop %d for %dx%d frames", rate, crop, H, W);
  const int F = (n + rate - 1) / rate;
  // ceil((H - crop) / 2) like transform.py:228-229 (math.ceil of a float division)
  const int y0 = (H - crop + 1) / 2, x0 = (W - crop + 1) / 2;
  const long long total = static_cast<long long>(F) * crop * crop;
  ProfScope prof(PROF_ELEMENTWISE, static_cast<double>(total) * 15.0, st);
  preprocess_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      video, H, W, rate, F, crop, y0, x0, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : set_error("preprocess launch: %s", cudaGetErrorString(e));
}

}  // namespace dvst
