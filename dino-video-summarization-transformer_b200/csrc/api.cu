// C-ABI layer of libdvst.so. Declarations (with the reference interfaces they replace) are in include/dvst.h.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

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

}  // namespace dvst

using namespace dvst;

extern "C" {

const char* dvst_last_error(void) { return g_err; }
long long dvst_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int dvst_abi_version(void) { return DVST_ABI_VERSION; }

int dvst_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias, int M, int N, int K,
                   int split, const float* residual, long long ldr, void* out, long long ldo, int out_kind,
                   long long lo_off, int act, void* stream) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias;
  g.M = M; g.N = N; g.K = K; g.split = split;
  g.residual = residual; g.ldr = ldr; g.out = out; g.ldo = ldo; g.out_kind = out_kind; g.lo_off = lo_off; g.act = act;
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
