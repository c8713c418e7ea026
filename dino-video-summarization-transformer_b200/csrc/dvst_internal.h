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
int num_sms();

struct GemmArgs {
  const void* A = nullptr;  // bf16 [M, lda]  (split: [hi | lo], lo at column K)
  long long lda = 0;
  const void* W = nullptr;  // bf16 [N, ldw]  (nn.Linear layout; split: [hi | lo])
  long long ldw = 0;
  const float* bias = nullptr;      // [N] or null
  const float* residual = nullptr;  // fp32 [M, ldr] or null (may alias out)
  long long ldr = 0;
  void* out = nullptr;
  long long ldo = 0;
  int out_kind = 0;  // 0 bf16, 1 fp32, 2 bf16 hi/lo pair
  long long lo_off = 0;
  int act = 0;  // 0 none, 1 GELU(erf)
  int M = 0, N = 0, K = 0;
  int split = 0;
};
int launch_gemm(const GemmArgs& g, cudaStream_t stream);

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols);

}  // namespace dvst
