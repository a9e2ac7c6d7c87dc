// 3xTF32 tcgen05 GEMM of the float32 opt-in (gemm_tf32.cu).
#pragma once
#include "gemm_f64.h"

namespace gpfx {

// float32 workspace elements for the split operands of one product
size_t gemm_tf32_workspace_floats(int batch, int M, int N, int K);
// C = alpha op(A) op(B) + beta C (tril optional) from float64 operands described like launch_gemm's
// (plain products only: no fused column reductions / add-in terms / split-K / k-segments)
int launch_gemm_tf32x3(const GemmArgs& a, float* workspace, cudaStream_t stream);

// true when `a` is a product the 3xTF32 kernel takes (large, plain, per-batch operands)
bool gemm_tf32_eligible(const GemmArgs& a);
// launch_gemm, except that eligible products go through the 3xTF32 kernel when the current handle's
// contraction mode asks for it and its workspace is large enough
int launch_gemm_auto(const GemmArgs& a, cudaStream_t stream);
// true when launch_gemm_auto would take the 3xTF32 kernel for `a` (ignoring fused-epilogue fields)
bool gemm_tf32_active(const GemmArgs& a);

}  // namespace gpfx
