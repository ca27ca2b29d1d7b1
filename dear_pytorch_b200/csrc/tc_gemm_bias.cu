// linear_bias: GEMM + bias  (one translation unit per GEMM flavour so the three instantiations compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

using FusionBias = cutlass::epilogue::fusion::LinCombPerColBiasEltAct<Ident, bf16, float, bf16>;
using GemmBias = TcGemm<ColMajor, FusionBias>;

// Y = X W^T + b
at::Tensor linear_bias(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias) {
  check_operand(x, "x"); check_operand(w, "w"); check_operand(bias, "bias");
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && bias.numel() == w.size(0), "linear_bias: shape mismatch");
  c10::cuda::CUDAGuard guard(x.device());
  int M = x.size(0), K = x.size(1), N = w.size(0);
  auto y = at::empty({M, N}, x.options());
  typename GemmBias::FusionArgs f{};
  f.alpha = 1.0f; f.beta = 0.0f;
  f.bias_ptr = reinterpret_cast<const bf16*>(bias.data_ptr());
  run<GemmBias>(M, N, K, reinterpret_cast<const bf16*>(x.data_ptr()), reinterpret_cast<const bf16*>(w.data_ptr()),
                reinterpret_cast<bf16*>(y.data_ptr()), f, x.get_device());
  return y;
}


}  // namespace dear_tc
