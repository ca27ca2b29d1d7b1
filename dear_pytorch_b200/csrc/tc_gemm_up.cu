// ffn_up: GEMM + bias + GELU, pre-activation kept for the backward  (one translation unit per GEMM flavour so the three instantiations compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

using FusionUp = cutlass::epilogue::fusion::LinCombPerColBiasEltActAux<RowMajor, GeluErf, bf16, float, bf16, bf16>;
using GemmUp = TcGemm<ColMajor, FusionUp>;          // B = W [N,K] row-major == K x N column-major

// H = gelu(Z), Z = X W^T + b.   x [M,K], w [N,K], bias [N]  ->  (H [M,N], Z [M,N])
std::vector<at::Tensor> ffn_up(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias) {
  check_operand(x, "x"); check_operand(w, "w"); check_operand(bias, "bias");
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && bias.numel() == w.size(0), "ffn_up: shape mismatch");
  c10::cuda::CUDAGuard guard(x.device());
  int M = x.size(0), K = x.size(1), N = w.size(0);
  auto h = at::empty({M, N}, x.options());
  auto z = at::empty({M, N}, x.options());
  typename GemmUp::FusionArgs f{};
  f.alpha = 1.0f; f.beta = 0.0f;
  f.bias_ptr = reinterpret_cast<const bf16*>(bias.data_ptr());
  f.aux_ptr = reinterpret_cast<bf16*>(z.data_ptr());
  f.dAux = cutlass::make_cute_packed_stride(typename GemmUp::StrideD{}, cute::make_shape(M, N, 1));
  run<GemmUp>(M, N, K, reinterpret_cast<const bf16*>(x.data_ptr()), reinterpret_cast<const bf16*>(w.data_ptr()),
              reinterpret_cast<bf16*>(h.data_ptr()), f, x.get_device());
  return {h, z};
}


}  // namespace dear_tc
