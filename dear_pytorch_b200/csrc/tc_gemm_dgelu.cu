// ffn_dgelu: dgrad GEMM with the GELU backward in the epilogue  (one translation unit per GEMM flavour so the three instantiations compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

using FusionDGelu = cutlass::epilogue::fusion::LinCombDeEltAct<RowMajor, DGeluErf, bf16, float, bf16>;
using GemmDGelu = TcGemm<RowMajor, FusionDGelu>;    // B = W [K,N] row-major (dY [M,K] x W [K,N])

// dZ = (dY W) * gelu'(Z).   dy [M,K], w [K,N] (the down projection's weight as stored: [out=K, in=N]), z [M,N]
at::Tensor ffn_dgelu(const at::Tensor& dy, const at::Tensor& w, const at::Tensor& z) {
  check_operand(dy, "dy"); check_operand(w, "w"); check_operand(z, "z");
  TORCH_CHECK(dy.dim() == 2 && w.dim() == 2 && z.dim() == 2 && dy.size(1) == w.size(0) && z.size(0) == dy.size(0) &&
              z.size(1) == w.size(1), "ffn_dgelu: shape mismatch");
  c10::cuda::CUDAGuard guard(dy.device());
  int M = dy.size(0), K = dy.size(1), N = w.size(1);
  auto dz = at::empty({M, N}, dy.options());
  typename GemmDGelu::FusionArgs f{};
  f.alpha = 1.0f; f.beta = 0.0f;
  f.aux_ptr = reinterpret_cast<const bf16*>(z.data_ptr());
  f.dAux = cutlass::make_cute_packed_stride(typename GemmDGelu::StrideD{}, cute::make_shape(M, N, 1));
  run<GemmDGelu>(M, N, K, reinterpret_cast<const bf16*>(dy.data_ptr()), reinterpret_cast<const bf16*>(w.data_ptr()),
                 reinterpret_cast<bf16*>(dz.data_ptr()), f, dy.get_device());
  return dz;
}


}  // namespace dear_tc
