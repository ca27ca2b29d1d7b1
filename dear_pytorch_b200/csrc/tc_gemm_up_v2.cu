// ffn_up variant 2: MMA tile / cluster / 2-SM = (128, 256, 1, 1, false), scheduler = void
// (one translation unit per instantiation so they compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

std::vector<at::Tensor> ffn_up_v2(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias) {
  using G = TcGemm<ColMajor, FusionUp, 128, 256, 1, 1, false, void>;
  return ffn_up_impl<G>(x, w, bias);
}

}  // namespace dear_tc
