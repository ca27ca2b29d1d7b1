// ffn_up variant 4: MMA tile / cluster / 2-SM = (256, 128, 2, 2, true), scheduler = void
// (one translation unit per instantiation so they compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

std::vector<at::Tensor> ffn_up_v4(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias) {
  using G = TcGemm<ColMajor, FusionUp, 256, 128, 2, 2, true, void>;
  return ffn_up_impl<G>(x, w, bias);
}

}  // namespace dear_tc
