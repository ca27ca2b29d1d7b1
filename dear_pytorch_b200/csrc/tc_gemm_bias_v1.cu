// linear_bias variant 1: MMA tile / cluster / 2-SM = (256, 256, 2, 1, true), scheduler = void
// (one translation unit per instantiation so they compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

at::Tensor linear_bias_v1(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias) {
  using G = TcGemm<ColMajor, FusionBias, 256, 256, 2, 1, true, void>;
  return linear_bias_impl<G>(x, w, bias);
}

}  // namespace dear_tc
