// linear_bias variant 2: MMA tile / cluster / 2-SM = (128, 128, 1, 1, false), scheduler = void
// (one translation unit per instantiation so they compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

at::Tensor linear_bias_v2(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias) {
  using G = TcGemm<ColMajor, FusionBias, 128, 128, 1, 1, false, void>;
  return linear_bias_impl<G>(x, w, bias);
}

}  // namespace dear_tc
