// ffn_dgelu variant 1: MMA tile / cluster / 2-SM = (256, 256, 2, 1, true), scheduler = void
// (one translation unit per instantiation so they compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

at::Tensor ffn_dgelu_v1(const at::Tensor& dy, const at::Tensor& w, const at::Tensor& z) {
  using G = TcGemm<RowMajor, FusionDGelu, 256, 256, 2, 1, true, void>;
  return ffn_dgelu_impl<G>(dy, w, z);
}

}  // namespace dear_tc
