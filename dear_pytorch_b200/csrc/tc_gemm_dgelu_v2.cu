// ffn_dgelu variant 2: MMA tile / cluster / 2-SM = (128, 256, 1, 1, false), scheduler = void
// (one translation unit per instantiation so they compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

at::Tensor ffn_dgelu_v2(const at::Tensor& dy, const at::Tensor& w, const at::Tensor& z) {
  using G = TcGemm<RowMajor, FusionDGelu, 128, 256, 1, 1, false, void>;
  return ffn_dgelu_impl<G>(dy, w, z);
}

}  // namespace dear_tc
