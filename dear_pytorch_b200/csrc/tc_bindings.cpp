// Python bindings and shared state of the tensor-core GEMM extension (dear_pytorch_b200._tc).
#include <torch/extension.h>

#include <atomic>
#include <map>
#include <mutex>

namespace dear_tc {

static std::atomic<long> g_launches{0};
static std::mutex g_ws_mu;
static std::map<int, at::Tensor> g_workspace;       // per device, grown on demand, never shrunk (CUDA-graph safe)

void* workspace(size_t bytes, int device) {
  if (bytes == 0) return nullptr;
  std::lock_guard<std::mutex> lk(g_ws_mu);
  auto it = g_workspace.find(device);
  if (it == g_workspace.end() || static_cast<size_t>(it->second.numel()) < bytes) {
    g_workspace[device] =
        at::empty({static_cast<long>(bytes)}, at::TensorOptions().dtype(at::kByte).device(at::kCUDA, device));
    it = g_workspace.find(device);
  }
  return it->second.data_ptr();
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long launches() { return g_launches.load(); }

std::vector<at::Tensor> ffn_up(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias);
at::Tensor linear_bias(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias);
at::Tensor ffn_dgelu(const at::Tensor& dy, const at::Tensor& w, const at::Tensor& z);

}  // namespace dear_tc

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "tcgen05/TMEM/TMA GEMMs with fused bias / GELU / dGELU epilogues (sm_100a)";
  m.def("ffn_up", &dear_tc::ffn_up, "H, Z = gelu(X W^T + b), X W^T + b");
  m.def("linear_bias", &dear_tc::linear_bias, "Y = X W^T + b");
  m.def("ffn_dgelu", &dear_tc::ffn_dgelu, "dZ = (dY W) * gelu'(Z)");
  m.def("launches", &dear_tc::launches);
}
