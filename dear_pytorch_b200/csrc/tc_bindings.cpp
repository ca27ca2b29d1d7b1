// Python bindings and shared state of the tensor-core GEMM extension (dear_pytorch_b200._tc).
#include <torch/extension.h>

#include <atomic>
#include <map>
#include <mutex>

namespace py = pybind11;

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

std::vector<at::Tensor> ffn_up_hw(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias);   // tc_ffn_hw.cu
at::Tensor ffn_dgelu_hw(const at::Tensor& dy, const at::Tensor& wt, const at::Tensor& z);
at::Tensor ffn_dgelu_hw_nt(const at::Tensor& dy, const at::Tensor& w, const at::Tensor& z);

}  // namespace dear_tc

#include "tc_variants.inc"   // declarations + tables of the generated configurations (tools/gen_tc_variants.py)

namespace {

template <class V, size_t N>
const V& pick(const V (&table)[N], int variant, const char* op) {
  TORCH_CHECK(variant >= 0 && variant < static_cast<int>(N), op, ": variant ", variant, " out of range [0, ", N, ")");
  return table[variant];
}

template <class V, size_t N>
std::vector<std::string> configs(const V (&table)[N]) {
  std::vector<std::string> out;
  for (const auto& v : table) out.emplace_back(v.config);
  return out;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "tcgen05/TMEM/TMA GEMMs with fused bias / GELU / dGELU epilogues (sm_100a)";
  m.def("ffn_up", [](const at::Tensor& x, const at::Tensor& w, const at::Tensor& b, int variant) {
    return pick(k_ffn_up, variant, "ffn_up").fn(x, w, b); }, py::arg("x"), py::arg("w"), py::arg("bias"), py::arg("variant") = 0,
    "H, Z = gelu(X W^T + b), X W^T + b");
  m.def("linear_bias", [](const at::Tensor& x, const at::Tensor& w, const at::Tensor& b, int variant) {
    return pick(k_linear_bias, variant, "linear_bias").fn(x, w, b); }, py::arg("x"), py::arg("w"), py::arg("bias"), py::arg("variant") = 0,
    "Y = X W^T + b");
  m.def("ffn_dgelu", [](const at::Tensor& dy, const at::Tensor& w, const at::Tensor& z, int variant) {
    return pick(k_ffn_dgelu, variant, "ffn_dgelu").fn(dy, w, z); }, py::arg("dy"), py::arg("w"), py::arg("z"), py::arg("variant") = 0,
    "dZ = (dY W) * gelu'(Z)");
  m.def("variants", []() {
    return std::map<std::string, std::vector<std::string>>{
        {"ffn_up", configs(k_ffn_up)}, {"linear_bias", configs(k_linear_bias)}, {"ffn_dgelu", configs(k_ffn_dgelu)}}; },
    "kernel configurations compiled for each op (index = `variant`)");
  m.def("ffn_up_hw", &dear_tc::ffn_up_hw, py::arg("x"), py::arg("w"), py::arg("bias"),
        "hand-written tcgen05 kernel (two-warpgroup epilogue, staged coalesced stores): H, Z = gelu(X W^T + b), X W^T + b");
  m.def("ffn_dgelu_hw", &dear_tc::ffn_dgelu_hw, py::arg("dy"), py::arg("wt"), py::arg("z"),
        "hand-written tcgen05 kernel: dZ = (dY Wt^T) * gelu'(Z), Wt = transposed down-projection weight [N, K]");
  m.def("ffn_dgelu_hw_nt", &dear_tc::ffn_dgelu_hw_nt, py::arg("dy"), py::arg("w"), py::arg("z"),
        "hand-written tcgen05 kernel: dZ = (dY W) * gelu'(Z) with W [K, N] as stored (MN-major B operand, no transpose)");
  m.def("launches", &dear_tc::launches);
}
