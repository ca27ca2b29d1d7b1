#include <cstdio>
#include <cstdint>
#include <cute/tensor.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
using namespace cute;
static uint64_t mine(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
int main() {
  auto id = UMMA::make_instr_desc<cutlass::bfloat16_t, cutlass::bfloat16_t, float, 128, 256, UMMA::Major::K, UMMA::Major::K>();
  uint32_t my = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(256 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  printf("idesc cute=%08x mine=%08x %s\n", uint32_t(id), my, uint32_t(id) == my ? "MATCH" : "DIFF");
  UMMA::SmemDescriptor sd;
  sd.desc_ = 0;
  uint32_t addr = 0x12400;   // 1024-aligned example
  sd.start_address_ = addr >> 4; sd.leading_byte_offset_ = 1; sd.stride_byte_offset_ = 64; sd.version_ = 1;
  sd.base_offset_ = 0; sd.lbo_mode_ = 0; sd.layout_type_ = uint8_t(UMMA::LayoutType::SWIZZLE_128B);
  printf("sdesc cute=%016llx mine=%016llx %s\n", (unsigned long long)sd.desc_, (unsigned long long)mine(addr), sd.desc_ == mine(addr) ? "MATCH" : "DIFF");
  // what CuTe itself builds for a K-major SW128 bf16 tile 128x64 (checks LBO/SBO values)
  using Atom = UMMA::Layout_K_SW128_Atom<cutlass::bfloat16_t>;
  auto layout = tile_to_shape(Atom{}, Shape<_128,_64>{});
  print(layout); printf("\n");
  return 0;
}
