#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
using namespace cute;
int main() {
  using Atom = UMMA::Layout_MN_SW128_Atom<cutlass::bfloat16_t>;
  print(Atom{}); printf("\n");
  auto l = tile_to_shape(Atom{}, Shape<_256,_64>{});            // (N, K) with N contiguous
  print(l); printf("\n");
  auto l2 = tile_to_shape(Atom{}, Shape<_256,_64>{}, Step<_2,_1>{});   // K-blocks outer?
  print(l2); printf("\n");
  // canonical form in uint128 units, as make_umma_desc<Major::MN> derives it
  auto u = recast_layout<cutlass::bfloat16_t, uint128_t>(l.layout_b());
  print(u); printf("\n");
  auto canon = logical_divide(u, Tile<Layout<_8>, Layout<_8>>{});
  print(canon); printf("\n");
  return 0;
}
