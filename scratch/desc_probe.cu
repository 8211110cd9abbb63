#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
using namespace cute;
template <UMMA::Major MJ, class L> void probe(const char* name, L layout) {
  auto t = make_tensor(make_smem_ptr((tfloat32_t*)nullptr), layout);
  UMMA::SmemDescriptor d = UMMA::make_umma_desc<MJ>(t);
  printf("%s: lbo=%u (x16B) sbo=%u (x16B) layout_type=%u version=%u\n", name, (unsigned)d.leading_byte_offset_, (unsigned)d.stride_byte_offset_, (unsigned)d.layout_type_, (unsigned)d.version_);
}
int main() {
  probe<UMMA::Major::MN>("MN SW128 (128 x 8)  ", tile_to_shape(UMMA::Layout_MN_SW128_Atom<tfloat32_t>{}, Shape<_128,_8>{}));
  probe<UMMA::Major::MN>("MN SW128 (128 x 16) ", tile_to_shape(UMMA::Layout_MN_SW128_Atom<tfloat32_t>{}, Shape<_128,_16>{}));
  probe<UMMA::Major::MN>("MN SW128 (128 x 32) ", tile_to_shape(UMMA::Layout_MN_SW128_Atom<tfloat32_t>{}, Shape<_128,_32>{}));
  probe<UMMA::Major::MN>("MN SW128 (128 x 16) k-first", tile_to_shape(UMMA::Layout_MN_SW128_Atom<tfloat32_t>{}, Shape<_128,_16>{}, Step<_2,_1>{}));
  probe<UMMA::Major::MN>("MN SW64  (128 x 16) ", tile_to_shape(UMMA::Layout_MN_SW64_Atom<tfloat32_t>{}, Shape<_128,_16>{}));
  probe<UMMA::Major::MN>("MN INTER (128 x 16) ", tile_to_shape(UMMA::Layout_MN_INTER_Atom<tfloat32_t>{}, Shape<_128,_16>{}));
  probe<UMMA::Major::K>("K  SW64  (32 x 16)  ", tile_to_shape(UMMA::Layout_K_SW64_Atom<tfloat32_t>{}, Shape<_32,_16>{}));
  probe<UMMA::Major::K>("K  SW128 (128 x 32) ", tile_to_shape(UMMA::Layout_K_SW128_Atom<tfloat32_t>{}, Shape<_128,_32>{}));
  probe<UMMA::Major::MN>("MN SW128_32B (128 x 8) ", tile_to_shape(UMMA::Layout_MN_SW128_32B_Atom<tfloat32_t>{}, Shape<_128,_8>{}));
  probe<UMMA::Major::MN>("MN SW128_32B (128 x 16)", tile_to_shape(UMMA::Layout_MN_SW128_32B_Atom<tfloat32_t>{}, Shape<_128,_16>{}));
  probe<UMMA::Major::MN>("MN SW128_32B (128 x 16) k-first", tile_to_shape(UMMA::Layout_MN_SW128_32B_Atom<tfloat32_t>{}, Shape<_128,_16>{}, Step<_2,_1>{}));
  print(tile_to_shape(UMMA::Layout_MN_SW128_32B_Atom<tfloat32_t>{}, Shape<_128,_16>{})); printf("\n");
  print(tile_to_shape(UMMA::Layout_MN_SW128_32B_Atom<tfloat32_t>{}, Shape<_128,_16>{}, Step<_2,_1>{})); printf("\n");
  { auto L = tile_to_shape(UMMA::Layout_MN_SW128_32B_Atom<tfloat32_t>{}, Shape<_128,_8>{}); for (int k = 0; k < 8; k++) { printf("k=%d:", k); for (int m = 0; m < 40; m += 4) printf(" %4d", (int)L(m, k)); printf("\n"); } }
  print(tile_to_shape(UMMA::Layout_MN_SW128_Atom<tfloat32_t>{}, Shape<_128,_16>{})); printf("\n");
  return 0;
}
