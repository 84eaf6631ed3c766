// Host build of the engine's bit-parallel OSA routine (pclean_b200/csrc/osa_bitpar.cuh) so the
// exact code the CUDA kernel runs is checked on the CPU against a plain dynamic programme.
#include <cstdint>
#include <vector>
#include "../pclean_b200/csrc/osa_bitpar.cuh"
extern "C" int osa_bitpar(const uint8_t* a, int n, const uint8_t* b0, int l0, const uint8_t* b1, int l1, const uint8_t* b2, int l2) {
  int words = n > 64 ? (n + 63) / 64 : 1;
  std::vector<uint64_t> peq(256 * words, 0);
  osa_build_peq(a, n, words, peq.data());
  OsaText t; t.seg[0] = b0; t.len[0] = l0; t.seg[1] = b1; t.len[1] = l1; t.seg[2] = b2; t.len[2] = l2;
  return osa_distance(peq.data(), n, words, t);
}
