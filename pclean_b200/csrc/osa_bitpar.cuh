// osa_bitpar.cuh — bit-parallel optimal-string-alignment (restricted Damerau-Levenshtein)
// distance, Hyyrö 2003, one thread per string pair, 64-bit words, multi-word blocks for
// patterns longer than 64 symbols.  Computes what `evaluate(DamerauLevenshtein(), a, b)`
// computes in the reference (add_typos.jl:56) — see DESIGN.md on the OSA/true-DL ambiguity.
//
// Strings are sequences of uint8 *symbol ids* (the engine compacts the codepoints of the
// dictionary into an alphabet of <= 256 symbols at load time).  The pattern's match masks
// (PEq) live in shared memory (device) or a plain array (host): peq[sym * words + w].
//
// Compiles for host too, so the routine itself is unit-tested on the CPU (tests/test_osa.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OSA_HD __host__ __device__ __forceinline__
#else
#define OSA_HD static inline
#endif

#define OSA_MAX_WORDS 4          /* patterns up to 256 symbols */

struct OsaText {                 /* a clean string given as up to three concatenated segments */
  const uint8_t* seg[3];
  int len[3];
};

/* Build PEq for `pat` (length m, words = ceil(m/64)) into peq[alphabet * words] (zeroed by caller). */
OSA_HD void osa_build_peq(const uint8_t* pat, int m, int words, uint64_t* peq) {
  for (int i = 0; i < m; ++i) peq[(int)pat[i] * words + (i >> 6)] |= (uint64_t)1 << (i & 63);
}

/* Single-word kernel: pattern length 1..64. */
OSA_HD int osa_distance_1w(const uint64_t* peq, int m, const OsaText& t) {
  uint64_t VP = ~(uint64_t)0, VN = 0, D0 = 0, PMold = 0;
  const uint64_t mask = (uint64_t)1 << (m - 1);
  int dist = m;
  for (int s = 0; s < 3; ++s) {
    const uint8_t* p = t.seg[s];
    for (int j = 0; j < t.len[s]; ++j) {
      const uint64_t PM = peq[p[j]];
      const uint64_t TR = (((~D0) & PM) << 1) & PMold;
      D0 = (((PM & VP) + VP) ^ VP) | PM | VN;
      D0 |= TR;
      uint64_t HP = VN | ~(D0 | VP);
      uint64_t HN = D0 & VP;
      dist += (HP & mask) != 0;
      dist -= (HN & mask) != 0;
      HP = (HP << 1) | 1;
      HN = HN << 1;
      VP = HN | ~(D0 | HP);
      VN = HP & D0;
      PMold = PM;
    }
  }
  return dist;
}

/* Multi-word kernel (2..OSA_MAX_WORDS words). */
OSA_HD int osa_distance_mw(const uint64_t* peq, int m, int words, const OsaText& t) {
  uint64_t VP[OSA_MAX_WORDS], VN[OSA_MAX_WORDS], D0[OSA_MAX_WORDS], PMo[OSA_MAX_WORDS];
  for (int w = 0; w < words; ++w) { VP[w] = ~(uint64_t)0; VN[w] = 0; D0[w] = 0; PMo[w] = 0; }
  const uint64_t last = (uint64_t)1 << ((m - 1) & 63);
  int dist = m;
  for (int s = 0; s < 3; ++s) {
    const uint8_t* p = t.seg[s];
    for (int j = 0; j < t.len[s]; ++j) {
      uint64_t HPc = 1, HNc = 0;
      uint64_t D0_prev_word_old = 0, PM_prev_word = 0;
      const uint64_t* pm_row = peq + (int)p[j] * words;
      for (int w = 0; w < words; ++w) {
        const uint64_t PM = pm_row[w];
        const uint64_t d0_old = D0[w];
        const uint64_t TR = ((((~d0_old) & PM) << 1) | (w ? (((~D0_prev_word_old) & PM_prev_word) >> 63) : 0)) & PMo[w];
        const uint64_t X = PM | HNc;
        uint64_t d0 = (((X & VP[w]) + VP[w]) ^ VP[w]) | X | VN[w] | TR;
        uint64_t HP = VN[w] | ~(d0 | VP[w]);
        uint64_t HN = d0 & VP[w];
        if (w == words - 1) { dist += (HP & last) != 0; dist -= (HN & last) != 0; }
        const uint64_t HPc_in = HPc, HNc_in = HNc;
        HPc = HP >> 63; HNc = HN >> 63;
        HP = (HP << 1) | HPc_in;
        HN = (HN << 1) | HNc_in;
        VP[w] = HN | ~(d0 | HP);
        VN[w] = HP & d0;
        D0_prev_word_old = d0_old;
        PM_prev_word = PM;
        D0[w] = d0;
        PMo[w] = PM;
      }
    }
  }
  return dist;
}

/* distance(pattern, text); m may be 0. */
OSA_HD int osa_distance(const uint64_t* peq, int m, int words, const OsaText& t) {
  const int n = t.len[0] + t.len[1] + t.len[2];
  if (m == 0) return n;
  if (n == 0) return m;
  if (words == 1) return osa_distance_1w(peq, m, t);
  return osa_distance_mw(peq, m, words, t);
}
