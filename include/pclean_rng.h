/*
 * pclean_rng.h — the random-number CONTRACT of the boundary.
 *
 * The reference draws from Julia's global RNG (no seeds anywhere), so bit-exact
 * RNG parity with it is neither possible nor meaningful.  The contract instead is "same
 * uniforms ⇒ same choices": every random decision on the path is a pure function of
 * (seed, sweep, class, row key, particle, block, site vertex, purpose, sub-counter) through
 * Philox4x32-10, so the CPU oracle (sequential) and the CUDA engine (row-parallel) consume
 * identical numbers regardless of execution order.  Header-only, host + device.
 */
#ifndef PCLEAN_RNG_H
#define PCLEAN_RNG_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PCLEAN_HD __host__ __device__ __forceinline__
#else
#define PCLEAN_HD static inline
#endif

/* purposes */
enum {
  PCLEAN_RNG_ENUM = 1,      /* categorical draw inside a compiled enumeration (proposal_compiler.jl:119,237) */
  PCLEAN_RNG_PRIOR = 2,     /* discrete_proposal draw in propose_non_enumerable! (block_proposal.jl:47)      */
  PCLEAN_RNG_RANDOM = 3,    /* random(dist, args...) (block_proposal.jl:60)                                 */
  PCLEAN_RNG_FKPRIOR = 4,   /* CRP prior draw for an unconstrained reference slot (block_proposal.jl:80)     */
  PCLEAN_RNG_RESAMPLE = 5,  /* multinomial resampling (row_inference.jl:99)                                  */
  PCLEAN_RNG_FINAL = 6,     /* final particle selection / MH accept (row_inference.jl:162,164)               */
  PCLEAN_RNG_PARAM = 7,     /* resample_value!(parameter)                                                    */
  PCLEAN_RNG_PY = 8,        /* resample_py_params! (trace.jl:65-107)                                         */
  PCLEAN_RNG_PARAM_INIT = 9 /* initialize_parameter / first param_value                                      */
};

typedef struct pclean_rng_key {
  uint64_t seed;
  uint32_t sweep;     /* 0 = initialize_trace, 1.. = sweeps */
  uint32_t cls;
  int64_t  row;       /* row key (or parameter slot / class id for PARAM / PY) */
  uint32_t particle;
  uint32_t block;
  uint32_t site;      /* vertex id of the choice being drawn (0 if n/a) */
  uint32_t purpose;
} pclean_rng_key;

PCLEAN_HD void pclean_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The idx-th uniform in [0,1) of the stream identified by `k` (53-bit resolution). */
PCLEAN_HD double pclean_uniform(const pclean_rng_key* k, uint32_t idx) {
  uint32_t c0 = (uint32_t)(uint64_t)k->row;
  uint32_t c1 = (uint32_t)((uint64_t)k->row >> 32) ^ (k->particle << 8) ^ (k->block << 24);
  uint32_t c2 = (k->site & 0xFFFFu) | ((k->cls & 0xFFu) << 16) | ((k->purpose & 0xFFu) << 24);
  uint32_t c3 = idx;
  uint32_t k0 = (uint32_t)k->seed ^ (k->sweep * 0x9E3779B1u);
  uint32_t k1 = (uint32_t)(k->seed >> 32) ^ 0x85EBCA77u;
  uint32_t o[4];
  pclean_philox4x32_10(c0, c1, c2, c3, k0, k1, o);
  uint64_t hi = (uint64_t)(o[0] >> 5), lo = (uint64_t)(o[1] >> 6);   /* 27 + 26 bits */
  return (double)((hi << 26) | lo) * (1.0 / 9007199254740992.0);
}

/* A tiny sequential view over one stream, for samplers that need several uniforms. */
typedef struct pclean_stream { pclean_rng_key key; uint32_t idx; } pclean_stream;

PCLEAN_HD double pclean_next(pclean_stream* s) { return pclean_uniform(&s->key, s->idx++); }

PCLEAN_HD double pclean_next_normal(pclean_stream* s) {   /* Box–Muller, one value per call */
  double u1 = pclean_next(s), u2 = pclean_next(s);
  if (u1 < 1e-300) u1 = 1e-300;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

PCLEAN_HD double pclean_next_gamma(pclean_stream* s, double shape) {   /* Marsaglia–Tsang, scale 1 */
  double boost = 1.0;
  if (shape < 1.0) {
    double u = pclean_next(s);
    if (u < 1e-300) u = 1e-300;
    boost = pow(u, 1.0 / shape);
    shape += 1.0;
  }
  double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int it = 0; it < 1000; ++it) {
    double x = pclean_next_normal(s);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = pclean_next(s);
    if (u < 1e-300) u = 1e-300;
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
  }
  return boost * d;
}

PCLEAN_HD double pclean_next_beta(pclean_stream* s, double a, double b) {
  double x = pclean_next_gamma(s, a), y = pclean_next_gamma(s, b);
  return x / (x + y);
}

#endif /* PCLEAN_RNG_H */
