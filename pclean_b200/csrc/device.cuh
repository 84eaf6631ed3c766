// device.cuh — device-side data model and kernels of the sweep engine (sm_100a).
//
// Layout in HBM (all SoA, column-major):
//   dictionary      : uint8 symbol ids (codepoints compacted to an alphabet <= 256), offsets
//   observations    : per dataset column  int32 sid[N], int32 uobs[N] (index into the column's
//                     unique-string list), int32 ulist[U]
//   latent tables   : per class  int32 cells[vertex][cap] (string id | slot of referenced row),
//                     int32 refcnt[cap], f64 logcnt[cap] = log(refcnt - discount)
//   distance matrices: uint8 D[u][element]  — OSA edit distance between the u-th unique observed
//                     string of a column and the element's clean string (element = table slot or
//                     option index); this is the device form of the reference's memo dictionary
//                     add_typos_density_dict (add_typos.jl:47).  AddTypos log-densities are
//                     evaluated from (distance, clean length) with fp64 table arithmetic that is
//                     bit-identical to the oracle's formula.
//   particles       : int32 choice[block][N][K], f64 weight[N][K] (row-major: a warp moves one row, lane = particle)
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <math.h>

#include "../../include/pclean_b200.h"
#include "../../include/pclean_rng.h"
#include "osa_bitpar.cuh"

namespace pcl {

#ifndef PCL_NI1
#define PCL_NI1 __noinline__
#endif
#ifndef PCL_NI2
#define PCL_NI2 __noinline__
#endif
#ifndef PCL_NI3
#define PCL_NI3 __noinline__
#endif
#ifndef PCL_NI4
#define PCL_NI4 __noinline__
#endif
#ifndef PCL_PRUNED_INLINE
#define PCL_PRUNED_INLINE __noinline__
#endif
#define PCL_MAX_STARS 24
#define PCL_MAX_TERMS 40
#define PCL_MAX_EX 8
#define PCL_NEWSTR_MAX 64               /* longest string random(StringPrior) may generate here */
#define PCL_MAX_K 64                   /* particles per row: two passes of 32 lanes */
#define PCL_LG_N 320
#define PCL_WARPS_PER_CTA 8
#define PCL_KB_WARPS 16                  /* k_block: 16 warps per CTA share one score table; 2 CTAs per SM = 32 warps at <= 64 registers */
#define PCL_LUT_N 64                 /* smem score table covers distance, length < 64 */
#define PCL_EXP_CUTOFF (-50.0)       /* exp(x) for x below this is dropped from sums (< 2e-22 relative) */
#define PCL_SURV_MAX 96            /* candidates that survive pruning, per star and row */
#define PCL_PRUNE_MARGIN 45.0      /* skipped candidates are below e^-45 of the best one */
#define PCL_TYPO_COST 3.93         /* every typo costs at least this many nats */
#define PCL_NEG_INF (-CUDART_INF)
#define PCL_CHOICE_NEW_BASE (-2)     /* choice = -(pool_idx + 2) encodes a proposed new row */
#define PCL_UNSET (-3)                   /* table / scratch cell without a value */
#define PCL_CHOICE_UNSET (-2147483647 - 1) /* particle choice not made (distinct from every new-row handle) */


struct MatD { const uint8_t* d; long long stride; const uint8_t* elen; };

struct RefCellD { int block, col, table; };   // a cell of the referring observation row: tables[table].cells[col][assign[block][r]]

struct TermD {
  int kind;        // TERM_CAND / TERM_OPT / TERM_JOIN_CAND / TERM_JOIN_OPT / TERM_JOIN_INLINE
  int obs_col;     // dataset column index
  int mat;         // matrix index (non-join) or join table id (join)
  int max_typos;
  int external;    // summed over the observation rows referring to the latent row
  int a_kind, a_ref, b_kind, b_ref, sep;   // TERM_JOIN_INLINE: a ++ sep ++ b (OP_* operand kinds)
  RefCellD a_cell, b_cell;
  int ptable, pcol;  // candidate terms: the table column holding the clean string (pruning order), else -1
  int grp;           // latent programs, external string terms: id of the referrer group set (distinct observed strings per latent row), else -1
  int lmat;          // >= 0: the term reads the per-list distance blocks Dev.lmats[lmat] (choice over a row-dependent option list), mat = -1
};

#define PCL_MAX_INNER_CH 3
#define PCL_MAX_LOCAL 4                  /* local cells of an observation row written by one block (inner choices or sampled MaybeSwap cells) */
struct InnerArgD { int kind, ref; };        // ARG_*: ref = value id | dataset column | table column | - | inner choice index
struct InnerChoiceD { int vertex; int list_off; int n; };      // uniform choice over values innervals[list_off .. +n)
struct InnerGaussD { int obs_col; int func; int nargs; InnerArgD args[4]; double mean_const; double stdev; InnerArgD xform; };
struct InnerConstD { int kind; int obs_col; int optmap; int logp_off; double value; };
struct InnerD { int nchoice, ngauss, nconst; InnerChoiceD ch[PCL_MAX_INNER_CH]; InnerGaussD g[2]; InnerConstD c[3]; };
// tabulated function: open-addressing table keyed by up to 3 value ids
struct LookupD { const int* keys; const int* vals; unsigned mask; int nkey; };
// Distances of a choice whose option list is looked up from a cell of the row (rents: possibilities[countykey]).
// One block per list — (unique observed strings that occur together with the list) x (its options, then the
// dummy placeholder) — instead of one matrix over the union of all lists, which is quadratic in the number of
// distinct strings although a string only ever meets the list of its own key.  rows: (unique-string index, list)
// -> block row; a pair the dataset never showed is scored with an inline DP.
struct ListMatD { const uint8_t* d; const long long* row_off; const uint8_t* elen; const long long* elen_off; LookupD rows; };

// where a value comes from when an observation row is read outside a row move (statistics,
// external likelihoods of latent moves)
struct TraceArgD { int kind, a, b, c; };   // 0 constant (a = value id) | 1 observed / local cell of the row (a = dataset column or -1, b = vertex) | 2 table cell reached from the row (a = block, b = table, c = column) | 3 the enumerated option | 4 column a of the enumerated candidate | 5 column a of the latent row being moved
// Gaussian likelihood of a referring row inside a latent move (ExternalLikelihoodNode of a TransformedGaussian)
struct GaussExtD { int obs_col; int lookup; int nargs; TraceArgD args[3]; TraceArgD xform; double stdev; };

// MaybeSwap(val, options, prob) terms (maybe_swap.jl:13-28): flights
struct LookupRefD { int lookup; int nargs; TraceArgD args[3]; };      // lookup < 0: none
struct MswapD {
  int obs_col, vertex;          // dataset column (or -1) and observation-class vertex of the MaybeSwap node
  int val_kind;                 // 0: cell of the row an earlier block chose (val_cell / val_vertex) | 1: the enumerated option
  RefCellD val_cell; int val_vertex;
  int list_const; LookupRefD list;
  int prob_kind; double prob_const; int prob_slot; LookupRefD prob;    // 0 constant | 1 parameter slot | 2 lookup (slot >= 0, or constant -2-k)
};
// cell of a new row sampled from its discrete proposal when the row is created (block_proposal.jl:42-60)
struct FillD { int vertex; int dist; int list_const; LookupRefD list; int dummy_sid; };

struct StarD {
  int kind, vertex, parent, table, tvertex;
  int term0, nterm;
  int child0, nchild;         // into children[]
  int hoist;                  // >=0: value = hoist_val[hoist][uobs] (single-term choice star)
  int hoist_col;              // dataset column whose uobs indexes the hoist array
  int nopt, has_dummy;        // choice stars (nopt includes the dummy)
  int prior_off;              // into prior_pool (choice stars)
  int opt_off;                // into optsid_pool: string id per option (dummy placeholder last)
  int copy0, ncopy;           // into copies[] (pairs: obs-class vertex, table column)
  int bucket, bucket_col, bucket_obs_col;   // @guaranteed hash-bucket enumeration: candidates = rows whose key equals the observed one
  int list_func, list_obs_col;              // option list looked up from an observed value (lists pool), splp = per-string prior
  int list_own_col;                         // latent moves: the lookup key is this column of the row being moved (-1: dataset column)
  int splp_off;                              // into splp_pool: StringPrior log-density of every dictionary string for this star's (min, max)
  int univ_off;                              // into univ_col: matrix column of every dictionary string in this star's option universe
  int inner_elems, inner_new;               // into inners[] or -1
  int fill0, nfill;                         // into fills[]: cells of a new row sampled from their proposals
  int has_eq;                               // carries equality terms: elements are scored one by one (the 4-wide path knows distance terms only)
  int dummy_time;                           // the dummy option is replaced by TimePrior.random (a string of the pre-interned time table)
  int sp_min, sp_max;                       // StringPrior stars: length range (random() draws of the dummy)
  int optidx_off;                           // latent choice stars over a constant list: into optmap_pool, option index of every dictionary string (-1: not an option); else -1
};

#define PCL_MAX_SITES 12
struct ProgD {
  int latent, cls, nroots, roots[PCL_MAX_SITES];   // latent-class programs: one root per independent site
  int nstar, root, norder;
  int order[PCL_MAX_STARS];
  int star0, term0;            // offsets of this program's stars/terms in the global arrays
  int n_local; int local_vertex[PCL_MAX_LOCAL];   // observation-class choices enumerated inside elements (rents: br, unit)
  int base_prog;               // first program of this missingness pattern (programs of a pattern are consecutive per block)
  int ms0, n_rterm, n_rsamp;   // rootless blocks: mswaps[ms0 .. +n_rterm) observed terms, then n_rsamp sampled cells (local_vertex order)
  int n_earlier;               // 0 or 1 particle-dependent input
  int earlier_vertex, earlier_block, earlier_col, earlier_table;
  int nterm;
};

struct TableD {
  int* cells;                  // [n_normal][cap]
  int* refcnt;                 // [cap]
  double* logcnt;              // [cap]
  double* logcnt1;             // [cap] log(count - 1 - discount): what the row that holds one of the references sees (its own removed)
  const long long* keys;       // [cap] row keys (RNG streams of latent-row moves are keyed by row key)
  uint8_t* alive;              // [cap] 1 = referenced row (packed for the SIMD pruning pass)
  double max_logcnt;           // max over live slots of logcnt (upper bound of the CRP term)
  double log_new, log_den;     // log(strength + discount * n_alive), log(total_refs + strength): what a row without exclusions on this table sees
  double log_new_x[PCL_MAX_EX + 1], log_den_x[PCL_MAX_EX + 1];   // the same with x rows / x references of the moving row removed
  int cap, n_slots, n_normal;
  long long total_refs;
  int n_alive;
  double strength, discount;
  int nfk; int fk_col[4]; int fk_table[4];
  int* div;                    // [n_normal] distinct values per column among the live rows (hashed estimate, <= 65536): pruning order
};

struct Dev {
  // dictionary
  const uint8_t* sym; const int* str_off; const int* str_len; int n_strings;
  // score tables
  const double* LG; const double* LOGN; const double* LUT;   // LUT[L * 64 + k] = AddTypos score
  // observation class
  long long N; int n_cols; int nvC;
  int* const* uobs;            // [n_cols] -> int32[N]
  int* const* ulist;           // [n_cols] -> int32[U] string id of each unique observed string
  // latent-class sweep working set (valid during a class sweep)
  const int* lref_off; const int* lref_rows;   // CSR: observation rows referring to each slot of the class
  // referrers of a latent row grouped by what they observe (valid during a class sweep): per group set g,
  // sorted distinct keys (slot << 44 | other-half string id + 1 << 22 | unique observed string + 1) with their multiplicities
  const unsigned long long* const* lgrp_key; const int* const* lgrp_cnt; const int* lgrp_n;
  int* lchoice;                // [PCL_MAX_SITES][cap] selected particle's element per site (or new-row handle)
  int* lsel;                   // [cap]
  double* llogml;              // [cap]
  int* lflags;                 // [cap]
  // programs
  const ProgD* progs; const StarD* stars; const TermD* terms; const int* children; const int2* copies;
  const MatD* mats;
  const int* join_mat;         // [n_join][max_a]: matrix index per (join term id, a slot), -1 = not built
  int max_a;
  const int* a_slot_of_sid;    // [n_strings] dense slot of an earlier-block string value, -1 = unknown
  const double* prior_pool; const int* optsid_pool;
  const InnerD* inners; const LookupD* lookups; const int* innervals;   // inner enumerations, tabulated functions, their value lists
  const ListMatD* lmats;       // per-list distance blocks of row-dependent option lists
  const GaussExtD* gext;       // Gaussian external terms of latent programs
  const MswapD* mswaps; const FillD* fills; const double* lkconst;   // MaybeSwap terms, new-row fill-ins, real constants returned by lookups
  const uint8_t* time_ok;      // [n_strings] 1 = matches TimePrior's pattern (time_prior.jl:8-14)
  const int* time_sid;         // [12 * 60 * 2] string id of "h:m a.m." / "h:m p.m." (TimePrior.random, time_prior.jl:20-22)
  const double* param_real;    // current value of every real-valued parameter slot (MeanParameter)
  const double* xform_scale;
  double* const* obs_real;     // [n_cols] -> f64[N] (real-valued dataset columns) or nullptr
  int* const* obs_sid;         // [n_cols] -> int32[N] string id of the observed cell (-1 missing)
  const int* vcol;             // [nvC] dataset column of an observation-class vertex, -1 = none
  const int* lists_off; const int* lists_sid;   // string lists of the model (row-dependent option lists)
  const double* splp_pool; const int* univ_col; const int* optmap_pool;   // per-dictionary-string side tables
  const int* const* bkt_off; const int* const* bkt_slots;   // [n_tables] hash-bucket CSR by key string id
  int* const* rowcell;         // [nvC] -> int32[N] local discrete cells of the observation rows (or nullptr)
  int* const* pinner;          // [n_blocks] -> int32[K][N] packed inner choices of each particle
  double* const* hoist_val;    // [n_hoist] -> double[U]
  TableD* tables;
  // particles
  int K, n_blocks;
  int* const* assign;          // [n_blocks] -> int32[N] current slot per row
  int* const* pchoice;         // [n_blocks] -> int32[K][N]
  double* pweight;             // [K][N]
  double* plogml;              // [N]
  int* sel;                    // [N]
  double* row_logml;           // [N]
  int* row_flags;              // [N]
  unsigned long long* row_bad; // [N] bit k: particle k carries a placeholder (StringPrior dummy) or lost its state: scored, but never selected
  int* pool; int pool_cap; int* pool_count;   // new-row scratch: int32[pool_cap][nvC]
  int* needed_a;               // [n_strings] flag: join matrices needed for this a value
  int* needed_any;             // set when some needed_a flag was raised (the host reads the list only then)
  // random(StringPrior) (string_prior.jl:27-38): letter model, dictionary symbol of each of its 28 letters,
  // and the pool the strings generated during a launch go to (id = newstr_base + index; the host interns them)
  const double* lm_uni; const double* lm_big; const int* lm_sym;
  uint8_t* newstr_chars; int* newstr_len; int* newstr_count; int newstr_cap; int newstr_base;
  int* err;                    // device error word
  int* dbg;                    // [32] debug counters (which path the latent pruning took), or nullptr
  // star-marginal memo (mask 0 = disabled): [0] entries valid for one launch (reference-table stars:
  // they depend on the counts), [1] entries of choice stars, which depend only on option lists,
  // priors and distance matrices and persist until one of those changes
  unsigned long long* memo_keys[2]; ulonglong2* memo_vals[2]; unsigned memo_mask;     // vals: (value bits, high half of the key)
  int prune;                   // 1: integer-bound pruning of far candidates (default), 0: exact path only
  int opts;                    // PCL_OPT_* switches (A/B measurements, tests)
  const int* term_order;       // [n_terms] per star: its terms in the order the pruning pass reads them (most selective first)
  const float* col_meanlen;    // [n_cols] mean length of the observed strings of a dataset column
  const long long* row_order;  // optional processing order of the rows (L2 reuse), or nullptr
};
enum { PCL_OPT_PROGRESSIVE = 1, PCL_OPT_PMEMO = 2, PCL_OPT_FASTEXCL = 4, PCL_OPT_PARHINT = 8, PCL_OPT_LAZYNEW = 16 };
// particle arrays are row-major: the K particles of a row are contiguous (one warp moves one row, lane = particle)
#define PCL_PK(E_, k_, r_) ((long long)(r_) * (E_).K + (k_))
#define PCL_PINNER(E_, q_, k_, r_) (((long long)(r_) * PCL_MAX_LOCAL + (q_)) * (E_).K + (k_))

#define PCL_ALIGN16(n) (((n) + 15) / 16 * 16)
#define PCL_KBLOCK_SMEM_W(W_) (PCL_OFF_W + (W_) * sizeof(WarpState))
#define PCL_KBLOCK_SMEM PCL_KBLOCK_SMEM_W(PCL_KB_WARPS)

enum { ROWFLAG_DUMMY = 1, ROWFLAG_NOJOIN = 2, ROWFLAG_POOL = 4, ROWFLAG_CHANGED = 8 };

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// AddTypos log-density from (distance, clean length): add_typos.jl:58-64 with the same
// operation order as the oracle (oracle/pclean_oracle.cpp addtypos_score) — no FMA contraction.
__device__ __forceinline__ double addtypos_score(int k, int L, int max_typos, const double* LG, const double* LOGN) {
  if (max_typos >= 0 && k > max_typos) return -1e5;
  const int r = (L + 4) / 5;
  double l = __dsub_rn(LG[k + r], LG[k + 1]);
  l = __dsub_rn(l, LG[r]);
  l = __dadd_rn(l, __dmul_rn((double)r, -0.10536051565782630123));   // log(0.9)
  l = __dadd_rn(l, __dmul_rn((double)k, -2.30258509299404568402));   // log(0.1)
  l = __dsub_rn(l, __dmul_rn(LOGN[L], (double)k));
  l = __dsub_rn(l, __dmul_rn(__dmul_rn(3.25809653802148204862, (double)k), 0.5));   // log(26) k / 2
  return l;
}

// out-of-table (distance or length >= 64) scores are rare: keep their code out of the hot loops
__device__ __noinline__ double score_slow(int k, int L, const double* LG, const double* LOGN) { return addtypos_score(k, L, -1, LG, LOGN); }
__device__ __forceinline__ double score_fast(int k, int L, int max_typos, const double* LG, const double* LOGN, const double* LUT) {
  if (max_typos >= 0 && k > max_typos) return -1e5;
  if ((k | L) < PCL_LUT_N) return LUT[L * PCL_LUT_N + k];
  return score_slow(k, L, LG, LOGN);
}

// exp / log are ~100 SASS instructions each when inlined; the row kernels call them from dozens of
// sites (only for the few candidates that survive pruning), so they live out of line: the
// kernels are bound by instruction fetch, not by these calls (profiles/kblock_r1_summary.md)
__device__ __noinline__ double exp_nl(double x) { return exp(x); }
__device__ __noinline__ double log_nl(double x) { return log(x); }
struct Lse { double m, s; };
__device__ __forceinline__ void lse_add(Lse& a, double x) {
  if (x == PCL_NEG_INF) return;
  if (x > a.m) { a.s = ((a.m == PCL_NEG_INF || a.m - x < PCL_EXP_CUTOFF) ? 0.0 : a.s * exp_nl(a.m - x)) + 1.0; a.m = x; }
  else { const double d = x - a.m; if (d > PCL_EXP_CUTOFF) a.s += exp_nl(d); }
}
// warp-wide log-sum-exp of the per-lane partial (max, sum) pairs: butterfly max, one rescale per
// lane, butterfly sum (every lane ends with the same bits)
__device__ __noinline__ double lse_warp(Lse a) {
  double M = a.m;
  #pragma unroll
  for (int o = 16; o; o >>= 1) M = fmax(M, shfl_xor_d(M, o));
  if (M == PCL_NEG_INF) return PCL_NEG_INF;
  double sum = a.m == PCL_NEG_INF ? 0.0 : (a.m == M ? a.s : a.s * exp_nl(a.m - M));
  #pragma unroll
  for (int o = 16; o; o >>= 1) sum += shfl_xor_d(sum, o);
  return M + log_nl(sum);
}

// per-warp working state (shared memory)
struct WarpState {
  double V[PCL_MAX_STARS];        // marginal of each star for the current upstream state
  double aux[PCL_MAX_STARS];      // per-star per-row scalar (dummy mass of a row-dependent option list)
  int lst[PCL_MAX_STARS];         // per-star per-row option list id (row-dependent lists), -1 otherwise
  int bkt0[PCL_MAX_STARS], bktn[PCL_MAX_STARS];   // per-star hash bucket of this row: first entry / size
  int u[PCL_MAX_TERMS];           // unique-obs index per term (-1 = explicit missing)
  int tmat[PCL_MAX_TERMS];        // resolved matrix per term for the current upstream state
  int ex_table[PCL_MAX_EX], ex_slot[PCL_MAX_EX], ex_gc[PCL_MAX_EX];
  int n_ex;
  const uint8_t* rowp[PCL_MAX_TERMS];   // distance-matrix row of each term for this row (nullptr = missing)
  const uint8_t* elenp[PCL_MAX_TERMS];  // clean-string length per element of each term's matrix
  int sv_star;                          // star whose survivors are currently in sv_* (-1 none)
  const uint8_t* act[PCL_MAX_TERMS];    // compacted non-missing row pointers of the star being pruned
  int nact;
  int sv_idx[PCL_SURV_MAX + 1];         // surviving elements (ascending), new-row branch last
  double sv_ll[PCL_SURV_MAX + 1];
  double sv_cs[PCL_SURV_MAX + 1];       // running sum of the survivors' probabilities (inverse-CDF draws search it)
  int sv_n;
  unsigned long long mk_hi[PCL_MAX_STARS];   // memo entries this row owns (claimed, to be published): high key half,
  int mk_slot[PCL_MAX_STARS];                // slot (-1: none) and table, indexed like P.order
  int mk_tbl[PCL_MAX_STARS];
  int glo[PCL_MAX_TERMS], ghi[PCL_MAX_TERMS];   // latent moves: this row's range of referrer groups per term (glo == ghi: none / term not grouped)
  long long row;                             // the row being moved (observation row index / latent slot)
  const int* refs; int nref;                 // latent moves: the observation rows referring to the row (nref < 0: observation-class move)
  int lazy_ok, lazy_fail;                    // root-first attempts of this warp that dropped the new-row branch / had to evaluate the children after all
};

// The row context.  Everything a phase needs beside its arguments sits at a fixed place in the CTA's
// dynamic shared memory — score tables, a copy of the launch descriptor and of the program descriptor
// (staged there by the kernels), the per-warp state, which also holds the row being moved — so the
// context types carry no data: a phase derives the pointers from the shared-memory base and its warp
// index.  They are LDS/STS with constant offsets; nothing is reloaded from a context in local memory
// and no store through them can alias anything the compiler keeps in registers.
extern __shared__ __align__(16) unsigned char pcl_smem[];
#define PCL_OFF_LG ((size_t)PCL_LUT_N * PCL_LUT_N * sizeof(double))
#define PCL_OFF_LOGN (PCL_OFF_LG + PCL_LG_N * sizeof(double))
#define PCL_OFF_DEV (PCL_OFF_LOGN + 256 * sizeof(double))
#define PCL_OFF_PROG (PCL_OFF_DEV + PCL_ALIGN16(sizeof(Dev)))
#define PCL_OFF_W (PCL_OFF_PROG + PCL_ALIGN16(sizeof(ProgD)))
#define PCL_CTX(c) \
  WarpState* const cW = reinterpret_cast<WarpState*>(pcl_smem + PCL_OFF_W) + (threadIdx.x >> 5); \
  const Dev* const cE = reinterpret_cast<const Dev*>(pcl_smem + PCL_OFF_DEV); \
  const ProgD* const cP = reinterpret_cast<const ProgD*>(pcl_smem + PCL_OFF_PROG); \
  const double* const cLUT = reinterpret_cast<const double*>(pcl_smem); \
  const double* const cLG = reinterpret_cast<const double*>(pcl_smem + PCL_OFF_LG); \
  const double* const cLOGN = reinterpret_cast<const double*>(pcl_smem + PCL_OFF_LOGN); \
  const int cLane = threadIdx.x & 31; \
  (void)c; (void)cW; (void)cE; (void)cP; (void)cLG; (void)cLOGN; (void)cLUT; (void)cLane
#define cR (cW->row)
#define cRefs (cW->refs)
#define cNref (cW->nref)
/* k_latent keeps per-warp match masks for inline joins behind the warp states */
#define cPeq (reinterpret_cast<unsigned long long*>(reinterpret_cast<WarpState*>(pcl_smem + PCL_OFF_W) + PCL_WARPS_PER_CTA) + (threadIdx.x >> 5) * 256)
__device__ __forceinline__ int lookup_ref(const Dev& E, const LookupRefD& L, long long r, int esid);   // defined with the trace readers below
struct RowCtx { static constexpr bool rich = true; };
// programs without hash buckets, row-dependent option lists, inner enumerations or equality terms
// (hospital): the same code with those branches folded away, which keeps k_block's registers
struct LeanCtx : RowCtx { static constexpr bool rich = false; };

__device__ __forceinline__ int excl_count(const WarpState* W, int table, int slot) {
  int c = 0;
  #pragma unroll 1
  for (int i = 0; i < W->n_ex; ++i) c += (W->ex_table[i] == table && W->ex_slot[i] == slot);
  return c;
}
__device__ __forceinline__ int excl_refs(const WarpState* W, int table) {
  int c = 0;
  #pragma unroll 1
  for (int i = 0; i < W->n_ex; ++i) c += (W->ex_table[i] == table);
  return c;
}
__device__ __forceinline__ int excl_rows(const WarpState* W, int table) {
  int c = 0;
  #pragma unroll 1
  for (int i = 0; i < W->n_ex; ++i) c += (W->ex_table[i] == table && W->ex_gc[i]);
  return c;
}


// ---- tabulated functions (JuliaNode closures tabulated by the host), inner enumerations ------
__device__ __forceinline__ unsigned long long hmix(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
#define PCL_LOOKUP_EMPTY (-2147483647 - 1)
__device__ __forceinline__ int lookup_find(const LookupD& L, int k0, int k1, int k2) {
  const unsigned long long key = hmix((unsigned long long)(unsigned)k0 * 0x9E3779B97F4A7C15ULL ^ ((unsigned long long)(unsigned)k1 << 20) ^ ((unsigned long long)(unsigned)k2 << 41));
  unsigned h = (unsigned)key & L.mask;
  for (int p = 0; p < 64; ++p, h = (h + 1) & L.mask) {
    const int a = L.keys[3 * h];
    if (a == PCL_LOOKUP_EMPTY) return PCL_LOOKUP_EMPTY;
    if (a == k0 && L.keys[3 * h + 1] == k1 && L.keys[3 * h + 2] == k2) return L.vals[h];
  }
  return PCL_LOOKUP_EMPTY;
}

__device__ __noinline__ int osa_plain(const uint8_t* A, int m, const uint8_t* B, int nb);    // plain two-row DP, strings <= PCL_NEWSTR_MAX (defined below)
// row of the (unique observed string u, list l) pair in the list blocks of term matrix lm, or nullptr
__device__ __forceinline__ const uint8_t* lmat_row(const Dev* E, int lm, int u, int l) {
  const ListMatD& M = E->lmats[lm];
  const int r = lookup_find(M.rows, u, l, 0);
  return r == PCL_LOOKUP_EMPTY ? nullptr : M.d + M.row_off[r];
}
// the distance a pair without a tabulated row costs: observed string (by id) against an option (by id)
__device__ __forceinline__ int lmat_inline(const Dev* E, int obs_sid, int opt_sid) {
  const int m = E->str_len[obs_sid], nb = E->str_len[opt_sid];
  if (m > PCL_NEWSTR_MAX || nb > PCL_NEWSTR_MAX) { atomicExch(E->err, PCLEAN_ERR_UNSUPPORTED); return 255; }
  return min(255, osa_plain(E->sym + E->str_off[obs_sid], m, E->sym + E->str_off[opt_sid], nb));
}

struct ElemRef { int table; int slot; int esid; };    // the enumerated element: a table row or an option string

template <class C> __device__ __forceinline__ int inner_arg(const C& c, const InnerArgD& a, const ElemRef& e, const InnerD& I, const int* pick) { PCL_CTX(c);
  switch (a.kind) {
    case 0: return a.ref;                                                       // ARG_CONST
    case 1: return cE->obs_sid[a.ref][cR];                                    // ARG_OBS
    case 2: { const TableD& T = cE->tables[e.table]; return T.cells[(long long)a.ref * T.cap + e.slot]; }   // ARG_ELEM_COL
    case 3: return e.esid;                                                      // ARG_ELEM_OPT
    default: return cE->innervals[I.ch[a.ref].list_off + pick[a.ref]];         // ARG_INNER
  }
}

// log-likelihood of one combination of inner choices
template <class C> __device__ double inner_combo(const C& c, const InnerD& I, const ElemRef& e, const int* pick) { PCL_CTX(c);
  double lp = 0.0;
  for (int i = 0; i < I.nchoice; ++i) lp -= log((double)I.ch[i].n);
  for (int g = 0; g < I.ngauss; ++g) {
    const InnerGaussD& G = I.g[g];
    double mean = G.mean_const;
    if (G.func >= 0) {
      int k[3] = {0, 0, 0};
      for (int a = 0; a < G.nargs && a < 3; ++a) k[a] = inner_arg(c, G.args[a], e, I, pick);
      const int slot = lookup_find(cE->lookups[G.func], k[0], k[1], k[2]);
      if (slot == PCL_LOOKUP_EMPTY) { atomicExch(cE->err, PCLEAN_ERR_LOOKUP); return PCL_NEG_INF; }
      mean = cE->param_real[slot];
    } else if (G.func == -2) mean = cE->param_real[(int)G.mean_const];
    const int xf = inner_arg(c, G.xform, e, I, pick);
    const double sc = cE->xform_scale[xf];
    const double x = cE->obs_real[G.obs_col][cR] * sc;
    const double z = (x - mean) / G.stdev;
    lp += -0.5 * z * z - log(G.stdev) - 0.91893853320467274178 - log(fabs(1.0 / sc));    // transformed_gaussian.jl:15-16
  }
  return lp;
}

// marginal over the inner choices (+ constant prior terms); with `u` != nullptr also samples the
// choices hierarchically (first choice from its marginal, then the next given it, ...), one
// uniform per choice site, exactly like the nested enumeration of the reference.
template <class C> __device__ double inner_eval(const C& c, const InnerD& I, const ElemRef& e, const double* u, int* picked) { PCL_CTX(c);
  double base = 0.0;
  for (int k = 0; k < I.nconst; ++k) {
    const InnerConstD& cp = I.c[k];
    if (cp.kind == 0) base += cp.value;
    else if (cp.kind == 2) { const int sid = cE->obs_sid[cp.obs_col][cR]; if (sid >= 0) base += cE->splp_pool[cp.logp_off + sid]; }
    else {
      const int sid = cE->obs_sid[cp.obs_col][cR];
      const int idx = sid >= 0 ? cE->optmap_pool[cp.optmap + sid] : -1;
      base += idx >= 0 ? cE->prior_pool[cp.logp_off + idx] : PCL_NEG_INF;    // ChooseProportionally.logdensity
    }
  }
  if (I.nchoice == 0 && I.ngauss == 0) return base;
  int pick[PCL_MAX_INNER_CH] = {0, 0, 0};
  const int n0 = I.nchoice > 0 ? I.ch[0].n : 1, n1 = I.nchoice > 1 ? I.ch[1].n : 1, n2 = I.nchoice > 2 ? I.ch[2].n : 1;
  // total
  Lse tot; tot.m = PCL_NEG_INF; tot.s = 0.0;
  for (pick[0] = 0; pick[0] < n0; ++pick[0]) for (pick[1] = 0; pick[1] < n1; ++pick[1]) for (pick[2] = 0; pick[2] < n2; ++pick[2])
    lse_add(tot, inner_combo(c, I, e, pick));
  const double L = tot.m == PCL_NEG_INF ? PCL_NEG_INF : tot.m + log(tot.s);
  if (u && picked) {
    int fix[PCL_MAX_INNER_CH] = {-1, -1, -1};
    for (int lvl = 0; lvl < I.nchoice; ++lvl) {
      const int nl = I.ch[lvl].n;
      double w[16]; double wt = PCL_NEG_INF;
      for (int i = 0; i < nl && i < 16; ++i) {
        Lse a; a.m = PCL_NEG_INF; a.s = 0.0;
        for (pick[0] = 0; pick[0] < n0; ++pick[0]) for (pick[1] = 0; pick[1] < n1; ++pick[1]) for (pick[2] = 0; pick[2] < n2; ++pick[2]) {
          bool ok = pick[lvl] == i;
          for (int q = 0; q < lvl; ++q) ok = ok && pick[q] == fix[q];
          if (ok) lse_add(a, inner_combo(c, I, e, pick));
        }
        w[i] = a.m == PCL_NEG_INF ? PCL_NEG_INF : a.m + log(a.s);
        wt = wt == PCL_NEG_INF ? w[i] : (w[i] == PCL_NEG_INF ? wt : fmax(wt, w[i]) + log1p(exp(-fabs(wt - w[i]))));
      }
      double cum = 0.0; int ch = -1, last = -1;
      for (int i = 0; i < nl && i < 16; ++i) {
        const double p = w[i] == PCL_NEG_INF ? 0.0 : exp(w[i] - wt);
        if (p > 0.0) last = i;
        cum += p;
        if (ch < 0 && u[lvl] < cum) ch = i;
      }
      fix[lvl] = ch >= 0 ? ch : last;
      picked[lvl] = fix[lvl];
    }
  }
  return base + L;
}

// number of enumerated elements of a star (excluding the new-row branch)
template <class C> __device__ __forceinline__ int star_index(const C& c, const StarD& s) { PCL_CTX(c); return (int)(&s - (cE->stars + cP->star0)); }
template <class C> __device__ __forceinline__ int star_nelem(const C& c, const StarD& s) { PCL_CTX(c);
  if (s.kind == 0) return (C::rich && s.bucket) ? cW->bktn[star_index(c, s)] : cE->tables[s.table].n_slots;
  if (C::rich && s.list_func >= 0) { const int l = cW->lst[star_index(c, s)]; return l >= 0 ? cE->lists_off[l + 1] - cE->lists_off[l] + 1 : 1; }
  return s.nopt;
}
// table slot of element j of an FK star (identity unless the star enumerates a hash bucket)
template <class C> __device__ __forceinline__ int star_slot(const C& c, const StarD& s, int j) { PCL_CTX(c);
  return (C::rich && s.bucket) ? cE->bkt_slots[s.table][cW->bkt0[star_index(c, s)] + j] : j;
}
// string id of option j of a choice star
template <class C> __device__ __forceinline__ int star_option_sid(const C& c, const StarD& s, int j) { PCL_CTX(c);
  if (!C::rich || s.list_func < 0) return cE->optsid_pool[s.opt_off + j];
  const int l = cW->lst[star_index(c, s)];
  const int n = l >= 0 ? cE->lists_off[l + 1] - cE->lists_off[l] : 0;
  return j < n ? cE->lists_sid[cE->lists_off[l] + j] : cE->optsid_pool[s.opt_off];       // last = dummy placeholder
}
// per-row preparation of a star: hash bucket / option list of this row, dummy mass of the list
template <class C> __device__ void star_prepare(const C& c, const StarD& s) { PCL_CTX(c);
  const int sidx = star_index(c, s);
  if (s.kind == 0 && C::rich && s.bucket) {
    if (cLane == 0) {
      const int key = cE->obs_sid[s.bucket_obs_col][cR];
      const int* off = cE->bkt_off[s.table];
      cW->bkt0[sidx] = key >= 0 ? off[key] : 0;
      cW->bktn[sidx] = key >= 0 ? off[key + 1] - off[key] : 0;
    }
  } else if (s.kind == 1 && s.list_func >= 0) {
    int l = -1;
    const int key = cE->obs_sid[s.list_obs_col][cR];
    if (key >= 0) { l = lookup_find(cE->lookups[s.list_func], key, 0, 0); if (l == PCL_LOOKUP_EMPTY) l = -1; }
    const int n = l >= 0 ? cE->lists_off[l + 1] - cE->lists_off[l] : 0;
    Lse a; a.m = PCL_NEG_INF; a.s = 0.0;
    for (int j = cLane; j < n; j += 32) lse_add(a, cE->splp_pool[s.splp_off + cE->lists_sid[cE->lists_off[l] + j]]);
    const double tot = lse_warp(a);
    if (cLane == 0) { cW->lst[sidx] = l; cW->aux[sidx] = log1p(-exp(tot)); }       // string_prior.jl:19-20
    // the terms of this star read the blocks of list l: row of each term's observed string
    const TermD* terms = cE->terms + cP->term0;
    for (int t = s.term0 + cLane; t < s.term0 + s.nterm; t += 32) {
      if (terms[t].lmat < 0) continue;
      const int u = cW->u[t];
      const ListMatD& M = cE->lmats[terms[t].lmat];
      cW->rowp[t] = (u >= 0 && l >= 0) ? lmat_row(cE, terms[t].lmat, u, l) : nullptr;
      cW->elenp[t] = l >= 0 ? M.elen + M.elen_off[l] : nullptr;
    }
  }
  __syncwarp();
}

// log-score of element j of star s for the current row / upstream state
template <class C> __device__ double star_elem(const C& c, const StarD& s, int j) { PCL_CTX(c);
  double l;
  ElemRef er; er.table = s.table; er.slot = -1; er.esid = -1;
  int col_index = j;                          // column of the distance matrices for this element
  if (s.kind == 0) {
    const int slot = star_slot(c, s, j);
    er.slot = slot; col_index = slot;
    const TableD& T = cE->tables[s.table];
    int cnt = T.refcnt[slot];
    if (cW->n_ex) {
      const int e = excl_count(cW, s.table, slot);
      if (e) { cnt -= e; l = cnt <= 0 ? PCL_NEG_INF : (e == 1 ? T.logcnt1[slot] : log_nl((double)cnt - T.discount)); }
      else l = T.logcnt[slot];
    } else l = T.logcnt[slot];
    if (cnt <= 0) return PCL_NEG_INF;
  } else if (C::rich && s.list_func >= 0) {
    const int sid = star_option_sid(c, s, j);
    er.esid = sid;
    const int n = star_nelem(c, s);
    l = j < n - 1 ? cE->splp_pool[s.splp_off + sid] : cW->aux[star_index(c, s)];
    col_index = cE->univ_col[s.univ_off + sid];
  } else {
    l = cE->prior_pool[s.prior_off + j];
    er.esid = cE->optsid_pool[s.opt_off + j];
  }
  const TermD* terms = cE->terms + cP->term0;
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) {
    if (C::rich && terms[t].kind == 5) {               // TERM_EQ: the candidate must agree with the observed cell (proposal_compiler.jl:282-291)
      const TableD& T = cE->tables[s.table];
      if (T.cells[(long long)terms[t].mat * T.cap + er.slot] != cE->obs_sid[terms[t].obs_col][cR]) return PCL_NEG_INF;
      continue;
    }
    const int u = cW->u[t];
    if (u < 0) continue;                         // explicit missing observation: log-density 0
    if (C::rich && terms[t].lmat >= 0) {         // per-list blocks: column = position in the row's list (the dummy last)
      const uint8_t* rp = cW->rowp[t];
      const int k = rp ? rp[j] : lmat_inline(cE, cE->ulist[terms[t].obs_col][u], er.esid);
      const int L = rp ? cW->elenp[t][j] : min(255, cE->str_len[er.esid]);
      l += score_fast(k, L, terms[t].max_typos, cLG, cLOGN, cLUT);
      continue;
    }
    const int k = cW->rowp[t][col_index];
    l += score_fast(k, cW->elenp[t][col_index], terms[t].max_typos, cLG, cLOGN, cLUT);
  }
  if (C::rich && s.inner_elems >= 0) l += inner_eval(c, cE->inners[s.inner_elems], er, nullptr, nullptr);
  return l;
}

// four consecutive elements j0..j0+3 (j0 % 4 == 0) with 32-bit loads of the distance / length bytes
template <class C> __device__ __forceinline__ void star_elem4(const C& c, const StarD& s, int j0, int J, double l[4]) { PCL_CTX(c);
  if (s.kind == 0) {
    const TableD& T = cE->tables[s.table];
    const int4 cnt4 = *reinterpret_cast<const int4*>(T.refcnt + j0);
    const double2 a = *reinterpret_cast<const double2*>(T.logcnt + j0), b = *reinterpret_cast<const double2*>(T.logcnt + j0 + 2);
    int cnt[4] = {cnt4.x, cnt4.y, cnt4.z, cnt4.w};
    l[0] = a.x; l[1] = a.y; l[2] = b.x; l[3] = b.y;
    for (int i = 0; i < cW->n_ex; ++i) {                 // at most a handful of (table, slot) exclusions per row
      const int q = cW->ex_slot[i] - j0;
      if (cW->ex_table[i] == s.table && q >= 0 && q < 4) {
        cnt[q] -= 1;
        l[q] = cnt[q] <= 0 ? PCL_NEG_INF : (cnt[q] + 1 == (&cnt4.x)[q] ? T.logcnt1[j0 + q] : log_nl((double)cnt[q] - T.discount));
      }
    }
    #pragma unroll
    for (int q = 0; q < 4; ++q) if (cnt[q] <= 0 || j0 + q >= J) l[q] = PCL_NEG_INF;
  } else {
    #pragma unroll
    for (int q = 0; q < 4; ++q) l[q] = j0 + q < J ? cE->prior_pool[s.prior_off + j0 + q] : PCL_NEG_INF;
  }
  const TermD* terms = cE->terms + cP->term0;
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) {
    const uint8_t* rp = cW->rowp[t];
    if (!rp) continue;
    const unsigned x = *reinterpret_cast<const unsigned*>(rp + j0);
    const unsigned L4 = *reinterpret_cast<const unsigned*>(cW->elenp[t] + j0);
    const int mt = terms[t].max_typos;
    #pragma unroll
    for (int q = 0; q < 4; ++q)
      l[q] += score_fast((x >> (8 * q)) & 255u, (L4 >> (8 * q)) & 255u, mt, cLG, cLOGN, cLUT);
  }
}

// log-score of the new-row branch of an FK star (without the common -log(n + s))
template <class C> __device__ __noinline__ double star_extra(const C& c, const StarD& s) { PCL_CTX(c);
  if (s.kind != 0) return PCL_NEG_INF;
  const TableD& T = cE->tables[s.table];
  const int xr = excl_rows(cW, s.table);
  double l = xr <= PCL_MAX_EX ? T.log_new_x[xr] : log_nl(T.strength + T.discount * (double)(T.n_alive - xr));      // same bits: k_table_stats evaluates the same expression
  const int* ch = cE->children + s.child0;
  for (int i = 0; i < s.nchild; ++i) l += cW->V[ch[i]];
  if (C::rich && s.inner_new >= 0) { ElemRef er; er.table = s.table; er.slot = -1; er.esid = -1; l += inner_eval(c, cE->inners[s.inner_new], er, nullptr, nullptr); }
  return l;
}
template <class C> __device__ __forceinline__ double star_logden(const C& c, const StarD& s) { PCL_CTX(c);
  if (s.kind != 0) return 0.0;
  const TableD& T = cE->tables[s.table];
  const int xf = excl_refs(cW, s.table);
  return xf <= PCL_MAX_EX ? T.log_den_x[xf] : log_nl((double)(T.total_refs - xf) + T.strength);
}

// LSE over all elements (+ extra), raw (before subtracting logden)
template <class C> __device__ PCL_NI4 double star_lse_raw(const C& c, const StarD& s) { PCL_CTX(c);
  const int J = star_nelem(c, s);
  const int J4 = (J + 3) & ~3;
  Lse acc; acc.m = PCL_NEG_INF; acc.s = 0.0;
  if (C::rich && (s.bucket || s.list_func >= 0 || s.inner_elems >= 0 || s.has_eq)) {       // irregular stars: scalar elements
    for (int j = cLane; j < J; j += 32) lse_add(acc, star_elem(c, s, j));
    if (cLane == 0) lse_add(acc, star_extra(c, s));
    return lse_warp(acc);
  }
  for (int j0 = cLane * 4; j0 < J4; j0 += 128) {
    double l[4];
    star_elem4(c, s, j0, J, l);
    #pragma unroll
    for (int q = 0; q < 4; ++q) lse_add(acc, l[q]);
  }
  if (cLane == 0) lse_add(acc, star_extra(c, s));
  return lse_warp(acc);
}

// ------------------------------------------------------------------------------------------
// Pruned evaluation.  Every AddTypos term satisfies score(k, L) <= -PCL_TYPO_COST * k
// (DESIGN.md §"pruning bound"), so a candidate's log-score is at most
//   Bmax - PCL_TYPO_COST * sum_t min(k_t, clamp)
// where Bmax bounds the CRP / prior term.  The byte sums are computed 4 candidates per 32-bit
// SIMD op straight from the distance rows; only candidates whose bound reaches within
// PCL_PRUNE_MARGIN nats of an exactly evaluated candidate get the fp64 evaluation.  The skipped
// mass is < n * e^-45 relative: far below the 1e-9 parity tolerance.
// ------------------------------------------------------------------------------------------
// Sums of the distance bytes of 16 consecutive candidates (j0 % 16 == 0) over the star's terms:
// one 128-bit load per term, bytes widened to 16-bit lanes with PRMT and added with plain
// integer adds (nterms * 255 < 65536: no clamping).  out[w] holds candidates j0+2w (low half)
// and j0+2w+1 (high half).  Dead slots / the tail beyond J get 0xFFFF.
#define PCL_ACC16(V_)                                                                         \
  out[0] += __byte_perm((V_).x, 0u, 0x4140); out[1] += __byte_perm((V_).x, 0u, 0x4342);           \
  out[2] += __byte_perm((V_).y, 0u, 0x4140); out[3] += __byte_perm((V_).y, 0u, 0x4342);           \
  out[4] += __byte_perm((V_).z, 0u, 0x4140); out[5] += __byte_perm((V_).z, 0u, 0x4342);           \
  out[6] += __byte_perm((V_).w, 0u, 0x4140); out[7] += __byte_perm((V_).w, 0u, 0x4342);

template <class C> __device__ __forceinline__ void star_sum16(const C& c, const StarD& s, const TableD* T, int j0, int J, unsigned out[8]) { PCL_CTX(c);
  #pragma unroll
  for (int w = 0; w < 8; ++w) out[w] = 0;
  // W->act[] = row pointers of the non-missing terms of this star, compacted by star_eval_pruned
  const int na = cW->nact;
  int t = 0;
  for (; t + 4 <= na; t += 4) {            // four independent 128-bit loads in flight per lane
    const uint4 x0 = *reinterpret_cast<const uint4*>(cW->act[t] + j0);
    const uint4 x1 = *reinterpret_cast<const uint4*>(cW->act[t + 1] + j0);
    const uint4 x2 = *reinterpret_cast<const uint4*>(cW->act[t + 2] + j0);
    const uint4 x3 = *reinterpret_cast<const uint4*>(cW->act[t + 3] + j0);
    PCL_ACC16(x0) PCL_ACC16(x1) PCL_ACC16(x2) PCL_ACC16(x3)
  }
  for (; t < na; ++t) {
    const uint4 x = *reinterpret_cast<const uint4*>(cW->act[t] + j0);
    PCL_ACC16(x)
  }
  if (T) {
    const uint4 a = *reinterpret_cast<const uint4*>(T->alive + j0);      // bytes 0/1
    const unsigned aw[4] = {a.x, a.y, a.z, a.w};
    #pragma unroll
    for (int w = 0; w < 4; ++w) {
      out[2 * w] |= (__byte_perm(aw[w], 0u, 0x4140) ^ 0x00010001u) * 0xFFFFu;
      out[2 * w + 1] |= (__byte_perm(aw[w], 0u, 0x4342) ^ 0x00010001u) * 0xFFFFu;
    }
  }
  if (j0 + 16 > J) {
    #pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (j0 + 2 * w >= J) out[w] |= 0x0000FFFFu;
      if (j0 + 2 * w + 1 >= J) out[w] |= 0xFFFF0000u;
    }
  }
}

// true if one of the 16 halfword sums in v[] is <= tau (all sums < 0x8000; tau2 = 0x8000 + tau in both halves)
__device__ __forceinline__ bool live16(const unsigned v[8], unsigned tau2) {
  unsigned m = 0;
  #pragma unroll
  for (int w = 0; w < 8; ++w) m |= (tau2 - v[w]);
  return (m & 0x80008000u) != 0u;
}

// Progressive form of star_sum16: the terms are read most selective first (W->act[] is in
// E.term_order), and a lane stops reading as soon as none of its 16 candidates can still come
// within `tau` (a partial sum only grows).  Returns whether this lane still holds live candidates;
// only then out[] holds their exact sums (dead slots / tail = 0xFFFF).  A random candidate is
// usually out after its first long term, so a stride costs ~1 instead of nterm 128-bit loads.
template <class C> __device__ __forceinline__ bool star_sum16_prog(const C& c, const TableD* T, int j0, int J, int J16, unsigned tau, unsigned out[8]) { PCL_CTX(c);
  #pragma unroll
  for (int w = 0; w < 8; ++w) out[w] = 0;
  const int na = cW->nact;
  const unsigned tau2 = (0x8000u + tau) * 0x00010001u;
  bool live = j0 < J16;
  int t = 0, step = 1;
  while (t < na) {
    if (live) {
      const uint4 x0 = *reinterpret_cast<const uint4*>(cW->act[t] + j0);
      if (step == 2 && t + 1 < na) {
        const uint4 x1 = *reinterpret_cast<const uint4*>(cW->act[t + 1] + j0);
        PCL_ACC16(x1)
      }
      PCL_ACC16(x0)
      live = live16(out, tau2);
    }
    t += step; step = 2;
    if (!__any_sync(0xffffffffu, live)) return false;
  }
  if (live) {
    if (T) {
      const uint4 a = *reinterpret_cast<const uint4*>(T->alive + j0);      // bytes 0/1
      const unsigned aw[4] = {a.x, a.y, a.z, a.w};
      #pragma unroll
      for (int w = 0; w < 4; ++w) {
        out[2 * w] |= (__byte_perm(aw[w], 0u, 0x4140) ^ 0x00010001u) * 0xFFFFu;
        out[2 * w + 1] |= (__byte_perm(aw[w], 0u, 0x4342) ^ 0x00010001u) * 0xFFFFu;
      }
    }
    if (j0 + 16 > J) {
      #pragma unroll
      for (int w = 0; w < 8; ++w) {
        if (j0 + 2 * w >= J) out[w] |= 0x0000FFFFu;
        if (j0 + 2 * w + 1 >= J) out[w] |= 0xFFFF0000u;
      }
    }
  }
  return live;
}

// element j of a plain star (prior / CRP term + distance terms only) with the terms spread over
// the lanes: one round of loads instead of nterm dependent ones.  Every lane returns the same bits.
template <class C> __device__ double star_elem_par(const C& c, const StarD& s, int j) { PCL_CTX(c);
  double l;
  if (s.kind == 0) {
    const TableD& T = cE->tables[s.table];
    int cnt = T.refcnt[j];
    const int e = cW->n_ex ? excl_count(cW, s.table, j) : 0;
    cnt -= e;
    if (cnt <= 0) return PCL_NEG_INF;
    l = e == 0 ? T.logcnt[j] : (e == 1 ? T.logcnt1[j] : log_nl((double)cnt - T.discount));
  } else l = cE->prior_pool[s.prior_off + j];
  const TermD* terms = cE->terms + cP->term0;
  double part = 0.0;
  for (int t = s.term0 + cLane; t < s.term0 + s.nterm; t += 32) {
    const uint8_t* rp = cW->rowp[t];
    if (rp) part += score_fast(rp[j], cW->elenp[t][j], terms[t].max_typos, cLG, cLOGN, cLUT);
  }
  #pragma unroll
  for (int o = 16; o; o >>= 1) part += shfl_xor_d(part, o);
  return l + part;
}

// Returns the raw log-sum-exp (new-row branch included, logden not subtracted) and leaves the
// surviving elements in W->sv_* (ascending element index; the new-row branch, if any, last with
// index J).  Returns false if pruning is not applicable (caller uses the exact path).
// `lazy_ub` (root stars): an upper bound of the new-row branch computed from the hoisted children only
// (every other child marginal is <= 0: normalised priors x likelihoods <= 1).  If even the bound is
// PCL_PRUNE_MARGIN nats below the best existing candidate, the branch is dropped without evaluating
// the child stars at all (they exist only to score it); otherwise the function returns false and the
// caller evaluates the children and comes back without `lazy_ub`.
template <class C> __device__ PCL_PRUNED_INLINE bool star_eval_pruned(const C& c, const StarD& s, double* Lraw_out, int hint = -1, const double* lazy_ub = nullptr) { PCL_CTX(c);
  WarpState* W = cW;
  if (cLane == 0) W->sv_star = -1;
  __syncwarp();
  const int J = star_nelem(c, s);
  const TableD* T = s.kind == 0 ? &cE->tables[s.table] : nullptr;
  const bool prog = (cE->opts & PCL_OPT_PROGRESSIVE) != 0;
  int nt = 0;
  if (s.nterm <= 32) {
    // one lane per term: the non-missing row pointers, most selective term first, compacted in order
    const uint8_t* rp = nullptr;
    if (cLane < s.nterm) rp = W->rowp[s.term0 + (prog ? cE->term_order[cP->term0 + s.term0 + cLane] : cLane)];
    const unsigned have = __ballot_sync(0xffffffffu, rp != nullptr);
    if (rp) W->act[__popc(have & ((1u << cLane) - 1u))] = rp;
    nt = __popc(have);
  } else {
    const int* ord = cE->term_order + cP->term0 + s.term0;     // this star's terms, most selective first
    for (int i = 0; i < s.nterm; ++i) {
      const int t = s.term0 + (prog ? ord[i] : i);
      if (W->rowp[t]) { if (cLane == 0) W->act[nt] = W->rowp[t]; ++nt; }
    }
  }
  if (cLane == 0) W->nact = nt;
  __syncwarp();
  if (nt == 0 && J > PCL_SURV_MAX) return false;
  if (C::rich && (s.bucket || s.list_func >= 0 || s.inner_elems >= 0 || s.has_eq) && J > PCL_SURV_MAX) return false;
  const int lane = cLane;
  int nsv = 0;
  if (J <= PCL_SURV_MAX) {
    // small star: everything "survives"
    for (int j = lane; j < J; j += 32) { W->sv_idx[j] = j; W->sv_ll[j] = star_elem(c, s, j); }
    nsv = J;
  } else {
    const int J16 = (J + 15) & ~15;
    const double Bmax = T ? T->max_logcnt : 0.0;
    const int cap = 255 * nt;
    // A known good candidate (the row's current reference, i.e. the retained particle) gives the
    // lower bound up front: one collection pass instead of min-search + collection.
    int tau = -1;
    if (hint >= 0 && hint < J) {
      const double l0 = (cE->opts & PCL_OPT_PARHINT) ? star_elem_par(c, s, hint) : star_elem(c, s, hint);
      if (l0 != PCL_NEG_INF) { const double need0 = (Bmax - l0 + PCL_PRUNE_MARGIN) / PCL_TYPO_COST; if (need0 < (double)cap) tau = (int)need0 + 1; }
    }
    unsigned best = 0;
    if (tau < 0) {
      // pass 1: smallest distance sum over live candidates (branch and bound: once some candidate
      // is known, a stride is abandoned as soon as nobody in it can beat it by the 18-typo window)
      best = 0xFFFFu;
      unsigned tau_run = 0x7FFFu;
      for (int jb = 0; jb < J16; jb += 512) {
        const int j0 = jb + lane * 16;
        unsigned v[8];
        bool live;
        if (prog) live = star_sum16_prog(c, T, j0, J, J16, tau_run, v);
        else { live = j0 < J16; if (live) star_sum16(c, s, T, j0, J, v); }
        if (!__any_sync(0xffffffffu, live)) continue;
        unsigned mine = 0xFFFFu;
        if (live) {
          #pragma unroll
          for (int w = 0; w < 8; ++w) mine = min(mine, min(v[w] & 0xFFFFu, v[w] >> 16));
        }
        for (int o = 16; o; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
        best = min(best, mine);
        if (best != 0xFFFFu) tau_run = min(0x7FFFu, best + 18u);
      }
      tau = (int)best + 18;
    }
    if (best == 0xFFFFu) { nsv = 0; }
    else {
      for (int round = 0; round < 2; ++round) {
        // collect candidates with distance sum <= tau (order preserving)
        nsv = 0;
        bool overflow = false;
        for (int jb = 0; jb < J16; jb += 512) {
          const int j0 = jb + lane * 16;
          unsigned v[8];
          bool live;
          if (prog) live = star_sum16_prog(c, T, j0, J, J16, (unsigned)min(tau, 0x7FFF), v);
          else {
            live = j0 < J16;
            if (live) star_sum16(c, s, T, j0, J, v);
          }
          unsigned keep = 0;
          if (live) {
            #pragma unroll
            for (int w = 0; w < 8; ++w) {
              keep |= ((int)(v[w] & 0xFFFFu) <= tau ? 1u : 0u) << (2 * w);
              keep |= ((int)(v[w] >> 16) <= tau ? 1u : 0u) << (2 * w + 1);
            }
          }
          const unsigned anyv = __ballot_sync(0xffffffffu, keep != 0);
          if (!anyv) continue;
          const int cnt = __popc(keep);
          int incl = cnt;
          for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
          const int tot = __shfl_sync(0xffffffffu, incl, 31);
          int pos = nsv + incl - cnt;
          if (nsv + tot > PCL_SURV_MAX) { overflow = true; break; }
          while (keep) { const int q = __ffs(keep) - 1; keep &= keep - 1; W->sv_idx[pos++] = j0 + q; }
          nsv += tot;
        }
        if (overflow) { if (hint >= 0) return star_eval_pruned(c, s, Lraw_out, -1, lazy_ub); return false; }
        __syncwarp();
        if (nsv <= 4 && (cE->opts & PCL_OPT_PARHINT) && !(C::rich && (s.inner_elems >= 0 || s.has_eq))) {
          for (int i = 0; i < nsv; ++i) { const double v = star_elem_par(c, s, W->sv_idx[i]); if (lane == 0) W->sv_ll[i] = v; }
        } else
          for (int i = lane; i < nsv; i += 32) W->sv_ll[i] = star_elem(c, s, W->sv_idx[i]);
        __syncwarp();
        double lb = PCL_NEG_INF;
        for (int i = lane; i < nsv; i += 32) lb = fmax(lb, W->sv_ll[i]);
        for (int o = 16; o; o >>= 1) lb = fmax(lb, shfl_xor_d(lb, o));
        if (!lazy_ub) lb = fmax(lb, star_extra(c, s));
        if (lb == PCL_NEG_INF) { if (lazy_ub) return false; if (tau >= cap) break; tau = cap; continue; }
        const double need = (Bmax - lb + PCL_PRUNE_MARGIN) / PCL_TYPO_COST;
        if (need <= (double)tau) break;              // every candidate that matters is already in the list
        if (need >= (double)cap || round == 1) return false;   // bound too weak: exact path
        tau = (int)need + 1;
      }
    }
  }
  __syncwarp();
  // append the new-row branch and reduce
  double ex;
  if (lazy_ub) {
    double lbx = PCL_NEG_INF;
    for (int i = lane; i < nsv; i += 32) lbx = fmax(lbx, W->sv_ll[i]);
    for (int o = 16; o; o >>= 1) lbx = fmax(lbx, shfl_xor_d(lbx, o));
    if (!(*lazy_ub < lbx - PCL_PRUNE_MARGIN)) return false;     // the new-row branch may matter: it needs its children
    ex = PCL_NEG_INF;                                            // below e^-45 of the best candidate: dropped like any pruned candidate
  } else ex = star_extra(c, s);
  if (s.kind == 0) { if (lane == 0) { W->sv_idx[nsv] = J; W->sv_ll[nsv] = ex; } nsv += 1; }
  if (lane == 0) { W->sv_n = nsv; W->sv_star = (int)(&s - (cE->stars + cP->star0)); }
  __syncwarp();
  Lse acc; acc.m = PCL_NEG_INF; acc.s = 0.0;
  for (int i = lane; i < nsv; i += 32) lse_add(acc, W->sv_ll[i]);
  *Lraw_out = lse_warp(acc);
  return true;
}

// Inverse-CDF draw over the survivor list (same order as the full enumeration): the running sums
// go to shared memory once, then every lane (= particle) binary-searches its own uniform — the
// first index whose running sum exceeds u, which is what a linear scan would return.
template <class C> __device__ PCL_NI2 int surv_sample(const C& c, double Lraw, double u, bool active) { PCL_CTX(c);
  WarpState* W = cW;
  const int n = W->sv_n;
  double carry = 0.0; int lastpos = -1;
  for (int base = 0; base < n; base += 32) {
    const int i = base + cLane;
    double p = 0.0;
    if (i < n) { const double l = W->sv_ll[i]; p = l - Lraw < PCL_EXP_CUTOFF ? 0.0 : exp_nl(l - Lraw); }
    double cs = p;
    for (int o = 1; o < 32; o <<= 1) { const double t = shfl_up_d(cs, o); if (cLane >= o) cs += t; }
    const double tot = shfl_d(cs, 31);
    const unsigned pos = __ballot_sync(0xffffffffu, p > 0.0);
    if (pos) lastpos = base + 31 - __clz(pos);
    if (i < n) W->sv_cs[i] = carry + cs;
    carry += tot;
  }
  __syncwarp();
  int idx = -1;
  if (active) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (u < W->sv_cs[mid]) hi = mid; else lo = mid + 1; }
    idx = lo < n ? lo : lastpos;
  }
  const int out = idx >= 0 ? W->sv_idx[idx] : -1;
  __syncwarp();
  return out;
}

// Inverse-CDF draw (oracle: Oracle::categorical) for up to 32 uniforms at once: lane i holds
// uniform `u` (active lanes only).  Returns the chosen element (J = new-row branch).
template <class C> __device__ PCL_NI3 int star_sample(const C& c, const StarD& s, double Lraw, double u, bool active) { PCL_CTX(c);
  const int J = star_nelem(c, s);
  const int Jx = J + (s.kind == 0 ? 1 : 0);
  double carry = 0.0;
  bool found = !active;
  int idx = -1, lastpos = -1;
  for (int base = 0; base < Jx; base += 32) {
    const int j = base + cLane;
    double p = 0.0;
    if (j < J) { const double l = star_elem(c, s, j); p = l - Lraw < PCL_EXP_CUTOFF ? 0.0 : exp_nl(l - Lraw); }
    else if (j == J && j < Jx) { const double l = star_extra(c, s); p = l - Lraw < PCL_EXP_CUTOFF ? 0.0 : exp_nl(l - Lraw); }
    double cs = p;
    for (int o = 1; o < 32; o <<= 1) { const double t = shfl_up_d(cs, o); if (cLane >= o) cs += t; }
    const double tot = shfl_d(cs, 31);
    const unsigned pos = __ballot_sync(0xffffffffu, p > 0.0);
    if (pos) lastpos = base + 31 - __clz(pos);
    const bool hit = !found && (u < carry + tot);
    if (__any_sync(0xffffffffu, hit)) {
      #pragma unroll 1
      for (int i = 0; i < 32; ++i) {
        const double ci = carry + shfl_d(cs, i);
        if (hit && !found && u < ci) { idx = base + i; found = true; }
      }
    }
    carry += tot;
  }
  if (active && idx < 0) idx = lastpos;
  return idx;
}

__device__ __forceinline__ double row_uniform(uint64_t seed, uint32_t sweep, uint32_t cls, long long r, int particle,
                                               int block, int site, int purpose) {
  pclean_rng_key k; k.seed = seed; k.sweep = sweep; k.cls = cls; k.row = r; k.particle = (uint32_t)particle;
  k.block = (uint32_t)block; k.site = (uint32_t)site; k.purpose = (uint32_t)purpose;
  return pclean_uniform(&k, 0);
}

// sample the inner choices of one element for particle k; vals[pos] = value id per local choice position
template <class C> __device__ void inner_sample(const C& c, const InnerD& I, const ElemRef& e, int k, int block, uint64_t seed, uint32_t sweep, uint32_t cls, int* vals) { PCL_CTX(c);
  double u[PCL_MAX_INNER_CH]; int picked[PCL_MAX_INNER_CH] = {0, 0, 0};
  for (int i = 0; i < I.nchoice; ++i) u[i] = row_uniform(seed, sweep, cls, cR, k, block, I.ch[i].vertex, PCLEAN_RNG_ENUM);
  inner_eval(c, I, e, u, picked);
  for (int i = 0; i < I.nchoice; ++i)
    for (int p = 0; C::rich && p < cP->n_local; ++p)
      if (cP->local_vertex[p] == I.ch[i].vertex) vals[p] = cE->innervals[I.ch[i].list_off + picked[i]];
}

// resolve per-term matrices for an upstream a-slot; returns false if a join matrix is missing
template <class C> __device__ bool resolve_terms(const C& c, int a_slot) { PCL_CTX(c);
  const TermD* terms = cE->terms + cP->term0;
  bool ok = true;
  for (int t = cLane; t < cP->nterm; t += 32) {
    int m = terms[t].mat;
    if (C::rich && (terms[t].kind == 5 || terms[t].lmat >= 0)) { cW->tmat[t] = 0; cW->rowp[t] = nullptr; cW->elenp[t] = nullptr; continue; }   // list blocks: set by star_prepare
    if (terms[t].kind >= 2) {
      m = a_slot >= 0 ? cE->join_mat[(long long)terms[t].mat * cE->max_a + a_slot] : -1;
      if (m < 0) { ok = false; m = 0; }
    }
    cW->tmat[t] = m;
    const int u = cW->u[t];
    const MatD M = cE->mats[m];
    cW->rowp[t] = u >= 0 ? M.d + (long long)u * M.stride : nullptr;
    cW->elenp[t] = M.elen;
  }
  ok = __all_sync(0xffffffffu, ok);
  __syncwarp();
  return ok;
}

// ---- star-marginal memo ---------------------------------------------------------------------
// The marginal of a non-root star depends on the row only through the unique observed strings its
// terms read (+ the upstream value): rows that share them share the value.  The key is EXACT: the
// tuple (star, upstream slot, up to four 22-bit unique-string indices) packed into 128 bits — `lo`
// claims the slot with a CAS, `hi` travels with the value in one 16-byte store, so a reader takes a
// value only if both halves it read belong to its own tuple (a torn read can only look like a miss
// or a pending entry).  A reader that finds the entry pending computes the value itself (no
// waiting).  Table 0 (reference-table stars: counts change every sweep) is cleared per launch;
// table 1 (choice stars: option lists, priors and matrices only) persists until a prior changes.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
struct MemoKey { unsigned long long lo, hi; int tbl; };
template <class C> __device__ bool memo_key(const C& c, const StarD& s, int sidx, int a_slot, MemoKey* key) { PCL_CTX(c);
  if (s.kind == 0) {       // FK star: values also depend on the row's own exclusions on that table
    for (int i = 0; i < cW->n_ex; ++i) if (cW->ex_table[i] == s.table) return false;
  }
  if (s.nterm == 0 || s.nterm > 4) return false;
  if (C::rich && (s.bucket || s.inner_elems >= 0 || s.inner_new >= 0)) return false;
  const int gs = cP->star0 + sidx;
  if (gs >= 512 || a_slot + 1 >= 1024) return false;
  unsigned long long w[4] = {0, 0, 0, 0};
  if (C::rich && s.list_func >= 0) {
    // a choice over a row-dependent option list: the marginal is a function of the observed strings AND of
    // which list the row's key selects — the list id takes the fourth slot of the key (observation rows only)
    if (s.nterm > 3 || s.list_obs_col < 0) return false;
    const int kv = cE->obs_sid[s.list_obs_col][cR];
    int l = kv >= 0 ? lookup_find(cE->lookups[s.list_func], kv, 0, 0) : -1;
    if (l == PCL_LOOKUP_EMPTY) l = -1;
    if (l + 1 >= (1 << 22)) return false;
    w[3] = (unsigned long long)(l + 1);
  }
  for (int i = 0; i < s.nterm; ++i) {
    const int u1 = cW->u[s.term0 + i] + 1;                     // 0 = explicit missing
    if (u1 >= (1 << 22)) return false;
    w[i] = (unsigned long long)u1;
  }
  key->lo = (1ull << 63) | ((unsigned long long)gs << 54) | ((unsigned long long)(a_slot + 1) << 44) | (w[1] << 22) | w[0];
  key->hi = (w[3] << 22) | w[2];
  key->tbl = (s.kind == 1 && (cE->opts & PCL_OPT_PMEMO)) ? 1 : 0;
  return true;
}
#define PCL_MEMO_PENDING 0x7FF8DEADBEEF0001ULL
__device__ int memo_probe(const Dev* E, const MemoKey& key, double* val, bool* hit) {
  unsigned long long* keys = E->memo_keys[key.tbl];
  const ulonglong2* vals = E->memo_vals[key.tbl];
  unsigned h = (unsigned)(mix64(key.lo ^ mix64(key.hi + 0x9E3779B97F4A7C15ULL)) & E->memo_mask);
  *hit = false;
  for (int p = 0; p < 8; ++p, h = (h + 1) & E->memo_mask) {
    unsigned long long k = keys[h];
    if (k == 0) {
      const unsigned long long old = atomicCAS(&keys[h], 0ull, key.lo);
      if (old == 0) return (int)h;                 // we own this slot: compute and publish
      k = old;
    }
    if (k == key.lo) {
      const ulonglong2 e = __ldcg(&vals[h]);         // L2 is the coherence point: never a stale L1 line
      if (e.x == PCL_MEMO_PENDING) return -1;       // not published yet (by whoever owns it): compute, do not wait
      if (e.y == key.hi) { *val = __longlong_as_double((long long)e.x); *hit = true; return -1; }
      // same low half, different tuple: keep probing
    }
  }
  return -1;
}
__device__ __forceinline__ void memo_publish(const Dev* E, int tbl, unsigned long long hi, int slot, double v) {
  ulonglong2 e; e.x = (unsigned long long)__double_as_longlong(v); e.y = hi;
  if (e.x == PCL_MEMO_PENDING) return;              // (a NaN payload nobody produces) never publish the sentinel
  __stcg(E->memo_vals[tbl] + slot, e);              // one 16-byte store: value and key half appear together
}
__global__ void k_memo_reset(unsigned long long* keys, ulonglong2* vals, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = 0ull;
  ulonglong2 e; e.x = PCL_MEMO_PENDING; e.y = 0ull;
  vals[i] = e;
}

// Evaluate every star bottom-up for the current upstream state.
template <class C> __device__ PCL_NI1 void eval_program(const C& c, int a_slot, int root_hint) { PCL_CTX(c);
  const StarD* stars = cE->stars + cP->star0;
  // Hoisted stars are two dependent loads each (unique-string index, then its value) and memo
  // probes two or three: one lane per star (PCL_MAX_STARS <= 32: lane oi <-> order[oi]), so these
  // latencies overlap instead of adding up.  Stars settled this way do not enter the loop below
  // (none of them reads a child's value).
  bool done = false;
  if (cLane < cP->norder) {
    const int sidx = cP->order[cLane];
    const StarD& s = stars[sidx];
    if (s.hoist >= 0) {
      const int u = cE->uobs[s.hoist_col][cR];
      if (u >= 0) { cW->V[sidx] = cE->hoist_val[s.hoist][u]; done = true; }
    }
  }
  // Root first: the other stars exist only to score the root's new-row branch.  With the hoisted
  // children alone that branch is bounded from above; when the bound is already negligible next to
  // an existing candidate (nearly every row: a fresh row would have to draw all its strings from
  // their priors), none of the remaining stars is evaluated — no memo probes, no enumerations.
  {
    const StarD& root = stars[cP->root];
    // (lean programs only: every likelihood term there is a probability, so child marginals are <= 0;
    // Gaussian densities of the rents shapes may exceed 1)
    // A program whose new-row branch is a priori plausible (uniform / proportional priors over few
    // options) fails the test row after row: a warp that mostly fails stops trying.
    const bool worth = !(cW->lazy_fail >= 8 && cW->lazy_fail > 2 * cW->lazy_ok);
    if (!C::rich && (cE->opts & PCL_OPT_LAZYNEW) && cE->prune && root.kind == 0 && worth) {
      double part = 0.0;
      if (cLane < cP->norder && done && stars[cP->order[cLane]].parent == cP->root) part = cW->V[cP->order[cLane]];
      __syncwarp();
      #pragma unroll
      for (int o = 16; o; o >>= 1) part += shfl_xor_d(part, o);
      const TableD& T = cE->tables[root.table];
      const double ub = part + T.log_new;           // log(strength + discount * live rows): >= the branch's prior for any exclusion (fewer rows)
      double raw;
      if (star_eval_pruned(c, root, &raw, root_hint, &ub)) {
        if (cLane == 0) { cW->V[cP->root] = raw - star_logden(c, root); cW->lazy_ok += 1; }
        __syncwarp();
        return;
      }
      if (cLane == 0) cW->lazy_fail += 1;
      __syncwarp();
    }
  }
  if (cLane < cP->norder && !done) {
    const int sidx = cP->order[cLane];
    const StarD& s = stars[sidx];
    int mslot = -1;
    if (cE->memo_mask && sidx != cP->root) {
      MemoKey mkey;
      if (memo_key(c, s, sidx, a_slot, &mkey)) {
        bool hit = false; double mval = 0.0;
        mslot = memo_probe(cE, mkey, &mval, &hit);
        if (hit) { cW->V[sidx] = mval; done = true; }
        cW->mk_hi[cLane] = mkey.hi; cW->mk_tbl[cLane] = mkey.tbl;
      }
    }
    cW->mk_slot[cLane] = mslot;
  }
  unsigned pending = __ballot_sync(0xffffffffu, cLane < cP->norder && !done);
  __syncwarp();
  while (pending) {
    const int oi = __ffs(pending) - 1; pending &= pending - 1;
    const int sidx = cP->order[oi];
    const StarD& s = stars[sidx];
    if (C::rich && (s.bucket || s.list_func >= 0)) star_prepare(c, s);
    double v;
    if (s.hoist >= 0) v = star_lse_raw(c, s);               // explicit missing observation: prior mass only
    else {
      double raw;
      if (!(cE->prune && star_eval_pruned(c, s, &raw, sidx == cP->root ? root_hint : -1))) raw = star_lse_raw(c, s);
      v = raw - star_logden(c, s);
      if (cLane == 0) { const int slot = cW->mk_slot[oi]; if (slot >= 0) memo_publish(cE, cW->mk_tbl[oi], cW->mk_hi[oi], slot, v); }
    }
    if (cLane == 0) cW->V[sidx] = v;
    __syncwarp();
  }
}

// Sample the contents of a proposed new row under star `s` (an FK star whose new-row branch
// was chosen) for particle `k`, writing the cells into scratch (obs-class vertex numbering).
// Iterative pre-order walk with an explicit stack (depth <= PCL_MAX_STARS).
// a cell of the new row that nothing informs: drawn from the choice's discrete proposal (TimePrior:
// the atoms that look like times get 1/1440 each, the dummy the rest; a dummy draw is replaced by
// random(), time_prior.jl:8-22).  Returns the weight the draw contributes (p - q_cont).
template <class C> __device__ double fill_new_cell(const C& c, const FillD& f, int k, int block, int* scratch, uint64_t seed, uint32_t sweep, uint32_t cls) { PCL_CTX(c);
  const Dev& E = *cE;
  const int list = f.list_const >= 0 ? f.list_const : lookup_ref(E, f.list, cR, -1);
  const int n = list >= 0 && list != PCL_LOOKUP_EMPTY ? E.lists_off[list + 1] - E.lists_off[list] : 0;
  int cnt = 0;
  for (int i = 0; i < n; ++i) cnt += E.time_ok[E.lists_sid[E.lists_off[list] + i]];
  const double la = -log(1440.0);
  const double tot_atoms = cnt > 0 ? la + log((double)cnt) : PCL_NEG_INF;
  const double ld = cnt > 0 ? log1p(-exp(tot_atoms)) : 0.0;                      // time_prior.jl:12-13
  const double tot = cnt > 0 ? fmax(tot_atoms, ld) + log(exp(tot_atoms - fmax(tot_atoms, ld)) + exp(ld - fmax(tot_atoms, ld))) : ld;
  const double u = row_uniform(seed, sweep, cls, cR, k, block, f.vertex, PCLEAN_RNG_PRIOR);
  double cum = 0.0; int chosen = -1;
  for (int i = 0; i < n && chosen < 0; ++i) {
    if (!E.time_ok[E.lists_sid[E.lists_off[list] + i]]) continue;
    cum += exp(la - tot);
    if (u < cum) chosen = i;
  }
  if (chosen >= 0) { scratch[f.vertex] = E.lists_sid[E.lists_off[list] + chosen]; return 0.0; }
  pclean_stream st; st.key.seed = seed; st.key.sweep = sweep; st.key.cls = cls; st.key.row = cR; st.key.particle = (uint32_t)k;
  st.key.block = (uint32_t)block; st.key.site = (uint32_t)f.vertex; st.key.purpose = PCLEAN_RNG_RANDOM; st.idx = 0;
  const int hh = min(11, (int)(pclean_next(&st) * 12)), mi = min(59, (int)(pclean_next(&st) * 60));
  const int pm = pclean_next(&st) < 0.5 ? 0 : 1;
  scratch[f.vertex] = E.time_sid[(hh * 60 + mi) * 2 + pm];
  return -ld;
}

// restricted Damerau-Levenshtein (optimal string alignment) of a dictionary string against a freshly
// generated one, plain two-row programme (rare path: only a particle that drew a StringPrior dummy)
__device__ __noinline__ int osa_plain(const uint8_t* A, int m, const uint8_t* B, int nb) {
  int pp[PCL_NEWSTR_MAX + 1], p[PCL_NEWSTR_MAX + 1], cur[PCL_NEWSTR_MAX + 1];
  for (int j = 0; j <= nb; ++j) { p[j] = j; pp[j] = 0; }
  for (int x = 1; x <= m; ++x) {
    cur[0] = x;
    for (int j = 1; j <= nb; ++j) {
      const int cost = A[x - 1] == B[j - 1] ? 0 : 1;
      int v = min(min(p[j] + 1, cur[j - 1] + 1), p[j - 1] + cost);
      if (x > 1 && j > 1 && A[x - 1] == B[j - 2] && A[x - 2] == B[j - 1]) v = min(v, pp[j - 2] + 1);
      cur[j] = v;
    }
    for (int j = 0; j <= nb; ++j) { pp[j] = p[j]; p[j] = cur[j]; }
  }
  return m == 0 ? nb : p[nb];
}

// A particle drew the dummy of a StringPrior choice: the reference replaces it by random(StringPrior)
// (block_proposal.jl:58-60, string_prior.jl:27-38) — the value then scores the observations instead of
// the placeholder, and the choice itself adds no prior term.  The string is a pure function of the
// keyed stream (the oracle draws the same one); it goes to the new-string pool, its provisional id
// into the scratch record.  Returns the weight it adds on top of the enumeration's marginal:
//   sum_t [ AddTypos(obs_t | random) - AddTypos(obs_t | placeholder) ] - log prior(dummy).
// Lane 0 only.  Returns NaN when the pool is full / unavailable (the caller marks the particle unusable).
template <class C> __device__ __noinline__ double dummy_string_draw(const C& c, const StarD& cs, int k, int block, int* scratch, uint64_t seed, uint32_t sweep, uint32_t cls) { PCL_CTX(c);
  const Dev& E = *cE;
  if (E.newstr_cap <= 0 || cs.sp_max > PCL_NEWSTR_MAX || cs.sp_max < cs.sp_min) return CUDART_NAN;
  const int idx = atomicAdd(E.newstr_count, 1);
  if (idx >= E.newstr_cap) return CUDART_NAN;
  pclean_stream st; st.key.seed = seed; st.key.sweep = sweep; st.key.cls = cls; st.key.row = cR; st.key.particle = (uint32_t)k;
  st.key.block = (uint32_t)block; st.key.site = (uint32_t)cs.vertex; st.key.purpose = PCLEAN_RNG_RANDOM; st.idx = 0;
  const int mn = cs.sp_min, mx = cs.sp_max;
  const int len = mn + min(mx - mn, (int)(pclean_next(&st) * (mx - mn + 1)));
  uint8_t* out = E.newstr_chars + (long long)idx * PCL_NEWSTR_MAX;
  uint8_t symb[PCL_NEWSTR_MAX];
  int prev = -1;
  for (int i = 0; i < len; ++i) {
    double tot = 0.0;
    for (int q = 0; q < 28; ++q) tot += prev < 0 ? E.lm_uni[q] : E.lm_big[q * 28 + prev];
    const double u = pclean_next(&st);
    double acc = 0.0; int pick = -1, last = -1;
    for (int q = 0; q < 28; ++q) {
      const double pr = (prev < 0 ? E.lm_uni[q] : E.lm_big[q * 28 + prev]) / tot;
      if (pr > 0.0) last = q;
      acc += pr;
      if (u < acc) { pick = q; break; }
    }
    if (pick < 0) pick = last;
    out[i] = (uint8_t)pick; symb[i] = (uint8_t)E.lm_sym[pick]; prev = pick;
  }
  E.newstr_len[idx] = len;
  scratch[cs.vertex] = E.newstr_base + idx;
  const int J = star_nelem(c, cs);
  const int sidx = star_index(c, cs);
  int dcol = J - 1; double prior_dummy;
  if (C::rich && cs.list_func >= 0) { dcol = E.univ_col[cs.univ_off + star_option_sid(c, cs, J - 1)]; prior_dummy = cW->aux[sidx]; }
  else prior_dummy = E.prior_pool[cs.prior_off + J - 1];
  double delta = -prior_dummy;
  const TermD* terms = E.terms + cP->term0;
  for (int t = cs.term0; t < cs.term0 + cs.nterm; ++t) {
    const int u = cW->u[t];
    if (u < 0) continue;                                      // explicit missing observation: log-density 0 either way
    const int mt = terms[t].max_typos;
    const int osid = E.ulist[terms[t].obs_col][u];
    if (C::rich && terms[t].lmat >= 0) {                      // per-list blocks: the placeholder is the list's last column
      const int psid = star_option_sid(c, cs, J - 1);
      const uint8_t* rp = cW->rowp[t];
      delta -= score_fast(rp ? rp[J - 1] : lmat_inline(cE, osid, psid), rp ? cW->elenp[t][J - 1] : min(255, E.str_len[psid]), mt, cLG, cLOGN, cLUT);
    } else {
      if (!cW->rowp[t]) continue;
      delta -= score_fast(cW->rowp[t][dcol], cW->elenp[t][dcol], mt, cLG, cLOGN, cLUT);
    }
    const int d = osa_plain(E.sym + E.str_off[osid], E.str_len[osid], symb, len);
    delta += score_fast(min(d, 255), len, mt, cLG, cLOGN, cLUT);
  }
  return delta;
}

template <class C> __device__ __noinline__ void expand_new(const C& c, int sroot, int k, int block, int* scratch, uint64_t seed, uint32_t sweep, uint32_t cls, int* inner_vals, double* wdelta, int* bad) { PCL_CTX(c);
  const StarD* stars = cE->stars + cP->star0;
  int stack[PCL_MAX_STARS]; int sp = 0;
  stack[sp++] = sroot;
  while (sp > 0) {
    const StarD& ps = stars[stack[--sp]];
    if (cLane == 0) scratch[ps.vertex] = -1;            // this reference slot points at a new row
    if (C::rich && ps.nfill > 0 && cLane == 0)
      for (int q = 0; q < ps.nfill; ++q) *wdelta += fill_new_cell(c, cE->fills[ps.fill0 + q], k, block, scratch, seed, sweep, cls);
    if (C::rich && ps.inner_new >= 0 && cLane == 0) {              // choices enumerated inside the new-row branch itself
      ElemRef er; er.table = ps.table; er.slot = -1; er.esid = -1;
      inner_sample(c, cE->inners[ps.inner_new], er, k, block, seed, sweep, cls, inner_vals);
    }
    const int* ch = cE->children + ps.child0;
    for (int i = 0; i < ps.nchild; ++i) {
      const int cidx = ch[i];
      const StarD& cs = stars[cidx];
      const double u = row_uniform(seed, sweep, cls, cR, k, block, cs.vertex, PCLEAN_RNG_ENUM);
      if (C::rich && (cs.bucket || cs.list_func >= 0)) star_prepare(c, cs);     // a marginal that came from the memo left the star unprepared
      double Lraw;
      if (cs.hoist >= 0) Lraw = star_lse_raw(c, cs);      // hoisted marginal -> recompute raw LSE
      else Lraw = cW->V[cidx] + star_logden(c, cs);
      int e;
      {
        double raw2;
        if (cE->prune && star_eval_pruned(c, cs, &raw2)) e = surv_sample(c, raw2, u, true);
        else e = star_sample(c, cs, Lraw, u, true);
      }
      const int J = star_nelem(c, cs);
      if (cs.kind == 1) {
        const int sid = star_option_sid(c, cs, e);
        if (cLane == 0) {
          scratch[cs.vertex] = sid;
          if (cs.has_dummy && e == J - 1) {
            atomicOr(&cE->row_flags[cR], ROWFLAG_DUMMY);
            const double dw = cs.dummy_time ? CUDART_NAN : dummy_string_draw(c, cs, k, block, scratch, seed, sweep, cls);
            if (dw == dw) *wdelta += dw; else *bad = 1;        // no pool (sharded engine / full): the placeholder stays and the particle is never selected
          }
          if (C::rich && cs.inner_elems >= 0) { ElemRef er; er.table = -1; er.slot = -1; er.esid = sid; inner_sample(c, cE->inners[cs.inner_elems], er, k, block, seed, sweep, cls, inner_vals); }
        }
      } else {
        if (e >= J) stack[sp++] = cidx;                   // nested new row
        else {
          const int slot = star_slot(c, cs, e);
          const TableD& T = cE->tables[cs.table];
          const int2* cp = cE->copies + cs.copy0;
          for (int q = cLane; q < cs.ncopy; q += 32) scratch[cp[q].x] = T.cells[(long long)cp[q].y * T.cap + slot];
          __syncwarp();
          if (cLane == 0) {
            scratch[cs.vertex] = slot;
            if (C::rich && cs.inner_elems >= 0) { ElemRef er; er.table = cs.table; er.slot = slot; er.esid = -1; inner_sample(c, cE->inners[cs.inner_elems], er, k, block, seed, sweep, cls, inner_vals); }
          }
        }
      }
      __syncwarp();
    }
  }
}

template <class C> __device__ void block_move_row(const Dev& E, const ProgD& P, int block, long long r, WarpState* W, const double* sLG,
                               const double* sLOGN, const double* sLUT, int lane, uint64_t seed, uint32_t sweep, uint32_t cls, int csmc) {
  const StarD* stars = E.stars + P.star0;
  const TermD* terms = E.terms + P.term0;
  C c;
  if (lane == 0) { W->row = r; W->refs = nullptr; W->nref = -1; }
  const int K = E.K;
  const long long N = E.N;

  for (int t = lane; t < P.nterm; t += 32) W->u[t] = E.uobs[terms[t].obs_col][r];
  // self-exclusion = unincorporate_row! (dependency_tracking.jl:26-66) done arithmetically:
  // the row's own references are removed from the counts; if one was the last reference the
  // target row is garbage-collected, cascading through that row's own reference slots (:162-202).
  bool fast = false;
  if (csmc && (E.opts & PCL_OPT_FASTEXCL)) {
    // common case, one lane per reference slot: no target is about to lose its last reference
    // (count > number of slots of the row), so there is no cascade to walk
    int t = -1, sl = -1, cnt = 0x7fffffff;
    if (lane < E.n_blocks) {
      const ProgD& P2 = E.progs[P.base_prog + lane];
      if (P2.root >= 0) { t = E.stars[P2.star0 + P2.root].table; sl = E.assign[lane][r]; cnt = E.tables[t].refcnt[sl]; }
    }
    const unsigned valid = __ballot_sync(0xffffffffu, t >= 0);
    fast = __all_sync(0xffffffffu, cnt > E.n_blocks) && __popc(valid) <= PCL_MAX_EX;
    if (fast) {
      if (t >= 0) { const int i = __popc(valid & ((1u << lane) - 1u)); W->ex_table[i] = t; W->ex_slot[i] = sl; W->ex_gc[i] = 0; }
      if (lane == 0) { W->n_ex = __popc(valid); W->sv_star = -1; }
    }
  }
  if (!fast && lane == 0) {
    int n = 0;
    if (csmc) {
      int qt[PCL_MAX_EX], qs[PCL_MAX_EX]; int qh = 0, qn = 0;
      for (int b2 = 0; b2 < E.n_blocks && qn < PCL_MAX_EX; ++b2) {   // every reference slot of the row
        const ProgD& P2 = E.progs[P.base_prog + b2];
        if (P2.root < 0) continue;                                   // block without a reference slot
        qt[qn] = E.stars[P2.star0 + P2.root].table; qs[qn] = E.assign[b2][r]; ++qn;
      }
      while (qh < qn && n < PCL_MAX_EX) {
        const int t = qt[qh], s = qs[qh]; ++qh;
        int prev = 0;
        for (int i = 0; i < n; ++i) prev += (W->ex_table[i] == t && W->ex_slot[i] == s);
        const TableD& T = E.tables[t];
        const int gc = (T.refcnt[s] - prev - 1) <= 0;
        W->ex_table[n] = t; W->ex_slot[n] = s; W->ex_gc[n] = gc; ++n;
        if (gc) for (int g = 0; g < T.nfk && qn < PCL_MAX_EX; ++g) { qt[qn] = T.fk_table[g]; qs[qn] = T.cells[(long long)T.fk_col[g] * T.cap + s]; ++qn; }
      }
    }
    W->n_ex = n;
    W->sv_star = -1;
  }
  __syncwarp();

  // Particles are handled 32 at a time (lane <-> particle pass * 32 + lane; K <= PCL_MAX_K).  The
  // program is evaluated once per distinct upstream value: a later group or pass with the same
  // value reuses the star marginals and the root's survivor list still in shared memory.
  bool have_eval = false; int eval_a = 0;
  for (int pass = 0; pass * 32 < K; ++pass) {
  const int kk = pass * 32 + lane;              // this lane's particle
  const bool mine = kk < K;
  // upstream (earlier-block) value per particle
  int a_sid = -1;
  if (P.n_earlier && mine) {
    const int ch = E.pchoice[P.earlier_block][PCL_PK(E, kk, r)];
    if (ch == PCL_CHOICE_UNSET) a_sid = -1;
    else if (ch >= 0) { const TableD& T = E.tables[P.earlier_table]; a_sid = T.cells[(long long)P.earlier_col * T.cap + ch]; }
    else a_sid = E.pool[(long long)(-(ch) - 2) * E.nvC + P.earlier_vertex];
  }
  unsigned todo = __ballot_sync(0xffffffffu, mine);
  int my_choice = PCL_CHOICE_UNSET; double my_w = 0.0;
  bool my_bad = false;
  int my_inner[PCL_MAX_INNER_CH] = {PCL_UNSET, PCL_UNSET, PCL_UNSET};
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int a = __shfl_sync(0xffffffffu, a_sid, leader);
    const unsigned members = __ballot_sync(0xffffffffu, mine && a_sid == a);
    todo &= ~members;
    if (!(have_eval && eval_a == a)) {
      have_eval = false;
      int a_slot = -1;
      if (P.n_earlier) a_slot = (a >= 0 && a < E.n_strings) ? E.a_slot_of_sid[a] : -1;
      if (!resolve_terms(c, a_slot)) {
        // no join matrices for this upstream value: these particles cannot be scored and are
        // never selected (weight -inf); the row is counted in ROWFLAG_NOJOIN
        if (lane == 0) atomicOr(&E.row_flags[r], ROWFLAG_NOJOIN);
        if ((members >> lane) & 1u) { my_w = PCL_NEG_INF; my_bad = true; }
        continue;
      }
      eval_program(c, a_slot, csmc ? E.assign[block][r] : -1);
      have_eval = true; eval_a = a;
    }
    const StarD& root = stars[P.root];
    const double L = W->V[P.root];
    const double Lraw = L + star_logden(c, root);
    const bool member = (members >> lane) & 1u;
    const bool forced = csmc && kk == 0;
    const bool draws = member && !forced;
    double u = 0.0;
    if (draws) u = row_uniform(seed, sweep, cls, r, kk, block, root.vertex, PCLEAN_RNG_ENUM);
    int e;
    if (E.prune && W->sv_star == P.root) e = surv_sample(c, Lraw, u, draws);      // survivors of the root are still in smem
    else e = star_sample(c, root, Lraw, u, draws);
    const int J = star_nelem(c, root);
    if (member) { my_w = L; my_choice = forced ? E.assign[block][r] : (e >= 0 && e < J ? star_slot(c, root, e) : e); }
    if (draws && e >= 0 && e < J && C::rich && root.inner_elems >= 0) {           // the choices enumerated inside the chosen candidate
      ElemRef er; er.table = root.table; er.slot = my_choice; er.esid = -1;
      inner_sample(c, E.inners[root.inner_elems], er, kk, block, seed, sweep, cls, my_inner);
    }
    // new-row proposals: expand one particle at a time (whole warp cooperates)
    unsigned newmask = __ballot_sync(0xffffffffu, draws && e >= J);
    if (newmask) have_eval = false;               // the expansions below reuse the survivor list for the stars they sample
    while (newmask) {
      const int k = __ffs(newmask) - 1; newmask &= newmask - 1;
      int pidx = 0;
      if (lane == 0) pidx = atomicAdd(E.pool_count, 1);
      pidx = __shfl_sync(0xffffffffu, pidx, 0);
      if (pidx >= E.pool_cap) {
        if (lane == 0) { atomicOr(&E.row_flags[r], ROWFLAG_POOL); atomicExch(E.err, PCLEAN_ERR_CAPACITY); }
        if (lane == k) { my_choice = E.assign[block][r]; my_w = PCL_NEG_INF; my_bad = true; }
        continue;
      }
      int* scratch = E.pool + (long long)pidx * E.nvC;
      // the record starts from the cells the row observes directly (refer_to_row! builds the new
      // row from the vmapped cells of the referring row, dependency_tracking.jl:205-236)
      for (int v = lane; v < E.nvC; v += 32) {
        const int col = E.vcol[v];
        const int sid = col >= 0 ? E.obs_sid[col][r] : -1;
        scratch[v] = sid >= 0 ? sid : PCL_UNSET;
      }
      __syncwarp();
      int iv[PCL_MAX_INNER_CH] = {PCL_UNSET, PCL_UNSET, PCL_UNSET};
      double wd = 0.0; int bad = 0;
      expand_new(c, P.root, pass * 32 + k, block, scratch, seed, sweep, cls, iv, &wd, &bad);
      wd = shfl_d(wd, 0); if (lane == k) my_w += wd;
      // a particle that drew a StringPrior dummy carries a placeholder, not a value (the reference
      // would draw a random string, block_proposal.jl:58-60): its weight is the same marginal as in
      // the reference and counts in the log-ML estimate, but it is never selected — per particle,
      // the other particles of the row are unaffected
      bad = __shfl_sync(0xffffffffu, bad, 0);
      if (bad && lane == k) my_bad = true;
      for (int q = 0; q < PCL_MAX_INNER_CH; ++q) { const int v = __shfl_sync(0xffffffffu, iv[q], 0); if (lane == k) my_inner[q] = v; }
      if (lane == k) my_choice = -(pidx + 2);
    }
  }
  {
    const unsigned badmask = __ballot_sync(0xffffffffu, my_bad && mine);
    if (badmask && lane == 0) atomicOr(&E.row_bad[r], (unsigned long long)badmask << (32 * pass));
  }
  if (mine) {
    E.pchoice[block][PCL_PK(E, kk, r)] = my_choice;
    E.pweight[PCL_PK(E, kk, r)] += my_w;
    for (int q = 0; C::rich && q < P.n_local; ++q) E.pinner[block][PCL_PINNER(E, q, kk, r)] = my_inner[q];
  }
  }   // pass
}

// k_block: persistent warps, one row per warp per iteration.  One SMC step (block) for all K
// particles of the row: make_block_proposal! (block_proposal.jl:160-191); particles that share
// their upstream state share one enumeration (SURVEY App. B "consequence worth exploiting").
// The device descriptor travels by value (__grid_constant__: constant bank), so table / column
// pointers are one constant-cache read away instead of a dependent global load through `Dev*`.
// WARPS x MINB = resident warps per SM and the register budget ptxas works with: 16 x 2 = 32 warps
// at <= 64 registers, 12 x 2 = 24 warps at <= 80, 16 x 1 = 16 warps at <= 128 (option "kb_variant").
template <bool RICH, int WARPS, int MINB> __global__ void __launch_bounds__(32 * WARPS, MINB)
k_block(const __grid_constant__ Dev E, int prog_id, int block, long long row0, long long nrows, uint64_t seed,
        uint32_t sweep, uint32_t cls, int csmc, const long long* __restrict__ row_list) {
  // dynamic shared memory (> 48 KB: opt-in), layout PCL_OFF_*: score tables, then the launch descriptor and
  // this launch's program — staged here because every phase reaches them through the row context
  // (PCL_CTX) — then one WarpState per warp
  double* sLUT = reinterpret_cast<double*>(pcl_smem);
  double* sLG = reinterpret_cast<double*>(pcl_smem + PCL_OFF_LG);
  double* sLOGN = reinterpret_cast<double*>(pcl_smem + PCL_OFF_LOGN);
  Dev* sE = reinterpret_cast<Dev*>(pcl_smem + PCL_OFF_DEV);
  ProgD* sP = reinterpret_cast<ProgD*>(pcl_smem + PCL_OFF_PROG);
  WarpState* sW = reinterpret_cast<WarpState*>(pcl_smem + PCL_OFF_W);
  for (int i = threadIdx.x; i < PCL_LG_N; i += blockDim.x) sLG[i] = E.LG[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sLOGN[i] = E.LOGN[i];
  for (int i = threadIdx.x; i < PCL_LUT_N * PCL_LUT_N; i += blockDim.x) sLUT[i] = E.LUT[i];
  for (int i = threadIdx.x; i < (int)(sizeof(Dev) / 4); i += blockDim.x) reinterpret_cast<int*>(sE)[i] = reinterpret_cast<const int*>(&E)[i];
  for (int i = threadIdx.x; i < (int)(sizeof(ProgD) / 4); i += blockDim.x) reinterpret_cast<int*>(sP)[i] = reinterpret_cast<const int*>(E.progs + prog_id)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sW[warp].lazy_ok = 0; sW[warp].lazy_fail = 0; }
  __syncwarp();
  const long long total_warps = (long long)gridDim.x * WARPS;
  for (long long wid = (long long)blockIdx.x * WARPS + warp; wid < nrows; wid += total_warps) {
    const long long r = row_list ? row_list[row0 + wid] : row0 + wid;
    if (RICH) block_move_row<RowCtx>(*sE, *sP, block, r, &sW[warp], sLG, sLOGN, sLUT, lane, seed, sweep, cls, csmc);
    else block_move_row<LeanCtx>(*sE, *sP, block, r, &sW[warp], sLG, sLOGN, sLUT, lane, seed, sweep, cls, csmc);
    __syncwarp();
  }
}

// sweep statistics: number of rows that drew a dummy, sum of the per-row log-ML estimates.  Each
// block sums one contiguous slice in a fixed order and the host adds the per-block partials in
// order, so the result does not depend on scheduling.
__global__ void k_sweep_stats(const int* __restrict__ row_flags, const double* __restrict__ row_logml, long long r0, long long r1, int dummy_bit, double* out) {
  __shared__ double ssum[256]; __shared__ double scnt[256];
  const long long n = r1 - r0, per = (n + gridDim.x - 1) / gridDim.x;
  const long long a = r0 + per * blockIdx.x, b = a + per < r1 ? a + per : r1;
  double s = 0.0, c = 0.0;
  for (long long r = a + threadIdx.x; r < b; r += blockDim.x) { s += row_logml[r]; c += (row_flags[r] & dummy_bit) ? 1.0 : 0.0; }
  ssum[threadIdx.x] = s; scnt[threadIdx.x] = c;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) { if ((int)threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; } __syncthreads(); }
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = ssum[0]; out[2 * blockIdx.x + 1] = scnt[0]; }
}

// zero the per-row particle state of a list of rows
__global__ void k_reset_rows(const Dev* __restrict__ Ep, const long long* __restrict__ rows, long long n) {
  const Dev& E = *Ep;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = rows[i];
  for (int k = 0; k < E.K; ++k) E.pweight[PCL_PK(E, k, r)] = 0.0;
  E.plogml[r] = 0.0; E.row_flags[r] = 0; E.row_bad[r] = 0ull;
}
__global__ void k_rows_to_int(const long long* __restrict__ rows, long long n, int* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)rows[i];
}

// Record which upstream string values block `prog_id` will need join matrices for.
__global__ void k_collect_a(const Dev* __restrict__ Ep, int prog_id, long long row0, long long nrows, const long long* __restrict__ rows) {
  const Dev& E = *Ep;
  const ProgD& P = E.progs[prog_id];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * E.K) return;
  const long long r = rows ? rows[i / E.K] : row0 + i / E.K; const int k = (int)(i % E.K);
  const int ch = E.pchoice[P.earlier_block][PCL_PK(E, k, r)];
  int a;
  if (ch == PCL_CHOICE_UNSET) return;
  if (ch >= 0) { const TableD& T = E.tables[P.earlier_table]; a = T.cells[(long long)P.earlier_col * T.cap + ch]; }
  else a = E.pool[(long long)(-(ch) - 2) * E.nvC + P.earlier_vertex];
  if (a >= 0 && a < E.n_strings && E.a_slot_of_sid[a] < 0) { E.needed_a[a] = 1; *E.needed_any = 1; }
}

// maybe_resample (row_inference.jl:87-105) between blocks, particle Gibbs only.
__global__ void k_resample(const Dev* __restrict__ Ep, int block, long long row0, long long nrows, uint64_t seed,
                           uint32_t sweep, uint32_t cls, int csmc, const long long* __restrict__ rows) {
  const Dev& E = *Ep;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const long long r = rows ? rows[i] : row0 + i; const int K = E.K; const long long N = E.N;
  double w[PCL_MAX_K]; double m = PCL_NEG_INF;
  for (int k = 0; k < K; ++k) { w[k] = E.pweight[PCL_PK(E, k, r)]; m = fmax(m, w[k]); }
  if (m == PCL_NEG_INF) return;                 // no usable particle: k_select reports it
  double s = 0.0; for (int k = 0; k < K; ++k) s += exp(w[k] - m);
  const double tot = m + log(s);
  double m2 = PCL_NEG_INF; for (int k = 0; k < K; ++k) m2 = fmax(m2, 2.0 * (w[k] - tot));
  double s2 = 0.0; for (int k = 0; k < K; ++k) s2 += exp(2.0 * (w[k] - tot) - m2);
  const double ess = exp(-(m2 + log(s2)));
  if (!(ess < K / 2.0)) return;
  int idx[PCL_MAX_K]; int old[PCL_MAX_K];
  for (int j = 0; j < K; ++j) {
    if (j == 0 && csmc) { idx[j] = 0; continue; }
    const double u = row_uniform(seed, sweep, cls, r, j, block, 0, PCLEAN_RNG_RESAMPLE);
    double c = 0.0; int pick = -1, last = -1;
    for (int k = 0; k < K; ++k) { const double p = exp(w[k] - tot); if (p > 0.0) last = k; c += p; if (u < c) { pick = k; break; } }
    idx[j] = pick >= 0 ? pick : last;
  }
  for (int b = 0; b <= block; ++b) {
    for (int k = 0; k < K; ++k) old[k] = E.pchoice[b][PCL_PK(E, k, r)];
    for (int k = 0; k < K; ++k) E.pchoice[b][PCL_PK(E, k, r)] = old[idx[k]];
    for (int q = 0; q < PCL_MAX_LOCAL; ++q) {
      if (!E.pinner[b]) break;
      for (int k = 0; k < K; ++k) old[k] = E.pinner[b][PCL_PINNER(E, q, k, r)];
      for (int k = 0; k < K; ++k) E.pinner[b][PCL_PINNER(E, q, k, r)] = old[idx[k]];
    }
  }
  {
    const unsigned long long ob = E.row_bad[r];
    if (ob) { unsigned long long nb = 0ull; for (int k = 0; k < K; ++k) nb |= ((ob >> idx[k]) & 1ull) << k; E.row_bad[r] = nb; }
  }
  for (int k = 0; k < K; ++k) E.pweight[PCL_PK(E, k, r)] = 0.0;
  E.plogml[r] += tot - log((double)K);
}

// final selection (row_inference.jl:157-165) + return value (:186)
__global__ void k_select(const Dev* __restrict__ Ep, long long row0, long long nrows, uint64_t seed, uint32_t sweep,
                         uint32_t cls, int csmc, int use_mh, const long long* __restrict__ rows) {
  const Dev& E = *Ep;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const long long r = rows ? rows[i] : row0 + i; const int K = E.K; const long long N = E.N;
  double w[PCL_MAX_K]; double m = PCL_NEG_INF;
  for (int k = 0; k < K; ++k) { w[k] = E.pweight[PCL_PK(E, k, r)]; m = fmax(m, w[k]); }
  const unsigned long long bad = E.row_bad[r];
  if (m == PCL_NEG_INF) {
    // no particle could be scored (missing join matrices / scratch pool full): a sweep keeps the
    // retained row; initialisation has nothing to fall back to
    if (!csmc) atomicExch(E.err, PCLEAN_ERR_UNSUPPORTED);
    E.sel[r] = 0; E.row_logml[r] = PCL_NEG_INF;
    return;
  }
  double s = 0.0; for (int k = 0; k < K; ++k) s += exp(w[k] - m);
  const double tot = m + log(s);
  const double u = row_uniform(seed, sweep, cls, r, 0, E.n_blocks, 0, PCLEAN_RNG_FINAL);
  int chosen;
  if (use_mh && csmc) {
    const double w0 = exp(w[0] - tot), w1 = exp(w[1] - tot);
    chosen = (u < fmin(1.0, w1 / (1e-10 + w0))) ? 1 : 0;
    if ((bad >> 1) & 1ull) chosen = 0;
  } else if (!bad) {
    double c = 0.0; int pick = -1, last = -1;
    for (int k = 0; k < K; ++k) { const double p = exp(w[k] - tot); if (p > 0.0) last = k; c += p; if (u < c) { pick = k; break; } }
    chosen = pick >= 0 ? pick : last;
  } else {
    // some particles carry placeholders: the draw is over the others, renormalised
    double mg = PCL_NEG_INF;
    for (int k = 0; k < K; ++k) if (!((bad >> k) & 1ull)) mg = fmax(mg, w[k]);
    if (mg == PCL_NEG_INF) { if (!csmc) atomicExch(E.err, PCLEAN_ERR_UNSUPPORTED); chosen = 0; }
    else {
      double sg = 0.0; for (int k = 0; k < K; ++k) if (!((bad >> k) & 1ull)) sg += exp(w[k] - mg);
      const double totg = mg + log(sg);
      double c = 0.0; int pick = -1, last = -1;
      for (int k = 0; k < K; ++k) { if ((bad >> k) & 1ull) continue; const double p = exp(w[k] - totg); if (p > 0.0) last = k; c += p; if (u < c) { pick = k; break; } }
      chosen = pick >= 0 ? pick : (last >= 0 ? last : 0);
    }
  }
  E.sel[r] = chosen;
  E.row_logml[r] = E.plogml[r] + tot - log((double)K);
}

// ------------------------------------------------------------------------------------------
// table maintenance
// ------------------------------------------------------------------------------------------
// write the selected particle's choices back: existing slot -> assignment, new row -> request
__global__ void k_apply(const Dev* __restrict__ Ep, int block, long long row0, long long nrows, int csmc, int* req,
                        int* changed_count, const int* prog_of_row, const long long* __restrict__ rows) {
  const Dev& E = *Ep;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const long long r = rows ? rows[i] : row0 + i;
  const int s = E.sel[r];
  req[i] = -1;
  if (csmc && s == 0) return;
  const int ch = E.pchoice[block][PCL_PK(E, s, r)];
  {
    const ProgD& P = E.progs[prog_of_row ? prog_of_row[r] * E.n_blocks + block : block];
    for (int q = 0; q < P.n_local; ++q) {
      const int v = E.pinner[block][PCL_PINNER(E, q, s, r)];
      if (v != PCL_UNSET && E.rowcell[P.local_vertex[q]]) E.rowcell[P.local_vertex[q]][r] = v;
    }
  }
  if (ch == PCL_CHOICE_UNSET) return;                                // block without a reference slot: local cells only
  if (ch >= 0) {
    if (E.assign[block][r] != ch) { E.assign[block][r] = ch; if (block == 0) atomicAdd(changed_count, 1); }
  } else { req[i] = -(ch) - 2; changed_count[1] = 1; if (block == 0) atomicAdd(changed_count, 1); }     // [1]: some row of this block asks for a new row
}

// flag rows that create a new row at star `sidx` of program `prog_id`
__global__ void k_create_flags(const Dev* __restrict__ Ep, int prog_id, int sidx, long long nrows, const int* req, int* flags) {
  const Dev& E = *Ep;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nrows) return;
  if (i == nrows) { flags[i] = 0; return; }        // sentinel so that the exclusive scan yields the total
  const StarD& s = E.stars[E.progs[prog_id].star0 + sidx];
  int f = 0;
  if (req[i] >= 0) f = E.pool[(long long)req[i] * E.nvC + s.vertex] == -1;
  flags[i] = f;
}

// materialise the new rows of star `sidx` (slot = base + exclusive rank)
__global__ void k_create_rows(const Dev* __restrict__ Ep, int prog_id, int sidx, int block, long long row0, long long nrows,
                              const int* req, const int* flags, const int* rank, int base, int is_root, const int* row_ids) {
  const Dev& E = *Ep;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows || !flags[i]) return;
  const StarD& s = E.stars[E.progs[prog_id].star0 + sidx];
  TableD& T = E.tables[s.table];
  const int slot = base + rank[i];
  if (slot >= T.cap) { atomicExch(E.err, PCLEAN_ERR_CAPACITY); return; }
  int* scratch = E.pool + (long long)req[i] * E.nvC;
  const int2* cp = E.copies + s.copy0;
  for (int q = 0; q < s.ncopy; ++q) T.cells[(long long)cp[q].y * T.cap + slot] = scratch[cp[q].x];
  T.refcnt[slot] = 0;
  scratch[s.vertex] = slot;
  if (is_root) E.assign[block][row_ids ? (long long)row_ids[i] : row0 + i] = slot;
}

// pack the scratch records of the rows that requested a new row: rec[rank][0] = row id, rec[rank][1..] = cells
__global__ void k_pack_requests(const Dev* __restrict__ Ep, long long row0, long long nrows, const int* req, const int* rank, int* rec) {
  const Dev& E = *Ep;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows || req[i] < 0) return;
  int* out = rec + (long long)rank[i] * (E.nvC + 1);
  out[0] = (int)(row0 + i);
  const int* src = E.pool + (long long)req[i] * E.nvC;
  for (int v = 0; v < E.nvC; ++v) out[1 + v] = src[v];
}
// unpack gathered records into the scratch pool: pool[j] <- rec[src[j]], req[j] = j, row_ids[j] = row id
__global__ void k_unpack_requests(const Dev* __restrict__ Ep, int n, int pool_base, const int* rec, const int* src, int* req, int* row_ids) {
  const Dev& E = *Ep;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int* in = rec + (long long)src[j] * (E.nvC + 1);
  row_ids[j] = in[0];
  int* dst = E.pool + (long long)(pool_base + j) * E.nvC;      // tail of the pool: live proposals of later blocks stay intact
  for (int v = 0; v < E.nvC; ++v) dst[v] = in[1 + v];
  req[j] = pool_base + j;
}
__global__ void k_req_flags(long long nrows, const int* req, int* flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= nrows) flags[i] = (i < nrows && req[i] >= 0) ? 1 : 0;
}
// sufficient statistics of a ProportionsParameter (choose_proportionally.jl:57-68, batch form):
// number of live rows of the owning class whose choice equals each option
__global__ void k_option_counts(const TableD* tables, int t, int col, const int* optsid, int nopt, int* counts) {
  const TableD& T = tables[t];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T.n_slots || T.refcnt[j] <= 0) return;
  const int v = T.cells[(long long)col * T.cap + j];
  for (int o = 0; o < nopt; ++o) if (optsid[o] == v) { atomicAdd(&counts[o], 1); return; }
}
// sufficient statistics of a MeanParameter (add_noise.jl:48-71, transformed_gaussian.jl:26-33, batch
// form): every observation row contributes x * scale to the parameter slot its mean resolves to.
struct GaussSiteD { int obs_col; int lookup; int nargs; int direct_slot; TraceArgD args[3]; TraceArgD xform; };
__device__ __forceinline__ int trace_arg(const Dev& E, const TraceArgD& a, long long r) {
  if (a.kind == 0) return a.a;
  if (a.kind == 1) {
    int s = a.a >= 0 ? E.obs_sid[a.a][r] : -1;
    if (s < 0 && E.rowcell[a.b]) s = E.rowcell[a.b][r];
    return s;
  }
  const TableD& T = E.tables[a.b];
  const int slot = E.assign[a.a][r];
  return slot >= 0 ? T.cells[(long long)a.c * T.cap + slot] : -1;
}
__global__ void k_gauss_site(const Dev* Ep, GaussSiteD S, long long r0, long long r1, long long N, int* slot_of_row, double* x_of_row, int* iota) {
  const Dev& E = *Ep;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= N) return;
  iota[r] = (int)r;
  int slot = 0x7fffffff; double x = 0.0;
  if (r >= r0 && r < r1) {
    const double v = E.obs_real[S.obs_col][r];
    const int xf = trace_arg(E, S.xform, r);
    if (v == v && xf >= 0) {
      if (S.lookup >= 0) {
        int k[3] = {0, 0, 0};
        bool ok = true;
        for (int a = 0; a < S.nargs; ++a) { k[a] = trace_arg(E, S.args[a], r); ok = ok && k[a] >= 0; }
        const int f = ok ? lookup_find(E.lookups[S.lookup], k[0], k[1], k[2]) : PCL_LOOKUP_EMPTY;
        if (f != PCL_LOOKUP_EMPTY) slot = f;
      } else slot = S.direct_slot;
      x = v * E.xform_scale[xf];
    }
  }
  slot_of_row[r] = slot; x_of_row[r] = x;
}
// rows sorted by slot (stable: ascending row inside a slot): the head of each run sums its run
// sequentially, so the moments do not depend on the launch geometry.
__global__ void k_segment_moments(const int* keys, const int* rows, long long n, const double* x_of_row, double* msum, double* mcnt) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = keys[i];
  if (k == 0x7fffffff || (i > 0 && keys[i - 1] == k)) return;
  double s = 0.0; long long c = 0;
  for (long long j = i; j < n && keys[j] == k; ++j) { s += x_of_row[rows[j]]; ++c; }
  msum[k] += s; mcnt[k] += (double)c;
}
// ---- MaybeSwap --------------------------------------------------------------------------------
__device__ __forceinline__ int lookup_ref(const Dev& E, const LookupRefD& L, long long r, int esid) {
  int k[3] = {0, 0, 0};
  for (int a = 0; a < L.nargs; ++a) {
    k[a] = L.args[a].kind == 3 ? esid : trace_arg(E, L.args[a], r);
    if (k[a] < 0) return PCL_LOOKUP_EMPTY;
  }
  return lookup_find(E.lookups[L.lookup], k[0], k[1], k[2]);
}
__device__ __forceinline__ int mswap_list(const Dev& E, const MswapD& M, long long r, int esid) {
  return M.list_const >= 0 ? M.list_const : lookup_ref(E, M.list, r, esid);
}
__device__ __forceinline__ double mswap_prob(const Dev& E, const MswapD& M, long long r, int esid, int* slot_out) {
  if (slot_out) *slot_out = -1;
  if (M.prob_kind == 0) return M.prob_const;
  int v = M.prob_kind == 1 ? M.prob_slot : lookup_ref(E, M.prob, r, esid);
  if (v == PCL_LOOKUP_EMPTY) { atomicExch(E.err, PCLEAN_ERR_LOOKUP); return 0.5; }
  if (v >= 0) { if (slot_out) *slot_out = v; return E.param_real[v]; }
  return E.lkconst[-2 - v];
}
__device__ __forceinline__ double mswap_logdensity(const Dev& E, int obs, int val, int list, double p) {
  const int n = list >= 0 ? E.lists_off[list + 1] - E.lists_off[list] : 0;
  if (obs < 0) {                                  // explicit missing observation
    for (int i = 0; i < n; ++i) if (E.lists_sid[E.lists_off[list] + i] == val) return 0.0;
    return -1000.0;
  }
  if (val == obs) return log1p(-p);
  return log(p) - log((double)n);
}
// cell `vertex` / `cell` of the row particle k chose in an earlier block
__device__ __forceinline__ int particle_cell(const Dev& E, const RefCellD& cell, int vertex, int k, long long r) {
  const int ch = E.pchoice[cell.block][PCL_PK(E, k, r)];
  if (ch == PCL_CHOICE_UNSET) return -1;
  if (ch >= 0) { const TableD& T = E.tables[cell.table]; return T.cells[(long long)cell.col * T.cap + ch]; }
  return E.pool[(long long)(-(ch) - 2) * E.nvC + vertex];
}
// A block without any enumeration (flights block 3): the incremental weight of a particle is the
// likelihood of the observed MaybeSwap cells given what its earlier blocks chose; absent cells are
// sampled with random() (block_proposal.jl:58-66).  One thread per (row, particle).
__global__ void k_rootless(const Dev* __restrict__ Ep, int prog_id, int block, long long row0, long long nrows, const long long* __restrict__ row_list,
                           uint64_t seed, uint32_t sweep, uint32_t cls, int csmc) {
  const Dev& E = *Ep;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * E.K) return;
  const long long ri = i / E.K; const int k = (int)(i % E.K);
  const long long r = row_list ? row_list[row0 + ri] : row0 + ri;
  const ProgD& P = E.progs[prog_id];
  const MswapD* ms = E.mswaps + P.ms0;
  double w = 0.0;
  for (int t = 0; t < P.n_rterm; ++t) {
    const MswapD& M = ms[t];
    const int obs = E.obs_sid[M.obs_col][r];
    const int val = particle_cell(E, M.val_cell, M.val_vertex, k, r);
    w += mswap_logdensity(E, obs, val, mswap_list(E, M, r, -1), mswap_prob(E, M, r, -1, nullptr));
  }
  for (int q = 0; q < P.n_rsamp; ++q) {
    const MswapD& M = ms[P.n_rterm + q];
    int v;
    if (csmc && k == 0) v = E.rowcell[M.vertex][r];                    // the retained particle keeps its value
    else {
      const int val = particle_cell(E, M.val_cell, M.val_vertex, k, r);
      const int list = mswap_list(E, M, r, -1);
      const int n = list >= 0 ? E.lists_off[list + 1] - E.lists_off[list] : 0;
      pclean_stream st; st.key.seed = seed; st.key.sweep = sweep; st.key.cls = cls; st.key.row = r; st.key.particle = (uint32_t)k;
      st.key.block = (uint32_t)block; st.key.site = (uint32_t)M.vertex; st.key.purpose = PCLEAN_RNG_RANDOM; st.idx = 0;
      if (pclean_next(&st) < mswap_prob(E, M, r, -1, nullptr) && n > 0) {             // maybe_swap.jl:30-33
        const int j = min(n - 1, (int)(pclean_next(&st) * n));
        v = E.lists_sid[E.lists_off[list] + j];
      } else v = val;
    }
    E.pinner[block][PCL_PINNER(E, q, k, r)] = v;
  }
  E.pchoice[block][PCL_PK(E, k, r)] = PCL_CHOICE_UNSET;
  E.pweight[PCL_PK(E, k, r)] += w;
}
// sufficient statistics of the ProbParameters behind MaybeSwap nodes (maybe_swap.jl:65-85, batch
// form): counts[2 * slot] = observations that differ from the clean value, [2 * slot + 1] = equal
__global__ void k_mswap_counts(const Dev* __restrict__ Ep, MswapD M, long long r0, long long r1, int* counts) {
  const Dev& E = *Ep;
  const long long r = r0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= r1) return;
  int obs = M.obs_col >= 0 ? E.obs_sid[M.obs_col][r] : -1;
  if (obs < 0 && E.rowcell[M.vertex]) obs = E.rowcell[M.vertex][r];     // sampled cells count like observations (they live in the row)
  if (obs < 0 || E.assign[M.val_cell.block][r] < 0) return;
  const TableD& T = E.tables[M.val_cell.table];
  const int val = T.cells[(long long)M.val_cell.col * T.cap + E.assign[M.val_cell.block][r]];
  int slot = -1;
  mswap_prob(E, M, r, -1, &slot);
  if (slot >= 0) atomicAdd(&counts[2 * slot + (obs == val ? 1 : 0)], 1);
}

// keys of the hash index: key string id of every live slot (dead slots sort last), bucket sizes
__global__ void k_bucket_keys(const TableD* tables, int t, int col, int n_strings, int* keys, int* counts, int* iota) {
  const TableD& T = tables[t];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T.n_slots) return;
  int k = T.refcnt[j] > 0 ? T.cells[(long long)col * T.cap + j] : -1;
  if (k < 0 || k >= n_strings) k = n_strings;
  keys[j] = k; iota[j] = j;
  atomicAdd(&counts[k], 1);
}
__global__ void k_fill_u64(unsigned long long* p, long long n, unsigned long long v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_zero_int(int* p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}
// reference counts of the targets of the observation class (dependency_tracking.jl:227-228)
__global__ void k_count_assign(const int* assign, long long n, int* refcnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && assign[i] >= 0) atomicAdd(&refcnt[assign[i]], 1);      // < 0: row not initialised yet
}
// reference counts contributed by the live rows of a latent table through one of its slots
__global__ void k_count_table(const TableD* tables, int t, int g) {
  const TableD& T = tables[t];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T.n_slots || T.refcnt[j] <= 0) return;
  atomicAdd(&tables[T.fk_table[g]].refcnt[T.cells[(long long)T.fk_col[g] * T.cap + j]], 1);
}
__global__ void k_zero_refcnt(TableD* tables) {
  const TableD& T = tables[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (T.refcnt && j < T.cap) T.refcnt[j] = 0;
}
__global__ void k_table_stats(TableD* tables) {
  TableD& T = tables[blockIdx.x];
  if (!T.refcnt || T.cap <= 0) return;             // class without a loaded table
  __shared__ int s_alive; __shared__ long long s_refs; __shared__ int s_maxc;
  if (threadIdx.x == 0) { s_alive = 0; s_refs = 0; s_maxc = 0; }
  __syncthreads();
  int alive = 0; long long refs = 0; int maxc = 0;
  for (int j = threadIdx.x; j < T.cap; j += blockDim.x) {
    const int c = j < T.n_slots ? T.refcnt[j] : 0;
    T.logcnt[j] = c > 0 ? log((double)c - T.discount) : PCL_NEG_INF;
    T.logcnt1[j] = c > 1 ? log((double)(c - 1) - T.discount) : PCL_NEG_INF;
    T.alive[j] = c > 0 ? 1 : 0;
    alive += c > 0; refs += c; maxc = max(maxc, c);
  }
  atomicAdd(&s_alive, alive); atomicAdd((unsigned long long*)&s_refs, (unsigned long long)refs); atomicMax(&s_maxc, maxc);
  __syncthreads();
  if (threadIdx.x == 0) {
    T.n_alive = s_alive; T.total_refs = s_refs; T.max_logcnt = s_maxc > 0 ? log((double)s_maxc - T.discount) : 0.0;
    T.log_new = log(T.strength + T.discount * (double)s_alive); T.log_den = log((double)s_refs + T.strength);
    for (int x = 0; x <= PCL_MAX_EX; ++x) {
      T.log_new_x[x] = log(T.strength + T.discount * (double)(s_alive - x));
      T.log_den_x[x] = log((double)(s_refs - x) + T.strength);
    }
  }
}

// ------------------------------------------------------------------------------------------
// distance matrices: one CTA per pattern (unique observed string) and tile of elements
// ------------------------------------------------------------------------------------------
struct DpArgs {
  const uint8_t* sym; const int* str_off; const int* str_len;
  const int* pat_ids; int n_pat; int pat0;          // patterns pat0 .. pat0 + gridDim.y
  const int* elem_ids; int elem0; int n_elem;       // elements elem0 .. elem0 + n_elem (ids may be < 0: dead)
  int prefix_a, prefix_sep;                         // >= 0: clean string = a ++ sep ++ elem
  uint8_t* out; long long stride;                   // out[pat * stride + elem]
  uint8_t* elem_len;                                // optional (written by pattern row 0 of the launch)
  int words;                                        // max words over patterns in this launch (<= OSA_MAX_WORDS)
  const int* col_list;                              // optional: the n_elem columns to (re)compute (else elem0 .. elem0 + n_elem)
};

__global__ void __launch_bounds__(128) k_dp_matrix(DpArgs A) {
  extern __shared__ uint64_t s_peq[];               // [256 * words]
  const int p = A.pat0 + blockIdx.y;
  const int pid = A.pat_ids[p];
  const int m = pid >= 0 ? A.str_len[pid] : 0;
  const int words = m > 64 ? (m + 63) >> 6 : 1;      // per-pattern word count (<= A.words)
  for (int i = threadIdx.x; i < 256 * words; i += blockDim.x) s_peq[i] = 0;
  __syncthreads();
  if (pid >= 0) {
    const uint8_t* ps = A.sym + A.str_off[pid];
    for (int i = threadIdx.x; i < m; i += blockDim.x)
      atomicOr((unsigned long long*)&s_peq[(int)ps[i] * words + (i >> 6)], 1ull << (i & 63));
  }
  __syncthreads();
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < A.n_elem; e += gridDim.x * blockDim.x) {
    const int colx = A.col_list ? A.col_list[e] : A.elem0 + e;
    const int eid = A.elem_ids[colx];
    OsaText t; t.seg[0] = t.seg[1] = t.seg[2] = A.sym; t.len[0] = t.len[1] = t.len[2] = 0;
    int s = 0;
    if (A.prefix_a >= 0) {
      t.seg[0] = A.sym + A.str_off[A.prefix_a]; t.len[0] = A.str_len[A.prefix_a];
      t.seg[1] = A.sym + A.str_off[A.prefix_sep]; t.len[1] = A.str_len[A.prefix_sep];
      s = 2;
    }
    if (eid >= 0) { t.seg[s] = A.sym + A.str_off[eid]; t.len[s] = A.str_len[eid]; }
    const int n = t.len[0] + t.len[1] + t.len[2];
    int d = (pid >= 0 && eid >= 0) ? osa_distance(s_peq, m, words, t) : 255;
    if (d > 255) d = 255;
    A.out[(long long)p * A.stride + colx] = (uint8_t)d;
    if (A.elem_len && blockIdx.y == 0) A.elem_len[colx] = (uint8_t)(n > 255 ? 255 : n);
  }
}

// distinct values per column among the live rows of table t, hashed into 65536 bits (an estimate
// that saturates): ranks the likelihood terms of a star by how selective they are
__global__ void __launch_bounds__(256) k_col_diversity(TableD* tables, int t) {
  __shared__ unsigned bits[2048];
  __shared__ int total;
  TableD& T = tables[t];
  const int col = blockIdx.x;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) bits[i] = 0u;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < T.n_slots; j += blockDim.x) {
    if (T.refcnt[j] <= 0) continue;
    unsigned h = (unsigned)T.cells[(long long)col * T.cap + j] * 2654435761u;
    h ^= h >> 15; h &= 65535u;
    atomicOr(&bits[h >> 5], 1u << (h & 31u));
  }
  __syncthreads();
  int cnt = 0;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) cnt += __popc(bits[i]);
  atomicAdd(&total, cnt);
  __syncthreads();
  if (threadIdx.x == 0) T.div[col] = total;
}
// Order in which the pruning pass of k_block reads the terms of each star: expected distance of a
// random candidate ~ (mean observed length) x (share of candidates that differ), largest first.
// One warp per program; order_out holds star-local term indices.
__global__ void k_term_order(const Dev* __restrict__ Ep, int n_progs, int* order_out) {
  const Dev& E = *Ep;
  if ((int)blockIdx.x >= n_progs) return;
  const ProgD& P = E.progs[blockIdx.x];
  const int lane = threadIdx.x;
  for (int si = 0; si < P.nstar; ++si) {
    const StarD& s = E.stars[P.star0 + si];
    const int n = s.nterm;
    if (n <= 0) continue;
    if (n > 32) { for (int i = lane; i < n; i += 32) order_out[P.term0 + s.term0 + i] = i; continue; }
    float prio = -1.0f;
    if (lane < n) {
      const TermD& tm = E.terms[P.term0 + s.term0 + lane];
      prio = E.col_meanlen[tm.obs_col];
      if (tm.ptable >= 0) { const int d = E.tables[tm.ptable].div[tm.pcol]; prio *= (float)min(d, 256) * (1.0f / 256.0f); }
    }
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float pj = __shfl_sync(0xffffffffu, prio, j);
      rank += (pj > prio || (pj == prio && j < lane)) ? 1 : 0;
    }
    if (lane < n) order_out[P.term0 + s.term0 + rank] = lane;
  }
}

// hoisted choice-star marginals: hoist[u] = LSE_o(prior[o] + AddTypos(obs_u | option o))
__global__ void __launch_bounds__(128) k_hoist(const Dev* __restrict__ Ep, int prog_id, int sidx, int n_u, double* out) {
  __shared__ double sLG[PCL_LG_N];
  __shared__ double sLOGN[256];
  __shared__ double s_part[4];
  const Dev& E = *Ep;
  for (int i = threadIdx.x; i < PCL_LG_N; i += blockDim.x) sLG[i] = E.LG[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sLOGN[i] = E.LOGN[i];
  __syncthreads();
  const ProgD& P = E.progs[prog_id];
  const StarD& s = E.stars[P.star0 + sidx];
  const TermD& t = E.terms[P.term0 + s.term0];
  const MatD M = E.mats[t.mat];
  for (int u = blockIdx.x; u < n_u; u += gridDim.x) {
    Lse acc; acc.m = PCL_NEG_INF; acc.s = 0.0;
    for (int o = threadIdx.x; o < s.nopt; o += blockDim.x) {
      const int k = M.d[(long long)u * M.stride + o];
      lse_add(acc, E.prior_pool[s.prior_off + o] + addtypos_score(k, M.elen[o], t.max_typos, sLG, sLOGN));
    }
    const double v = lse_warp(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double mm = fmax(fmax(s_part[0], s_part[1]), fmax(s_part[2], s_part[3]));
      double ss = 0.0;
      for (int w = 0; w < 4; ++w) if (s_part[w] != PCL_NEG_INF) ss += exp(s_part[w] - mm);
      out[u] = mm == PCL_NEG_INF ? PCL_NEG_INF : mm + log(ss);
    }
    __syncthreads();
  }
}

// log of a proportions parameter into the prior pool (ChooseProportionally prior, utils.jl:33-36)
__global__ void k_log_prior(const double* value, int n, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = log(value[i]);
}

// plain pair kernel behind pclean_addtypos_pairs
__global__ void k_pairs(const uint8_t* sym, const int* str_off, const int* str_len, long long n, const int* obs,
                        const int* clean, int max_typos, const double* LG, const double* LOGN, int* dist, double* logd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = obs[i], b = clean[i];
  const int m = str_len[a], nb = str_len[b];
  // direct two-row DP (tiny helper kernel; the production path is k_dp_matrix)
  const uint8_t* A = sym + str_off[a]; const uint8_t* B = sym + str_off[b];
  int pp[257], p[257], c[257];
  for (int j = 0; j <= nb; ++j) p[j] = j;
  for (int x = 1; x <= m; ++x) {
    c[0] = x;
    for (int j = 1; j <= nb; ++j) {
      const int cost = A[x - 1] == B[j - 1] ? 0 : 1;
      int v = min(min(p[j] + 1, c[j - 1] + 1), p[j - 1] + cost);
      if (x > 1 && j > 1 && A[x - 1] == B[j - 2] && A[x - 2] == B[j - 1]) v = min(v, pp[j - 2] + 1);
      c[j] = v;
    }
    for (int j = 0; j <= nb; ++j) { pp[j] = p[j]; p[j] = c[j]; }
  }
  const int d = m == 0 ? nb : p[nb];
  dist[i] = d;
  logd[i] = addtypos_score(d > 255 ? 255 : d, nb > 255 ? 255 : nb, max_typos, LG, LOGN);
}

}  // namespace pcl
