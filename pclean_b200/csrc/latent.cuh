// latent.cuh — moves of latent-class rows (run_smc! on a class with incoming references:
// row_inference.jl:108-187 with collect_referring_rows :23-47, ExternalLikelihood enumeration
// proposal_compiler.jl:306-350 and block_proposal.jl:119-155).
//
// One warp per latent row.  The block's plan is a forest of independent *sites* (a discrete
// choice or a reference slot), each lowered to a star whose terms are summed over the
// observation rows that (transitively) refer to the row.  Exact enumeration makes every
// particle's weight the product of the site marginals, so particles differ only in the values
// they sample; lane k <-> particle k, particle 0 keeps the retained row.
#pragma once
#include "device.cuh"

namespace pcl {

__device__ __forceinline__ int refcell_sid(const Dev& E, const RefCellD& rc, long long r) {
  const TableD& T = E.tables[rc.table];
  return T.cells[(long long)rc.col * T.cap + E.assign[rc.block][r]];
}

// ---- referrers grouped by what they observe ----------------------------------------------------
// An external AddTypos term of a latent move is a sum over the observation rows referring to the
// latent row; rows that observe the same string contribute the same score, so the sum runs over
// the DISTINCT observed strings with their multiplicities (exact: score x count).  One group set per
// (dataset column [, cell of the referring row a string join reads]); keys are sorted, so the groups
// of a latent row are a contiguous range found by binary search.
#define PCL_GRP_SLOT_SHIFT 44
#define PCL_GRP_REF_SHIFT 22
#define PCL_GRP_MASK22 0x3FFFFFull
struct GroupSetD { int obs_col; int has_ref; RefCellD ref; };
__global__ void k_group_keys(const Dev* __restrict__ Ep, GroupSetD G, const int* __restrict__ slot_of_row, long long n, unsigned long long* keys) {
  const Dev& E = *Ep;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int slot = slot_of_row[r];
  unsigned long long k = ~0ull;                                  // rows that refer to nothing sort last
  if (slot != 0x7fffffff) {
    const int u = E.uobs[G.obs_col][r];
    unsigned long long ref = 0;
    if (G.has_ref) ref = (unsigned long long)(unsigned)(refcell_sid(E, G.ref, r) + 1) & PCL_GRP_MASK22;
    k = ((unsigned long long)(unsigned)slot << PCL_GRP_SLOT_SHIFT) | (ref << PCL_GRP_REF_SHIFT) | ((unsigned long long)(unsigned)(u + 1) & PCL_GRP_MASK22);
  }
  keys[r] = k;
}
// first group of `slot` in group set g (lower bound of slot << 44)
__device__ __forceinline__ int grp_lower(const Dev& E, int g, int slot) {
  const unsigned long long* keys = E.lgrp_key[g];
  const unsigned long long want = (unsigned long long)(unsigned)slot << PCL_GRP_SLOT_SHIFT;
  int lo = 0, hi = E.lgrp_n[g];
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
  return lo;
}

// per-row preparation of a choice star whose option list depends on a cell of the row being moved
// (rents County: possibilities[countykey]): list id and dummy mass (string_prior.jl:19-20)
__device__ void lstar_prepare(const RowCtx& c, const StarD& s) { PCL_CTX(c);
  if (s.kind != 1 || s.list_func < 0) return;
  const Dev& E = *cE;
  const TableD& TT = E.tables[cP->cls];
  const int key = s.list_own_col >= 0 ? TT.cells[(long long)s.list_own_col * TT.cap + cR] : -1;
  int l = -1;
  if (key >= 0) { l = lookup_find(E.lookups[s.list_func], key, 0, 0); if (l == PCL_LOOKUP_EMPTY) l = -1; }
  const int n = l >= 0 ? E.lists_off[l + 1] - E.lists_off[l] : 0;
  Lse a; a.m = PCL_NEG_INF; a.s = 0.0;
  for (int j = cLane; j < n; j += 32) lse_add(a, E.splp_pool[s.splp_off + E.lists_sid[E.lists_off[l] + j]]);
  const double tot = lse_warp(a);
  const int sidx = star_index(c, s);
  if (cLane == 0) { cW->lst[sidx] = l; cW->aux[sidx] = log1p(-exp(tot)); }
  __syncwarp();
}

// value id of an argument of an external lookup for referring row r and enumerated element (esid | candidate slot)
__device__ __forceinline__ int ext_arg(const RowCtx& c, const StarD& s, const TraceArgD& a, long long r, int esid, int slot) { PCL_CTX(c);
  if (a.kind <= 2) return trace_arg(*cE, a, r);
  if (a.kind == 3) return esid;
  if (a.kind == 4) { const TableD& T = cE->tables[s.table]; return T.cells[(long long)a.a * T.cap + slot]; }
  const TableD& TT = cE->tables[cP->cls];
  return TT.cells[(long long)a.a * TT.cap + cR];
}
// sum over the referring rows of the TransformedGaussian log-density (transformed_gaussian.jl:15-16)
__device__ __noinline__ double gauss_ext_sum(const RowCtx& c, const StarD& s, const GaussExtD& G, int esid, int slot) { PCL_CTX(c);
  const Dev& E = *cE;
  double acc = 0.0;
  for (int ri = 0; ri < cNref; ++ri) {
    const long long r = cRefs[ri];
    const double v = E.obs_real[G.obs_col][r];
    if (!(v == v)) continue;                                  // missing observation: log-density 0
    int k[3] = {0, 0, 0};
    bool ok = true;
    for (int a = 0; a < G.nargs; ++a) { k[a] = ext_arg(c, s, G.args[a], r, esid, slot); ok = ok && k[a] >= 0; }
    const int ps = ok ? lookup_find(E.lookups[G.lookup], k[0], k[1], k[2]) : PCL_LOOKUP_EMPTY;
    if (ps == PCL_LOOKUP_EMPTY) { atomicExch(E.err, PCLEAN_ERR_LOOKUP); return PCL_NEG_INF; }
    const int xf = ext_arg(c, s, G.xform, r, esid, slot);
    const double sc = E.xform_scale[xf];
    const double z = (v * sc - E.param_real[ps]) / G.stdev;
    acc += -0.5 * z * z - log(G.stdev) - 0.91893853320467274178 - log(fabs(1.0 / sc));
  }
  return acc;
}

// one element of a star whose elements do not sit in consecutive matrix columns (row-dependent
// option lists) or that carries Gaussian external terms: one lane per element
__device__ __noinline__ double lstar_elem_generic(const RowCtx& c, const StarD& s, int j, int J) { PCL_CTX(c);
  const Dev& E = *cE;
  if (j >= J) return PCL_NEG_INF;
  double l; int esid = -1, slot = -1, col_index = j;
  if (s.kind == 0) {
    const TableD& T = E.tables[s.table];
    slot = j;
    int cnt = T.refcnt[j];
    const int e = cW->n_ex ? excl_count(cW, s.table, j) : 0;
    cnt -= e;
    if (cnt <= 0) return PCL_NEG_INF;
    l = e ? log((double)cnt - T.discount) : T.logcnt[j];
  } else if (s.list_func >= 0) {
    esid = star_option_sid(c, s, j);
    l = j < J - 1 ? E.splp_pool[s.splp_off + esid] : cW->aux[star_index(c, s)];
    col_index = E.univ_col[s.univ_off + esid];
  } else {
    l = E.prior_pool[s.prior_off + j];
    esid = E.optsid_pool[s.opt_off + j];
  }
  const TermD* terms = E.terms + cP->term0;
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) {
    const TermD& tm = terms[t];
    if (tm.kind == 6) { l += gauss_ext_sum(c, s, E.gext[tm.mat], esid, slot); continue; }
    if (tm.kind == 7) {                       // MaybeSwap likelihood of every referring row given this option (maybe_swap.jl:13-28)
      const MswapD& M = E.mswaps[tm.mat];
      for (int ri = 0; ri < cNref; ++ri) {
        const long long r = cRefs[ri];
        int obs = M.obs_col >= 0 ? E.obs_sid[M.obs_col][r] : -1;
        if (obs < 0 && E.rowcell[M.vertex]) obs = E.rowcell[M.vertex][r];      // absent in the dataset: the value the row sampled
        // obs < 0 now means an explicit `missing` observation: 0 if the value is one of the options, else -1000
        l += mswap_logdensity(E, obs, esid, mswap_list(E, M, r, esid), mswap_prob(E, M, r, esid, nullptr));
      }
      continue;
    }
    if (tm.kind == TERM_JOIN_INLINE) { atomicExch(E.err, PCLEAN_ERR_UNSUPPORTED); continue; }
    if (tm.lmat >= 0) {
      // per-list distance blocks (device.cuh ListMatD): column = position of the option in the row's list;
      // the row of each referrer's observed string is looked up, a pair the dataset never showed costs a DP
      const int lid = cW->lst[star_index(c, s)];
      const ListMatD& LM = E.lmats[tm.lmat];
      const int Lj = lid >= 0 ? LM.elen[LM.elen_off[lid] + j] : min(255, E.str_len[esid]);
      if (tm.grp >= 0) {
        const unsigned long long* gk = E.lgrp_key[tm.grp]; const int* gc = E.lgrp_cnt[tm.grp];
        for (int gi = cW->glo[t]; gi < cW->ghi[t]; ++gi) {
          const int u = (int)(gk[gi] & PCL_GRP_MASK22) - 1;
          if (u < 0) continue;
          const uint8_t* rp = lid >= 0 ? lmat_row(&E, tm.lmat, u, lid) : nullptr;
          const int k = rp ? rp[j] : lmat_inline(&E, E.ulist[tm.obs_col][u], esid);
          l += (double)gc[gi] * score_fast(k, Lj, tm.max_typos, cLG, cLOGN, cLUT);
        }
      } else {
        for (int ri = 0; ri < cNref; ++ri) {
          const int u = E.uobs[tm.obs_col][cRefs[ri]];
          if (u < 0) continue;
          const uint8_t* rp = lid >= 0 ? lmat_row(&E, tm.lmat, u, lid) : nullptr;
          const int k = rp ? rp[j] : lmat_inline(&E, E.ulist[tm.obs_col][u], esid);
          l += score_fast(k, Lj, tm.max_typos, cLG, cLOGN, cLUT);
        }
      }
      continue;
    }
    const MatD M = E.mats[tm.mat];
    const int L = M.elen[col_index];
    if (tm.grp >= 0) {                         // distinct observed strings x multiplicity
      const unsigned long long* gk = E.lgrp_key[tm.grp]; const int* gc = E.lgrp_cnt[tm.grp];
      for (int gi = cW->glo[t]; gi < cW->ghi[t]; ++gi) {
        const int u = (int)(gk[gi] & PCL_GRP_MASK22) - 1;
        if (u < 0) continue;
        l += (double)gc[gi] * score_fast(M.d[(long long)u * M.stride + col_index], L, tm.max_typos, cLG, cLOGN, cLUT);
      }
      continue;
    }
    for (int ri = 0; ri < cNref; ++ri) {
      const int u = E.uobs[tm.obs_col][cRefs[ri]];
      if (u < 0) continue;
      l += score_fast(M.d[(long long)u * M.stride + col_index], L, tm.max_typos, cLG, cLOGN, cLUT);
    }
  }
  return l;
}
__device__ __forceinline__ bool lstar_is_generic(const RowCtx& c, const StarD& s) { PCL_CTX(c);
  if (s.kind == 1 && s.list_func >= 0) return true;
  const TermD* terms = cE->terms + cP->term0;
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) if (terms[t].kind == 6 || terms[t].kind == 7) return true;
  return false;
}

// log-scores of elements j0..j0+3 of star s for the latent row of `c` (all 32 lanes must call:
// inline joins build their match masks cooperatively)
// `want`: bit q set = element j0 + q is needed (inline joins cost one DP per element and referrer group; the
// caller that scores a single survivor does not pay for its three neighbours, whose l[] is then meaningless)
__device__ __noinline__ void lstar_tile4(const RowCtx& c, const StarD& s, int j0, int J, double l[4], unsigned want = 0xFu) { PCL_CTX(c);
  const Dev& E = *cE;
  if (lstar_is_generic(c, s)) {
    #pragma unroll
    for (int q = 0; q < 4; ++q) l[q] = ((want >> q) & 1u) ? lstar_elem_generic(c, s, j0 + q, J) : PCL_NEG_INF;
    return;
  }
  const TableD* T = s.kind == 0 ? &E.tables[s.table] : nullptr;
  #pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = j0 + q;
    if (j >= J) { l[q] = PCL_NEG_INF; continue; }
    if (T) {
      int cnt = T->refcnt[j];
      const int e = cW->n_ex ? excl_count(cW, s.table, j) : 0;
      cnt -= e;
      l[q] = cnt > 0 ? (e ? log((double)cnt - T->discount) : T->logcnt[j]) : PCL_NEG_INF;
    } else l[q] = E.prior_pool[s.prior_off + j];
  }
  const TermD* terms = E.terms + cP->term0;
  const int jj = min(j0, max(0, ((J + 3) & ~3) - 4));     // safe aligned address for out-of-range lanes
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) {
    const TermD& tm = terms[t];
    if (tm.kind == TERM_JOIN_INLINE) {
      // one pass per referring row, or — with a group set — per distinct (observed string, other half) with its multiplicity
      const bool grouped = tm.grp >= 0;
      const unsigned long long* gk = grouped ? E.lgrp_key[tm.grp] : nullptr; const int* gc = grouped ? E.lgrp_cnt[tm.grp] : nullptr;
      const int i0 = grouped ? cW->glo[t] : 0, i1 = grouped ? cW->ghi[t] : cNref;
      for (int ri = i0; ri < i1; ++ri) {
        int u, refsid = -1; double mult = 1.0; long long r = 0;
        if (grouped) { const unsigned long long k = gk[ri]; u = (int)(k & PCL_GRP_MASK22) - 1; refsid = (int)((k >> PCL_GRP_REF_SHIFT) & PCL_GRP_MASK22) - 1; mult = (double)gc[ri]; }
        else { r = cRefs[ri]; u = E.uobs[tm.obs_col][r]; }
        if (u < 0) continue;                                 // explicit missing observation
        const int psid = E.ulist[tm.obs_col][u];
        const int m = E.str_len[psid];
        __syncwarp();
        for (int i = cLane; i < 256; i += 32) cPeq[i] = 0ull;
        __syncwarp();
        const uint8_t* ps = E.sym + E.str_off[psid];
        for (int i = cLane; i < min(m, 64); i += 32) atomicOr(&cPeq[ps[i]], 1ull << i);
        __syncwarp();
        if (m > 64) { if (cLane == 0) atomicExch(E.err, PCLEAN_ERR_UNSUPPORTED); continue; }
        const int fa = tm.a_kind == OP_REFROW ? (grouped ? refsid : refcell_sid(E, tm.a_cell, r)) : -1;
        const int fb = tm.b_kind == OP_REFROW ? (grouped ? refsid : refcell_sid(E, tm.b_cell, r)) : -1;
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = j0 + q;
          if (j >= J || l[q] == PCL_NEG_INF || !((want >> q) & 1u)) continue;
          int esid;
          if (s.kind == 0) esid = -1; else esid = E.optsid_pool[s.opt_off + j];
          int a = fa, b = fb;
          if (tm.a_kind == OP_ELEM_OPT) a = esid; else if (tm.a_kind == OP_ELEM_COL) a = T->cells[(long long)tm.a_ref * T->cap + j];
          if (tm.b_kind == OP_ELEM_OPT) b = esid; else if (tm.b_kind == OP_ELEM_COL) b = T->cells[(long long)tm.b_ref * T->cap + j];
          OsaText tx;
          tx.seg[0] = E.sym + E.str_off[a]; tx.len[0] = E.str_len[a];
          tx.seg[1] = E.sym + E.str_off[tm.sep]; tx.len[1] = E.str_len[tm.sep];
          tx.seg[2] = E.sym + E.str_off[b]; tx.len[2] = E.str_len[b];
          int d = osa_distance((const uint64_t*)cPeq, m, 1, tx);
          const int L = min(255, tx.len[0] + tx.len[1] + tx.len[2]);
          l[q] += mult * score_fast(min(d, 255), L, tm.max_typos, cLG, cLOGN, cLUT);
        }
      }
      continue;
    }
    const MatD M = E.mats[tm.mat];
    const unsigned L4 = *reinterpret_cast<const unsigned*>(M.elen + jj);
    if (tm.grp >= 0) {                         // distinct observed strings x multiplicity
      const unsigned long long* gk = E.lgrp_key[tm.grp]; const int* gc = E.lgrp_cnt[tm.grp];
      for (int gi = cW->glo[t]; gi < cW->ghi[t]; ++gi) {
        const int u = (int)(gk[gi] & PCL_GRP_MASK22) - 1;
        if (u < 0) continue;
        const double mult = (double)gc[gi];
        const unsigned x = *reinterpret_cast<const unsigned*>(M.d + (long long)u * M.stride + jj);
        #pragma unroll
        for (int q = 0; q < 4; ++q)
          if (j0 + q < J) l[q] += mult * score_fast((x >> (8 * q)) & 255u, (L4 >> (8 * q)) & 255u, tm.max_typos, cLG, cLOGN, cLUT);
      }
      continue;
    }
    for (int ri = 0; ri < cNref; ++ri) {
      const long long r = cRefs[ri];
      const int u = E.uobs[tm.obs_col][r];
      if (u < 0) continue;
      const unsigned x = *reinterpret_cast<const unsigned*>(M.d + (long long)u * M.stride + jj);
      #pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j0 + q < J) l[q] += score_fast((x >> (8 * q)) & 255u, (L4 >> (8 * q)) & 255u, tm.max_typos, cLG, cLOGN, cLUT);
    }
  }
}

// One element of a (non-generic) star scored by the WHOLE warp: the referrer groups of every term are spread
// over the lanes (lane <-> group) instead of being walked by each lane for its own elements — the shape of
// the pruned path, where one or two survivors face the thousands of groups of a popular row.  For an inline
// join the CLEAN string is the DP's pattern (the distance is symmetric): groups are sorted by (other half of
// the join, observed string), so one set of match masks serves the whole run of groups that share the other
// half and the lanes take one observed string each.  Every lane returns the same value (same bits).
__device__ __noinline__ double lstar_elem_coop(const RowCtx& c, const StarD& s, int j, int J) { PCL_CTX(c);
  const Dev& E = *cE;
  if (j >= J) return PCL_NEG_INF;
  const TableD* T = s.kind == 0 ? &E.tables[s.table] : nullptr;
  double base;
  if (T) {
    int cnt = T->refcnt[j];
    const int e = cW->n_ex ? excl_count(cW, s.table, j) : 0;
    cnt -= e;
    if (cnt <= 0) return PCL_NEG_INF;
    base = e ? log((double)cnt - T->discount) : T->logcnt[j];
  } else base = E.prior_pool[s.prior_off + j];
  const TermD* terms = E.terms + cP->term0;
  const int esid = s.kind == 0 ? -1 : E.optsid_pool[s.opt_off + j];
  double part = 0.0;
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) {
    const TermD& tm = terms[t];
    if (tm.kind == TERM_JOIN_INLINE) {
      const unsigned long long* gk = E.lgrp_key[tm.grp]; const int* gc = E.lgrp_cnt[tm.grp];
      int run0 = cW->glo[t]; const int i1 = cW->ghi[t];
      while (run0 < i1) {
        const unsigned long long k0 = gk[run0];
        const unsigned long long next = ((k0 >> PCL_GRP_REF_SHIFT) + 1ull) << PCL_GRP_REF_SHIFT;     // first key of the next (slot, other half)
        int lo = run0 + 1, hi = i1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (gk[mid] < next) lo = mid + 1; else hi = mid; }
        const int run1 = lo;
        const int refsid = (int)((k0 >> PCL_GRP_REF_SHIFT) & PCL_GRP_MASK22) - 1;
        int a = tm.a_kind == OP_REFROW ? refsid : -1, b = tm.b_kind == OP_REFROW ? refsid : -1;
        if (tm.a_kind == OP_ELEM_OPT) a = esid; else if (tm.a_kind == OP_ELEM_COL) a = T->cells[(long long)tm.a_ref * T->cap + j];
        if (tm.b_kind == OP_ELEM_OPT) b = esid; else if (tm.b_kind == OP_ELEM_COL) b = T->cells[(long long)tm.b_ref * T->cap + j];
        const uint8_t* seg[3] = {E.sym + E.str_off[a], E.sym + E.str_off[tm.sep], E.sym + E.str_off[b]};
        const int len[3] = {E.str_len[a], E.str_len[tm.sep], E.str_len[b]};
        const int m = len[0] + len[1] + len[2];
        if (m > 64) { if (cLane == 0) atomicExch(E.err, PCLEAN_ERR_UNSUPPORTED); run0 = run1; continue; }
        __syncwarp();
        for (int i = cLane; i < 256; i += 32) cPeq[i] = 0ull;
        __syncwarp();
        for (int i = cLane; i < m; i += 32) {
          const uint8_t sym = i < len[0] ? seg[0][i] : (i < len[0] + len[1] ? seg[1][i - len[0]] : seg[2][i - len[0] - len[1]]);
          atomicOr(&cPeq[sym], 1ull << i);
        }
        __syncwarp();
        for (int g = run0 + cLane; g < run1; g += 32) {
          const int u = (int)(gk[g] & PCL_GRP_MASK22) - 1;
          if (u < 0) continue;                                 // explicit missing observation
          const int psid = E.ulist[tm.obs_col][u];
          OsaText tx;
          tx.seg[0] = E.sym + E.str_off[psid]; tx.len[0] = E.str_len[psid];
          tx.seg[1] = tx.seg[0]; tx.len[1] = 0; tx.seg[2] = tx.seg[0]; tx.len[2] = 0;
          const int d = osa_distance((const uint64_t*)cPeq, m, 1, tx);
          part += (double)gc[g] * score_fast(min(d, 255), min(255, m), tm.max_typos, cLG, cLOGN, cLUT);
        }
        run0 = run1;
      }
      __syncwarp();
      continue;
    }
    const MatD M = E.mats[tm.mat];
    const int Lj = M.elen[j];
    if (tm.grp >= 0) {
      const unsigned long long* gk = E.lgrp_key[tm.grp]; const int* gc = E.lgrp_cnt[tm.grp];
      for (int gi = cW->glo[t] + cLane; gi < cW->ghi[t]; gi += 32) {
        const int u = (int)(gk[gi] & PCL_GRP_MASK22) - 1;
        if (u < 0) continue;
        part += (double)gc[gi] * score_fast(M.d[(long long)u * M.stride + j], Lj, tm.max_typos, cLG, cLOGN, cLUT);
      }
    } else {
      for (int ri = cLane; ri < cNref; ri += 32) {
        const int u = E.uobs[tm.obs_col][cRefs[ri]];
        if (u < 0) continue;
        part += score_fast(M.d[(long long)u * M.stride + j], Lj, tm.max_typos, cLG, cLOGN, cLUT);
      }
    }
  }
  #pragma unroll
  for (int o = 16; o; o >>= 1) part += shfl_xor_d(part, o);
  return base + part;
}

__device__ __noinline__ double lstar_lse_raw(const RowCtx& c, const StarD& s) { PCL_CTX(c);
  const int J = star_nelem(c, s);
  Lse acc; acc.m = PCL_NEG_INF; acc.s = 0.0;
  for (int jb = 0; jb < J; jb += 128) {
    double l[4];
    lstar_tile4(c, s, jb + cLane * 4, J, l);
    #pragma unroll
    for (int q = 0; q < 4; ++q) lse_add(acc, l[q]);
  }
  if (cLane == 0) lse_add(acc, star_extra(c, s));
  return lse_warp(acc);
}

// ------------------------------------------------------------------------------------------
// Pruned evaluation of a choice star over a long constant option list (hospital: the unique observed
// values of a column, tens of thousands) inside a latent move.  With M grouped observations in all,
//   score(o) <= log prior(o) + sum_g m_g S(d_g(o), L_o) <= 0 - 0.10536 M - PCL_TYPO_COST W(o),
//   W(o) = sum_g m_g d(u_g, o)      (S(d, L) <= S(0, L) - 3.93 d  and  S(0, L) <= log 0.9),
// so once some option h (the row's current value, else the most observed string) has been scored
// exactly, only options with W(o) <= tau = (45 - 0.10536 M - score(h)) / 3.93 can matter.  W is an
// integer sum over the distance rows of the groups, largest group first: for a row with many
// referrers every other option is out after that one row of bytes.  Survivors are scored by the same
// code as the exhaustive path (lstar_tile4), so both paths give the same bits for the same option.
// ------------------------------------------------------------------------------------------
#define PCL_DBG(i_) do { if (cLane == 0 && cE->dbg) atomicAdd(&cE->dbg[(i_)], 1); } while (0)
__device__ __noinline__ bool lstar_eval_pruned(const RowCtx& c, const StarD& s, double* Lraw_out) { PCL_CTX(c);
  const Dev& E = *cE;
  WarpState* W = cW;
  const bool fk = s.kind == 0;
  if (!fk && (s.kind != 1 || s.list_func >= 0 || s.optidx_off < 0)) { PCL_DBG(1); return false; }
  if (fk && s.bucket) { PCL_DBG(2); return false; }
  const TableD* T = fk ? &E.tables[s.table] : nullptr;
  const int J = fk ? T->n_slots : s.nopt;
  const TermD* terms = E.terms + cP->term0;
  const int lane = cLane;
  const int plain_kind = fk ? TERM_CAND : TERM_OPT;
  // A short list is cheaper to score outright than to prune — unless an inline join hangs on it: those
  // run one edit-distance DP per (element, referrer group), and the most popular rows have thousands of
  // groups (H1M: a Hospital's State against the joined state_measure strings of up to 10^5 Records).
  if (J <= 2 * PCL_SURV_MAX) {
    long long inl = 0;
    for (int t = s.term0; t < s.term0 + s.nterm; ++t) if (terms[t].kind == TERM_JOIN_INLINE && terms[t].grp >= 0) inl += W->ghi[t] - W->glo[t];
    if (inl * J < 4096 || J <= 4) { PCL_DBG(3); return false; }
  }
  // every term must be a grouped distance-matrix term (joins are scored on the survivors only: dropping them keeps the bound valid)
  long long M_tot = 0; int best_cnt = 0, best_t = -1, best_gi = -1;
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) {
    const TermD& tm = terms[t];
    if (tm.kind == TERM_JOIN_INLINE && tm.grp >= 0) continue;
    if (tm.kind != plain_kind || tm.grp < 0) { PCL_DBG(4); return false; }
    const unsigned long long* gk = E.lgrp_key[tm.grp]; const int* gc = E.lgrp_cnt[tm.grp];
    for (int gi = W->glo[t] + lane; gi < W->ghi[t]; gi += 32) {
      if ((int)(gk[gi] & PCL_GRP_MASK22) == 0) continue;       // explicit missing observations score 0 for every element
      const int m = gc[gi];
      M_tot += m;
      if (m > best_cnt) { best_cnt = m; best_t = t; best_gi = gi; }
    }
  }
  for (int o = 16; o; o >>= 1) {
    M_tot += __shfl_xor_sync(0xffffffffu, M_tot, o);
    const int oc = __shfl_xor_sync(0xffffffffu, best_cnt, o), ot = __shfl_xor_sync(0xffffffffu, best_t, o), og = __shfl_xor_sync(0xffffffffu, best_gi, o);
    if (oc > best_cnt || (oc == best_cnt && oc > 0 && (ot < best_t || (ot == best_t && og < best_gi)))) { best_cnt = oc; best_t = ot; best_gi = og; }
  }
  if (best_cnt <= 0) { PCL_DBG(5); return false; }                              // nothing observed: prior mass only (exhaustive path)
  // hint: the element the row holds now (option equal to its current string / the row it references);
  // for options, else, the most observed string
  int hint = -1;
  {
    const TableD& TT = E.tables[cP->cls];
    const int cur = TT.cells[(long long)s.vertex * TT.cap + cR];
    if (fk) hint = cur;
    else {
      if (cur >= 0 && cur < E.n_strings) hint = E.optmap_pool[s.optidx_off + cur];
      if (hint < 0) {
        const TermD& tm = terms[best_t];
        const int u = (int)(E.lgrp_key[tm.grp][best_gi] & PCL_GRP_MASK22) - 1;
        hint = E.optmap_pool[s.optidx_off + E.ulist[tm.obs_col][u]];
      }
    }
  }
  // the score the others are measured against: the hinted element, or — reference-table stars — the
  // new-row branch when it is better (a row that is the last reference of its target has no usable hint:
  // its own exclusion empties the target; every other candidate is then measured against "a fresh row")
  double l4[4];
  double l0 = PCL_NEG_INF;
  // few elements against many groups: the warp scores one element at a time, groups over the lanes
  long long n_groups = 0;
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) n_groups += terms[t].grp >= 0 ? W->ghi[t] - W->glo[t] : (terms[t].kind == TERM_JOIN_INLINE ? 0 : cNref);
  bool coop = n_groups >= 64 && !lstar_is_generic(c, s);
  for (int t = s.term0; t < s.term0 + s.nterm; ++t) if (terms[t].kind == TERM_JOIN_INLINE && terms[t].grp < 0) coop = false;
  if (hint >= 0 && hint < J) {
    if (coop) l0 = lstar_elem_coop(c, s, hint, J);
    else {
      lstar_tile4(c, s, hint & ~3, J, l4, 1u << (hint & 3));    // all lanes: the join terms build their masks cooperatively
      l0 = l4[hint & 3];
    }
  } else if (!fk) { PCL_DBG(6); return false; }
  if (fk) l0 = fmax(l0, star_extra(c, s));
  if (l0 == PCL_NEG_INF) { PCL_DBG(7); return false; }
  const double Bmax = fk ? T->max_logcnt : 0.0;                 // log prior(o) <= 0; CRP term <= log(max count - discount)
  const double need = (Bmax + PCL_PRUNE_MARGIN - 0.10536051565782628 * (double)M_tot - l0) / PCL_TYPO_COST;
  if (!(need < 1.0e9)) { PCL_DBG(8); return false; }
  const unsigned tau = need < 0.0 ? 0u : (unsigned)need + 1u;
  const unsigned CLAMP = 1u << 30;
  // collect the elements with W(o) <= tau (ascending)
  if (lane == 0) W->sv_star = -1;
  int nsv = 0; bool overflow = false;
  const MatD Mb = E.mats[terms[best_t].mat];
  const int ub = (int)(E.lgrp_key[terms[best_t].grp][best_gi] & PCL_GRP_MASK22) - 1;
  const uint8_t* rowb = Mb.d + (long long)ub * Mb.stride;
  const unsigned mb = (unsigned)min(best_cnt, 1 << 20);
  const int J4 = (J + 3) & ~3;
  for (int jb = 0; jb < J4 && !overflow; jb += 128) {
    const int j0 = jb + lane * 4;
    unsigned acc[4] = {0u, 0u, 0u, 0u};
    bool live = j0 < J4;
    if (live) {
      const unsigned x = *reinterpret_cast<const unsigned*>(rowb + j0);
      #pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = mb * ((x >> (8 * q)) & 255u);
      live = acc[0] <= tau || acc[1] <= tau || acc[2] <= tau || acc[3] <= tau;
    }
    if (!__any_sync(0xffffffffu, live)) continue;
    // the other groups, until nobody in the chunk can still qualify
    for (int t = s.term0; t < s.term0 + s.nterm; ++t) {
      const TermD& tm = terms[t];
      if (tm.kind != plain_kind) continue;
      const MatD M = E.mats[tm.mat];
      const unsigned long long* gk = E.lgrp_key[tm.grp]; const int* gc = E.lgrp_cnt[tm.grp];
      bool dead = false;
      for (int gi = W->glo[t]; gi < W->ghi[t]; ++gi) {
        if (t == best_t && gi == best_gi) continue;
        const int u = (int)(gk[gi] & PCL_GRP_MASK22) - 1;
        if (u < 0) continue;
        if (live) {
          const unsigned m = (unsigned)min(gc[gi], 1 << 20);
          const unsigned x = *reinterpret_cast<const unsigned*>(M.d + (long long)u * M.stride + j0);
          #pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] = min(acc[q] + m * ((x >> (8 * q)) & 255u), CLAMP);
          live = acc[0] <= tau || acc[1] <= tau || acc[2] <= tau || acc[3] <= tau;
        }
        if (((gi - W->glo[t]) & 7) == 7 && !__any_sync(0xffffffffu, live)) { dead = true; break; }
      }
      if (dead) { live = false; break; }
    }
    unsigned keep = 0;
    if (live) {
      unsigned al = 0x01010101u;
      if (fk) al = *reinterpret_cast<const unsigned*>(T->alive + j0);       // bytes 0/1: rows nobody references are not candidates
      #pragma unroll
      for (int q = 0; q < 4; ++q) if (acc[q] <= tau && j0 + q < J && ((al >> (8 * q)) & 1u)) keep |= 1u << q;
    }
    const unsigned anyv = __ballot_sync(0xffffffffu, keep != 0);
    if (!anyv) continue;
    const int cnt = __popc(keep);
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    if (nsv + tot > PCL_SURV_MAX) { overflow = true; break; }
    int pos = nsv + incl - cnt;
    while (keep) { const int q = __ffs(keep) - 1; keep &= keep - 1; W->sv_idx[pos++] = j0 + q; }
    nsv += tot;
  }
  if (overflow) { PCL_DBG(9); return false; }
  __syncwarp();
  // exact scores of the survivors, by the code of the exhaustive path
  if (coop && (long long)nsv * 32 <= n_groups) {
    for (int i = 0; i < nsv; ++i) {
      const double v = lstar_elem_coop(c, s, W->sv_idx[i], J);
      if (lane == 0) W->sv_ll[i] = v;
    }
  } else for (int base = 0; base < nsv; base += 32) {
    const int i = base + lane;
    const int j = i < nsv ? W->sv_idx[i] : 0;
    lstar_tile4(c, s, j & ~3, J, l4, i < nsv ? 1u << (j & 3) : 0u);
    if (i < nsv) W->sv_ll[i] = l4[j & 3];
  }
  __syncwarp();
  if (fk) {                                                     // the new-row branch (its child stars were evaluated before: post-order)
    const double ex = star_extra(c, s);
    if (lane == 0) { W->sv_idx[nsv] = J; W->sv_ll[nsv] = ex; }
    nsv += 1;
  }
  if (lane == 0) { W->sv_n = nsv; W->sv_star = star_index(c, s); }
  __syncwarp();
  Lse acc2; acc2.m = PCL_NEG_INF; acc2.s = 0.0;
  for (int i = lane; i < nsv; i += 32) lse_add(acc2, W->sv_ll[i]);
  *Lraw_out = lse_warp(acc2);
  PCL_DBG(0);
  return true;
}

#undef PCL_DBG
// inverse-CDF draws: lane i holds uniform u (active lanes).  Elements in ascending order, the
// new-row branch (FK stars) last with index J.
__device__ __noinline__ int lstar_sample(const RowCtx& c, const StarD& s, double Lraw, double u, bool active) { PCL_CTX(c);
  const int J = star_nelem(c, s);
  double carry = 0.0; bool found = !active; int idx = -1, lastpos = -1;
  for (int jb = 0; jb < J; jb += 128) {
    double l[4], cs[4];
    lstar_tile4(c, s, jb + cLane * 4, J, l);
    double run = 0.0;
    #pragma unroll
    for (int q = 0; q < 4; ++q) { const double p = l[q] - Lraw < PCL_EXP_CUTOFF ? 0.0 : exp(l[q] - Lraw); run += p; cs[q] = run; l[q] = p; }
    double incl = run;
    for (int o = 1; o < 32; o <<= 1) { const double t = shfl_up_d(incl, o); if (cLane >= o) incl += t; }
    const double tot = shfl_d(incl, 31);
    const double before = incl - run;
    #pragma unroll
    for (int q = 0; q < 4; ++q) cs[q] += before;
    const unsigned pos = __ballot_sync(0xffffffffu, run > 0.0);
    if (pos) {
      const int hl = 31 - __clz(pos);
      int lq = -1;
      #pragma unroll
      for (int q = 0; q < 4; ++q) if (l[q] > 0.0) lq = q;
      lastpos = jb + hl * 4 + __shfl_sync(0xffffffffu, lq, hl);
    }
    const bool hit = !found && (u < carry + tot);
    if (__any_sync(0xffffffffu, hit)) {
      for (int i = 0; i < 32; ++i) {
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double ci = carry + shfl_d(cs[q], i);
          const double pi = shfl_d(l[q], i);
          if (hit && !found && pi > 0.0 && u < ci) { idx = jb + i * 4 + q; found = true; }
        }
      }
    }
    carry += tot;
  }
  if (s.kind == 0) {                       // new-row branch
    const double ex = star_extra(c, s);
    const double p = ex - Lraw < PCL_EXP_CUTOFF ? 0.0 : exp(ex - Lraw);
    if (p > 0.0) lastpos = J;
    if (!found && u < carry + p) { idx = J; found = true; }
  }
  if (active && idx < 0) idx = lastpos;
  return idx;
}

// evaluate the subtree of root `ridx` bottom-up (post-order segment of the program order)
__device__ void leval_site(const RowCtx& c, int o0, int o1) { PCL_CTX(c);
  const StarD* stars = cE->stars + cP->star0;
  for (int oi = o0; oi < o1; ++oi) {
    const int sidx = cP->order[oi];
    const StarD& s = stars[sidx];
    lstar_prepare(c, s);
    double raw;
    if (!(cE->prune && lstar_eval_pruned(c, s, &raw))) { if (cLane == 0) cW->sv_star = -1; __syncwarp(); raw = lstar_lse_raw(c, s); }
    const double v = raw - star_logden(c, s);
    if (cLane == 0) cW->V[sidx] = v;
    __syncwarp();
  }
}

// sample the contents of a new row under FK star `sroot` for particle k into scratch
__device__ __noinline__ void lexpand_new(const RowCtx& c, int sroot, int k, int block, int* scratch, uint64_t seed, uint32_t sweep, uint32_t cls, long long key, int* bad) { PCL_CTX(c);
  const StarD* stars = cE->stars + cP->star0;
  int stack[PCL_MAX_STARS]; int sp = 0;
  stack[sp++] = sroot;
  while (sp > 0) {
    const StarD& ps = stars[stack[--sp]];
    if (cLane == 0) scratch[ps.vertex] = -1;
    const int* ch = cE->children + ps.child0;
    for (int i = 0; i < ps.nchild; ++i) {
      const int cidx = ch[i];
      const StarD& cs = stars[cidx];
      const double u = row_uniform(seed, sweep, cls, key, k, block, cs.vertex, PCLEAN_RNG_ENUM);
      // sampled from the survivors of the pruned evaluation when the star qualifies (what the exhaustive scan would return)
      double Lraw;
      int e;
      if (cE->prune && lstar_eval_pruned(c, cs, &Lraw)) e = surv_sample(c, Lraw, u, true);
      else { Lraw = cW->V[cidx] + star_logden(c, cs); e = lstar_sample(c, cs, Lraw, u, true); }
      if (cs.kind == 1) {
        if (cLane == 0) {
          scratch[cs.vertex] = cE->optsid_pool[cs.opt_off + e];
          if (cs.has_dummy && e == cs.nopt - 1) { atomicOr(&cE->lflags[cR], ROWFLAG_DUMMY); *bad = 1; }
        }
      } else {
        const int J = cE->tables[cs.table].n_slots;
        if (e >= J) stack[sp++] = cidx;
        else {
          const TableD& T = cE->tables[cs.table];
          const int2* cp = cE->copies + cs.copy0;
          for (int q = cLane; q < cs.ncopy; q += 32) scratch[cp[q].x] = T.cells[(long long)cp[q].y * T.cap + e];
          __syncwarp();
          if (cLane == 0) scratch[cs.vertex] = e;
        }
      }
      __syncwarp();
    }
  }
}

#define PCL_KLATENT_SMEM (PCL_OFF_W + PCL_WARPS_PER_CTA * (sizeof(WarpState) + 256 * sizeof(unsigned long long) + PCL_MAX_SITES * PCL_MAX_K * sizeof(int)))

// k_latent: one warp per slot of latent class P.cls (persistent).  slot0/nslots select the range
// (debug: a single slot).
__global__ void __launch_bounds__(32 * PCL_WARPS_PER_CTA, 2)
k_latent(const Dev* __restrict__ Ep, int prog_id, int block, int n_blocks, int slot0, int nslots, const int* __restrict__ slot_list, uint64_t seed, uint32_t sweep, int use_mh) {
  double* sLUT = reinterpret_cast<double*>(pcl_smem);                          // layout: device.cuh PCL_OFF_*
  double* sLG = reinterpret_cast<double*>(pcl_smem + PCL_OFF_LG);
  double* sLOGN = reinterpret_cast<double*>(pcl_smem + PCL_OFF_LOGN);
  Dev* sE = reinterpret_cast<Dev*>(pcl_smem + PCL_OFF_DEV);
  ProgD* sP = reinterpret_cast<ProgD*>(pcl_smem + PCL_OFF_PROG);
  WarpState* sW = reinterpret_cast<WarpState*>(pcl_smem + PCL_OFF_W);
  unsigned long long* sPeq = reinterpret_cast<unsigned long long*>(sW + PCL_WARPS_PER_CTA);
  int* sChoice = reinterpret_cast<int*>(sPeq + PCL_WARPS_PER_CTA * 256);
  for (int i = threadIdx.x; i < (int)(sizeof(Dev) / 4); i += blockDim.x) reinterpret_cast<int*>(sE)[i] = reinterpret_cast<const int*>(Ep)[i];
  __syncthreads();
  const Dev& E = *sE;
  for (int i = threadIdx.x; i < PCL_LG_N; i += blockDim.x) sLG[i] = E.LG[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sLOGN[i] = E.LOGN[i];
  for (int i = threadIdx.x; i < PCL_LUT_N * PCL_LUT_N; i += blockDim.x) sLUT[i] = E.LUT[i];
  for (int i = threadIdx.x; i < (int)(sizeof(ProgD) / 4); i += blockDim.x) reinterpret_cast<int*>(sP)[i] = reinterpret_cast<const int*>(E.progs + prog_id)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ProgD& P = *sP;
  const StarD* stars = E.stars + P.star0;
  const TableD& TT = E.tables[P.cls];
  WarpState* W = &sW[warp];
  int* myChoice = sChoice + warp * PCL_MAX_SITES * PCL_MAX_K;       // [site][particle]
  const int K = E.K;
  const long long total_warps = (long long)gridDim.x * PCL_WARPS_PER_CTA;
  for (long long wid = (long long)blockIdx.x * PCL_WARPS_PER_CTA + warp; wid < nslots; wid += total_warps) {
    const int t = slot_list ? slot_list[wid] : slot0 + (int)wid;
    if (TT.refcnt[t] <= 0) { if (lane == 0) { E.lsel[t] = 0; E.lflags[t] = 0; E.llogml[t] = 0.0; } continue; }
    const long long key = TT.keys[t];
    RowCtx c;
    if (lane == 0) { W->row = t; W->refs = E.lref_rows + E.lref_off[t]; W->nref = E.lref_off[t + 1] - E.lref_off[t]; }
    // self-exclusion: the row's own outgoing references (unincorporate_row!, with the GC cascade)
    if (lane == 0) {
      int n = 0;
      int qt[PCL_MAX_EX], qs[PCL_MAX_EX]; int qh = 0, qn = 0;
      for (int g = 0; g < TT.nfk && qn < PCL_MAX_EX; ++g) { qt[qn] = TT.fk_table[g]; qs[qn] = TT.cells[(long long)TT.fk_col[g] * TT.cap + t]; ++qn; }
      while (qh < qn && n < PCL_MAX_EX) {
        const int tb = qt[qh], sl = qs[qh]; ++qh;
        int prev = 0;
        for (int i = 0; i < n; ++i) prev += (W->ex_table[i] == tb && W->ex_slot[i] == sl);
        const TableD& T2 = E.tables[tb];
        const int gc = (T2.refcnt[sl] - prev - 1) <= 0;
        W->ex_table[n] = tb; W->ex_slot[n] = sl; W->ex_gc[n] = gc; ++n;
        if (gc) for (int g = 0; g < T2.nfk && qn < PCL_MAX_EX; ++g) { qt[qn] = T2.fk_table[g]; qs[qn] = T2.cells[(long long)T2.fk_col[g] * T2.cap + sl]; ++qn; }
      }
      W->n_ex = n; W->sv_star = -1;
      E.lflags[t] = 0;
    }
    for (int tt = lane; tt < P.nterm; tt += 32) {
      const int g = E.terms[P.term0 + tt].grp;
      int lo = 0, hi = 0;
      if (g >= 0) { lo = grp_lower(E, g, t); hi = grp_lower(E, g, t + 1); }
      W->glo[tt] = lo; W->ghi[tt] = hi;
    }
    __syncwarp();
    double wsum = 0.0;
    int o0 = 0;
    unsigned my_badbits = 0;                        // bit p: this lane's particle of pass p carries a placeholder / lost its scratch record: never selected
    for (int si = 0; si < P.nroots; ++si) {
      const int ridx = P.roots[si];
      int o1 = o0;
      while (P.order[o1] != ridx) ++o1;
      ++o1;                                         // order[o0..o1) = subtree of this site, root last
      leval_site(c, o0, o1);
      o0 = o1;
      const StarD& root = stars[ridx];
      const double L = W->V[ridx];
      wsum += L;
      const double Lraw = L + star_logden(c, root);
      for (int pass = 0; pass * 32 < K; ++pass) {     // lane <-> particle pass * 32 + lane
      const int kk = pass * 32 + lane;
      bool my_bad = false;
      const bool draws = kk >= 1 && kk < K;          // particle 0 keeps the retained row
      double u = 0.0;
      if (draws) u = row_uniform(seed, sweep, (uint32_t)P.cls, key, kk, block, root.vertex, PCLEAN_RNG_ENUM);
      // survivors of the pruned evaluation still in shared memory: draw from them (what the exhaustive scan would return)
      const int e = W->sv_star == ridx ? surv_sample(c, Lraw, u, draws) : lstar_sample(c, root, Lraw, u, draws);
      int mine = e;
      if (root.kind == 1 && root.list_func >= 0 && e >= 0) mine = star_option_sid(c, root, e);   // row-dependent list: keep the value, not its position
      if (root.kind == 1 && root.has_dummy && draws && e == star_nelem(c, root) - 1) {
        if (root.dummy_time && root.list_func >= 0) {
          // the dummy stands for "some other time": random(TimePrior) (block_proposal.jl:58-60, time_prior.jl:20-22)
          pclean_stream st; st.key.seed = seed; st.key.sweep = sweep; st.key.cls = (uint32_t)P.cls; st.key.row = key; st.key.particle = (uint32_t)kk;
          st.key.block = (uint32_t)block; st.key.site = (uint32_t)root.vertex; st.key.purpose = PCLEAN_RNG_RANDOM; st.idx = 0;
          const int hh = min(11, (int)(pclean_next(&st) * 12)), mi = min(59, (int)(pclean_next(&st) * 60));
          const int pm = pclean_next(&st) < 0.5 ? 0 : 1;
          mine = E.time_sid[(hh * 60 + mi) * 2 + pm];
        } else { atomicOr(&E.lflags[t], ROWFLAG_DUMMY); my_bad = true; }
      }
      if (root.kind == 0) {
        const int J = E.tables[root.table].n_slots;
        unsigned newmask = __ballot_sync(0xffffffffu, draws && e >= J);
        while (newmask) {
          const int k = __ffs(newmask) - 1; newmask &= newmask - 1;
          int pidx = 0;
          if (lane == 0) pidx = atomicAdd(E.pool_count, 1);
          pidx = __shfl_sync(0xffffffffu, pidx, 0);
          if (pidx >= E.pool_cap) { if (lane == 0) { atomicOr(&E.lflags[t], ROWFLAG_POOL); atomicExch(E.err, PCLEAN_ERR_CAPACITY); } if (lane == k) { mine = -1; my_bad = true; } continue; }
          int* scratch = E.pool + (long long)pidx * E.nvC;
          for (int v = lane; v < E.nvC; v += 32) scratch[v] = PCL_UNSET;
          __syncwarp();
          int bad = 0;
          lexpand_new(c, ridx, pass * 32 + k, block, scratch, seed, sweep, (uint32_t)P.cls, key, &bad);
          bad = __shfl_sync(0xffffffffu, bad, 0);
          if (lane == k) { mine = -(pidx + 2); if (bad) my_bad = true; }
        }
      }
      if (kk < PCL_MAX_K) myChoice[si * PCL_MAX_K + kk] = mine;
      if (my_bad) my_badbits |= 1u << pass;
      }   // pass
      __syncwarp();
    }
    // final selection (row_inference.jl:157-165): every usable particle has the same weight; a
    // particle that drew a dummy placeholder (or lost its scratch record) has weight zero
    unsigned long long badmask = 0ull;
    for (int pass = 0; pass * 32 < K; ++pass)
      badmask |= (unsigned long long)__ballot_sync(0xffffffffu, ((my_badbits >> pass) & 1u) && pass * 32 + lane < K) << (32 * pass);
    if (lane == 0) {
      const double u = row_uniform(seed, sweep, (uint32_t)P.cls, key, 0, n_blocks, 0, PCLEAN_RNG_FINAL);
      int chosen;
      const int ngood = K - __popcll(badmask);
      const double w = badmask ? 1.0 / (double)ngood : exp(-log((double)K));
      if (use_mh) chosen = ((badmask >> 1) & 1u) ? 0 : ((u < fmin(1.0, w / (1e-10 + w))) ? 1 : 0);
      else { double cc = 0.0; chosen = -1; int last = 0; for (int k = 0; k < K; ++k) { if ((badmask >> k) & 1u) continue; last = k; cc += w; if (u < cc) { chosen = k; break; } } if (chosen < 0) chosen = last; }
      E.lsel[t] = chosen;
      E.llogml[t] = wsum;
      for (int si = 0; si < P.nroots; ++si) E.lchoice[(long long)si * TT.cap + t] = chosen == 0 ? PCL_CHOICE_UNSET : myChoice[si * PCL_MAX_K + chosen];
    }
    __syncwarp();
  }
}

// reference chain: slot of latent class reached from observation row r through `n_links` cells
struct RefChainD { int n_links; int block0; int col[4]; int table[4]; };
__global__ void k_ref_slots(const Dev* __restrict__ Ep, RefChainD ch, long long n, int* slot_of_row, int* counts) {
  const Dev& E = *Ep;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int s = E.assign[ch.block0][r];
  if (s < 0) { slot_of_row[r] = 0x7fffffff; return; }         // row not initialised yet: sorts last, refers to nothing
  for (int i = 0; i < ch.n_links; ++i) { const TableD& T = E.tables[ch.table[i]]; s = T.cells[(long long)ch.col[i] * T.cap + s]; }
  slot_of_row[r] = s;
  atomicAdd(&counts[s], 1);
}
__global__ void k_iota(int* p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}

// which cells of a latent row are observed (incorporate_observations!, dependency_tracking.jl:102-158):
// bit q of pat[slot] is set when some referring observation row observes column ocol[q] directly
struct ObsCellsD { int n; int data_col[8]; int block[8]; int bit[8]; };
__global__ void k_obs_pattern(const Dev* __restrict__ Ep, ObsCellsD oc, long long n, int* pat) {
  const Dev& E = *Ep;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int q = 0; q < oc.n; ++q) {
    const bool present = E.obs_real[oc.data_col[q]] ? (E.obs_real[oc.data_col[q]][r] == E.obs_real[oc.data_col[q]][r]) : E.obs_sid[oc.data_col[q]][r] >= 0;
    if (present && E.assign[oc.block[q]][r] >= 0) atomicOr(&pat[E.assign[oc.block[q]][r]], 1 << oc.bit[q]);
  }
}

// write the selected values of one site into the latent table (existing option / existing target row)
__global__ void k_lapply_site(const Dev* __restrict__ Ep, int prog_id, int site, int nslots, const int* pat_of_slot, int pat, int* req, int* changed) {
  const Dev& E = *Ep;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  req[t] = -1;
  if (pat_of_slot && pat_of_slot[t] != pat) return;
  const ProgD& P = E.progs[prog_id];
  const StarD& s = E.stars[P.star0 + P.roots[site]];
  TableD& T = E.tables[P.cls];
  if (T.refcnt[t] <= 0 || E.lsel[t] == 0) return;
  const int e = E.lchoice[(long long)site * T.cap + t];
  if (e == PCL_CHOICE_UNSET) return;
  if (s.kind == 1) {
    const int sid = s.list_func >= 0 ? e : E.optsid_pool[s.opt_off + e];
    if (T.cells[(long long)s.vertex * T.cap + t] != sid) { T.cells[(long long)s.vertex * T.cap + t] = sid; atomicAdd(changed, 1); }
  } else if (e >= 0) {
    const TableD& TG = E.tables[s.table];
    if (T.cells[(long long)s.vertex * T.cap + t] != e) atomicAdd(changed, 1);
    T.cells[(long long)s.vertex * T.cap + t] = e;
    const int2* cp = E.copies + s.copy0;
    for (int q = 0; q < s.ncopy; ++q) T.cells[(long long)cp[q].x * T.cap + t] = TG.cells[(long long)cp[q].y * TG.cap + e];
  } else { req[t] = -(e) - 2; atomicAdd(changed, 1); }
}
// after the proposed target rows were created: copy the (now complete) scratch record into the row
__global__ void k_lapply_new(const Dev* __restrict__ Ep, int prog_id, int site, int nslots, const int* req) {
  const Dev& E = *Ep;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots || req[t] < 0) return;
  const ProgD& P = E.progs[prog_id];
  const StarD& s = E.stars[P.star0 + P.roots[site]];
  TableD& T = E.tables[P.cls];
  const int* scratch = E.pool + (long long)req[t] * E.nvC;
  T.cells[(long long)s.vertex * T.cap + t] = scratch[s.vertex];
  const int2* cp = E.copies + s.copy0;
  for (int q = 0; q < s.ncopy; ++q) T.cells[(long long)cp[q].x * T.cap + t] = scratch[cp[q].x];
}
// denormalised copies of a referenced table's cells (update_referring_rows..., dependency_tracking.jl:239-258)
__global__ void k_refresh_copies(TableD* tables, int t, int g, const int2* copies, int ncopy) {
  TableD& T = tables[t];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T.n_slots) return;
  const TableD& TG = tables[T.fk_table[g]];
  const int tgt = T.cells[(long long)T.fk_col[g] * T.cap + j];
  if (tgt < 0) return;
  for (int q = 0; q < ncopy; ++q) T.cells[(long long)copies[q].x * T.cap + j] = TG.cells[(long long)copies[q].y * TG.cap + tgt];
}
// columns of a candidate matrix whose clean string changed since they were computed
__global__ void k_diff_cols(const int* cells, int* shadow, int n, int* flags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  flags[j] = (j < n && cells[j] != shadow[j]) ? 1 : 0;
}
__global__ void k_compact_cols(const int* cells, int* shadow, int n, const int* flags, const int* rank, int* list) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !flags[j]) return;
  list[rank[j]] = j;
  shadow[j] = cells[j];
}

}  // namespace pcl
