/*
 * pclean_oracle.cpp — CPU ORACLE: a restatement of probcomp/PClean's inference path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pclean_b200/) may include, link or
 * call this file; it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg as the checker and the timed CPU baseline ("port").
 *
 * PARITY UNPINNED: the reference ships no golden vectors or tests and cannot be
 * executed here (no Julia toolchain), so this restatement is pinned only by (a) known-answer
 * values derived from the reference's formulas (tests/test_oracle_kat.py) and
 * (b) end-to-end F1 on the three shipped datasets computed with the restated
 * evaluate_accuracy (analysis.jl:36-88).
 *
 * What it follows (all paths relative to /root/reference/src):
 *   inference/inference.jl:3-88            initialize_trace, pgibbs_sweep!, run_inference!
 *   inference/row_inference.jl:1-187       particles, ESS/resampling, run_smc!
 *   inference/block_proposal.jl:3-191      prune_plan, propose_non_enumerable!, make_block_proposal!
 *   inference/proposal_compiler.jl:5-422   semantics of the generated enumeration code
 *   inference/proposal_row_state.jl:2-66   overlay state for external likelihoods
 *   model/dependency_tracking.jl:1-258     incorporate/unincorporate, reference counting, GC
 *   model/trace.jl:53-107                  Pitman-Yor prior + hyper-parameter MH
 *   distributions/{...}.jl                  log-densities, proposals, sufficient statistics
 *   utils.jl:16-36                         logsumexp, logprobs
 * Third-party arithmetic restated from published closed forms (not vendored, unpinned):
 *   StringDistances.jl DamerauLevenshtein (OSA variant by default), Distributions.jl logpdf of
 *   NegativeBinomial / Normal / Gamma(1,1).
 *
 * Randomness follows include/pclean_rng.h ("same uniforms => same choices").
 */
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/pclean_b200.h"
#include "../include/pclean_rng.h"

namespace {

constexpr double NEG_INF = -std::numeric_limits<double>::infinity();
constexpr int TAG_NOTHING = 99;   // overlay marker: explicitly `nothing` (proposal_row_state.jl:63)

typedef pclean_value Val;
typedef std::vector<Val> Row;

inline Val mk(int tag, int i = 0, double d = 0.0) { Val v; v.tag = tag; v.i = i; v.d = d; return v; }
inline Val mk_str(int id) { return mk(PCLEAN_VAL_STR, id); }
inline Val mk_real(double d) { return mk(PCLEAN_VAL_REAL, 0, d); }
inline Val mk_key(int64_t k) { return mk(PCLEAN_VAL_KEY, 0, (double)k); }
inline int64_t key_of(const Val& v) { return (int64_t)v.d; }
inline bool present(const Val& v) { return v.tag != PCLEAN_VAL_ABSENT; }

// Julia `==` on the value kinds we carry
inline bool val_eq(const Val& a, const Val& b) {
  if (a.tag != b.tag) {
    // Int == Float compare numerically in Julia
    if ((a.tag == PCLEAN_VAL_INT && b.tag == PCLEAN_VAL_REAL)) return (double)a.i == b.d;
    if ((a.tag == PCLEAN_VAL_REAL && b.tag == PCLEAN_VAL_INT)) return a.d == (double)b.i;
    return false;
  }
  switch (a.tag) {
    case PCLEAN_VAL_REAL: return a.d == b.d;
    case PCLEAN_VAL_KEY: return a.d == b.d;
    case PCLEAN_VAL_MISSING: return false;   // missing == missing is `missing`, not true
    case PCLEAN_VAL_DUMMY: return true;
    default: return a.i == b.i;
  }
}

// utils.jl:16-25
double logsumexp(const std::vector<double>& x) {
  if (x.empty()) return NEG_INF;
  double m = *std::max_element(x.begin(), x.end());
  if (m == NEG_INF) return NEG_INF;
  double s = 0.0;
  for (double v : x) s += std::exp(v - m);
  return m + std::log(s);
}

struct OracleError : std::runtime_error { using std::runtime_error::runtime_error; };

// ------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------
struct Node {
  int kind = -1, wrap = 0, dist = -1, func = -1, target = -1, param = -1, path = -1, extv = -1;
  std::vector<int> wfk, wsub, args, vmap;
};
struct PlanNode { int v; std::vector<PlanNode> kids; };
typedef std::vector<PlanNode> Plan;

struct ClassM {
  int nv = 0, n_normal = 0;
  std::vector<Node> nodes;
  std::vector<std::vector<int>> blocks;
  std::vector<Plan> plans;
  std::vector<int> hash_keys;
  std::vector<int> paths;          // global ids of incoming paths
  double py_strength = 1.0, py_discount = 0.0;
};
struct PathM {
  int target;
  std::vector<std::pair<int, int>> links;   // (class, fk vertex); links.back() = ultimately referring class
  std::vector<int> vmap;                    // target-class vertex -> vertex in links.back().class
};
struct FuncM {
  int kind;
  Val cst;
  std::vector<int> keyargs;
  std::map<std::vector<int>, Val> table;
};

struct Model {
  std::vector<ClassM> classes;
  std::vector<PathM> paths;
  std::vector<FuncM> funcs;
  std::vector<int> param_kind, param_indexed, slot_param;
  std::vector<double> param_prior0, param_prior1;
  std::vector<std::vector<Val>> lists;
  std::vector<double> xform_scale;
  double lm_uni[28], lm_big[28 * 28];
};

static Plan parse_plan(const int32_t* pv, const int32_t* pn, int& pos, int nchild) {
  Plan out;
  for (int c = 0; c < nchild; ++c) {
    PlanNode n; n.v = pv[pos]; int k = pn[pos]; ++pos;
    n.kids = parse_plan(pv, pn, pos, k);
    out.push_back(std::move(n));
  }
  return out;
}

// ------------------------------------------------------------------------------------------
// parameters (choose_proportionally.jl:31-74, add_noise.jl:21-82, maybe_swap.jl:41-89)
// ------------------------------------------------------------------------------------------
struct ParamSlot {
  int spec = 0;
  std::vector<double> value;        // proportions: lazily sized; mean / prob: [v]
  std::vector<int64_t> counts;      // proportions: per-option counts; prob: {heads, tails}
  std::vector<int64_t> mcounts;     // mean: per std-group
  std::vector<double> msums, mstds;
  uint32_t epoch = 0;               // number of resample_value! calls so far
};

struct TableTrace {
  double strength = 1.0, discount = 0.0;
  std::map<int64_t, Row> rows;
  std::map<int64_t, Row> observations;
  std::map<int64_t, std::map<int, int>> observation_counts;
  std::map<std::vector<std::pair<int, int64_t>>, std::set<int64_t>> hashed_keys;
  // key -> (class, vertex) slot -> referring keys
  std::map<int64_t, std::map<std::pair<int, int>, std::set<int64_t>>> direct_incoming;
  std::map<int64_t, int64_t> reference_counts;
  int64_t total_references = 0;
  std::vector<std::pair<int, Val>> parameters;   // (vertex, PARAM/IPARAM value)
  uint32_t py_epoch = 0;
};

typedef std::map<int, std::vector<int64_t>> ReferringRows;   // path id -> sorted keys

struct Oracle;

struct RowState {
  int cls;
  Row row;
  int64_t key;
  const ReferringRows* referring = nullptr;
  const Row* retained = nullptr;
  // overlay (proposal_row_state.jl)
  const Row* active_parent = nullptr;
  Row recomputed;

  bool has(int i) const {
    if (!active_parent) return present(row[i]);
    const Val& r = recomputed[i];
    if (r.tag == PCLEAN_VAL_ABSENT) return present((*active_parent)[i]);
    return r.tag != TAG_NOTHING;
  }
  const Val& get(int i) const {
    if (active_parent) {
      const Val& r = recomputed[i];
      if (r.tag != PCLEAN_VAL_ABSENT) return r;
      return (*active_parent)[i];
    }
    return row[i];
  }
  void set(int i, const Val& v) {
    if (active_parent) { recomputed[i] = v; return; }
    row[i] = v;
  }
};

struct Particle { RowState state; double weight = 0.0; int block_index = 0; };

struct MoveRecord {          // what pclean_row_move_debug reports
  std::vector<int64_t> chosen_keys;   // [K][n_blocks]; -1 = new row, -2 = block has no FK root
  std::vector<double> weights;        // final weights [K]
  int selected = 0;
  double log_ml = 0.0;
};

struct Res { double p = 0.0, q = 0.0; std::vector<std::pair<int, Val>> t; };

struct Oracle {
  Model m;
  pclean_config cfg;
  uint64_t seed = 0;
  std::vector<std::u32string> strings;
  std::unordered_map<std::u32string, int> string_ids;
  std::vector<ParamSlot> slots;
  std::vector<TableTrace> tables;
  int64_t gensym = 0;                    // gensym_counter.jl (keys of latent rows)
  bool true_damerau = false;
  std::unordered_map<uint64_t, double> typo_memo;       // add_typos.jl:47 (key ignores max_typos)
  std::unordered_map<uint64_t, double> sprior_memo;     // string_prior.jl:42
  std::string last_error;
  // observed datasets
  struct Obs { int cls; int64_t n; std::vector<int> vertex_of_col; std::vector<Val> cells; };
  std::vector<Obs> datasets;
  // counters for the report
  int64_t n_dp_cells = 0, n_typo_evals = 0, n_typo_misses = 0;
  uint32_t cur_sweep = 0;
  MoveRecord* record = nullptr;

  // ---------------------------------------------------------------- strings
  int intern(const std::u32string& s) {
    auto it = string_ids.find(s);
    if (it != string_ids.end()) return it->second;
    int id = (int)strings.size();
    strings.push_back(s);
    string_ids.emplace(s, id);
    return id;
  }
  int intern_ascii(const std::string& s) { return intern(std::u32string(s.begin(), s.end())); }

  // ---------------------------------------------------------------- AddTypos (add_typos.jl:50-66)
  int edit_distance(const std::u32string& a, const std::u32string& b) {
    const int n = (int)a.size(), mlen = (int)b.size();
    n_dp_cells += (int64_t)n * mlen;
    if (n == 0) return mlen;
    if (mlen == 0) return n;
    if (!true_damerau) {
      // optimal string alignment (restricted Damerau-Levenshtein): StringDistances <= 0.10
      std::vector<int> pp(mlen + 1), p(mlen + 1), c(mlen + 1);
      for (int j = 0; j <= mlen; ++j) p[j] = j;
      for (int i = 1; i <= n; ++i) {
        c[0] = i;
        for (int j = 1; j <= mlen; ++j) {
          int cost = a[i - 1] == b[j - 1] ? 0 : 1;
          int v = std::min(std::min(p[j] + 1, c[j - 1] + 1), p[j - 1] + cost);
          if (i > 1 && j > 1 && a[i - 1] == b[j - 2] && a[i - 2] == b[j - 1]) v = std::min(v, pp[j - 2] + 1);
          c[j] = v;
        }
        std::swap(pp, p); std::swap(p, c);
      }
      return p[mlen];
    }
    // unrestricted Damerau-Levenshtein (Lowrance-Wagner)
    std::map<char32_t, int> da;
    const int maxd = n + mlen;
    std::vector<std::vector<int>> d(n + 2, std::vector<int>(mlen + 2, 0));
    d[0][0] = maxd;
    for (int i = 0; i <= n; ++i) { d[i + 1][0] = maxd; d[i + 1][1] = i; }
    for (int j = 0; j <= mlen; ++j) { d[0][j + 1] = maxd; d[1][j + 1] = j; }
    for (int i = 1; i <= n; ++i) {
      int db = 0;
      for (int j = 1; j <= mlen; ++j) {
        int k = da.count(b[j - 1]) ? da[b[j - 1]] : 0, l = db, cost = 1;
        if (a[i - 1] == b[j - 1]) { cost = 0; db = j; }
        d[i + 1][j + 1] = std::min(std::min(d[i][j] + cost, d[i + 1][j] + 1),
                                   std::min(d[i][j + 1] + 1, d[k][l] + (i - k - 1) + 1 + (j - l - 1)));
      }
      da[a[i - 1]] = i;
    }
    return d[n + 1][mlen + 1];
  }

  static double addtypos_score(int k, int word_len) {
    // logpdf(NegativeBinomial(ceil(len/5), 0.9), k) - k log(len) - k log(26)/2
    double r = std::ceil(word_len / 5.0);
    double l = std::lgamma(k + r) - std::lgamma(k + 1.0) - std::lgamma(r) + r * std::log(0.9) + k * std::log(0.1);
    l -= std::log((double)word_len) * k;
    l -= std::log(26.0) * k / 2.0;
    return l;
  }

  double addtypos_logdensity(const Val& observed, const Val& word, int max_typos) {
    if (observed.tag == PCLEAN_VAL_MISSING) return 0.0;
    if (observed.tag != PCLEAN_VAL_STR || word.tag != PCLEAN_VAL_STR) throw OracleError("AddTypos: non-string argument");
    ++n_typo_evals;
    uint64_t mk_ = ((uint64_t)(uint32_t)observed.i << 32) | (uint32_t)word.i;
    auto it = typo_memo.find(mk_);
    if (it != typo_memo.end()) return it->second;
    ++n_typo_misses;
    const std::u32string& o = strings[observed.i];
    const std::u32string& w = strings[word.i];
    int k = edit_distance(o, w);
    double l;
    if (max_typos >= 0 && k > max_typos) l = -1e5;            // IMPOSSIBLE, add_typos.jl:34
    else l = addtypos_score(k, (int)w.size());
    typo_memo.emplace(mk_, l);
    return l;
  }

  // ---------------------------------------------------------------- StringPrior (string_prior.jl:43-61)
  static int alphabet_index(char32_t c) {
    if (c >= U'A' && c <= U'Z') c = c - U'A' + U'a';
    if (c >= U'a' && c <= U'z') return (int)(c - U'a');
    if (c == U' ') return 26;
    if (c == U'.') return 27;
    return -1;
  }
  double stringprior_logdensity(int sid, int minl, int maxl) {
    uint64_t key = ((uint64_t)(uint32_t)sid << 24) ^ ((uint64_t)minl << 12) ^ (uint64_t)maxl;
    auto it = sprior_memo.find(key);
    if (it != sprior_memo.end()) return it->second;
    const std::u32string& s = strings[sid];
    double score;
    int len = (int)s.size();
    if (len < minl || len > maxl) score = NEG_INF;
    else {
      score = -std::log((double)(maxl - minl + 1));
      int prev = -1;
      for (char32_t ch : s) {
        int cur = alphabet_index(ch);
        if (cur < 0) score += -std::log(28.0);
        else {
          double pr = prev < 0 ? m.lm_uni[cur] : m.lm_big[cur * 28 + prev];
          score += std::max(std::log(pr), -1000.0);
        }
        prev = cur;
      }
    }
    sprior_memo.emplace(key, score);
    return score;
  }
  static bool time_regex(const std::u32string& s) {    // ^\d?\d:\d\d [ap]\.m\.$
    size_t n = s.size();
    auto dig = [&](size_t i) { return i < n && s[i] >= U'0' && s[i] <= U'9'; };
    size_t p = 0;
    if (!dig(p)) return false;
    ++p;
    if (dig(p)) ++p;
    if (p >= n || s[p] != U':') return false;
    ++p;
    if (!dig(p) || !dig(p + 1)) return false;
    p += 2;
    if (p + 5 != n) return false;
    return s[p] == U' ' && (s[p + 1] == U'a' || s[p + 1] == U'p') && s[p + 2] == U'.' && s[p + 3] == U'm' && s[p + 4] == U'.';
  }

  // ---------------------------------------------------------------- parameters
  pclean_stream param_stream(int slot, uint32_t epoch, int purpose) const {
    pclean_stream s; s.key.seed = seed; s.key.sweep = epoch; s.key.cls = 0; s.key.row = slot;
    s.key.particle = 0; s.key.block = 0; s.key.site = 0; s.key.purpose = (uint32_t)purpose; s.idx = 0;
    return s;
  }
  void init_slot(int slot) {
    ParamSlot& p = slots[slot];
    int spec = p.spec;
    pclean_stream s = param_stream(slot, 0, PCLEAN_RNG_PARAM_INIT);
    switch (m.param_kind[spec]) {
      case PCLEAN_PARAM_PROPORTIONS: break;   // sized on first param_value (choose_proportionally.jl:48-55)
      case PCLEAN_PARAM_MEAN:
        p.value = {m.param_prior0[spec] + m.param_prior1[spec] * pclean_next_normal(&s)};   // add_noise.jl:44
        break;
      case PCLEAN_PARAM_PROB:
        p.value = {pclean_next_beta(&s, m.param_prior0[spec], m.param_prior1[spec])};       // maybe_swap.jl:58
        p.counts = {0, 0};
        break;
    }
  }
  const std::vector<double>& proportions_value(int slot, size_t n_options) {
    ParamSlot& p = slots[slot];
    if (p.value.empty()) {
      p.counts.assign(n_options, 0);
      p.value.resize(n_options);
      pclean_stream s = param_stream(slot, 0, PCLEAN_RNG_PARAM_INIT);
      double conc = m.param_prior0[p.spec], tot = 0.0;
      for (size_t i = 0; i < n_options; ++i) { p.value[i] = pclean_next_gamma(&s, conc); tot += p.value[i]; }
      for (double& v : p.value) v /= tot;
    }
    return p.value;
  }
  void resample_slot(int slot) {
    ParamSlot& p = slots[slot];
    int spec = p.spec;
    ++p.epoch;
    pclean_stream s = param_stream(slot, p.epoch, PCLEAN_RNG_PARAM);
    switch (m.param_kind[spec]) {
      case PCLEAN_PARAM_PROPORTIONS: {       // choose_proportionally.jl:70-74
        if (p.value.empty()) return;
        double conc = m.param_prior0[spec], tot = 0.0;
        for (size_t i = 0; i < p.value.size(); ++i) { p.value[i] = pclean_next_gamma(&s, conc + (double)p.counts[i]); tot += p.value[i]; }
        for (double& v : p.value) v /= tot;
        break;
      }
      case PCLEAN_PARAM_MEAN: {              // add_noise.jl:74-82
        double mean = m.param_prior0[spec], var = m.param_prior1[spec] * m.param_prior1[spec];
        for (size_t g = 0; g < p.mcounts.size(); ++g) {
          double sd = p.mstds[g];
          double new_var = 1.0 / (1.0 / var + (double)p.mcounts[g] / (sd * sd));
          mean = new_var * (mean / var + p.msums[g] / (sd * sd));
          var = new_var;
        }
        p.value[0] = mean + std::sqrt(var) * pclean_next_normal(&s);
        break;
      }
      case PCLEAN_PARAM_PROB:                // maybe_swap.jl:87-89
        p.value[0] = pclean_next_beta(&s, m.param_prior0[spec] + (double)p.counts[0], m.param_prior1[spec] + (double)p.counts[1]);
        break;
    }
  }
  void resample_parameters_of_class(int cls) {
    for (auto& pr : tables[cls].parameters) {
      const Val& v = pr.second;
      if (v.tag == PCLEAN_VAL_PARAM) resample_slot(v.i);
      else if (v.tag == PCLEAN_VAL_IPARAM) {       // distributions.jl:57-61
        for (size_t s = 0; s < slots.size(); ++s) if (slots[s].spec == v.i) resample_slot((int)s);
      }
    }
  }
  double real_of(const Val& v) {
    if (v.tag == PCLEAN_VAL_REAL) return v.d;
    if (v.tag == PCLEAN_VAL_INT) return (double)v.i;
    if (v.tag == PCLEAN_VAL_PARAM) return slots[v.i].value.at(0);
    throw OracleError("expected a real-valued argument");
  }
  static bool isapprox(double a, double b) {   // Base.isapprox default rtol = sqrt(eps)
    return a == b || std::fabs(a - b) <= 1.4901161193847656e-8 * std::max(std::fabs(a), std::fabs(b));
  }

  // ---------------------------------------------------------------- distribution protocol
  const std::vector<Val>& list_of(const Val& v) {
    if (v.tag != PCLEAN_VAL_LIST) throw OracleError("expected an option list");
    return m.lists[v.i];
  }
  int int_of(const Val& v) {
    if (v.tag == PCLEAN_VAL_INT) return v.i;
    if (v.tag == PCLEAN_VAL_REAL) return (int)v.d;
    throw OracleError("expected an integer argument");
  }
  std::vector<double> proportions_logprobs(const Val& opts, const Val& probs) {
    const std::vector<Val>& o = list_of(opts);
    std::vector<double> lp(o.size());
    if (probs.tag == PCLEAN_VAL_PARAM) {
      const std::vector<double>& v = proportions_value(probs.i, o.size());
      for (size_t i = 0; i < o.size(); ++i) lp[i] = std::log(v[i]);     // utils.jl:33-36 (no normalisation)
    } else throw OracleError("ChooseProportionally: literal probability vectors are not lowered");
    return lp;
  }

  double logdensity(int dist, const Val& obs, const std::vector<Val>& a) {
    switch (dist) {
      case PCLEAN_DIST_ADD_TYPOS:
        return addtypos_logdensity(obs, a.at(0), a.size() > 1 ? int_of(a[1]) : -1);
      case PCLEAN_DIST_CHOOSE_PROPORTIONALLY: {       // choose_proportionally.jl:7-11
        const std::vector<Val>& o = list_of(a.at(0));
        std::vector<double> lp = proportions_logprobs(a[0], a.at(1)), rel;
        for (size_t i = 0; i < o.size(); ++i) if (val_eq(o[i], obs)) rel.push_back(lp[i]);
        if (rel.empty()) return NEG_INF;
        return logsumexp(rel);
      }
      case PCLEAN_DIST_CHOOSE_UNIFORMLY:              // choose_uniformly.jl:7-10
        return -std::log((double)list_of(a.at(0)).size());
      case PCLEAN_DIST_STRING_PRIOR:
        if (obs.tag != PCLEAN_VAL_STR) throw OracleError("StringPrior: non-string value");
        return stringprior_logdensity(obs.i, int_of(a.at(0)), int_of(a.at(1)));
      case PCLEAN_DIST_TIME_PRIOR: return -std::log(1440.0);      // time_prior.jl:24-26
      case PCLEAN_DIST_MAYBE_SWAP: {                  // maybe_swap.jl:13-28
        const Val& val = a.at(0);
        const std::vector<Val>& o = list_of(a.at(1));
        double prob = real_of(a.at(2));
        if (obs.tag == PCLEAN_VAL_MISSING) {
          for (const Val& x : o) if (val_eq(x, val)) return 0.0;
          return -1000.0;
        }
        if (val_eq(val, obs)) return std::log1p(-prob);
        return std::log(prob) - std::log((double)o.size());
      }
      case PCLEAN_DIST_TRANSFORMED_GAUSSIAN: {        // transformed_gaussian.jl:15-16
        double mean = real_of(a.at(0)), sd = real_of(a.at(1));
        if (a.at(2).tag != PCLEAN_VAL_XFORM) throw OracleError("TransformedGaussian: bad transformation");
        double sc = m.xform_scale[a[2].i];
        double x = real_of(obs) * sc;                  // backward
        double z = (x - mean) / sd;
        return -0.5 * z * z - std::log(sd) - 0.91893853320467274178 - std::log(std::fabs(1.0 / sc));
      }
      case PCLEAN_DIST_ADD_NOISE: {                   // add_noise.jl:7
        double mean = real_of(a.at(0)), sd = real_of(a.at(1));
        double z = (real_of(obs) - mean) / sd;
        return -0.5 * z * z - std::log(sd) - 0.91893853320467274178;
      }
      case PCLEAN_DIST_UNMODELED: return 0.0;         // unmodeled.jl:7-10
    }
    throw OracleError("unknown distribution");
  }

  static bool has_discrete_proposal(int dist) {
    return dist == PCLEAN_DIST_CHOOSE_PROPORTIONALLY || dist == PCLEAN_DIST_CHOOSE_UNIFORMLY ||
           dist == PCLEAN_DIST_STRING_PRIOR || dist == PCLEAN_DIST_TIME_PRIOR;
  }

  // discrete_proposal: options (DUMMY as last entry for open-support priors) and log-probs
  void discrete_proposal(int dist, const std::vector<Val>& a, std::vector<Val>& options, std::vector<double>& lp) {
    options.clear(); lp.clear();
    switch (dist) {
      case PCLEAN_DIST_CHOOSE_PROPORTIONALLY:
        options = list_of(a.at(0)); lp = proportions_logprobs(a[0], a.at(1)); return;
      case PCLEAN_DIST_CHOOSE_UNIFORMLY: {
        options = list_of(a.at(0)); lp.assign(options.size(), -std::log((double)options.size())); return;
      }
      case PCLEAN_DIST_STRING_PRIOR: {               // string_prior.jl:16-22
        options = list_of(a.at(2));
        int mn = int_of(a.at(0)), mx = int_of(a.at(1));
        for (const Val& o : options) lp.push_back(stringprior_logdensity(o.i, mn, mx));
        double total = logsumexp(lp);
        options.push_back(mk(PCLEAN_VAL_DUMMY));
        lp.push_back(std::log1p(-std::exp(total)));
        return;
      }
      case PCLEAN_DIST_TIME_PRIOR: {                 // time_prior.jl:8-14
        options = list_of(a.at(0));
        for (const Val& o : options) lp.push_back(time_regex(strings[o.i]) ? -std::log(1440.0) : NEG_INF);
        double total = logsumexp(lp);
        options.push_back(mk(PCLEAN_VAL_DUMMY));
        lp.push_back(std::log1p(-std::exp(total)));
        return;
      }
    }
    throw OracleError("distribution has no discrete proposal");
  }
  Val dummy_value(int dist, const std::vector<Val>& a) {
    if (dist == PCLEAN_DIST_STRING_PRIOR) {           // string_prior.jl:24-26
      int n = (int_of(a.at(0)) + int_of(a.at(1))) / 2;
      return mk_str(intern(std::u32string((size_t)n, U'*')));
    }
    if (dist == PCLEAN_DIST_TIME_PRIOR) return mk_str(intern_ascii("**:** p.m."));   // time_prior.jl:16-18
    throw OracleError("no dummy value");
  }
  static int categorical(const std::vector<double>& probs, double u) {   // inverse CDF
    double c = 0.0; int last = -1;
    for (size_t i = 0; i < probs.size(); ++i) {
      if (probs[i] > 0.0) last = (int)i;
      c += probs[i];
      if (u < c) return (int)i;
    }
    if (last < 0) throw OracleError("Categorical: all-zero probability vector");
    return last;
  }
  Val random_value(int dist, const std::vector<Val>& a, pclean_stream& s) {
    switch (dist) {
      case PCLEAN_DIST_CHOOSE_UNIFORMLY: {
        const std::vector<Val>& o = list_of(a.at(0));
        return o[std::min(o.size() - 1, (size_t)(pclean_next(&s) * o.size()))];
      }
      case PCLEAN_DIST_CHOOSE_PROPORTIONALLY: {
        const std::vector<Val>& o = list_of(a.at(0));
        std::vector<double> lp = proportions_logprobs(a[0], a.at(1)), pr(lp.size());
        double tot = 0.0;
        for (size_t i = 0; i < lp.size(); ++i) { pr[i] = std::exp(lp[i]); tot += pr[i]; }
        for (double& v : pr) v /= tot;
        return o[categorical(pr, pclean_next(&s))];
      }
      case PCLEAN_DIST_STRING_PRIOR: {               // string_prior.jl:28-39
        int mn = int_of(a.at(0)), mx = int_of(a.at(1));
        int len = mn + std::min(mx - mn, (int)(pclean_next(&s) * (mx - mn + 1)));
        std::u32string out;
        int prev = -1;
        static const char32_t alpha[] = U"abcdefghijklmnopqrstuvwxyz .";
        for (int i = 0; i < len; ++i) {
          std::vector<double> pr(28);
          double tot = 0.0;
          for (int c = 0; c < 28; ++c) { pr[c] = prev < 0 ? m.lm_uni[c] : m.lm_big[c * 28 + prev]; tot += pr[c]; }
          for (double& v : pr) v /= tot;
          prev = categorical(pr, pclean_next(&s));
          out.push_back(alpha[prev]);
        }
        return mk_str(intern(out));
      }
      case PCLEAN_DIST_TIME_PRIOR: {                 // time_prior.jl:20-22
        int h = 1 + std::min(11, (int)(pclean_next(&s) * 12)), mi = 1 + std::min(59, (int)(pclean_next(&s) * 60));
        bool am = pclean_next(&s) < 0.5;
        return mk_str(intern_ascii(std::to_string(h) + ":" + std::to_string(mi) + (am ? " a.m." : " p.m.")));
      }
      case PCLEAN_DIST_TRANSFORMED_GAUSSIAN: {
        double mean = real_of(a.at(0)), sd = real_of(a.at(1)), sc = m.xform_scale[a.at(2).i];
        return mk_real((mean + sd * pclean_next_normal(&s)) / sc);
      }
      case PCLEAN_DIST_ADD_NOISE:
        return mk_real(real_of(a.at(0)) + real_of(a.at(1)) * pclean_next_normal(&s));
      case PCLEAN_DIST_MAYBE_SWAP: {
        const std::vector<Val>& o = list_of(a.at(1));
        if (pclean_next(&s) < real_of(a.at(2))) return o[std::min(o.size() - 1, (size_t)(pclean_next(&s) * o.size()))];
        return a.at(0);
      }
      case PCLEAN_DIST_ADD_TYPOS: {                  // add_typos.jl:9-45
        std::u32string w = strings[a.at(0).i];
        double r = std::ceil(w.size() / 5.0);
        // NegativeBinomial(r, 0.9) by inversion
        double u = pclean_next(&s), c = 0.0; int k = 0;
        for (; k < 1000; ++k) {
          c += std::exp(std::lgamma(k + r) - std::lgamma(k + 1.0) - std::lgamma(r) + r * std::log(0.9) + k * std::log(0.1));
          if (u < c) break;
        }
        if (a.size() > 1) k = std::min(k, int_of(a[1]));
        for (int t = 0; t < k; ++t) {
          int typo = std::min(3, (int)(pclean_next(&s) * 4));
          char32_t letter = U'a' + std::min(25, (int)(pclean_next(&s) * 26));
          int L = (int)w.size();
          if (typo == 0) { int idx = std::min(L, (int)(pclean_next(&s) * (L + 1))); w.insert(w.begin() + idx, letter); }
          else if (typo == 1 && L > 0) { int idx = std::min(L - 1, (int)(pclean_next(&s) * L)); w.erase(w.begin() + idx); }
          else if (typo == 2 && L > 1) { int idx = std::min(L - 2, (int)(pclean_next(&s) * (L - 1))); std::swap(w[idx], w[idx + 1]); }
          else if (typo == 3 && L > 0) { int idx = std::min(L - 1, (int)(pclean_next(&s) * L)); w[idx] = letter; }
        }
        return mk_str(intern(w));
      }
    }
    throw OracleError("random(): unsupported distribution (Unmodeled values must be observed)");
  }

  // incorporate_choice! / unincorporate_choice!
  void incorporate_choice(int dist, const Val& obs, const std::vector<Val>& a, int sign) {
    switch (dist) {
      case PCLEAN_DIST_CHOOSE_PROPORTIONALLY: {       // choose_proportionally.jl:57-68
        if (a.at(1).tag != PCLEAN_VAL_PARAM) return;
        const std::vector<Val>& o = list_of(a.at(0));
        proportions_value(a[1].i, o.size());
        for (size_t i = 0; i < o.size(); ++i) if (val_eq(o[i], obs)) { slots[a[1].i].counts[i] += sign; return; }
        throw OracleError("ChooseProportionally: incorporated value is not an option");
      }
      case PCLEAN_DIST_TRANSFORMED_GAUSSIAN:
      case PCLEAN_DIST_ADD_NOISE: {                   // add_noise.jl:48-71, transformed_gaussian.jl:26-33
        if (a.at(0).tag != PCLEAN_VAL_PARAM) return;
        double x = real_of(obs);
        if (dist == PCLEAN_DIST_TRANSFORMED_GAUSSIAN) x *= m.xform_scale[a.at(2).i];
        double sd = real_of(a.at(1));
        ParamSlot& p = slots[a[0].i];
        int g = -1;
        for (size_t i = 0; i < p.mstds.size(); ++i) if (isapprox(p.mstds[i], sd)) { g = (int)i; break; }
        if (sign > 0) {
          if (g < 0) { p.mstds.push_back(sd); p.msums.push_back(x); p.mcounts.push_back(1); return; }
          p.mcounts[g] += 1; p.msums[g] += x;
        } else {
          if (g < 0) throw OracleError("MeanParameter: unincorporate of unseen std group");
          p.mcounts[g] -= 1; p.msums[g] -= x;
          if (p.mcounts[g] == 0) { p.mcounts.erase(p.mcounts.begin() + g); p.msums.erase(p.msums.begin() + g); p.mstds.erase(p.mstds.begin() + g); }
        }
        return;
      }
      case PCLEAN_DIST_MAYBE_SWAP: {                  // maybe_swap.jl:65-85
        if (a.at(2).tag != PCLEAN_VAL_PARAM) return;
        if (obs.tag == PCLEAN_VAL_MISSING) return;
        ParamSlot& p = slots[a[2].i];
        if (val_eq(obs, a.at(0))) p.counts[1] += sign; else p.counts[0] += sign;
        return;
      }
      default: return;
    }
  }

  // ---------------------------------------------------------------- JuliaNode evaluation
  Val eval_func(int f, const std::vector<Val>& a) {
    const FuncM& fn = m.funcs[f];
    switch (fn.kind) {
      case PCLEAN_FUNC_CONST: return fn.cst;
      case PCLEAN_FUNC_TABLE: {
        std::vector<int> key;
        for (int pos : fn.keyargs) {
          const Val& v = a.at(pos);
          if (v.tag == PCLEAN_VAL_REAL || v.tag == PCLEAN_VAL_MISSING || v.tag == PCLEAN_VAL_DUMMY || v.tag == PCLEAN_VAL_ABSENT)
            throw OracleError("tabulated JuliaNode called with a non-discrete argument");
          key.push_back(v.i);
        }
        auto it = fn.table.find(key);
        if (it == fn.table.end()) throw OracleError("tabulated JuliaNode: argument outside its tabulated support");
        return it->second;
      }
      case PCLEAN_FUNC_ROUND_BACKWARD: {
        double x = real_of(a.at(1)) * m.xform_scale[a.at(0).i];
        return mk_real(std::nearbyint(x));            // Julia round = ties-to-even
      }
      case PCLEAN_FUNC_JOIN: {
        if (a.at(0).tag != PCLEAN_VAL_STR || a.at(1).tag != PCLEAN_VAL_STR) throw OracleError("join: non-string argument");
        std::u32string s = strings[a[0].i];
        s += strings[fn.cst.i];
        s += strings[a[1].i];
        return mk_str(intern(s));
      }
    }
    throw OracleError("unknown function kind");
  }

  // ---------------------------------------------------------------- dependency tracking
  void update_sufficient_statistics(int cls, Row& row, int sign, bool reevaluate_jns = false) {
    const ClassM& cm = m.classes[cls];                // dependency_tracking.jl:6-21
    std::vector<Val> args;
    for (int i = 0; i < cm.nv; ++i) {
      const Node& n = cm.nodes[i];
      if (n.wrap != PCLEAN_WRAP_NONE) continue;
      if (reevaluate_jns && n.kind == PCLEAN_NODE_JULIA) {
        args.clear();
        for (int a : n.args) args.push_back(row[a]);
        row[i] = eval_func(n.func, args);
      }
      if (n.kind == PCLEAN_NODE_CHOICE) {
        args.clear();
        for (int a : n.args) args.push_back(row[a]);
        incorporate_choice(n.dist, row[i], args, sign);
      }
    }
  }
  std::vector<std::pair<int, int64_t>> hash_key_of(int cls, const Row& row) {
    std::vector<std::pair<int, int64_t>> k;
    for (int h : m.classes[cls].hash_keys) {
      const Val& v = row[h];
      k.emplace_back(v.tag, v.tag == PCLEAN_VAL_REAL || v.tag == PCLEAN_VAL_KEY ? (int64_t)v.d : (int64_t)v.i);
    }
    return k;
  }
  void unincorporate_observations(int cls, int64_t key, const std::vector<int>& to_delete) {
    TableTrace& t = tables[cls];                      // dependency_tracking.jl:102-129
    const ClassM& cm = m.classes[cls];
    std::set<int> gone;
    for (int v : to_delete) {
      int& c = t.observation_counts[key][v];
      c -= 1;
      if (c == 0) { gone.insert(v); t.observations[key][v] = mk(PCLEAN_VAL_ABSENT); }
    }
    Row& row = t.rows.at(key);
    for (int i = 0; i < cm.nv; ++i) {
      const Node& n = cm.nodes[i];
      if (n.wrap != PCLEAN_WRAP_NONE || n.kind != PCLEAN_NODE_FK) continue;
      std::vector<int> sub;
      for (size_t tv = 0; tv < n.vmap.size(); ++tv) if (gone.count(n.vmap[tv])) sub.push_back((int)tv);
      unincorporate_observations(n.target, key_of(row[i]), sub);
    }
  }
  void incorporate_observations(int cls, int64_t key, const std::vector<std::pair<int, Val>>& obs) {
    TableTrace& t = tables[cls];                      // dependency_tracking.jl:132-158
    const ClassM& cm = m.classes[cls];
    Row& existing = t.observations.at(key);
    std::set<int> fresh;
    for (auto& pr : obs) {
      if (present(existing[pr.first])) t.observation_counts[key][pr.first] += 1;
      else { existing[pr.first] = pr.second; fresh.insert(pr.first); t.observation_counts[key][pr.first] = 1; }
    }
    Row& row = t.rows.at(key);
    for (int i = 0; i < cm.nv; ++i) {
      const Node& n = cm.nodes[i];
      if (n.wrap != PCLEAN_WRAP_NONE || n.kind != PCLEAN_NODE_FK) continue;
      std::vector<std::pair<int, Val>> sub;
      for (size_t tv = 0; tv < n.vmap.size(); ++tv)
        if (fresh.count(n.vmap[tv])) {
          for (auto& pr : obs) if (pr.first == n.vmap[tv]) { sub.emplace_back((int)tv, pr.second); break; }
        }
      incorporate_observations(n.target, key_of(row[i]), sub);
    }
  }
  void unrefer_to_row(int tcls, int64_t tkey, std::pair<int, int> slot, int64_t referring_key, const std::vector<int>& obs_to_delete) {
    TableTrace& t = tables[tcls];                     // dependency_tracking.jl:162-202
    auto& inc = t.direct_incoming.at(tkey);
    inc[slot].erase(referring_key);
    if (inc[slot].empty()) inc.erase(slot);
    unincorporate_observations(tcls, tkey, obs_to_delete);
    t.total_references -= 1;
    if (t.reference_counts.at(tkey) > 1) { t.reference_counts[tkey] -= 1; return; }
    unincorporate_row(tcls, tkey);
    update_sufficient_statistics(tcls, t.rows.at(tkey), -1);
    t.reference_counts.erase(tkey);
    t.rows.erase(tkey);
    t.observations.erase(tkey);
    t.observation_counts.erase(tkey);
    t.direct_incoming.erase(tkey);
  }
  void unincorporate_row(int cls, int64_t key) {
    TableTrace& t = tables[cls];                      // dependency_tracking.jl:26-66
    const ClassM& cm = m.classes[cls];
    const Row row = t.rows.at(key);
    const Row obs = t.observations.at(key);
    if (!cm.hash_keys.empty()) {
      auto hk = hash_key_of(cls, row);
      auto it = t.hashed_keys.find(hk);
      if (it == t.hashed_keys.end()) throw OracleError("hash bucket missing on unincorporate");
      it->second.erase(key);
      if (it->second.empty()) t.hashed_keys.erase(it);
    }
    for (int i = 0; i < cm.nv; ++i) {
      const Node& n = cm.nodes[i];
      if (n.wrap != PCLEAN_WRAP_NONE || n.kind != PCLEAN_NODE_FK) continue;
      std::vector<int> del;
      for (size_t tv = 0; tv < n.vmap.size(); ++tv) if (present(obs[n.vmap[tv]])) del.push_back((int)tv);
      unrefer_to_row(n.target, key_of(row[i]), {cls, i}, key, del);
    }
  }
  void refer_to_row(int tcls, int64_t tkey, std::pair<int, int> slot, int64_t referring_key, Row&& row_trace,
                    const std::vector<std::pair<int, Val>>& obs) {
    TableTrace& t = tables[tcls];                     // dependency_tracking.jl:205-236
    if (!t.rows.count(tkey)) {
      t.rows[tkey] = std::move(row_trace);
      t.reference_counts[tkey] = 0;
      t.observations[tkey] = Row(m.classes[tcls].nv, mk(PCLEAN_VAL_ABSENT));
      t.observation_counts[tkey];
      t.direct_incoming[tkey][slot];
      incorporate_row(tcls, tkey);
      update_sufficient_statistics(tcls, t.rows.at(tkey), +1);
    }
    t.reference_counts[tkey] += 1;
    t.total_references += 1;
    t.direct_incoming[tkey][slot].insert(referring_key);
    incorporate_observations(tcls, tkey, obs);
  }
  void incorporate_row(int cls, int64_t key) {
    TableTrace& t = tables[cls];                      // dependency_tracking.jl:71-99
    const ClassM& cm = m.classes[cls];
    const Row row = t.rows.at(key);
    const Row obs = t.observations.at(key);
    if (!cm.hash_keys.empty()) t.hashed_keys[hash_key_of(cls, row)].insert(key);
    for (int i = 0; i < cm.nv; ++i) {
      const Node& n = cm.nodes[i];
      if (n.wrap != PCLEAN_WRAP_NONE || n.kind != PCLEAN_NODE_FK) continue;
      const int tnv = m.classes[n.target].nv;
      Row trow(tnv, mk(PCLEAN_VAL_ABSENT));
      std::vector<std::pair<int, Val>> tobs;
      for (size_t tv = 0; tv < n.vmap.size(); ++tv) {
        trow[tv] = row[n.vmap[tv]];
        if (present(obs[n.vmap[tv]])) tobs.emplace_back((int)tv, obs[n.vmap[tv]]);
      }
      refer_to_row(n.target, key_of(row[i]), {cls, i}, key, std::move(trow), tobs);
    }
  }
  void update_referring_rows(int cls, const Row& new_values, const ReferringRows& referring) {
    const ClassM& cm = m.classes[cls];                // dependency_tracking.jl:239-258
    for (int pid : cm.paths) {
      const PathM& path = m.paths[pid];
      int rcls = path.links.back().first;
      TableTrace& rt = tables[rcls];
      auto it = referring.find(pid);
      if (it == referring.end()) continue;
      for (int64_t rkey : it->second) {
        Row& rrow = rt.rows.at(rkey);
        update_sufficient_statistics(rcls, rrow, -1);
        for (size_t tv = 0; tv < path.vmap.size(); ++tv) if (path.vmap[tv] >= 0) rrow[path.vmap[tv]] = new_values[tv];
        update_sufficient_statistics(rcls, rrow, +1, true);
      }
    }
  }

  // ---------------------------------------------------------------- Pitman-Yor (trace.jl:53-107)
  static double pitman_yor_score(double strength, double discount, const std::vector<int64_t>& counts) {
    double lp = 0.0; int64_t nref = 0; int64_t nobj = 0;
    for (int64_t size : counts) {
      ++nobj;
      lp += std::log(nobj * discount + strength) - std::log(nref + strength);
      for (int64_t i = 1; i <= size - 1; ++i) lp += std::log(i - discount) - std::log(nref + i + strength);
      nref += size;
    }
    return lp;
  }
  void resample_py_params(int cls) {
    TableTrace& t = tables[cls];
    std::vector<int64_t> counts;
    for (auto& pr : t.reference_counts) counts.push_back(pr.second);
    ++t.py_epoch;
    pclean_stream s; s.key.seed = seed; s.key.sweep = t.py_epoch; s.key.cls = (uint32_t)cls; s.key.row = cls;
    s.key.particle = 0; s.key.block = 0; s.key.site = 0; s.key.purpose = PCLEAN_RNG_PY; s.idx = 0;
    double cs = t.strength, cd = t.discount;
    double old_score = pitman_yor_score(cs, cd, counts);
    double u = pclean_next(&s); if (u < 1e-300) u = 1e-300;
    double proposed = -std::log(u);                       // Gamma(1,1)
    double new_score = pitman_yor_score(proposed, cd, counts);
    double old_q = -cs, new_q = -proposed;                // logpdf(Gamma(1,1), x) = -x
    double alpha = new_score + old_q - old_score - new_q;
    double u2 = pclean_next(&s);
    if (std::log(u2) < alpha) { cs = proposed; old_score = new_score; }
    double pd = pclean_next(&s);
    new_score = pitman_yor_score(cs, pd, counts);
    double u3 = pclean_next(&s);
    if (std::log(u3) < new_score - old_score) cd = pd;
    t.strength = cs; t.discount = cd;
  }

  // ---------------------------------------------------------------- enumeration (proposal_compiler.jl)
  struct EnumCtx {
    Oracle* o; RowState* st; int cls; const ClassM* cm;
    std::vector<char> obs;              // observation_indices = keys(state.row_trace) at call time
    std::vector<Val> bound;             // variable_names that carry a value during the walk
    std::vector<char> is_bound;
    std::map<int, const Row*> active_child;       // fk vertex -> candidate row
    bool in_external = false;
    const Row* active_parent = nullptr;
    std::vector<Val> recomputed; std::vector<char> recomputed_set;
    pclean_rng_key rk;

    bool avail(int k) const { return obs[k] || is_bound[k]; }
    bool any_unavailable(const std::vector<int>& a) const { for (int k : a) if (!avail(k)) return true; return false; }
    Val value(int k) const { return is_bound[k] ? bound[k] : st->row[k]; }
    void bind(int k, const Val& v) { bound[k] = v; is_bound[k] = 1; }
    void unbind(int k) { is_bound[k] = 0; }
    Val retained(int idx) const { return st->retained ? (*st->retained)[idx] : mk(PCLEAN_VAL_ABSENT); }
    std::vector<Val> args_of(const Node& n) const { std::vector<Val> a; for (int k : n.args) a.push_back(value(k)); return a; }
    double uniform_at(int site) { rk.site = (uint32_t)site; rk.purpose = PCLEAN_RNG_ENUM; return pclean_uniform(&rk, 0); }

    Res plan(const Plan& steps) {                       // process_plan! :363-388
      if (steps.empty()) return Res();
      if (steps.size() == 1) return step(steps[0]);
      Res out;
      for (const PlanNode& s : steps) {
        Res r = step(s);
        out.p += r.p; out.q += r.q;
        for (auto& e : r.t) out.t.push_back(e);
      }
      return out;
    }
    Res step(const PlanNode& s) {
      const Node& n = cm->nodes[s.v];
      if (n.wrap == PCLEAN_WRAP_EXTERNAL) return external(n, s.v, s.kids);
      if (n.wrap == PCLEAN_WRAP_SUBMODEL) return submodel(n, 0, s.v, s.kids);
      return base(n, s.v, s.kids);
    }
    Res base(const Node& n, int idx, const Plan& rest) {
      switch (n.kind) {
        case PCLEAN_NODE_JULIA: return julia(n, idx, rest);
        case PCLEAN_NODE_CHOICE: return choice(n, idx, rest);
        case PCLEAN_NODE_FK: return foreign_key(n, idx, rest);
        default: return plan(rest);
      }
    }
    Res julia(const Node& n, int idx, const Plan& rest) {           // :40-52
      if (any_unavailable(n.args)) return plan(rest);
      bind(idx, o->eval_func(n.func, args_of(n)));
      Res r = plan(rest);
      unbind(idx);
      return r;
    }
    Res choice(const Node& n, int idx, const Plan& rest) {          // :55-129
      if (!obs[idx] && !has_discrete_proposal(n.dist)) return plan(rest);
      if (any_unavailable(n.args)) return plan(rest);
      if (obs[idx]) {
        Res r = plan(rest);
        r.p += o->logdensity(n.dist, st->row[idx], args_of(n));
        return r;
      }
      std::vector<Val> a = args_of(n), options; std::vector<double> prior;
      o->discrete_proposal(n.dist, a, options, prior);
      Val ret = retained(idx);
      int chosen = -1;
      std::vector<double> plist; std::vector<double> qlist; std::vector<std::vector<std::pair<int, Val>>> tlist;
      for (size_t it = 0; it < options.size(); ++it) {
        Val x = options[it];
        if (x.tag == PCLEAN_VAL_DUMMY) x = o->dummy_value(n.dist, a);
        if (present(ret) && val_eq(x, ret)) chosen = (int)it;
        bind(idx, x);
        Res r = plan(rest);
        plist.push_back(r.p + prior[it]); qlist.push_back(r.q); tlist.push_back(std::move(r.t));
      }
      unbind(idx);
      Res out;
      out.p = logsumexp(plist);
      for (double& v : plist) v -= out.p;
      if (chosen < 0) {
        std::vector<double> pr(plist.size());
        for (size_t i = 0; i < pr.size(); ++i) pr[i] = std::exp(plist[i]);
        if (!(out.p > NEG_INF)) throw OracleError("Categorical: enumeration has zero total mass");
        chosen = categorical(pr, uniform_at(idx));
      }
      out.t = std::move(tlist[chosen]);
      out.t.emplace_back(idx, options[chosen]);
      out.q = qlist[chosen] + plist[chosen];
      return out;
    }
    Res foreign_key(const Node& n, int idx, const Plan& rest) {     // :131-247
      TableTrace& table = o->tables[n.target];
      const ClassM& tm = o->m.classes[n.target];
      bool can_hash = !tm.hash_keys.empty();
      for (int h : tm.hash_keys) if (!obs[n.vmap[h]]) can_hash = false;
      std::vector<int64_t> keys;
      if (can_hash) {
        std::vector<std::pair<int, int64_t>> hk;
        for (int h : tm.hash_keys) {
          const Val& v = st->row[n.vmap[h]];
          hk.emplace_back(v.tag, v.tag == PCLEAN_VAL_REAL || v.tag == PCLEAN_VAL_KEY ? (int64_t)v.d : (int64_t)v.i);
        }
        auto it = table.hashed_keys.find(hk);
        if (it != table.hashed_keys.end()) keys.assign(it->second.begin(), it->second.end());
      } else {
        for (auto& pr : table.rows) keys.push_back(pr.first);
      }
      const size_t nk = keys.size();
      double logden = std::log(table.total_references + table.strength);
      std::vector<double> py;
      for (int64_t k : keys) py.push_back(std::log(table.reference_counts.at(k) - table.discount) - logden);
      py.push_back(std::log(table.strength + table.discount * (double)table.rows.size()) - logden);
      Val ret = retained(idx);
      bool ret_is_key = ret.tag == PCLEAN_VAL_KEY;
      int64_t new_key;
      if (!ret_is_key || table.rows.count(key_of(ret))) new_key = ++o->gensym;      // pclean_gensym!, :186-192
      else new_key = key_of(ret);
      int chosen = -1;
      std::vector<double> plist, qlist; std::vector<std::vector<std::pair<int, Val>>> tlist;
      for (size_t it = 0; it < nk; ++it) {
        if (ret_is_key && keys[it] == key_of(ret)) chosen = (int)it;
        active_child[idx] = &table.rows.at(keys[it]);
        bind(idx, mk_key(keys[it]));
        Res r = plan(rest);
        plist.push_back(r.p + py[it]); qlist.push_back(r.q); tlist.push_back(std::move(r.t));
      }
      active_child.erase(idx);
      bind(idx, mk_key(new_key));
      if (ret_is_key && new_key == key_of(ret)) chosen = (int)nk;
      {
        Res r = plan(rest);
        plist.push_back(r.p + py.back()); qlist.push_back(r.q); tlist.push_back(std::move(r.t));
      }
      unbind(idx);
      keys.push_back(new_key);
      Res out;
      out.p = logsumexp(plist);
      for (double& v : plist) v -= out.p;
      if (chosen < 0) {
        if (!(out.p > NEG_INF)) throw OracleError("Categorical: foreign-key enumeration has zero total mass");
        std::vector<double> pr(plist.size());
        for (size_t i = 0; i < pr.size(); ++i) pr[i] = std::exp(plist[i]);
        chosen = categorical(pr, uniform_at(idx));
      }
      out.t = std::move(tlist[chosen]);
      out.t.emplace_back(idx, mk_key(keys[chosen]));
      out.q = qlist[chosen] + plist[chosen];
      return out;
    }
    bool can_process_base(const Node& n, int idx) const {           // :249-252
      if (n.kind == PCLEAN_NODE_JULIA) return !any_unavailable(n.args);
      if (n.kind == PCLEAN_NODE_CHOICE) return !any_unavailable(n.args) && (obs[idx] || has_discrete_proposal(n.dist));
      if (n.kind == PCLEAN_NODE_FK) return true;
      return false;
    }
    Res submodel(const Node& n, size_t level, int idx, const Plan& rest) {   // :254-300
      if (level >= n.wfk.size()) return base(n, idx, rest);
      if (!(obs[idx] || can_process_base(n, idx))) return plan(rest);
      auto it = active_child.find(n.wfk[level]);
      if (it == active_child.end()) return submodel(n, level + 1, idx, rest);       // case 1
      const Val& chosen_value = (*it->second)[n.wsub[level]];
      if (obs[idx]) {                                                               // case 2
        const Val& mine = st->row[idx];
        bool close = (mine.tag == PCLEAN_VAL_MISSING && chosen_value.tag == PCLEAN_VAL_MISSING) ||
                     (chosen_value.tag == PCLEAN_VAL_REAL && (mine.tag == PCLEAN_VAL_REAL || mine.tag == PCLEAN_VAL_INT) &&
                      isapprox(chosen_value.d, mine.tag == PCLEAN_VAL_REAL ? mine.d : (double)mine.i)) ||
                     (chosen_value.tag != PCLEAN_VAL_MISSING && mine.tag != PCLEAN_VAL_MISSING && val_eq(chosen_value, mine));
        if (!close) { Res r; r.p = NEG_INF; r.q = NEG_INF; return r; }
        return plan(rest);
      }
      bind(idx, chosen_value);                                                      // case 3
      Res r = plan(rest);
      unbind(idx);
      return r;
    }
    Val ext_value(int i) const { return recomputed_set[i] ? recomputed[i] : (*active_parent)[i]; }
    Res external(const Node& n, int idx, const Plan& rest) {        // :306-350
      if (in_external) {
        std::vector<Val> a;
        for (int k : n.args) a.push_back(ext_value(k));
        if (n.kind == PCLEAN_NODE_JULIA) {
          recomputed[n.extv] = o->eval_func(n.func, a); recomputed_set[n.extv] = 1;
          return plan(rest);
        }
        if (n.kind == PCLEAN_NODE_CHOICE) {
          Res r = plan(rest);
          r.p += o->logdensity(n.dist, (*active_parent)[n.extv], a);
          return r;
        }
        throw OracleError("ExternalLikelihoodNode{ForeignKeyNode} is not supported");
      }
      const PathM& path = o->m.paths[n.path];
      const int src = path.links.back().first;
      TableTrace& stab = o->tables[src];
      Res out;
      auto rit = st->referring->find(n.path);
      if (rit == st->referring->end()) return out;
      in_external = true;
      const int snv = o->m.classes[src].nv;
      for (int64_t pk : rit->second) {
        active_parent = &stab.rows.at(pk);
        recomputed.assign(snv, mk(PCLEAN_VAL_ABSENT)); recomputed_set.assign(snv, 0);
        for (size_t tv = 0; tv < path.vmap.size(); ++tv)
          if (path.vmap[tv] >= 0 && is_bound[tv]) { recomputed[path.vmap[tv]] = bound[tv]; recomputed_set[path.vmap[tv]] = 1; }
        Res r = external(n, idx, rest);
        out.p += r.p; out.q += r.q;
      }
      in_external = false;
      active_parent = nullptr;
      return out;
    }
  };

  // prune_plan (block_proposal.jl:3-22)
  Plan prune_plan(const Plan& plan, const RowState& st, const ClassM& cm) {
    Plan out;
    for (const PlanNode& s : plan) {
      Plan sub = prune_plan(s.kids, st, cm);
      if (!sub.empty()) { PlanNode n; n.v = s.v; n.kids = std::move(sub); out.push_back(std::move(n)); }
      else if (st.has(s.v) || cm.nodes[s.v].wrap == PCLEAN_WRAP_EXTERNAL) { PlanNode n; n.v = s.v; out.push_back(std::move(n)); }
    }
    return out;
  }

  // propose_non_enumerable! (block_proposal.jl:24-157)
  struct NonEnum {
    Oracle* o; RowState* st; const ClassM* cm; pclean_rng_key rk; double p = 0.0, q_cont = 0.0;

    void node(const Node& n, size_t level, int idx, const ClassM* arg_class) {
      (void)arg_class;
      if (level < n.wfk.size()) {                                   // SubmodelNode, :99-110
        Val fkv = st->get(n.wfk[level]);
        const Node& fkn = cm->nodes[n.wfk[level]];
        TableTrace& tt = o->tables[fkn.target];
        auto it = tt.rows.find(key_of(fkv));
        if (it == tt.rows.end()) node(n, level + 1, idx, arg_class);
        else if (!st->has(idx)) st->set(idx, it->second[n.wsub[level]]);
        return;
      }
      switch (n.kind) {
        case PCLEAN_NODE_JULIA: {                                   // :32-36
          std::vector<Val> a; for (int k : n.args) a.push_back(st->get(k));
          st->set(idx, o->eval_func(n.func, a));
          return;
        }
        case PCLEAN_NODE_CHOICE: {                                  // :38-66
          std::vector<Val> a; for (int k : n.args) a.push_back(st->get(k));
          if (!st->has(idx) && has_discrete_proposal(n.dist)) {
            std::vector<Val> options; std::vector<double> lp;
            o->discrete_proposal(n.dist, a, options, lp);
            int chosen = -1;
            if (!st->retained) {
              double tot = logsumexp(lp);
              std::vector<double> pr(lp.size());
              for (size_t i = 0; i < lp.size(); ++i) pr[i] = std::exp(lp[i] - tot);
              rk.site = (uint32_t)idx; rk.purpose = PCLEAN_RNG_PRIOR;
              chosen = categorical(pr, pclean_uniform(&rk, 0));
            } else {
              const Val& rv = (*st->retained)[idx];
              for (size_t i = 0; i < options.size(); ++i) if (val_eq(options[i], rv)) { chosen = (int)i; break; }
              if (chosen < 0) for (size_t i = 0; i < options.size(); ++i) if (options[i].tag == PCLEAN_VAL_DUMMY) { chosen = (int)i; break; }
              if (chosen < 0) throw OracleError("retained value is not among the proposal options");
            }
            st->set(idx, options[chosen]);
            q_cont += lp[chosen];
          }
          if (!st->has(idx) || st->get(idx).tag == PCLEAN_VAL_DUMMY) {
            if (!st->retained) {
              pclean_stream s; s.key = rk; s.key.site = (uint32_t)idx; s.key.purpose = PCLEAN_RNG_RANDOM; s.idx = 0;
              st->set(idx, o->random_value(n.dist, a, s));
            } else st->set(idx, (*st->retained)[idx]);
          } else {
            p += o->logdensity(n.dist, st->get(idx), a);
          }
          return;
        }
        case PCLEAN_NODE_FK: {                                      // :68-97
          TableTrace& tt = o->tables[n.target];
          if (!st->has(idx)) {
            if (!st->retained) {
              std::vector<int64_t> keys; std::vector<double> w;
              double logden = std::log(tt.total_references + tt.strength);
              for (auto& pr : tt.reference_counts) { keys.push_back(pr.first); w.push_back(std::exp(std::log(pr.second - tt.discount) - logden)); }
              w.push_back(std::exp(std::log(keys.size() * tt.discount + tt.strength) - logden));
              rk.site = (uint32_t)idx; rk.purpose = PCLEAN_RNG_FKPRIOR;
              int c = categorical(w, pclean_uniform(&rk, 0));
              st->set(idx, mk_key(c < (int)keys.size() ? keys[c] : ++o->gensym));
            } else st->set(idx, (*st->retained)[idx]);
          } else {
            int64_t fk = key_of(st->get(idx));
            double logden = std::log(tt.total_references + tt.strength);
            if (tt.rows.count(fk)) p += std::log(tt.reference_counts.at(fk) - tt.discount) - logden;
            else p += std::log(tt.discount * (double)tt.rows.size() + tt.strength) - logden;
          }
          return;
        }
        default: return;
      }
    }

    void run(const std::vector<int>& order) {
      size_t i = 0;
      while (i < order.size() && cm->nodes[order[i]].wrap != PCLEAN_WRAP_EXTERNAL) {
        node(cm->nodes[order[i]], 0, order[i], cm);
        ++i;
      }
      while (i < order.size()) {                                    // :119-155
        const Node& first = cm->nodes[order[i]];
        const int pid = first.path;
        const PathM& path = o->m.paths[pid];
        const int src = path.links.back().first;
        TableTrace& stab = o->tables[src];
        const int snv = o->m.classes[src].nv;
        size_t next_i = i + 1;
        auto rit = st->referring->find(pid);
        if (rit != st->referring->end()) {
          for (int64_t rk_ : rit->second) {
            st->active_parent = nullptr;
            st->recomputed.assign(snv, mk(PCLEAN_VAL_ABSENT));
            for (size_t tv = 0; tv < path.vmap.size(); ++tv)
              if (path.vmap[tv] >= 0) st->recomputed[path.vmap[tv]] = st->has((int)tv) ? st->get((int)tv) : mk(TAG_NOTHING);
            st->active_parent = &stab.rows.at(rk_);
            size_t j = i;
            while (j < order.size() && cm->nodes[order[j]].wrap == PCLEAN_WRAP_EXTERNAL && cm->nodes[order[j]].path == pid) {
              const Node& en = cm->nodes[order[j]];
              Node plain = en; plain.wrap = PCLEAN_WRAP_NONE; plain.wfk.clear(); plain.wsub.clear();
              node(plain, 0, en.extv, nullptr);
              ++j;
            }
            next_i = j;
          }
        } else {
          size_t j = i;
          while (j < order.size() && cm->nodes[order[j]].wrap == PCLEAN_WRAP_EXTERNAL && cm->nodes[order[j]].path == pid) ++j;
          next_i = j;
        }
        i = next_i;
        st->active_parent = nullptr;
        st->recomputed.clear();
      }
    }
  };

  // make_block_proposal! (block_proposal.jl:160-191)
  double make_block_proposal(RowState& st, int block_index, int particle) {
    const ClassM& cm = m.classes[st.cls];
    double q_disc = 0.0;
    pclean_rng_key rk; rk.seed = seed; rk.sweep = cur_sweep; rk.cls = (uint32_t)st.cls; rk.row = st.key;
    rk.particle = (uint32_t)particle; rk.block = (uint32_t)block_index; rk.site = 0; rk.purpose = 0;
    int64_t chosen_key = -2;
    if (cfg.use_dd_proposals) {
      Plan pruned = prune_plan(cm.plans[block_index], st, cm);
      EnumCtx cx; cx.o = this; cx.st = &st; cx.cls = st.cls; cx.cm = &cm; cx.rk = rk;
      cx.obs.resize(cm.nv); cx.bound.assign(cm.nv, mk(PCLEAN_VAL_ABSENT)); cx.is_bound.assign(cm.nv, 0);
      for (int i = 0; i < cm.nv; ++i) cx.obs[i] = present(st.row[i]);
      Res r = cx.plan(pruned);
      q_disc = r.q;
      for (auto& e : r.t) st.row[e.first] = e.second;
      if (!pruned.empty()) {
        int root = pruned[0].v;
        const Node& rn = cm.nodes[root];
        if (rn.wrap == PCLEAN_WRAP_NONE && rn.kind == PCLEAN_NODE_FK && present(st.row[root])) {
          int64_t k = key_of(st.row[root]);
          chosen_key = tables[rn.target].rows.count(k) ? k : -1;
        }
      }
    }
    NonEnum ne; ne.o = this; ne.st = &st; ne.cm = &cm; ne.rk = rk;
    ne.run(cm.blocks[block_index]);
    if (record) record->chosen_keys.push_back(chosen_key);
    return ne.p - q_disc - ne.q_cont;
  }

  // ---------------------------------------------------------------- run_smc! (row_inference.jl)
  void fill_parameters(int cls, Row& row, const std::function<int(int)>& vmapf) {     // :49-59
    for (auto& pr : tables[cls].parameters) row[vmapf(pr.first)] = pr.second;
    const ClassM& cm = m.classes[cls];
    for (int i = 0; i < cm.nv; ++i) {
      const Node& n = cm.nodes[i];
      if (n.wrap != PCLEAN_WRAP_NONE || n.kind != PCLEAN_NODE_FK) continue;
      const Node* np = &n;
      fill_parameters(n.target, row, [np, &vmapf](int j) { return vmapf(np->vmap[j]); });
    }
  }
  ReferringRows collect_referring_rows(int cls, int64_t key) {                        // :23-47
    ReferringRows out;
    TableTrace& t = tables[cls];
    auto dit = t.direct_incoming.find(key);
    if (dit == t.direct_incoming.end()) return out;
    const ClassM& cm = m.classes[cls];
    std::vector<int> pids = cm.paths;
    std::stable_sort(pids.begin(), pids.end(), [&](int a, int b) { return m.paths[a].links.size() < m.paths[b].links.size(); });
    for (int pid : pids) {
      const PathM& path = m.paths[pid];
      std::pair<int, int> last = path.links.back();
      std::set<int64_t> acc;
      if (path.links.size() == 1) {
        auto sit = dit->second.find(last);
        if (sit != dit->second.end()) acc = sit->second;
      } else {
        // find the path with links[0..n-2]
        int prev = -1;
        for (int q : cm.paths) {
          const PathM& pq = m.paths[q];
          if (pq.links.size() + 1 == path.links.size() && std::equal(pq.links.begin(), pq.links.end(), path.links.begin())) { prev = q; break; }
        }
        if (prev < 0) throw OracleError("collect_referring_rows: prefix path missing");
        TableTrace& lt = tables[m.paths[prev].links.back().first];
        for (int64_t k : out[prev]) {
          auto kit = lt.direct_incoming.find(k);
          if (kit == lt.direct_incoming.end()) continue;
          auto sit = kit->second.find(last);
          if (sit != kit->second.end()) acc.insert(sit->second.begin(), sit->second.end());
        }
      }
      out[pid].assign(acc.begin(), acc.end());
    }
    return out;
  }

  double run_smc(int cls, int64_t key) {                                              // :108-187
    TableTrace& table = tables[cls];
    const ClassM& cm = m.classes[cls];
    const int K = cfg.num_particles;
    const bool is_csmc = table.rows.count(key) > 0;
    Row retained_row;
    if (is_csmc) { retained_row = table.rows.at(key); unincorporate_row(cls, key); }
    Row start = table.observations.at(key);
    fill_parameters(cls, start, [](int i) { return i; });
    ReferringRows referring = collect_referring_rows(cls, key);
    std::vector<Particle> particles(K);
    for (int j = 0; j < K; ++j) {
      particles[j].state.cls = cls; particles[j].state.row = start; particles[j].state.key = key;
      particles[j].state.referring = &referring; particles[j].state.retained = nullptr;
    }
    double log_ml = 0.0;
    const int nb = (int)cm.blocks.size();
    std::vector<int64_t> chosen_by_block;      // record layout [K][nb]
    MoveRecord* rec = record;
    std::vector<std::vector<int64_t>> rec_keys(K, std::vector<int64_t>(nb, -2));
    for (int b = 0; b < nb; ++b) {
      for (int j = 0; j < K; ++j) {
        if (j == 0) particles[j].state.retained = is_csmc ? &retained_row : nullptr;
        MoveRecord tmp; record = rec ? &tmp : nullptr;
        double w = make_block_proposal(particles[j].state, b, j);
        if (rec && !tmp.chosen_keys.empty()) rec_keys[j][b] = tmp.chosen_keys[0];
        record = rec;
        particles[j].weight += w;
        particles[j].block_index += 1;
      }
      if (!cfg.use_mh_instead_of_pg && b < nb - 1) {                                   // maybe_resample :87-105
        std::vector<double> lw(K);
        for (int j = 0; j < K; ++j) lw[j] = particles[j].weight;
        double tot = logsumexp(lw);
        std::vector<double> ln(K), l2(K);
        for (int j = 0; j < K; ++j) { ln[j] = lw[j] - tot; l2[j] = 2.0 * ln[j]; }
        double ess = std::exp(-logsumexp(l2));
        if (ess < K / 2.0) {
          std::vector<double> w(K);
          for (int j = 0; j < K; ++j) w[j] = std::exp(ln[j]);
          std::vector<int> idx(K);
          pclean_rng_key rk; rk.seed = seed; rk.sweep = cur_sweep; rk.cls = (uint32_t)cls; rk.row = key;
          rk.block = (uint32_t)b; rk.site = 0; rk.purpose = PCLEAN_RNG_RESAMPLE;
          for (int j = 0; j < K; ++j) {
            if (j == 0 && is_csmc) { idx[j] = 0; continue; }
            rk.particle = (uint32_t)j;
            idx[j] = categorical(w, pclean_uniform(&rk, 0));
          }
          std::vector<Particle> np(K);
          std::vector<std::vector<int64_t>> nk(K);
          for (int j = 0; j < K; ++j) {
            np[j].state = particles[idx[j]].state;          // clone_with_zero_weight :17-21 (retained_trace = nothing)
            np[j].state.retained = nullptr;
            np[j].weight = 0.0; np[j].block_index = particles[idx[j]].block_index;
            nk[j] = rec_keys[idx[j]];
          }
          particles.swap(np); rec_keys.swap(nk);
          log_ml += tot - std::log((double)K);
        }
      }
    }
    std::vector<double> lw(K);
    for (int j = 0; j < K; ++j) lw[j] = particles[j].weight;
    double tot = logsumexp(lw);
    std::vector<double> w(K);
    for (int j = 0; j < K; ++j) w[j] = std::exp(lw[j] - tot);
    pclean_rng_key rk; rk.seed = seed; rk.sweep = cur_sweep; rk.cls = (uint32_t)cls; rk.row = key;
    rk.particle = 0; rk.block = (uint32_t)nb; rk.site = 0; rk.purpose = PCLEAN_RNG_FINAL;
    double u = pclean_uniform(&rk, 0);
    int chosen;
    if (cfg.use_mh_instead_of_pg && is_csmc) chosen = (u < std::min(1.0, w[1] / (1e-10 + w[0]))) ? 1 : 0;
    else chosen = categorical(w, u);
    Row chosen_row = particles[chosen].state.row;
    table.rows[key] = chosen_row;
    incorporate_row(cls, key);
    if (is_csmc) {
      if (chosen != 0) {
        update_sufficient_statistics(cls, retained_row, -1);
        update_sufficient_statistics(cls, table.rows.at(key), +1);
        update_referring_rows(cls, chosen_row, referring);
      }
    } else update_sufficient_statistics(cls, table.rows.at(key), +1);
    double ret = log_ml + tot - std::log((double)K);
    if (rec) {
      rec->chosen_keys.clear();
      for (int j = 0; j < K; ++j) for (int b = 0; b < nb; ++b) rec->chosen_keys.push_back(rec_keys[j][b]);
      rec->weights = lw; rec->selected = chosen; rec->log_ml = ret;
    }
    return ret;
  }

  // ---------------------------------------------------------------- trace installation (test/bench harness)
  // Install a latent row given its denormalised cells, as refer_to_row! would have created it
  // (dependency_tracking.jl:205-236), classes in class order so targets exist first.
  void install_latent_row(int cls, int64_t key, const Row& cells) {
    TableTrace& t = tables[cls];
    Row row = cells;
    row.resize(m.classes[cls].nv, mk(PCLEAN_VAL_ABSENT));
    fill_parameters(cls, row, [](int i) { return i; });
    // zero-argument JuliaNodes (option lists, literals) are part of every row
    const ClassM& cm = m.classes[cls];
    for (int i = 0; i < cm.nv; ++i) {
      const Node& n = cm.nodes[i];
      if (n.wrap == PCLEAN_WRAP_EXTERNAL || present(row[i])) continue;
      if (n.kind == PCLEAN_NODE_JULIA) {
        bool ok = true; std::vector<Val> a;
        for (int k : n.args) { if (!present(row[k])) ok = false; a.push_back(row[k]); }
        if (ok) row[i] = eval_func(n.func, a);
      }
    }
    t.rows[key] = row;
    t.reference_counts[key] = 0;
    t.observations[key] = Row(cm.nv, mk(PCLEAN_VAL_ABSENT));
    t.observation_counts[key];
    t.direct_incoming[key];
    incorporate_row(cls, key);
    update_sufficient_statistics(cls, t.rows.at(key), +1);
    gensym = std::max(gensym, key);
  }
  // Install an observation row that references the given target keys (top-level slots in
  // vertex order); the rest of the row is filled in as propose_non_enumerable! would.
  void install_obs_row(int cls, int64_t key, const std::vector<int64_t>& fk_keys) {
    TableTrace& t = tables[cls];
    const ClassM& cm = m.classes[cls];
    RowState st; st.cls = cls; st.key = key; st.row = t.observations.at(key);
    fill_parameters(cls, st.row, [](int i) { return i; });
    size_t f = 0;
    for (int i = 0; i < cm.nv; ++i) {
      const Node& n = cm.nodes[i];
      if (n.wrap == PCLEAN_WRAP_NONE && n.kind == PCLEAN_NODE_FK) st.row[i] = mk_key(fk_keys.at(f++));
    }
    ReferringRows none; st.referring = &none;
    for (size_t b = 0; b < cm.blocks.size(); ++b) {
      NonEnum ne; ne.o = this; ne.st = &st; ne.cm = &cm;
      ne.rk.seed = seed; ne.rk.sweep = 0; ne.rk.cls = (uint32_t)cls; ne.rk.row = key; ne.rk.particle = 0; ne.rk.block = (uint32_t)b; ne.rk.site = 0; ne.rk.purpose = 0;
      ne.run(cm.blocks[b]);
    }
    t.rows[key] = st.row;
    incorporate_row(cls, key);
    update_sufficient_statistics(cls, t.rows.at(key), +1);
  }
  void bump_refcount(int cls, int64_t key, int64_t extra) {
    TableTrace& t = tables[cls];
    t.reference_counts.at(key) += extra;
    t.total_references += extra;
  }

  // ---------------------------------------------------------------- drivers (inference.jl)
  void create_tables() {
    const int nc = (int)m.classes.size();
    tables.assign(nc, TableTrace());
    slots.assign(m.slot_param.size(), ParamSlot());
    for (size_t s = 0; s < slots.size(); ++s) { slots[s].spec = m.slot_param[s]; init_slot((int)s); }
    for (int c = 0; c < nc; ++c) {
      tables[c].strength = m.classes[c].py_strength; tables[c].discount = m.classes[c].py_discount;
      for (int v = 0; v < m.classes[c].nv; ++v) {
        const Node& n = m.classes[c].nodes[v];
        if (n.wrap == PCLEAN_WRAP_NONE && n.kind == PCLEAN_NODE_PARAM) {
          if (m.param_indexed[n.param]) tables[c].parameters.emplace_back(v, mk(PCLEAN_VAL_IPARAM, n.param));
          else {
            int slot = -1;
            for (size_t s = 0; s < slots.size(); ++s) if (slots[s].spec == n.param) { slot = (int)s; break; }
            tables[c].parameters.emplace_back(v, mk(PCLEAN_VAL_PARAM, slot));
          }
        }
      }
    }
  }
  void initialize_trace(int64_t limit = -1) {                                          // inference.jl:3-58 (limit: tests stop after `limit` rows)
    create_tables();
    cur_sweep = 0;
    for (const Obs& ds : datasets) {
      TableTrace& t = tables[ds.cls];
      const int nv = m.classes[ds.cls].nv;
      for (int64_t i = 0; i < ds.n; ++i) {
        Row obs(nv, mk(PCLEAN_VAL_ABSENT));
        for (size_t c = 0; c < ds.vertex_of_col.size(); ++c) {
          const Val& v = ds.cells[c * ds.n + i];
          if (present(v)) obs[ds.vertex_of_col[c]] = v;
        }
        t.observations[i] = std::move(obs);
        if (limit >= 0 && i >= limit) continue;
        run_smc(ds.cls, i);
        if ((i + 1) % cfg.rejuv_frequency == 0) {
          for (size_t c = 0; c < m.classes.size(); ++c) { resample_parameters_of_class((int)c); resample_py_params((int)c); }
        }
      }
    }
  }
  void sweep_class(int cls, int64_t row_begin = 0, int64_t row_end = -1) {             // inference.jl:60-81
    TableTrace& t = tables[cls];
    std::vector<int64_t> keys;
    for (auto& pr : t.rows) keys.push_back(pr.first);
    int64_t i = 0;
    for (int64_t key : keys) {
      ++i;
      if (i <= row_begin) continue;
      if (row_end >= 0 && i > row_end) break;
      if (!t.rows.count(key)) continue;          // row was garbage-collected earlier in this sweep
      if (i % cfg.rejuv_frequency == 0) { resample_parameters_of_class(cls); resample_py_params(cls); }
      run_smc(cls, key);
    }
  }
  void sweep() {
    ++cur_sweep;
    for (size_t c = 0; c < m.classes.size(); ++c) sweep_class((int)c);
  }
};

// ------------------------------------------------------------------------------------------
// IR parsing
// ------------------------------------------------------------------------------------------
static void parse_ir(const pclean_model_ir* ir, Oracle& o) {
  Model& m = o.m;
  m.classes.resize(ir->n_classes);
  for (int c = 0; c < ir->n_classes; ++c) {
    ClassM& cm = m.classes[c];
    int v0 = ir->class_voff[c], v1 = ir->class_voff[c + 1];
    cm.nv = v1 - v0;
    cm.py_strength = ir->py_strength[c]; cm.py_discount = ir->py_discount[c];
    cm.nodes.resize(cm.nv);
    for (int g = v0; g < v1; ++g) {
      Node& n = cm.nodes[g - v0];
      n.kind = ir->v_kind[g]; n.wrap = ir->v_wrap[g]; n.dist = ir->v_dist[g]; n.func = ir->v_func[g];
      n.target = ir->v_target[g]; n.param = ir->v_param[g]; n.path = ir->v_path[g]; n.extv = ir->v_extv[g];
      for (int k = ir->v_wrap_off[g]; k < ir->v_wrap_off[g + 1]; ++k) { n.wfk.push_back(ir->wrap_fk[k]); n.wsub.push_back(ir->wrap_subid[k]); }
      for (int k = ir->v_args_off[g]; k < ir->v_args_off[g + 1]; ++k) n.args.push_back(ir->v_args[k]);
      for (int k = ir->v_vmap_off[g]; k < ir->v_vmap_off[g + 1]; ++k) n.vmap.push_back(ir->v_vmap[k]);
      if (n.wrap != PCLEAN_WRAP_EXTERNAL) cm.n_normal = g - v0 + 1;
    }
    for (int b = ir->class_block_off[c]; b < ir->class_block_off[c + 1]; ++b) {
      std::vector<int> blk;
      for (int k = ir->block_voff[b]; k < ir->block_voff[b + 1]; ++k) blk.push_back(ir->block_v[k]);
      cm.blocks.push_back(blk);
      int pos = ir->plan_off[b];
      int nroots = ir->plan_nchild[pos]; ++pos;
      cm.plans.push_back(parse_plan(ir->plan_vertex, ir->plan_nchild, pos, nroots));
    }
    for (int k = ir->class_hash_off[c]; k < ir->class_hash_off[c + 1]; ++k) cm.hash_keys.push_back(ir->hash_v[k]);
  }
  m.paths.resize(ir->n_paths);
  for (int p = 0; p < ir->n_paths; ++p) {
    PathM& pm = m.paths[p];
    pm.target = ir->path_target[p];
    for (int k = ir->path_len_off[p]; k < ir->path_len_off[p + 1]; ++k) pm.links.emplace_back(ir->path_class[k], ir->path_vertex[k]);
    for (int k = ir->path_vmap_off[p]; k < ir->path_vmap_off[p + 1]; ++k) pm.vmap.push_back(ir->path_vmap[k]);
    m.classes[pm.target].paths.push_back(p);
  }
  m.funcs.resize(ir->n_funcs);
  for (int f = 0; f < ir->n_funcs; ++f) {
    FuncM& fm = m.funcs[f];
    fm.kind = ir->func_kind[f]; fm.cst = ir->func_const[f];
    for (int k = ir->func_keyarg_off[f]; k < ir->func_keyarg_off[f + 1]; ++k) fm.keyargs.push_back(ir->func_keyargs[k]);
    for (int e = ir->func_tab_off[f]; e < ir->func_tab_off[f + 1]; ++e) {
      std::vector<int> key(ir->tab_keys + ir->tab_key_off[e], ir->tab_keys + ir->tab_key_off[e + 1]);
      fm.table[key] = ir->tab_vals[e];
    }
  }
  m.param_kind.assign(ir->param_kind, ir->param_kind + ir->n_params);
  m.param_indexed.assign(ir->param_indexed, ir->param_indexed + ir->n_params);
  m.param_prior0.assign(ir->param_prior0, ir->param_prior0 + ir->n_params);
  m.param_prior1.assign(ir->param_prior1, ir->param_prior1 + ir->n_params);
  m.slot_param.assign(ir->slot_param, ir->slot_param + ir->n_param_slots);
  m.lists.resize(ir->n_lists);
  for (int l = 0; l < ir->n_lists; ++l) m.lists[l].assign(ir->list_vals + ir->list_off[l], ir->list_vals + ir->list_off[l + 1]);
  m.xform_scale.assign(ir->xform_scale, ir->xform_scale + ir->n_xforms);
  std::memcpy(m.lm_uni, ir->lm_unigram, sizeof(m.lm_uni));
  std::memcpy(m.lm_big, ir->lm_bigram, sizeof(m.lm_big));
  o.strings.clear(); o.string_ids.clear();
  for (int s = 0; s < ir->n_strings; ++s) {
    std::u32string str(ir->str_cp + ir->str_off[s], ir->str_cp + ir->str_off[s + 1]);
    o.strings.push_back(str);
    o.string_ids.emplace(str, s);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C API (ctypes)
// ------------------------------------------------------------------------------------------
#define ORACLE_TRY(h, body)                                             \
  try { body; return 0; }                                               \
  catch (const std::exception& e) { (h)->last_error = e.what(); return -1; }

extern "C" {

void* oracle_create(const pclean_model_ir* ir, const pclean_config* cfg, uint64_t seed) {
  Oracle* o = new Oracle();
  try { parse_ir(ir, *o); } catch (const std::exception& e) { fprintf(stderr, "oracle_create: %s\n", e.what()); delete o; return nullptr; }
  o->cfg = *cfg; o->seed = seed;
  o->create_tables();
  return o;
}
void oracle_destroy(void* h) { delete (Oracle*)h; }
const char* oracle_last_error(void* h) { return ((Oracle*)h)->last_error.c_str(); }
void* oracle_clone(void* h) { return new Oracle(*(Oracle*)h); }
void oracle_set_config(void* h, const pclean_config* cfg) { ((Oracle*)h)->cfg = *cfg; }
void oracle_set_seed(void* h, uint64_t seed) { ((Oracle*)h)->seed = seed; }
void oracle_set_true_damerau(void* h, int flag) { ((Oracle*)h)->true_damerau = flag != 0; ((Oracle*)h)->typo_memo.clear(); }

int oracle_load_observations(void* h, const pclean_observations* obs) {
  Oracle* o = (Oracle*)h;
  ORACLE_TRY(o, {
    Oracle::Obs d; d.cls = obs->cls; d.n = obs->n_rows;
    d.vertex_of_col.assign(obs->vertex_of_col, obs->vertex_of_col + obs->n_cols);
    d.cells.assign(obs->cells, obs->cells + (size_t)obs->n_cols * obs->n_rows);
    o->datasets.push_back(std::move(d));
  });
}
int oracle_initialize_trace(void* h) { Oracle* o = (Oracle*)h; ORACLE_TRY(o, o->initialize_trace()); }
int oracle_initialize_prefix(void* h, int64_t n_rows) { Oracle* o = (Oracle*)h; ORACLE_TRY(o, o->initialize_trace(n_rows)); }
int oracle_sweep(void* h) { Oracle* o = (Oracle*)h; ORACLE_TRY(o, o->sweep()); }
int oracle_begin_sweep(void* h) { ((Oracle*)h)->cur_sweep += 1; return 0; }
int oracle_sweep_class(void* h, int cls, int64_t row_begin, int64_t row_end) {
  Oracle* o = (Oracle*)h; ORACLE_TRY(o, o->sweep_class(cls, row_begin, row_end));
}
int oracle_run_inference(void* h) {
  Oracle* o = (Oracle*)h;
  ORACLE_TRY(o, { for (int it = 0; it < o->cfg.num_iters; ++it) o->sweep(); });
}
// one run_smc! on the live trace, reporting what happened
int oracle_row_move(void* h, int cls, int64_t key, int64_t* chosen_keys, double* weights, int* selected, double* log_ml) {
  Oracle* o = (Oracle*)h;
  ORACLE_TRY(o, {
    MoveRecord rec; o->record = &rec;
    try { o->run_smc(cls, key); } catch (...) { o->record = nullptr; throw; }
    o->record = nullptr;
    std::copy(rec.chosen_keys.begin(), rec.chosen_keys.end(), chosen_keys);
    std::copy(rec.weights.begin(), rec.weights.end(), weights);
    *selected = rec.selected; *log_ml = rec.log_ml;
  });
}
int oracle_install_table(void* h, int cls, int64_t n_rows, int n_cols, const int64_t* keys, const pclean_value* cells) {
  Oracle* o = (Oracle*)h;
  ORACLE_TRY(o, {
    for (int64_t r = 0; r < n_rows; ++r) {
      Row row(n_cols);
      for (int v = 0; v < n_cols; ++v) row[v] = cells[(size_t)v * n_rows + r];
      o->install_latent_row(cls, keys[r], row);
    }
  });
}
// observations of rows [0, n_rows) must have been loaded; fk_keys is [n_fk][n_rows]
int oracle_install_obs_rows(void* h, int cls, int64_t n_rows, int n_fk, const int64_t* fk_keys, int64_t stride) {
  Oracle* o = (Oracle*)h;
  ORACLE_TRY(o, {
    const Oracle::Obs* ds = nullptr;
    for (const Oracle::Obs& d : o->datasets) if (d.cls == cls) ds = &d;
    if (!ds) throw OracleError("no observations loaded for this class");
    const int nv = o->m.classes[cls].nv;
    for (int64_t i = 0; i < n_rows; ++i) {
      Row obs(nv, mk(PCLEAN_VAL_ABSENT));
      for (size_t c = 0; c < ds->vertex_of_col.size(); ++c) { const Val& v = ds->cells[c * ds->n + i]; if (present(v)) obs[ds->vertex_of_col[c]] = v; }
      o->tables[cls].observations[i] = std::move(obs);
      std::vector<int64_t> fk(n_fk);
      for (int f = 0; f < n_fk; ++f) fk[f] = fk_keys[(size_t)f * stride + i];
      o->install_obs_row(cls, i, fk);
    }
  });
}
int oracle_bump_refcount(void* h, int cls, int64_t key, int64_t extra) { Oracle* o = (Oracle*)h; ORACLE_TRY(o, o->bump_refcount(cls, key, extra)); }
void oracle_reset_tables(void* h) { ((Oracle*)h)->create_tables(); }
int64_t oracle_table_size(void* h, int cls) { return (int64_t)((Oracle*)h)->tables[cls].rows.size(); }
int oracle_table_keys(void* h, int cls, int64_t* keys, int64_t* refcounts) {
  Oracle* o = (Oracle*)h;
  size_t k = 0;
  for (auto& pr : o->tables[cls].rows) {
    keys[k] = pr.first;
    auto it = o->tables[cls].reference_counts.find(pr.first);
    if (refcounts) refcounts[k] = it == o->tables[cls].reference_counts.end() ? 0 : it->second;
    ++k;
  }
  return 0;
}
// cells[vi * n_rows + r] for the rows in ascending key order
int oracle_get_cells(void* h, int cls, int n_vertices, const int32_t* vertices, pclean_value* out) {
  Oracle* o = (Oracle*)h;
  const size_t n = o->tables[cls].rows.size();
  size_t r = 0;
  for (auto& pr : o->tables[cls].rows) {
    for (int vi = 0; vi < n_vertices; ++vi) out[(size_t)vi * n + r] = pr.second[vertices[vi]];
    ++r;
  }
  return 0;
}
void oracle_get_py(void* h, int cls, double* strength, double* discount, int64_t* total_refs) {
  Oracle* o = (Oracle*)h;
  *strength = o->tables[cls].strength; *discount = o->tables[cls].discount; *total_refs = o->tables[cls].total_references;
}
void oracle_set_py(void* h, int cls, double strength, double discount) {
  Oracle* o = (Oracle*)h; o->tables[cls].strength = strength; o->tables[cls].discount = discount;
}
int oracle_string_count(void* h) { return (int)((Oracle*)h)->strings.size(); }
int oracle_get_string(void* h, int id, int cap, uint32_t* cp) {
  const std::u32string& s = ((Oracle*)h)->strings.at(id);
  int n = (int)std::min<size_t>(s.size(), (size_t)cap);
  for (int i = 0; i < n; ++i) cp[i] = s[i];
  return (int)s.size();
}
int oracle_intern_string(void* h, int n, const uint32_t* cp) {
  return ((Oracle*)h)->intern(std::u32string(cp, cp + n));
}
int oracle_param_get(void* h, int slot, int cap, double* values, int64_t* counts) {
  Oracle* o = (Oracle*)h;
  const ParamSlot& p = o->slots.at(slot);
  int n = (int)std::min<size_t>(p.value.size(), (size_t)cap);
  for (int i = 0; i < n; ++i) { values[i] = p.value[i]; if (counts) counts[i] = i < (int)p.counts.size() ? p.counts[i] : 0; }
  return (int)p.value.size();
}
int oracle_param_set(void* h, int slot, int n, const double* values) {
  Oracle* o = (Oracle*)h;
  ParamSlot& p = o->slots.at(slot);
  p.value.assign(values, values + n);
  if (p.counts.size() < (size_t)n && o->m.param_kind[p.spec] == PCLEAN_PARAM_PROPORTIONS) p.counts.resize(n, 0);
  return 0;
}
int oracle_n_slots(void* h) { return (int)((Oracle*)h)->slots.size(); }
int oracle_edit_distance(void* h, int a, int b) {
  Oracle* o = (Oracle*)h; return o->edit_distance(o->strings.at(a), o->strings.at(b));
}
double oracle_addtypos(void* h, int observed, int word, int max_typos) {
  Oracle* o = (Oracle*)h;
  try { return o->addtypos_logdensity(mk_str(observed), mk_str(word), max_typos); } catch (...) { return NAN; }
}
double oracle_stringprior(void* h, int sid, int minl, int maxl) { return ((Oracle*)h)->stringprior_logdensity(sid, minl, maxl); }
double oracle_logdensity(void* h, int dist, const pclean_value* obs, int nargs, const pclean_value* args) {
  Oracle* o = (Oracle*)h;
  try { return o->logdensity(dist, *obs, std::vector<Val>(args, args + nargs)); }
  catch (const std::exception& e) { o->last_error = e.what(); return NAN; }
}
void oracle_counters(void* h, int64_t* dp_cells, int64_t* typo_evals, int64_t* typo_misses) {
  Oracle* o = (Oracle*)h; *dp_cells = o->n_dp_cells; *typo_evals = o->n_typo_evals; *typo_misses = o->n_typo_misses;
}
double oracle_crp_logprior(int64_t count, double discount, double strength, int64_t total) {
  return std::log(count - discount) - std::log(total + strength);
}
double oracle_logsumexp(int n, const double* x) { return logsumexp(std::vector<double>(x, x + n)); }
/* parameter-move parity tests: put every resampling counter (the `sweep` field of the keyed
   PARAM / PY streams, include/pclean_rng.h) at a known value, then run exactly the rejuvenation
   step of pgibbs_sweep! (inference.jl:72-77) for one class */
void oracle_set_epochs(void* h, uint32_t epoch) {
  Oracle* o = (Oracle*)h;
  for (auto& p : o->slots) p.epoch = epoch;
  for (auto& t : o->tables) t.py_epoch = epoch;
}
int oracle_resample_class(void* h, int cls) {
  Oracle* o = (Oracle*)h;
  ORACLE_TRY(o, { o->resample_parameters_of_class(cls); o->resample_py_params(cls); });
}

}  // extern "C"
