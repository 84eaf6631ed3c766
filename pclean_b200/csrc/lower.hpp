// lower.hpp — host-side lowering of the model IR to device enumeration schedules.
//
// The reference JIT-compiles, per (class, block, missingness pattern), Julia source that
// enumerates discrete choices and candidate reference keys (proposal_compiler.jl:5-422).
// Here the same Plan forest is walked *symbolically* once at load time, following exactly
// the case analysis of that generator (JuliaNode :40-52, RandomChoiceNode :55-129,
// ForeignKeyNode :131-247, SubmodelNode :254-300), and emitted as a tree of "stars":
//
//   FK star      — enumerate the rows of a latent table (+ the new-row branch)
//   choice star  — enumerate the options of a discrete choice of a not-yet-existing row
//
// each with a list of likelihood terms.  The device evaluates stars bottom-up per row; no
// closures, no dictionaries.  Shapes the three benchmark programs do not need yet are
// rejected loudly (PCLEAN_ERR_UNSUPPORTED) instead of being silently approximated.
#pragma once
#include <cmath>
#include <functional>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pclean_b200.h"

namespace pcl {

struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };
struct BadArg : std::runtime_error { using std::runtime_error::runtime_error; };

typedef pclean_value Val;

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
  double py_strength = 1.0, py_discount = 0.0;
  int n_incoming = 0;
};
struct FuncM { int kind; Val cst; std::vector<int> keyargs; std::map<std::vector<int>, Val> table; };
struct Model {
  std::vector<ClassM> classes;
  std::vector<FuncM> funcs;
  std::vector<int> param_kind, param_indexed, slot_param;
  std::vector<double> param_prior0, param_prior1;
  std::vector<std::vector<Val>> lists;
  std::vector<double> xform_scale;
  std::vector<std::u32string> strings;
  double lm_uni[28], lm_big[28 * 28];
};

inline Plan parse_plan(const int32_t* pv, const int32_t* pn, int& pos, int nchild) {
  Plan out;
  for (int c = 0; c < nchild; ++c) {
    PlanNode n; n.v = pv[pos]; int k = pn[pos]; ++pos;
    n.kids = parse_plan(pv, pn, pos, k);
    out.push_back(std::move(n));
  }
  return out;
}

inline void parse_model(const pclean_model_ir* ir, Model& m) {
  m.classes.assign(ir->n_classes, ClassM());
  for (int c = 0; c < ir->n_classes; ++c) {
    ClassM& cm = m.classes[c];
    const int v0 = ir->class_voff[c], v1 = ir->class_voff[c + 1];
    cm.nv = v1 - v0; cm.py_strength = ir->py_strength[c]; cm.py_discount = ir->py_discount[c];
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
      const int nroots = ir->plan_nchild[pos]; ++pos;
      cm.plans.push_back(parse_plan(ir->plan_vertex, ir->plan_nchild, pos, nroots));
    }
    for (int k = ir->class_hash_off[c]; k < ir->class_hash_off[c + 1]; ++k) cm.hash_keys.push_back(ir->hash_v[k]);
  }
  for (int p = 0; p < ir->n_paths; ++p) m.classes[ir->path_target[p]].n_incoming += 1;
  m.funcs.resize(ir->n_funcs);
  for (int f = 0; f < ir->n_funcs; ++f) {
    FuncM& fm = m.funcs[f];
    fm.kind = ir->func_kind[f]; fm.cst = ir->func_const[f];
    for (int k = ir->func_keyarg_off[f]; k < ir->func_keyarg_off[f + 1]; ++k) fm.keyargs.push_back(ir->func_keyargs[k]);
    for (int e = ir->func_tab_off[f]; e < ir->func_tab_off[f + 1]; ++e)
      fm.table[std::vector<int>(ir->tab_keys + ir->tab_key_off[e], ir->tab_keys + ir->tab_key_off[e + 1])] = ir->tab_vals[e];
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
  m.strings.clear();
  for (int s = 0; s < ir->n_strings; ++s)
    m.strings.emplace_back(ir->str_cp + ir->str_off[s], ir->str_cp + ir->str_off[s + 1]);
}

// ------------------------------------------------------------------------------------------
// symbolic values
// ------------------------------------------------------------------------------------------
enum SymKind { S_NONE = 0, S_CONST, S_OBS, S_EARLIER, S_CAND, S_OPT, S_KEYOF, S_JOIN_EARLIER_CAND, S_JOIN_EARLIER_OPT, S_REFROW, S_JOIN_GENERIC, S_INNER, S_LOOKUP };
// operand of a generic join (latent-class moves): the enumerated element (column of the candidate / the option) or a cell of the referring row
enum { OP_ELEM_COL = 0, OP_ELEM_OPT = 1, OP_REFROW = 2 };
// argument of a tabulated-function lookup / Gaussian term evaluated per element
enum { ARG_CONST = 0, ARG_OBS = 1, ARG_ELEM_COL = 2, ARG_ELEM_OPT = 3, ARG_INNER = 4, ARG_REFROW = 5 };
struct ArgL { int kind = ARG_CONST; int ref = -1; };   // ref: value id | obs vertex (latent mode: the row's own cell) | column | - | inner choice index | referring-class vertex
struct Sym {
  int kind = S_NONE;
  Val cst{};          // S_CONST
  int vertex = -1;    // S_OBS / S_EARLIER: vertex in the observation class
  int star = -1;      // S_CAND / S_OPT / S_KEYOF: star id
  int col = -1;       // S_CAND: column (vertex) of the star's table
  int sep = -1;       // joins: separator string id
  int a_vertex = -1;  // joins: the earlier-block vertex supplying the left operand
  // S_JOIN_GENERIC: operands a ++ sep ++ b
  int a_kind = -1, a_ref = -1, b_kind = -1, b_ref = -1;   // ref = column (OP_ELEM_COL) or referring-class vertex (OP_REFROW)
  int inner = -1;                 // S_INNER: index of the inner choice
  int func = -1;                  // S_LOOKUP: tabulated function
  std::vector<ArgL> largs;        // S_LOOKUP: its key arguments
};

enum { ST_FK = 0, ST_CHOICE = 1 };
enum { TERM_CAND = 0, TERM_OPT = 1, TERM_JOIN_CAND = 2, TERM_JOIN_OPT = 3, TERM_JOIN_INLINE = 4, TERM_EQ = 5, TERM_GAUSS_EXT = 6, TERM_MSWAP_EXT = 7 };
// which part of a star a scope refers to
enum { SCOPE_ELEMS = 0, SCOPE_NEW = 1 };
enum { PRIOR_STATIC = 0, PRIOR_PROPORTIONS = 1 };

struct TermL {
  int obs_vertex;     // observed AddTypos leaf (vertex in the observation class)
  int kind;           // TERM_*
  int star;           // star whose element supplies the clean string
  int col;            // TERM_CAND / TERM_JOIN_CAND: column of star's table
  int sep = -1, a_vertex = -1;   // joins
  int max_typos = -1;
  bool external = false;         // summed over the rows referring to the latent row (ExternalLikelihoodNode)
  int a_kind = -1, a_ref = -1, b_kind = -1, b_ref = -1;   // TERM_JOIN_INLINE operands
  int gauss = -1;                // TERM_GAUSS_EXT: index into BlockProgram::gauss_ext
  int mswap = -1;                // TERM_MSWAP_EXT: index into BlockProgram::mswaps
};
struct InnerChoiceL { int vertex; int dist; int list; bool observed; int obs_vertex; };   // ChooseUniformly over a constant list
struct GaussL { int obs_vertex; ArgL mean_args[4]; int n_mean_args = 0; int mean_func = -1; double mean_const = 0; double stdev = 1; ArgL xform; };
// MaybeSwap(val, options, prob) (maybe_swap.jl:13-28): where its three arguments come from
struct LookupL { int func = -1; std::vector<ArgL> args; };                 // tabulated function of observed / referring-row values
struct MswapL {
  int obs_vertex = -1;          // the MaybeSwap node (observation-class vertex)
  int val_kind = 0;             // 0: cell of an earlier block's chosen row (val_vertex) | 1: the enumerated option (latent moves)
  int val_vertex = -1;
  int list_const = -1; LookupL list;      // options: constant list id, or a lookup returning a list
  int prob_kind = 0;            // 0 constant | 1 parameter slot | 2 lookup returning a parameter slot or a constant
  double prob_const = 0.0; int prob_slot = -1; LookupL prob;
};
struct ConstPriorL { int kind; int obs_vertex; int list; int slot; double value; };  // 0 constant, 1 log p[obs] of a proportions parameter, 2 StringPrior(min = list, max = slot) log-density of the observed string
struct InnerL {                  // per-element enumeration of small dependent choices + their likelihood terms
  std::vector<InnerChoiceL> choices;
  std::vector<GaussL> gauss;
  std::vector<ConstPriorL> consts;
  bool empty() const { return choices.empty() && gauss.empty() && consts.empty(); }
};
struct StarL {
  int kind;
  int vertex;         // vertex of the observation class this star assigns
  int parent = -1;    // star whose new-row branch contains this one (-1: block root)
  int level = 0;
  int table = -1;     // ST_FK: latent class enumerated
  int tvertex = -1;   // vertex in parent's table class that this star assigns (-1 for root)
  std::vector<int> terms;     // indices into BlockProgram::terms
  std::vector<int> children;  // stars evaluated in the new-row branch (ST_FK)
  // ST_CHOICE
  int dist = -1;
  int list = -1;              // option list id (string options)
  bool has_dummy = false;
  int dummy_string = -1;      // placeholder string id (interned by the engine)
  int prior_kind = PRIOR_STATIC;
  int prior_slot = -1;        // PRIOR_PROPORTIONS: parameter slot
  std::vector<double> static_prior;   // length n_options (+1 with dummy)
  int sp_min = 0, sp_max = 0;
  // (vertex in obs class, column in this star's table) pairs copied when an existing row is chosen
  std::vector<std::pair<int, int>> copies;
  // @guaranteed hash-bucket enumeration (proposal_compiler.jl:142-151)
  bool bucket = false; int bucket_col = -1, bucket_obs_vertex = -1;
  // option list that depends on the row (e.g. possibilities[countykey]): tabulated function + its key argument
  int list_func = -1; ArgL list_arg;
  InnerL inner_elems, inner_new;     // nested enumerations inside each element / inside the new-row branch
  // cells of a new row that no observation informs: sampled from the choice's discrete proposal in
  // propose_non_enumerable! (block_proposal.jl:42-60) when the new-row branch is taken
  struct FillL { int vertex; int dist; int list = -1; int list_func = -1; ArgL list_arg; int dummy_string = -1; };
  std::vector<FillL> fillins;
};
struct BlockProgram {
  int cls, block;
  int root = -1;                // the (single) root of an observation-class block
  std::vector<int> roots;       // latent-class blocks: one root per independent site
  bool latent = false;
  std::vector<StarL> stars;     // index = star id; children precede parents is NOT required
  std::vector<TermL> terms;
  std::vector<GaussL> gauss_ext;    // Gaussian likelihoods of the referring rows (TERM_GAUSS_EXT)
  std::vector<MswapL> mswaps;       // MaybeSwap terms: external (indexed by TermL::mswap) or of a rootless block
  bool rootless = false;            // observation-class block without any enumeration: only likelihood terms of earlier choices
  std::vector<int> root_terms;      // rootless block: observed MaybeSwap nodes (indices into mswaps)
  std::vector<int> root_sampled;    // rootless block: absent MaybeSwap nodes, sampled with random() (block_proposal.jl:58-60)
  std::vector<int> order;       // post-order evaluation (root last)
  std::set<int> earlier_vertices;   // particle-dependent inputs
};

// ------------------------------------------------------------------------------------------
// the symbolic walk
// ------------------------------------------------------------------------------------------
struct Lowerer {
  const Model& m;
  int cls;
  const ClassM& cm;
  std::vector<char> obs;          // vertex observed by the dataset
  std::vector<char> earlier;      // vertex assigned by an earlier block (or a parameter)
  std::vector<Sym> bound;
  std::vector<char> is_bound;
  std::map<int, int> active_child;    // fk vertex -> star
  BlockProgram prog;
  int scope_star = -1;            // star whose scope we are in
  bool scope_new = false;         // inside the new-row branch of scope_star (ST_FK)
  bool latent = false;            // lowering a latent class (External nodes allowed, several roots)
  int data_cls = -1;              // the observed class (source of every supported External path)
  const std::vector<char>* data_obs = nullptr;   // dataset columns of the observed class
  const pclean_model_ir* ir = nullptr;
  bool in_external = false;
  int ext_path = -1;
  InnerL* inner_scope() {
    if (scope_star < 0) return nullptr;
    StarL& st = prog.stars[scope_star];
    return (scope_new && st.kind == ST_FK) ? &st.inner_new : &st.inner_elems;
  }
  ArgL arg_of(const Sym& x) {
    ArgL a;
    switch (x.kind) {
      case S_CONST: a.kind = ARG_CONST; a.ref = x.cst.i; return a;
      case S_OBS: a.kind = ARG_OBS; a.ref = x.vertex; return a;
      case S_CAND: if (x.star != scope_star) break; a.kind = ARG_ELEM_COL; a.ref = x.col; return a;
      case S_OPT: if (x.star != scope_star) break; a.kind = ARG_ELEM_OPT; return a;
      case S_INNER: a.kind = ARG_INNER; a.ref = x.inner; return a;
      case S_REFROW: a.kind = ARG_REFROW; a.ref = x.vertex; return a;
      default: break;
    }
    throw Unsupported("tabulated function argument that is neither constant, observed, the enumerated value nor an inner choice");
  }
  std::map<int, Sym> recomputed;  // referring-class vertex -> symbolic value inside the external loop
  std::function<int(const std::u32string&)> intern;

  Lowerer(const Model& model, int c) : m(model), cls(c), cm(model.classes[c]) {}

  bool avail(int k) const { return obs[k] || earlier[k] || is_bound[k]; }
  bool any_unavailable(const std::vector<int>& a) const { for (int k : a) if (!avail(k)) return true; return false; }
  Sym value(int k) const {
    if (is_bound[k]) return bound[k];
    Sym s;
    if (obs[k]) { s.kind = S_OBS; s.vertex = k; return s; }
    const Node& n = cm.nodes[k];
    if (n.wrap == PCLEAN_WRAP_NONE && n.kind == PCLEAN_NODE_PARAM) { s.kind = S_CONST; s.cst.tag = PCLEAN_VAL_PARAM; s.cst.i = -1; return s; }
    s.kind = S_EARLIER; s.vertex = k; return s;
  }
  static bool has_discrete_proposal(int dist) {
    return dist == PCLEAN_DIST_CHOOSE_PROPORTIONALLY || dist == PCLEAN_DIST_CHOOSE_UNIFORMLY ||
           dist == PCLEAN_DIST_STRING_PRIOR || dist == PCLEAN_DIST_TIME_PRIOR;
  }

  void walk(const Plan& steps) { for (const PlanNode& s : steps) step(s); }
  void step(const PlanNode& s) {
    const Node& n = cm.nodes[s.v];
    if (n.wrap == PCLEAN_WRAP_EXTERNAL) return external(n, s.v, s.kids);
    if (n.wrap == PCLEAN_WRAP_SUBMODEL) return submodel(n, 0, s.v, s.kids);
    base(n, s.v, s.kids);
  }
  std::vector<std::pair<int, StarL::FillL>> fill_todo;
  bool node_wrapped = false;      // the node being dispatched is the base of a SubmodelNode (a cell of a referenced row)
  bool bucket_pending = false;
  void base(const Node& n, int idx, const Plan& rest) {
    switch (n.kind) {
      case PCLEAN_NODE_JULIA: return julia(n, idx, rest);
      case PCLEAN_NODE_CHOICE: return choice(n, idx, rest);
      case PCLEAN_NODE_FK: return foreign_key(n, idx, rest);
      default: return walk(rest);
    }
  }
  void julia(const Node& n, int idx, const Plan& rest) {
    node_wrapped = false;
    if (any_unavailable(n.args)) return walk(rest);
    const FuncM& f = m.funcs[n.func];
    Sym out;
    if (f.kind == PCLEAN_FUNC_CONST) { out.kind = S_CONST; out.cst = f.cst; }
    else if (f.kind == PCLEAN_FUNC_JOIN) {
      Sym a = value(n.args.at(0)), b = value(n.args.at(1));
      if (a.kind == S_EARLIER && b.kind == S_CAND) { out.kind = S_JOIN_EARLIER_CAND; out.star = b.star; out.col = b.col; }
      else if (a.kind == S_EARLIER && b.kind == S_OPT) { out.kind = S_JOIN_EARLIER_OPT; out.star = b.star; }
      else throw Unsupported("string join with operands other than (earlier-block value, enumerated value)");
      out.a_vertex = a.vertex; out.sep = f.cst.i;
    } else if (f.kind == PCLEAN_FUNC_TABLE) {
      std::vector<int> key; bool all_const = true;
      for (int pos : f.keyargs) { Sym a = value(n.args.at(pos)); if (a.kind != S_CONST) all_const = false; else key.push_back(a.cst.i); }
      if (all_const) {
        auto it = f.table.find(key);
        if (it == f.table.end()) throw BadArg("tabulated JuliaNode: constant argument outside its support");
        out.kind = S_CONST; out.cst = it->second;
      } else {
        out.kind = S_LOOKUP; out.func = n.func; out.star = scope_star;
        for (int pos : f.keyargs) out.largs.push_back(arg_of(value(n.args.at(pos))));
      }
    } else if (f.kind == PCLEAN_FUNC_ROUND_BACKWARD) {
      out.kind = S_NONE;       // output-only node (corrected = round(unit.backward(rent))): evaluated at download time
    } else throw Unsupported("JuliaNode builtin not supported on a scoring path");
    bound[idx] = out; is_bound[idx] = 1;
    walk(rest);
    is_bound[idx] = 0;
  }
  int new_star(int kind, int vertex) {
    StarL s; s.kind = kind; s.vertex = vertex; s.parent = scope_star;
    s.level = scope_star < 0 ? 0 : prog.stars[scope_star].level + 1;
    prog.stars.push_back(s);
    const int id = (int)prog.stars.size() - 1;
    if (scope_star >= 0) prog.stars[scope_star].children.push_back(id);
    else {
      if (prog.root >= 0 && !latent) throw Unsupported("block plan with more than one enumeration root");
      if (prog.root < 0) prog.root = id;
      prog.roots.push_back(id);
    }
    return id;
  }
  void add_term(int obs_vertex, const Sym& clean, int max_typos) {
    if (scope_star < 0) throw Unsupported("likelihood term outside any enumeration");
    TermL t; t.obs_vertex = obs_vertex; t.max_typos = max_typos; t.star = clean.star; t.col = clean.col;
    t.sep = clean.sep; t.a_vertex = clean.a_vertex;
    switch (clean.kind) {
      case S_CAND: t.kind = TERM_CAND; break;
      case S_OPT: t.kind = TERM_OPT; break;
      case S_JOIN_EARLIER_CAND: t.kind = TERM_JOIN_CAND; prog.earlier_vertices.insert(clean.a_vertex); break;
      case S_JOIN_EARLIER_OPT: t.kind = TERM_JOIN_OPT; prog.earlier_vertices.insert(clean.a_vertex); break;
      default: throw Unsupported("AddTypos whose clean argument does not depend on the enumerated value");
    }
    if (t.star != scope_star) throw Unsupported("likelihood term depends on an outer enumeration variable (nested dependent enumeration)");
    const bool in_new = scope_new && prog.stars[scope_star].kind == ST_FK;
    if (in_new) throw Unsupported("likelihood term directly inside a new-row branch");
    prog.terms.push_back(t);
    prog.stars[scope_star].terms.push_back((int)prog.terms.size() - 1);
  }
  void choice(const Node& n, int idx, const Plan& rest) {
    const bool wrapped_here = node_wrapped; node_wrapped = false;
    const bool observed = obs[idx] || earlier[idx];
    if (!observed && n.dist == PCLEAN_DIST_MAYBE_SWAP && !latent && scope_star < 0 && !any_unavailable(n.args)) {
      // absent observation: the reference samples it with random() (no weight) and keeps it in the row
      prog.mswaps.push_back(mswap_of(n, idx));
      prog.root_sampled.push_back((int)prog.mswaps.size() - 1);
      return walk(rest);
    }
    if (!observed && !has_discrete_proposal(n.dist)) return walk(rest);
    if (any_unavailable(n.args)) return walk(rest);
    if (observed) {
      walk(rest);
      if (n.dist == PCLEAN_DIST_ADD_TYPOS) {
        if (!obs[idx]) throw Unsupported("AddTypos leaf that is not a dataset column");
        int max_typos = -1;
        if (n.args.size() > 1) {
          Sym mt = value(n.args[1]);
          if (mt.kind != S_CONST) throw Unsupported("non-constant max_typos");
          max_typos = mt.cst.tag == PCLEAN_VAL_INT ? mt.cst.i : (int)mt.cst.d;
        }
        add_term(idx, value(n.args.at(0)), max_typos);
        return;
      }
      InnerL* in = inner_scope();
      if (n.dist == PCLEAN_DIST_UNMODELED) return;                       // log-density 0
      if (latent && !in && n.dist != PCLEAN_DIST_TRANSFORMED_GAUSSIAN && n.dist != PCLEAN_DIST_ADD_NOISE && n.dist != PCLEAN_DIST_MAYBE_SWAP)
        return;      // observed cell of a latent row outside any enumeration: the same factor for every particle
      if (n.dist == PCLEAN_DIST_MAYBE_SWAP && !latent && scope_star < 0) {
        if (!obs[idx]) throw Unsupported("MaybeSwap leaf that is not a dataset column");
        prog.mswaps.push_back(mswap_of(n, idx));
        prog.root_terms.push_back((int)prog.mswaps.size() - 1);
        return;
      }
      if (!in) throw Unsupported("observed choice outside any enumeration");
      if (n.dist == PCLEAN_DIST_CHOOSE_UNIFORMLY) {
        Sym l = value(n.args.at(0));
        if (l.kind != S_CONST) throw Unsupported("ChooseUniformly over a non-constant list");
        ConstPriorL c{0, idx, l.cst.i, -1, -std::log((double)m.lists.at(l.cst.i).size())};
        in->consts.push_back(c);
        return;
      }
      if (n.dist == PCLEAN_DIST_CHOOSE_PROPORTIONALLY) {
        Sym l = value(n.args.at(0));
        if (l.kind != S_CONST || !obs[idx]) throw Unsupported("observed ChooseProportionally with non-constant options");
        ConstPriorL c{1, idx, l.cst.i, param_slot_of(n.args.at(1)), 0.0};
        in->consts.push_back(c);
        return;
      }
      if (n.dist == PCLEAN_DIST_TRANSFORMED_GAUSSIAN) {
        if (!obs[idx]) throw Unsupported("Gaussian leaf that is not a dataset column");
        GaussL g; g.obs_vertex = idx;
        Sym mu = value(n.args.at(0)), sd = value(n.args.at(1)), xf = value(n.args.at(2));
        if (sd.kind != S_CONST) throw Unsupported("non-constant standard deviation");
        g.stdev = sd.cst.tag == PCLEAN_VAL_REAL ? sd.cst.d : (double)sd.cst.i;
        if (mu.kind == S_LOOKUP) { g.mean_func = mu.func; g.n_mean_args = (int)mu.largs.size(); if (g.n_mean_args > 4) throw Unsupported("lookup with more than 4 key arguments"); for (int i = 0; i < g.n_mean_args; ++i) g.mean_args[i] = mu.largs[i]; }
        else if (mu.kind == S_CONST && mu.cst.tag == PCLEAN_VAL_REAL) g.mean_const = mu.cst.d;
        else if (mu.kind == S_CONST && mu.cst.tag == PCLEAN_VAL_PARAM) { g.mean_func = -2; g.mean_const = (double)mu.cst.i; }
        else throw Unsupported("Gaussian mean that is neither a constant nor a tabulated parameter");
        g.xform = arg_of(xf);
        in->gauss.push_back(g);
        return;
      }
      if (n.dist == PCLEAN_DIST_STRING_PRIOR) {
        // observed string in a new-row branch (flights flight_id): StringPrior.logdensity of the observed value
        if (!obs[idx]) throw Unsupported("StringPrior value fixed by an earlier block");
        Sym a0 = value(n.args.at(0)), a1 = value(n.args.at(1));
        if (a0.kind != S_CONST || a1.kind != S_CONST) throw Unsupported("StringPrior with non-constant length bounds");
        ConstPriorL c{2, idx, a0.cst.i, a1.cst.i, 0.0};      // list / slot fields carry (min, max)
        in->consts.push_back(c);
        return;
      }
      throw Unsupported("observed choice with a likelihood that is not lowered yet");
    }
    // unobserved with a discrete proposal: a choice star
    if (!((scope_star >= 0 && scope_new && prog.stars[scope_star].kind == ST_FK) || (latent && scope_star < 0)) ||
        (n.dist == PCLEAN_DIST_CHOOSE_UNIFORMLY && !latent && !wrapped_here)) {
      // a small dependent choice enumerated inside every element of the enclosing enumeration
      // (rents: br, unit): ChooseUniformly over a constant list
      InnerL* in = inner_scope();
      if (!in || n.dist != PCLEAN_DIST_CHOOSE_UNIFORMLY) throw Unsupported("discrete choice enumerated per element of another enumeration (only uniform choices over constant lists are lowered)");
      Sym l = value(n.args.at(0));
      if (l.kind != S_CONST) throw Unsupported("inner choice over a non-constant list");
      InnerChoiceL c{idx, n.dist, l.cst.i, false, -1};
      in->choices.push_back(c);
      Sym me; me.kind = S_INNER; me.inner = (int)in->choices.size() - 1;
      bound[idx] = me; is_bound[idx] = 1;
      walk(rest);
      is_bound[idx] = 0;
      return;
    }
    const int sid = new_star(ST_CHOICE, idx);
    StarL& s = prog.stars[sid];
    s.dist = n.dist;
    s.tvertex = prog.stars[sid].parent >= 0 ? tvertex_of(idx, prog.stars[sid].parent) : idx;
    auto const_arg = [&](int pos) { Sym a = value(n.args.at(pos)); if (a.kind != S_CONST) throw Unsupported("choice with non-constant arguments"); return a.cst; };
    if (n.dist == PCLEAN_DIST_CHOOSE_UNIFORMLY) {
      s.list = const_arg(0).i;
      s.static_prior.assign(m.lists.at(s.list).size(), -std::log((double)m.lists[s.list].size()));
    } else if (n.dist == PCLEAN_DIST_CHOOSE_PROPORTIONALLY) {
      s.list = const_arg(0).i;
      s.prior_kind = PRIOR_PROPORTIONS;
      s.prior_slot = param_slot_of(n.args.at(1));
    } else if (n.dist == PCLEAN_DIST_STRING_PRIOR) {
      s.sp_min = const_arg(0).i; s.sp_max = const_arg(1).i;
      Sym la = value(n.args.at(2));
      if (la.kind == S_CONST) s.list = la.cst.i;
      else if (la.kind == S_LOOKUP && la.largs.size() == 1 && la.largs[0].kind == ARG_OBS) { s.list = -1; s.list_func = la.func; s.list_arg = la.largs[0]; }
      else throw Unsupported("StringPrior atoms that are neither constant nor a lookup on an observed value");
      s.has_dummy = true;
      s.dummy_string = intern(std::u32string((size_t)((s.sp_min + s.sp_max) / 2), U'*'));
    } else if (n.dist == PCLEAN_DIST_TIME_PRIOR) {
      Sym la = value(n.args.at(0));
      if (la.kind == S_CONST) s.list = la.cst.i;
      else if (la.kind == S_LOOKUP && la.largs.size() == 1 && la.largs[0].kind == ARG_OBS) { s.list = -1; s.list_func = la.func; s.list_arg = la.largs[0]; }
      else throw Unsupported("TimePrior atoms that are neither constant nor a lookup on an observed value");
      s.has_dummy = true;
      std::string d = "**:** p.m.";
      s.dummy_string = intern(std::u32string(d.begin(), d.end()));
    }
    if (s.list >= 0) for (const Val& v : m.lists.at(s.list)) if (v.tag != PCLEAN_VAL_STR) throw Unsupported("choice over non-string options");
    Sym me; me.kind = S_OPT; me.star = sid;
    bound[idx] = me; is_bound[idx] = 1;
    const int save_star = scope_star; const bool save_new = scope_new;
    scope_star = sid; scope_new = false;
    walk(rest);
    scope_star = save_star; scope_new = save_new;
    is_bound[idx] = 0;
  }
  LookupL lookup_of(const Sym& x) const { LookupL l; l.func = x.func; l.args = x.largs; return l; }
  // arguments of a MaybeSwap node of the observation class (values resolved symbolically)
  MswapL mswap_of(const Node& n, int idx) {
    MswapL ms; ms.obs_vertex = idx;
    const Sym v = value(n.args.at(0)), l = value(n.args.at(1)), p = value(n.args.at(2));
    if (v.kind != S_EARLIER) throw Unsupported("MaybeSwap whose value is not a cell chosen by an earlier block");
    ms.val_kind = 0; ms.val_vertex = v.vertex;
    if (l.kind == S_CONST && l.cst.tag == PCLEAN_VAL_LIST) ms.list_const = l.cst.i;
    else if (l.kind == S_LOOKUP) ms.list = lookup_of(l);
    else throw Unsupported("MaybeSwap options that are neither constant nor a lookup");
    set_prob(ms, p);
    return ms;
  }
  void set_prob(MswapL& ms, const Sym& p) const {
    if (p.kind == S_CONST && p.cst.tag == PCLEAN_VAL_REAL) { ms.prob_kind = 0; ms.prob_const = p.cst.d; }
    else if (p.kind == S_CONST && p.cst.tag == PCLEAN_VAL_PARAM && p.cst.i >= 0) { ms.prob_kind = 1; ms.prob_slot = p.cst.i; }
    else if (p.kind == S_LOOKUP) { ms.prob_kind = 2; ms.prob = lookup_of(p); }
    else throw Unsupported("MaybeSwap probability that is neither a constant, a parameter nor a lookup");
  }
  int param_slot_of(int vertex) const {
    // parameter vertices resolve (through submodel wrappers) to the one basic slot of their spec
    const Node& n = cm.nodes[vertex];
    if (n.kind != PCLEAN_NODE_PARAM) throw Unsupported("ChooseProportionally with a literal probability vector");
    if (m.param_indexed[n.param]) throw Unsupported("indexed parameter as a proportions vector");
    for (size_t s = 0; s < m.slot_param.size(); ++s) if (m.slot_param[s] == n.param) return (int)s;
    throw BadArg("parameter has no slot");
  }
  // vertex of the parent's table class corresponding to obs-class vertex `idx`
  int tvertex_of(int idx, int parent_star) const {
    if (parent_star < 0) return -1;
    const StarL& ps = prog.stars[parent_star];
    const Node& fk = cm.nodes[ps.vertex];
    for (size_t tv = 0; tv < fk.vmap.size(); ++tv) if (fk.vmap[tv] == idx) return (int)tv;
    throw BadArg("vertex is not in the parent reference slot's vmap");
  }
  void foreign_key(const Node& n, int idx, const Plan& rest) {
    node_wrapped = false;
    const ClassM& tm = m.classes[n.target];
    if (!tm.hash_keys.empty()) {
      bool all = true;
      for (int h : tm.hash_keys) if (!(obs[n.vmap[h]] || earlier[n.vmap[h]])) all = false;
      if (all && tm.hash_keys.size() != 1) throw Unsupported("more than one @guaranteed key");
      if (all && !obs[n.vmap[tm.hash_keys[0]]]) throw Unsupported("@guaranteed key supplied by an earlier block");
      bucket_pending = all;
    }
    if (scope_star >= 0 && !(scope_new && prog.stars[scope_star].kind == ST_FK))
      throw Unsupported("reference slot enumerated per candidate of another enumeration");
    if (scope_star < 0 && latent && prog.root >= 0) { /* further independent site of a latent block */ }
    const int sid = new_star(ST_FK, idx);
    if (bucket_pending) { prog.stars[sid].bucket = true; prog.stars[sid].bucket_col = tm.hash_keys[0]; prog.stars[sid].bucket_obs_vertex = n.vmap[tm.hash_keys[0]]; bucket_pending = false; }
    prog.stars[sid].table = n.target;
    prog.stars[sid].tvertex = prog.stars[sid].parent >= 0 ? tvertex_of(idx, prog.stars[sid].parent) : idx;
    // copies: every obs-class vertex that is a submodel cell of this slot
    for (size_t tv = 0; tv < n.vmap.size(); ++tv) prog.stars[sid].copies.emplace_back(n.vmap[tv], (int)tv);
    const int save_star = scope_star; const bool save_new = scope_new;
    Sym key; key.kind = S_KEYOF; key.star = sid;
    bound[idx] = key; is_bound[idx] = 1;
    // existing candidates
    active_child[idx] = sid;
    scope_star = sid; scope_new = false;
    walk(rest);
    active_child.erase(idx);
    // new-row branch
    scope_new = true;
    walk(rest);
    scope_star = save_star; scope_new = save_new;
    is_bound[idx] = 0;
  }
  // ExternalLikelihoodNode (proposal_compiler.jl:306-350): the likelihood of the rows that
  // (transitively) refer to the latent row being moved.  Lowered to terms summed over those rows.
  Sym ext_value(int k) const {
    auto it = recomputed.find(k);
    if (it != recomputed.end()) return it->second;
    Sym s; s.kind = S_REFROW; s.vertex = k;
    // a JuliaNode of the referring class holds a function of other cells of that row: constants and
    // tabulated functions are re-expressed over those cells (the stored value is the same thing)
    const Node& dn = m.classes[data_cls].nodes[k];
    if (dn.wrap == PCLEAN_WRAP_NONE && dn.kind == PCLEAN_NODE_JULIA) {
      const FuncM& f = m.funcs[dn.func];
      if (f.kind == PCLEAN_FUNC_CONST) { Sym c; c.kind = S_CONST; c.cst = f.cst; return c; }
      if (f.kind == PCLEAN_FUNC_TABLE) {
        Sym out; out.kind = S_LOOKUP; out.func = dn.func; out.star = -1;
        for (int pos : f.keyargs) {
          const Sym a = ext_value(dn.args.at(pos));
          if (a.kind == S_OPT || a.kind == S_CAND) out.star = a.star;
          ArgL al;
          switch (a.kind) {
            case S_CONST: al.kind = ARG_CONST; al.ref = a.cst.i; break;
            case S_REFROW: al.kind = ARG_REFROW; al.ref = a.vertex; break;
            case S_OPT: al.kind = ARG_ELEM_OPT; break;
            case S_CAND: al.kind = ARG_ELEM_COL; al.ref = a.col; break;
            default: return s;
          }
          out.largs.push_back(al);
        }
        return out;
      }
    }
    return s;
  }
  void external(const Node& n, int idx, const Plan& rest) {
    if (!latent) throw Unsupported("external likelihood node in an observation class");
    if (in_external) {
      if (n.kind == PCLEAN_NODE_JULIA && m.funcs[n.func].kind == PCLEAN_FUNC_TABLE) {
        // recomputed tabulated function of the referring row (rents: rent_base = avg_rent[state_key_br])
        const FuncM& f = m.funcs[n.func];
        Sym out; out.kind = S_LOOKUP; out.func = n.func; out.star = -1;
        for (int pos : f.keyargs) {
          const Sym a = ext_value(n.args.at(pos));
          if ((a.kind == S_OPT || a.kind == S_CAND) && a.star != scope_star) throw Unsupported("external lookup over values of an outer enumeration");
          if (a.kind == S_OPT || a.kind == S_CAND) out.star = scope_star;
          out.largs.push_back(arg_of(a));
        }
        recomputed[n.extv] = out;
        walk(rest);
        recomputed.erase(n.extv);
        return;
      }
      if (n.kind == PCLEAN_NODE_JULIA) {
        const FuncM& f = m.funcs[n.func];
        if (f.kind != PCLEAN_FUNC_JOIN) throw Unsupported("external JuliaNode other than a string join or a tabulated function");
        Sym a = ext_value(n.args.at(0)), b = ext_value(n.args.at(1)), out;
        auto op = [&](const Sym& x, int& kind, int& ref) {
          if (x.kind == S_REFROW) { kind = OP_REFROW; ref = x.vertex; }
          else if (x.kind == S_OPT && x.star == scope_star) { kind = OP_ELEM_OPT; ref = -1; }
          else if (x.kind == S_CAND && x.star == scope_star) { kind = OP_ELEM_COL; ref = x.col; }
          else throw Unsupported("string join over values of an outer enumeration");
        };
        out.kind = S_JOIN_GENERIC; out.sep = f.cst.i; out.star = scope_star;
        op(a, out.a_kind, out.a_ref); op(b, out.b_kind, out.b_ref);
        recomputed[n.extv] = out;
        walk(rest);
        recomputed.erase(n.extv);
        return;
      }
      if (n.kind == PCLEAN_NODE_CHOICE && n.dist == PCLEAN_DIST_MAYBE_SWAP) {
        walk(rest);
        const Sym v = ext_value(n.args.at(0)), l = ext_value(n.args.at(1)), pr = ext_value(n.args.at(2));
        const bool on_elem = (v.kind == S_OPT || v.kind == S_CAND) && v.star == scope_star && scope_star >= 0;
        const bool prob_on_elem = pr.kind == S_LOOKUP && pr.star >= 0;
        if (!on_elem && !prob_on_elem) return;        // same factor for every option
        if (!on_elem) throw Unsupported("external MaybeSwap whose probability (not its value) depends on the enumerated value");
        if (v.kind != S_OPT) throw Unsupported("external MaybeSwap over a candidate's column");
        if (scope_new && prog.stars[scope_star].kind == ST_FK) throw Unsupported("external likelihood directly inside a new-row branch");
        MswapL ms; ms.obs_vertex = n.extv; ms.val_kind = 1;
        if (l.kind == S_CONST && l.cst.tag == PCLEAN_VAL_LIST) ms.list_const = l.cst.i;
        else if (l.kind == S_LOOKUP) ms.list = lookup_of(l);
        else throw Unsupported("MaybeSwap options that are neither constant nor a lookup");
        set_prob(ms, pr);
        TermL t; t.obs_vertex = n.extv; t.kind = TERM_MSWAP_EXT; t.external = true; t.star = scope_star; t.col = -1;
        t.mswap = (int)prog.mswaps.size();
        prog.mswaps.push_back(ms);
        prog.terms.push_back(t);
        prog.stars[scope_star].terms.push_back((int)prog.terms.size() - 1);
        return;
      }
      if (n.kind == PCLEAN_NODE_CHOICE && n.dist == PCLEAN_DIST_TRANSFORMED_GAUSSIAN) {
        walk(rest);
        if (!(*data_obs)[n.extv]) throw Unsupported("external Gaussian leaf that is not a dataset column");
        const Sym mu = ext_value(n.args.at(0));
        if (mu.kind != S_LOOKUP) { if (mu.kind == S_REFROW || mu.kind == S_CONST) return; throw Unsupported("external Gaussian whose mean is not a tabulated parameter"); }
        if (mu.star < 0) return;                     // does not depend on the enumerated value: same for every option
        if (mu.star != scope_star || (scope_new && prog.stars[scope_star].kind == ST_FK)) throw Unsupported("external Gaussian depending on an outer enumeration variable");
        auto ext_const = [&](int v) -> Val {
          const Node& an = m.classes[data_cls].nodes[v];
          if (an.kind == PCLEAN_NODE_JULIA && an.wrap == PCLEAN_WRAP_NONE && m.funcs[an.func].kind == PCLEAN_FUNC_CONST) return m.funcs[an.func].cst;
          throw Unsupported("external Gaussian with a non-constant standard deviation");
        };
        GaussL g; g.obs_vertex = n.extv;
        g.mean_func = mu.func; g.n_mean_args = (int)mu.largs.size();
        if (g.n_mean_args > 3) throw Unsupported("lookup with more than 3 key arguments");
        for (int i = 0; i < g.n_mean_args; ++i) g.mean_args[i] = mu.largs[i];
        const Val sd = ext_const(n.args.at(1));
        g.stdev = sd.tag == PCLEAN_VAL_REAL ? sd.d : (double)sd.i;
        const Sym xf = ext_value(n.args.at(2));
        if (xf.kind == S_REFROW) {
          const Node& xn = m.classes[data_cls].nodes[xf.vertex];
          if (xn.kind == PCLEAN_NODE_JULIA && xn.wrap == PCLEAN_WRAP_NONE && m.funcs[xn.func].kind == PCLEAN_FUNC_CONST) { g.xform.kind = ARG_CONST; g.xform.ref = m.funcs[xn.func].cst.i; }
          else g.xform = arg_of(xf);
        } else g.xform = arg_of(xf);
        TermL t; t.obs_vertex = n.extv; t.kind = TERM_GAUSS_EXT; t.external = true; t.star = scope_star; t.col = -1;
        t.gauss = (int)prog.gauss_ext.size();
        prog.gauss_ext.push_back(g);
        prog.terms.push_back(t);
        prog.stars[scope_star].terms.push_back((int)prog.terms.size() - 1);
        return;
      }
      if (n.kind == PCLEAN_NODE_CHOICE) {
        walk(rest);
        if (n.dist != PCLEAN_DIST_ADD_TYPOS) throw Unsupported("external likelihood other than AddTypos / TransformedGaussian (flights shapes) is not lowered yet");
        if (!(*data_obs)[n.extv]) throw Unsupported("external AddTypos leaf that is not a dataset column");
        int max_typos = -1;
        if (n.args.size() > 1) {
          Sym mt = ext_value(n.args[1]);
          // literal arguments of the referring class are zero-argument JuliaNodes stored in its rows
          const Node& an = m.classes[data_cls].nodes[n.args[1]];
          if (an.kind == PCLEAN_NODE_JULIA && m.funcs[an.func].kind == PCLEAN_FUNC_CONST) {
            const Val& c = m.funcs[an.func].cst; max_typos = c.tag == PCLEAN_VAL_INT ? c.i : (int)c.d;
          } else throw Unsupported("non-constant max_typos");
          (void)mt;
        }
        Sym clean = ext_value(n.args.at(0));
        if (scope_star < 0) throw Unsupported("external likelihood outside any enumeration");
        const bool in_new = scope_new && prog.stars[scope_star].kind == ST_FK;
        if (in_new) throw Unsupported("external likelihood directly inside a new-row branch");
        TermL t; t.obs_vertex = n.extv; t.max_typos = max_typos; t.external = true; t.star = scope_star; t.col = clean.col; t.sep = clean.sep;
        if (clean.kind == S_CAND && clean.star == scope_star) t.kind = TERM_CAND;
        else if (clean.kind == S_OPT && clean.star == scope_star) t.kind = TERM_OPT;
        else if (clean.kind == S_JOIN_GENERIC && clean.star == scope_star) {
          t.kind = TERM_JOIN_INLINE; t.a_kind = clean.a_kind; t.a_ref = clean.a_ref; t.b_kind = clean.b_kind; t.b_ref = clean.b_ref;
        } else if (clean.kind == S_REFROW) return;   // does not depend on the enumerated value: same for every option
        else throw Unsupported("external likelihood depending on an outer enumeration variable");
        prog.terms.push_back(t);
        prog.stars[scope_star].terms.push_back((int)prog.terms.size() - 1);
        return;
      }
      throw Unsupported("ExternalLikelihoodNode{ForeignKeyNode}");
    }
    // open the loop over referring rows
    const int p = n.path;
    const int n_links = ir->path_len_off[p + 1] - ir->path_len_off[p];
    const int src = ir->path_class[ir->path_len_off[p] + n_links - 1];
    if (src != data_cls) throw Unsupported("external likelihood whose source is not the observed class");
    in_external = true; ext_path = p;
    recomputed.clear();
    const int v0 = ir->path_vmap_off[p], v1 = ir->path_vmap_off[p + 1];
    for (int tv = 0; tv < v1 - v0; ++tv) {
      const int j = ir->path_vmap[v0 + tv];
      if (j >= 0 && tv < (int)is_bound.size() && is_bound[tv]) recomputed[j] = bound[tv];
    }
    external(n, idx, rest);
    in_external = false; ext_path = -1;
    recomputed.clear();
  }

  bool can_process_base(const Node& n, int idx) const {
    if (n.kind == PCLEAN_NODE_JULIA) return !any_unavailable(n.args);
    if (n.kind == PCLEAN_NODE_CHOICE) return !any_unavailable(n.args) && (obs[idx] || earlier[idx] || has_discrete_proposal(n.dist));
    return n.kind == PCLEAN_NODE_FK;
  }
  void submodel(const Node& n, size_t level, int idx, const Plan& rest) {
    if (level >= n.wfk.size()) { node_wrapped = true; base(n, idx, rest); node_wrapped = false; return; }
    if (!(obs[idx] || earlier[idx] || can_process_base(n, idx))) return walk(rest);
    auto it = active_child.find(n.wfk[level]);
    if (it == active_child.end()) return submodel(n, level + 1, idx, rest);
    if (obs[idx] || earlier[idx]) {
      // case 2 (proposal_compiler.jl:277-292): the candidate must agree with the observed cell
      if (!obs[idx]) throw Unsupported("equality constraint against an earlier-block value");
      if (it->second != scope_star || scope_new) throw Unsupported("equality constraint outside the candidate's own scope");
      TermL t; t.obs_vertex = idx; t.kind = TERM_EQ; t.star = scope_star; t.col = n.wsub[level];
      prog.terms.push_back(t);
      prog.stars[scope_star].terms.push_back((int)prog.terms.size() - 1);
      walk(rest);
      return;
    }
    Sym s; s.kind = S_CAND; s.star = it->second; s.col = n.wsub[level];
    bound[idx] = s; is_bound[idx] = 1;
    walk(rest);
    is_bound[idx] = 0;
  }

  static Plan prune(const Plan& plan, const std::vector<char>& has, const ClassM& cm) {
    Plan out;
    for (const PlanNode& s : plan) {
      Plan sub = prune(s.kids, has, cm);
      if (!sub.empty()) { PlanNode n; n.v = s.v; n.kids = std::move(sub); out.push_back(std::move(n)); }
      else if (has[s.v] || cm.nodes[s.v].wrap == PCLEAN_WRAP_EXTERNAL) { PlanNode n; n.v = s.v; out.push_back(std::move(n)); }
    }
    return out;
  }

  void postorder(int s) {
    for (int c : prog.stars[s].children) postorder(c);
    prog.order.push_back(s);
  }

  BlockProgram lower_block(int block, const std::vector<char>& observed_vertices) {
    prog = BlockProgram(); prog.cls = cls; prog.block = block;
    obs = observed_vertices;
    earlier.assign(cm.nv, 0);
    for (int v = 0; v < cm.nv; ++v) if (cm.nodes[v].kind == PCLEAN_NODE_PARAM) earlier[v] = 1;   // fill_parameters!
    for (int b = 0; b < block; ++b) for (int v : cm.blocks[b]) if (!obs[v]) earlier[v] = 1;
    bound.assign(cm.nv, Sym()); is_bound.assign(cm.nv, 0);
    std::vector<char> has(cm.nv);
    for (int v = 0; v < cm.nv; ++v) has[v] = obs[v] || earlier[v];
    // JuliaNodes of earlier blocks that are functions of observed cells only (flights error_prob):
    // their value is a row constant, re-expressed symbolically instead of read from the particle
    auto bind_row_function = [&](int v) {
      const Node& en = cm.nodes[v];
      if (en.wrap != PCLEAN_WRAP_NONE || en.kind != PCLEAN_NODE_JULIA || obs[v] || is_bound[v]) return;
      const FuncM& f = m.funcs[en.func];
      bool ok = f.kind == PCLEAN_FUNC_CONST || f.kind == PCLEAN_FUNC_TABLE;
      if (f.kind == PCLEAN_FUNC_TABLE) for (int pos : f.keyargs) { const int a = en.args.at(pos); ok = ok && (obs[a] || is_bound[a]) ; }
      if (!ok) return;
      Sym out;
      if (f.kind == PCLEAN_FUNC_CONST) { out.kind = S_CONST; out.cst = f.cst; }
      else {
        out.kind = S_LOOKUP; out.func = en.func; out.star = -1;
        bool fine = true;
        for (int pos : f.keyargs) {
          const Sym a = value(en.args.at(pos));
          if (a.kind == S_OBS) { ArgL al; al.kind = ARG_OBS; al.ref = a.vertex; out.largs.push_back(al); }
          else if (a.kind == S_CONST) { ArgL al; al.kind = ARG_CONST; al.ref = a.cst.i; out.largs.push_back(al); }
          else fine = false;
        }
        if (!fine) return;
      }
      bound[v] = out; is_bound[v] = 1;
    };
    for (int b = 0; b < block; ++b) for (int v : cm.blocks[b]) bind_row_function(v);
    Plan pruned = prune(cm.plans[block], has, cm);
    if (pruned.empty() && latent) throw Unsupported("block with nothing to enumerate");
    scope_star = -1; scope_new = false;
    walk(pruned);
    prog.latent = latent;
    if (!latent) {
      // absent MaybeSwap observations are pruned from the plan; propose_non_enumerable! samples them
      // with random() (block_proposal.jl:58-60) after evaluating the JuliaNodes they depend on
      for (int v : cm.blocks[block]) {
        const Node& bn = cm.nodes[v];
        if (bn.wrap != PCLEAN_WRAP_NONE) continue;
        if (bn.kind == PCLEAN_NODE_JULIA) { bind_row_function(v); continue; }
        if (bn.kind != PCLEAN_NODE_CHOICE || bn.dist != PCLEAN_DIST_MAYBE_SWAP || obs[v] || any_unavailable(bn.args)) continue;
        bool have = false;
        for (int i : prog.root_sampled) have = have || prog.mswaps[i].obs_vertex == v;
        if (have) continue;
        prog.mswaps.push_back(mswap_of(bn, v));
        prog.root_sampled.push_back((int)prog.mswaps.size() - 1);
      }
    }
    if (prog.root < 0 && !latent && (!prog.root_terms.empty() || !prog.root_sampled.empty())) { prog.rootless = true; return prog; }
    if (prog.root < 0) throw Unsupported(pruned.empty() ? "block with nothing to enumerate" : "block without an enumeration root");
    prog.latent = latent;
    if (!latent && prog.stars[prog.root].kind != ST_FK) throw Unsupported("block whose root is not a reference slot");
    // every choice vertex of every creatable table must be covered by a star: otherwise the
    // reference samples it from its prior in propose_non_enumerable! (block_proposal.jl:42-56)
    for (const StarL& s : prog.stars) {
      if (s.kind != ST_FK) continue;
      const ClassM& tm = m.classes[s.table];
      for (int tv = 0; tv < tm.n_normal; ++tv) {
        const Node& tn = tm.nodes[tv];
        if (tn.wrap != PCLEAN_WRAP_NONE) continue;
        if (tn.kind != PCLEAN_NODE_CHOICE && tn.kind != PCLEAN_NODE_FK) continue;
        bool covered = false;
        for (int c : s.children) if (prog.stars[c].tvertex == tv) covered = true;
        {   // a cell the row observes directly (e.g. rents countykey / state) needs no enumeration
          const Node& fk = cm.nodes[s.vertex];
          if (tv < (int)fk.vmap.size() && obs[fk.vmap[tv]]) covered = true;
        }
        if (!covered) {
          // sampled from its discrete proposal when the row is created (block_proposal.jl:42-60)
          if (tn.kind != PCLEAN_NODE_CHOICE || tn.dist != PCLEAN_DIST_TIME_PRIOR || latent)
            throw Unsupported("latent class with a choice that no observation informs (prior-sampled fill-in other than TimePrior)");
          const Node& fk = cm.nodes[s.vertex];
          const int ov = fk.vmap.at(tv);                       // the cell in the referring row
          const Node& on = cm.nodes[ov];
          StarL::FillL f; f.vertex = ov; f.dist = tn.dist;
          // its list argument, re-expressed over observed cells of the referring row
          const Node& ln = cm.nodes[on.args.at(0)];
          if (ln.kind != PCLEAN_NODE_JULIA) throw Unsupported("fill-in whose option list is not a JuliaNode");
          const FuncM& lf = m.funcs[ln.func];
          if (lf.kind == PCLEAN_FUNC_CONST && lf.cst.tag == PCLEAN_VAL_LIST) f.list = lf.cst.i;
          else if (lf.kind == PCLEAN_FUNC_TABLE && lf.keyargs.size() == 1 && obs[ln.args.at(lf.keyargs[0])]) {
            f.list_func = ln.func; f.list_arg.kind = ARG_OBS; f.list_arg.ref = ln.args.at(lf.keyargs[0]);
          } else throw Unsupported("fill-in whose option list is neither constant nor a lookup on an observed cell");
          std::string d = "**:** p.m.";
          f.dummy_string = intern(std::u32string(d.begin(), d.end()));
          fill_todo.push_back(std::make_pair((int)(&s - &prog.stars[0]), f));
        }
      }
    }
    for (auto& ft : fill_todo) prog.stars[ft.first].fillins.push_back(ft.second);
    fill_todo.clear();
    for (int r : prog.roots) postorder(r);
    return prog;
  }
};

}  // namespace pcl
