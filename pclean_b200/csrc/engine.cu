// engine.cu — C-ABI implementation of include/pclean_b200.h: host orchestration of the sweep.
//
// Product path only: no oracle code, no CPU fallback.  Every entry point fails loudly
// (status code + pclean_last_error) when CUDA is unavailable or a model shape is unsupported.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include <cub/cub.cuh>

#include "device.cuh"
#include "lower.hpp"
#include "latent.cuh"

using namespace pcl;

namespace {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_));    \
  } while (0)

template <class T> struct DBuf {
  T* p = nullptr; size_t n = 0;
  DBuf() {}
  DBuf(const DBuf&) = delete; DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  ~DBuf() { if (p) cudaFree(p); }
  void alloc(size_t count) {
    if (p) cudaFree(p);
    p = nullptr; n = count;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    const cudaError_t e_ = cudaMalloc(&p, bytes);
    if (e_ != cudaSuccess) {
      size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot);
      p = nullptr; n = 0;
      throw CudaError("cudaMalloc of " + std::to_string(bytes >> 20) + " MiB failed (" + std::to_string(fr >> 20) + " of " + std::to_string(tot >> 20) +
                      " MiB free): " + cudaGetErrorString(e_));
    }
    if (bytes >= ((size_t)1 << 30) && std::getenv("PCLEAN_LOG_ALLOC")) std::fprintf(stderr, "[pclean_b200] device allocation %.2f GiB\n", (double)bytes / (double)((size_t)1 << 30));
  }
  void upload(const std::vector<T>& h) { alloc(h.size()); if (!h.empty()) CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
  void zero() { if (p && n) CK(cudaMemset(p, 0, n * sizeof(T))); }
  std::vector<T> download(size_t count = (size_t)-1) const {
    if (count == (size_t)-1) count = n;
    std::vector<T> h(count);
    if (count) CK(cudaMemcpy(h.data(), p, count * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
  }
};

struct ObsCol { int vertex; bool is_real = false; std::vector<int> sid, uobs, ulist; std::vector<double> real; std::vector<char> absent;
                DBuf<int> d_uobs, d_ulist, d_sid; DBuf<double> d_real; int max_len = 0;
                DBuf<int> d_u_of_sid; };   // string id -> index in ulist (-1: not a value of this column); built on the first pclean_update_observations

struct TableH {
  int cls = -1, n_normal = 0, cap = 0, n_slots = 0, min_cap = 0, reserve = 0;
  bool loaded = false;
  std::vector<int64_t> keys;
  std::unordered_map<int64_t, int> slot_of_key;
  std::vector<pclean_value> raw;      // [n_cols][n_rows] as loaded
  int raw_cols = 0;
  DBuf<int> cells, refcnt, div; DBuf<double> logcnt, logcnt1; DBuf<uint8_t> alive; DBuf<long long> d_keys;
  std::vector<int> fk_col, fk_table;
  double strength = 1.0, discount = 0.0;
  uint32_t py_epoch = 0;
};

struct MatH {
  DBuf<uint8_t> d, elen; long long stride = 0; int rows = 0, cols = 0;
  int obs_col = -1; int table = -1, col = -1; int prefix_a = -1, prefix_sep = -1;
  int cols_done = 0;
  DBuf<int> shadow;              // string id each column was computed for (candidate matrices)
};

typedef struct { char internal[128]; } NcclUniqueId;
struct pinned_cols_t { std::vector<int*> host; size_t bytes_per_col = 0; };
struct Nccl {
  void* lib = nullptr; void* comm = nullptr; int rank = 0, world = 1; bool owned = false;   // owned: created by pclean_nccl_init (destroyed with the engine)
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
};

struct JoinTerm { int prog, term, kind, obs_col, table, col, opt_off, nopt, sep; };
struct Hoist { int prog, star, obs_col; std::unique_ptr<DBuf<double>> val; bool dynamic; };
struct UnivEntry { int opt_off, ids_off, count, splp_off, univ_off; };   // option universe of a list function (finalize)
struct ParamH { int spec = 0; std::vector<double> value; std::vector<int> prior_offs; int nopt = 0; uint32_t epoch = 0; };

}  // namespace

struct pclean_engine {
  pclean_config cfg{};
  int device = 0;
  std::string err;
  Model m;
  bool model_loaded = false, finalized = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  cudaEvent_t evb[16] = {};          // per-block k_block start/stop
  float block_ms[8] = {};
  pinned_cols_t* pinned = nullptr;
  // dictionary
  std::vector<std::u32string> strings;
  std::unordered_map<std::u32string, int> string_ids;
  int n_dev_strings = 0;
  DBuf<uint8_t> d_sym; DBuf<int> d_str_off, d_str_len;
  // strings random(StringPrior) generates on the device (device.cuh dummy_string_draw): pool + dictionary head-room
  int newstr_cap = 16384; size_t sym_used = 0, sym_cap = 0; int str_cap = 0; int lm_sym[28] = {};
  DBuf<uint8_t> d_newstr_chars; DBuf<int> d_newstr_len, d_newstr_count, d_lm_sym, d_newstr_map; DBuf<double> d_lm_uni, d_lm_big;
  std::map<char32_t, int> alphabet;
  DBuf<double> d_LG, d_LOGN, d_LUT;
  // observations
  int obs_cls = -1; int64_t N = 0;
  std::vector<int> pat_of_row; std::vector<std::vector<char>> pat_cols; std::vector<std::vector<long long>> pat_rows;
  std::vector<std::unique_ptr<DBuf<long long>>> d_pat_rows; DBuf<int> d_pat_of_row;
  std::vector<std::unique_ptr<ObsCol>> cols; std::map<int, int> col_of_vertex;
  DBuf<int*> d_uobs_ptrs;
  // tables
  std::vector<TableH> tables; DBuf<TableD> d_tables; std::vector<TableD> h_tables;
  int64_t next_key = 1;
  std::map<int, std::vector<int64_t>> assign_keys;   // fk vertex -> keys per row
  std::vector<std::unique_ptr<DBuf<int>>> d_assign; DBuf<int*> d_assign_ptrs;
  // programs
  std::vector<BlockProgram> progs, lprogs; std::vector<int> lprog_cls, lprog_mask;
  std::vector<char> prog_rich;              // program uses hash buckets / row-dependent lists / inner enumerations / equality terms
  std::map<int, std::vector<int>> lobs_cols;        // latent class -> its columns that observation rows can observe directly
  std::map<int, ObsCellsD> lobs_cells;
  std::map<std::pair<int, int>, int> lprog_of_pat;  // (class, observed-cell mask) -> program id
  std::map<std::pair<int, int>, std::string> lprog_pat_error;
  bool mats_dirty = true;            // a table cell changed since the candidate matrices were last refreshed
  bool row_state_synced = false;     // row-sharded engines: assignments / local cells of the other ranks are current
  DBuf<int> d_vcol;
  DBuf<long long> d_rowlist, d_rowlist_pat; DBuf<int> d_rowlist_int;    // pclean_init_trace: the rows of the current batch (any order)
  DBuf<int> d_lpat, d_lslots; std::vector<int> h_lpat; std::vector<int> lpats_present; bool lpat_active = false;
  std::vector<GaussExtD> h_gext; DBuf<GaussExtD> d_gext;
  std::vector<MswapD> h_mswaps; DBuf<MswapD> d_mswaps;
  std::vector<FillD> h_fills; DBuf<FillD> d_fills;
  std::vector<double> h_lkconst; DBuf<double> d_lkconst;
  std::vector<int> h_time_sid; DBuf<int> d_time_sid; DBuf<uint8_t> d_time_ok;
  std::vector<char> prog_rootless;
  std::vector<int> lprog_block; std::map<std::pair<int, int>, int> lprog_trivial;   // (class, mask) whose move has nothing to enumerate
  struct MswapSite { MswapD d; };
  std::vector<MswapD> msites;               // MaybeSwap nodes of the observation class (ProbParameter statistics)
  // path arrays of the IR (copied at load: the caller owns the IR buffers)
  int ir_n_paths = 0; std::vector<int> ir_path_target, ir_path_len_off, ir_path_class, ir_path_vertex, ir_path_vmap_off, ir_path_vmap;
  pclean_model_ir ir_view{};
  std::vector<ProgD> h_progs; std::vector<StarD> h_stars; std::vector<TermD> h_terms; std::vector<int> h_children;
  std::vector<int2> h_copies; std::vector<double> h_prior; std::vector<int> h_optsid;
  DBuf<ProgD> d_progs; DBuf<StarD> d_stars; DBuf<TermD> d_terms; DBuf<int> d_children; DBuf<int2> d_copies;
  DBuf<double> d_prior; DBuf<int> d_optsid;
  std::vector<std::unique_ptr<MatH>> mats; DBuf<MatD> d_mats; std::vector<MatD> h_mats;
  std::vector<JoinTerm> joins; int max_a = 512; std::vector<int> a_sids;
  DBuf<int> d_join_mat, d_a_slot, d_needed_a, d_needed_any; std::vector<int> h_join_mat;
  DBuf<double> d_stats2;
  std::vector<Hoist> hoists; DBuf<double*> d_hoist_ptrs;
  std::vector<ParamH> params;
  // particles
  int K = 0, n_blocks = 0, nvC = 0;
  std::vector<std::unique_ptr<DBuf<int>>> d_pchoice; DBuf<int*> d_pchoice_ptrs;
  DBuf<double> d_pweight, d_plogml, d_row_logml;
  DBuf<unsigned long long> d_row_bad; DBuf<int> d_dbg;
  DBuf<int> d_sel, d_row_flags, d_pool, d_pool_count, d_err, d_req, d_flags, d_rank, d_counter;
  int pool_cap = 0;
  DBuf<uint8_t> d_cub_tmp;
  DBuf<unsigned long long> d_memo_keys[2]; DBuf<ulonglong2> d_memo_vals[2]; int memo_log2 = 22;
  bool pmemo_dirty = true;           // the persistent (choice-star) memo must be cleared before the next launch
  int opts = PCL_OPT_PROGRESSIVE | PCL_OPT_PMEMO | PCL_OPT_FASTEXCL | PCL_OPT_PARHINT | PCL_OPT_LAZYNEW;
  DBuf<int> d_term_order; DBuf<float> d_col_meanlen;
  DBuf<Dev> d_dev; Dev h_dev{};
  int64_t shard_begin = 0, shard_end = -1;
  Nccl nccl;
  int launches = 0;
  int64_t total_new_rows = 0;
  int max_batch_new = 0;             // most rows one batch has appended to a table so far (head-room the compaction trigger keeps)
  std::map<std::tuple<int, int, int, int, int>, UnivEntry> univ_cache;
  bool obs_host_stale = false;       // pclean_update_observations changed the device copy of the observed cells: the host mirror is refreshed before it is read
  bool compact_now = false;          // option "compact_now": pack the tables before the next class sweep whatever their fill
  int compact_head = 64;             // option "compact_headroom": least free slots a table keeps before it is packed
  int compactions = 0; long long compact_futile_at = -1;   // total slot count at which the last check found nothing to pack
  int prune = 1;
  int block_grid = 148 * 2;
  int kb_variant = 0;                // k_block geometry: 0 = 16 warps x 2 CTAs (<= 64 registers), 1 = 12 x 2 (<= 80), 2 = 16 x 1 (<= 128)
  uint64_t param_seed = 0;           // seed of the keyed prior draws that initialise parameters nobody set (initialize_parameter)
  int init_divisor = 8;              // pclean_init_trace: a batch holds done / init_divisor rows (smaller batches = fewer duplicate entities, more launches)
  int64_t init_rows = 0;             // > 0: pclean_init_trace stops after this many rows (tests)
  int table_cap = 65536;             // pclean_init_trace: capacity reserved per latent table (rows the run may create)
  int64_t batch_rows = 0;            // > 0: observation rows are moved in consecutive batches of this size (1 = the reference's sequential Gibbs order)
  int resample_params = 1;           // resample @learned parameters + PY hyper-parameters at each latent class sweep
  int exchange_path = 0;             // 1: create rows through the gathered-record path even on one GPU (tests)
  // latent-class programs
  std::map<int, int> lprog_of_class;               // class -> program id (single-block latent classes)
  std::map<int, std::string> lprog_error;          // class -> why it could not be lowered
  std::map<int, RefChainD> ref_chain;
  std::map<std::tuple<int, int, int>, int> cand_mats;
  std::map<std::tuple<int, int>, int> opt_mats;
  std::vector<std::vector<int>> prog_opt_off;
  DBuf<int> d_lref_off, d_lref_rows, d_slot_of_row, d_iota, d_lchoice, d_lsel, d_lflags, d_collist;
  DBuf<double> d_llogml;
  // referrer group sets of latent moves (latent.cuh): one per (dataset column [, referring-row cell of a string join])
  std::vector<GroupSetD> gsets; std::vector<std::vector<int>> gsets_of_class;
  std::vector<std::unique_ptr<DBuf<unsigned long long>>> d_grp_key; std::vector<std::unique_ptr<DBuf<int>>> d_grp_cnt;
  DBuf<unsigned long long> d_grp_tmp; DBuf<int> d_grp_n; DBuf<unsigned long long*> d_grp_key_ptrs; DBuf<int*> d_grp_cnt_ptrs; DBuf<uint8_t> d_grp_cub;
  std::vector<std::unique_ptr<DBuf<long long>>> d_keys;
  DBuf<int*> d_ulist_ptrs;
  DBuf<uint8_t> d_sort_tmp;
  int max_cap = 0;
  // inner enumerations, tabulated functions, per-string side tables (rents shapes)
  std::vector<InnerD> h_inners; DBuf<InnerD> d_inners;
  std::vector<int> h_innervals; DBuf<int> d_innervals;
  struct LookupH { DBuf<int> keys, vals; unsigned mask = 0; int nkey = 0; };
  // per-list distance blocks of a row-dependent option list (device.cuh ListMatD): host description + device buffers
  struct ListMatH {
    int obs_col = -1, words = 1; size_t bytes = 0;
    std::vector<long long> row_off, elen_off, eoff; std::vector<int> pat_sids, esids, ncols, rkeys, rvals; std::vector<uint8_t> elen;
    std::vector<std::pair<int, std::pair<long long, long long>>> ranges;   // list -> rows [first, second) in pat_sids
    unsigned rmask = 0;
    DBuf<uint8_t> d, d_elen; DBuf<long long> d_row_off, d_elen_off; DBuf<int> d_rkeys, d_rvals;
  };
  std::vector<std::unique_ptr<ListMatH>> lmats; std::map<std::tuple<int, int>, int> lmat_of; DBuf<ListMatD> d_lmats;
  std::map<int, int> lookup_of_func; std::vector<std::unique_ptr<LookupH>> lookups; DBuf<LookupD> d_lookups;
  struct GaussSiteH { GaussSiteD d; double stdev; std::vector<int> slots; };
  std::vector<GaussSiteH> gsites;           // observed Gaussian nodes of the observation class feeding MeanParameters
  DBuf<double> d_x_of_row, d_msum, d_mcnt;
  DBuf<int> d_lists_off, d_lists_sid;
  std::vector<double> h_splp; DBuf<double> d_splp; std::vector<int> h_univ; DBuf<int> d_univ; std::vector<int> h_optmap; DBuf<int> d_optmap;
  DBuf<double> d_param_real, d_xform;
  std::vector<std::unique_ptr<DBuf<int>>> d_bkt_off, d_bkt_slots; DBuf<int*> d_bkt_off_ptrs, d_bkt_slots_ptrs; std::vector<int> bucket_col_of_table;
  DBuf<int> d_bkt_keys, d_bkt_keys_out, d_bkt_iota;
  std::map<int, std::vector<int>> rowcell_init;          // vertex -> value ids as loaded
  std::vector<std::unique_ptr<DBuf<int>>> d_rowcell; DBuf<int*> d_rowcell_ptrs;
  std::vector<std::unique_ptr<DBuf<int>>> d_pinner; DBuf<int*> d_pinner_ptrs;
  DBuf<int*> d_obs_sid_ptrs; DBuf<double*> d_obs_real_ptrs;
  int n_patterns = 1;
  std::map<std::pair<int, int>, std::unique_ptr<DBuf<int2>>> fk_copies;   // (class, fk index) -> (local vertex, target column)
  std::map<std::pair<int, int>, int> fk_ncopies;
  DBuf<int> d_rec_local, d_rec_all, d_src, d_row_ids, d_counts;

  int intern(const std::u32string& s) {
    auto it = string_ids.find(s);
    if (it != string_ids.end()) return it->second;
    const int id = (int)strings.size();
    strings.push_back(s); string_ids.emplace(s, id);
    return id;
  }
};

namespace {

typedef pclean_engine Eng;

int guard(Eng* h, const std::function<void()>& f) {
  try { f(); return PCLEAN_OK; }
  catch (const Unsupported& e) { h->err = std::string("unsupported: ") + e.what(); return PCLEAN_ERR_UNSUPPORTED; }
  catch (const BadArg& e) { h->err = std::string("bad argument: ") + e.what(); return PCLEAN_ERR_ARG; }
  catch (const CudaError& e) { h->err = std::string("cuda: ") + e.what(); return PCLEAN_ERR_CUDA; }
  catch (const std::exception& e) { h->err = e.what(); return PCLEAN_ERR_STATE; }
}

// StringPrior log-density (string_prior.jl:43-61)
double stringprior_logdensity(const Model& m, const std::u32string& s, int minl, int maxl) {
  const int len = (int)s.size();
  if (len < minl || len > maxl) return -INFINITY;
  double score = -std::log((double)(maxl - minl + 1));
  int prev = -1;
  for (char32_t ch : s) {
    char32_t c = ch;
    if (c >= U'A' && c <= U'Z') c = c - U'A' + U'a';
    int cur = -1;
    if (c >= U'a' && c <= U'z') cur = (int)(c - U'a'); else if (c == U' ') cur = 26; else if (c == U'.') cur = 27;
    if (cur < 0) score += -std::log(28.0);
    else { const double pr = prev < 0 ? m.lm_uni[cur] : m.lm_big[cur * 28 + prev]; score += std::max(std::log(pr), -1000.0); }
    prev = cur;
  }
  return score;
}
bool time_regex(const std::u32string& s) {     // ^\d?\d:\d\d [ap]\.m\.$  (time_prior.jl:10)
  const size_t n = s.size();
  auto dig = [&](size_t i) { return i < n && s[i] >= U'0' && s[i] <= U'9'; };
  size_t p = 0;
  if (!dig(p)) return false;
  ++p; if (dig(p)) ++p;
  if (p >= n || s[p] != U':') return false;
  ++p;
  if (!dig(p) || !dig(p + 1)) return false;
  p += 2;
  if (p + 5 != n) return false;
  return s[p] == U' ' && (s[p + 1] == U'a' || s[p + 1] == U'p') && s[p + 2] == U'.' && s[p + 3] == U'm' && s[p + 4] == U'.';
}
double lse_host(const std::vector<double>& x) {
  double mx = -INFINITY; for (double v : x) mx = std::max(mx, v);
  if (mx == -INFINITY) return mx;
  double s = 0; for (double v : x) s += std::exp(v - mx);
  return mx + std::log(s);
}
inline int nblk(int64_t n, int t) { return (int)std::max<int64_t>(1, (n + t - 1) / t); }

void upload_dev(Eng* h) { CK(cudaMemcpy(h->d_dev.p, &h->h_dev, sizeof(Dev), cudaMemcpyHostToDevice)); }
void upload_tables(Eng* h) {
  for (size_t c = 0; c < h->tables.size(); ++c) h->h_tables[c].n_slots = h->tables[c].n_slots;
  CK(cudaMemcpy(h->d_tables.p, h->h_tables.data(), h->h_tables.size() * sizeof(TableD), cudaMemcpyHostToDevice));
}
void upload_mats(Eng* h) {
  h->h_mats.resize(h->mats.size());
  for (size_t i = 0; i < h->mats.size(); ++i) { h->h_mats[i].d = h->mats[i]->d.p; h->h_mats[i].stride = h->mats[i]->stride; h->h_mats[i].elen = h->mats[i]->elen.p; }
  if (h->d_mats.n < h->h_mats.size()) { h->d_mats.alloc(h->h_mats.size() + 256); h->h_dev.mats = h->d_mats.p; upload_dev(h); }
  if (!h->h_mats.empty()) CK(cudaMemcpy(h->d_mats.p, h->h_mats.data(), h->h_mats.size() * sizeof(MatD), cudaMemcpyHostToDevice));
}

// bit-parallel DP for columns [e0, e1) of matrix M; column e <-> string id d_elem_ids[e]
void run_dp(Eng* h, MatH& M, const int* d_elem_ids, int e0, int e1, const int* d_col_list = nullptr) {
  if (e1 <= e0 || M.rows == 0) return;
  const ObsCol& oc = *h->cols[M.obs_col];
  DpArgs A{};
  A.sym = h->d_sym.p; A.str_off = h->d_str_off.p; A.str_len = h->d_str_len.p;
  A.pat_ids = oc.d_ulist.p; A.n_pat = M.rows;
  A.elem_ids = d_elem_ids; A.elem0 = e0; A.n_elem = e1 - e0;
  A.prefix_a = M.prefix_a; A.prefix_sep = M.prefix_sep;
  A.out = M.d.p; A.stride = M.stride; A.words = std::max(1, (oc.max_len + 63) / 64); A.col_list = d_col_list;
  if (A.words > OSA_MAX_WORDS) throw Unsupported("observed string longer than 256 symbols");
  const int gx = std::min(nblk(e1 - e0, 128), 128);
  for (int p0 = 0; p0 < M.rows; p0 += 65535) {
    A.pat0 = p0; A.elem_len = p0 == 0 ? M.elen.p : nullptr;
    dim3 grid(gx, std::min(65535, M.rows - p0));
    k_dp_matrix<<<grid, 128, 256 * A.words * sizeof(uint64_t), h->stream>>>(A);
    ++h->launches;
  }
  CK(cudaGetLastError());
}

int new_mat(Eng* h, int obs_col, int rows, int cols_cap) {
  std::unique_ptr<MatH> M(new MatH());
  M->obs_col = obs_col; M->rows = rows; M->cols = cols_cap; M->stride = ((long long)cols_cap + 15) / 16 * 16;
  M->d.alloc((size_t)std::max(1, rows) * M->stride); M->elen.alloc(M->stride);
  CK(cudaMemset(M->elen.p, 0, M->stride));
  h->mats.push_back(std::move(M));
  return (int)h->mats.size() - 1;
}

// Per-list distance blocks (device.cuh ListMatD).  plan: which (list, unique observed string) pairs the dataset
// shows — rows of the observation class whose `key_col` value selects list l and whose `obs_col` value is u —,
// block layout, element ids and lengths, the (u, l) -> row hash.  build: device buffers + one DP launch per list.
int plan_list_mat(Eng* h, int obs_col, int list_func, int key_col, int dummy_sid) {
  const Model& m = h->m;
  const FuncM& lf = m.funcs.at(list_func);
  const ObsCol& oc = *h->cols[obs_col];
  std::unique_ptr<pclean_engine::ListMatH> LM(new pclean_engine::ListMatH());
  LM->obs_col = obs_col; LM->words = std::max(1, (oc.max_len + 63) / 64);
  const int n_lists = (int)m.lists.size();
  auto slen = [&](int sid) { return (int)std::min<size_t>(255, h->strings[sid].size()); };
  LM->elen_off.assign(n_lists, -1); LM->eoff.assign(n_lists, -1); LM->ncols.assign(n_lists, 0);
  std::unordered_map<int, int> list_of_key;
  for (auto& kv : lf.table) {
    if (kv.second.tag != PCLEAN_VAL_LIST) continue;
    const int l = kv.second.i;
    if (kv.first.size() == 1) list_of_key[kv.first[0]] = l;
    if (LM->eoff[l] >= 0) continue;
    LM->eoff[l] = (long long)LM->esids.size(); LM->elen_off[l] = (long long)LM->elen.size();
    for (const Val& v : m.lists.at(l)) if (v.tag == PCLEAN_VAL_STR) { LM->esids.push_back(v.i); LM->elen.push_back((uint8_t)slen(v.i)); }
    LM->esids.push_back(dummy_sid); LM->elen.push_back((uint8_t)slen(dummy_sid));       // the placeholder: last column of every block
    LM->ncols[l] = (int)LM->esids.size() - (int)LM->eoff[l];
  }
  std::vector<unsigned long long> pairs;
  if (key_col >= 0) {
    const ObsCol& kc = *h->cols[key_col];
    for (int64_t r = 0; r < h->N; ++r) {
      const int key = kc.sid[r], u = oc.uobs[r];
      if (key < 0 || u < 0) continue;
      auto it = list_of_key.find(key);
      if (it != list_of_key.end()) pairs.push_back(((unsigned long long)(unsigned)it->second << 32) | (unsigned)u);
    }
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
  }
  const size_t np = pairs.size();
  LM->row_off.resize(np); LM->pat_sids.resize(np);
  size_t cap = 16; while (cap < np * 2 + 2) cap <<= 1;
  LM->rkeys.assign(cap * 3, PCL_LOOKUP_EMPTY); LM->rvals.assign(cap, 0); LM->rmask = (unsigned)(cap - 1);
  auto hmix = [](unsigned long long x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; };
  for (size_t i = 0; i < np; ++i) {
    const int l = (int)(pairs[i] >> 32), u = (int)(pairs[i] & 0xFFFFFFFFull);
    if (LM->ranges.empty() || LM->ranges.back().first != l) LM->ranges.push_back({l, {(long long)i, (long long)i}});
    LM->ranges.back().second.second = (long long)i + 1;
    LM->row_off[i] = (long long)LM->bytes; LM->bytes += (size_t)LM->ncols[l];
    LM->pat_sids[i] = oc.ulist[u];
    const unsigned long long key = hmix((unsigned long long)(unsigned)u * 0x9E3779B97F4A7C15ULL ^ ((unsigned long long)(unsigned)l << 20));      // lookup_find(u, l, 0)
    size_t hh = (size_t)((unsigned)key & (unsigned)(cap - 1));
    while (LM->rkeys[3 * hh] != PCL_LOOKUP_EMPTY) hh = (hh + 1) & (cap - 1);
    LM->rkeys[3 * hh] = u; LM->rkeys[3 * hh + 1] = l; LM->rkeys[3 * hh + 2] = 0; LM->rvals[hh] = (int)i;
  }
  if (np >= ((size_t)1 << 31)) throw Unsupported("more than 2^31 (observed string, option list) pairs in one column");
  h->lmats.push_back(std::move(LM));
  return (int)h->lmats.size() - 1;
}
void build_list_mats(Eng* h) {
  std::vector<ListMatD> dl;
  for (auto& LMp : h->lmats) {
    pclean_engine::ListMatH& LM = *LMp;
    LM.d.alloc(std::max<size_t>(16, LM.bytes)); LM.d_elen.upload(LM.elen); LM.d_row_off.upload(LM.row_off); LM.d_elen_off.upload(LM.elen_off);
    LM.d_rkeys.upload(LM.rkeys); LM.d_rvals.upload(LM.rvals);
    DBuf<int> d_pat, d_es; d_pat.upload(LM.pat_sids); d_es.upload(LM.esids);
    if (LM.words > OSA_MAX_WORDS) throw Unsupported("observed string longer than 256 symbols");
    for (auto& rg : LM.ranges) {
      const int l = rg.first; const long long i0 = rg.second.first, i1 = rg.second.second;
      DpArgs A{};
      A.sym = h->d_sym.p; A.str_off = h->d_str_off.p; A.str_len = h->d_str_len.p;
      A.pat_ids = d_pat.p + i0; A.n_pat = (int)(i1 - i0);
      A.elem_ids = d_es.p + LM.eoff[l]; A.elem0 = 0; A.n_elem = LM.ncols[l];
      A.prefix_a = -1; A.prefix_sep = -1;
      A.out = LM.d.p + LM.row_off[i0]; A.stride = LM.ncols[l]; A.words = LM.words; A.col_list = nullptr; A.elem_len = nullptr;
      const int gx = std::min(nblk(A.n_elem, 128), 128);
      for (int p0 = 0; p0 < A.n_pat; p0 += 65535) {
        A.pat0 = p0;
        dim3 grid(gx, std::min(65535, A.n_pat - p0));
        k_dp_matrix<<<grid, 128, 256 * A.words * sizeof(uint64_t), h->stream>>>(A);
        ++h->launches;
      }
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));            // d_pat / d_es are freed at scope exit
    ListMatD D{};
    D.d = LM.d.p; D.row_off = LM.d_row_off.p; D.elen = LM.d_elen.p; D.elen_off = LM.d_elen_off.p;
    D.rows = LookupD{LM.d_rkeys.p, LM.d_rvals.p, LM.rmask, 2};
    dl.push_back(D);
  }
  if (dl.empty()) dl.push_back(ListMatD{});
  h->d_lmats.upload(dl);
}

// Recompute the columns of every candidate matrix whose clean string changed (new slots, values
// rewritten by a latent-class move) — a shadow copy of the string ids tells which.
void refresh_one_mat(Eng* h, MatH& M) {
  TableH& T = h->tables[M.table];
  const int n = T.n_slots;
  if (n == 0) return;
  if (M.shadow.n == 0) { M.shadow.alloc(T.cap); CK(cudaMemsetAsync(M.shadow.p, 0xFF, (size_t)T.cap * sizeof(int), h->stream)); }
  const int* col = T.cells.p + (size_t)M.col * T.cap;
  k_diff_cols<<<nblk(n + 1, 256), 256, 0, h->stream>>>(col, M.shadow.p, n, h->d_flags.p); ++h->launches;
  size_t tmp = h->d_cub_tmp.n;
  CK(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, h->d_flags.p, h->d_rank.p, n + 1, h->stream)); ++h->launches;
  int total = 0;
  CK(cudaMemcpyAsync(&total, h->d_rank.p + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (total == 0) return;
  k_compact_cols<<<nblk(n, 256), 256, 0, h->stream>>>(col, M.shadow.p, n, h->d_flags.p, h->d_rank.p, h->d_collist.p); ++h->launches;
  run_dp(h, M, col, 0, total, h->d_collist.p);
}
// Candidate matrices follow the strings in their table column; cells only change when rows are
// created or a latent class is applied (both set mats_dirty), so a plain observation sweep skips
// the whole scan (≈ 100 launches + host syncs per sweep on the hospital schema).
void refresh_candidate_mats(Eng* h) {
  if (!h->mats_dirty) return;
  for (auto& Mp : h->mats) if (Mp->table >= 0) refresh_one_mat(h, *Mp);
  h->mats_dirty = false;
}

void recount(Eng* h) {
  {   // all tables in one launch (grid.y = class); an unloaded class has cap 0 in its descriptor
    dim3 grid(nblk(std::max(1, h->max_cap), 256), (unsigned)h->tables.size());
    k_zero_refcnt<<<grid, 256, 0, h->stream>>>(h->d_tables.p); ++h->launches;
  }
  const int64_t r0 = h->shard_begin, r1 = h->shard_end < 0 ? h->N : h->shard_end;
  for (int b = 0; b < h->n_blocks; ++b) {
    if (h->progs[b].root < 0) continue;                         // block without a reference slot
    const int t = h->progs[b].stars[h->progs[b].root].table;
    k_count_assign<<<nblk(r1 - r0, 256), 256, 0, h->stream>>>(h->d_assign[b]->p + r0, r1 - r0, h->tables[t].refcnt.p);
    ++h->launches;
  }
  if (h->nccl.comm) {   // the one collective of the sweep: row shards -> global reference counts
    std::vector<int> roots;                                     // two blocks may root the same table (src ~ Airport, dst ~ Airport): reduce it once
    for (int b = 0; b < h->n_blocks; ++b) {
      if (h->progs[b].root < 0) continue;
      const int t = h->progs[b].stars[h->progs[b].root].table;
      if (std::find(roots.begin(), roots.end(), t) == roots.end()) roots.push_back(t);
    }
    for (int t : roots) {
      TableH& T = h->tables[t];
      if (h->nccl.AllReduce(T.refcnt.p, T.refcnt.p, (size_t)T.cap, /*ncclInt32*/ 2, /*ncclSum*/ 0, h->nccl.comm, h->stream) != 0)
        throw std::runtime_error("ncclAllReduce failed");
    }
  }
  for (int c = (int)h->tables.size() - 1; c >= 0; --c) {
    TableH& T = h->tables[c];
    if (!T.loaded || T.n_slots == 0) continue;
    for (size_t g = 0; g < T.fk_col.size(); ++g) { k_count_table<<<nblk(T.n_slots, 256), 256, 0, h->stream>>>(h->d_tables.p, c, (int)g); ++h->launches; }
  }
  k_table_stats<<<(unsigned)h->tables.size(), 256, 0, h->stream>>>(h->d_tables.p); ++h->launches;      // one block per class
  // how selective each candidate column is now, and from it the order in which k_block's pruning pass reads a star's terms
  if (h->opts & PCL_OPT_PROGRESSIVE) {
    for (int b = 0; b < h->n_blocks; ++b) {
      if (h->progs[b].root < 0) continue;
      const int t = h->progs[b].stars[h->progs[b].root].table;
      bool seen = false;
      for (int b2 = 0; b2 < b; ++b2) seen = seen || (h->progs[b2].root >= 0 && h->progs[b2].stars[h->progs[b2].root].table == t);
      if (seen || h->tables[t].n_normal <= 0) continue;
      k_col_diversity<<<h->tables[t].n_normal, 256, 0, h->stream>>>(h->d_tables.p, t); ++h->launches;
    }
    const int np = h->n_patterns * h->n_blocks;
    k_term_order<<<np, 32, 0, h->stream>>>(h->d_dev.p, np, h->d_term_order.p); ++h->launches;
  }
  CK(cudaGetLastError());
}

void upload_param_priors(Eng* h) {
  h->pmemo_dirty = true;               // choice-star marginals were computed from the old priors
  for (auto& P : h->params) {
    if (P.prior_offs.empty() || P.value.empty()) continue;
    std::vector<double> lp(P.nopt);
    for (int i = 0; i < P.nopt; ++i) lp[i] = std::log(P.value[i]);     // utils.jl:33-36
    for (int off : P.prior_offs) CK(cudaMemcpy(h->d_prior.p + off, lp.data(), P.nopt * sizeof(double), cudaMemcpyHostToDevice));
  }
}

void compute_hoists(Eng* h, bool only_dynamic) {
  for (auto& H : h->hoists) {
    if (only_dynamic && !H.dynamic) continue;
    const int U = (int)h->cols[H.obs_col]->ulist.size();
    if (U == 0) continue;
    k_hoist<<<std::min(U, 4096), 128, 0, h->stream>>>(h->d_dev.p, H.prog, H.star, U, H.val->p);
    ++h->launches;
  }
  CK(cudaGetLastError());
}

void build_join_mats_for(Eng* h, int a_sid) {
  if ((int)h->a_sids.size() >= h->max_a) throw std::runtime_error("too many distinct upstream string values (join matrices)");
  const int slot = (int)h->a_sids.size();
  h->a_sids.push_back(a_sid);
  for (size_t j = 0; j < h->joins.size(); ++j) {
    const JoinTerm& J = h->joins[j];
    const ObsCol& oc = *h->cols[J.obs_col];
    int mi;
    if (J.kind == TERM_JOIN_CAND) {
      TableH& T = h->tables[J.table];
      mi = new_mat(h, J.obs_col, (int)oc.ulist.size(), T.cap);
      MatH& M = *h->mats[mi];
      M.table = J.table; M.col = J.col; M.prefix_a = a_sid; M.prefix_sep = J.sep;
      refresh_one_mat(h, M);
    } else {
      mi = new_mat(h, J.obs_col, (int)oc.ulist.size(), J.nopt);
      MatH& M = *h->mats[mi];
      M.prefix_a = a_sid; M.prefix_sep = J.sep;
      run_dp(h, M, h->d_optsid.p + J.opt_off, 0, J.nopt);
    }
    h->h_join_mat[j * h->max_a + slot] = mi;
  }
  CK(cudaMemcpy(h->d_join_mat.p, h->h_join_mat.data(), h->h_join_mat.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_a_slot.p + a_sid, &slot, sizeof(int), cudaMemcpyHostToDevice));
  upload_mats(h);
}

// ------------------------------------------------------------------------------------------
// finalize: everything that needs model + observations + tables
// ------------------------------------------------------------------------------------------
void finalize(Eng* h) {
  if (h->finalized) return;
  if (!h->model_loaded) throw std::runtime_error("pclean_load_model has not been called");
  if (h->obs_cls < 0) throw std::runtime_error("pclean_load_observations has not been called");
  const Model& m = h->m;
  const ClassM& cm = m.classes[h->obs_cls];
  if (cm.n_incoming) throw BadArg("observation class has incoming references (inference.jl:1-2)");
  h->K = h->cfg.num_particles;
  if (h->K < 1 || h->K > PCL_MAX_K) throw Unsupported("num_particles must be in 1.." + std::to_string(PCL_MAX_K) + " in this build");
  // block_proposal.jl:168: without data-driven proposals the reference proposes every unobserved
  // cell from its prior; only the data-driven (compiled enumeration) path is built here.
  // use_lo_sweeps is read by the reference's instrumented driver only (instrumented_inference.jl:98,329):
  // pgibbs_sweep! (inference.jl:60-81) sweeps every class regardless, and so does the engine.
  if (!h->cfg.use_dd_proposals) throw Unsupported("InferenceConfig.use_dd_proposals = false (prior proposals, block_proposal.jl:168) is not built; the engine runs data-driven proposals only");
  h->n_blocks = (int)cm.blocks.size();
  h->nvC = cm.nv;
  h->max_cap = 0;

  // ---- programs, one set of blocks per missingness pattern (may intern dummy placeholder strings)
  h->n_patterns = (int)std::max<size_t>(1, h->pat_cols.size());
  h->progs.clear();
  for (int pt = 0; pt < h->n_patterns; ++pt) {
    std::vector<char> obsv(cm.nv, 0);
    for (size_t c = 0; c < h->cols.size(); ++c) if (h->pat_cols.empty() || h->pat_cols[pt][c]) obsv[h->cols[c]->vertex] = 1;
    for (int b = 0; b < h->n_blocks; ++b) {
      Lowerer L(m, h->obs_cls);
      L.intern = [h](const std::u32string& s) { return h->intern(s); };
      h->progs.push_back(L.lower_block(b, obsv));
    }
  }
  // ---- latent-class programs (lowered here so that placeholder strings enter the dictionary)
  h->lprogs.clear(); h->lprog_cls.clear(); h->lprog_mask.clear(); h->lprog_block.clear(); h->lprog_trivial.clear(); h->lprog_error.clear(); h->lprog_pat_error.clear(); h->lobs_cols.clear();
  {
    std::vector<char> dobs(cm.nv, 0);
    for (auto& c : h->cols) dobs[c->vertex] = 1;
    // cells of latent rows that the dataset observes directly (rents: county.countykey, county.state)
    for (auto& c : h->cols) {
      const Node& n = cm.nodes[c->vertex];
      if (n.wrap != PCLEAN_WRAP_SUBMODEL || n.wfk.empty()) continue;
      const Node& fk = cm.nodes[n.wfk[0]];
      if (n.wfk.size() != 1) throw Unsupported("observed cell of a nested reference");
      h->lobs_cols[fk.target].push_back(n.wsub[0]);
    }
    for (int c = 0; c < (int)m.classes.size(); ++c) {
      if (c == h->obs_cls || !h->tables[c].loaded) continue;
      std::vector<int>& oc = h->lobs_cols[c];
      std::sort(oc.begin(), oc.end()); oc.erase(std::unique(oc.begin(), oc.end()), oc.end());
      if (oc.size() > 6) { h->lprog_error[c] = "latent class with more than 6 directly observed columns"; continue; }
      for (int mask = 0; mask < (1 << oc.size()); ++mask) {
        try {
          // blocks whose plan enumerates nothing contribute the same factor to every particle; a
          // class may have one block that does enumerate (flights Flight: block 2)
          std::vector<char> own(m.classes[c].nv, 0);
          for (size_t q = 0; q < oc.size(); ++q) if (mask >> q & 1) own[oc[q]] = 1;
          int found = -1; BlockProgram keep;
          for (int b = 0; b < (int)m.classes[c].blocks.size(); ++b) {
            Lowerer L(m, c);
            L.intern = [h](const std::u32string& s) { return h->intern(s); };
            L.latent = true; L.data_cls = h->obs_cls; L.data_obs = &dobs; L.ir = &h->ir_view;
            try { BlockProgram bp = L.lower_block(b, own); if (found >= 0) throw BadArg("latent class with several enumerating blocks"); found = b; keep = std::move(bp); }
            catch (const Unsupported& e) {
              const std::string w = e.what();
              if (w != "block without an enumeration root" && w != "block with nothing to enumerate") throw;
            }
          }
          if (found < 0) { h->lprog_trivial[std::make_pair(c, mask)] = 1; continue; }
          h->lprogs.push_back(std::move(keep));
          h->lprog_cls.push_back(c); h->lprog_mask.push_back(mask); h->lprog_block.push_back(found);
        } catch (const BadArg& e) { h->lprog_pat_error[std::make_pair(c, mask)] = e.what(); if (oc.empty()) h->lprog_error[c] = e.what(); }
          catch (const Unsupported& e) { h->lprog_pat_error[std::make_pair(c, mask)] = e.what(); if (oc.empty()) h->lprog_error[c] = e.what(); }
      }
    }
  }
  // TimePrior.random draws "h:m a.m." strings (time_prior.jl:20-22): all 1440 of them enter the dictionary
  h->h_time_sid.clear();
  {
    bool any_time = false;
    for (const ClassM& c2 : m.classes) for (const Node& nd : c2.nodes) any_time = any_time || (nd.kind == PCLEAN_NODE_CHOICE && nd.dist == PCLEAN_DIST_TIME_PRIOR);
    if (any_time)
      for (int hh = 1; hh <= 12; ++hh) for (int mi = 1; mi <= 60; ++mi) for (int pm = 0; pm < 2; ++pm) {
        const std::string t = std::to_string(hh) + ":" + std::to_string(mi) + (pm ? " p.m." : " a.m.");
        h->h_time_sid.push_back(h->intern(std::u32string(t.begin(), t.end())));
      }
  }
  for (const ClassM& c2 : m.classes) h->nvC = std::max(h->nvC, c2.nv);      // scratch records hold a row of any class

  // ---- dictionary
  std::map<char32_t, int>& alphabet = h->alphabet;
  alphabet.clear();
  {
    // the 28 letters random(StringPrior) can produce get dictionary symbols up front (string_prior.jl:10)
    static const char32_t lm_letters[] = U"abcdefghijklmnopqrstuvwxyz .";
    bool any_sp = false;
    for (const ClassM& c2 : m.classes) for (const Node& nd : c2.nodes) any_sp = any_sp || (nd.kind == PCLEAN_NODE_CHOICE && nd.dist == PCLEAN_DIST_STRING_PRIOR);
    for (int q = 0; q < 28; ++q) {
      h->lm_sym[q] = 255;
      if (!any_sp) continue;
      auto it = alphabet.find(lm_letters[q]);
      if (it == alphabet.end() && alphabet.size() < 256) it = alphabet.emplace(lm_letters[q], (int)alphabet.size()).first;
      if (it != alphabet.end()) h->lm_sym[q] = it->second;
    }
  }
  std::vector<uint8_t> sym; std::vector<int> off, len;
  for (const std::u32string& s : h->strings) {
    off.push_back((int)sym.size()); len.push_back((int)s.size());
    if (s.size() > 255) throw Unsupported("string longer than 255 codepoints");
    for (char32_t c : s) {
      auto it = alphabet.find(c);
      if (it == alphabet.end()) {
        if (alphabet.size() >= 256) throw Unsupported("more than 256 distinct codepoints in the dictionary");
        it = alphabet.emplace(c, (int)alphabet.size()).first;
      }
      sym.push_back((uint8_t)it->second);
    }
  }
  sym.push_back(0);
  h->n_dev_strings = (int)h->strings.size();
  {
    // head-room for strings generated later: appended in place, the pointers never change
    h->sym_used = sym.size() - 1; h->sym_cap = sym.size() + (size_t)h->newstr_cap * PCL_NEWSTR_MAX; h->str_cap = (int)off.size() + h->newstr_cap;
    std::vector<uint8_t> sym2 = sym; sym2.resize(h->sym_cap, 0);
    std::vector<int> off2 = off, len2 = len; off2.resize(h->str_cap, 0); len2.resize(h->str_cap, 0);
    h->d_sym.upload(sym2); h->d_str_off.upload(off2); h->d_str_len.upload(len2);
    h->d_newstr_chars.alloc((size_t)h->newstr_cap * PCL_NEWSTR_MAX); h->d_newstr_len.alloc(h->newstr_cap); h->d_newstr_count.alloc(1); h->d_newstr_count.zero();
    h->d_newstr_map.alloc(h->newstr_cap);
    h->d_lm_sym.upload(std::vector<int>(h->lm_sym, h->lm_sym + 28));
    h->d_lm_uni.upload(std::vector<double>(m.lm_uni, m.lm_uni + 28)); h->d_lm_big.upload(std::vector<double>(m.lm_big, m.lm_big + 28 * 28));
  }
  for (auto& c : h->cols) { c->max_len = 0; for (int s : c->ulist) c->max_len = std::max(c->max_len, len[s]); }
  std::vector<double> LG(PCL_LG_N, 0.0), LOGN(256, 0.0);
  for (int i = 1; i < PCL_LG_N; ++i) LG[i] = std::lgamma((double)i);
  LG[0] = INFINITY;
  for (int i = 1; i < 256; ++i) LOGN[i] = std::log((double)i);
  LOGN[0] = -INFINITY;
  h->d_LG.upload(LG); h->d_LOGN.upload(LOGN);
  {
    // AddTypos score table, same operation order as the device routine (and the oracle)
    std::vector<double> LUT((size_t)PCL_LUT_N * PCL_LUT_N, 0.0);
    for (int L = 1; L < PCL_LUT_N; ++L)
      for (int k = 0; k < PCL_LUT_N; ++k) {
        const int r = (L + 4) / 5;
        volatile double l = LG[k + r] - LG[k + 1];
        l = l - LG[r];
        l = l + (double)r * -0.10536051565782630123;
        l = l + (double)k * -2.30258509299404568402;
        l = l - LOGN[L] * (double)k;
        l = l - (3.25809653802148204862 * (double)k) * 0.5;
        LUT[(size_t)L * PCL_LUT_N + k] = l;
      }
    for (int k = 0; k < PCL_LUT_N; ++k) LUT[k] = k == 0 ? 0.0 : -INFINITY;   // L = 0 never occurs (string_prior min lengths)
    h->d_LUT.upload(LUT);
  }

  // ---- tables
  const int nc = (int)m.classes.size();
  h->h_tables.assign(nc, TableD{});
  for (int c = 0; c < nc; ++c) {
    TableH& T = h->tables[c];
    if (!T.loaded) continue;
    const ClassM& tm = m.classes[c];
    T.n_normal = tm.n_normal;
    T.fk_col.clear(); T.fk_table.clear();
    for (int v = 0; v < tm.n_normal; ++v)
      if (tm.nodes[v].wrap == PCLEAN_WRAP_NONE && tm.nodes[v].kind == PCLEAN_NODE_FK) { T.fk_col.push_back(v); T.fk_table.push_back(tm.nodes[v].target); }
    if (T.fk_col.size() > 4) throw Unsupported("latent class with more than 4 reference slots");
    T.n_slots = (int)T.keys.size();
    // an explicit reservation (pclean_reserve_table) is taken at its word; otherwise room for the table to double
    T.cap = ((T.reserve > 0 ? std::max(T.reserve, T.n_slots + 16) : std::max(T.n_slots * 2 + 1024, T.min_cap)) + 15) / 16 * 16;
  }
  for (int c = 0; c < nc; ++c) {
    TableH& T = h->tables[c];
    if (!T.loaded) continue;
    const ClassM& tm = m.classes[c];
    std::vector<int> cells((size_t)T.n_normal * T.cap, PCL_UNSET);
    const int64_t nr = T.n_slots;
    for (int v = 0; v < std::min(T.n_normal, T.raw_cols); ++v) {
      const Node& nd = tm.nodes[v];
      const bool is_fk = nd.kind == PCLEAN_NODE_FK;
      for (int64_t r = 0; r < nr; ++r) {
        const pclean_value& x = T.raw[(size_t)v * nr + r];
        int out = PCL_UNSET;
        if (x.tag == PCLEAN_VAL_STR) out = x.i;
        else if (x.tag == PCLEAN_VAL_KEY && is_fk) {
          const TableH& TT = h->tables[nd.target];
          auto it = TT.slot_of_key.find((int64_t)x.d);
          if (it == TT.slot_of_key.end()) throw BadArg("table snapshot references a key that is not in the target table");
          out = it->second;
        }
        cells[(size_t)v * T.cap + r] = out;
      }
    }
    T.cells.upload(cells); T.refcnt.alloc(T.cap); T.refcnt.zero(); T.logcnt.alloc(T.cap); T.logcnt1.alloc(T.cap);
    T.alive.alloc(T.cap + 16); T.alive.zero();
    T.div.alloc(std::max(1, T.n_normal)); T.div.zero();
    TableD& D = h->h_tables[c];
    D.cells = T.cells.p; D.refcnt = T.refcnt.p; D.logcnt = T.logcnt.p; D.logcnt1 = T.logcnt1.p; D.alive = T.alive.p; D.max_logcnt = 0.0; D.div = T.div.p;
    {
      std::vector<long long> kk(T.cap, 0);
      for (size_t i = 0; i < T.keys.size(); ++i) kk[i] = T.keys[i];
      T.d_keys.upload(kk); D.keys = T.d_keys.p;
    }
    h->max_cap = std::max(h->max_cap, T.cap);
    D.cap = T.cap; D.n_slots = T.n_slots; D.n_normal = T.n_normal; D.total_refs = 0; D.n_alive = 0;
    D.strength = T.strength; D.discount = T.discount; D.nfk = (int)T.fk_col.size();
    for (size_t g = 0; g < T.fk_col.size(); ++g) { D.fk_col[g] = T.fk_col[g]; D.fk_table[g] = T.fk_table[g]; }
    for (int64_t k : T.keys) h->next_key = std::max(h->next_key, k + 1);
  }
  h->d_tables.alloc(nc);
  h->fk_copies.clear(); h->fk_ncopies.clear();
  for (int c = 0; c < nc; ++c) {
    TableH& T = h->tables[c];
    if (!T.loaded) continue;
    for (size_t g = 0; g < T.fk_col.size(); ++g) {
      const Node& fk = m.classes[c].nodes[T.fk_col[g]];
      std::vector<int2> cp;
      for (size_t tv = 0; tv < fk.vmap.size(); ++tv) cp.push_back(make_int2(fk.vmap[tv], (int)tv));
      std::unique_ptr<DBuf<int2>> b(new DBuf<int2>()); b->upload(cp);
      h->fk_ncopies[{c, (int)g}] = (int)cp.size();
      h->fk_copies[{c, (int)g}] = std::move(b);
    }
  }

  // ---- assignment
  h->d_assign.clear();
  std::vector<int*> aptrs;
  for (int b = 0; b < h->n_blocks; ++b) {
    if (h->progs[b].root < 0) {                                 // keeps assign[] indexed by block
      h->d_assign.emplace_back(new DBuf<int>()); h->d_assign.back()->alloc(1); aptrs.push_back(nullptr);
      continue;
    }
    const StarL& root = h->progs[b].stars[h->progs[b].root];
    auto it = h->assign_keys.find(root.vertex);
    if (it == h->assign_keys.end()) throw std::runtime_error("pclean_load_assignment: missing reference slot of a block root");
    const TableH& T = h->tables[root.table];
    if (!T.loaded) throw std::runtime_error("pclean_load_table: a referenced latent table was not loaded");
    std::vector<int> slots(h->N);
    for (int64_t r = 0; r < h->N; ++r) {
      if (it->second[r] == INT64_MIN) { slots[r] = -1; continue; }        // row not initialised yet (pclean_init_trace)
      auto sk = T.slot_of_key.find(it->second[r]);
      if (sk == T.slot_of_key.end()) throw BadArg("assignment references a key that is not in the table");
      slots[r] = sk->second;
    }
    h->d_assign.emplace_back(new DBuf<int>()); h->d_assign.back()->upload(slots);
    aptrs.push_back(h->d_assign.back()->p);
  }
  h->d_assign_ptrs.upload(aptrs);

  // ---- parameters (initialize_parameter: keyed prior draws, include/pclean_rng.h)
  if (h->params.size() != m.slot_param.size()) h->params.assign(m.slot_param.size(), ParamH());
  for (size_t s = 0; s < h->params.size(); ++s) h->params[s].spec = m.slot_param[s];

  // ---- flatten programs (observation-class blocks first, then one program per latent class)
  h->h_progs.clear(); h->h_stars.clear(); h->h_terms.clear(); h->h_children.clear(); h->h_copies.clear();
  h->h_prior.clear(); h->h_optsid.clear(); h->joins.clear(); h->hoists.clear(); h->mats.clear();
  h->cand_mats.clear(); h->opt_mats.clear(); h->univ_cache.clear(); h->lmats.clear(); h->lmat_of.clear(); h->lprog_of_class.clear(); h->lprog_of_pat.clear(); h->ref_chain.clear(); h->h_gext.clear();
  h->h_mswaps.clear(); h->h_fills.clear(); h->h_lkconst.clear(); h->prog_rootless.clear();
  struct PendingMat { int mat; int opt_off; int nopt; };
  std::vector<PendingMat> pending_opt;
  std::map<std::pair<int, int>, int> opt_pool;            // (list id, dummy string) -> offset into the option pool
  // a cell of the referring observation row: which block's reference slot reaches it
  auto refcell = [&](int v) {
    RefCellD rc{-1, -1, -1};
    const Node& an = cm.nodes[v];
    for (int b2 = 0; b2 < h->n_blocks; ++b2) {
      if (h->progs[b2].root < 0) continue;
      const StarL& r2 = h->progs[b2].stars[h->progs[b2].root];
      if (!an.wfk.empty() && an.wfk[0] == r2.vertex) { rc.block = b2; rc.col = an.wsub[0]; rc.table = r2.table; }
    }
    if (rc.block < 0) throw Unsupported("referring-row value that is not a cell of a top-level reference slot");
    return rc;
  };
  // tabulated function -> device hash table (same hash as lookup_find in device.cuh)
  auto lookup_index = [&](int func) {
    auto it = h->lookup_of_func.find(func);
    if (it != h->lookup_of_func.end()) return it->second;
    const FuncM& f = m.funcs.at(func);
    if (f.kind != PCLEAN_FUNC_TABLE || f.keyargs.size() > 3) throw Unsupported("lookup of a function that is not a table over <= 3 key arguments");
    size_t cap = 16; while (cap < f.table.size() * 2 + 2) cap <<= 1;
    std::vector<int> keys(cap * 3, PCL_LOOKUP_EMPTY), vals(cap, 0);
    auto hmix = [](unsigned long long x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; };
    for (auto& kv : f.table) {
      int k[3] = {0, 0, 0};
      for (size_t i = 0; i < kv.first.size(); ++i) k[i] = kv.first[i];
      int val;
      if (kv.second.tag == PCLEAN_VAL_PARAM || kv.second.tag == PCLEAN_VAL_LIST || kv.second.tag == PCLEAN_VAL_STR || kv.second.tag == PCLEAN_VAL_INT) val = kv.second.i;
      else if (kv.second.tag == PCLEAN_VAL_REAL) {          // a real constant (flights: error_prob = 1e-5): code -2 - index into the constant pool
        size_t ci = 0;
        while (ci < h->h_lkconst.size() && h->h_lkconst[ci] != kv.second.d) ++ci;
        if (ci == h->h_lkconst.size()) h->h_lkconst.push_back(kv.second.d);
        val = -2 - (int)ci;
      }
      else throw Unsupported("tabulated function returning a value that is neither a parameter, a list, a string nor a real");
      const unsigned long long key = hmix((unsigned long long)(unsigned)k[0] * 0x9E3779B97F4A7C15ULL ^ ((unsigned long long)(unsigned)k[1] << 20) ^ ((unsigned long long)(unsigned)k[2] << 41));
      size_t hh = (size_t)((unsigned)key & (unsigned)(cap - 1));
      while (keys[3 * hh] != PCL_LOOKUP_EMPTY) hh = (hh + 1) & (cap - 1);
      keys[3 * hh] = k[0]; keys[3 * hh + 1] = k[1]; keys[3 * hh + 2] = k[2]; vals[hh] = val;
    }
    std::unique_ptr<pclean_engine::LookupH> L(new pclean_engine::LookupH());
    L->keys.upload(keys); L->vals.upload(vals); L->mask = (unsigned)(cap - 1); L->nkey = (int)f.keyargs.size();
    h->lookups.push_back(std::move(L));
    const int idx = (int)h->lookups.size() - 1;
    h->lookup_of_func[func] = idx;
    return idx;
  };
  auto dataset_col = [&](int vertex) {
    auto cit = h->col_of_vertex.find(vertex);
    if (cit == h->col_of_vertex.end()) throw Unsupported("value of a vertex that is not a dataset column");
    return cit->second;
  };
  auto lower_arg = [&](const ArgL& a) {
    InnerArgD d{a.kind, a.ref};
    if (a.kind == ARG_OBS) d.ref = dataset_col(a.ref);
    return d;
  };
  std::map<int, int> optmap_of_list;       // list id -> offset into optmap_pool
  std::map<int, int> optidx_of_off;        // option-pool offset of a constant list -> offset into optmap_pool of its (string -> option index) map
  auto lower_inner = [&](const InnerL& in, std::vector<int>& local_vertices) {
    if (in.empty()) return -1;
    if (in.choices.size() > PCL_MAX_INNER_CH || in.gauss.size() > 2 || in.consts.size() > 3) throw Unsupported("inner enumeration too large");
    InnerD I{};
    I.nchoice = (int)in.choices.size(); I.ngauss = (int)in.gauss.size(); I.nconst = (int)in.consts.size();
    for (int i = 0; i < I.nchoice; ++i) {
      const InnerChoiceL& c = in.choices[i];
      const std::vector<Val>& vals = m.lists.at(c.list);
      if (vals.size() > 16) throw Unsupported("inner choice with more than 16 options");
      I.ch[i].vertex = c.vertex; I.ch[i].list_off = (int)h->h_innervals.size(); I.ch[i].n = (int)vals.size();
      for (const Val& v : vals) h->h_innervals.push_back(v.i);
      if (std::find(local_vertices.begin(), local_vertices.end(), c.vertex) == local_vertices.end()) local_vertices.push_back(c.vertex);
    }
    for (int g = 0; g < I.ngauss; ++g) {
      const GaussL& G = in.gauss[g];
      InnerGaussD& D = I.g[g];
      D.obs_col = dataset_col(G.obs_vertex);
      if (!h->cols[D.obs_col]->is_real) throw Unsupported("Gaussian likelihood on a non-numeric column");
      D.func = G.mean_func >= 0 ? lookup_index(G.mean_func) : G.mean_func; D.nargs = G.n_mean_args;
      for (int a2 = 0; a2 < G.n_mean_args; ++a2) D.args[a2] = lower_arg(G.mean_args[a2]);
      D.mean_const = G.mean_const; D.stdev = G.stdev; D.xform = lower_arg(G.xform);
    }
    for (int k = 0; k < I.nconst; ++k) {
      const ConstPriorL& C = in.consts[k];
      InnerConstD& D = I.c[k];
      D.kind = C.kind; D.value = C.value; D.obs_col = -1; D.optmap = -1; D.logp_off = -1;
      if (C.kind == 2) {        // StringPrior.logdensity of the observed string: table over the dictionary (only this column's strings are filled)
        D.obs_col = dataset_col(C.obs_vertex);
        D.logp_off = (int)h->h_splp.size();
        h->h_splp.resize(h->h_splp.size() + h->strings.size(), 0.0);
        for (int sid : h->cols[D.obs_col]->ulist) h->h_splp[D.logp_off + sid] = stringprior_logdensity(m, h->strings[sid], C.list, C.slot);
      }
      if (C.kind == 1) {
        D.obs_col = dataset_col(C.obs_vertex);
        const std::vector<Val>& opts = m.lists.at(C.list);
        auto om = optmap_of_list.find(C.list);
        if (om == optmap_of_list.end()) {
          const int off = (int)h->h_optmap.size();
          h->h_optmap.resize(off + h->strings.size(), -1);
          for (size_t i = 0; i < opts.size(); ++i) if (opts[i].tag == PCLEAN_VAL_STR && h->h_optmap[off + opts[i].i] < 0) h->h_optmap[off + opts[i].i] = (int)i;
          om = optmap_of_list.emplace(C.list, off).first;
        }
        D.optmap = om->second;
        ParamH& PR = h->params.at(C.slot);
        if (PR.value.empty()) {
          pclean_stream st{}; st.key.seed = h->param_seed; st.key.row = C.slot; st.key.purpose = PCLEAN_RNG_PARAM_INIT;
          PR.value.resize(opts.size()); double tot = 0;
          for (double& v : PR.value) { v = pclean_next_gamma(&st, m.param_prior0[PR.spec]); tot += v; }
          for (double& v : PR.value) v /= tot;
        }
        if (PR.value.size() != opts.size()) throw BadArg("proportions parameter has the wrong length");
        D.logp_off = (int)h->h_prior.size(); PR.prior_offs.push_back(D.logp_off); PR.nopt = (int)opts.size();
        for (double v : PR.value) h->h_prior.push_back(std::log(v));
      }
    }
    h->h_inners.push_back(I);
    return (int)h->h_inners.size() - 1;
  };
  auto trace_arg_of = [&](int v) {
    const Node& n = cm.nodes[v];
    TraceArgD a{0, 0, 0, 0};
    if (n.wrap != PCLEAN_WRAP_NONE) { const RefCellD rc = refcell(v); a.kind = 2; a.a = rc.block; a.b = rc.table; a.c = rc.col; return a; }
    if (n.kind == PCLEAN_NODE_JULIA && m.funcs[n.func].kind == PCLEAN_FUNC_CONST) { a.kind = 0; a.a = m.funcs[n.func].cst.i; return a; }
    if (n.kind == PCLEAN_NODE_CHOICE) { auto cit = h->col_of_vertex.find(v); a.kind = 1; a.a = cit == h->col_of_vertex.end() ? -1 : cit->second; a.b = v; return a; }
    throw Unsupported("MeanParameter statistics: argument that is neither a constant, a choice nor a reference-table cell");
  };
  auto lookup_ref_of = [&](const LookupL& L, bool latent_prog_) {
    LookupRefD R{}; R.lookup = -1; R.nargs = 0;
    if (L.func < 0) return R;
    R.lookup = lookup_index(L.func);
    if (L.args.size() > 3) throw Unsupported("lookup with more than 3 key arguments");
    for (const ArgL& a : L.args) {
      TraceArgD t{0, 0, 0, 0};
      switch (a.kind) {
        case ARG_CONST: t = TraceArgD{0, a.ref, 0, 0}; break;
        case ARG_OBS:
          if (latent_prog_) t = TraceArgD{5, a.ref, 0, 0};
          else { auto cit = h->col_of_vertex.find(a.ref); t = TraceArgD{1, cit == h->col_of_vertex.end() ? -1 : cit->second, a.ref, 0}; }
          break;
        case ARG_REFROW: t = trace_arg_of(a.ref); break;
        case ARG_ELEM_OPT: t = TraceArgD{3, 0, 0, 0}; break;
        default: throw Unsupported("lookup argument kind");
      }
      R.args[R.nargs++] = t;
    }
    return R;
  };
  auto mswap_of = [&](const MswapL& ms, bool latent_prog_) {
    MswapD M{};
    auto cit = h->col_of_vertex.find(ms.obs_vertex);
    M.obs_col = cit == h->col_of_vertex.end() ? -1 : cit->second;
    M.vertex = ms.obs_vertex;
    M.val_kind = ms.val_kind; M.val_vertex = ms.val_vertex; M.val_cell = RefCellD{-1, -1, -1};
    if (ms.val_kind == 0) M.val_cell = refcell(ms.val_vertex);
    M.list_const = ms.list_const; M.list = lookup_ref_of(ms.list, latent_prog_);
    M.prob_kind = ms.prob_kind; M.prob_const = ms.prob_const; M.prob_slot = ms.prob_slot; M.prob = lookup_ref_of(ms.prob, latent_prog_);
    return M;
  };
  // referrer group keys pack (slot, string id, unique-string index) into 20 + 22 + 22 bits
  const bool groups_ok = h->N < (1ll << 31) - 2 && h->strings.size() < (size_t)(1 << 22) - 2;
  auto flatten = [&](const BlockProgram& bp, int b, int latent_cls, int prog_id, int base_prog) {
    if ((int)bp.stars.size() > PCL_MAX_STARS || (int)bp.terms.size() > PCL_MAX_TERMS) throw Unsupported("block program too large");
    ProgD P{};
    P.latent = latent_cls >= 0; P.cls = latent_cls >= 0 ? latent_cls : h->obs_cls;
    P.base_prog = base_prog; P.n_local = 0;
    std::vector<int> local_vertices;
    P.nroots = (int)bp.roots.size();
    if (P.nroots > PCL_MAX_SITES) throw Unsupported("latent block with too many independent sites");
    for (int i = 0; i < P.nroots; ++i) P.roots[i] = bp.roots[i];
    P.ms0 = 0; P.n_rterm = 0; P.n_rsamp = 0;
    if ((int)h->prog_rootless.size() <= prog_id) h->prog_rootless.resize(prog_id + 1, 0);
    h->prog_rootless[prog_id] = bp.rootless ? 1 : 0;
    if (bp.rootless) {
      // no enumeration: observed MaybeSwap terms, then the absent ones in ascending vertex order (their local cell index)
      P.nroots = 0; P.nstar = 0; P.root = -1; P.norder = 0; P.star0 = (int)h->h_stars.size(); P.term0 = (int)h->h_terms.size(); P.nterm = 0; P.n_earlier = 0;
      P.ms0 = (int)h->h_mswaps.size(); P.n_rterm = (int)bp.root_terms.size(); P.n_rsamp = (int)bp.root_sampled.size();
      for (int i : bp.root_terms) h->h_mswaps.push_back(mswap_of(bp.mswaps[i], false));
      std::vector<int> samp = bp.root_sampled;
      std::sort(samp.begin(), samp.end(), [&](int x, int y) { return bp.mswaps[x].obs_vertex < bp.mswaps[y].obs_vertex; });
      if (samp.size() > PCL_MAX_LOCAL) throw Unsupported("too many sampled cells in a block");
      P.n_local = (int)samp.size();
      for (size_t q = 0; q < samp.size(); ++q) { h->h_mswaps.push_back(mswap_of(bp.mswaps[samp[q]], false)); P.local_vertex[q] = bp.mswaps[samp[q]].obs_vertex; }
      if ((int)h->prog_rich.size() <= prog_id) h->prog_rich.resize(prog_id + 1, 1);
      h->h_progs.push_back(P);
      return;
    }
    P.nstar = (int)bp.stars.size(); P.root = bp.root; P.norder = (int)bp.order.size();
    for (int i = 0; i < P.norder; ++i) P.order[i] = bp.order[i];
    P.star0 = (int)h->h_stars.size(); P.term0 = (int)h->h_terms.size(); P.nterm = (int)bp.terms.size();
    if (bp.earlier_vertices.size() > 1) throw Unsupported("block depending on more than one earlier-block value");
    P.n_earlier = (int)bp.earlier_vertices.size();
    if (P.n_earlier) {
      const int av = *bp.earlier_vertices.begin();
      P.earlier_vertex = av; P.earlier_block = -1;
      const Node& an = cm.nodes[av];
      for (int b2 = 0; b2 < b; ++b2) {
        if (h->progs[b2].root < 0) continue;
        const StarL& r2 = h->progs[b2].stars[h->progs[b2].root];
        if (!an.wfk.empty() && an.wfk[0] == r2.vertex) { P.earlier_block = b2; P.earlier_col = an.wsub[0]; P.earlier_table = r2.table; }
      }
      if (P.earlier_block < 0) throw Unsupported("earlier-block value that is not a cell of an earlier reference slot");
    }
    // stars
    std::map<int, std::pair<int, int>> univ_ids;      // star -> (offset, count) of its option universe in the id pool
    for (size_t si = 0; si < bp.stars.size(); ++si) {
      const StarL& s = bp.stars[si];
      StarD D{};
      D.kind = s.kind; D.vertex = s.vertex; D.parent = s.parent; D.table = s.table; D.tvertex = s.tvertex;
      D.term0 = -1; D.nterm = 0; D.hoist = -1; D.hoist_col = -1;
      D.list_func = -1; D.list_obs_col = -1; D.list_own_col = -1; D.splp_off = -1; D.univ_off = -1; D.inner_elems = -1; D.inner_new = -1; D.optidx_off = -1;
      D.child0 = (int)h->h_children.size(); D.nchild = (int)s.children.size();
      for (int c : s.children) h->h_children.push_back(c);
      D.copy0 = (int)h->h_copies.size(); D.ncopy = (int)s.copies.size();
      for (auto& pr : s.copies) h->h_copies.push_back(make_int2(pr.first, pr.second));
      if (s.kind == ST_CHOICE && s.list < 0) {
        // option list looked up from an observed value (rents: possibilities[countykey]): the universe of
        // all lists the function can return gets one distance matrix; per-string prior table
        // (every program that meets the same list function — one per block, missingness pattern and latent
        // class — shares the universe, its prior table and, through the pool offset, its distance matrices)
        const FuncM& lf = m.funcs.at(s.list_func);
        D.list_func = lookup_index(s.list_func);
        D.nopt = 0; D.has_dummy = 1; D.prior_off = -1;
        const auto ukey = std::make_tuple(s.list_func, s.dummy_string, (int)s.dist, s.sp_min, s.sp_max);
        auto uit = h->univ_cache.find(ukey);
        if (uit == h->univ_cache.end()) {
          std::vector<int> universe; std::vector<char> in_univ(h->strings.size(), 0);
          for (auto& kv : lf.table) {
            if (kv.second.tag != PCLEAN_VAL_LIST) throw Unsupported("option-list lookup returning a non-list");
            for (const Val& o : m.lists.at(kv.second.i)) { if (o.tag != PCLEAN_VAL_STR) throw Unsupported("choice over non-string options"); if (!in_univ[o.i]) { in_univ[o.i] = 1; universe.push_back(o.i); } }
          }
          if (!in_univ[s.dummy_string]) { in_univ[s.dummy_string] = 1; universe.push_back(s.dummy_string); }
          UnivEntry ue{};
          ue.opt_off = (int)h->h_optsid.size();
          h->h_optsid.push_back(s.dummy_string);
          while (h->h_optsid.size() % 4) h->h_optsid.push_back(-1);
          ue.ids_off = (int)h->h_optsid.size(); ue.count = (int)universe.size();
          for (int sid : universe) h->h_optsid.push_back(sid);
          while (h->h_optsid.size() % 4) h->h_optsid.push_back(-1);
          ue.splp_off = (int)h->h_splp.size();
          h->h_splp.resize(h->h_splp.size() + h->strings.size(), 0.0);
          ue.univ_off = (int)h->h_univ.size();
          h->h_univ.resize(h->h_univ.size() + h->strings.size(), -1);
          for (size_t ui = 0; ui < universe.size(); ++ui) {
            h->h_univ[ue.univ_off + universe[ui]] = (int)ui;
            h->h_splp[ue.splp_off + universe[ui]] = s.dist == PCLEAN_DIST_TIME_PRIOR
                ? (time_regex(h->strings[universe[ui]]) ? -std::log(1440.0) : -INFINITY)                  // time_prior.jl:8-14
                : stringprior_logdensity(m, h->strings[universe[ui]], s.sp_min, s.sp_max);
          }
          uit = h->univ_cache.emplace(ukey, ue).first;
        }
        const UnivEntry& ue = uit->second;
        D.opt_off = ue.opt_off; D.splp_off = ue.splp_off; D.univ_off = ue.univ_off;
        univ_ids[(int)si] = std::make_pair(ue.ids_off, ue.count);
        if (latent_cls >= 0) { D.list_obs_col = -1; D.list_own_col = s.list_arg.ref; }
        else D.list_obs_col = dataset_col(s.list_arg.ref);
      } else if (s.kind == ST_CHOICE) {
        const std::vector<Val>& opts = m.lists.at(s.list);
        auto pk = std::make_pair(s.list, s.has_dummy ? s.dummy_string : -1);
        auto pit = opt_pool.find(pk);
        if (pit == opt_pool.end()) {
          const int off = (int)h->h_optsid.size();
          for (const Val& o : opts) h->h_optsid.push_back(o.i);
          if (s.has_dummy) h->h_optsid.push_back(s.dummy_string);
          while (h->h_optsid.size() % 4) h->h_optsid.push_back(-1);
          pit = opt_pool.emplace(pk, off).first;
        }
        D.opt_off = pit->second;
        D.nopt = (int)opts.size() + (s.has_dummy ? 1 : 0); D.has_dummy = s.has_dummy;
        if (latent_cls >= 0 && D.nopt > 2 * PCL_SURV_MAX) {
          // long constant option list enumerated by a latent move: option index of every dictionary string
          // (the pruned evaluation starts from the option equal to the row's current / most observed string)
          auto oi = optidx_of_off.find(D.opt_off);
          if (oi == optidx_of_off.end()) {
            const int off = (int)h->h_optmap.size();
            h->h_optmap.resize(off + h->strings.size(), -1);
            for (size_t i = 0; i < opts.size(); ++i) if (opts[i].tag == PCLEAN_VAL_STR && h->h_optmap[off + opts[i].i] < 0) h->h_optmap[off + opts[i].i] = (int)i;
            oi = optidx_of_off.emplace(D.opt_off, off).first;
          }
          D.optidx_off = oi->second;
        }
        D.prior_off = (int)h->h_prior.size();
        std::vector<double> lp;
        if (s.dist == PCLEAN_DIST_STRING_PRIOR) {
          for (const Val& o : opts) lp.push_back(stringprior_logdensity(m, h->strings[o.i], s.sp_min, s.sp_max));
          lp.push_back(std::log1p(-std::exp(lse_host(lp))));                         // string_prior.jl:19-20
        } else if (s.dist == PCLEAN_DIST_TIME_PRIOR) {
          for (const Val& o : opts) lp.push_back(time_regex(h->strings[o.i]) ? -std::log(1440.0) : -INFINITY);
          lp.push_back(std::log1p(-std::exp(lse_host(lp))));
        } else if (s.prior_kind == PRIOR_PROPORTIONS) {
          ParamH& PR = h->params.at(s.prior_slot);
          PR.prior_offs.push_back(D.prior_off); PR.nopt = D.nopt;
          if (PR.value.empty()) {           // first param_value: Dirichlet draw (choose_proportionally.jl:48-55)
            pclean_stream st{}; st.key.seed = h->param_seed; st.key.row = s.prior_slot; st.key.purpose = PCLEAN_RNG_PARAM_INIT;
            PR.value.resize(D.nopt); double tot = 0;
            for (double& v : PR.value) { v = pclean_next_gamma(&st, m.param_prior0[PR.spec]); tot += v; }
            for (double& v : PR.value) v /= tot;
          }
          if ((int)PR.value.size() != D.nopt) throw BadArg("proportions parameter has the wrong length");
          for (double v : PR.value) lp.push_back(std::log(v));
        } else lp = s.static_prior;
        if ((int)lp.size() != D.nopt) throw std::runtime_error("internal: prior length mismatch");
        h->h_prior.insert(h->h_prior.end(), lp.begin(), lp.end());
      }
      D.bucket = s.bucket ? 1 : 0; D.bucket_col = s.bucket_col; D.bucket_obs_col = s.bucket ? dataset_col(s.bucket_obs_vertex) : -1;
      if (s.bucket) {
        if ((int)h->bucket_col_of_table.size() < nc) h->bucket_col_of_table.assign(nc, -1);
        if (h->bucket_col_of_table[s.table] >= 0 && h->bucket_col_of_table[s.table] != s.bucket_col) throw Unsupported("table bucketed on two different keys");
        h->bucket_col_of_table[s.table] = s.bucket_col;
      }
      D.dummy_time = (s.kind == ST_CHOICE && s.dist == PCLEAN_DIST_TIME_PRIOR) ? 1 : 0;
      D.sp_min = s.sp_min; D.sp_max = s.sp_max;
      D.has_eq = 0;
      for (int ti : s.terms) if (bp.terms[ti].kind == TERM_EQ) D.has_eq = 1;
      D.fill0 = (int)h->h_fills.size(); D.nfill = (int)s.fillins.size();
      for (const StarL::FillL& f : s.fillins) {
        FillD F{}; F.vertex = f.vertex; F.dist = f.dist; F.list_const = f.list; F.dummy_sid = f.dummy_string;
        LookupL L; L.func = f.list_func; if (f.list_func >= 0) L.args.push_back(f.list_arg);
        F.list = lookup_ref_of(L, latent_cls >= 0);
        h->h_fills.push_back(F);
      }
      D.inner_elems = lower_inner(s.inner_elems, local_vertices);
      D.inner_new = lower_inner(s.inner_new, local_vertices);
      h->h_stars.push_back(D);
    }
    if (local_vertices.size() > PCL_MAX_INNER_CH) throw Unsupported("too many local choices in a block");
    for (int q = 0; q < PCL_MAX_LOCAL; ++q) P.local_vertex[q] = -1;
    std::sort(local_vertices.begin(), local_vertices.end());
    P.n_local = (int)local_vertices.size();
    for (int q = 0; q < P.n_local; ++q) P.local_vertex[q] = local_vertices[q];
    // terms, grouped per star (contiguous)
    for (size_t si = 0; si < bp.stars.size(); ++si) {
      const StarL& s = bp.stars[si];
      StarD& D = h->h_stars[P.star0 + si];
      D.term0 = (int)h->h_terms.size() - P.term0; D.nterm = (int)s.terms.size();
      for (int ti : s.terms) {
        const TermL& t = bp.terms[ti];
        TermD T{}; T.kind = t.kind; T.max_typos = t.max_typos; T.external = t.external ? 1 : 0;
        T.ptable = -1; T.pcol = -1; T.lmat = -1;
        if (t.kind == TERM_CAND || t.kind == TERM_JOIN_CAND) { T.ptable = s.table; T.pcol = t.col; }
        T.a_kind = t.a_kind; T.a_ref = t.a_ref; T.b_kind = t.b_kind; T.b_ref = t.b_ref; T.sep = t.sep;
        auto cit = h->col_of_vertex.find(t.obs_vertex);
        if (cit == h->col_of_vertex.end()) throw std::runtime_error("internal: term on a non-dataset vertex");
        T.obs_col = cit->second;
        const int U = (int)h->cols[T.obs_col]->ulist.size();
        if (t.kind == TERM_EQ) { T.mat = t.col; h->h_terms.push_back(T); continue; }
        if (t.kind == TERM_MSWAP_EXT) {
          T.mat = (int)h->h_mswaps.size();
          h->h_mswaps.push_back(mswap_of(bp.mswaps.at(t.mswap), latent_cls >= 0));
          h->h_terms.push_back(T); continue;
        }
        if (t.kind == TERM_GAUSS_EXT) {
          const GaussL& g = bp.gauss_ext.at(t.gauss);
          GaussExtD G{};
          G.obs_col = T.obs_col;
          if (!h->cols[G.obs_col]->is_real) throw Unsupported("Gaussian likelihood on a non-numeric column");
          G.lookup = lookup_index(g.mean_func); G.nargs = g.n_mean_args; G.stdev = g.stdev;
          auto ext_arg_of = [&](const ArgL& a) {
            switch (a.kind) {
              case ARG_CONST: return TraceArgD{0, a.ref, 0, 0};
              case ARG_OBS: return TraceArgD{5, a.ref, 0, 0};          // latent mode: the moved row's own cell
              case ARG_ELEM_COL: return TraceArgD{4, a.ref, 0, 0};
              case ARG_ELEM_OPT: return TraceArgD{3, 0, 0, 0};
              case ARG_REFROW: return trace_arg_of(a.ref);
              default: throw Unsupported("external lookup argument kind");
            }
          };
          for (int a2 = 0; a2 < G.nargs; ++a2) G.args[a2] = ext_arg_of(g.mean_args[a2]);
          G.xform = ext_arg_of(g.xform);
          T.mat = (int)h->h_gext.size();
          h->h_gext.push_back(G);
          h->h_terms.push_back(T); continue;
        }
        if (t.kind == TERM_OPT && univ_ids.count((int)si)) {
          // a choice over a row-dependent option list: distances live in per-list blocks (ListMatD), not in one
          // matrix over the union of all lists.  The block rows come from the dataset: the observation-class
          // program knows which column selects the list; a latent program reuses the blocks built for it.
          auto key = std::make_tuple(T.obs_col, univ_ids[(int)si].first);
          auto lit = h->lmat_of.find(key);
          if (lit == h->lmat_of.end()) {
            const int key_col = latent_cls < 0 ? dataset_col(s.list_arg.ref) : -1;
            lit = h->lmat_of.emplace(key, plan_list_mat(h, T.obs_col, s.list_func, key_col, s.dummy_string)).first;
          }
          T.lmat = lit->second; T.mat = -1; h->h_terms.push_back(T); continue;
        }
        if (t.kind == TERM_CAND) {
          auto key = std::make_tuple(T.obs_col, s.table, t.col);
          auto mit = h->cand_mats.find(key);
          if (mit == h->cand_mats.end()) {
            const int mi = new_mat(h, T.obs_col, U, h->tables[s.table].cap);
            h->mats[mi]->table = s.table; h->mats[mi]->col = t.col;
            mit = h->cand_mats.emplace(key, mi).first;
          }
          T.mat = mit->second;
        } else if (t.kind == TERM_OPT) {
          auto key = std::make_tuple(T.obs_col, D.opt_off);
          auto mit = h->opt_mats.find(key);
          if (mit == h->opt_mats.end()) {
            const int mi = new_mat(h, T.obs_col, U, D.nopt);
            pending_opt.push_back({mi, D.opt_off, D.nopt});
            mit = h->opt_mats.emplace(key, mi).first;
          }
          T.mat = mit->second;
        } else if (t.kind == TERM_JOIN_INLINE) {
          if (t.a_kind == OP_REFROW) T.a_cell = refcell(t.a_ref);
          if (t.b_kind == OP_REFROW) T.b_cell = refcell(t.b_ref);
          T.mat = -1;
        } else {
          JoinTerm J{}; J.prog = prog_id; J.term = (int)h->h_terms.size() - P.term0; J.kind = t.kind; J.obs_col = T.obs_col;
          J.table = s.table; J.col = t.col; J.opt_off = D.opt_off; J.nopt = D.nopt; J.sep = t.sep;
          T.mat = (int)h->joins.size();
          h->joins.push_back(J);
        }
        h->h_terms.push_back(T);
      }
      // hoisting: a choice star with a single option-indexed term depends on the row only
      // through that column's unique observed string
      if (latent_cls < 0 && s.kind == ST_CHOICE && s.list >= 0 && s.inner_elems.empty() && s.terms.size() == 1 && bp.terms[s.terms[0]].kind == TERM_OPT) {
        Hoist H; H.prog = prog_id; H.star = (int)si; H.obs_col = h->h_terms.back().obs_col;
        H.val.reset(new DBuf<double>()); H.val->alloc(h->cols[H.obs_col]->ulist.size());
        H.dynamic = s.prior_kind == PRIOR_PROPORTIONS;
        D.hoist = (int)h->hoists.size(); D.hoist_col = H.obs_col;
        h->hoists.push_back(std::move(H));
      }
    }
    bool rich = false;
    for (const StarL& s2 : bp.stars) rich = rich || s2.bucket || s2.list_func >= 0 || !s2.inner_elems.empty() || !s2.inner_new.empty() || !s2.fillins.empty();
    for (const TermL& t2 : bp.terms) rich = rich || t2.kind == TERM_EQ || t2.kind == TERM_GAUSS_EXT;
    if ((int)h->prog_rich.size() <= prog_id) h->prog_rich.resize(prog_id + 1, 1);
    h->prog_rich[prog_id] = rich ? 1 : 0;
    // latent programs: which referrer group set each external string term sums over (latent.cuh)
    for (int ti = P.term0; ti < (int)h->h_terms.size(); ++ti) {
      TermD& T = h->h_terms[ti];
      T.grp = -1;
      if (latent_cls < 0 || !T.external || !groups_ok) continue;
      GroupSetD G{}; G.obs_col = T.obs_col; G.has_ref = 0; G.ref = RefCellD{-1, -1, -1};
      if (T.kind == TERM_JOIN_INLINE) {
        if ((T.a_kind == OP_REFROW) == (T.b_kind == OP_REFROW)) continue;      // both or neither half from the referring row: not grouped
        G.has_ref = 1; G.ref = T.a_kind == OP_REFROW ? T.a_cell : T.b_cell;
      } else if (T.kind != TERM_CAND && T.kind != TERM_OPT) continue;
      int id = -1;
      for (size_t g = 0; g < h->gsets.size(); ++g) {
        const GroupSetD& X = h->gsets[g];
        if (X.obs_col == G.obs_col && X.has_ref == G.has_ref && X.ref.block == G.ref.block && X.ref.col == G.ref.col && X.ref.table == G.ref.table) id = (int)g;
      }
      if (id < 0) { id = (int)h->gsets.size(); h->gsets.push_back(G); }
      T.grp = id;
      if ((int)h->gsets_of_class.size() <= latent_cls) h->gsets_of_class.resize(latent_cls + 1);
      std::vector<int>& L = h->gsets_of_class[latent_cls];
      if (std::find(L.begin(), L.end(), id) == L.end()) L.push_back(id);
    }
    h->h_progs.push_back(P);
  };
  for (auto& PR : h->params) PR.prior_offs.clear();
  h->gsets.clear(); h->gsets_of_class.clear();

  h->h_inners.clear(); h->h_innervals.clear(); h->lookup_of_func.clear(); h->lookups.clear();
  h->h_splp.clear(); h->h_univ.clear(); h->h_optmap.clear(); h->bucket_col_of_table.assign(nc, -1);
  for (int pt = 0; pt < h->n_patterns; ++pt)
    for (int b = 0; b < h->n_blocks; ++b) flatten(h->progs[pt * h->n_blocks + b], b, -1, pt * h->n_blocks + b, pt * h->n_blocks);
  // latent classes: flatten the programs lowered above
  for (size_t li = 0; li < h->lprogs.size(); ++li) {
    const int c = h->lprog_cls[li];
    try {
      // the one chain of reference slots from the observed class down to this class
      int chain_path = -1;
      for (int pth = 0; pth < h->ir_n_paths; ++pth) {
        if (h->ir_path_target[pth] != c) continue;
        const int l1 = h->ir_path_len_off[pth + 1];
        if (h->ir_path_class[l1 - 1] != h->obs_cls) continue;
        if (chain_path >= 0) throw Unsupported("latent class reachable from the observed class through several paths");
        chain_path = pth;
      }
      if (chain_path < 0) throw Unsupported("latent class not reachable from the observed class");
      RefChainD ch{}; ch.n_links = 0; ch.block0 = -1;
      const int l0 = h->ir_path_len_off[chain_path], l1 = h->ir_path_len_off[chain_path + 1];
      const int topv = h->ir_path_vertex[l1 - 1];
      for (int b2 = 0; b2 < h->n_blocks; ++b2) if (h->progs[b2].root >= 0 && h->progs[b2].stars[h->progs[b2].root].vertex == topv) ch.block0 = b2;
      if (ch.block0 < 0) throw Unsupported("reference chain does not start at a block root");
      for (int l = l1 - 2; l >= l0; --l) {
        if (ch.n_links >= 4) throw Unsupported("reference chain longer than 4 links");
        ch.table[ch.n_links] = h->ir_path_class[l]; ch.col[ch.n_links] = h->ir_path_vertex[l]; ++ch.n_links;
      }
      const int pid = (int)h->h_progs.size();
      flatten(h->lprogs[li], 0, c, pid, pid);
      if (h->lprog_mask[li] == 0 || !h->lprog_of_class.count(c)) h->lprog_of_class[c] = pid;
      h->lprog_of_pat[std::make_pair(c, h->lprog_mask[li])] = pid;
      h->ref_chain[c] = ch;
    } catch (const Unsupported& e) { h->lprog_pat_error[std::make_pair(c, h->lprog_mask[li])] = e.what(); if (h->lobs_cols[c].empty()) h->lprog_error[c] = e.what(); }
  }
  // which dataset columns set which observed-cell bit of which latent class
  h->lobs_cells.clear();
  for (auto& kv : h->lobs_cols) {
    if (kv.second.empty()) continue;
    ObsCellsD oc{}; oc.n = 0;
    for (size_t ci = 0; ci < h->cols.size(); ++ci) {
      const Node& n = cm.nodes[h->cols[ci]->vertex];
      if (n.wrap != PCLEAN_WRAP_SUBMODEL || n.wfk.size() != 1 || cm.nodes[n.wfk[0]].target != kv.first) continue;
      const RefCellD rc = refcell(h->cols[ci]->vertex);
      if (oc.n >= 8) throw Unsupported("more than 8 directly observed cells of one latent class");
      oc.data_col[oc.n] = (int)ci; oc.block[oc.n] = rc.block;
      oc.bit[oc.n] = (int)(std::find(kv.second.begin(), kv.second.end(), n.wsub[0]) - kv.second.begin());
      ++oc.n;
    }
    h->lobs_cells[kv.first] = oc;
  }
  h->d_gext.upload(h->h_gext);
  // MaybeSwap nodes of the observation class: where the ProbParameter statistics come from
  h->msites.clear();
  for (int v = 0; v < cm.n_normal; ++v) {
    const Node& n = cm.nodes[v];
    if (n.wrap != PCLEAN_WRAP_NONE || n.kind != PCLEAN_NODE_CHOICE || n.dist != PCLEAN_DIST_MAYBE_SWAP) continue;
    for (int pid = 0; pid < h->n_patterns * h->n_blocks; ++pid) {
      const BlockProgram& bp = h->progs[pid];
      bool done = false;
      for (const MswapL& ms : bp.mswaps) if (ms.obs_vertex == v && !done) { h->msites.push_back(mswap_of(ms, false)); done = true; }
      if (done) break;
    }
  }
  h->d_mswaps.upload(h->h_mswaps); h->d_fills.upload(h->h_fills);
  { std::vector<double> lk = h->h_lkconst; if (lk.empty()) lk.push_back(0.0); h->d_lkconst.upload(lk); }
  {
    std::vector<int> ts = h->h_time_sid; if (ts.empty()) ts.push_back(-1);
    h->d_time_sid.upload(ts);
    std::vector<uint8_t> ok(std::max<size_t>(1, h->strings.size()), 0);
    if (!h->h_time_sid.empty()) for (size_t i = 0; i < h->strings.size(); ++i) ok[i] = time_regex(h->strings[i]) ? 1 : 0;
    h->d_time_ok.upload(ok);
  }
  h->d_progs.upload(h->h_progs); h->d_stars.upload(h->h_stars); h->d_terms.upload(h->h_terms);
  h->d_children.upload(h->h_children); h->d_copies.upload(h->h_copies);
  h->d_prior.upload(h->h_prior); h->d_optsid.upload(h->h_optsid);
  std::vector<double*> hp; for (auto& H : h->hoists) hp.push_back(H.val->p);
  h->d_hoist_ptrs.upload(hp);
  h->h_join_mat.assign(std::max<size_t>(1, h->joins.size()) * h->max_a, -1);
  h->d_join_mat.upload(h->h_join_mat);
  std::vector<int> aslot(std::max(1, h->n_dev_strings), -1);
  h->d_a_slot.upload(aslot);
  h->d_needed_a.alloc(std::max(1, h->n_dev_strings)); h->d_needed_a.zero();
  h->d_needed_any.alloc(1); h->d_needed_any.zero(); h->d_stats2.alloc(2 * 296);
  h->a_sids.clear();

  // ---- particles
  const int64_t N = h->N; const int K = h->K;
  h->d_pchoice.clear(); std::vector<int*> pp;
  for (int b = 0; b < h->n_blocks; ++b) { h->d_pchoice.emplace_back(new DBuf<int>()); h->d_pchoice.back()->alloc((size_t)K * N); pp.push_back(h->d_pchoice.back()->p); }
  h->d_pchoice_ptrs.upload(pp);
  h->d_pweight.alloc((size_t)K * N); h->d_plogml.alloc(N); h->d_row_logml.alloc(N); h->d_sel.alloc(N); h->d_row_flags.alloc(N);
  h->d_sel.zero(); h->d_row_logml.zero();
  h->d_row_bad.alloc(N); h->d_row_bad.zero();
  // scratch records of particles that propose a new row: a quarter of all particles may do so at once
  h->pool_cap = (int)std::min<int64_t>(std::max<int64_t>(65536, N * K / 4), 8 * 1024 * 1024);
  h->d_pool.alloc((size_t)h->pool_cap * h->nvC); h->d_pool_count.alloc(1); h->d_pool_count.zero();
  if (!h->d_err.p) { h->d_err.alloc(1); h->d_err.zero(); }     // (kept across re-finalisation: pclean_update_observations may have flagged a value)
  h->d_dbg.alloc(32); h->d_dbg.zero();
  const int64_t NB = std::max<int64_t>(N, h->max_cap) + 2;
  h->d_req.alloc(NB); h->d_flags.alloc(NB + 1); h->d_rank.alloc(NB + 1); h->d_counter.alloc(4); h->d_counter.zero();
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->d_flags.p, h->d_rank.p, (int)(NB + 1));
  h->d_cub_tmp.alloc(tmp_bytes + 256);

  // ---- device descriptor
  std::vector<int*> uptrs; for (auto& c : h->cols) uptrs.push_back(c->d_uobs.p);
  h->d_uobs_ptrs.upload(uptrs);
  Dev& D = h->h_dev;
  D.sym = h->d_sym.p; D.str_off = h->d_str_off.p; D.str_len = h->d_str_len.p; D.n_strings = h->n_dev_strings;
  D.LG = h->d_LG.p; D.LOGN = h->d_LOGN.p; D.LUT = h->d_LUT.p;
  D.N = N; D.n_cols = (int)h->cols.size(); D.nvC = h->nvC; D.uobs = h->d_uobs_ptrs.p;
  D.progs = h->d_progs.p; D.stars = h->d_stars.p; D.terms = h->d_terms.p; D.children = h->d_children.p; D.copies = h->d_copies.p;
  D.mats = nullptr; D.join_mat = h->d_join_mat.p; D.max_a = h->max_a; D.a_slot_of_sid = h->d_a_slot.p;
  D.prior_pool = h->d_prior.p; D.optsid_pool = h->d_optsid.p; D.hoist_val = h->d_hoist_ptrs.p; D.tables = h->d_tables.p;
  D.K = K; D.n_blocks = h->n_blocks; D.assign = h->d_assign_ptrs.p; D.pchoice = h->d_pchoice_ptrs.p;
  D.pweight = h->d_pweight.p; D.plogml = h->d_plogml.p; D.sel = h->d_sel.p; D.row_logml = h->d_row_logml.p; D.row_flags = h->d_row_flags.p; D.row_bad = h->d_row_bad.p;
  {
    std::vector<int*> lp; for (auto& c : h->cols) lp.push_back(c->d_ulist.p);
    h->d_ulist_ptrs.upload(lp); D.ulist = h->d_ulist_ptrs.p;
    const size_t mc = (size_t)std::max(16, h->max_cap);
    h->d_lref_off.alloc(mc + 2); h->d_lref_rows.alloc(std::max<int64_t>(1, N)); h->d_slot_of_row.alloc(std::max<int64_t>(1, N)); h->d_iota.alloc(std::max<int64_t>(1, N));
    h->d_lpat.alloc(mc + 2); h->d_lslots.alloc(mc + 2);
    h->d_lchoice.alloc((size_t)PCL_MAX_SITES * mc); h->d_lsel.alloc(mc); h->d_lflags.alloc(mc); h->d_llogml.alloc(mc);
    h->d_collist.alloc(mc + 2);
    {
      const size_t ng = std::max<size_t>(1, h->gsets.size());
      const int64_t NN = std::max<int64_t>(1, N);
      h->d_grp_key.clear(); h->d_grp_cnt.clear();
      std::vector<unsigned long long*> kp(ng, nullptr); std::vector<int*> cp(ng, nullptr);
      for (size_t g = 0; g < h->gsets.size(); ++g) {
        h->d_grp_key.emplace_back(new DBuf<unsigned long long>()); h->d_grp_key.back()->alloc(NN + 1);
        h->d_grp_cnt.emplace_back(new DBuf<int>()); h->d_grp_cnt.back()->alloc(NN + 1);
        kp[g] = h->d_grp_key.back()->p; cp[g] = h->d_grp_cnt.back()->p;
      }
      h->d_grp_key_ptrs.upload(kp); h->d_grp_cnt_ptrs.upload(cp);
      h->d_grp_n.alloc(ng); h->d_grp_n.zero();
      h->d_grp_tmp.alloc(2 * (NN + 1));
      size_t b1 = 0, b2 = 0;
      cub::DeviceRadixSort::SortKeys(nullptr, b1, h->d_grp_tmp.p, h->d_grp_tmp.p + NN + 1, (int)NN);
      cub::DeviceRunLengthEncode::Encode(nullptr, b2, h->d_grp_tmp.p, h->d_grp_tmp.p, h->d_grp_n.p, h->d_grp_n.p, (int)NN);
      h->d_grp_cub.alloc(std::max(b1, b2) + 256);
      D.lgrp_key = h->d_grp_key_ptrs.p; D.lgrp_cnt = h->d_grp_cnt_ptrs.p; D.lgrp_n = h->d_grp_n.p;
    }
    D.lref_off = h->d_lref_off.p; D.lref_rows = h->d_lref_rows.p; D.lchoice = h->d_lchoice.p; D.lsel = h->d_lsel.p; D.llogml = h->d_llogml.p; D.lflags = h->d_lflags.p;
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, h->d_slot_of_row.p, h->d_slot_of_row.p, h->d_iota.p, h->d_lref_rows.p, (int)std::max<int64_t>(1, N));
    h->d_sort_tmp.alloc(sb + 256);
  }
  {
    // observed Gaussian nodes of the observation class whose mean is a learned MeanParameter
    h->gsites.clear();
    for (int v = 0; v < cm.n_normal; ++v) {
      const Node& n = cm.nodes[v];
      if (n.wrap != PCLEAN_WRAP_NONE || n.kind != PCLEAN_NODE_CHOICE) continue;
      if (n.dist != PCLEAN_DIST_TRANSFORMED_GAUSSIAN && n.dist != PCLEAN_DIST_ADD_NOISE) continue;
      auto cit = h->col_of_vertex.find(v);
      if (cit == h->col_of_vertex.end() || !h->cols[cit->second]->is_real) continue;
      const Node& mn = cm.nodes[n.args.at(0)];
      pclean_engine::GaussSiteH G{}; G.d.obs_col = cit->second; G.d.lookup = -1; G.d.direct_slot = -1; G.d.nargs = 0;
      if (mn.kind == PCLEAN_NODE_PARAM && !m.param_indexed[mn.param]) {
        for (size_t s2 = 0; s2 < m.slot_param.size(); ++s2) if (m.slot_param[s2] == mn.param) G.d.direct_slot = (int)s2;
        if (G.d.direct_slot < 0) continue;
        G.slots.push_back(G.d.direct_slot);
      } else if (mn.kind == PCLEAN_NODE_JULIA && m.funcs[mn.func].kind == PCLEAN_FUNC_TABLE) {
        const FuncM& f = m.funcs[mn.func];
        bool params = !f.table.empty();
        for (auto& kv : f.table) params = params && kv.second.tag == PCLEAN_VAL_PARAM;
        if (!params) continue;
        G.d.lookup = lookup_index(mn.func);
        G.d.nargs = (int)f.keyargs.size();
        for (int a2 = 0; a2 < G.d.nargs; ++a2) G.d.args[a2] = trace_arg_of(mn.args.at(f.keyargs[a2]));
        for (auto& kv : f.table) G.slots.push_back(kv.second.i);
        std::sort(G.slots.begin(), G.slots.end()); G.slots.erase(std::unique(G.slots.begin(), G.slots.end()), G.slots.end());
      } else continue;
      const Node& sn = cm.nodes[n.args.at(1)];
      if (sn.kind != PCLEAN_NODE_JULIA || m.funcs[sn.func].kind != PCLEAN_FUNC_CONST) throw Unsupported("MeanParameter statistics: non-constant standard deviation");
      G.stdev = m.funcs[sn.func].cst.tag == PCLEAN_VAL_REAL ? m.funcs[sn.func].cst.d : (double)m.funcs[sn.func].cst.i;
      if (n.dist == PCLEAN_DIST_TRANSFORMED_GAUSSIAN) G.d.xform = trace_arg_of(n.args.at(2));
      else { G.d.xform = TraceArgD{0, 0, 0, 0}; if (m.xform_scale.empty() || m.xform_scale[0] != 1.0) throw Unsupported("AddNoise statistics need the identity transformation at index 0"); }
      h->gsites.push_back(G);
    }
    if (!h->gsites.empty()) {
      h->d_x_of_row.alloc(std::max<int64_t>(1, N));
      h->d_msum.alloc(std::max<size_t>(1, h->params.size())); h->d_mcnt.alloc(std::max<size_t>(1, h->params.size()));
    }
  }
  {
    // rents shapes: inner enumerations, lookups, side tables, buckets, local row cells
    h->d_inners.upload(h->h_inners); h->d_innervals.upload(h->h_innervals);
    std::vector<LookupD> lk;
    for (auto& L : h->lookups) lk.push_back(LookupD{L->keys.p, L->vals.p, L->mask, L->nkey});
    h->d_lookups.upload(lk);
    std::vector<int> loff{0}, lsid;
    for (const auto& lst : m.lists) { for (const Val& v : lst) if (v.tag == PCLEAN_VAL_STR) lsid.push_back(v.i); loff.push_back((int)lsid.size()); }
    h->d_lists_off.upload(loff); h->d_lists_sid.upload(lsid);
    h->d_splp.upload(h->h_splp); h->d_univ.upload(h->h_univ); h->d_optmap.upload(h->h_optmap);
    std::vector<double> preal(std::max<size_t>(1, h->params.size()), 0.0);
    for (size_t sl = 0; sl < h->params.size(); ++sl) {
      ParamH& PR = h->params[sl];
      if (m.param_kind[PR.spec] == PCLEAN_PARAM_MEAN || m.param_kind[PR.spec] == PCLEAN_PARAM_PROB) {
        if (PR.value.empty()) {         // initialize_parameter: keyed prior draw (add_noise.jl:43-45, maybe_swap.jl:57-59)
          pclean_stream st{}; st.key.seed = h->param_seed; st.key.row = (int64_t)sl; st.key.purpose = PCLEAN_RNG_PARAM_INIT;
          if (m.param_kind[PR.spec] == PCLEAN_PARAM_MEAN) PR.value = {m.param_prior0[PR.spec] + m.param_prior1[PR.spec] * pclean_next_normal(&st)};
          else PR.value = {pclean_next_beta(&st, m.param_prior0[PR.spec], m.param_prior1[PR.spec])};
        }
        preal[sl] = PR.value[0];
      }
    }
    h->d_param_real.upload(preal);
    std::vector<double> xf = m.xform_scale; if (xf.empty()) xf.push_back(1.0);
    h->d_xform.upload(xf);
    h->d_bkt_off.clear(); h->d_bkt_slots.clear();
    std::vector<int*> bo(nc, nullptr), bs(nc, nullptr);
    int maxslots = 16;
    for (int t = 0; t < nc; ++t) {
      h->d_bkt_off.emplace_back(new DBuf<int>()); h->d_bkt_slots.emplace_back(new DBuf<int>());
      if (t < (int)h->bucket_col_of_table.size() && h->bucket_col_of_table[t] >= 0) {
        h->d_bkt_off[t]->alloc(h->n_dev_strings + 2); h->d_bkt_off[t]->zero(); h->d_bkt_slots[t]->alloc(h->tables[t].cap);
        bo[t] = h->d_bkt_off[t]->p; bs[t] = h->d_bkt_slots[t]->p; maxslots = std::max(maxslots, h->tables[t].cap);
      }
    }
    h->d_bkt_off_ptrs.upload(bo); h->d_bkt_slots_ptrs.upload(bs);
    h->d_bkt_keys.alloc(maxslots); h->d_bkt_keys_out.alloc(maxslots); h->d_bkt_iota.alloc(maxslots);
    std::vector<int*> sp; std::vector<double*> rp;
    for (auto& c : h->cols) { sp.push_back(c->d_sid.p); rp.push_back(c->is_real ? c->d_real.p : nullptr); }
    h->d_obs_sid_ptrs.upload(sp); h->d_obs_real_ptrs.upload(rp);
    // local discrete cells of the observation rows + per-particle inner choices
    h->d_rowcell.clear(); std::vector<int*> rc(h->nvC, nullptr);
    std::vector<char> is_local(h->nvC, 0);
    for (int pid = 0; pid < h->n_patterns * h->n_blocks; ++pid) for (int q = 0; q < h->h_progs[pid].n_local; ++q) is_local[h->h_progs[pid].local_vertex[q]] = 1;
    for (int v = 0; v < h->nvC; ++v) {
      h->d_rowcell.emplace_back(new DBuf<int>());
      if (!is_local[v]) continue;
      std::vector<int> init(N, PCL_UNSET);
      auto it = h->rowcell_init.find(v);
      if (it != h->rowcell_init.end()) init = it->second;
      h->d_rowcell[v]->upload(init); rc[v] = h->d_rowcell[v]->p;
    }
    h->d_rowcell_ptrs.upload(rc);
    h->d_pinner.clear(); std::vector<int*> pi(h->n_blocks, nullptr);
    for (int b = 0; b < h->n_blocks; ++b) {
      h->d_pinner.emplace_back(new DBuf<int>());
      int nl = 0;
      for (int pt = 0; pt < h->n_patterns; ++pt) nl = std::max(nl, h->h_progs[pt * h->n_blocks + b].n_local);
      if (nl > 0) { h->d_pinner[b]->alloc((size_t)PCL_MAX_LOCAL * K * N); pi[b] = h->d_pinner[b]->p; }
    }
    h->d_pinner_ptrs.upload(pi);
    h->d_pat_rows.clear();
    for (auto& L : h->pat_rows) { h->d_pat_rows.emplace_back(new DBuf<long long>()); h->d_pat_rows.back()->upload(L); }
    h->d_pat_of_row.upload(h->pat_of_row);
    D.inners = h->d_inners.p; D.lookups = h->d_lookups.p; D.innervals = h->d_innervals.p; D.param_real = h->d_param_real.p; D.xform_scale = h->d_xform.p; D.gext = h->d_gext.p;
    D.mswaps = h->d_mswaps.p; D.fills = h->d_fills.p; D.lkconst = h->d_lkconst.p; D.time_sid = h->d_time_sid.p; D.time_ok = h->d_time_ok.p;
    {
      std::vector<int> vc(h->nvC, -1);
      for (size_t ci = 0; ci < h->cols.size(); ++ci)      // only cells of referenced rows (SubmodelNodes) the dataset observes directly
        if (!h->cols[ci]->is_real && cm.nodes[h->cols[ci]->vertex].wrap == PCLEAN_WRAP_SUBMODEL) vc[h->cols[ci]->vertex] = (int)ci;
      h->d_vcol.upload(vc);
    }
    D.vcol = h->d_vcol.p;
    D.obs_real = h->d_obs_real_ptrs.p; D.obs_sid = h->d_obs_sid_ptrs.p; D.lists_off = h->d_lists_off.p; D.lists_sid = h->d_lists_sid.p;
    D.splp_pool = h->d_splp.p; D.univ_col = h->d_univ.p; D.optmap_pool = h->d_optmap.p;
    D.bkt_off = h->d_bkt_off_ptrs.p; D.bkt_slots = h->d_bkt_slots_ptrs.p; D.rowcell = h->d_rowcell_ptrs.p; D.pinner = h->d_pinner_ptrs.p;
  }
  D.prune = h->prune; D.row_order = nullptr;
  if (h->memo_log2 > 0) {
    for (int tb = 0; tb < 2; ++tb) {
      h->d_memo_keys[tb].alloc((size_t)1 << h->memo_log2); h->d_memo_vals[tb].alloc((size_t)1 << h->memo_log2);
      D.memo_keys[tb] = h->d_memo_keys[tb].p; D.memo_vals[tb] = h->d_memo_vals[tb].p;
    }
    D.memo_mask = (1u << h->memo_log2) - 1u;
  } else { for (int tb = 0; tb < 2; ++tb) { D.memo_keys[tb] = nullptr; D.memo_vals[tb] = nullptr; } D.memo_mask = 0; }
  h->pmemo_dirty = true;
  D.opts = h->opts;
  {
    // pruning order of every star's terms: identity until recount() ranks them (k_term_order)
    std::vector<int> ord(std::max<size_t>(1, h->h_terms.size()), 0);
    for (size_t pi = 0; pi < h->h_progs.size(); ++pi) {
      const ProgD& P = h->h_progs[pi];
      for (int si = 0; si < P.nstar; ++si) { const StarD& S = h->h_stars[P.star0 + si]; for (int i = 0; i < S.nterm; ++i) ord[P.term0 + S.term0 + i] = i; }
    }
    h->d_term_order.upload(ord); D.term_order = h->d_term_order.p;
    std::vector<float> ml(std::max<size_t>(1, h->cols.size()), 0.0f);
    for (size_t ci = 0; ci < h->cols.size(); ++ci) {
      double tot = 0; for (int sid : h->cols[ci]->ulist) tot += (double)len[sid];
      ml[ci] = h->cols[ci]->ulist.empty() ? 0.0f : (float)(tot / (double)h->cols[ci]->ulist.size());
    }
    h->d_col_meanlen.upload(ml); D.col_meanlen = h->d_col_meanlen.p;
  }
  D.pool = h->d_pool.p; D.pool_cap = h->pool_cap; D.pool_count = h->d_pool_count.p; D.needed_a = h->d_needed_a.p; D.needed_any = h->d_needed_any.p; D.err = h->d_err.p; D.dbg = h->d_dbg.p;
  D.lm_uni = h->d_lm_uni.p; D.lm_big = h->d_lm_big.p; D.lm_sym = h->d_lm_sym.p;
  D.newstr_chars = h->d_newstr_chars.p; D.newstr_len = h->d_newstr_len.p; D.newstr_count = h->d_newstr_count.p;
  D.newstr_cap = h->newstr_cap; D.newstr_base = (int)h->strings.size();
  build_list_mats(h);
  D.lmats = h->d_lmats.p;
  h->d_dev.alloc(1);
  upload_dev(h);
  upload_tables(h);
  upload_mats(h);

  // ---- distance matrices (the device form of the reference's AddTypos memo)
  for (const PendingMat& pm : pending_opt) run_dp(h, *h->mats[pm.mat], h->d_optsid.p + pm.opt_off, 0, pm.nopt);
  h->mats_dirty = true;
  refresh_candidate_mats(h);
  compute_hoists(h, false);
  CK(cudaStreamSynchronize(h->stream));
  h->finalized = true;
}

// hash index of @guaranteed keys (TableTrace.hashed_keys, trace.jl:33; dependency_tracking.jl:77-84):
// CSR key string id -> live slots, rebuilt from the table cells
void build_buckets(Eng* h) {
  for (int t = 0; t < (int)h->tables.size(); ++t) {
    if (t >= (int)h->bucket_col_of_table.size() || h->bucket_col_of_table[t] < 0) continue;
    TableH& T = h->tables[t];
    const int ns = h->n_dev_strings;
    k_zero_int<<<nblk(ns + 2, 256), 256, 0, h->stream>>>(h->d_bkt_off[t]->p, ns + 2); ++h->launches;
    if (T.n_slots > 0) {
      k_bucket_keys<<<nblk(T.n_slots, 256), 256, 0, h->stream>>>(h->d_tables.p, t, h->bucket_col_of_table[t], ns, h->d_bkt_keys.p, h->d_bkt_off[t]->p, h->d_bkt_iota.p); ++h->launches;
    }
    size_t tmp = h->d_cub_tmp.n;
    DBuf<int> scanned; scanned.alloc(ns + 2);
    CK(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, h->d_bkt_off[t]->p, scanned.p, ns + 2, h->stream)); ++h->launches;
    CK(cudaMemcpyAsync(h->d_bkt_off[t]->p, scanned.p, (size_t)(ns + 2) * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    if (T.n_slots > 0) {
      size_t sb = h->d_sort_tmp.n;
      CK(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, sb, h->d_bkt_keys.p, h->d_bkt_keys_out.p, h->d_bkt_iota.p, h->d_bkt_slots[t]->p, T.n_slots, 0, 32, h->stream)); ++h->launches;
    }
    CK(cudaStreamSynchronize(h->stream));     // `scanned` is freed at scope exit
  }
}

// ------------------------------------------------------------------------------------------
// k_block geometry variants (option "kb_variant"): resident warps per SM against registers per thread
// ------------------------------------------------------------------------------------------
struct KbVariant { int warps, minb; };
static const KbVariant kKbVariants[3] = {{16, 2}, {12, 2}, {16, 1}};
template <bool RICH, int WARPS, int MINB> void kb_prepare(Eng* h, int device, int* grid_out) {
  cudaFuncSetAttribute(k_block<RICH, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCL_KBLOCK_SMEM_W(WARPS));
  cudaDeviceProp prop{}; int per_sm = 0;
  if (grid_out && cudaGetDeviceProperties(&prop, device) == cudaSuccess &&
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_block<RICH, WARPS, MINB>, 32 * WARPS, PCL_KBLOCK_SMEM_W(WARPS)) == cudaSuccess && per_sm > 0)
    *grid_out = prop.multiProcessorCount * per_sm;        // persistent: every resident CTA slot of every SM
}
template <bool RICH, int WARPS, int MINB> void kb_launch(Eng* h, int prog, int block, long long row0, long long cnt, uint64_t seed, uint32_t sweep, uint32_t cls, int csmc, const long long* list) {
  const int grid = std::min(nblk(cnt, WARPS), h->block_grid);
  k_block<RICH, WARPS, MINB><<<grid, 32 * WARPS, PCL_KBLOCK_SMEM_W(WARPS), h->stream>>>(h->h_dev, prog, block, row0, cnt, seed, sweep, cls, csmc, list);
}
void launch_k_block(Eng* h, bool rich, int prog, int block, long long row0, long long cnt, uint64_t seed, uint32_t sweep, uint32_t cls, int csmc, const long long* list) {
  switch (h->kb_variant * 2 + (rich ? 1 : 0)) {
    case 0: kb_launch<false, 16, 2>(h, prog, block, row0, cnt, seed, sweep, cls, csmc, list); break;
    case 1: kb_launch<true, 16, 2>(h, prog, block, row0, cnt, seed, sweep, cls, csmc, list); break;
    case 2: kb_launch<false, 12, 2>(h, prog, block, row0, cnt, seed, sweep, cls, csmc, list); break;
    case 3: kb_launch<true, 12, 2>(h, prog, block, row0, cnt, seed, sweep, cls, csmc, list); break;
    case 4: kb_launch<false, 16, 1>(h, prog, block, row0, cnt, seed, sweep, cls, csmc, list); break;
    default: kb_launch<true, 16, 1>(h, prog, block, row0, cnt, seed, sweep, cls, csmc, list); break;
  }
}
void prepare_k_block(Eng* h) {
  int g = 0;
  switch (h->kb_variant) {
    case 0: kb_prepare<true, 16, 2>(h, h->device, nullptr); kb_prepare<false, 16, 2>(h, h->device, &g); break;
    case 1: kb_prepare<true, 12, 2>(h, h->device, nullptr); kb_prepare<false, 12, 2>(h, h->device, &g); break;
    default: kb_prepare<true, 16, 1>(h, h->device, nullptr); kb_prepare<false, 16, 1>(h, h->device, &g); break;
  }
  if (g > 0) h->block_grid = g;
}

// ------------------------------------------------------------------------------------------
// the row move kernels for rows [r0, r1)
// ------------------------------------------------------------------------------------------
void run_row_moves(Eng* h, int64_t r0, int64_t r1, uint64_t seed, uint32_t sweep, bool csmc, const std::vector<long long>* rows = nullptr) {
  const int64_t n = rows ? (int64_t)rows->size() : r1 - r0;
  if (n <= 0) return;
  const uint32_t cls = (uint32_t)h->obs_cls;
  const int K = h->K; const int64_t N = h->N;
  const long long* drows = nullptr;
  std::vector<long long> by_pat; std::vector<int64_t> pat_off;
  if (rows) {
    // an explicit list of rows (initialisation order): grouped per missingness pattern for the block kernels
    if (h->d_rowlist.n < (size_t)n) { h->d_rowlist.alloc(n + 1024); h->d_rowlist_pat.alloc(n + 1024); h->d_rowlist_int.alloc(n + 1024); }
    CK(cudaMemcpyAsync(h->d_rowlist.p, rows->data(), n * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    drows = h->d_rowlist.p;
    pat_off.assign(h->n_patterns + 1, 0);
    for (int pt = 0; pt < h->n_patterns; ++pt) {
      for (long long r : *rows) if (h->n_patterns == 1 || h->pat_of_row[r] == pt) by_pat.push_back(r);
      pat_off[pt + 1] = (int64_t)by_pat.size();
    }
    CK(cudaMemcpyAsync(h->d_rowlist_pat.p, by_pat.data(), by_pat.size() * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    k_reset_rows<<<nblk(n, 256), 256, 0, h->stream>>>(h->d_dev.p, drows, n); ++h->launches;
  } else {
    // zero per-row particle state of the range (weights are [K][N]: strided memsets)
    CK(cudaMemsetAsync(h->d_pweight.p + (size_t)r0 * K, 0, (size_t)n * K * sizeof(double), h->stream));      // weights are [N][K]
    CK(cudaMemsetAsync(h->d_plogml.p + r0, 0, n * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->d_row_flags.p + r0, 0, n * sizeof(int), h->stream));
    CK(cudaMemsetAsync(h->d_row_bad.p + r0, 0, n * sizeof(unsigned long long), h->stream));
  }
  CK(cudaMemsetAsync(h->d_pool_count.p, 0, sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->d_newstr_count.p, 0, sizeof(int), h->stream));
  // strings generated by this call get the ids after everything interned so far; a sharded engine cannot
  // hand them to the other replicas (ids are per process): there the dummy particle stays unusable
  h->h_dev.newstr_base = (int)h->strings.size();
  {
    const long long room = std::min<long long>((long long)h->str_cap - (long long)h->strings.size(), (long long)((h->sym_cap - h->sym_used) / PCL_NEWSTR_MAX));
    h->h_dev.newstr_cap = h->nccl.comm ? 0 : (int)std::max<long long>(0, std::min<long long>(h->newstr_cap, room));
  }
  if (h->h_dev.memo_mask) {
    // table 0 (reference-table stars) lives for this call; table 1 (choice stars) until a prior changes
    for (int tb = 0; tb < 2; ++tb) {
      if (tb == 1 && !h->pmemo_dirty) continue;
      k_memo_reset<<<nblk((int64_t)h->d_memo_keys[tb].n, 256), 256, 0, h->stream>>>(h->d_memo_keys[tb].p, h->d_memo_vals[tb].p, (long long)h->d_memo_keys[tb].n);
      ++h->launches;
    }
    h->pmemo_dirty = false;
  }
  build_buckets(h);
  for (int b = 0; b < h->n_blocks; ++b) {
    if (h->h_progs[b].n_earlier) {
      if (h->n_patterns > 1) throw Unsupported("earlier-block joins together with several missingness patterns");
      // which upstream string values does this block need join matrices for?
      k_collect_a<<<nblk(n * K, 256), 256, 0, h->stream>>>(h->d_dev.p, b, r0, n, drows); ++h->launches;
      int any_needed = 0;
      CK(cudaMemcpyAsync(&any_needed, h->d_needed_any.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (any_needed) {                 // rare after the first sweeps: a state value nobody held before
        std::vector<int> need = h->d_needed_a.download();
        for (int s = 0; s < (int)need.size(); ++s) if (need[s]) build_join_mats_for(h, s);
        h->d_needed_a.zero(); h->d_needed_any.zero();
      }
    }
    if (b < 8) CK(cudaEventRecord(h->evb[2 * b], h->stream));
    for (int pt = 0; pt < h->n_patterns; ++pt) {
      long long row0 = r0, cnt = n; const long long* list = nullptr;
      if (rows) { row0 = pat_off[pt]; cnt = pat_off[pt + 1] - pat_off[pt]; list = h->d_rowlist_pat.p; }
      else if (h->n_patterns > 1) {
        const std::vector<long long>& L = h->pat_rows[pt];
        const long long lo = std::lower_bound(L.begin(), L.end(), (long long)r0) - L.begin();
        const long long hi = std::lower_bound(L.begin(), L.end(), (long long)r1) - L.begin();
        row0 = lo; cnt = hi - lo; list = h->d_pat_rows[pt]->p;
      }
      if (cnt <= 0) continue;
      if (h->prog_rootless.at(pt * h->n_blocks + b)) {
        k_rootless<<<nblk(cnt * K, 256), 256, 0, h->stream>>>(h->d_dev.p, pt * h->n_blocks + b, b, row0, cnt, list, seed, sweep, cls, csmc ? 1 : 0);
        ++h->launches;
        continue;
      }
      launch_k_block(h, h->prog_rich.at(pt * h->n_blocks + b) != 0, pt * h->n_blocks + b, b, row0, cnt, seed, sweep, cls, csmc ? 1 : 0, list);
      ++h->launches;
    }
    if (b < 8) CK(cudaEventRecord(h->evb[2 * b + 1], h->stream));
    if (!h->cfg.use_mh_instead_of_pg && b < h->n_blocks - 1) {
      k_resample<<<nblk(n, 128), 128, 0, h->stream>>>(h->d_dev.p, b, r0, n, seed, sweep, cls, csmc ? 1 : 0, drows); ++h->launches;
    }
  }
  k_select<<<nblk(n, 128), 128, 0, h->stream>>>(h->d_dev.p, r0, n, seed, sweep, cls, csmc ? 1 : 0, h->cfg.use_mh_instead_of_pg, drows); ++h->launches;
  if (rows) CK(cudaStreamSynchronize(h->stream));      // by_pat (host vector) must outlive its copy
  CK(cudaGetLastError());
}

// Strings random(StringPrior) generated during the row moves just applied (device.cuh
// dummy_string_draw): interned on the host, appended to the device dictionary in its head-room, and —
// where interning found an equal string already there, so that provisional id (base + pool index) and
// real id differ — fixed up in the table cells.
__global__ void k_remap_cells(int* cells, long long n, int base, int cnt, const int* map) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = cells[i];
  if (v >= base && v < base + cnt) cells[i] = map[v - base];
}
void intern_new_strings(Eng* h) {
  if (h->h_dev.newstr_cap <= 0) return;
  int cnt = 0;
  CK(cudaMemcpyAsync(&cnt, h->d_newstr_count.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cnt = std::min(cnt, h->h_dev.newstr_cap);
  if (cnt <= 0) return;
  std::vector<int> lens((size_t)cnt); std::vector<uint8_t> chars((size_t)cnt * PCL_NEWSTR_MAX);
  CK(cudaMemcpy(lens.data(), h->d_newstr_len.p, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(chars.data(), h->d_newstr_chars.p, chars.size(), cudaMemcpyDeviceToHost));
  static const char32_t lm_letters[] = U"abcdefghijklmnopqrstuvwxyz .";
  const int base = h->h_dev.newstr_base;
  std::vector<int> map((size_t)cnt); bool remap = false;
  std::vector<uint8_t> sym_add; std::vector<int> off_add, len_add;
  const size_t first_new = h->strings.size();
  for (int i = 0; i < cnt; ++i) {
    std::u32string s2;
    for (int q = 0; q < lens[i]; ++q) s2.push_back(lm_letters[chars[(size_t)i * PCL_NEWSTR_MAX + q] % 28]);
    const size_t before = h->strings.size();
    const int id = h->intern(s2);
    map[i] = id;
    if (id != base + i) remap = true;
    if (h->strings.size() > before) {
      off_add.push_back((int)(h->sym_used + sym_add.size())); len_add.push_back(lens[i]);
      for (int q = 0; q < lens[i]; ++q) sym_add.push_back((uint8_t)h->lm_sym[chars[(size_t)i * PCL_NEWSTR_MAX + q] % 28]);
    }
  }
  if (!off_add.empty()) {
    if (first_new + off_add.size() > (size_t)h->str_cap || h->sym_used + sym_add.size() + 1 > h->sym_cap) throw std::runtime_error("device dictionary head-room exhausted (PCLEAN_ERR_CAPACITY)");
    if (!sym_add.empty()) CK(cudaMemcpy(h->d_sym.p + h->sym_used, sym_add.data(), sym_add.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_str_off.p + first_new, off_add.data(), off_add.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_str_len.p + first_new, len_add.data(), len_add.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->sym_used += sym_add.size();
  }
  if (remap) {
    CK(cudaMemcpy(h->d_newstr_map.p, map.data(), (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice));
    for (TableH& T : h->tables) {
      if (!T.loaded || T.n_slots == 0) continue;
      const long long n = (long long)T.n_normal * T.cap;
      k_remap_cells<<<nblk(n, 256), 256, 0, h->stream>>>(T.cells.p, n, base, cnt, h->d_newstr_map.p); ++h->launches;
    }
    h->mats_dirty = true;
  }
  CK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------
// Slot compaction.  A row whose last reference went away stays in its slot as a dead column (the
// reference deletes it from the table's Dict, dependency_tracking.jl:162-202) and new rows are
// appended, so a long run walks towards the table's capacity.  When a table is close to it, the live
// rows are packed to the front IN ORDER (slot order = key order = the order the categorical draws
// enumerate, include/pclean_rng.h), every reference to them is renumbered and the columns of the
// table's distance matrices move along with their rows — no distance is recomputed.  All replicas of
// a row-sharded engine take the same decision from the same (all-reduced) counts.
// ------------------------------------------------------------------------------------------
__global__ void k_fill_int(int* p, long long n, int v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_gather_slots(const int* in, int* out, const int* map, int n_old) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_old && map[j] >= 0) out[map[j]] = in[j];
}
__global__ void k_renumber_refs(int* a, long long n, const int* map, int n_old) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = a[i];
  if (v >= 0 && v < n_old) a[i] = map[v];
}
// one CTA per matrix row: the row is staged in shared memory, then every kept column is written to
// its new place (new index <= old index, the staging makes the order of the writes irrelevant)
__global__ void __launch_bounds__(256) k_pack_matrix_cols(uint8_t* d, long long stride, const int* __restrict__ map, int n_old) {
  extern __shared__ __align__(16) uint8_t s_row[];
  uint8_t* row = d + (long long)blockIdx.x * stride;
  const int n16 = (n_old + 15) / 16;                // stride is a multiple of 16 >= n_old
  for (int j = threadIdx.x; j < n16; j += blockDim.x) reinterpret_cast<uint4*>(s_row)[j] = reinterpret_cast<const uint4*>(row)[j];
  __syncthreads();
  for (int j = threadIdx.x; j < n_old; j += blockDim.x) { const int m = map[j]; if (m >= 0 && m != j) row[m] = s_row[j]; }
}
bool compact_tables(Eng* h, bool force) {
  bool near = force;
  for (int c = 0; c < (int)h->tables.size() && !near; ++c) {
    const TableH& T = h->tables[c];
    if (c == h->obs_cls || !T.loaded || T.n_slots == 0) continue;
    const int head = std::max(std::max(h->compact_head, T.cap / 8), 2 * h->max_batch_new);
    near = T.n_slots + head > T.cap;
  }
  long long total_slots = 0;
  for (const TableH& T : h->tables) total_slots += T.loaded ? T.n_slots : 0;
  if (!near || (!force && total_slots == h->compact_futile_at)) return false;
  recount(h);
  CK(cudaStreamSynchronize(h->stream));
  struct Plan { int c, n_old, n_new; std::vector<int> map; DBuf<int> d_map; };
  std::vector<std::unique_ptr<Plan>> plans;
  for (int c = 0; c < (int)h->tables.size(); ++c) {
    TableH& T = h->tables[c];
    if (c == h->obs_cls || !T.loaded || T.n_slots == 0) continue;
    const std::vector<int> rc = T.refcnt.download(T.n_slots);
    std::unique_ptr<Plan> P(new Plan());
    P->c = c; P->n_old = T.n_slots; P->n_new = 0; P->map.assign(T.n_slots, -1);
    for (int j = 0; j < T.n_slots; ++j) if (rc[j] > 0) P->map[j] = P->n_new++;
    if (P->n_new == P->n_old) continue;
    P->d_map.upload(P->map);
    plans.push_back(std::move(P));
  }
  if (plans.empty()) { h->compact_futile_at = total_slots; return false; }
  // (1) move the rows: cells, keys, matrix columns (+ the shadow ids and element lengths that describe them)
  for (auto& P : plans) {
    TableH& T = h->tables[P->c];
    DBuf<int> tmp; tmp.alloc(T.cap);
    auto pack_ints = [&](int* col, int fill) {
      k_fill_int<<<nblk(T.cap, 256), 256, 0, h->stream>>>(tmp.p, T.cap, fill);
      k_gather_slots<<<nblk(P->n_old, 256), 256, 0, h->stream>>>(col, tmp.p, P->d_map.p, P->n_old);
      CK(cudaMemcpyAsync(col, tmp.p, (size_t)T.cap * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
      h->launches += 2;
    };
    for (int v = 0; v < T.n_normal; ++v) pack_ints(T.cells.p + (size_t)v * T.cap, PCL_UNSET);
    std::vector<int64_t> nk(P->n_new);
    for (int j = 0; j < P->n_old; ++j) if (P->map[j] >= 0) nk[P->map[j]] = T.keys[j];
    T.keys.swap(nk);
    T.slot_of_key.clear();
    std::vector<long long> kk(T.cap, 0);
    for (int j = 0; j < P->n_new; ++j) { T.slot_of_key[T.keys[j]] = j; kk[j] = T.keys[j]; }
    CK(cudaMemcpyAsync(T.d_keys.p, kk.data(), kk.size() * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    const size_t smem = (size_t)((P->n_old + 15) / 16) * 16;
    for (auto& Mp : h->mats) {
      MatH& M = *Mp;
      if (M.table != P->c || M.shadow.n == 0 || M.rows == 0) continue;
      if (smem > 200 * 1024) continue;                 // too wide to stage: the shadow comparison recomputes what moved
      if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_pack_matrix_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_pack_matrix_cols<<<M.rows, 256, smem, h->stream>>>(M.d.p, M.stride, P->d_map.p, P->n_old);
      k_pack_matrix_cols<<<1, 256, smem, h->stream>>>(M.elen.p, M.stride, P->d_map.p, P->n_old);
      h->launches += 2;
      pack_ints(M.shadow.p, -1);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));              // tmp / kk go out of scope
  }
  // (2) renumber what points at the rows: reference slots of other latent tables and of the observation rows
  for (auto& P : plans) {
    for (int c2 = 0; c2 < (int)h->tables.size(); ++c2) {
      TableH& T2 = h->tables[c2];
      if (c2 == h->obs_cls || !T2.loaded) continue;
      for (size_t g = 0; g < T2.fk_col.size(); ++g)
        if (T2.fk_table[g] == P->c) { k_renumber_refs<<<nblk(T2.cap, 256), 256, 0, h->stream>>>(T2.cells.p + (size_t)T2.fk_col[g] * T2.cap, T2.cap, P->d_map.p, P->n_old); ++h->launches; }
    }
    for (int b = 0; b < h->n_blocks; ++b) {
      if (h->progs[b].root < 0 || h->progs[b].stars[h->progs[b].root].table != P->c) continue;
      k_renumber_refs<<<nblk(h->N, 256), 256, 0, h->stream>>>(h->d_assign[b]->p, h->N, P->d_map.p, P->n_old); ++h->launches;
    }
  }
  for (auto& P : plans) h->tables[P->c].n_slots = P->n_new;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  upload_tables(h);
  // rows also carry denormalised copies of their targets' cells — reference slots of deeper tables among
  // them (Hospital keeps loc.county): rewritten from the renumbered targets, upward in class order
  for (int c2 = 0; c2 < (int)h->tables.size(); ++c2) {
    TableH& T2 = h->tables[c2];
    if (!T2.loaded || c2 == h->obs_cls || T2.n_slots == 0) continue;
    for (size_t g = 0; g < T2.fk_col.size(); ++g) {
      auto key = std::make_pair(c2, (int)g);
      k_refresh_copies<<<nblk(T2.n_slots, 256), 256, 0, h->stream>>>(h->d_tables.p, c2, (int)g, h->fk_copies.at(key)->p, h->fk_ncopies.at(key));
      ++h->launches;
    }
  }
  h->mats_dirty = true; h->pmemo_dirty = true;
  recount(h);
  refresh_candidate_mats(h);
  build_buckets(h);
  ++h->compactions;
  return true;
}

// apply the selected particles of rows [r0, r1): assignments + creation of proposed rows
void apply_moves(Eng* h, int64_t r0, int64_t r1, bool csmc, int64_t* n_changed, int64_t* n_new, const std::vector<long long>* rows = nullptr) {
  const int64_t n = rows ? (int64_t)rows->size() : r1 - r0;
  *n_changed = 0; *n_new = 0;
  if (n <= 0) return;
  h->d_counter.zero();
  const long long* drows = rows ? h->d_rowlist.p : nullptr;          // uploaded by run_row_moves for the same list
  if (rows) {
    if (h->nccl.comm || h->exchange_path) throw Unsupported("row lists together with the new-row exchange path");
    k_rows_to_int<<<nblk(n, 256), 256, 0, h->stream>>>(drows, n, h->d_rowlist_int.p); ++h->launches;
  }
  for (int b = 0; b < h->n_blocks; ++b) {
    CK(cudaMemsetAsync(h->d_counter.p + 1, 0, sizeof(int), h->stream));
    k_apply<<<nblk(n, 256), 256, 0, h->stream>>>(h->d_dev.p, b, r0, n, csmc ? 1 : 0, h->d_req.p, h->d_counter.p, h->n_patterns > 1 ? h->d_pat_of_row.p : nullptr, drows); ++h->launches;
    const BlockProgram& bp = h->progs[b];
    if (bp.root < 0) continue;                                  // no reference slot: k_apply wrote the local cells
    {
      // nobody (on any rank) proposed a new row — the common case in later sweeps: nothing to create,
      // one 4-byte all-reduce and one host read instead of the scan + gather of the exchange path
      int any_req = 0;
      if (h->nccl.comm && h->nccl.AllReduce(h->d_counter.p + 1, h->d_counter.p + 1, 1, /*ncclInt32*/ 2, /*ncclMax*/ 2, h->nccl.comm, h->stream) != 0)
        throw std::runtime_error("ncclAllReduce failed");
      CK(cudaMemcpyAsync(&any_req, h->d_counter.p + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (!any_req) continue;
    }
    // rows to create come either from this rank's rows directly, or (multi-GPU / exchange path)
    // from the records of ALL ranks, gathered and replayed in (rank, row) order on every replica
    const int* req = h->d_req.p; const int* row_ids = rows ? h->d_rowlist_int.p : nullptr; int64_t nlist = n; int64_t list_row0 = r0;
    if (h->nccl.comm || h->exchange_path) {
      const int world = h->nccl.comm ? h->nccl.world : 1, rank = h->nccl.comm ? h->nccl.rank : 0;
      const int recw = h->nvC + 1;
      k_req_flags<<<nblk(n + 1, 256), 256, 0, h->stream>>>(n, h->d_req.p, h->d_flags.p); ++h->launches;
      size_t tmp = h->d_cub_tmp.n;
      CK(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, h->d_flags.p, h->d_rank.p, (int)(n + 1), h->stream)); ++h->launches;
      int nloc = 0;
      CK(cudaMemcpyAsync(&nloc, h->d_rank.p + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      std::vector<int> counts(world, 0);
      counts[rank] = nloc;
      if (h->nccl.comm) {
        if (h->d_counts.n < (size_t)world) h->d_counts.alloc(world);
        CK(cudaMemcpyAsync(h->d_counts.p + rank, &nloc, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        if (h->nccl.AllGather(h->d_counts.p + rank, h->d_counts.p, 1, /*ncclInt32*/ 2, h->nccl.comm, h->stream) != 0) throw std::runtime_error("ncclAllGather failed");
        CK(cudaMemcpyAsync(counts.data(), h->d_counts.p, world * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
      }
      int maxc = 0, total = 0;
      for (int c : counts) { maxc = std::max(maxc, c); total += c; }
      if (total == 0) continue;
      int used = 0;
      CK(cudaMemcpy(&used, h->d_pool_count.p, sizeof(int), cudaMemcpyDeviceToHost));
      if ((int64_t)used + total > h->pool_cap) throw std::runtime_error("new-row exchange exceeds the scratch pool (PCLEAN_ERR_CAPACITY)");
      const int pool_base = h->pool_cap - total;
      if (h->d_rec_local.n < (size_t)maxc * recw) h->d_rec_local.alloc((size_t)maxc * recw);
      if (h->d_rec_all.n < (size_t)maxc * recw * world) h->d_rec_all.alloc((size_t)maxc * recw * world);
      k_pack_requests<<<nblk(n, 256), 256, 0, h->stream>>>(h->d_dev.p, r0, n, h->d_req.p, h->d_rank.p, h->d_rec_local.p); ++h->launches;
      if (h->nccl.comm) {
        if (h->nccl.AllGather(h->d_rec_local.p, h->d_rec_all.p, (size_t)maxc * recw, 2, h->nccl.comm, h->stream) != 0) throw std::runtime_error("ncclAllGather failed");
      } else CK(cudaMemcpyAsync(h->d_rec_all.p, h->d_rec_local.p, (size_t)maxc * recw * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
      std::vector<int> src;
      for (int rk = 0; rk < world; ++rk) for (int i = 0; i < counts[rk]; ++i) src.push_back(rk * maxc + i);
      if (h->d_src.n < src.size()) { h->d_src.alloc(src.size() + 1024); h->d_row_ids.alloc(src.size() + 1024); }
      CK(cudaMemcpyAsync(h->d_src.p, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      if (h->d_req.n < (size_t)total) throw std::runtime_error("new-row exchange exceeds the request buffer");
      k_unpack_requests<<<nblk(total, 256), 256, 0, h->stream>>>(h->d_dev.p, total, pool_base, h->d_rec_all.p, h->d_src.p, h->d_req.p, h->d_row_ids.p); ++h->launches;
      CK(cudaStreamSynchronize(h->stream));      // src (host vector) must outlive the copy
      req = h->d_req.p; row_ids = h->d_row_ids.p; nlist = total; list_row0 = 0;
    }
    for (int sidx : bp.order) {                          // post-order: nested rows first
      const StarL& s = bp.stars[sidx];
      if (s.kind != ST_FK) continue;
      k_create_flags<<<nblk(nlist + 1, 256), 256, 0, h->stream>>>(h->d_dev.p, b, sidx, nlist, req, h->d_flags.p); ++h->launches;
      size_t tmp = h->d_cub_tmp.n;
      CK(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, h->d_flags.p, h->d_rank.p, (int)(nlist + 1), h->stream)); ++h->launches;
      int total = 0;
      CK(cudaMemcpyAsync(&total, h->d_rank.p + nlist, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (total == 0) continue;
      TableH& T = h->tables[s.table];
      if (csmc) h->max_batch_new = std::max(h->max_batch_new, total);     // sweeps only: initialisation batches append far more
      if (T.n_slots + total > T.cap) throw std::runtime_error("latent table capacity exceeded: reserve more rows with pclean_reserve_table / option table_cap (PCLEAN_ERR_CAPACITY)");
      k_create_rows<<<nblk(nlist, 256), 256, 0, h->stream>>>(h->d_dev.p, b, sidx, b, list_row0, nlist, req, h->d_flags.p, h->d_rank.p, T.n_slots, sidx == bp.root ? 1 : 0, row_ids);
      ++h->launches;
      {
        std::vector<long long> nk(total);
        for (int i = 0; i < total; ++i) { T.slot_of_key[h->next_key] = T.n_slots + i; nk[i] = h->next_key; T.keys.push_back(h->next_key++); }
        CK(cudaMemcpy(T.d_keys.p + T.n_slots, nk.data(), total * sizeof(long long), cudaMemcpyHostToDevice));
      }
      T.n_slots += total; *n_new += total;
      h->mats_dirty = true;
      upload_tables(h);
    }
  }
  std::vector<int> cnt = h->d_counter.download(1);
  *n_changed = cnt[0];
  CK(cudaGetLastError());
}


// ------------------------------------------------------------------------------------------
// latent-class moves
// ------------------------------------------------------------------------------------------
int latent_prog(Eng* h, int cls) {
  auto it = h->lprog_of_class.find(cls);
  if (it != h->lprog_of_class.end()) return it->second;
  auto er = h->lprog_error.find(cls);
  throw Unsupported(er != h->lprog_error.end() ? er->second : std::string("class has no latent program (not loaded?)"));
}

// CSR of the observation rows that (transitively) refer to each slot of `cls` (collect_referring_rows,
// row_inference.jl:23-47): stable radix sort of (slot, row) — rows stay ascending within a slot.
void build_ref_csr(Eng* h, int cls) {
  TableH& T = h->tables[cls];
  const RefChainD ch = h->ref_chain.at(cls);
  const int64_t N = h->N;
  k_zero_int<<<nblk(T.cap + 1, 256), 256, 0, h->stream>>>(h->d_flags.p, T.cap + 1); ++h->launches;
  k_ref_slots<<<nblk(N, 256), 256, 0, h->stream>>>(h->d_dev.p, ch, N, h->d_slot_of_row.p, h->d_flags.p); ++h->launches;
  size_t tmp = h->d_cub_tmp.n;
  CK(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, h->d_flags.p, h->d_lref_off.p, T.cap + 1, h->stream)); ++h->launches;
  k_iota<<<nblk(N, 256), 256, 0, h->stream>>>(h->d_iota.p, N); ++h->launches;
  size_t sb = h->d_sort_tmp.n;
  CK(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, sb, h->d_slot_of_row.p, h->d_req.p, h->d_iota.p, h->d_lref_rows.p, (int)N, 0, 32, h->stream)); ++h->launches;
  // the same referrers grouped by what they observe: distinct (slot, [other half,] observed string) with multiplicities
  if (cls < (int)h->gsets_of_class.size() && !h->gsets_of_class[cls].empty()) {
    if (T.cap >= (1 << 20)) throw Unsupported("latent table with 2^20 or more slots (referrer group keys hold 20 slot bits)");
    for (int g : h->gsets_of_class[cls]) {
      k_group_keys<<<nblk(N, 256), 256, 0, h->stream>>>(h->d_dev.p, h->gsets[g], h->d_slot_of_row.p, N, h->d_grp_tmp.p); ++h->launches;
      size_t tb = h->d_grp_cub.n;
      CK(cub::DeviceRadixSort::SortKeys(h->d_grp_cub.p, tb, h->d_grp_tmp.p, h->d_grp_tmp.p + N + 1, (int)N, 0, 64, h->stream)); ++h->launches;
      tb = h->d_grp_cub.n;
      CK(cub::DeviceRunLengthEncode::Encode(h->d_grp_cub.p, tb, h->d_grp_tmp.p + N + 1, h->d_grp_key[g]->p, h->d_grp_cnt[g]->p, h->d_grp_n.p + g, (int)N, h->stream)); ++h->launches;
    }
  }
  CK(cudaGetLastError());
}

// program of a latent class for a mask of directly observed cells
int latent_prog_pat(Eng* h, int cls, int mask) {
  auto it = h->lprog_of_pat.find(std::make_pair(cls, mask));
  if (it != h->lprog_of_pat.end()) return it->second;
  auto er = h->lprog_pat_error.find(std::make_pair(cls, mask));
  if (er != h->lprog_pat_error.end()) throw Unsupported(er->second);
  return latent_prog(h, cls);
}
const BlockProgram* latent_bp(Eng* h, int cls, int mask) {
  for (size_t li = 0; li < h->lprogs.size(); ++li) if (h->lprog_cls[li] == cls && h->lprog_mask[li] == mask) return &h->lprogs[li];
  throw Unsupported("class has no latent program for this pattern of observed cells");
}

void run_latent_moves(Eng* h, int cls, int slot0, int nslots, uint64_t seed, uint32_t sweep) {
  recount(h);
  refresh_candidate_mats(h);
  CK(cudaMemsetAsync(h->d_pool_count.p, 0, sizeof(int), h->stream));
  TableH& T = h->tables[cls];
  const int nb = (int)h->m.classes[cls].blocks.size();
  auto oc = h->lobs_cells.find(cls);
  h->lpat_active = oc != h->lobs_cells.end();
  h->lpats_present.clear();
  if (!h->lpat_active) {
    if (h->lprog_trivial.count(std::make_pair(cls, 0))) { h->lpat_active = true; return; }    // nothing to enumerate: apply sees no pattern to write back
    const int pid = latent_prog(h, cls);
    build_ref_csr(h, cls);
    int lblock = 0;
    for (size_t li = 0; li < h->lprogs.size(); ++li) if (h->lprog_cls[li] == cls) { lblock = h->lprog_block[li]; break; }
    const int grid = std::min(nblk(nslots, PCL_WARPS_PER_CTA), 148 * 2);
    const int* order = nullptr;
    std::vector<int> by_weight;
    if (nslots > grid * PCL_WARPS_PER_CTA / 4) {
      // heaviest rows first: a row's cost grows with its referrers, popular rows have orders of magnitude
      // more of them, and a warp that meets one late in its stride is the tail of the launch
      CK(cudaStreamSynchronize(h->stream));
      const std::vector<int> rc = T.refcnt.download(T.n_slots);
      by_weight.resize(nslots);
      for (int i = 0; i < nslots; ++i) by_weight[i] = slot0 + i;
      std::stable_sort(by_weight.begin(), by_weight.end(), [&](int a, int b) { return rc[a] > rc[b]; });
      CK(cudaMemcpyAsync(h->d_lslots.p, by_weight.data(), by_weight.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      order = h->d_lslots.p;
    }
    k_latent<<<grid, 32 * PCL_WARPS_PER_CTA, PCL_KLATENT_SMEM, h->stream>>>(h->d_dev.p, pid, lblock, nb, order ? 0 : slot0, nslots, order, seed, sweep, h->cfg.use_mh_instead_of_pg);
    ++h->launches;
    CK(cudaGetLastError());
    if (order) CK(cudaStreamSynchronize(h->stream));      // by_weight (host vector) must outlive its copy
    return;
  }
  // rows are grouped by which of their cells the dataset observes (the reference compiles one
  // proposal per set of present vertices, block_proposal.jl:169-174)
  CK(cudaMemsetAsync(h->d_lpat.p, 0, (size_t)T.cap * sizeof(int), h->stream));
  k_obs_pattern<<<nblk(h->N, 256), 256, 0, h->stream>>>(h->d_dev.p, oc->second, h->N, h->d_lpat.p); ++h->launches;
  CK(cudaStreamSynchronize(h->stream));
  h->h_lpat = h->d_lpat.download(T.n_slots);
  const std::vector<int> rc = T.refcnt.download(T.n_slots);
  std::map<int, std::vector<int>> groups;
  for (int t = slot0; t < slot0 + nslots; ++t) if (rc[t] > 0 && !h->lprog_trivial.count(std::make_pair(cls, h->h_lpat[t]))) groups[h->h_lpat[t]].push_back(t);
  if (groups.empty()) return;
  for (auto& g : groups) std::stable_sort(g.second.begin(), g.second.end(), [&](int a, int b) { return rc[a] > rc[b]; });   // heaviest rows first (see above)                               // every live row has nothing to enumerate (flights TrackingWebsite)
  build_ref_csr(h, cls);
  std::vector<int> all; std::vector<std::pair<int, std::pair<int, int>>> launches;
  for (auto& g : groups) { launches.push_back({g.first, {(int)all.size(), (int)g.second.size()}}); all.insert(all.end(), g.second.begin(), g.second.end()); }
  CK(cudaMemcpyAsync(h->d_lslots.p, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  for (auto& L : launches) {
    const int pid = latent_prog_pat(h, cls, L.first);
    h->lpats_present.push_back(L.first);
    int lblock = 0;
    for (size_t li = 0; li < h->lprogs.size(); ++li) if (h->lprog_cls[li] == cls && h->lprog_mask[li] == L.first) lblock = h->lprog_block[li];
    const int grid = std::min(nblk(L.second.second, PCL_WARPS_PER_CTA), 148 * 2);
    k_latent<<<grid, 32 * PCL_WARPS_PER_CTA, PCL_KLATENT_SMEM, h->stream>>>(h->d_dev.p, pid, lblock, nb, 0, L.second.second, h->d_lslots.p + L.second.first, seed, sweep, h->cfg.use_mh_instead_of_pg);
    ++h->launches;
  }
  CK(cudaStreamSynchronize(h->stream));      // `all` must outlive the copy
  CK(cudaGetLastError());
}

// write the selected particles of a latent class back (table.rows[key] = chosen row), create the
// rows they proposed, then refresh the denormalised copies held by the classes above
// (update_referring_rows_with_new_values_for_updated_row!, dependency_tracking.jl:239-258)
void apply_latent_moves(Eng* h, int cls, int64_t* n_changed, int64_t* n_new) {
  h->mats_dirty = true;
  TableH& T = h->tables[cls];
  const int n = T.n_slots;
  *n_changed = 0; *n_new = 0;
  h->d_counter.zero();
  std::vector<int> pats = h->lpat_active ? h->lpats_present : std::vector<int>{0};
  for (int mask : pats) {
  const int pid = h->lpat_active ? latent_prog_pat(h, cls, mask) : latent_prog(h, cls);
  const BlockProgram* bp = h->lpat_active ? latent_bp(h, cls, mask) : nullptr;
  if (!bp) for (size_t li = 0; li < h->lprogs.size(); ++li) if (h->lprog_cls[li] == cls) { bp = &h->lprogs[li]; break; }
  for (int site = 0; site < (int)bp->roots.size(); ++site) {
    const int ridx = bp->roots[site];
    k_lapply_site<<<nblk(n, 256), 256, 0, h->stream>>>(h->d_dev.p, pid, site, n, h->lpat_active ? h->d_lpat.p : nullptr, mask, h->d_req.p, h->d_counter.p); ++h->launches;
    if (bp->stars[ridx].kind != ST_FK) continue;
    // post-order over the subtree of this site: nested rows first
    std::vector<int> sub;
    std::function<void(int)> po = [&](int s) { for (int c : bp->stars[s].children) po(c); sub.push_back(s); };
    po(ridx);
    bool any = false;
    for (int sidx : sub) {
      const StarL& s = bp->stars[sidx];
      if (s.kind != ST_FK) continue;
      k_create_flags<<<nblk(n + 1, 256), 256, 0, h->stream>>>(h->d_dev.p, pid, sidx, n, h->d_req.p, h->d_flags.p); ++h->launches;
      size_t tmp = h->d_cub_tmp.n;
      CK(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, h->d_flags.p, h->d_rank.p, n + 1, h->stream)); ++h->launches;
      int total = 0;
      CK(cudaMemcpyAsync(&total, h->d_rank.p + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (total == 0) continue;
      any = true;
      TableH& TG = h->tables[s.table];
      h->max_batch_new = std::max(h->max_batch_new, total);
      if (TG.n_slots + total > TG.cap) throw std::runtime_error("latent table capacity exceeded: reserve more rows with pclean_reserve_table / option table_cap (PCLEAN_ERR_CAPACITY)");
      k_create_rows<<<nblk(n, 256), 256, 0, h->stream>>>(h->d_dev.p, pid, sidx, 0, 0, n, h->d_req.p, h->d_flags.p, h->d_rank.p, TG.n_slots, 0, nullptr);
      ++h->launches;
      std::vector<long long> nk(total);
      for (int i = 0; i < total; ++i) { TG.slot_of_key[h->next_key] = TG.n_slots + i; nk[i] = h->next_key; TG.keys.push_back(h->next_key++); }
      CK(cudaMemcpy(TG.d_keys.p + TG.n_slots, nk.data(), total * sizeof(long long), cudaMemcpyHostToDevice));
      TG.n_slots += total; *n_new += total;
      upload_tables(h);
    }
    if (any) { k_lapply_new<<<nblk(n, 256), 256, 0, h->stream>>>(h->d_dev.p, pid, site, n, h->d_req.p); ++h->launches; }
  }
  }
  // classes are defined before their referrers: refresh copies upward in class order
  for (int c2 = 0; c2 < (int)h->tables.size(); ++c2) {
    TableH& T2 = h->tables[c2];
    if (!T2.loaded || c2 == h->obs_cls || T2.n_slots == 0) continue;
    for (size_t g = 0; g < T2.fk_col.size(); ++g) {
      auto key = std::make_pair(c2, (int)g);
      k_refresh_copies<<<nblk(T2.n_slots, 256), 256, 0, h->stream>>>(h->d_tables.p, c2, (int)g, h->fk_copies.at(key)->p, h->fk_ncopies.at(key));
      ++h->launches;
    }
  }
  std::vector<int> cnt = h->d_counter.download(1);
  *n_changed = cnt[0];
  CK(cudaGetLastError());
}


// resample_value!(ProportionsParameter) (choose_proportionally.jl:70-74) for the parameters declared
// in class `cls`, and resample_py_params! (trace.jl:83-107) for its table — once per class sweep
// (the reference does it every rejuv_frequency rows of the sequential scan).
void resample_class_parameters(Eng* h, int cls, uint64_t seed) {
  const Model& m = h->m;
  const ClassM& cm = m.classes[cls];
  TableH& T = h->tables[cls];
  bool changed = false;
  if (cls != h->obs_cls && T.loaded) {
    for (int v = 0; v < cm.n_normal; ++v) {
      const Node& n = cm.nodes[v];
      if (n.wrap != PCLEAN_WRAP_NONE || n.kind != PCLEAN_NODE_CHOICE || n.dist != PCLEAN_DIST_CHOOSE_PROPORTIONALLY) continue;
      const Node& pn = cm.nodes[n.args.at(1)];
      if (pn.kind != PCLEAN_NODE_PARAM || m.param_indexed[pn.param]) continue;
      int slot = -1;
      for (size_t s2 = 0; s2 < m.slot_param.size(); ++s2) if (m.slot_param[s2] == pn.param) slot = (int)s2;
      const Node& ln = cm.nodes[n.args.at(0)];
      if (slot < 0 || ln.kind != PCLEAN_NODE_JULIA || m.funcs[ln.func].kind != PCLEAN_FUNC_CONST) continue;
      const std::vector<Val>& opts = m.lists.at(m.funcs[ln.func].cst.i);
      ParamH& P = h->params[slot];
      if (P.value.size() != opts.size()) continue;
      std::vector<int> ids; for (const Val& o : opts) ids.push_back(o.i);
      DBuf<int> d_ids, d_cnt; d_ids.upload(ids); d_cnt.alloc(ids.size()); d_cnt.zero();
      k_option_counts<<<nblk(T.n_slots, 256), 256, 0, h->stream>>>(h->d_tables.p, cls, v, d_ids.p, (int)ids.size(), d_cnt.p); ++h->launches;
      CK(cudaStreamSynchronize(h->stream));
      std::vector<int> cnt = d_cnt.download();
      ++P.epoch;
      pclean_stream st{}; st.key.seed = seed; st.key.sweep = P.epoch; st.key.row = slot; st.key.purpose = PCLEAN_RNG_PARAM;
      double tot = 0;
      for (size_t i = 0; i < P.value.size(); ++i) { P.value[i] = pclean_next_gamma(&st, m.param_prior0[P.spec] + (double)cnt[i]); tot += P.value[i]; }
      for (double& x : P.value) x /= tot;
      changed = true;
    }
  }
  if (changed) { upload_param_priors(h); compute_hoists(h, true); }
  if (cls == h->obs_cls && !h->gsites.empty()) {
    // MeanParameter.resample_value! (add_noise.jl:74-82): conjugate normal update from the moments
    // of the rows that currently use each slot; moments are all-reduced over the row shards
    const int64_t N = h->N, r0 = h->shard_begin, r1 = h->shard_end < 0 ? h->N : h->shard_end;
    const size_t ns = h->params.size();
    std::vector<double> post_mean(ns), post_var(ns);
    std::vector<char> touched(ns, 0);
    for (auto& G : h->gsites) {
      CK(cudaMemsetAsync(h->d_msum.p, 0, ns * sizeof(double), h->stream));
      CK(cudaMemsetAsync(h->d_mcnt.p, 0, ns * sizeof(double), h->stream));
      k_gauss_site<<<nblk(N, 256), 256, 0, h->stream>>>(h->d_dev.p, G.d, r0, r1, N, h->d_slot_of_row.p, h->d_x_of_row.p, h->d_iota.p); ++h->launches;
      size_t sb = h->d_sort_tmp.n;
      CK(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, sb, h->d_slot_of_row.p, h->d_req.p, h->d_iota.p, h->d_lref_rows.p, (int)N, 0, 32, h->stream)); ++h->launches;
      k_segment_moments<<<nblk(N, 256), 256, 0, h->stream>>>(h->d_req.p, h->d_lref_rows.p, N, h->d_x_of_row.p, h->d_msum.p, h->d_mcnt.p); ++h->launches;
      if (h->nccl.comm) {
        if (h->nccl.AllReduce(h->d_msum.p, h->d_msum.p, ns, /*ncclFloat64*/ 8, 0, h->nccl.comm, h->stream) != 0 ||
            h->nccl.AllReduce(h->d_mcnt.p, h->d_mcnt.p, ns, 8, 0, h->nccl.comm, h->stream) != 0) throw std::runtime_error("ncclAllReduce failed");
      }
      CK(cudaStreamSynchronize(h->stream));
      const std::vector<double> ms = h->d_msum.download(ns), mc = h->d_mcnt.download(ns);
      for (int slot : G.slots) {
        ParamH& P = h->params[slot];
        if (!touched[slot]) { touched[slot] = 1; post_mean[slot] = m.param_prior0[P.spec]; post_var[slot] = m.param_prior1[P.spec] * m.param_prior1[P.spec]; }
        if (mc[slot] <= 0) continue;
        const double sd2 = G.stdev * G.stdev;
        const double nv = 1.0 / (1.0 / post_var[slot] + mc[slot] / sd2);
        post_mean[slot] = nv * (post_mean[slot] / post_var[slot] + ms[slot] / sd2);
        post_var[slot] = nv;
      }
    }
    std::vector<double> preal = h->d_param_real.download(std::max<size_t>(1, ns));
    for (size_t slot = 0; slot < ns; ++slot) {
      if (!touched[slot]) continue;
      ParamH& P = h->params[slot];
      ++P.epoch;
      pclean_stream st{}; st.key.seed = seed; st.key.sweep = P.epoch; st.key.row = (int64_t)slot; st.key.purpose = PCLEAN_RNG_PARAM;
      P.value = {post_mean[slot] + std::sqrt(post_var[slot]) * pclean_next_normal(&st)};
      preal[slot] = P.value[0];
    }
    CK(cudaMemcpy(h->d_param_real.p, preal.data(), preal.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (cls == h->obs_cls && !h->msites.empty()) {
    // ProbParameter.resample_value! (maybe_swap.jl:87-89): Beta(a + differing, b + equal) per slot
    const int64_t r0 = h->shard_begin, r1 = h->shard_end < 0 ? h->N : h->shard_end;
    const size_t ns = h->params.size();
    DBuf<int> d_cnt; d_cnt.alloc(2 * ns); d_cnt.zero();
    for (const MswapD& M : h->msites) { k_mswap_counts<<<nblk(r1 - r0, 256), 256, 0, h->stream>>>(h->d_dev.p, M, r0, r1, d_cnt.p); ++h->launches; }
    if (h->nccl.comm && h->nccl.AllReduce(d_cnt.p, d_cnt.p, 2 * ns, /*ncclInt32*/ 2, 0, h->nccl.comm, h->stream) != 0) throw std::runtime_error("ncclAllReduce failed");
    CK(cudaStreamSynchronize(h->stream));
    const std::vector<int> cnt = d_cnt.download();
    std::vector<double> preal = h->d_param_real.download(std::max<size_t>(1, ns));
    for (size_t slot = 0; slot < ns; ++slot) {
      ParamH& P = h->params[slot];
      if (m.param_kind[P.spec] != PCLEAN_PARAM_PROB) continue;
      ++P.epoch;
      pclean_stream st{}; st.key.seed = seed; st.key.sweep = P.epoch; st.key.row = (int64_t)slot; st.key.purpose = PCLEAN_RNG_PARAM;
      P.value = {pclean_next_beta(&st, m.param_prior0[P.spec] + (double)cnt[2 * slot], m.param_prior1[P.spec] + (double)cnt[2 * slot + 1])};
      preal[slot] = P.value[0];
    }
    CK(cudaMemcpy(h->d_param_real.p, preal.data(), preal.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  // Pitman-Yor hyper-parameters: independent MH on strength (proposal Gamma(1,1)) then discount (Uniform)
  if (cls != h->obs_cls && T.loaded && T.n_slots > 0) {
    std::vector<int> rc = T.refcnt.download(T.n_slots);
    std::vector<long long> counts; long long N = 0;
    for (int c : rc) if (c > 0) { counts.push_back(c); N += c; }
    auto score = [&](double s, double d) {      // pitman_yor_score (trace.jl:65-81), sums in closed form
      double lp = 0; long long j = 0;
      for (long long size : counts) { ++j; lp += std::log(j * d + s) + std::lgamma((double)size - d) - std::lgamma(1.0 - d); }
      return lp - (std::lgamma((double)N + s) - std::lgamma(s));
    };
    ++T.py_epoch;
    pclean_stream st{}; st.key.seed = seed; st.key.sweep = T.py_epoch; st.key.cls = (uint32_t)cls; st.key.row = cls; st.key.purpose = PCLEAN_RNG_PY;
    double cs = T.strength, cd = T.discount;
    double old_score = score(cs, cd);
    double u = pclean_next(&st); if (u < 1e-300) u = 1e-300;
    const double proposed = -std::log(u);
    double new_score = score(proposed, cd);
    if (std::log(pclean_next(&st)) < new_score + (-cs) - old_score - (-proposed)) { cs = proposed; old_score = new_score; }
    const double pd = pclean_next(&st);
    new_score = score(cs, pd);
    if (std::log(pclean_next(&st)) < new_score - old_score) cd = pd;
    T.strength = cs; T.discount = cd;
    h->h_tables[cls].strength = cs; h->h_tables[cls].discount = cd;
    // only the two scalars change on the device (counts are recomputed by recount())
    CK(cudaMemcpy((char*)(h->d_tables.p + cls) + offsetof(TableD, strength), &cs, sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy((char*)(h->d_tables.p + cls) + offsetof(TableD, discount), &cd, sizeof(double), cudaMemcpyHostToDevice));
  }
}

void check_device_error(Eng* h) {
  int e = 0;
  CK(cudaMemcpy(&e, h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (e != 0) { h->d_err.zero(); throw std::runtime_error("device reported error code " + std::to_string(e)); }
}

}  // namespace

// ---- on-disk IR ("PCLIRv1", pclean_b200/irfile.py and julia/PCleanB200.jl write it): named arrays,
// one per pointer field of pclean_model_ir / pclean_observations, scalars as int32[1]
namespace {
struct IrFile {
  std::vector<char> buf;
  std::map<std::string, std::pair<const char*, uint64_t>> e;      // name -> (payload, count)
  std::map<std::string, uint32_t> dtype;
  void read(const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) throw BadArg(std::string("cannot open ") + path);
    std::fseek(f, 0, SEEK_END); const long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    buf.resize((size_t)std::max<long>(n, 0) + 8);
    const size_t got = n > 0 ? std::fread(buf.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    if (got != (size_t)n || n < 12 || std::memcmp(buf.data(), "PCLIRv1\n", 8) != 0) throw BadArg(std::string(path) + ": not a PCLIRv1 file");
    static const size_t item[6] = {4, 8, 8, 4, 16, 1};
    uint32_t cnt; std::memcpy(&cnt, buf.data() + 8, 4);
    size_t pos = 12;
    for (uint32_t i = 0; i < cnt; ++i) {
      if (pos + 4 > (size_t)n) throw BadArg("truncated IR file");
      uint32_t ln; std::memcpy(&ln, buf.data() + pos, 4); pos += 4;
      if (pos + ln + 12 > (size_t)n) throw BadArg("truncated IR file");
      std::string name(buf.data() + pos, ln); pos += ln;
      uint32_t dt; uint64_t count; std::memcpy(&dt, buf.data() + pos, 4); std::memcpy(&count, buf.data() + pos + 4, 8); pos += 12;
      pos = (pos + 7) & ~(size_t)7;
      if (dt > 5 || pos + count * item[dt] > (size_t)n) throw BadArg("corrupt IR file entry " + name);
      e[name] = std::make_pair(buf.data() + pos, count); dtype[name] = dt;
      pos += count * item[dt];
    }
  }
  template <class T> const T* arr(const char* name, uint32_t want) const {
    auto it = e.find(name);
    if (it == e.end()) throw BadArg(std::string("IR file lacks entry ") + name);
    if (dtype.at(name) != want) throw BadArg(std::string("IR file entry of the wrong type: ") + name);
    return reinterpret_cast<const T*>(it->second.first);
  }
  int32_t scalar(const char* name) const { return *arr<int32_t>(name, 0); }
};
}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* pclean_version(void) { return "pclean_b200 0.1 (sm_100a)"; }

int32_t pclean_create(const pclean_config* cfg, int32_t device, pclean_engine** out) {
  if (!cfg || !out) return PCLEAN_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return PCLEAN_ERR_CUDA;      // no CPU fallback: fail loudly
  if (device < 0 || device >= n) return PCLEAN_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return PCLEAN_ERR_CUDA;
  pclean_engine* h = new pclean_engine();
  h->cfg = *cfg; h->device = device;
  if (h->cfg.use_mh_instead_of_pg) h->cfg.num_particles = 2;   // infer_config.jl:11-13
  if (cudaStreamCreate(&h->stream) != cudaSuccess) { delete h; return PCLEAN_ERR_CUDA; }
  prepare_k_block(h);
  cudaFuncSetAttribute(k_latent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCL_KLATENT_SMEM);   // per device, hence per engine
  cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1); cudaEventCreate(&h->ev2); cudaEventCreate(&h->ev3);
  for (int i = 0; i < 16; ++i) cudaEventCreate(&h->evb[i]);
  *out = h;
  return PCLEAN_OK;
}

int32_t pclean_destroy(pclean_engine* h) {
  if (!h) return PCLEAN_ERR_ARG;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->ev0) { cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->ev2); cudaEventDestroy(h->ev3); }
  for (int i = 0; i < 16; ++i) if (h->evb[i]) cudaEventDestroy(h->evb[i]);
  if (h->pinned) { for (int* p : h->pinned->host) cudaFreeHost(p); delete h->pinned; }
  if (h->nccl.comm && h->nccl.owned && h->nccl.lib) {
    typedef int (*destroy_t)(void*);
    destroy_t f = (destroy_t)dlsym(h->nccl.lib, "ncclCommDestroy");
    if (f) f(h->nccl.comm);
  }
  if (h->nccl.lib) dlclose(h->nccl.lib);
  delete h;
  return PCLEAN_OK;
}

const char* pclean_last_error(const pclean_engine* h) { return h ? h->err.c_str() : "null handle"; }

int32_t pclean_load_model(pclean_engine* h, const pclean_model_ir* ir) {
  if (!h || !ir) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    parse_model(ir, h->m);
    h->ir_n_paths = ir->n_paths;
    h->ir_path_target.assign(ir->path_target, ir->path_target + ir->n_paths);
    h->ir_path_len_off.assign(ir->path_len_off, ir->path_len_off + ir->n_paths + 1);
    h->ir_path_class.assign(ir->path_class, ir->path_class + ir->path_len_off[ir->n_paths]);
    h->ir_path_vertex.assign(ir->path_vertex, ir->path_vertex + ir->path_len_off[ir->n_paths]);
    h->ir_path_vmap_off.assign(ir->path_vmap_off, ir->path_vmap_off + ir->n_paths + 1);
    h->ir_path_vmap.assign(ir->path_vmap, ir->path_vmap + ir->path_vmap_off[ir->n_paths]);
    std::memset(&h->ir_view, 0, sizeof(h->ir_view));
    h->ir_view.n_paths = h->ir_n_paths; h->ir_view.path_target = h->ir_path_target.data(); h->ir_view.path_len_off = h->ir_path_len_off.data();
    h->ir_view.path_class = h->ir_path_class.data(); h->ir_view.path_vertex = h->ir_path_vertex.data();
    h->ir_view.path_vmap_off = h->ir_path_vmap_off.data(); h->ir_view.path_vmap = h->ir_path_vmap.data();
    h->strings = h->m.strings; h->string_ids.clear();
    for (size_t i = 0; i < h->strings.size(); ++i) h->string_ids.emplace(h->strings[i], (int)i);
    h->tables.clear(); h->tables.resize(h->m.classes.size());
    for (size_t c = 0; c < h->tables.size(); ++c) { h->tables[c].cls = (int)c; h->tables[c].strength = h->m.classes[c].py_strength; h->tables[c].discount = h->m.classes[c].py_discount; }
    h->params.assign(h->m.slot_param.size(), ParamH());
    h->model_loaded = true; h->finalized = false;
  });
}

/* load_model from a PCLIRv1 file (the flat IR a host wrote with pclean_b200/irfile.py or julia/PCleanB200.jl) */
int32_t pclean_load_model_file(pclean_engine* h, const char* path) {
  if (!h || !path) return PCLEAN_ERR_ARG;
  IrFile F;
  const int32_t rc = guard(h, [&] { F.read(path); });
  if (rc != PCLEAN_OK) return rc;
  pclean_model_ir ir; std::memset(&ir, 0, sizeof(ir));
  const int32_t rc2 = guard(h, [&] {
#define PCL_S(name) ir.name = F.scalar(#name)
#define PCL_A(name, T, code) ir.name = F.arr<T>(#name, code)
    PCL_S(n_classes); PCL_S(n_vertices); PCL_S(n_blocks); PCL_S(n_paths); PCL_S(n_funcs); PCL_S(n_params); PCL_S(n_param_slots);
    PCL_S(n_lists); PCL_S(n_xforms); PCL_S(n_strings);
    PCL_A(class_voff, int32_t, 0); PCL_A(py_strength, double, 2); PCL_A(py_discount, double, 2);
    PCL_A(v_kind, int32_t, 0); PCL_A(v_wrap, int32_t, 0); PCL_A(v_wrap_off, int32_t, 0); PCL_A(wrap_fk, int32_t, 0); PCL_A(wrap_subid, int32_t, 0);
    PCL_A(v_dist, int32_t, 0); PCL_A(v_args_off, int32_t, 0); PCL_A(v_args, int32_t, 0); PCL_A(v_func, int32_t, 0); PCL_A(v_target, int32_t, 0);
    PCL_A(v_vmap_off, int32_t, 0); PCL_A(v_vmap, int32_t, 0); PCL_A(v_param, int32_t, 0); PCL_A(v_path, int32_t, 0); PCL_A(v_extv, int32_t, 0);
    PCL_A(class_block_off, int32_t, 0); PCL_A(block_voff, int32_t, 0); PCL_A(block_v, int32_t, 0);
    PCL_A(plan_off, int32_t, 0); PCL_A(plan_vertex, int32_t, 0); PCL_A(plan_nchild, int32_t, 0);
    PCL_A(class_hash_off, int32_t, 0); PCL_A(hash_v, int32_t, 0);
    PCL_A(path_target, int32_t, 0); PCL_A(path_len_off, int32_t, 0); PCL_A(path_class, int32_t, 0); PCL_A(path_vertex, int32_t, 0);
    PCL_A(path_vmap_off, int32_t, 0); PCL_A(path_vmap, int32_t, 0);
    PCL_A(func_kind, int32_t, 0); PCL_A(func_const, pclean_value, 4); PCL_A(func_keyarg_off, int32_t, 0); PCL_A(func_keyargs, int32_t, 0);
    PCL_A(func_tab_off, int32_t, 0); PCL_A(tab_keys, int32_t, 0); PCL_A(tab_key_off, int64_t, 1); PCL_A(tab_vals, pclean_value, 4);
    PCL_A(param_kind, int32_t, 0); PCL_A(param_indexed, int32_t, 0); PCL_A(param_prior0, double, 2); PCL_A(param_prior1, double, 2);
    PCL_A(slot_param, int32_t, 0); PCL_A(list_off, int64_t, 1); PCL_A(list_vals, pclean_value, 4); PCL_A(xform_scale, double, 2);
    PCL_A(str_off, int64_t, 1); PCL_A(str_cp, uint32_t, 3); PCL_A(lm_unigram, double, 2); PCL_A(lm_bigram, double, 2);
#undef PCL_S
#undef PCL_A
  });
  if (rc2 != PCLEAN_OK) return rc2;
  return pclean_load_model(h, &ir);         // copies everything it keeps: F may go away
}

/* load_observations from the "obs.*" entries of a PCLIRv1 file */
int32_t pclean_load_observations_file(pclean_engine* h, const char* path) {
  if (!h || !path) return PCLEAN_ERR_ARG;
  IrFile F;
  pclean_observations obs; std::memset(&obs, 0, sizeof(obs));
  const int32_t rc = guard(h, [&] {
    F.read(path);
    obs.cls = F.scalar("obs.cls"); obs.n_rows = *F.arr<int64_t>("obs.n_rows", 1); obs.n_cols = F.scalar("obs.n_cols");
    obs.vertex_of_col = F.arr<int32_t>("obs.vertex_of_col", 0); obs.cells = F.arr<pclean_value>("obs.cells", 4);
    if (F.e.at("obs.cells").second != (uint64_t)obs.n_rows * (uint64_t)obs.n_cols || F.e.at("obs.vertex_of_col").second != (uint64_t)obs.n_cols)
      throw BadArg("obs.cells / obs.vertex_of_col do not match obs.n_rows x obs.n_cols");
  });
  if (rc != PCLEAN_OK) return rc;
  return pclean_load_observations(h, &obs);
}

int32_t pclean_load_observations(pclean_engine* h, const pclean_observations* obs) {
  if (!h || !obs) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (!h->model_loaded) throw std::runtime_error("load the model first");
    if (h->obs_cls >= 0) throw Unsupported("more than one observed dataset");
    CK(cudaSetDevice(h->device));
    h->obs_cls = obs->cls; h->N = obs->n_rows;
    for (int c = 0; c < obs->n_cols; ++c) {
      std::unique_ptr<ObsCol> oc(new ObsCol());
      oc->vertex = obs->vertex_of_col[c];
      oc->sid.assign(h->N, -1); oc->uobs.assign(h->N, -1); oc->absent.assign(h->N, 0);
      std::unordered_map<int, int> uniq;
      for (int64_t r = 0; r < h->N; ++r) {
        const pclean_value& v = obs->cells[(size_t)c * h->N + r];
        if (v.tag == PCLEAN_VAL_STR) {
          oc->sid[r] = v.i;
          auto it = uniq.find(v.i);
          if (it == uniq.end()) { it = uniq.emplace(v.i, (int)oc->ulist.size()).first; oc->ulist.push_back(v.i); }
          oc->uobs[r] = it->second;
        } else if (v.tag == PCLEAN_VAL_REAL || v.tag == PCLEAN_VAL_INT) {
          if (!oc->is_real) { oc->is_real = true; oc->real.assign(h->N, std::nan("")); }
          oc->real[r] = v.tag == PCLEAN_VAL_REAL ? v.d : (double)v.i;
        } else if (v.tag == PCLEAN_VAL_MISSING) { /* explicit missing observation: sid = uobs = -1 */ }
        else if (v.tag == PCLEAN_VAL_ABSENT) oc->absent[r] = 1;            // not an observation of this row
        else throw Unsupported("observation cell of an unsupported kind");
      }
      if (oc->is_real && !oc->ulist.empty()) throw Unsupported("dataset column mixing strings and numbers");
      oc->d_uobs.upload(oc->uobs); oc->d_ulist.upload(oc->ulist); oc->d_sid.upload(oc->sid);
      if (oc->is_real) oc->d_real.upload(oc->real);
      h->col_of_vertex[oc->vertex] = (int)h->cols.size();
      h->cols.push_back(std::move(oc));
    }
    // missingness patterns (block_proposal.jl:169-170: one compiled proposal per set of present vertices)
    h->pat_of_row.assign(h->N, 0); h->pat_cols.clear(); h->pat_rows.clear();
    {
      std::map<std::vector<char>, int> ids;
      for (int64_t r = 0; r < h->N; ++r) {
        std::vector<char> key(h->cols.size());
        for (size_t c = 0; c < h->cols.size(); ++c) key[c] = !h->cols[c]->absent[r];
        auto it = ids.find(key);
        if (it == ids.end()) { it = ids.emplace(key, (int)h->pat_cols.size()).first; h->pat_cols.push_back(key); h->pat_rows.emplace_back(); }
        h->pat_of_row[r] = it->second;
        h->pat_rows[it->second].push_back(r);
      }
      if (h->pat_cols.size() > 32) throw Unsupported("more than 32 distinct missingness patterns");
    }
    h->finalized = false;
  });
}

int32_t pclean_load_table(pclean_engine* h, const pclean_table_snapshot* t) {
  if (!h || !t) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (!h->model_loaded) throw std::runtime_error("load the model first");
    if (t->cls < 0 || t->cls >= (int)h->tables.size()) throw BadArg("class index out of range");
    TableH& T = h->tables[t->cls];
    {
      const ClassM& tm = h->m.classes[t->cls];
      if (t->n_rows < 0 || (t->n_rows > 0 && (!t->keys || !t->cells))) throw BadArg("table snapshot without keys / cells");
      if (t->n_rows > 0 && (t->n_cols < tm.n_normal || t->n_cols > tm.nv)) throw BadArg("table snapshot must carry one column per (non-external) vertex of the class");
    }
    T.keys.assign(t->keys, t->keys + t->n_rows);
    T.slot_of_key.clear();
    for (int64_t r = 0; r < t->n_rows; ++r) T.slot_of_key[t->keys[r]] = (int)r;
    T.raw.assign(t->cells, t->cells + (size_t)t->n_cols * t->n_rows);
    T.raw_cols = t->n_cols;
    T.strength = t->py_strength; T.discount = t->py_discount;
    T.loaded = true; h->finalized = false;
  });
}

int32_t pclean_load_assignment(pclean_engine* h, int32_t cls, int64_t n_rows, int32_t n_fk, const int32_t* fk_vertices, const int64_t* keys) {
  if (!h || !fk_vertices || !keys) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (cls != h->obs_cls || n_rows != h->N) throw BadArg("assignment does not match the observed dataset");
    for (int f = 0; f < n_fk; ++f) h->assign_keys[fk_vertices[f]].assign(keys + (size_t)f * n_rows, keys + (size_t)(f + 1) * n_rows);
    h->finalized = false;
  });
}

int32_t pclean_set_param_values(pclean_engine* h, int32_t slot, int32_t n, const double* values) {
  if (!h || !values) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (slot < 0 || slot >= (int)h->params.size()) throw BadArg("parameter slot out of range");
    if (n < 0) throw BadArg("negative value count");
    h->params[slot].value.assign(values, values + n);
    if (h->finalized) {
      CK(cudaSetDevice(h->device)); upload_param_priors(h); compute_hoists(h, true);
      if (n == 1 && h->d_param_real.n > (size_t)slot) CK(cudaMemcpy(h->d_param_real.p + slot, values, sizeof(double), cudaMemcpyHostToDevice));
    }
  });
}

int32_t pclean_get_param_values(pclean_engine* h, int32_t slot, int32_t cap, double* values, int32_t* n) {
  if (!h || !values || !n) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (slot < 0 || slot >= (int)h->params.size()) throw BadArg("parameter slot out of range");
    const auto& v = h->params[slot].value;
    *n = (int)v.size();
    for (int i = 0; i < std::min<int>(cap, (int)v.size()); ++i) values[i] = v[i];
  });
}

/* rows to reserve for one latent table before pclean_init_trace (candidate distance matrices are
   sized by it: unique observed strings x reserved rows bytes per likelihood term) */
int32_t pclean_reserve_table(pclean_engine* h, int32_t cls, int32_t rows) {
  if (!h || cls < 0 || cls >= (int)h->tables.size() || rows < 16) return PCLEAN_ERR_ARG;
  h->tables[cls].reserve = rows;
  return PCLEAN_OK;
}

/* initialize_trace (inference.jl:3-58) as batched SMC: the reference adds the rows one at a time
   (run_smc! without a retained particle) to tables that start empty; here rows [done, b) are moved
   together against the tables built from rows [0, done), b - done = max(1, done / 2), with the
   same row kernel as a sweep (csmc = 0).  Option "batch_rows" = 1 gives the reference's sequential
   order exactly.  Parameters / PY hyper-parameters are rejuvenated whenever a multiple of
   rejuv_frequency rows has been passed (inference.jl:37-44). */
int32_t pclean_init_trace(pclean_engine* h, uint64_t seed) {
  if (!h) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (!h->model_loaded || h->N <= 0) throw std::runtime_error("load the model and the observations first");
    CK(cudaSetDevice(h->device));
    // On a row-sharded engine the initialisation runs REPLICATED: every rank adds all the rows with the
    // same keyed uniforms (the kernels are deterministic), so the replicas end with identical traces and
    // nothing is exchanged — like the latent-class sweeps.  The shard and the
    // communicator are put back afterwards.
    struct Restore {
      pclean_engine* h; void* comm; int64_t b, e;
      ~Restore() { h->nccl.comm = comm; h->shard_begin = b; h->shard_end = e; }
    } restore{h, h->nccl.comm, h->shard_begin, h->shard_end};
    h->nccl.comm = nullptr; h->shard_begin = 0; h->shard_end = -1;
    const Model& m = h->m;
    const ClassM& cm = m.classes[h->obs_cls];
    for (int c = 0; c < (int)h->tables.size(); ++c) {
      if (c == h->obs_cls) continue;
      TableH& T = h->tables[c];
      T.keys.clear(); T.slot_of_key.clear(); T.raw.clear(); T.raw_cols = 0; T.strength = 1.0; T.discount = 0.0;   // builder.jl:39
      T.py_epoch = 0;
      T.min_cap = (int)std::min<int64_t>(h->N + 1024, T.reserve > 0 ? T.reserve : h->table_cap);
      T.loaded = true;
    }
    h->assign_keys.clear();
    for (int v = 0; v < cm.n_normal; ++v)
      if (cm.nodes[v].wrap == PCLEAN_WRAP_NONE && cm.nodes[v].kind == PCLEAN_NODE_FK) h->assign_keys[v].assign(h->N, INT64_MIN);
    // a fresh trace: every parameter starts from its keyed prior draw (trace.jl:28, distributions.jl:45-61)
    h->param_seed = seed;
    for (auto& P : h->params) { P.value.clear(); P.epoch = 0; }
    h->finalized = false;
    finalize(h);
    h->launches = 0;
    const int64_t N = h->init_rows > 0 ? std::min<int64_t>(h->N, h->init_rows) : h->N;
    const int64_t rejuv = std::max(1, h->cfg.rejuv_frequency);
    // batched mode visits the rows in a scattered order (i * stride mod N, stride coprime with N near
    // the golden ratio): datasets sorted by entity would otherwise put many rows of one new entity
    // into the same batch, where they cannot see each other and each create their own copy
    int64_t stride = 1;
    if (h->batch_rows != 1 && N > 2) {
      stride = std::max<int64_t>(1, (int64_t)(0.6180339887 * (double)N));
      auto gcd = [](int64_t a, int64_t b2) { while (b2) { const int64_t t = a % b2; a = b2; b2 = t; } return a; };
      while (gcd(stride, N) != 1) ++stride;
    }
    int64_t done = 0;
    std::vector<long long> rowsv;
    while (done < N) {
      const int64_t step = h->batch_rows > 0 ? h->batch_rows : std::max<int64_t>(1, done / h->init_divisor);
      const int64_t b = std::min(N, done + step);
      rowsv.clear();
      for (int64_t i = done; i < b; ++i) rowsv.push_back((long long)((i * stride) % N));
      std::sort(rowsv.begin(), rowsv.end());
      const std::vector<long long>* rl = stride == 1 ? nullptr : &rowsv;
      recount(h);
      refresh_candidate_mats(h);
      run_row_moves(h, done, b, seed, 0, false, rl);
      int64_t ch = 0, cr = 0;
      apply_moves(h, done, b, false, &ch, &cr, rl);
      if (cr) intern_new_strings(h);
      CK(cudaStreamSynchronize(h->stream));
      check_device_error(h);
      h->total_new_rows += cr;
      if (h->resample_params && b / rejuv > done / rejuv) {
        for (int c = 0; c < (int)h->tables.size(); ++c) {
          if (c != h->obs_cls && !h->tables[c].loaded) continue;
          recount(h); CK(cudaStreamSynchronize(h->stream));
          const int64_t sb = h->shard_begin, se = h->shard_end;
          h->shard_begin = 0; h->shard_end = h->N;              // statistics over the rows initialised so far (the others are skipped: no assignment yet)
          try { resample_class_parameters(h, c, seed); } catch (...) { h->shard_begin = sb; h->shard_end = se; throw; }
          h->shard_begin = sb; h->shard_end = se;
        }
      }
      done = b;
    }
    recount(h);
    refresh_candidate_mats(h);
    CK(cudaStreamSynchronize(h->stream));
    h->row_state_synced = true;          // every replica holds every row's state
  });
}

static void sweep_obs_class(pclean_engine* h, uint64_t seed, uint32_t sweep_idx, pclean_sweep_stats* out) {
  h->row_state_synced = false;
  const int64_t r0 = h->shard_begin, r1 = h->shard_end < 0 ? h->N : h->shard_end;
  // sequential mode issues one set of collectives per batch: uneven shards would issue different
  // numbers of them and deadlock, and a sequential scan cannot be row-sharded anyway
  if (h->batch_rows > 0 && h->nccl.comm) throw Unsupported("batch_rows > 0 (sequential order) on a row-sharded engine");
  const int64_t step = h->batch_rows > 0 ? h->batch_rows : std::max<int64_t>(1, r1 - r0);
  int64_t changed = 0, created = 0;
  float kernel_ms = 0, total_ms = 0;
  for (int64_t a = r0; a < r1; a += step) {
    const int64_t b = std::min(r1, a + step);
    CK(cudaEventRecord(h->ev0, h->stream));
    recount(h);
    if (h->resample_params && a == r0) resample_class_parameters(h, h->obs_cls, seed);
    refresh_candidate_mats(h);
    CK(cudaEventRecord(h->ev1, h->stream));
    run_row_moves(h, a, b, seed, sweep_idx, true);
    CK(cudaEventRecord(h->ev2, h->stream));
    int64_t ch = 0, cr = 0;
    apply_moves(h, a, b, true, &ch, &cr);
    if (cr) { intern_new_strings(h); refresh_candidate_mats(h); }
    CK(cudaEventRecord(h->ev3, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    changed += ch; created += cr;
    float ms = 0; cudaEventElapsedTime(&ms, h->ev1, h->ev2); kernel_ms += ms;
    cudaEventElapsedTime(&ms, h->ev0, h->ev3); total_ms += ms;
  }
  h->total_new_rows += created;
  if (out) {
    out->rows += r1 - r0; out->particles += (r1 - r0) * h->K; out->new_rows += created; out->changed_rows += changed;
    out->kernel_ms += kernel_ms; out->total_ms += total_ms;
    for (int b = 0; b < std::min(8, h->n_blocks); ++b) cudaEventElapsedTime(&h->block_ms[b], h->evb[2 * b], h->evb[2 * b + 1]);
    const int SB = 296;
    k_sweep_stats<<<SB, 256, 0, h->stream>>>(h->d_row_flags.p, h->d_row_logml.p, r0, r1, ROWFLAG_DUMMY, h->d_stats2.p); ++h->launches;
    double st2[2 * 296];
    CK(cudaMemcpyAsync(st2, h->d_stats2.p, sizeof(st2), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < SB; ++b) { out->sum_log_ml += st2[2 * b]; out->dummy_draws += (int64_t)st2[2 * b + 1]; }
  }
}

// Row-sharded engines: every rank moved only its own observation rows, so before a latent class is
// swept (replicated, identically on every rank: the kernels and the keyed RNG are deterministic)
// the per-row state — reference slots and local cells — is all-gathered over NVLink.
static void gather_row_state(pclean_engine* h) {
  if (!h->nccl.comm || h->row_state_synced) return;
  const int world = h->nccl.world, rank = h->nccl.rank;
  const int64_t N = h->N, r0 = h->shard_begin, r1 = h->shard_end < 0 ? N : h->shard_end;
  // every rank's range
  DBuf<long long> d_rng; d_rng.alloc(2 * world);
  const long long mine[2] = {(long long)r0, (long long)r1};
  CK(cudaMemcpyAsync(d_rng.p + 2 * rank, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  if (h->nccl.AllGather(d_rng.p + 2 * rank, d_rng.p, 2, /*ncclInt64*/ 4, h->nccl.comm, h->stream) != 0) throw std::runtime_error("ncclAllGather failed");
  CK(cudaStreamSynchronize(h->stream));
  const std::vector<long long> rng = d_rng.download();
  int64_t chunk = 0;
  for (int k = 0; k < world; ++k) chunk = std::max<int64_t>(chunk, rng[2 * k + 1] - rng[2 * k]);
  if (chunk == 0) return;
  DBuf<int> send, recv; send.alloc(chunk); recv.alloc((size_t)chunk * world);
  auto gather = [&](int* arr) {
    if (!arr) return;
    CK(cudaMemcpyAsync(send.p, arr + r0, (size_t)(r1 - r0) * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    if (h->nccl.AllGather(send.p, recv.p, (size_t)chunk, /*ncclInt32*/ 2, h->nccl.comm, h->stream) != 0) throw std::runtime_error("ncclAllGather failed");
    for (int k = 0; k < world; ++k) {
      const int64_t a = rng[2 * k], b = rng[2 * k + 1];
      if (k == rank || b <= a) continue;
      CK(cudaMemcpyAsync(arr + a, recv.p + (size_t)k * chunk, (size_t)(b - a) * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    }
  };
  for (int b = 0; b < h->n_blocks; ++b) if (h->progs[b].root >= 0) gather(h->d_assign[b]->p);
  for (auto& rc : h->d_rowcell) if (rc && rc->p && rc->n >= (size_t)N) gather(rc->p);
  CK(cudaStreamSynchronize(h->stream));
  h->row_state_synced = true;
}

static void sweep_latent_class(pclean_engine* h, int cls, uint64_t seed, uint32_t sweep_idx, pclean_sweep_stats* out) {
  gather_row_state(h);
  TableH& T = h->tables[cls];
  CK(cudaEventRecord(h->ev0, h->stream));
  if (h->resample_params) { recount(h); CK(cudaStreamSynchronize(h->stream)); resample_class_parameters(h, cls, seed); }
  run_latent_moves(h, cls, 0, T.n_slots, seed, sweep_idx);
  CK(cudaEventRecord(h->ev2, h->stream));
  int64_t changed = 0, created = 0;
  apply_latent_moves(h, cls, &changed, &created);
  refresh_candidate_mats(h);
  CK(cudaEventRecord(h->ev3, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  check_device_error(h);
  if (out) {
    out->new_rows += created; out->changed_rows += changed;
    float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev2); out->kernel_ms += ms;
    cudaEventElapsedTime(&ms, h->ev0, h->ev3); out->total_ms += ms;
    std::vector<int> fl = h->d_lflags.download(T.n_slots);
    for (int f : fl) out->dummy_draws += (f & ROWFLAG_DUMMY) ? 1 : 0;
  }
}

int32_t pclean_sweep(pclean_engine* h, int32_t cls, uint64_t seed, uint32_t sweep_idx, pclean_sweep_stats* out) {
  if (!h) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    h->launches = 0;
    if (out) std::memset(out, 0, sizeof(*out));
    auto pack = [&] { const bool force = h->compact_now; h->compact_now = false; compact_tables(h, force); };
    if (cls >= 0) {
      pack();
      if (cls == h->obs_cls) sweep_obs_class(h, seed, sweep_idx, out);
      else sweep_latent_class(h, cls, seed, sweep_idx, out);
    } else {
      // pgibbs_sweep! (inference.jl:60-81): every class in class_order
      for (int c = 0; c < (int)h->tables.size(); ++c) {
        if (c == h->obs_cls || h->tables[c].loaded) pack();
        if (c == h->obs_cls) sweep_obs_class(h, seed, sweep_idx, out);
        else if (h->tables[c].loaded) sweep_latent_class(h, c, seed, sweep_idx, out);
      }
    }
    if (out) out->launches = h->launches;
  });
}

int32_t pclean_run_inference(pclean_engine* h, uint64_t seed, pclean_sweep_stats* out_total) {
  if (!h) return PCLEAN_ERR_ARG;
  pclean_sweep_stats tot{}; std::memset(&tot, 0, sizeof(tot));
  for (int it = 0; it < h->cfg.num_iters; ++it) {
    pclean_sweep_stats s{};
    const int32_t rc = pclean_sweep(h, -1, seed, (uint32_t)(it + 1), &s);
    if (rc != PCLEAN_OK) return rc;
    tot.rows += s.rows; tot.particles += s.particles; tot.new_rows += s.new_rows; tot.dummy_draws += s.dummy_draws;
    tot.changed_rows += s.changed_rows; tot.sum_log_ml += s.sum_log_ml; tot.kernel_ms += s.kernel_ms; tot.total_ms += s.total_ms; tot.launches += s.launches;
  }
  if (out_total) *out_total = tot;
  return PCLEAN_OK;
}

int32_t pclean_row_move_debug(pclean_engine* h, int32_t cls, int64_t row, uint64_t seed, uint32_t sweep_idx,
                              int64_t* chosen_keys, double* weights, int32_t* selected, double* log_ml) {
  if (!h) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls != h->obs_cls) throw Unsupported("row moves of latent classes are not built yet");
    if (row < 0 || row >= h->N) throw BadArg("row out of range");
    recount(h);
    refresh_candidate_mats(h);
    int cur = 0;
    CK(cudaMemcpy(&cur, h->d_assign[0]->p + row, sizeof(int), cudaMemcpyDeviceToHost));
    run_row_moves(h, row, row + 1, seed, sweep_idx, cur >= 0);      // a row that is not in the trace yet has no retained particle
    CK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    const int K = h->K;
    for (int b = 0; b < h->n_blocks; ++b) {
      if (h->progs[b].root < 0) { for (int k = 0; k < K; ++k) if (chosen_keys) chosen_keys[(size_t)k * h->n_blocks + b] = -2; continue; }
      const TableH& T = h->tables[h->progs[b].stars[h->progs[b].root].table];
      for (int k = 0; k < K; ++k) {
        int ch = 0;
        CK(cudaMemcpy(&ch, h->d_pchoice[b]->p + (size_t)row * K + k, sizeof(int), cudaMemcpyDeviceToHost));
        if (chosen_keys) chosen_keys[(size_t)k * h->n_blocks + b] = ch >= 0 ? T.keys.at(ch) : -1;
      }
    }
    if (weights) CK(cudaMemcpy(weights, h->d_pweight.p + (size_t)row * K, K * sizeof(double), cudaMemcpyDeviceToHost));
    if (selected) CK(cudaMemcpy(selected, h->d_sel.p + row, sizeof(int), cudaMemcpyDeviceToHost));
    if (log_ml) CK(cudaMemcpy(log_ml, h->d_row_logml.p + row, sizeof(double), cudaMemcpyDeviceToHost));
    int flags = 0;
    CK(cudaMemcpy(&flags, h->d_row_flags.p + row, sizeof(int), cudaMemcpyDeviceToHost));
    // a dummy draw (ROWFLAG_DUMMY) is reported through pclean_download_row_flags: its particle is scored but never selected
    if (flags & ~(ROWFLAG_CHANGED | ROWFLAG_DUMMY)) throw std::runtime_error("row move hit an unsupported path (flags " + std::to_string(flags) + ")");
  });
}

int32_t pclean_download_assignment(pclean_engine* h, int32_t cls, int32_t fk_vertex, int64_t n_rows, int64_t* keys) {
  if (!h || !keys) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls != h->obs_cls || n_rows != h->N) throw BadArg("bad class / row count");
    for (int b = 0; b < h->n_blocks; ++b) {
      if (h->progs[b].root < 0) continue;
      const StarL& root = h->progs[b].stars[h->progs[b].root];
      if (root.vertex != fk_vertex) continue;
      std::vector<int> slots = h->d_assign[b]->download();
      const TableH& T = h->tables[root.table];
      for (int64_t r = 0; r < n_rows; ++r) keys[r] = slots[r] >= 0 ? T.keys.at(slots[r]) : -1;      // -1: row without an assignment yet
      return;
    }
    throw BadArg("vertex is not a top-level reference slot");
  });
}

int32_t pclean_download_cells(pclean_engine* h, int32_t cls, int32_t n_vertices, const int32_t* vertices, int64_t n_rows, pclean_value* out) {
  if (!h || !vertices || !out) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls != h->obs_cls || n_rows != h->N) throw BadArg("bad class / row count");
    if (h->obs_host_stale) {             // the observed cells were replaced on the device (pclean_update_observations)
      CK(cudaStreamSynchronize(h->stream));
      check_device_error(h);
      for (auto& c : h->cols) { if (c->is_real) c->real = c->d_real.download(); else { c->sid = c->d_sid.download(); c->uobs = c->d_uobs.download(); } }
      h->obs_host_stale = false;
    }
    const ClassM& cm = h->m.classes[cls];
    std::vector<std::vector<int>> slots(h->n_blocks);
    for (int b = 0; b < h->n_blocks; ++b) slots[b] = h->d_assign[b]->download();
    std::map<int, std::vector<int>> table_cells;
    auto cells_of = [&](int t) -> const std::vector<int>& {
      auto it = table_cells.find(t);
      if (it == table_cells.end()) it = table_cells.emplace(t, h->tables[t].cells.download()).first;
      return it->second;
    };
    // value of obs-class vertex v for row r as a string id (or -1)
    std::map<int, std::vector<int>> rowcells;
    auto rowcell_of = [&](int v) -> const std::vector<int>* {
      if (v < 0 || v >= (int)h->d_rowcell.size() || !h->d_rowcell[v]->p) return nullptr;
      auto it = rowcells.find(v);
      if (it == rowcells.end()) it = rowcells.emplace(v, h->d_rowcell[v]->download()).first;
      return &it->second;
    };
    std::function<int(int, int64_t)> sid_of = [&](int v, int64_t r) -> int {
      auto cit = h->col_of_vertex.find(v);
      if (cit != h->col_of_vertex.end() && !h->cols[cit->second]->absent[r]) return h->cols[cit->second]->sid[r];
      if (const std::vector<int>* rc = rowcell_of(v)) return (*rc)[r] >= 0 ? (*rc)[r] : -1;
      const Node& n = cm.nodes[v];
      if (n.wrap == PCLEAN_WRAP_SUBMODEL) {
        for (int b = 0; b < h->n_blocks; ++b) {
          if (h->progs[b].root < 0) continue;
          const StarL& root = h->progs[b].stars[h->progs[b].root];
          if (n.wfk[0] != root.vertex) continue;
          const TableH& T = h->tables[root.table];
          if (slots[b][r] < 0) return -1;                 // row without an assignment yet (init_rows / unsupported row): ABSENT
          return cells_of(root.table)[(size_t)n.wsub[0] * T.cap + slots[b][r]];
        }
        return -1;
      }
      if (n.kind == PCLEAN_NODE_JULIA && h->m.funcs[n.func].kind == PCLEAN_FUNC_JOIN) {
        const int a = sid_of(n.args[0], r), b2 = sid_of(n.args[1], r);
        if (a < 0 || b2 < 0) return -1;
        std::u32string s = h->strings[a]; s += h->strings[h->m.funcs[n.func].cst.i]; s += h->strings[b2];
        return h->intern(s);
      }
      return -1;
    };
    for (int vi = 0; vi < n_vertices; ++vi) if (vertices[vi] < 0 || vertices[vi] >= cm.nv) throw BadArg("vertex out of range");
    for (int vi = 0; vi < n_vertices; ++vi) {
      const int v = vertices[vi];
      const Node& vn = cm.nodes[v];
      auto cit = h->col_of_vertex.find(v);
      const bool real_col = cit != h->col_of_vertex.end() && h->cols[cit->second]->is_real;
      const bool round_back = vn.wrap == PCLEAN_WRAP_NONE && vn.kind == PCLEAN_NODE_JULIA && h->m.funcs[vn.func].kind == PCLEAN_FUNC_ROUND_BACKWARD;
      for (int64_t r = 0; r < n_rows; ++r) {
        pclean_value& o = out[(size_t)vi * n_rows + r];
        if (real_col) { o.tag = h->cols[cit->second]->absent[r] ? PCLEAN_VAL_ABSENT : PCLEAN_VAL_REAL; o.i = 0; o.d = h->cols[cit->second]->real[r]; continue; }
        if (round_back) {          // corrected = round(unit.backward(rent))  (experiments/rents/run.jl:25)
          const std::vector<int>* uc = rowcell_of(vn.args.at(0));
          auto xc = h->col_of_vertex.find(vn.args.at(1));
          if (!uc || (*uc)[r] < 0 || xc == h->col_of_vertex.end()) { o.tag = PCLEAN_VAL_ABSENT; o.i = 0; o.d = 0; continue; }
          o.tag = PCLEAN_VAL_REAL; o.i = 0; o.d = std::nearbyint(h->cols[xc->second]->real[r] * h->m.xform_scale.at((*uc)[r]));
          continue;
        }
        const int s = sid_of(v, r);
        o.tag = s >= 0 ? PCLEAN_VAL_STR : PCLEAN_VAL_ABSENT; o.i = s; o.d = 0.0;
      }
    }
  });
}

int32_t pclean_download_logweights(pclean_engine* h, int32_t cls, int64_t n_rows, double* out) {
  if (!h || !out) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls != h->obs_cls || n_rows != h->N) throw BadArg("bad class / row count");
    CK(cudaMemcpy(out, h->d_row_logml.p, n_rows * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

/* rows [row_begin, row_end) only (a row-sharded engine owns just that range): keys of the rows
   referenced through `fk_vertex` / per-row log-weights */
int32_t pclean_download_assignment_range(pclean_engine* h, int32_t cls, int32_t fk_vertex, int64_t row_begin, int64_t row_end, int64_t* keys) {
  if (!h || !keys) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls != h->obs_cls || row_begin < 0 || row_end > h->N || row_begin > row_end) throw BadArg("bad class / row range");
    for (int b = 0; b < h->n_blocks; ++b) {
      if (h->progs[b].root < 0) continue;
      const StarL& root = h->progs[b].stars[h->progs[b].root];
      if (root.vertex != fk_vertex) continue;
      const int64_t n = row_end - row_begin;
      std::vector<int> slots((size_t)n);
      if (n) CK(cudaMemcpy(slots.data(), h->d_assign[b]->p + row_begin, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
      const TableH& T = h->tables[root.table];
      for (int64_t r = 0; r < n; ++r) keys[r] = slots[r] >= 0 ? T.keys.at(slots[r]) : -1;
      return;
    }
    throw BadArg("vertex is not a top-level reference slot");
  });
}
int32_t pclean_download_logweights_range(pclean_engine* h, int32_t cls, int64_t row_begin, int64_t row_end, double* out) {
  if (!h || !out) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls != h->obs_cls || row_begin < 0 || row_end > h->N || row_begin > row_end) throw BadArg("bad class / row range");
    if (row_end > row_begin) CK(cudaMemcpy(out, h->d_row_logml.p + row_begin, (size_t)(row_end - row_begin) * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

/* per-row flags of the last row moves (1 = some particle drew a StringPrior dummy placeholder and was
   excluded from the selection, 2 = missing join matrices, 4 = scratch pool full) */
int32_t pclean_download_row_flags(pclean_engine* h, int32_t cls, int64_t row_begin, int64_t row_end, int32_t* out) {
  if (!h || !out) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls != h->obs_cls || row_begin < 0 || row_end > h->N || row_begin > row_end) throw BadArg("bad class / row range");
    if (row_end > row_begin) CK(cudaMemcpy(out, h->d_row_flags.p + row_begin, (size_t)(row_end - row_begin) * sizeof(int), cudaMemcpyDeviceToHost));
  });
}

int32_t pclean_table_size(pclean_engine* h, int32_t cls, int64_t* n_rows) {
  if (!h || !n_rows || cls < 0 || cls >= (int)h->tables.size()) return PCLEAN_ERR_ARG;
  *n_rows = h->tables[cls].n_slots ? h->tables[cls].n_slots : (int64_t)h->tables[cls].keys.size();
  return PCLEAN_OK;
}

int32_t pclean_download_table(pclean_engine* h, int32_t cls, int64_t cap_rows, int64_t* keys, int32_t* refcounts,
                              pclean_value* cells, int64_t* n_rows) {
  if (!h || !n_rows) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls < 0 || cls >= (int)h->tables.size() || !h->tables[cls].loaded) throw BadArg("class not loaded");
    recount(h);
    CK(cudaStreamSynchronize(h->stream));
    TableH& T = h->tables[cls];
    *n_rows = T.n_slots;
    const int64_t n = std::min<int64_t>(cap_rows, T.n_slots);
    std::vector<int> rc = T.refcnt.download(), cl = T.cells.download();
    for (int64_t r = 0; r < n; ++r) { if (keys) keys[r] = T.keys[r]; if (refcounts) refcounts[r] = rc[r]; }
    if (cells) {
      const ClassM& tm = h->m.classes[cls];
      for (int v = 0; v < T.n_normal; ++v)
        for (int64_t r = 0; r < n; ++r) {
          pclean_value& o = cells[(size_t)v * n + r];
          const int x = cl[(size_t)v * T.cap + r];
          const bool is_fk = tm.nodes[v].kind == PCLEAN_NODE_FK;
          if (x < 0) { o.tag = PCLEAN_VAL_ABSENT; o.i = 0; o.d = 0; }
          else if (is_fk) { o.tag = PCLEAN_VAL_KEY; o.i = 0; o.d = (double)h->tables[tm.nodes[v].target].keys.at(x); }
          else { o.tag = PCLEAN_VAL_STR; o.i = x; o.d = 0; }
        }
    }
  });
}

int32_t pclean_string_count(pclean_engine* h, int32_t* n) {
  if (!h || !n) return PCLEAN_ERR_ARG;
  *n = (int)h->strings.size();
  return PCLEAN_OK;
}

int32_t pclean_get_string(pclean_engine* h, int32_t id, int32_t cap, uint32_t* cp, int32_t* len) {
  if (!h || !len || id < 0 || id >= (int)h->strings.size()) return PCLEAN_ERR_ARG;
  const std::u32string& s = h->strings[id];
  *len = (int)s.size();
  for (int i = 0; i < std::min<int>(cap, (int)s.size()); ++i) cp[i] = s[i];
  return PCLEAN_OK;
}

int32_t pclean_addtypos_pairs(pclean_engine* h, int64_t n, const int32_t* observed_ids, const int32_t* clean_ids,
                              int32_t max_typos, int32_t* distances, double* logdensities) {
  if (!h || !observed_ids || !clean_ids) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    for (int64_t i = 0; i < n; ++i)
      if (observed_ids[i] < 0 || observed_ids[i] >= h->n_dev_strings || clean_ids[i] < 0 || clean_ids[i] >= h->n_dev_strings)
        throw BadArg("string id outside the uploaded dictionary");
    DBuf<int> a, b, d; DBuf<double> l;
    a.upload(std::vector<int>(observed_ids, observed_ids + n)); b.upload(std::vector<int>(clean_ids, clean_ids + n));
    d.alloc(n); l.alloc(n);
    k_pairs<<<nblk(n, 64), 64, 0, h->stream>>>(h->d_sym.p, h->d_str_off.p, h->d_str_len.p, n, a.p, b.p, max_typos, h->d_LG.p, h->d_LOGN.p, d.p, l.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    if (distances) CK(cudaMemcpy(distances, d.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (logdensities) CK(cudaMemcpy(logdensities, l.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

/* read one cell of a distance matrix family: used by the parity tests to check the
   bit-parallel kernel against the oracle's plain DP */
int32_t pclean_debug_distance(pclean_engine* h, int32_t obs_col, int32_t u, int32_t table, int32_t col, int32_t slot, int32_t* out) {
  if (!h || !out) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    for (auto& Mp : h->mats) {
      MatH& M = *Mp;
      if (M.obs_col == obs_col && M.table == table && M.col == col && M.prefix_a < 0) {
        uint8_t v = 0;
        CK(cudaMemcpy(&v, M.d.p + (size_t)u * M.stride + slot, 1, cudaMemcpyDeviceToHost));
        *out = v; return;
      }
    }
    throw BadArg("no such distance matrix");
  });
}

int32_t pclean_attach_nccl(pclean_engine* h, void* nccl_comm, int32_t rank, int32_t world) {
  if (!h || !nccl_comm) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw std::runtime_error(std::string("dlopen libnccl.so.2: ") + dlerror());
    h->nccl.lib = lib;
    h->nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    if (!h->nccl.AllReduce) throw std::runtime_error("ncclAllReduce not found");
    h->nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    if (!h->nccl.AllGather) throw std::runtime_error("ncclAllGather not found");
    h->nccl.comm = nccl_comm; h->nccl.rank = rank; h->nccl.world = world;
  });
}

/* create a communicator from a unique id the host distributed (rank 0 obtains it with
   pclean_nccl_unique_id and broadcasts the 128 bytes, e.g. through torch.distributed) */
int32_t pclean_nccl_unique_id(void* out128) {
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib || !out128) return PCLEAN_ERR_NCCL;
  typedef int (*fn_t)(NcclUniqueId*);
  fn_t f = (fn_t)dlsym(lib, "ncclGetUniqueId");
  if (!f) return PCLEAN_ERR_NCCL;
  return f((NcclUniqueId*)out128) == 0 ? PCLEAN_OK : PCLEAN_ERR_NCCL;
}
int32_t pclean_nccl_init(pclean_engine* h, const void* id128, int32_t rank, int32_t world) {
  if (!h || !id128) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw std::runtime_error(std::string("dlopen libnccl.so.2: ") + dlerror());
    typedef int (*init_t)(void**, int, NcclUniqueId, int);
    init_t init = (init_t)dlsym(lib, "ncclCommInitRank");
    if (!init) throw std::runtime_error("ncclCommInitRank not found");
    NcclUniqueId id; std::memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    if (init(&comm, world, id, rank) != 0) throw std::runtime_error("ncclCommInitRank failed");
    h->nccl.lib = lib; h->nccl.comm = comm; h->nccl.rank = rank; h->nccl.world = world; h->nccl.owned = true;
    h->nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    h->nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    if (!h->nccl.AllReduce || !h->nccl.AllGather) throw std::runtime_error("ncclAllReduce / ncclAllGather not found");
  });
}

int32_t pclean_set_row_shard(pclean_engine* h, int32_t cls, int64_t row_begin, int64_t row_end) {
  if (!h) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (cls != h->obs_cls) throw BadArg("row shards apply to the observation class");
    if (row_begin < 0 || row_end > h->N || row_begin > row_end) throw BadArg("bad shard range");
    h->shard_begin = row_begin; h->shard_end = row_end;
  });
}


/* per-block figures of the last sweep and of the lowered programs (bench.py roofline):
   out[0] = device ms of k_block for `block`, out[1] = algorithmic distance bytes per row
   (sum over enumerated stars of elements x terms x 1 B), out[2] = enumerated elements per row,
   out[3] = likelihood terms evaluated per row, out[4] = candidates of the block's reference
   table (|C_b| - 1 of SURVEY 8d), out[5] = likelihood terms per candidate (F_b) */
int32_t pclean_block_metrics(pclean_engine* h, int32_t block, double* out4) {
  if (!h || !out4) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    finalize(h);
    if (block < 0 || block >= h->n_blocks) throw BadArg("block out of range");
    const BlockProgram& bp = h->progs[block];
    double bytes = 0, elems = 0, terms = 0;
    for (size_t si = 0; si < bp.stars.size(); ++si) {
      const StarL& s = bp.stars[si];
      const StarD& D = h->h_stars[h->h_progs[block].star0 + si];
      if (D.hoist >= 0) continue;
      const double ne = s.kind == ST_FK ? (double)h->tables[s.table].n_slots : (double)D.nopt;
      elems += ne; terms += ne * s.terms.size(); bytes += ne * s.terms.size();
    }
    out4[0] = block < 8 ? h->block_ms[block] : 0.0; out4[1] = bytes; out4[2] = elems; out4[3] = terms;
    out4[4] = 0.0; out4[5] = 0.0;
    if (bp.root >= 0) {
      const StarL& rs = bp.stars[bp.root];
      out4[4] = rs.kind == ST_FK ? (double)h->tables[rs.table].n_slots : (double)h->h_stars[h->h_progs[block].star0 + bp.root].nopt;
      out4[5] = (double)rs.terms.size();
    }
  });
}

/* total bytes of device memory held by distance matrices */
int32_t pclean_matrix_bytes(pclean_engine* h, int64_t* out) {
  if (!h || !out) return PCLEAN_ERR_ARG;
  int64_t b = 0;
  for (auto& M : h->mats) b += (int64_t)std::max(1, M->rows) * M->stride;
  for (auto& LM : h->lmats) b += (int64_t)LM->bytes;
  *out = b;
  return PCLEAN_OK;
}

/* re-send the encoded observation columns host->device from pinned memory (the per-step
   input transfer of the end-to-end measurement); returns bytes copied */
// string id -> unique-value index of the column (what k_block's matrices are indexed by); a value the
// column never held at load time has no distance row: the row is flagged and the call reports it
__global__ void k_sid_to_u(const int* __restrict__ sid, const int* __restrict__ u_of_sid, int n_map, int* __restrict__ uobs, long long n, int* err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = sid[i];
  int u = -1;
  if (s >= 0) { u = s < n_map ? u_of_sid[s] : -1; if (u < 0) atomicExch(err, PCLEAN_ERR_ARG); }
  uobs[i] = u;
}
/* The observed cells of rows [row_begin, row_end) again, from CALLER memory (the buffers a host keeps
   after encoding its table once, julia/PCleanB200.jl encode_observations): per dataset column either
   int32 string ids of the engine's dictionary (-1 = missing) or doubles, in the column order of
   pclean_load_observations; a column's pointer may be null (left as it is).  Pointers address element 0
   of the column (not row_begin).  Copies are enqueued on the engine's stream; ids are mapped to the
   unique-value indices on the device. */
int32_t pclean_update_observations(pclean_engine* h, int32_t n_cols, const int32_t* const* sid_cols, const double* const* real_cols,
                                   int64_t row_begin, int64_t row_end, int64_t* bytes) {
  if (!h || (!sid_cols && !real_cols)) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    if (h->obs_cls < 0) throw std::runtime_error("load the observations first");
    if (!h->d_err.p) { h->d_err.alloc(1); h->d_err.zero(); }          // usable before the trace exists (the columns live on the device since pclean_load_observations)
    if (n_cols != (int)h->cols.size()) throw BadArg("column count differs from the loaded dataset");
    if (row_begin < 0 || row_end > h->N || row_begin > row_end) throw BadArg("row range out of bounds");
    const int64_t n = row_end - row_begin;
    int64_t total = 0;
    for (int c = 0; c < n_cols && n > 0; ++c) {
      ObsCol& oc = *h->cols[c];
      if (oc.is_real) {
        if (!real_cols || !real_cols[c]) continue;
        CK(cudaMemcpyAsync(oc.d_real.p + row_begin, real_cols[c] + row_begin, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        total += n * (int64_t)sizeof(double);
        continue;
      }
      if (!sid_cols || !sid_cols[c]) continue;
      if (oc.d_u_of_sid.n == 0) {
        std::vector<int> map((size_t)std::max<int>(1, (int)h->strings.size()), -1);
        for (size_t u = 0; u < oc.ulist.size(); ++u) map[oc.ulist[u]] = (int)u;
        oc.d_u_of_sid.upload(map);
      }
      CK(cudaMemcpyAsync(oc.d_sid.p + row_begin, sid_cols[c] + row_begin, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      k_sid_to_u<<<nblk(n, 256), 256, 0, h->stream>>>(oc.d_sid.p + row_begin, oc.d_u_of_sid.p, (int)oc.d_u_of_sid.n, oc.d_uobs.p + row_begin, n, h->d_err.p);
      ++h->launches;
      total += n * (int64_t)sizeof(int);
    }
    CK(cudaGetLastError());
    if (total) h->obs_host_stale = true;
    if (bytes) *bytes = total;
  });
}

int32_t pclean_resync_observations(pclean_engine* h, int64_t* bytes) {
  if (!h) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (!h->pinned) {
      h->pinned = new pinned_cols_t();
      h->pinned->bytes_per_col = (size_t)h->N * sizeof(int);
      for (auto& c : h->cols) {
        int* p = nullptr;
        CK(cudaMallocHost(&p, std::max<size_t>(4, h->pinned->bytes_per_col)));
        std::memcpy(p, c->uobs.data(), h->pinned->bytes_per_col);
        h->pinned->host.push_back(p);
      }
    }
    // a row-sharded engine scores only its own rows: only their cells travel
    const int64_t r0 = h->shard_begin, r1 = h->shard_end < 0 ? h->N : h->shard_end;
    int64_t total = 0;
    for (size_t c = 0; c < h->cols.size(); ++c) {
      CK(cudaMemcpyAsync(h->cols[c]->d_uobs.p + r0, h->pinned->host[c] + r0, (size_t)(r1 - r0) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      total += (int64_t)(r1 - r0) * (int64_t)sizeof(int);
    }
    if (bytes) *bytes = total;
  });
}


/* engine options: "prune" (1 default: integer-bound pruning of far candidates; 0: exact
   enumeration of every candidate) */
int32_t pclean_set_option(pclean_engine* h, const char* name, int32_t value) {
  if (!h || !name) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (std::string(name) == "exchange_path") { h->exchange_path = value ? 1 : 0; }
    else if (std::string(name) == "resample_params") { h->resample_params = value ? 1 : 0; }
    else if (std::string(name) == "batch_rows") { h->batch_rows = value > 0 ? value : 0; }
    else if (std::string(name) == "compact_now") { h->compact_now = value != 0; }
    else if (std::string(name) == "compact_headroom") { h->compact_head = std::max(0, value); }
    else if (std::string(name) == "init_rows") { h->init_rows = value > 0 ? value : 0; }
    else if (std::string(name) == "init_divisor") { if (value < 1) throw BadArg("init_divisor must be >= 1"); h->init_divisor = value; }
    else if (std::string(name) == "table_cap") { if (value < 16) throw BadArg("table_cap too small"); h->table_cap = value; }
    else if (std::string(name) == "memo") {
      if (h->finalized) { h->h_dev.memo_mask = value && h->d_memo_keys[0].p ? (1u << h->memo_log2) - 1u : 0; h->pmemo_dirty = true; CK(cudaSetDevice(h->device)); upload_dev(h); }
      else if (!value) h->memo_log2 = 0;
    } else if (std::string(name) == "param_seed") {
      // seed of the keyed prior draws that initialise parameters nobody set (include/pclean_rng.h, PCLEAN_RNG_PARAM_INIT):
      // with the oracle's seed both sides start from the same parameter values (tests on traces that carry none)
      h->param_seed = (uint64_t)(uint32_t)value; h->finalized = false;
    } else if (std::string(name) == "kb_variant") {
      if (value < 0 || value > 2) throw BadArg("kb_variant must be 0, 1 or 2");
      h->kb_variant = value; CK(cudaSetDevice(h->device)); prepare_k_block(h);
    } else if (std::string(name) == "opts") {          // PCL_OPT_* bit mask (A/B measurements; results do not depend on it beyond rounding)
      h->opts = value; h->pmemo_dirty = true;
      if (h->finalized) { h->h_dev.opts = value; CK(cudaSetDevice(h->device)); upload_dev(h); }
    } else if (std::string(name) == "prune") { h->prune = value ? 1 : 0; if (h->finalized) { h->h_dev.prune = h->prune; CK(cudaSetDevice(h->device)); upload_dev(h); } }
    else throw BadArg("unknown option");
  });
}


/* debug: per-particle choice of `block` for `row` after the last row-move kernels, and the
   scratch record (obs-class vertex numbering) of particles that proposed a new row */
int32_t pclean_debug_particles(pclean_engine* h, int64_t row, int32_t block, int32_t* choices, int32_t* scratch /* [K][nvC] */) {
  if (!h || !choices) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    for (int k = 0; k < h->K; ++k) {
      CK(cudaMemcpy(&choices[k], h->d_pchoice[block]->p + (size_t)row * h->K + k, sizeof(int), cudaMemcpyDeviceToHost));
      if (scratch) {
        for (int v = 0; v < h->nvC; ++v) scratch[(size_t)k * h->nvC + v] = PCL_UNSET;
        if (choices[k] <= -2 && choices[k] != PCL_CHOICE_UNSET)
          CK(cudaMemcpy(scratch + (size_t)k * h->nvC, h->d_pool.p + (size_t)(-(choices[k]) - 2) * h->nvC, h->nvC * sizeof(int), cudaMemcpyDeviceToHost));
      }
    }
  });
}


/* run_smc! of one latent row as a pure function of the snapshot: the row the selected particle
   would install (cells of the class's vertices: STR id, KEY of a referenced row or KEY -1 for a
   proposed new row), the selected particle and the return value.  Parity tests. */
int32_t pclean_latent_move_debug(pclean_engine* h, int32_t cls, int64_t key, uint64_t seed, uint32_t sweep_idx,
                                 pclean_value* out_cells, int32_t* selected, double* log_ml) {
  if (!h || !out_cells) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    finalize(h);
    if (cls < 0 || cls >= (int)h->tables.size() || cls == h->obs_cls || !h->tables[cls].loaded) throw BadArg("not a loaded latent class");
    TableH& T = h->tables[cls];
    auto sk = T.slot_of_key.find(key);
    if (sk == T.slot_of_key.end()) throw BadArg("no such row key");
    const int slot = sk->second;
    run_latent_moves(h, cls, slot, 1, seed, sweep_idx);
    CK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    const int mask = h->lpat_active ? h->h_lpat.at(slot) : 0;
    const int pid = h->lpat_active ? latent_prog_pat(h, cls, mask) : latent_prog(h, cls);
    int sel = 0, flags = 0; double ml = 0;
    CK(cudaMemcpy(&sel, h->d_lsel.p + slot, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&flags, h->d_lflags.p + slot, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&ml, h->d_llogml.p + slot, sizeof(double), cudaMemcpyDeviceToHost));
    if (flags) throw std::runtime_error("latent move hit an unsupported path (flags " + std::to_string(flags) + ")");
    if (selected) *selected = sel;
    if (log_ml) *log_ml = ml;
    const ClassM& tm = h->m.classes[cls];
    std::vector<int> cells = T.cells.download();
    std::vector<int> row(T.n_normal);
    for (int v = 0; v < T.n_normal; ++v) row[v] = cells[(size_t)v * T.cap + slot];
    std::vector<char> is_new(T.n_normal, 0);
    const BlockProgram* bp = nullptr;
    if (h->lpat_active) bp = latent_bp(h, cls, mask);
    else for (size_t li = 0; li < h->lprogs.size(); ++li) if (h->lprog_cls[li] == cls) { bp = &h->lprogs[li]; break; }
    if (sel != 0) {
      std::vector<int> optsid = h->h_optsid;
      for (int site = 0; site < (int)bp->roots.size(); ++site) {
        const StarL& s = bp->stars[bp->roots[site]];
        int e = 0;
        CK(cudaMemcpy(&e, h->d_lchoice.p + (size_t)site * T.cap + slot, sizeof(int), cudaMemcpyDeviceToHost));
        const StarD& D = h->h_stars[h->h_progs[pid].star0 + bp->roots[site]];
        if (s.kind == ST_CHOICE) row[s.vertex] = D.list_func >= 0 ? e : optsid[D.opt_off + e];
        else if (e >= 0) {
          std::vector<int> tc = h->tables[s.table].cells.download();
          row[s.vertex] = e;
          for (auto& pr : s.copies) row[pr.first] = tc[(size_t)pr.second * h->tables[s.table].cap + e];
        } else {
          std::vector<int> sc(h->nvC);
          CK(cudaMemcpy(sc.data(), h->d_pool.p + (size_t)(-(e) - 2) * h->nvC, h->nvC * sizeof(int), cudaMemcpyDeviceToHost));
          row[s.vertex] = -1; is_new[s.vertex] = 1;
          for (auto& pr : s.copies) { row[pr.first] = sc[pr.first]; if (sc[pr.first] == -1) is_new[pr.first] = 1; }
        }
      }
    }
    for (int v = 0; v < T.n_normal; ++v) {
      pclean_value& o = out_cells[v];
      const Node& nd = tm.nodes[v];
      const bool is_fk = nd.kind == PCLEAN_NODE_FK;
      if (is_fk && (row[v] >= 0 || is_new[v])) { o.tag = PCLEAN_VAL_KEY; o.i = 0; o.d = is_new[v] ? -1.0 : (double)h->tables[nd.target].keys.at(row[v]); }
      else if (row[v] >= 0 && !is_fk) { o.tag = PCLEAN_VAL_STR; o.i = row[v]; o.d = 0; }
      else { o.tag = PCLEAN_VAL_ABSENT; o.i = 0; o.d = 0; }
    }
  });
}


/* debug: 32 counters of which path the pruned latent evaluation took ([0] = pruned, [i] = i-th bail-out); reset on read */
int32_t pclean_debug_counters(pclean_engine* h, int32_t* out32) {
  if (!h || !out32) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    CK(cudaSetDevice(h->device));
    if (!h->d_dbg.p) throw std::runtime_error("not finalized");
    CK(cudaMemcpy(out32, h->d_dbg.p, 32 * sizeof(int), cudaMemcpyDeviceToHost));
    h->d_dbg.zero();
  });
}

/* Pitman-Yor hyper-parameters of a class table (trace.jl:1-5) */
int32_t pclean_get_py_params(pclean_engine* h, int32_t cls, double* strength, double* discount) {
  if (!h || !strength || !discount || cls < 0 || cls >= (int)h->tables.size()) return PCLEAN_ERR_ARG;
  *strength = h->tables[cls].strength; *discount = h->tables[cls].discount;
  return PCLEAN_OK;
}


/* local discrete cells of the observation rows that are not reference slots (rents: br, unit) —
   part of TableTrace.rows of the observed class (trace.jl:30); values: STR id / XFORM id */
int32_t pclean_load_row_cells(pclean_engine* h, int32_t cls, int32_t vertex, int64_t n_rows, const pclean_value* values) {
  if (!h || !values) return PCLEAN_ERR_ARG;
  return guard(h, [&] {
    if (cls != h->obs_cls || n_rows != h->N) throw BadArg("row cells must cover the observed dataset");
    std::vector<int> v(n_rows, PCL_UNSET);
    for (int64_t r = 0; r < n_rows; ++r) if (values[r].tag == PCLEAN_VAL_STR || values[r].tag == PCLEAN_VAL_XFORM || values[r].tag == PCLEAN_VAL_INT) v[r] = values[r].i;
    h->rowcell_init[vertex] = v;
    h->finalized = false;
  });
}

}  // extern "C"
