/*
 * pclean_b200.h — C ABI of the B200-native PClean sweep engine.
 *
 * The reference (probcomp/PClean) has no FFI: its hot path is ordinary Julia
 * (`initialize_trace` / `run_inference!`, src/inference/inference.jl:3,83).  This header is
 * the boundary a Julia shim binds with `ccall` (see INTEGRATION.md and julia/PCleanB200.jl):
 * plain pointers and sizes, int32 status returns, no exceptions across the boundary, caller
 * owns every host buffer (the engine copies during the call and retains nothing).
 *
 * Conventions: all indices 0-based on the C side (Julia vertex ids are 1-based; the shim
 * subtracts 1).  One handle = one host thread at a time.  The engine owns its CUDA stream(s).
 */
#ifndef PCLEAN_B200_H
#define PCLEAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- values ------------- */
/* A cell of a row trace (reference RowTrace = Dict{Int,Any}, src/model/trace.jl:16). */
enum {
  PCLEAN_VAL_ABSENT  = 0,  /* no entry in the row trace                                  */
  PCLEAN_VAL_MISSING = 1,  /* Julia `missing` held as an explicit observation             */
  PCLEAN_VAL_STR     = 2,  /* i = string id in the dictionary                             */
  PCLEAN_VAL_REAL    = 3,  /* d                                                           */
  PCLEAN_VAL_INT     = 4,  /* i                                                           */
  PCLEAN_VAL_LIST    = 5,  /* i = list id (option lists)                                  */
  PCLEAN_VAL_XFORM   = 6,  /* i = transformation id (transformed_gaussian.jl:5-9)         */
  PCLEAN_VAL_PARAM   = 7,  /* i = parameter slot id (a BasicParameter)                    */
  PCLEAN_VAL_IPARAM  = 8,  /* i = parameter spec id of an IndexedParameter                */
  PCLEAN_VAL_KEY     = 9,  /* i = row key in the target table of a reference slot         */
  PCLEAN_VAL_DUMMY   = 10  /* ProposalDummyValue (distributions.jl:7-8)                   */
};
typedef struct pclean_value { int32_t tag; int32_t i; double d; } pclean_value;

/* ---------------------------------------------------------------- node kinds --------- */
enum { PCLEAN_NODE_JULIA = 0, PCLEAN_NODE_CHOICE = 1, PCLEAN_NODE_PARAM = 2, PCLEAN_NODE_FK = 3 };
enum { PCLEAN_WRAP_NONE = 0, PCLEAN_WRAP_SUBMODEL = 1, PCLEAN_WRAP_EXTERNAL = 2 };
/* distributions (reference src/distributions/, one file each) */
enum {
  PCLEAN_DIST_ADD_TYPOS = 0, PCLEAN_DIST_CHOOSE_PROPORTIONALLY = 1, PCLEAN_DIST_CHOOSE_UNIFORMLY = 2,
  PCLEAN_DIST_STRING_PRIOR = 3, PCLEAN_DIST_TIME_PRIOR = 4, PCLEAN_DIST_MAYBE_SWAP = 5,
  PCLEAN_DIST_TRANSFORMED_GAUSSIAN = 6, PCLEAN_DIST_UNMODELED = 7, PCLEAN_DIST_ADD_NOISE = 8
};
enum { PCLEAN_PARAM_PROPORTIONS = 0, PCLEAN_PARAM_MEAN = 1, PCLEAN_PARAM_PROB = 2 };
/* JuliaNode lowering: constants, tabulated closures, and natively understood builtins. */
enum {
  PCLEAN_FUNC_CONST = 0,          /* zero-argument node: func_const                         */
  PCLEAN_FUNC_TABLE = 1,          /* closure tabulated over the supports of its key args    */
  PCLEAN_FUNC_ROUND_BACKWARD = 2, /* round(unit.backward(x)); args (XFORM, REAL)            */
  PCLEAN_FUNC_JOIN = 3            /* "$(a)<sep>$(b)"; args (STR, STR); func_const = sep STR */
};

/*
 * Flat model IR — the App. D flattening of PCleanModel (src/model/model.jl:87-188).
 * Vertices of all classes are concatenated: class c owns global vertices
 * [class_voff[c], class_voff[c+1]).  Vertex ids stored *inside* arrays are class-local.
 * CSR arrays have n+1 offsets.
 */
typedef struct pclean_model_ir {
  int32_t n_classes;                 /* in model.class_order                                */
  const int32_t* class_voff;         /* [n_classes+1]                                       */
  const double*  py_strength;        /* [n_classes]  initial PitmanYorParams (builder.jl:39)*/
  const double*  py_discount;
  int32_t n_vertices;                /* = class_voff[n_classes]                             */
  const int32_t* v_kind;             /* base node kind after stripping wrappers             */
  const int32_t* v_wrap;             /* PCLEAN_WRAP_*                                       */
  const int32_t* v_wrap_off;         /* CSR → wrap_fk/wrap_subid: SubmodelNode chain,       */
  const int32_t* wrap_fk;            /*   outermost first: foreign_key_node_id (local)      */
  const int32_t* wrap_subid;         /*   subnode_id (vertex in the target class)           */
  const int32_t* v_dist;             /* CHOICE: PCLEAN_DIST_*                               */
  const int32_t* v_args_off;         /* CSR → v_args (local ids; EXTERNAL: referring-class) */
  const int32_t* v_args;
  const int32_t* v_func;             /* JULIA: function id                                  */
  const int32_t* v_target;           /* FK: target class                                    */
  const int32_t* v_vmap_off;         /* FK: CSR → v_vmap[target vertex] = local vertex      */
  const int32_t* v_vmap;
  const int32_t* v_param;            /* PARAM: parameter spec id                            */
  const int32_t* v_path;             /* EXTERNAL: path id (global)                          */
  const int32_t* v_extv;             /* EXTERNAL: external_node_id (vertex in path source)  */
  /* blocks and plans */
  const int32_t* class_block_off;    /* [n_classes+1] → global block ids                    */
  int32_t n_blocks;
  const int32_t* block_voff;         /* [n_blocks+1] → block_v (ordered vertex lists)       */
  const int32_t* block_v;
  const int32_t* plan_off;           /* [n_blocks+1] → preorder plan arrays                 */
  const int32_t* plan_vertex;
  const int32_t* plan_nchild;
  /* @guaranteed hash keys */
  const int32_t* class_hash_off;     /* [n_classes+1] → hash_v                              */
  const int32_t* hash_v;
  /* incoming references (model.jl:122-126): one path = chain of (class, FK vertex),
     path[last] = ultimately referring class; path_target = class the path ends in */
  int32_t n_paths;
  const int32_t* path_target;        /* [n_paths]                                           */
  const int32_t* path_len_off;       /* [n_paths+1] → path_class/path_vertex                */
  const int32_t* path_class;
  const int32_t* path_vertex;
  const int32_t* path_vmap_off;      /* [n_paths+1] → path_vmap[target vertex] = vertex in  */
  const int32_t* path_vmap;          /*   the ultimately referring class (or -1)            */
  /* functions */
  int32_t n_funcs;
  const int32_t* func_kind;
  const pclean_value* func_const;
  const int32_t* func_keyarg_off;    /* [n_funcs+1] → positions of key args                 */
  const int32_t* func_keyargs;
  const int32_t* func_tab_off;       /* [n_funcs+1] → entries                               */
  const int32_t* tab_keys;           /* entry e: keys at tab_keys[tab_key_off[e] ..]        */
  const int64_t* tab_key_off;        /* [n_entries+1]                                       */
  const pclean_value* tab_vals;      /* [n_entries]                                         */
  /* parameters */
  int32_t n_params;                  /* parameter specs (one per ParameterNode)             */
  const int32_t* param_kind;
  const int32_t* param_indexed;
  const double*  param_prior0;       /* concentration | mean | a                            */
  const double*  param_prior1;       /*               | std  | b                            */
  int32_t n_param_slots;             /* BasicParameter instances                            */
  const int32_t* slot_param;         /* [n_param_slots] → spec                              */
  /* option lists */
  int32_t n_lists;
  const int64_t* list_off;           /* [n_lists+1] → list_vals                             */
  const pclean_value* list_vals;
  /* transformations: backward(x) = x*scale */
  int32_t n_xforms;
  const double* xform_scale;
  /* string dictionary (UTF-32 codepoints) */
  int32_t n_strings;
  const int64_t* str_off;            /* [n_strings+1]                                       */
  const uint32_t* str_cp;
  /* StringPrior language model (string_prior.jl:8-9): 28 unigram + 28x28 bigram,
     bigram[next*28 + prev] (column = previous letter)                                     */
  const double* lm_unigram;
  const double* lm_bigram;
} pclean_model_ir;

/* Observed dataset bound to an observation class (`ObservedDataset`, query.jl:40-43) —
   column-major cells, `vertex_of_col` from Query.obsmap.  A cell with tag ABSENT is not an
   observation (inference.jl:23-33 decides ABSENT vs explicit MISSING on the host side). */
typedef struct pclean_observations {
  int32_t cls;
  int64_t n_rows;
  int32_t n_cols;
  const int32_t* vertex_of_col;      /* [n_cols] class-local                                */
  const pclean_value* cells;         /* [n_cols][n_rows] column-major                       */
} pclean_observations;

/* InferenceConfig (src/inference/infer_config.jl:1-16) */
typedef struct pclean_config {
  int32_t num_iters;
  int32_t num_particles;
  int32_t use_dd_proposals;
  int32_t use_lo_sweeps;
  int32_t use_mh_instead_of_pg;
  int32_t rejuv_frequency;
  int32_t reporting_frequency;
} pclean_config;

/* Table snapshot exchanged with the host (TableTrace, trace.jl:24-44).  Rows are
   denormalised exactly like the reference's: one cell per non-external vertex. */
typedef struct pclean_table_snapshot {
  int32_t cls;
  int64_t n_rows;
  int32_t n_cols;                    /* = number of vertices of the class                   */
  const int64_t* keys;               /* [n_rows] row keys                                   */
  const pclean_value* cells;         /* [n_cols][n_rows] column-major                       */
  double py_strength, py_discount;
} pclean_table_snapshot;

typedef struct pclean_sweep_stats {
  int64_t rows;                      /* observation rows moved                              */
  int64_t particles;                 /* rows * K                                            */
  int64_t new_rows;                  /* latent rows created by this sweep                   */
  int64_t dummy_draws;               /* ProposalDummyValue selections                       */
  int64_t changed_rows;              /* rows whose selected particle != retained            */
  double  sum_log_ml;                /* Σ run_smc! return values (row_inference.jl:186)     */
  float   kernel_ms;                 /* device time of the row-move kernels                 */
  float   total_ms;                  /* device time of the whole sweep                      */
  int32_t launches;                  /* kernel launches issued by this sweep                */
} pclean_sweep_stats;

typedef struct pclean_engine pclean_engine;

/* status codes */
enum {
  PCLEAN_OK = 0, PCLEAN_ERR_ARG = -1, PCLEAN_ERR_CUDA = -2, PCLEAN_ERR_UNSUPPORTED = -3,
  PCLEAN_ERR_STATE = -4, PCLEAN_ERR_CAPACITY = -5, PCLEAN_ERR_LOOKUP = -6, PCLEAN_ERR_NCCL = -7
};

/* lifecycle — replaces nothing in the reference (it has no engine object); the handle plays
   the role of PCleanTrace (trace.jl:47-50) */
int32_t pclean_create(const pclean_config* cfg, int32_t device, pclean_engine** out);
int32_t pclean_destroy(pclean_engine* h);
const char* pclean_last_error(const pclean_engine* h);
const char* pclean_version(void);

/* model + data upload — replaces the in-process sharing of PCleanModel / ObservedDataset
   (inference.jl:3-5, query.jl:40-43) */
int32_t pclean_load_model(pclean_engine* h, const pclean_model_ir* ir);
int32_t pclean_load_observations(pclean_engine* h, const pclean_observations* obs);
/* the same two calls fed from a "PCLIRv1" file (named arrays, one per field of pclean_model_ir /
   pclean_observations; written by pclean_b200/irfile.py and julia/PCleanB200.jl `write_ir`): a Julia
   host and a C / Python harness can share inputs without sharing a process */
int32_t pclean_load_model_file(pclean_engine* h, const char* path);
int32_t pclean_load_observations_file(pclean_engine* h, const char* path);

/* trace state.  `pclean_load_table` installs the rows of one latent class
   (TableTrace.rows, trace.jl:30); `pclean_load_assignment` gives, for every observation row,
   the key it references through each top-level reference slot (n_fk int64 per row,
   column-major [n_fk][n_rows]) — reference counts, hash buckets and sufficient statistics
   are rebuilt on the device (dependency_tracking.jl:71-99,205-236). */
int32_t pclean_load_table(pclean_engine* h, const pclean_table_snapshot* t);
int32_t pclean_load_assignment(pclean_engine* h, int32_t cls, int64_t n_rows, int32_t n_fk,
                               const int32_t* fk_vertices, const int64_t* keys);
int32_t pclean_set_param_values(pclean_engine* h, int32_t slot, int32_t n, const double* values);
int32_t pclean_get_param_values(pclean_engine* h, int32_t slot, int32_t cap, double* values, int32_t* n);

/* the hot path */
/* initialize_trace (inference.jl:3-58): batched SMC initialisation on the device into empty
   tables (rows visited in a scattered order; option "batch_rows" = 1 gives the reference's
   sequential order exactly).  Parameters start from keyed prior draws under `seed`. */
int32_t pclean_init_trace(pclean_engine* h, uint64_t seed);
/* rows to reserve for one latent table before pclean_init_trace (default: option "table_cap");
   candidate distance matrices take (unique observed strings x reserved rows) bytes per term */
int32_t pclean_reserve_table(pclean_engine* h, int32_t cls, int32_t rows);
/* pgibbs_sweep! restricted to one class (inference.jl:60-81); cls = -1: every supported class */
int32_t pclean_sweep(pclean_engine* h, int32_t cls, uint64_t seed, uint32_t sweep_idx,
                     pclean_sweep_stats* out);
/* run_inference! (inference.jl:83-88): num_iters sweeps */
int32_t pclean_run_inference(pclean_engine* h, uint64_t seed, pclean_sweep_stats* out_total);
/* run_smc! for one row as a pure function of the current snapshot
   (row_inference.jl:108-187): used by the parity tests; does not mutate tables. Outputs:
   per particle per block chosen key (-1 = new row), per particle weight, selected particle,
   return value log_ml + log_total_weight - log K. */
int32_t pclean_row_move_debug(pclean_engine* h, int32_t cls, int64_t row, uint64_t seed,
                              uint32_t sweep_idx, int64_t* chosen_keys /* [K][n_blocks] */,
                              double* weights /* [K] */, int32_t* selected, double* log_ml);

/* results — replaces reading trace.tables[c].rows[key][vertex] (analysis.jl:44-56) */
int32_t pclean_download_cells(pclean_engine* h, int32_t cls, int32_t n_vertices,
                              const int32_t* vertices, int64_t n_rows, pclean_value* out_colmajor);
int32_t pclean_download_assignment(pclean_engine* h, int32_t cls, int32_t fk_vertex,
                                   int64_t n_rows, int64_t* keys);
int32_t pclean_download_logweights(pclean_engine* h, int32_t cls, int64_t n_rows, double* out);
/* the same for rows [row_begin, row_end) only — what a row-sharded engine (pclean_set_row_shard) owns */
int32_t pclean_download_assignment_range(pclean_engine* h, int32_t cls, int32_t fk_vertex,
                                         int64_t row_begin, int64_t row_end, int64_t* keys);
int32_t pclean_download_logweights_range(pclean_engine* h, int32_t cls, int64_t row_begin, int64_t row_end, double* out);
/* per-row flags of the last row moves: 1 = some particle of the row drew a StringPrior dummy placeholder
   (scored like the reference scores it, but excluded from the selection: DESIGN.md section 1), 2 = missing
   join matrices, 4 = scratch pool full */
int32_t pclean_download_row_flags(pclean_engine* h, int32_t cls, int64_t row_begin, int64_t row_end, int32_t* out);
int32_t pclean_table_size(pclean_engine* h, int32_t cls, int64_t* n_rows);
int32_t pclean_download_table(pclean_engine* h, int32_t cls, int64_t cap_rows, int64_t* keys,
                              int32_t* refcounts, pclean_value* cells_colmajor, int64_t* n_rows);
/* strings created on the device (StringPrior.random draws) can be read back */
int32_t pclean_string_count(pclean_engine* h, int32_t* n);
int32_t pclean_get_string(pclean_engine* h, int32_t id, int32_t cap, uint32_t* cp, int32_t* len);

/* the AddTypos kernel on its own (add_typos.jl:50-66): OSA edit distance + log-density for
   n string pairs given as dictionary ids; max_typos < 0 = none */
int32_t pclean_addtypos_pairs(pclean_engine* h, int64_t n, const int32_t* observed_ids,
                              const int32_t* clean_ids, int32_t max_typos,
                              int32_t* distances, double* logdensities);

/* multi-GPU (one process per GPU): attach an NCCL communicator created by the host
   (torch.distributed in the Python harness; NCCL.jl on the Julia side).  comm is an
   `ncclComm_t`; after attachment pclean_sweep all-reduces the sufficient statistics. */
int32_t pclean_attach_nccl(pclean_engine* h, void* nccl_comm, int32_t rank, int32_t world);
int32_t pclean_set_row_shard(pclean_engine* h, int32_t cls, int64_t row_begin, int64_t row_end);
/* the engine can also create the communicator itself (libnccl.so.2 is dlopen'ed): rank 0 asks
   for the 128-byte unique id, the host broadcasts it, every rank calls pclean_nccl_init */
int32_t pclean_nccl_unique_id(void* out128);
int32_t pclean_nccl_init(pclean_engine* h, const void* id128, int32_t rank, int32_t world);

/* more trace state */
/* local discrete cells of the observation rows that are not reference slots (rents: br, unit):
   the rest of TableTrace.rows of the observed class (trace.jl:30); values STR id / XFORM id */
int32_t pclean_load_row_cells(pclean_engine* h, int32_t cls, int32_t vertex, int64_t n_rows,
                              const pclean_value* values);
/* Pitman-Yor hyper-parameters of a class table (trace.jl:1-5, resample_py_params! :83-107) */
int32_t pclean_get_py_params(pclean_engine* h, int32_t cls, double* strength, double* discount);
/* re-send the encoded observation columns host->device from pinned memory (what a host that
   keeps the DataFrame does before a sweep); returns the bytes copied */
int32_t pclean_resync_observations(pclean_engine* h, int64_t* bytes);
/* the observed cells of rows [row_begin, row_end) again, from CALLER memory — the encoded columns a
   host keeps after `encode_observations` (julia/PCleanB200.jl; the reference keeps the rows in
   TableTrace.observations, trace.jl:31): per dataset column (order of pclean_load_observations) int32
   string ids of the engine dictionary (-1 = missing) for string columns, doubles for numeric ones; a null
   column pointer leaves that column as it is; pointers address row 0.  Host->device copies and the
   id -> unique-value mapping run on the engine's stream; a value the column did not hold at load time
   has no distance row and fails the next sweep with PCLEAN_ERR_ARG.  `bytes` = bytes copied. */
int32_t pclean_update_observations(pclean_engine* h, int32_t n_cols, const int32_t* const* sid_cols,
                                   const double* const* real_cols, int64_t row_begin, int64_t row_end, int64_t* bytes);

/* engine options (name, value):
     "prune"            1 (default) integer-bound pruning of candidates that cannot matter at
                        fp64 resolution; 0 = score every candidate exactly
     "memo"             1 (default) share identical block marginals between rows; 0 = off
     "exchange_path"    1 = route new rows through the pack / all-gather / replay path even on
                        one GPU (what a row-sharded run does)
     "resample_params"  1 (default) resample @learned parameters and Pitman-Yor
                        hyper-parameters at the start of each class sweep (inference.jl:72-77)
     "batch_rows"       n > 0: observation rows are moved in consecutive batches of n
                        (1 = the reference's sequential Gibbs order); 0 = the whole shard
     "init_divisor"     pclean_init_trace moves done / init_divisor rows per batch (default 8)
     "init_rows"        pclean_init_trace stops after this many rows (tests)
     "table_cap"        rows reserved per latent table by pclean_init_trace (default 65536)
     "opts"             mask of k_block evaluation strategies, all on by default (31): 1 progressive
                        pruning, 2 persistent memo of choice marginals, 4 lane-parallel exclusions,
                        8 parallel hint scoring, 16 lazy new-row branch (A/B measurements; results
                        do not depend on it)
     "kb_variant"       k_block geometry: 0 = 16 warps x 2 CTAs per SM (default), 1 = 12 x 2, 2 = 16 x 1
     "param_seed"       seed of the keyed prior draws that initialise parameters nobody set
     "compact_now"      1 = pack the dead slots of every latent table before the next class sweep
     "compact_headroom" least free slots a table keeps before it is packed (default 64; the trigger
                        also keeps an eighth of the capacity and twice the largest batch of new rows) */
int32_t pclean_set_option(pclean_engine* h, const char* name, int32_t value);

/* measurement: per-block figures of the last sweep and of the lowered programs:
   out6[0] = device ms of the block kernel, out6[1] = algorithmic distance bytes per row,
   out6[2] = enumerated elements per row, out6[3] = likelihood terms per row,
   out6[4] = candidates of the block's reference table, out6[5] = likelihood terms per candidate
   (|C_b| - 1 and F_b of SURVEY.md section 8d) */
int32_t pclean_block_metrics(pclean_engine* h, int32_t block, double* out6);
int32_t pclean_matrix_bytes(pclean_engine* h, int64_t* out);

/* parity-test entry points (pure functions of the current snapshot) */
/* run_smc! of one latent row (row_inference.jl:108-187 with ExternalLikelihood terms,
   proposal_compiler.jl:306-350): cells the selected particle would install */
int32_t pclean_latent_move_debug(pclean_engine* h, int32_t cls, int64_t key, uint64_t seed,
                                 uint32_t sweep_idx, pclean_value* out_cells, int32_t* selected,
                                 double* log_ml);
/* one cell of a device distance matrix (the memo of add_typos.jl:47,52-58) */
int32_t pclean_debug_distance(pclean_engine* h, int32_t obs_col, int32_t u, int32_t table,
                              int32_t col, int32_t slot, int32_t* out);
/* per-particle choices of `block` for `row` after the last row-move kernels */
int32_t pclean_debug_particles(pclean_engine* h, int64_t row, int32_t block, int32_t* choices,
                               int32_t* scratch /* [K][nvC] */);

#ifdef __cplusplus
}
#endif
#endif /* PCLEAN_B200_H */
