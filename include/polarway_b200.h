/*
 * polarway_b200.h — the drop-in C ABI of libpolarway_b200.so
 *
 * One hot path of Polarway (Polars 0.52 fork) re-implemented as sm_100a CUDA:
 *     LazyFrame.filter(p).group_by(k).agg(a)   and   LazyFrame.group_by_dynamic(...).agg(a)
 *
 * The reference has no per-operator C ABI; what it has is
 *   (1) the Arrow C Data Interface it uses for every buffer hand-off
 *       (crates/polars-arrow/src/ffi/generated.rs:4-34, ffi/array.rs, ffi/schema.rs), and
 *   (2) the expression-plugin ABI (crates/polars-ffi/src/version_0.rs:7-16 `SeriesExport`,
 *       crates/polars-plan/src/plans/aexpr/function_expr/plugin.rs:16-142 loader).
 * This header therefore declares
 *   - an operator-level entry whose arguments are exactly the fields of the executors it replaces
 *     (GroupByExec: polars-mem-engine/src/executors/group_by.rs:23-31, FilterExec: executors/
 *     filter.rs:5-121, GroupByDynamicExec + DynamicGroupOptions: executors/group_by_dynamic.rs:4-17,
 *     polars-time/src/group_by/dynamic.rs:19-39) with columns passed as ArrowArray/ArrowSchema, and
 *   - the `_polars_plugin_*` symbol set, so the same .so loads through the unmodified plugin loader.
 * INTEGRATION.md shows the Rust binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * PwStatus on failure, with a thread-local NUL-terminated message from pw_b200_last_error()
 * (same convention as `_polars_plugin_get_last_error_message`, plugin.rs:63-72).  All entry points
 * are re-entrant; device work is issued on the calling thread's stream (pw_b200_set_stream).
 * There is no CPU fallback: without a CUDA device every compute entry fails with PW_ERR_CUDA.
 */
#ifndef POLARWAY_B200_H
#define POLARWAY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; this header is the export list */
#endif

/* ---- Arrow C Data Interface (verbatim ABI; polars-arrow/src/ffi/generated.rs:4-34) ---------- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif

/* ---- status ------------------------------------------------------------------------------- */
typedef enum PwStatus {
  PW_OK = 0,
  PW_ERR_INVALID = -1,      /* malformed query / schema mismatch */
  PW_ERR_UNSUPPORTED = -2,  /* dtype or expression outside this path (SURVEY §8f "next") */
  PW_ERR_CUDA = -3,         /* no device, launch or allocation failure */
  PW_ERR_NOT_SORTED = -4,   /* group_by_dynamic index column not ascending inside a key */
  PW_ERR_INTERNAL = -5
} PwStatus;

/* ---- query description ---------------------------------------------------------------------- */
/* comparison operators: TotalOrdKernel::tot_{eq,ne,lt,le,gt,ge}_kernel_broadcast
 * (polars-compute/src/comparisons/mod.rs:56-76).  A null input row compares to "false"
 * (polars-compute/src/filter/mod.rs:18-28). */
typedef enum PwCmpOp { PW_EQ = 0, PW_NE = 1, PW_LT = 2, PW_LE = 3, PW_GT = 4, PW_GE = 5 } PwCmpOp;

/* aggregations: the pre-aggregatable reductions of polars-expr/src/reduce/convert.rs:23-168 that this
 * path covers (sum.rs, mean.rs, min_max.rs, count.rs, len.rs, first_last.rs). */
typedef enum PwAggKind {
  PW_SUM = 0, PW_MEAN = 1, PW_MIN = 2, PW_MAX = 3, PW_COUNT = 4, PW_LEN = 5, PW_FIRST = 6, PW_LAST = 7,
  /* the rest of the pre-aggregatable reductions (convert.rs:46-150): */
  PW_VAR = 8, PW_STD = 9,                       /* reduce/var_std.rs; PwAgg.ddof */
  PW_FIRST_NON_NULL = 10, PW_LAST_NON_NULL = 11,/* first/last(ignore_nulls=True): reduce/first_last_nonnull.rs */
  PW_NULL_COUNT = 12,                           /* reduce/count.rs NullCountReduce */
  PW_BIT_AND = 13, PW_BIT_OR = 14, PW_BIT_XOR = 15, /* reduce/bitwise.rs (integers and booleans) */
  PW_ANY = 16, PW_ALL = 17                      /* reduce/any_all.rs, ignore_nulls = true (boolean input) */
} PwAggKind;

/* polars-time/src/windows/group_by.rs:24-44 */
typedef enum PwClosedWindow { PW_CLOSED_LEFT = 0, PW_CLOSED_RIGHT = 1, PW_CLOSED_BOTH = 2, PW_CLOSED_NONE = 3 } PwClosedWindow;
typedef enum PwLabel { PW_LABEL_LEFT = 0, PW_LABEL_RIGHT = 1, PW_LABEL_DATAPOINT = 2 } PwLabel;

/* `column <op> scalar`; the scalar is given in the column's physical unit. */
typedef struct PwPredicate {
  int32_t column;          /* index into the input columns */
  int32_t op;              /* PwCmpOp */
  int32_t scalar_is_float; /* 1: use .f, 0: use .i (bit pattern for unsigned columns) */
  int32_t reserved;
  union { int64_t i; uint64_t u; double f; } scalar;
} PwPredicate;

/* one affine factor a + b*column, evaluated in f64 with one rounding per operation (no FMA) */
typedef struct PwFactor {
  double a, b;
  int32_t column;
  int32_t reserved;
} PwFactor;

#define PW_MAX_FACTORS 4
typedef struct PwAgg {
  int32_t kind;      /* PwAggKind */
  int32_t column;    /* plain input column, or -1 when `factors` is used / for PW_LEN */
  int32_t n_factors; /* 0: plain column; >0: product of factors[0..n) (f64) */
  int32_t ddof;      /* PW_VAR / PW_STD: delta degrees of freedom (Polars default 1); 0 for every other kind */
  PwFactor factors[PW_MAX_FACTORS];
  const char* name;  /* output column name (borrowed) */
} PwAgg;

/* DynamicGroupOptions (polars-time/src/group_by/dynamic.rs:19-39); fixed durations only, expressed
 * in the index column's own unit. */
typedef struct PwDynamic {
  int32_t index_column;
  int32_t closed;             /* PwClosedWindow */
  int32_t label;              /* PwLabel */
  int32_t include_boundaries; /* emit _lower_boundary/_upper_boundary */
  int64_t every, period, offset;
} PwDynamic;

/* strategy overrides, the analogue of POLARS_FORCE_PARTITION / POLARS_NO_PARTITION /
 * POLARS_HOT_TABLE_SIZE (group_by_streaming.rs:178-197, nodes/group_by.rs:487-489): they let
 * unit-sized inputs exercise the eviction, spill and merge paths. */
#define PW_FLAG_FORCE_HOT_TABLE (1ull << 0)    /* shared-memory hot table + HBM spill tier */
#define PW_FLAG_FORCE_GLOBAL_TABLE (1ull << 1) /* HBM open-addressing table only */
#define PW_FLAG_FORCE_SEGMENTED (1ull << 2)    /* sorted-run segmented reduction (dynamic, no keys) */
#define PW_FLAG_NO_SEGMENTED (1ull << 3)       /* dynamic without keys through the hash path */
#define PW_FLAG_FORCE_PARTITION (1ull << 4)    /* radix-partition the rows by key hash first (high-cardinality tier) */
#define PW_FLAG_NO_PARTITION (1ull << 5)       /* never partition (POLARS_NO_PARTITION) */
#define PW_FLAG_NO_DENSE_IDS (1ull << 6)       /* never map a small integer key range to dense ids: always the hash index */
#define PW_FLAG_KEYS_SORTED (1ull << 8)        /* the key columns are sorted (the IsSorted flag of the reference's Series): groups are
                                                 contiguous runs -> run-boundary + segmented-reduce scan, pw_runs.cuh
                                                 (polars-core/src/frame/group_by/into_groups.rs:65-129).  A wrong promise costs time,
                                                 not correctness: the table still merges by key */
#define PW_FLAG_NO_BUCKETS (1ull << 7)         /* dense ids through the per-cell hot table, never the bucket tier (pw_bucket.cuh) */

#define PW_ABI_VERSION 1u
typedef struct PwQuery {
  uint32_t abi_version;  /* PW_ABI_VERSION */
  int32_t maintain_order;
  int32_t n_predicates;
  int32_t n_keys;
  int32_t n_aggs;
  int32_t hot_table_slots;    /* 0 = default; power of two */
  const PwPredicate* predicates; /* conjunction */
  const int32_t* key_columns;
  const PwAgg* aggs;
  const PwDynamic* dynamic;   /* NULL for a plain group_by */
  uint64_t flags;
  int64_t row_offset;         /* global index of row 0 of this shard (multi-GPU first/last order) */
  int64_t initial_table_slots;/* 0 = estimate from a key sample */
} PwQuery;

/* ---- library / device -------------------------------------------------------------------------- */
uint32_t pw_b200_abi_version(void);
const char* pw_b200_last_error(void);          /* thread-local, valid until the next call on this thread */
int pw_b200_device_count(void);                /* 0 when no CUDA device is usable */
int pw_b200_set_device(int device);            /* calling thread's device (default 0) */
int pw_b200_set_stream(void* cuda_stream);     /* calling thread's cudaStream_t (default: legacy stream 0) */

/* per-call phase timings measured with CUDA events on the call's stream — the analogue of the
 * reference's NodeTimer (start,end,name) triples (polars-expr/src/state/node_timer.rs:14-72). */
typedef struct PwTimings {
  float h2d_ms, estimate_ms, scan_ms, finalize_ms, d2h_ms, total_device_ms;
  int64_t n_rows, n_groups, table_slots;
  int32_t strategy;      /* 1 hot table, 2 global table, 3 segmented, 4 hot table with dense ids (small integer key range), 5 partitioned,
                            6 the same per (key, window) for group_by_dynamic by one dense key,
                            7 dense ids bucketed per tile, accumulators in registers (pw_bucket.cuh), 8 sorted-key runs (pw_runs.cuh),
                            9 the bucket tier with ids from a shared-memory key index (sparse integers, strings, several keys),
                            10 two-level radix partition + one shared-memory table per partition (pw_radix.cuh; high cardinality) */
  int32_t retries;       /* table growth re-runs */
  int64_t kernel_launches; /* launches of this library's kernels in the last call */
  int64_t spilled_rows;  /* rows that bypassed the hot table (spill tier) */
  float scan_kernel_ms;  /* the dominant kernel alone (events immediately around its launch) */
  float reserved;         /* 1 = the query-shape specialised (NVRTC) kernel ran, 0 = the ahead-of-time kernel */
  float host_ms;          /* wall-clock time spent inside the last pw_b200_frame_groupby call (host + device) */
  float partition_ms;     /* strategies 5 and 10: histogram + scatter passes (part of scan_ms) */
  int32_t jit_compiles;   /* NVRTC compilations during the last call */
  int32_t jit_cache_hits; /* specialised kernels loaded from the on-disk cubin cache during the last call */
} PwTimings;
int pw_b200_last_timings(PwTimings* out);

/* ---- resident frames (inputs already in HBM) ---------------------------------------------------- */
typedef struct PwFrame PwFrame; /* opaque: columns resident on one device */

/* Copy `n_cols` host Arrow arrays to the device.  Arrays are borrowed for the duration of the call
 * (not released).  All columns must have the same length.  Supported formats: b c C s S i I l L f g,
 * tdD, ts{s,m,u,n}:*, tD{s,m,u,n}, vu, vz (views of any length: buffers = [validity, views, data..., variadic sizes];
 * values longer than 12 bytes are canonicalised when the frame is created, pw_views.cu). */
int pw_b200_frame_upload(const struct ArrowArray* const* cols, const struct ArrowSchema* const* schemas,
                         size_t n_cols, PwFrame** out);
/* Wrap buffers that already live on the current device (zero copy).  `cols[i]->buffers` hold DEVICE
 * pointers laid out as the Arrow spec prescribes; they must stay alive until pw_b200_frame_free. */
int pw_b200_frame_from_device(const struct ArrowArray* const* cols, const struct ArrowSchema* const* schemas,
                              size_t n_cols, PwFrame** out);
int64_t pw_b200_frame_num_rows(const PwFrame* f);
int pw_b200_frame_free(PwFrame* f);

/* ---- the operator --------------------------------------------------------------------------------- */
/* filter -> group_by -> agg (or group_by_dynamic) over a resident frame.  Results are returned as host
 * Arrow arrays the caller owns (call ->release): out_cols/out_schemas must have room for
 * *n_out entries on input; on output *n_out is the number of result columns
 * ([keys..., (_lower_boundary,_upper_boundary), index] + aggs). */
int pw_b200_frame_groupby(const PwQuery* q, const PwFrame* frame, struct ArrowArray* out_cols,
                          struct ArrowSchema* out_schemas, size_t* n_out);

/* Same, taking host Arrow columns: upload + operator + free.  This is the call a GroupByExec /
 * GroupByNode / GroupByDynamicExec replacement makes. */
int pw_b200_filter_groupby_agg(const PwQuery* q, const struct ArrowArray* const* cols,
                               const struct ArrowSchema* const* schemas, size_t n_cols,
                               struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out);

/* FilterExec alone (polars-mem-engine/src/executors/filter.rs:93-121): predicate -> selection vector ->
 * every column compacted.  out_* sized n_cols. */
int pw_b200_filter(const PwPredicate* predicates, int32_t n_predicates, const struct ArrowArray* const* cols,
                   const struct ArrowSchema* const* schemas, size_t n_cols, struct ArrowArray* out_cols,
                   struct ArrowSchema* out_schemas);
/* predicate -> compacted selection vector (ascending u32 row ids) on a resident frame; returns the
 * selected count through *n_selected and the ids as one UInt32 Arrow array. */
int pw_b200_frame_filter_select(const PwPredicate* predicates, int32_t n_predicates, const PwFrame* frame,
                                struct ArrowArray* out_ids, struct ArrowSchema* out_schema, int64_t* n_selected);

/* GroupsIdx / GroupsSlice construction (polars-core/src/frame/group_by/position.rs:16-20, 252-267):
 * first[g] (u32), offsets[g+1] (u32... as Int64) and all row ids (u32) ordered by group, ascending inside a
 * group; groups ordered by first occurrence when maintain_order, else unspecified. */
int pw_b200_frame_group_tuples(const PwFrame* frame, const int32_t* key_columns, int32_t n_keys,
                               int32_t maintain_order, struct ArrowArray* out_first, struct ArrowArray* out_offsets,
                               struct ArrowArray* out_row_ids, struct ArrowSchema* out_schemas /*[3]*/);

/* GroupsSlice of SORTED key columns: [first, len] per run of equal keys, in row order — partition_to_groups
 * (polars-arrow/src/legacy/kernels/sort_partition.rs:168) / create_groups_from_sorted
 * (polars-core/src/frame/group_by/into_groups.rs:65-129).  Null keys form runs like values.  The caller vouches for the
 * order (as the reference's sorted flag does); a key that reappears later starts another slice.  Both outputs UInt32. */
int pw_b200_frame_group_slices(const PwFrame* frame, const int32_t* key_columns, int32_t n_keys,
                               struct ArrowArray* out_first, struct ArrowArray* out_len, struct ArrowSchema* out_schemas /*[2]*/);

/* ---- multi-GPU partial aggregates ------------------------------------------------------ */
/* Phase 1 on every GPU: build the local partial-aggregate table and export it as fixed-width packed rows
 * partitioned by owner = hash(key) -> [0, n_parts) (the role of HashPartitioner, polars-utils/src/
 * hashing.rs:72-121).  The packed buffer is DEVICE memory owned by the returned handle; part_offsets
 * (host, n_parts+1 entries, in rows) delimit each owner's slice.  row_bytes is the packed row size. */
typedef struct PwPartial PwPartial;
int pw_b200_frame_groupby_partial(const PwQuery* q, const PwFrame* frame, int32_t n_parts, PwPartial** out);
int64_t pw_b200_partial_row_bytes(const PwPartial* p);
const void* pw_b200_partial_device_rows(const PwPartial* p);
/* device-to-device copy of all packed rows into a caller-owned buffer (e.g. a torch tensor handed to NCCL) */
int pw_b200_partial_copy_rows(const PwPartial* p, void* dst_device);
int pw_b200_partial_offsets(const PwPartial* p, int64_t* part_offsets /* n_parts+1 */);
int pw_b200_partial_free(PwPartial* p);
/* Phase 2 on the owner: merge packed rows received from every peer (device pointer, n_rows rows of
 * row_bytes) with GroupedReduction::combine semantics (polars-expr/src/reduce/mod.rs:94-105) and
 * finalise.  `schema_from` supplies dtypes/names (any rank's frame with the same schema). */
int pw_b200_merge_partials(const PwQuery* q, const PwFrame* schema_from, const void* device_rows, int64_t n_rows,
                           struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out);
/* Small-result exchange (one collective, no host round trip between export and merge): every rank writes
 * [n_rows | cap_rows x (owner, packed row)] into `send_device` (1 + cap_rows * (row_words + 1) 64-bit words), the
 * buffers are all-gathered, and each rank merges the rows it owns straight from the gathered buffer.
 * partial_into returns 1 when this rank has more than cap_rows groups (its header then says so to every peer);
 * merge_gathered returns 1 when any rank overflowed: all ranks then repeat through the general path above. */
int64_t pw_b200_partial_row_words(const PwQuery* q, const PwFrame* frame);
int pw_b200_frame_groupby_partial_into(const PwQuery* q, const PwFrame* frame, int32_t n_parts, void* send_device, int64_t cap_rows);
int pw_b200_merge_gathered(const PwQuery* q, const PwFrame* schema_from, const void* gathered_device, int32_t world, int64_t cap_rows,
                           int32_t my_rank, struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out);

/* ---- expression-plugin compatibility shim (polars-ffi/src/version_0.rs, plugin.rs:75-142) ----------- */
typedef struct SeriesExport {
  struct ArrowSchema* field;
  struct ArrowArray** arrays;
  size_t len; /* number of chunks */
  void (*release)(struct SeriesExport*);
  void* private_data;
} SeriesExport;
typedef struct CallerContext { uint64_t bitflags; } CallerContext;

uint32_t _polars_plugin_get_version(void);                 /* (major<<16)|minor = (0,1) */
const char* _polars_plugin_get_last_error_message(void);
/* inputs = every column of the frame; kwargs = pickled dict describing the query;
 * returns a Struct series whose fields are the result columns. */
void _polars_plugin_filter_groupby_agg(const SeriesExport* inputs, size_t n_inputs, const uint8_t* kwargs,
                                       size_t kwargs_len, SeriesExport* return_value, const CallerContext* ctx);
void _polars_plugin_field_filter_groupby_agg(const struct ArrowSchema* fields, size_t n_fields,
                                             struct ArrowSchema* out, const uint8_t* kwargs, size_t kwargs_len);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* POLARWAY_B200_H */
