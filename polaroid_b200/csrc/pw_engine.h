// pw_engine.h — host-side structures of libpolarway_b200 (internal; the public ABI is include/polarway_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/polarway_b200.h"
#include "pw_finalize.cuh"
#include "pw_plan.h"
#include "pw_segmented.cuh"

namespace pw {

struct ThreadCtx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  PwTimings timings{};
  cudaEvent_t ev[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_ready = false;
  int sm_count = 0;
  bool pool_ready = false;
  void* staging = nullptr;  // pinned host block for small results (one D2H copy per query), lives with the thread
};
static const size_t STAGING_BYTES = 1u << 20;
ThreadCtx& ctx();
int fail(int code, const char* fmt, ...);

#define PW_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return ::pw::fail(PW_ERR_CUDA, "%s failed: %s (%s:%d)", #call,          \
                                             cudaGetErrorString(_e), __FILE__, __LINE__);          \
  } while (0)
#define PW_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)

struct FrameColumn {
  std::string format, name;
  int32_t dtype = DT_I64;       // physical DType
  int64_t null_count = 0;
  const void* values = nullptr; // device
  const uint8_t* validity = nullptr;
  int32_t bit_offset = 0;
  void* owned_values = nullptr; // freed with the frame
  void* owned_validity = nullptr;
  // view columns: the variadic data buffers behind views longer than 12 bytes (pw_views.cu)
  std::vector<const void*> var_bufs;   // device pointers
  std::vector<void*> owned_var;        // the uploaded ones, freed with the frame
  const void** d_var_ptrs = nullptr;   // device copy of var_bufs (owned)
  bool has_long = false;               // long views present: `values` is the canonicalised copy (owned)
  // var/std shift (a sample mean of the column), computed the first time a query needs it (guarded by PwFrame::mu)
  mutable bool shift_known = false;
  mutable double var_shift = 0.0;
};

}  // namespace pw

namespace pw {
// What the key-sample pilot of a query learnt about a frame's key columns (run_groupby).  Cached on the frame per key
// set: a resident frame pays the pilot launch and its host synchronisation once.  The statistics are hints only — a
// stale entry (zero-copy frames can change under us) costs speed, never correctness: the table grows on overflow and
// keys outside a dense range take the HBM table.
struct PilotStats {
  unsigned long long distinct_strided = 0, distinct_block = 0;
  unsigned long long kmax_u = 0, kmin_n = 0;
  int32_t overflow = 0;
};
}  // namespace pw

struct PwFrame {
  int device = 0;
  int64_t n_rows = 0;
  std::vector<pw::FrameColumn> cols;
  mutable std::mutex mu;
  mutable std::map<std::string, pw::PilotStats> pilot;
  mutable std::map<std::string, int> sorted_cache;   // check_sorted_within_keys: 1 ascending, 2 not (per index column / per keys + index)
};

namespace pw {

// one result column: how to emit it + its Arrow schema
struct OutCol {
  std::string name, format;
  const FrameColumn* src_col = nullptr;  // view keys: the frame column whose data buffers long values reference
  EmitDesc emit{};
  int32_t out_dtype = DT_I64;  // buffer dtype (DT_VIEW for string keys)
  bool nullable = false;
};

struct Lowered {
  ScanPlan plan{};
  std::vector<OutCol> outs;
  std::vector<SortSpec> sort;  // least-significant first
  int null_word = -1;
  int single_key_null = 0;
  int n_nc = 0;                // raw slots in use
  bool tumbling = true;
};

int launch_scan_nc4(const ScanPlan& P, int sm, cudaStream_t st);
int launch_scan_nc12(const ScanPlan& P, int sm, cudaStream_t st);
int launch_scan_jit(const ScanPlan& P, int nc, int kw, bool hot, int threads, int sm_count, cudaStream_t st);
int launch_seg_jit(const ScanPlan& P, const SegParams& sp, int nc, int threads, size_t smem, int sm_count, cudaStream_t st);
int launch_bucket_jit(const ScanPlan& P, int nc, int kw, int sm_count, cudaStream_t st);
int launch_runs_jit(const ScanPlan& P, int nc, int kw, int sm_count, cudaStream_t st);
int launch_part_jit(const ScanPlan& P, const PartParams& pp, int nc, int kw, int sm_count, cudaStream_t st);
int launch_overlap_jit(const ScanPlan& P, int nc, int kw, int sm_count, cudaStream_t st);
int launch_radix_jit(const ScanPlan& P, const RadixParams& rp, int nc, int kw, size_t smem, int64_t work_ctas, int sm_count, cudaStream_t st);
bool jit_available();
struct PilotParams;
int launch_pilot_jit(const ScanPlan& P, const PilotParams& pp, int nc, int kw, cudaStream_t st);
int jit_selftest_compile(const ScanPlan& P, int nc, int kw, bool hot, int threads, std::string* err);
int launch_scan_aot(ScanPlan P, int sm, cudaStream_t st);
int ensure_device();
int dev_alloc(void** p, size_t bytes);
void dev_free(void* p);

int lower_query(const PwQuery* q, const PwFrame* f, Lowered* out);
// Small results ("count on the device" mode): when the table is small, run_groupby does not wait for the group count.
// It allocates the result block up front — [Control | every result column sized for the table's capacity] — uses the
// block's header as the scan's control block, and leaves the count on the device; ordering and emission are launched
// for the capacity and read the count on the device; emit_results copies the block to the host and synchronises ONCE.
struct RunOpts {
  uint64_t min_cap = 0;      // lower bound of the HBM table size (retry after an overflow seen late)
  bool allow_deferred = false;
  bool control_only = false; // deferred, but the caller does not emit columns: the block is just the control header
};
struct RunState {
  bool deferred = false;     // *n_groups_out is the capacity bound, the count is dctl->counter (device)
  Control* dctl = nullptr;   // deferred: header of `block`
  char* block = nullptr;     // deferred: result block (freed by emit_results)
  uint64_t cap = 0;          // table slots used
};
constexpr int PW_RETRY = 1;  // emit_results (deferred): the scan overflowed its table — run again with a larger one
int run_groupby(const PwQuery* q, const PwFrame* f, Lowered& L, Table* table_out, uint32_t** slot_list_out, uint64_t* n_groups_out,
                const RunOpts* opts = nullptr, RunState* state = nullptr);
int emit_results(const Lowered& L, const Table& T, const uint32_t* slot_list, uint64_t n_groups,
                 struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out, RunState* state = nullptr);
void free_table(Table& T);
int alloc_table_raw(Table* T, int n_kw, int n_acc, uint64_t cap, int32_t* overflow, unsigned long long* spilled);
// deferred result block for `bound` result rows: [Control | columns]; *fits = false when it would not fit the staging copy
int alloc_result_block(const Lowered& L, uint64_t bound, char** block, bool* fits);
int order_groups(const Lowered& L, const Table& T, int kw, uint32_t** slots_io, uint64_t G, const unsigned long long* g_dev = nullptr);

// segmented (sorted-run) dynamic path, pw_segmented.cu
int run_dynamic_segmented(const PwQuery* q, const PwFrame* f, struct ArrowArray* out_cols, struct ArrowSchema* out_schemas,
                          size_t* n_out, bool* handled);

// Arrow helpers (pw_arrow.cpp)
int parse_format(const char* fmt, int32_t* dtype);
void* host_alloc(size_t bytes);  // result buffers: pinned pool for large ones; release with host_free
void host_free(void* p);
int make_host_array(int64_t length, int64_t null_count, void* validity, void* values, size_t n_extra_buffers,
                    struct ArrowArray* out);
int make_schema(const char* format, const char* name, bool nullable, struct ArrowSchema* out);
// view array with one variadic data buffer (takes ownership of all four blocks)
int make_host_view_array(int64_t length, int64_t null_count, void* validity, void* values, void* data, int64_t data_bytes, struct ArrowArray* out);
// group_by_dynamic with keys: is the index ascending inside every key? (pw_filter.cu)
int check_sorted_within_keys(const PwQuery* q, const PwFrame* frame);
// long string keys (pw_views.cu)
int views_intern(FrameColumn* col, int64_t n);
int views_gather_long(const FrameColumn* col, void* h_views, const void* h_validity, uint64_t G, void** h_data, int64_t* data_bytes);

}  // namespace pw
