// pw_dynamic.cu — group_by_dynamic without keys over a sorted index column: window construction in closed
// form + segmented (slice) reduction.  Covers every `closed` mode and overlapping windows (period > every).
//
// Reference: polars-time/src/windows/group_by.rs:79-246 (group_by_windows / update_groups_and_bounds: a
// sequential two-pointer sweep that emits [start,len] for every NON-EMPTY window), windows/window.rs:25-53,
// 115-170 (first window), windows/bounds.rs:33-76 (membership), then slice aggregation
// (polars-core/src/frame/group_by/aggregations/mod.rs:184-191).  On the GPU the sweep becomes:
//   windows are  w_i = [s0 + i*every, s0 + i*every + period)   i = 0 .. W-1   (s0 from the first row)
//   row j is a member of the contiguous index range [lo_j, hi_j]; it is the FIRST member of the windows
//   [max(lo_j, hi_{j-1}+1), hi_j] because the rows are sorted.  Counting those per row + an exclusive scan
//   gives every non-empty window its output position and start row with no per-window loop (sparse data
//   costs nothing); the end row is a binary search; one warp then reduces each [start,end) slice with the
//   same row evaluation as the hash path and warp shuffles.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "pw_engine.h"
#include "pw_scan.cuh"
#include "pw_segmented.cuh"

namespace pw {

struct WinParams {
  int64_t s0, every, period;
  int64_t n_windows;  // candidate windows (start < boundary.stop)
  int32_t closed;
};

__device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return -floor_div(-a, b); }

// index range of the windows that contain t (may be empty: lo > hi)
__device__ __forceinline__ void member_range(const WinParams& w, int64_t t, int64_t& lo, int64_t& hi) {
  const int64_t rel = t - w.s0;
  const bool right_closed = w.closed == PW_CLOSED_RIGHT || w.closed == PW_CLOSED_BOTH;  // t <= stop allowed
  const bool left_closed = w.closed == PW_CLOSED_LEFT || w.closed == PW_CLOSED_BOTH;    // t >= start allowed
  // exit condition: t < s_i + period (or <=)   ->  smallest admissible i
  lo = right_closed ? ceil_div(rel - w.period, w.every) : floor_div(rel - w.period, w.every) + 1;
  // entry condition: t >= s_i (or >)           ->  largest admissible i
  hi = left_closed ? floor_div(rel, w.every) : ceil_div(rel, w.every) - 1;
  if (lo < 0) lo = 0;
  if (hi > w.n_windows - 1) hi = w.n_windows - 1;
}

__device__ __forceinline__ int64_t load_time(const RawSlot& ts, int64_t row) {
  const uint4 r = load_pair(ts.values, ts.dtype, row, row + 1, false);
  return (int64_t)decode(r, ts.dtype, 0);
}

// per row: how many windows have this row as their first member
static __global__ void dyn_count_kernel(RawSlot ts, int64_t n, WinParams w, uint32_t* counts, int32_t* not_sorted) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t t = load_time(ts, j);
  int64_t lo, hi;
  member_range(w, t, lo, hi);
  int64_t first = lo;
  if (j > 0) {
    const int64_t tp = load_time(ts, j - 1);
    if (tp > t) *not_sorted = 1;
    int64_t plo, phi;
    member_range(w, tp, plo, phi);
    if (phi + 1 > first) first = phi + 1;
  }
  const int64_t c = hi - first + 1;
  counts[j] = c > 0 ? (uint32_t)c : 0u;
}

// window list: index i, start row
static __global__ void dyn_emit_kernel(RawSlot ts, int64_t n, WinParams w, const uint32_t* counts, const uint64_t* offsets,
                                       int64_t* win_index, int64_t* win_start) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t c = counts[j];
  if (c == 0u) return;
  int64_t lo, hi;
  member_range(w, load_time(ts, j), lo, hi);
  const int64_t first = hi - (int64_t)c + 1;
  const uint64_t o = offsets[j];
  for (uint32_t r = 0; r < c; ++r) { win_index[o + r] = first + r; win_start[o + r] = j; }
}

// end row of each window: first row >= start that fails the exit condition (binary search; rows sorted)
static __global__ void dyn_end_kernel(RawSlot ts, int64_t n, WinParams w, int64_t n_out, const int64_t* win_index,
                                      const int64_t* win_start, int64_t* win_end, uint64_t* key_word, int64_t k0) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n_out) return;
  const int64_t i = win_index[x];
  const int64_t stop = w.s0 + i * w.every + w.period;
  const bool right_closed = w.closed == PW_CLOSED_RIGHT || w.closed == PW_CLOSED_BOTH;
  int64_t lo = win_start[x] + 1, hi = n;  // answer in [lo, hi]
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    const int64_t t = load_time(ts, mid);
    const bool inside = right_closed ? t <= stop : t < stop;
    if (inside) lo = mid + 1; else hi = mid;
  }
  win_end[x] = lo;
  key_word[x] = (uint64_t)(k0 + i);  // window index on the grid origin + k*every used by the emit kernels
}

// one warp per window: lanes stride over the slice, accumulate into lane-private shared cells, then a
// warp-shuffle reduction per accumulator word
struct LaneSink {
  uint64_t* cells;  // [n_acc][32] for this warp
  int lane;
  template <int OP>
  __device__ __forceinline__ void add(const ScanPlan&, int a, uint64_t x) const {
    uint64_t* q = cells + a * 32 + lane;
    *q = acc_combine(OP, *q, x);
  }
};

template <int NC>
static __global__ void __launch_bounds__(256) slice_agg_kernel(const __grid_constant__ ScanPlan P, int64_t n_out, const int64_t* win_start,
                                                               const int64_t* win_end) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NV = NVof<NC>::value;
  using CT = RtCtl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  uint64_t* cells = (uint64_t*)smem_raw + (size_t)warp * P.n_acc * 32;
  for (int64_t x = (int64_t)blockIdx.x * warps + warp; x < n_out; x += (int64_t)gridDim.x * warps) {
    for (int a = 0; a < P.n_acc; ++a) cells[a * 32 + lane] = acc_init(P.accs[a].op);
    __syncwarp();
    const int64_t lo = win_start[x], hi = win_end[x];
    const LaneSink sink{cells, lane};
    for (int64_t row = lo + lane; row < hi; row += 32) {
      Row<NC> r;
      r.in_valid = 0;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        r.in[c] = 0;
        if (c < P.n_slots) {
          const uint4 raw = load_row(P.slots[c], row);
          r.in[c] = decode(raw, P.slots[c].dtype, 0);
          r.in_valid |= (load_valid_pair(P.slots[c], row, row + 1) & 1u) << c;
        }
      }
      RowOut<1, NV> o;
      o.k[0] = 0; o.alive = true; o.sentinel_free = true; o.row = row;
      row_vexprs<CT, NC, NV>(P, r, o.v, o.v_valid);
      o.tval = pick<NC>(r.in, P.dyn.slot);
      accumulate_row<CT, NV, 1>(P, o, (uint64_t)(row + P.row_offset), sink);
    }
    __syncwarp();
    for (int a = 0; a < P.n_acc; ++a) {
      const int op = P.accs[a].op;
      uint64_t v = cells[a * 32 + lane];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v = acc_combine(op, v, __shfl_xor_sync(0xffffffffu, v, d));
      if (lane == 0) tacc(P.table, a, (uint64_t)x) = v;
    }
    __syncwarp();
  }
}


static int64_t floor_div_h(int64_t a, int64_t b) {
  int64_t q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

// dense window table helpers of the tumbling fast path
static __global__ void seg_fill_keys_kernel(uint64_t* keys, int64_t k0, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (uint64_t)(k0 + (int64_t)i);
}
static __global__ void seg_flags_kernel(Table T, int acc_len, uint64_t n, unsigned char* flags) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = tacc(T, acc_len, i) != 0ull;
}

// Tumbling windows (every row in at most one window) and a dense window range that fits memory: the one-pass
// segmented kernel of pw_segmented.cuh.  Returns handled = false when the shape does not qualify.
static int run_segmented_fast(const PwQuery* q, const PwFrame* f, Lowered& L, struct ArrowArray* out_cols, struct ArrowSchema* out_schemas,
                              size_t* n_out, bool* handled) {
  *handled = false;
  ThreadCtx& c = ctx();
  ScanPlan& P = L.plan;
  const PwDynamic& d = *q->dynamic;
  const int64_t n = f->n_rows;
  if (n == 0) return 0;
  const RawSlot ts = P.slots[P.dyn.slot];
  int64_t t_first = 0, t_last = 0;
  {
    const int w = ts.dtype == DT_I32 ? 4 : 8;
    int64_t a = 0, b = 0;
    PW_CUDA(cudaMemcpyAsync(&a, (const char*)ts.values, w, cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaMemcpyAsync(&b, (const char*)ts.values + (size_t)(n - 1) * w, w, cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaStreamSynchronize(c.stream));
    t_first = w == 4 ? (int64_t)(int32_t)a : a;
    t_last = w == 4 ? (int64_t)(int32_t)b : b;
  }
  if (t_last < t_first) return fail(PW_ERR_NOT_SORTED, "argument in operation 'group_by_dynamic' is not sorted, please sort the 'expr/series/column' first");
  // a window that contains t has index floor((t - origin) / every) or one less (right-closed edges)
  const int64_t k_lo = floor_div_h(t_first - d.offset, d.every) - 1;
  const int64_t k_hi = floor_div_h(t_last - d.offset, d.every) + 1;
  const int64_t n_dense = k_hi - k_lo + 1;
  if (n_dense > (int64_t)1 << 25 || n_dense > 4 * n + 1024) return 0;  // sparse index: the general path lists only non-empty windows
  *handled = true;
  PwTimings& tm = c.timings;
  tm.n_rows = n; tm.strategy = 3;
  struct Ctl { int32_t not_sorted; int32_t overflow; unsigned long long spilled; int32_t n_sel; int32_t pad; } hctl{};
  Ctl* dctl = nullptr;
  void* p = nullptr;
  PW_TRY(dev_alloc(&p, sizeof(Ctl))); dctl = (Ctl*)p;
  PW_CUDA(cudaMemsetAsync(dctl, 0, sizeof(Ctl), c.stream));
  PW_CUDA(cudaEventRecord(c.ev[2], c.stream));
  Table T{};
  const uint64_t nn = (uint64_t)n_dense + 2;
  PW_TRY(dev_alloc(&p, nn * 8)); T.keys = (uint64_t*)p;
  PW_TRY(dev_alloc(&p, nn * 4)); T.state = (uint32_t*)p;
  PW_TRY(dev_alloc(&p, nn * 8 * (uint64_t)P.n_acc)); T.accs = (uint64_t*)p;
  T.cap = (uint64_t)n_dense; T.key_sw = T.acc_sw = nn; T.key_ss = T.acc_ss = 1;
  T.overflow = &dctl->overflow; T.spilled = &dctl->spilled;
  AccOps ops{};
  ops.n = P.n_acc;
  for (int a = 0; a < P.n_acc; ++a) ops.op[a] = P.accs[a].op;
  table_init_kernel<<<(int)std::min<uint64_t>((nn + 255) / 256, 148 * 8), 256, 0, c.stream>>>(T, 2 /*no key sentinel*/, ops);
  PW_CUDA(cudaGetLastError());
  seg_fill_keys_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, c.stream>>>(T.keys, k_lo, nn);
  PW_CUDA(cudaGetLastError());
  P.table = T; P.not_sorted = &dctl->not_sorted; P.check_sorted = 1; P.hot_slots = 0;
  SegParams sp{k_lo, n_dense};
  const size_t smem = (size_t)8 * P.n_acc * 32 * 8;
  const bool narrow = P.n_slots <= 4 && P.n_vexpr <= NVof<4>::value;
  int per_sm = 1;
  if (narrow) PW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_kernel<4>, 256, smem));
  else PW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_kernel<12>, 256, smem));
  const int64_t n_steps = (n + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  int64_t grid = std::min<int64_t>((int64_t)c.sm_count * std::max(per_sm, 1), (n_steps + 7) / 8);
  if (grid < 1) grid = 1;
  PW_CUDA(cudaEventRecord(c.ev[8], c.stream));
  const int jrc = launch_seg_jit(P, sp, narrow ? 4 : 12, 256, smem, c.sm_count, c.stream);  // query-shape specialised (NVRTC)
  if (jrc < 0) return jrc;
  if (jrc > 0) {
    if (narrow) seg_kernel<4><<<(unsigned)grid, 256, smem, c.stream>>>(P, sp);
    else seg_kernel<12><<<(unsigned)grid, 256, smem, c.stream>>>(P, sp);
    PW_CUDA(cudaGetLastError());
  }
  PW_CUDA(cudaEventRecord(c.ev[9], c.stream));
  // non-empty windows, ascending
  int acc_len = -1;
  for (int a = 0; a < P.n_acc; ++a) if (P.accs[a].src == SRC_ONE) acc_len = a;
  unsigned char* flags = nullptr; uint32_t* slots = nullptr;
  PW_TRY(dev_alloc(&p, (size_t)n_dense)); flags = (unsigned char*)p;
  PW_TRY(dev_alloc(&p, (size_t)n_dense * 4)); slots = (uint32_t*)p;
  seg_flags_kernel<<<(unsigned)((n_dense + 255) / 256), 256, 0, c.stream>>>(T, acc_len, (uint64_t)n_dense, flags);
  PW_CUDA(cudaGetLastError());
  cub::CountingInputIterator<uint32_t> iota(0u);
  size_t tmp_bytes = 0;
  cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, flags, slots, &dctl->n_sel, (int)n_dense, c.stream);
  void* tmp = nullptr;
  PW_TRY(dev_alloc(&tmp, tmp_bytes));
  PW_CUDA(cub::DeviceSelect::Flagged(tmp, tmp_bytes, iota, flags, slots, &dctl->n_sel, (int)n_dense, c.stream));
  PW_CUDA(cudaMemcpyAsync(&hctl, dctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c.stream));
  PW_CUDA(cudaStreamSynchronize(c.stream));
  tm.kernel_launches += 6;
  dev_free(tmp); dev_free(flags);
  int rc = 0;
  if (hctl.not_sorted) rc = fail(PW_ERR_NOT_SORTED, "argument in operation 'group_by_dynamic' is not sorted, please sort the 'expr/series/column' first");
  const uint64_t n_win = (uint64_t)hctl.n_sel;
  tm.n_groups = (int64_t)n_win; tm.table_slots = n_dense;
  PW_CUDA(cudaEventRecord(c.ev[3], c.stream));
  if (!rc) rc = emit_results(L, T, slots, n_win, out_cols, out_schemas, n_out);
  dev_free(slots); dev_free(dctl);
  free_table(T);
  if (rc) return rc;
  float ms;
  if (cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]) == cudaSuccess) tm.scan_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]) == cudaSuccess) tm.finalize_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[4], c.ev[5]) == cudaSuccess) tm.d2h_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[0], c.ev[5]) == cudaSuccess) tm.total_device_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[8], c.ev[9]) == cudaSuccess) tm.scan_kernel_ms = ms;
  return 0;
}

int run_dynamic_segmented(const PwQuery* q, const PwFrame* f, struct ArrowArray* out_cols, struct ArrowSchema* out_schemas,
                          size_t* n_out, bool* handled) {
  *handled = false;
  if (!q->dynamic || q->n_keys != 0) return 0;
  const PwDynamic& d = *q->dynamic;
  const bool overlapping = d.closed == PW_CLOSED_BOTH ? d.period >= d.every : d.period > d.every;
  if (q->flags & PW_FLAG_NO_SEGMENTED) return 0;       // windows through the hash path (test hook)
  if (q->n_predicates > 0 && overlapping) return 0;    // filter + overlapping windows: the hash path (pw_overlap.cuh) applies the predicate in registers
  ThreadCtx& c = ctx();
  Lowered L;
  PW_TRY(lower_query(q, f, &L));
  if (!overlapping && !(q->flags & PW_FLAG_FORCE_SEGMENTED)) {
    // (PW_FLAG_FORCE_SEGMENTED here means: use the general window-list path even for tumbling windows — test hook)
    int rc = run_segmented_fast(q, f, L, out_cols, out_schemas, n_out, handled);
    if (rc || *handled) return rc;
  }
  if (q->n_predicates > 0) return 0;  // tumbling + filter that did not qualify: the hash path applies the predicate in registers
  ScanPlan& P = L.plan;
  const int64_t n = f->n_rows;
  *handled = true;
  PwTimings& tm = c.timings;
  tm.n_rows = n; tm.strategy = 3;

  struct Ctl { int32_t not_sorted; int32_t pad; unsigned long long total; } hctl{};
  Ctl* dctl = nullptr;
  { void* p = nullptr; PW_TRY(dev_alloc(&p, sizeof(Ctl))); dctl = (Ctl*)p; }
  PW_CUDA(cudaMemsetAsync(dctl, 0, sizeof(Ctl), c.stream));
  PW_CUDA(cudaEventRecord(c.ev[2], c.stream));

  Table T{};
  uint64_t n_win = 0;
  int64_t *win_index = nullptr, *win_start = nullptr, *win_end = nullptr;
  uint32_t* slots = nullptr;
  if (n > 0) {
    const RawSlot ts = P.slots[P.dyn.slot];
    // first and last index value -> first window (window.rs:25-53,115-170) and the candidate count
    int64_t t_first = 0, t_last = 0;
    {
      const int w = ts.dtype == DT_I32 ? 4 : 8;
      int64_t a = 0, b = 0;
      PW_CUDA(cudaMemcpyAsync(&a, (const char*)ts.values, w, cudaMemcpyDeviceToHost, c.stream));
      PW_CUDA(cudaMemcpyAsync(&b, (const char*)ts.values + (size_t)(n - 1) * w, w, cudaMemcpyDeviceToHost, c.stream));
      PW_CUDA(cudaStreamSynchronize(c.stream));
      t_first = w == 4 ? (int64_t)(int32_t)a : a;
      t_last = w == 4 ? (int64_t)(int32_t)b : b;
    }
    if (t_last < t_first) { dev_free(dctl); return fail(PW_ERR_NOT_SORTED, "argument in operation 'group_by_dynamic' is not sorted, please sort the 'expr/series/column' first"); }
    int64_t rem = t_first % d.every; if (rem < 0) rem += d.every;
    int64_t start = t_first - rem + d.offset;
    auto is_past = [&](int64_t s) { return (d.closed == PW_CLOSED_LEFT || d.closed == PW_CLOSED_BOTH) ? s > t_first : s >= t_first; };
    while (is_past(start)) {
      int64_t gap = start - t_first;
      if (d.closed == PW_CLOSED_RIGHT || d.closed == PW_CLOSED_NONE) gap += 1;
      int64_t stride = (gap + d.every - 1) / d.every;
      if (stride < 1) stride = 1;
      start -= d.every * stride;
    }
    const int64_t boundary_stop = (n > 1 ? t_last : t_first) + 1;
    WinParams w{};
    w.s0 = start; w.every = d.every; w.period = d.period; w.closed = d.closed;
    w.n_windows = start < boundary_stop ? (boundary_stop - start + d.every - 1) / d.every : 0;
    const int64_t k0 = floor_div_h(start - d.offset, d.every);

    uint32_t* counts = nullptr; uint64_t* offsets = nullptr;
    void* p = nullptr;
    PW_TRY(dev_alloc(&p, (size_t)n * 4)); counts = (uint32_t*)p;
    PW_TRY(dev_alloc(&p, (size_t)n * 8)); offsets = (uint64_t*)p;
    const int grid_n = (int)((n + 255) / 256);
    dyn_count_kernel<<<grid_n, 256, 0, c.stream>>>(ts, n, w, counts, &dctl->not_sorted);
    PW_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offsets, n, c.stream);
    void* tmp = nullptr;
    PW_TRY(dev_alloc(&tmp, tmp_bytes));
    PW_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, offsets, n, c.stream));
    tm.kernel_launches += 2;
    uint32_t last_c = 0; uint64_t last_o = 0;
    PW_CUDA(cudaMemcpyAsync(&last_c, counts + (n - 1), 4, cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaMemcpyAsync(&last_o, offsets + (n - 1), 8, cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaMemcpyAsync(&hctl, dctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaStreamSynchronize(c.stream));
    dev_free(tmp);
    if (hctl.not_sorted) {
      dev_free(counts); dev_free(offsets); dev_free(dctl);
      return fail(PW_ERR_NOT_SORTED, "argument in operation 'group_by_dynamic' is not sorted, please sort the 'expr/series/column' first");
    }
    n_win = last_o + last_c;
    if (n_win > 0xFFFFFFF0ull) { dev_free(counts); dev_free(offsets); dev_free(dctl); return fail(PW_ERR_UNSUPPORTED, "more than 2^32 windows"); }
    PW_TRY(dev_alloc(&p, std::max<uint64_t>(n_win, 1) * 8)); win_index = (int64_t*)p;
    PW_TRY(dev_alloc(&p, std::max<uint64_t>(n_win, 1) * 8)); win_start = (int64_t*)p;
    PW_TRY(dev_alloc(&p, std::max<uint64_t>(n_win, 1) * 8)); win_end = (int64_t*)p;
    // the result "table": slot x = x-th non-empty window; key word 0 = window index on the label grid
    T.cap = n_win;
    PW_TRY(dev_alloc(&p, (n_win + 2) * 8 * 1)); T.keys = (uint64_t*)p;
    PW_TRY(dev_alloc(&p, (n_win + 2) * 4)); T.state = (uint32_t*)p;
    PW_TRY(dev_alloc(&p, (n_win + 2) * 8 * (uint64_t)P.n_acc)); T.accs = (uint64_t*)p;
    T.overflow = &dctl->not_sorted; T.spilled = &dctl->total;
    T.key_sw = T.acc_sw = n_win + 2; T.key_ss = T.acc_ss = 1;  // dense, sequentially written: struct of arrays
    if (n_win) {
      dyn_emit_kernel<<<grid_n, 256, 0, c.stream>>>(ts, n, w, counts, offsets, win_index, win_start);
      PW_CUDA(cudaGetLastError());
      const int grid_w = (int)((n_win + 255) / 256);
      dyn_end_kernel<<<grid_w, 256, 0, c.stream>>>(ts, n, w, (int64_t)n_win, win_index, win_start, win_end, T.keys, k0);
      PW_CUDA(cudaGetLastError());
      P.table = T;
      const size_t smem = (size_t)8 * P.n_acc * 32 * 8;
      int64_t grid_s = std::min<int64_t>(((int64_t)n_win + 7) / 8, (int64_t)c.sm_count * 8);
      PW_CUDA(cudaEventRecord(c.ev[8], c.stream));
      slice_agg_kernel<12><<<(unsigned)grid_s, 256, smem, c.stream>>>(P, (int64_t)n_win, win_start, win_end);
      PW_CUDA(cudaGetLastError());
      PW_CUDA(cudaEventRecord(c.ev[9], c.stream));
      tm.kernel_launches += 3;
    }
    dev_free(counts); dev_free(offsets);
  } else {
    void* p = nullptr;
    T.cap = 0;
    PW_TRY(dev_alloc(&p, 16)); T.keys = (uint64_t*)p;
    PW_TRY(dev_alloc(&p, 8)); T.state = (uint32_t*)p;
    PW_TRY(dev_alloc(&p, 16 * (uint64_t)std::max(1, P.n_acc))); T.accs = (uint64_t*)p;
    T.key_sw = T.acc_sw = 2; T.key_ss = T.acc_ss = 1;
  }
  { void* p = nullptr; PW_TRY(dev_alloc(&p, std::max<uint64_t>(n_win, 1) * 4)); slots = (uint32_t*)p; }
  if (n_win) {
    iota_kernel<<<(int)((n_win + 255) / 256), 256, 0, c.stream>>>(slots, n_win);
    PW_CUDA(cudaGetLastError());
    tm.kernel_launches++;
  }
  tm.n_groups = (int64_t)n_win; tm.table_slots = (int64_t)n_win;
  PW_CUDA(cudaEventRecord(c.ev[3], c.stream));
  // key word 0 (the window index) is not a nullable key: emit through the plain table path with n_kw = 2 so that
  // slot occupancy is never consulted via the single-word sentinel rule
  int rc = emit_results(L, T, slots, n_win, out_cols, out_schemas, n_out);
  dev_free(slots); dev_free(win_index); dev_free(win_start); dev_free(win_end);
  free_table(T);
  dev_free(dctl);
  if (rc) return rc;
  float ms;
  if (cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]) == cudaSuccess) tm.scan_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]) == cudaSuccess) tm.finalize_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[4], c.ev[5]) == cudaSuccess) tm.d2h_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[0], c.ev[5]) == cudaSuccess) tm.total_device_ms = ms;
  if (n_win && cudaEventElapsedTime(&ms, c.ev[8], c.ev[9]) == cudaSuccess) tm.scan_kernel_ms = ms;
  return 0;
}

}  // namespace pw
