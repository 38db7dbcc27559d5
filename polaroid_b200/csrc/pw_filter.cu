// pw_filter.cu — the operators either side of the fused path:
//   * predicate -> compacted selection vector (ascending u32 row ids) and FilterExec-style column compaction
//     (polars-mem-engine/src/executors/filter.rs:93-121, polars-compute/src/filter/mod.rs:18-110).  The fused
//     group-by never materialises these; they exist for a bare `lf.filter(p).collect()` and for callers that
//     want the selection vector.
//   * GroupsIdx construction (polars-core/src/frame/group_by/hashing.rs:75-231, position.rs:16-20): first row,
//     offsets and all row ids per group, ascending inside a group.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "pw_engine.h"
#include "pw_scan.cuh"

namespace pw {

constexpr int SEL_THREADS = 256;
constexpr int SEL_ROWS = 8;  // rows per thread, consecutive

struct PredPlan {
  int32_t n_preds;
  int32_t pad;
  RawSlot slot[MAX_PREDS];
  Pred pred[MAX_PREDS];
};

__device__ __forceinline__ bool eval_preds(const PredPlan& pp, int64_t row) {
  bool alive = true;
  for (int q = 0; q < pp.n_preds; ++q) {
    const RawSlot& s = pp.slot[q];
    const bool ok = load_valid_pair(s, row, row + 1) & 1u;
    const uint64_t v = decode(load_row(s, row), s.dtype, 0);
    alive = alive && ok && compare(v, pp.pred[q].scalar, pp.pred[q].cls, pp.pred[q].op);
  }
  return alive;
}

// pass 0: selected rows per CTA tile; pass 1: write ascending ids at tile_base + in-tile rank
static __global__ void __launch_bounds__(SEL_THREADS) select_kernel(PredPlan pp, int64_t n, unsigned long long* tile_counts,
                                                                    const unsigned long long* tile_base, uint32_t* ids, int pass) {
  __shared__ uint32_t warp_tot[SEL_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile0 = (int64_t)blockIdx.x * SEL_THREADS * SEL_ROWS;
  const int64_t base = tile0 + (int64_t)threadIdx.x * SEL_ROWS;
  uint32_t mask = 0;
#pragma unroll
  for (int r = 0; r < SEL_ROWS; ++r)
    if (base + r < n && eval_preds(pp, base + r)) mask |= 1u << r;
  const uint32_t mine = __popc(mask);
  uint32_t incl = mine;  // inclusive scan inside the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t before = 0, total = 0;
  for (int w = 0; w < SEL_THREADS / 32; ++w) { if (w < warp) before += warp_tot[w]; total += warp_tot[w]; }
  if (pass == 0) { if (threadIdx.x == 0) tile_counts[blockIdx.x] = total; return; }
  unsigned long long pos = tile_base[blockIdx.x] + before + incl - mine;
#pragma unroll
  for (int r = 0; r < SEL_ROWS; ++r)
    if (mask & (1u << r)) ids[pos++] = (uint32_t)(base + r);
}

// gather one column by selection vector; validity is ballot-packed
static __global__ void gather_kernel(RawSlot src, int width, const uint32_t* ids, uint64_t n, unsigned char* out, uint32_t* out_valid,
                                     unsigned long long* null_count, int32_t* long_view) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = true;
  if (i < n) {
    const int64_t row = ids[i];
    const unsigned char* p = (const unsigned char*)src.values + (size_t)row * width;
    switch (width) {
      case 1: out[i] = *p; break;
      case 2: ((uint16_t*)out)[i] = *(const uint16_t*)p; break;
      case 4: ((uint32_t*)out)[i] = *(const uint32_t*)p; break;
      case 8: ((uint64_t*)out)[i] = *(const uint64_t*)p; break;
      default: {
        const uint4 v = *(const uint4*)p;
        ((uint4*)out)[i] = v;
        if (v.x > 12u) *long_view = 1;
        break; }
    }
    valid = load_valid_pair(src, row, row + 1) & 1u;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, valid && i < n);
  const uint32_t in = __ballot_sync(0xffffffffu, i < n);
  if ((threadIdx.x & 31) == 0 && in) {
    out_valid[i >> 5] = m;
    const int nulls = __popc(in & ~m);
    if (nulls) atomicAdd(null_count, (unsigned long long)nulls);
  }
}

// rank of every slot in the ordered group list + group sizes in rank order (sizes[G] = 0 closes the scan)
static __global__ void rank_kernel(const uint32_t* slots, uint32_t* slot_rank, uint64_t* sizes, Table T, int acc_len, uint64_t G) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G) { const uint32_t s = slots[i]; slot_rank[s] = (uint32_t)i; sizes[i] = tacc(T, acc_len, s); }
  else if (i == G) sizes[i] = 0;
}
static __global__ void first_kernel(const uint32_t* row_ids_sorted, const uint64_t* offsets, uint32_t* first, uint64_t G) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G) first[i] = row_ids_sorted[offsets[i]];
}

static int width_of(int dt) {
  switch (dt) {
    case DT_I8: case DT_U8: return 1;
    case DT_I16: case DT_U16: return 2;
    case DT_I32: case DT_U32: case DT_F32: return 4;
    case DT_VIEW: return 16;
    default: return 8;
  }
}

static int build_pred_plan(const PwPredicate* preds, int32_t n_preds, const PwFrame* f, PredPlan* pp) {
  memset(pp, 0, sizeof *pp);
  if (n_preds < 0 || n_preds > MAX_PREDS) return fail(PW_ERR_UNSUPPORTED, "more than %d predicate conjuncts", MAX_PREDS);
  pp->n_preds = n_preds;
  for (int i = 0; i < n_preds; ++i) {
    const PwPredicate& p = preds[i];
    if (p.column < 0 || p.column >= (int)f->cols.size()) return fail(PW_ERR_INVALID, "predicate column %d out of range", p.column);
    const FrameColumn& c = f->cols[p.column];
    if (c.dtype == DT_VIEW || c.dtype == DT_BOOL) return fail(PW_ERR_UNSUPPORTED, "predicate on column '%s' of format %s", c.name.c_str(), c.format.c_str());
    pp->slot[i].values = c.values; pp->slot[i].validity = c.null_count ? c.validity : nullptr;
    pp->slot[i].dtype = c.dtype; pp->slot[i].bit_offset = c.bit_offset;
    pp->pred[i].op = p.op;
    const int cls = (c.dtype == DT_F32 || c.dtype == DT_F64) ? CLS_F64 : ((c.dtype == DT_U8 || c.dtype == DT_U16 || c.dtype == DT_U32 || c.dtype == DT_U64) ? CLS_U64 : CLS_I64);
    pp->pred[i].cls = cls;
    if (cls == CLS_F64) { double v = p.scalar_is_float ? p.scalar.f : (double)p.scalar.i; memcpy(&pp->pred[i].scalar, &v, 8); }
    else {
      if (p.scalar_is_float) return fail(PW_ERR_UNSUPPORTED, "float scalar compared with integer column '%s'", c.name.c_str());
      pp->pred[i].scalar = p.scalar.u;
    }
  }
  return 0;
}

// device selection vector; caller frees *ids_out
static int select_rows(const PredPlan& pp, int64_t n, uint32_t** ids_out, int64_t* n_sel) {
  ThreadCtx& c = ctx();
  // row ids are IdxSize = u32 (polars-utils/src/index.rs:9-11)
  if (n > 0xFFFFFFF0ll) return fail(PW_ERR_UNSUPPORTED, "IdxSize is u32: more than 2^32 rows in a filter selection");
  const int64_t tile = (int64_t)SEL_THREADS * SEL_ROWS;
  const int64_t n_tiles = std::max<int64_t>(1, (n + tile - 1) / tile);
  unsigned long long *counts = nullptr, *base = nullptr;
  void* v = nullptr;
  PW_TRY(dev_alloc(&v, (size_t)n_tiles * 8)); counts = (unsigned long long*)v;
  PW_TRY(dev_alloc(&v, (size_t)n_tiles * 8)); base = (unsigned long long*)v;
  select_kernel<<<(unsigned)n_tiles, SEL_THREADS, 0, c.stream>>>(pp, n, counts, base, nullptr, 0);
  PW_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, base, (int)n_tiles, c.stream);
  void* tmp = nullptr;
  PW_TRY(dev_alloc(&tmp, tmp_bytes));
  PW_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, base, (int)n_tiles, c.stream));
  unsigned long long last_c = 0, last_b = 0;
  PW_CUDA(cudaMemcpyAsync(&last_c, counts + (n_tiles - 1), 8, cudaMemcpyDeviceToHost, c.stream));
  PW_CUDA(cudaMemcpyAsync(&last_b, base + (n_tiles - 1), 8, cudaMemcpyDeviceToHost, c.stream));
  PW_CUDA(cudaStreamSynchronize(c.stream));
  *n_sel = (int64_t)(last_c + last_b);
  PW_TRY(dev_alloc(&v, (size_t)std::max<int64_t>(*n_sel, 1) * 4)); *ids_out = (uint32_t*)v;
  select_kernel<<<(unsigned)n_tiles, SEL_THREADS, 0, c.stream>>>(pp, n, counts, base, *ids_out, 1);
  PW_CUDA(cudaGetLastError());
  c.timings.kernel_launches += 3;
  dev_free(tmp); dev_free(counts); dev_free(base);
  return 0;
}

static int device_to_arrow(const void* d_vals, size_t val_bytes, const void* d_valid, size_t valid_bytes, int64_t n, int64_t nulls,
                           bool is_view, const char* format, const char* name, struct ArrowArray* out, struct ArrowSchema* schema) {
  ThreadCtx& c = ctx();
  void* hv = host_alloc(val_bytes);
  void* hb = host_alloc(valid_bytes);
  if (!hv || !hb) return fail(PW_ERR_INTERNAL, "out of host memory");
  if (val_bytes) PW_CUDA(cudaMemcpyAsync(hv, d_vals, val_bytes, cudaMemcpyDeviceToHost, c.stream));
  if (valid_bytes && d_valid) PW_CUDA(cudaMemcpyAsync(hb, d_valid, valid_bytes, cudaMemcpyDeviceToHost, c.stream));
  PW_CUDA(cudaStreamSynchronize(c.stream));
  PW_TRY(make_host_array(n, nulls, hb, hv, is_view ? 1 : 0, out));
  return make_schema(format, name, true, schema);
}

}  // namespace pw

using namespace pw;

// ---- GroupsSlice of sorted keys: run boundaries (partition_to_groups, sort_partition.rs:168) --------------------------
// head[i] = row i starts a run: i == 0 or its key (values and null-ness of every key column) differs from row i - 1
static __device__ __forceinline__ void keys_of_row(const ScanPlan& P, int64_t row, uint64_t (&k)[6], bool& sf) {
  uint4 raw[12];
  uint32_t vbits[12];
  Row<12> r;
  r.in_valid = 0;
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    raw[c] = make_uint4(0u, 0u, 0u, 0u); vbits[c] = 0u; r.in[c] = 0;
    if (c < P.n_slots) {
      raw[c] = load_row(P.slots[c], row);
      vbits[c] = load_valid_pair(P.slots[c], row, row + 1) & 1u;
      r.in[c] = decode(raw[c], P.slots[c].dtype, 0);
      r.in_valid |= vbits[c] << c;
    }
  }
  row_keys<RtCtl, 12, 6>(P, r, raw, vbits, 0, true, k, sf);
}
static __global__ void __launch_bounds__(256) run_heads_kernel(const __grid_constant__ ScanPlan P, unsigned char* head) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    bool h = i == 0;
    if (!h) {
      uint64_t a[6], b[6];
      bool sa, sb;
      keys_of_row(P, i, a, sa);
      keys_of_row(P, i - 1, b, sb);
      h = sa != sb;
#pragma unroll
      for (int w = 0; w < 6; ++w) h = h || a[w] != b[w];
    }
    head[i] = h ? 1 : 0;
  }
}
static __global__ void run_lens_kernel(const int64_t* first64, const unsigned long long* n_runs, int64_t n_rows, uint32_t* first, uint32_t* len, uint64_t cap) {
  const uint64_t G = *n_runs;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G && g < cap; g += (uint64_t)gridDim.x * blockDim.x) {
    const int64_t a = first64[g], b = g + 1 < G ? first64[g + 1] : n_rows;
    first[g] = (uint32_t)a;
    len[g] = (uint32_t)(b - a);
  }
}

extern "C" {

int pw_b200_frame_filter_select(const PwPredicate* predicates, int32_t n_predicates, const PwFrame* frame,
                                struct ArrowArray* out_ids, struct ArrowSchema* out_schema, int64_t* n_selected) {
  PW_TRY(ensure_device());
  if (!frame || !out_ids || !out_schema || !n_selected) return fail(PW_ERR_INVALID, "null argument");
  PredPlan pp;
  PW_TRY(build_pred_plan(predicates, n_predicates, frame, &pp));
  uint32_t* ids = nullptr;
  PW_TRY(select_rows(pp, frame->n_rows, &ids, n_selected));
  int rc = device_to_arrow(ids, (size_t)*n_selected * 4, nullptr, 0, *n_selected, 0, false, "I", "row_id", out_ids, out_schema);
  dev_free(ids);
  return rc;
}

int pw_b200_filter(const PwPredicate* predicates, int32_t n_predicates, const struct ArrowArray* const* cols,
                   const struct ArrowSchema* const* schemas, size_t n_cols, struct ArrowArray* out_cols, struct ArrowSchema* out_schemas) {
  PwFrame* f = nullptr;
  PW_TRY(pw_b200_frame_upload(cols, schemas, n_cols, &f));
  ThreadCtx& c = ctx();
  PredPlan pp;
  int rc = build_pred_plan(predicates, n_predicates, f, &pp);
  uint32_t* ids = nullptr;
  int64_t n_sel = 0;
  if (!rc) rc = select_rows(pp, f->n_rows, &ids, &n_sel);
  unsigned long long* d_nulls = nullptr;
  int32_t* d_long = nullptr;
  void* v = nullptr;
  if (!rc) rc = dev_alloc(&v, 8 * (n_cols + 1));
  if (!rc) { d_nulls = (unsigned long long*)v; cudaMemsetAsync(d_nulls, 0, 8 * (n_cols + 1), c.stream); d_long = (int32_t*)(d_nulls + n_cols); }
  for (size_t i = 0; i < n_cols && !rc; ++i) {
    const FrameColumn& col = f->cols[i];
    const int w = width_of(col.dtype);
    RawSlot s{col.values, col.null_count ? col.validity : nullptr, col.dtype, col.bit_offset};
    void *dv = nullptr, *db = nullptr;
    const size_t vb = (size_t)n_sel * w, bb = (size_t)((n_sel + 31) / 32) * 4;
    rc = dev_alloc(&dv, vb);
    if (!rc) rc = dev_alloc(&db, bb);
    if (!rc && n_sel) {
      cudaMemsetAsync(d_long, 0, 4, c.stream);   // per column
      gather_kernel<<<(unsigned)((n_sel + 255) / 256), 256, 0, c.stream>>>(s, w, ids, (uint64_t)n_sel, (unsigned char*)dv, (uint32_t*)db, d_nulls + i, d_long);
      if (cudaGetLastError() != cudaSuccess) rc = fail(PW_ERR_CUDA, "gather launch failed");
      c.timings.kernel_launches++;
    }
    unsigned long long nulls = 0;
    int32_t is_long = 0;
    if (!rc) {
      cudaMemcpyAsync(&nulls, d_nulls + i, 8, cudaMemcpyDeviceToHost, c.stream);
      cudaMemcpyAsync(&is_long, d_long, 4, cudaMemcpyDeviceToHost, c.stream);
      cudaStreamSynchronize(c.stream);
      if (is_long && col.dtype == DT_VIEW && !col.has_long) rc = fail(PW_ERR_INVALID, "column '%s' holds values longer than 12 bytes but no data buffer", col.name.c_str());
    }
    if (!rc && col.dtype == DT_VIEW && col.has_long) {
      // values longer than 12 bytes: the selected views still point into the frame's data buffers — gather their
      // bytes into a data buffer of the result's own (pw_views.cu), as the group-by does for its keys
      void* hv = host_alloc(vb);
      void* hb = host_alloc(bb);
      void* h_data = nullptr;
      int64_t data_bytes = 0;
      if (!hv || !hb) rc = fail(PW_ERR_INTERNAL, "out of host memory");
      if (!rc && vb && cudaMemcpyAsync(hv, dv, vb, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess) rc = fail(PW_ERR_CUDA, "copy failed");
      if (!rc && bb && cudaMemcpyAsync(hb, db, bb, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess) rc = fail(PW_ERR_CUDA, "copy failed");
      if (!rc && cudaStreamSynchronize(c.stream) != cudaSuccess) rc = fail(PW_ERR_CUDA, "copy failed");
      if (!rc) rc = views_gather_long(&col, hv, nulls ? hb : nullptr, (uint64_t)n_sel, &h_data, &data_bytes);
      if (!rc) rc = h_data ? make_host_view_array(n_sel, (int64_t)nulls, hb, hv, h_data, data_bytes, &out_cols[i])
                           : make_host_array(n_sel, (int64_t)nulls, hb, hv, 1, &out_cols[i]);
      if (rc) { host_free(hv); host_free(hb); host_free(h_data); }
      else rc = make_schema(col.format.c_str(), col.name.c_str(), true, &out_schemas[i]);
    } else
    if (!rc) rc = device_to_arrow(dv, vb, db, bb, n_sel, (int64_t)nulls, col.dtype == DT_VIEW, col.format.c_str(), col.name.c_str(), &out_cols[i], &out_schemas[i]);
    dev_free(dv); dev_free(db);
    if (rc) for (size_t k = 0; k < i; ++k) { if (out_cols[k].release) out_cols[k].release(&out_cols[k]); if (out_schemas[k].release) out_schemas[k].release(&out_schemas[k]); }
  }
  dev_free(ids); dev_free(d_nulls);
  pw_b200_frame_free(f);
  return rc;
}

int pw_b200_frame_group_tuples(const PwFrame* frame, const int32_t* key_columns, int32_t n_keys, int32_t maintain_order,
                               struct ArrowArray* out_first, struct ArrowArray* out_offsets, struct ArrowArray* out_row_ids,
                               struct ArrowSchema* out_schemas) {
  PW_TRY(ensure_device());
  if (!frame || !out_first || !out_offsets || !out_row_ids || !out_schemas) return fail(PW_ERR_INVALID, "null argument");
  ThreadCtx& c = ctx();
  const int64_t n = frame->n_rows;
  if (n > 0xFFFFFFF0ll) return fail(PW_ERR_UNSUPPORTED, "IdxSize is u32: more than 2^32 rows");
  // pass 1: the ordinary group-by with a `len` word (group sizes) and the first-occurrence word
  PwAgg len_agg{};
  len_agg.kind = PW_LEN; len_agg.column = -1; len_agg.name = "len";
  PwQuery q{};
  q.abi_version = PW_ABI_VERSION; q.maintain_order = maintain_order ? 1 : 0; q.n_keys = n_keys; q.key_columns = key_columns;
  q.n_aggs = 1; q.aggs = &len_agg; q.flags = PW_FLAG_FORCE_GLOBAL_TABLE;
  Lowered L;
  PW_TRY(lower_query(&q, frame, &L));
  Table T{};
  uint32_t* slots = nullptr;
  uint64_t G = 0;
  PW_TRY(run_groupby(&q, frame, L, &T, &slots, &G));
  const uint64_t nn = T.cap + 2;
  // rank of every slot in the (ordered) group list, sizes in rank order -> offsets
  uint32_t* slot_rank = nullptr; uint64_t *sizes = nullptr, *offsets = nullptr;
  void* v = nullptr;
  PW_TRY(dev_alloc(&v, nn * 4)); slot_rank = (uint32_t*)v;
  PW_TRY(dev_alloc(&v, (G + 1) * 8)); sizes = (uint64_t*)v;
  PW_TRY(dev_alloc(&v, (G + 1) * 8)); offsets = (uint64_t*)v;
  int acc_len = -1;
  for (int a = 0; a < L.plan.n_acc; ++a) if (L.plan.accs[a].src == SRC_ONE) acc_len = a;
  rank_kernel<<<(unsigned)std::max<uint64_t>(1, (G + 1 + 255) / 256), 256, 0, c.stream>>>(slots, slot_rank, sizes, T, acc_len, G);
  PW_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, sizes, offsets, (int)(G + 1), c.stream);
  void* tmp = nullptr;
  PW_TRY(dev_alloc(&tmp, tmp_bytes));
  PW_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, sizes, offsets, (int)(G + 1), c.stream));
  dev_free(tmp);
  // pass 2: every row looks its group up -> rank; stable radix sort of (rank, row) gives ascending ids per group
  uint32_t *row_rank = nullptr, *row_rank_sorted = nullptr, *row_ids = nullptr, *row_ids_sorted = nullptr;
  PW_TRY(dev_alloc(&v, (size_t)std::max<int64_t>(n, 1) * 4)); row_rank = (uint32_t*)v;
  PW_TRY(dev_alloc(&v, (size_t)std::max<int64_t>(n, 1) * 4)); row_rank_sorted = (uint32_t*)v;
  PW_TRY(dev_alloc(&v, (size_t)std::max<int64_t>(n, 1) * 4)); row_ids = (uint32_t*)v;
  PW_TRY(dev_alloc(&v, (size_t)std::max<int64_t>(n, 1) * 4)); row_ids_sorted = (uint32_t*)v;
  if (n) {
    ScanPlan P = L.plan;
    P.table = T; P.hot_slots = 0; P.row_group_out = row_rank; P.slot_rank = slot_rank;
    int32_t* d_flag = nullptr;
    PW_TRY(dev_alloc(&v, 64)); d_flag = (int32_t*)v;
    PW_CUDA(cudaMemsetAsync(d_flag, 0, 64, c.stream));
    P.not_sorted = d_flag; P.table.overflow = d_flag + 1; P.table.spilled = (unsigned long long*)(d_flag + 2);
    PW_TRY(launch_scan_aot(P, c.sm_count, c.stream));
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(row_ids, (uint64_t)n);
    PW_CUDA(cudaGetLastError());
    int bits = 1;
    while ((1ull << bits) < G + 1 && bits < 32) ++bits;
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, row_rank, row_rank_sorted, row_ids, row_ids_sorted, (int64_t)n, 0, bits, c.stream);
    void* st = nullptr;
    PW_TRY(dev_alloc(&st, sb));
    PW_CUDA(cub::DeviceRadixSort::SortPairs(st, sb, row_rank, row_rank_sorted, row_ids, row_ids_sorted, (int64_t)n, 0, bits, c.stream));
    dev_free(st); dev_free(d_flag);
    c.timings.kernel_launches += 4;
  }
  // first[g] = row_ids_sorted[offsets[g]]
  uint32_t* first = nullptr;
  PW_TRY(dev_alloc(&v, std::max<uint64_t>(G, 1) * 4)); first = (uint32_t*)v;
  if (G) {
    first_kernel<<<(unsigned)((G + 255) / 256), 256, 0, c.stream>>>(row_ids_sorted, offsets, first, G);
    PW_CUDA(cudaGetLastError());
  }
  int rc = device_to_arrow(first, (size_t)G * 4, nullptr, 0, (int64_t)G, 0, false, "I", "first", out_first, &out_schemas[0]);
  if (!rc) rc = device_to_arrow(offsets, (size_t)(G + 1) * 8, nullptr, 0, (int64_t)G + 1, 0, false, "L", "offsets", out_offsets, &out_schemas[1]);
  if (!rc) rc = device_to_arrow(row_ids_sorted, (size_t)n * 4, nullptr, 0, n, 0, false, "I", "row_ids", out_row_ids, &out_schemas[2]);
  dev_free(first); dev_free(row_rank); dev_free(row_rank_sorted); dev_free(row_ids); dev_free(row_ids_sorted);
  dev_free(slot_rank); dev_free(sizes); dev_free(offsets); dev_free(slots);
  free_table(T);
  return rc;
}

}  // extern "C"  (the sortedness check below is internal)

namespace pw {
namespace {
__device__ __forceinline__ int64_t index_at(const void* values, int w, int64_t row) {
  return w == 4 ? (int64_t)((const int32_t*)values)[row] : ((const int64_t*)values)[row];
}
static __global__ void index_sorted_kernel(const void* values, int w, int64_t n, int32_t* flag) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bad = bad || index_at(values, w, i) < index_at(values, w, i - 1);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *flag = 1;
}
static __global__ void gather_u32_kernel(const uint32_t* src, const uint32_t* ids, int64_t n, uint32_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[ids[i]];
}
// (group, row) pairs in group-major, row-ascending order: a neighbour of the same group with a smaller index value
static __global__ void sorted_within_groups_kernel(const uint32_t* group, const uint32_t* rows, int64_t n, const void* values, int w, int32_t* flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (i < n && group[i] == group[i - 1] && index_at(values, w, rows[i]) < index_at(values, w, rows[i - 1])) *flag = 1;
}
}  // namespace

// group_by_dynamic with `group_by=` keys: the reference sorts the frame by the keys (stable) and raises when the index
// is not ascending inside a key slice (polars-time/src/group_by/dynamic.rs:77-80, 327).  Neither the hash path nor the
// windowed bucket tier needs that order, so the check is its own step:
//   1. is the index column ascending over ALL rows?  (one pass, remembered on the frame) — then it is inside every key;
//   2. otherwise the exact check: every row's group (the group-by on the keys alone, as pw_b200_frame_group_tuples
//      does), stable radix sort of (group, row) over the rows that pass the predicate, neighbours compared.  The verdict
//      for the unfiltered frame is remembered per (keys, index) — a subsequence of an ascending sequence is ascending.
// Returns 0 sorted, PW_ERR_NOT_SORTED, or another error.
int check_sorted_within_keys(const PwQuery* q, const PwFrame* frame) {
  ThreadCtx& c = ctx();
  const int64_t n = frame->n_rows;
  const int icol = q->dynamic->index_column;
  if (n < 2 || icol < 0 || icol >= (int)frame->cols.size()) return 0;
  const FrameColumn& ic = frame->cols[icol];
  if (ic.dtype != DT_I32 && ic.dtype != DT_I64) return 0;   // lower_query refuses the query with the reference's message
  const int w = ic.dtype == DT_I32 ? 4 : 8;
  auto not_sorted = [&]() { return fail(PW_ERR_NOT_SORTED, "argument in operation 'group_by_dynamic' is not sorted, please sort the 'expr/series/column' first"); };
  std::string gkey = "g:" + std::to_string(icol), kkey = "k:" + std::to_string(icol);
  for (int i = 0; i < q->n_keys; ++i) kkey += ":" + std::to_string(q->key_columns[i]);
  int global_state = 0, keyed_state = 0;   // 0 unknown, 1 sorted, 2 not sorted
  {
    std::lock_guard<std::mutex> lk(frame->mu);
    auto it = frame->sorted_cache.find(gkey);
    if (it != frame->sorted_cache.end()) global_state = it->second;
    it = frame->sorted_cache.find(kkey);
    if (it != frame->sorted_cache.end()) keyed_state = it->second;
  }
  if (global_state == 1 || keyed_state == 1) return 0;
  if (keyed_state == 2 && q->n_predicates == 0) return not_sorted();
  void* v = nullptr;
  int32_t* d_flag = nullptr;
  PW_TRY(dev_alloc(&v, 64)); d_flag = (int32_t*)v;
  auto read_flag = [&](int32_t* out) -> int {
    if (cudaMemcpyAsync(out, d_flag, 4, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess || cudaStreamSynchronize(c.stream) != cudaSuccess)
      return fail(PW_ERR_CUDA, "sortedness check failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
  };
  if (global_state == 0) {
    int32_t h = 0;
    int rc = cudaMemsetAsync(d_flag, 0, 64, c.stream) == cudaSuccess ? 0 : fail(PW_ERR_CUDA, "memset failed");
    if (!rc) {
      index_sorted_kernel<<<148 * 8, 256, 0, c.stream>>>(ic.values, w, n, d_flag);
      c.timings.kernel_launches++;
      rc = read_flag(&h);
    }
    if (rc) { dev_free(d_flag); return rc; }
    global_state = h ? 2 : 1;
    std::lock_guard<std::mutex> lk(frame->mu);
    if (frame->sorted_cache.size() > 256) frame->sorted_cache.clear();
    frame->sorted_cache[gkey] = global_state;
    if (global_state == 1) { dev_free(d_flag); return 0; }
  }
  // ---- exact check per key
  if (n > 0xFFFFFFF0ll) { dev_free(d_flag); return fail(PW_ERR_UNSUPPORTED, "sortedness check inside keys: more than 2^32 rows"); }
  PwAgg len_agg{};
  len_agg.kind = PW_LEN; len_agg.column = -1; len_agg.name = "len";
  PwQuery gq{};
  gq.abi_version = PW_ABI_VERSION; gq.n_keys = q->n_keys; gq.key_columns = q->key_columns;
  gq.n_aggs = 1; gq.aggs = &len_agg; gq.flags = PW_FLAG_FORCE_GLOBAL_TABLE;
  Lowered L;
  Table T{};
  uint32_t *slots = nullptr, *slot_rank = nullptr, *row_rank = nullptr, *ids = nullptr, *keys_in = nullptr, *keys_out = nullptr, *ids_out = nullptr;
  uint64_t* sizes = nullptr;
  void* tmp = nullptr;
  uint64_t G = 0;
  int64_t n_sel = n;
  int32_t h = 0;
  auto run = [&]() -> int {
    PW_TRY(lower_query(&gq, frame, &L));
    PW_TRY(run_groupby(&gq, frame, L, &T, &slots, &G));
    const uint64_t nn = T.cap + 2;
    PW_TRY(dev_alloc(&v, nn * 4)); slot_rank = (uint32_t*)v;
    PW_TRY(dev_alloc(&v, (G + 1) * 8)); sizes = (uint64_t*)v;
    int acc_len = -1;
    for (int a = 0; a < L.plan.n_acc; ++a) if (L.plan.accs[a].src == SRC_ONE) acc_len = a;
    rank_kernel<<<(unsigned)std::max<uint64_t>(1, (G + 1 + 255) / 256), 256, 0, c.stream>>>(slots, slot_rank, sizes, T, acc_len, G);
    PW_CUDA(cudaGetLastError());
    PW_TRY(dev_alloc(&v, (size_t)n * 4)); row_rank = (uint32_t*)v;
    ScanPlan P = L.plan;
    P.table = T; P.hot_slots = 0; P.row_group_out = row_rank; P.slot_rank = slot_rank;
    PW_CUDA(cudaMemsetAsync(d_flag, 0, 64, c.stream));
    P.not_sorted = d_flag + 4; P.table.overflow = d_flag + 5; P.table.spilled = (unsigned long long*)(d_flag + 6);
    PW_TRY(launch_scan_aot(P, c.sm_count, c.stream));
    if (q->n_predicates > 0) {
      PredPlan pp;
      PW_TRY(build_pred_plan(q->predicates, q->n_predicates, frame, &pp));
      PW_TRY(select_rows(pp, n, &ids, &n_sel));
      if (n_sel < 2) return 0;
      PW_TRY(dev_alloc(&v, (size_t)n_sel * 4)); keys_in = (uint32_t*)v;
      gather_u32_kernel<<<(unsigned)((n_sel + 255) / 256), 256, 0, c.stream>>>(row_rank, ids, n_sel, keys_in);
      PW_CUDA(cudaGetLastError());
    } else {
      PW_TRY(dev_alloc(&v, (size_t)n * 4)); ids = (uint32_t*)v;
      iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(ids, (uint64_t)n);
      PW_CUDA(cudaGetLastError());
    }
    const uint32_t* kin = keys_in ? keys_in : row_rank;
    PW_TRY(dev_alloc(&v, (size_t)n_sel * 4)); keys_out = (uint32_t*)v;
    PW_TRY(dev_alloc(&v, (size_t)n_sel * 4)); ids_out = (uint32_t*)v;
    int bits = 1;
    while ((1ull << bits) < G + 1 && bits < 32) ++bits;
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, kin, keys_out, ids, ids_out, n_sel, 0, bits, c.stream);
    PW_TRY(dev_alloc(&tmp, sb));
    PW_CUDA(cub::DeviceRadixSort::SortPairs(tmp, sb, kin, keys_out, ids, ids_out, n_sel, 0, bits, c.stream));
    sorted_within_groups_kernel<<<(unsigned)((n_sel + 255) / 256), 256, 0, c.stream>>>(keys_out, ids_out, n_sel, ic.values, w, d_flag);
    PW_CUDA(cudaGetLastError());
    c.timings.kernel_launches += 5;
    return read_flag(&h);
  };
  const int rc = run();
  dev_free(slots); dev_free(slot_rank); dev_free(sizes); dev_free(row_rank); dev_free(ids); dev_free(keys_in); dev_free(keys_out); dev_free(ids_out);
  dev_free(tmp); dev_free(d_flag);
  if (T.keys) free_table(T);
  if (rc) return rc;
  if (q->n_predicates == 0) {
    std::lock_guard<std::mutex> lk(frame->mu);
    frame->sorted_cache[kkey] = h ? 2 : 1;
  }
  return h ? not_sorted() : 0;
}
}  // namespace pw

extern "C" {
int pw_b200_frame_group_slices(const PwFrame* frame, const int32_t* key_columns, int32_t n_keys, struct ArrowArray* out_first,
                               struct ArrowArray* out_len, struct ArrowSchema* out_schemas) {
  PW_TRY(ensure_device());
  if (!frame || !key_columns || n_keys < 1 || !out_first || !out_len || !out_schemas) return fail(PW_ERR_INVALID, "null argument");
  ThreadCtx& c = ctx();
  const int64_t n = frame->n_rows;
  if (n > 0xFFFFFFF0ll) return fail(PW_ERR_UNSUPPORTED, "IdxSize is u32: more than 2^32 rows");
  PwAgg len_agg{};
  len_agg.kind = PW_LEN; len_agg.column = -1; len_agg.name = "len";
  PwQuery q{};
  q.abi_version = PW_ABI_VERSION; q.maintain_order = 1; q.n_keys = n_keys; q.key_columns = key_columns; q.n_aggs = 1; q.aggs = &len_agg;
  Lowered L;
  PW_TRY(lower_query(&q, frame, &L));
  ScanPlan P = L.plan;
  if (P.n_slots > 12 || P.n_kw > 6) return fail(PW_ERR_UNSUPPORTED, "too many key columns");
  P.n_rows = n;
  void* v = nullptr;
  unsigned char* head = nullptr; int64_t* first64 = nullptr; unsigned long long* d_runs = nullptr; uint32_t *first = nullptr, *len = nullptr;
  PW_TRY(dev_alloc(&v, (size_t)std::max<int64_t>(n, 1))); head = (unsigned char*)v;
  PW_TRY(dev_alloc(&v, (size_t)std::max<int64_t>(n, 1) * 8)); first64 = (int64_t*)v;
  PW_TRY(dev_alloc(&v, 16)); d_runs = (unsigned long long*)v;
  PW_CUDA(cudaMemsetAsync(d_runs, 0, 16, c.stream));
  unsigned long long G = 0;
  if (n > 0) {
    run_heads_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)c.sm_count * 16), 256, 0, c.stream>>>(P, head);
    PW_CUDA(cudaGetLastError());
    cub::CountingInputIterator<int64_t> rows(0);
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, rows, head, first64, d_runs, n, c.stream);
    void* tmp = nullptr;
    PW_TRY(dev_alloc(&tmp, tb));
    PW_CUDA(cub::DeviceSelect::Flagged(tmp, tb, rows, head, first64, d_runs, n, c.stream));
    PW_CUDA(cudaMemcpyAsync(&G, d_runs, 8, cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaStreamSynchronize(c.stream));
    dev_free(tmp);
    c.timings.kernel_launches += 2;
  }
  PW_TRY(dev_alloc(&v, std::max<uint64_t>(G, 1) * 4)); first = (uint32_t*)v;
  PW_TRY(dev_alloc(&v, std::max<uint64_t>(G, 1) * 4)); len = (uint32_t*)v;
  if (G) {
    run_lens_kernel<<<(unsigned)std::min<uint64_t>((G + 255) / 256, 65535), 256, 0, c.stream>>>(first64, d_runs, n, first, len, G);
    PW_CUDA(cudaGetLastError());
    c.timings.kernel_launches++;
  }
  int rc = device_to_arrow(first, (size_t)G * 4, nullptr, 0, (int64_t)G, 0, false, "I", "first", out_first, &out_schemas[0]);
  if (!rc) rc = device_to_arrow(len, (size_t)G * 4, nullptr, 0, (int64_t)G, 0, false, "I", "len", out_len, &out_schemas[1]);
  dev_free(head); dev_free(first64); dev_free(d_runs); dev_free(first); dev_free(len);
  c.timings.n_rows = n; c.timings.n_groups = (int64_t)G; c.timings.strategy = 8;
  return rc;
}

}  // extern "C"
