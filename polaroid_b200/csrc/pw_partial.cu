// pw_partial.cu — multi-GPU partial aggregates.
//
// Rows are sharded contiguously over the GPUs of one box.  Phase 1 (every GPU): the ordinary fused scan
// builds the local aggregate table; it is exported as fixed-width packed rows
//     [flags][key words][accumulator words][first/last value words]
// counting-sorted by owner = (mix(hash(key)) * n_parts) >> 64 — the same multiply-shift routing as the
// reference's HashPartitioner (polars-utils/src/hashing.rs:100-109).  The host side exchanges the slices with
// one NCCL all-to-all (polaroid_b200/multigpu.py).  Phase 2 (owner): packed rows are merged into a fresh table
// with the accumulators' own associative ops — GroupedReduction::combine semantics
// (polars-expr/src/reduce/mod.rs:94-105): sums add, min/max/first/last keep the better (row-index, value) pair —
// and finalised by the same emit kernels.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "pw_engine.h"
#include "pw_scan.cuh"

struct PwPartial {
  void* rows = nullptr;
  int64_t row_words = 0;
  int64_t n_rows = 0;
  std::vector<int64_t> offsets;  // n_parts + 1
};

namespace pw {

struct PackLayout {
  int32_t kw, n_acc, n_fl;
  int32_t fl_acc[MAX_ACC];  // accumulator index of each first/last word
  RawSlot fl_src[MAX_ACC];  // its source column (phase 1 gather)
  int32_t row_words;        // 1 + kw + n_acc + n_fl
};

static PackLayout make_layout(const Lowered& L, int kw) {
  PackLayout pl{};
  pl.kw = kw; pl.n_acc = L.plan.n_acc;
  for (int a = 0; a < L.plan.n_acc; ++a)
    if (L.plan.accs[a].src == SRC_ROWIDX) {
      pl.fl_acc[pl.n_fl] = a;
      pl.fl_src[pl.n_fl] = L.plan.slots[L.plan.vexprs[L.plan.accs[a].vexpr].slot];
      pl.n_fl++;
    }
  pl.row_words = 1 + pl.kw + pl.n_acc + pl.n_fl;
  return pl;
}

__device__ __forceinline__ uint32_t owner_of(const uint64_t* k, int kw, int n_parts) {
  uint64_t h = mix64(k[0]);
  for (int w = 1; w < kw; ++w) h = mix64(h ^ (k[w] + 0x9E3779B97F4A7C15ull * (uint64_t)w));
  return (uint32_t)__umul64hi(mix64(h ^ 0x5851F42D4C957F2Dull), (uint64_t)n_parts);
}

// pass 1: owner histogram; pass 2: scatter packed rows
static __global__ void export_kernel(Table T, PackLayout pl, const uint32_t* slot_list, uint64_t n, int n_parts, int64_t row_offset,
                                     unsigned long long* part_count, const unsigned long long* part_base, uint64_t* rows, int pass) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t slot = slot_list[i];
  const uint64_t stride = T.cap + 2;
  uint64_t k[MAX_KW];
  for (int w = 0; w < pl.kw; ++w) k[w] = tkey(T, w, slot);
  const uint32_t owner = owner_of(k, pl.kw, n_parts);
  if (pass == 0) { atomicAdd(&part_count[owner], 1ull); return; }
  const unsigned long long pos = part_base[owner] + atomicAdd(&part_count[owner], 1ull);
  uint64_t* r = rows + pos * (uint64_t)pl.row_words;
  r[0] = (pl.kw == 1 && slot >= T.cap) ? 1ull : 0ull;  // escape slot: the key word is raw data that aliases a sentinel
  for (int w = 0; w < pl.kw; ++w) r[1 + w] = k[w];
  for (int a = 0; a < pl.n_acc; ++a) r[1 + pl.kw + a] = tacc(T, a, slot);
  for (int f = 0; f < pl.n_fl; ++f) {
    const uint64_t packed = tacc(T, pl.fl_acc[f], slot);
    uint64_t bits = 0;
    if (packed & 1ull) {
      const int64_t row = (int64_t)(packed >> 1) - row_offset;
      bits = decode(load_row(pl.fl_src[f], row), pl.fl_src[f].dtype, 0);
    }
    r[1 + pl.kw + pl.n_acc + f] = bits;
  }
}

// small-result exchange: rows go straight into the caller's send buffer, unsorted, each tagged with its owner
// ctl (optional): the group count is still on the device (deferred run): n is the capacity of the send buffer, the count
// comes from ctl->counter; a count beyond the capacity — or a scan that overflowed its table — writes the overflow header
static __global__ void export_gather_kernel(Table T, PackLayout pl, const uint32_t* slot_list, uint64_t n, int n_parts, int64_t row_offset,
                                            uint64_t* buf, uint64_t header, const Control* ctl = nullptr) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ctl) {
    const uint64_t cnt = ctl->counter;
    header = (cnt > n || ctl->overflow != 0 || ctl->not_sorted != 0) ? ~0ull : cnt;
    n = cnt < n ? cnt : n;
  }
  if (i == 0) buf[0] = header;
  if (i >= n || header == ~0ull) return;
  const uint64_t slot = slot_list[i];
  uint64_t k[MAX_KW];
  for (int w = 0; w < pl.kw; ++w) k[w] = tkey(T, w, slot);
  uint64_t* r = buf + 1 + i * (uint64_t)(pl.row_words + 1);
  r[0] = owner_of(k, pl.kw, n_parts);
  r[1] = (pl.kw == 1 && slot >= T.cap) ? 1ull : 0ull;
  for (int w = 0; w < pl.kw; ++w) r[2 + w] = k[w];
  for (int a = 0; a < pl.n_acc; ++a) r[2 + pl.kw + a] = tacc(T, a, slot);
  for (int f = 0; f < pl.n_fl; ++f) {
    const uint64_t packed = tacc(T, pl.fl_acc[f], slot);
    uint64_t bits = 0;
    if (packed & 1ull) {
      const int64_t row = (int64_t)(packed >> 1) - row_offset;
      bits = decode(load_row(pl.fl_src[f], row), pl.fl_src[f].dtype, 0);
    }
    r[2 + pl.kw + pl.n_acc + f] = bits;
  }
}

struct AccOpsK { int32_t n; int32_t op[MAX_ACC]; };

// Where the packed rows of a merge live.  Plain: n rows back to back.  Gathered (small-result exchange, one
// all-gather): `world` segments of [n_rows | cap_rows x (owner, packed row)]; a rank merges the rows it owns.  The row
// counts stay on the device (no host round trip between the collective and the merge); a segment whose header is
// GATHER_OVERFLOW (its rank had more than cap_rows groups) raises flag 3 and the caller repeats the exchange through
// the general two-step path.
constexpr uint64_t GATHER_OVERFLOW = ~0ull;
struct RowSrc {
  const uint64_t* base;
  uint64_t seg_words, cap_rows;
  int32_t gathered, my_rank;
};
__device__ __forceinline__ const uint64_t* locate_row(const RowSrc& src, const PackLayout& pl, uint64_t i, int32_t* overflow) {
  if (!src.gathered) return src.base + i * (uint64_t)pl.row_words;
  const uint64_t seg = i / src.cap_rows, j = i % src.cap_rows;
  const uint64_t* sb = src.base + seg * src.seg_words;
  const uint64_t n = sb[0];
  if (n == GATHER_OVERFLOW) { if (j == 0) *overflow = 3; return nullptr; }
  if (j >= n) return nullptr;
  const uint64_t* r = sb + 1 + j * (uint64_t)(pl.row_words + 1);
  return r[0] == (uint64_t)src.my_rank ? r + 1 : nullptr;
}

template <int KW>
static __global__ void merge_kernel(Table T, PackLayout pl, AccOpsK ops, RowSrc src, uint64_t n, uint32_t* row_slot) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t* r = locate_row(src, pl, i, T.overflow);
  if (!r) { row_slot[i] = 0xFFFFFFFFu; return; }
  uint64_t k[KW];
#pragma unroll
  for (int w = 0; w < KW; ++w) k[w] = w < pl.kw ? r[1 + w] : 0ull;
  const uint64_t slot = table_upsert<KW>(T, k, hash_words<KW>(k), r[0] == 0ull);
  row_slot[i] = slot == ~0ull ? 0xFFFFFFFFu : (uint32_t)slot;
  if (slot == ~0ull) return;
  const uint64_t stride = T.cap + 2;
  for (int a = 0; a < pl.n_acc; ++a) {
    const uint64_t v = r[1 + pl.kw + a];
    if (v != acc_init(ops.op[a])) acc_apply_global(&tacc(T, a, slot), ops.op[a], v);
  }
}

// second phase of first/last: the row whose packed index won the merge publishes its value word
static __global__ void merge_values_kernel(Table T, PackLayout pl, RowSrc src, uint64_t n, const uint32_t* row_slot, uint64_t* fl_values) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = row_slot[i];
  if (slot == 0xFFFFFFFFu) return;
  int32_t dummy = 0;
  const uint64_t* r = locate_row(src, pl, i, &dummy);
  if (!r) return;
  const uint64_t stride = T.cap + 2;
  for (int f = 0; f < pl.n_fl; ++f) {
    const uint64_t mine = r[1 + pl.kw + pl.fl_acc[f]];
    if (tacc(T, pl.fl_acc[f], slot) == mine) fl_values[(uint64_t)f * (T.cap + 2) + slot] = r[1 + pl.kw + pl.n_acc + f];
  }
}

static int kw_class(int n_kw) { return n_kw <= 1 ? 1 : (n_kw <= 2 ? 2 : (n_kw <= 4 ? 4 : 6)); }

}  // namespace pw

using namespace pw;

extern "C" {

int pw_b200_frame_groupby_partial(const PwQuery* q, const PwFrame* frame, int32_t n_parts, PwPartial** out) {
  PW_TRY(ensure_device());
  if (!q || !frame || !out || n_parts < 1) return fail(PW_ERR_INVALID, "bad argument");
  ThreadCtx& c = ctx();
  memset(&c.timings, 0, sizeof c.timings);
  PW_CUDA(cudaEventRecord(c.ev[0], c.stream));
  Lowered L;
  PW_TRY(lower_query(q, frame, &L));
  for (int i = 0; i < q->n_keys; ++i)   // the views of a long value point into THIS rank's data buffers
    if (frame->cols[q->key_columns[i]].has_long) return fail(PW_ERR_UNSUPPORTED, "partial aggregates by a string key longer than 12 bytes cannot be exchanged between GPUs");
  if (q->dynamic && !L.tumbling) return fail(PW_ERR_UNSUPPORTED, "overlapping windows are not shardable by row range");
  L.sort.clear();  // ordering happens after the merge
  Table T{};
  uint32_t* slots = nullptr;
  uint64_t G = 0;
  PW_TRY(run_groupby(q, frame, L, &T, &slots, &G));
  const int kw = kw_class(L.plan.n_kw);
  PackLayout pl = make_layout(L, kw);
  PwPartial* p = new PwPartial();
  p->row_words = pl.row_words; p->n_rows = (int64_t)G; p->offsets.assign(n_parts + 1, 0);
  unsigned long long *d_count = nullptr, *d_base = nullptr;
  void* v = nullptr;
  PW_TRY(dev_alloc(&v, (size_t)n_parts * 8)); d_count = (unsigned long long*)v;
  PW_TRY(dev_alloc(&v, (size_t)n_parts * 8)); d_base = (unsigned long long*)v;
  PW_TRY(dev_alloc(&v, std::max<uint64_t>(G, 1) * pl.row_words * 8)); p->rows = v;
  PW_CUDA(cudaMemsetAsync(d_count, 0, (size_t)n_parts * 8, c.stream));
  std::vector<unsigned long long> h_count(n_parts, 0), h_base(n_parts, 0);
  if (G) {
    const int grid = (int)((G + 255) / 256);
    export_kernel<<<grid, 256, 0, c.stream>>>(T, pl, slots, G, n_parts, q->row_offset, d_count, d_base, (uint64_t*)p->rows, 0);
    PW_CUDA(cudaGetLastError());
    PW_CUDA(cudaMemcpyAsync(h_count.data(), d_count, (size_t)n_parts * 8, cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < n_parts; ++i) { p->offsets[i + 1] = p->offsets[i] + (int64_t)h_count[i]; h_base[i] = (unsigned long long)p->offsets[i]; }
    PW_CUDA(cudaMemcpyAsync(d_base, h_base.data(), (size_t)n_parts * 8, cudaMemcpyHostToDevice, c.stream));
    PW_CUDA(cudaMemsetAsync(d_count, 0, (size_t)n_parts * 8, c.stream));
    export_kernel<<<grid, 256, 0, c.stream>>>(T, pl, slots, G, n_parts, q->row_offset, d_count, d_base, (uint64_t*)p->rows, 1);
    PW_CUDA(cudaGetLastError());
    c.timings.kernel_launches += 2;
  }
  PW_CUDA(cudaEventRecord(c.ev[5], c.stream));
  PW_CUDA(cudaStreamSynchronize(c.stream));
  float ms;
  if (cudaEventElapsedTime(&ms, c.ev[0], c.ev[5]) == cudaSuccess) c.timings.total_device_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[8], c.ev[9]) == cudaSuccess) c.timings.scan_kernel_ms = ms;
  dev_free(d_count); dev_free(d_base); dev_free(slots);
  free_table(T);
  *out = p;
  return 0;
}

int64_t pw_b200_partial_row_bytes(const PwPartial* p) { return p ? p->row_words * 8 : -1; }
const void* pw_b200_partial_device_rows(const PwPartial* p) { return p ? p->rows : nullptr; }
int pw_b200_partial_copy_rows(const PwPartial* p, void* dst_device) {
  if (!p || !dst_device) return fail(PW_ERR_INVALID, "null argument");
  if (p->n_rows) PW_CUDA(cudaMemcpyAsync(dst_device, p->rows, (size_t)p->n_rows * p->row_words * 8, cudaMemcpyDeviceToDevice, ctx().stream));
  PW_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}
int pw_b200_partial_offsets(const PwPartial* p, int64_t* part_offsets) {
  if (!p || !part_offsets) return fail(PW_ERR_INVALID, "null argument");
  for (size_t i = 0; i < p->offsets.size(); ++i) part_offsets[i] = p->offsets[i];
  return 0;
}
int pw_b200_partial_free(PwPartial* p) {
  if (!p) return 0;
  dev_free(p->rows);
  delete p;
  return 0;
}

static int merge_impl(const PwQuery* q, const PwFrame* schema_from, RowSrc src, int64_t n_rows, uint64_t cap,
                      struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out);

int pw_b200_merge_partials(const PwQuery* q, const PwFrame* schema_from, const void* device_rows, int64_t n_rows,
                           struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out) {
  PW_TRY(ensure_device());
  if (!q || !schema_from || !out_cols || !out_schemas || !n_out || n_rows < 0) return fail(PW_ERR_INVALID, "bad argument");
  RowSrc src{};
  src.base = (const uint64_t*)device_rows;
  return merge_impl(q, schema_from, src, n_rows, (uint64_t)std::max<int64_t>(2 * n_rows, 64), out_cols, out_schemas, n_out);
}

int64_t pw_b200_partial_row_words(const PwQuery* q, const PwFrame* frame) {
  if (!q || !frame) return fail(PW_ERR_INVALID, "bad argument");
  Lowered L;
  PW_TRY(lower_query(q, frame, &L));
  for (int i = 0; i < q->n_keys; ++i)   // the views of a long value point into THIS rank's data buffers
    if (frame->cols[q->key_columns[i]].has_long) return fail(PW_ERR_UNSUPPORTED, "partial aggregates by a string key longer than 12 bytes cannot be exchanged between GPUs");
  return make_layout(L, kw_class(L.plan.n_kw)).row_words;
}

int pw_b200_frame_groupby_partial_into(const PwQuery* q, const PwFrame* frame, int32_t n_parts, void* send_device, int64_t cap_rows) {
  PW_TRY(ensure_device());
  if (!q || !frame || !send_device || n_parts < 1 || cap_rows < 1) return fail(PW_ERR_INVALID, "bad argument");
  ThreadCtx& c = ctx();
  memset(&c.timings, 0, sizeof c.timings);
  PW_CUDA(cudaEventRecord(c.ev[0], c.stream));
  Lowered L;
  PW_TRY(lower_query(q, frame, &L));
  for (int i = 0; i < q->n_keys; ++i)   // the views of a long value point into THIS rank's data buffers
    if (frame->cols[q->key_columns[i]].has_long) return fail(PW_ERR_UNSUPPORTED, "partial aggregates by a string key longer than 12 bytes cannot be exchanged between GPUs");
  if (q->dynamic && !L.tumbling) return fail(PW_ERR_UNSUPPORTED, "overlapping windows are not shardable by row range");
  L.sort.clear();
  Table T{};
  uint32_t* slots = nullptr;
  uint64_t G = 0;
  RunOpts ro;
  ro.allow_deferred = true; ro.control_only = true;   // small tables: the count never visits the host
  RunState rs;
  PW_TRY(run_groupby(q, frame, L, &T, &slots, &G, &ro, &rs));
  PackLayout pl = make_layout(L, kw_class(L.plan.n_kw));
  bool fits = true;
  if (rs.deferred) {
    const uint64_t n = std::min<uint64_t>(G, (uint64_t)cap_rows);   // G = the table's capacity bound
    export_gather_kernel<<<(int)std::max<uint64_t>(1, (n + 255) / 256), 256, 0, c.stream>>>(T, pl, slots, (uint64_t)cap_rows, n_parts, q->row_offset,
                                                                                            (uint64_t*)send_device, 0, rs.dctl);
    PW_CUDA(cudaGetLastError());
    dev_free(rs.block);
  } else {
    fits = G <= (uint64_t)cap_rows;
    const uint64_t n = fits ? G : 0;
    export_gather_kernel<<<(int)std::max<uint64_t>(1, (n + 255) / 256), 256, 0, c.stream>>>(T, pl, slots, n, n_parts, q->row_offset, (uint64_t*)send_device,
                                                                                            fits ? G : GATHER_OVERFLOW);
    PW_CUDA(cudaGetLastError());
    float ms;
    if (cudaEventElapsedTime(&ms, c.ev[8], c.ev[9]) == cudaSuccess) c.timings.scan_kernel_ms = ms;
  }
  c.timings.kernel_launches++;
  dev_free(slots);
  free_table(T);
  return fits ? 0 : 1;
}

int pw_b200_merge_gathered(const PwQuery* q, const PwFrame* schema_from, const void* gathered_device, int32_t world, int64_t cap_rows,
                           int32_t my_rank, struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out) {
  PW_TRY(ensure_device());
  if (!q || !schema_from || !gathered_device || !out_cols || !out_schemas || !n_out || world < 1 || cap_rows < 1) return fail(PW_ERR_INVALID, "bad argument");
  Lowered L;
  PW_TRY(lower_query(q, schema_from, &L));
  const int64_t rw = make_layout(L, kw_class(L.plan.n_kw)).row_words;
  RowSrc src{};
  src.base = (const uint64_t*)gathered_device;
  src.seg_words = 1 + (uint64_t)cap_rows * (uint64_t)(rw + 1);
  src.cap_rows = (uint64_t)cap_rows;
  src.gathered = 1; src.my_rank = my_rank;
  const int rc = merge_impl(q, schema_from, src, (int64_t)world * cap_rows, (uint64_t)std::max<int64_t>(4 * cap_rows, 64), out_cols, out_schemas, n_out);
  // the local scan of this exchange (pw_b200_frame_groupby_partial_into, deferred) has finished by now: its kernel time
  ThreadCtx& c = ctx();
  float ms;
  if (cudaEventElapsedTime(&ms, c.ev[8], c.ev[9]) == cudaSuccess) c.timings.scan_kernel_ms = ms;
  return rc;
}

// returns 0, an error, or 1 = "repeat through the general exchange" (gathered input only: a rank overflowed its
// segment, or this rank owns more groups than the optimistic table holds)
static int merge_impl(const PwQuery* q, const PwFrame* schema_from, RowSrc src, int64_t n_rows, uint64_t cap,
                      struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out) {
  ThreadCtx& c = ctx();
  Lowered L;
  PW_TRY(lower_query(q, schema_from, &L));
  const int kw = kw_class(L.plan.n_kw);
  PackLayout pl = make_layout(L, kw);
  const uint64_t nn = cap + 2;
  // Small merge tables keep the group count on the device: the control block is the header of the result block and
  // the whole merge (upsert, compaction, ordering, emission, copy) needs ONE host synchronisation (emit_results).
  char* block = nullptr;
  bool deferred = false;
  PW_TRY(alloc_result_block(L, nn, &block, &deferred));
  Control* dctl = nullptr;
  void* v = nullptr;
  if (deferred) dctl = (Control*)block;
  else { PW_TRY(dev_alloc(&v, sizeof(Control))); dctl = (Control*)v; }
  Table T{};
  uint32_t *row_slot = nullptr, *slots = nullptr;
  uint64_t* fl_values = nullptr;
  auto drop = [&]() { dev_free(slots); dev_free(row_slot); dev_free(fl_values); if (T.keys) free_table(T); };
  int rc = alloc_table_raw(&T, kw, L.plan.n_acc, cap, &dctl->overflow, &dctl->spilled);
  if (rc) { dev_free(deferred ? (void*)block : (void*)dctl); return rc; }
  AccOps ops{};
  AccOpsK opsk{};
  ops.n = opsk.n = L.plan.n_acc;
  for (int a = 0; a < L.plan.n_acc; ++a) ops.op[a] = opsk.op[a] = L.plan.accs[a].op;
  table_init_kernel<<<(int)std::min<uint64_t>((nn + 255) / 256, 148 * 8), 256, 0, c.stream>>>(T, kw, ops, dctl);  // clears the control block too
  if ((rc = dev_alloc(&v, (size_t)std::max<int64_t>(n_rows, 1) * 4)) == 0) row_slot = (uint32_t*)v;
  if (!rc && (rc = dev_alloc(&v, nn * 8 * (uint64_t)std::max(1, pl.n_fl))) == 0) fl_values = (uint64_t*)v;
  if (!rc && (rc = dev_alloc(&v, nn * 4)) == 0) slots = (uint32_t*)v;
  if (rc) { drop(); dev_free(deferred ? (void*)block : (void*)dctl); return rc; }
  if (n_rows) {
    const int grid = (int)((n_rows + 255) / 256);
    switch (kw) {
      case 1: merge_kernel<1><<<grid, 256, 0, c.stream>>>(T, pl, opsk, src, (uint64_t)n_rows, row_slot); break;
      case 2: merge_kernel<2><<<grid, 256, 0, c.stream>>>(T, pl, opsk, src, (uint64_t)n_rows, row_slot); break;
      case 4: merge_kernel<4><<<grid, 256, 0, c.stream>>>(T, pl, opsk, src, (uint64_t)n_rows, row_slot); break;
      default: merge_kernel<6><<<grid, 256, 0, c.stream>>>(T, pl, opsk, src, (uint64_t)n_rows, row_slot); break;
    }
    if (pl.n_fl) merge_values_kernel<<<grid, 256, 0, c.stream>>>(T, pl, src, (uint64_t)n_rows, row_slot, fl_values);
    c.timings.kernel_launches += 3;
  }
  // compact + order + emit (first/last values come from fl_values instead of a column gather)
  compact_kernel<<<(int)std::min<uint64_t>((nn + 255) / 256, 148 * 8), 256, 0, c.stream>>>(T, kw, slots, &dctl->counter);
  if (cudaGetLastError() != cudaSuccess) { drop(); dev_free(deferred ? (void*)block : (void*)dctl); return fail(PW_ERR_CUDA, "merge launch failed"); }
  int f_idx = 0;
  for (OutCol& o : L.outs)
    if (o.emit.kind == EMIT_FIRSTLAST) {
      for (int f = 0; f < pl.n_fl; ++f) if (pl.fl_acc[f] == o.emit.acc) f_idx = f;
      o.emit.fl_values = fl_values + (uint64_t)f_idx * nn;
    }
  if (deferred) {
    RunState rs;
    rs.deferred = true; rs.dctl = dctl; rs.block = block; rs.cap = cap;
    rc = order_groups(L, T, kw, &slots, nn, &dctl->counter);
    if (rc) { drop(); dev_free(block); return rc; }
    rc = emit_results(L, T, slots, nn, out_cols, out_schemas, n_out, &rs);   // frees the block; PW_RETRY on overflow
    drop();
    if (rc == PW_RETRY) return src.gathered ? 1 : fail(PW_ERR_INTERNAL, "merge table overflow");
    return rc;
  }
  Control hctl{};
  if (cudaMemcpyAsync(&hctl, dctl, sizeof(Control), cudaMemcpyDeviceToHost, c.stream) != cudaSuccess || cudaStreamSynchronize(c.stream) != cudaSuccess) {
    drop(); dev_free(dctl);
    return fail(PW_ERR_CUDA, "merge failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  dev_free(dctl);
  if (hctl.overflow) {
    drop();
    if (src.gathered) return 1;
    return fail(PW_ERR_INTERNAL, "merge table overflow");
  }
  const uint64_t G = hctl.counter;
  rc = order_groups(L, T, kw, &slots, G);
  if (!rc) rc = emit_results(L, T, slots, G, out_cols, out_schemas, n_out);
  drop();
  return rc;
}

}  // extern "C"
