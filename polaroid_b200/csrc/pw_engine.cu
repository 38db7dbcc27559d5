// pw_engine.cu — host orchestration of the filter -> group_by -> agg operator:
//   frame upload (Arrow buffers -> HBM), query lowering (PwQuery -> ScanPlan + result plan), strategy
//   selection from key samples (the analogue of can_run_partitioned / estimate_unique_count,
//   polars-mem-engine/src/executors/group_by_streaming.rs:114-244), kernel launches, finalisation and
//   Arrow result construction.  One process per GPU; every call runs on the calling thread's stream.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "pw_engine.h"
#include "pw_pilot.cuh"
#include "pw_radix.cuh"

namespace pw {

// ------------------------------------------------------------------------------------------------------
// thread context / errors / device memory
// ------------------------------------------------------------------------------------------------------
static thread_local ThreadCtx g_ctx;
ThreadCtx& ctx() { return g_ctx; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_ctx.last_error = buf;
  return code;
}

static std::mutex g_pool_mu;
static bool g_pool_done[64] = {false};

int ensure_device() {
  ThreadCtx& c = ctx();
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(PW_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (c.device >= n) return fail(PW_ERR_CUDA, "device %d out of range (%d devices)", c.device, n);
  PW_CUDA(cudaSetDevice(c.device));
  if (!c.pool_ready) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (c.device < 64 && !g_pool_done[c.device]) {
      cudaMemPool_t pool;
      PW_CUDA(cudaDeviceGetDefaultMemPool(&pool, c.device));
      uint64_t thr = UINT64_MAX;  // keep freed blocks cached: allocation stays off the hot path
      PW_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
      g_pool_done[c.device] = true;
    }
    PW_CUDA(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, c.device));
    c.pool_ready = true;
  }
  if (!c.ev_ready) {
    for (auto& e2 : c.ev) PW_CUDA(cudaEventCreate(&e2));
    c.ev_ready = true;
  }
  return 0;
}

int dev_alloc(void** p, size_t bytes) {
  if (bytes == 0) bytes = 16;
  PW_CUDA(cudaMallocAsync(p, bytes, ctx().stream));
  return 0;
}
void dev_free(void* p) {
  if (p) cudaFreeAsync(p, ctx().stream);
}

static inline int dtype_bytes(int dt) {
  switch (dt) {
    case DT_I8: case DT_U8: return 1;
    case DT_I16: case DT_U16: return 2;
    case DT_I32: case DT_U32: case DT_F32: return 4;
    case DT_VIEW: return 16;
    default: return 8;
  }
}
// bytes of a result column of G rows (Boolean columns are bit-packed in 32-bit words)
static inline size_t out_col_bytes(int dt, uint64_t G) {
  return dt == DT_BOOL ? (size_t)((G + 31) / 32) * 4 : (size_t)G * dtype_bytes(dt);
}
static inline int dtype_class(int dt) {
  switch (dt) {
    case DT_U8: case DT_U16: case DT_U32: case DT_U64: case DT_BOOL: return CLS_U64;
    case DT_F32: case DT_F64: return CLS_F64;
    default: return CLS_I64;
  }
}

// ------------------------------------------------------------------------------------------------------
// query lowering
// ------------------------------------------------------------------------------------------------------
namespace {
struct Lowerer {
  const PwQuery* q;
  const PwFrame* f;
  Lowered* L;
  int slot_of_col[256];

  int slot_for(int col) {
    if (slot_of_col[col] >= 0) return slot_of_col[col];
    const FrameColumn& c = f->cols[col];
    ScanPlan& P = L->plan;
    const int need = c.dtype == DT_VIEW ? 2 : 1;
    if (P.n_slots + need > MAX_SLOTS) return -1;
    const int s = P.n_slots;
    P.slots[s].values = c.values;
    P.slots[s].validity = c.null_count ? c.validity : nullptr;
    P.slots[s].dtype = c.dtype;
    P.slots[s].bit_offset = c.bit_offset;
    if (need == 2) {
      P.slots[s + 1] = P.slots[s];
      P.slots[s + 1].dtype = DT_VIEW_HI;
      P.slots[s + 1].validity = nullptr;
    }
    P.n_slots += need;
    slot_of_col[col] = s;
    return s;
  }
  int vexpr_plain(int col) {
    ScanPlan& P = L->plan;
    const int s = slot_for(col);
    if (s < 0) return -1;
    for (int e = 0; e < P.n_vexpr; ++e)
      if (P.vexprs[e].n_factors == 0 && P.vexprs[e].slot == s) return e;
    if (P.n_vexpr >= MAX_VEXPR) return -1;
    VExpr& ve = P.vexprs[P.n_vexpr];
    ve.n_factors = 0; ve.slot = s; ve.cls = dtype_class(f->cols[col].dtype);
    return P.n_vexpr++;
  }
  int vexpr_product(const PwAgg& a) {
    ScanPlan& P = L->plan;
    VExpr ve{};
    ve.n_factors = a.n_factors; ve.slot = 0; ve.cls = CLS_F64;
    for (int i = 0; i < a.n_factors; ++i) {
      const int s = slot_for(a.factors[i].column);
      if (s < 0) return -1;
      ve.f[i].a = a.factors[i].a; ve.f[i].b = a.factors[i].b; ve.f[i].slot = s;
    }
    for (int e = 0; e < P.n_vexpr; ++e) {
      const VExpr& o = P.vexprs[e];
      if (o.n_factors != ve.n_factors) continue;
      bool same = true;
      for (int i = 0; i < ve.n_factors; ++i)
        same = same && o.f[i].a == ve.f[i].a && o.f[i].b == ve.f[i].b && o.f[i].slot == ve.f[i].slot;
      if (same) return e;
    }
    if (P.n_vexpr >= MAX_VEXPR) return -1;
    P.vexprs[P.n_vexpr] = ve;
    return P.n_vexpr++;
  }
  int acc(int op, int src, int vexpr) {
    ScanPlan& P = L->plan;
    for (int a = 0; a < P.n_acc; ++a)
      if (P.accs[a].op == op && P.accs[a].src == src && P.accs[a].vexpr == vexpr) return a;
    if (P.n_acc >= MAX_ACC) return -1;
    P.accs[P.n_acc].op = op; P.accs[P.n_acc].src = src; P.accs[P.n_acc].vexpr = vexpr;
    return P.n_acc++;
  }
};

}  // namespace

// mean of a strided sample (<= 4096 non-null rows) of a numeric column: the shift of its var/std words
static __global__ void shift_sample_kernel(RawSlot s, int64_t n, int64_t stride, double* sum, unsigned long long* cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = i * stride;
  double v = 0.0;
  bool ok = row < n;
  if (ok && s.validity) { const int64_t b = (int64_t)s.bit_offset + row; ok = (s.validity[b >> 3] >> (b & 7)) & 1; }
  if (ok) {
    v = bits_to_f64(decode(load_row(s, row), s.dtype, 0), slot_class(s.dtype));
    ok = v == v && fabs(v) < 1.0e300;
  }
  if (ok) { atomicAdd(sum, v); atomicAdd(cnt, 1ull); }
}
static int column_shift(const PwFrame* f, int col, const RawSlot& slot, double* out) {
  const FrameColumn& c = f->cols[col];
  std::lock_guard<std::mutex> lk(f->mu);
  if (!c.shift_known) {
    struct { double sum; unsigned long long cnt; } h{0.0, 0ull};
    if (f->n_rows > 0) {   // (schema-only frames — output dtype inference — never touch the device)
      PW_TRY(ensure_device());
      ThreadCtx& t = ctx();
      void* d = nullptr;
      PW_TRY(dev_alloc(&d, 16));
      PW_CUDA(cudaMemsetAsync(d, 0, 16, t.stream));
      const int64_t stride = std::max<int64_t>(1, f->n_rows / 4096);
      const int64_t m = (f->n_rows + stride - 1) / stride;
      shift_sample_kernel<<<(unsigned)((m + 255) / 256), 256, 0, t.stream>>>(slot, f->n_rows, stride, (double*)d, (unsigned long long*)d + 1);
      PW_CUDA(cudaMemcpyAsync(&h, d, 16, cudaMemcpyDeviceToHost, t.stream));
      PW_CUDA(cudaStreamSynchronize(t.stream));
      dev_free(d);
    }
    c.var_shift = h.cnt ? h.sum / (double)h.cnt : 0.0;
    c.shift_known = true;
  }
  *out = c.var_shift;
  return 0;
}

int lower_query(const PwQuery* q, const PwFrame* f, Lowered* L) {
  if (!q || !f) return fail(PW_ERR_INVALID, "null query or frame");
  if (q->abi_version != PW_ABI_VERSION) return fail(PW_ERR_INVALID, "PwQuery.abi_version %u != %u", q->abi_version, PW_ABI_VERSION);
  if (f->cols.size() > 256) return fail(PW_ERR_UNSUPPORTED, "more than 256 columns");
  const int ncols = (int)f->cols.size();
  Lowerer lw{q, f, L, {}};
  for (int i = 0; i < 256; ++i) lw.slot_of_col[i] = -1;
  ScanPlan& P = L->plan;
  memset(&P, 0, sizeof P);
  P.n_rows = f->n_rows;
  P.row_begin = 0; P.row_stride = 1;
  P.row_offset = q->row_offset;
  auto col_ok = [&](int c) { return c >= 0 && c < ncols; };

  // ---- predicates (a1/a2)
  if (q->n_predicates > MAX_PREDS) return fail(PW_ERR_UNSUPPORTED, "more than %d predicate conjuncts", MAX_PREDS);
  for (int i = 0; i < q->n_predicates; ++i) {
    const PwPredicate& p = q->predicates[i];
    if (!col_ok(p.column)) return fail(PW_ERR_INVALID, "predicate column %d out of range", p.column);
    const FrameColumn& c = f->cols[p.column];
    if (c.dtype == DT_VIEW) return fail(PW_ERR_UNSUPPORTED, "predicate on column '%s' of format %s", c.name.c_str(), c.format.c_str());
    const int s = lw.slot_for(p.column);
    if (s < 0) return fail(PW_ERR_UNSUPPORTED, "query touches more than %d column slots", MAX_SLOTS);
    Pred& d = P.preds[P.n_preds++];
    d.slot = s; d.op = p.op; d.cls = dtype_class(c.dtype);
    if (d.cls == CLS_F64) {
      double v = p.scalar_is_float ? p.scalar.f : (double)p.scalar.i;
      memcpy(&d.scalar, &v, 8);
    } else {
      if (p.scalar_is_float) return fail(PW_ERR_UNSUPPORTED, "float scalar compared with integer column '%s'", c.name.c_str());
      d.scalar = p.scalar.u;
    }
  }

  // ---- dynamic options
  const PwDynamic* dyn = q->dynamic;
  int dyn_slot = -1;
  if (dyn) {
    if (!col_ok(dyn->index_column)) return fail(PW_ERR_INVALID, "index column out of range");
    const FrameColumn& c = f->cols[dyn->index_column];
    // polars-time/src/group_by/dynamic.rs:210-258 accepts Date, Datetime, Int32 and Int64 only (physical i32 / i64)
    {
      const bool fmt_ok = c.format == "i" || c.format == "l" || c.format == "tdD" || c.format.rfind("ts", 0) == 0;
      if (!fmt_ok || (c.dtype != DT_I32 && c.dtype != DT_I64))
        return fail(PW_ERR_INVALID, "expected any of the following dtypes: { Date, Datetime, Int32, Int64 }, got column '%s' of format %s", c.name.c_str(), c.format.c_str());
    }
    if (c.null_count) return fail(PW_ERR_INVALID, "null values in `group_by_dynamic` index column are not supported");
    if (dyn->every <= 0) return fail(PW_ERR_INVALID, "'every' argument must be positive");
    if (dyn->period <= 0) return fail(PW_ERR_INVALID, "'period' argument must be positive");
    dyn_slot = lw.slot_for(dyn->index_column);
    if (dyn_slot < 0) return fail(PW_ERR_UNSUPPORTED, "too many column slots");
    const bool overlapping = dyn->closed == PW_CLOSED_BOTH ? dyn->period >= dyn->every : dyn->period > dyn->every;  // dynamic.rs:312-315
    L->tumbling = !overlapping;
    P.overlap = overlapping ? 1 : 0;   // hash path: every row joins all the windows that contain it (pw_overlap.cuh)
    P.dyn.enabled = 1; P.dyn.slot = dyn_slot; P.dyn.closed = dyn->closed;
    P.dyn.every = dyn->every; P.dyn.period = dyn->period;
    // window starts lie on the grid offset + k*every: truncate(t0, every) + offset (window.rs:115-170)
    P.dyn.origin = dyn->offset;
    div_prepare((uint64_t)dyn->every, &P.dyn.div_magic, &P.dyn.div_more);
    P.check_sorted = q->n_keys == 0;
  }

  // ---- keys (a3/a4/a8)
  if (q->n_keys > MAX_KEYS) return fail(PW_ERR_UNSUPPORTED, "more than %d key columns (multi-column row encoding is SURVEY 8f rank 1)", MAX_KEYS);
  int n_words = 0;
  bool any_nullable = false;
  for (int i = 0; i < q->n_keys; ++i) {
    const int kcol = q->key_columns[i];
    if (!col_ok(kcol)) return fail(PW_ERR_INVALID, "key column %d out of range", kcol);
    const FrameColumn& c = f->cols[kcol];
    const int s = lw.slot_for(kcol);   // Boolean keys: the key word is 0 / 1 (the reference row-encodes them, hash_keys.rs:37,114-141)
    if (s < 0) return fail(PW_ERR_UNSUPPORTED, "too many column slots");
    KeyCol& k = P.keys[P.n_keys++];
    k.slot = s; k.dtype = c.dtype; k.n_words = c.dtype == DT_VIEW ? 2 : 1; k.nullable = c.null_count != 0;
    // windows are emitted key slice by key slice in key order, and the view words only order values of up to 12 bytes
    if (dyn && c.has_long) return fail(PW_ERR_UNSUPPORTED, "group_by_dynamic by a string key longer than 12 bytes (column '%s')", c.name.c_str());
    any_nullable = any_nullable || k.nullable;
    n_words += k.n_words;
  }
  if (dyn) n_words += 1;
  const bool single_plain_key = q->n_keys == 1 && !dyn && P.keys[0].n_words == 1;
  if (any_nullable && !single_plain_key) { P.has_null_word = 1; L->null_word = n_words; n_words += 1; }
  else L->null_word = -1;
  L->single_key_null = (any_nullable && single_plain_key) ? 1 : 0;
  if (n_words == 0) n_words = 1;  // global aggregation: one constant key word
  if (n_words > MAX_KW) return fail(PW_ERR_UNSUPPORTED, "key wider than %d 64-bit words", MAX_KW);
  P.n_kw = n_words;

  // ---- output plan: keys first
  int word = 0;
  for (int i = 0; i < q->n_keys; ++i) {
    const FrameColumn& c = f->cols[q->key_columns[i]];
    OutCol o;
    o.name = c.name; o.format = c.format; o.nullable = c.null_count != 0;
    o.emit.kind = c.dtype == DT_VIEW ? EMIT_KEY_VIEW : EMIT_KEY_INT;
    if (c.dtype == DT_VIEW) o.src_col = &c;
    o.emit.out_dtype = c.dtype; o.out_dtype = c.dtype;
    o.emit.word = word;
    o.emit.null_word = P.keys[i].nullable ? L->null_word : -1;
    o.emit.nullbit = i;
    o.emit.single_key_null = (P.keys[i].nullable && L->null_word < 0) ? 1 : 0;
    L->outs.push_back(o);
    word += P.keys[i].n_words;
  }
  const int dyn_word = word;
  if (dyn) {
    const FrameColumn& c = f->cols[dyn->index_column];
    auto bound_col = [&](const char* name, int kind) {
      OutCol o; o.name = name; o.format = c.format; o.nullable = false;
      o.emit.kind = kind; o.emit.out_dtype = c.dtype; o.out_dtype = c.dtype; o.emit.word = dyn_word;
      o.emit.every = dyn->every; o.emit.period = dyn->period; o.emit.origin = dyn->offset;
      o.emit.null_word = -1;
      L->outs.push_back(o);
    };
    if (dyn->include_boundaries) { bound_col("_lower_boundary", EMIT_DYN_LOWER); bound_col("_upper_boundary", EMIT_DYN_UPPER); }
    if (dyn->label == PW_LABEL_LEFT) bound_col(c.name.c_str(), EMIT_DYN_LOWER);
    else if (dyn->label == PW_LABEL_RIGHT) bound_col(c.name.c_str(), EMIT_DYN_UPPER);
    else {
      OutCol o; o.name = c.name; o.format = c.format; o.nullable = false;
      o.emit.kind = EMIT_ACC_I64; o.emit.acc = -1 /* patched below */; o.emit.out_dtype = c.dtype; o.out_dtype = c.dtype; o.emit.null_word = -1;
      L->outs.push_back(o);
    }
  }

  // ---- aggregations (a6/a10): first collect, per value expression, which aggregate words are needed
  struct AggTmp { OutCol o; int ve; int kind; bool ve_nullable; bool is_float; int cls; };
  std::vector<AggTmp> tmp;
  int gflags = 0;
  for (int i = 0; i < q->n_aggs; ++i) {
    const PwAgg& a = q->aggs[i];
    AggTmp t{};
    OutCol& o = t.o;
    o.name = a.name ? a.name : "";
    o.emit.null_word = -1;
    o.emit.row_offset = q->row_offset;
    t.kind = a.kind; t.ve = -1;
    if (a.kind == PW_LEN) {
      gflags |= GF_LEN;
      o.format = "I"; o.out_dtype = DT_U32; o.emit.kind = EMIT_COUNT; o.emit.out_dtype = DT_U32;
      tmp.push_back(t);
      continue;
    }
    int ve, in_dtype = DT_F64;
    std::string in_format = "g";
    if (a.n_factors > 0) {
      if (a.n_factors > PW_MAX_FACTORS) return fail(PW_ERR_INVALID, "too many factors");
      for (int k = 0; k < a.n_factors; ++k) {
        if (!col_ok(a.factors[k].column)) return fail(PW_ERR_INVALID, "factor column out of range");
        const int dt = f->cols[a.factors[k].column].dtype;
        if (dt == DT_VIEW) return fail(PW_ERR_UNSUPPORTED, "arithmetic on a non-numeric column");
      }
      ve = lw.vexpr_product(a);
    } else {
      if (!col_ok(a.column)) return fail(PW_ERR_INVALID, "aggregation column %d out of range", a.column);
      const FrameColumn& c = f->cols[a.column];
      if (c.dtype == DT_VIEW && a.kind != PW_COUNT && a.kind != PW_NULL_COUNT)
        return fail(PW_ERR_UNSUPPORTED, "aggregation over column '%s' of format %s (string aggregations are SURVEY 8f)", c.name.c_str(), c.format.c_str());
      in_dtype = c.dtype; in_format = c.format;
      ve = lw.vexpr_plain(a.column);
    }
    if (ve < 0) return fail(PW_ERR_UNSUPPORTED, "query needs more than %d column slots / %d value expressions", MAX_SLOTS, MAX_VEXPR);
    VExpr& V = P.vexprs[ve];
    const int cls = V.cls;
    const bool is_float = cls == CLS_F64;
    const bool temporal = in_format.size() > 1 && in_format[0] == 't';
    // non-null count of this value: an expression without nulls shares the single per-group row counter
    bool ve_nullable = false;
    if (V.n_factors == 0) ve_nullable = P.slots[V.slot].validity != nullptr;
    else for (int k = 0; k < V.n_factors; ++k) ve_nullable = ve_nullable || P.slots[V.f[k].slot].validity != nullptr;
    auto need_count = [&]() { if (ve_nullable) V.flags |= VF_COUNT; else gflags |= GF_LEN; };
    t.ve = ve; t.ve_nullable = ve_nullable; t.is_float = is_float; t.cls = cls;
    const bool is_bool = in_dtype == DT_BOOL;
    switch (a.kind) {
      case PW_SUM:
        if (temporal && in_format[1] != 'D') return fail(PW_ERR_UNSUPPORTED, "`sum` operation not supported for dtype %s", in_format.c_str());
        if (is_bool) {   // Boolean -> IDX_DTYPE (sum.rs:43)
          V.flags |= VF_SUM_I;
          o.emit.kind = EMIT_SUM_INT; o.out_dtype = DT_U32; o.format = "I";
        } else if (is_float) {
          V.flags |= VF_SUM_F;
          o.emit.kind = EMIT_SUM_F64; o.out_dtype = in_dtype == DT_F32 ? DT_F32 : DT_F64; o.format = in_dtype == DT_F32 ? "f" : "g";
        } else {
          V.flags |= VF_SUM_I;
          o.emit.kind = EMIT_SUM_INT;
          // i8/i16/u8/u16 -> Int64, others keep their dtype (sum.rs:40-47)
          if (in_dtype == DT_I8 || in_dtype == DT_I16 || in_dtype == DT_U8 || in_dtype == DT_U16) { o.out_dtype = DT_I64; o.format = "l"; }
          else { o.out_dtype = in_dtype; o.format = in_format; }
        }
        break;
      case PW_MEAN:
        V.flags |= VF_SUM_F; need_count();
        o.emit.kind = EMIT_MEAN; o.nullable = true;
        if (in_dtype == DT_F32 && !temporal) { o.emit.mean_out = MEAN_F32; o.out_dtype = DT_F32; o.format = "f"; }
        else if (in_format == "tdD") { o.emit.mean_out = MEAN_DATE_US; o.out_dtype = DT_I64; o.format = "tsu:"; }
        else if (temporal) { o.emit.mean_out = MEAN_I64; o.out_dtype = DT_I64; o.format = in_format; }
        else { o.emit.mean_out = MEAN_F64; o.out_dtype = DT_F64; o.format = "g"; }
        break;
      case PW_MIN: case PW_MAX:
        V.flags |= (a.kind == PW_MIN ? VF_MIN : VF_MAX); need_count();
        o.emit.kind = is_float ? EMIT_MINMAX_F64 : EMIT_MINMAX_INT;
        o.out_dtype = in_dtype; o.format = in_format; o.nullable = true;
        break;
      case PW_COUNT:
        need_count();
        o.emit.kind = EMIT_COUNT; o.out_dtype = DT_U32; o.format = "I";
        break;
      case PW_FIRST: case PW_LAST:
        if (a.n_factors > 0) return fail(PW_ERR_UNSUPPORTED, "first/last of a computed expression");
        V.flags |= (a.kind == PW_FIRST ? VF_FIRST : VF_LAST);
        o.emit.kind = EMIT_FIRSTLAST; o.out_dtype = in_dtype; o.format = in_format; o.nullable = true;
        o.emit.src = P.slots[V.slot];
        break;
      case PW_FIRST_NON_NULL: case PW_LAST_NON_NULL:   // reduce/first_last_nonnull.rs: nulls never replace a value
        if (a.n_factors > 0) return fail(PW_ERR_UNSUPPORTED, "first/last of a computed expression");
        V.flags |= (a.kind == PW_FIRST_NON_NULL ? VF_FIRST_NN : VF_LAST_NN);
        o.emit.kind = EMIT_FIRSTLAST; o.out_dtype = in_dtype; o.format = in_format; o.nullable = true;
        o.emit.src = P.slots[V.slot];
        o.emit.period = 1;   // "still at init" means no non-null row: null
        o.emit.every = (int64_t)acc_init(a.kind == PW_FIRST_NON_NULL ? OP_MIN_U64 : OP_MAX_U64);
        break;
      case PW_VAR: case PW_STD:
        // reduce/var_std.rs:9-48: numeric and Boolean inputs; the result is Float64 (Float32 stays Float32)
        if (temporal) return fail(PW_ERR_UNSUPPORTED, "`%s` operation not supported for dtype %s", a.kind == PW_STD ? "std" : "var", in_format.c_str());
        if (a.ddof < 0 || a.ddof > 255) return fail(PW_ERR_INVALID, "ddof must be in [0, 255]");
        V.flags |= VF_SUMD | VF_SUMD2; need_count();
        o.emit.kind = EMIT_VAR; o.nullable = true; o.emit.pad = a.ddof; o.emit.src_cls = a.kind == PW_STD ? 1 : 0;
        if (in_dtype == DT_F32) { o.emit.mean_out = MEAN_F32; o.out_dtype = DT_F32; o.format = "f"; }
        else { o.emit.mean_out = MEAN_F64; o.out_dtype = DT_F64; o.format = "g"; }
        break;
      case PW_NULL_COUNT:
        gflags |= GF_LEN;
        if (ve_nullable) V.flags |= VF_COUNT;
        o.emit.kind = EMIT_NULL_COUNT; o.out_dtype = DT_U32; o.format = "I";
        break;
      case PW_BIT_AND: case PW_BIT_OR: case PW_BIT_XOR:
        if (is_float || temporal || a.n_factors > 0) return fail(PW_ERR_UNSUPPORTED, "bitwise aggregation of a non-integer column");
        V.flags |= (a.kind == PW_BIT_AND ? VF_AND : (a.kind == PW_BIT_OR ? VF_OR : VF_XOR)); need_count();
        o.emit.kind = EMIT_BITWISE; o.out_dtype = in_dtype; o.format = in_format; o.nullable = true;
        break;
      case PW_ANY: case PW_ALL:
        if (!is_bool) return fail(PW_ERR_INVALID, "any/all need a Boolean column, got %s", in_format.c_str());
        V.flags |= (a.kind == PW_ANY ? VF_MAX : VF_MIN);
        o.emit.kind = EMIT_ANYALL; o.out_dtype = DT_BOOL; o.format = "b"; o.nullable = false;
        break;
      default: return fail(PW_ERR_INVALID, "unknown aggregation kind %d", a.kind);
    }
    o.emit.out_dtype = o.out_dtype;
    if (o.emit.kind != EMIT_VAR) o.emit.src_cls = cls;   // EMIT_VAR keeps its var/std switch there
    tmp.push_back(t);
  }
  if (dyn) gflags |= GF_LEN;  // the sorted fast path finds the non-empty windows through the row counter
  const bool want_row = !dyn && q->maintain_order;
  const bool want_tmin = dyn && dyn->label == PW_LABEL_DATAPOINT;
  if (want_row) gflags |= GF_ROW;
  if (want_tmin) gflags |= GF_TMIN;

  // ---- accumulator words: consecutive per value expression in VFlag order, then LEN, ROW, TMIN
  auto push_acc = [&](int op, int src, int ve) { P.accs[P.n_acc].op = op; P.accs[P.n_acc].src = src; P.accs[P.n_acc].vexpr = ve; return P.n_acc++; };
  {
    int words = 0;
    for (int e = 0; e < P.n_vexpr; ++e) words += __builtin_popcount(P.vexprs[e].flags);
    words += __builtin_popcount(gflags);
    if (words > MAX_ACC) return fail(PW_ERR_UNSUPPORTED, "query needs more than %d accumulator words", MAX_ACC);
  }
  for (int e = 0; e < P.n_vexpr; ++e) {
    VExpr& V = P.vexprs[e];
    V.acc_base = P.n_acc;
    const bool u = V.cls == CLS_U64, fl = V.cls == CLS_F64;
    if (V.flags & VF_SUM_I) push_acc(OP_ADD_I64, SRC_BITS, e);
    if (V.flags & VF_SUM_F) push_acc(OP_ADD_F64, SRC_F64, e);
    if (V.flags & VF_COUNT) push_acc(OP_ADD_I64, SRC_VALID, e);
    if (V.flags & VF_MIN) push_acc(u ? OP_MIN_U64 : OP_MIN_I64, fl ? SRC_F64_ORD : SRC_BITS, e);
    if (V.flags & VF_MAX) push_acc(u ? OP_MAX_U64 : OP_MAX_I64, fl ? SRC_F64_ORD : SRC_BITS, e);
    if (V.flags & VF_FIRST) push_acc(OP_MIN_U64, SRC_ROWIDX, e);
    if (V.flags & VF_LAST) push_acc(OP_MAX_U64, SRC_ROWIDX, e);
    if (V.flags & VF_SUMD) push_acc(OP_ADD_F64, SRC_F64_D, e);
    if (V.flags & VF_SUMD2) push_acc(OP_ADD_F64, SRC_F64_D2, e);
    if (V.flags & VF_FIRST_NN) push_acc(OP_MIN_U64, SRC_ROWIDX_NN, e);
    if (V.flags & VF_LAST_NN) push_acc(OP_MAX_U64, SRC_ROWIDX_NN, e);
    if (V.flags & VF_AND) push_acc(OP_AND_U64, SRC_BITS, e);
    if (V.flags & VF_OR) push_acc(OP_OR_U64, SRC_BITS, e);
    if (V.flags & VF_XOR) push_acc(OP_XOR_U64, SRC_BITS, e);
  }
  for (int e = 0; e < P.n_vexpr; ++e) {
    P.var_shift[e] = 0.0;
    if ((P.vexprs[e].flags & VF_SUMD) && P.vexprs[e].n_factors == 0) {
      int col = -1;
      for (int c2 = 0; c2 < ncols; ++c2) if (lw.slot_of_col[c2] == P.vexprs[e].slot) col = c2;
      if (col >= 0) PW_TRY(column_shift(f, col, P.slots[P.vexprs[e].slot], &P.var_shift[e]));
    }
  }
  if (P.n_acc == 0 && gflags == 0) gflags |= GF_LEN;  // a table needs at least one word per group
  P.gflags = gflags;
  P.acc_gbase = P.n_acc;
  int acc_len = -1, acc_row = -1, acc_tmin = -1;
  if (gflags & GF_LEN) acc_len = push_acc(OP_ADD_I64, SRC_ONE, 0);
  if (gflags & GF_ROW) acc_row = push_acc(OP_MIN_U64, SRC_ROW, 0);
  if (gflags & GF_TMIN) acc_tmin = push_acc(OP_MIN_I64, SRC_INDEX_T, 0);
  auto acc_of = [&](int e, int flag) {
    const VExpr& V = P.vexprs[e];
    int a = V.acc_base;
    for (int b2 = 1; b2 < flag; b2 <<= 1) if (V.flags & b2) ++a;
    return a;
  };
  for (AggTmp& t : tmp) {
    OutCol& o = t.o;
    auto cnt = [&]() { return t.ve_nullable ? acc_of(t.ve, VF_COUNT) : acc_len; };
    switch (t.kind) {
      case PW_LEN: o.emit.acc = acc_len; break;
      case PW_SUM: o.emit.acc = acc_of(t.ve, t.is_float ? VF_SUM_F : VF_SUM_I); break;
      case PW_MEAN: o.emit.acc = acc_of(t.ve, VF_SUM_F); o.emit.acc_cnt = cnt(); break;
      case PW_MIN: case PW_MAX:
        o.emit.acc = acc_of(t.ve, t.kind == PW_MIN ? VF_MIN : VF_MAX); o.emit.acc_cnt = cnt();
        o.emit.every = (int64_t)acc_init(P.accs[o.emit.acc].op);  // "no non-NaN value seen" marker for floats
        break;
      case PW_COUNT: o.emit.acc = cnt(); break;
      case PW_FIRST: o.emit.acc = acc_of(t.ve, VF_FIRST); break;
      case PW_LAST: o.emit.acc = acc_of(t.ve, VF_LAST); break;
      case PW_FIRST_NON_NULL: o.emit.acc = acc_of(t.ve, VF_FIRST_NN); break;
      case PW_LAST_NON_NULL: o.emit.acc = acc_of(t.ve, VF_LAST_NN); break;
      case PW_VAR: case PW_STD: o.emit.acc = acc_of(t.ve, VF_SUMD); o.emit.acc_nn = acc_of(t.ve, VF_SUMD2); o.emit.acc_cnt = cnt(); break;
      case PW_NULL_COUNT: o.emit.acc = acc_len; o.emit.acc_cnt = t.ve_nullable ? acc_of(t.ve, VF_COUNT) : -1; break;
      case PW_BIT_AND: o.emit.acc = acc_of(t.ve, VF_AND); o.emit.acc_cnt = cnt(); break;
      case PW_BIT_OR: o.emit.acc = acc_of(t.ve, VF_OR); o.emit.acc_cnt = cnt(); break;
      case PW_BIT_XOR: o.emit.acc = acc_of(t.ve, VF_XOR); o.emit.acc_cnt = cnt(); break;
      case PW_ANY: o.emit.acc = acc_of(t.ve, VF_MAX); break;
      case PW_ALL: o.emit.acc = acc_of(t.ve, VF_MIN); break;
      default: return fail(PW_ERR_INVALID, "unknown aggregation kind %d", t.kind);
    }
    L->outs.push_back(o);
  }
  if (want_tmin) {
    // label = datapoint: the index value of the first row of the window = its minimum (sorted input)
    for (OutCol& o : L->outs) if (o.emit.kind == EMIT_ACC_I64) o.emit.acc = acc_tmin;
  }

  // ---- ordering
  if (dyn) {
    // key slices ascending (nulls first) then windows ascending: group_by_rolling.rs:17-58 + dynamic.rs:317-362
    SortSpec w{}; w.src = SORT_WORD_I64; w.word = dyn_word; w.nullbit = 63; L->sort.push_back(w);  // least significant
    int wd = dyn_word;
    for (int i = q->n_keys - 1; i >= 0; --i) {
      wd -= P.keys[i].n_words;
      SortSpec sp{}; sp.word = wd; sp.nullbit = i; sp.single_key_null = 0;
      const int dt = P.keys[i].dtype;
      const int nullable = P.keys[i].nullable;
      auto push = [&](int src) { SortSpec t2 = sp; t2.src = src; if (!nullable) t2.nullbit = 63; L->sort.push_back(t2); };
      if (dt == DT_VIEW) { push(SORT_VIEW_LO); push(SORT_VIEW_HI); }
      else if (dtype_class(dt) == CLS_F64) push(SORT_WORD_F64);
      else if (dtype_class(dt) == CLS_U64) push(SORT_WORD_U64);
      else push(SORT_WORD_I64);
      if (nullable) push(SORT_NULLBIT);
    }
  } else if (want_row) {
    SortSpec sp{}; sp.src = SORT_ACC_U64; sp.acc = acc_row; L->sort.push_back(sp);
  }

  // vector loads need 16-byte aligned column bases
  P.vec_ok = 1;
  for (int s = 0; s < P.n_slots; ++s)
    if (((uintptr_t)P.slots[s].values & 15u) != 0) P.vec_ok = 0;
  L->n_nc = P.n_slots;
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------------

// the narrow kernel class holds <= 4 raw slots and <= 2 value expressions in registers
static bool narrow_class(const ScanPlan& P);
// ahead-of-time kernels only (the group_tuples lookup pass carries per-call pointers that must not enter the JIT key)
int launch_scan_aot(ScanPlan P, int sm, cudaStream_t st) {
  if (P.n_kw == 3) P.n_kw = 4;
  if (P.n_kw == 5) P.n_kw = 6;
  if (narrow_class(P)) return launch_scan_nc4(P, sm, st);
  return launch_scan_nc12(P, sm, st);
}
static bool narrow_class(const ScanPlan& P) { return P.n_slots <= 4 && P.n_vexpr <= NVof<4>::value; }
// CTA size of the scan launch (the specialised kernels take it as a launch parameter; PW_SCAN_THREADS overrides the
// narrow class for experiments)
static int scan_threads(const ScanPlan& P) {
  if (P.hot_slots > 0 && P.hot.threads > 0) return P.hot.threads;
  return narrow_class(P) ? ScanCfg<4>::THREADS : ScanCfg<12>::THREADS;
}
// CTA sizes worth trying for the hot table, best first: 16 warps hide the shared-memory round trips of the narrow
// class better than 12 (C2: 0.674 -> 0.635 ms) when 16 private regions still fit next to the full set of ids
static int wide_cta_threads() {
  static const int env = getenv("PW_SCAN_THREADS") ? atoi(getenv("PW_SCAN_THREADS")) : 0;
  return env >= 64 && env <= 1024 && env % 32 == 0 ? env : 512;
}
static int launch_scan(ScanPlan P, int sm, cudaStream_t st) {
  // the kernel is compiled for a few (slots, key words) classes; round the key width up (extra words are 0)
  if (P.n_kw == 3) P.n_kw = 4;
  if (P.n_kw == 5) P.n_kw = 6;
  const bool narrow = narrow_class(P);
  // query-shape specialised kernel (NVRTC); falls back to the ahead-of-time kernel of the same class
  const int kwc = P.n_kw <= 1 ? 1 : (P.n_kw <= 2 ? 2 : (P.n_kw <= 4 ? 4 : 6));
  if (P.overlap) {
    const int orc = launch_overlap_jit(P, narrow ? 4 : 12, kwc, sm, st);
    if (orc > 0) return fail(PW_ERR_UNSUPPORTED, "overlapping windows through the hash path need the run-time compiler (libnvrtc)");
    return orc;
  }
  if (P.runs) {
    // sorted keys: run-combining scan (specialised build only; otherwise the regular tiers below)
    const int rrc = launch_runs_jit(P, narrow ? 4 : 12, kwc, sm, st);
    if (rrc <= 0) { if (rrc == 0) ctx().timings.reserved = 3.0f; return rrc; }
    P.runs = 0;
  }
  if (P.hot_slots > 0 && P.hot.bucket) {
    // bucket tier (specialised build only); without NVRTC the per-cell geometry in the same HotGeom takes over
    const int brc = launch_bucket_jit(P, narrow ? 4 : 12, kwc, sm, st);
    if (brc <= 0) { if (brc == 0) ctx().timings.reserved = 2.0f; return brc; }
    P.hot.bucket = 0;
  }
  const int rc = launch_scan_jit(P, narrow ? 4 : 12, kwc, P.hot_slots > 0, scan_threads(P), sm, st);
  if (rc <= 0) { if (rc == 0) ctx().timings.reserved = 1.0f; return rc; }
  if (narrow) return launch_scan_nc4(P, sm, st);
  return launch_scan_nc12(P, sm, st);
}
static int padded_kw(int n_kw) { return n_kw == 3 ? 4 : (n_kw == 5 ? 6 : n_kw); }

// Shared-memory hot table geometry for `groups_hint` live groups (see pw_scan.cuh).  Returns false when no
// useful table fits.
// dense_range > 0: dense ids (id = key - dense_min) over exactly that many ids, no key index.
static bool plan_hot_threads(ScanPlan& P, int64_t groups_hint, int requested_gcap, int64_t dense_range, int threads, bool exact);
static bool plan_bucket(ScanPlan& P, int64_t groups, int64_t dense_range, bool sentinels, bool windowed, bool assume_jit = false);
static bool plan_hot(ScanPlan& P, int64_t groups_hint, int requested_gcap, int64_t dense_range = 0) {
  const int base = narrow_class(P) ? ScanCfg<4>::THREADS : ScanCfg<12>::THREADS;
  if (!plan_hot_threads(P, groups_hint, requested_gcap, dense_range, base, false)) return false;
  // one CTA per SM anyway (big table): more warps, if their private regions still fit next to the same ids
  if (narrow_class(P) && wide_cta_threads() != base && P.hot.total_bytes > 110 * 1024 && requested_gcap == 0) {
    const HotGeom keep = P.hot;
    if (!plan_hot_threads(P, groups_hint, keep.gcap, dense_range, wide_cta_threads(), true) || P.hot.n_mm != keep.n_mm || P.hot.replicas < keep.replicas)
      P.hot = keep;
  }
  return true;
}
// exact: fail instead of shrinking the id capacity
static bool plan_hot_threads(ScanPlan& P, int64_t groups_hint, int requested_gcap, int64_t dense_range, int threads, bool exact) {
  HotGeom& g = P.hot;
  memset(&g, 0, sizeof g);
  g.threads = threads;
  const bool dense = dense_range > 0;
  const int kw = padded_kw(P.n_kw);
  const int warps = threads / 32;
  int n_priv64 = 0, n_priv32 = 0, n_mm = 0;
  // words updated by a plain read-modify-write under exclusive ownership: sums and the bitwise folds
  auto is_add = [&](int a) { const int op = P.accs[a].op; return op == OP_ADD_F64 || op == OP_ADD_I64 || op == OP_AND_U64 || op == OP_OR_U64 || op == OP_XOR_U64; };
  auto is_count = [&](int a) { return P.accs[a].op == OP_ADD_I64 && (P.accs[a].src == SRC_ONE || P.accs[a].src == SRC_VALID); };
  for (int a = 0; a < P.n_acc; ++a) { if (is_count(a)) n_priv32++; else if (is_add(a)) n_priv64++; else n_mm++; }
  // dense ids: a little head-room over the live-group estimate
  int gcap = requested_gcap > 0 ? requested_gcap : (int)std::min<int64_t>(4096, std::max<int64_t>(8, groups_hint + groups_hint / 8 + 4));
  if (dense) gcap = (int)dense_range;
  const int mm_stride_all = n_mm > 1 ? ((n_mm + 1) & ~1) : n_mm;
  // 32-bit shadow (pw_scan.cuh, HotTable::shadow): the first (min, max) pair over one value expression whose two
  // CTA-shared words form an aligned 16-byte cell.  Shared words are numbered in accumulator order.
  int shadow_acc = -1;
  // (not over partitioned input: its groups live for a handful of rows, most of which improve an extremum — measured
  // 8.5 vs 7.1 ms on C3)
  if (!getenv("PW_NO_SHADOW") && (mm_stride_all & 1) == 0 && P.rowid_slot_p1 == 0) {
    int mm_i = 0;
    for (int a = 0; a < P.n_acc && shadow_acc < 0; ++a) {
      if (is_count(a) || is_add(a)) continue;
      if (a + 1 < P.acc_gbase && (mm_i & 1) == 0 && P.accs[a].vexpr == P.accs[a + 1].vexpr && P.accs[a].src == P.accs[a + 1].src &&
          (P.accs[a].src == SRC_BITS || P.accs[a].src == SRC_F64_ORD) &&
          ((P.accs[a].op == OP_MIN_I64 && P.accs[a + 1].op == OP_MAX_I64) || (P.accs[a].op == OP_MIN_U64 && P.accs[a + 1].op == OP_MAX_U64)))
        shadow_acc = a;
      ++mm_i;
    }
  }
  for (bool first = true;; gcap = gcap * 3 / 4, first = false) {
    if (gcap < 4 || ((dense || exact) && !first)) return false;  // a dense range is all or nothing
    // key index: buckets of four tags.  At 4 slots per id (25 % load) a row finds its key in the HOME bucket with
    // probability > 0.999, so the probe is one LDS.128 + one key compare and the neighbour bucket is only looked at
    // on the (rare, warp-uniformly branched) slow path; 2 slots per id is the fallback when memory is short.
    int S2 = 8;
    while (S2 < 2 * gcap) S2 <<= 1;
    if (dense) S2 = 4;  // one unused bucket keeps the "hot table on" switch (idx_slots != 0)
    // private bytes per cell: 8 per sum-like word (+ min/max words when they are private), 4 per counter; the claim
    // byte lives in the top byte of the first counter (a dedicated 4-byte word when the query has no counter)
    auto cell_bytes = [&](int mmp) { return (size_t)(n_priv64 + mmp) * 8 + (size_t)n_priv32 * 4 + (n_priv32 ? 0 : 4); };
    auto per_warp = [&](int R, int mmp) { return (((size_t)gcap * R * cell_bytes(mmp)) + 15) & ~(size_t)15; };
    auto total = [&](int S, int R, int mmp) {
      const size_t shared = (size_t)S * 4 + (dense ? 0 : (size_t)gcap * 8 * kw) + 48;
      return shared + (size_t)(mmp ? 0 : mm_stride_all) * gcap * 8 + ((mmp == 0 && shadow_acc >= 0) ? (size_t)gcap * 8 : 0) + per_warp(R, mmp) * warps;
    };
    // Preference order: every word warp-private (plain read-modify-write; first/last words improve on almost every
    // row, so CTA-shared atomics for them are very slow) with two CTAs per SM, then with one CTA per SM, then
    // min/max words CTA-shared (2 CTAs, 1 CTA), else fewer dense ids.  Inside a tier the larger key index wins.
    const size_t budget2 = 110 * 1024, budget1 = 224 * 1024;
    int mmp = -1, S = 0;
    size_t budget = 0;
    for (int tier = 0; tier < 4 && mmp < 0; ++tier) {
      const int cand_mmp = tier < 2 ? n_mm : 0;
      const size_t cand_budget = (tier & 1) ? budget1 : budget2;
      for (int mult = dense ? 1 : 2; mult >= 1; --mult)
        if (total(S2 * mult, 1, cand_mmp) <= cand_budget) { mmp = cand_mmp; budget = cand_budget; S = S2 * mult; break; }
    }
    if (mmp < 0) continue;  // fewer ids
    int R = 32;
    while (R > 1 && total(S, R, mmp) > budget) R >>= 1;
    if (requested_gcap > 0 && requested_gcap <= 64) R = std::min(R, 2);  // test hook: exercise the claim path
    g.idx_slots = S; g.gcap = gcap; g.replicas = R; g.n_mm = n_mm - mmp; g.dense = dense ? 1 : 0;
    g.mm_stride = mmp ? 0 : mm_stride_all;
    size_t off = (size_t)S * 4;
    off = (off + 15) & ~(size_t)15; g.keys_off = (int32_t)off; off += dense ? 0 : (size_t)gcap * 8 * kw;
    off = (off + 15) & ~(size_t)15; g.mm_off = (int32_t)off; off += (size_t)g.mm_stride * gcap * 8;
    g.count_off = (int32_t)off; off += 32;  // [count, full flag, -, -]
    g.guard_acc = (mmp == 0) ? shadow_acc : -1;
    g.shadow_off = (int32_t)off;
    if (g.guard_acc >= 0) off += (((size_t)gcap * 8) + 15) & ~(size_t)15;
    g.warp_off = (int32_t)off;
    size_t woff = 0;
    int mm_idx = 0;
    g.claim_acc = -1;
    for (int pass = 0; pass < 2; ++pass) {  // 8-byte words first, then 4-byte counters
      for (int a = 0; a < P.n_acc; ++a) {
        if (pass == 0) {
          if (is_count(a)) continue;
          if (!is_add(a) && mmp == 0) { g.acc_kind[a] = HOT_SHARED_MM; g.acc_off[a] = mm_idx++; continue; }
          g.acc_kind[a] = HOT_PRIV64; g.acc_off[a] = (int32_t)woff; woff += (size_t)gcap * R * 8;
        } else if (is_count(a)) {
          g.acc_kind[a] = HOT_PRIV32; g.acc_off[a] = (int32_t)woff; woff += (size_t)gcap * R * 4;
          if (g.claim_acc < 0) { g.claim_acc = a; g.claim_off = g.acc_off[a]; }
        }
      }
    }
    // duplicate rows of a warp instruction: combined in registers through warp votes (no claim byte) unless switched off
    if (g.claim_acc < 0) { g.claim_off = (int32_t)woff; woff += (size_t)gcap * R * 4; }
    woff = (woff + 15) & ~(size_t)15;
    g.warp_bytes = (int32_t)woff;
    g.total_bytes = (int32_t)(off + woff * warps);
    return true;
  }
}


// Bucket tier geometry (pw_bucket.cuh) on top of a dense-id hot plan.  `groups` = populated ids expected per tile.
// Returns false (P.hot.bucket stays 0) when the shape is not eligible or nothing fits.
// The per-cell geometry already in P.hot (dense ids or the hash index) stays valid: it is what runs when the specialised
// build is not available.  dense_range = ids [0, dense_range) after subtracting P.dense_min.
static bool plan_bucket(ScanPlan& P, int64_t groups, int64_t dense_range, bool sentinels, bool windowed, bool assume_jit) {
  HotGeom& g = P.hot;
  g.bucket = 0;
  static const bool off = getenv("PW_NO_BUCKET") != nullptr;
  // dense_range == 0: INDEXED ids — the CTA maps the key words to ids through a shared-memory index (any key shape)
  const bool indexed = dense_range == 0;
  if (indexed) {
    if (windowed || groups < 48 || groups > 1600 || getenv("PW_NO_BUCKET_INDEX")) return false;
    dense_range = 64;
    while (dense_range < groups + groups / 64 + 8) dense_range <<= 1;   // id capacity: a little head-room over the live-group estimate
                                                                          // (both pilot samples have seen a low-cardinality key set whole)
  }
  if (off || (!assume_jit && !jit_available()) || dense_range < (windowed ? 16 : 48) || dense_range > 2048) return false;
  g.b_range = (int32_t)dense_range; g.b_sent = sentinels ? 1 : 0; g.b_win = windowed ? 1 : 0; g.b_idx = indexed ? 1 : 0;
  if ((P.dyn.enabled != 0) != windowed || P.row_group_out || P.rowid_slot_p1 || P.check_sorted || !P.vec_ok || P.row_begin != 0 || P.row_stride != 1) return false;
  if ((!indexed && P.n_kw != (windowed ? 2 : 1)) || P.n_kw > 6 || P.n_vexpr > 8 || P.n_acc < 1 || (P.gflags & GF_TMIN)) return false;
  const int kw_pad = P.n_kw <= 1 ? 1 : (P.n_kw <= 2 ? 2 : (P.n_kw <= 4 ? 4 : 6));   // the kernel's key-word class (launch_scan)
  // the meta plane ((row << 8) | validity bits per bucketed row) is needed when a value can be null; first / last /
  // first-occurrence words alone are served by two row positions per group and tile (ROWPOS, pw_bucket.cuh)
  bool needs_valid = false, needs_row = (P.gflags & GF_ROW) != 0;
  for (int e = 0; e < P.n_vexpr; ++e) {
    const VExpr& V = P.vexprs[e];
    if (V.flags & (VF_FIRST | VF_LAST | VF_FIRST_NN | VF_LAST_NN)) needs_row = true;
    if (V.n_factors == 0) needs_valid = needs_valid || P.slots[V.slot].validity != nullptr;
    else for (int k = 0; k < V.n_factors; ++k) needs_valid = needs_valid || P.slots[V.f[k].slot].validity != nullptr;
  }
  static const bool no_rowpos = getenv("PW_NO_ROWPOS") != nullptr;
  const bool meta = needs_valid || (needs_row && no_rowpos);
  const bool rowpos = needs_row && !meta;
  g.b_rowpos = rowpos ? 1 : 0;
  const int planes = P.n_vexpr + (meta ? 1 : 0);
  if (planes < 1) return false;   // len-only queries: nothing to bucket, the per-cell counters are already cheap
  int gcap = windowed ? 16 : 64;
  while (gcap < dense_range) gcap <<= 1;
  // bucket depth from the Poisson tail of rows per id per tile: expected share of rows beyond depth J below eps
  auto depth = [&](double lambda, double eps) {
    double p = exp(-lambda), cdf = 0.0, mean_le = 0.0;   // P(X = k), running sums over k <= J
    for (int k = 0; k < 4096; ++k) {
      if (k > 0) p *= lambda / k;
      cdf += p; mean_le += k * p;
      // rows beyond depth k: sum_{x > k} (x - k) p(x) = (lambda - mean_le) - k (1 - cdf)
      const double beyond = (lambda - mean_le) - k * (1.0 - cdf);
      if (k >= 2 && beyond <= eps * lambda) return k;
    }
    return 4096;
  };
  const int64_t pop = std::max<int64_t>(1, std::min<int64_t>(groups > 0 ? groups : dense_range, dense_range));
  // Input staging by bulk async copies (TMA): every slot a plain 1/2/4/8-byte column without validity, 16-byte aligned.
  // The copies of the next `stages` tiles are in flight while a tile is scattered and folded — bytes in flight no
  // longer cost registers or L1 lines (the register pipeline holds one tile = 32 KB per SM in flight on the C2 shape,
  // about 70 % of what the HBM latency-bandwidth product asks for).
  static const int stages_env = getenv("PW_BUCKET_STAGES") ? atoi(getenv("PW_BUCKET_STAGES")) : -1;
  static const int j_env = getenv("PW_BUCKET_J") ? atoi(getenv("PW_BUCKET_J")) : 0;
  int row_bytes = 0;
  bool tma_ok = stages_env != 0;
  for (int c = 0; c < P.n_slots; ++c) {
    const int dt = P.slots[c].dtype;
    const int w = (dt == DT_I8 || dt == DT_U8) ? 1 : (dt == DT_I16 || dt == DT_U16) ? 2 : (dt == DT_I32 || dt == DT_U32 || dt == DT_F32) ? 4 :
                  (dt == DT_I64 || dt == DT_U64 || dt == DT_F64) ? 8 : 0;
    if (w == 0 || P.slots[c].validity != nullptr || ((uintptr_t)P.slots[c].values & 15u)) tma_ok = false;
    row_bytes += w;
  }
  // in order of measured preference on the C2 shape: one 32-warp CTA with two bucket buffers (one barrier per tile),
  // two 16-warp CTAs with one buffer each, then whatever fits
  // (windowed: more live state per thread — two more accumulators, the window bookkeeping — than 64 registers hold:
  // 16 warps with 128 registers each)
  struct Cand { int threads, nbuf, cps; };
  static const Cand cand_plain[5] = {{1024, 2, 1}, {512, 1, 2}, {512, 2, 1}, {1024, 1, 1}, {512, 1, 1}};
  static const Cand cand_win[5] = {{512, 2, 1}, {512, 1, 1}, {1024, 2, 1}, {1024, 1, 1}, {256, 2, 1}};
  const Cand* cand = windowed ? cand_win : cand_plain;
  static const int only = getenv("PW_BUCKET_CAND") ? atoi(getenv("PW_BUCKET_CAND")) : -1;   // experiments: one geometry only
  for (int ci = 0; ci < 5; ++ci) {
    const Cand& cd = cand[ci];
    if (only >= 0 && ci != only % 5) continue;
    const int tile = cd.threads / 32 * 64;
    const int j_full = depth((double)tile / (double)pop, 1e-4), j_min = depth((double)tile / (double)pop, 1e-3);
    const int ncnt = gcap < cd.threads ? cd.nbuf + 1 : cd.nbuf;
    const size_t ovf = (16 + (size_t)cd.nbuf * 32 * (4 + 8 * planes) + 127) & ~(size_t)127;   // overflow list (pw_bucket.cuh OVF_BYTES)
    static const int idx_mul = getenv("PW_BUCKET_IDXMUL") ? atoi(getenv("PW_BUCKET_IDXMUL")) : 8;   // measured on C2 without dense ids: 8 -> 0.73 ms, 4 -> 1.01 ms (old per-cell table 0.91 ms)
    g.b_idx_mul = idx_mul;
    const size_t idx_bytes = indexed ? (size_t)idx_mul * gcap * 4 + (size_t)gcap * kw_pad * 8 + 16 : 0;   // key index (pw_bucket.cuh IDX_BYTES)
    const size_t pos_bytes = rowpos ? (size_t)cd.nbuf * 2 * gcap * 4 : 0;   // ROWPOS: first / last position per id and buffer
    const size_t fixed = (size_t)ncnt * gcap * 4 + 128 + ovf + idx_bytes + pos_bytes;   // counters + one dummy counter per lane + overflow list + index
    const size_t per_j = (size_t)cd.nbuf * planes * gcap * 8;
    if (tma_ok && cd.cps == 1) {
      // staged: shared memory holds the buckets AND the tiles in flight; L1 is not needed for the stream
      const size_t stage_bytes = (size_t)tile * row_bytes, limit = 226 * 1024;
      for (int stages = stages_env > 0 ? stages_env : 3; stages >= (stages_env > 0 ? stages_env : 2); --stages) {
        const size_t tail = (size_t)stages * stage_bytes + (size_t)stages * 8;
        // full-depth buckets with three tiles in flight if that fits, else two tiles in flight and whatever depth is left
        const int j_need = stages > 2 && stages_env <= 0 ? j_full : j_min;
        if (fixed + tail + per_j * j_need > limit) continue;
        int J = (int)std::min<size_t>((size_t)j_full, (limit - fixed - tail) / per_j);
        if (j_env > 0) J = j_env;
        const size_t bytes = per_j * J + fixed + tail;
        if (bytes > limit) continue;
        g.bucket = 1; g.b_threads = cd.threads; g.b_gcap = gcap; g.b_j = J; g.b_nbuf = cd.nbuf; g.b_stages = stages; g.b_meta = meta ? 1 : 0;
        g.b_cps = 1; g.b_stage_bytes = (int32_t)stage_bytes; g.b_bytes = (int32_t)bytes;
        return true;
      }
    }
    const size_t budget = 172 * 1024;   // per SM; leaves ~56 KB of L1 for the loads in flight (see pw_bucket.cuh)
    const int J = j_env > 0 ? j_env : j_full;
    const size_t bytes = per_j * J + fixed;
    if (bytes * cd.cps > budget) continue;
    g.bucket = 1; g.b_threads = cd.threads; g.b_gcap = gcap; g.b_j = J; g.b_nbuf = cd.nbuf; g.b_stages = 0; g.b_meta = meta ? 1 : 0;
    g.b_cps = cd.cps; g.b_stage_bytes = 0;
    g.b_bytes = (int32_t)bytes;
    return true;
  }
  return false;
}
}  // namespace pw
// Diagnostics: NVRTC-compiles the specialised scan kernel for a C2-shaped plan (int64 key, f64 value,
// sum/mean/min/max).  Needs no GPU; used by build() as the "does the JIT path build" check.
extern "C" __attribute__((visibility("default"))) int pw_b200_jit_selftest(char* log, size_t log_len) {
  using namespace pw;
  ScanPlan P;
  memset(&P, 0, sizeof P);
  P.n_rows = 1 << 20; P.row_stride = 1; P.n_slots = 2; P.vec_ok = 1;
  P.slots[0].dtype = DT_I64; P.slots[1].dtype = DT_F64;
  P.n_keys = 1; P.keys[0].slot = 0; P.keys[0].n_words = 1; P.keys[0].dtype = DT_I64; P.n_kw = 1;
  P.n_vexpr = 1; P.vexprs[0].slot = 1; P.vexprs[0].cls = CLS_F64; P.vexprs[0].flags = VF_SUM_F | VF_MIN | VF_MAX; P.vexprs[0].acc_base = 0;
  P.accs[0].op = OP_ADD_F64; P.accs[0].src = SRC_F64; P.accs[1].op = OP_MIN_I64; P.accs[1].src = SRC_F64_ORD;
  P.accs[2].op = OP_MAX_I64; P.accs[2].src = SRC_F64_ORD; P.accs[3].op = OP_ADD_I64; P.accs[3].src = SRC_ONE;
  P.n_acc = 4; P.gflags = GF_LEN; P.acc_gbase = 3;
  const bool win = getenv("PW_SELFTEST_WIN") != nullptr;   // OHLCV shape: (time, symbol, price) -> windowed bucket tier
  if (win) {
    P.n_slots = 3; P.slots[2].dtype = DT_I64;   // slot 2 = time
    P.dyn.enabled = 1; P.dyn.slot = 2; P.dyn.every = P.dyn.period = 60; P.dyn.closed = 0; P.n_kw = 2;
    P.keys[0].dtype = DT_I32; P.slots[0].dtype = DT_I32;
    P.vexprs[0].flags = VF_SUM_F | VF_MIN | VF_MAX | VF_FIRST | VF_LAST;
    P.accs[3].op = OP_MIN_U64; P.accs[3].src = SRC_ROW; P.accs[4].op = OP_MAX_U64; P.accs[4].src = SRC_ROW;
    P.accs[5].op = OP_ADD_I64; P.accs[5].src = SRC_ONE; P.n_acc = 6; P.acc_gbase = 5;
  }
  if (getenv("PW_SELFTEST_RADIX")) {   // C3 shape over the radix tier's records: (key, value, row id), null-aware sum/count/min/max/first/last
    P.n_slots = 3; P.slots[2].dtype = DT_U64; P.rowid_slot_p1 = 3;
    P.vexprs[0].flags = VF_SUM_F | VF_COUNT | VF_MIN | VF_MAX | VF_FIRST | VF_LAST;
    P.accs[1].op = OP_ADD_I64; P.accs[1].src = SRC_VALID; P.accs[2].op = OP_MIN_I64; P.accs[2].src = SRC_F64_ORD;
    P.accs[3].op = OP_MAX_I64; P.accs[3].src = SRC_F64_ORD; P.accs[4].op = OP_MIN_U64; P.accs[4].src = SRC_ROWIDX;
    P.accs[5].op = OP_MAX_U64; P.accs[5].src = SRC_ROWIDX; P.n_acc = 6; P.gflags = 0; P.acc_gbase = 6;
  }
  if (!plan_hot(P, 1000, 0, getenv("PW_SELFTEST_DENSE") && !win ? 1000 : 0)) return -1;
  P.hot_slots = P.hot.idx_slots;
  if (win) {
    if (!plan_bucket(P, 100, 100, false, true, true)) return -4;
  } else if (getenv("PW_SELFTEST_INDEX")) {   // ids from the shared-memory key index
    if (!plan_bucket(P, 1000, 0, false, false, true)) return -5;
  } else if (getenv("PW_SELFTEST_DENSE")) {   // the bucket tier on top of the dense ids (NVRTC exists whenever this function can succeed)
    if (!plan_bucket(P, 1000, 1000, false, false, true) && !getenv("PW_NO_BUCKET")) return -3;
  }
  std::string err;
  if (getenv("PW_SELFTEST_THREADS")) {  // the geometry a 16-warp CTA would get
    if (!plan_hot_threads(P, 1000, P.hot.gcap, getenv("PW_SELFTEST_DENSE") ? 1000 : 0, atoi(getenv("PW_SELFTEST_THREADS")), true)) return -2;
    P.hot_slots = P.hot.idx_slots;
  }
  const int rc = jit_selftest_compile(P, 4, win ? 2 : 1, true, P.hot.threads, &err);
  if (log && log_len) { strncpy(log, err.c_str(), log_len - 1); log[log_len - 1] = 0; }
  return rc;
}
// Diagnostics without a GPU (tests/test_abi_cpu.py): the multiply-shift window division against native division, and
// the result-buffer pool's fall-back to pageable memory when nothing can be pinned.  Returns the number of mismatches.
extern "C" __attribute__((visibility("default"))) int64_t pw_b200_host_selftest(void) {
  using namespace pw;
  int64_t bad = 0;
  const uint64_t ds[] = {1, 2, 3, 7, 10, 60, 1000, 60000000ull, 86400000000ull, 3600000000000ull, (1ull << 40) + 1, (1ull << 62) + 12345,
                         0x7FFFFFFFFFFFFFFFull, 999999937ull, 641, 4294967297ull};
  uint64_t x = 0x9E3779B97F4A7C15ull;
  for (uint64_t d : ds) {
    uint64_t magic; int32_t more;
    div_prepare(d, &magic, &more);
    for (int i = 0; i < 200000; ++i) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;   // xorshift64
      uint64_t n = x >> (i % 61);
      if (i % 7 == 0) n = d * (n % 1000) + (uint64_t)(i % 3) - 1;   // around multiples of d
      if (i == 0) n = ~0ull;
      if (i == 1) n = 0;
      if (div_apply(n, magic, more) != n / d) ++bad;
    }
  }
  void* p = host_alloc(3u << 20);
  if (!p) ++bad; else { memset(p, 0x5A, 3u << 20); host_free(p); }
  void* q2 = host_alloc(3u << 20);   // a pinned block would come back from the pool; a pageable one is simply fresh
  if (!q2) ++bad; else host_free(q2);
  void* small = host_alloc(100);
  if (!small) ++bad; else host_free(small);
  return bad;
}
namespace pw {

// value range of a single integer key column over a row sample (row = begin + i * stride): decides whether
// dense ids apply (the analogue of a perfect-hash / direct-address aggregate over a small key domain)
__global__ void key_range_kernel(RawSlot key, int64_t begin, int64_t stride, int64_t n, unsigned long long* kmax_u, unsigned long long* kmin_n) {
  unsigned long long hi = 0, lo = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = begin + i * stride;
    if (key.validity) {
      const int64_t b = (int64_t)key.bit_offset + row;
      if (!((key.validity[b >> 3] >> (b & 7)) & 1)) continue;
    }
    int64_t v;
    switch (key.dtype) {
      case DT_I8: v = ((const int8_t*)key.values)[row]; break;
      case DT_U8: v = ((const uint8_t*)key.values)[row]; break;
      case DT_I16: v = ((const int16_t*)key.values)[row]; break;
      case DT_U16: v = ((const uint16_t*)key.values)[row]; break;
      case DT_I32: v = ((const int32_t*)key.values)[row]; break;
      case DT_U32: v = ((const uint32_t*)key.values)[row]; break;
      default: v = ((const int64_t*)key.values)[row]; break;
    }
    const unsigned long long u = (unsigned long long)v ^ 0x8000000000000000ull;
    hi = u > hi ? u : hi;
    lo = ~u > lo ? ~u : lo;
  }
  for (int o = 16; o; o >>= 1) {
    const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, o), l2 = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = h2 > hi ? h2 : hi;
    lo = l2 > lo ? l2 : lo;
  }
  if ((threadIdx.x & 31) == 0 && (hi | lo)) { atomicMax(kmax_u, hi); atomicMax(kmin_n, lo); }
}
// may this query use dense ids at all?  One plain signed (or narrow unsigned) integer key, a per-group row counter
static bool dense_eligible(const PwQuery* q, const ScanPlan& P) {
  if (getenv("PW_NO_DENSE") || (q->flags & PW_FLAG_NO_DENSE_IDS)) return false;
  if (q->n_keys != 1 || P.dyn.enabled || P.n_kw != 1 || !(P.gflags & GF_LEN) || P.row_group_out) return false;
  switch (P.keys[0].dtype) {
    case DT_I8: case DT_U8: case DT_I16: case DT_U16: case DT_I32: case DT_U32: case DT_I64: return true;
    default: return false;
  }
}

// group_by_dynamic by ONE plain integer key (OHLCV bars by symbol): the windowed form of the bucket tier needs that key's
// value range from the pilot; the group key is then (id, window)
static bool wbucket_eligible(const PwQuery* q, const ScanPlan& P) {
  if (getenv("PW_NO_DENSE") || getenv("PW_NO_WBUCKET") || (q->flags & (PW_FLAG_NO_DENSE_IDS | PW_FLAG_NO_BUCKETS))) return false;
  if (q->n_keys != 1 || !P.dyn.enabled || P.n_kw != 2 || P.has_null_word || P.row_group_out || (P.gflags & GF_TMIN)) return false;
  if (P.slots[P.keys[0].slot].validity != nullptr) return false;
  switch (P.keys[0].dtype) {
    case DT_I8: case DT_U8: case DT_I16: case DT_U16: case DT_I32: case DT_U32: case DT_I64: return true;
    default: return false;
  }
}

// Array-of-structs hash table: one row [key words | accumulator words | pad] per slot, padded to 4/8/16/32 words
int alloc_table_raw(Table* T, int n_kw, int n_acc, uint64_t cap, int32_t* overflow, unsigned long long* spilled) {
  const uint64_t n = cap + 2;
  int row = 4;
  while (row < n_kw + n_acc) row <<= 1;
  void* p = nullptr;
  PW_TRY(dev_alloc(&p, n * 8 * (uint64_t)row));
  T->keys = (uint64_t*)p;
  T->accs = T->keys + n_kw;
  PW_TRY(dev_alloc(&p, n * 4)); T->state = (uint32_t*)p;
  T->cap = cap;
  T->key_sw = T->acc_sw = 1; T->key_ss = T->acc_ss = (uint64_t)row;
  T->overflow = overflow;
  T->spilled = spilled;
  return 0;
}
static int alloc_table(Table* T, int n_kw, int n_acc, uint64_t cap, Control* dctl) {
  return alloc_table_raw(T, n_kw, n_acc, cap, &dctl->overflow, &dctl->spilled);
}
void free_table(Table& T) {
  const bool aos = T.key_ss != 1;  // accs points into the row buffer
  dev_free(T.keys); dev_free(T.state);
  if (!aos) dev_free(T.accs);
  T.keys = nullptr; T.state = nullptr; T.accs = nullptr;
}
static int init_table(const Table& T, const ScanPlan& P, cudaStream_t st, Control* zero = nullptr) {
  AccOps ops{};
  ops.n = P.n_acc;
  for (int a = 0; a < P.n_acc; ++a) ops.op[a] = P.accs[a].op;
  const uint64_t n = T.cap + 2;
  int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 8);
  table_init_kernel<<<grid, 256, 0, st>>>(T, padded_kw(P.n_kw), ops, zero);
  PW_CUDA(cudaGetLastError());
  ctx().timings.kernel_launches++;
  return 0;
}

// distinct keys among `n_sample` rows starting at row_begin with the given stride (pilot launch)
// distinct keys among `n_sample` rows starting at row_begin with the given stride (pilot launch).  Launch and
// collect are split so that both pilots of a query run back to back with ONE host synchronisation.
struct Pilot {
  Table T{};
  Control* dctl = nullptr;
  uint32_t* dummy = nullptr;
};
static int pilot_launch(const Lowered& L, int64_t row_begin, int64_t stride, int64_t n_sample, Pilot* pl) {
  ThreadCtx& c = ctx();
  ScanPlan P = L.plan;
  P.n_rows = n_sample; P.row_begin = row_begin; P.row_stride = stride;
  P.vec_ok = (L.plan.vec_ok && stride == 1 && (row_begin % 2) == 0) ? 1 : 0;
  P.hot_slots = 0; P.check_sorted = 0; P.n_preds = 0;
  // only group identity matters: one len accumulator
  P.n_acc = 1; P.accs[0].op = OP_ADD_I64; P.accs[0].src = SRC_ONE; P.accs[0].vexpr = 0; P.n_vexpr = 0;
  P.gflags = GF_LEN; P.acc_gbase = 0;
  void* p = nullptr;
  PW_TRY(dev_alloc(&p, sizeof(Control))); pl->dctl = (Control*)p;
  PW_CUDA(cudaMemsetAsync(pl->dctl, 0, sizeof(Control), c.stream));
  PW_TRY(alloc_table(&pl->T, padded_kw(P.n_kw), 1, (uint64_t)n_sample * 2 + 64, pl->dctl));
  PW_TRY(init_table(pl->T, P, c.stream));
  P.table = pl->T; P.not_sorted = &pl->dctl->not_sorted;
  PW_TRY(launch_scan(P, c.sm_count, c.stream));
  PW_TRY(dev_alloc(&p, (pl->T.cap + 2) * 4)); pl->dummy = (uint32_t*)p;
  int grid = (int)std::min<uint64_t>((pl->T.cap + 2 + 255) / 256, 148 * 8);
  compact_kernel<<<grid, 256, 0, c.stream>>>(pl->T, padded_kw(P.n_kw), pl->dummy, &pl->dctl->counter);
  PW_CUDA(cudaGetLastError());
  c.timings.kernel_launches++;
  return 0;
}
static void pilot_free(Pilot* pl) {
  dev_free(pl->dummy); dev_free(pl->dctl);
  if (pl->T.keys) free_table(pl->T);
}

// d distinct values seen in a uniform sample of n rows -> number of groups, assuming equally likely
// groups: d = G (1 - exp(-n/G)).  (The reference extrapolates from a sqrt(N) sample the same way,
// group_by_streaming.rs:114-153.)
static double solve_groups(double d, double n) {
  if (d >= 0.995 * n) return 1e18;
  double lo = d, hi = 1e15;
  for (int i = 0; i < 200; ++i) {
    double mid = sqrt(lo * hi);
    double seen = mid * (1.0 - exp(-n / mid));
    if (seen < d) lo = mid; else hi = mid;
    if (hi / lo < 1.0001) break;
  }
  return hi;
}

// order the compacted group list by the plan's sort words (LSD radix passes, least significant first)
int order_groups(const Lowered& L, const Table& T, int kw, uint32_t** slots_io, uint64_t G, const unsigned long long* g_dev) {
  ThreadCtx& c = ctx();
  PwTimings& tm = c.timings;
  uint32_t* slots = *slots_io;
  if (G > 1 && !L.sort.empty()) {
    uint64_t *k_in = nullptr, *k_out = nullptr;
    uint32_t *v_out = nullptr;
    void* p = nullptr;
    PW_TRY(dev_alloc(&p, G * 8)); k_in = (uint64_t*)p;
    PW_TRY(dev_alloc(&p, G * 8)); k_out = (uint64_t*)p;
    PW_TRY(dev_alloc(&p, G * 4)); v_out = (uint32_t*)p;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, slots, v_out, (int64_t)G, 0, 64, c.stream);
    void* tmp = nullptr;
    PW_TRY(dev_alloc(&tmp, tmp_bytes));
    const int grid = (int)((G + 255) / 256);
    for (const SortSpec& sp : L.sort) {  // LSD: least significant word first, every pass stable
      sort_key_kernel<<<grid, 256, 0, c.stream>>>(T, kw, L.null_word, sp, slots, G, k_in, g_dev);
      PW_CUDA(cudaGetLastError());
      PW_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, slots, v_out, (int64_t)G, 0, 64, c.stream));
      std::swap(slots, v_out);
      tm.kernel_launches += 2;
    }
    dev_free(tmp); dev_free(k_in); dev_free(k_out); dev_free(v_out);
  }
  *slots_io = slots;
  return 0;
}

// ---- high-cardinality tier: radix-partition the rows by key hash, then scan the partitioned copy (pw_partition.cuh)
struct PartTemp {
  uint64_t* words = nullptr;
  uint32_t* hist = nullptr;
  uint32_t* cursor = nullptr;
  void* scan_tmp = nullptr;
};
static void part_free(PartTemp& t) { dev_free(t.words); dev_free(t.hist); dev_free(t.cursor); dev_free(t.scan_tmp); t = PartTemp{}; }
static bool part_eligible(const PwQuery* q, const ScanPlan& P) {
  if ((q->flags & PW_FLAG_NO_PARTITION) || getenv("PW_NO_PARTITION") || !jit_available()) return false;
  if (P.dyn.enabled || P.row_group_out || P.n_slots + 1 > 4 || !narrow_class(P) || P.n_slots > 8) return false;
  if (P.n_rows >= ((int64_t)1 << 32) - 64) return false;  // 32-bit partition cursors: larger inputs keep the plain HBM-table path
  for (int s = 0; s < P.n_slots; ++s)
    if (P.slots[s].dtype == DT_VIEW || P.slots[s].dtype == DT_VIEW_HI || P.slots[s].dtype == DT_BOOL) return false;
  return true;
}
// Builds the partitioned copy of the rows that pass the predicate and the plan P2 that scans it.  g_hint = expected
// number of groups.  Returns PW_OK with P2->n_rows = surviving rows.
static int partition_input(const ScanPlan& P, double g_hint, ScanPlan* P2out, PartTemp* tmp) {
  ThreadCtx& c = ctx();
  const int64_t N = P.n_rows;
  if (N >= (int64_t)1 << 32) return fail(PW_ERR_UNSUPPORTED, "partitioned path: more than 2^32 rows per device");
  ScanPlan P2 = P;
  const uint64_t stride = ((uint64_t)N + 31) & ~(uint64_t)31;
  { void* p = nullptr; PW_TRY(dev_alloc(&p, (size_t)(P.n_slots + 1) * stride * 8)); tmp->words = (uint64_t*)p; }
  for (int s = 0; s < P.n_slots; ++s) {
    RawSlot& d = P2.slots[s];
    d.values = tmp->words + (uint64_t)s * stride;
    d.validity = nullptr; d.bit_offset = 0;
    const int cls = dtype_class(P.slots[s].dtype);
    d.dtype = cls == CLS_F64 ? DT_F64 : (cls == CLS_U64 ? DT_U64 : DT_I64);  // canonical 64-bit image of the class
  }
  RawSlot& rid = P2.slots[P.n_slots];
  rid.values = tmp->words + (uint64_t)P.n_slots * stride; rid.validity = nullptr; rid.bit_offset = 0; rid.dtype = DT_U64;
  P2.n_slots = P.n_slots + 1;
  P2.rowid_slot_p1 = P.n_slots + 1;
  for (int k = 0; k < P2.n_keys; ++k) P2.keys[k].dtype = P2.slots[P2.keys[k].slot].dtype;
  P2.n_preds = 0; P2.check_sorted = 0; P2.vec_ok = 1; P2.row_begin = 0; P2.row_stride = 1;
  // partitions: about an eighth of the hot table's ids each, so that several of them are resident side by side
  if (!plan_hot(P2, 4096, 0)) return 1;  // no hot table fits: caller keeps the plain HBM-table path
  static const double part_div = getenv("PW_PART_DIV") ? atof(getenv("PW_PART_DIV")) : 8.0;  // measured on C3: 4 -> 17.5 ms, 8 -> 16.7 ms, 32 -> 18.1 ms (scan phase)
  const double per_part = std::max(8.0, (double)P2.hot.gcap / part_div);
  uint64_t n_parts = (uint64_t)std::min(4.0e6, std::max(4.0, ceil(std::max(g_hint, 1.0) * 1.1 / per_part)));
  if (getenv("PW_DEBUG"))
    fprintf(stderr, "[pw] partition: rows=%lld groups~%.0f parts=%llu gcap=%d S=%d R=%d n_mm=%d smem=%d threads=%d\n", (long long)N, g_hint,
            (unsigned long long)n_parts, P2.hot.gcap, P2.hot.idx_slots, P2.hot.replicas, P2.hot.n_mm, P2.hot.total_bytes, P2.hot.threads);
  { void* p = nullptr; PW_TRY(dev_alloc(&p, n_parts * 4)); tmp->hist = (uint32_t*)p; }
  { void* p = nullptr; PW_TRY(dev_alloc(&p, n_parts * 4)); tmp->cursor = (uint32_t*)p; }
  PW_CUDA(cudaMemsetAsync(tmp->hist, 0, n_parts * 4, c.stream));
  PartParams pp{};
  pp.hist = tmp->hist; pp.cursor = tmp->cursor; pp.out = tmp->words; pp.out_stride = stride; pp.n_parts = (uint32_t)n_parts;
  const int kwc = padded_kw(P.n_kw) <= 1 ? 1 : (padded_kw(P.n_kw) <= 2 ? 2 : (padded_kw(P.n_kw) <= 4 ? 4 : 6));
  ScanPlan PA = P;  // passes 1 and 2 read the original frame
  PA.n_kw = padded_kw(P.n_kw); PA.hot_slots = 0; memset(&PA.hot, 0, sizeof PA.hot);
  pp.mode = 1;
  if (int rc = launch_part_jit(PA, pp, 4, kwc, c.sm_count, c.stream)) return rc;
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, tmp->hist, tmp->cursor, (int)n_parts, c.stream);
  PW_TRY(dev_alloc(&tmp->scan_tmp, tb));
  PW_CUDA(cub::DeviceScan::ExclusiveSum(tmp->scan_tmp, tb, tmp->hist, tmp->cursor, (int)n_parts, c.stream));
  c.timings.kernel_launches++;
  int64_t n2 = N;
  if (P.n_preds > 0 || P.n_keys == 0) {  // rows dropped by the predicate never reach the copy: count them
    uint32_t last[2] = {0, 0};
    PW_CUDA(cudaMemcpyAsync(&last[0], tmp->cursor + n_parts - 1, 4, cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaMemcpyAsync(&last[1], tmp->hist + n_parts - 1, 4, cudaMemcpyDeviceToHost, c.stream));
    PW_CUDA(cudaStreamSynchronize(c.stream));
    n2 = (int64_t)last[0] + (int64_t)last[1];
  }
  pp.mode = 2;
  if (int rc = launch_part_jit(PA, pp, 4, kwc, c.sm_count, c.stream)) return rc;
  P2.n_rows = n2;
  *P2out = P2;
  return 0;
}

// ---- high-cardinality tier, second form (strategy 10): two-level radix partition with staged writes + one CTA per
// partition aggregating in shared memory (pw_radix.cuh)
struct RadixTemp {
  uint4* buf_a = nullptr;
  uint4* buf_b = nullptr;
  uint32_t* small = nullptr;   // hist[P] | offs[P + 1] | cursor1[256] | cursor2[P] | tile_first[257] | pad | dense_count (u64)
  unsigned long long* dense_count = nullptr;
  size_t smem3 = 0;
  int kwc = 1;
};
static void radix_free(RadixTemp& t) { dev_free(t.buf_a); dev_free(t.buf_b); dev_free(t.small); t = RadixTemp{}; }
static bool radix_eligible(const PwQuery* q, const ScanPlan& P) {
  static const bool off = getenv("PW_NO_RADIX") != nullptr;
  return !off && part_eligible(q, P) && P.n_slots <= 3 && P.n_acc >= 1;
}
// exclusive prefix of the partition histogram, the cursors of both levels and the tile list of the level-2 pass
static __global__ void __launch_bounds__(1024) radix_setup_kernel(const uint32_t* hist, uint32_t* offs, uint32_t* cursor1, uint32_t* cursor2,
                                                                  uint32_t* tile_first, uint32_t n_parts, uint32_t log2_p2) {
  __shared__ uint32_t wsum[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t chunk = (n_parts + 1023u) / 1024u;
  const uint32_t b = min(tid * chunk, n_parts), e = min(b + chunk, n_parts);
  uint32_t sum = 0;
  for (uint32_t i = b; i < e; ++i) sum += hist[i];
  uint32_t incl = sum;
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += t; }
  if (lane == 31u) wsum[warp] = incl;
  __syncthreads();
  uint32_t off = 0;
  for (uint32_t w = 0; w < warp; ++w) off += wsum[w];
  uint32_t run = off + incl - sum;
  for (uint32_t i = b; i < e; ++i) { offs[i] = run; cursor2[i] = run; run += hist[i]; }
  if (tid == 1023u) offs[n_parts] = off + incl;
  __syncthreads();
  const uint32_t n_l1 = n_parts >> log2_p2;  // <= 256
  uint32_t t = 0;
  if (tid < n_l1) {
    const uint32_t first = offs[tid << log2_p2], last = offs[(tid + 1u) << log2_p2];
    cursor1[tid] = first;
    t = (last - first + (uint32_t)RADIX_TILE - 1u) / (uint32_t)RADIX_TILE;
  }
  incl = t;
  for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += u; }
  if (lane == 31u) wsum[warp] = incl;
  __syncthreads();
  off = 0;
  for (uint32_t w = 0; w < warp; ++w) off += wsum[w];
  if (tid < n_l1) tile_first[tid] = off + incl - t;
  if (tid == n_l1 - 1u) tile_first[n_l1] = off + incl;
}
// Partitions the rows that pass the predicate into 32-byte records grouped by final partition.  Returns 0 (done: *P2out
// scans records, *rp carries the partition table and the record buffer as `src`), 1 (not applicable: caller keeps
// another tier), < 0 error.
static int radix_prepare(const ScanPlan& P, double g_hint, ScanPlan* P2out, RadixParams* rp, RadixTemp* tmp) {
  ThreadCtx& c = ctx();
  const int64_t N = P.n_rows;
  const int kw = padded_kw(P.n_kw);
  const int kwc = kw <= 1 ? 1 : (kw <= 2 ? 2 : (kw <= 4 ? 4 : 6));
  tmp->kwc = kwc;
  // shared memory of the aggregation pass: the staging area of one round + as many table slots as the rest holds.
  // Two 512-thread CTAs per SM with half-size tables (one CTA's barriers and latencies hide behind the other's work)
  // or one 1024-thread CTA: PW_RADIX_AGG_THREADS
  const char* at_env = getenv("PW_RADIX_AGG_THREADS");
  const int agg_threads = at_env && atoi(at_env) == 1024 ? 1024 : 512;
  const size_t slot_bytes = (size_t)(kwc + P.n_acc) * 8 + 8 + 2 + 2 + (kwc > 1 ? 4 : 0);
  const size_t stage_bytes = (size_t)(P.n_slots + 1) * 8 * RADIX_ROUND * agg_threads;
  const size_t budget = (agg_threads == 1024 ? 220 : 110) * 1024 - 512;
  if (stage_bytes + (slot_bytes << 6) > budget) return 1;
  int log2_slots = agg_threads == 1024 ? 12 : 11;   // at most four slots per thread
  while (log2_slots > 6 && stage_bytes + (slot_bytes << log2_slots) > budget) --log2_slots;
  const double per_part = 0.4 * (double)(1u << log2_slots);   // groups per partition the table is planned for
  const double want = std::max(g_hint, 1.0) * 1.1 / per_part;
  int log2_parts = 1;
  while (log2_parts < 15 && (double)(1u << log2_parts) < want) ++log2_parts;
  if (const char* e = getenv("PW_RADIX_LOG2_PARTS")) log2_parts = std::max(1, std::min(15, atoi(e)));   // tests: both levels on small inputs
  else if ((double)(1u << log2_parts) * per_part * 1.75 < std::max(g_hint, 1.0)) return 1;  // more groups than 2^15 tables hold
  const int log2_p1 = log2_parts <= 8 ? log2_parts : (log2_parts + 1) / 2;
  const int log2_p2 = log2_parts - log2_p1;
  const uint32_t n_parts = 1u << log2_parts;
  if (getenv("PW_DEBUG"))
    fprintf(stderr, "[pw] radix tier: rows=%lld groups~%.0f parts=2^%d (2^%d x 2^%d) table slots=2^%d (%zu B each) accs=%d\n", (long long)N, g_hint, log2_parts,
            log2_p1, log2_p2, log2_slots, slot_bytes, P.n_acc);
  ScanPlan P2 = P;
  for (int s = 0; s < P.n_slots; ++s) {
    RawSlot& d = P2.slots[s];
    d.values = nullptr; d.validity = nullptr; d.bit_offset = 0;
    const int cls = dtype_class(P.slots[s].dtype);
    d.dtype = cls == CLS_F64 ? DT_F64 : (cls == CLS_U64 ? DT_U64 : DT_I64);  // canonical 64-bit image of the class
  }
  RawSlot& rid = P2.slots[P.n_slots];
  rid.values = nullptr; rid.validity = nullptr; rid.bit_offset = 0; rid.dtype = DT_U64;
  P2.n_slots = P.n_slots + 1;
  P2.rowid_slot_p1 = P.n_slots + 1;
  for (int k = 0; k < P2.n_keys; ++k) P2.keys[k].dtype = P2.slots[P2.keys[k].slot].dtype;
  P2.n_preds = 0; P2.check_sorted = 0; P2.vec_ok = 1; P2.row_begin = 0; P2.row_stride = 1;
  P2.n_kw = kw; P2.hot_slots = 0; memset(&P2.hot, 0, sizeof P2.hot);
  ScanPlan PA = P;  // the histogram and the first scatter read the original frame
  PA.n_kw = kw; PA.hot_slots = 0; memset(&PA.hot, 0, sizeof PA.hot);

  const size_t n_small = (size_t)n_parts * 2 + 1 + 256 + n_parts + 257 + 8;
  { void* p = nullptr; PW_TRY(dev_alloc(&p, n_small * 4)); tmp->small = (uint32_t*)p; }
  uint32_t* hist = tmp->small;
  uint32_t* offs = hist + n_parts;
  uint32_t* cursor1 = offs + n_parts + 1;
  uint32_t* cursor2 = cursor1 + 256;
  uint32_t* tile_first = cursor2 + n_parts;
  tmp->dense_count = (unsigned long long*)(((uintptr_t)(tile_first + 257) + 7) & ~(uintptr_t)7);
  const size_t rec_bytes = ((size_t)N + 64) * 32;
  {
    // the record copies (one per level) must fit next to the frame and the result table: otherwise another tier
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && rec_bytes * (log2_p2 > 0 ? 2 : 1) > free_b / 10 * 7) return 1;
  }
  { void* p = nullptr; PW_TRY(dev_alloc(&p, rec_bytes)); tmp->buf_a = (uint4*)p; }
  if (log2_p2 > 0) { void* p = nullptr; PW_TRY(dev_alloc(&p, rec_bytes)); tmp->buf_b = (uint4*)p; }
  PW_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_parts * 4, c.stream));
  RadixParams r{};
  r.log2_parts = log2_parts; r.log2_p2 = log2_p2; r.log2_slots = log2_slots; r.probe_limit = 32; r.agg_threads = agg_threads;
  r.hist = hist; r.offs = offs; r.tile_first = tile_first;
  const int64_t n_steps = (N + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  const int64_t n_tiles = (N + RADIX_TILE - 1) / RADIX_TILE;
  r.mode = 0;
  if (int rc = launch_radix_jit(PA, r, 4, kwc, (size_t)n_parts * 4, (n_steps + 7) / 8, c.sm_count, c.stream)) return rc;
  radix_setup_kernel<<<1, 1024, 0, c.stream>>>(hist, offs, cursor1, cursor2, tile_first, n_parts, (uint32_t)log2_p2);
  PW_CUDA(cudaGetLastError());
  c.timings.kernel_launches++;
  r.mode = 1; r.cursor = cursor1; r.dst = tmp->buf_a;
  if (int rc = launch_radix_jit(PA, r, 4, kwc, RADIX_SCATTER_SMEM, n_tiles, c.sm_count, c.stream)) return rc;
  const uint4* records = tmp->buf_a;
  if (log2_p2 > 0) {
    r.mode = 2; r.cursor = cursor2; r.src = tmp->buf_a; r.dst = tmp->buf_b;
    if (int rc = launch_radix_jit(P2, r, 4, kwc, RADIX_SCATTER_SMEM, n_tiles + (1 << log2_p1), c.sm_count, c.stream)) return rc;
    records = tmp->buf_b;
  }
  r.mode = 3; r.src = records; r.dst = nullptr; r.cursor = nullptr;
  tmp->smem3 = stage_bytes + ((size_t)slot_bytes << log2_slots) + 512;
  *rp = r;
  *P2out = P2;
  return 0;
}
// group list of the radix tier: occupied slots of the overflow region and the escape slots, then the dense region
// (its first *dense_count slots are groups by construction: no occupancy marker to initialise or to read)
static __global__ void radix_compact_kernel(Table T, int n_kw, uint64_t ovf_cap, const unsigned long long* dense_count, uint64_t dense_cap,
                                            uint32_t* slot_list, unsigned long long* counter) {
  const uint64_t nd = min((uint64_t)*dense_count, dense_cap);
  const uint64_t n = ovf_cap + 2 + nd;
  const int lane = threadIdx.x & 31;
  for (uint64_t s0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; s0 < n; s0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = s0 + lane;
    uint64_t s = 0;
    bool occ = false;
    if (i < ovf_cap) { s = i; occ = slot_occupied(T, n_kw, s); }
    else if (i < ovf_cap + 2) { s = T.cap + (i - ovf_cap); occ = slot_occupied(T, n_kw, s); }
    else if (i < n) { s = ovf_cap + (i - ovf_cap - 2); occ = true; }
    const uint32_t m = __ballot_sync(0xffffffffu, occ);
    if (m == 0u) continue;
    unsigned long long basei = 0;
    if (lane == 0) basei = atomicAdd(counter, (unsigned long long)__popc(m));
    basei = __shfl_sync(0xffffffffu, basei, 0);
    if (occ) slot_list[basei + __popc(m & ((1u << lane) - 1u))] = (uint32_t)s;
  }
}
// table split of the aggregation pass: [overflow region | dense region]
static void radix_split(uint64_t cap, RadixParams* rp) {
  rp->ovf_cap = std::max<uint64_t>(std::min<uint64_t>(8192, cap / 2), cap / 16);
  rp->dense_cap = cap - rp->ovf_cap;
}

// ---- result block layout: [header][values 0][validity 0][values 1] ... every piece 256-byte aligned
struct BlockLayout {
  std::vector<size_t> val_bytes, valid_bytes, val_off, valid_off;
  size_t total = 0;
};
static void plan_block(const Lowered& L, uint64_t G, size_t header_bytes, BlockLayout* b) {
  const size_t ncol = L.outs.size();
  b->val_bytes.resize(ncol); b->valid_bytes.resize(ncol); b->val_off.resize(ncol); b->valid_off.resize(ncol);
  size_t total = (header_bytes + 255) / 256 * 256;
  auto place = [&](size_t bytes) { size_t o = total; total += (bytes + 255) / 256 * 256; return o; };
  for (size_t i = 0; i < ncol; ++i) {
    b->val_bytes[i] = out_col_bytes(L.outs[i].out_dtype, G);
    b->valid_bytes[i] = ((G + 31) / 32) * 4;
    b->val_off[i] = place(b->val_bytes[i]);
    b->valid_off[i] = place(b->valid_bytes[i]);
  }
  b->total = total;
}

int alloc_result_block(const Lowered& L, uint64_t bound, char** block, bool* fits) {
  BlockLayout bl;
  plan_block(L, bound, sizeof(Control), &bl);
  *fits = bl.total <= STAGING_BYTES && getenv("PW_NO_DEFERRED") == nullptr;
  *block = nullptr;
  if (!*fits) return 0;
  void* p = nullptr;
  PW_TRY(dev_alloc(&p, bl.total));
  *block = (char*)p;
  return 0;
}

// cache key of the pilot statistics: the key columns (and the window grid) are all the pilot looks at
static std::string pilot_key(const PwQuery* q) {
  std::string k;
  auto put = [&](const void* p, size_t n) { k.append((const char*)p, n); };
  put(&q->n_keys, 4);
  for (int i = 0; i < q->n_keys; ++i) put(&q->key_columns[i], 4);
  if (q->dynamic) { const PwDynamic& d = *q->dynamic; put(&d.index_column, 4); put(&d.closed, 4); put(&d.every, 8); put(&d.period, 8); put(&d.offset, 8); }
  return k;
}

int run_groupby(const PwQuery* q, const PwFrame* f, Lowered& L, Table* table_out, uint32_t** slot_list_out, uint64_t* n_groups_out,
                const RunOpts* opts, RunState* state) {
  ThreadCtx& c = ctx();
  ScanPlan& P = L.plan;
  const int64_t N = f->n_rows;
  PwTimings& tm = c.timings;
  tm.n_rows = N;
  if (state) *state = RunState{};

  Control* dctl = nullptr;  // allocated once the table size is known (deferred mode: header of the result block)
  Control hctl{};

  // ---- strategy + table size -------------------------------------------------------------------
  PW_CUDA(cudaEventRecord(c.ev[1], c.stream));
  uint64_t cap = 0;
  bool use_hot = true;
  int64_t live_groups = 0;  // distinct keys among consecutive rows (sizes the hot table)
  double g_est = 0;         // estimated number of groups (0: no pilot ran)
  int64_t dense_range = 0;  // > 0: single integer key whose sampled values span this many ids
  bool dense_sentinels = false;
  const int64_t SMALL = 1 << 18;
  if (q->initial_table_slots > 0) cap = (uint64_t)q->initial_table_slots;
  auto take_range = [&](unsigned long long kmax_u, unsigned long long kmin_n, int64_t live) {
    if (!(kmax_u | kmin_n)) return;
    const uint64_t lo = ~kmin_n, hi = kmax_u;  // biased images, lo <= hi
    const uint64_t range = hi - lo + 1;
    // dense ids pay when the range is small and mostly populated (rows outside it still aggregate, through the
    // HBM table)
    if (range <= 4096 && range <= (uint64_t)std::max<int64_t>(64, 4 * live)) {
      dense_range = (int64_t)range;
      P.dense_min = (int64_t)(lo ^ 0x8000000000000000ull);
      dense_sentinels = P.dense_min <= -1 && P.dense_min + (int64_t)range > -2;  // -1 / -2 inside the range
    }
  };
  if (N <= SMALL) {
    if (!cap) cap = (uint64_t)std::max<int64_t>(2 * N, 64);
    if ((q->flags & PW_FLAG_FORCE_HOT_TABLE) && q->hot_table_slots == 0 && N > 0 && (dense_eligible(q, P) || wbucket_eligible(q, P))) {
      // unit-sized inputs reach the dense-id path through the force flag: range of ALL keys, one extra sync
      Control* kctl = nullptr;
      { void* p = nullptr; PW_TRY(dev_alloc(&p, sizeof(Control))); kctl = (Control*)p; }
      PW_CUDA(cudaMemsetAsync(kctl, 0, sizeof(Control), c.stream));
      key_range_kernel<<<64, 256, 0, c.stream>>>(P.slots[P.keys[0].slot], 0, 1, N, &kctl->kmax_u, &kctl->kmin_n);
      tm.kernel_launches++;
      PW_CUDA(cudaMemcpyAsync(&hctl, kctl, sizeof(Control), cudaMemcpyDeviceToHost, c.stream));
      PW_CUDA(cudaStreamSynchronize(c.stream));
      dev_free(kctl);
      take_range(hctl.kmax_u, hctl.kmin_n, 1024);
    }
  } else if (!cap || !(q->flags & (PW_FLAG_FORCE_HOT_TABLE | PW_FLAG_FORCE_GLOBAL_TABLE))) {
    // (1) strided sample over the whole input -> table size; (2) a contiguous block from the middle -> do
    // consecutive rows share few groups (then the hot table pays)?  Both pilots are queued, then one sync.
    // A frame remembers what the pilot found per key set (PwFrame::pilot): resident frames pay this phase once.
    const int64_t n_s = N >= (32ll << 20) ? (1 << 17) : (1 << 16);
    const int64_t n_b = 1 << 16;
    int64_t mid = ((N / 2) / ROWS_PER_STEP) * ROWS_PER_STEP;
    if (mid + n_b > N) mid = 0;
    PilotStats ps;
    bool cached = false;
    const std::string pkey = pilot_key(q);
    static const bool no_pilot_cache = getenv("PW_NO_PILOT_CACHE") != nullptr;
    if (!no_pilot_cache) {
      std::lock_guard<std::mutex> lk(f->mu);
      auto it = f->pilot.find(pkey);
      if (it != f->pilot.end()) { ps = it->second; cached = true; }
    }
    if (!cached) {
      Pilot p1, p2;
      Control h1{}, h2{};
      // fused pilot (one launch, pw_pilot.cuh); the two-scan form below is the fallback without NVRTC
      int frc = 1;
      if (jit_available() && !getenv("PW_NO_FUSED_PILOT")) {
        ScanPlan PP = L.plan;
        PP.row_begin = 0; PP.row_stride = 2; PP.vec_ok = 0;  // shape: strided scalar loads
        PP.hot_slots = 0; memset(&PP.hot, 0, sizeof PP.hot); PP.check_sorted = 0; PP.n_preds = 0;
        PP.n_acc = 1; PP.accs[0].op = OP_ADD_I64; PP.accs[0].src = SRC_ONE; PP.accs[0].vexpr = 0; PP.n_vexpr = 0;
        PP.gflags = GF_LEN; PP.acc_gbase = 0;
        const int kw = padded_kw(PP.n_kw);
        PP.n_kw = kw;
        int row = 4;
        while (row < kw + 1) row <<= 1;
        const uint64_t cap1 = (uint64_t)n_s * 2 + 64, cap2 = (uint64_t)n_b * 2 + 64;
        const size_t rows_bytes = ((cap1 + 2) + (cap2 + 2)) * 8 * (size_t)row;
        const size_t ctl_bytes = (sizeof(Control) + 15) & ~(size_t)15;
        const size_t zero_bytes = ctl_bytes + ((cap1 + 2) + (cap2 + 2)) * 4;
        char* arena = nullptr;
        { void* p = nullptr; PW_TRY(dev_alloc(&p, rows_bytes + zero_bytes)); arena = (char*)p; }
        PW_CUDA(cudaMemsetAsync(arena, 0xFF, rows_bytes, c.stream));
        PW_CUDA(cudaMemsetAsync(arena + rows_bytes, 0, zero_bytes, c.stream));
        Control* pctl = (Control*)(arena + rows_bytes);
        PilotParams pp{};
        pp.begin[0] = 0; pp.stride[0] = N / n_s; pp.n[0] = n_s;
        pp.begin[1] = mid; pp.stride[1] = 1; pp.n[1] = n_b;
        uint64_t* rows0 = (uint64_t*)arena;
        uint32_t* state0 = (uint32_t*)(arena + rows_bytes + ctl_bytes);
        for (int t = 0; t < 2; ++t) {
          Table& T = pp.table[t];
          T.keys = t == 0 ? rows0 : rows0 + (cap1 + 2) * (uint64_t)row;
          T.accs = T.keys + kw;
          T.state = t == 0 ? state0 : state0 + (cap1 + 2);
          T.cap = t == 0 ? cap1 : cap2;
          T.key_sw = T.acc_sw = 1; T.key_ss = T.acc_ss = (uint64_t)row;
          T.overflow = &pctl->overflow; T.spilled = &pctl->spilled;
        }
        pp.distinct[0] = &pctl->counter; pp.distinct[1] = &pctl->null_counts[0];
        if (dense_eligible(q, P) || wbucket_eligible(q, P)) { pp.kmax_u = &pctl->kmax_u; pp.kmin_n = &pctl->kmin_n; }
        PP.table = pp.table[0]; PP.not_sorted = &pctl->not_sorted;
        const int kwc = kw <= 1 ? 1 : (kw <= 2 ? 2 : (kw <= 4 ? 4 : 6));
        frc = launch_pilot_jit(PP, pp, narrow_class(PP) ? 4 : 12, kwc, c.stream);
        if (frc == 0) {
          cudaMemcpyAsync(&h1, pctl, sizeof(Control), cudaMemcpyDeviceToHost, c.stream);
          if (cudaStreamSynchronize(c.stream) != cudaSuccess) frc = fail(PW_ERR_CUDA, "pilot failed: %s", cudaGetErrorString(cudaGetLastError()));
          h2.counter = h1.null_counts[0];
          h2.overflow = h1.overflow;
        }
        dev_free(arena);
        if (frc < 0) return frc;
      }
      int prc = 0;
      if (frc != 0) {
        prc = pilot_launch(L, 0, N / n_s, n_s, &p1);
        if (!prc && dense_eligible(q, P)) {
          const RawSlot& ks = P.slots[P.keys[0].slot];
          key_range_kernel<<<64, 256, 0, c.stream>>>(ks, 0, N / n_s, n_s, &p1.dctl->kmax_u, &p1.dctl->kmin_n);
          key_range_kernel<<<64, 256, 0, c.stream>>>(ks, mid, 1, n_b, &p1.dctl->kmax_u, &p1.dctl->kmin_n);
          tm.kernel_launches += 2;
        }
        if (!prc) {
          cudaMemcpyAsync(&h1, p1.dctl, sizeof(Control), cudaMemcpyDeviceToHost, c.stream);
          prc = pilot_launch(L, mid, 1, n_b, &p2);
        }
        if (!prc) {
          cudaMemcpyAsync(&h2, p2.dctl, sizeof(Control), cudaMemcpyDeviceToHost, c.stream);
          if (cudaStreamSynchronize(c.stream) != cudaSuccess) prc = fail(PW_ERR_CUDA, "pilot failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
      }
      pilot_free(&p1); pilot_free(&p2);
      if (prc) return prc;
      ps.distinct_strided = h1.counter; ps.distinct_block = h2.counter;
      ps.kmax_u = h1.kmax_u; ps.kmin_n = h1.kmin_n;
      ps.overflow = (h1.overflow == 2 || h2.overflow == 2) ? 2 : 0;
      if (!no_pilot_cache) {
        std::lock_guard<std::mutex> lk(f->mu);
        if (f->pilot.size() > 256) f->pilot.clear();
        f->pilot[pkey] = ps;
      }
    }
    if (ps.overflow == 2) return fail(PW_ERR_UNSUPPORTED, "string key longer than 12 bytes (long views need the data buffers: SURVEY 8f rank 1)");
    double g = solve_groups((double)ps.distinct_strided, (double)n_s);
    g = std::min(g, (double)N);
    g_est = g;
    if (!cap) cap = (uint64_t)std::max(1024.0, std::min(2.0 * (double)N + 64.0, 2.5 * g + 1024.0));
    live_groups = (int64_t)ps.distinct_block;
    use_hot = ps.distinct_block <= 2048;
    if (use_hot) take_range(ps.kmax_u, ps.kmin_n, live_groups);
  }
  if (opts && opts->min_cap > cap) cap = opts->min_cap;
  if (q->flags & PW_FLAG_FORCE_HOT_TABLE) use_hot = true;
  if (q->flags & PW_FLAG_FORCE_GLOBAL_TABLE) use_hot = false;
  if (live_groups == 0) live_groups = std::min<int64_t>(std::max<int64_t>(N, 4), 1024);
  const bool windowed = P.dyn.enabled != 0;
  if (q->flags & PW_FLAG_NO_DENSE_IDS) dense_range = 0;   // (the frame's cached pilot statistics may carry a range)
  // windows several tiles long (the contiguous pilot block of 65 536 rows saw at most ~4 windows), else the hash path
  if (windowed && dense_range > 0 && N > SMALL && live_groups > 4 * dense_range) dense_range = 0;
  if (use_hot && dense_range > 0 && !windowed && q->hot_table_slots == 0 && plan_hot(P, live_groups, 0, dense_range)) {
    if (dense_sentinels) P.hot.dense = 2;
    if (!(q->flags & PW_FLAG_NO_BUCKETS)) plan_bucket(P, live_groups, dense_range, dense_sentinels, false);
  }
  else if (use_hot && !plan_hot(P, live_groups, q->hot_table_slots)) use_hot = false;
  // the bucket tier does not need the dense per-cell table to fit (many accumulators): it only needs the id range
  else if (use_hot && dense_range > 0 && q->hot_table_slots == 0 && !(q->flags & (PW_FLAG_NO_BUCKETS | PW_FLAG_NO_DENSE_IDS)))
    plan_bucket(P, windowed ? dense_range : live_groups, dense_range, dense_sentinels, windowed);
  // no small dense range (sparse integers, strings, several key columns) but few live groups: ids from a key index
  else if (use_hot && dense_range == 0 && !windowed && q->hot_table_slots == 0 && !(q->flags & PW_FLAG_NO_BUCKETS) && N > SMALL)
    plan_bucket(P, live_groups, 0, false, false);
  if (cap > 0xFFFFFFF0ull) return fail(PW_ERR_UNSUPPORTED, "table larger than 2^32 slots");
  if (getenv("PW_DEBUG"))
    fprintf(stderr, "[pw] rows=%lld kw=%d slots=%d accs=%d cap=%llu hot=%d live=%lld gcap=%d S=%d R=%d n_mm=%d smem=%d dense=%d min=%lld\n", (long long)N, P.n_kw,
            P.n_slots, P.n_acc, (unsigned long long)cap, (int)use_hot, (long long)live_groups, P.hot.gcap, P.hot.idx_slots, P.hot.replicas,
            P.hot.n_mm, P.hot.total_bytes, P.hot.dense, (long long)P.dense_min);
  if (getenv("PW_DEBUG") && P.hot.bucket)
    fprintf(stderr, "[pw] bucket tier: threads=%d gcap=%d J=%d nbuf=%d stages=%d meta=%d smem=%d idx=%d win=%d range=%d\n", P.hot.b_threads, P.hot.b_gcap, P.hot.b_j, P.hot.b_nbuf,
            P.hot.b_stages, P.hot.b_meta, P.hot.b_bytes, P.hot.b_idx, P.hot.b_win, P.hot.b_range);
  PW_CUDA(cudaEventRecord(c.ev[2], c.stream));

  // ---- overlapping windows: plain HBM table, sized for the windows every row joins
  uint64_t cap_max = (uint64_t)2 * (uint64_t)N + 64;
  if (P.overlap) {
    use_hot = false; memset(&P.hot, 0, sizeof P.hot);
    const uint64_t m = (uint64_t)((P.dyn.period + P.dyn.every - 1) / P.dyn.every) + 1;
    cap_max = (uint64_t)2 * (uint64_t)N * m + 64;
    cap = std::min<uint64_t>(cap * m, cap_max);
    if (cap > 0xFFFFFFF0ull) return fail(PW_ERR_UNSUPPORTED, "table larger than 2^32 slots");
  }
  // ---- sorted keys (the caller's flag, as the reference's IsSorted): groups are runs of equal keys
  const bool runs = (q->flags & PW_FLAG_KEYS_SORTED) && !P.dyn.enabled && !P.row_group_out && P.rowid_slot_p1 == 0 && N > 0 && jit_available() &&
                    !getenv("PW_NO_RUNS") && !(q->flags & (PW_FLAG_FORCE_HOT_TABLE | PW_FLAG_FORCE_GLOBAL_TABLE | PW_FLAG_FORCE_PARTITION));
  P.runs = runs ? 1 : 0;
  // ---- high-cardinality tier: many groups, several rows each, no locality -> partition the rows by key hash first
  // (the analogue of the reference's partitioned group-by; POLARS_FORCE_PARTITION / POLARS_NO_PARTITION = the flags)
  PartTemp ptmp;
  ScanPlan PP{};  // plan over the partitioned copy
  bool partitioned = false;
  RadixTemp rtmp;
  RadixParams rprm{};
  bool radix = false;
  {
    const bool forced = (q->flags & PW_FLAG_FORCE_PARTITION) != 0;
    const bool pays = !use_hot && g_est >= 262144.0 && N >= (4ll << 20) && (double)N >= 3.0 * g_est &&
                      !(q->flags & (PW_FLAG_FORCE_HOT_TABLE | PW_FLAG_FORCE_GLOBAL_TABLE));
    // second form first (pw_radix.cuh); a retry after an overflow seen late (opts->min_cap) keeps the older tiers
    if ((forced || pays) && N > 0 && !runs && radix_eligible(q, P) && !(opts && opts->min_cap) && !getenv("PW_NO_RADIX_NOW")) {
      const double g_hint = g_est > 0 ? g_est : (double)N;
      PW_CUDA(cudaEventRecord(c.ev[10], c.stream));
      const int prc = radix_prepare(P, g_hint, &PP, &rprm, &rtmp);
      PW_CUDA(cudaEventRecord(c.ev[11], c.stream));
      if (prc < 0) { radix_free(rtmp); return prc; }
      radix = prc == 0;
      if (!radix) radix_free(rtmp);
      else cap = std::min<uint64_t>((uint64_t)2 * (uint64_t)N + 64, (uint64_t)(1.4 * g_hint) + 16384);
    }
    if (!radix && (forced || pays) && N > 0 && !runs && part_eligible(q, P)) {
      PW_CUDA(cudaEventRecord(c.ev[10], c.stream));
      const int prc = partition_input(P, g_est > 0 ? g_est : std::max<double>((double)N / 4.0, 64.0), &PP, &ptmp);
      PW_CUDA(cudaEventRecord(c.ev[11], c.stream));
      if (prc < 0) { part_free(ptmp); return prc; }
      partitioned = prc == 0;
      if (!partitioned) part_free(ptmp);
    }
  }

  // ---- small table: leave the group count on the device (one host synchronisation per query, in emit_results)
  static const bool no_deferred = getenv("PW_NO_DEFERRED") != nullptr;
  BlockLayout bl;
  bool deferred = false;
  if (opts && opts->allow_deferred && state && !no_deferred && cap + 2 <= (1u << 16)) {
    if (opts->control_only) { bl.total = sizeof(Control); deferred = true; }
    else {
      plan_block(L, cap + 2, sizeof(Control), &bl);
      deferred = bl.total <= STAGING_BYTES;
    }
  }
  char* block = nullptr;
  if (deferred) { void* p = nullptr; PW_TRY(dev_alloc(&p, bl.total)); block = (char*)p; dctl = (Control*)block; }
  else { void* p = nullptr; PW_TRY(dev_alloc(&p, sizeof(Control))); dctl = (Control*)p; }

  // ---- scan (with growth retries) ------------------------------------------------------------------
  Table T{};
  Table T0{};   // overlapping windows: earliest index value per key slice
  uint32_t* slots = nullptr;
  if (!(opts && opts->min_cap)) tm.retries = 0;
  auto drop = [&]() { part_free(ptmp); radix_free(rtmp); if (T0.keys) { free_table(T0); T0 = Table{}; } if (T.keys) free_table(T); dev_free(slots); dev_free(deferred ? (void*)block : (void*)dctl); };
  for (;;) {
    int rc = alloc_table(&T, padded_kw(P.n_kw), P.n_acc, cap, dctl);
    if (!rc && radix) {
      // only the overflow region and the escape slots are probed: the dense region is written whole by its CTAs
      radix_split(cap, &rprm);
      Table V = T; V.cap = rprm.ovf_cap;
      rc = init_table(V, P, c.stream, dctl);
      Table E = T; E.keys = T.keys + cap * T.key_ss; E.accs = T.accs + cap * T.acc_ss; E.state = T.state + cap; E.cap = 0;
      if (!rc) rc = init_table(E, P, c.stream, nullptr);
    } else
    if (!rc) rc = init_table(T, P, c.stream, dctl);  // also clears the control block
    if (rc) { drop(); return rc; }
    P.table = T;
    P.not_sorted = &dctl->not_sorted;
    P.hot_slots = use_hot ? P.hot.idx_slots : 0;
    cudaEventRecord(c.ev[8], c.stream);
    if (radix) {
      PP.table = T; PP.not_sorted = &dctl->not_sorted;
      radix_split(cap, &rprm);
      rprm.dense_count = rtmp.dense_count;
      if (cudaMemsetAsync(rtmp.dense_count, 0, 16, c.stream) != cudaSuccess) { drop(); return fail(PW_ERR_CUDA, "radix tier: memset failed"); }
      rc = launch_radix_jit(PP, rprm, 4, rtmp.kwc, rtmp.smem3, (int64_t)1 << rprm.log2_parts, c.sm_count, c.stream);
      if (rc > 0) rc = fail(PW_ERR_INTERNAL, "radix tier: the aggregation kernel is not available");
    } else if (P.overlap && N > 0) {
      // overlapping windows: first the earliest index value of every key slice (its own table, one MIN word), then the
      // rows join their windows (pw_overlap.cuh)
      if (T0.keys) free_table(T0);
      const uint64_t cap0 = std::min<uint64_t>(cap, (uint64_t)2 * (uint64_t)N + 64);
      rc = alloc_table_raw(&T0, padded_kw(P.n_kw), 1, cap0, &dctl->overflow, &dctl->spilled);
      if (!rc) {
        AccOps ops0{}; ops0.n = 1; ops0.op[0] = OP_MIN_I64;
        const int g0 = (int)std::min<uint64_t>((cap0 + 2 + 255) / 256, 148 * 8);
        table_init_kernel<<<g0, 256, 0, c.stream>>>(T0, padded_kw(P.n_kw), ops0, nullptr);
        if (cudaGetLastError() != cudaSuccess) rc = fail(PW_ERR_CUDA, "table_init_kernel launch failed");
        tm.kernel_launches++;
      }
      if (!rc) { P.t0 = T0; P.overlap = 2; rc = launch_scan(P, c.sm_count, c.stream); P.overlap = 1; }
      if (!rc) rc = launch_scan(P, c.sm_count, c.stream);
    } else if (partitioned) {
      PP.table = T; PP.not_sorted = &dctl->not_sorted; PP.hot_slots = PP.hot.idx_slots;
      if (PP.n_rows > 0) rc = launch_scan(PP, c.sm_count, c.stream);
    } else if (N > 0) rc = launch_scan(P, c.sm_count, c.stream);
    if (rc) { drop(); return rc; }
    cudaEventRecord(c.ev[9], c.stream);
    // queue the compaction right behind the scan: its result is simply discarded when the scan overflowed
    { void* p = nullptr; rc = dev_alloc(&p, (cap + 2) * 4); if (rc) { drop(); return rc; } slots = (uint32_t*)p; }
    if (radix) {
      int grid = (int)std::min<uint64_t>((cap + 2 + 255) / 256, 148 * 8);
      radix_compact_kernel<<<grid, 256, 0, c.stream>>>(T, padded_kw(P.n_kw), rprm.ovf_cap, rtmp.dense_count, rprm.dense_cap, slots, &dctl->counter);
      if (cudaGetLastError() != cudaSuccess) { drop(); return fail(PW_ERR_CUDA, "radix_compact_kernel launch failed"); }
      tm.kernel_launches++;
    } else {
      int grid = (int)std::min<uint64_t>((cap + 2 + 255) / 256, 148 * 8);
      compact_kernel<<<grid, 256, 0, c.stream>>>(T, padded_kw(P.n_kw), slots, &dctl->counter);
      if (cudaGetLastError() != cudaSuccess) { drop(); return fail(PW_ERR_CUDA, "compact_kernel launch failed"); }
      tm.kernel_launches++;
    }
    if (deferred) break;
    if (cudaMemcpyAsync(&hctl, dctl, sizeof(Control), cudaMemcpyDeviceToHost, c.stream) != cudaSuccess ||
        cudaStreamSynchronize(c.stream) != cudaSuccess) {
      drop();
      return fail(PW_ERR_CUDA, "scan failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (hctl.overflow == 2) { drop(); return fail(PW_ERR_UNSUPPORTED, "string key longer than 12 bytes (long views need the data buffers: SURVEY 8f rank 1)"); }
    if (hctl.overflow == 1) {
      free_table(T); dev_free(slots); slots = nullptr;
      if (cap >= cap_max) { drop(); return fail(PW_ERR_INTERNAL, "hash table overflow at maximum size"); }
      cap = std::min<uint64_t>(cap * 4, cap_max);
      tm.retries++;
      if (radix) { radix = false; radix_free(rtmp); }  // group estimate too low or a skewed partition: the plain HBM table takes over
      continue;
    }
    break;
  }
  tm.strategy = runs ? 8 : radix ? 10 : partitioned ? 5 : (use_hot ? (P.hot.bucket ? (P.hot.b_win ? 6 : (P.hot.b_idx ? 9 : 7)) : (P.hot.dense ? 4 : 1)) : 2);
  tm.table_slots = (int64_t)cap;
  tm.partition_ms = 0.0f;
  if (deferred) {
    // count, overflow and sortedness are looked at by emit_results, after the single synchronisation
    part_free(ptmp);
    radix_free(rtmp);
    if (T0.keys) { free_table(T0); T0 = Table{}; }
    cudaEventRecord(c.ev[3], c.stream);
    const uint64_t bound = cap + 2;
    int rc = order_groups(L, T, padded_kw(P.n_kw), &slots, bound, &dctl->counter);
    if (rc) { ptmp = PartTemp{}; drop(); return rc; }
    state->deferred = true; state->dctl = dctl; state->block = block; state->cap = cap;
    *table_out = T;
    *slot_list_out = slots;
    *n_groups_out = bound;
    return 0;
  }
  if (hctl.not_sorted) {
    drop();
    return fail(PW_ERR_NOT_SORTED, "argument in operation 'group_by_dynamic' is not sorted, please sort the 'expr/series/column' first");
  }
  part_free(ptmp);
  if (T0.keys) { free_table(T0); T0 = Table{}; }
  const bool was_radix = radix || rtmp.small != nullptr;
  radix_free(rtmp);
  if (partitioned || was_radix) { float ms = 0; if (cudaEventElapsedTime(&ms, c.ev[10], c.ev[11]) == cudaSuccess) tm.partition_ms = ms; }
  tm.spilled_rows = (int64_t)hctl.spilled;
  PW_CUDA(cudaEventRecord(c.ev[3], c.stream));
  const uint64_t G = hctl.counter;
  tm.n_groups = (int64_t)G;
  if (state) state->cap = cap;
  { int rc = order_groups(L, T, padded_kw(P.n_kw), &slots, G); if (rc) { ptmp = PartTemp{}; drop(); return rc; } }
  dev_free(dctl);
  *table_out = T;
  *slot_list_out = slots;
  *n_groups_out = G;
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// result emission: one kernel per output column, then D2H into malloc'd Arrow buffers
// ------------------------------------------------------------------------------------------------------
int emit_results(const Lowered& L, const Table& T, const uint32_t* slot_list, uint64_t G,
                 struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out, RunState* state) {
  ThreadCtx& c = ctx();
  const size_t ncol = L.outs.size();
  const bool deferred = state && state->deferred;  // G is a bound; the count is in the block's header (device)
  auto bail = [&](int rc) { if (deferred) { dev_free(state->block); state->block = nullptr; } return rc; };
  if (*n_out < ncol) return bail(fail(PW_ERR_INVALID, "output capacity %zu < %zu result columns", *n_out, ncol));
  if (ncol > 64) return bail(fail(PW_ERR_UNSUPPORTED, "more than 64 result columns"));
  const int kw = padded_kw(L.plan.n_kw);
  const int grid = (int)std::max<uint64_t>(1, (G + 255) / 256);
  BlockLayout bl;
  plan_block(L, G, deferred ? sizeof(Control) : 64 * 8, &bl);
  const size_t total = bl.total;
  char* d_block = nullptr;
  unsigned long long* d_nulls = nullptr;
  if (deferred) { d_block = state->block; d_nulls = state->dctl->null_counts; }
  else {
    void* p = nullptr;
    PW_TRY(dev_alloc(&p, total));
    d_block = (char*)p;
    d_nulls = (unsigned long long*)d_block;
    if (cudaMemsetAsync(d_nulls, 0, 64 * 8, c.stream) != cudaSuccess) { dev_free(d_block); return fail(PW_ERR_CUDA, "memset failed"); }
  }
  for (size_t i0 = 0; i0 < ncol && G; i0 += EMIT_BATCH) {
    EmitBatch batch;
    const size_t nb = std::min<size_t>(EMIT_BATCH, ncol - i0);
    for (size_t j = 0; j < nb; ++j) {
      EmitDesc d = L.outs[i0 + j].emit;
      d.out_values = d_block + bl.val_off[i0 + j];
      d.out_validity = (uint32_t*)(d_block + bl.valid_off[i0 + j]);
      d.null_count = d_nulls + i0 + j;
      batch.d[j] = d;
    }
    emit_kernel<<<(unsigned)((uint64_t)grid * nb), 256, 0, c.stream>>>(T, kw, batch, (int)nb, slot_list, G, deferred ? state->dctl : nullptr);
    if (cudaGetLastError() != cudaSuccess) { dev_free(d_block); if (deferred) state->block = nullptr; return fail(PW_ERR_CUDA, "emit_kernel launch failed"); }
    c.timings.kernel_launches++;
  }
  cudaEventRecord(c.ev[4], c.stream);
  auto cuda_fail = [&](const char* what) {
    dev_free(d_block);
    if (deferred) state->block = nullptr;
    return fail(PW_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(cudaGetLastError()));
  };
  unsigned long long h_nulls[64] = {0};
  const char* h = nullptr;  // staged copy of the block (small results)
  if (total <= STAGING_BYTES) {
    // small result: ONE copy into this thread's pinned staging block, then split on the host.  Copies into
    // pageable memory cost a driver round trip each (~10 us x 2 x columns: most of a Q1-sized query).
    if (!c.staging && cudaHostAlloc(&c.staging, STAGING_BYTES, cudaHostAllocPortable) != cudaSuccess) return cuda_fail("cudaHostAlloc");
    if (cudaMemcpyAsync(c.staging, d_block, total, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess) return cuda_fail("cudaMemcpyAsync");
    cudaEventRecord(c.ev[5], c.stream);
    if (cudaStreamSynchronize(c.stream) != cudaSuccess) return cuda_fail("cudaStreamSynchronize");
    h = (const char*)c.staging;
  }
  uint64_t n_final = G;
  if (deferred) {
    const Control* hc = (const Control*)h;
    c.timings.spilled_rows = (int64_t)hc->spilled;
    if (hc->overflow == 2) return bail(fail(PW_ERR_UNSUPPORTED, "string key longer than 12 bytes (long views need the data buffers: SURVEY 8f rank 1)"));
    if (hc->overflow == 1 || hc->overflow == 3) { dev_free(d_block); state->block = nullptr; return PW_RETRY; }  // 3: a gathered segment overflowed (pw_partial.cu)
    if (hc->not_sorted) return bail(fail(PW_ERR_NOT_SORTED, "argument in operation 'group_by_dynamic' is not sorted, please sort the 'expr/series/column' first"));
    n_final = hc->counter;
    c.timings.n_groups = (int64_t)n_final;
    memcpy(h_nulls, hc->null_counts, 64 * 8);
  }
  const uint64_t Gf = n_final;
  std::vector<size_t> vb(ncol), qb(ncol);
  std::vector<void*> h_vals(ncol, nullptr), h_valid(ncol, nullptr);
  for (size_t i = 0; i < ncol; ++i) {
    vb[i] = out_col_bytes(L.outs[i].out_dtype, Gf);
    qb[i] = ((Gf + 31) / 32) * 4;
    h_vals[i] = host_alloc(vb[i]);
    h_valid[i] = host_alloc(qb[i]);
    if (!h_vals[i] || !h_valid[i]) { dev_free(d_block); if (deferred) state->block = nullptr; return fail(PW_ERR_INTERNAL, "out of host memory"); }
  }
  if (h) {
    if (!deferred) memcpy(h_nulls, h, 64 * 8);
    for (size_t i = 0; i < ncol; ++i) {
      memcpy(h_vals[i], h + bl.val_off[i], vb[i]);
      memcpy(h_valid[i], h + bl.valid_off[i], qb[i]);
    }
  } else {
    if (cudaMemcpyAsync(h_nulls, d_nulls, 64 * 8, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess) return cuda_fail("cudaMemcpyAsync");
    for (size_t i = 0; i < ncol; ++i) {
      if (cudaMemcpyAsync(h_vals[i], d_block + bl.val_off[i], vb[i], cudaMemcpyDeviceToHost, c.stream) != cudaSuccess ||
          cudaMemcpyAsync(h_valid[i], d_block + bl.valid_off[i], qb[i], cudaMemcpyDeviceToHost, c.stream) != cudaSuccess)
        return cuda_fail("cudaMemcpyAsync");
    }
    cudaEventRecord(c.ev[5], c.stream);
    if (cudaStreamSynchronize(c.stream) != cudaSuccess) return cuda_fail("cudaStreamSynchronize");
  }
  dev_free(d_block);
  if (deferred) state->block = nullptr;
  for (size_t i = 0; i < ncol; ++i) {
    const OutCol& o = L.outs[i];
    // a validity bitmap shorter than the bound's may carry bits past the count in its last word: harmless (Arrow readers
    // look at `length` bits)
    if (o.out_dtype == DT_VIEW && o.src_col && o.src_col->has_long) {
      // long keys: their bytes leave the device in one data buffer of the result's own
      void* h_data = nullptr;
      int64_t data_bytes = 0;
      PW_TRY(views_gather_long(o.src_col, h_vals[i], h_nulls[i] ? h_valid[i] : nullptr, Gf, &h_data, &data_bytes));
      if (h_data) {
        PW_TRY(make_host_view_array((int64_t)Gf, (int64_t)h_nulls[i], h_valid[i], h_vals[i], h_data, data_bytes, &out_cols[i]));
        PW_TRY(make_schema(o.format.c_str(), o.name.c_str(), true, &out_schemas[i]));
        continue;
      }
    }
    PW_TRY(make_host_array((int64_t)Gf, (int64_t)h_nulls[i], h_valid[i], h_vals[i], o.out_dtype == DT_VIEW ? 1 : 0, &out_cols[i]));
    PW_TRY(make_schema(o.format.c_str(), o.name.c_str(), true, &out_schemas[i]));
  }
  *n_out = ncol;
  return 0;
}

}  // namespace pw
