// pw_capi.cu — the extern "C" surface declared in include/polarway_b200.h
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include "pw_engine.h"

using namespace pw;

namespace {

int upload_column(const struct ArrowArray* a, const struct ArrowSchema* s, bool zero_copy, FrameColumn* out) {
  if (!a || !s) return fail(PW_ERR_INVALID, "null ArrowArray/ArrowSchema");
  out->format = s->format ? s->format : "";
  out->name = s->name ? s->name : "";
  PW_TRY(parse_format(s->format, &out->dtype));
  if (a->n_buffers < 2) return fail(PW_ERR_INVALID, "column '%s': expected at least 2 buffers", out->name.c_str());
  if (out->dtype == DT_BOOL) {
    // Boolean values are a bitmap like validity: both are taken from the byte that holds bit `offset`, so that one
    // residual bit offset serves both
    const int64_t n = a->length;
    const unsigned char* vals = (const unsigned char*)a->buffers[1];
    const unsigned char* valid = (const unsigned char*)a->buffers[0];
    int64_t nulls = a->null_count;
    if (!valid) nulls = 0;
    else if (nulls < 0) nulls = 1;
    out->null_count = nulls;
    out->bit_offset = (int32_t)(a->offset & 7);
    const size_t first = (size_t)(a->offset >> 3);
    const size_t bytes = (size_t)(((a->offset & 7) + n + 7) >> 3);
    if (zero_copy) {
      out->values = vals + first;
      out->validity = nulls ? valid + first : nullptr;
      return 0;
    }
    cudaStream_t st = ctx().stream;
    void* d = nullptr;
    PW_TRY(dev_alloc(&d, bytes + 32));
    if (bytes) PW_CUDA(cudaMemcpyAsync(d, vals + first, bytes, cudaMemcpyHostToDevice, st));
    out->values = d; out->owned_values = d;
    if (nulls) {
      void* dv = nullptr;
      PW_TRY(dev_alloc(&dv, bytes + 32));
      PW_CUDA(cudaMemcpyAsync(dv, valid + first, bytes, cudaMemcpyHostToDevice, st));
      out->validity = (const uint8_t*)dv; out->owned_validity = dv;
    }
    return 0;
  }
  const int w = out->dtype == DT_VIEW ? 16 : (out->dtype == DT_I8 || out->dtype == DT_U8 ? 1 : (out->dtype == DT_I16 || out->dtype == DT_U16 ? 2 : (out->dtype == DT_I32 || out->dtype == DT_U32 || out->dtype == DT_F32 ? 4 : 8)));
  const int64_t n = a->length;
  const unsigned char* vals = (const unsigned char*)a->buffers[1] + (size_t)a->offset * w;
  const unsigned char* valid = (const unsigned char*)a->buffers[0];
  int64_t nulls = a->null_count;
  if (!valid) nulls = 0;
  else if (nulls < 0) nulls = 1;  // unknown: treat as nullable
  out->null_count = nulls;
  cudaStream_t st = ctx().stream;
  // view columns: the variadic data buffers behind values longer than 12 bytes — buffers = [validity, views, data...,
  // sizes] (polars-arrow/src/ffi: the variadic buffer sizes are the last buffer)
  if (out->dtype == DT_VIEW && a->n_buffers >= 4) {
    const int64_t n_var = a->n_buffers - 3;
    const int64_t* sizes = (const int64_t*)a->buffers[a->n_buffers - 1];
    for (int64_t v = 0; v < n_var; ++v) {
      const void* src = a->buffers[2 + v];
      if (zero_copy) { out->var_bufs.push_back(src); continue; }
      const size_t bytes = sizes && sizes[v] > 0 ? (size_t)sizes[v] : 0;
      void* dv = nullptr;
      PW_TRY(dev_alloc(&dv, bytes + 32));
      out->owned_var.push_back(dv);
      out->var_bufs.push_back(dv);
      if (bytes && src) PW_CUDA(cudaMemcpyAsync(dv, src, bytes, cudaMemcpyHostToDevice, st));
    }
  }
  if (zero_copy) {
    out->values = vals;
    out->validity = nulls ? valid + (a->offset >> 3) : nullptr;
    out->bit_offset = (int32_t)(a->offset & 7);
    return views_intern(out, n);
  }
  void* d = nullptr;
  PW_TRY(dev_alloc(&d, (size_t)n * w + 32));
  if (n) PW_CUDA(cudaMemcpyAsync(d, vals, (size_t)n * w, cudaMemcpyHostToDevice, st));
  out->values = d; out->owned_values = d;
  if (nulls) {
    const size_t first = (size_t)(a->offset >> 3);
    const size_t bytes = (size_t)(((a->offset & 7) + n + 7) >> 3);
    void* dv = nullptr;
    PW_TRY(dev_alloc(&dv, bytes + 32));
    PW_CUDA(cudaMemcpyAsync(dv, valid + first, bytes, cudaMemcpyHostToDevice, st));
    out->validity = (const uint8_t*)dv; out->owned_validity = dv;
    out->bit_offset = (int32_t)(a->offset & 7);
  }
  return views_intern(out, n);
}

int make_frame(const struct ArrowArray* const* cols, const struct ArrowSchema* const* schemas, size_t n_cols,
               bool zero_copy, PwFrame** out) {
  if (!out) return fail(PW_ERR_INVALID, "null output pointer");
  *out = nullptr;
  PW_TRY(ensure_device());
  PwFrame* f = new PwFrame();
  f->device = ctx().device;
  f->cols.resize(n_cols);
  for (size_t i = 0; i < n_cols; ++i) {
    int rc = upload_column(cols[i], schemas[i], zero_copy, &f->cols[i]);
    if (rc == 0 && i > 0 && cols[i]->length != cols[0]->length) rc = fail(PW_ERR_INVALID, "column %zu has length %lld, expected %lld", i, (long long)cols[i]->length, (long long)cols[0]->length);
    if (rc) { pw_b200_frame_free(f); return rc; }
  }
  f->n_rows = n_cols ? cols[0]->length : 0;
  *out = f;
  return 0;
}

}  // namespace

extern "C" {

uint32_t pw_b200_abi_version(void) { return PW_ABI_VERSION; }
const char* pw_b200_last_error(void) { return ctx().last_error.c_str(); }

int pw_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int pw_b200_set_device(int device) {
  ThreadCtx& c = ctx();
  if (device != c.device) { c.device = device; c.pool_ready = false; c.ev_ready = false; }
  return ensure_device();
}
int pw_b200_set_stream(void* s) { ctx().stream = (cudaStream_t)s; return 0; }
int pw_b200_last_timings(PwTimings* out) {
  if (!out) return fail(PW_ERR_INVALID, "null output");
  *out = ctx().timings;
  return 0;
}

int pw_b200_frame_upload(const struct ArrowArray* const* cols, const struct ArrowSchema* const* schemas, size_t n_cols, PwFrame** out) {
  ThreadCtx& c = ctx();
  int rc = ensure_device();
  if (rc) return rc;
  cudaEventRecord(c.ev[6], c.stream);
  rc = make_frame(cols, schemas, n_cols, false, out);
  if (rc) return rc;
  cudaEventRecord(c.ev[7], c.stream);
  PW_CUDA(cudaStreamSynchronize(c.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c.ev[6], c.ev[7]);
  c.timings.h2d_ms = ms;
  return 0;
}
int pw_b200_frame_from_device(const struct ArrowArray* const* cols, const struct ArrowSchema* const* schemas, size_t n_cols, PwFrame** out) {
  return make_frame(cols, schemas, n_cols, true, out);
}
int64_t pw_b200_frame_num_rows(const PwFrame* f) { return f ? f->n_rows : -1; }
int pw_b200_frame_free(PwFrame* f) {
  if (!f) return 0;
  for (auto& c : f->cols) {
    dev_free(c.owned_values); dev_free(c.owned_validity); dev_free((void*)c.d_var_ptrs);
    for (void* v : c.owned_var) dev_free(v);
  }
  delete f;
  return 0;
}

int pw_b200_frame_groupby(const PwQuery* q, const PwFrame* frame, struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out) {
  PW_TRY(ensure_device());
  if (!q || !frame || !out_cols || !out_schemas || !n_out) return fail(PW_ERR_INVALID, "null argument");
  ThreadCtx& c = ctx();
  const auto t_begin = std::chrono::steady_clock::now();
  auto host_ms = [&]() { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  const float h2d = c.timings.h2d_ms;
  memset(&c.timings, 0, sizeof c.timings);
  c.timings.h2d_ms = h2d;
  PW_CUDA(cudaEventRecord(c.ev[0], c.stream));
  if (q->dynamic) {
    bool handled = false;
    int rc = run_dynamic_segmented(q, frame, out_cols, out_schemas, n_out, &handled);
    if (rc) return rc;
    if (handled) { c.timings.host_ms = host_ms(); return 0; }
  }
  Lowered L;
  PW_TRY(lower_query(q, frame, &L));
  if (q->dynamic && q->n_keys > 0) PW_TRY(check_sorted_within_keys(q, frame));   // dynamic.rs:77-80, 327
  RunOpts ro;
  ro.allow_deferred = true;
  for (;;) {
    Table T{};
    uint32_t* slots = nullptr;
    uint64_t G = 0;
    RunState rs;
    PW_TRY(run_groupby(q, frame, L, &T, &slots, &G, &ro, &rs));
    int rc = emit_results(L, T, slots, G, out_cols, out_schemas, n_out, &rs);
    dev_free(slots);
    free_table(T);
    if (rc == PW_RETRY) {
      // the table was sized from a sample and overflowed, which the deferred path only learns here: go again with a
      // larger table, synchronously (the classic path grows on its own)
      ro.min_cap = rs.cap * 4; ro.allow_deferred = false;
      c.timings.retries++;
      continue;
    }
    if (rc) return rc;
    break;
  }
  float ms;
  if (cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]) == cudaSuccess) c.timings.estimate_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]) == cudaSuccess) c.timings.scan_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]) == cudaSuccess) c.timings.finalize_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[4], c.ev[5]) == cudaSuccess) c.timings.d2h_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[0], c.ev[5]) == cudaSuccess) c.timings.total_device_ms = ms;
  if (cudaEventElapsedTime(&ms, c.ev[8], c.ev[9]) == cudaSuccess) c.timings.scan_kernel_ms = ms;
  c.timings.host_ms = host_ms();
  return 0;
}

int pw_b200_filter_groupby_agg(const PwQuery* q, const struct ArrowArray* const* cols, const struct ArrowSchema* const* schemas,
                               size_t n_cols, struct ArrowArray* out_cols, struct ArrowSchema* out_schemas, size_t* n_out) {
  PwFrame* f = nullptr;
  PW_TRY(pw_b200_frame_upload(cols, schemas, n_cols, &f));
  int rc = pw_b200_frame_groupby(q, f, out_cols, out_schemas, n_out);
  pw_b200_frame_free(f);
  return rc;
}

}  // extern "C"
