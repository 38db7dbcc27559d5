// pw_plugin.cu — the `_polars_plugin_*` compatibility shim.
//
// With these four symbols the library loads through the reference's UNMODIFIED expression-plugin loader
// (crates/polars-plan/src/plans/aexpr/function_expr/plugin.rs:16-142):
//   _polars_plugin_get_version                -> (major<<16)|minor = (0,1)           plugin.rs:85,139-141
//   _polars_plugin_get_last_error_message     -> thread-local C string              plugin.rs:63-72
//   _polars_plugin_filter_groupby_agg         (inputs, n, kwargs, len, *out, *ctx)  plugin.rs:93-125
//   _polars_plugin_field_filter_groupby_agg   (fields, n, *out, kwargs, len)        plugin.rs:183-205
// inputs = every column of the frame as SeriesExport (crates/polars-ffi/src/version_0.rs:7-16; the callee owns
// them and must release each one, plugin.rs:127-130); kwargs = pickle protocol-5 bytes of a dict
// (py-polars/src/polars/plugins.py:111-120) describing the query; the result is ONE Struct series whose fields are
// the result columns.  Failure: *out left empty (private_data == NULL) + message (plugin.rs:132-138).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "pw_engine.h"

using namespace pw;

namespace {

// ---- a pickle reader for the value shapes plugins.py produces (dict/list/tuple/str/int/float/bool/None) --------
struct PV {
  enum Kind { NONE, BOOL, INT, FLOAT, STR, LIST, DICT } kind = NONE;
  long long i = 0;
  double f = 0;
  std::string s;
  std::vector<PV> items;                        // LIST / tuple
  std::vector<std::pair<std::string, PV>> kv;   // DICT (string keys)
  const PV* get(const char* key) const {
    for (auto& p : kv) if (p.first == key) return &p.second;
    return nullptr;
  }
  double num() const { return kind == FLOAT ? f : (double)i; }
};

struct Unpickler {
  const uint8_t* p;
  const uint8_t* end;
  std::vector<PV> stack, memo;
  std::vector<size_t> marks;
  std::string err;
  bool need(size_t n) { if ((size_t)(end - p) < n) { err = "truncated pickle"; return false; } return true; }
  uint64_t le(int n) { uint64_t v = 0; for (int k = 0; k < n; ++k) v |= (uint64_t)p[k] << (8 * k); p += n; return v; }
  // false (err set) when there is no pending MARK or the mark lies above the stack: malformed bytes must not index
  // out of the vectors (these bytes cross the C ABI from the host process)
  bool pop_mark(std::vector<PV>* out) {
    if (marks.empty() || marks.back() > stack.size()) { err = "stack underflow"; return false; }
    const size_t m = marks.back(); marks.pop_back();
    out->assign(stack.begin() + m, stack.end());
    stack.resize(m);
    return true;
  }
  bool run(PV* out) {
    while (p < end) {
      const uint8_t op = *p++;
      switch (op) {
        case 0x80: if (!need(1)) return false; p += 1; break;                 // PROTO
        case 0x95: if (!need(8)) return false; p += 8; break;                 // FRAME
        case '}': { PV v; v.kind = PV::DICT; stack.push_back(v); break; }
        case ']': case ')': { PV v; v.kind = PV::LIST; stack.push_back(v); break; }
        case 0x94: if (stack.empty()) { err = "stack underflow"; return false; } memo.push_back(stack.back()); break;  // MEMOIZE
        case 'h': { if (!need(1)) return false; size_t k = *p++; if (k >= memo.size()) { err = "bad memo"; return false; } stack.push_back(memo[k]); break; }
        case 'j': { if (!need(4)) return false; size_t k = (size_t)le(4); if (k >= memo.size()) { err = "bad memo"; return false; } stack.push_back(memo[k]); break; }
        case 0x8c: { if (!need(1)) return false; size_t n = *p++; if (!need(n)) return false; PV v; v.kind = PV::STR; v.s.assign((const char*)p, n); p += n; stack.push_back(v); break; }
        case 'X': { if (!need(4)) return false; size_t n = (size_t)le(4); if (!need(n)) return false; PV v; v.kind = PV::STR; v.s.assign((const char*)p, n); p += n; stack.push_back(v); break; }
        case 'C': { if (!need(1)) return false; size_t n = *p++; if (!need(n)) return false; PV v; v.kind = PV::STR; v.s.assign((const char*)p, n); p += n; stack.push_back(v); break; }
        case 'K': { if (!need(1)) return false; PV v; v.kind = PV::INT; v.i = *p++; stack.push_back(v); break; }
        case 'M': { if (!need(2)) return false; PV v; v.kind = PV::INT; v.i = (long long)le(2); stack.push_back(v); break; }
        case 'J': { if (!need(4)) return false; PV v; v.kind = PV::INT; v.i = (int32_t)le(4); stack.push_back(v); break; }
        case 0x8a: { if (!need(1)) return false; int n = *p++; if (!need(n) || n > 8) { err = "long too wide"; return false; }
                     uint64_t u = 0; for (int k = 0; k < n; ++k) u |= (uint64_t)p[k] << (8 * k);
                     if (n && n < 8 && (p[n - 1] & 0x80)) u |= ~0ull << (8 * n);
                     p += n; PV v; v.kind = PV::INT; v.i = (long long)u; stack.push_back(v); break; }
        case 'G': { if (!need(8)) return false; uint64_t u = 0; for (int k = 0; k < 8; ++k) u = (u << 8) | p[k]; p += 8;
                    PV v; v.kind = PV::FLOAT; memcpy(&v.f, &u, 8); stack.push_back(v); break; }
        case 0x88: case 0x89: { PV v; v.kind = PV::BOOL; v.i = op == 0x88; stack.push_back(v); break; }
        case 'N': stack.push_back(PV()); break;
        case '(': marks.push_back(stack.size()); break;
        case 't': { PV v; v.kind = PV::LIST; if (!pop_mark(&v.items)) return false; stack.push_back(v); break; }
        case 0x85: case 0x86: case 0x87: {
          const size_t n = op - 0x84; if (stack.size() < n) { err = "stack underflow"; return false; }
          PV v; v.kind = PV::LIST; v.items.assign(stack.end() - n, stack.end()); stack.resize(stack.size() - n); stack.push_back(v); break; }
        case 'a': { if (stack.size() < 2) { err = "stack underflow"; return false; } PV x = stack.back(); stack.pop_back();
                    if (stack.back().kind != PV::LIST) { err = "APPEND to a non-list"; return false; } stack.back().items.push_back(x); break; }
        case 'e': { std::vector<PV> xs; if (!pop_mark(&xs)) return false; if (stack.empty()) { err = "stack underflow"; return false; }
                    if (stack.back().kind != PV::LIST) { err = "APPENDS to a non-list"; return false; } for (auto& x : xs) stack.back().items.push_back(x); break; }
        case 's': { if (stack.size() < 3) { err = "stack underflow"; return false; } PV v = stack.back(); stack.pop_back(); PV k = stack.back(); stack.pop_back();
                    if (stack.back().kind != PV::DICT || k.kind != PV::STR) { err = "bad SETITEM"; return false; } stack.back().kv.emplace_back(k.s, v); break; }
        case 'u': { std::vector<PV> xs; if (!pop_mark(&xs)) return false; if (stack.empty() || xs.size() % 2 || stack.back().kind != PV::DICT) { err = "bad SETITEMS"; return false; }
                    for (size_t k = 0; k < xs.size(); k += 2) { if (xs[k].kind != PV::STR) { err = "bad SETITEMS"; return false; } stack.back().kv.emplace_back(xs[k].s, xs[k + 1]); } break; }
        case '.': if (stack.empty()) { err = "empty pickle"; return false; } *out = stack.back(); return true;
        default: { char b[64]; snprintf(b, sizeof b, "unsupported pickle opcode 0x%02x", op); err = b; return false; }
      }
    }
    err = "pickle without STOP";
    return false;
  }
};

// owns every buffer a PwQuery built from kwargs points to
struct QueryHolder {
  PwQuery q{};
  std::vector<PwPredicate> preds;
  std::vector<int32_t> keys;
  std::vector<PwAgg> aggs;
  std::vector<std::string> names;
  PwDynamic dyn{};
};

int query_from_kwargs(const uint8_t* kw, size_t n, QueryHolder* h) {
  PV root;
  Unpickler u{kw, kw + n, {}, {}, {}, ""};
  if (!kw || !n) return fail(PW_ERR_INVALID, "the plugin needs kwargs describing the query (INTEGRATION.md)");
  if (!u.run(&root) || root.kind != PV::DICT) return fail(PW_ERR_INVALID, "kwargs: %s", u.err.empty() ? "not a dict" : u.err.c_str());
  h->q.abi_version = PW_ABI_VERSION;
  if (const PV* v = root.get("maintain_order")) h->q.maintain_order = (int)v->i;
  if (const PV* v = root.get("flags")) h->q.flags = (uint64_t)v->i;
  if (const PV* v = root.get("hot_table_slots")) h->q.hot_table_slots = (int)v->i;
  if (const PV* v = root.get("predicates"))
    for (const PV& p : v->items) {  // (column, op, value)
      if (p.items.size() != 3) return fail(PW_ERR_INVALID, "kwargs.predicates: expected (column, op, value)");
      PwPredicate q{};
      q.column = (int)p.items[0].i; q.op = (int)p.items[1].i;
      if (p.items[2].kind == PV::FLOAT) { q.scalar_is_float = 1; q.scalar.f = p.items[2].f; } else q.scalar.i = p.items[2].i;
      h->preds.push_back(q);
    }
  if (const PV* v = root.get("keys")) for (const PV& k : v->items) h->keys.push_back((int)k.i);
  if (const PV* v = root.get("aggs")) {
    h->names.reserve(v->items.size());
    for (const PV& a : v->items) {  // (name, kind, column | None, [(a, b, column), ...] | None[, ddof])
      if (a.items.size() < 3) return fail(PW_ERR_INVALID, "kwargs.aggs: expected (name, kind, column[, factors])");
      PwAgg g{};
      h->names.push_back(a.items[0].s);
      g.kind = (int)a.items[1].i;
      g.column = a.items[2].kind == PV::NONE ? -1 : (int)a.items[2].i;
      if (a.items.size() > 3 && a.items[3].kind == PV::LIST) {
        if (a.items[3].items.size() > PW_MAX_FACTORS) return fail(PW_ERR_UNSUPPORTED, "more than %d factors", PW_MAX_FACTORS);
        for (const PV& f : a.items[3].items) {
          if (f.items.size() != 3) return fail(PW_ERR_INVALID, "kwargs.aggs factor: expected (a, b, column)");
          g.factors[g.n_factors].a = f.items[0].num(); g.factors[g.n_factors].b = f.items[1].num(); g.factors[g.n_factors].column = (int)f.items[2].i;
          g.n_factors++;
        }
        if (g.n_factors) g.column = -1;
      }
      if (a.items.size() > 4 && a.items[4].kind == PV::INT) g.ddof = (int)a.items[4].i;   // var / std
      h->aggs.push_back(g);
    }
    for (size_t i = 0; i < h->aggs.size(); ++i) h->aggs[i].name = h->names[i].c_str();
  }
  if (const PV* v = root.get("dynamic"))
    if (v->kind == PV::DICT) {
      auto geti = [&](const char* k, long long d) { const PV* x = v->get(k); return x ? x->i : d; };
      h->dyn.index_column = (int)geti("index_column", 0); h->dyn.closed = (int)geti("closed", 0); h->dyn.label = (int)geti("label", 0);
      h->dyn.include_boundaries = (int)geti("include_boundaries", 0);
      h->dyn.every = geti("every", 0); h->dyn.period = geti("period", h->dyn.every); h->dyn.offset = geti("offset", 0);
      h->q.dynamic = &h->dyn;
    }
  h->q.n_predicates = (int)h->preds.size(); h->q.predicates = h->preds.data();
  h->q.n_keys = (int)h->keys.size(); h->q.key_columns = h->keys.data();
  h->q.n_aggs = (int)h->aggs.size(); h->q.aggs = h->aggs.data();
  return 0;
}

// ---- struct series construction ---------------------------------------------------------------------------------
struct StructPriv {
  std::vector<ArrowArray> child_arrays;
  std::vector<ArrowArray*> child_array_ptrs;
  std::vector<ArrowSchema> child_schemas;
  std::vector<ArrowSchema*> child_schema_ptrs;
  const void* buffers[1] = {nullptr};
  ArrowArray array{};
  ArrowArray* array_ptr = nullptr;
  ArrowSchema schema{};
};
void release_struct_array(ArrowArray* a) {
  if (!a || !a->release) return;
  for (int64_t i = 0; i < a->n_children; ++i) if (a->children[i]->release) a->children[i]->release(a->children[i]);
  a->release = nullptr;
}
void release_struct_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  for (int64_t i = 0; i < s->n_children; ++i) if (s->children[i]->release) s->children[i]->release(s->children[i]);
  s->release = nullptr;
}
void release_series(SeriesExport* e) {
  if (!e || !e->private_data) return;
  StructPriv* p = (StructPriv*)e->private_data;
  release_struct_array(&p->array);
  release_struct_schema(&p->schema);
  delete p;
  e->private_data = nullptr;
  e->release = nullptr;
}
void fill_struct_schema(StructPriv* p, size_t n) {
  p->child_schema_ptrs.resize(n);
  for (size_t i = 0; i < n; ++i) p->child_schema_ptrs[i] = &p->child_schemas[i];
  memset(&p->schema, 0, sizeof p->schema);
  p->schema.format = "+s"; p->schema.name = "filter_groupby_agg"; p->schema.flags = 0;
  p->schema.n_children = (int64_t)n; p->schema.children = p->child_schema_ptrs.data();
  p->schema.release = release_struct_schema;
}

// one contiguous host ArrowArray per input series (chunks concatenated on the host when there are several)
struct HostColumn {
  ArrowArray arr{};
  const void* bufs[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<uint8_t> values, validity;
};
int width_of_format(const char* f) {
  int32_t dt = 0;
  if (parse_format(f, &dt)) return -1;
  switch (dt) { case DT_I8: case DT_U8: return 1; case DT_I16: case DT_U16: return 2; case DT_I32: case DT_U32: case DT_F32: return 4; case DT_VIEW: return 16; case DT_BOOL: return -1; default: return 8; }
}
int concat_series(const SeriesExport& s, HostColumn* out) {
  if (s.len == 1) { out->arr = *s.arrays[0]; out->arr.release = nullptr; return 0; }
  const int w = width_of_format(s.field->format);
  if (w < 0) return fail(PW_ERR_UNSUPPORTED, "column format '%s'", s.field->format);
  int64_t total = 0, nulls = 0;
  for (size_t c = 0; c < s.len; ++c) { total += s.arrays[c]->length; nulls += s.arrays[c]->buffers[0] ? std::max<int64_t>(s.arrays[c]->null_count, 0) + (s.arrays[c]->null_count < 0) : 0; }
  out->values.resize((size_t)total * w + 16);
  if (nulls) out->validity.assign((size_t)(total + 7) / 8 + 8, 0);
  int64_t at = 0;
  for (size_t c = 0; c < s.len; ++c) {
    const ArrowArray* a = s.arrays[c];
    if (w == 16 && a->n_buffers > 3) {
      const uint8_t* views = (const uint8_t*)a->buffers[1] + (size_t)a->offset * 16;
      for (int64_t r = 0; r < a->length; ++r) { uint32_t len; memcpy(&len, views + r * 16, 4); if (len > 12) return fail(PW_ERR_UNSUPPORTED, "string longer than 12 bytes in a multi-chunk column (SURVEY 8f rank 1)"); }
    }
    memcpy(out->values.data() + (size_t)at * w, (const uint8_t*)a->buffers[1] + (size_t)a->offset * w, (size_t)a->length * w);
    if (nulls) {
      const uint8_t* vb = (const uint8_t*)a->buffers[0];
      for (int64_t r = 0; r < a->length; ++r) {
        const int64_t b = a->offset + r;
        const bool ok = !vb || ((vb[b >> 3] >> (b & 7)) & 1);
        if (ok) out->validity[(size_t)((at + r) >> 3)] |= (uint8_t)(1u << ((at + r) & 7));
      }
    }
    at += a->length;
  }
  out->bufs[0] = nulls ? out->validity.data() : nullptr;
  out->bufs[1] = out->values.data();
  memset(&out->arr, 0, sizeof out->arr);
  out->arr.length = total; out->arr.null_count = nulls; out->arr.n_buffers = 2; out->arr.buffers = out->bufs;
  return 0;
}

}  // namespace

extern "C" {

uint32_t _polars_plugin_get_version(void) { return (0u << 16) | 1u; }
const char* _polars_plugin_get_last_error_message(void) { return pw_b200_last_error(); }

void _polars_plugin_filter_groupby_agg(const SeriesExport* inputs, size_t n_inputs, const uint8_t* kwargs, size_t kwargs_len,
                                       SeriesExport* return_value, const CallerContext* /*ctx*/) {
  if (return_value) { return_value->private_data = nullptr; return_value->release = nullptr; return_value->field = nullptr; return_value->arrays = nullptr; return_value->len = 0; }
  QueryHolder h;
  int rc = query_from_kwargs(kwargs, kwargs_len, &h);
  std::vector<HostColumn> host(n_inputs);
  std::vector<const ArrowArray*> arrs(n_inputs);
  std::vector<const ArrowSchema*> schs(n_inputs);
  for (size_t i = 0; i < n_inputs && !rc; ++i) {
    rc = concat_series(inputs[i], &host[i]);
    arrs[i] = &host[i].arr; schs[i] = inputs[i].field;
  }
  StructPriv* p = nullptr;
  if (!rc && return_value) {
    p = new StructPriv();
    const size_t cap = h.keys.size() + h.aggs.size() + 4;
    p->child_arrays.resize(cap); p->child_schemas.resize(cap);
    size_t n_out = cap;
    rc = pw_b200_filter_groupby_agg(&h.q, arrs.data(), schs.data(), n_inputs, p->child_arrays.data(), p->child_schemas.data(), &n_out);
    if (!rc) {
      p->child_arrays.resize(n_out); p->child_schemas.resize(n_out);
      p->child_array_ptrs.resize(n_out);
      for (size_t i = 0; i < n_out; ++i) p->child_array_ptrs[i] = &p->child_arrays[i];
      fill_struct_schema(p, n_out);
      memset(&p->array, 0, sizeof p->array);
      p->array.length = n_out ? p->child_arrays[0].length : 0;
      p->array.n_buffers = 1; p->array.buffers = p->buffers;
      p->array.n_children = (int64_t)n_out; p->array.children = p->child_array_ptrs.data();
      p->array.release = release_struct_array;
      p->array_ptr = &p->array;
      return_value->field = &p->schema; return_value->arrays = &p->array_ptr; return_value->len = 1;
      return_value->release = release_series; return_value->private_data = p;
    } else delete p;
  }
  // the callee owns the inputs (plugin.rs:127-130)
  for (size_t i = 0; i < n_inputs; ++i) if (inputs[i].release) inputs[i].release(const_cast<SeriesExport*>(&inputs[i]));
}

void _polars_plugin_field_filter_groupby_agg(const struct ArrowSchema* fields, size_t n_fields, struct ArrowSchema* out,
                                             const uint8_t* kwargs, size_t kwargs_len) {
  if (!out) return;
  out->release = nullptr;
  QueryHolder h;
  if (query_from_kwargs(kwargs, kwargs_len, &h)) return;
  PwFrame f;  // schema only: lowering needs dtypes/names, not data
  f.cols.resize(n_fields);
  for (size_t i = 0; i < n_fields; ++i) {
    f.cols[i].format = fields[i].format ? fields[i].format : ""; f.cols[i].name = fields[i].name ? fields[i].name : "";
    if (parse_format(fields[i].format, &f.cols[i].dtype)) return;
  }
  Lowered L;
  if (lower_query(&h.q, &f, &L)) return;
  StructPriv* p = new StructPriv();  // owned by the schema's private_data
  const size_t n = L.outs.size();
  p->child_schemas.resize(n);
  for (size_t i = 0; i < n; ++i) make_schema(L.outs[i].format.c_str(), L.outs[i].name.c_str(), true, &p->child_schemas[i]);
  fill_struct_schema(p, n);
  *out = p->schema;
  out->private_data = p;
  out->release = [](ArrowSchema* s) {
    if (!s || !s->release) return;
    StructPriv* q = (StructPriv*)s->private_data;
    for (auto& c : q->child_schemas) if (c.release) c.release(&c);
    delete q;
    s->release = nullptr;
  };
}

}  // extern "C"
