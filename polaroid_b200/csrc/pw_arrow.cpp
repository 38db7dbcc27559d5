// pw_arrow.cpp — Arrow C Data Interface plumbing: format parsing, host result arrays with release
// callbacks (ownership rules of crates/polars-ffi/src/version_0.rs:42-55: the producer's `release`
// frees everything the struct points to; `private_data == NULL` / `release == NULL` marks an empty struct).
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <map>
#include <mutex>

#include "pw_engine.h"

namespace pw {

int parse_format(const char* fmt, int32_t* dtype) {
  if (!fmt) return fail(PW_ERR_INVALID, "schema without format");
  switch (fmt[0]) {
    case 'c': *dtype = DT_I8; return 0;
    case 'C': *dtype = DT_U8; return 0;
    case 's': *dtype = DT_I16; return 0;
    case 'S': *dtype = DT_U16; return 0;
    case 'i': *dtype = DT_I32; return 0;
    case 'I': *dtype = DT_U32; return 0;
    case 'l': *dtype = DT_I64; return 0;
    case 'L': *dtype = DT_U64; return 0;
    case 'f': *dtype = DT_F32; return 0;
    case 'g': *dtype = DT_F64; return 0;
    case 'b': *dtype = DT_BOOL; return 0;
    case 'v':
      if (fmt[1] == 'u' || fmt[1] == 'z') { *dtype = DT_VIEW; return 0; }
      break;
    case 't':
      // tdD date32 | tdm date64 | tts/ttm time32 | ttu/ttn time64 | ts?:tz timestamp | tD? duration
      if (fmt[1] == 'd') { *dtype = fmt[2] == 'D' ? DT_I32 : DT_I64; return 0; }
      if (fmt[1] == 't') { *dtype = (fmt[2] == 's' || fmt[2] == 'm') ? DT_I32 : DT_I64; return 0; }
      if (fmt[1] == 's' || fmt[1] == 'D') { *dtype = DT_I64; return 0; }
      break;
    default: break;
  }
  return fail(PW_ERR_UNSUPPORTED, "Arrow format '%s' is outside this path (SURVEY 8f: long/offset strings, nested types)", fmt);
}

// Result buffers.  Large ones come from a process-wide pool of PINNED blocks: the device-to-host copy of a
// 1e7-group result into pageable memory ran at ~5 GB/s (C3: 119 ms for 0.6 GB, most of the query); into pinned
// memory it runs at PCIe speed.  Pinning is expensive (~0.3 ms per MB), so a released result returns its blocks to the
// pool instead of unpinning them; the pool is capped (PW_PINNED_POOL_MB, default 8192) and anything beyond the cap is
// unpinned on release.  Small buffers stay on malloc.
namespace {
struct PinnedPool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;  // size -> block
  std::map<void*, size_t> live;              // every pinned block handed out
  size_t pooled = 0, cap = 0;
  bool disabled = false;
  PinnedPool() {
    const char* e = getenv("PW_PINNED_POOL_MB");
    cap = (size_t)(e ? atoll(e) : 8192) << 20;
    disabled = cap == 0;
  }
};
PinnedPool& pinned_pool() { static PinnedPool* p = new PinnedPool(); return *p; }  // never destroyed (library is never unloaded)
const size_t PINNED_MIN = 1u << 20;
}  // namespace

void* host_alloc(size_t bytes) {
  if (bytes < PINNED_MIN) return malloc(bytes + 64);
  PinnedPool& pool = pinned_pool();
  const size_t gran = 2u << 20;
  const size_t rounded = (bytes + gran - 1) / gran * gran;
  if (!pool.disabled) {
    {
      std::lock_guard<std::mutex> lk(pool.mu);
      auto it = pool.free_blocks.lower_bound(rounded);
      if (it != pool.free_blocks.end() && it->first <= rounded + rounded / 2) {  // best fit, at most 1.5x oversized
        void* p = it->second;
        pool.pooled -= it->first;
        pool.live[p] = it->first;
        pool.free_blocks.erase(it);
        return p;
      }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, rounded, cudaHostAllocPortable) == cudaSuccess && p) {
      std::lock_guard<std::mutex> lk(pool.mu);
      pool.live[p] = rounded;
      return p;
    }
    cudaGetLastError();  // pinning refused (ulimit, no device): plain pages
  }
  void* p = aligned_alloc(gran, rounded);
  if (p) madvise(p, rounded, MADV_HUGEPAGE);
  return p;
}

void host_free(void* p) {
  if (!p) return;
  PinnedPool& pool = pinned_pool();
  size_t size = 0;
  {
    std::lock_guard<std::mutex> lk(pool.mu);
    auto it = pool.live.find(p);
    if (it == pool.live.end()) { size = 0; }
    else {
      size = it->second;
      pool.live.erase(it);
      if (pool.pooled + size <= pool.cap) {
        pool.free_blocks.emplace(size, p);
        pool.pooled += size;
        return;
      }
    }
  }
  if (size) cudaFreeHost(p);
  else free(p);
}

namespace {
struct ArrayPrivate {
  const void* buffers[4];
  void* owned[4];
};
void release_array(struct ArrowArray* a) {
  if (!a || !a->release) return;
  ArrayPrivate* p = (ArrayPrivate*)a->private_data;
  if (p) {
    for (int i = 0; i < 4; ++i) host_free(p->owned[i]);
    free(p);
  }
  a->release = nullptr;
  a->private_data = nullptr;
}
struct SchemaPrivate { char* format; char* name; };
void release_schema(struct ArrowSchema* s) {
  if (!s || !s->release) return;
  SchemaPrivate* p = (SchemaPrivate*)s->private_data;
  if (p) { free(p->format); free(p->name); free(p); }
  s->release = nullptr;
  s->private_data = nullptr;
}
}  // namespace

// Takes ownership of `validity` and `values` (malloc'd).  n_extra_buffers = 1 for view arrays (the
// trailing variadic-sizes buffer, empty because every emitted view is inline).
int make_host_array(int64_t length, int64_t null_count, void* validity, void* values, size_t n_extra_buffers,
                    struct ArrowArray* out) {
  ArrayPrivate* p = (ArrayPrivate*)calloc(1, sizeof(ArrayPrivate));
  if (!p) return fail(PW_ERR_INTERNAL, "out of host memory");
  p->owned[0] = validity;
  p->owned[1] = values;
  p->buffers[0] = null_count ? validity : nullptr;
  p->buffers[1] = values;
  if (n_extra_buffers) {
    p->owned[2] = calloc(1, 8);
    p->buffers[2] = p->owned[2];
  }
  memset(out, 0, sizeof(*out));
  out->length = length;
  out->null_count = null_count;
  out->offset = 0;
  out->n_buffers = 2 + (int64_t)n_extra_buffers;
  out->n_children = 0;
  out->buffers = p->buffers;
  out->release = release_array;
  out->private_data = p;
  return 0;
}

// View array whose long values live in one variadic data buffer: buffers = [validity, views, data, sizes]
// (the layout polars-arrow's FFI reads: the variadic buffer sizes travel as the last buffer).
int make_host_view_array(int64_t length, int64_t null_count, void* validity, void* values, void* data, int64_t data_bytes, struct ArrowArray* out) {
  ArrayPrivate* p = (ArrayPrivate*)calloc(1, sizeof(ArrayPrivate));
  if (!p) return fail(PW_ERR_INTERNAL, "out of host memory");
  p->owned[0] = validity;
  p->owned[1] = values;
  p->owned[2] = data;
  p->owned[3] = calloc(1, 8);
  if (!p->owned[3]) { free(p); return fail(PW_ERR_INTERNAL, "out of host memory"); }
  *(int64_t*)p->owned[3] = data_bytes;
  p->buffers[0] = null_count ? validity : nullptr;
  p->buffers[1] = values;
  p->buffers[2] = data;
  p->buffers[3] = p->owned[3];
  memset(out, 0, sizeof(*out));
  out->length = length;
  out->null_count = null_count;
  out->offset = 0;
  out->n_buffers = 4;
  out->n_children = 0;
  out->buffers = p->buffers;
  out->release = release_array;
  out->private_data = p;
  return 0;
}

int make_schema(const char* format, const char* name, bool nullable, struct ArrowSchema* out) {
  SchemaPrivate* p = (SchemaPrivate*)calloc(1, sizeof(SchemaPrivate));
  if (!p) return fail(PW_ERR_INTERNAL, "out of host memory");
  p->format = strdup(format);
  p->name = strdup(name ? name : "");
  memset(out, 0, sizeof(*out));
  out->format = p->format;
  out->name = p->name;
  out->metadata = nullptr;
  out->flags = nullable ? ARROW_FLAG_NULLABLE : 0;
  out->release = release_schema;
  out->private_data = p;
  return 0;
}

}  // namespace pw
