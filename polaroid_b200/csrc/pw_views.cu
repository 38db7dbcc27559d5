// pw_views.cu — string / binary keys longer than 12 bytes.
//
// A Utf8View / BinaryView value of more than 12 bytes is [len u32 | first 4 bytes | buffer index u32 | offset u32]
// (polars-arrow/src/array/binview/view.rs:19-55): two equal strings generally carry different (buffer, offset) pairs,
// so the 16 view bytes are not a key.  The reference hashes and compares the bytes (row encoding, hash_keys.rs:114-141;
// BinviewKeyIdxTable).  Here the column is CANONICALISED once, when the frame is created: a device hash set over
// the long strings of the column (hash of the bytes, full byte compare on a candidate) picks one representative per
// distinct string, and every long view is rewritten to point at its representative's bytes.  After that, equality of
// the 16 view bytes is equality of the strings — for inline and long values alike — and every group-by tier (hash
// words, hot table, HBM table, buckets, radix partitions) takes string keys of any length unchanged.
// When a result carries long keys, the referenced bytes are gathered into one fresh data buffer and the emitted views
// are re-pointed at it (views_gather_long), so the result owns its bytes like any Arrow array.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "pw_engine.h"

namespace pw {
namespace {

__device__ __forceinline__ const unsigned char* view_bytes(const uint4& v, const void* const* bufs) { return (const unsigned char*)bufs[v.z] + v.w; }

__device__ __forceinline__ uint64_t bytes_hash(const unsigned char* p, uint32_t len) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ len;
  uint32_t i = 0;
  for (; i + 8 <= len; i += 8) {
    uint64_t w = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) w |= (uint64_t)p[i + b] << (8 * b);
    h = (h ^ w) * 0xd6e8feb86659fd93ull;
    h ^= h >> 32;
  }
  uint64_t w = 0;
  for (int b = 0; i < len; ++i, ++b) w |= (uint64_t)p[i] << (8 * b);
  h = (h ^ w) * 0xd6e8feb86659fd93ull;
  h ^= h >> 29;
  return h * 0x55fbfd6bfc5458e9ull;
}

__device__ __forceinline__ bool bytes_equal(const unsigned char* a, const unsigned char* b, uint32_t len) {
  if (a == b) return true;
  for (uint32_t i = 0; i < len; ++i)
    if (a[i] != b[i]) return false;
  return true;
}

__device__ __forceinline__ bool row_valid(const uint8_t* validity, int32_t bit_offset, int64_t i) {
  if (!validity) return true;
  const int64_t b = i + bit_offset;
  return (validity[b >> 3] >> (b & 7)) & 1;
}

__global__ void count_long_kernel(const uint4* views, const uint8_t* validity, int32_t bit_offset, int64_t n, unsigned long long* n_long) {
  unsigned long long mine = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (views[i].x > 12u && row_valid(validity, bit_offset, i)) ++mine;
  for (int d = 16; d >= 1; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_long, mine);
}

// set[slot] = row + 1 of the representative (0 = empty).  The representative of a string is whichever of its rows wins
// the slot; the input views are never written, so a reader that meets a fresh entry compares against stable bytes.
__global__ void intern_kernel(const uint4* views, const uint8_t* validity, int32_t bit_offset, int64_t n, const void* const* bufs,
                              uint32_t* set, uint64_t cap, uint4* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 v = views[i];
    if (v.x > 12u && row_valid(validity, bit_offset, i)) {
      const unsigned char* p = view_bytes(v, bufs);
      uint64_t slot = __umul64hi(bytes_hash(p, v.x), cap);
      for (;;) {
        uint32_t cur = *(volatile uint32_t*)&set[slot];
        if (cur == 0u) {
          cur = atomicCAS(&set[slot], 0u, (uint32_t)i + 1u);
          if (cur == 0u) break;   // this row represents the string
        }
        const uint4 r = views[cur - 1u];
        if (r.x == v.x && r.y == v.y && bytes_equal(p, view_bytes(r, bufs), v.x)) { v.z = r.z; v.w = r.w; break; }
        slot = slot + 1 == cap ? 0 : slot + 1;
      }
    } else if (v.x > 12u) {
      v = make_uint4(0u, 0u, 0u, 0u);   // a null row's view bytes are unspecified: keep them harmless
    }
    out[i] = v;
  }
}

struct Seg { const unsigned char* src; uint64_t dst; uint32_t len; uint32_t pad; };
__global__ void gather_kernel(const Seg* segs, uint64_t n_segs, unsigned char* out) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_segs) return;
  const Seg s = segs[w];
  for (uint32_t i = lane; i < s.len; i += 32) out[s.dst + i] = s.src[i];
}

}  // namespace

// Canonicalises the long views of a freshly imported view column (see the header comment).  `col->values` must be the
// device views, `col->var_bufs` the device data buffers.  No-op (and no synchronisation) for columns without data buffers.
int views_intern(FrameColumn* col, int64_t n) {
  if (col->dtype != DT_VIEW || col->var_bufs.empty() || n == 0) return 0;
  if (n >= ((int64_t)1 << 32) - 1) return fail(PW_ERR_UNSUPPORTED, "column '%s': long string keys in frames of 2^32 rows or more", col->name.c_str());
  ThreadCtx& c = ctx();
  void* p = nullptr;
  PW_TRY(dev_alloc(&p, (col->var_bufs.size() + 2) * 8));
  col->d_var_ptrs = (const void**)p;
  unsigned long long* d_count = (unsigned long long*)p + col->var_bufs.size();
  PW_CUDA(cudaMemcpyAsync(p, col->var_bufs.data(), col->var_bufs.size() * sizeof(void*), cudaMemcpyHostToDevice, c.stream));
  PW_CUDA(cudaMemsetAsync(d_count, 0, 8, c.stream));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  count_long_kernel<<<grid, 256, 0, c.stream>>>((const uint4*)col->values, col->validity, col->bit_offset, n, d_count);
  PW_CUDA(cudaGetLastError());
  unsigned long long n_long = 0;
  PW_CUDA(cudaMemcpyAsync(&n_long, d_count, 8, cudaMemcpyDeviceToHost, c.stream));
  PW_CUDA(cudaStreamSynchronize(c.stream));
  c.timings.kernel_launches++;
  if (n_long == 0) return 0;
  for (const void* b : col->var_bufs)
    if (!b) return fail(PW_ERR_INVALID, "column '%s' holds values longer than 12 bytes but a data buffer is missing", col->name.c_str());
  const uint64_t cap = 2 * (uint64_t)n_long + 64;
  void* set = nullptr;
  void* canon = nullptr;
  PW_TRY(dev_alloc(&set, cap * 4));
  if (int rc = dev_alloc(&canon, (size_t)n * 16 + 32)) { dev_free(set); return rc; }
  if (cudaMemsetAsync(set, 0, cap * 4, c.stream) != cudaSuccess) { dev_free(set); dev_free(canon); return fail(PW_ERR_CUDA, "memset failed"); }
  intern_kernel<<<grid, 256, 0, c.stream>>>((const uint4*)col->values, col->validity, col->bit_offset, n, col->d_var_ptrs, (uint32_t*)set, cap, (uint4*)canon);
  const cudaError_t e = cudaGetLastError();
  const cudaError_t e2 = cudaStreamSynchronize(c.stream);
  dev_free(set);
  if (e != cudaSuccess || e2 != cudaSuccess) { dev_free(canon); return fail(PW_ERR_CUDA, "intern_kernel failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2)); }
  c.timings.kernel_launches++;
  dev_free(col->owned_values);   // the uploaded copy (nullptr for zero-copy frames: the caller's buffer is left alone)
  col->values = canon; col->owned_values = canon;
  col->has_long = true;
  return 0;
}

// Result side: `h_views` = G emitted views on the host.  Long ones reference the frame's device data buffers; their
// bytes are gathered into one host buffer (*h_data, host_alloc'd, *data_bytes long) and the views re-pointed at it
// (buffer 0).  *h_data stays nullptr when the result holds no long value.
int views_gather_long(const FrameColumn* col, void* h_views, const void* h_validity, uint64_t G, void** h_data, int64_t* data_bytes) {
  *h_data = nullptr; *data_bytes = 0;
  if (!col || !col->has_long || G == 0) return 0;
  ThreadCtx& c = ctx();
  uint32_t* v = (uint32_t*)h_views;
  const uint32_t* valid = (const uint32_t*)h_validity;
  std::vector<Seg> segs;
  uint64_t total = 0;
  for (uint64_t i = 0; i < G; ++i) {
    const uint32_t len = v[4 * i];
    if (len <= 12u) continue;
    if (valid && !((valid[i >> 5] >> (i & 31)) & 1u)) continue;
    const uint32_t b = v[4 * i + 2], off = v[4 * i + 3];
    if (b >= col->var_bufs.size()) return fail(PW_ERR_INTERNAL, "view references data buffer %u of %zu", b, col->var_bufs.size());
    segs.push_back(Seg{(const unsigned char*)col->var_bufs[b] + off, total, len, 0u});
    v[4 * i + 2] = 0u; v[4 * i + 3] = (uint32_t)total;
    total += len;
    if (total > 0xFFFFFFFFull) return fail(PW_ERR_UNSUPPORTED, "more than 4 GiB of long string keys in one result");
  }
  if (segs.empty()) return 0;
  void* d_segs = nullptr;
  void* d_out = nullptr;
  PW_TRY(dev_alloc(&d_segs, segs.size() * sizeof(Seg)));
  if (int rc = dev_alloc(&d_out, total)) { dev_free(d_segs); return rc; }
  void* h = host_alloc(total);
  auto drop = [&]() { dev_free(d_segs); dev_free(d_out); };
  if (!h) { drop(); return fail(PW_ERR_INTERNAL, "out of host memory"); }
  bool ok = cudaMemcpyAsync(d_segs, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice, c.stream) == cudaSuccess;
  if (ok) {
    const uint64_t threads = (uint64_t)segs.size() * 32;
    gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, c.stream>>>((const Seg*)d_segs, segs.size(), (unsigned char*)d_out);
    ok = cudaGetLastError() == cudaSuccess && cudaMemcpyAsync(h, d_out, total, cudaMemcpyDeviceToHost, c.stream) == cudaSuccess &&
         cudaStreamSynchronize(c.stream) == cudaSuccess;
    c.timings.kernel_launches++;
  }
  drop();
  if (!ok) { host_free(h); return fail(PW_ERR_CUDA, "gathering long string keys failed: %s", cudaGetErrorString(cudaGetLastError())); }
  *h_data = h; *data_bytes = (int64_t)total;
  return 0;
}

}  // namespace pw
