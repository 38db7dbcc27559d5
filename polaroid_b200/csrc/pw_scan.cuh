// pw_scan.cuh — the fused predicate + hash-probe + aggregate pass (SURVEY §8 a1,a2,a3,a4,a6,a8-a11).
//
// One pass over the referenced columns.  Each warp step covers 128 consecutive rows as two 64-row
// halves; inside a half lane l owns the row pair (2l, 2l+1), so every 8-byte column is read with one
// perfectly coalesced 128-bit load per half (narrower types use the matching narrower vector load and
// keep the same row<->lane map).  The predicate is applied in registers (no mask, no compaction), the
// key words are hashed, and the row is aggregated either
//   * into the CTA's shared-memory hot table (the GPU analogue of the reference's FixedIndexTable hot
//     grouper, polars-expr/src/hot_groups/fixed_index_table.rs) which is flushed into the HBM table
//     when it fills up and at the end of the CTA's contiguous row range, or
//   * straight into the HBM open-addressing table (the cold / spill tier: rows whose key does not fit
//     the hot table, and the whole input when the estimated cardinality is high).
// CTAs take contiguous row ranges so that time-sorted inputs keep few live groups per CTA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pw_plan.h"

namespace pw {

constexpr int SCAN_THREADS = 256;
constexpr int ROWS_PER_STEP = 128;  // per warp

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 32; x *= 0xd6e8feb86659fd93ull;
  x ^= x >> 32; x *= 0xd6e8feb86659fd93ull;
  x ^= x >> 32;
  return x;
}
template <int KW>
__device__ __forceinline__ uint64_t hash_words(const uint64_t (&k)[KW]) {
  uint64_t h = mix64(k[0]);
#pragma unroll
  for (int w = 1; w < KW; ++w) h = mix64(h ^ (k[w] + 0x9E3779B97F4A7C15ull * (uint64_t)w));
  return h;
}

// order-preserving map f64 -> int64 (total order, -0 < +0); NaNs never enter (callers skip them)
__device__ __forceinline__ int64_t f64_to_ordered(double d) {
  int64_t b = __double_as_longlong(d);
  return b ^ ((b >> 63) & 0x7FFFFFFFFFFFFFFFll);
}
__host__ __device__ __forceinline__ uint64_t ordered_to_f64_bits(int64_t o) {
  return (uint64_t)(o ^ ((o >> 63) & 0x7FFFFFFFFFFFFFFFll));
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; }

// "Dynamic index" into a register array, written as branch-free mask arithmetic over EVERY element.
// A select chain `(idx == i) ? a[i] : v` gets folded by LLVM back into an indexed load a[idx], which
// forces the array into local memory (seen in SASS as STL.128/LDL.128 on the raw column data); the
// and/or form reads every element unconditionally, so the array stays in registers.
__device__ __forceinline__ uint64_t mask64(bool c) { return (uint64_t)0 - (uint64_t)c; }
__device__ __forceinline__ uint32_t mask32(bool c) { return (uint32_t)0 - (uint32_t)c; }
template <int N>
__device__ __forceinline__ uint64_t pick(const uint64_t (&a)[N], int idx) {
  uint64_t v = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) v |= a[i] & mask64(idx == i);
  return v;
}
template <int N>
__device__ __forceinline__ uint32_t pick32(const uint32_t (&a)[N], int idx) {
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) v |= a[i] & mask32(idx == i);
  return v;
}
template <int N>
__device__ __forceinline__ uint4 pick128(const uint4 (&a)[N], int idx) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const uint32_t m = mask32(idx == i);
    v.x |= a[i].x & m; v.y |= a[i].y & m; v.z |= a[i].z & m; v.w |= a[i].w & m;
  }
  return v;
}
template <int N>
__device__ __forceinline__ void put(uint64_t (&a)[N], int idx, uint64_t x) {
#pragma unroll
  for (int i = 0; i < N; ++i) { const uint64_t m = mask64(idx == i); a[i] = (a[i] & ~m) | (x & m); }
}

template <int V> struct IC { static constexpr int value = V; };

__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

// ---------------------------------------------------------------------------------------------------
// accumulator application (shared and global memory)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void acc_apply_global(uint64_t* p, int op, uint64_t x) {
  switch (op) {
    case OP_ADD_F64: atomicAdd((double*)p, __longlong_as_double((long long)x)); break;
    case OP_ADD_I64: atomicAdd((unsigned long long*)p, (unsigned long long)x); break;
    case OP_MIN_I64: if ((long long)x < (long long)__ldcg((const unsigned long long*)p)) atomicMin((long long*)p, (long long)x); break;
    case OP_MAX_I64: if ((long long)x > (long long)__ldcg((const unsigned long long*)p)) atomicMax((long long*)p, (long long)x); break;
    case OP_MIN_U64: if (x < __ldcg((const unsigned long long*)p)) atomicMin((unsigned long long*)p, (unsigned long long)x); break;
    default: if (x > __ldcg((const unsigned long long*)p)) atomicMax((unsigned long long*)p, (unsigned long long)x); break;
  }
}
__device__ __forceinline__ void acc_apply_shared(uint64_t* p, int op, uint64_t x) {
  switch (op) {
    case OP_ADD_F64: atomicAdd((double*)p, __longlong_as_double((long long)x)); break;
    case OP_ADD_I64: atomicAdd((unsigned long long*)p, (unsigned long long)x); break;
    case OP_MIN_I64: if ((long long)x < *(volatile long long*)p) atomicMin((long long*)p, (long long)x); break;
    case OP_MAX_I64: if ((long long)x > *(volatile long long*)p) atomicMax((long long*)p, (long long)x); break;
    case OP_MIN_U64: if (x < *(volatile unsigned long long*)p) atomicMin((unsigned long long*)p, (unsigned long long)x); break;
    default: if (x > *(volatile unsigned long long*)p) atomicMax((unsigned long long*)p, (unsigned long long)x); break;
  }
}

// ---------------------------------------------------------------------------------------------------
// HBM table: find-or-insert.  Returns the slot, or ~0 when the table is full (overflow flag raised).
// n_kw == 1: the key word itself is the occupancy marker (CAS from KEY_EMPTY); a real key equal to a
//            sentinel lives in an escape slot past `cap`.
// n_kw  > 1: per-slot state word 0 -> 1 (busy, keys being written) -> 2 (ready).  A reader that finds a
//            busy slot retries on the next iteration of a warp-convergent loop (no lane ever waits on
//            another lane inside a divergent branch).
// ---------------------------------------------------------------------------------------------------
template <int KW>
__device__ __forceinline__ uint64_t table_upsert(const Table& T, const uint64_t (&k)[KW], uint64_t h, bool key0_is_sentinel_free) {
  const uint64_t cap = T.cap;
  if (KW == 1) {
    const uint64_t k0 = k[0];
    if (!key0_is_sentinel_free) {
      // k0 is a raw data value that collides with a sentinel -> escape slots
      if (k0 == KEY_EMPTY) { st_volatile_u32(&T.state[cap], 2u); T.keys[cap] = k0; return cap; }
      if (k0 == KEY_NULL) { st_volatile_u32(&T.state[cap + 1], 2u); T.keys[cap + 1] = k0; return cap + 1; }
    }
    uint64_t slot = __umul64hi(h, cap);
    for (uint64_t probes = 0; probes < cap; ++probes) {
      unsigned long long old = __ldcg((const unsigned long long*)&T.keys[slot]);
      if (old == k0) return slot;
      if (old == KEY_EMPTY) {
        old = atomicCAS((unsigned long long*)&T.keys[slot], (unsigned long long)KEY_EMPTY, (unsigned long long)k0);
        if (old == KEY_EMPTY || old == k0) return slot;
      }
      slot = (slot + 1 == cap) ? 0 : slot + 1;
    }
    *T.overflow = 1;
    return ~0ull;
  } else {
    uint64_t slot = __umul64hi(h, cap);
    uint64_t probes = 0;
    uint64_t result = ~0ull;
    bool done = false;
    while (!done) {
      uint32_t s = ld_volatile_u32(&T.state[slot]);
      if (s == 0u) s = atomicCAS(&T.state[slot], 0u, 1u) == 0u ? 3u : 1u;  // 3: we own the slot
      if (s == 3u) {
#pragma unroll
        for (int w = 0; w < KW; ++w) T.keys[(uint64_t)w * (cap + 2) + slot] = k[w];
        __threadfence();
        st_volatile_u32(&T.state[slot], 2u);
        result = slot; done = true;
      } else if (s == 2u) {
        __threadfence();
        bool eq = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) eq &= (__ldcg((const unsigned long long*)&T.keys[(uint64_t)w * (cap + 2) + slot]) == k[w]);
        if (eq) { result = slot; done = true; }
        else {
          slot = (slot + 1 == cap) ? 0 : slot + 1;
          if (++probes >= cap) { *T.overflow = 1; done = true; }
        }
      }
      // s == 1: busy -> look again next iteration
    }
    return result;
  }
}

// ---------------------------------------------------------------------------------------------------
// shared-memory hot table
// layout (bytes): keys[KW][S] u64 | accs[n_acc][S] u64 | state[S] u32 | count u32
// ---------------------------------------------------------------------------------------------------
template <int KW>
struct HotTable {
  uint64_t* keys;
  uint64_t* accs;
  uint32_t* state;
  uint32_t* count;
  int S;
  __device__ __forceinline__ void bind(unsigned char* smem, int slots, int n_acc) {
    S = slots;
    keys = (uint64_t*)smem;
    accs = keys + (size_t)KW * S;
    state = (uint32_t*)(accs + (size_t)n_acc * S);
    count = state + S;
  }
  static __host__ __device__ size_t bytes(int slots, int n_acc) {
    return (size_t)slots * (KW * 8 + n_acc * 8 + 4) + 16;
  }
  __device__ __forceinline__ void clear(const ScanPlan& P) {
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      state[s] = 0u;
      for (int a = 0; a < P.n_acc; ++a) accs[(size_t)a * S + s] = acc_init(P.accs[a].op);
    }
    if (threadIdx.x == 0) *count = 0u;
  }
  // returns slot or -1 when the table is at its load limit / probe limit (row goes to the spill tier)
  __device__ __forceinline__ int upsert(const uint64_t (&k)[KW], uint64_t h) {
    int slot = (int)(h & (uint64_t)(S - 1));
    int probes = 0;
    int result = -1;
    bool done = false;
    const uint32_t limit = (uint32_t)(S - (S >> 2));  // 75 % load
    while (!done) {
      uint32_t s = ld_volatile_u32(&state[slot]);
      if (s == 0u) {
        if (*(volatile uint32_t*)count >= limit) { done = true; continue; }
        s = atomicCAS(&state[slot], 0u, 1u) == 0u ? 3u : 1u;
      }
      if (s == 3u) {
#pragma unroll
        for (int w = 0; w < KW; ++w) keys[(size_t)w * S + slot] = k[w];
        __threadfence_block();
        st_volatile_u32(&state[slot], 2u);
        atomicAdd(count, 1u);
        result = slot; done = true;
      } else if (s == 2u) {
        __threadfence_block();
        bool eq = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) eq &= (*(volatile uint64_t*)&keys[(size_t)w * S + slot] == k[w]);
        if (eq) { result = slot; done = true; }
        else {
          slot = (slot + 1) & (S - 1);
          if (++probes >= 16) done = true;
        }
      }
    }
    return result;
  }
  // move every group into the HBM table and reset (CTA-wide; callers bracket with __syncthreads)
  __device__ __forceinline__ void flush(const ScanPlan& P, bool sentinel_free) {
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      if (state[s] != 2u) continue;
      uint64_t k[KW];
#pragma unroll
      for (int w = 0; w < KW; ++w) k[w] = keys[(size_t)w * S + s];
      uint64_t g = table_upsert<KW>(P.table, k, hash_words<KW>(k), sentinel_free);
      for (int a = 0; a < P.n_acc; ++a) {
        uint64_t v = accs[(size_t)a * S + s];
        int op = P.accs[a].op;
        if (g != ~0ull && v != acc_init(op)) acc_apply_global(&P.table.accs[(uint64_t)a * (P.table.cap + 2) + g], op, v);
        accs[(size_t)a * S + s] = acc_init(op);
      }
      state[s] = 0u;
    }
    if (threadIdx.x == 0) *count = 0u;
  }
};

// ---------------------------------------------------------------------------------------------------
// raw loads: one uint4 per slot per row pair
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dtype_width(int dt) {
  switch (dt) {
    case DT_I8: case DT_U8: return 1;
    case DT_I16: case DT_U16: return 2;
    case DT_I32: case DT_U32: case DT_F32: return 4;
    case DT_VIEW: case DT_VIEW_HI: return 16;
    default: return 8;
  }
}

// p = even LOGICAL row index (pair base); logical rows p and p+1 map to physical rows rb + p*rs.
// `full` = both rows exist, the stride is 1 and vector loads are legal.
__device__ __forceinline__ uint4 load_pair(const RawSlot& s, int64_t p, int64_t n_rows, bool full, int64_t rb = 0, int64_t rs = 1) {
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  const unsigned char* base = (const unsigned char*)s.values;
  const int dt = s.dtype;
  if (dt == DT_VIEW || dt == DT_VIEW_HI) {
    const int64_t lrow = p + (dt == DT_VIEW_HI ? 1 : 0);
    if (lrow < n_rows) {
      const int64_t row = rb + lrow * rs;
      if (full) r = __ldg((const uint4*)(base + row * 16));
      else {
        const uint32_t* q = (const uint32_t*)(base + row * 16);
        r = make_uint4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
      }
    }
    return r;
  }
  const int w = dtype_width(dt);
  if (full) {
    const int64_t pp = rb + p;
    switch (w) {
      case 8: r = __ldg((const uint4*)(base + pp * 8)); break;
      case 4: { uint2 t = __ldg((const uint2*)(base + pp * 4)); r.x = t.x; r.y = t.y; break; }
      case 2: r.x = __ldg((const uint32_t*)(base + pp * 2)); break;
      default: r.x = __ldg((const unsigned short*)(base + pp)); break;
    }
    return r;
  }
  // guarded scalar path (tail rows, strided pilot, unaligned zero-copy frames)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (p + j >= n_rows) break;
    const int64_t row = rb + (p + j) * rs;
    switch (w) {
      case 8: {
        const uint32_t* q = (const uint32_t*)(base + row * 8);
        uint32_t lo = __ldg(q), hi = __ldg(q + 1);
        if (j == 0) { r.x = lo; r.y = hi; } else { r.z = lo; r.w = hi; }
        break; }
      case 4: { uint32_t v = __ldg((const uint32_t*)(base + row * 4)); if (j == 0) r.x = v; else r.y = v; break; }
      case 2: { uint32_t v = __ldg((const unsigned short*)(base + row * 2)); r.x |= v << (16 * j); break; }
      default: { uint32_t v = __ldg(base + row); r.x |= v << (8 * j); break; }
    }
  }
  return r;
}

// validity bits of rows p, p+1 -> bit0, bit1
__device__ __forceinline__ uint32_t load_valid_pair(const RawSlot& s, int64_t p, int64_t n_rows, int64_t rb = 0, int64_t rs = 1) {
  if (s.validity == nullptr) return 3u;
  uint32_t out = 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int64_t b = (int64_t)s.bit_offset + rb + (p + j) * rs;
    if (p + j < n_rows) out |= ((uint32_t)(__ldg(s.validity + (b >> 3)) >> (b & 7)) & 1u) << j;
  }
  return out;
}

// decode row j (0/1) of a raw pair into canonical 64-bit bits (ints sign/zero extended, f32 -> f64)
__device__ __forceinline__ uint64_t decode(const uint4& r, int dt, int j) {
  switch (dt) {
    case DT_I64: case DT_U64: case DT_F64: return j ? ((uint64_t)r.w << 32 | r.z) : ((uint64_t)r.y << 32 | r.x);
    case DT_I32: return (uint64_t)(int64_t)(int32_t)(j ? r.y : r.x);
    case DT_U32: return (uint64_t)(j ? r.y : r.x);
    case DT_F32: return (uint64_t)__double_as_longlong((double)__uint_as_float(j ? r.y : r.x));
    case DT_I16: return (uint64_t)(int64_t)(int16_t)(r.x >> (16 * j));
    case DT_U16: return (uint64_t)((r.x >> (16 * j)) & 0xFFFFu);
    case DT_I8: return (uint64_t)(int64_t)(int8_t)(r.x >> (8 * j));
    case DT_U8: return (uint64_t)((r.x >> (8 * j)) & 0xFFu);
    default: return 0;
  }
}

__device__ __forceinline__ bool compare(uint64_t a, uint64_t b, int cls, int op) {
  int c;
  if (cls == CLS_I64) c = ((int64_t)a > (int64_t)b) - ((int64_t)a < (int64_t)b);
  else if (cls == CLS_U64) c = (a > b) - (a < b);
  else {
    // TotalOrd: NaN == NaN, NaN is the largest value (polars-utils/src/total_ord.rs)
    double x = __longlong_as_double((long long)a), y = __longlong_as_double((long long)b);
    bool xn = x != x, yn = y != y;
    c = (xn || yn) ? ((int)xn - (int)yn) : ((x > y) - (x < y));
  }
  switch (op) {
    case 0: return c == 0;
    case 1: return c != 0;
    case 2: return c < 0;
    case 3: return c <= 0;
    case 4: return c > 0;
    default: return c >= 0;
  }
}

__device__ __forceinline__ int slot_class(int dt) {
  switch (dt) {
    case DT_U8: case DT_U16: case DT_U32: case DT_U64: return CLS_U64;
    case DT_F32: case DT_F64: return CLS_F64;
    default: return CLS_I64;
  }
}

// ---------------------------------------------------------------------------------------------------
// one row: predicate -> key words -> probe -> aggregate.  hf/j are compile-time so that the raw
// register arrays are never indexed dynamically (they must stay in registers).
// ---------------------------------------------------------------------------------------------------
template <int NC, int KW, bool HOT, int hf, int j>
__device__ __forceinline__ void process_row(const ScanPlan& P, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC],
                                            HotTable<KW>& hot, const int64_t base, const int lane, const int64_t n_rows,
                                            unsigned long long& spilled) {
  const bool single_key_sentinel_guard = (KW == 1);
    const int64_t row = base + hf * 64 + 2 * lane + j;
    bool alive = row < n_rows;
    // canonical inputs
    uint64_t in[NC];
    uint32_t in_valid = 0;  // bit per slot
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int dt = (c < P.n_slots) ? P.slots[c].dtype : DT_I64;
      in[c] = decode(raw[hf][c], dt, j);
      in_valid |= ((vbits[hf][c] >> j) & 1u) << c;
    }
    // predicate: conjunction, null => false   (polars-compute/src/filter/mod.rs:18-28)
    for (int q = 0; q < P.n_preds && alive; ++q) {
      const Pred& pr = P.preds[q];
      const bool ok = (in_valid >> pr.slot) & 1u;
      alive = ok && compare(pick<NC>(in, pr.slot), pr.scalar, pr.cls, pr.op);
    }
    // key words
    uint64_t k[KW];
#pragma unroll
    for (int w = 0; w < KW; ++w) k[w] = 0;
    uint64_t nullmask = 0;
    bool sentinel_free = true;
    {
      int w = 0;
      for (int q = 0; q < P.n_keys; ++q) {
        const KeyCol& kc = P.keys[q];
        uint64_t w0, w1 = 0;
        bool ok;
        if (kc.dtype == DT_VIEW) {
          // row j of the pair lives in slot kc.slot + j (even / odd view)
          const uint4 v = pick128<NC>(raw[hf], kc.slot + j);
          w0 = (uint64_t)v.y << 32 | v.x;
          w1 = (uint64_t)v.w << 32 | v.z;
          const uint32_t vb = pick32<NC>(vbits[hf], kc.slot);
          ok = (vb >> j) & 1u;
          if ((uint32_t)w0 > 12u) { if (alive && ok) *P.table.overflow = 2; }  // long string: unsupported here
        } else {
          w0 = pick<NC>(in, kc.slot);
          ok = (in_valid >> kc.slot) & 1u;
          if (kc.dtype == DT_F64 || kc.dtype == DT_F32) {
            double d = __longlong_as_double((long long)w0);
            if (d == 0.0) w0 = 0;                                  // -0.0 == 0.0
            if (d != d) w0 = 0x7FF8000000000000ull;               // one NaN
          }
        }
        if (!ok) { w0 = 0; w1 = 0; nullmask |= 1ull << q; }
put<KW>(k, w, w0);
        if (kc.n_words == 2) put<KW>(k, w + 1, w1);
        w += kc.n_words;
      }
      if (P.dyn.enabled) {
        // tumbling window index (each row belongs to at most one window on this path)
        const int64_t t = (int64_t)pick<NC>(in, P.dyn.slot);
        const int64_t rel = t - P.dyn.origin;
        int64_t kk;
        bool member;
        if (P.dyn.closed == 1) {  // right: (s, s+period]
          kk = floor_div(rel - 1, P.dyn.every);
          member = rel - kk * P.dyn.every <= P.dyn.period;
        } else {
          kk = floor_div(rel, P.dyn.every);
          const int64_t off = rel - kk * P.dyn.every;
          if (P.dyn.closed == 0) member = off < P.dyn.period;              // left  [s, s+period)
          else if (P.dyn.closed == 3) member = off > 0 && off < P.dyn.period;  // none (s, s+period)
          else member = off <= P.dyn.period;                               // both [s, s+period], period < every
        }
        alive = alive && member;
put<KW>(k, w, (uint64_t)kk);
        w += 1;
      }
      if (P.has_null_word) {
put<KW>(k, w, nullmask);
      } else if (KW == 1 && P.n_keys == 1 && !P.dyn.enabled) {
        if (nullmask) k[0] = KEY_NULL;
        else if (k[0] >= KEY_NULL) sentinel_free = false;
      } else if (KW == 1) {
        if (k[0] >= KEY_NULL) sentinel_free = false;
      }
    }
    if (!alive) return;
    const uint64_t h = hash_words<KW>(k);
    // value expressions
    uint64_t v[MAX_VEXPR];
    uint32_t v_valid = 0;
#pragma unroll
    for (int e = 0; e < MAX_VEXPR; ++e) {
      v[e] = 0;
      if (e < P.n_vexpr) {
        const VExpr& ve = P.vexprs[e];
        if (ve.n_factors == 0) {
          v[e] = pick<NC>(in, ve.slot);
          v_valid |= ((in_valid >> ve.slot) & 1u) << e;
        } else {
          double prod = 1.0;
          bool ok = true;
          for (int f = 0; f < ve.n_factors; ++f) {
            const Factor& fc = ve.f[f];
            const uint64_t xb = pick<NC>(in, fc.slot);
            const int cls = slot_class(P.slots[fc.slot].dtype);
            double x = cls == CLS_F64 ? __longlong_as_double((long long)xb)
                                      : (cls == CLS_U64 ? __ull2double_rn(xb) : __ll2double_rn((long long)xb));
            ok = ok && ((in_valid >> fc.slot) & 1u);
            double t = fc.b == 1.0 ? x : __dmul_rn(fc.b, x);
            double u = fc.a == 0.0 ? t : __dadd_rn(fc.a, t);
            prod = f == 0 ? u : __dmul_rn(prod, u);
          }
          v[e] = (uint64_t)__double_as_longlong(prod);
          v_valid |= (ok ? 1u : 0u) << e;
        }
      }
    }
    // probe
    int hslot = -1;
    uint64_t gslot = ~0ull;
    // rows whose raw key aliases a sentinel bypass the hot table so that a hot KEY_NULL is always a
    // true null
    if (HOT && sentinel_free) hslot = hot.upsert(k, h);
    if (hslot < 0) {
      gslot = table_upsert<KW>(P.table, k, h, sentinel_free || !single_key_sentinel_guard);
      if (HOT) ++spilled;
      if (gslot == ~0ull) return;
    }
    // accumulate
    const uint64_t grow = (uint64_t)(P.row_begin + row * P.row_stride + P.row_offset);
    for (int a = 0; a < P.n_acc; ++a) {
      const Acc& ac = P.accs[a];
      const uint64_t bits = pick<MAX_VEXPR>(v, ac.vexpr);
      const bool ok = (v_valid >> ac.vexpr) & 1u;
      uint64_t x = 0;
      bool apply = true;
      switch (ac.src) {
        case SRC_BITS: x = bits; apply = ok; break;
        case SRC_F64: {
          const int cls = P.vexprs[ac.vexpr].cls;
          double d = cls == CLS_F64 ? __longlong_as_double((long long)bits)
                                    : (cls == CLS_U64 ? __ull2double_rn(bits) : __ll2double_rn((long long)bits));
          x = (uint64_t)__double_as_longlong(d); apply = ok; break; }
        case SRC_F64_ORD: {
          double d = __longlong_as_double((long long)bits);
          apply = ok && (d == d);
          x = (uint64_t)f64_to_ordered(d); break; }
        case SRC_VALID: x = ok ? 1ull : 0ull; apply = ok; break;
        case SRC_NOT_NAN: {
          double d = __longlong_as_double((long long)bits);
          apply = ok && (d == d); x = 1ull; break; }
        case SRC_ONE: x = 1ull; break;
        case SRC_ROWIDX: x = (grow << 1) | (ok ? 1ull : 0ull); break;
        case SRC_ROW: x = grow; break;
        default: x = pick<NC>(in, P.dyn.slot); break;  // SRC_INDEX_T
      }
      if (!apply) continue;
      if (hslot >= 0) acc_apply_shared(&hot.accs[(size_t)a * hot.S + hslot], ac.op, x);
      else acc_apply_global(&P.table.accs[(uint64_t)a * (P.table.cap + 2) + gslot], ac.op, x);
    }
}

// ---------------------------------------------------------------------------------------------------
// the scan kernel
// ---------------------------------------------------------------------------------------------------
template <int NC, int KW, bool HOT>
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const __grid_constant__ ScanPlan P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HotTable<KW> hot;
  if (HOT) {
    hot.bind(smem_raw, P.hot_slots, P.n_acc);
    hot.clear(P);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = SCAN_THREADS / 32;
  const int64_t n_rows = P.n_rows;
  const int64_t n_steps = (n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  const int64_t n_tiles = (n_steps + warps - 1) / warps;
  const int64_t tile_lo = n_tiles * blockIdx.x / gridDim.x;
  const int64_t tile_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;
  unsigned long long spilled = 0;

  for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
    const int64_t step = tile * warps + warp;
    const int64_t base = step * ROWS_PER_STEP;
    if (step < n_steps) {
      // ---- phase 1: issue every load of this step (two halves x NC slots) -------------------------
      uint4 raw[2][NC];
      uint32_t vbits[2][NC];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int64_t p = base + hf * 64 + 2 * lane;
        const bool full = P.vec_ok && (p + 1 < n_rows);  // host clears vec_ok unless stride == 1 and row_begin is even
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if (c < P.n_slots && p < n_rows) {
            raw[hf][c] = load_pair(P.slots[c], p, n_rows, full, P.row_begin, P.row_stride);
            vbits[hf][c] = load_valid_pair(P.slots[c], p, n_rows, P.row_begin, P.row_stride);
          } else {
            raw[hf][c] = make_uint4(0u, 0u, 0u, 0u);
            vbits[hf][c] = 0u;
          }
        }
      }
      // ---- sortedness of the dynamic index (every adjacent row pair of this step, plus the row
      //      before the step) — polars-time/src/group_by/dynamic.rs:77-80 raises when it is violated
      if (P.check_sorted) {
        int64_t carry = INT64_MIN;  // t of the row just before this half (lane 0)
        if (lane == 0 && base > 0) {
          const RawSlot& ts = P.slots[P.dyn.slot];
          uint4 r1 = load_pair(ts, base - 1, n_rows, false, P.row_begin, P.row_stride);
          carry = (int64_t)decode(r1, ts.dtype, 0);
        }
        bool bad = false;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int64_t p = base + hf * 64 + 2 * lane;
          const uint4 r = pick128<NC>(raw[hf], P.dyn.slot);
          const int dt = P.slots[P.dyn.slot].dtype;
          const int64_t t0 = (int64_t)decode(r, dt, 0), t1 = (int64_t)decode(r, dt, 1);
          int64_t prev = __shfl_up_sync(0xffffffffu, t1, 1);
          if (lane == 0) prev = carry;
          if (p < n_rows && t0 < prev) bad = true;
          if (p + 1 < n_rows && t1 < t0) bad = true;
          carry = __shfl_sync(0xffffffffu, (p + 1 < n_rows) ? t1 : ((p < n_rows) ? t0 : prev), 31);
        }
        if (bad) *P.not_sorted = 1;
      }
      // ---- phase 2: per row: predicate -> key -> probe -> aggregate ---------------------------------
      process_row<NC, KW, HOT, 0, 0>(P, raw, vbits, hot, base, lane, n_rows, spilled);
      process_row<NC, KW, HOT, 0, 1>(P, raw, vbits, hot, base, lane, n_rows, spilled);
      process_row<NC, KW, HOT, 1, 0>(P, raw, vbits, hot, base, lane, n_rows, spilled);
      process_row<NC, KW, HOT, 1, 1>(P, raw, vbits, hot, base, lane, n_rows, spilled);
    }
    if (HOT) {
      // evict everything when the hot table is getting full (time-sorted inputs drift through groups)
      __syncthreads();
      const bool need_flush = *(volatile uint32_t*)hot.count >= (uint32_t)(hot.S >> 1);
      __syncthreads();
      if (need_flush && tile + 1 < tile_hi) {
        hot.flush(P, true);
        __syncthreads();
      }
    }
  }
  if (HOT) {
    __syncthreads();
    hot.flush(P, true);
    if (spilled) atomicAdd(P.table.spilled, spilled);
  }
}

}  // namespace pw
