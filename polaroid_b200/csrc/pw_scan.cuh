// pw_scan.cuh — the fused predicate + hash-probe + aggregate pass (SURVEY §8 a1,a2,a3,a4,a6,a8-a11).
//
// One pass over the referenced columns.  Each warp step covers 128 consecutive rows as two 64-row
// halves; inside a half lane l owns the row pair (2l, 2l+1), so every 8-byte column is read with one
// perfectly coalesced 128-bit load per half (narrower types use the matching narrower vector load and
// keep the same row<->lane map).  The predicate is applied in registers (no mask, no compaction), the
// key words are hashed, and the row is aggregated either
//   * into the CTA's shared-memory HOT TABLE — the GPU analogue of the reference's FixedIndexTable hot
//     grouper (polars-expr/src/hot_groups/fixed_index_table.rs) — or
//   * straight into the HBM open-addressing table (the cold / spill tier: rows whose key does not fit
//     the hot table, and the whole input when consecutive rows do not share groups).
//
// Hot table layout (why it looks the way it does: a shared-memory atomic costs ~64 cycles per warp on
// this part, a plain LDS/STS 2-6, so the per-row updates must not be atomics):
//   CTA-shared  key index   buckets of four 32-bit tags (fingerprint | dense id) read with one LDS.128,
//                           keys stored per dense id -> a probe costs ~1 iteration for every lane (linear
//                           probing cost the max chain length over the 32 lanes, ~6.5 iterations)
//   CTA-shared  min/max     one 64-bit word per (acc, id): read, compare, atomic only when it improves
//                           (used for min/max words when private copies would not fit)
//   warp-private words      R replicas per (acc, id): replica = lane % R.  A row's lane first CLAIMS
//                           (id, replica) by writing its lane number and reading it back; winners do a
//                           plain read-modify-write, losers retry.  R = 32 (tiny group counts, e.g.
//                           TPC-H Q1's 4 groups) makes every lane its own replica: no claims at all.
// The hot table is flushed into the HBM table when it fills up (time-sorted inputs drift through
// groups) and at the end of the CTA's contiguous row range.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "pw_ctl.h"
#include "pw_plan.h"

namespace pw {

// CTA size per raw-slot class: the narrow class (<= 4 slots, ~125 registers) runs 12 warps so that one
// CTA per SM still hides HBM latency when the hot table needs most of the shared memory; the wide class
// (<= 12 slots, ~240 registers) is register-limited to 8 warps.
template <int NC> struct ScanCfg { static constexpr int THREADS = NC <= 4 ? 384 : 256; };
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
// One key word: a single multiply, consumers take the TOP bits (bucket = bits 32.., fingerprint = bits 48..,
// HBM slot = umulhi(h, cap)) — the reference's DirtyHash + hash_to_partition construction
// (polars-utils/src/hashing.rs:62-69, 124-151).  Several words: a full mix per word.
template <int KW>
__device__ __forceinline__ uint64_t hash_words(const uint64_t (&k)[KW]) {
  if (KW == 1) return k[0] * 0x55fbfd6bfc5458e9ull;
  // several words: one multiply per word (xor-fold, add the next word, multiply) and one finishing round — the first
  // version ran a full two-multiply mixer per word (8 64-bit multiplies per row for Q1's two string keys)
  uint64_t h = k[0] * 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int w = 1; w < KW; ++w) h = ((h ^ (h >> 32)) + k[w]) * 0xd6e8feb86659fd93ull;
  h ^= h >> 32; h *= 0x55fbfd6bfc5458e9ull;
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

__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

// ---------------------------------------------------------------------------------------------------
// accumulator application
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t acc_combine(int op, uint64_t a, uint64_t b) {
  switch (op) {
    case OP_ADD_F64: return (uint64_t)__double_as_longlong(__longlong_as_double((long long)a) + __longlong_as_double((long long)b));
    case OP_ADD_I64: return a + b;
    case OP_MIN_I64: return (int64_t)b < (int64_t)a ? b : a;
    case OP_MAX_I64: return (int64_t)b > (int64_t)a ? b : a;
    case OP_MIN_U64: return b < a ? b : a;
    case OP_AND_U64: return a & b;
    case OP_OR_U64: return a | b;
    case OP_XOR_U64: return a ^ b;
    default: return b > a ? b : a;
  }
}
__device__ __forceinline__ void acc_apply_global(uint64_t* p, int op, uint64_t x) {
  switch (op) {
    case OP_ADD_F64: atomicAdd((double*)p, __longlong_as_double((long long)x)); break;
    case OP_ADD_I64: atomicAdd((unsigned long long*)p, (unsigned long long)x); break;
    case OP_MIN_I64: if ((long long)x < (long long)__ldcg((const unsigned long long*)p)) atomicMin((long long*)p, (long long)x); break;
    case OP_MAX_I64: if ((long long)x > (long long)__ldcg((const unsigned long long*)p)) atomicMax((long long*)p, (long long)x); break;
    case OP_MIN_U64: if (x < __ldcg((const unsigned long long*)p)) atomicMin((unsigned long long*)p, (unsigned long long)x); break;
    case OP_AND_U64: atomicAnd((unsigned long long*)p, (unsigned long long)x); break;
    case OP_OR_U64: atomicOr((unsigned long long*)p, (unsigned long long)x); break;
    case OP_XOR_U64: atomicXor((unsigned long long*)p, (unsigned long long)x); break;
    default: if (x > __ldcg((const unsigned long long*)p)) atomicMax((unsigned long long*)p, (unsigned long long)x); break;
  }
}
// shared-memory min/max word: read, compare, atomic only when the row improves the extremum
__device__ __forceinline__ void minmax_apply_shared(uint64_t* p, int op, uint64_t x) {
  switch (op) {
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
//            busy slot retries on the next iteration of its loop (it never waits inside a branch that
//            the slot owner has to leave first).
// ---------------------------------------------------------------------------------------------------
template <int KW>
__device__ __forceinline__ uint64_t table_upsert(const Table& T, const uint64_t (&k)[KW], uint64_t h, bool key0_is_sentinel_free,
                                                 uint64_t max_probes = ~0ull) {
  const uint64_t cap = T.cap;
  const uint64_t lim = max_probes < cap ? max_probes : cap;   // pw_radix.cuh bounds the probe sequence of its overflow region
  if (KW == 1) {
    const uint64_t k0 = k[0];
    if (!key0_is_sentinel_free) {
      // k0 is a raw data value that collides with a sentinel -> escape slots
      if (k0 == KEY_EMPTY) { st_volatile_u32(&T.state[cap], 2u); tkey(T, 0, cap) = k0; return cap; }
      if (k0 == KEY_NULL) { st_volatile_u32(&T.state[cap + 1], 2u); tkey(T, 0, cap + 1) = k0; return cap + 1; }
    }
    uint64_t slot = __umul64hi(h, cap);
    for (uint64_t probes = 0; probes < lim; ++probes) {
      unsigned long long old = __ldcg((const unsigned long long*)&tkey(T, 0, slot));
      if (old == k0) return slot;
      if (old == KEY_EMPTY) {
        old = atomicCAS((unsigned long long*)&tkey(T, 0, slot), (unsigned long long)KEY_EMPTY, (unsigned long long)k0);
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
        for (int w = 0; w < KW; ++w) tkey(T, w, slot) = k[w];
        __threadfence();
        st_volatile_u32(&T.state[slot], 2u);
        result = slot; done = true;
      } else if (s == 2u) {
        __threadfence();
        bool eq = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) eq &= (__ldcg((const unsigned long long*)&tkey(T, w, slot)) == k[w]);
        if (eq) { result = slot; done = true; }
        else {
          slot = (slot + 1 == cap) ? 0 : slot + 1;
          if (++probes >= lim) { *T.overflow = 1; done = true; }
        }
      }
      // s == 1: busy -> look again next iteration
    }
    return result;
  }
}

// ---------------------------------------------------------------------------------------------------
// shared-memory hot table (see the header comment).  Geometry comes from ScanPlan::hot via CT:
//   [tag u32 x S (buckets of 4)][keys u64 x KW x G][minmax u64 x n_mm x G][count, full-flag]   CTA-shared
//   per warp: [private words: G x R x (8|4) bytes each][claim u8 x G x R]                      warp-private
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t TAG_BUSY = 0xFFFFFFFFu;

__device__ __forceinline__ uint4 lds128_volatile(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

template <class CT, int KW>
struct HotTable {
  uint32_t* tag;
  uint64_t* keys;
  uint64_t* mm;
  uint32_t* count;       // [0] dense ids handed out, [1] "a row found the table full"
  unsigned char* wbase;  // this warp's private region
  unsigned char* wall;   // warp 0's private region
  // 32-bit SHADOW of one CTA-shared (min, max) pair: per id the high words of the two 64-bit images, always on the
  // safe side of the truth (shadow min >= true min, shadow max <= true max in the high word), so
  // `hi(x) <= shadow.min || hi(x) >= shadow.max` catches every row that could improve an extremum.  The common row
  // reads 8 bytes and does two 32-bit compares instead of 16 bytes and two 64-bit compares; the rare row takes the
  // exact path and tightens the shadow with native 32-bit shared-memory atomics.
  int2* shadow;
  __device__ __forceinline__ void bind(unsigned char* smem, const ScanPlan& P, int warp) {
    tag = (uint32_t*)smem;
    keys = (uint64_t*)(smem + CT::h_keys_off(P));
    mm = (uint64_t*)(smem + CT::h_mm_off(P));
    count = (uint32_t*)(smem + CT::h_count_off(P));
    wall = smem + CT::h_warp_off(P);
    wbase = wall + (size_t)warp * CT::h_warp_bytes(P);
    shadow = (int2*)(smem + CT::h_shadow_off(P));
  }
  // reset everything (CTA-wide; caller syncs)
  __device__ __forceinline__ void clear(const ScanPlan& P) {
    const int S = CT::h_slots(P), G = CT::h_gcap(P), R = CT::h_rep(P), W = blockDim.x >> 5;
    for (int s = threadIdx.x; s < S; s += blockDim.x) tag[s] = 0u;
    for (int a = 0; a < CT::n_acc(P); ++a) {
      const uint64_t init = acc_init(CT::acc_op(P, a));
      const int kind = CT::h_kind(P, a);
      if (kind == HOT_SHARED_MM) {
        uint64_t* p = mm + CT::h_off(P, a);
        for (int i = threadIdx.x; i < G; i += blockDim.x) p[(size_t)i * CT::h_mm_stride(P)] = init;
      } else {
        const int cells = G * R;
        for (int w = 0; w < W; ++w) {
          unsigned char* base = wall + (size_t)w * CT::h_warp_bytes(P) + CT::h_off(P, a);
          for (int i = threadIdx.x; i < cells; i += blockDim.x) {
            if (kind == HOT_PRIV64) ((uint64_t*)base)[i] = init;
            else ((uint32_t*)base)[i] = 0u;
          }
        }
      }
    }
    if (CT::h_guard_acc(P) >= 0)
      for (int i = threadIdx.x; i < G; i += blockDim.x) shadow[i] = make_int2(0x7FFFFFFF, (int)0x80000000);
    if (threadIdx.x == 0) { count[0] = 0u; count[1] = 0u; }
  }
  __device__ __forceinline__ bool key_equals(const ScanPlan& P, int id, const uint64_t (&k)[KW]) const {
    const int G = CT::h_gcap(P);
    bool eq = true;
#pragma unroll
    for (int w = 0; w < KW; ++w) eq &= (*(volatile uint64_t*)&keys[(size_t)w * G + id] == k[w]);
    return eq;
  }
  // key -> dense group id.  FAST PATH (every row once its group exists): one LDS.128 of the home bucket's four
  // tags, a branch-free fingerprint match, one key compare; -1 = not there (caller takes the slow path for the
  // warp's missing rows).  Fingerprints live in [1, 0xFFFE] so that neither an empty (0) nor a busy (all ones)
  // tag can match.
  // The lookup is split in two so that a lane's rows can issue all their tag loads, then all their key loads
  // (volatile shared-memory loads keep program order: back-to-back whole lookups would serialise the round trips).
  __device__ __forceinline__ uint4 lookup_tags(const ScanPlan& P, uint64_t h) const {
    const int bmask = (CT::h_slots(P) >> 2) - 1;
    const int bucket = (int)((uint32_t)(h >> 32) & (uint32_t)bmask);
    return lds128_volatile(tag + bucket * 4);
  }
  // fingerprint of a hash: 15 bits + 1, i.e. [1, 0x8000] — never an empty (0) or busy (all ones) tag
  static __device__ __forceinline__ uint32_t fingerprint(uint64_t h) { return (uint32_t)(h >> 49) + 1u; }
  // -> matching tag (fingerprint << 16 | dense id) or 0
  __device__ __forceinline__ uint32_t lookup_match(const uint4& ta, uint64_t h) const {
    const uint32_t f16 = fingerprint(h);
    uint32_t ca = 0;
    ca = ((ta.w >> 16) == f16) ? ta.w : ca; ca = ((ta.z >> 16) == f16) ? ta.z : ca;
    ca = ((ta.y >> 16) == f16) ? ta.y : ca; ca = ((ta.x >> 16) == f16) ? ta.x : ca;
    return ca;
  }
  // SLOW PATH: insertion, fingerprint collisions, keys that live in a neighbour of an overflowing bucket.
  // -1 when the table is full or the probe budget is spent (row goes cold).
  struct Key { uint64_t w[KW]; };  // by value: a reference would force the caller's key registers onto the stack
  __device__ __forceinline__ int upsert_slow(const ScanPlan& P, const uint64_t (&k)[KW], uint64_t h) const {
    Key kv;
#pragma unroll
    for (int w = 0; w < KW; ++w) kv.w[w] = k[w];
    return upsert_slow_fn(*this, P, kv, h);
  }
  static __device__ __noinline__ int upsert_slow_fn(HotTable hot, const ScanPlan& P, Key kv, uint64_t h) {
    uint32_t* const tag = hot.tag;
    uint64_t* const keys = hot.keys;
    uint32_t* const count = hot.count;
    uint64_t k[KW];
#pragma unroll
    for (int w = 0; w < KW; ++w) k[w] = kv.w[w];
    const int S = CT::h_slots(P), G = CT::h_gcap(P);
    const uint32_t fp = fingerprint(h) << 16;
    const int bmask = (S >> 2) - 1;
    int bucket = (int)((uint32_t)(h >> 32) & (uint32_t)bmask);
    int probes = 0;
    int result = -1;
    bool done = false;
    while (!done) {
      uint32_t* tb = tag + bucket * 4;
      const uint4 t4 = lds128_volatile(tb);
      const uint32_t t[4] = {t4.x, t4.y, t4.z, t4.w};
      bool busy = false;
      int empty = -1;
#pragma unroll
      for (int i = 3; i >= 0; --i) {
        if (t[i] == TAG_BUSY) busy = true;
        else if (t[i] == 0u) empty = i;
        else if ((t[i] & 0xFFFF0000u) == fp && result < 0) {
          const int id = (int)(t[i] & 0xFFFFu);
          if (hot.key_equals(P, id, k)) result = id;
        }
      }
      if (result >= 0) { done = true; }
      else if (busy) { /* someone is inserting into this bucket (maybe our key): look again */ }
      else if (empty >= 0) {
        if (*(volatile uint32_t*)count >= (uint32_t)G) { count[1] = 1u; done = true; }  // full: ask for an eviction
        else if (atomicCAS(&tb[empty], 0u, TAG_BUSY) == 0u) {
          const uint32_t id = atomicAdd(count, 1u);
          if (id < (uint32_t)G) {
#pragma unroll
            for (int w = 0; w < KW; ++w) keys[(size_t)w * G + id] = k[w];
            __threadfence_block();
            st_volatile_u32(&tb[empty], fp | id);
            result = (int)id;
          } else {
            st_volatile_u32(&tb[empty], 0u);  // lost the race for the last ids: give the slot back
            count[1] = 1u;
          }
          done = true;
        }
        // CAS lost: look again
      } else {
        bucket = (bucket + 1) & bmask;
        if (++probes >= 8) done = true;
      }
    }
    return result;
  }
  // move every group into the HBM table (CTA-wide; callers bracket with __syncthreads and clear())
  __device__ __forceinline__ void flush(const ScanPlan& P) {
    const int G = CT::h_gcap(P), R = CT::h_rep(P), W = blockDim.x >> 5;
    const int n = CT::h_dense(P) ? G : min((int)count[0], G);
    for (int id = threadIdx.x; id < n; id += blockDim.x) {
      uint64_t k[KW];
      if (CT::h_dense(P)) {
        // dense ids: the id exists iff a row was counted into it (the planner guarantees a LEN word)
        const int a = CT::acc_gbase(P);
        uint32_t rows = 0;
        for (int w = 0; w < W; ++w) {
          const uint32_t* base = (const uint32_t*)(wall + (size_t)w * CT::h_warp_bytes(P) + CT::h_off(P, a));
          for (int r = 0; r < R; ++r) rows += base[(size_t)id * R + r] & (a == CT::h_claim_acc(P) ? 0x00FFFFFFu : 0xFFFFFFFFu);
        }
        if (rows == 0u) continue;
#pragma unroll
        for (int w = 0; w < KW; ++w) k[w] = w == 0 ? (uint64_t)P.dense_min + (uint64_t)id : 0ull;
      } else {
#pragma unroll
        for (int w = 0; w < KW; ++w) k[w] = keys[(size_t)w * G + id];
      }
      // a hot KEY_NULL is always a true null (raw sentinel-valued keys bypass the hot table)
      const uint64_t gs = table_upsert<KW>(P.table, k, hash_words<KW>(k), true);
      for (int a = 0; a < CT::n_acc(P); ++a) {
        const int op = CT::acc_op(P, a);
        const int kind = CT::h_kind(P, a);
        uint64_t v;
        if (kind == HOT_SHARED_MM) {
          v = mm[(size_t)id * CT::h_mm_stride(P) + CT::h_off(P, a)];
        } else {
          v = acc_init(op);  // combine the replicas of every warp
          for (int w = 0; w < W; ++w) {
            const unsigned char* base = wall + (size_t)w * CT::h_warp_bytes(P) + CT::h_off(P, a);
            for (int r = 0; r < R; ++r) {
              if (kind == HOT_PRIV64) v = acc_combine(op, v, ((const uint64_t*)base)[(size_t)id * R + r]);
              else v += ((const uint32_t*)base)[(size_t)id * R + r] & (a == CT::h_claim_acc(P) ? 0x00FFFFFFu : 0xFFFFFFFFu);
            }
          }
        }
        if (gs != ~0ull && v != acc_init(op)) acc_apply_global(&tacc(P.table, a, gs), op, v);
      }
    }
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
    case DT_BOOL: return 0;  // bit-packed: see load_bool_pair
    default: return 8;
  }
}

// p = even LOGICAL row index (pair base); logical rows p and p+1 map to physical rows rb + p*rs.
// `full` = both rows exist, the stride is 1 and vector loads are legal.
__device__ __forceinline__ uint4 load_pair(const void* values, int dt, int64_t p, int64_t n_rows, bool full, int64_t rb = 0, int64_t rs = 1) {
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  const unsigned char* base = (const unsigned char*)values;
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
__device__ __forceinline__ uint4 load_pair(const RawSlot& s, int64_t p, int64_t n_rows, bool full, int64_t rb = 0, int64_t rs = 1) {
  return load_pair(s.values, s.dtype, p, n_rows, full, rb, rs);
}

// boolean VALUES are bit-packed like validity (LSB first, same bit offset as the column's validity): rows p, p+1 -> bit0, bit1
__device__ __forceinline__ uint4 load_bool_pair(const RawSlot& s, int64_t p, int64_t n_rows, int64_t rb = 0, int64_t rs = 1) {
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  const uint8_t* bits = (const uint8_t*)s.values;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int64_t b = (int64_t)s.bit_offset + rb + (p + j) * rs;
    if (p + j < n_rows) r.x |= ((uint32_t)(__ldg(bits + (b >> 3)) >> (b & 7)) & 1u) << j;
  }
  return r;
}

// one row of any column type (scalar path)
__device__ __forceinline__ uint4 load_row(const RawSlot& s, int64_t row) {
  return s.dtype == DT_BOOL ? load_bool_pair(s, row, row + 1) : load_pair(s.values, s.dtype, row, row + 1, false);
}

// validity bits of logical rows p, p+1 -> bit0, bit1
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
    case DT_BOOL: return (uint64_t)((r.x >> j) & 1u);
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
    case DT_U8: case DT_U16: case DT_U32: case DT_U64: case DT_BOOL: return CLS_U64;
    case DT_F32: case DT_F64: return CLS_F64;
    default: return CLS_I64;
  }
}
__device__ __forceinline__ double bits_to_f64(uint64_t b, int cls) {
  return cls == CLS_F64 ? __longlong_as_double((long long)b) : (cls == CLS_U64 ? __ull2double_rn(b) : __ll2double_rn((long long)b));
}

// ---------------------------------------------------------------------------------------------------
// row evaluation shared by the scan, slice and segmented kernels.
//
// FRONT END (one copy): select the row's raw registers, decode, apply the predicate, build the key words,
// evaluate the value expressions -> RowOut (a handful of scalars).
// BACK END (one copy): hash, probe, and one pass over the per-expression aggregate flags.  Every register
// array is indexed with compile-time constants only.  History: the first version selected operands per
// accumulator with runtime indices and inlined the whole row four times; it executed ~57 warp instructions
// per ROW and stalled on instruction fetch (profiles/r01_*.txt).
// ---------------------------------------------------------------------------------------------------
template <int NC> struct NVof { static constexpr int value = NC <= 4 ? 2 : MAX_VEXPR; };

template <int NC>
struct Row {
  uint64_t in[NC];      // canonical inputs per raw slot
  uint32_t in_valid;    // bit per slot
};

template <int KW, int NV>
struct RowOut {
  uint64_t k[KW];
  uint64_t v[NV];
  uint32_t v_valid;
  bool alive, sentinel_free;
  int64_t row;    // logical row
  uint64_t tval;  // dynamic index value
};

template <class CT, int NC>
__device__ __forceinline__ void row_decode(const ScanPlan& P, const uint4 (&raw)[NC], const uint32_t (&vbits)[NC], int j, Row<NC>& r) {
  r.in_valid = 0;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < CT::n_slots(P)) {
      r.in[c] = decode(raw[c], CT::slot_dtype(P, c), j);
      // a slot without a validity bitmap is valid by construction (rows past the end are never alive), which lets
      // the compiler drop every null check downstream
      r.in_valid |= (CT::slot_nullable(P, c) ? ((vbits[c] >> j) & 1u) : 1u) << c;
    } else r.in[c] = 0;
  }
  // partitioned input: the validity bits of the original row travel in the low byte of the row-id word
  if (CT::rowid_slot(P) >= 0) r.in_valid = (uint32_t)pick<NC>(r.in, CT::rowid_slot(P)) & 0xFFu;
}

// predicate: conjunction, null => false   (polars-compute/src/filter/mod.rs:18-28)
template <class CT, int NC>
__device__ __forceinline__ bool row_predicate(const ScanPlan& P, const Row<NC>& r) {
  bool alive = true;
#pragma unroll
  for (int q = 0; q < MAX_PREDS; ++q) {
    if (q >= CT::n_preds(P)) break;
    const int slot = CT::pred_slot(P, q);
    const bool ok = (r.in_valid >> slot) & 1u;
    alive = alive && ok && compare(pick<NC>(r.in, slot), P.preds[q].scalar, CT::pred_cls(P, q), CT::pred_op(P, q));
  }
  return alive;
}

template <class CT, int NC, int NV>
__device__ __forceinline__ void row_vexprs(const ScanPlan& P, const Row<NC>& r, uint64_t (&v)[NV], uint32_t& v_valid) {
  v_valid = 0;
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    v[e] = 0;
    if (e < CT::n_vexpr(P)) {
      if (CT::ve_nf(P, e) == 0) {
        const int slot = CT::ve_slot(P, e);
        v[e] = pick<NC>(r.in, slot);
        v_valid |= ((r.in_valid >> slot) & 1u) << e;
      } else {
        double prod = 1.0;
        bool ok = true;
#pragma unroll
        for (int f = 0; f < MAX_FACTORS; ++f) {
          if (f >= CT::ve_nf(P, e)) break;
          const int slot = CT::fac_slot(P, e, f);
          const double x = bits_to_f64(pick<NC>(r.in, slot), slot_class(CT::slot_dtype(P, slot)));
          ok = ok && ((r.in_valid >> slot) & 1u);
          // one rounding per operation, no FMA contraction (matches the CPU engines bit for bit)
          const double fb = CT::fac_b(P, e, f), fa = CT::fac_a(P, e, f);
          const double t = fb == 1.0 ? x : __dmul_rn(fb, x);
          const double u = fa == 0.0 ? t : __dadd_rn(fa, t);
          prod = f == 0 ? u : __dmul_rn(prod, u);
        }
        v[e] = (uint64_t)__double_as_longlong(prod);
        v_valid |= (ok ? 1u : 0u) << e;
      }
    }
  }
}

// tumbling window index of t and membership (each row belongs to at most one window on this path)
// floor(a / every) with the host-prepared multiplier when there is one (every > 0)
__device__ __forceinline__ int64_t floor_div_every(const Dyn& d, int64_t a) {
  if (d.div_magic == 0 && !(d.div_more & 0x80)) return floor_div(a, d.every);
  return a >= 0 ? (int64_t)div_apply((uint64_t)a, d.div_magic, d.div_more) : ~(int64_t)div_apply(~(uint64_t)a, d.div_magic, d.div_more);
}
__device__ __forceinline__ bool window_of(const Dyn& d, int closed, int64_t t, int64_t& kk) {
  const int64_t rel = t - d.origin;
  if (closed == 1) {  // right: (s, s+period]
    kk = floor_div_every(d, rel - 1);
    return rel - kk * d.every <= d.period;
  }
  kk = floor_div_every(d, rel);
  const int64_t off = rel - kk * d.every;
  if (closed == 0) return off < d.period;               // left  [s, s+period)
  if (closed == 3) return off > 0 && off < d.period;    // none  (s, s+period)
  return off <= d.period;                               // both  [s, s+period], period < every
}

// the same with `every` / `period` as compile-time constants of the query-shape specialised build (CT::kDynEvery > 0):
// division by a constant instead of the run-time (magic, shift, flags) interpretation — ~12 instead of ~45 instructions
template <class CT>
__device__ __forceinline__ bool window_of_ct(const ScanPlan& P, int64_t t, int64_t& kk) {
  if (CT::kDynEvery <= 0) return window_of(P.dyn, CT::dyn_closed(P), t, kk);
  constexpr long long E = CT::kDynEvery > 0 ? CT::kDynEvery : 1, PERIOD = CT::kDynPeriod;
  const int closed = CT::dyn_closed(P);
  const int64_t rel = t - P.dyn.origin;
  const int64_t a = closed == 1 ? rel - 1 : rel;
  const int64_t q = a / E, r = a - q * E;
  kk = r < 0 ? q - 1 : q;
  const int64_t off = rel - kk * E;
  if (closed == 1) return off <= PERIOD;              // right: (s, s+period]
  if (closed == 0) return off < PERIOD;               // left  [s, s+period)
  if (closed == 3) return off > 0 && off < PERIOD;    // none  (s, s+period)
  return off <= PERIOD;                               // both  [s, s+period], period < every
}

// key words of the row; returns false when the row belongs to no window (dynamic)
template <class CT, int NC, int KW>
__device__ __forceinline__ bool row_keys(const ScanPlan& P, const Row<NC>& r, const uint4 (&raw)[NC], const uint32_t (&vbits)[NC], int j,
                                         bool alive, uint64_t (&k)[KW], bool& sentinel_free) {
#pragma unroll
  for (int w = 0; w < KW; ++w) k[w] = 0;
  uint64_t nullmask = 0;
  sentinel_free = true;
  bool member = true;
  int w = 0;
#pragma unroll
  for (int q = 0; q < MAX_KEYS; ++q) {
    if (q >= CT::n_keys(P)) break;
    const int kslot = CT::key_slot(P, q), kdt = CT::key_dtype(P, q);
    uint64_t w0, w1 = 0;
    bool ok;
    if (kdt == DT_VIEW) {
      // row j of the pair lives in slot kslot + j (even / odd view)
      const uint4 v = pick128<NC>(raw, kslot + j);
      w0 = (uint64_t)v.y << 32 | v.x;
      w1 = (uint64_t)v.w << 32 | v.z;
      ok = (pick32<NC>(vbits, kslot) >> j) & 1u;
      // (values longer than 12 bytes were canonicalised when the frame was created, pw_views.cu: equal strings carry
      // equal view bytes, so the two words are the key for every length)
    } else {
      w0 = pick<NC>(r.in, kslot);
      ok = (r.in_valid >> kslot) & 1u;
      if (kdt == DT_F64 || kdt == DT_F32) {
        const double d = __longlong_as_double((long long)w0);
        if (d == 0.0) w0 = 0;                     // -0.0 == 0.0
        if (d != d) w0 = 0x7FF8000000000000ull;   // one NaN
      }
    }
    if (!ok) { w0 = 0; w1 = 0; nullmask |= 1ull << q; }
    put<KW>(k, w, w0);
    if (CT::key_words(P, q) == 2) put<KW>(k, w + 1, w1);
    w += CT::key_words(P, q);
  }
  if (CT::dyn_enabled(P)) {
    int64_t kk;
    member = window_of_ct<CT>(P, (int64_t)pick<NC>(r.in, CT::dyn_slot(P)), kk);
    put<KW>(k, w, (uint64_t)kk);
    w += 1;
  }
  if (CT::has_null_word(P)) put<KW>(k, w, nullmask);
  else if (KW == 1) {
    if (CT::n_keys(P) == 1 && !CT::dyn_enabled(P) && nullmask) k[0] = KEY_NULL;
    else if (k[0] >= KEY_NULL) sentinel_free = false;
  }
  return member;
}

// FRONT END for the row at (half HF, pair element j) of this lane.  The half is a template parameter (two
// copies of the front end) so that the raw arrays see compile-time indices only; j stays a run-time value.
template <class CT, int NC, int KW, int NV, int HF>
__device__ __forceinline__ void row_front(const ScanPlan& P, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC],
                                          int j, int64_t base, int lane, int rem, RowOut<KW, NV>& o) {
  // rem = rows of this warp step that exist (<= ROWS_PER_STEP): bounds checks stay in 32 bits
  o.row = base + HF * 64 + 2 * lane + j;
  Row<NC> r;
  row_decode<CT, NC>(P, raw[HF], vbits[HF], j, r);
  if (CT::rowid_slot(P) >= 0) o.row = (int64_t)(pick<NC>(r.in, CT::rowid_slot(P)) >> 8);  // row of the original frame
  bool alive = (HF * 64 + 2 * lane + j) < rem && row_predicate<CT, NC>(P, r);
  alive = row_keys<CT, NC, KW>(P, r, raw[HF], vbits[HF], j, alive, o.k, o.sentinel_free) && alive;
  o.alive = alive;
  row_vexprs<CT, NC, NV>(P, r, o.v, o.v_valid);
  o.tval = CT::dyn_enabled(P) ? pick<NC>(r.in, CT::dyn_slot(P)) : 0ull;
}

// ---- accumulator sinks ------------------------------------------------------------------------------------
struct ColdSink {  // HBM table, atomics
  const Table& T;
  uint64_t slot;
  template <int OP>
  __device__ __forceinline__ void add(const ScanPlan&, int a, uint64_t x) const { acc_apply_global(&tacc(T, a, slot), OP, x); }
};
// Shared-memory hot table, B rows of one lane at a time: private cells under a claim, shared min/max words.
// Every accumulator word is updated for all B rows together — B independent loads, then B stores — so that the
// shared-memory round trips of a lane's rows overlap (the caller guarantees that enabled rows of one lane hit
// DISTINCT cells).  `cw` = the cell's claim/counter word as read under the claim.
// PART selects which words a pass touches: the CTA-shared min/max words need no claim, so they are updated once,
// outside the claim loop (whose every extra round would otherwise repeat them).
enum HotPart : int { PART_ALL = 0, PART_PRIVATE = 1, PART_SHARED = 2 };
template <class CT, int KW, int B, int PART>
struct HotSinkB {
  const HotTable<CT, KW>& hot;
  const int (&id)[B];
  const int (&cell)[B];
  const uint32_t (&cw)[B];
  template <int OP>
  __device__ __forceinline__ void add(const ScanPlan& P, int a, const uint64_t (&x)[B], const bool (&en)[B]) const {
    const int kind = CT::h_kind(P, a);
    if (PART == PART_PRIVATE && kind == HOT_SHARED_MM) return;
    if (PART == PART_SHARED && kind != HOT_SHARED_MM) return;
    if (kind == HOT_PRIV64) {
      uint64_t* base = (uint64_t*)(hot.wbase + CT::h_off(P, a));
      uint64_t old[B];
#pragma unroll
      for (int i = 0; i < B; ++i) old[i] = en[i] ? base[cell[i]] : 0ull;
#pragma unroll
      for (int i = 0; i < B; ++i) if (en[i]) base[cell[i]] = acc_combine(OP, old[i], x[i]);
    } else if (kind == HOT_PRIV32) {
      uint32_t* base = (uint32_t*)(hot.wbase + CT::h_off(P, a));
      if (CT::h_rep(P) < 32 && a == CT::h_claim_acc(P)) {
        // the counter that hosts the claim byte was read by the claim check: store only (count < 2^24 between
        // flushes, so the add never carries into the claim byte)
#pragma unroll
        for (int i = 0; i < B; ++i) if (en[i]) base[cell[i]] = cw[i] + (uint32_t)x[i];
      } else {
        uint32_t old[B];
#pragma unroll
        for (int i = 0; i < B; ++i) old[i] = en[i] ? base[cell[i]] : 0u;
#pragma unroll
        for (int i = 0; i < B; ++i) if (en[i]) base[cell[i]] = old[i] + (uint32_t)x[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < B; ++i)
        if (en[i]) minmax_apply_shared(&hot.mm[(size_t)id[i] * CT::h_mm_stride(P) + CT::h_off(P, a)], OP, x[i]);
    }
  }
  // min word at accumulator a, max word at a + 1, same operand: when both are CTA-shared and form an aligned
  // 16-byte pair, one LDS.128 reads both extrema
  // `x` = 64-bit image in the op's order (floats: f64_to_ordered), `ok` = row enabled and value non-null.  With
  // is_f64 the NaN rows (skipped by min/max) are still in `ok`: they are filtered here, from the image.
  static __device__ __forceinline__ bool image_is_nan(uint64_t x) {
    return (long long)x > 0x7FF0000000000000ll || (long long)x < (long long)0x800FFFFFFFFFFFFFull;
  }
  template <int OPMIN, int OPMAX>
  __device__ __forceinline__ void minmax(const ScanPlan& P, int a, const uint64_t (&x)[B], const bool (&ok)[B], bool is_f64) const {
    if (CT::h_kind(P, a) == HOT_SHARED_MM && CT::h_kind(P, a + 1) == HOT_SHARED_MM && (CT::h_off(P, a) & 1) == 0 &&
        CT::h_off(P, a + 1) == CT::h_off(P, a) + 1 && (CT::h_mm_stride(P) & 1) == 0) {
      minmax_pair<OPMIN>(P, a, x, ok, is_f64, part_tag<PART != PART_PRIVATE>{});
    } else {
      bool en[B];
#pragma unroll
      for (int i = 0; i < B; ++i) en[i] = ok[i] && !(is_f64 && image_is_nan(x[i]));
      add<OPMIN>(P, a, x, en);
      add<OPMAX>(P, a + 1, x, en);
    }
  }
  template <bool ON> struct part_tag {};
  template <int OPMIN>
  __device__ __forceinline__ void minmax_pair(const ScanPlan&, int, const uint64_t (&)[B], const bool (&)[B], bool, part_tag<false>) const {}
  template <int OPMIN>
  __device__ __forceinline__ void minmax_pair(const ScanPlan& P, int a, const uint64_t (&x)[B], const bool (&ok)[B], bool is_f64, part_tag<true>) const {
    bool en[B];
    if (a == CT::h_guard_acc(P) && B > 1) {
      int2 sh[B];
#pragma unroll
      for (int i = 0; i < B; ++i) {
        asm volatile("ld.volatile.shared.v2.s32 {%0,%1}, [%2];" : "=r"(sh[i].x), "=r"(sh[i].y)
                     : "r"((uint32_t)__cvta_generic_to_shared(&hot.shadow[ok[i] ? id[i] : 0])));
      }
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < B; ++i) {
        const int32_t xh = (int32_t)((uint32_t)(x[i] >> 32) ^ (OPMIN == OP_MIN_I64 ? 0u : 0x80000000u));  // signed domain
        m |= ((ok[i] && (xh <= sh[i].x || xh >= sh[i].y)) ? 1u : 0u) << i;  // NaN images lie outside every interval
      }
      // few rows get here once a group has seen some rows: one row per lane per round through a single-row body
      while (m) {
        const int j = __ffs((int)m) - 1;
        m &= m - 1u;
        uint64_t xv = x[0];
        int idv = id[0];
#pragma unroll
        for (int i = 1; i < B; ++i) if (j == i) { xv = x[i]; idv = id[i]; }
        if (is_f64 && image_is_nan(xv)) continue;
        uint64_t* q = &hot.mm[(size_t)idv * CT::h_mm_stride(P) + CT::h_off(P, a)];
        const uint4 cur = lds128_volatile((const uint32_t*)q);
        const uint64_t lo = (uint64_t)cur.y << 32 | cur.x, hi = (uint64_t)cur.w << 32 | cur.z;
        const int32_t xh = (int32_t)((uint32_t)(xv >> 32) ^ (OPMIN == OP_MIN_I64 ? 0u : 0x80000000u));
        if (OPMIN == OP_MIN_I64) {
          if ((long long)xv < (long long)lo) { atomicMin((long long*)q, (long long)xv); atomicMin(&hot.shadow[idv].x, xh); }
          if ((long long)xv > (long long)hi) { atomicMax((long long*)(q + 1), (long long)xv); atomicMax(&hot.shadow[idv].y, xh); }
        } else {
          if (xv < lo) { atomicMin((unsigned long long*)q, (unsigned long long)xv); atomicMin(&hot.shadow[idv].x, xh); }
          if (xv > hi) { atomicMax((unsigned long long*)(q + 1), (unsigned long long)xv); atomicMax(&hot.shadow[idv].y, xh); }
        }
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < B; ++i) en[i] = ok[i] && !(is_f64 && image_is_nan(x[i]));
    {
      uint4 cur[B];
#pragma unroll
      for (int i = 0; i < B; ++i) {
        const uint64_t* q = &hot.mm[(size_t)(en[i] ? id[i] : 0) * CT::h_mm_stride(P) + CT::h_off(P, a)];
        cur[i] = lds128_volatile((const uint32_t*)q);
      }
      // the common case (no row improves an extremum) leaves through ONE branch
      bool lower[B], higher[B], need = false;
#pragma unroll
      for (int i = 0; i < B; ++i) {
        const uint64_t lo = (uint64_t)cur[i].y << 32 | cur[i].x, hi = (uint64_t)cur[i].w << 32 | cur[i].z;
        if (OPMIN == OP_MIN_I64) { lower[i] = en[i] && (long long)x[i] < (long long)lo; higher[i] = en[i] && (long long)x[i] > (long long)hi; }
        else { lower[i] = en[i] && x[i] < lo; higher[i] = en[i] && x[i] > hi; }
        need = need || lower[i] || higher[i];
      }
      if (need) {
#pragma unroll
        for (int i = 0; i < B; ++i) {
          uint64_t* q = &hot.mm[(size_t)id[i] * CT::h_mm_stride(P) + CT::h_off(P, a)];
          if (OPMIN == OP_MIN_I64) {
            if (lower[i]) atomicMin((long long*)q, (long long)x[i]);
            if (higher[i]) atomicMax((long long*)(q + 1), (long long)x[i]);
          } else {
            if (lower[i]) atomicMin((unsigned long long*)q, (unsigned long long)x[i]);
            if (higher[i]) atomicMax((unsigned long long*)(q + 1), (unsigned long long)x[i]);
          }
        }
      }
    }
  }
};

// B-row form of accumulate_row (same walk over the aggregate flags; `en` = rows that take part)
template <class CT, int NV, int KW, int B, class Sink>
__device__ __forceinline__ void accumulate_rows(const ScanPlan& P, const RowOut<KW, NV> (&o)[B], const uint64_t (&grow)[B],
                                                const bool (&en)[B], const Sink& s) {
  uint64_t ones[B];
#pragma unroll
  for (int i = 0; i < B; ++i) ones[i] = 1ull;
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    if (e >= CT::n_vexpr(P)) break;
    const int fl = CT::ve_flags(P, e), cls = CT::ve_cls(P, e);
    int a = CT::ve_acc(P, e);
    uint64_t bits[B];
    bool ok[B];
#pragma unroll
    for (int i = 0; i < B; ++i) { bits[i] = o[i].v[e]; ok[i] = en[i] && ((o[i].v_valid >> e) & 1u); }
    if (fl & VF_SUM_I) { s.template add<OP_ADD_I64>(P, a, bits, ok); ++a; }
    if (fl & VF_SUM_F) {
      uint64_t xs[B];
#pragma unroll
      for (int i = 0; i < B; ++i) xs[i] = (uint64_t)__double_as_longlong(bits_to_f64(bits[i], cls));
      s.template add<OP_ADD_F64>(P, a, xs, ok); ++a;
    }
    if (fl & VF_COUNT) { s.template add<OP_ADD_I64>(P, a, ones, ok); ++a; }
    if (fl & (VF_MIN | VF_MAX)) {
      uint64_t x[B];
      bool use[B];
      const bool both = (fl & VF_MIN) && (fl & VF_MAX);
#pragma unroll
      for (int i = 0; i < B; ++i) {
        x[i] = bits[i]; use[i] = ok[i];
        if (cls == CLS_F64) {
          const double d = __longlong_as_double((long long)bits[i]);
          // NaN is skipped; an all-NaN group keeps the init word -> NaN at emit time.  (min AND max: the sink filters
          // NaN from the image, after its cheap guard test)
          if (!both) use[i] = ok[i] && (d == d);
          x[i] = (uint64_t)f64_to_ordered(d);
        }
      }
      if (both) {
        if (cls == CLS_U64) s.template minmax<OP_MIN_U64, OP_MAX_U64>(P, a, x, use, false);
        else s.template minmax<OP_MIN_I64, OP_MAX_I64>(P, a, x, use, cls == CLS_F64);
        a += 2;
      } else if (fl & VF_MIN) {
        if (cls == CLS_U64) s.template add<OP_MIN_U64>(P, a, x, use); else s.template add<OP_MIN_I64>(P, a, x, use);
        ++a;
      } else {
        if (cls == CLS_U64) s.template add<OP_MAX_U64>(P, a, x, use); else s.template add<OP_MAX_I64>(P, a, x, use);
        ++a;
      }
    }
    if (fl & (VF_FIRST | VF_LAST)) {
      uint64_t x[B];
#pragma unroll
      for (int i = 0; i < B; ++i) x[i] = (grow[i] << 1) | (ok[i] ? 1ull : 0ull);
      if (fl & VF_FIRST) { s.template add<OP_MIN_U64>(P, a, x, en); ++a; }
      if (fl & VF_LAST) { s.template add<OP_MAX_U64>(P, a, x, en); ++a; }
    }
    if (fl & (VF_SUMD | VF_SUMD2)) {
      // var / std: shifted sums, one rounding per operation (see VFlag)
      uint64_t d1[B], d2[B];
#pragma unroll
      for (int i = 0; i < B; ++i) {
        const double d = __dsub_rn(bits_to_f64(bits[i], cls), P.var_shift[e]);
        d1[i] = (uint64_t)__double_as_longlong(d);
        d2[i] = (uint64_t)__double_as_longlong(__dmul_rn(d, d));
      }
      if (fl & VF_SUMD) { s.template add<OP_ADD_F64>(P, a, d1, ok); ++a; }
      if (fl & VF_SUMD2) { s.template add<OP_ADD_F64>(P, a, d2, ok); ++a; }
    }
    if (fl & (VF_FIRST_NN | VF_LAST_NN)) {
      uint64_t x[B];
#pragma unroll
      for (int i = 0; i < B; ++i) x[i] = (grow[i] << 1) | 1ull;
      if (fl & VF_FIRST_NN) { s.template add<OP_MIN_U64>(P, a, x, ok); ++a; }
      if (fl & VF_LAST_NN) { s.template add<OP_MAX_U64>(P, a, x, ok); ++a; }
    }
    if (fl & VF_AND) { s.template add<OP_AND_U64>(P, a, bits, ok); ++a; }
    if (fl & VF_OR) { s.template add<OP_OR_U64>(P, a, bits, ok); ++a; }
    if (fl & VF_XOR) { s.template add<OP_XOR_U64>(P, a, bits, ok); ++a; }
  }
  int a = CT::acc_gbase(P);
  const int gf = CT::gflags(P);
  if (gf & GF_LEN) { s.template add<OP_ADD_I64>(P, a, ones, en); ++a; }
  if (gf & GF_ROW) { s.template add<OP_MIN_U64>(P, a, grow, en); ++a; }
  if (gf & GF_TMIN) {
    uint64_t t[B];
#pragma unroll
    for (int i = 0; i < B; ++i) t[i] = o[i].tval;
    s.template add<OP_MIN_I64>(P, a, t, en); ++a;
  }
}

// a sink may keep some f64 min / max accumulators as plain doubles (pw_bucket.cuh); every other sink sees ordered images
template <class S>
struct SinkNative {
  static __device__ __forceinline__ constexpr bool on(int) { return false; }
  static __device__ __forceinline__ void fmin(const S&, int, uint64_t) {}
  static __device__ __forceinline__ void fmax(const S&, int, uint64_t) {}
};

// one pass over the aggregate flags of every value expression; accumulator words are consecutive per
// expression in VFlag order (host: lower_query), then LEN, ROW, TMIN
template <class CT, int NV, int KW, class Sink>
__device__ __forceinline__ void accumulate_row(const ScanPlan& P, const RowOut<KW, NV>& o, uint64_t grow, const Sink& s) {
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    if (e >= CT::n_vexpr(P)) break;
    const int fl = CT::ve_flags(P, e), cls = CT::ve_cls(P, e);
    int a = CT::ve_acc(P, e);
    const uint64_t bits = o.v[e];
    const bool ok = (o.v_valid >> e) & 1u;
    if (fl & VF_SUM_I) { if (ok) s.template add<OP_ADD_I64>(P, a, bits); ++a; }
    if (fl & VF_SUM_F) { if (ok) s.template add<OP_ADD_F64>(P, a, (uint64_t)__double_as_longlong(bits_to_f64(bits, cls))); ++a; }
    if (fl & VF_COUNT) { if (ok) s.template add<OP_ADD_I64>(P, a, 1ull); ++a; }
    if ((fl & (VF_MIN | VF_MAX)) && SinkNative<Sink>::on(a)) {
      // plain doubles: the sink's rows are never null, NaN or -0.0 here (they took another path)
      if (fl & VF_MIN) { SinkNative<Sink>::fmin(s, a, bits); ++a; }
      if (fl & VF_MAX) { SinkNative<Sink>::fmax(s, a, bits); ++a; }
    } else if (fl & (VF_MIN | VF_MAX)) {
      uint64_t x = bits;
      bool use = ok;
      if (cls == CLS_F64) {
        const double d = __longlong_as_double((long long)bits);
        use = ok && (d == d);  // NaN is skipped; an all-NaN group keeps the init word -> NaN at emit time
        x = (uint64_t)f64_to_ordered(d);
      }
      if (fl & VF_MIN) { if (use) { if (cls == CLS_U64) s.template add<OP_MIN_U64>(P, a, x); else s.template add<OP_MIN_I64>(P, a, x); } ++a; }
      if (fl & VF_MAX) { if (use) { if (cls == CLS_U64) s.template add<OP_MAX_U64>(P, a, x); else s.template add<OP_MAX_I64>(P, a, x); } ++a; }
    }
    if (fl & VF_FIRST) { s.template add<OP_MIN_U64>(P, a, (grow << 1) | (ok ? 1ull : 0ull)); ++a; }
    if (fl & VF_LAST) { s.template add<OP_MAX_U64>(P, a, (grow << 1) | (ok ? 1ull : 0ull)); ++a; }
    if (fl & (VF_SUMD | VF_SUMD2)) {
      const double d = __dsub_rn(bits_to_f64(bits, cls), P.var_shift[e]);
      if (fl & VF_SUMD) { if (ok) s.template add<OP_ADD_F64>(P, a, (uint64_t)__double_as_longlong(d)); ++a; }
      if (fl & VF_SUMD2) { if (ok) s.template add<OP_ADD_F64>(P, a, (uint64_t)__double_as_longlong(__dmul_rn(d, d))); ++a; }
    }
    if (fl & VF_FIRST_NN) { if (ok) s.template add<OP_MIN_U64>(P, a, (grow << 1) | 1ull); ++a; }
    if (fl & VF_LAST_NN) { if (ok) s.template add<OP_MAX_U64>(P, a, (grow << 1) | 1ull); ++a; }
    if (fl & VF_AND) { if (ok) s.template add<OP_AND_U64>(P, a, bits); ++a; }
    if (fl & VF_OR) { if (ok) s.template add<OP_OR_U64>(P, a, bits); ++a; }
    if (fl & VF_XOR) { if (ok) s.template add<OP_XOR_U64>(P, a, bits); ++a; }
  }
  int a = CT::acc_gbase(P);
  const int gf = CT::gflags(P);
  if (gf & GF_LEN) { s.template add<OP_ADD_I64>(P, a, 1ull); ++a; }
  if (gf & GF_ROW) { s.template add<OP_MIN_U64>(P, a, grow); ++a; }
  if (gf & GF_TMIN) { s.template add<OP_MIN_I64>(P, a, o.tval); ++a; }
}

template <class CT>
__device__ __forceinline__ uint64_t global_row(const ScanPlan& P, int64_t row) {
  return CT::unit_stride(P) ? (uint64_t)(row + P.row_offset) : (uint64_t)(P.row_begin + row * P.row_stride + P.row_offset);
}

// BACK END, phase 1: hash + hot-table probe of B rows per lane.  The home-bucket lookups are branch-free and
// independent, so their shared-memory round trips overlap; rows that miss (first sight of a group, keys that
// live next to an overflowing bucket) take the slow path under one warp-uniform branch.  Called convergently.
template <class CT, int KW, int NV, bool HOT, int B>
__device__ __forceinline__ void rows_probe(const ScanPlan& P, HotTable<CT, KW>& hot, const RowOut<KW, NV> (&o)[B], uint64_t (&h)[B], int (&id)[B]) {
  bool miss = false;
  if (HOT && CT::h_dense(P)) {
    // dense ids: the key IS the id (minus the range base): no index, no key compare, no insertion.  Null keys and
    // raw keys that alias a sentinel (>= KEY_NULL) and keys outside the range go to the HBM table.
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const uint64_t d = o[i].k[0] - (uint64_t)P.dense_min;
      h[i] = 0;  // the cold path hashes on demand
      const bool plain = !CT::h_dense_sentinels(P) || o[i].k[0] < KEY_NULL;  // -1 / -2 inside the range: sentinels go cold
      id[i] = (o[i].alive && plain && d < (uint64_t)CT::h_gcap(P)) ? (int)d : -1;
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < B; ++i) { h[i] = hash_words<KW>(o[i].k); id[i] = -1; }
  if (HOT) {
    uint4 tags[B];
    uint32_t cand[B];
    bool eq[B];
#pragma unroll
    for (int i = 0; i < B; ++i) tags[i] = hot.lookup_tags(P, h[i]);
#pragma unroll
    for (int i = 0; i < B; ++i) cand[i] = hot.lookup_match(tags[i], h[i]);
#pragma unroll
    for (int i = 0; i < B; ++i) eq[i] = hot.key_equals(P, (int)(cand[i] & 0xFFFFu), o[i].k);  // no match: id 0, ignored
#pragma unroll
    for (int i = 0; i < B; ++i) {
      // rows whose raw key aliases a sentinel bypass the hot table so that a hot KEY_NULL is always a true null
      const bool want = o[i].alive && o[i].sentinel_free;
      const bool hit = eq[i] && cand[i] != 0u;
      id[i] = (want && hit) ? (int)(cand[i] & 0xFFFFu) : -1;
      miss = miss || (want && !hit);
    }
  }
  if (HOT && __any_sync(0xffffffffu, miss)) {
#pragma unroll
    for (int i = 0; i < B; ++i)
      if (o[i].alive && o[i].sentinel_free && id[i] < 0) id[i] = hot.upsert_slow(P, o[i].k, h[i]);
  }
}

// BACK END, phase 2: aggregate B rows per lane.  Called convergently by all 32 lanes (dead rows keep
// `alive == false`) because the claim loop uses warp-wide votes.
template <class CT, int KW, int NV, bool HOT, int B>
__device__ __forceinline__ void rows_accumulate(const ScanPlan& P, HotTable<CT, KW>& hot, const RowOut<KW, NV> (&o)[B], const uint64_t (&h)[B],
                                                const int (&id)[B], int lane, unsigned long long& spilled) {
  uint64_t grow[B];
#pragma unroll
  for (int i = 0; i < B; ++i) grow[i] = global_row<CT>(P, o[i].row);
  if (HOT) {
    const int R = CT::h_rep(P);
    int cell[B];
    uint32_t cw[B];
    bool pend[B];
#pragma unroll
    for (int i = 0; i < B; ++i) { pend[i] = id[i] >= 0; cell[i] = pend[i] ? id[i] * R + (lane & (R - 1)) : 0; cw[i] = 0u; }
    const HotSinkB<CT, KW, B, PART_ALL> sink{hot, id, cell, cw};
    const HotSinkB<CT, KW, B, PART_PRIVATE> sink_private{hot, id, cell, cw};
    const HotSinkB<CT, KW, B, PART_SHARED> sink_shared{hot, id, cell, cw};
    if (R == 32) {
      // every lane owns its replica: no claims.  Two rows of one lane may share a cell -> one row at a time.
#pragma unroll
      for (int i = 0; i < B; ++i) {
        bool en[B];
#pragma unroll
        for (int j = 0; j < B; ++j) en[j] = (j == i) && pend[j];
        accumulate_rows<CT, NV, KW, B>(P, o, grow, en, sink);
      }
    } else {
      // Lanes of this warp that target the same private cell: every pending row writes its lane number into the
      // cell's claim byte (top byte of the cell's counter word), the warp syncs, and the row whose lane number
      // survived owns the cell for this round: plain read-modify-write, no atomics.  Losers (and a lane's second row
      // for the same cell) go again.  All B rows of a lane claim in the same round.  (A MATCH.ANY based ranking was
      // measured 15 % slower on B200: 1.93 ms vs 1.68 ms on C2.)
      uint32_t* cwords = (uint32_t*)(hot.wbase + CT::h_claim_off(P));
      bool any = false;
#pragma unroll
      for (int i = 0; i < B; ++i) any = any || pend[i];
      if (CT::h_n_mm(P) > 0) accumulate_rows<CT, NV, KW, B>(P, o, grow, pend, sink_shared);
      // first round: all B rows of the lane
      bool go = __any_sync(0xffffffffu, any);
      while (go) {
#pragma unroll
        for (int i = 0; i < B; ++i)
          if (pend[i]) ((volatile unsigned char*)(cwords + cell[i]))[3] = (unsigned char)lane;
        __syncwarp();
        bool win[B];
#pragma unroll
        for (int i = 0; i < B; ++i) {
          cw[i] = pend[i] ? ld_volatile_u32(cwords + cell[i]) : 0u;
          win[i] = pend[i] && (cw[i] >> 24) == (uint32_t)lane;
#pragma unroll
          for (int j = 0; j < i; ++j) win[i] = win[i] && !(win[j] && cell[j] == cell[i]);
        }
        accumulate_rows<CT, NV, KW, B>(P, o, grow, win, sink_private);
        any = false;
#pragma unroll
        for (int i = 0; i < B; ++i) { pend[i] = pend[i] && !win[i]; any = any || pend[i]; }
        go = (B == 1) && __any_sync(0xffffffffu, any);
      }
      if (B > 1) {
        // later rounds: few rows are left (the losers of a collision), so each lane retries ONE row per round — its
        // first pending one — through a single-row body that costs a fraction of the B-wide round
        uint32_t pm = 0;  // pending rows of this lane, one bit each
#pragma unroll
        for (int i = 0; i < B; ++i) pm |= (pend[i] ? 1u : 0u) << i;
        while (__any_sync(0xffffffffu, pm != 0u)) {
          const int j = __ffs((int)pm) - 1;  // -1: nothing pending in this lane
          RowOut<KW, NV> so[1];
          uint64_t sgrow[1];
          int sid[1], scell[1];
          uint32_t scw[1];
          so[0] = o[0]; sgrow[0] = grow[0]; sid[0] = id[0]; scell[0] = cell[0];
#pragma unroll
          for (int i = 1; i < B; ++i)
            if (j == i) { so[0] = o[i]; sgrow[0] = grow[i]; sid[0] = id[i]; scell[0] = cell[i]; }
          const bool mine = j >= 0;
          if (mine) ((volatile unsigned char*)(cwords + scell[0]))[3] = (unsigned char)lane;
          __syncwarp();
          scw[0] = mine ? ld_volatile_u32(cwords + scell[0]) : 0u;
          const bool swin[1] = {mine && (scw[0] >> 24) == (uint32_t)lane};
          const HotSinkB<CT, KW, 1, PART_PRIVATE> s1{hot, sid, scell, scw};
          accumulate_rows<CT, NV, KW, 1>(P, so, sgrow, swin, s1);
          if (swin[0]) pm &= pm - 1u;  // retire the row (lowest set bit)
        }
      }
    }
  }
  bool cold = CT::group_out(P);
#pragma unroll
  for (int i = 0; i < B; ++i) cold = cold || (o[i].alive && id[i] < 0);
  if (cold) {
#pragma unroll
    for (int i = 0; i < B; ++i) {
      if (o[i].alive && id[i] < 0) {
        // cold / spill tier: straight into the HBM table
        const uint64_t hh = (HOT && CT::h_dense(P)) ? hash_words<KW>(o[i].k) : h[i];
        const uint64_t gslot = table_upsert<KW>(P.table, o[i].k, hh, o[i].sentinel_free || KW != 1);
        if (HOT) ++spilled;
        if (gslot != ~0ull) {
          if (CT::group_out(P)) P.row_group_out[o[i].row] = P.slot_rank[gslot];  // lookup pass of group_tuples
          else {
            const ColdSink sink{P.table, gslot};
            accumulate_row<CT, NV, KW>(P, o[i], grow[i], sink);
          }
        }
      } else if (!o[i].alive && CT::group_out(P) && o[i].row < P.n_rows) P.row_group_out[o[i].row] = 0xFFFFFFFFu;
    }
  }
}

// sortedness of the dynamic index over one warp step (every adjacent row pair plus the row before the
// step) — polars-time/src/group_by/dynamic.rs:77-80 raises when it is violated
template <class CT, int NC>
__device__ __forceinline__ void check_sorted_step(const ScanPlan& P, const uint4 (&raw)[2][NC], int64_t base, int lane, int64_t n_rows) {
  int64_t carry = INT64_MIN;  // t of the row just before this half (lane 0)
  const int tslot = CT::dyn_slot(P), tdt = CT::slot_dtype(P, tslot);
  if (lane == 0 && base > 0) {
    const uint4 r1 = load_pair(P.slots[tslot].values, tdt, base - 1, n_rows, false, P.row_begin, P.row_stride);  // (rare path)
    carry = (int64_t)decode(r1, tdt, 0);
  }
  bool bad = false;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int64_t p = base + hf * 64 + 2 * lane;
    const uint4 r = pick128<NC>(raw[hf], tslot);
    const int64_t t0 = (int64_t)decode(r, tdt, 0), t1 = (int64_t)decode(r, tdt, 1);
    int64_t prev = __shfl_up_sync(0xffffffffu, t1, 1);
    if (lane == 0) prev = carry;
    if (p < n_rows && t0 < prev) bad = true;
    if (p + 1 < n_rows && t1 < t0) bad = true;
    carry = __shfl_sync(0xffffffffu, (p + 1 < n_rows) ? t1 : ((p < n_rows) ? t0 : prev), 31);
  }
  if (bad) *P.not_sorted = 1;
}

// phase 1 of a warp step: issue every load (two halves x NC slots)
template <class CT, int NC>
__device__ __forceinline__ void load_step(const ScanPlan& P, int64_t base, int lane, int64_t n_rows, uint4 (&raw)[2][NC], uint32_t (&vbits)[2][NC]) {
  if (CT::vec_ok(P) && CT::unit_stride(P) && base + ROWS_PER_STEP <= n_rows) {
    // the whole step exists (every step but the last): no per-lane bounds logic, straight vector loads
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int64_t p = base + hf * 64 + 2 * lane;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c < CT::n_slots(P)) {
          raw[hf][c] = CT::slot_dtype(P, c) == DT_BOOL ? load_bool_pair(P.slots[c], p, n_rows, 0, 1)
                                                        : load_pair(P.slots[c].values, CT::slot_dtype(P, c), p, n_rows, true, 0, 1);
          vbits[hf][c] = CT::slot_nullable(P, c) ? load_valid_pair(P.slots[c], p, n_rows, 0, 1) : 3u;
        } else {
          raw[hf][c] = make_uint4(0u, 0u, 0u, 0u);
          vbits[hf][c] = 0u;
        }
      }
    }
    return;
  }
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int64_t p = base + hf * 64 + 2 * lane;
    const bool full = CT::vec_ok(P) && (p + 1 < n_rows);  // host clears vec_ok unless stride == 1 and row_begin is even
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < CT::n_slots(P) && p < n_rows) {
        if (CT::unit_stride(P)) {
          raw[hf][c] = CT::slot_dtype(P, c) == DT_BOOL ? load_bool_pair(P.slots[c], p, n_rows, 0, 1)
                                                        : load_pair(P.slots[c].values, CT::slot_dtype(P, c), p, n_rows, full, 0, 1);
          vbits[hf][c] = CT::slot_nullable(P, c) ? load_valid_pair(P.slots[c], p, n_rows, 0, 1) : 3u;
        } else {
          raw[hf][c] = CT::slot_dtype(P, c) == DT_BOOL ? load_bool_pair(P.slots[c], p, n_rows, P.row_begin, P.row_stride)
                                                        : load_pair(P.slots[c].values, CT::slot_dtype(P, c), p, n_rows, full, P.row_begin, P.row_stride);
          vbits[hf][c] = CT::slot_nullable(P, c) ? load_valid_pair(P.slots[c], p, n_rows, P.row_begin, P.row_stride) : 3u;
        }
      } else {
        raw[hf][c] = make_uint4(0u, 0u, 0u, 0u);
        vbits[hf][c] = 0u;
      }
    }
  }
}

// one row of the wide class: front end, probe, aggregate
template <class CT, int NC, int KW, int NV, bool HOT, int HF>
__device__ __forceinline__ void wide_row(const ScanPlan& P, HotTable<CT, KW>& hot, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC], int j,
                                         int64_t base, int lane, int rem, unsigned long long& spilled) {
  RowOut<KW, NV> o[1];
  uint64_t h[1];
  int id[1];
  row_front<CT, NC, KW, NV, HF>(P, raw, vbits, j, base, lane, rem, o[0]);
  rows_probe<CT, KW, NV, HOT, 1>(P, hot, o, h, id);
  rows_accumulate<CT, KW, NV, HOT, 1>(P, hot, o, h, id, lane, spilled);
}

// ---------------------------------------------------------------------------------------------------
// the scan kernel body: CTAs take contiguous row ranges (time-sorted inputs keep few live groups per CTA)
// ---------------------------------------------------------------------------------------------------
template <class CT, int NC, int KW, bool HOT>
__device__ __forceinline__ void scan_body(const ScanPlan& P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NV = NVof<NC>::value;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  HotTable<CT, KW> hot;
  if (HOT) {
    hot.bind(smem_raw, P, warp);
    hot.clear(P);
    __syncthreads();
  }
  const int64_t n_rows = P.n_rows;
  const int64_t n_steps = (n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  const int64_t n_tiles = (n_steps + warps - 1) / warps;
  const int64_t tile_lo = n_tiles * blockIdx.x / gridDim.x;
  const int64_t tile_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;
  unsigned long long spilled = 0;
  bool stable_set = false, flushed_once = false;  // eviction heuristics (CTA-uniform)
  int tiles_since_flush = 0;

  // eviction check after a tile (CTA-wide: every warp calls it for the same tiles).  Every 4 tiles (8192 rows at 16
  // warps): two CTA barriers.  Partitioned input walks through disjoint group sets (one per partition, a tile or two
  // long): check every tile and evict at half full, so that the next partitions always find room
  auto evict_check = [&](int64_t tile) {
    if (!(HOT && (CT::rowid_slot(P) >= 0 || ((tile - tile_lo) & 3) == 3))) return;
    __syncthreads();
    // Evict everything (FixedIndexTable evicts per slot; a wholesale flush keeps the per-row path free of
    // eviction logic) when a row found the table full, or — for group sets that drift (time-sorted input) —
    // before that happens: at 7/8 full, unless the table refilled right after the previous eviction (a stable
    // group set that simply needs most of the table: then only a full table evicts).
    const uint32_t cnt = *(volatile uint32_t*)hot.count;
    const bool full = *(volatile uint32_t*)(hot.count + 1) != 0u;
    const uint32_t G = (uint32_t)CT::h_gcap(P);
    const bool nearly = CT::rowid_slot(P) >= 0 ? cnt >= (G >> 1) : (cnt >= G - (G >> 3) && !stable_set);
    __syncthreads();
    ++tiles_since_flush;
    const bool wrap = tiles_since_flush >= 30000;  // private counters share their word with the claim byte: 24 bits
    if ((full || nearly || wrap) && tile + 1 < tile_hi) {
      if (!full && !wrap && flushed_once && tiles_since_flush <= 2 && CT::rowid_slot(P) < 0) stable_set = true;  // refilled at once: same groups again
      else {
        hot.flush(P);
        __syncthreads();
        hot.clear(P);
        __syncthreads();
        flushed_once = true;
        tiles_since_flush = 0;
      }
    }
  };

  if constexpr (NC <= 4) {
    // Narrow class: evaluate, PROBE and AGGREGATE the lane's four rows together (independent chains -> instruction-
    // level parallelism; one claim round for all four), with the loads of the NEXT tile in flight while the current
    // one is processed (software pipeline: 12-16 warps per SM cannot hide an HBM round trip per tile otherwise).  The
    // tile loop is unrolled by two so that the two register buffers alternate in place (the first version rotated
    // them with ~20 register moves per step), and the row index advances by a constant stride.
    auto rows4 = [&](const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC], int64_t base, int rem) {
      if (CT::check_sorted(P)) check_sorted_step<CT, NC>(P, raw, base, lane, n_rows);
      RowOut<KW, NV> o[4];
      uint64_t h[4];
      int id[4];
      row_front<CT, NC, KW, NV, 0>(P, raw, vbits, 0, base, lane, rem, o[0]);
      row_front<CT, NC, KW, NV, 0>(P, raw, vbits, 1, base, lane, rem, o[1]);
      row_front<CT, NC, KW, NV, 1>(P, raw, vbits, 0, base, lane, rem, o[2]);
      row_front<CT, NC, KW, NV, 1>(P, raw, vbits, 1, base, lane, rem, o[3]);
      rows_probe<CT, KW, NV, HOT, 4>(P, hot, o, h, id);
      rows_accumulate<CT, KW, NV, HOT, 4>(P, hot, o, h, id, lane, spilled);
    };
    const int64_t stride = (int64_t)warps * ROWS_PER_STEP;
    int64_t base = (tile_lo * warps + warp) * ROWS_PER_STEP;
    int64_t partial_base = -1;  // the input's last, incomplete step (one warp of one CTA): handled after the loop
    // a step that lies completely inside the input (all but the last) runs with a compile-time row count: no
    // per-row bounds checks
    auto step = [&](const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC], int64_t b) {
      const int64_t left = n_rows - b;
      if (left >= ROWS_PER_STEP) rows4(raw, vbits, b, ROWS_PER_STEP);
      else if (left > 0) partial_base = b;
    };
    uint4 rawA[2][NC], rawB[2][NC];
    uint32_t vbA[2][NC], vbB[2][NC];
    if (tile_lo < tile_hi) load_step<CT, NC>(P, base, lane, n_rows, rawA, vbA);
    for (int64_t tile = tile_lo; tile < tile_hi; tile += 2) {
      load_step<CT, NC>(P, base + stride, lane, tile + 1 < tile_hi ? n_rows : 0, rawB, vbB);
      step(rawA, vbA, base);
      evict_check(tile);
      if (tile + 1 < tile_hi) {
        load_step<CT, NC>(P, base + 2 * stride, lane, tile + 2 < tile_hi ? n_rows : 0, rawA, vbA);
        step(rawB, vbB, base + stride);
        evict_check(tile + 1);
      }
      base += 2 * stride;
    }
    if (partial_base >= 0) {
      load_step<CT, NC>(P, partial_base, lane, n_rows, rawA, vbA);
      rows4(rawA, vbA, partial_base, (int)(n_rows - partial_base));
    }
  } else {
    // wide class (register-bound, no prefetch): one row at a time; the half is unrolled.  The pair element is a real
    // loop in the ahead-of-time kernels (code size); the specialised build inlines both (compile-time j: no selects
    // on the raw words — the same change took the sorted-window kernel from 75 % to 87 % of the HBM roofline)
    uint4 raw[2][NC];
    uint32_t vbits[2][NC];
    for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
      const int64_t step = tile * warps + warp;
      const int64_t base = step * ROWS_PER_STEP;
      if (step < n_steps) {
        const int64_t left = n_rows - base;
        const int rem = left >= ROWS_PER_STEP ? ROWS_PER_STEP : (int)left;
        load_step<CT, NC>(P, base, lane, n_rows, raw, vbits);
        if (CT::check_sorted(P)) check_sorted_step<CT, NC>(P, raw, base, lane, n_rows);
        if constexpr (CT::kJit) {
          wide_row<CT, NC, KW, NV, HOT, 0>(P, hot, raw, vbits, 0, base, lane, rem, spilled);
          wide_row<CT, NC, KW, NV, HOT, 0>(P, hot, raw, vbits, 1, base, lane, rem, spilled);
          wide_row<CT, NC, KW, NV, HOT, 1>(P, hot, raw, vbits, 0, base, lane, rem, spilled);
          wide_row<CT, NC, KW, NV, HOT, 1>(P, hot, raw, vbits, 1, base, lane, rem, spilled);
        } else {
#pragma unroll 1
          for (int j = 0; j < 2; ++j) wide_row<CT, NC, KW, NV, HOT, 0>(P, hot, raw, vbits, j, base, lane, rem, spilled);
#pragma unroll 1
          for (int j = 0; j < 2; ++j) wide_row<CT, NC, KW, NV, HOT, 1>(P, hot, raw, vbits, j, base, lane, rem, spilled);
        }
      }
      evict_check(tile);
    }
  }
  if (HOT) {
    __syncthreads();
    hot.flush(P);
    if (spilled) atomicAdd(P.table.spilled, spilled);
  }
}

#ifndef __CUDACC_RTC__
template <int NC, int KW, bool HOT>
__global__ void __launch_bounds__(ScanCfg<NC>::THREADS, 1) scan_kernel(const __grid_constant__ ScanPlan P) {
  scan_body<RtCtl, NC, KW, HOT>(P);
}
#endif

}  // namespace pw
