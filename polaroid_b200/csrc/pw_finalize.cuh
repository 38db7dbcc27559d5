// pw_finalize.cuh — table initialisation, group compaction, ordering keys and result-column emission.
// Output dtype rules follow the reference (SURVEY §8a "Output dtype rules"): sum.rs:40-47 out_dtype,
// mean.rs:29-80 finish_output, min_max.rs / first_last.rs keep the input dtype, count/len -> IdxSize u32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pw_plan.h"
#include "pw_scan.cuh"

namespace pw {

struct AccOps { int32_t n; int32_t op[MAX_ACC]; };

// device control block of one query (zeroed per run).  Its scalar fields travel back to the host in front of the
// result block, so a query whose result is small needs ONE device->host copy and ONE synchronisation.
struct Control {
  int32_t overflow;
  int32_t not_sorted;
  unsigned long long spilled;
  unsigned long long counter;          // number of groups (compact_kernel)
  unsigned long long null_counts[64];  // per result column (emit_kernel)
  // key range of the pilot samples, as maxima of order-preserving unsigned images (zero-initialised = empty):
  // kmax_u = max(key ^ 2^63), kmin_n = max(~(key ^ 2^63))
  unsigned long long kmax_u, kmin_n;
};

// `zero` (optional): control block cleared by the same launch (saves a memset per query)
static __global__ void table_init_kernel(Table T, int n_kw, AccOps ops, Control* zero = nullptr) {
  const uint64_t n = T.cap + 2;
  if (zero && blockIdx.x == 0)
    for (int i = threadIdx.x; i < (int)(sizeof(Control) / 4); i += blockDim.x) ((uint32_t*)zero)[i] = 0u;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (uint64_t)gridDim.x * blockDim.x) {
    if (n_kw == 1) tkey(T, 0, s) = KEY_EMPTY;
    T.state[s] = 0u;
    for (int a = 0; a < ops.n; ++a) tacc(T, a, s) = acc_init(ops.op[a]);
  }
}

__device__ __forceinline__ bool slot_occupied(const Table& T, int n_kw, uint64_t s) {
  if (n_kw == 1 && s < T.cap) return tkey(T, 0, s) != KEY_EMPTY;
  return T.state[s] == 2u;
}

// occupied slots -> dense list (order unspecified)
static __global__ void compact_kernel(Table T, int n_kw, uint32_t* slot_list, unsigned long long* counter) {
  const uint64_t n = T.cap + 2;
  const int lane = threadIdx.x & 31;
  for (uint64_t s0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; s0 < n; s0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t s = s0 + lane;
    const bool occ = s < n && slot_occupied(T, n_kw, s);
    const uint32_t m = __ballot_sync(0xffffffffu, occ);
    if (m == 0u) continue;
    unsigned long long basei = 0;
    if (lane == 0) basei = atomicAdd(counter, (unsigned long long)__popc(m));
    basei = __shfl_sync(0xffffffffu, basei, 0);
    if (occ) slot_list[basei + __popc(m & ((1u << lane) - 1u))] = (uint32_t)s;
  }
}

// ---- ordering -----------------------------------------------------------------------------------------
enum SortSrc : int32_t {
  SORT_ACC_U64 = 0,   // accumulator word as unsigned (first occurrence row)
  SORT_WORD_I64,      // key word as signed integer
  SORT_WORD_U64,      // key word as unsigned integer
  SORT_WORD_F64,      // key word as f64 (total order)
  SORT_VIEW_HI,       // inline view: bytes 0..7 of the string, big endian
  SORT_VIEW_LO,       // inline view: bytes 8..11 big endian, then length
  SORT_NULLBIT        // 0 for null (nulls first), 1 otherwise
};
struct SortSpec { int32_t src, word, acc, nullbit; int32_t single_key_null; int32_t pad; };

__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
  return ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
}

__device__ __forceinline__ bool key_is_null(const Table& T, int n_kw, uint64_t slot, int null_word, int nullbit, int single_key_null) {
  if (null_word >= 0) return (tkey(T, null_word, slot) >> nullbit) & 1ull;
  if (single_key_null) return slot < T.cap && tkey(T, 0, slot) == KEY_NULL;
  return false;
}

// n_dev (optional): the group count is still on the device and `n` is only an upper bound — entries past the count get
// the largest key, which a STABLE sort leaves behind every real entry
static __global__ void sort_key_kernel(Table T, int n_kw, int null_word, SortSpec sp, const uint32_t* slot_list, uint64_t n, uint64_t* out,
                                       const unsigned long long* n_dev = nullptr) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (n_dev && i >= *n_dev) { out[i] = ~0ull; return; }
  const uint64_t slot = slot_list[i];
  uint64_t v = 0;
  const bool isnull = key_is_null(T, n_kw, slot, null_word, sp.nullbit, sp.single_key_null);
  switch (sp.src) {
    case SORT_ACC_U64: v = tacc(T, sp.acc, slot); break;
    case SORT_WORD_I64: v = isnull ? 0 : (tkey(T, sp.word, slot) ^ 0x8000000000000000ull); break;
    case SORT_WORD_U64: v = isnull ? 0 : tkey(T, sp.word, slot); break;
    case SORT_WORD_F64: {
      uint64_t b = tkey(T, sp.word, slot);
      int64_t o = (int64_t)b ^ (((int64_t)b >> 63) & 0x7FFFFFFFFFFFFFFFll);
      v = isnull ? 0 : ((uint64_t)o ^ 0x8000000000000000ull); break; }
    case SORT_VIEW_HI: {
      uint64_t w0 = tkey(T, sp.word, slot), w1 = tkey(T, sp.word + 1, slot);
      uint64_t first8 = (w0 >> 32) | (w1 << 32);  // string bytes 0..7 (little endian in memory)
      v = isnull ? 0 : bswap64(first8); break; }
    case SORT_VIEW_LO: {
      uint64_t w0 = tkey(T, sp.word, slot), w1 = tkey(T, sp.word + 1, slot);
      uint64_t last4 = w1 >> 32;  // bytes 8..11
      v = isnull ? 0 : ((bswap64(last4) & 0xFFFFFFFF00000000ull) | (w0 & 0xFFFFFFFFull)); break; }
    default: v = isnull ? 0 : 1; break;
  }
  out[i] = v;
}

static __global__ void iota_kernel(uint32_t* p, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}
static __global__ void gather_u32_kernel(const uint32_t* src, const uint32_t* idx, uint32_t* dst, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// ---- emission -----------------------------------------------------------------------------------------
enum EmitKind : int32_t {
  EMIT_KEY_INT = 0, EMIT_KEY_VIEW, EMIT_SUM_INT, EMIT_SUM_F64, EMIT_MEAN, EMIT_MINMAX_INT, EMIT_MINMAX_F64,
  EMIT_COUNT, EMIT_FIRSTLAST, EMIT_DYN_LOWER, EMIT_DYN_UPPER, EMIT_ACC_I64,
  EMIT_VAR,         // acc = sum of d, acc_nn = sum of d*d, acc_cnt = count; pad = ddof, mean_out = MEAN_F64/MEAN_F32, src_cls = 1 -> std
  EMIT_NULL_COUNT,  // acc = row counter, acc_cnt = non-null counter (-1: the column has no nulls)
  EMIT_BITWISE,     // acc = and/or/xor word, acc_cnt = non-null counter: all-null group -> null (reduce/bitwise.rs)
  EMIT_ANYALL       // acc = max / min word over 0/1: never null (reduce/any_all.rs, ignore_nulls)
};
enum MeanOut : int32_t { MEAN_F64 = 0, MEAN_F32, MEAN_DATE_US, MEAN_I64 };

struct EmitDesc {
  int32_t kind;
  int32_t out_dtype;      // DType of the output buffer
  int32_t word;           // key word index
  int32_t acc, acc_cnt, acc_nn;
  int32_t null_word, nullbit, single_key_null;
  int32_t mean_out;
  int32_t src_cls;        // class of the aggregated values (mean of u64 etc.)
  int32_t pad;
  RawSlot src;            // FIRST/LAST gather source
  const uint64_t* fl_values;  // FIRST/LAST after a multi-GPU merge: value bits per table slot (else nullptr)
  int64_t row_offset;
  int64_t every, period, origin;
  void* out_values;
  uint32_t* out_validity; // bitmap words (LSB first) or nullptr
  unsigned long long* null_count;
};

// DT_BOOL results are bit-packed by the caller (emit_kernel: one ballot per warp), not stored here
__device__ __forceinline__ void store_typed(void* base, uint64_t i, int dt, uint64_t bits) {
  switch (dt) {
    case DT_BOOL: break;
    case DT_I8: case DT_U8: ((uint8_t*)base)[i] = (uint8_t)bits; break;
    case DT_I16: case DT_U16: ((uint16_t*)base)[i] = (uint16_t)bits; break;
    case DT_I32: case DT_U32: ((uint32_t*)base)[i] = (uint32_t)bits; break;
    case DT_F32: ((float*)base)[i] = (float)__longlong_as_double((long long)bits); break;
    default: ((uint64_t*)base)[i] = bits; break;
  }
}

// One launch emits up to EMIT_BATCH result columns.  Blocks are numbered column-fastest (block b: column b % n_cols, rows
// of block b / n_cols), so the blocks that read the same table rows — one per column — run next to each other and the
// rows come from L2 after the first touch (a (rows, columns) grid walked a 1e7-group table once per column from HBM:
// C3's emit 1.7 ms).
constexpr int EMIT_BATCH = 16;
struct EmitBatch { EmitDesc d[EMIT_BATCH]; };
// ctl (optional, "count on the device" mode): n is an upper bound, the real count is ctl->counter.  In that mode the
// control block is the header of the result block itself, so the null counts this kernel accumulates, the group count
// and the overflow / not-sorted flags reach the host with the same copy as the columns.
static __global__ void emit_kernel(Table T, int n_kw, const __grid_constant__ EmitBatch batch, int n_cols, const uint32_t* slot_list, uint64_t n,
                                   const Control* ctl = nullptr) {
  const EmitDesc& d = batch.d[blockIdx.x % (unsigned)n_cols];
  const uint64_t i = (uint64_t)(blockIdx.x / (unsigned)n_cols) * blockDim.x + threadIdx.x;
  if (ctl) {
    const uint64_t cnt = ctl->counter;
    n = cnt < n ? cnt : n;
  }
  bool valid = true;
  uint64_t bval = 0;  // the value when the output column is Boolean (bit-packed below)
  if (i < n) {
    const uint64_t slot = slot_list[i];
    uint64_t bits = 0;
    switch (d.kind) {
      case EMIT_KEY_INT:
        valid = !key_is_null(T, n_kw, slot, d.null_word, d.nullbit, d.single_key_null);
        bits = valid ? tkey(T, d.word, slot) : 0;
        bval = bits;
        store_typed(d.out_values, i, d.out_dtype, bits);
        break;
      case EMIT_KEY_VIEW: {
        valid = !key_is_null(T, n_kw, slot, d.null_word, d.nullbit, d.single_key_null);
        uint64_t w0 = valid ? tkey(T, d.word, slot) : 0;
        uint64_t w1 = valid ? tkey(T, d.word + 1, slot) : 0;
        ((uint64_t*)d.out_values)[2 * i] = w0;
        ((uint64_t*)d.out_values)[2 * i + 1] = w1;
        break; }
      case EMIT_SUM_INT: case EMIT_ACC_I64:
        store_typed(d.out_values, i, d.out_dtype, tacc(T, d.acc, slot));
        break;
      case EMIT_SUM_F64:
        store_typed(d.out_values, i, d.out_dtype, tacc(T, d.acc, slot));
        break;
      case EMIT_MEAN: {
        const uint64_t cnt = tacc(T, d.acc_cnt, slot);
        valid = cnt != 0;
        double m = valid ? __longlong_as_double((long long)tacc(T, d.acc, slot)) / (double)cnt : 0.0;
        switch (d.mean_out) {
          case MEAN_F64: ((double*)d.out_values)[i] = m; break;
          case MEAN_F32: ((float*)d.out_values)[i] = (float)m; break;
          case MEAN_DATE_US: {
            // (s * US_IN_DAY / c) as i64   (mean.rs:62-69)
            double s = __longlong_as_double((long long)tacc(T, d.acc, slot));
            double us = valid ? __dmul_rn(s, 86400000000.0) / (double)cnt : 0.0;
            ((int64_t*)d.out_values)[i] = (int64_t)us; break; }
          default: ((int64_t*)d.out_values)[i] = (int64_t)m; break;
        }
        break; }
      case EMIT_MINMAX_INT: case EMIT_BITWISE:
        valid = tacc(T, d.acc_cnt, slot) != 0;
        bval = valid ? tacc(T, d.acc, slot) : 0;
        store_typed(d.out_values, i, d.out_dtype, bval);
        break;
      case EMIT_ANYALL:
        bval = tacc(T, d.acc, slot) & 1ull;   // all: the MIN word's init (all ones) reads as true; any: MAX's init 0 as false
        break;
      case EMIT_NULL_COUNT:
        ((uint32_t*)d.out_values)[i] = d.acc_cnt >= 0 ? (uint32_t)(tacc(T, d.acc, slot) - tacc(T, d.acc_cnt, slot)) : 0u;
        break;
      case EMIT_VAR: {
        // var = (sum d^2 - (sum d)^2 / n) / (n - ddof) over d = x - shift; n <= ddof -> null (moment.rs:126-133)
        const double cnt = (double)tacc(T, d.acc_cnt, slot);
        valid = cnt > (double)d.pad;
        double out = 0.0;
        if (valid) {
          const double s1 = __longlong_as_double((long long)tacc(T, d.acc, slot)), s2 = __longlong_as_double((long long)tacc(T, d.acc_nn, slot));
          double var = (s2 - s1 * s1 / cnt) / (cnt - (double)d.pad);
          if (var < 0.0) var = 0.0;   // rounding can leave a tiny negative residue for constant groups
          out = d.src_cls == 1 ? sqrt(var) : var;
        }
        if (d.mean_out == MEAN_F32) ((float*)d.out_values)[i] = (float)out; else ((double*)d.out_values)[i] = out;
        break; }
      case EMIT_MINMAX_F64: {
        valid = tacc(T, d.acc_cnt, slot) != 0;
        // the ordered image of a real value never equals the init sentinel (INT64_MAX/MIN map to NaN payloads)
        const uint64_t raw = tacc(T, d.acc, slot);
        const bool any_num = raw != (uint64_t)d.every;  // d.every carries acc_init(op) for this emit kind
        uint64_t b = any_num ? ordered_to_f64_bits((int64_t)raw) : 0x7FF8000000000000ull;
        store_typed(d.out_values, i, d.out_dtype, valid ? b : 0);
        break; }
      case EMIT_COUNT:
        ((uint32_t*)d.out_values)[i] = (uint32_t)tacc(T, d.acc, slot);
        break;
      case EMIT_FIRSTLAST: {
        const uint64_t packed = tacc(T, d.acc, slot);
        valid = packed & 1ull;
        if (d.period == 1 && packed == (uint64_t)d.every) valid = false;  // ignore_nulls form: the word never left its init value
        const int64_t row = (int64_t)(packed >> 1) - d.row_offset;
        uint64_t b = 0;
        if (valid) {
          if (d.fl_values) b = d.fl_values[slot];
          else {
            uint4 r = load_row(d.src, row);  // scalar path, row only
            b = decode(r, d.src.dtype, 0);
          }
        }
        bval = b;
        store_typed(d.out_values, i, d.out_dtype, b);
        break; }
      case EMIT_DYN_LOWER: case EMIT_DYN_UPPER: {
        const int64_t k = (int64_t)tkey(T, d.word, slot);
        int64_t t = d.origin + k * d.every + (d.kind == EMIT_DYN_UPPER ? d.period : 0);
        store_typed(d.out_values, i, d.out_dtype, (uint64_t)t);
        break; }
    }
  }
  if (d.out_dtype == DT_BOOL) {   // Boolean column: LSB-first bitmap, one word per warp
    const uint32_t bm = __ballot_sync(0xffffffffu, i < n && valid && (bval & 1ull));
    if ((threadIdx.x & 31) == 0 && i < n) ((uint32_t*)d.out_values)[i >> 5] = bm;
  }
  if (d.out_validity) {
    const uint32_t m = __ballot_sync(0xffffffffu, valid && i < n);
    const uint32_t inrange = __ballot_sync(0xffffffffu, i < n);
    if ((threadIdx.x & 31) == 0 && inrange) {
      d.out_validity[i >> 5] = m;
      const int nulls = __popc(inrange & ~m);
      if (nulls) atomicAdd(d.null_count, (unsigned long long)nulls);
    }
  }
}

}  // namespace pw
