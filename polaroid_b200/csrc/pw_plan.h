// pw_plan.h — device-side query plan shared by the host engine and the kernels.
//
// The host lowers a PwQuery (include/polarway_b200.h) over a resident frame into a ScanPlan:
//   raw slots  -> the distinct columns the kernels must read (each 8-byte-or-narrower column is one slot,
//                 a 16-byte string view column takes two slots: even rows / odd rows of a row pair)
//   predicates -> conjunction of `slot <op> scalar`, null => false           
//   key words  -> canonical 64-bit words per row (+ optional null-mask word) (a3/a4/a8)
//   vexprs     -> value expressions (plain column or product of affine factors)
//   accs       -> 64-bit accumulator words per group with an associative op   (a6/a10)
// Everything is runtime data so one compiled kernel serves every query of a given shape class
// (template parameters: number of raw slots, number of key words, hot table on/off).
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#else
// NVRTC (pw_jit.cu): no host headers
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#ifndef INT64_MIN
#define INT64_MIN (-9223372036854775807ll - 1)
#define INT64_MAX 9223372036854775807ll
#endif
#endif

namespace pw {

enum DType : int32_t {
  DT_I8 = 0, DT_I16, DT_I32, DT_I64, DT_U8, DT_U16, DT_U32, DT_U64, DT_F32, DT_F64,
  DT_VIEW,     // 16-byte Utf8View/BinaryView (inline strings only)
  DT_VIEW_HI,  // second raw slot of a view column (odd row of the pair)
  DT_BOOL
};
enum ValClass : int32_t { CLS_I64 = 0, CLS_U64 = 1, CLS_F64 = 2 };

// accumulator ops: every one is a native 64-bit atomic on sm_100a (global) and an associative,
// commutative merge, which is what makes partial aggregates exchangeable.
enum AccOp : int32_t { OP_ADD_F64 = 0, OP_ADD_I64, OP_MIN_I64, OP_MAX_I64, OP_MIN_U64, OP_MAX_U64, OP_AND_U64, OP_OR_U64, OP_XOR_U64 };
// what a row contributes to an accumulator
enum AccSrc : int32_t {
  SRC_BITS = 0,   // raw 64-bit value (int sum / int min / int max); skipped when null
  SRC_F64,        // value as f64 (f64 sum, mean numerator); skipped when null
  SRC_F64_ORD,    // order-preserving int64 image of the f64 value (float min/max); null and NaN skipped
                  // (an accumulator still at its init value with count > 0 means "all NaN" -> NaN)
  SRC_VALID,      // 1 when the value is non-null (count, mean denominator, min/max validity)
  SRC_ONE,        // 1 per row (len)
  SRC_ROWIDX,     // (global_row << 1) | valid   (first = MIN, last = MAX; nulls included)
  SRC_ROW,        // global row index (group first occurrence for maintain_order / key gather)
  SRC_INDEX_T,    // the dynamic index value itself (label = datapoint -> MIN_I64)
  SRC_F64_D,      // value - shift as f64 (var/std: shifted sum)
  SRC_F64_D2,     // (value - shift)^2 (var/std: shifted sum of squares)
  SRC_ROWIDX_NN   // (global_row << 1) | 1, non-null rows only (first/last with ignore_nulls)
};

constexpr int MAX_SLOTS = 12;  // raw column slots
constexpr int MAX_PREDS = 4;
constexpr int MAX_KEYS = 4;    // key columns
constexpr int MAX_KW = 6;      // key words (4 cols x up to 2 words is capped by the templates below)
constexpr int MAX_VEXPR = 8;
constexpr int MAX_FACTORS = 4;
constexpr int MAX_ACC = 24;

struct RawSlot {
  const void* values;       // device pointer, element 0 = row 0 of the frame
  const uint8_t* validity;  // LSB-first bitmap or nullptr
  int32_t dtype;            // DType
  int32_t bit_offset;       // validity bit offset of row 0
};

struct Pred {
  int32_t slot, op, cls, pad;
  uint64_t scalar;  // bit pattern in `cls`
};

struct KeyCol {
  int32_t slot;     // first raw slot
  int32_t n_words;  // 1, or 2 for views
  int32_t dtype;
  int32_t nullable; // contributes a bit to the null-mask word
};

struct Factor { double a, b; int32_t slot, pad; };
// aggregates requested over one value expression.  Their accumulator words are consecutive, starting at
// acc_base, in this order (only the requested ones exist).
enum VFlag : int32_t {
  VF_SUM_I = 1,   // ADD_I64 of the raw integer value
  VF_SUM_F = 2,   // ADD_F64 of the value as f64 (f64 sum, mean numerator)
  VF_COUNT = 4,   // ADD_I64 of "non-null" (only when the expression is nullable; otherwise GF_LEN is shared)
  VF_MIN = 8, VF_MAX = 16,      // class-dependent: ordered-f64 / i64 / u64
  VF_FIRST = 32, VF_LAST = 64,  // MIN/MAX_U64 over (global_row << 1 | valid)
  // var / std (reduce/var_std.rs keeps Welford's (weight, mean, dp); here, so that the state stays a set of words with
  // a commutative atomic op: count + sum of d + sum of d*d with d = value - shift, shift = a sample mean of the
  // column (ScanPlan::var_shift) — the textbook shifted-data algorithm, which removes the cancellation of the plain
  // sum-of-squares form as long as the shift is within a few standard deviations of the group means)
  VF_SUMD = 128, VF_SUMD2 = 256,
  VF_FIRST_NN = 512, VF_LAST_NN = 1024,   // MIN/MAX_U64 over (global_row << 1 | 1), non-null rows only
  VF_AND = 2048, VF_OR = 4096, VF_XOR = 8192
};
enum GFlag : int32_t { GF_LEN = 1, GF_ROW = 2, GF_TMIN = 4 };  // per-group words not tied to a value expression
struct VExpr {
  int32_t n_factors;  // 0 => plain column `slot`
  int32_t slot;
  int32_t cls;        // class of the result (plain: class of the column; product: F64)
  int32_t flags;      // VFlag set
  int32_t acc_base;
  int32_t pad;
  Factor f[MAX_FACTORS];
};

struct Acc { int32_t op, src, vexpr, pad; };

struct Dyn {
  int32_t enabled, slot, closed;
  int32_t div_more;               // division by `every` as multiply + shift: shift | 0x40 (add step) | 0x80 (power of two)
  int64_t every, period, origin;  // window k = [origin + k*every, origin + k*every + period)
  uint64_t div_magic;             // 0 = not prepared (plain 64-bit division)
};
// floor(n / d) for unsigned n through the precomputed (magic, more) of d — the round-up method of Granlund & Montgomery
// ("Division by invariant integers using multiplication"), as libdivide's u64 branch-free variant lays it out.  A
// 64-bit division is ~70 instructions per row on the GPU; this is 4-6.
#ifndef __CUDACC_RTC__
__host__ inline void div_prepare(uint64_t d, uint64_t* magic, int32_t* more) {
  if ((d & (d - 1)) == 0) { int sh = 0; while ((1ull << sh) < d) ++sh; *magic = 0; *more = sh | 0x80; return; }
  int fl = 63; while (!((d >> fl) & 1)) --fl;            // floor(log2 d)
  const unsigned __int128 num = (unsigned __int128)1 << (64 + fl);
  uint64_t m = (uint64_t)(num / d);
  const uint64_t rem = (uint64_t)(num % d);
  const uint64_t e = d - rem;
  if (e < (1ull << fl)) { *more = fl; }
  else {
    m += m;
    const uint64_t twice = rem + rem;
    if (twice >= d || twice < rem) m += 1;
    *more = fl | 0x40;
  }
  *magic = m + 1;
}
#endif
__host__ __device__ inline uint64_t div_apply(uint64_t n, uint64_t magic, int32_t more) {
  if (more & 0x80) return n >> (more & 0x3F);
#ifdef __CUDA_ARCH__
  const uint64_t q = __umul64hi(magic, n);
#else
  const uint64_t q = (uint64_t)(((unsigned __int128)magic * n) >> 64);
#endif
  if (more & 0x40) return (((n - q) >> 1) + q) >> (more & 0x3F);
  return q >> (more & 0x3F);
}

// HBM open-addressing table (also the partial-aggregate state exchanged between GPUs)
// Layout: hash tables are ARRAY OF STRUCTS — one row [key words | accumulator words | pad] of 4/8/16/32 words per
// slot, so a probe plus the updates of one input row touch one or two 32-byte sectors (the first, struct-of-arrays
// layout touched one sector per word: ~7 random sectors per row on the high-cardinality path).  The window list of
// the sorted dynamic path is a dense, sequentially written table and stays struct-of-arrays.  Both are addressed
// through tkey()/tacc() with (word stride, slot stride).
struct Table {
  uint64_t* keys;    // word 0 doubles as the occupancy marker when n_kw == 1
  uint32_t* state;   // [cap + 2]: 0 empty, 1 busy, 2 ready   (n_kw > 1, and the escape slots)
  uint64_t* accs;
  uint64_t cap;      // probe range; slots cap and cap+1 are the escape slots for sentinel-valued keys
  int32_t* overflow; // set to 1 when a probe sequence exhausts the table (host retries bigger)
  unsigned long long* spilled; // rows that bypassed the hot table
  uint64_t key_sw, key_ss, acc_sw, acc_ss;  // strides (in words) per word index / per slot
};
__host__ __device__ inline uint64_t& tkey(const Table& T, int w, uint64_t slot) { return T.keys[(uint64_t)w * T.key_sw + slot * T.key_ss]; }
__host__ __device__ inline uint64_t& tacc(const Table& T, int a, uint64_t slot) { return T.accs[(uint64_t)a * T.acc_sw + slot * T.acc_ss]; }

// shared-memory hot table geometry (host-computed; see pw_scan.cuh)
enum HotKind : int32_t {
  HOT_SHARED_MM = 0,  // CTA-shared min/max word, atomic only when a row improves it
  HOT_PRIV64 = 1,     // warp-private 64-bit word (any op), plain read-modify-write under a claim
  HOT_PRIV32 = 2      // warp-private 32-bit counter (len / count)
};
struct HotGeom {
  int32_t idx_slots;   // key index slots (power of two), 0 = hot table off
  int32_t gcap;        // dense group ids
  int32_t replicas;    // R: replicas of every private cell inside a warp (power of two, <= 32)
  int32_t n_mm;        // number of CTA-shared min/max words per group
  int32_t keys_off, mm_off, count_off, warp_off, warp_bytes, claim_off, total_bytes;
  int32_t mm_stride;   // words per group in the CTA-shared min/max block (n_mm rounded up to even: 16-byte pairs)
  int32_t claim_acc;   // accumulator whose 32-bit private counter hosts the claim byte (bits 24..31), -1 = dedicated words
  int32_t dense;       // != 0: single integer key with a small value range -> dense id = key - ScanPlan::dense_min
                       // (no key index, no key compare; the per-group row counter marks the ids that exist);
                       // 2 = the range contains -1 / -2, whose bit patterns are the table's key sentinels
  int32_t guard_acc;   // min word of the CTA-shared (min, max) pair that has a 32-bit shadow, -1 = none
  int32_t shadow_off;  // byte offset of the shadow array (int2 per id)
  int32_t pad4;
  int32_t threads;     // CTA size the geometry was planned for (warp-private regions = threads / 32)
  // BUCKET tier (pw_bucket.cuh; dense ids only, specialised build only).  The fields above stay valid: when the
  // specialised build is not available the scan falls back to the per-cell table they describe.
  int32_t bucket;      // != 0: rows are bucketed by dense id and folded in the owner threads' registers
  int32_t b_threads;   // CTA size
  int32_t b_gcap;      // ids per CTA: power of two >= gcap
  int32_t b_j;         // bucket depth (rows per id per tile; deeper rows take the HBM path)
  int32_t b_nbuf;      // bucket buffers (2: one barrier per tile)
  int32_t b_stages;    // > 0: input tiles staged in shared memory by bulk async copies (TMA), this many tiles ahead; 0: register pipeline
  int32_t b_meta;      // != 0: extra plane (global row << 8 | validity bits of the value expressions)
  int32_t b_bytes;     // dynamic shared memory of the bucket kernel
  int32_t b_cps;       // CTAs per SM the geometry was planned for (launch bound)
  int32_t b_range;     // dense ids the buckets cover: [0, b_range) after subtracting ScanPlan::dense_min
  int32_t b_sent;      // the range contains the key images -1 / -2 (table sentinels): those rows take the HBM path
  int32_t b_win;       // group_by_dynamic by one dense key: buckets and registers hold one window at a time
  int32_t b_idx_mul;   // index slots per id (power of two)
  int32_t b_rowpos;    // first / last / first-occurrence words from two row positions per id and tile (no meta plane)
  int32_t b_idx;       // ids come from a CTA-local key index in shared memory (keys without a small dense range)
  int32_t b_stage_bytes;  // bytes of one staged tile (every slot's TILE rows, 16-byte aligned parts)
  int32_t acc_kind[MAX_ACC];
  int32_t acc_off[MAX_ACC];  // HOT_SHARED_MM: word inside the group's min/max block; private kinds: byte offset inside the warp region
};

struct ScanPlan {
  int64_t n_rows;          // logical rows scanned by this launch
  int64_t row_begin;       // physical row of logical row 0
  int64_t row_stride;      // physical = row_begin + logical * row_stride (1 except for the key-sample pilot)
  int64_t row_offset;      // global index of physical row 0 (multi-GPU shards)
  int32_t n_slots, n_preds, n_keys, n_kw, n_vexpr, n_acc;
  int32_t has_null_word;   // last key word = null mask
  int32_t vec_ok;          // every slot pointer 16-byte aligned -> 128-bit loads
  int32_t hot_slots;       // != 0: use the shared-memory hot table described by `hot`
  int32_t check_sorted;    // dynamic without keys: flag descending index
  int32_t gflags;          // GFlag set; their words follow the value expressions' words: LEN, ROW, TMIN
  int32_t acc_gbase;
  RawSlot slots[MAX_SLOTS];
  Pred preds[MAX_PREDS];
  KeyCol keys[MAX_KEYS];
  VExpr vexprs[MAX_VEXPR];
  Acc accs[MAX_ACC];
  Dyn dyn;
  Table table;
  int32_t* not_sorted;     // device flag
  // group_tuples second pass (GroupsIdx construction): look every row's group up and record its rank
  uint32_t* row_group_out; // [n_rows] or nullptr
  const uint32_t* slot_rank; // [cap + 2] rank of every occupied slot in the ordered group list
  int64_t dense_min;       // hot.dense: key value of dense id 0 (data, not part of the JIT shape)
  double var_shift[MAX_VEXPR];  // per value expression: the shift of its var/std words (data, not part of the JIT shape)
  // Partitioned input (high-cardinality path, pw_partition.cuh): the scan runs over a temporary frame whose rows were
  // scattered into key-hash partitions.  Every slot of that frame is a 64-bit word column; the extra slot
  // `rowid_slot_p1 - 1` holds (original row << 8) | validity bit of every slot.  0 = plain input.
  int32_t rowid_slot_p1;
  int32_t runs;            // != 0: sorted keys — the run-combining scan (pw_runs.cuh) instead of the hot table
  int32_t overlap;         // overlapping dynamic windows (pw_overlap.cuh): 2 = first pass (earliest index value per key slice
                           // into `t0`), 1 = main pass (every row joins all its windows from the slice's first window on)
  int32_t pad5;
  Table t0;                // overlap: (key words, window word = 0) -> MIN_I64 of the index value
  HotGeom hot;
};

// radix partitioning pass (pw_partition.cuh): mode 1 counts rows per partition, mode 2 scatters them
struct PartParams {
  uint32_t* hist;       // [n_parts] mode 1: rows per partition
  uint32_t* cursor;     // [n_parts] mode 2: next free row of every partition (starts at the partition's offset)
  uint64_t* out;        // mode 2: (n_slots + 1) word columns, `out_stride` words apart
  uint64_t out_stride;
  uint32_t n_parts;
  int32_t mode;
};

// two-level radix partitioning + per-partition aggregation in shared memory (pw_radix.cuh).  A record is 32 bytes:
// one 64-bit word per raw slot, then (row << 8 | validity bits).
constexpr int RADIX_TILE = 2048;        // records a scatter CTA stages per tile (two CTAs per SM: one loads while the other copies out)
constexpr int RADIX_SC_THREADS = 512;   // scatter passes: 4 rows per thread
constexpr int RADIX_THREADS = 1024;     // aggregation pass: largest CTA
struct RadixParams {
  int32_t mode;          // 0 histogram, 1 scatter (frame -> level-1 partitions), 2 scatter (level 1 -> final), 3 aggregate
  int32_t log2_parts;    // final partitions P = 2^log2_parts; partition = top bits of the key hash
  int32_t log2_p2;       // level-2 fan-out (0: one level, mode 1 writes the final partitions)
  int32_t log2_slots;    // mode 3: slots of the shared-memory table
  int32_t probe_limit;   // mode 3: probes before a key leaves for the overflow region of the HBM table
  int32_t agg_threads;   // mode 3: CTA size (1024: one CTA per SM; 512: two per SM with half-size tables)
  uint32_t* hist;        // [P]      mode 0 out: records per final partition
  const uint32_t* offs;  // [P + 1]  first record of every final partition
  uint32_t* cursor;      // mode 1: [P >> log2_p2], mode 2: [P] next free record
  const uint32_t* tile_first;  // mode 2: [(P >> log2_p2) + 1] first tile of every level-1 partition
  const uint4* src;      // modes 2, 3
  uint4* dst;            // modes 1, 2
  uint64_t ovf_cap;      // mode 3: slots [0, ovf_cap) of ScanPlan::table = open-addressing overflow region
  uint64_t dense_cap;    //         slots [ovf_cap, ovf_cap + dense_cap) = groups appended by the CTAs
  unsigned long long* dense_count;  // [0] groups appended so far, [1] next partition to hand out
};

constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;  // n_kw == 1 occupancy sentinel
constexpr uint64_t KEY_NULL = 0xFFFFFFFFFFFFFFFEull;   // n_kw == 1 image of a null key

__host__ __device__ inline uint64_t acc_init(int32_t op) {
  switch (op) {
    case OP_MIN_I64: return 0x7FFFFFFFFFFFFFFFull;
    case OP_MAX_I64: return 0x8000000000000000ull;
    case OP_MIN_U64: case OP_AND_U64: return 0xFFFFFFFFFFFFFFFFull;
    default: return 0ull;  // ADD_*, MAX_U64, OR, XOR
  }
}

}  // namespace pw
