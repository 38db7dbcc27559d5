// pw_bucket.cuh — the BUCKET tier of the hot path: dense group ids, accumulators in the owner thread's registers.
//
// Why: on the per-cell hot table (pw_scan.cuh) every row pays five or six random shared-memory accesses — claim byte,
// claim check, sum load, sum store, counter store, min/max shadow — about 34 shared-memory wavefronts per 32 rows, and
// the LSU data pipe, not HBM, bounds the kernel (C2: 41 % of the measured HBM peak).  A row's accumulator update is a
// read-modify-write on a cell that any lane of any warp may touch, which is what forces ownership protocols.
//
// Here the read-modify-write never happens in shared memory.  A CTA works tile by tile (64 rows per warp):
//   SCATTER  every live row takes a RANK inside its group's bucket with one native 32-bit shared-memory atomic
//            (ATOMS.ADD on cnt[id], 2.7 pipe cycles per warp instruction measured on B200 — the same as a plain
//            random STS.32) and stores its value words at buf[plane][rank][id] (one random STS.64 per word);
//   barrier;
//   FOLD     thread g reads bucket g — buf[plane][j][g] for j < cnt[g]: consecutive threads read consecutive words, no
//            bank conflicts — and folds the rows into accumulators that live in ITS REGISTERS for the whole kernel
//            (the query-shape specialised build turns every accumulator index into a constant, as in pw_segmented.cuh).
// Per 32 rows that is ~13 wavefronts (atomic 2.7 + store 5 per word + ~5 for the fold) instead of ~34, min/max cost
// nothing extra in shared memory, and nothing is claimed or retried.  Buckets are double-buffered (the scatter of tile
// t+1 overlaps the fold of tile t: one barrier per tile) when two buffers fit next to ~60 KB of L1 — in-flight global
// loads land in L1, and a CTA that takes all 227 KB of shared memory starves its own loads (measured: 0.42 ms with
// 221 KB of buckets, 0.32 ms with 172 KB on the C2 shape).
// Groups narrower than the CTA get several owner threads each (ranks interleaved), wider ones several groups per thread.
// Rows the buckets cannot take — a rank beyond the bucket depth J (sized from the Poisson tail of rows-per-group per tile:
// ~1e-4 of the rows), null / out-of-range / sentinel keys — go to the HBM table with atomics, as in the other tiers.
// At the end every owner publishes its registers into the HBM table.
//
// Reference shape: polars-expr/src/hot_groups/fixed_index_table.rs:63-160 feeding GroupedReduction::update_groups
// (polars-expr/src/reduce/mod.rs:292-312) — there, too, the rows of a morsel are first turned into (group index, value)
// pairs and each reduction then sweeps its values by group index.
#pragma once
#include "pw_scan.cuh"

namespace pw {

// shared memory through 32-bit window addresses and predicated instructions: no generic-address arithmetic and no
// branches around the per-row atomics and stores (the kernel is bound by warp-instruction issue, not by a data pipe)
__device__ __forceinline__ uint32_t sh_rank(uint32_t addr) {   // atomicAdd(cnt, 1)
  uint32_t r;
  asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t sh_rank_if(uint32_t addr, bool p) {   // p ? atomicAdd(cnt, 1) : ~0
  uint32_t r = 0xFFFFFFFFu;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q atom.shared.add.u32 %0, [%1], 1;\n\t}" : "+r"(r) : "r"(addr), "r"((uint32_t)p) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t atomicAdd_shared_u32(uint32_t addr) { return sh_rank(addr); }
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x) {   // a value the compiler may not recompute (it lives in a register)
  uint32_t y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void sh_st64_if(uint32_t addr, uint64_t v, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.b64 [%0], %1;\n\t}" ::"r"(addr), "l"(v), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ uint64_t sh_ld64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t sh_ld32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sh_st32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// ---- input staging: bulk async copies (TMA, cp.async.bulk -> UBLKCP) completing on an mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "PW_MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra PW_MBAR_DONE;\n\t"
      "bra PW_MBAR_WAIT;\n\t"
      "PW_MBAR_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// this lane's two rows of a staged slot, in the register format of load_pair (pw_scan.cuh)
__device__ __forceinline__ uint4 staged_pair(uint32_t addr, int w) {
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  if (w == 8) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  else if (w == 4) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr) : "memory");
  else if (w == 2) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r.x) : "r"(addr) : "memory");
  else asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r.x) : "r"(addr) : "memory");
  return r;
}

// accumulators of ONE group in registers; SKIP = accumulators handled outside the per-row walk (the row counter, and
// under ROWPOS the row-index words of first / last / maintain_order);
// NATIVE = bit per accumulator: an f64 min / max kept as the plain double and compared with DSETP (the rows that need the
// total order — NaN, -0.0 — never reach the buckets), turned into the ordered image once, when the registers are published
template <int NACC, uint32_t SKIP, uint32_t NATIVE, bool SELECT_MM = false>
struct BucketRegSink {
  uint64_t (&acc)[NACC];
  template <int OP>
  __device__ __forceinline__ void add(const ScanPlan&, int a, uint64_t x) const {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      if (i == a && !((SKIP >> i) & 1u)) acc[i] = acc_combine(OP, acc[i], x);
  }
  // SELECT_MM: compare-and-select on both halves (measured faster: 0.419 vs 0.465 ms on C2); otherwise compare and
  // skip the update (fewer instructions, but the predicated moves wait on the FP64 compare)
  __device__ __forceinline__ void fmin(int a, uint64_t bits) const {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      if (i == a) {
        if (SELECT_MM) acc[i] = __longlong_as_double((long long)bits) < __longlong_as_double((long long)acc[i]) ? bits : acc[i];
        else if (__longlong_as_double((long long)bits) < __longlong_as_double((long long)acc[i])) { asm volatile("" ::: "memory"); acc[i] = bits; }
      }
  }
  __device__ __forceinline__ void fmax(int a, uint64_t bits) const {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      if (i == a) {
        if (SELECT_MM) acc[i] = __longlong_as_double((long long)bits) > __longlong_as_double((long long)acc[i]) ? bits : acc[i];
        else if (__longlong_as_double((long long)bits) > __longlong_as_double((long long)acc[i])) { asm volatile("" ::: "memory"); acc[i] = bits; }
      }
  }
};
template <int NACC, uint32_t SKIP, uint32_t NATIVE, bool SELECT_MM>
struct SinkNative<BucketRegSink<NACC, SKIP, NATIVE, SELECT_MM>> {
  using S = BucketRegSink<NACC, SKIP, NATIVE, SELECT_MM>;
  static __device__ __forceinline__ constexpr bool on(int a) { return ((NATIVE >> a) & 1u) != 0; }
  static __device__ __forceinline__ void fmin(const S& s, int a, uint64_t bits) { s.fmin(a, bits); }
  static __device__ __forceinline__ void fmax(const S& s, int a, uint64_t bits) { s.fmax(a, bits); }
};

struct BucketWholeTag { static constexpr bool value = true; };     // a tile without a ragged end: no row bounds anywhere
struct BucketRaggedTag { static constexpr bool value = false; };

// the two rows of every slot this lane owns in the tile at `base` (lane l: rows base + 2l, base + 2l + 1)
template <class CT, int NC, bool WHOLE>
__device__ __forceinline__ void bucket_load(const ScanPlan& P, int64_t base, int lane, int64_t n_rows, uint4 (&raw)[NC], uint32_t (&vbits)[NC]) {
  const int64_t p = base + 2 * lane;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    raw[c] = make_uint4(0u, 0u, 0u, 0u);
    vbits[c] = 0u;
    if (c < CT::n_slots(P) && (WHOLE || p < n_rows)) {
      const bool full = WHOLE || p + 1 < n_rows;
      raw[c] = CT::slot_dtype(P, c) == DT_BOOL ? load_bool_pair(P.slots[c], p, n_rows, 0, 1)
                                                : load_pair(P.slots[c].values, CT::slot_dtype(P, c), p, n_rows, full, 0, 1);
      vbits[c] = CT::slot_nullable(P, c) ? load_valid_pair(P.slots[c], p, n_rows, 0, 1) : 3u;
    }
  }
}

// the plan without its window word: the windowed tier classifies rows against the tile's one or two windows itself
// (two compares per window instead of a 64-bit division per row); o.k[KW - 1] is filled in only for the HBM path
template <class CT>
struct NoDynCtl : CT {
  static __device__ __forceinline__ constexpr bool dyn_enabled(const ScanPlan&) { return false; }
};

template <class CT, int NC, int KW, int NV>
__device__ __forceinline__ void bucket_row_front(const ScanPlan& P, const uint4 (&raw)[NC], const uint32_t (&vbits)[NC], int j, int64_t row,
                                                 bool in_range, RowOut<KW, NV>& o) {
  o.row = row;
  Row<NC> r;
  row_decode<CT, NC>(P, raw, vbits, j, r);
  bool alive = in_range && row_predicate<CT, NC>(P, r);
  alive = row_keys<NoDynCtl<CT>, NC, KW>(P, r, raw, vbits, j, alive, o.k, o.sentinel_free) && alive;
  o.alive = alive;
  row_vexprs<CT, NC, NV>(P, r, o.v, o.v_valid);
  o.tval = CT::dyn_enabled(P) ? pick<NC>(r.in, CT::dyn_slot(P)) : 0ull;   // index value (windowed tier)
}

// a row the buckets cannot take: straight into the HBM table (kept out of line: it is rare and register-hungry)
template <class CT, int KW, int NV>
__device__ __noinline__ void bucket_cold_row(const ScanPlan& P, RowOut<KW, NV> o, uint64_t grow) {
  const uint64_t gslot = table_upsert<KW>(P.table, o.k, hash_words<KW>(o.k), o.sentinel_free || KW != 1);
  if (gslot != ~0ull) {
    const ColdSink sink{P.table, gslot};
    accumulate_row<CT, NV, KW>(P, o, grow, sink);
  }
}

// ---- INDEXED ids: keys that do not span a small integer range (sparse integers, short strings, several key columns) ----
// A CTA-local open-addressing index in shared memory maps the key words to an id in [0, GCAP), assigned in order of
// first sight; everything after that — ranks, buckets, the owner's registers — is the dense-id tier.  Slot word:
// 0 empty, 0xFFFFFFFF being written, 0xFFFFFFFE closed (the index was full when this slot was wanted), else id + 1.
__device__ __forceinline__ uint32_t sh_ld32_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sh_st32_volatile(uint32_t addr, uint32_t v) { asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t sh_cas32(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
// -> id of the key, or ~0 when the index is full (the row takes the HBM path).
// Slot word = (20-bit hash tag << 12) | (id + 1): a probe reads one word and goes to the key words only on a tag match
// (a foreign key passes the tag test once in 2^20), the index runs at a load factor <= 0.25 so that the first probe
// almost always ends the search — the warp pays the LONGEST probe sequence among its lanes.
template <int KW>
__device__ __forceinline__ uint32_t bucket_index_hash(const uint64_t (&k)[KW]) {
  uint32_t x = 0x9E3779B1u;
#pragma unroll
  for (int w = 0; w < KW; ++w) {
    x = (x ^ (uint32_t)k[w]) * 0x85EBCA6Bu;
    x = (x ^ (uint32_t)(k[w] >> 32)) * 0xC2B2AE35u;
    x ^= x >> 15;
  }
  x *= 0x27D4EB2Fu;
  return x ^ (x >> 16);
}
// The index is BUCKETIZED: four slot words per 16-byte bucket, read with one LDS.128.  With <= 0.25 keys per slot word a
// key sits in its home bucket with probability > 0.99, so the common lookup is: hash, one 16-byte read, four tag
// compares, one key read — straight-line code, no probe loop (a loop costs the warp its SLOWEST lane).  Everything else
// — first sight of a key (insert), a full home bucket, a tag collision — goes through this out-of-line routine.
// -> id, or 0xFFFFFFFF when the index is full.
template <int KW>
struct KeyWords { uint64_t w[KW]; };   // by value: the caller's key words stay in registers
template <int KW, int NB, int GCAP>
__device__ __noinline__ uint32_t bucket_index_slow(uint32_t sidx, uint32_t skeys, uint32_t scount, KeyWords<KW> k, uint32_t hx) {
  const uint32_t tag = ((hx & 0x7FFFF000u) | 0x1000u);
  uint32_t b = hx & (uint32_t)(NB - 1);
  for (int walked = 0; walked < NB; ++walked, b = (b + 1) & (uint32_t)(NB - 1)) {
    for (int j = 0; j < 4; ++j) {
      const uint32_t a = sidx + 16u * b + 4u * (uint32_t)j;
      uint32_t e = sh_ld32_volatile(a);
      if (e == 0u) {
        e = sh_cas32(a, 0u, 0xFFFFFFFFu);
        if (e == 0u) {   // ours to fill
          const uint32_t id = sh_rank(scount);
          if (id >= (uint32_t)GCAP) { sh_st32_volatile(a, 0xFFFFFFFEu); return 0xFFFFFFFFu; }
#pragma unroll
          for (int w = 0; w < KW; ++w) sh_st64_if(skeys + ((uint32_t)w * GCAP + id) * 8u, k.w[w], true);
          __threadfence_block();
          sh_st32_volatile(a, tag | (id + 1u));
          return id;
        }
      }
      // another thread is writing this word's key — possibly a lane of this very warp: sleeping hands the scheduler to the
      // other divergent paths (a bare spin inside an out-of-line function starved the writer: measured as a hang), and
      // the wait is bounded — a row that gives up takes the HBM path, which is always correct
      for (int spins = 0; e == 0xFFFFFFFFu && spins < 4096; ++spins) { __nanosleep(64); e = sh_ld32_volatile(a); }
      if (e >= 0xFFFFFFFEu) return 0xFFFFFFFFu;            // closed: the index was full when this word was wanted (or gave up)
      if ((e & 0xFFFFF000u) == tag) {
        const uint32_t id = (e & 0xFFFu) - 1u;
        bool eq = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) eq = eq && sh_ld64(skeys + ((uint32_t)w * GCAP + id) * 8u) == k.w[w];
        if (eq) return id;
      }
    }
  }
  return 0xFFFFFFFFu;
}

template <class CT, int NC, int KW>
__device__ __forceinline__ void bucket_body(const ScanPlan& P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NV = NVof<NC>::value;
  constexpr int THREADS = CT::kBThreads, WARPS = THREADS / 32;
  constexpr int GCAP = CT::kBGcap;        // power of two >= the dense id range
  constexpr int J = CT::kBJ;              // bucket depth
  constexpr int NBUF = CT::kBNbuf;        // 2: one barrier per tile, 1: two
  constexpr int NVE = CT::kNVexpr;        // value planes
  constexpr bool META = CT::kBMeta;       // extra plane: (global row << 8) | validity bits of the value expressions
  constexpr int PLANES = NVE + (META ? 1 : 0);
  constexpr int NACC = CT::kNAcc;
  constexpr uint32_t NATIVE = CT::kBNativeAcc;   // accumulators kept as plain doubles
  constexpr uint32_t ODD = CT::kBOddVexpr;       // value expressions whose NaN / -0.0 rows take the HBM path
  constexpr int GPT = GCAP > THREADS ? GCAP / THREADS : 1;   // groups per owner thread
  constexpr int SUB = GCAP < THREADS ? THREADS / GCAP : 1;   // owner threads per group
  // one owner per group: the owner zeroes its own counter while it folds, one counter array per buffer.  Several owners
  // per group: none of them may zero it while the others still read, so the arrays rotate one step slower than the
  // buffers and the previous tile's array is cleared after the barrier.
  constexpr int NCNT = SUB == 1 ? NBUF : NBUF + 1;
  constexpr int TILE = WARPS * 64;        // rows per tile: two per lane
  constexpr int LEN_ACC = CT::kLenAcc;    // the row counter is bumped once per fold, not once per row
  constexpr int STAGES = CT::kBStages;    // > 0: input tiles arrive in shared memory by bulk async copies, this many tiles ahead
  // WINDOWED: group_by_dynamic with ONE dense integer key (OHLCV bars by symbol).  The group is (id, window); the
  // buckets and the registers hold the ids of ONE window at a time.  A tile of time-ordered rows lies in one window, or
  // two at a window change: then it is scattered and folded twice, and the registers are published in between.  Rows of
  // any other window (unsorted input, windows shorter than a tile) take the HBM path — correct, just slow — which is
  // why the host asks for windows several tiles long before it plans this tier.
  constexpr bool WIN = CT::kBWin;
  // ROWPOS: first / last / first-occurrence words need only the smallest and largest ROW of a group, not a word per
  // row: when no value can be null (the valid bit is constant) every bucketed row registers its position inside the
  // tile with a native shared-memory min / max, the meta plane disappears (a third of the bucket memory and of the
  // fold's loads for OHLCV bars) and the owner turns the two positions into row words once per tile.
  constexpr bool ROWPOS = CT::kBRowPos;
  constexpr uint32_t FIRSTW = CT::kBFirstAcc, LASTW = CT::kBLastAcc;   // accumulators: MIN / MAX over a row word
  constexpr uint32_t ROWTAG = CT::kBRowTagAcc;                         // of those: word = (row << 1) | 1 (else the plain row)
  constexpr uint32_t POS_BYTES = ROWPOS ? (uint32_t)NBUF * 2u * GCAP * 4u : 0u;
  constexpr bool IDX = CT::kBIdx;             // ids from the CTA's key index instead of key - dense_min
  constexpr int ISLOTS = CT::kBIdxMul * GCAP; // index slot words (8 per id by default: <= 0.125 keys per word)
  constexpr int INB = ISLOTS / 4;             // buckets of four words
  constexpr uint32_t IDX_BYTES = IDX ? (uint32_t)(ISLOTS * 4 + GCAP * KW * 8 + 16) : 0u;
  static_assert(!(IDX && WIN), "indexed ids and the windowed tier are not combined");
  constexpr uint32_t VAR = CT::kBVar;     // experiment switches (PW_BUCKET_VAR): 1 branch-form min / max, 2 predicated rank atomics
  // OVERFLOW LIST: a row whose rank is beyond the bucket depth is appended to a short CTA-wide list (id + value words)
  // that every owner scans after its bucket — a broadcast read per entry.  The HBM path is ~1 us per row (dependent,
  // contended global atomics; measured: 79 k such rows of 1e8 doubled the kernel time), so it must stay exceptional even
  // when the host's depth estimate (Poisson tail, 1e-4 of the rows) is off; only a full list falls through to HBM.
  constexpr int OVF = 32, NOVF = NBUF + 1;
  constexpr uint32_t OVF_BYTES = (16u + (uint32_t)NBUF * OVF * (4u + 8u * PLANES) + 127u) & ~127u;
  constexpr uint32_t PLANE_BYTES = (uint32_t)J * GCAP * 8u, BUF_BYTES = PLANE_BYTES * PLANES;
  static_assert((GCAP & (GCAP - 1)) == 0 && (THREADS % 32) == 0, "bucket geometry");

  const uint32_t sbuf = opaque_u32((uint32_t)__cvta_generic_to_shared(smem_raw));   // [NBUF][PLANES][J][GCAP] x 8 bytes
  const uint32_t scnt = sbuf + (uint32_t)NBUF * BUF_BYTES;              // [NCNT][GCAP] x 4 bytes, then 32 dummy cells
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // a row that takes no bucket still runs the rank atomic — on its lane's own dummy cell (one bank each) — so that the
  // scatter has no branches; the result is ignored
  const uint32_t dummy = scnt + (uint32_t)NCNT * (GCAP * 4u) + 4u * lane;
  // staged input: [STAGES][slot][TILE rows], then one mbarrier per stage
  uint32_t slot_off[NC], stage_bytes = 0;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    slot_off[c] = stage_bytes;
    if (c < CT::n_slots(P)) stage_bytes += (uint32_t)TILE * (uint32_t)dtype_width(CT::slot_dtype(P, c));
  }
  const uint32_t sovf = scnt + (uint32_t)NCNT * (GCAP * 4u) + 128u;     // [NOVF] counters | [NBUF][OVF] ids | [NBUF][PLANES][OVF] words
  const uint32_t sovf_id = sovf + 16u, sovf_val = sovf_id + (uint32_t)NBUF * OVF * 4u;
  const uint32_t sindex = sovf + OVF_BYTES;                 // IDX: [ISLOTS] slot words | [KW][GCAP] key words | id counter
  const uint32_t sikeys = sindex + (uint32_t)ISLOTS * 4u, sicount = sikeys + (uint32_t)GCAP * KW * 8u;
  const uint32_t spos = sindex + IDX_BYTES;                  // ROWPOS: [NBUF][first | last][GCAP] positions inside the tile
  const uint32_t sstage = spos + POS_BYTES;
  const uint32_t sbar = sstage + (uint32_t)STAGES * stage_bytes;
  for (int i = tid; i < NCNT * GCAP; i += THREADS) sh_st32(scnt + 4u * i, 0u);
  if (tid < 4) sh_st32(sovf + 4u * tid, 0u);
  if (ROWPOS)
    for (int i = tid; i < NBUF * 2 * GCAP; i += THREADS) sh_st32(spos + 4u * i, ((i / GCAP) & 1) ? 0u : 0xFFFFFFFFu);
  if (IDX) {
    for (int i = tid; i < ISLOTS; i += THREADS) sh_st32(sindex + 4u * i, 0u);
    if (tid == 0) sh_st32(sicount, 0u);
  }
  if (STAGES > 0 && tid == 0) {
    for (int st = 0; st < STAGES; ++st) mbar_init(sbar + 8u * st, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  uint64_t acc[GPT][NACC];
  bool seen[GPT];
  auto reset_acc = [&]() {
#pragma unroll
    for (int gi = 0; gi < GPT; ++gi) {
      seen[gi] = false;
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        const int op = CT::acc_op(P, a);
        acc[gi][a] = !((NATIVE >> a) & 1u) ? acc_init(op) : (op == OP_MIN_I64 ? 0x7FF0000000000000ull : 0xFFF0000000000000ull);   // +inf / -inf
      }
    }
  };
  reset_acc();
  const int my_g0 = GCAP >= THREADS ? tid : (tid & (GCAP - 1));   // first (or only) group of this thread
  const int my_sub = GCAP >= THREADS ? 0 : tid / GCAP;

  const int64_t n_rows = P.n_rows;
  const int64_t n_tiles = (n_rows + TILE - 1) / TILE, n_whole = n_rows / TILE;
  const int64_t tile_lo = n_tiles * blockIdx.x / gridDim.x;
  const int64_t tile_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int64_t whole_hi = tile_hi < n_whole ? tile_hi : n_whole;   // [tile_lo, whole_hi): tiles without a ragged end
  unsigned long long spilled = 0;
  const int64_t lane_row = (int64_t)warp * 64 + 2 * lane;
  int64_t cur_w = 0;      // WINDOWED: the window the registers belong to
  bool have_w = false;

  // the registers of this thread's groups go to the HBM table (at the end; WINDOWED: whenever the window changes)
  auto publish = [&](int64_t w) {
#pragma unroll
    for (int gi = 0; gi < GPT; ++gi) {
      if (!seen[gi]) continue;
      const int g = my_g0 + gi * THREADS;
      uint64_t k[KW];
#pragma unroll
      for (int x = 0; x < KW; ++x) {
        if (IDX) k[x] = sh_ld64(sikeys + ((uint32_t)x * GCAP + (uint32_t)g) * 8u);
        else k[x] = x == 0 ? (uint64_t)P.dense_min + (uint64_t)g : (WIN && x == 1 ? (uint64_t)w : 0ull);
      }
      const uint64_t gs = table_upsert<KW>(P.table, k, hash_words<KW>(k), true);
      if (gs == ~0ull) continue;
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        const int op = CT::acc_op(P, a);
        if ((NATIVE >> a) & 1u) acc_apply_global(&tacc(P.table, a, gs), op, (uint64_t)f64_to_ordered(__longlong_as_double((long long)acc[gi][a])));
        else if (acc[gi][a] != acc_init(op)) acc_apply_global(&tacc(P.table, a, gs), op, acc[gi][a]);
      }
    }
  };

  // one thread starts the copies of tile t into stage `st`: every slot's TILE rows, completing on the stage's mbarrier
  auto issue = [&](int64_t t, int st) {
    const uint32_t bar = sbar + 8u * st, dst = sstage + (uint32_t)st * stage_bytes;
    mbar_expect_tx(bar, stage_bytes);
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c < CT::n_slots(P)) {
        const uint32_t w = (uint32_t)dtype_width(CT::slot_dtype(P, c));
        bulk_g2s(dst + slot_off[c], (const unsigned char*)P.slots[c].values + (size_t)t * TILE * w, (uint32_t)TILE * w, bar);
      }
  };

  int b = 0, ci = 0, oi = 0;   // buffer / counter array / overflow counter of the current tile
  int st_cur = 0;      // staged input: stage of the current tile
  // WINDOWED: is the row at `rel` (index value - origin) a member of the window that starts at `lo` (= k * every)?
  // Same membership as window_of (pw_scan.cuh) for each `closed`.
  constexpr long long EVERY = CT::kDynEvery > 0 ? CT::kDynEvery : 1, PERIOD = CT::kDynPeriod > 0 ? CT::kDynPeriod : 2;
  auto in_window_at = [&](int64_t rel, int64_t lo) -> bool {
    const int64_t off = rel - lo;
    const int closed = CT::dyn_closed(P);
    if (closed == 0) return (uint64_t)off < (uint64_t)PERIOD;                 // left  [s, s+period)
    if (closed == 1) return (uint64_t)(off - 1) < (uint64_t)PERIOD;           // right (s, s+period]
    if (closed == 3) return (uint64_t)(off - 1) < (uint64_t)(PERIOD - 1);     // none  (s, s+period)
    return (uint64_t)off <= (uint64_t)PERIOD;                                 // both  [s, s+period], period < every
  };
  // w0 / w1: WINDOWED — the windows of the tile's first and last row; pass 0 handles w0, pass 1 (only when w1 != w0) w1
  auto process = [&](int64_t t, const uint4 (&raw)[NC], const uint32_t (&vb)[NC], auto whole_tag, int64_t w0, int64_t w1, int pass) {
    constexpr bool WHOLE = decltype(whole_tag)::value;
    const int64_t wp = pass == 0 ? w0 : w1;
    const uint32_t bb = sbuf + (uint32_t)b * BUF_BYTES;
    const uint32_t cc = scnt + (uint32_t)ci * (GCAP * 4u);
    // ---- scatter ----
    {
      RowOut<KW, NV> o[2];
      const int64_t row0 = t * TILE + lane_row;
      bucket_row_front<CT, NC, KW, NV>(P, raw, vb, 0, row0, WHOLE || row0 < n_rows, o[0]);
      bucket_row_front<CT, NC, KW, NV>(P, raw, vb, 1, row0 + 1, WHOLE || row0 + 1 < n_rows, o[1]);
      uint32_t id[2], rk[2];
      bool take[2], cand[2], cold[2];
      // IDX: ids of both rows.  The home-slot probes of the two rows are issued together (slot words, then key words:
      // two dependent shared-memory round trips for the pair instead of four); only a row whose home slot does not
      // hold its key walks the probe sequence.
      uint32_t idx_id[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
      bool idx_plain[2] = {false, false};
      if (IDX) {
        uint32_t hx[2], tagw[2];
        uint4 e[2], f[2];   // home bucket and its successor: a key displaced by a full home bucket sits next door
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          // (a single-word key whose VALUE is one of the table sentinels is told apart by sentinel_free: HBM path)
          idx_plain[i] = o[i].alive && (KW != 1 || o[i].sentinel_free);
          hx[i] = bucket_index_hash<KW>(o[i].k);
          const uint32_t b0 = hx[i] & (uint32_t)(INB - 1), b1 = (b0 + 1u) & (uint32_t)(INB - 1);
          const uint32_t a = sindex + 16u * b0, a1 = sindex + 16u * b1;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e[i].x), "=r"(e[i].y), "=r"(e[i].z), "=r"(e[i].w) : "r"(a) : "memory");
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(f[i].x), "=r"(f[i].y), "=r"(f[i].z), "=r"(f[i].w) : "r"(a1) : "memory");
        }
        uint64_t kw0[2][KW];
        bool any[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t tag = ((hx[i] & 0x7FFFF000u) | 0x1000u);
          const bool m0 = (e[i].x & 0xFFFFF000u) == tag, m1 = (e[i].y & 0xFFFFF000u) == tag, m2 = (e[i].z & 0xFFFFF000u) == tag,
                     m3 = (e[i].w & 0xFFFFF000u) == tag, m4 = (f[i].x & 0xFFFFF000u) == tag, m5 = (f[i].y & 0xFFFFF000u) == tag,
                     m6 = (f[i].z & 0xFFFFF000u) == tag, m7 = (f[i].w & 0xFFFFF000u) == tag;
          tagw[i] = m0 ? e[i].x : (m1 ? e[i].y : (m2 ? e[i].z : (m3 ? e[i].w : (m4 ? f[i].x : (m5 ? f[i].y : (m6 ? f[i].z : f[i].w))))));
          any[i] = m0 || m1 || m2 || m3 || m4 || m5 || m6 || m7;
          const uint32_t cid = ((tagw[i] & 0xFFFu) - 1u) & (uint32_t)(GCAP - 1);   // always a readable key slot
#pragma unroll
          for (int w = 0; w < KW; ++w) kw0[i][w] = sh_ld64(sikeys + ((uint32_t)w * GCAP + cid) * 8u);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          bool hit = idx_plain[i] && any[i];
#pragma unroll
          for (int w = 0; w < KW; ++w) hit = hit && kw0[i][w] == o[i].k[w];
          if (hit) idx_id[i] = (tagw[i] & 0xFFFu) - 1u;
          else if (idx_plain[i]) {
            KeyWords<KW> kv;
#pragma unroll
            for (int w = 0; w < KW; ++w) kv.w[w] = o[i].k[w];
            idx_id[i] = bucket_index_slow<KW, INB, GCAP>(sindex, sikeys, sicount, kv, hx[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint64_t d;
        bool plain;
        if (IDX) {
          plain = idx_plain[i];
          d = idx_id[i];
        } else {
          d = o[i].k[0] - (uint64_t)P.dense_min;
          plain = !CT::kBSent || o[i].k[0] < KEY_NULL;   // -1 / -2 inside the range: sentinels go cold
        }
        bool odd = false;   // a value the plain-double min / max cannot order
#pragma unroll
        for (int e = 0; e < NVE; ++e)
          if ((ODD >> e) & 1u) {
            const double x = __longlong_as_double((long long)o[i].v[e]);
            odd = odd || x != x || o[i].v[e] == 0x8000000000000000ull;
          }
        const bool elig = o[i].alive && plain && !odd && d < (uint64_t)(IDX ? GCAP : CT::kBRange);
        if (WIN) {
          const int64_t rel = (int64_t)o[i].tval - P.dyn.origin;
          const bool in0 = in_window_at(rel, w0 * EVERY), in1 = w1 != w0 && in_window_at(rel, w1 * EVERY);
          take[i] = elig && (pass == 0 ? in0 : in1);
          cold[i] = pass == 0 && o[i].alive && !(elig && (in0 || in1));   // another window, or none: sorted out below
        } else {
          take[i] = elig;
          cold[i] = o[i].alive && !elig;
        }
        id[i] = (uint32_t)d & (uint32_t)(GCAP - 1);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) rk[i] = (VAR & 2u) ? sh_rank_if(cc + 4u * id[i], take[i]) : sh_rank(take[i] ? cc + 4u * id[i] : dummy);
      if (ROWPOS) {
        // position inside the tile (+ 1 for the maximum: 0 = none); rows that take no bucket aim at the lane's dummy cell
        const uint32_t pb = spos + (uint32_t)b * (2u * GCAP * 4u);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t r = (uint32_t)lane_row + (uint32_t)i;
          if (FIRSTW) asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(take[i] ? pb + 4u * id[i] : dummy), "r"(r) : "memory");
          if (LASTW) asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(take[i] ? pb + (GCAP + id[i]) * 4u : dummy), "r"(r + 1u) : "memory");
        }
      }
      bool late = false;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool fits = take[i] && rk[i] < (uint32_t)J;
        cand[i] = take[i];
        take[i] = fits;
        const uint32_t q = bb + (rk[i] * GCAP + id[i]) * 8u;
#pragma unroll
        for (int e = 0; e < NVE; ++e) sh_st64_if(q + (uint32_t)e * PLANE_BYTES, o[i].v[e], fits);
        if (META) sh_st64_if(q + (uint32_t)NVE * PLANE_BYTES, (global_row<CT>(P, o[i].row) << 8) | (uint64_t)(o[i].v_valid & 0xFFu), fits);
        late = late || (cand[i] && !fits) || cold[i];
      }
      if (late) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if ((cand[i] && !take[i]) || cold[i]) {
            uint32_t slot = OVF;
            if (cand[i]) slot = atomicAdd_shared_u32(sovf + 4u * oi);
            if (slot < (uint32_t)OVF) {
              sh_st32(sovf_id + ((uint32_t)b * OVF + slot) * 4u, id[i]);
              const uint32_t q = sovf_val + ((uint32_t)b * PLANES * OVF + slot) * 8u;
#pragma unroll
              for (int e = 0; e < NVE; ++e) sh_st64_if(q + (uint32_t)e * OVF * 8u, o[i].v[e], true);
              if (META) sh_st64_if(q + (uint32_t)NVE * OVF * 8u, (global_row<CT>(P, o[i].row) << 8) | (uint64_t)(o[i].v_valid & 0xFFu), true);
            } else {
              bool member = true;
              if (WIN) {   // the HBM path needs the row's own window word
                int64_t kk;
                member = window_of_ct<CT>(P, (int64_t)o[i].tval, kk);
                o[i].k[KW - 1] = (uint64_t)kk;
              }
              if (member) { bucket_cold_row<CT, KW, NV>(P, o[i], global_row<CT>(P, o[i].row)); ++spilled; }
            }
          }
      }
    }
    __syncthreads();
    // every thread has taken its rows out of the stage: the tile STAGES ahead may land there
    if (STAGES > 0 && WHOLE && tid == 0 && pass == 0 && t + STAGES < whole_hi) issue(t + STAGES, st_cur);
    if (WIN && (!have_w || wp != cur_w)) {   // a new window: the finished one leaves the registers
      if (have_w) { publish(cur_w); reset_acc(); }
      cur_w = wp; have_w = true;
    }
    // ---- fold ----
    if (SUB > 1) {
      // the PREVIOUS tile's counters: every thread finished reading them before it arrived at the barrier above, and
      // their next scatter comes after the next barrier
      const uint32_t cprev = scnt + (uint32_t)(ci == 0 ? NCNT - 1 : ci - 1) * (GCAP * 4u);
      for (int i = tid; i < GCAP; i += THREADS) sh_st32(cprev + 4u * i, 0u);
    }
    // the overflow counter two tiles ahead was last read in the fold before the barrier above
    if (tid == 0) sh_st32(sovf + 4u * (uint32_t)(oi == 0 ? NOVF - 1 : oi - 1), 0u);
    uint32_t n_ovf = sh_ld32(sovf + 4u * oi);
    n_ovf = n_ovf < (uint32_t)OVF ? n_ovf : (uint32_t)OVF;
#pragma unroll
    for (int gi = 0; gi < GPT; ++gi) {
      const int g = my_g0 + gi * THREADS;
      uint32_t c = sh_ld32(cc + 4u * g);
      if (SUB == 1 && c) sh_st32(cc + 4u * g, 0u);
      c = c < (uint32_t)J ? c : (uint32_t)J;
      uint32_t q = bb + ((uint32_t)my_sub * GCAP + g) * 8u;
      uint32_t mine = 0;
      constexpr uint32_t SKIPW = (LEN_ACC >= 0 ? (1u << LEN_ACC) : 0u) | (ROWPOS ? (FIRSTW | LASTW) : 0u);
      const BucketRegSink<NACC, SKIPW, NATIVE, (VAR & 1u) == 0> sink{acc[gi]};
#pragma unroll 1
      for (uint32_t j = my_sub; j < c; j += SUB, q += (uint32_t)SUB * GCAP * 8u) {
        RowOut<KW, NV> o;
#pragma unroll
        for (int e = 0; e < NV; ++e) o.v[e] = e < NVE ? sh_ld64(q + (uint32_t)e * PLANE_BYTES) : 0ull;
        uint64_t grow = 0;
        o.v_valid = 0xFFFFFFFFu;
        if (META) { const uint64_t m = sh_ld64(q + (uint32_t)NVE * PLANE_BYTES); grow = m >> 8; o.v_valid = (uint32_t)(m & 0xFFu); }
        o.tval = 0ull;
        accumulate_row<CT, NV, KW>(P, o, grow, sink);
        ++mine;
      }
#pragma unroll 1
      for (uint32_t i = 0; i < n_ovf; ++i) {
        if (my_sub != 0 || sh_ld32(sovf_id + ((uint32_t)b * OVF + i) * 4u) != (uint32_t)g) continue;
        const uint32_t qo = sovf_val + ((uint32_t)b * PLANES * OVF + i) * 8u;
        RowOut<KW, NV> o;
#pragma unroll
        for (int e = 0; e < NV; ++e) o.v[e] = e < NVE ? sh_ld64(qo + (uint32_t)e * OVF * 8u) : 0ull;
        uint64_t grow = 0;
        o.v_valid = 0xFFFFFFFFu;
        if (META) { const uint64_t m = sh_ld64(qo + (uint32_t)NVE * OVF * 8u); grow = m >> 8; o.v_valid = (uint32_t)(m & 0xFFu); }
        o.tval = 0ull;
        accumulate_row<CT, NV, KW>(P, o, grow, sink);
        ++mine;
      }
      if (ROWPOS && my_sub == 0 && mine + c > 0u) {
        // (the list scan above only adds to `mine` for sub-owner 0, and every bucketed or listed row registered itself)
        const uint32_t pa = spos + (uint32_t)b * (2u * GCAP * 4u) + 4u * (uint32_t)g;
        const uint64_t row0 = global_row<CT>(P, t * TILE);
        if (FIRSTW) {
          const uint32_t fp = sh_ld32(pa);
          if (fp != 0xFFFFFFFFu) {
            sh_st32(pa, 0xFFFFFFFFu);
            const uint64_t row = row0 + fp;
#pragma unroll
            for (int a = 0; a < NACC; ++a)
              if ((FIRSTW >> a) & 1u) { const uint64_t w = ((ROWTAG >> a) & 1u) ? ((row << 1) | 1ull) : row; acc[gi][a] = w < acc[gi][a] ? w : acc[gi][a]; }
          }
        }
        if (LASTW) {
          const uint32_t lp = sh_ld32(pa + GCAP * 4u);
          if (lp != 0u) {
            sh_st32(pa + GCAP * 4u, 0u);
            const uint64_t row = row0 + (lp - 1u);
#pragma unroll
            for (int a = 0; a < NACC; ++a)
              if ((LASTW >> a) & 1u) { const uint64_t w = ((ROWTAG >> a) & 1u) ? ((row << 1) | 1ull) : row; acc[gi][a] = w > acc[gi][a] ? w : acc[gi][a]; }
          }
        }
      }
      if (mine) {
        seen[gi] = true;
        if (LEN_ACC >= 0) {
#pragma unroll
          for (int a = 0; a < NACC; ++a) if (a == LEN_ACC) acc[gi][a] += (uint64_t)mine;
        }
      }
    }
    if (NBUF == 1) __syncthreads();
    b = (NBUF == 2) ? (b ^ 1) : 0;
    ci = ci + 1 == NCNT ? 0 : ci + 1;
    oi = oi + 1 == NOVF ? 0 : oi + 1;
  };

  // WINDOWED: window of one row of the time column (scalar read: from the stage when the tile is staged).  The window
  // of the previous tile's last row, or its successor, is almost always the answer: two range checks before a division.
  int64_t w_hint = 0;
  auto window_at = [&](int64_t row, uint32_t stage_base) -> int64_t {
    const int ts = CT::dyn_slot(P), dt = CT::slot_dtype(P, ts);
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (STAGES > 0 && stage_base != 0u) {
      const uint32_t a = stage_base + pick32<NC>(slot_off, ts) + (uint32_t)(row % TILE) * (uint32_t)dtype_width(dt);
      if (dtype_width(dt) == 8) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a) : "memory");
      else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r.x) : "r"(a) : "memory");
    } else r = load_pair(P.slots[ts].values, dt, row, row + 1, false);
    const int64_t t = (int64_t)decode(r, dt, 0), rel = t - P.dyn.origin;
    int64_t kk = w_hint;
    if (!in_window_at(rel, kk * EVERY)) {
      kk = w_hint + 1;
      if (!in_window_at(rel, kk * EVERY)) window_of_ct<CT>(P, t, kk);
    }
    w_hint = kk;
    return kk;
  };
  auto run_tile = [&](int64_t t, const uint4 (&raw)[NC], const uint32_t (&vb)[NC], auto whole_tag, uint32_t stage_base) {
    if (!WIN) { process(t, raw, vb, whole_tag, 0, 0, 0); return; }
    const int64_t last = (t + 1) * TILE <= n_rows ? (t + 1) * TILE - 1 : n_rows - 1;
    const int64_t w0 = window_at(t * TILE, stage_base), w1 = window_at(last, stage_base);
    process(t, raw, vb, whole_tag, w0, w1, 0);
    if (w1 != w0) process(t, raw, vb, whole_tag, w0, w1, 1);
  };

  // software pipeline over the whole tiles: the loads of the next tile are in flight while this one is scattered and
  // folded; the loop is unrolled by two so that the two register buffers alternate in place
  uint4 rawA[NC], rawB[NC];
  uint32_t vbA[NC], vbB[NC];
  if (STAGES > 0) {
    // staged pipeline: the copies run STAGES tiles ahead; a tile's rows come out of shared memory (conflict-free
    // consecutive 16-byte reads) once its mbarrier phase completes
    if (tid == 0)
      for (int k = 0; k < STAGES; ++k)
        if (tile_lo + k < whole_hi) issue(tile_lo + k, k);
    uint32_t parity = 0;
    for (int64_t t = tile_lo; t < whole_hi; ++t) {
      mbar_wait(sbar + 8u * st_cur, parity);
      const uint32_t base = sstage + (uint32_t)st_cur * stage_bytes;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        rawA[c] = make_uint4(0u, 0u, 0u, 0u);
        vbA[c] = 3u;
        if (c < CT::n_slots(P)) {
          const int w = dtype_width(CT::slot_dtype(P, c));
          rawA[c] = staged_pair(base + slot_off[c] + (uint32_t)lane_row * (uint32_t)w, w);
        }
      }
      run_tile(t, rawA, vbA, BucketWholeTag{}, base);
      if (++st_cur == STAGES) { st_cur = 0; parity ^= 1u; }
    }
  } else {
  if (tile_lo < whole_hi) bucket_load<CT, NC, true>(P, tile_lo * TILE + warp * 64, lane, n_rows, rawA, vbA);
  for (int64_t t = tile_lo; t < whole_hi; t += 2) {
    if (t + 1 < whole_hi) bucket_load<CT, NC, true>(P, (t + 1) * TILE + warp * 64, lane, n_rows, rawB, vbB);
    run_tile(t, rawA, vbA, BucketWholeTag{}, 0u);
    if (t + 1 < whole_hi) {
      if (t + 2 < whole_hi) bucket_load<CT, NC, true>(P, (t + 2) * TILE + warp * 64, lane, n_rows, rawA, vbA);
      run_tile(t + 1, rawB, vbB, BucketWholeTag{}, 0u);
    }
  }
  }
  // the ragged last tile (one CTA at most)
  if (whole_hi < tile_hi) {
    bucket_load<CT, NC, false>(P, whole_hi * TILE + warp * 64, lane, n_rows, rawA, vbA);
    run_tile(whole_hi, rawA, vbA, BucketRaggedTag{}, 0u);
  }

  // ---- publish the registers ----
  if (!WIN || have_w) publish(cur_w);
  if (spilled) atomicAdd(P.table.spilled, spilled);
}

}  // namespace pw
