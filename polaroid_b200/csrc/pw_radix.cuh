// pw_radix.cuh — the high-cardinality tier, second form: two-level radix partitioning by key hash with shared-memory
// staged, coalesced writes, then ONE CTA per partition aggregates it in a shared-memory table and appends the finished
// groups to the result table (strategy 10).
//
// The first form (pw_partition.cuh: one ~70 000-way scatter with a returning global atomic and three 8-byte stores per
// row, then the ordinary scan with its HBM table behind the hot table) ran at the L2's request rate, not at HBM speed:
// C3 = 8.3 ms partitioning + 6.3 ms scan for 2.0 GB of algorithmic traffic.  Here every pass streams:
//   mode 0  histogram of the FINAL partition of every surviving row (shared-memory counters, flushed once per CTA)
//   mode 1  frame -> level-1 partitions (<= 256): a 512-thread CTA ranks a 2048-row tile with shared-memory atomics, takes one
//           global cursor step per (tile, partition), sorts the tile's records by partition in shared memory and copies
//           them out as runs of consecutive 32-byte records (two threads per record: 16-byte stores, full sectors)
//   mode 2  level-1 partition -> final partitions (<= 256 per level-1 partition), same tile machinery
//   mode 3  one CTA per final partition: every row of a group is in the partition, so the shared-memory table holds
//           finished groups — no HBM hash table, no atomics on HBM; the groups are appended to a dense region of the
//           result table.  Keys that do not find a slot within `probe_limit` probes (skewed partitions, group estimate
//           too low) go to an open-addressing overflow region of the same table — a key is in exactly one of the two
//           (slots never empty again, so "no slot within the probe limit" is a stable property of a key).
// The reference's counterpart: HashKeys::gen_idxs_per_partition (polars-expr/src/hash_keys.rs:263-314) feeding one
// hash table per partition (polars-stream/src/nodes/group_by.rs:216-441).
#pragma once
#include "pw_scan.cuh"

namespace pw {

__device__ __forceinline__ uint32_t radix_part(uint64_t h, int log2_parts) { return log2_parts > 0 ? (uint32_t)(h >> (64 - log2_parts)) : 0u; }

// ---- the tile machinery shared by modes 1 and 2 -------------------------------------------------------------------------
struct RadixSmem {
  uint4* stage;      // [2 * RADIX_TILE]
  uint32_t* cnt;     // [256] records of the tile per bin
  uint32_t* pre;     // [256] exclusive prefix of cnt
  uint32_t* gb;      // [256] first global record of the tile's run per bin
  uint32_t* wsum;    // [8]
  uint32_t* misc;    // [4]
  uint16_t* binid;   // [RADIX_TILE]
};
__device__ __forceinline__ RadixSmem radix_smem(unsigned char* raw) {
  RadixSmem s;
  s.stage = (uint4*)raw;
  s.cnt = (uint32_t*)(s.stage + 2 * RADIX_TILE);
  s.pre = s.cnt + 256; s.gb = s.pre + 256; s.wsum = s.gb + 256; s.misc = s.wsum + 8;
  s.binid = (uint16_t*)(s.misc + 4);
  return s;
}
constexpr int RADIX_SCATTER_SMEM = 2 * RADIX_TILE * 16 + (3 * 256 + 8 + 4) * 4 + RADIX_TILE * 2;

// Called by every thread of the CTA after the tile's ranks were taken (cnt complete behind a barrier).  Thread t holds up
// to 4 records (lo/hi), their bin (0xFFFF = none) and rank inside the bin.
__device__ __forceinline__ void radix_flush_tile(const RadixSmem& s, uint32_t* cursor, uint32_t cursor_base, uint4* dst,
                                                 const uint4 (&lo)[4], const uint4 (&hi)[4], const uint32_t (&bin)[4], const uint32_t (&rank)[4]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t v = 0, incl = 0;
  if (tid < 256) {
    v = s.cnt[tid];
    incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s.wsum[warp] = incl;
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s.wsum[w];
    s.pre[tid] = off + incl - v;
    if (v) s.gb[tid] = atomicAdd(cursor + cursor_base + tid, v);
    if (tid == 255) s.misc[0] = off + incl;  // records of the tile
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (bin[i] == 0xFFFFu) continue;
    const uint32_t pos = s.pre[bin[i]] + rank[i];
    s.stage[2 * pos] = lo[i];
    s.stage[2 * pos + 1] = hi[i];
    s.binid[pos] = (uint16_t)bin[i];
  }
  __syncthreads();
  const uint32_t total = s.misc[0];
  for (uint32_t i = tid; i < 2 * total; i += RADIX_SC_THREADS) {
    const uint32_t rec = i >> 1, b = s.binid[rec];
    dst[2 * ((uint64_t)s.gb[b] + (rec - s.pre[b])) + (i & 1u)] = s.stage[i];
  }
  __syncthreads();
}

// one row of the frame: decode, predicate, key words -> hash; false when the row is dropped
template <class CT, int NC, int KW, int HF>
__device__ __forceinline__ bool radix_frame_row(const ScanPlan& P, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC], int j, int lane, int rem,
                                                Row<NC>& r, uint64_t& h) {
  row_decode<CT, NC>(P, raw[HF], vbits[HF], j, r);
  bool alive = (HF * 64 + 2 * lane + j) < rem && row_predicate<CT, NC>(P, r);
  uint64_t k[KW];
  bool sentinel_free;
  alive = row_keys<CT, NC, KW>(P, r, raw[HF], vbits[HF], j, alive, k, sentinel_free) && alive;
  h = hash_words<KW>(k);
  return alive;
}

// a 32-byte record -> Row (slots are canonical 64-bit words; the last used word carries row << 8 | validity bits)
template <class CT, int NC>
__device__ __forceinline__ void radix_record_row(const ScanPlan& P, const uint4& lo, const uint4& hi, Row<NC>& r, uint64_t& rowid) {
  const uint64_t w[4] = {(uint64_t)lo.y << 32 | lo.x, (uint64_t)lo.w << 32 | lo.z, (uint64_t)hi.y << 32 | hi.x, (uint64_t)hi.w << 32 | hi.z};
#pragma unroll
  for (int c = 0; c < NC; ++c) r.in[c] = c < 4 ? w[c] : 0ull;
  rowid = pick<4>(w, CT::rowid_slot(P));
  r.in_valid = (uint32_t)rowid & 0xFFu;
}
template <class CT, int NC, int KW>
__device__ __forceinline__ uint64_t radix_record_hash(const ScanPlan& P, const Row<NC>& r, uint64_t (&k)[KW], bool& sentinel_free) {
  const uint4 noraw[NC] = {};
  const uint32_t novb[NC] = {};
  row_keys<CT, NC, KW>(P, r, noraw, novb, 0, true, k, sentinel_free);
  return hash_words<KW>(k);
}

// ---- mode 0: histogram -------------------------------------------------------------------------------------------------
template <class CT, int NC, int KW>
__device__ __forceinline__ void radix_hist_body(const ScanPlan& P, const RadixParams& rp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = (uint32_t*)smem_raw;
  const uint32_t n_parts = 1u << rp.log2_parts;
  for (uint32_t i = threadIdx.x; i < n_parts; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_rows = P.n_rows;
  const int64_t n_steps = (n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  for (int64_t step = warp; step < n_steps; step += n_warps) {
    const int64_t base = step * ROWS_PER_STEP;
    const int64_t left = n_rows - base;
    const int rem = left >= ROWS_PER_STEP ? ROWS_PER_STEP : (int)left;
    uint4 raw[2][NC];
    uint32_t vbits[2][NC];
    load_step<CT, NC>(P, base, lane, n_rows, raw, vbits);
    Row<NC> r;
    uint64_t h;
    if (radix_frame_row<CT, NC, KW, 0>(P, raw, vbits, 0, lane, rem, r, h)) atomicAdd(&hist[radix_part(h, rp.log2_parts)], 1u);
    if (radix_frame_row<CT, NC, KW, 0>(P, raw, vbits, 1, lane, rem, r, h)) atomicAdd(&hist[radix_part(h, rp.log2_parts)], 1u);
    if (radix_frame_row<CT, NC, KW, 1>(P, raw, vbits, 0, lane, rem, r, h)) atomicAdd(&hist[radix_part(h, rp.log2_parts)], 1u);
    if (radix_frame_row<CT, NC, KW, 1>(P, raw, vbits, 1, lane, rem, r, h)) atomicAdd(&hist[radix_part(h, rp.log2_parts)], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n_parts; i += blockDim.x) {
    const uint32_t v = hist[i];
    if (v) atomicAdd(rp.hist + i, v);
  }
}

// ---- mode 1: frame -> level-1 partitions -------------------------------------------------------------------------------
template <class CT, int NC, int KW, int HF>
__device__ __forceinline__ void radix_take_frame_row(const ScanPlan& P, const RadixParams& rp, const RadixSmem& s, const uint4 (&raw)[2][NC],
                                                     const uint32_t (&vbits)[2][NC], int j, int lane, int rem, int64_t base, uint4& lo, uint4& hi,
                                                     uint32_t& bin, uint32_t& rank) {
  Row<NC> r;
  uint64_t h;
  bin = 0xFFFFu; rank = 0u;
  lo = make_uint4(0u, 0u, 0u, 0u); hi = lo;
  if (!radix_frame_row<CT, NC, KW, HF>(P, raw, vbits, j, lane, rem, r, h)) return;
  bin = radix_part(h, rp.log2_parts) >> rp.log2_p2;
  rank = atomicAdd(&s.cnt[bin], 1u);
  uint64_t w[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < 3 && c < CT::n_slots(P)) w[c] = r.in[c];
  const uint64_t row = (uint64_t)(base + HF * 64 + 2 * lane + j);
  put<4>(w, CT::n_slots(P), (row << 8) | (uint64_t)(r.in_valid & 0xFFu));
  lo = make_uint4((uint32_t)w[0], (uint32_t)(w[0] >> 32), (uint32_t)w[1], (uint32_t)(w[1] >> 32));
  hi = make_uint4((uint32_t)w[2], (uint32_t)(w[2] >> 32), (uint32_t)w[3], (uint32_t)(w[3] >> 32));
}

template <class CT, int NC, int KW>
__device__ __forceinline__ void radix_scatter_frame_body(const ScanPlan& P, const RadixParams& rp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RadixSmem s = radix_smem(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_rows = P.n_rows;
  const int64_t n_tiles = (n_rows + RADIX_TILE - 1) / RADIX_TILE;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (tid < 256) s.cnt[tid] = 0u;
    __syncthreads();
    const int64_t base = tile * RADIX_TILE + (int64_t)warp * ROWS_PER_STEP;
    uint4 lo[4], hi[4];
    uint32_t bin[4], rank[4];
    if (base < n_rows) {
      const int64_t left = n_rows - base;
      const int rem = left >= ROWS_PER_STEP ? ROWS_PER_STEP : (int)left;
      uint4 raw[2][NC];
      uint32_t vbits[2][NC];
      load_step<CT, NC>(P, base, lane, n_rows, raw, vbits);
      radix_take_frame_row<CT, NC, KW, 0>(P, rp, s, raw, vbits, 0, lane, rem, base, lo[0], hi[0], bin[0], rank[0]);
      radix_take_frame_row<CT, NC, KW, 0>(P, rp, s, raw, vbits, 1, lane, rem, base, lo[1], hi[1], bin[1], rank[1]);
      radix_take_frame_row<CT, NC, KW, 1>(P, rp, s, raw, vbits, 0, lane, rem, base, lo[2], hi[2], bin[2], rank[2]);
      radix_take_frame_row<CT, NC, KW, 1>(P, rp, s, raw, vbits, 1, lane, rem, base, lo[3], hi[3], bin[3], rank[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { bin[i] = 0xFFFFu; rank[i] = 0u; lo[i] = make_uint4(0u, 0u, 0u, 0u); hi[i] = lo[i]; }
    }
    __syncthreads();
    radix_flush_tile(s, rp.cursor, 0u, rp.dst, lo, hi, bin, rank);
  }
}

// ---- mode 2: level-1 partition -> final partitions ---------------------------------------------------------------------
template <class CT, int NC, int KW>
__device__ __forceinline__ void radix_scatter_records_body(const ScanPlan& P, const RadixParams& rp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RadixSmem s = radix_smem(smem_raw);
  const int tid = threadIdx.x;
  const uint32_t n_l1 = 1u << (rp.log2_parts - rp.log2_p2);
  const uint32_t n_tiles = rp.tile_first[n_l1];
  const uint32_t mask2 = (1u << rp.log2_p2) - 1u;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (tid < 256) s.cnt[tid] = 0u;
    if (tid == 0) {
      uint32_t lo_ = 0, hi_ = n_l1;   // last level-1 partition whose first tile <= tile
      while (hi_ - lo_ > 1) { const uint32_t mid = (lo_ + hi_) >> 1; if (rp.tile_first[mid] <= tile) lo_ = mid; else hi_ = mid; }
      s.misc[1] = lo_;
    }
    __syncthreads();
    const uint32_t l1 = s.misc[1];
    const uint32_t begin = rp.offs[l1 << rp.log2_p2] + (tile - rp.tile_first[l1]) * (uint32_t)RADIX_TILE;
    const uint32_t end = min(rp.offs[(l1 + 1) << rp.log2_p2], begin + (uint32_t)RADIX_TILE);
    uint4 lo[4], hi[4];
    uint32_t bin[4], rank[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t idx = begin + (uint32_t)i * RADIX_SC_THREADS + tid;
      bin[i] = 0xFFFFu; rank[i] = 0u;
      lo[i] = make_uint4(0u, 0u, 0u, 0u); hi[i] = lo[i];
      if (idx < end) { lo[i] = __ldcs(rp.src + 2 * (uint64_t)idx); hi[i] = __ldcs(rp.src + 2 * (uint64_t)idx + 1); bin[i] = 0u; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (bin[i] == 0xFFFFu) continue;
      Row<NC> r;
      uint64_t rowid, k[KW];
      bool sf;
      radix_record_row<CT, NC>(P, lo[i], hi[i], r, rowid);
      const uint64_t h = radix_record_hash<CT, NC, KW>(P, r, k, sf);
      bin[i] = radix_part(h, rp.log2_parts) & mask2;
      rank[i] = atomicAdd(&s.cnt[bin[i]], 1u);
    }
    __syncthreads();
    radix_flush_tile(s, rp.cursor, l1 << rp.log2_p2, rp.dst, lo, hi, bin, rank);
  }
}

// ---- mode 3: one CTA per final partition ---------------------------------------------------------------------------------
// Shared-memory 64-bit arithmetic atomics are CAS loops on sm_100a (ATOMS.CAST.SPIN): a first version that applied every
// accumulator with one of them per row spent 6.1 ms on C3.  Here a row costs one key probe and ONE native 32-bit atomic
// (its rank among the rows of its group in this round); the round's rows are then laid out group by group in a
// shared-memory staging area (counting sort by table slot) and every slot's OWNER thread folds its rows into registers —
// plain loads, no atomics — and writes the accumulators back to the table.
constexpr int RADIX_ROUND = 3;   // rows a thread stages per round (round = RADIX_ROUND * blockDim.x rows)
struct SmemTable {
  uint64_t* keys;    // [KW][S]
  uint64_t* accs;    // [n_acc][S]
  uint32_t* state;   // [S]   (KW > 1)
  uint32_t* cnt;     // [S]   rows of the current round
  uint32_t* start;   // [S]   first staged row of the slot
  uint64_t* stage;   // [n_slots + 1][RADIX_ROUND * blockDim.x]
  uint16_t* occ;     // [S]   occupied slots in insertion order
  uint32_t S;
};
template <int NACC>
struct RadixRegSink {
  uint64_t (&acc)[NACC];
  template <int OP>
  __device__ __forceinline__ void add(const ScanPlan&, int a, uint64_t x) const {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      if (i == a) acc[i] = acc_combine(OP, acc[i], x);
  }
};

// find-or-insert in the CTA's table; ~0 when no slot within the probe limit
template <int KW>
__device__ __forceinline__ uint32_t smem_upsert(const SmemTable& T, const uint64_t (&k)[KW], uint32_t s0, int limit, uint32_t* n_ins, bool& inserted) {
  const uint32_t mask = T.S - 1u;
  uint32_t s = s0 & mask;
  if (KW == 1) {
    const uint64_t k0 = k[0];
    for (int probes = 0; probes < limit; ++probes) {
      unsigned long long old = *(volatile unsigned long long*)&T.keys[s];
      if (old == k0) return s;
      if (old == KEY_EMPTY) {
        old = atomicCAS((unsigned long long*)&T.keys[s], (unsigned long long)KEY_EMPTY, (unsigned long long)k0);
        if (old == KEY_EMPTY) { T.occ[atomicAdd(n_ins, 1u)] = (uint16_t)s; inserted = true; return s; }
        if (old == k0) return s;
      }
      s = (s + 1u) & mask;
    }
    return ~0u;
  } else {
    int probes = 0;
    uint32_t result = ~0u;
    bool done = false;
    while (!done) {
      uint32_t st = *(volatile uint32_t*)&T.state[s];
      if (st == 0u) st = atomicCAS(&T.state[s], 0u, 1u) == 0u ? 3u : 1u;  // 3: we own the slot
      if (st == 3u) {
#pragma unroll
        for (int w = 0; w < KW; ++w) *(volatile uint64_t*)&T.keys[(uint32_t)w * T.S + s] = k[w];
        __threadfence_block();
        *(volatile uint32_t*)&T.state[s] = 2u;
        T.occ[atomicAdd(n_ins, 1u)] = (uint16_t)s;
        inserted = true;
        result = s; done = true;
      } else if (st == 2u) {
        __threadfence_block();
        bool eq = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) eq &= (*(volatile uint64_t*)&T.keys[(uint32_t)w * T.S + s] == k[w]);
        if (eq) { result = s; done = true; }
        else {
          s = (s + 1u) & mask;
          if (++probes >= limit) done = true;
        }
      }
      // st == 1: the owner is writing the key words -> look again
    }
    return result;
  }
}

// staged rows q_begin, q_begin + q_step, ... < q_end folded into acc
template <class CT, int NC, int KW, int NACC>
__device__ __forceinline__ void radix_fold(const ScanPlan& P, const SmemTable& T, uint32_t cap_rows, int n_words, uint64_t (&acc)[NACC],
                                           uint32_t q_begin, uint32_t q_end, uint32_t q_step) {
  constexpr int NV = NVof<NC>::value;
  const RadixRegSink<NACC> sink{acc};
  for (uint32_t q = q_begin; q < q_end; q += q_step) {
    Row<NC> r;
#pragma unroll
    for (int c = 0; c < NC; ++c) r.in[c] = (c < 4 && c < n_words) ? T.stage[(uint32_t)c * cap_rows + q] : 0ull;
    const uint64_t rowid = pick<NC>(r.in, CT::rowid_slot(P));
    r.in_valid = (uint32_t)rowid & 0xFFu;
    RowOut<KW, NV> o;
    o.alive = true; o.sentinel_free = true;
    o.row = (int64_t)(rowid >> 8);
    o.tval = 0ull;
    row_vexprs<CT, NC, NV>(P, r, o.v, o.v_valid);
    accumulate_row<CT, NV, KW, RadixRegSink<NACC>>(P, o, global_row<CT>(P, o.row), sink);
  }
}

template <class CT, int NC, int KW>
__device__ __forceinline__ void radix_aggregate_body(const ScanPlan& P, const RadixParams& rp) {
  constexpr int NV = NVof<NC>::value;
  constexpr int NACC = CT::kNAcc;
  const uint32_t NT = blockDim.x;                 // 1024 (one CTA per SM) or 512 (two per SM, half-size tables): RadixParams::agg_threads
  const uint32_t CAP = RADIX_ROUND * NT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_acc = CT::n_acc(P), n_words = CT::n_slots(P);   // n_slots of the record plan counts the row-id word
  SmemTable T;
  T.S = 1u << rp.log2_slots;
  T.keys = (uint64_t*)smem_raw;
  T.accs = T.keys + (size_t)KW * T.S;
  T.stage = T.accs + (size_t)n_acc * T.S;
  T.cnt = (uint32_t*)(T.stage + (size_t)n_words * CAP);
  T.start = T.cnt + T.S;
  T.state = T.start + T.S;
  uint32_t* ctl = T.state + (KW > 1 ? T.S : 0u);   // [0] inserted groups, [1] emit rank, [2..3] dense base, [4] fits, [5] overflow seen, [6] slots with rows
  uint32_t* wsum = ctl + 8;      // [32] block scan
  uint32_t* chist = ctl + 40;    // [64] slots per row count (descending)
  uint16_t* order = (uint16_t*)(ctl + 104);   // [S] slots that have rows in this round, longest first
  T.occ = order + T.S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_parts = 1u << rp.log2_parts;
  const Table& G = P.table;          // host view: [overflow region | dense region | two escape slots]
  Table ovf = G; ovf.cap = rp.ovf_cap;
  unsigned long long spilled = 0;
  for (;;) {
    // partitions are handed out one at a time (a partition that holds a heavy group takes several times the average)
    __syncthreads();
    if (tid == 0) ctl[7] = (uint32_t)atomicAdd(rp.dense_count + 1, 1ull);
    __syncthreads();
    const uint32_t part = ctl[7];
    if (part >= n_parts) break;
    const uint32_t begin = rp.offs[part], end = rp.offs[part + 1];
    if (begin == end) continue;
    for (uint32_t s = tid; s < T.S; s += blockDim.x) {
      if (KW == 1) T.keys[s] = KEY_EMPTY; else T.state[s] = 0u;
      T.cnt[s] = 0u;   // (the accumulators of a slot are initialised by the thread that inserts its key)
    }
    if (tid < 2) ctl[tid] = 0u;
    if (tid == 2) ctl[5] = (uint32_t)*(volatile int32_t*)G.overflow;
    __syncthreads();
    if (ctl[5] != 0u) break;   // (CTA-uniform) the table overflowed somewhere: the result is discarded anyway
    for (uint64_t r0 = begin; r0 < end; r0 += CAP) {
      // ---- phase A: key probe + rank of every row of the round
      uint4 lo[RADIX_ROUND], hi[RADIX_ROUND];
      uint32_t slot[RADIX_ROUND], rank[RADIX_ROUND];
#pragma unroll
      for (int u = 0; u < RADIX_ROUND; ++u) {
        const uint64_t idx = r0 + (uint64_t)u * NT + tid;
        slot[u] = ~0u; rank[u] = 0u;
        if (idx < end) { lo[u] = __ldcs(rp.src + 2 * idx); hi[u] = __ldcs(rp.src + 2 * idx + 1); slot[u] = 0u; }
      }
#pragma unroll
      for (int u = 0; u < RADIX_ROUND; ++u) {
        if (slot[u] == ~0u) continue;
        Row<NC> r;
        uint64_t rowid;
        RowOut<KW, NV> o;
        radix_record_row<CT, NC>(P, lo[u], hi[u], r, rowid);
        const uint64_t h = radix_record_hash<CT, NC, KW>(P, r, o.k, o.sentinel_free);
        slot[u] = ~0u;
        bool inserted = false;
        if (o.sentinel_free) slot[u] = smem_upsert<KW>(T, o.k, (uint32_t)((h << rp.log2_parts) >> (64 - rp.log2_slots)), rp.probe_limit, &ctl[0], inserted);
        if (inserted) {
#pragma unroll
          for (int a = 0; a < NACC; ++a) T.accs[(uint32_t)a * T.S + slot[u]] = acc_init(CT::acc_op(P, a));
        }
        if (slot[u] != ~0u) rank[u] = atomicAdd(&T.cnt[slot[u]], 1u);
        else {
          // a data value equal to a key sentinel lives in the escape slots of the whole table; everything else that found
          // no slot goes to the overflow region (hash re-mixed: the partition's keys share their top bits; bounded
          // probing: a filling overflow region raises the overflow flag instead of crawling — the host repeats the
          // query on the plain HBM table)
          o.alive = true;
          o.row = (int64_t)(rowid >> 8);
          o.tval = 0ull;
          row_vexprs<CT, NC, NV>(P, r, o.v, o.v_valid);
          uint64_t gs = ~0ull;
          if (!o.sentinel_free) gs = table_upsert<KW>(G, o.k, h, false);
          else if (*(volatile int32_t*)G.overflow == 0) gs = table_upsert<KW>(ovf, o.k, mix64(h), true, 128);
          if (gs != ~0ull) {
            const ColdSink sink{G, gs};
            accumulate_row<CT, NV, KW, ColdSink>(P, o, global_row<CT>(P, o.row), sink);
          }
          ++spilled;
        }
      }
      __syncthreads();
      // ---- phase B: exclusive prefix of the slots' row counts
      {
        const uint32_t per = T.S / NT > 0 ? T.S / NT : 1u;   // consecutive slots per thread
        const uint32_t s_b = (uint32_t)tid * per;
        uint32_t sum = 0;
        if (s_b < T.S)
          for (uint32_t i = 0; i < per; ++i) sum += T.cnt[s_b + i];
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t off = lane < warp ? wsum[lane] : 0u;   // totals of the warps before this one (32 warps: one per lane)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) off += __shfl_xor_sync(0xffffffffu, off, d);
        uint32_t run = off + incl - sum;
        if (s_b < T.S)
          for (uint32_t i = 0; i < per; ++i) { T.start[s_b + i] = run; run += T.cnt[s_b + i]; }
      }
      __syncthreads();
      // ---- phase C: the round's rows, group by group
#pragma unroll
      for (int u = 0; u < RADIX_ROUND; ++u) {
        if (slot[u] == ~0u) continue;
        const uint32_t pos = T.start[slot[u]] + rank[u];
        const uint64_t w[4] = {(uint64_t)lo[u].y << 32 | lo[u].x, (uint64_t)lo[u].w << 32 | lo[u].z, (uint64_t)hi[u].y << 32 | hi[u].x, (uint64_t)hi[u].w << 32 | hi[u].z};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < n_words) T.stage[(uint32_t)c * CAP + pos] = w[c];
      }
      __syncthreads();
      // ---- slots that have rows, ordered by row count (counting sort, longest first): the lanes of a warp then fold
      // runs of (nearly) the same length, and warps without work skip the phase — folding slot tid, tid + 1024, ...
      // in place ran at ~20 % lane efficiency (9.0 ms on C3)
      {
        uint32_t my_b[4], my_r[4];
        if (tid < 64) chist[tid] = 0u;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t s = (uint32_t)tid + (uint32_t)i * NT;
          my_b[i] = ~0u; my_r[i] = 0u;
          if (s < T.S) {
            const uint32_t n = T.cnt[s];
            if (n) { my_b[i] = 63u - min(n, 63u); my_r[i] = atomicAdd(&chist[my_b[i]], 1u); }
          }
        }
        __syncthreads();
        if (warp == 0) {
          const uint32_t v0 = chist[2 * lane], v1 = chist[2 * lane + 1];
          uint32_t incl = v0 + v1;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
          chist[2 * lane] = incl - v0 - v1;
          chist[2 * lane + 1] = incl - v1;
          if (lane == 31) ctl[6] = incl;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (my_b[i] != ~0u) order[chist[my_b[i]] + my_r[i]] = (uint16_t)((uint32_t)tid + (uint32_t)i * NT);
        __syncthreads();
      }
      // ---- phase D: every slot's rows are folded in registers — by its owner thread, or, for a slot with 63 or more
      // rows in the round (heavy hitters: the null-key group of C3 holds 1e5 rows, and one thread folding them alone
      // was two thirds of the kernel's time), by a whole warp whose lanes then combine by shuffles
      const uint32_t n_busy = ctl[6], n_heavy = chist[1];
      for (uint32_t t = warp; t < n_heavy; t += NT / 32) {
        const uint32_t s = order[t];
        const uint32_t n = T.cnt[s], first = T.start[s];
        uint64_t acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = acc_init(CT::acc_op(P, a));
        radix_fold<CT, NC, KW, NACC>(P, T, CAP, n_words, acc, first + lane, first + n, 32u);
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) acc[a] = acc_combine(CT::acc_op(P, a), acc[a], __shfl_down_sync(0xffffffffu, (unsigned long long)acc[a], d));
        }
        if (lane == 0) {
#pragma unroll
          for (int a = 0; a < NACC; ++a) T.accs[(uint32_t)a * T.S + s] = acc_combine(CT::acc_op(P, a), T.accs[(uint32_t)a * T.S + s], acc[a]);
          T.cnt[s] = 0u;
        }
      }
      for (uint32_t t = n_heavy + tid; t < n_busy; t += blockDim.x) {
        const uint32_t s = order[t];
        const uint32_t n = T.cnt[s], first = T.start[s];
        uint64_t acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = T.accs[(uint32_t)a * T.S + s];
        radix_fold<CT, NC, KW, NACC>(P, T, CAP, n_words, acc, first, first + n, 1u);
#pragma unroll
        for (int a = 0; a < NACC; ++a) T.accs[(uint32_t)a * T.S + s] = acc[a];
        T.cnt[s] = 0u;
      }
      __syncthreads();
    }
    const uint32_t n_groups = ctl[0];
    if (tid == 0 && n_groups) {
      const unsigned long long b = atomicAdd(rp.dense_count, (unsigned long long)n_groups);
      ctl[2] = (uint32_t)b; ctl[3] = (uint32_t)(b >> 32);
      ctl[4] = b + n_groups <= rp.dense_cap ? 1u : 0u;
      if (!ctl[4]) *G.overflow = 1;
    }
    __syncthreads();
    if (n_groups && ctl[4]) {
      const uint64_t dbase = rp.ovf_cap + ((uint64_t)ctl[3] << 32 | ctl[2]);
      for (uint32_t t = tid; t < n_groups; t += blockDim.x) {
        const uint32_t s = T.occ[t];
        const uint64_t d = dbase + t;
        uint64_t* kp = &tkey(G, 0, d);
        uint64_t* ap = &tacc(G, 0, d);
#pragma unroll
        for (int w = 0; w < KW; ++w) kp[(uint64_t)w * G.key_sw] = T.keys[(uint32_t)w * T.S + s];
#pragma unroll
        for (int a = 0; a < NACC; ++a) ap[(uint64_t)a * G.acc_sw] = T.accs[(uint32_t)a * T.S + s];
        if (KW > 1) G.state[d] = 2u;
      }
    }
    __syncthreads();
  }
  if (spilled) atomicAdd(G.spilled, spilled);
}

template <class CT, int NC, int KW, int MODE>
__device__ __forceinline__ void radix_body(const ScanPlan& P, const RadixParams& rp) {
  if (MODE == 0) radix_hist_body<CT, NC, KW>(P, rp);
  else if (MODE == 1) radix_scatter_frame_body<CT, NC, KW>(P, rp);
  else if (MODE == 2) radix_scatter_records_body<CT, NC, KW>(P, rp);
  else radix_aggregate_body<CT, NC, KW>(P, rp);
}

}  // namespace pw
