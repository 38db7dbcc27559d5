// pw_partition.cuh — radix partitioning of the input rows by key hash (the high-cardinality tier).
//
// With ~1e7 groups the HBM table is far larger than the 126 MB L2 and every input row costs a random
// read-modify-write of its table row plus one L2 atomic per accumulator word (C3: 19 ms for 1e8 rows, 1.6 % of the
// HBM roofline).  The reference meets the same wall on the CPU and answers with partitioning
// (HashKeys::gen_idxs_per_partition, polars-expr/src/hash_keys.rs:263-314; partitioned group-by,
// polars-mem-engine/src/executors/group_by_partitioned.rs).  Here:
//   pass 1  histogram: rows per partition            (reads the key columns)
//   pass 2  scatter:   every surviving row is written, as 64-bit words (one per raw slot + a row-id word that also
//                      carries the validity bits), to its partition's range of a temporary frame.  Writes of one
//                      partition are consecutive, so L2 merges them into full sectors.
//   pass 3  the ordinary scan kernel over the temporary frame (ScanPlan::rowid_slot_p1): consecutive rows now share
//           a few hundred groups, so they aggregate in the shared-memory hot table and each group reaches the HBM
//           table once or twice instead of once per row.
// partition = umulhi(hash, n_parts) — the same top-bits map as the HBM table slot (umulhi(hash, cap)), so the
// flushes of one partition land in one contiguous piece of the table.
#pragma once
#include "pw_scan.cuh"

namespace pw {

// one row of the lane: decode, predicate, key words -> partition (or ~0 when the row is dropped)
template <class CT, int NC, int KW, int HF>
__device__ __forceinline__ uint32_t part_of_row(const ScanPlan& P, const PartParams& pp, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC],
                                                int j, int lane, int rem, Row<NC>& r) {
  row_decode<CT, NC>(P, raw[HF], vbits[HF], j, r);
  bool alive = (HF * 64 + 2 * lane + j) < rem && row_predicate<CT, NC>(P, r);
  uint64_t k[KW];
  bool sentinel_free;
  alive = row_keys<CT, NC, KW>(P, r, raw[HF], vbits[HF], j, alive, k, sentinel_free) && alive;
  return alive ? (uint32_t)__umul64hi(hash_words<KW>(k), (uint64_t)pp.n_parts) : 0xFFFFFFFFu;
}

// grid-stride over 128-row warp steps; no shared memory, as many CTAs as fit.  The four rows of a lane are handled
// together: their four position atomics are in flight at the same time (one at a time, the scatter pass was bound
// by the round trip of an atomic with return: 7.3 ms for 1e8 rows).
template <class CT, int NC, int KW>
__device__ __forceinline__ void part_body(const ScanPlan& P, const PartParams& pp) {
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
    Row<NC> r[4];
    uint32_t part[4];
    part[0] = part_of_row<CT, NC, KW, 0>(P, pp, raw, vbits, 0, lane, rem, r[0]);
    part[1] = part_of_row<CT, NC, KW, 0>(P, pp, raw, vbits, 1, lane, rem, r[1]);
    part[2] = part_of_row<CT, NC, KW, 1>(P, pp, raw, vbits, 0, lane, rem, r[2]);
    part[3] = part_of_row<CT, NC, KW, 1>(P, pp, raw, vbits, 1, lane, rem, r[3]);
    if (pp.mode == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (part[i] != 0xFFFFFFFFu) atomicAdd(pp.hist + part[i], 1u);
      continue;
    }
    uint32_t pos[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pos[i] = part[i] != 0xFFFFFFFFu ? atomicAdd(pp.cursor + part[i], 1u) : 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (part[i] == 0xFFFFFFFFu) continue;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < CT::n_slots(P)) pp.out[(uint64_t)c * pp.out_stride + pos[i]] = r[i].in[c];
      const uint64_t row = (uint64_t)(base + (i >> 1) * 64 + 2 * lane + (i & 1));
      pp.out[(uint64_t)CT::n_slots(P) * pp.out_stride + pos[i]] = (row << 8) | (uint64_t)(r[i].in_valid & 0xFFu);
    }
  }
}

}  // namespace pw
