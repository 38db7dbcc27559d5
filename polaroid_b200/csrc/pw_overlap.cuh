// pw_overlap.cuh — group_by_dynamic with OVERLAPPING windows (period > every) through the hash path: with `group_by=`
// keys, and / or with a filter (the keys-empty, unfiltered case is pw_dynamic.cu's closed-form slice path).
//
// Window starts lie on the grid offset + i * every for every key slice alike, and group_by_windows emits the NON-EMPTY
// windows of a slice, beginning with the slice's FIRST window: truncate(t0, every) + offset for the slice's earliest
// index value t0, moved back by whole steps while it starts after t0 (polars-time/src/windows/window.rs:115-170
// get_earliest_bounds + ensure_t_in_or_in_front_of_window, 342-438 BoundsIter; windows/group_by.rs:79-246) — earlier grid
// windows that also contain t0 do not exist.  Two passes over the rows:
//   pass 1 (ScanPlan::overlap == 2)  t0 per key slice: (key words, window word 0) -> MIN of the index value, in its own
//                                    small table (the rows of a slice are ascending, but the filter may drop its head)
//   pass 2 (ScanPlan::overlap == 1)  every row looks its slice's t0 up, derives the first window i0 and joins ALL the
//                                    windows [max(lo, i0), hi] that contain its index value: one find-or-insert of
//                                    (key words, window index) in the HBM table and one set of atomics per (row, window)
// The groups that exist afterwards are the reference's windows.  first / last stay exact through the global row index;
// the result is ordered by key slice, then window (lower_query's sort words), like the reference's.
#pragma once
#include "pw_scan.cuh"

namespace pw {

template <class CT, int NC, int KW, int HF>
__device__ __forceinline__ void overlap_row(const ScanPlan& P, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC], int j, int64_t base,
                                            int lane, int rem) {
  constexpr int NV = NVof<NC>::value;
  Row<NC> r;
  row_decode<CT, NC>(P, raw[HF], vbits[HF], j, r);
  const bool alive = (HF * 64 + 2 * lane + j) < rem && row_predicate<CT, NC>(P, r);
  if (!alive) return;
  RowOut<KW, NV> o;
  (void)row_keys<CT, NC, KW>(P, r, raw[HF], vbits[HF], j, true, o.k, o.sentinel_free);   // the window word is set below
  o.alive = true;
  o.row = base + HF * 64 + 2 * lane + j;
  row_vexprs<CT, NC, NV>(P, r, o.v, o.v_valid);
  const int64_t t = (int64_t)pick<NC>(r.in, CT::dyn_slot(P));
  o.tval = (uint64_t)t;
  int wpos = 0;   // the window word follows the key columns' words
#pragma unroll
  for (int q = 0; q < MAX_KEYS; ++q)
    if (q < CT::n_keys(P)) wpos += CT::key_words(P, q);
  const Dyn& d = P.dyn;
  const int closed = CT::dyn_closed(P);
  const bool right_closed = closed == 1 || closed == 2, left_closed = closed == 0 || closed == 2;
  const int64_t rel = t - d.origin;
  // exit condition t < start + period (or <=) -> smallest window; entry condition t >= start (or >) -> largest window
  const int64_t lo = right_closed ? -floor_div(-(rel - d.period), d.every) : floor_div(rel - d.period, d.every) + 1;
  const int64_t hi = left_closed ? floor_div(rel, d.every) : -floor_div(-rel, d.every) - 1;
  // the slice's earliest index value
  put<KW>(o.k, wpos, 0ull);
  bool sf0 = true;
  if (KW == 1) sf0 = o.k[0] < KEY_NULL;
  const uint64_t s0 = table_upsert<KW>(P.t0, o.k, hash_words<KW>(o.k), sf0);
  if (s0 == ~0ull) return;       // table full: the overflow flag is up, the host grows the tables and repeats
  if (P.overlap == 2) {
    if (t < (int64_t)__ldcg((const unsigned long long*)&tacc(P.t0, 0, s0))) atomicMin((long long*)&tacc(P.t0, 0, s0), (long long)t);
    return;
  }
  const int64_t t0 = (int64_t)__ldcg((const unsigned long long*)&tacc(P.t0, 0, s0));
  // first window of the slice: truncate(t0, every) + offset, or the last grid window that does not start after t0
  const int64_t i_trunc = floor_div(t0, d.every);
  const int64_t i_front = left_closed ? floor_div(t0 - d.origin, d.every) : -floor_div(-(t0 - d.origin), d.every) - 1;
  const int64_t i0 = i_trunc < i_front ? i_trunc : i_front;
  const uint64_t grow = global_row<CT>(P, o.row);
  for (int64_t w = lo > i0 ? lo : i0; w <= hi; ++w) {
    put<KW>(o.k, wpos, (uint64_t)w);
    bool sf = true;
    if (KW == 1) sf = o.k[0] < KEY_NULL;   // window -1 / -2 alias the single-word key sentinels: escape slots
    const uint64_t slot = table_upsert<KW>(P.table, o.k, hash_words<KW>(o.k), sf);
    if (slot == ~0ull) return;   // table full: the overflow flag is up, the host grows the table and repeats
    const ColdSink sink{P.table, slot};
    accumulate_row<CT, NV, KW, ColdSink>(P, o, grow, sink);
  }
}

template <class CT, int NC, int KW>
__device__ __forceinline__ void overlap_body(const ScanPlan& P) {
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
    if (CT::check_sorted(P) && P.overlap == 2) check_sorted_step<CT, NC>(P, raw, base, lane, n_rows);
    overlap_row<CT, NC, KW, 0>(P, raw, vbits, 0, base, lane, rem);
    overlap_row<CT, NC, KW, 0>(P, raw, vbits, 1, base, lane, rem);
    overlap_row<CT, NC, KW, 1>(P, raw, vbits, 0, base, lane, rem);
    overlap_row<CT, NC, KW, 1>(P, raw, vbits, 1, base, lane, rem);
  }
}

}  // namespace pw
