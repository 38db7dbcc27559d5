// pw_runs.cuh — SORTED KEYS: the run-boundary + segmented-reduce path.
//
// When the key columns are sorted, the rows of a group are one contiguous run ("GroupsSlice": [first, len] per group —
// polars-core/src/frame/group_by/into_groups.rs:65-129 create_groups_from_sorted, polars-arrow/src/legacy/kernels/
// sort_partition.rs:168 partition_to_groups), and the reference reduces each slice on its own.  Here a warp walks its
// rows in order and keeps the accumulators of its CURRENT run in registers (as pw_segmented.cuh does for the current
// window): a 64-row slot whose rows all carry the current key costs one ALU op per accumulator word and row — no hashing,
// no shared-memory table, no atomics.  Only where a run ends inside the slot does the warp reduce its registers with
// shuffles, publish them under the finished key (one find-or-insert + one atomic per word) and send the rows of the
// slot straight to the HBM table; the last row's key becomes the new current run.
// The HBM table merges by key, so the result is correct whatever the order of the rows — a key that comes back later
// (the caller's "sorted" promise was wrong, or the runs come from clustered, not sorted, data) simply merges; the tier is
// chosen for speed when runs are long (caller flag PW_FLAG_KEYS_SORTED), not for correctness.
#pragma once
#include "pw_segmented.cuh"

namespace pw {

template <class CT, int KW>
__device__ __forceinline__ void runs_flush(const ScanPlan& P, uint64_t (&regs)[MAX_ACC], int lane, const uint64_t (&ck)[KW], bool csf, bool any) {
  if (!any) return;
  uint64_t gs = ~0ull;
  if (lane == 0) gs = table_upsert<KW>(P.table, ck, hash_words<KW>(ck), csf || KW != 1);
#pragma unroll
  for (int a = 0; a < MAX_ACC; ++a) {
    if (a < CT::n_acc(P)) {
      const int op = CT::acc_op(P, a);
      uint64_t v = regs[a];
      regs[a] = acc_init(op);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v = acc_combine(op, v, __shfl_xor_sync(0xffffffffu, v, d));
      if (lane == 0 && gs != ~0ull && v != acc_init(op)) acc_apply_global(&tacc(P.table, a, gs), op, v);
    }
  }
}

template <class CT, int NC, int KW, int HF>
__device__ __forceinline__ void runs_slot(const ScanPlan& P, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC], int j, int64_t base, int lane,
                                          int rem, uint64_t (&regs)[MAX_ACC], uint64_t (&ck)[KW], bool& csf, bool& have, bool& dirty) {
  constexpr int NV = NVof<NC>::value;
  RowOut<KW, NV> o;
  row_front<CT, NC, KW, NV, HF>(P, raw, vbits, j, base, lane, rem, o);
  bool same = have && o.sentinel_free == csf;
#pragma unroll
  for (int w = 0; w < KW; ++w) same = same && o.k[w] == ck[w];
  const uint64_t grow = global_row<CT>(P, o.row);
  if (__all_sync(0xffffffffu, !o.alive || same)) {
    if (o.alive) {
      const RegSink s{regs};
      accumulate_row<CT, NV, KW>(P, o, grow, s);
      dirty = true;
    }
    return;
  }
  // a run ends inside this slot
  runs_flush<CT, KW>(P, regs, lane, ck, csf, __any_sync(0xffffffffu, dirty));
  dirty = false;
  if (o.alive) {
    const uint64_t gslot = table_upsert<KW>(P.table, o.k, hash_words<KW>(o.k), o.sentinel_free || KW != 1);
    if (gslot != ~0ull) {
      const ColdSink sink{P.table, gslot};
      accumulate_row<CT, NV, KW>(P, o, grow, sink);
    }
  }
  // the key of the last live row of the slot starts the new current run
  const uint32_t live = __ballot_sync(0xffffffffu, o.alive);
  if (live) {
    const int src = 31 - __clz((int)live);
#pragma unroll
    for (int w = 0; w < KW; ++w) ck[w] = __shfl_sync(0xffffffffu, o.k[w], src);
    csf = __shfl_sync(0xffffffffu, (int)o.sentinel_free, src) != 0;
    have = true;
  }
}

template <class CT, int NC, int KW>
__device__ __forceinline__ void runs_body(const ScanPlan& P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  uint64_t regs[MAX_ACC];
#pragma unroll
  for (int a = 0; a < MAX_ACC; ++a) regs[a] = a < CT::n_acc(P) ? acc_init(CT::acc_op(P, a)) : 0ull;
  const int64_t n_rows = P.n_rows;
  const int64_t n_steps = (n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  // every warp owns a CONTIGUOUS range of steps: runs stay inside one warp as long as possible
  const int64_t n_warps = (int64_t)gridDim.x * warps, w_id = (int64_t)blockIdx.x * warps + warp;
  const int64_t step_lo = n_steps * w_id / n_warps, step_hi = n_steps * (w_id + 1) / n_warps;
  uint64_t ck[KW];
#pragma unroll
  for (int w = 0; w < KW; ++w) ck[w] = 0ull;
  bool csf = true, have = false, dirty = false;
  uint4 rawA[2][NC], rawB[2][NC];
  uint32_t vbA[2][NC], vbB[2][NC];
  auto work = [&](int64_t step, const uint4 (&raw)[2][NC], const uint32_t (&vb)[2][NC]) {
    const int64_t base = step * ROWS_PER_STEP;
    const int rem = (int)(n_rows - base < ROWS_PER_STEP ? n_rows - base : ROWS_PER_STEP);
    runs_slot<CT, NC, KW, 0>(P, raw, vb, 0, base, lane, rem, regs, ck, csf, have, dirty);
    runs_slot<CT, NC, KW, 0>(P, raw, vb, 1, base, lane, rem, regs, ck, csf, have, dirty);
    runs_slot<CT, NC, KW, 1>(P, raw, vb, 0, base, lane, rem, regs, ck, csf, have, dirty);
    runs_slot<CT, NC, KW, 1>(P, raw, vb, 1, base, lane, rem, regs, ck, csf, have, dirty);
  };
  // two register buffers: the loads of the next step are in flight while this one is folded
  if (step_lo < step_hi) load_step<CT, NC>(P, step_lo * ROWS_PER_STEP, lane, n_rows, rawA, vbA);
  for (int64_t s = step_lo; s < step_hi; s += 2) {
    if (s + 1 < step_hi) load_step<CT, NC>(P, (s + 1) * ROWS_PER_STEP, lane, n_rows, rawB, vbB);
    work(s, rawA, vbA);
    if (s + 1 < step_hi) {
      if (s + 2 < step_hi) load_step<CT, NC>(P, (s + 2) * ROWS_PER_STEP, lane, n_rows, rawA, vbA);
      work(s + 1, rawB, vbB);
    }
  }
  runs_flush<CT, KW>(P, regs, lane, ck, csf, __any_sync(0xffffffffu, dirty));
}

}  // namespace pw
