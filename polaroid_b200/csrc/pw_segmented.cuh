// pw_segmented.cuh — the sorted-time fast path: tumbling windows over a sorted index column, no group keys.
//
// Because the index is sorted, the rows of a window are contiguous and a 128-row warp step almost always lies
// inside ONE window (1-minute bars over 1 ms ticks: ~60 000 rows per window).  Each lane keeps lane-private
// accumulator cells for the warp's current window in shared memory ([word][lane]: conflict-free, no claims, no
// atomics, no hashing); only when a row slot crosses a window boundary does the warp reduce its cells with
// shuffles and push them into the dense window table with one atomic per word.  The window index of a row is
// found with two compares against the current window's bounds; the 64-bit division runs only at boundaries.
// Reference: the sequential sweep of polars-time/src/windows/group_by.rs:79-151 + slice aggregation
// (polars-core/src/frame/group_by/aggregations/mod.rs:184-191, float_sum.rs blocked sums).
#pragma once
#include "pw_scan.cuh"

namespace pw {

struct SegParams {
  int64_t k0;       // window index of dense slot 0
  int64_t n_dense;  // dense slots
};

struct CellSink {  // lane-private cells of the current window
  uint64_t* cells;
  int lane;
  template <int OP>
  __device__ __forceinline__ void add(const ScanPlan&, int a, uint64_t x) const {
    uint64_t* q = cells + a * 32 + lane;
    *q = acc_combine(OP, *q, x);
  }
};
// The same cells in REGISTERS: under the query-shape specialised build every accumulator index is a compile-time
// constant after unrolling, so the array never leaves the register file and a row costs one ALU op per word instead of a
// shared-memory load and store (the ahead-of-time kernels index at run time and keep the shared-memory cells).
struct RegSink {
  uint64_t (&acc)[MAX_ACC];
  template <int OP>
  __device__ __forceinline__ void add(const ScanPlan&, int a, uint64_t x) const {
#pragma unroll
    for (int i = 0; i < MAX_ACC; ++i)
      if (i == a) acc[i] = acc_combine(OP, acc[i], x);
  }
};
struct DenseSink {  // straight into the dense window table (boundary rows)
  const Table& T;
  uint64_t slot;
  template <int OP>
  __device__ __forceinline__ void add(const ScanPlan&, int a, uint64_t x) const { acc_apply_global(&tacc(T, a, slot), OP, x); }
};

template <class CT>
__device__ __forceinline__ void seg_flush(const ScanPlan& P, const SegParams& sp, uint64_t* cells, uint64_t (&regs)[MAX_ACC], int lane, int64_t k, bool any) {
  if (!any) return;
  // The host sizes the dense table from the first and last index value only.  A window outside it can only come from
  // an unsorted index (an interior value beyond both ends): report it instead of writing out of bounds.
  const bool inside = (uint64_t)(k - sp.k0) < (uint64_t)sp.n_dense;
  if (!inside && lane == 0) *P.not_sorted = 1;
  if constexpr (CT::kJit) {
#pragma unroll
    for (int a = 0; a < MAX_ACC; ++a) {
      if (a < CT::n_acc(P)) {
        const int op = CT::acc_op(P, a);
        uint64_t v = regs[a];
        regs[a] = acc_init(op);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = acc_combine(op, v, __shfl_xor_sync(0xffffffffu, v, d));
        if (lane == 0 && inside && v != acc_init(op)) acc_apply_global(&tacc(P.table, a, (uint64_t)(k - sp.k0)), op, v);
      }
    }
  } else {
    for (int a = 0; a < CT::n_acc(P); ++a) {
      const int op = CT::acc_op(P, a);
      uint64_t v = cells[a * 32 + lane];
      cells[a * 32 + lane] = acc_init(op);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v = acc_combine(op, v, __shfl_xor_sync(0xffffffffu, v, d));
      if (lane == 0 && inside && v != acc_init(op)) acc_apply_global(&tacc(P.table, a, (uint64_t)(k - sp.k0)), op, v);
    }
  }
}

__device__ __forceinline__ bool in_window(int closed, int64_t t, int64_t s, int64_t e) {
  switch (closed) {
    case 0: return t >= s && t < e;
    case 1: return t > s && t <= e;
    case 3: return t > s && t < e;
    default: return t >= s && t <= e;
  }
}

template <class CT, int NC, int HF>
__device__ __forceinline__ void seg_row(const ScanPlan& P, const SegParams& sp, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC], int j,
                                        int64_t base, int lane, int64_t n_rows, uint64_t* cells, uint64_t (&regs)[MAX_ACC], int64_t& cur_k,
                                        int64_t& cur_s, int64_t& cur_e, bool& dirty);

template <class CT, int NC, int HF>
__device__ __forceinline__ void seg_rows(const ScanPlan& P, const SegParams& sp, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC],
                                         int64_t base, int lane, int64_t n_rows, uint64_t* cells, uint64_t (&regs)[MAX_ACC], int64_t& cur_k,
                                         int64_t& cur_s, int64_t& cur_e, bool& dirty) {
  if constexpr (CT::kJit) {  // specialised build: both pair elements inline (compile-time j: no selects on the raw words)
    seg_row<CT, NC, HF>(P, sp, raw, vbits, 0, base, lane, n_rows, cells, regs, cur_k, cur_s, cur_e, dirty);
    seg_row<CT, NC, HF>(P, sp, raw, vbits, 1, base, lane, n_rows, cells, regs, cur_k, cur_s, cur_e, dirty);
  } else {
#pragma unroll 1
    for (int j = 0; j < 2; ++j) seg_row<CT, NC, HF>(P, sp, raw, vbits, j, base, lane, n_rows, cells, regs, cur_k, cur_s, cur_e, dirty);
  }
}

template <class CT, int NC, int HF>
__device__ __forceinline__ void seg_row(const ScanPlan& P, const SegParams& sp, const uint4 (&raw)[2][NC], const uint32_t (&vbits)[2][NC], int j,
                                        int64_t base, int lane, int64_t n_rows, uint64_t* cells, uint64_t (&regs)[MAX_ACC], int64_t& cur_k,
                                        int64_t& cur_s, int64_t& cur_e, bool& dirty) {
  constexpr int NV = NVof<NC>::value;
  {
    const int64_t row = base + HF * 64 + 2 * lane + j;
    Row<NC> r;
    row_decode<CT, NC>(P, raw[HF], vbits[HF], j, r);
    bool alive = row < n_rows && row_predicate<CT, NC>(P, r);
    const int64_t t = (int64_t)pick<NC>(r.in, CT::dyn_slot(P));
    int64_t k = cur_k;
    if (alive && !in_window(CT::dyn_closed(P), t, cur_s, cur_e)) alive = window_of_ct<CT>(P, t, k);
    RowOut<1, NV> o;
    o.k[0] = 0; o.alive = alive; o.sentinel_free = true; o.row = row; o.tval = (uint64_t)t;
    row_vexprs<CT, NC, NV>(P, r, o.v, o.v_valid);
    const uint64_t grow = (uint64_t)(row + P.row_offset);
    if (__all_sync(0xffffffffu, !alive || k == cur_k)) {
      if (alive) {
        if constexpr (CT::kJit) { const RegSink s{regs}; accumulate_row<CT, NV, 1>(P, o, grow, s); }
        else { const CellSink s{cells, lane}; accumulate_row<CT, NV, 1>(P, o, grow, s); }
        dirty = true;
      }
    } else {
      // a window boundary inside this row slot: publish the current window, send these rows straight to the table,
      // continue with the latest window seen
      seg_flush<CT>(P, sp, cells, regs, lane, cur_k, __any_sync(0xffffffffu, dirty));
      dirty = false;
      if (alive) {
        if ((uint64_t)(k - sp.k0) < (uint64_t)sp.n_dense) { const DenseSink s{P.table, (uint64_t)(k - sp.k0)}; accumulate_row<CT, NV, 1>(P, o, grow, s); }
        else *P.not_sorted = 1;  // outside [first, last]: the index is not sorted (see seg_flush)
      }
      const int64_t kmax = alive ? k : INT64_MIN;
      int64_t km = kmax;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) { const int64_t o2 = __shfl_xor_sync(0xffffffffu, km, d); km = o2 > km ? o2 : km; }
      cur_k = km;
      cur_s = P.dyn.origin + km * P.dyn.every;
      cur_e = cur_s + P.dyn.period;
    }
  }
}

template <class CT, int NC>
__device__ __forceinline__ void seg_body(const ScanPlan& P, const SegParams& sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  uint64_t* cells = (uint64_t*)smem_raw + (size_t)warp * CT::n_acc(P) * 32;
  for (int a = 0; a < CT::n_acc(P); ++a) cells[a * 32 + lane] = acc_init(CT::acc_op(P, a));
  uint64_t regs[MAX_ACC];
#pragma unroll
  for (int a = 0; a < MAX_ACC; ++a) regs[a] = a < CT::n_acc(P) ? acc_init(CT::acc_op(P, a)) : 0ull;
  __syncwarp();
  const int64_t n_rows = P.n_rows;
  const int64_t n_steps = (n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  const int64_t n_tiles = (n_steps + warps - 1) / warps;
  const int64_t tile_lo = n_tiles * blockIdx.x / gridDim.x;
  const int64_t tile_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;
  int64_t cur_k = sp.k0 - 1, cur_s = INT64_MAX, cur_e = INT64_MIN;  // an empty current window
  bool dirty = false;
  for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
    const int64_t step = tile * warps + warp;
    if (step >= n_steps) break;
    const int64_t base = step * ROWS_PER_STEP;
    uint4 raw[2][NC];
    uint32_t vbits[2][NC];
    load_step<CT, NC>(P, base, lane, n_rows, raw, vbits);
    check_sorted_step<CT, NC>(P, raw, base, lane, n_rows);
    seg_rows<CT, NC, 0>(P, sp, raw, vbits, base, lane, n_rows, cells, regs, cur_k, cur_s, cur_e, dirty);
    seg_rows<CT, NC, 1>(P, sp, raw, vbits, base, lane, n_rows, cells, regs, cur_k, cur_s, cur_e, dirty);
  }
  seg_flush<CT>(P, sp, cells, regs, lane, cur_k, __any_sync(0xffffffffu, dirty));
}

#ifndef __CUDACC_RTC__
template <int NC>
__global__ void __launch_bounds__(256) seg_kernel(const __grid_constant__ ScanPlan P, const SegParams sp) {
  seg_body<RtCtl, NC>(P, sp);
}
#endif

}  // namespace pw
