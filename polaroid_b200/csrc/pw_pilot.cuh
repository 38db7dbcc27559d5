// pw_pilot.cuh — the key-sample pilot as ONE launch.
//
// Strategy selection needs three numbers (pw_engine.cu, run_groupby): distinct keys in a strided sample of the whole
// input (sizes the HBM table), distinct keys in a contiguous block (are consecutive rows sharing few groups?), and —
// for a single integer key — the value range of both samples (dense ids).  The first version ran two scans through the
// regular path plus two range kernels: ~20 host calls and 8 launches whose launch latencies, not their work, cost
// ~0.1 ms per query.  Here both samples are read by one kernel; distinct keys are counted at insertion time.
#pragma once
#include "pw_scan.cuh"

namespace pw {

struct PilotParams {
  int64_t begin[2], stride[2], n[2];   // sample s reads rows begin + i * stride, i < n
  Table table[2];                      // one scratch table per sample (keys preset to KEY_EMPTY, state to 0)
  unsigned long long* distinct[2];
  unsigned long long *kmax_u, *kmin_n; // key range as maxima of biased images (see Control), or nullptr
};

// find-or-insert that reports a fresh insertion; same slot protocol as table_upsert
template <int KW>
__device__ __forceinline__ bool pilot_insert(const Table& T, const uint64_t (&k)[KW], uint64_t h, bool sentinel_free) {
  const uint64_t cap = T.cap;
  if (KW == 1) {
    const uint64_t k0 = k[0];
    if (!sentinel_free) return atomicExch(&T.state[k0 == KEY_EMPTY ? cap : cap + 1], 2u) != 2u;
    uint64_t slot = __umul64hi(h, cap);
    for (uint64_t probes = 0; probes < cap; ++probes) {
      unsigned long long old = __ldcg((const unsigned long long*)&tkey(T, 0, slot));
      if (old == k0) return false;
      if (old == KEY_EMPTY) {
        old = atomicCAS((unsigned long long*)&tkey(T, 0, slot), (unsigned long long)KEY_EMPTY, (unsigned long long)k0);
        if (old == KEY_EMPTY) return true;
        if (old == k0) return false;
      }
      slot = (slot + 1 == cap) ? 0 : slot + 1;
    }
    return false;
  } else {
    uint64_t slot = __umul64hi(h, cap);
    uint64_t probes = 0;
    for (;;) {
      uint32_t s = ld_volatile_u32(&T.state[slot]);
      if (s == 0u) s = atomicCAS(&T.state[slot], 0u, 1u) == 0u ? 3u : 1u;
      if (s == 3u) {
#pragma unroll
        for (int w = 0; w < KW; ++w) tkey(T, w, slot) = k[w];
        __threadfence();
        st_volatile_u32(&T.state[slot], 2u);
        return true;
      }
      if (s == 2u) {
        __threadfence();
        bool eq = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) eq &= (__ldcg((const unsigned long long*)&tkey(T, w, slot)) == k[w]);
        if (eq) return false;
        slot = (slot + 1 == cap) ? 0 : slot + 1;
        if (++probes >= cap) return false;
      }
      // s == 1: being written -> look again
    }
  }
}

// blockIdx.y = sample.  Rows are read one per thread with the guarded scalar loads (the strided sample has no two
// neighbouring rows anyway).
template <class CT, int NC, int KW>
__device__ __forceinline__ void pilot_body(const ScanPlan& P, const PilotParams& pp) {
  const int s = blockIdx.y;
  const int64_t n = pp.n[s], rb = pp.begin[s], rs = pp.stride[s];
  unsigned long long hi = 0, lo = 0, fresh = 0;
  for (int64_t p = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); p < n; p += 2 * (int64_t)gridDim.x * blockDim.x) {
    uint4 raw[NC];
    uint32_t vbits[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < CT::n_slots(P)) {
        raw[c] = CT::slot_dtype(P, c) == DT_BOOL ? load_bool_pair(P.slots[c], p, n, rb, rs) : load_pair(P.slots[c].values, CT::slot_dtype(P, c), p, n, false, rb, rs);
        vbits[c] = CT::slot_nullable(P, c) ? load_valid_pair(P.slots[c], p, n, rb, rs) : 3u;
      } else { raw[c] = make_uint4(0u, 0u, 0u, 0u); vbits[c] = 0u; }
    }
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      if (p + j >= n) break;
      Row<NC> r;
      row_decode<CT, NC>(P, raw, vbits, j, r);
      uint64_t k[KW];
      bool sentinel_free;
      if (!row_keys<CT, NC, KW>(P, r, raw, vbits, j, true, k, sentinel_free)) continue;  // row outside every window
      if (pilot_insert<KW>(pp.table[s], k, hash_words<KW>(k), sentinel_free || KW != 1)) ++fresh;
      // single integer key, non-null (raw -1 / -2 included); under group_by_dynamic k[0] is that key and k[1] the window
      if (pp.kmax_u && (KW == 1 ? !(sentinel_free && k[0] == KEY_NULL) : CT::dyn_enabled(P))) {
        const unsigned long long u = k[0] ^ 0x8000000000000000ull;
        hi = u > hi ? u : hi;
        lo = ~u > lo ? ~u : lo;
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
    const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, o), l2 = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = h2 > hi ? h2 : hi;
    lo = l2 > lo ? l2 : lo;
  }
  if ((threadIdx.x & 31) == 0) {
    if (fresh) atomicAdd(pp.distinct[s], fresh);
    if (pp.kmax_u && (hi | lo)) { atomicMax(pp.kmax_u, hi); atomicMax(pp.kmin_n, lo); }
  }
}

}  // namespace pw
