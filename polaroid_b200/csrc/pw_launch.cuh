// pw_launch.cuh — scan-kernel instantiation + launch for one raw-slot class (compiled in its own TU so the
// template instantiations build in parallel)
#pragma once
#include "pw_engine.h"
#include "pw_scan.cuh"

namespace pw {

template <int NC, int KW, bool HOT>
static int launch_scan_t(const ScanPlan& P, int sm_count, cudaStream_t st) {
  auto kern = scan_kernel<NC, KW, HOT>;
  size_t smem = HOT ? (size_t)P.hot.total_bytes : 0;
  if (smem > 48 * 1024) PW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  constexpr int THREADS = ScanCfg<NC>::THREADS;
  PW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
  if (per_sm < 1) return fail(PW_ERR_CUDA, "scan kernel does not fit on an SM (smem %zu)", smem);
  const int64_t n_steps = (P.n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  const int64_t n_tiles = (n_steps + (THREADS / 32) - 1) / (THREADS / 32);
  int64_t grid = (int64_t)sm_count * per_sm;  // one wave of resident CTAs, each owning a contiguous row range
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, THREADS, smem, st>>>(P);
  PW_CUDA(cudaGetLastError());
  ctx().timings.kernel_launches++;
  return 0;
}

template <int NC, int KW>
static int launch_scan_nk(const ScanPlan& P, int sm, cudaStream_t st) {
  return P.hot_slots > 0 ? launch_scan_t<NC, KW, true>(P, sm, st) : launch_scan_t<NC, KW, false>(P, sm, st);
}
}  // namespace pw
