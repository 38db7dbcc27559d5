// dispatch over the key-word classes of the 4-slot scan kernels (one translation unit per class: parallel builds)
#include "pw_engine.h"
namespace pw {
int launch_scan_nc4_kw1(const ScanPlan& P, int sm, cudaStream_t st);
int launch_scan_nc4_kw2(const ScanPlan& P, int sm, cudaStream_t st);
int launch_scan_nc4_kw4(const ScanPlan& P, int sm, cudaStream_t st);
int launch_scan_nc4_kw6(const ScanPlan& P, int sm, cudaStream_t st);
int launch_scan_nc4(const ScanPlan& P, int sm, cudaStream_t st) {
  if (P.n_kw <= 1) return launch_scan_nc4_kw1(P, sm, st);
  if (P.n_kw <= 2) return launch_scan_nc4_kw2(P, sm, st);
  if (P.n_kw <= 4) return launch_scan_nc4_kw4(P, sm, st);
  return launch_scan_nc4_kw6(P, sm, st);
}
}  // namespace pw
