#include "pw_launch.cuh"
namespace pw { int launch_scan_nc12_kw6(const ScanPlan& P, int sm, cudaStream_t st) { return launch_scan_nk<12, 6>(P, sm, st); } }
