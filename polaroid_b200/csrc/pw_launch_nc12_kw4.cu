#include "pw_launch.cuh"
namespace pw { int launch_scan_nc12_kw4(const ScanPlan& P, int sm, cudaStream_t st) { return launch_scan_nk<12, 4>(P, sm, st); } }
