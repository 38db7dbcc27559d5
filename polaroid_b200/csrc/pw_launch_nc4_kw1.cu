#include "pw_launch.cuh"
namespace pw { int launch_scan_nc4_kw1(const ScanPlan& P, int sm, cudaStream_t st) { return launch_scan_nk<4, 1>(P, sm, st); } }
