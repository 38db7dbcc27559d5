#include "pw_launch.cuh"
namespace pw { int launch_scan_nc4_kw6(const ScanPlan& P, int sm, cudaStream_t st) { return launch_scan_nk<4, 6>(P, sm, st); } }
