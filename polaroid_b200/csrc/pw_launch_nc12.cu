#include "pw_launch.cuh"
namespace pw { int launch_scan_nc12(const ScanPlan& P, int sm, cudaStream_t st) { return launch_scan_n<12>(P, sm, st); } }
