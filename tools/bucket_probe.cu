// bucket_probe.cu — stand-alone prototype of the "bucket rows by dense id, fold in the owner's registers" scheme for the
// C2 shape (1e8 rows, int64 key in [0, 1000), f64 value -> sum / count / min / max per key), and a micro-probe of
// shared-memory atomics on this part.  Development tool: measures what the scheme can reach before it goes into
// pw_scan.cuh.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bucket_probe tools/bucket_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int GCAP = 1024;

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
__global__ void gen_kernel(int64_t* keys, double* vals, int64_t n, int groups) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t h = mix((uint64_t)i * 0x9E3779B97F4A7C15ull + 12345);
    keys[i] = (int64_t)(h % (uint64_t)groups);
    vals[i] = (double)(mix(h) >> 11) * (1.0 / 9007199254740992.0) * 1000.0 - 500.0;
  }
}

struct Part { double sum; unsigned long long cnt; double mn, mx; };

// reference: global atomics
__global__ void ref_kernel(const int64_t* keys, const double* vals, int64_t n, double* sum, unsigned long long* cnt, long long* mn, long long* mx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = keys[i];
    const double v = vals[i];
    atomicAdd(&sum[k], v);
    atomicAdd(&cnt[k], 1ull);
    long long o = __double_as_longlong(v);
    o ^= (o >> 63) & 0x7FFFFFFFFFFFFFFFll;   // ordered image
    atomicMin(&mn[k], o);
    atomicMax(&mx[k], o);
  }
}

// ---- streaming ceiling: same loads, no table ------------------------------------------------------------------------
template <int RPT>
__global__ void __launch_bounds__(1024, 1) stream_kernel(const int64_t* __restrict__ keys, const double* __restrict__ vals, int64_t n, double* out) {
  const int tid = threadIdx.x;
  const int64_t tile_rows = 1024 * RPT;
  const int64_t n_tiles = n / tile_rows;
  double s = 0.0;
  int64_t ks = 0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t base = t * tile_rows;
#pragma unroll
    for (int r = 0; r < RPT / 2; ++r) {
      const int64_t p = base + (int64_t)r * 2048 + tid * 2;
      const longlong2 k2 = *(const longlong2*)(keys + p);
      const double2 v2 = *(const double2*)(vals + p);
      ks += k2.x ^ k2.y;
      s += v2.x + v2.y;
    }
  }
  if (s == 1.2345 && ks == 77) out[0] = s;
}

// ---- the bucket scheme ------------------------------------------------------------------------------------------------
// THREADS threads per CTA, GCAP / THREADS groups per thread.  Tile = THREADS * RPT rows.  Per tile: every row takes a rank
// inside its id's bucket (shared atomic on cnt[id]) and stores its value at buf[rank][id]; after a barrier thread g folds
// bucket g into its registers (reads buf[j][g]: consecutive threads, consecutive words: conflict-free).  NBUF = 2: the
// scatter of tile t+1 may start before every thread has folded tile t (one barrier per tile); NBUF = 1: two barriers.
template <int THREADS, int J, int RPT, int NBUF, int CPS>
__global__ void __launch_bounds__(THREADS, CPS) bucket_kernel(const int64_t* __restrict__ keys, const double* __restrict__ vals, int64_t n,
                                                          int64_t kmin, Part* parts, double* ov_sum, unsigned long long* ov_cnt,
                                                          long long* ov_mn, long long* ov_mx, unsigned long long* ov_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int GPT = GCAP / THREADS;
  double* buf = (double*)smem_raw;                          // [NBUF][J][GCAP]
  uint32_t* cnt = (uint32_t*)(buf + NBUF * J * GCAP);       // [NBUF][GCAP]
  const int tid = threadIdx.x;
  for (int i = tid; i < NBUF * GCAP; i += THREADS) cnt[i] = 0;
  __syncthreads();
  constexpr int64_t tile_rows = THREADS * RPT;
  const int64_t n_tiles = n / tile_rows;   // full tiles; the tail goes through the overflow path below
  double sum[GPT], mn[GPT], mx[GPT];
  unsigned long long c_total[GPT];
#pragma unroll
  for (int g = 0; g < GPT; ++g) { sum[g] = 0.0; mn[g] = INFINITY; mx[g] = -INFINITY; c_total[g] = 0; }
  unsigned long long ov = 0;

  auto spill = [&](uint64_t id, double v) {
    atomicAdd(&ov_sum[id], v); atomicAdd(&ov_cnt[id], 1ull);
    long long o = __double_as_longlong(v); o ^= (o >> 63) & 0x7FFFFFFFFFFFFFFFll;
    atomicMin(&ov_mn[id], o); atomicMax(&ov_mx[id], o);
    ++ov;
  };

  longlong2 k2[RPT / 2], kn[RPT / 2];
  double2 v2[RPT / 2], vn[RPT / 2];
  int64_t t = blockIdx.x;
  if (t < n_tiles) {
#pragma unroll
    for (int r = 0; r < RPT / 2; ++r) {
      const int64_t p = t * tile_rows + (int64_t)r * 2 * THREADS + tid * 2;
      kn[r] = *(const longlong2*)(keys + p);
      vn[r] = *(const double2*)(vals + p);
    }
  }
  int b = 0;
  for (; t < n_tiles; t += gridDim.x) {
#pragma unroll
    for (int r = 0; r < RPT / 2; ++r) { k2[r] = kn[r]; v2[r] = vn[r]; }
    if (t + gridDim.x < n_tiles) {
#pragma unroll
      for (int r = 0; r < RPT / 2; ++r) {
        const int64_t p = (t + gridDim.x) * tile_rows + (int64_t)r * 2 * THREADS + tid * 2;
        kn[r] = *(const longlong2*)(keys + p);
        vn[r] = *(const double2*)(vals + p);
      }
    }
    double* bb = buf + (size_t)b * J * GCAP;
    uint32_t* cc = cnt + b * GCAP;
    uint64_t id[RPT];
    double v[RPT];
    uint32_t rk[RPT];
#pragma unroll
    for (int r = 0; r < RPT / 2; ++r) {
      id[2 * r] = (uint64_t)(k2[r].x - kmin); id[2 * r + 1] = (uint64_t)(k2[r].y - kmin);
      v[2 * r] = v2[r].x; v[2 * r + 1] = v2[r].y;
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) rk[r] = id[r] < (uint64_t)GCAP ? atomicAdd(&cc[id[r]], 1u) : 0xFFFFFFFFu;
    bool late = false;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (rk[r] < (uint32_t)J) bb[rk[r] * GCAP + (uint32_t)id[r]] = v[r];
      else late = true;
    }
    if (late) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) if (rk[r] >= (uint32_t)J && rk[r] != 0xFFFFFFFFu) spill(id[r], v[r]);
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < GPT; ++g) {
      const int gid = g * THREADS + tid;
      uint32_t c = cc[gid];
      if (c) cc[gid] = 0;
      c = c < (uint32_t)J ? c : (uint32_t)J;
      c_total[g] += c;
      const double* q = bb + gid;
      double s_ = sum[g], lo_ = mn[g], hi_ = mx[g];
#pragma unroll 1
      for (uint32_t j = 0; j < c; ++j, q += GCAP) {
        const double x = *q;
        s_ += x;
        lo_ = x < lo_ ? x : lo_;
        hi_ = x > hi_ ? x : hi_;
      }
      sum[g] = s_; mn[g] = lo_; mx[g] = hi_;
    }
    if (NBUF == 2) b ^= 1; else __syncthreads();
  }
  // tail rows (less than a tile): CTA 0, through the spill path
  if (blockIdx.x == 0) {
    for (int64_t p = n_tiles * tile_rows + tid; p < n; p += THREADS) {
      const uint64_t idt = (uint64_t)(keys[p] - kmin);
      if (idt < (uint64_t)GCAP) spill(idt, vals[p]);
    }
  }
#pragma unroll
  for (int g = 0; g < GPT; ++g) {
    Part p; p.sum = sum[g]; p.cnt = c_total[g]; p.mn = mn[g]; p.mx = mx[g];
    parts[(size_t)blockIdx.x * GCAP + g * THREADS + tid] = p;
  }
  if (ov) atomicAdd(ov_rows, ov);
}

// ---- shared-memory atomic micro-probe ---------------------------------------------------------------------------------
// MODE 0: atomicAdd u32 with return, 1: red (no return) u32, 2: atomicAdd u64 with return, 3: atomicAdd f64 (CAS loop?),
// 4: plain STS.32 for comparison.  Random cells in a CTA-shared table of 1024 cells.
template <int MODE>
__global__ void atoms_probe(unsigned long long* out) {
  __shared__ __align__(16) unsigned long long tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = 0;
  __syncthreads();
  uint32_t s = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x * 7919u;
  uint32_t acc = 0;
  const int ITERS = 2048;
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    s = s * 1664525u + 1013904223u;
    const uint32_t cell = (s >> 10) & 1023u;
    if (MODE == 0) acc += atomicAdd((uint32_t*)tab + cell, 1u);
    else if (MODE == 1) { const uint32_t a = (uint32_t)__cvta_generic_to_shared((uint32_t*)tab + cell); asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
    else if (MODE == 2) acc += (uint32_t)atomicAdd(tab + cell, 1ull);
    else if (MODE == 3) acc += (uint32_t)__double_as_longlong(atomicAdd((double*)tab + cell, 1.0));
    else { const uint32_t a = (uint32_t)__cvta_generic_to_shared((uint32_t*)tab + cell); asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(s) : "memory"); }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[1] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
}
template <int MODE>
void run_atoms(const char* name) {
  unsigned long long* d;
  CK(cudaMalloc(&d, 16));
  for (int warps : {8, 32}) {
    atoms_probe<MODE><<<148, warps * 32>>>(d);
    atoms_probe<MODE><<<148, warps * 32>>>(d);
    unsigned long long h[2];
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("%-34s warps/SM %2d: pipe cycles per warp instruction %6.2f\n", name, warps, (double)h[0] / 2048 / warps);
  }
  cudaFree(d);
}

template <int THREADS, int J, int RPT, int NBUF, int CPS>
void run_bucket(const int64_t* keys, const double* vals, int64_t n, int groups, const std::vector<double>& rsum, const std::vector<unsigned long long>& rcnt,
                const std::vector<long long>& rmn, const std::vector<long long>& rmx) {
  Part* parts;
  double* ov_sum; unsigned long long *ov_cnt, *ov_rows; long long *ov_mn, *ov_mx;
  const int grid = 148 * CPS;
  CK(cudaMalloc(&parts, sizeof(Part) * grid * GCAP));
  CK(cudaMalloc(&ov_sum, 8 * GCAP)); CK(cudaMalloc(&ov_cnt, 8 * GCAP)); CK(cudaMalloc(&ov_mn, 8 * GCAP)); CK(cudaMalloc(&ov_mx, 8 * GCAP)); CK(cudaMalloc(&ov_rows, 8));
  const size_t smem = (size_t)NBUF * J * GCAP * 8 + NBUF * GCAP * 4;
  CK(cudaFuncSetAttribute(bucket_kernel<THREADS, J, RPT, NBUF, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f, tot = 0;
  const int reps = 12;
  for (int it = 0; it < reps + 3; ++it) {
    CK(cudaMemset(ov_sum, 0, 8 * GCAP)); CK(cudaMemset(ov_cnt, 0, 8 * GCAP)); CK(cudaMemset(ov_rows, 0, 8));
    CK(cudaMemset(ov_mn, 0x7f, 8 * GCAP)); CK(cudaMemset(ov_mx, 0x80, 8 * GCAP));
    cudaEventRecord(e0);
    bucket_kernel<THREADS, J, RPT, NBUF, CPS><<<grid, THREADS, smem>>>(keys, vals, n, 0, parts, ov_sum, ov_cnt, ov_mn, ov_mx, ov_rows);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 3) { best = ms < best ? ms : best; tot += ms; }
  }
  std::vector<Part> hp((size_t)grid * GCAP);
  std::vector<double> os(GCAP); std::vector<unsigned long long> oc(GCAP); std::vector<long long> omn(GCAP), omx(GCAP);
  unsigned long long ovr = 0;
  CK(cudaMemcpy(hp.data(), parts, sizeof(Part) * grid * GCAP, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(os.data(), ov_sum, 8 * GCAP, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(oc.data(), ov_cnt, 8 * GCAP, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(omn.data(), ov_mn, 8 * GCAP, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(omx.data(), ov_mx, 8 * GCAP, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&ovr, ov_rows, 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int g = 0; g < groups; ++g) {
    double s = os[g]; unsigned long long c = oc[g];
    auto img = [](double v) { long long o; memcpy(&o, &v, 8); o ^= (o >> 63) & 0x7FFFFFFFFFFFFFFFll; return o; };
    long long mn = omn[g], mx = omx[g];
    for (int b = 0; b < grid; ++b) {
      const Part& p = hp[(size_t)b * GCAP + g];
      s += p.sum; c += p.cnt;
      if (p.cnt) { mn = std::min(mn, img(p.mn)); mx = std::max(mx, img(p.mx)); }
    }
    if (c != rcnt[g] || mn != rmn[g] || mx != rmx[g] || fabs(s - rsum[g]) > 1e-9 * (fabs(rsum[g]) + 1e3)) {
      if (bad < 5) printf("  MISMATCH g=%d cnt %llu/%llu sum %.6f/%.6f\n", g, c, rcnt[g], s, rsum[g]);
      ++bad;
    }
  }
  const double gb = (double)n * 16 / 1e9;
  printf("bucket T=%4d J=%2d RPT=%d NBUF=%d CTAs/SM=%d smem=%6zu B: best %.4f ms avg %.4f ms  -> %.0f GB/s (best), overflow rows %llu, %s\n", THREADS, J, RPT, NBUF, CPS, smem, best, tot / reps,
         gb / (best * 1e-3), ovr, bad ? "WRONG" : "results ok");
  cudaFree(parts); cudaFree(ov_sum); cudaFree(ov_cnt); cudaFree(ov_mn); cudaFree(ov_mx); cudaFree(ov_rows);
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 100000000ll;
  const int groups = argc > 2 ? atoi(argv[2]) : 1000;
  run_atoms<0>("ATOMS.ADD.32 return, random cells");
  run_atoms<1>("RED.shared.add.u32, random cells");
  run_atoms<2>("ATOMS.ADD.64 return, random cells");
  run_atoms<3>("atomicAdd f64 shared, random cells");
  run_atoms<4>("STS.32 random cells");
  CK(cudaDeviceSynchronize());

  int64_t* keys; double* vals;
  CK(cudaMalloc(&keys, n * 8)); CK(cudaMalloc(&vals, n * 8));
  gen_kernel<<<148 * 8, 256>>>(keys, vals, n, groups);
  CK(cudaDeviceSynchronize());
  double* rs; unsigned long long* rc; long long *rmn, *rmx;
  CK(cudaMalloc(&rs, 8 * GCAP)); CK(cudaMalloc(&rc, 8 * GCAP)); CK(cudaMalloc(&rmn, 8 * GCAP)); CK(cudaMalloc(&rmx, 8 * GCAP));
  CK(cudaMemset(rs, 0, 8 * GCAP)); CK(cudaMemset(rc, 0, 8 * GCAP)); CK(cudaMemset(rmn, 0x7f, 8 * GCAP)); CK(cudaMemset(rmx, 0x80, 8 * GCAP));
  ref_kernel<<<148 * 8, 256>>>(keys, vals, n, rs, rc, rmn, rmx);
  CK(cudaDeviceSynchronize());
  std::vector<double> hs(GCAP); std::vector<unsigned long long> hc(GCAP); std::vector<long long> hmn(GCAP), hmx(GCAP);
  CK(cudaMemcpy(hs.data(), rs, 8 * GCAP, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc.data(), rc, 8 * GCAP, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hmn.data(), rmn, 8 * GCAP, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hmx.data(), rmx, 8 * GCAP, cudaMemcpyDeviceToHost));

  // streaming ceiling
  {
    double* o; CK(cudaMalloc(&o, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 8; ++it) {
      cudaEventRecord(e0);
      stream_kernel<4><<<148, 1024>>>(keys, vals, n, o);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2) best = ms < best ? ms : best;
    }
    printf("stream (148 x 1024 threads, 64 B per thread per tile): best %.4f ms -> %.0f GB/s\n", best, (double)n * 16 / 1e9 / (best * 1e-3));
  }
  run_bucket<512, 10, 4, 1, 2>(keys, vals, n, groups, hs, hc, hmn, hmx);
  run_bucket<512, 7, 2, 1, 3>(keys, vals, n, groups, hs, hc, hmn, hmx);
  run_bucket<256, 7, 4, 1, 3>(keys, vals, n, groups, hs, hc, hmn, hmx);
  run_bucket<1024, 10, 2, 2, 1>(keys, vals, n, groups, hs, hc, hmn, hmx);
  run_bucket<1024, 13, 4, 2, 1>(keys, vals, n, groups, hs, hc, hmn, hmx);
  run_bucket<512, 9, 4, 1, 2>(keys, vals, n, groups, hs, hc, hmn, hmx);
  return 0;
}
