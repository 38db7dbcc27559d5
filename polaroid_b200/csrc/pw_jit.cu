// pw_jit.cu — query-shape specialisation of the scan kernel with NVRTC.
//
// The ahead-of-time kernels interpret the query plan at run time (RtCtl).  The first time a query SHAPE
// (dtypes, key layout, aggregate flags, hot-table geometry — no pointers, no row counts) is seen, this
// module generates a `JitCtl` policy whose accessors are constexpr, compiles `scan_body<JitCtl,...>` from
// the very same pw_scan.cuh for sm_100a, and caches the CUfunction for the life of the process.  NVRTC and
// the driver API are dlopen'ed lazily so the library still loads on a machine without a GPU; when either is
// missing the engine keeps using the AOT kernels (still on the GPU — there is no CPU path).
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "pw_engine.h"
#include "pw_partition.cuh"
#include "pw_overlap.cuh"
#include "pw_pilot.cuh"
#include "pw_radix.cuh"
#include "pw_scan.cuh"

namespace pw {
namespace {

// ---- minimal NVRTC / driver API surface (resolved with dlsym) -------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;
typedef int CUresult;
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
typedef struct CUstream_st* CUstream;

struct Api {
  bool ok = false;
  std::string why;
  nvrtcResult (*nvrtcCreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*nvrtcCompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*nvrtcGetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*nvrtcGetProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*nvrtcGetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*nvrtcGetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*nvrtcDestroyProgram)(nvrtcProgram*);
  CUresult (*cuModuleLoadData)(CUmodule*, const void*);
  CUresult (*cuModuleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*cuFuncSetAttribute)(CUfunction, int, int);
  CUresult (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t);
  CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* rtc = nullptr;
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"})
      if ((rtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!rtc) { a.why = "libnvrtc not found"; return; }
    void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!drv) { a.why = "libcuda.so.1 not found"; return; }
#define PW_SYM(lib, name)                                                        \
  *(void**)(&a.name) = dlsym(lib, #name);                                        \
  if (!a.name) { a.why = std::string("missing symbol ") + #name; return; }
    PW_SYM(rtc, nvrtcCreateProgram) PW_SYM(rtc, nvrtcCompileProgram) PW_SYM(rtc, nvrtcGetProgramLogSize)
    PW_SYM(rtc, nvrtcGetProgramLog) PW_SYM(rtc, nvrtcGetCUBINSize) PW_SYM(rtc, nvrtcGetCUBIN) PW_SYM(rtc, nvrtcDestroyProgram)
    PW_SYM(drv, cuModuleLoadData) PW_SYM(drv, cuModuleGetFunction) PW_SYM(drv, cuFuncSetAttribute)
    PW_SYM(drv, cuOccupancyMaxActiveBlocksPerMultiprocessor) PW_SYM(drv, cuLaunchKernel)
#undef PW_SYM
    a.ok = true;
  });
  return a;
}

std::string csrc_dir() {
  if (const char* e = getenv("PW_CSRC_DIR")) return e;
  Dl_info info;
  if (dladdr((void*)&csrc_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;  // .../polaroid_b200/lib/libpolarway_b200.so
    size_t k = p.rfind('/');
    if (k != std::string::npos) p = p.substr(0, k);
    k = p.rfind('/');
    if (k != std::string::npos) return p.substr(0, k) + "/csrc";
  }
  return "polaroid_b200/csrc";
}

// ---- JitCtl generation -------------------------------------------------------------------------------------
struct Gen {
  std::ostringstream o;
  void scalar(const char* type, const char* name, long long v) {
    o << "  static __device__ __forceinline__ constexpr " << type << " " << name << "(const ScanPlan&) { return " << v << "; }\n";
  }
  template <class F>
  void table1(const char* type, const char* name, int n, F f) {
    o << "  static __device__ __forceinline__ constexpr " << type << " " << name << "(const ScanPlan&, int i) { return ";
    for (int i = 0; i < n; ++i) o << "i == " << i << " ? " << f(i) << " : ";
    o << "0; }\n";
  }
};

std::string hexdouble(double d) {
  char buf[64];
  snprintf(buf, sizeof buf, "%a", d);
  return buf;
}

std::string jit_ctl(const ScanPlan& P) {
  Gen g;
  g.o << "struct JitCtl {\n  static constexpr bool kJit = true;\n";
  g.scalar("int", "n_slots", P.n_slots);
  g.table1("int", "slot_dtype", P.n_slots, [&](int i) { return P.slots[i].dtype; });
  g.table1("bool", "slot_nullable", P.n_slots, [&](int i) { return P.slots[i].validity != nullptr ? 1 : 0; });
  g.scalar("int", "n_preds", P.n_preds);
  g.table1("int", "pred_slot", P.n_preds, [&](int i) { return P.preds[i].slot; });
  g.table1("int", "pred_op", P.n_preds, [&](int i) { return P.preds[i].op; });
  g.table1("int", "pred_cls", P.n_preds, [&](int i) { return P.preds[i].cls; });
  g.scalar("int", "n_keys", P.n_keys);
  g.table1("int", "key_slot", P.n_keys, [&](int i) { return P.keys[i].slot; });
  g.table1("int", "key_dtype", P.n_keys, [&](int i) { return P.keys[i].dtype; });
  g.table1("int", "key_words", P.n_keys, [&](int i) { return P.keys[i].n_words; });
  g.scalar("bool", "has_null_word", P.has_null_word);
  g.scalar("bool", "dyn_enabled", P.dyn.enabled);
  g.scalar("int", "dyn_slot", P.dyn.slot);
  g.scalar("int", "dyn_closed", P.dyn.closed);
  g.o << "  static constexpr long long kDynEvery = " << (P.dyn.enabled && P.dyn.every > 0 ? (long long)P.dyn.every : 0ll) << "ll, kDynPeriod = "
      << (long long)P.dyn.period << "ll;\n";
  g.scalar("int", "n_vexpr", P.n_vexpr);
  g.table1("int", "ve_nf", P.n_vexpr, [&](int i) { return P.vexprs[i].n_factors; });
  g.table1("int", "ve_slot", P.n_vexpr, [&](int i) { return P.vexprs[i].slot; });
  g.table1("int", "ve_cls", P.n_vexpr, [&](int i) { return P.vexprs[i].cls; });
  g.table1("int", "ve_flags", P.n_vexpr, [&](int i) { return P.vexprs[i].flags; });
  g.table1("int", "ve_acc", P.n_vexpr, [&](int i) { return P.vexprs[i].acc_base; });
  auto table2 = [&](const char* type, const char* name, auto f) {
    g.o << "  static __device__ __forceinline__ constexpr " << type << " " << name << "(const ScanPlan&, int e, int f) { return ";
    for (int e = 0; e < P.n_vexpr; ++e)
      for (int k = 0; k < P.vexprs[e].n_factors; ++k) g.o << "(e == " << e << " && f == " << k << ") ? " << f(e, k) << " : ";
    g.o << "0; }\n";
  };
  table2("int", "fac_slot", [&](int e, int k) { return std::to_string(P.vexprs[e].f[k].slot); });
  table2("double", "fac_a", [&](int e, int k) { return hexdouble(P.vexprs[e].f[k].a); });
  table2("double", "fac_b", [&](int e, int k) { return hexdouble(P.vexprs[e].f[k].b); });
  g.scalar("int", "gflags", P.gflags);
  g.scalar("int", "acc_gbase", P.acc_gbase);
  g.scalar("int", "n_acc", P.n_acc);
  g.table1("int", "acc_op", P.n_acc, [&](int i) { return P.accs[i].op; });
  g.scalar("bool", "vec_ok", P.vec_ok);
  g.scalar("bool", "check_sorted", P.check_sorted);
  g.scalar("bool", "unit_stride", P.row_begin == 0 && P.row_stride == 1);
  g.scalar("int", "h_slots", P.hot.idx_slots);
  g.scalar("int", "h_gcap", P.hot.gcap);
  g.scalar("int", "h_rep", P.hot.replicas > 0 ? P.hot.replicas : 1);
  g.scalar("int", "h_keys_off", P.hot.keys_off);
  g.scalar("int", "h_mm_off", P.hot.mm_off);
  g.scalar("int", "h_count_off", P.hot.count_off);
  g.scalar("int", "h_warp_off", P.hot.warp_off);
  g.scalar("int", "h_warp_bytes", P.hot.warp_bytes);
  g.scalar("int", "h_claim_off", P.hot.claim_off);
  g.scalar("int", "h_claim_acc", P.hot.claim_acc);
  g.scalar("int", "h_n_mm", P.hot.n_mm);
  g.scalar("int", "rowid_slot", P.rowid_slot_p1 - 1);
  g.scalar("bool", "h_dense", P.hot.dense != 0);
  g.scalar("bool", "h_dense_sentinels", P.hot.dense == 2);
  g.scalar("int", "h_guard_acc", P.hot.guard_acc);
  g.scalar("int", "h_shadow_off", P.hot.shadow_off);
  g.scalar("int", "h_mm_stride", P.hot.mm_stride > 0 ? P.hot.mm_stride : 1);
  g.scalar("bool", "group_out", P.row_group_out != nullptr);
  g.table1("int", "h_kind", P.n_acc, [&](int i) { return P.hot.acc_kind[i]; });
  g.table1("int", "h_off", P.n_acc, [&](int i) { return P.hot.acc_off[i]; });
  // array extents and geometry of the bucket tier (pw_bucket.cuh) have to be constant expressions of their own
  g.o << "  static constexpr int kNAcc = " << (P.n_acc > 0 ? P.n_acc : 1) << ", kNVexpr = " << P.n_vexpr
      << ", kLenAcc = " << ((P.gflags & GF_LEN) ? P.acc_gbase : -1) << ";\n";
  g.o << "  static constexpr int kBThreads = " << (P.hot.b_threads > 0 ? P.hot.b_threads : 1024) << ", kBGcap = " << (P.hot.b_gcap > 0 ? P.hot.b_gcap : 1024)
      << ", kBJ = " << (P.hot.b_j > 0 ? P.hot.b_j : 1) << ", kBNbuf = " << (P.hot.b_nbuf > 0 ? P.hot.b_nbuf : 1)
      << ", kBStages = " << P.hot.b_stages << ";\n";
  g.o << "  static constexpr int kBRange = " << P.hot.b_range << ";\n  static constexpr bool kBSent = " << (P.hot.b_sent ? "true" : "false") << ";\n";
  {
    uint32_t firstw = 0, lastw = 0, tagw = 0;
    for (int a = 0; a < P.n_acc && P.hot.b_rowpos; ++a) {
      const int src = P.accs[a].src, op = P.accs[a].op;
      if (src != SRC_ROWIDX && src != SRC_ROWIDX_NN && src != SRC_ROW) continue;
      if (op == OP_MIN_U64) firstw |= 1u << a; else lastw |= 1u << a;
      if (src != SRC_ROW) tagw |= 1u << a;
    }
    g.o << "  static constexpr bool kBRowPos = " << (P.hot.b_rowpos ? "true" : "false") << ";\n  static constexpr unsigned kBFirstAcc = " << firstw
        << "u, kBLastAcc = " << lastw << "u, kBRowTagAcc = " << tagw << "u;\n";
  }
  g.o << "  static constexpr int kBIdxMul = " << (P.hot.b_idx_mul > 0 ? P.hot.b_idx_mul : 8) << ";\n";
  g.o << "  static constexpr bool kBIdx = " << (P.hot.b_idx ? "true" : "false") << ";\n";
  g.o << "  static constexpr bool kBWin = " << (P.hot.b_win ? "true" : "false") << ";\n";
  g.o << "  static constexpr unsigned kBVar = " << (getenv("PW_BUCKET_VAR") ? atoi(getenv("PW_BUCKET_VAR")) : 0) << "u;\n";
  g.o << "  static constexpr bool kBMeta = " << (P.hot.b_meta ? "true" : "false") << ";\n";
  {
    // f64 min / max of a value that cannot be null: kept as plain doubles by the bucket tier (rows holding NaN or -0.0
    // take the HBM path).  Accumulator order inside an expression: sum_i, sum_f, count, min, max (lower_query).
    uint32_t native = 0, odd = 0;
    for (int e = 0; e < P.n_vexpr && P.hot.bucket; ++e) {
      const VExpr& V = P.vexprs[e];
      if (V.cls != CLS_F64 || !(V.flags & (VF_MIN | VF_MAX))) continue;
      bool nullable = false;
      if (V.n_factors == 0) nullable = P.slots[V.slot].validity != nullptr;
      else for (int k = 0; k < V.n_factors; ++k) nullable = nullable || P.slots[V.f[k].slot].validity != nullptr;
      if (nullable || getenv("PW_NO_NATIVE_MINMAX")) continue;
      int a = V.acc_base + ((V.flags & VF_SUM_I) ? 1 : 0) + ((V.flags & VF_SUM_F) ? 1 : 0) + ((V.flags & VF_COUNT) ? 1 : 0);
      if (V.flags & VF_MIN) native |= 1u << a++;
      if (V.flags & VF_MAX) native |= 1u << a++;
      odd |= 1u << e;
    }
    g.o << "  static constexpr unsigned kBNativeAcc = " << native << "u, kBOddVexpr = " << odd << "u;\n";
  }
  g.o << "};\n";
  return g.o.str();
}

struct Compiled { CUfunction fn = nullptr; bool failed = false; int per_sm = 0; };
std::mutex g_mu;
std::map<std::string, Compiled> g_cache;

std::string scan_entry(int nc, int kw, bool hot, int threads, int min_blocks = 1) {
  std::ostringstream src;
  src << "extern \"C\" __global__ void __launch_bounds__(" << threads << ", " << min_blocks << ") pw_scan_jit(const __grid_constant__ pw::ScanPlan P) {\n"
      << "  pw::scan_body<pw::JitCtl, " << nc << ", " << kw << ", " << (hot ? "true" : "false") << ">(P);\n}\n";
  return src.str();
}
std::string bucket_entry(int nc, int kw, int threads, int cps) {
  std::ostringstream src;
  src << "extern \"C\" __global__ void __launch_bounds__(" << threads << ", " << (cps > 0 ? cps : 1) << ") pw_bucket_jit(const __grid_constant__ pw::ScanPlan P) {\n"
      << "  pw::bucket_body<pw::JitCtl, " << nc << ", " << kw << ">(P);\n}\n";
  return src.str();
}
std::string runs_entry(int nc, int kw, int threads) {
  std::ostringstream src;
  src << "extern \"C\" __global__ void __launch_bounds__(" << threads << ") pw_runs_jit(const __grid_constant__ pw::ScanPlan P) {\n"
      << "  pw::runs_body<pw::JitCtl, " << nc << ", " << kw << ">(P);\n}\n";
  return src.str();
}
std::string seg_entry(int nc, int threads) {
  std::ostringstream src;
  src << "extern \"C\" __global__ void __launch_bounds__(" << threads << ") pw_seg_jit(const __grid_constant__ pw::ScanPlan P, const pw::SegParams sp) {\n"
      << "  pw::seg_body<pw::JitCtl, " << nc << ">(P, sp);\n}\n";
  return src.str();
}

std::string part_entry(int nc, int kw, int threads) {
  std::ostringstream src;
  src << "extern \"C\" __global__ void __launch_bounds__(" << threads << ") pw_part_jit(const __grid_constant__ pw::ScanPlan P, const pw::PartParams pp) {\n"
      << "  pw::part_body<pw::JitCtl, " << nc << ", " << kw << ">(P, pp);\n}\n";
  return src.str();
}

std::string overlap_entry(int nc, int kw, int threads) {
  std::ostringstream src;
  src << "extern \"C\" __global__ void __launch_bounds__(" << threads << ") pw_overlap_jit(const __grid_constant__ pw::ScanPlan P) {\n"
      << "  pw::overlap_body<pw::JitCtl, " << nc << ", " << kw << ">(P);\n}\n";
  return src.str();
}
std::string radix_entry(int nc, int kw, int mode, int threads) {
  std::ostringstream src;
  src << "extern \"C\" __global__ void __launch_bounds__(" << threads << (mode == 1 || mode == 2 || (mode == 3 && threads <= 512) ? ", 2" : "") << ") pw_radix_jit_m" << mode << "(const __grid_constant__ pw::ScanPlan P, const pw::RadixParams rp) {\n"
      << "  pw::radix_body<pw::JitCtl, " << nc << ", " << kw << ", " << mode << ">(P, rp);\n}\n";
  return src.str();
}

std::string pilot_entry(int nc, int kw, int threads) {
  std::ostringstream src;
  src << "extern \"C\" __global__ void __launch_bounds__(" << threads << ") pw_pilot_jit(const __grid_constant__ pw::ScanPlan P, const pw::PilotParams pp) {\n"
      << "  pw::pilot_body<pw::JitCtl, " << nc << ", " << kw << ">(P, pp);\n}\n";
  return src.str();
}

// ---- persistent cubin cache -----------------------------------------------------------------------------------
// A fresh process used to pay ~0.7 s of NVRTC for the first query of every shape (a one-shot collect() pays it every
// time).  Compiled cubins are kept on disk under a name derived from the generated source text, the contents of the
// kernel headers it includes and the target architecture; PW_JIT_CACHE_DIR chooses the directory ("off" disables it).
uint64_t fnv1a(const void* p, size_t n, uint64_t h) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
  return h;
}
// hash of every header the generated translation unit includes (computed once per process)
void sources_hash(uint64_t* h0, uint64_t* h1) {
  static uint64_t a0 = 0, a1 = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    a0 = 0xcbf29ce484222325ull; a1 = 0x84222325cbf29ce4ull;
    for (const char* name : {"pw_plan.h", "pw_ctl.h", "pw_scan.cuh", "pw_segmented.cuh", "pw_partition.cuh", "pw_pilot.cuh", "pw_bucket.cuh", "pw_runs.cuh", "pw_radix.cuh", "pw_overlap.cuh"}) {
      const std::string path = csrc_dir() + "/" + name;
      FILE* f = fopen(path.c_str(), "rb");
      if (!f) continue;
      char buf[1 << 16];
      size_t n;
      while ((n = fread(buf, 1, sizeof buf, f)) > 0) { a0 = fnv1a(buf, n, a0); a1 = fnv1a(buf, n, a1 ^ 0x9E3779B97F4A7C15ull); }
      fclose(f);
    }
  });
  *h0 = a0; *h1 = a1;
}
std::string cache_dir() {
  static std::string dir;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* e = getenv("PW_JIT_CACHE_DIR");
    if (e && !strcmp(e, "off")) return;
    std::vector<std::string> cands;
    if (e && *e) cands.push_back(e);
    else {
      if (const char* x = getenv("XDG_CACHE_HOME")) cands.push_back(std::string(x) + "/polarway_b200");
      if (const char* hm = getenv("HOME")) cands.push_back(std::string(hm) + "/.cache/polarway_b200");
      cands.push_back("/tmp/polarway_b200_jit_" + std::to_string((long)getuid()));
    }
    for (const std::string& d : cands) {
      std::string acc;
      for (size_t i = 0; i <= d.size(); ++i)   // mkdir -p
        if (i == d.size() || (d[i] == '/' && i > 0)) { acc = d.substr(0, i); mkdir(acc.c_str(), 0700); }
      if (access(d.c_str(), W_OK | X_OK) == 0) { dir = d; return; }
    }
  });
  return dir;
}
std::string cache_path(const std::string& text) {
  const std::string d = cache_dir();
  if (d.empty()) return "";
  uint64_t h0, h1;
  sources_hash(&h0, &h1);
  h0 = fnv1a(text.data(), text.size(), h0);
  h1 = fnv1a(text.data(), text.size(), h1 ^ 0x9E3779B97F4A7C15ull);
  char name[96];
  snprintf(name, sizeof name, "/sm_100a_%016llx%016llx.cubin", (unsigned long long)h0, (unsigned long long)h1);
  return d + name;
}
bool cache_read(const std::string& path, std::vector<char>* out) {
  if (path.empty()) return false;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  bool ok = n > 64;
  if (ok) { out->resize((size_t)n); ok = fread(out->data(), 1, (size_t)n, f) == (size_t)n; }
  fclose(f);
  return ok && !memcmp(out->data(), "\x7f" "ELF", 4);
}
void cache_write(const std::string& path, const std::vector<char>& cubin) {
  if (path.empty()) return;
  char tmp[64];
  snprintf(tmp, sizeof tmp, ".tmp.%ld.%p", (long)getpid(), (void*)&cubin);
  const std::string t = path + tmp;
  FILE* f = fopen(t.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  fclose(f);
  if (ok) rename(t.c_str(), path.c_str()); else unlink(t.c_str());   // atomic publish: concurrent processes race benignly
}

Compiled compile(const std::string& ctl, const std::string& entry, const char* entry_name) {
  Api& a = api();
  Compiled c;
  const std::string text = "#include \"pw_segmented.cuh\"\n#include \"pw_partition.cuh\"\n#include \"pw_pilot.cuh\"\n#include \"pw_bucket.cuh\"\n#include \"pw_runs.cuh\"\n#include \"pw_radix.cuh\"\n#include \"pw_overlap.cuh\"\nnamespace pw {\n" + ctl + "}\n" + entry;
  const std::string cpath = getenv("PW_JIT_DUMP") || getenv("PW_DEBUG") ? std::string() : cache_path(text);
  {
    std::vector<char> cached;
    if (cache_read(cpath, &cached)) {
      CUmodule mod = nullptr;
      if (a.cuModuleLoadData(&mod, cached.data()) == 0 && a.cuModuleGetFunction(&c.fn, mod, entry_name) == 0) { ctx().timings.jit_cache_hits++; return c; }
      c.fn = nullptr;  // stale or damaged file: compile again (and overwrite it)
    }
  }
  nvrtcProgram prog = nullptr;
  if (a.nvrtcCreateProgram(&prog, text.c_str(), "pw_scan_jit.cu", 0, nullptr, nullptr) != 0) { c.failed = true; return c; }
  const std::string inc = "--include-path=" + csrc_dir();
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", inc.c_str(), "-lineinfo", "--device-as-default-execution-space"};
  const nvrtcResult rc = a.nvrtcCompileProgram(prog, 5, opts);
  if (rc != 0 || getenv("PW_DEBUG")) {
    size_t n = 0;
    a.nvrtcGetProgramLogSize(prog, &n);
    if (n > 1) {
      std::vector<char> log(n + 1);
      a.nvrtcGetProgramLog(prog, log.data());
      if (rc != 0 || n > 2) fprintf(stderr, "[pw jit] nvrtc rc=%d\n%s\n", rc, log.data());
    }
    if (getenv("PW_DEBUG") && rc == 0) fprintf(stderr, "[pw jit] compiled %s for shape:\n%s", entry_name, ctl.c_str());
  }
  if (rc != 0) { a.nvrtcDestroyProgram(&prog); c.failed = true; return c; }
  size_t sz = 0;
  a.nvrtcGetCUBINSize(prog, &sz);
  std::vector<char> cubin(sz);
  a.nvrtcGetCUBIN(prog, cubin.data());
  a.nvrtcDestroyProgram(&prog);
  if (const char* dump = getenv("PW_JIT_DUMP")) {
    FILE* f = fopen(dump, "wb");
    if (f) { fwrite(cubin.data(), 1, sz, f); fclose(f); }
  }
  CUmodule mod = nullptr;
  if (a.cuModuleLoadData(&mod, cubin.data()) != 0) { c.failed = true; return c; }
  if (a.cuModuleGetFunction(&c.fn, mod, entry_name) != 0) { c.failed = true; c.fn = nullptr; }
  else cache_write(cpath, cubin);
  ctx().timings.jit_compiles++;
  return c;
}

}  // namespace

// Cross-compile check used by build(): NVRTC needs no GPU.  Returns 0 when the specialised source compiles.
int jit_selftest_compile(const ScanPlan& P, int nc, int kw, bool hot, int threads, std::string* err) {
  Api& a = api();
  if (!a.nvrtcCreateProgram) {
    // driver may be missing on a CPU-only box: resolve NVRTC alone
    void* rtc = nullptr;
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"})
      if ((rtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!rtc) { *err = "libnvrtc not found"; return 1; }
    *(void**)(&a.nvrtcCreateProgram) = dlsym(rtc, "nvrtcCreateProgram");
    *(void**)(&a.nvrtcCompileProgram) = dlsym(rtc, "nvrtcCompileProgram");
    *(void**)(&a.nvrtcGetProgramLogSize) = dlsym(rtc, "nvrtcGetProgramLogSize");
    *(void**)(&a.nvrtcGetProgramLog) = dlsym(rtc, "nvrtcGetProgramLog");
    *(void**)(&a.nvrtcGetCUBINSize) = dlsym(rtc, "nvrtcGetCUBINSize");
    *(void**)(&a.nvrtcGetCUBIN) = dlsym(rtc, "nvrtcGetCUBIN");
    *(void**)(&a.nvrtcDestroyProgram) = dlsym(rtc, "nvrtcDestroyProgram");
  }
  const std::string text = "#include \"pw_segmented.cuh\"\n#include \"pw_partition.cuh\"\n#include \"pw_pilot.cuh\"\n#include \"pw_bucket.cuh\"\n#include \"pw_runs.cuh\"\n#include \"pw_radix.cuh\"\n#include \"pw_overlap.cuh\"\nnamespace pw {\n" + jit_ctl(P) + "}\n" + scan_entry(nc, kw, hot, threads) + seg_entry(nc, 256) + part_entry(nc, kw, 256) + pilot_entry(nc, kw, 256) + runs_entry(nc, kw, 256) +
                           (nc <= 4 && P.n_slots <= 3 && !P.dyn.enabled ? radix_entry(nc, kw, 0, 256) + radix_entry(nc, kw, 1, RADIX_SC_THREADS) + radix_entry(nc, kw, 2, RADIX_SC_THREADS) + radix_entry(nc, kw, 3, RADIX_THREADS) : std::string()) +
                           (P.dyn.enabled ? overlap_entry(nc, kw, 256) : std::string()) +
                           (P.hot.bucket ? bucket_entry(nc, kw, P.hot.b_threads, P.hot.b_cps) : std::string());
  if (const char* dump_src = getenv("PW_JIT_DUMP_SRC")) {   // the generated translation unit, for offline experiments
    FILE* f = fopen(dump_src, "wb");
    if (f) { fwrite(text.data(), 1, text.size(), f); fclose(f); }
  }
  nvrtcProgram prog = nullptr;
  if (a.nvrtcCreateProgram(&prog, text.c_str(), "pw_scan_jit.cu", 0, nullptr, nullptr) != 0) { *err = "nvrtcCreateProgram failed"; return 2; }
  const std::string inc = "--include-path=" + csrc_dir();
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", inc.c_str(), "-lineinfo", "--device-as-default-execution-space"};
  const nvrtcResult rc = a.nvrtcCompileProgram(prog, 5, opts);
  size_t n = 0;
  a.nvrtcGetProgramLogSize(prog, &n);
  std::vector<char> log(n + 1, 0);
  if (n > 1) a.nvrtcGetProgramLog(prog, log.data());
  *err = log.data();
  if (rc == 0) {
    if (const char* dump = getenv("PW_JIT_DUMP")) {
      size_t sz = 0;
      a.nvrtcGetCUBINSize(prog, &sz);
      std::vector<char> cubin(sz);
      a.nvrtcGetCUBIN(prog, cubin.data());
      FILE* f = fopen(dump, "wb");
      if (f) { fwrite(cubin.data(), 1, sz, f); fclose(f); }
    }
  }
  a.nvrtcDestroyProgram(&prog);
  return rc == 0 ? 0 : 3;
}

// returns 0 launched, 1 JIT unavailable (caller falls back to the AOT kernel), <0 error
// Binary cache key: exactly the fields jit_ctl() turns into constants (the text itself is only generated on a miss —
// building it costs more than a small kernel launch).
static std::string plan_key(const ScanPlan& P) {
  std::string k;
  k.reserve(1024);
  auto put = [&](const void* p, size_t n) { k.append((const char*)p, n); };
  auto i32 = [&](int32_t v) { put(&v, 4); };
  // a CUfunction belongs to the context it was loaded in (one primary context per device): the calling thread's
  // device is part of the key, so pw_b200_set_device(other) compiles/loads its own copy instead of launching a
  // handle from a foreign context
  i32(ctx().device);
  i32(P.n_slots);
  for (int i = 0; i < P.n_slots; ++i) { i32(P.slots[i].dtype); i32(P.slots[i].validity != nullptr); }
  i32(P.n_preds);
  for (int i = 0; i < P.n_preds; ++i) { i32(P.preds[i].slot); i32(P.preds[i].op); i32(P.preds[i].cls); }
  i32(P.n_keys);
  for (int i = 0; i < P.n_keys; ++i) { i32(P.keys[i].slot); i32(P.keys[i].dtype); i32(P.keys[i].n_words); }
  i32(P.has_null_word); i32(P.dyn.enabled); i32(P.dyn.slot); i32(P.dyn.closed);
  if (P.dyn.enabled) { put(&P.dyn.every, 8); put(&P.dyn.period, 8); }
  i32(P.n_vexpr);
  for (int e = 0; e < P.n_vexpr; ++e) {
    const VExpr& v = P.vexprs[e];
    i32(v.n_factors); i32(v.slot); i32(v.cls); i32(v.flags); i32(v.acc_base);
    for (int f = 0; f < v.n_factors; ++f) { i32(v.f[f].slot); put(&v.f[f].a, 8); put(&v.f[f].b, 8); }
  }
  i32(P.gflags); i32(P.acc_gbase); i32(P.n_acc);
  for (int a = 0; a < P.n_acc; ++a) i32(P.accs[a].op);
  i32(P.rowid_slot_p1); i32(P.runs);
  i32(P.vec_ok); i32(P.check_sorted); i32(P.row_begin == 0 && P.row_stride == 1); i32(P.row_group_out != nullptr);
  put(&P.hot, sizeof P.hot);
  return k;
}

// returns 0 launched, 1 JIT unavailable (caller falls back to the AOT kernel), <0 error
int launch_scan_jit(const ScanPlan& P, int nc, int kw, bool hot, int threads, int sm_count, cudaStream_t st) {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  if (disabled) return 1;
  Api& a = api();
  if (!a.ok) return 1;
  // two resident CTAs per SM when the hot table leaves room for them: the register allocator is told so
  const size_t smem = hot ? (size_t)P.hot.total_bytes : 0;
  const int min_blocks = (2 * (smem + 1024) <= 227 * 1024 && 2 * threads <= 1024) ? 2 : 1;
  std::string key = plan_key(P);
  const int32_t tail[5] = {nc, kw, (int32_t)hot, threads, min_blocks};
  key.append((const char*)tail, sizeof tail);
  Compiled c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      c = compile(jit_ctl(P), scan_entry(nc, kw, hot, threads, min_blocks), "pw_scan_jit");
      if (!c.failed && c.fn) {
        if (a.cuFuncSetAttribute(c.fn, 8 /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/, (int)smem) != 0 ||
            a.cuOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, c.fn, threads, smem) != 0 || c.per_sm < 1)
          c.failed = true;
      }
      g_cache[key] = c;
    } else c = it->second;
  }
  if (c.failed || !c.fn) return 1;
  const int64_t n_steps = (P.n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  const int64_t n_tiles = (n_steps + (threads / 32) - 1) / (threads / 32);
  int64_t grid = (int64_t)sm_count * c.per_sm;
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  ScanPlan copy = P;
  void* params[] = {&copy};
  const CUresult rc = a.cuLaunchKernel(c.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, (CUstream)st, params, nullptr);
  if (rc != 0) return fail(PW_ERR_CUDA, "cuLaunchKernel(pw_scan_jit) failed with %d", rc);
  ctx().timings.kernel_launches++;
  return 0;
}

// the bucket tier (pw_bucket.cuh): one CTA of P.hot.b_threads threads per SM; returns 0 launched, 1 unavailable
int launch_bucket_jit(const ScanPlan& P, int nc, int kw, int sm_count, cudaStream_t st) {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  if (disabled) return 1;
  Api& a = api();
  if (!a.ok) return 1;
  const int threads = P.hot.b_threads;
  const size_t smem = (size_t)P.hot.b_bytes;
  std::string key = plan_key(P);
  const int32_t tail[3] = {-4 /* bucket */, nc, kw};
  key.append((const char*)tail, sizeof tail);
  Compiled c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      c = compile(jit_ctl(P), bucket_entry(nc, kw, threads, P.hot.b_cps), "pw_bucket_jit");
      if (!c.failed && c.fn) {
        if (a.cuFuncSetAttribute(c.fn, 8 /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/, (int)smem) != 0 ||
            a.cuOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, c.fn, threads, smem) != 0 || c.per_sm < 1)
          c.failed = true;
      }
      g_cache[key] = c;
    } else c = it->second;
  }
  if (c.failed || !c.fn) return 1;
  const int64_t tile = (int64_t)(threads / 32) * 64;
  const int64_t n_tiles = (P.n_rows + tile - 1) / tile;
  int64_t grid = (int64_t)sm_count * c.per_sm;
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  ScanPlan copy = P;
  void* params[] = {&copy};
  const CUresult rc = a.cuLaunchKernel(c.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, (CUstream)st, params, nullptr);
  if (rc != 0) return fail(PW_ERR_CUDA, "cuLaunchKernel(pw_bucket_jit) failed with %d", rc);
  ctx().timings.kernel_launches++;
  return 0;
}

// the sorted-key run scan (pw_runs.cuh); returns 0 launched, 1 unavailable
int launch_runs_jit(const ScanPlan& P, int nc, int kw, int sm_count, cudaStream_t st) {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  if (disabled) return 1;
  Api& a = api();
  if (!a.ok) return 1;
  const int threads = 256;
  std::string key = plan_key(P);
  const int32_t tail[3] = {-5 /* runs */, nc, kw};
  key.append((const char*)tail, sizeof tail);
  Compiled c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      c = compile(jit_ctl(P), runs_entry(nc, kw, threads), "pw_runs_jit");
      if (!c.failed && c.fn && (a.cuOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, c.fn, threads, 0) != 0 || c.per_sm < 1)) c.failed = true;
      g_cache[key] = c;
    } else c = it->second;
  }
  if (c.failed || !c.fn) return 1;
  const int64_t n_steps = (P.n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  int64_t grid = std::min<int64_t>((int64_t)sm_count * c.per_sm, (n_steps + (threads / 32) - 1) / (threads / 32));
  if (grid < 1) grid = 1;
  ScanPlan copy = P;
  void* params[] = {&copy};
  const CUresult rc = a.cuLaunchKernel(c.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, 0, (CUstream)st, params, nullptr);
  if (rc != 0) return fail(PW_ERR_CUDA, "cuLaunchKernel(pw_runs_jit) failed with %d", rc);
  ctx().timings.kernel_launches++;
  return 0;
}

// the partitioning passes (histogram / scatter), specialised the same way; returns 0 launched, 1 unavailable
int launch_part_jit(const ScanPlan& P, const PartParams& pp, int nc, int kw, int sm_count, cudaStream_t st) {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  if (disabled) return 1;
  Api& a = api();
  if (!a.ok) return 1;
  const int threads = 256;
  std::string key = plan_key(P);
  const int32_t tail[3] = {-2 /* part */, nc, kw};
  key.append((const char*)tail, sizeof tail);
  Compiled c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      c = compile(jit_ctl(P), part_entry(nc, kw, threads), "pw_part_jit");
      if (!c.failed && c.fn) {
        if (a.cuOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, c.fn, threads, 0) != 0 || c.per_sm < 1) c.failed = true;
      }
      g_cache[key] = c;
    } else c = it->second;
  }
  if (c.failed || !c.fn) return 1;
  const int64_t n_steps = (P.n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  int64_t grid = std::min<int64_t>((int64_t)sm_count * c.per_sm, (n_steps + (threads / 32) - 1) / (threads / 32));
  if (grid < 1) grid = 1;
  ScanPlan copy = P;
  PartParams ppc = pp;
  void* params[] = {&copy, &ppc};
  const CUresult rc = a.cuLaunchKernel(c.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, 0, (CUstream)st, params, nullptr);
  if (rc != 0) return fail(PW_ERR_CUDA, "cuLaunchKernel(pw_part_jit) failed with %d", rc);
  ctx().timings.kernel_launches++;
  return 0;
}
// overlapping dynamic windows through the HBM table (pw_overlap.cuh); returns 0 launched, 1 unavailable
int launch_overlap_jit(const ScanPlan& P, int nc, int kw, int sm_count, cudaStream_t st) {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  if (disabled) return 1;
  Api& a = api();
  if (!a.ok) return 1;
  const int threads = 256;
  std::string key = plan_key(P);
  const int32_t tail[3] = {-5 /* overlap */, nc, kw};
  key.append((const char*)tail, sizeof tail);
  Compiled c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      c = compile(jit_ctl(P), overlap_entry(nc, kw, threads), "pw_overlap_jit");
      if (!c.failed && c.fn) {
        if (a.cuOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, c.fn, threads, 0) != 0 || c.per_sm < 1) c.failed = true;
      }
      g_cache[key] = c;
    } else c = it->second;
  }
  if (c.failed || !c.fn) return 1;
  const int64_t n_steps = (P.n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  int64_t grid = std::min<int64_t>((int64_t)sm_count * c.per_sm, (n_steps + (threads / 32) - 1) / (threads / 32));
  if (grid < 1) grid = 1;
  ScanPlan copy = P;
  void* params[] = {&copy};
  const CUresult rc = a.cuLaunchKernel(c.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, 0, (CUstream)st, params, nullptr);
  if (rc != 0) return fail(PW_ERR_CUDA, "cuLaunchKernel(pw_overlap_jit) failed with %d", rc);
  ctx().timings.kernel_launches++;
  ctx().timings.reserved = 1.0f;
  return 0;
}
// the radix tier's passes (pw_radix.cuh), one specialised kernel per mode; returns 0 launched, 1 unavailable
int launch_radix_jit(const ScanPlan& P, const RadixParams& rp, int nc, int kw, size_t smem, int64_t work_ctas, int sm_count, cudaStream_t st) {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  if (disabled) return 1;
  Api& a = api();
  if (!a.ok) return 1;
  const int threads = rp.mode == 0 ? (smem > 64 * 1024 ? 1024 : 256) : (rp.mode == 3 ? rp.agg_threads : RADIX_SC_THREADS);
  std::string key = plan_key(P);
  const int32_t tail[6] = {-4 /* radix */, nc, kw, rp.mode, (int32_t)smem, threads};
  key.append((const char*)tail, sizeof tail);
  Compiled c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      const std::string name = "pw_radix_jit_m" + std::to_string(rp.mode);
      c = compile(jit_ctl(P), radix_entry(nc, kw, rp.mode, threads), name.c_str());
      if (!c.failed && c.fn) {
        if ((smem > 48 * 1024 && a.cuFuncSetAttribute(c.fn, 8, (int)smem) != 0) ||
            a.cuOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, c.fn, threads, smem) != 0 || c.per_sm < 1)
          c.failed = true;
      }
      g_cache[key] = c;
    } else c = it->second;
  }
  if (c.failed || !c.fn) return 1;
  int64_t grid = std::min<int64_t>((int64_t)sm_count * c.per_sm, work_ctas);
  if (grid < 1) grid = 1;
  ScanPlan copy = P;
  RadixParams rpc = rp;
  void* params[] = {&copy, &rpc};
  const CUresult rc = a.cuLaunchKernel(c.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, (CUstream)st, params, nullptr);
  if (rc != 0) return fail(PW_ERR_CUDA, "cuLaunchKernel(pw_radix_jit mode %d) failed with %d", rp.mode, rc);
  ctx().timings.kernel_launches++;
  ctx().timings.reserved = 1.0f;
  return 0;
}
// the fused key-sample pilot; returns 0 launched, 1 unavailable
int launch_pilot_jit(const ScanPlan& P, const PilotParams& pp, int nc, int kw, cudaStream_t st) {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  if (disabled) return 1;
  Api& a = api();
  if (!a.ok) return 1;
  const int threads = 256;
  std::string key = plan_key(P);
  const int32_t tail[3] = {-3 /* pilot */, nc, kw};
  key.append((const char*)tail, sizeof tail);
  Compiled c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      c = compile(jit_ctl(P), pilot_entry(nc, kw, threads), "pw_pilot_jit");
      g_cache[key] = c;
    } else c = it->second;
  }
  if (c.failed || !c.fn) return 1;
  const int64_t n_max = std::max(pp.n[0], pp.n[1]);
  int64_t grid = std::max<int64_t>(1, (n_max / 2 + threads - 1) / threads);
  ScanPlan copy = P;
  PilotParams ppc = pp;
  void* params[] = {&copy, &ppc};
  const CUresult rc = a.cuLaunchKernel(c.fn, (unsigned)grid, 2, 1, (unsigned)threads, 1, 1, 0, (CUstream)st, params, nullptr);
  if (rc != 0) return fail(PW_ERR_CUDA, "cuLaunchKernel(pw_pilot_jit) failed with %d", rc);
  ctx().timings.kernel_launches++;
  return 0;
}
bool jit_available() {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  return !disabled && api().ok;
}

// the sorted-window kernel, specialised the same way; returns 0 launched, 1 unavailable
int launch_seg_jit(const ScanPlan& P, const SegParams& sp, int nc, int threads, size_t smem, int sm_count, cudaStream_t st) {
  static const bool disabled = getenv("PW_NO_JIT") != nullptr;
  if (disabled) return 1;
  Api& a = api();
  if (!a.ok) return 1;
  std::string key = plan_key(P);
  const int32_t tail[4] = {-1 /* seg */, nc, threads, (int32_t)smem};
  key.append((const char*)tail, sizeof tail);
  Compiled c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      c = compile(jit_ctl(P), seg_entry(nc, threads), "pw_seg_jit");
      if (!c.failed && c.fn) {
        if ((smem > 48 * 1024 && a.cuFuncSetAttribute(c.fn, 8, (int)smem) != 0) ||
            a.cuOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, c.fn, threads, smem) != 0 || c.per_sm < 1)
          c.failed = true;
      }
      g_cache[key] = c;
    } else c = it->second;
  }
  if (c.failed || !c.fn) return 1;
  const int per_sm = c.per_sm;
  const int64_t n_steps = (P.n_rows + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
  int64_t grid = std::min<int64_t>((int64_t)sm_count * per_sm, (n_steps + (threads / 32) - 1) / (threads / 32));
  if (grid < 1) grid = 1;
  ScanPlan copy = P;
  SegParams spc = sp;
  void* params[] = {&copy, &spc};
  const CUresult rc = a.cuLaunchKernel(c.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, (CUstream)st, params, nullptr);
  if (rc != 0) return fail(PW_ERR_CUDA, "cuLaunchKernel(pw_seg_jit) failed with %d", rc);
  ctx().timings.kernel_launches++;
  ctx().timings.reserved = 1.0f;
  return 0;
}

}  // namespace pw
