// jit.cu — specialises the hand-written scan kernel body (scan_body.cuh) on one query with NVRTC.
//
// The generated translation unit is tiny: it includes the very same headers the ahead-of-time interpreter kernels are
// built from (embedded in this library at build time, jit_embedded.inc) and spells the compiled program as one
// `exec_insn<R>(Insn{...literal fields...}, ...)` call per instruction, plus constexpr tables of the column types,
// validity-bitmap presence and accumulator kinds.  NVRTC inlines exec_insn with constant fields, so the evaluator's
// opcode / type switches fold away, every column load is typed and hoistable, duplicate loads of a column merge,
// and accumulators and the spill stack live in registers — the semantics stay those of the one evaluator definition
// (interp.cuh), which the parity tests run in both forms.
//
// libnvrtc is dlopen'ed on first use; the cubin is loaded with the driver API obtained through
// cudaGetDriverEntryPoint (no link-time dependency on libcuda, so the library still loads on a box without a driver).
// Kernels are cached per (device, source text) for the life of the process.  QGPU_JIT_DUMP=<dir> writes every
// generated source, NVRTC log and cubin there (for cuobjdump -sass / ptxas -v).
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <unordered_map>

#include "jit.h"

namespace qgpu {

namespace {

struct EmbeddedHeader {
  const char* name;
  const unsigned char* data;
  size_t size;
};
#include "jit_embedded.inc"  // static const EmbeddedHeader kHeaders[]; static const int kNumHeaders;

struct NvrtcApi {
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
  const char* (*GetErrorString)(nvrtcResult);
};

NvrtcApi* nvrtc() {
  static NvrtcApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* nm : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"}) {
      h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) {
      err = std::string("cannot load libnvrtc.so.12: ") + dlerror();
      return;
    }
#define QG_SYM(field, name)                                      \
  api.field = (decltype(api.field))dlsym(h, name);               \
  if (!api.field) err = std::string("libnvrtc lacks ") + name;
    QG_SYM(CreateProgram, "nvrtcCreateProgram")
    QG_SYM(CompileProgram, "nvrtcCompileProgram")
    QG_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    QG_SYM(GetCUBIN, "nvrtcGetCUBIN")
    QG_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    QG_SYM(GetProgramLog, "nvrtcGetProgramLog")
    QG_SYM(DestroyProgram, "nvrtcDestroyProgram")
    QG_SYM(GetErrorString, "nvrtcGetErrorString")
#undef QG_SYM
  });
  if (!err.empty()) throw_error(QGPU_ERR_CUDA, "%s", err.c_str());
  return &api;
}

struct DriverApi {
  CUresult (*ModuleLoadData)(CUmodule*, const void*);
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**,
                           void**);
  CUresult (*GetErrorString)(CUresult, const char**);
};

DriverApi* driver() {
  static DriverApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    auto get = [&](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult st;
      cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st);
      if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || !*fn) err = std::string("driver entry point missing: ") + name;
    };
    get("cuModuleLoadData", (void**)&api.ModuleLoadData);
    get("cuModuleGetFunction", (void**)&api.ModuleGetFunction);
    get("cuFuncSetAttribute", (void**)&api.FuncSetAttribute);
    get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&api.OccupancyMaxActiveBlocksPerMultiprocessor);
    get("cuLaunchKernel", (void**)&api.LaunchKernel);
    get("cuGetErrorString", (void**)&api.GetErrorString);
  });
  if (!err.empty()) throw_error(QGPU_ERR_CUDA, "%s", err.c_str());
  return &api;
}

#define QG_CU(call)                                                                                  \
  do {                                                                                               \
    CUresult _r = (call);                                                                            \
    if (_r != CUDA_SUCCESS) {                                                                        \
      const char* _s = nullptr;                                                                      \
      driver()->GetErrorString(_r, &_s);                                                             \
      ::qgpu::throw_error(QGPU_ERR_CUDA, "CUDA driver error at %s:%d: %s", __FILE__, __LINE__, _s ? _s : "?"); \
    }                                                                                                \
  } while (0)

std::mutex g_mu;
struct CacheKey {
  int device;
  std::string src;
  bool operator==(const CacheKey& o) const { return device == o.device && src == o.src; }
};
struct CacheKeyHash {
  size_t operator()(const CacheKey& k) const { return std::hash<std::string>()(k.src) * 31u + (size_t)k.device; }
};

}  // namespace

struct JitKernel {
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  size_t smem_limit = 48 * 1024;
  std::string log;
};

namespace {
std::unordered_map<CacheKey, std::unique_ptr<JitKernel>, CacheKeyHash> g_cache;
std::unordered_map<std::string, std::unique_ptr<std::string>> g_names;
int g_dump_seq = 0;
}  // namespace

const char* intern_name(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_names.find(s);
  if (it == g_names.end()) it = g_names.emplace(s, std::make_unique<std::string>(s)).first;
  return it->second->c_str();
}

// ---- source generation --------------------------------------------------------------------------
static void emit_insn(std::ostringstream& o, const Insn& in) {
  o << "    exec_insn<R>(Insn{" << (int)in.op << ", " << (int)in.ty << ", " << (int)in.mode << ", " << (int)in.flags << ", " << in.arg
    << ", 0x" << std::hex << in.imm << std::dec << "ull}, m, ld, st, sk);\n";
}

std::string jit_scan_source(int sink, int R, const ScanParams& p, const Program& prog, bool eager, int min_blocks) {
  std::ostringstream o;
  const int n_cols = (int)prog.col_of_slot.size();
  if (sink == SK_GROUP_PART) o << "#define QGPU_GP_PRELOAD " << (p.gp_preload ? 1 : 0) << "\n";
  o << "#include \"scan_body.cuh\"\n";
  o << "namespace qgpu {\nstruct JitProg {\n";
  o << "  static constexpr bool STATIC = true;\n";
  o << "  static constexpr bool EAGER = " << (eager ? "true" : "false") << ";\n";
  o << "  static constexpr int N_ACC = " << prog.acc_kinds.size() << ";\n";
  o << "  static constexpr int STACK = " << prog.max_stack << ";\n";
  o << "  static constexpr int N_COLS = " << n_cols << ";\n";
  o << "  static constexpr int N_KEYS = " << p.n_keys << ";\n";
  o << "  static QG_D int key_slot(int k) {\n    constexpr int t[] = {";
  for (int i = 0; i < p.n_keys; ++i) o << (i ? ", " : "") << p.key_slots[i];
  if (p.n_keys == 0) o << "0";
  o << "};\n    return t[k];\n  }\n";
  o << "  static QG_D int col_dtype(int slot, const ColRef*) {\n    constexpr int t[] = {";
  for (int i = 0; i < n_cols; ++i) o << (i ? ", " : "") << p.cols[i].dtype;
  if (n_cols == 0) o << "0";
  o << "};\n    return t[slot];\n  }\n";
  o << "  static QG_D bool col_has_valid(int slot, const ColRef*) {\n    constexpr bool t[] = {";
  for (int i = 0; i < n_cols; ++i) o << (i ? ", " : "") << (p.cols[i].valid != nullptr ? "true" : "false");
  if (n_cols == 0) o << "false";
  o << "};\n    return t[slot];\n  }\n";
  o << "  static QG_D uint8_t acc_kind(int s) {\n    constexpr uint8_t t[] = {";
  for (size_t i = 0; i < prog.acc_kinds.size(); ++i) o << (i ? ", " : "") << (int)prog.acc_kinds[i];
  if (prog.acc_kinds.empty()) o << "0";
  o << "};\n    return t[s];\n  }\n";
  const int n_pred = p.n_pred > 0 ? p.n_pred : 0;
  o << "  template <int R, class L, class S, class K>\n  QG_D void run_pred(Machine<R>& m, L& ld, S& st, K& sk) const {\n";
  for (int i = 0; i < n_pred && i < (int)prog.code.size(); ++i) emit_insn(o, prog.code[i]);
  o << "  }\n";
  o << "  template <int R, class L, class S, class K>\n  QG_D void run_rest(Machine<R>& m, L& ld, S& st, K& sk) const {\n";
  for (int i = n_pred; i < (int)prog.code.size(); ++i) emit_insn(o, prog.code[i]);
  o << "  }\n};\n}  // namespace qgpu\n";
  o << "extern \"C\" __global__ void __launch_bounds__(" << SCAN_THREADS << ", " << min_blocks << ") qgpu_jit_scan(const __grid_constant__ qgpu::ScanParams P) {\n";
  o << "  extern __shared__ __align__(16) unsigned char smem_raw[];\n";
  o << "  qgpu::scan_body<" << R << ", " << sink << ", qgpu::JitProg>(P, qgpu::JitProg{}, smem_raw);\n}\n";
  return o.str();
}

std::string jit_scan_key(int sink, int R, const ScanParams& p, const Program& prog) {
  std::string k;
  k.reserve(64 + prog.code.size() * sizeof(Insn));
  auto put = [&](const void* d, size_t n) { k.append((const char*)d, n); };
  const int32_t head[] = {sink, R, (int32_t)prog.col_of_slot.size(), (int32_t)prog.acc_kinds.size(), prog.max_stack, p.n_pred, p.n_keys,
                          sink == SK_GROUP_PART ? (int32_t)p.gp_preload : 0};
  put(head, sizeof head);
  for (size_t i = 0; i < prog.col_of_slot.size(); ++i) {
    const int8_t d[2] = {(int8_t)p.cols[i].dtype, (int8_t)(p.cols[i].valid != nullptr)};
    put(d, 2);
  }
  put(prog.acc_kinds.data(), prog.acc_kinds.size());
  put(p.key_slots, sizeof(int32_t) * (size_t)(p.n_keys > 0 ? p.n_keys : 0));
  put(prog.code.data(), prog.code.size() * sizeof(Insn));
  return k;
}

// ---- compilation ----------------------------------------------------------------------------------
std::string jit_compile(const std::string& src, std::string* log_out) {
  NvrtcApi* N = nvrtc();
  std::vector<const char*> names, texts;
  std::vector<std::string> bodies;
  bodies.reserve(kNumHeaders);
  for (int i = 0; i < kNumHeaders; ++i) {
    bodies.emplace_back((const char*)kHeaders[i].data, kHeaders[i].size);
    names.push_back(kHeaders[i].name);
  }
  for (auto& b : bodies) texts.push_back(b.c_str());
  nvrtcProgram prog = nullptr;
  nvrtcResult r = N->CreateProgram(&prog, src.c_str(), "qgpu_jit_scan.cu", kNumHeaders, texts.data(), names.data());
  if (r != NVRTC_SUCCESS) throw_error(QGPU_ERR_CUDA, "nvrtcCreateProgram: %s", N->GetErrorString(r));
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=false", "--ptxas-options=-v",
                        "-default-device"};
  r = N->CompileProgram(prog, (int)(sizeof opts / sizeof opts[0]), opts);
  std::string log;
  size_t ls = 0;
  if (N->GetProgramLogSize(prog, &ls) == NVRTC_SUCCESS && ls > 1) {
    log.resize(ls);
    N->GetProgramLog(prog, &log[0]);
  }
  if (log_out) *log_out = log;
  if (r != NVRTC_SUCCESS) {
    N->DestroyProgram(&prog);
    throw_error(QGPU_ERR_CUDA, "NVRTC failed to specialise the scan kernel: %s\n%.3000s", N->GetErrorString(r), log.c_str());
  }
  size_t cs = 0;
  std::string cubin;
  r = N->GetCUBINSize(prog, &cs);
  if (r == NVRTC_SUCCESS) {
    cubin.resize(cs);
    r = N->GetCUBIN(prog, &cubin[0]);
  }
  N->DestroyProgram(&prog);
  if (r != NVRTC_SUCCESS) throw_error(QGPU_ERR_CUDA, "nvrtcGetCUBIN: %s", N->GetErrorString(r));
  if (const char* dir = getenv("QGPU_JIT_DUMP")) {
    int seq;
    {
      std::lock_guard<std::mutex> lk(g_mu);
      seq = g_dump_seq++;
    }
    const std::string base = std::string(dir) + "/qgpu_jit_" + std::to_string(seq);
    std::ofstream(base + ".cu") << src;
    std::ofstream(base + ".log") << log;
    std::ofstream(base + ".cubin", std::ios::binary).write(cubin.data(), (std::streamsize)cubin.size());
  }
  return cubin;
}

bool jit_cached(const std::string& key) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  return g_cache.find(CacheKey{dev, key}) != g_cache.end();
}

JitKernel* jit_get(Ctx* c, const std::string& key, const std::function<std::string()>& make_source, const char* entry) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(CacheKey{c->device, key});
    if (it != g_cache.end()) return it->second.get();
  }
  auto k = std::make_unique<JitKernel>();
  const std::string cubin = jit_compile(make_source(), &k->log);
  DriverApi* D = driver();
  QG_CU(D->ModuleLoadData(&k->mod, cubin.data()));
  QG_CU(D->ModuleGetFunction(&k->fn, k->mod, entry));
  std::lock_guard<std::mutex> lk(g_mu);
  auto ins = g_cache.emplace(CacheKey{c->device, key}, std::move(k));
  return ins.first->second.get();
}

int jit_prepare(Ctx* c, JitKernel* k, int threads, size_t smem) {
  (void)c;
  DriverApi* D = driver();
  if (smem > k->smem_limit) {
    QG_CU(D->FuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
    k->smem_limit = smem;
  }
  int per_sm = 0;
  QG_CU(D->OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k->fn, threads, smem));
  return per_sm;
}

void jit_launch(Ctx* c, JitKernel* k, unsigned grid, unsigned block, size_t smem, void** args) {
  QG_CU(driver()->LaunchKernel(k->fn, grid, 1, 1, block, 1, 1, (unsigned)smem, (CUstream)c->stream, args, nullptr));
}

}  // namespace qgpu
