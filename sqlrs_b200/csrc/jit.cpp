// sqlrs_b200 — NVRTC specialisation + module cache (see jit.hpp).
#include "jit.hpp"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>

namespace sq {

struct JitKernel {
  CUfunction fn = nullptr;
  CUmodule mod = nullptr;
  size_t smem_opt_in = 0;
  CUdeviceptr rank_sym = 0;          // address of the module's sq_rank_table (0: the kernel compares no strings by order)
  const void** rank_value = nullptr; // persistent host cell the pointer is copied from
};

namespace {

struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**,
                           void**) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
  CUresult (*MemcpyHtoDAsync)(CUdeviceptr, const void*, size_t, CUstream) = nullptr;
};

std::mutex g_mu;
Driver g_drv;
bool g_drv_ready = false;
std::map<std::string, JitKernel*> g_kernels;  // key: device|kernel|hash(source)

template <typename F>
void resolve(const char* name, F& out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st);
  if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || !p)
    fail(SQLRS_ERR_CUDA, std::string("cannot resolve driver entry point ") + name + " (no NVIDIA driver?)");
  out = reinterpret_cast<F>(p);
}

void ensure_driver() {
  if (g_drv_ready) return;
  SQ_CUDA(cudaFree(nullptr));  // make sure the primary context exists and is current
  resolve("cuModuleLoadData", g_drv.ModuleLoadData);
  resolve("cuModuleGetFunction", g_drv.ModuleGetFunction);
  resolve("cuLaunchKernel", g_drv.LaunchKernel);
  resolve("cuFuncSetAttribute", g_drv.FuncSetAttribute);
  resolve("cuGetErrorString", g_drv.GetErrorString);
  resolve("cuOccupancyMaxActiveBlocksPerMultiprocessor", g_drv.OccupancyMaxActiveBlocksPerMultiprocessor);
  resolve("cuModuleGetGlobal", g_drv.ModuleGetGlobal);
  resolve("cuMemcpyHtoDAsync", g_drv.MemcpyHtoDAsync);
  g_drv_ready = true;
}

void cu_check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  fail(SQLRS_ERR_CUDA, std::string(what) + ": " + (s ? s : "unknown driver error"));
}

uint64_t fnv1a(const std::string& s, uint64_t h = 0xcbf29ce484222325ULL) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 0x100000001b3ULL;
  }
  return h;
}

std::string cache_dir() {
  const char* env = std::getenv("SQLRS_B200_JIT_CACHE");
  if (env && *env) return env;
  Dl_info info;
  if (dladdr((void*)&cache_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t slash = p.rfind('/');
    if (slash != std::string::npos) return p.substr(0, slash) + "/jit_cache";
  }
  return "/tmp/sqlrs_b200_jit_cache";
}

const char* kArch = "--gpu-architecture=sm_100a";

std::string hash_name(const std::string& full_source) {
  char buf[40];
  std::snprintf(buf, sizeof buf, "%016llx", (unsigned long long)fnv1a(full_source, fnv1a(kArch)));
  return buf;
}

}  // namespace

std::string jit_full_source(const std::string& skeleton, const std::string& generated) {
  std::string src;
  // tuning experiments: SQLRS_B200_JIT_DEFINES="SQ_JUNROLL=12;SQ_JMINB=5" puts #defines in front of every kernel source
  // (the skeletons guard their tunables with #ifndef); part of the source, hence of the cache key
  if (const char* defs = std::getenv("SQLRS_B200_JIT_DEFINES")) {
    std::string d(defs);
    size_t pos = 0;
    while (pos < d.size()) {
      size_t semi = d.find(';', pos);
      if (semi == std::string::npos) semi = d.size();
      std::string one = d.substr(pos, semi - pos);
      const size_t eq = one.find('=');
      if (!one.empty()) src += "#define " + (eq == std::string::npos ? one : one.substr(0, eq) + " " + one.substr(eq + 1)) + "\n";
      pos = semi + 1;
    }
  }
  src += embedded_source("prelude");
  if (generated.find("sq_str_rank(") != std::string::npos) src += embedded_source("strrank");  // Utf8 ordering comparisons
  src += "\n// ---- generated row program -------------------------------------------------\n";
  src += generated;
  src += "\n// ---- skeleton: ";
  src += skeleton;
  src += " -------------------------------------------------------\n";
  // "a+b": building blocks concatenated in order
  size_t pos = 0;
  while (pos <= skeleton.size()) {
    size_t plus = skeleton.find('+', pos);
    if (plus == std::string::npos) plus = skeleton.size();
    src += embedded_source(skeleton.substr(pos, plus - pos));
    src += "\n";
    pos = plus + 1;
  }
  return src;
}

std::string jit_compile_to_cubin(const std::string& skeleton, const std::string& generated, std::string* log) {
  std::string src = jit_full_source(skeleton, generated);
  std::string name = "sqlrs_jit_" + skeleton + "_" + hash_name(src) + ".cu";
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, src.c_str(), name.c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS)
    fail(SQLRS_ERR_INTERNAL, "nvrtcCreateProgram failed");
  const char* opts[] = {kArch, "--std=c++17", "-lineinfo", "--extra-device-vectorization", "-diag-suppress=177"};
  nvrtcResult r = nvrtcCompileProgram(prog, 5, opts);
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  std::string lg(log_size, '\0');
  if (log_size > 1) nvrtcGetProgramLog(prog, &lg[0]);
  if (log) *log = lg;
  if (r != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    const char* dump = std::getenv("SQLRS_B200_JIT_DUMP");
    if (dump && *dump) {
      std::ofstream f(std::string(dump) + "/" + name);
      f << src;
    }
    fail(SQLRS_ERR_INTERNAL, "NVRTC compilation of " + name + " failed:\n" + lg);
  }
  size_t sz = 0;
  nvrtcGetCUBINSize(prog, &sz);
  std::string cubin(sz, '\0');
  nvrtcGetCUBIN(prog, &cubin[0]);
  nvrtcDestroyProgram(&prog);
  const char* dump = std::getenv("SQLRS_B200_JIT_DUMP");
  if (dump && *dump) {
    std::ofstream f(std::string(dump) + "/" + name);
    f << src;
  }
  return cubin;
}

static std::string load_or_compile(const std::string& skeleton, const std::string& generated) {
  std::string src = jit_full_source(skeleton, generated);
  std::string dir = cache_dir();
  std::string path = dir + "/" + skeleton + "_" + hash_name(src) + ".cubin";
  {
    std::ifstream f(path, std::ios::binary);
    if (f) {
      std::stringstream ss;
      ss << f.rdbuf();
      std::string c = ss.str();
      if (!c.empty()) return c;
    }
  }
  std::string cubin = jit_compile_to_cubin(skeleton, generated, nullptr);
  mkdir(dir.c_str(), 0755);
  std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  {
    std::ofstream f(tmp, std::ios::binary);
    if (f) f.write(cubin.data(), (std::streamsize)cubin.size());
  }
  if (rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());
  return cubin;
}

JitKernel* jit_get(const std::string& skeleton, const std::string& generated, const std::string& kernel_name) {
  std::lock_guard<std::mutex> lock(g_mu);
  ensure_driver();
  int dev = 0;
  SQ_CUDA(cudaGetDevice(&dev));
  char keybuf[64];
  std::snprintf(keybuf, sizeof keybuf, "%d|%016llx|", dev, (unsigned long long)fnv1a(generated, fnv1a(skeleton)));
  std::string key = keybuf + kernel_name;
  auto it = g_kernels.find(key);
  if (it != g_kernels.end()) return it->second;
  // kernels of one module share the compile: look for a sibling with the same module
  std::string modkey = std::string(keybuf) + "#module";
  CUmodule mod = nullptr;
  auto mit = g_kernels.find(modkey);
  if (mit != g_kernels.end()) {
    mod = mit->second->mod;
  } else {
    std::string cubin = load_or_compile(skeleton, generated);
    cu_check(g_drv.ModuleLoadData(&mod, cubin.data()), "cuModuleLoadData");
    auto* holder = new JitKernel();
    holder->mod = mod;
    g_kernels[modkey] = holder;
  }
  auto* k = new JitKernel();
  k->mod = mod;
  cu_check(g_drv.ModuleGetFunction(&k->fn, mod, kernel_name.c_str()), ("cuModuleGetFunction " + kernel_name).c_str());
  if (generated.find("sq_str_rank(") != std::string::npos) {
    size_t bytes = 0;
    cu_check(g_drv.ModuleGetGlobal(&k->rank_sym, &bytes, mod, "sq_rank_table"), "cuModuleGetGlobal sq_rank_table");
    k->rank_value = new const void*(nullptr);
  }
  g_kernels[key] = k;
  return k;
}

void jit_launch(JitKernel* k, unsigned grid, unsigned block, size_t dyn_smem, cudaStream_t stream, void** args) {
  if (dyn_smem > 48 * 1024 && dyn_smem > k->smem_opt_in) {
    cu_check(g_drv.FuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn_smem),
             "cuFuncSetAttribute(max dynamic shared)");
    k->smem_opt_in = dyn_smem;
  }
  if (k->rank_sym) {  // the kernel compares strings by order: point it at the current rank table (stream-ordered)
    const void* table = string_rank_table(stream);
    if (*k->rank_value != table) {
      *k->rank_value = table;
      cu_check(g_drv.MemcpyHtoDAsync(k->rank_sym, k->rank_value, sizeof(void*), (CUstream)stream), "cuMemcpyHtoDAsync sq_rank_table");
    }
  }
  cu_check(g_drv.LaunchKernel(k->fn, grid, 1, 1, block, 1, 1, (unsigned)dyn_smem, (CUstream)stream, args, nullptr), "cuLaunchKernel");
  count_launch();
}

int jit_max_blocks_per_sm(JitKernel* k, int block, size_t dyn_smem) {
  if (dyn_smem > 48 * 1024 && dyn_smem > k->smem_opt_in) {
    cu_check(g_drv.FuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn_smem),
             "cuFuncSetAttribute(max dynamic shared)");
    k->smem_opt_in = dyn_smem;
  }
  int nb = 0;
  cu_check(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&nb, k->fn, block, dyn_smem), "cuOccupancyMaxActiveBlocksPerMultiprocessor");
  return nb;
}

int device_sm_count(int device) {
  static std::map<int, int> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(device);
  if (it != cache.end()) return it->second;
  int n = 0;
  SQ_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  cache[device] = n;
  return n;
}

}  // namespace sq
