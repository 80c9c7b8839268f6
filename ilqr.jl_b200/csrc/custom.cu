// custom.cu — run-time compilation (NVRTC) and launch of the kernels for user-defined dynamics
// (custom_kernels.cuh).  libnvrtc is opened with dlopen so that libilqr_b200.so loads on machines
// without it; the compiled cubin is loaded through the runtime's library API (no driver-API linkage).
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "internal.cuh"
#include "_embedded.inc"

namespace ilqr {

namespace {

struct Nvrtc {
  void* so = nullptr;
  decltype(&nvrtcCreateProgram) create = nullptr;
  decltype(&nvrtcCompileProgram) compile = nullptr;
  decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
  decltype(&nvrtcGetProgramLog) log = nullptr;
  decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
  decltype(&nvrtcGetCUBIN) cubin = nullptr;
  decltype(&nvrtcDestroyProgram) destroy = nullptr;
  decltype(&nvrtcGetErrorString) errstr = nullptr;
};

const Nvrtc* nvrtc(std::string& err) {
  static Nvrtc api;
  static std::once_flag once;
  static std::string load_err;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nme : names)
      if ((api.so = dlopen(nme, RTLD_NOW | RTLD_LOCAL))) break;
    if (!api.so) { load_err = "libnvrtc not found (needed for ILQR_MODEL_CUSTOM)"; return; }
#define SYM(field, name)                                                   \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.so, name)); \
  if (!api.field) load_err = std::string("libnvrtc lacks ") + name;
    SYM(create, "nvrtcCreateProgram") SYM(compile, "nvrtcCompileProgram") SYM(log_size, "nvrtcGetProgramLogSize")
    SYM(log, "nvrtcGetProgramLog") SYM(cubin_size, "nvrtcGetCUBINSize") SYM(cubin, "nvrtcGetCUBIN")
    SYM(destroy, "nvrtcDestroyProgram") SYM(errstr, "nvrtcGetErrorString")
#undef SYM
  });
  if (!load_err.empty()) { err = load_err; return nullptr; }
  return &api;
}

std::mutex g_cache_mu;
std::map<std::string, CustomModule> g_cache;   // key: device | n | m | source

inline unsigned grid_for(int n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

// Compile the user's snippet into a cubin for `arch` (e.g. "sm_100a").  No GPU needed.
int32_t custom_compile(const char* user_src, int n, int m, bool user_cost, const char* arch, std::vector<char>& cubin,
                       std::string& log) {
  std::string err;
  const Nvrtc* rt = nvrtc(err);
  if (!rt) { log = err; return -1; }
  const std::string src = std::string(kEmbeddedPre) + "\n// ---- user snippet ----\n" + user_src +
                          "\n// ---- end of user snippet ----\n" + kEmbeddedPost;
  nvrtcProgram prog;
  if (rt->create(&prog, src.c_str(), "ilqr_custom.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
    log = "nvrtcCreateProgram failed";
    return -1;
  }
  const std::string a = std::string("--gpu-architecture=") + arch, dn = "-DILQR_N=" + std::to_string(n),
                    dm = "-DILQR_M=" + std::to_string(m), dc = std::string("-DILQR_USER_COST=") + (user_cost ? "1" : "0");
  const char* opts[] = {a.c_str(), "-std=c++17", dn.c_str(), dm.c_str(), dc.c_str(), "-lineinfo"};
  const nvrtcResult rc = rt->compile(prog, 6, opts);
  size_t ls = 0;
  if (rt->log_size(prog, &ls) == NVRTC_SUCCESS && ls > 1) { log.resize(ls); rt->log(prog, &log[0]); }
  if (rc != NVRTC_SUCCESS) {
    log = std::string("NVRTC: ") + rt->errstr(rc) + "\n" + log;
    rt->destroy(&prog);
    return -1;
  }
  size_t cs = 0;
  rt->cubin_size(prog, &cs);
  cubin.resize(cs);
  rt->cubin(prog, cubin.data());
  rt->destroy(&prog);
  return 0;
}

// Compiled + loaded module for (device, n, m, source); cached for the life of the process.
int32_t custom_get(const char* user_src, int n, int m, bool user_cost, int device, CustomModule* out, std::string& err) {
  const std::string key = std::to_string(device) + "|" + std::to_string(n) + "|" + std::to_string(m) + "|" + (user_cost ? "c|" : "d|") + user_src;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) { *out = it->second; return 0; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { err = "cudaGetDeviceProperties failed"; return -1; }
  const std::string arch = "sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + (prop.major >= 9 ? "a" : "");
  std::vector<char> cubin;
  if (custom_compile(user_src, n, m, user_cost, arch.c_str(), cubin, err) != 0) return -1;
  CustomModule mod;
  cudaLibrary_t lib;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) { err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); return -1; }
  mod.lib = lib;
  struct { cudaKernel_t* k; const char* name; } want[] = {{&mod.bwd, "ilqr_bwd_custom"}, {&mod.fwd, "ilqr_fwd_custom"},
                                                           {&mod.rollout, "ilqr_rollout_init_custom"},
                                                           {&mod.advance, "ilqr_mpc_advance_custom"}};
  for (auto& w : want) {
    e = cudaLibraryGetKernel(w.k, lib, w.name);
    if (e != cudaSuccess) { err = std::string("cudaLibraryGetKernel(") + w.name + "): " + cudaGetErrorString(e); return -1; }
  }
  g_cache[key] = mod;
  *out = mod;
  return 0;
}

void launch_bwd_custom(const CustomModule& mod, const DevState& st, const CustomP& mp, const CostP& cost, cudaStream_t s) {
  if (st.nslots <= 0) return;
  void* args[] = {(void*)&st, (void*)&mp, (void*)&cost};
  cudaLaunchKernel((const void*)mod.bwd, dim3(grid_for(st.nslots, 4)), dim3(128), args, 0, s);
}
void launch_fwd_custom(const CustomModule& mod, const DevState& st, const CustomP& mp, const CostP& cost, cudaStream_t s) {
  if (st.nslots <= 0) return;
  void* args[] = {(void*)&st, (void*)&mp, (void*)&cost};
  cudaLaunchKernel((const void*)mod.fwd, dim3(grid_for(st.nslots, 128)), dim3(128), args, 0, s);
}
void launch_rollout_init_custom(const CustomModule& mod, const DevState& st, const CustomP& mp, const double* d_x0,
                                cudaStream_t s) {
  if (st.nslots <= 0) return;
  void* args[] = {(void*)&st, (void*)&mp, (void*)&d_x0};
  cudaLaunchKernel((const void*)mod.rollout, dim3(grid_for(st.nslots, 128)), dim3(128), args, 0, s);
}
void launch_mpc_advance_custom(const CustomModule& mod, const CustomP& mp, const double* out_u, double* plant,
                               double* u_applied, int B, int H, cudaStream_t s) {
  void* args[] = {(void*)&mp, (void*)&out_u, (void*)&plant, (void*)&u_applied, (void*)&B, (void*)&H};
  cudaLaunchKernel((const void*)mod.advance, dim3(grid_for(B, 128)), dim3(128), args, 0, s);
}

}  // namespace ilqr
