// User-supplied target densities at the speed of the built-in ones: the sampler kernel is compiled AT RUN TIME (NVRTC,
// sm_100a) around a target type written by the user, loaded with cudaLibraryLoadData, and launched through the very same
// host code as the built-in kernels (launch_warp_kernel / launch_kernel take any kernel handle).
//
// The reference accepts any Python callable `logp_dlogp_func(q) -> (logp, dlogp)` (base_hmc.py:34) and calls it once per
// leapfrog (integration.py:62,115).  On the GPU a Python callable means callback mode (lmc_callback.cu: one launch group
// per gradient); a density written as a Target (protocol: lmc_device.cuh, "built-in target densities") is evaluated
// INSIDE the sampler kernel instead -- registers to registers, no launch, no HBM round trip.
//
// NVRTC is loaded lazily (dlopen) so that the library itself loads on machines without it.
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "lmc_inst_warp.cuh"

namespace lmc {
int check_sampler_args(const lmc_sampler_args* a, int kind, bool check_target);  // lmc_sampler.cu

namespace {
thread_local std::string g_user_log;

struct Nvrtc {
  void* h = nullptr;
  decltype(&nvrtcCreateProgram) create = nullptr;
  decltype(&nvrtcDestroyProgram) destroy = nullptr;
  decltype(&nvrtcCompileProgram) compile = nullptr;
  decltype(&nvrtcAddNameExpression) add_name = nullptr;
  decltype(&nvrtcGetLoweredName) lowered = nullptr;
  decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
  decltype(&nvrtcGetProgramLog) log = nullptr;
  decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
  decltype(&nvrtcGetCUBIN) cubin = nullptr;
  decltype(&nvrtcGetErrorString) err = nullptr;
  bool load() {
    if (h) return true;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "libnvrtc.so.13"}) {
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) return false;
#define LMC_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(h, #sym)); if (!field) return false;
    LMC_SYM(create, nvrtcCreateProgram) LMC_SYM(destroy, nvrtcDestroyProgram) LMC_SYM(compile, nvrtcCompileProgram)
    LMC_SYM(add_name, nvrtcAddNameExpression) LMC_SYM(lowered, nvrtcGetLoweredName)
    LMC_SYM(log_size, nvrtcGetProgramLogSize) LMC_SYM(log, nvrtcGetProgramLog) LMC_SYM(cubin_size, nvrtcGetCUBINSize)
    LMC_SYM(cubin, nvrtcGetCUBIN) LMC_SYM(err, nvrtcGetErrorString)
#undef LMC_SYM
    return true;
  }
};
Nvrtc g_nvrtc;

bool read_file(const std::string& path, std::string* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char buf[1 << 16];
  size_t n;
  out->clear();
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, n);
  fclose(f);
  return true;
}
bool write_file_atomic(const std::string& path, const std::string& data) {
  const std::string tmp = path + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return false;
  const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  return ok && rename(tmp.c_str(), path.c_str()) == 0;
}
}  // namespace
}  // namespace lmc

struct lmc_user_kernel {
  cudaLibrary_t lib;
  cudaKernel_t kern;
  int warp;      // 1: sampler_warp_kernel<T, NP, B, 1, MINB, TAPE>; 0: sampler_kernel<T, G, NP, KIND>
  int kind, tape, NP, B, G;
  int stage_vecs;  // warp kernel: parameter vectors the kernel stages in shared memory (warp_stage_probe of the module)
};

extern "C" const char* lmc_user_kernel_log(void) { return lmc::g_user_log.c_str(); }

extern "C" int lmc_user_kernel_build(const char* source, const char* type_name, int32_t kind, int32_t ndim, int32_t chunk,
                                     int32_t tape, const char* const* include_dirs, int32_t n_include_dirs,
                                     const char* cache_path, lmc_user_kernel** out) {
  using namespace lmc;
  g_user_log.clear();
  if (!source || !type_name || !out || ndim < 1 || (kind != KIND_NUTS && kind != KIND_HMC)) return LMC_ERR_BADARG;
  lmc_user_kernel k = {};
  k.kind = kind;
  k.tape = tape ? 1 : 0;
  char name_expr[512], probe_expr[512] = "";
  const char* header;
  if (kind == KIND_NUTS && (ndim + 1) / 2 <= 128) {
    if (!pick_warp_shape(ndim, chunk, &k.NP, &k.B)) return LMC_ERR_UNSUPPORTED;
    int minb = 8;
#define LMC_X(n, bb, mb) if (k.NP == n && k.B == bb) minb = mb;
    LMC_WARP_SHAPES(LMC_X)
#undef LMC_X
    k.warp = 1;
    header = "lmc_sampler_warp.cuh";
    snprintf(name_expr, sizeof(name_expr), "lmc::sampler_warp_kernel<%s, %d, %d, 1, %d, %s>", type_name, k.NP, k.B, minb,
             k.tape ? "true" : "false");
    snprintf(probe_expr, sizeof(probe_expr), "&lmc::warp_stage_probe<%s, %d, %d>", type_name, k.NP, k.B);
  } else {
    Shape s;
    if (!pick_shape(ndim, 0, &s)) return LMC_ERR_UNSUPPORTED;
    k.G = s.G;
    k.NP = s.NP;
    k.warp = 0;
    header = "lmc_sampler.cuh";
    snprintf(name_expr, sizeof(name_expr), "lmc::sampler_kernel<%s, %d, %d, %d>", type_name, k.G, k.NP, kind);
  }

  std::string cubin, lowered, lowered_probe;  // the .name file of a cached cubin: kernel name [newline probe name]
  const std::string cpath = cache_path ? cache_path : "";
  if (cpath.empty() || !read_file(cpath, &cubin) || !read_file(cpath + ".name", &lowered) || cubin.empty()) {
    if (!g_nvrtc.load()) {
      g_user_log = "libnvrtc.so.12 could not be loaded (needed to compile user targets at run time)";
      return LMC_ERR_UNSUPPORTED;
    }
    std::string program = std::string("#include \"") + header + "\"\n" + source + "\n";
    nvrtcProgram prog;
    if (g_nvrtc.create(&prog, program.c_str(), "lmc_user_target.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
      return LMC_ERR_LAUNCH;
    g_nvrtc.add_name(prog, name_expr);
    if (probe_expr[0]) g_nvrtc.add_name(prog, probe_expr);
    // same code generation as the library build (__graft_entry__.py): un-fused multiply-add, sm_100a
    // -default-device: unannotated declarations (the C prototypes of lmc_b200.h, helpers in the user source) are device code
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "--generate-line-info",
                                     "-default-device"};
    for (int i = 0; i < n_include_dirs; ++i) opts.push_back(std::string("-I") + include_dirs[i]);
    std::vector<const char*> copts;
    for (auto& o : opts) copts.push_back(o.c_str());
    const nvrtcResult rc = g_nvrtc.compile(prog, (int)copts.size(), copts.data());
    size_t ls = 0;
    g_nvrtc.log_size(prog, &ls);
    if (ls > 1) {
      g_user_log.resize(ls);
      g_nvrtc.log(prog, &g_user_log[0]);
    }
    if (rc != NVRTC_SUCCESS) {
      g_user_log += std::string("\nnvrtc: ") + g_nvrtc.err(rc);
      g_nvrtc.destroy(&prog);
      return LMC_ERR_BADARG;
    }
    const char* low = nullptr;
    if (g_nvrtc.lowered(prog, name_expr, &low) != NVRTC_SUCCESS || !low) {
      g_nvrtc.destroy(&prog);
      return LMC_ERR_LAUNCH;
    }
    lowered = low;
    if (probe_expr[0]) {
      const char* lowp = nullptr;
      if (g_nvrtc.lowered(prog, probe_expr, &lowp) == NVRTC_SUCCESS && lowp) lowered += std::string("\n") + lowp;
    }
    size_t cs = 0;
    g_nvrtc.cubin_size(prog, &cs);
    cubin.resize(cs);
    g_nvrtc.cubin(prog, &cubin[0]);
    g_nvrtc.destroy(&prog);
    if (!cpath.empty()) {
      write_file_atomic(cpath, cubin);
      write_file_atomic(cpath + ".name", lowered);
    }
  }
  const size_t nl = lowered.find('\n');
  if (nl != std::string::npos) {
    lowered_probe = lowered.substr(nl + 1);
    lowered.resize(nl);
  }
  LMC_CUDA(cudaLibraryLoadData(&k.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  LMC_CUDA(cudaLibraryGetKernel(&k.kern, k.lib, lowered.c_str()));
  k.stage_vecs = 0;
  if (!lowered_probe.empty()) {
    void* dptr = nullptr;
    size_t bytes = 0;
    LMC_CUDA(cudaLibraryGetGlobal(&dptr, &bytes, k.lib, lowered_probe.c_str()));
    if (bytes != sizeof(int)) return LMC_ERR_LAUNCH;
    LMC_CUDA(cudaMemcpy(&k.stage_vecs, dptr, sizeof(int), cudaMemcpyDeviceToHost));
  }
  *out = new lmc_user_kernel(k);
  return LMC_OK;
}

extern "C" int lmc_user_kernel_destroy(lmc_user_kernel* k) {
  if (!k) return LMC_OK;
  cudaLibraryUnload(k->lib);
  delete k;
  return LMC_OK;
}

extern "C" int lmc_user_sample(lmc_user_kernel* k, const lmc_sampler_args* a, const void* target_bytes) {
  using namespace lmc;
  if (!k || !a || !target_bytes) return LMC_ERR_BADARG;
  const int rc = check_sampler_args(a, k->kind, false);
  if (rc != LMC_OK) return rc;
  if ((a->rng.mode == LMC_RNG_TAPE) != (k->tape != 0) && k->warp) return LMC_ERR_BADARG;  // built for the other RNG mode
  if (a->n_chains == 0 || a->n_trans == 0) return LMC_OK;
  const void* kern = reinterpret_cast<const void*>(k->kern);
  if (k->warp) {
    int NP = 0, B = 0;
    if (!pick_warp_shape(a->ndim, k->B, &NP, &B) || NP != k->NP) return LMC_ERR_BADARG;  // built for another ndim
#define LMC_X(n, bb, mb) \
  if (k->NP == n && k->B == bb) return launch_warp_kernel<n, bb, 1>(kern, *a, target_bytes, k->stage_vecs);
    LMC_WARP_SHAPES(LMC_X)
#undef LMC_X
    return LMC_ERR_UNSUPPORTED;
  }
  Shape s;
  if (!pick_shape(a->ndim, 0, &s) || s.G != k->G || s.NP != k->NP) return LMC_ERR_BADARG;
#define LMC_X(g, np, mc)                                                                         \
  if (k->G == g && k->NP == np)                                                                  \
    return k->kind == KIND_NUTS ? launch_kernel<g, np, KIND_NUTS>(kern, *a, target_bytes)        \
                                : launch_kernel<g, np, KIND_HMC>(kern, *a, target_bytes);
  LMC_SHAPES(LMC_X)
#undef LMC_X
  return LMC_ERR_UNSUPPORTED;
}
