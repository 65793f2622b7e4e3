// Run-time shape specialisation of the fused step + featurise kernel (MDX_CFG_JIT_SHAPE).
//
// Round 1 baked ONE shape (the C3 benchmark: map1, 5 robots, <= 56 packages, K = 4 / 5, MR = MP = 100) into a kernel variant whose
// loops unroll and whose index arithmetic folds (kFix = 1: 1 896 instead of 3 000 SASS instructions, +9 % on B200).  Here the same
// folding is available to ANY instance: at mdx_create the kernel source that sits next to libmdx.so is compiled by NVRTC for sm_100a
// with the instance's (H, W, R, P, T, K_o, K_p, MR, MP) and its shared-memory carve-up as macros (mdx_kernels.cuh: specialise<3>),
// the cubin is loaded through the driver API and cached on disk, and mdx_step_featurize launches it instead of the run-time-shape
// variant.  libnvrtc / libcuda are opened with dlopen, so libmdx.so has no link-time dependency on them; whenever anything here
// fails the instance silently keeps the precompiled kernels (mdx_jit_status tells which one runs).
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "mdx_host.h"

namespace mdx {
namespace {

struct Api {
  bool ok = false;
  // nvrtc
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  // driver
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
};

template <typename F>
bool sym(void* lib, const char* name, F& out) {
  out = reinterpret_cast<F>(dlsym(lib, name));
  return out != nullptr;
}

Api& api() {
  static Api a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  void* rtc = nullptr;
  for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"})
    if ((rtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
  void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!rtc || !drv) return a;
  a.ok = sym(rtc, "nvrtcCreateProgram", a.CreateProgram) && sym(rtc, "nvrtcDestroyProgram", a.DestroyProgram) &&
         sym(rtc, "nvrtcAddNameExpression", a.AddNameExpression) && sym(rtc, "nvrtcCompileProgram", a.CompileProgram) &&
         sym(rtc, "nvrtcGetLoweredName", a.GetLoweredName) && sym(rtc, "nvrtcGetCUBINSize", a.GetCUBINSize) &&
         sym(rtc, "nvrtcGetCUBIN", a.GetCUBIN) && sym(rtc, "nvrtcGetProgramLogSize", a.GetProgramLogSize) &&
         sym(rtc, "nvrtcGetProgramLog", a.GetProgramLog) && sym(drv, "cuModuleLoadData", a.ModuleLoadData) &&
         sym(drv, "cuModuleUnload", a.ModuleUnload) && sym(drv, "cuModuleGetFunction", a.ModuleGetFunction) &&
         sym(drv, "cuFuncSetAttribute", a.FuncSetAttribute) && sym(drv, "cuLaunchKernel", a.LaunchKernel);
  return a;
}

std::string source_dir() {
  Dl_info info;
  if (dladdr(reinterpret_cast<const void*>(&source_dir), &info) == 0 || info.dli_fname == nullptr) return ".";
  std::string p(info.dli_fname);
  const size_t k = p.rfind('/');
  return k == std::string::npos ? "." : p.substr(0, k);
}

uint64_t fnv(uint64_t h, const void* p, size_t n) {
  const unsigned char* c = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) h = (h ^ c[i]) * 1099511628211ull;
  return h;
}

std::string slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

}  // namespace

// Source of the specialised translation unit + the kernel's name expression for this instance.
void jit_source(const mdx_env* e, std::string* src, std::string* name_expr, int* blocks_out) {
  const KParams& b = e->base;
  const int n4p = b.HW >> 2;
  int ks = 0;
  if ((b.HW & 3) == 0 && !e->force_generic) {
    const int s = (n4p + 31) / 32;
    ks = s <= 1 ? 1 : s <= 2 ? 2 : s <= 4 ? 4 : 0;
  }
  const bool small = e->P <= 64;
  // resident CTAs per SM the variant is compiled for: what shared memory allows, capped like the precompiled variants
  // (four-slice planes with long package lists run best with 96 registers = 5 CTAs, everything else with 72 = 7)
  const size_t cta_smem = 256 + (size_t)4 * b.smem_per_warp + 1024;
  // measured on B200 (profiles/r2_experiments.txt): K = 100 rows 5 / 6 / 7 / 8 / 9 CTAs -> 0.830 / 0.862 / 0.885 / 0.886 / 0.873 of peak,
  // C4 (four-slice planes, 100 packages) 5 / 6 / 7 -> 0.903 / 0.925 / 0.883
  int blocks = (int)std::min<size_t>((ks == 4 && !small) ? 6 : 7, std::max<size_t>(1, (227 * 1024) / cta_smem));
  if (const char* s = getenv("MDX_JIT_BLOCKS")) blocks = std::max(1, std::min(9, atoi(s)));
  *blocks_out = blocks;
  // the interleaved package selection of the baked C3 variant (R independent arg-min chains in flight, winners kept in lanes) for every
  // shape it fits: few robots, R * slots <= 32 lanes, short package lists, all vec rows of an env staged at once
  const bool interleave = small && b.R <= 8 && b.sel_k >= 1 && b.R * b.sel_k <= 32 && b.vec_rows_per_flush == b.R && !getenv("MDX_JIT_NO_INTERLEAVE");
  if (interleave && ks == 1 && !getenv("MDX_JIT_BLOCKS")) blocks = (int)std::min<size_t>(9, std::max<size_t>(1, (227 * 1024) / cta_smem));  // like the baked variant: 56 registers
  *blocks_out = blocks;
  std::ostringstream o;
  o << "#define MDX_JIT_SHAPE 1\n#define MDX_JIT_BLOCKS " << blocks << "\n#define MDX_J_INTERLEAVE " << (interleave ? 1 : 0) << "\n#define MDX_J_NR " << b.R << "\n";
  auto def = [&](const char* k, long long v) { o << "#define MDX_J_" << k << " " << v << "\n"; };
  auto defu = [&](const char* k, uint32_t v) { o << "#define MDX_J_" << k << " " << v << "u\n"; };
  def("R", b.R), def("P", b.P), def("T", b.T), def("PP8", b.Pp8), def("PP16", b.Pp16), def("H", b.H_out), def("W", b.W_out);
  def("KO", b.K_o), def("KP", b.K_p), def("SELK", b.sel_k), def("MR", b.MR), def("MP", b.MP), def("RPF", b.vec_rows_per_flush), def("GCH", b.gvec_chunk);
  def("STAGEB", b.stage_b), def("WBITS", b.key_wbits), def("NRANKS", b.nranks_max), def("SELRP", b.sel_rp), def("TQLEN", b.tq_len);
  defu("MUL_WOUT", b.mul_wout), defu("MUL_R", b.mul_r), defu("MUL_KSEL", b.mul_ksel), defu("MUL_PLANE", b.mul_plane);
  def("OFF_SCHED", b.off_sched), def("OFF_STATUS", b.off_status), def("OFF_WLIST", b.off_wlist), def("OFF_WKEY", b.off_wkey);
  def("OFF_RINFO", b.off_rinfo), def("OFF_ALIST", b.off_alist), def("OFF_SEL", b.off_sel), def("OFF_MASKS", b.off_masks);
  def("OFF_STAGE", b.off_stage), def("SPW", b.smem_per_warp);
  o << "#include \"mdx_kernels.cuh\"\n";
  *src = o.str();
  std::ostringstream n;
  n << "mdx::mdx_env_kernel<3, 32, " << ks << ", " << (small ? "true" : "false") << ", 3, false>";
  *name_expr = n.str();  // (the featurise-only kernel of the same shape: the same expression with mode 2, see jit_build)
}

// Compiles (or fetches from the disk cache) and loads the specialised fused kernel of `e`.  MDX_OK: e->jit_func is set.
int jit_build(mdx_env* e) {
  if (e->jit_state != 0) return e->jit_state > 0 ? MDX_OK : MDX_E_INVALID;
  e->jit_state = -1;
  Api& a = api();
  if (!a.ok) return fail(MDX_E_INVALID, "run-time specialisation unavailable: libnvrtc.so.12 / libcuda.so.1 could not be opened");
  if (e->stale || e->base.shaping_bfs) return fail(MDX_E_INVALID, "run-time specialisation covers the default (env truth, Manhattan) path only");
  if (baked_c3_shape(e) && !getenv("MDX_JIT_FORCE")) {  // its hand-specialised variant (interleaved selection, 9 CTAs/SM) is faster: 0.89 vs 0.79
    e->jit_state = 0;
    return MDX_OK;
  }
  std::string src, name;
  int blocks = 7;
  jit_source(e, &src, &name, &blocks);
  std::string name2 = name, name5 = name;  // mdx_featurize (mode 2) and the fused IDQ iteration (mode 5) of the same shape
  name2.replace(name2.find("<3,"), 3, "<2,");
  name5.replace(name5.find("<3,"), 3, "<5,");
  const bool with5 = name.find("<3, 32, 0,") == std::string::npos;  // the fused IDQ kernel exists for plane-sliced maps only
  const std::string dir = source_dir();
  // cache key: the specialised source + every header it includes (by content) + the target
  uint64_t h = fnv(1469598103934665603ull, src.data(), src.size());
  for (const char* f : {"mdx_kernels.cuh", "mdx_internal.h", "mdx_sampler.cuh"}) {
    const std::string body = slurp(dir + "/" + f);
    if (body.empty()) return fail(MDX_E_INVALID, "run-time specialisation: kernel source %s/%s not found", dir.c_str(), f);
    h = fnv(h, body.data(), body.size());
  }
  h = fnv(h, name.data(), name.size());
  h = fnv(h, "v3", 2);  // (layout of the cached .name file: three lowered names)
  const char* cdir = getenv("MDX_JIT_CACHE");
  const std::string cache_dir = cdir ? cdir : "/tmp/mdx_jit_cache";
  mkdir(cache_dir.c_str(), 0777);
  char key[64];
  snprintf(key, sizeof(key), "%016llx", (unsigned long long)h);
  const std::string cubin_path = cache_dir + "/" + key + ".cubin", name_path = cache_dir + "/" + key + ".name";
  DeviceGuard g(e->device);
  cudaFree(nullptr);  // make sure the primary context exists and is current
  const size_t smem = 256 + (size_t)4 * e->base.smem_per_warp;
  // two attempts: a cached cubin first (if any), then -- should that one not load, e.g. a file truncated by a crash -- a fresh compilation
  for (int attempt = 0; attempt < 2; ++attempt) {
    std::string cubin = attempt == 0 ? slurp(cubin_path) : std::string(), lowered = attempt == 0 ? slurp(name_path) : std::string();
    const bool from_cache = !cubin.empty() && !lowered.empty();
    if (!from_cache) {
      nvrtcProgram prog = nullptr;
      if (a.CreateProgram(&prog, src.c_str(), "mdx_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return fail(MDX_E_INVALID, "nvrtcCreateProgram failed");
      a.AddNameExpression(prog, name.c_str());
      a.AddNameExpression(prog, name2.c_str());
      if (with5) a.AddNameExpression(prog, name5.c_str());
      const std::string inc = "-I" + dir;
      const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-lineinfo", inc.c_str()};
      const nvrtcResult rc = a.CompileProgram(prog, 5, opts);
      if (rc != NVRTC_SUCCESS) {
        size_t n = 0;
        a.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) a.GetProgramLog(prog, &log[0]);
        a.DestroyProgram(&prog);
        return fail(MDX_E_INVALID, "run-time specialisation: NVRTC compilation failed: %.380s", log.c_str());
      }
      const char *low = nullptr, *low2 = nullptr;
      size_t n = 0;
      if (a.GetLoweredName(prog, name.c_str(), &low) != NVRTC_SUCCESS || a.GetLoweredName(prog, name2.c_str(), &low2) != NVRTC_SUCCESS ||
          a.GetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) {
        a.DestroyProgram(&prog);
        return fail(MDX_E_INVALID, "run-time specialisation: no cubin / lowered name from NVRTC");
      }
      lowered = std::string(low) + "\n" + low2 + "\n";
      if (with5) {
        const char* low5 = nullptr;
        if (a.GetLoweredName(prog, name5.c_str(), &low5) == NVRTC_SUCCESS && low5) lowered += low5;
      }
      cubin.resize(n);
      a.GetCUBIN(prog, &cubin[0]);
      a.DestroyProgram(&prog);
      {  // best effort: a later process (or instance) of the same shape skips the compilation.  Written under a per-process name and
         // renamed, so that ranks compiling the same shape at the same time never see (or produce) a partial file.
        const std::string tag = "." + std::to_string((long long)getpid());
        std::ofstream(cubin_path + tag, std::ios::binary).write(cubin.data(), (std::streamsize)cubin.size());
        std::ofstream(name_path + tag, std::ios::binary) << lowered;
        rename((name_path + tag).c_str(), name_path.c_str());
        rename((cubin_path + tag).c_str(), cubin_path.c_str());
      }
    }
    CUmodule mod = nullptr;
    CUfunction fn = nullptr, fn2 = nullptr;
    std::vector<std::string> lows;
    {
      std::stringstream ls(lowered);
      std::string ln;
      while (std::getline(ls, ln, '\n')) lows.push_back(ln);
    }
    const std::string l3 = lows.size() > 0 ? lows[0] : std::string(), l2 = lows.size() > 1 ? lows[1] : std::string(),
                      l5 = lows.size() > 2 ? lows[2] : std::string();
    CUfunction fn5 = nullptr;
    const int smem_attr = (int)std::max<size_t>(smem, 48 * 1024);
    if (!l2.empty() && a.ModuleLoadData(&mod, cubin.data()) == CUDA_SUCCESS && a.ModuleGetFunction(&fn, mod, l3.c_str()) == CUDA_SUCCESS &&
        a.ModuleGetFunction(&fn2, mod, l2.c_str()) == CUDA_SUCCESS && a.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem_attr) == CUDA_SUCCESS &&
        a.FuncSetAttribute(fn2, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem_attr) == CUDA_SUCCESS &&
        (!with5 || (!l5.empty() && a.ModuleGetFunction(&fn5, mod, l5.c_str()) == CUDA_SUCCESS &&
                    a.FuncSetAttribute(fn5, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem_attr) == CUDA_SUCCESS))) {
      e->jit_func5 = fn5;
      e->jit_module = mod;
      e->jit_func = fn;
      e->jit_func2 = fn2;
      e->jit_blocks = blocks;
      e->jit_state = 1;
      return MDX_OK;
    }
    if (mod) a.ModuleUnload(mod);
    if (!from_cache) return fail(MDX_E_CUDA, "run-time specialisation: the compiled module could not be loaded (%s)", lowered.c_str());
    remove(cubin_path.c_str());  // a bad cache entry: compile afresh
    remove(name_path.c_str());
  }
  return fail(MDX_E_CUDA, "run-time specialisation: module load failed");
}

void jit_release(mdx_env* e) {
  if (e->jit_module && api().ok) api().ModuleUnload(static_cast<CUmodule>(e->jit_module));
  e->jit_module = nullptr;
  e->jit_func = nullptr;
  e->jit_func2 = nullptr;
  e->jit_func5 = nullptr;
}

int jit_launch(mdx_env* e, const KParams& kp, cudaStream_t st, int mode) {
  KParams q = kp;
  q.smem_per_warp = e->base.smem_per_warp;
  void* params[] = {&q};
  const size_t smem = 256 + (size_t)4 * e->base.smem_per_warp;
  const unsigned grid = (unsigned)((kp.n + 3) / 4);
  const CUresult rc = api().LaunchKernel(static_cast<CUfunction>(mode == 2 ? e->jit_func2 : mode == 5 ? e->jit_func5 : e->jit_func), grid, 1, 1, 128, 1, 1, (unsigned)smem, reinterpret_cast<CUstream>(st),
                                         params, nullptr);
  if (rc != CUDA_SUCCESS) return fail(MDX_E_CUDA, "launch of the run-time-specialised kernel failed (CUresult %d)", (int)rc);
  return MDX_OK;
}

}  // namespace mdx
