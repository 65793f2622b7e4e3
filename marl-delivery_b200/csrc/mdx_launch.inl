// Variant selection and launch of mdx_env_kernel for ONE kernel mode (MDX_MODE), so that the three modes
// compile in parallel translation units.
#include <atomic>

#include "mdx_host.h"
#include "mdx_kernels.cuh"

namespace mdx {
namespace {

#ifndef MDX_TRK
#define MDX_TRK 0
#endif
constexpr bool kTrkTU = MDX_TRK != 0;  // this translation unit holds the stale-tracker variants (whole warps, run-time shapes)

int smem_for_mode(const mdx_env* e, int mode) {
  return ((mode & 6) || kTrkTU) ? e->base.smem_per_warp : e->base.off_alist;
}

template <int kMode, int LPE, int kS, bool kSmall, int kFix>
int launch_variant(mdx_env* e, const KParams& kp, cudaStream_t st) {
  const int spe = smem_for_mode(e, kMode);  // shared memory per environment
  const int threads = (kFix == 1 && (kMode & 2)) ? MDX_FIX_THREADS : 128;
  const int epb = threads / LPE;            // environments per CTA
  const size_t smem = (size_t)kLutBytes + (size_t)epb * spe;
  if (smem > 227 * 1024) return fail(MDX_E_INVALID, "per-CTA working set (%zu bytes) exceeds shared memory", smem);
  // The dynamic shared-memory limit is an attribute of this kernel FUNCTION on a device, shared by every mdx_env of the process:
  // the high-water mark lives in a static of this instantiation (one per function), indexed by device.
  static std::atomic<size_t> high_water[64];
  if (smem > high_water[e->device & 63].load(std::memory_order_acquire)) {
    int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(mdx_env_kernel<kMode, LPE, kS, kSmall, kFix, kTrkTU>), smem, e->carveout);
    if (rc != MDX_OK) return rc;
    size_t prev = high_water[e->device & 63].load(std::memory_order_relaxed);
    while (prev < smem && !high_water[e->device & 63].compare_exchange_weak(prev, smem)) {
    }
  }
  KParams q = kp;
  q.smem_per_warp = spe;
  const unsigned grid = (unsigned)((kp.n + epb - 1) / epb);
  // step-only launches are a few microseconds long: let consecutive launches on a stream overlap their launch latency and ramp
  // (programmatic dependent launch; the kernel waits for its predecessor before it touches the env state).  MDX_PDL=0 disables it,
  // MDX_PDL=2 enables it for the featurising modes as well.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (e->pdl >= 2 || (e->pdl == 1 && kMode == 1)) ? 1 : 0;
  if (cfg.numAttrs != 0 && e->pdl == 1) {
    // measured on B200, 4 096 envs, step only: eager launches 10.3 -> 8.6 us per step with the attribute, but the same launches captured
    // in a CUDA graph 7.7 -> 8.1 us: a graph already removes the launch latency, the programmatic edge only adds to it
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) cfg.numAttrs = 0;
  }
  CU(cudaLaunchKernelEx(&cfg, mdx_env_kernel<kMode, LPE, kS, kSmall, kFix, kTrkTU>, q));
  return MDX_OK;
}

template <int kMode, int LPE>
int launch_lpe(mdx_env* e, const KParams& kp, cudaStream_t st, int ks) {
  const bool small = e->P <= 64;
  if constexpr (kMode == 5) {  // step + IDQ featurisers: whole warps, plane-sliced maps only (the caller falls back to two launches)
    if constexpr (LPE == 32) {
      switch (ks) {
        case 1: return small ? launch_variant<5, 32, 1, true, 0>(e, kp, st) : launch_variant<5, 32, 1, false, 0>(e, kp, st);
        case 2: return small ? launch_variant<5, 32, 2, true, 0>(e, kp, st) : launch_variant<5, 32, 2, false, 0>(e, kp, st);
        case 4: return small ? launch_variant<5, 32, 4, true, 0>(e, kp, st) : launch_variant<5, 32, 4, false, 0>(e, kp, st);
        default: break;
      }
    }
    return MDX_E_UNSUPPORTED_SHAPE;
  } else {
  if (kMode == 1) return small ? launch_variant<1, LPE, 0, true, 0>(e, kp, st) : launch_variant<1, LPE, 0, false, 0>(e, kp, st);
  switch (ks) {
    case 1: return small ? launch_variant<kMode, LPE, 1, true, 0>(e, kp, st) : launch_variant<kMode, LPE, 1, false, 0>(e, kp, st);
    case 2: return small ? launch_variant<kMode, LPE, 2, true, 0>(e, kp, st) : launch_variant<kMode, LPE, 2, false, 0>(e, kp, st);
    case 4:
      if (small) return launch_variant<kMode, LPE, 4, true, 0>(e, kp, st);
      // four-slice planes with a long package list (C4, C5): the build with the looser register cap (MDX_NOFAT=1: the capped one)
      if constexpr (!kTrkTU) {
        if (LPE == 32 && kMode != 5 && !e->no_fat)
          return launch_variant<kMode, (LPE == 32 ? 32 : LPE), 4, false, (LPE == 32 ? 2 : 0)>(e, kp, st);
      }
      return launch_variant<kMode, LPE, 4, false, 0>(e, kp, st);
    default: return small ? launch_variant<kMode, LPE, 0, true, 0>(e, kp, st) : launch_variant<kMode, LPE, 0, false, 0>(e, kp, st);
  }
  }
}

// Lanes per environment: the smallest group (8, 16, 32 lanes) that holds all robots and, for the
// featurising modes, still has a plane-sliced fast path (a plane = H*W/4 float4 in <= 4 slices of LPE).
template <int kMode>
int launch_env_kernel(mdx_env* e, const KParams& kp, cudaStream_t st) {
  const int HW = e->base.HW, n4p = HW >> 2;
  auto slices = [&](int lpe) {
    if (e->force_generic || (HW & 3) != 0) return 0;
    const int s = (n4p + lpe - 1) / lpe;
    return s <= 1 ? 1 : s <= 2 ? 2 : s <= 4 ? 4 : 0;
  };
  // Measured on B200 (C3, 5 robots): 32 lanes per env 221-229 M env-steps/s, 16 lanes 194 M, 8 lanes 143 M --
  // packing several envs into a warp lowers the warp count and lengthens each warp's dependent chain, which
  // costs more than the amortised uniform instructions save.  Whole warps are therefore the default;
  // MDX_LPE=8|16 selects the packed variants (kept and tested).
  if constexpr (kMode == 5) {
    return launch_lpe<kMode, 32>(e, kp, st, slices(32));
  } else {
  if constexpr (kTrkTU) {
    return launch_lpe<kMode, 32>(e, kp, st, kMode == 1 ? 0 : slices(32));
  } else {
  int lpe = 32;
  // Step-only launches have no store phase: they are bound by instruction issue once the batch fills the GPU a few times over, and
  // packing four environments into a warp amortises the warp-uniform work (65 536 envs: 78 -> 60 us per step); small batches are
  // latency-bound and keep one warp per environment (4 096 envs: 7.6 us vs 13.3 us packed).
  if (kMode == 1 && e->R <= 8 && kp.n >= 32768) lpe = 8;
  if (e->lpe_override) lpe = std::max(e->R <= 8 ? 8 : e->R <= 16 ? 16 : 32, e->lpe_override);
  if (kMode != 1 && !e->force_generic && (HW & 3) == 0)
    while (lpe < 32 && slices(lpe) == 0) lpe *= 2;
  const int ks = kMode == 1 ? 0 : slices(lpe);
  // the headline MAPPO shape has a fully specialised instantiation
  if ((kMode & 2) && lpe == 32 && ks == 1 && baked_c3_shape(e))
    return launch_variant<kMode, 32, 1, true, 1>(e, kp, st);
  if (lpe == 8) return launch_lpe<kMode, 8>(e, kp, st, ks);
  if (lpe == 16) return launch_lpe<kMode, 16>(e, kp, st, ks);
  return launch_lpe<kMode, 32>(e, kp, st, ks);
  }
  }
}

}  // namespace

#define MDX_CAT2(a, b) a##b
#define MDX_CAT(a, b) MDX_CAT2(a, b)
#if MDX_TRK
int MDX_CAT(launch_env_kernel_trk_, MDX_MODE)(mdx_env* e, const KParams& kp, cudaStream_t st) { return launch_env_kernel<MDX_MODE>(e, kp, st); }
#else
int MDX_CAT(launch_env_kernel_, MDX_MODE)(mdx_env* e, const KParams& kp, cudaStream_t st) { return launch_env_kernel<MDX_MODE>(e, kp, st); }
#endif

}  // namespace mdx
