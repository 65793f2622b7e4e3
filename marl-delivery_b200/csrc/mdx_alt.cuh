// The other featurisers of the same boundary: IDQ / legacy-qmix convert_state
// (IDQ/networks.py:112-217 == qmix/networks.py:243-348) and the legacy 7-channel global tensor
// (qmix/networks.py:350-468), for all listed envs and all robots in one launch.  One warp per environment;
// 0/1 channels are bit planes in shared memory, the urgency channel is a float plane built with an integer
// atomicMax (non-negative floats order like their bit patterns); outputs leave as 128-bit streaming stores.
#pragma once
#include "mdx_kernels.cuh"

namespace mdx {

struct AltArgs {
  float* obs6;   // [n, R, 6, H, W]
  float* g7;     // [n, 7, H, W]
  int32_t smem_per_warp;
};

// urgency of a waiting package whose start_time has passed (IDQ/networks.py:172-179): float64 quotient,
// clamped, then stored into a float32 array
__device__ __forceinline__ float alt_urgency(int t, int st, int dl) {
  if (dl > st) {
    double u = __ddiv_rn((double)(t - st), (double)(dl - st));
    u = u > 1.0 ? 1.0 : u;
    return __double2float_rn(u);
  }
  return dl == st ? 1.0f : 0.0f;
}

__device__ __forceinline__ float4 nib4(uint32_t n) {
  return make_float4((n & 1u) ? 1.f : 0.f, (n & 2u) ? 1.f : 0.f, (n & 4u) ? 1.f : 0.f, (n & 8u) ? 1.f : 0.f);
}

__global__ void __launch_bounds__(128) mdx_alt_featurize_kernel(const __grid_constant__ KParams p, const AltArgs a) {
  extern __shared__ __align__(16) unsigned char alt_smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int widx = blockIdx.x * (blockDim.x >> 5) + wib;
  if (widx >= p.n) return;
  const int env = p.env_ids ? __ldg(p.env_ids + widx) : widx;
  const int R = p.R, P = p.P, HW = p.HW, W32 = p.W32;
  float* urg = reinterpret_cast<float*>(alt_smem + (size_t)wib * a.smem_per_warp);  // [HW rounded up to 4]
  uint32_t* planes = reinterpret_cast<uint32_t*>(urg + ((HW + 3) & ~3));            // [6][W32]
  MapDev map = p.map0;
  if (p.env_map) map = p.maps[__ldg(p.env_map + env)];
  const int t = p.t[env];
  const uint16_t* S = p.cur + (size_t)env * 4 * p.Pp8;
  const uint16_t *G = S + p.Pp8, *T0 = S + 2 * p.Pp8, *DL = S + 3 * p.Pp8;
  const uint8_t* status = p.status + (size_t)env * p.Pp16;
  const bool active = lane < R;
  const uint32_t rob = active ? p.robots[(size_t)env * R + lane] : 0u;
  const int cell = active ? (int)(rob & 0xffffu) : 0xFFFF, carry = (int)(rob >> 16);
  for (int i = lane; i < ((HW + 3) & ~3); i += 32) urg[i] = 0.0f;
  for (int i = lane; i < W32; i += 32) {
    planes[i] = __ldg(map.grid_bits + i);
#pragma unroll
    for (int k = 1; k < 6; ++k) planes[k * W32 + i] = 0u;
  }
  __syncwarp();
  // planes: 0 grid, 1 robots, 2 robots that carry, 3 waiting starts, 4 waiting targets, 5 in-transit targets
  int own_tgt = 0xFFFF;
  if (active) {
    atomicOr(&planes[W32 + (cell >> 5)], 1u << (cell & 31));
    if (carry != 0) {
      atomicOr(&planes[2 * W32 + (cell >> 5)], 1u << (cell & 31));
      const uint8_t st = status[carry - 1];
      if (st == ST_WAITING || st == ST_IN_TRANSIT) own_tgt = __ldg(G + carry - 1);  // "id in persistent_packages"
    }
  }
  for (int j = lane; j < P; j += 32) {
    const uint8_t st = status[j];
    if (st != ST_WAITING && st != ST_IN_TRANSIT) continue;
    const int t0 = __ldg(T0 + j);
    if (t < t0) continue;
    const int tc = __ldg(G + j);
    if (st == ST_WAITING) {
      const int sc = __ldg(S + j);
      atomicOr(&planes[3 * W32 + (sc >> 5)], 1u << (sc & 31));
      atomicOr(&planes[4 * W32 + (tc >> 5)], 1u << (tc & 31));
      atomicMax(reinterpret_cast<int*>(urg + sc), __float_as_int(alt_urgency(t, t0, (int)__ldg(DL + j))));
    } else {
      atomicOr(&planes[5 * W32 + (tc >> 5)], 1u << (tc & 31));
    }
  }
  __syncwarp();
  const bool vec4 = (HW & 3) == 0;
  // ---- legacy global tensor: grid, robots, carrying robots, waiting starts, waiting targets, in-transit targets, urgency
  if (a.g7 != nullptr) {
    float* out = a.g7 + (size_t)widx * 7 * HW;
    if (vec4) {
      float4* out4 = reinterpret_cast<float4*>(out);
      for (int q = lane; q < (7 * HW) >> 2; q += 32) {
        const uint32_t e = (uint32_t)q << 2;
        const uint32_t ch = __umulhi(e, p.mul_plane), cl = e - ch * HW;
        float4 v;
        if (ch == 6) v = *reinterpret_cast<const float4*>(urg + cl);
        else v = nib4((planes[ch * W32 + (cl >> 5)] >> (cl & 31u)) & 15u);
        __stcs(out4 + q, v);
      }
    } else {
      for (int e = lane; e < 7 * HW; e += 32) {
        const uint32_t ch = __umulhi((uint32_t)e, p.mul_plane), cl = (uint32_t)e - ch * HW;
        __stcs(out + e, ch == 6 ? urg[cl] : (((planes[ch * W32 + (cl >> 5)] >> (cl & 31u)) & 1u) ? 1.0f : 0.0f));
      }
    }
  }
  // ---- per-robot observation: grid, urgency*, waiting starts* (* only when not carrying), others, self, own target
  if (a.obs6 != nullptr) {
    for (int i = 0; i < R; ++i) {
      const int ci = __shfl_sync(FULL, cell, i), ki = __shfl_sync(FULL, carry, i), ti = __shfl_sync(FULL, own_tgt, i);
      float* out = a.obs6 + ((size_t)widx * R + i) * 6 * HW;
      if (vec4) {
        float4* out4 = reinterpret_cast<float4*>(out);
        for (int q = lane; q < (6 * HW) >> 2; q += 32) {
          const uint32_t e = (uint32_t)q << 2;
          const uint32_t ch = __umulhi(e, p.mul_plane), cl = e - ch * HW;
          const uint32_t self = ((uint32_t)ci >> 2) == (cl >> 2) ? (1u << (ci & 3)) : 0u;
          float4 v;
          if (ch == 0) v = nib4((planes[cl >> 5] >> (cl & 31u)) & 15u);
          else if (ch == 1) v = ki == 0 ? *reinterpret_cast<const float4*>(urg + cl) : make_float4(0.f, 0.f, 0.f, 0.f);
          else if (ch == 2) v = nib4(ki == 0 ? (planes[3 * W32 + (cl >> 5)] >> (cl & 31u)) & 15u : 0u);
          else if (ch == 3) v = nib4(((planes[W32 + (cl >> 5)] >> (cl & 31u)) & 15u) & ~self);
          else if (ch == 4) v = nib4(self);
          else v = nib4(((uint32_t)ti >> 2) == (cl >> 2) ? (1u << (ti & 3)) : 0u);
          __stcs(out4 + q, v);
        }
      } else {
        for (int e = lane; e < 6 * HW; e += 32) {
          const uint32_t ch = __umulhi((uint32_t)e, p.mul_plane), cl = (uint32_t)e - ch * HW;
          float v;
          if (ch == 0) v = ((planes[cl >> 5] >> (cl & 31u)) & 1u) ? 1.0f : 0.0f;
          else if (ch == 1) v = ki == 0 ? urg[cl] : 0.0f;
          else if (ch == 2) v = (ki == 0 && ((planes[3 * W32 + (cl >> 5)] >> (cl & 31u)) & 1u)) ? 1.0f : 0.0f;
          else if (ch == 3) v = ((int)cl != ci && ((planes[W32 + (cl >> 5)] >> (cl & 31u)) & 1u)) ? 1.0f : 0.0f;
          else if (ch == 4) v = (int)cl == ci ? 1.0f : 0.0f;
          else v = (int)cl == ti ? 1.0f : 0.0f;
          __stcs(out + e, v);
        }
      }
    }
  }
}

}  // namespace mdx
