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

__device__ __forceinline__ float4 nib4(uint32_t n) {
  return make_float4((n & 1u) ? 1.f : 0.f, (n & 2u) ? 1.f : 0.f, (n & 4u) ? 1.f : 0.f, (n & 8u) ? 1.f : 0.f);
}

__global__ void __launch_bounds__(128) mdx_alt_featurize_kernel(const __grid_constant__ KParams p, const AltArgs a) {
  extern __shared__ __align__(16) unsigned char alt_smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int widx = blockIdx.x * (blockDim.x >> 5) + wib;
  if (widx >= p.n) return;
  const int env = p.env_ids ? __ldg(p.env_ids + widx) : p.env_base + widx;
  if ((unsigned)env >= (unsigned)p.n_envs) return;  // a bad caller-supplied id never becomes an out-of-bounds access (row left unwritten)
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

// Plane-sliced variant (H*W % 4 == 0 and H*W <= 4 * 32 * NS): a lane owns float4 number (lane + 32*s) of every plane.
// The six shared planes' nibbles and the urgency float4 are hoisted into registers once per env; a robot then costs six
// stores per slice, each a 16-entry nibble -> float4 table load (LDS.128) plus one streaming STG.128 -- no per-element
// channel dispatch, no divergence.  Same layout of the per-warp shared memory as the generic kernel, after the table.
template <int NS>
__global__ void __launch_bounds__(128) mdx_alt_featurize_sliced(const __grid_constant__ KParams p, const AltArgs a) {
  extern __shared__ __align__(16) unsigned char alt_smem[];
  float4* lut = reinterpret_cast<float4*>(alt_smem);
  if (threadIdx.x < 16) lut[threadIdx.x] = nib4(threadIdx.x);
  __syncthreads();
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int widx = blockIdx.x * (blockDim.x >> 5) + wib;
  if (widx >= p.n) return;
  const int env = p.env_ids ? __ldg(p.env_ids + widx) : p.env_base + widx;
  if ((unsigned)env >= (unsigned)p.n_envs) return;  // a bad caller-supplied id never becomes an out-of-bounds access (row left unwritten)
  const int R = p.R, P = p.P, HW = p.HW, W32 = p.W32;
  float* urg = reinterpret_cast<float*>(alt_smem + kLutBytes + (size_t)wib * a.smem_per_warp);
  uint32_t* planes = reinterpret_cast<uint32_t*>(urg + HW);
  MapDev map = p.map0;
  if (p.env_map) map = p.maps[__ldg(p.env_map + env)];
  // every global load of the env's state is issued before the first use
  const int t = p.t[env];
  const bool active = lane < R;
  const uint32_t rob = active ? p.robots[(size_t)env * R + lane] : 0u;
  const uint16_t* S = p.cur + (size_t)env * 4 * p.Pp8;
  const uint16_t *G = S + p.Pp8, *T0 = S + 2 * p.Pp8, *DL = S + 3 * p.Pp8;
  const uint8_t* status = p.status + (size_t)env * p.Pp16;
  const int cell = active ? (int)(rob & 0xffffu) : 0xFFFF, carry = (int)(rob >> 16);
  for (int i = lane; i < (HW >> 2); i += 32) reinterpret_cast<float4*>(urg)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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
  const int n4p = HW >> 2;
  uint32_t nib[NS][6];
  float4 u4[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int q = lane + 32 * s;
    const bool ok = q < n4p;
    const int wi = q >> 3;
    const uint32_t sh = (uint32_t)(q & 7) << 2;
#pragma unroll
    for (int k = 0; k < 6; ++k) nib[s][k] = ok ? (planes[k * W32 + wi] >> sh) & 15u : 0u;
    u4[s] = ok ? reinterpret_cast<const float4*>(urg)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // ---- legacy global tensor: grid, robots, carrying robots, waiting starts, waiting targets, in-transit targets, urgency
  if (a.g7 != nullptr) {
    float4* out4 = reinterpret_cast<float4*>(a.g7 + (size_t)widx * 7 * HW);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int q = lane + 32 * s;
      if (q < n4p) {
#pragma unroll
        for (int k = 0; k < 6; ++k) __stcs(out4 + k * n4p + q, lut[nib[s][k]]);
        __stcs(out4 + 6 * n4p + q, u4[s]);
      }
    }
  }
  // ---- per-robot observation: grid, urgency*, waiting starts* (* only when not carrying), others, self, own target
  if (a.obs6 != nullptr) {
    float4* base4 = reinterpret_cast<float4*>(a.obs6 + (size_t)widx * R * 6 * HW);
#pragma unroll 1
    for (int i = 0; i < R; ++i) {
      const uint32_t ci = (uint32_t)__shfl_sync(FULL, cell, i), ti = (uint32_t)__shfl_sync(FULL, own_tgt, i);
      const bool free_i = __shfl_sync(FULL, carry, i) == 0;
      float4* out4 = base4 + (size_t)i * 6 * n4p;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int q = lane + 32 * s;
        if (q < n4p) {
          const uint32_t oh_r = ((ci >> 2) == (uint32_t)q) ? (1u << (ci & 3u)) : 0u;
          const uint32_t oh_t = ((ti >> 2) == (uint32_t)q) ? (1u << (ti & 3u)) : 0u;  // 0xFFFF (none) matches no q
          __stcs(out4 + q, lut[nib[s][0]]);
          __stcs(out4 + n4p + q, free_i ? u4[s] : z4);
          __stcs(out4 + 2 * n4p + q, lut[free_i ? nib[s][3] : 0u]);
          __stcs(out4 + 3 * n4p + q, lut[nib[s][1] & ~oh_r]);
          __stcs(out4 + 4 * n4p + q, lut[oh_r]);
          __stcs(out4 + 5 * n4p + q, lut[oh_t]);
        }
      }
    }
  }
}

}  // namespace mdx
