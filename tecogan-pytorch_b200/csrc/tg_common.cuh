// Shared host/device helpers for libtecogan_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tecogan_b200.h"

// ---------------------------------------------------------------- error plumbing
void tg_set_error(const char* fmt, ...);

#define TG_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      tg_set_error(__VA_ARGS__);               \
      return (code);                           \
    }                                          \
  } while (0)

#define TG_CUDA_LAUNCH_CHECK(name)                                         \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      tg_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                     \
    }                                                                      \
  } while (0)

static inline int tg_ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- packed-weight geometry
// One weight tile = [cout_pad rows][64 k] fp16, 128-byte rows, 128B swizzle:
// byte offset of (row n, k) = n*128 + (((k>>3) ^ (n&7))<<4) + (k&7)*2.
__host__ __device__ static inline uint32_t tg_wtile_off(uint32_t n, uint32_t k) {
  return n * 128u + ((((k >> 3) ^ (n & 7u)) << 4) | ((k & 7u) << 1));
}

// conv3x3: group g = ky*3+kx reads input pixel (y+ky-1, x+kx-1).
// convT 3x3 s2 p1 op1: 9 groups ordered by output parity acc = py*2+px:
//   acc0: in[y,x]*Wt[1,1]
//   acc1: in[y,x]*Wt[1,2] + in[y,x+1]*Wt[1,0]
//   acc2: in[y,x]*Wt[2,1] + in[y+1,x]*Wt[0,1]
//   acc3: in[y,x]*Wt[2,2] + in[y,x+1]*Wt[2,0] + in[y+1,x]*Wt[0,2] + in[y+1,x+1]*Wt[0,0]
struct TgGroup { int8_t acc, dy, dx, ky, kx; };
__host__ __device__ static inline TgGroup tg_group(int kind, int g) {
  if (kind == TG_CONV_3X3) {
    TgGroup r = {0, (int8_t)(g / 3 - 1), (int8_t)(g % 3 - 1), (int8_t)(g / 3), (int8_t)(g % 3)};
    return r;
  }
  const int8_t T[9][5] = {{0, 0, 0, 1, 1}, {1, 0, 0, 1, 2}, {1, 0, 1, 1, 0}, {2, 0, 0, 2, 1},
                          {2, 1, 0, 0, 1}, {3, 0, 0, 2, 2}, {3, 0, 1, 2, 0}, {3, 1, 0, 0, 2},
                          {3, 1, 1, 0, 0}};
  TgGroup r = {T[g][0], T[g][1], T[g][2], T[g][3], T[g][4]};
  return r;
}

// ---------------------------------------------------------------- sampling helpers (device)
#ifdef __CUDACC__
// BicubicUpsampler taps (net_utils.py:116-131), a=-0.75, t = d/scale. Exact in fp32.
__device__ __forceinline__ void tg_cubic_taps(int d, int s, float k[4]) {
  const float a = -0.75f;
  const float t = (float)d / (float)s, t2 = t * t, t3 = t2 * t;
  k[0] = a * t - 2.f * a * t2 + a * t3;
  k[1] = 1.f - (a + 3.f) * t2 + (a + 2.f) * t3;
  k[2] = -a * t + (2.f * a + 3.f) * t2 - (a + 2.f) * t3;
  k[3] = a * t2 - a * t3;
}

// reflect index of F.pad(...,'reflect') on the bottom/right only: i in [0, L) over a source of
// length Ls <= L: i >= Ls -> 2*Ls-2-i   (tecogan_nets.py:239-241)
__device__ __forceinline__ int tg_reflect_hi(int i, int Ls) { return i < Ls ? i : 2 * Ls - 2 - i; }

__device__ __forceinline__ int tg_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// upsample_func evaluated at HR position (Y,X) of one plane of a (reflect-padded) LR image.
// src: plane [hs][ws]; logical padded size (h,w); bicubic = BicubicUpsampler (no half-pixel
// shift, replicate pad (1,2)); bilinear = F.interpolate(align_corners=False).
__device__ __forceinline__ float tg_upsample_at(const float* __restrict__ src, int hs, int ws,
                                                int h, int w, int s, int up_mode, int Y, int X) {
  if (up_mode == TG_UP_BICUBIC) {
    const int y = Y / s, dy = Y - y * s, x = X / s, dx = X - x * s;
    float ky[4], kx[4];
    tg_cubic_taps(dy, s, ky);
    tg_cubic_taps(dx, s, kx);
    float acc = 0.f;
    // vertical pass first then horizontal (net_utils.py:144-151); separable so order only
    // affects fp32 rounding.
    float col[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xx = tg_reflect_hi(tg_clampi(x - 1 + j, 0, w - 1), ws);
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int yy = tg_reflect_hi(tg_clampi(y - 1 + i, 0, h - 1), hs);
        v += ky[i] * __ldg(src + (size_t)yy * ws + xx);
      }
      col[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += kx[j] * col[j];
    return acc;
  } else {
    const float fs = (float)s;
    float sy = fmaxf(((float)Y + 0.5f) / fs - 0.5f, 0.f);
    float sx = fmaxf(((float)X + 0.5f) / fs - 0.5f, 0.f);
    int y0 = min((int)floorf(sy), h - 1), x0 = min((int)floorf(sx), w - 1);
    int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float fy = sy - (float)y0, fx = sx - (float)x0;
    y0 = tg_reflect_hi(y0, hs); y1 = tg_reflect_hi(y1, hs);
    x0 = tg_reflect_hi(x0, ws); x1 = tg_reflect_hi(x1, ws);
    const float v00 = __ldg(src + (size_t)y0 * ws + x0), v01 = __ldg(src + (size_t)y0 * ws + x1);
    const float v10 = __ldg(src + (size_t)y1 * ws + x0), v11 = __ldg(src + (size_t)y1 * ws + x1);
    const float top = v00 * (1.f - fx) + v01 * fx;
    const float bot = v10 * (1.f - fx) + v11 * fx;
    return top * (1.f - fy) + bot * fy;
  }
}

__device__ __forceinline__ float tg_act(float v, int act) {
  if (act == TG_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TG_ACT_LRELU02) return v >= 0.f ? v : 0.2f * v;
  return v;
}
#endif  // __CUDACC__
