// Epilogue math shared by the tcgen05 and the CUDA-core convolution kernels, so both produce
// identical values from identical accumulators.
#pragma once
#include "tg_common.cuh"

// TG_EPI_FLOW_NCHW_F32: torch.tanh(conv)*24 (tecogan_nets.py:80) -> flow NCHW fp32 [n,2,H,W]
__device__ __forceinline__ void tg_epi_flow(const tg_conv_desc& d, int n, int y, int x, int H,
                                            int W, int ch, float acc) {
  if (ch < d.cout_real)
    reinterpret_cast<float*>(d.y)[(((size_t)n * d.cout_real + ch) * H + y) * W + x] =
        24.f * tanhf(acc + __ldg(d.bias + ch));
}

// TG_EPI_OUT_NCHW_F32: out = conv_out(...) (tecogan_nets.py:144); the `+= upsample_func(lr_curr)`
// of :145 is applied afterwards by tg_upsample_nchw_f32(accumulate=1): (conv + bias) + up.
__device__ __forceinline__ void tg_epi_out(const tg_conv_desc& d, int n, int y, int x, int H,
                                           int W, int ch, float acc) {
  if (ch < d.cout_real)
    reinterpret_cast<float*>(d.y)[(((size_t)n * d.cout_real + ch) * H + y) * W + x] =
        acc + __ldg(d.bias + ch);
}

// TG_EPI_NHWC_F16 value: act(acc + bias) [+ residual]
__device__ __forceinline__ float tg_epi_val(float acc, float bias, int act) {
  return tg_act(acc + bias, act);
}

// CUDA-core kernel: 8 consecutive output channels c0..c0+7 of output pixel (n,oy,ox)
__device__ __forceinline__ void tg_epilogue_store8(const tg_conv_desc& d, int n, int oy, int ox,
                                                   int OH, int OW, int c0, const float a[8]) {
  if (d.epilogue == TG_EPI_NHWC_F16) {
    const size_t off = (((size_t)n * OH + oy) * OW + ox) * d.cout + c0;
    __align__(16) __half o[8];
    __align__(16) __half r[8];
    __align__(16) __half m[8];
    const bool bwd = d.act >= TG_ACT_DRELU;
    if (d.residual) *reinterpret_cast<uint4*>(r) = *reinterpret_cast<const uint4*>(
        reinterpret_cast<const __half*>(d.residual) + off);
    if (bwd) *reinterpret_cast<uint4*>(m) = *reinterpret_cast<const uint4*>(
        reinterpret_cast<const __half*>(d.mask) + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v;
      if (!bwd) {
        v = tg_epi_val(a[j], __ldg(d.bias + c0 + j), d.act);
        if (d.residual) v += __half2float(r[j]);
      } else {   // data gradient: (conv + bias [+ residual]) * act'(stored output)
        v = a[j] + __ldg(d.bias + c0 + j);
        if (d.residual) v += __half2float(r[j]);
        v *= tg_dact(__half2float(m[j]), d.act);
      }
      o[j] = __float2half(v);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(d.y) + off) = *reinterpret_cast<const uint4*>(o);
  } else if (d.epilogue == TG_EPI_FLOW_NCHW_F32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) tg_epi_flow(d, n, oy, ox, OH, OW, c0 + j, a[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) tg_epi_out(d, n, oy, ox, OH, OW, c0 + j, a[j]);
  }
}
