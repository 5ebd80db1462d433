/*
 * tecogan_b200.h -- C ABI of libtecogan_b200.so (sm_100a).
 *
 * The reference (skycrapers/TecoGAN-PyTorch @ 903b070) has NO native / FFI
 * boundary: its generator hot path is Python calling PyTorch library ops
 * (SURVEY.md 2.1, 8-b).  This ABI is therefore new; each entry point cites the
 * reference Python call site whose arithmetic it replaces.  The Python host
 * (tecogan-pytorch_b200/) binds it with ctypes and keeps the reference's
 * nn.Module surface (FRNet / FNet / SRNet / define_generator); INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns int: 0 = OK, <0 = invalid argument / unsupported
 *     shape (TG_E_*), >0 = cudaError_t.  Nothing throws.
 *   - every pointer is a caller-owned DEVICE pointer unless named host_*;
 *     no ownership transfer, no hidden allocation, no hidden synchronisation.
 *   - all work is enqueued on the cudaStream_t passed last (as void*), so the
 *     calls are CUDA-graph capturable.
 *   - activation layout inside the path: NHWC fp16, channel count a multiple
 *     of 64 ("c64"); module boundaries are NCHW fp32 like the reference.
 *   - tg_last_error_string() describes the most recent failure of this thread.
 */
#ifndef TECOGAN_B200_H_
#define TECOGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_ABI_VERSION 2   /* 2: tg_conv_desc.mask, backward (training) entry points; tg_conv_desc.reserved became
                              cin_real (0 keeps the old meaning: all stored input channels are used) */
#define TG_TAPN_ROWS 48   /* 9 taps x 4 output channels, padded to a multiple of 16 */

enum {
  TG_OK = 0,
  TG_E_INVALID = -1,      /* null pointer, non-positive size, bad enum       */
  TG_E_UNSUPPORTED = -2,  /* shape / channel count the kernels do not cover  */
  TG_E_DRIVER = -3        /* cuTensorMapEncodeTiled unavailable or failed    */
};

enum {
  TG_ACT_NONE = 0, TG_ACT_RELU = 1, TG_ACT_LRELU02 = 2,
  /* data-gradient epilogues: y = (conv + bias [+ residual]) * act'(mask), the derivative taken from
   * the STORED forward output `mask` of the layer the gradient flows into (sign(out) == sign(pre-act)) */
  TG_ACT_DRELU = 3,     /* * (mask > 0 ? 1 : 0)   */
  TG_ACT_DLRELU02 = 4   /* * (mask > 0 ? 1 : 0.2) */
};
enum {
  TG_CONV_3X3 = 0,
  TG_CONVT_3X3_S2 = 1,
  TG_CONV_3X3_S2 = 2    /* stride-2 conv, pad 1: y[oy,ox] = sum x[2oy+ky-1, 2ox+kx-1] * w[.,.,ky,kx] -- the data
                           gradient of TG_CONVT_3X3_S2 (tcgen05: tap mode over the four parity planes of x) */
};
enum { TG_UP_BICUBIC = 0, TG_UP_BILINEAR = 1 };
enum {
  TG_EPI_NHWC_F16 = 0,      /* y = act(conv + bias) [+ residual]  -> NHWC fp16        */
  TG_EPI_FLOW_NCHW_F32 = 1, /* y = 24*tanh(conv + bias)           -> NCHW fp32 [N,2,H,W] */
  TG_EPI_OUT_NCHW_F32 = 2,  /* y = conv + bias                    -> NCHW fp32 [N,C,H,W] */
  TG_EPI_NHWC_F16_POOL2 = 3 /* y = maxpool2x2(act(conv + bias))   -> NHWC fp16 [n,h/2,w/2,cout]: nn.MaxPool2d(2,2)
                               (tecogan_nets.py:28,35,42) folded into the producing conv's epilogue (tcgen05 kernel,
                               conv3x3 only, no residual); the full-resolution map is never written */
};
enum { TG_AMODE_AUTO = 0, TG_AMODE_HALO = 1, TG_AMODE_TAP = 2 };

int tg_version(void);
const char* tg_last_error_string(void);
/* number of SMs of the current device (148 on B200) */
int tg_device_sm_count(int* out_sm_count);

/* ------------------------------------------------------------------------
 * Weight packing (run once per optimizer step / checkpoint load).
 * Packed layout = the exact shared-memory image the tcgen05 kernel consumes:
 * tiles [group g][chunk c][cout_pad rows][64 k] fp16, 128-byte rows with the
 * 128B swizzle (16-byte chunk index XOR (row & 7)); g = ky*3+kx for conv3x3.
 * ---------------------------------------------------------------------- */
size_t tg_packed_weight_bytes(int cin_pad, int cout_pad);
/* nn.Conv2d(cin,cout,3,1,1).weight [cout,cin,3,3] fp32 (tecogan_nets.py:24-65,93-95,112,131) */
int tg_pack_conv3x3_weights(const float* w_oihw, int cout, int cin, void* packed,
                            int cout_pad, int cin_pad, void* stream);
/* nn.ConvTranspose2d(cin,cout,3,2,1,output_padding=1).weight [cin,cout,3,3] fp32
 * (tecogan_nets.py:119-126) -> 9 tiles grouped by output parity (1/2/2/4 taps) */
int tg_pack_convT3x3s2_weights(const float* w_iohw, int cin, int cout, void* packed,
                               int cout_pad, int cin_pad, void* stream);
/* Data-gradient operands (autograd of the above under loss.backward(), vsr_model.py:92):
 *  - conv3x3 dgrad = conv3x3 of dz with the taps flipped and cin/cout swapped: packs
 *    w'[ci][co][ky][kx] = w[co][ci][2-ky][2-kx] from the nn.Conv2d weight [cout,cin,3,3]; run it as a
 *    TG_CONV_3X3 layer with cin_pad(layer) = pad(cout), cout_pad(layer) = pad(cin).
 *  - convT dgrad = TG_CONV_3X3_S2 over dz with w'[ci][co][ky][kx] = wt[ci][co][ky][kx]: the
 *    nn.ConvTranspose2d weight [cin,cout,3,3] read as an OIHW conv weight (out = cin, in = cout). */
int tg_pack_conv3x3_weights_dgrad(const float* w_oihw, int cout, int cin, void* packed, int cin_as_cout_pad,
                                  int cout_as_cin_pad, void* stream);
int tg_pack_conv3x3s2_weights(const float* w_oihw, int cout, int cin, void* packed, int cout_pad, int cin_pad,
                              void* stream);
/* Thin heads (cout <= 4: FNet flow head 32->2, SRNet conv_out 64->3) use the "tap-major N"
 * layout: one tile per 64-ch chunk, [48 rows][64 k], row = tap*4 + co (rows >= 36 zero).  One
 * MMA group then yields all nine tap products of a pixel (N = 48) and the 3x3 shift-add happens
 * in the epilogue -- 4 MMAs per 128 pixels instead of 36.  Used with the two NCHW epilogues. */
size_t tg_packed_weight_bytes_tapn(int cin_pad);
int tg_pack_conv3x3_weights_tapn(const float* w_oihw, int cout, int cin, void* packed, int cin_pad,
                                 void* stream);

/* ------------------------------------------------------------------------
 * 3x3 convolution / stride-2 transposed convolution as tcgen05 implicit GEMM.
 * Replaces nn.Conv2d+activation (tecogan_nets.py:23-65, 92-98, 111-116, 131),
 * nn.ConvTranspose2d+ReLU (:119-126), torch.tanh(.)*24 (:80) and the add of
 * `out += upsample_func(lr_curr)` (:145): TG_EPI_OUT_NCHW_F32 stores conv+bias and the caller
 * then runs tg_upsample_nchw_f32(lr_curr, accumulate=1) on the same buffer (the epilogue is a pure
 * store: a read-modify-write there exposes a global-load round trip per tile).
 * ---------------------------------------------------------------------- */
typedef struct tg_conv_desc {
  const void* x;        /* NHWC fp16 [n,h,w,cin]  (TG_CONV_3X3_S2: [n,2h,2w,cin])                */
  const void* weights;  /* packed weights (tg_pack_*)                                    */
  const float* bias;    /* fp32 [cout] (zero padded)                                     */
  const void* residual; /* NHWC fp16 [n,h,w,cout] or NULL (TG_EPI_NHWC_F16, conv3x3 only) */
  void* y;              /* see epilogue; convT writes [n,2h,2w,cout]                     */
  int32_t n, h, w;      /* batch and the height / width the kernel tiles over: the input's (= the
                           output's for conv3x3; convT writes 2h x 2w), the OUTPUT's for TG_CONV_3X3_S2 */
  int32_t cin, cout;    /* stored channel counts: cin in {64,128,256}; cout in {64,128,256}
                           for TG_EPI_NHWC_F16, 48 (= TG_TAPN_ROWS, tap-major N packing)
                           for the two NCHW epilogues                                    */
  int32_t cout_real;    /* NCHW epilogues: channels actually written (2 resp. 3)         */
  int32_t kind;         /* TG_CONV_3X3 | TG_CONVT_3X3_S2                                 */
  int32_t act;          /* TG_ACT_*                                                      */
  int32_t epilogue;     /* TG_EPI_*                                                      */
  int32_t a_mode;       /* TG_AMODE_* (tcgen05 kernel only; AUTO = fastest validated)    */
  int32_t max_ctas;     /* 0 = one persistent CTA per SM                                 */
  int32_t cin_real;     /* input channels that can be non-zero (0 = cin): for cin = 64 the tcgen05 kernel skips the
                           UMMA k-steps of 16 channels at or beyond it -- the packed weights are zero there, so the
                           result is bit-identical (thin layers: FNet 6->32, 32->32, 32->64, 32->2)  */
  const void* mask;     /* TG_ACT_DRELU / TG_ACT_DLRELU02: NHWC fp16, shape of y; else NULL */
} tg_conv_desc;

int tg_conv_tcgen05(const tg_conv_desc* d, void* stream);
/* Same contract on CUDA cores (fp32 accumulate, reads the same packed weights):
 * bring-up / cross-check kernel used by the GPU tests, not by the hot path. */
int tg_conv_simt(const tg_conv_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * SRNet tail in one launch: last nn.ConvTranspose2d(64,64,3,2,1,op=1) + ReLU -> conv_out (64 -> out_nc)
 * -> + upsample_func(lr_curr) (tecogan_nets.py:119-131,143-145), optionally also float32_to_uint8 +
 * CHW->HWC (data_utils.py:80-87, tecogan_nets.py:278-281).  The 64-channel HR map only exists as one
 * 32x16-pixel tile in shared memory instead of a round trip through HBM.
 * ---------------------------------------------------------------------- */
typedef struct tg_tail_desc {
  const void* x;        /* input of the transposed conv, NHWC fp16 [n,h,w,64]                       */
  const void* w_up;     /* tg_pack_convT3x3s2_weights(cout_pad=64, cin_pad=64)                      */
  const float* b_up;    /* fp32 [64]                                                                */
  const void* w_out;    /* tg_pack_conv3x3_weights_tapn(cin_pad=64)                                 */
  const float* b_out;   /* fp32 [cout_real]                                                         */
  const float* lr;      /* lr_curr NCHW fp32 [n,cout_real,2h/lr_scale,2w/lr_scale] or NULL: the residual is
                           evaluated inside the kernel (16 gathers per pixel pair and channel)          */
  float* y;             /* NCHW fp32 [n,cout_real,2h,2w]; accumulate != 0: read-modify-write             */
  uint8_t* y_u8;        /* NHWC uint8 [n,2h,2w,cout_real] (round-half-even, clip) or NULL           */
  int32_t n, h, w;      /* of the transposed conv's input                                           */
  int32_t cout_real;    /* 1..3                                                                     */
  int32_t lr_scale;     /* 2 or 4: output size / lr size                                            */
  int32_t up_mode;      /* TG_UP_*                                                                  */
  int32_t max_ctas;     /* 0 = one persistent CTA per SM                                            */
  int32_t accumulate;   /* != 0: y already holds upsample_func(lr_curr) (tg_upsample_nchw_f32): out = y + conv +
                           bias -- one coalesced read per pixel instead of the in-kernel gathers (lr must be NULL) */
  int32_t reserved;     /* must be 0                                                                */
} tg_tail_desc;
int tg_convT_convout_tcgen05(const tg_tail_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * A chain of 64->64 3x3 convolutions (SRNet conv_in + the residual blocks,
 * tecogan_nets.py:92-100, 111-116, 139-141) as ONE persistent launch: every CTA walks all
 * layers over its fixed set of 16x8 tiles; a tile of layer l starts as soon as the (up to 9)
 * tiles of layer l-1 under its 18x10 halo have been published (per-tile progress flags in
 * `sync_ws`), so there is no launch, pipeline fill/drain or whole-grid barrier between layers,
 * and the next layer's weights stream into a ring of shared-memory tap slots behind the current
 * layer.  Bit-identical to n_layers calls of tg_conv_tcgen05.
 *   layers[l].x / y / residual : NHWC fp16 [n,h,w,64]; y[l] is normally x[l+1].  y[l] may alias
 *       residual[l] (in place) or a buffer last READ by layer <= l-1; it must not alias x[l].
 *       At most 4 distinct x buffers per chain.
 *   sync_ws : device memory of tg_conv_chain_workspace_bytes(n,h,w) bytes, zeroed ONCE by the
 *       caller before first use, then owned by the library (epoch-stamped; one chain launch in
 *       flight per workspace).
 * Needs every CTA co-resident (grid = min(#SM, tiles), 1 CTA/SM): launch at most ONE chain at a
 * time per device (two chains racing for SMs from different streams can starve each other) and
 * do not run it under an SM partition smaller than the device (waits are bounded and trap
 * instead of hanging).
 * ---------------------------------------------------------------------- */
typedef struct tg_chain_layer {
  const void* x;        /* NHWC fp16 [n,h,w,64]                          */
  const void* weights;  /* tg_pack_conv3x3_weights(cout_pad=64, cin_pad=64) */
  const float* bias;    /* fp32 [64]                                     */
  const void* residual; /* NHWC fp16 [n,h,w,64] or NULL                  */
  void* y;              /* NHWC fp16 [n,h,w,64]                          */
  int32_t act;          /* TG_ACT_*                                      */
  int32_t reserved;     /* must be 0                                     */
} tg_chain_layer;
#define TG_CHAIN_MAX_LAYERS 24

size_t tg_conv_chain_workspace_bytes(int n, int h, int w);
int tg_conv_chain_tcgen05(const tg_chain_layer* layers, int n_layers, int n, int h, int w,
                          void* sync_ws, int max_ctas, void* stream);

/* ------------------------------------------------------------------------
 * Fused  backward_warp + space_to_depth + concat  (HBM-bound).
 * Replaces net_utils.backward_warp (net_utils.py:50-82), space_to_depth
 * (:36-47) and torch.cat([lr_curr, hr_prev_tran]) (tecogan_nets.py:141).
 * out NHWC fp16 [n,h,w,cpad]: ch [0,c) = lr_curr, ch c+(sy*s+sx)*c+k =
 * warp(hr_prev)[k, y*s+sy, x*s+sx], remaining channels zero.
 * ---------------------------------------------------------------------- */
/* flow given at HR: hr_flow NCHW fp32 [n,2,s*h,s*w] (FRNet.forward_sequence, :201-212) */
int tg_warp_s2d_concat_hrflow(const float* hr_prev, const float* hr_flow, const float* lr_curr,
                              void* out, int n, int c, int h, int w, int s, int cpad,
                              void* stream);
/* flow given at LR: lr_flow NCHW fp32 [n,2,h8,w8]; reflect pad to (h,w) (:239-241),
 * upsample_func and the *scale (:244) are evaluated inline (FRNet.step, :227-252) */
int tg_warp_s2d_concat_lrflow(const float* hr_prev, const float* lr_flow, const float* lr_curr,
                              void* out, int n, int c, int h, int w, int h8, int w8, int s,
                              int up_mode, int cpad, void* stream);

/* ------------------------------------------------------------------------
 * Small NHWC fp16 helpers of FNet (tecogan_nets.py:28,35,42 and :74-79)
 * ---------------------------------------------------------------------- */
int tg_maxpool2x2_nhwc_f16(const void* x, void* y, int n, int h, int w, int c, void* stream);
int tg_upsample2x_bilinear_nhwc_f16(const void* x, void* y, int n, int h, int w, int c,
                                    void* stream);
/* cat([x1,x2],1) (tecogan_nets.py:71) + NCHW fp32 -> NHWC fp16, zero padded to cpad */
int tg_pack_pair_nhwc_f16(const float* x1, const float* x2, void* y, int n, int c, int h, int w,
                          int cpad, void* stream);

/* ------------------------------------------------------------------------
 * Module-boundary ops on NCHW fp32 (drop-in for codes/utils/net_utils.py)
 * ---------------------------------------------------------------------- */
int tg_backward_warp_nchw_f32(const float* x, const float* flow, float* y, int n, int c, int h,
                              int w, void* stream);                      /* net_utils.py:50-82  */
int tg_space_to_depth_nchw_f32(const float* x, float* y, int n, int c, int h, int w, int s,
                               void* stream);                            /* net_utils.py:36-47  */
/* y = [y +] mul * upsample(reflect_pad(x -> (h,w)))  ; x [n,c,hin,win], hin<=h, win<=w;
 * accumulate != 0 adds into y (fp32).
 * up_mode bicubic = BicubicUpsampler (net_utils.py:101-156), bilinear = F.interpolate
 * (net_utils.py:87-89).  hin==h, win==w, mul==1 gives the plain upsample_func. */
int tg_upsample_nchw_f32(const float* x, float* y, int n, int c, int hin, int win, int h, int w,
                         int s, int up_mode, float mul, int accumulate, void* stream);
int tg_nchw_f32_to_nhwc_f16(const float* x, void* y, int n, int c, int h, int w, int cpad,
                            int c_offset, void* stream);
int tg_nhwc_f16_to_nchw_f32(const void* x, float* y, int n, int c, int h, int w, int cpad,
                            void* stream);
/* float32_to_uint8 (data_utils.py:80-87) + CHW->HWC (tecogan_nets.py:281):
 * x NCHW fp32 [n,c,h,w] -> uint8 [n,h,w,c], round-half-even, clip [0,255] */
int tg_float_to_uint8_nhwc(const float* x, uint8_t* y, int n, int c, int h, int w, void* stream);

/* BD degradation of the data side (codes/utils/data_utils.py:30-53, called on GT frames by
 * base_model.py:75,115): optional reflect pad by (k-1)/2 | k-1-(k-1)/2, then a depthwise valid
 * correlation with the k x k kernel `k2d` (device, fp32, = create_kernel(sigma)[0,0]) and stride s.
 * x NCHW fp32 [n,c,H,W] -> y [n,c,h,w] with h = (Hp-k)/s+1, Hp = H (+k-1 when pad_data). */
int tg_downsample_bd_nchw_f32(const float* x, const float* k2d, float* y, int n, int c, int H, int W,
                              int k, int s, int pad_data, void* stream);

/* ========================================================================
 * Training: the generator backward.  Replaces autograd through
 * FRNet.forward_sequence (tecogan_nets.py:174-225) under loss_G.backward() (vsr_model.py:92,
 * vsrgan_model.py:273) and through backward_warp / upsample_func / fnet at the module boundary
 * (vsr_model.py:86, vsrgan_model.py:106-108,214-222, tecogan_nets.py:419-453).
 *
 * fp16 gradients between conv layers carry a power-of-two LOSS SCALE held in device memory:
 * `scale` points at two floats {scale, 1/scale} (tg_grad_scale_from_amax / tg_flow_head_bwd write
 * them); kernels producing fp32 results multiply by 1/scale.  scale == NULL means 1.
 * ====================================================================== */
size_t tg_grad_scale_workspace_bytes(void);   /* 16: {scale, 1/scale} fp32 + amax scratch (zero it once) */
/* scale = 2^floor(log2(target / max(|a|,|b|))) clamped to 2^+-24 (1 when all-zero); b may be NULL */
int tg_grad_scale_from_amax(const float* a, size_t na, const float* b, size_t nb, float target, void* ws,
                            void* stream);
/* (a [+ b]) * scale : NCHW fp32 [n,c,h,w] -> NHWC fp16 [n,h,w,cpad] (pad channels zero) */
int tg_grad_pack_nhwc_f16(const float* a, const float* b, const float* scale, void* y, int n, int c, int h, int w,
                          int cpad, void* stream);
/* channels [c_offset, c_offset+c) of NHWC fp16 -> NCHW fp32 * 1/scale ; accumulate != 0 adds into y */
int tg_grad_unpack_nchw_f32(const void* x, const float* scale, float* y, int n, int c, int h, int w, int cpad,
                            int c_offset, int accumulate, void* stream);
/* db[ch] += 1/scale * sum_pixels dz[p][ch], ch < c_real  (bias gradient of any conv layer) */
int tg_bias_grad_nhwc_f16(const void* dz, size_t npix, int c, int c_real, const float* scale, float* db,
                          void* stream);

/* Weight gradient of a conv3x3 / convT3x3s2 layer: dw += 1/scale * sum_p x[p+tap] (x) dz[p], written in
 * the parameter's own layout (nn.Conv2d [cout_real,cin_real,3,3]; nn.ConvTranspose2d
 * [cin_real,cout_real,3,3]) with fp32 atomics -- the caller zeroes dw (or accumulates on purpose).
 * tcgen05: GEMM over pixels (K), x and dz both channel-contiguous ("MN-major") operands. */
typedef struct tg_wgrad_desc {
  const void* x;        /* layer input,  NHWC fp16 [n,h,w,cin]                              */
  const void* dz;       /* gradient of the pre-activation output, NHWC fp16 [n,h,w,cout]
                           (convT: [n,2h,2w,cout]), loss-scaled                           */
  float* dw;            /* fp32 gradient, parameter layout                                  */
  const float* scale;   /* {scale, 1/scale} or NULL                                         */
  float* db;            /* conv3x3 only, may be NULL: bias gradient db[co] += 1/scale * sum_p dz[p][co], computed by
                           the same MMAs (the unused half of the last tap pair reads a block of ones) */
  int32_t n, h, w;      /* of the layer INPUT                                               */
  int32_t cin, cout;    /* stored channel counts (64/128/256)                               */
  int32_t cin_real, cout_real;
  int32_t kind;         /* TG_CONV_3X3 | TG_CONVT_3X3_S2                                    */
  int32_t max_ctas;     /* 0 = all SMs                                                      */
  int32_t reserved;     /* must be 0                                                        */
} tg_wgrad_desc;
int tg_wgrad_tcgen05(const tg_wgrad_desc* d, void* stream);
int tg_wgrad_simt(const tg_wgrad_desc* d, void* stream);    /* CUDA-core cross-check (tests only) */

/* grid_sample(bilinear, border, align_corners) backward = autograd of net_utils.backward_warp
 * (net_utils.py:50-82): gx += scatter(gy) (fp32 atomics, caller zeroes gx), gflow = gather; either
 * output may be NULL. */
int tg_backward_warp_bwd_nchw_f32(const float* x, const float* flow, const float* gy, float* gx, float* gflow,
                                  int n, int c, int h, int w, void* stream);
/* gradient of tg_warp_s2d_concat_hrflow w.r.t. hr_prev (atomic accumulate) and hr_flow (store), from the
 * loss-scaled NHWC fp16 gradient `gx` of the SRNet input; either output may be NULL */
int tg_warp_s2d_concat_bwd(const void* gx, const float* hr_prev, const float* hr_flow, const float* scale,
                           float* d_hr_prev, float* d_hr_flow, int n, int c, int h, int w, int s, int cpad,
                           void* stream);
/* gradient of y = mul * upsample_func(x): gy [n,c,s*h,s*w] -> gx [n,c,h,w] (net_utils.py:85-156) */
int tg_upsample_bwd_nchw_f32(const float* gy, float* gx, int n, int c, int h, int w, int s, int up_mode, float mul,
                             int accumulate, void* stream);
/* FNet helpers (tecogan_nets.py:28,35,42,74-79), each fused with the activation derivative of the conv
 * layer whose stored output is `x` / `m` (act = that layer's TG_ACT_*):
 *   maxpool:  gx[n,h,w,c] = route(gy[n,h/2,w/2,c]) * act'(x)      upsample2x: gx[n,h,w,c] = T(gy[n,2h,2w,c]) * act'(m) */
int tg_maxpool2x2_bwd_nhwc_f16(const void* x, const void* gy, void* gx, int n, int h, int w, int c, int act,
                               void* stream);
int tg_upsample2x_bilinear_bwd_nhwc_f16(const void* gy, const void* m, void* gx, int n, int h, int w, int c,
                                        int act, void* stream);
/* flow head: flow = 24*tanh(z) (tecogan_nets.py:80): dz = (gflow [+ gflow2]) * (24 - flow^2/24) * scale as NHWC
 * fp16 [n,h,w,cpad]; chooses the loss scale of the FNet backward from the amax of that product (scale_ws as
 * in tg_grad_scale_from_amax) */
int tg_flow_head_bwd(const float* gflow, const float* gflow2, const float* flow, void* scale_ws, float target,
                     void* dz, int n, int h, int w, int cpad, void* stream);
/* Input builder of the spatio-temporal discriminator (tecogan_nets.py:438-463): the three
 * backward_warps of every 3-frame clip + centre crop / zero pad + "rrrgggbbb" permutes + 27-channel concat
 * in one kernel.  data, bi: [n,t_full,c,h,w] fp32 (frames >= t are ignored, t % 3 == 0); flow:
 * hr_flow_merge [n*t,2,h,w] (clip-major, 3 per clip); out: [n*t/3, 9*c, h, w] = [orig | warp | cond].
 * pad = (spatial_size - c_size)/2, csize = c_size = int(spatial_size * crop_border_ratio). */
int tg_st_disc_input_nchw_f32(const float* data, const float* bi, const float* flow, float* out, int n, int t_full,
                              int t, int c, int h, int w, int pad, int csize, void* stream);
/* its gradient w.r.t. data: gdata [n,t_full,c,h,w] (zeroed by the caller) += orig part + warp scatter */
int tg_st_disc_input_bwd_nchw_f32(const float* gout, const float* flow, float* gdata, int n, int t_full, int t, int c,
                                  int h, int w, int pad, int csize, void* stream);
/* space_to_depth backward (net_utils.py:36-47): gy [n,c*s*s,h/s,w/s] -> gx [n,c,h,w] */
int tg_depth_to_space_nchw_f32(const float* gy, float* gx, int n, int c, int h, int w, int s, void* stream);

/* ------------------------------------------------------------------------
 * Diagnostics: when a device buffer of 16*gridDim uint64 is registered, every
 * tg_conv_tcgen05 launch writes per-CTA role timers (cycles spent by the TMA
 * producer / MMA issuer / epilogue in each wait and work phase) into it.
 * NULL (default) disables timing. Layout: tools/conv_timers.py.
 * ---------------------------------------------------------------------- */
int tg_debug_set_conv_timers(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* TECOGAN_B200_H_ */
