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

#define TG_ABI_VERSION 1
#define TG_TAPN_ROWS 48   /* 9 taps x 4 output channels, padded to a multiple of 16 */

enum {
  TG_OK = 0,
  TG_E_INVALID = -1,      /* null pointer, non-positive size, bad enum       */
  TG_E_UNSUPPORTED = -2,  /* shape / channel count the kernels do not cover  */
  TG_E_DRIVER = -3        /* cuTensorMapEncodeTiled unavailable or failed    */
};

enum { TG_ACT_NONE = 0, TG_ACT_RELU = 1, TG_ACT_LRELU02 = 2 };
enum { TG_CONV_3X3 = 0, TG_CONVT_3X3_S2 = 1 };
enum { TG_UP_BICUBIC = 0, TG_UP_BILINEAR = 1 };
enum {
  TG_EPI_NHWC_F16 = 0,      /* y = act(conv + bias) [+ residual]  -> NHWC fp16        */
  TG_EPI_FLOW_NCHW_F32 = 1, /* y = 24*tanh(conv + bias)           -> NCHW fp32 [N,2,H,W] */
  TG_EPI_OUT_NCHW_F32 = 2   /* y = conv + bias                    -> NCHW fp32 [N,C,H,W] */
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
  const void* x;        /* NHWC fp16 [n,h,w,cin]                                         */
  const void* weights;  /* packed weights (tg_pack_*)                                    */
  const float* bias;    /* fp32 [cout] (zero padded)                                     */
  const void* residual; /* NHWC fp16 [n,h,w,cout] or NULL (TG_EPI_NHWC_F16, conv3x3 only) */
  void* y;              /* see epilogue; convT writes [n,2h,2w,cout]                     */
  int32_t n, h, w;      /* input batch / height / width                                  */
  int32_t cin, cout;    /* stored channel counts: cin in {64,128,256}; cout in {64,128,256}
                           for TG_EPI_NHWC_F16, 48 (= TG_TAPN_ROWS, tap-major N packing)
                           for the two NCHW epilogues                                    */
  int32_t cout_real;    /* NCHW epilogues: channels actually written (2 resp. 3)         */
  int32_t kind;         /* TG_CONV_3X3 | TG_CONVT_3X3_S2                                 */
  int32_t act;          /* TG_ACT_*                                                      */
  int32_t epilogue;     /* TG_EPI_*                                                      */
  int32_t a_mode;       /* TG_AMODE_* (tcgen05 kernel only; AUTO = fastest validated)    */
  int32_t max_ctas;     /* 0 = one persistent CTA per SM                                 */
  int32_t reserved;     /* must be 0                                                     */
} tg_conv_desc;

int tg_conv_tcgen05(const tg_conv_desc* d, void* stream);
/* Same contract on CUDA cores (fp32 accumulate, reads the same packed weights):
 * bring-up / cross-check kernel used by the GPU tests, not by the hot path. */
int tg_conv_simt(const tg_conv_desc* d, void* stream);

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
