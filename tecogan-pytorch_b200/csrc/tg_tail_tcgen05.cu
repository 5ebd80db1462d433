// SRNet tail in ONE kernel (sm_100a): last ConvTranspose2d(64,64,3,2,1,op=1) + ReLU  ->  conv_out 3x3
// (64 -> out_nc <= 4)  ->  + upsample_func(lr_curr)  ->  fp32 NCHW frame (+ uint8 NHWC frame).
//
// Replaces nn.ConvTranspose2d+ReLU, nn.Conv2d and `out += self.upsample_func(lr_curr)`
// (tecogan_nets.py:119-131,143-145) and float32_to_uint8 + CHW->HWC (data_utils.py:80-87,
// tecogan_nets.py:278-281).  As separate kernels the 64-channel HR map (351 MB for four 536x1280
// frames) is written to and read back from HBM -- 24 % of the round-1 step for 16 % of its FLOPs;
// here it only ever exists as one 32x16-pixel tile in shared memory.
//
// One CTA walks tiles of 16x8 pixels of the transposed conv's INPUT (17x9 halo box by TMA):
//   1. transposed conv = 4 parity accumulators in TMEM (1/2/2/4 taps, K = 64 per tap), as in
//      tg_conv_tcgen05.cu; issued as two halves (parities 0,1 | 2,3) with their own barriers.
//   2. epilogue A (8 warps): TMEM -> +bias, ReLU, fp16 -> the HR tile in shared memory, stored directly
//      in the UMMA operand layout (K-major, 128B swizzle): block q = parity (py,px), row m = input
//      pixel (m>>3, m&7) = HR pixel (2*(m>>3)+py, 2*(m&7)+px); pixels outside the image are ZERO
//      (conv_out's zero padding).
//   3. conv_out as "tap-major N" MMAs on that tile: D2[pixel][tap*4+co] = x[pixel] . W[tap][co]
//      (N = 48, four K=16 MMAs per 128-pixel block).
//   4. epilogue B (8 warps): the 3x3 shift-add  out[P] = sum_taps D2[P+off(tap)][tap].  Lane m of every
//      block is the 2x2 HR quad of input pixel m; two warp groups each produce one row of the quad: the
//      terms that stay inside the quad row are summed in registers, the ones that cross to the left / right
//      quad travel by warp shuffle (lanes +-1), the ones that cross the top / bottom edge are pre-summed per
//      receiving pixel and handed over through 8 float4 slots per quad in shared memory (double-buffered by
//      tile parity: one barrier per tile); + bias + upsample_func(lr_curr) (compile-time taps, one 4x4 LR
//      neighbourhood per pixel pair) -> fp32 NCHW (8-byte stores) and uint8 NHWC (2-byte stores).
//      (Round-2 measurement that shaped this: with ONE 4-warp group doing all of it the kernel was bound by
//      that group's single-warp instruction latency -- 12.6k cycles per tile against 2.4k of MMA work,
//      profiles/tail_timers_r2d.log.)
//   Valid outputs per tile: 30x14 HR pixels (tiles advance by 15x7 input pixels; the transposed conv is
//   recomputed on the one-pixel ring, 1.22x).  Issue order of the MMA warp is software-pipelined --
//   ConvT(i).p0, conv_out(i-1).b23, ConvT(i).p1, conv_out(i).b01 -- so the tensor pipe always has the
//   other half's MMAs queued while an epilogue-A half converts, and the in-order pipe itself protects
//   the HR tile against being overwritten before conv_out has read it.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "tg_common.cuh"
#include "tg_tcgen05.cuh"

namespace {

constexpr int TH = 16, TW = 8;
constexpr int kStepY = 15, kStepX = 7;          // input pixels a tile advances by
constexpr int kThreads = 640;                    // 4 control warps + 8 epilogue-A + 8 epilogue-B
constexpr int kBoxW = TW + 1, kBoxH = TH + 1;    // 17 x 9 halo box, origin at the tile's first pixel
constexpr uint32_t kHaloBytes = kBoxW * kBoxH * 128;           // 19584
constexpr uint32_t kStageBytes = (kHaloBytes + 1023u) & ~1023u;  // 20480
constexpr int kStages = 2;
constexpr uint32_t kWtBytes = 9 * 64 * 128;      // transposed-conv weights, 9 tap tiles of [64][64]
constexpr uint32_t kWoBytes = TG_TAPN_ROWS * 128;  // conv_out weights, [48 rows = tap*4+co][64]
constexpr uint32_t kHrBlock = 128 * 128;         // one parity block of the HR tile: 128 pixels x 128 B
constexpr uint32_t kExBytes = 2 * 8 * 128 * 16; // exchange: 8 float4 slots per quad, double-buffered by tile parity
constexpr uint32_t kOffWt = 2048;
constexpr uint32_t kOffWo = kOffWt + kWtBytes;                 // 75776
constexpr uint32_t kOffStage = kOffWo + kWoBytes;              // 81920 (1024-aligned)
constexpr uint32_t kOffHr = kOffStage + kStages * kStageBytes; // 122880 (1024-aligned)
constexpr uint32_t kOffEx = kOffHr + 4 * kHrBlock;             // 188416
constexpr uint32_t kSmemBytes = kOffEx + kExBytes + 1024;      // + alignment slack = 222208
static_assert(kOffStage % 1024 == 0 && kOffHr % 1024 == 0, "swizzled regions need 1024-byte alignment");
static_assert(kSmemBytes <= 232448, "shared memory budget");
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kD2Col = 256;                 // D2 block q at columns 256 + 64*q (48 used)

struct TailParams {
  CUtensorMap map_x;
  const unsigned char* w_up;
  const unsigned char* w_out;
  const float* b_up;
  const float* b_out;
  const float* lr;           // lr_curr NCHW fp32 [n,cout_real,lh,lw] or null
  float* y;                  // NCHW fp32 [n,cout_real,2h,2w]
  uint8_t* y_u8;             // NHWC uint8 [n,2h,2w,cout_real] or null
  int n, h, w, cout_real;
  int lr_scale, up_mode, lh, lw;
  int tiles_x, tiles_y, num_tiles;
  uint32_t idesc_up, idesc_out;
  float taps[5][4];          // upsample_func taps per phase d (host-computed, = tg_up_taps): constant-bank operands
  int accumulate;            // y already holds the residual (upsample_func(lr_curr)): out = y + conv + bias
  int flags;                 // diagnostics (TG_TAIL_FLAGS): 1 = skip the residual, 2 = skip uint8, 4 = skip the exchange
  unsigned long long* dbg;   // optional per-CTA role timers (tg_debug_set_conv_timers), 16 slots per CTA
};
// role-timer slots (cycles per CTA): tools/conv_timers.py tail
enum { TT_MMA_WAIT_FULL = 0, TT_MMA_WAIT_TEMPTY, TT_MMA_WAIT_HRFULL, TT_MMA_WAIT_D2EMPTY, TT_MMA_TOTAL, TT_EA_WAIT, TT_EA_BUSY,
       TT_EB_WAIT, TT_EB_TMEM, TT_EB_EXCH, TT_EB_RESID, TT_EB_STORE, TT_EB_TOTAL, TT_KERNEL, TT_TILES, TT_EB_TOP, TT_SLOTS = 16 };
#define TT0() (TIMING ? clock64() : 0)
#define TTACC(var, t0) do { if (TIMING) var += clock64() - (t0); } while (0)

// tile -> (image, first input row / column) walked incrementally (tile += gridDim.x): no integer divisions in the
// per-tile loops of the five roles
struct TileWalk {
  int img, by, bx, tiles_x, tiles_y, step;
  __device__ __forceinline__ TileWalk(int tile0, int tx, int ty, int stp) : tiles_x(tx), tiles_y(ty), step(stp) {
    const int per_img = tx * ty;
    img = tile0 / per_img;
    const int r = tile0 - img * per_img;
    by = r / tx;
    bx = r - by * tx;
  }
  __device__ __forceinline__ void next() {
    bx += step;
    while (bx >= tiles_x) { bx -= tiles_x; ++by; }
    while (by >= tiles_y) { by -= tiles_y; ++img; }
  }
  __device__ __forceinline__ int y0() const { return by * kStepY - 1; }
  __device__ __forceinline__ int x0() const { return bx * kStepX - 1; }
};

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ uint32_t q8(float v) { return (uint32_t)fminf(fmaxf(rintf(v * 255.f), 0.f), 255.f); }

// LRS = output size / lr size of the fused residual (0: none, 2, 4); UPM = TG_UP_* of upsample_func
template <bool TIMING, int LRS, int UPM>
__global__ void __launch_bounds__(kThreads, 1)
tail_tcgen05_kernel(const __grid_constant__ TailParams p) {
  const long long t_kernel0 = TIMING ? clock64() : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  // (a shuffle-broadcast warp index, which helps the other tcgen05 kernels, makes this one 45 % slower: measured
  // 264 vs 182 us -- the epilogue roles lose registers to the uniform-path bookkeeping)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_full = base;              // [2] halo stage landed
  const uint32_t bar_empty = base + 16;        // [2] halo stage consumed
  const uint32_t bar_w = base + 32;            // weights resident
  const uint32_t bar_tfull = base + 40;        // [2] ConvT accumulator half complete
  const uint32_t bar_tempty = base + 56;       // [2] ... drained by epilogue A
  const uint32_t bar_hrfull = base + 72;       // [2] HR-tile half written by epilogue A
  const uint32_t bar_d2full = base + 88;       // conv_out accumulators complete
  const uint32_t bar_d2empty = base + 96;      // ... drained by epilogue B
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + 128);
  float* bias_up_s = reinterpret_cast<float*>(sm + 1024);
  float* bias_out_s = reinterpret_cast<float*>(sm + 1024 + 256);

  if (warp == 0 && lane == 0) tma_prefetch_desc(&p.map_x);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_w, 1);
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_tfull + 8 * h, 1);
      mbar_init(bar_tempty + 8 * h, 4);
      mbar_init(bar_hrfull + 8 * h, 4);
    }
    mbar_init(bar_d2full, 1);
    mbar_init(bar_d2empty, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_s)), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // resident weights do not depend on the previous kernel: load them before the PDL wait
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(bar_w, kWtBytes + kWoBytes);
    for (int g = 0; g < 9; ++g) bulk_load(base + kOffWt + g * 8192u, p.w_up + (size_t)g * 8192u, 8192u, bar_w);
    bulk_load(base + kOffWo, p.w_out, kWoBytes, bar_w);
  }
  tg_pdl_wait();
  tg_pdl_trigger();
  for (int i = threadIdx.x; i < 64; i += kThreads) bias_up_s[i] = p.b_up[i];
  if (threadIdx.x < 4) bias_out_s[threadIdx.x] = threadIdx.x < p.cout_real ? p.b_out[threadIdx.x] : 0.f;
  __syncthreads();

  const int H = 2 * p.h, W = 2 * p.w;

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      int it = 0;
      TileWalk tw(blockIdx.x, p.tiles_x, p.tiles_y, gridDim.x);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it, tw.next()) {
        const int stage = it & 1;
        const uint32_t phase = (uint32_t)(it >> 1) & 1u;
        const int img = tw.img, y0 = tw.y0(), x0 = tw.x0();
        mbar_wait(bar_empty + 8 * stage, phase ^ 1, 1);
        mbar_expect_tx(bar_full + 8 * stage, kHaloBytes);
        tma_load_4d(base + kOffStage + stage * kStageBytes, &p.map_x, bar_full + 8 * stage, 0, x0, y0, img);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    mbar_wait(bar_w, 0, 2);
    const uint64_t a_hi = make_sdesc(0, (uint32_t)kBoxW * 128u);     // halo views: 8-row groups one box row apart
    const uint64_t k_hi = make_sdesc(0, 1024u);                      // weights / HR tile blocks: dense 8-row groups
    const uint32_t wt16 = ((base + kOffWt) & 0x3FFFFu) >> 4;
    const uint32_t wo16 = ((base + kOffWo) & 0x3FFFFu) >> 4;
    const uint32_t hr16 = ((base + kOffHr) & 0x3FFFFu) >> 4;
    auto issue_convt = [&](uint32_t sa16, int g_lo, int g_hi) {
#pragma unroll
      for (int g = 0; g < 9; ++g) {
        if (g < g_lo || g >= g_hi) continue;
        const TgGroup gr = tg_group(TG_CONVT_3X3_S2, g);
        const bool first_of_acc = (g == 0) || (tg_group(TG_CONVT_3X3_S2, g > 0 ? g - 1 : 0).acc != gr.acc);
        const uint32_t off = (uint32_t)(gr.dy * kBoxW + gr.dx) * 8u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + (uint32_t)gr.acc * 64u, a_hi | (uint64_t)(sa16 + off + 2u * k),
                   k_hi | (uint64_t)(wt16 + (uint32_t)g * 512u + 2u * k), p.idesc_up, (first_of_acc && k == 0) ? 0u : 1u);
      }
    };
    auto issue_convout = [&](int q_lo) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + kD2Col + (uint32_t)(q_lo + q) * 64u, k_hi | (uint64_t)(hr16 + (uint32_t)(q_lo + q) * 1024u + 2u * k),
                   k_hi | (uint64_t)(wo16 + 2u * k), p.idesc_out, k == 0 ? 0u : 1u);
    };
    int it = 0;
    long long tw_full = 0, tw_tempty = 0, tw_hrfull = 0, tw_d2empty = 0;
    const long long t_mma0 = TT0();
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int stage = it & 1;
      const uint32_t sph = (uint32_t)(it >> 1) & 1u, tph = (uint32_t)it & 1u;
      const uint32_t sa16 = ((base + kOffStage + stage * kStageBytes) & 0x3FFFFu) >> 4;
      long long t0 = TT0();
      mbar_wait(bar_full + 8 * stage, sph, 3);
      TTACC(tw_full, t0); t0 = TT0();
      mbar_wait(bar_tempty + 0, tph ^ 1, 4);
      TTACC(tw_tempty, t0);
      tc_fence_after();
      if (elect_one_sync()) { issue_convt(sa16, 0, 3); umma_commit(bar_tfull + 0); }
      __syncwarp();
      if (it > 0) {                                   // second half of conv_out of the previous tile
        t0 = TT0();
        mbar_wait(bar_hrfull + 8, (uint32_t)(it - 1) & 1u, 5);
        TTACC(tw_hrfull, t0);
        tc_fence_after();
        if (elect_one_sync()) { issue_convout(2); umma_commit(bar_d2full); }
        __syncwarp();
      }
      t0 = TT0();
      mbar_wait(bar_tempty + 8, tph ^ 1, 4);
      TTACC(tw_tempty, t0);
      tc_fence_after();
      if (elect_one_sync()) { issue_convt(sa16, 3, 9); umma_commit(bar_tfull + 8); umma_commit(bar_empty + 8 * stage); }
      __syncwarp();
      t0 = TT0();
      mbar_wait(bar_hrfull + 0, tph, 5);
      TTACC(tw_hrfull, t0); t0 = TT0();
      mbar_wait(bar_d2empty, tph ^ 1, 6);
      TTACC(tw_d2empty, t0);
      tc_fence_after();
      if (elect_one_sync()) issue_convout(0);
      __syncwarp();
    }
    if (it > 0) {
      mbar_wait(bar_hrfull + 8, (uint32_t)(it - 1) & 1u, 5);
      tc_fence_after();
      if (elect_one_sync()) { issue_convout(2); umma_commit(bar_d2full); }
      __syncwarp();
    }
    if (TIMING && lane == 0) {
      unsigned long long* o = p.dbg + blockIdx.x * TT_SLOTS;
      o[TT_MMA_WAIT_FULL] = tw_full; o[TT_MMA_WAIT_TEMPTY] = tw_tempty; o[TT_MMA_WAIT_HRFULL] = tw_hrfull;
      o[TT_MMA_WAIT_D2EMPTY] = tw_d2empty; o[TT_MMA_TOTAL] = clock64() - t_mma0; o[TT_TILES] = it;
    }
  } else if (warp >= 4 && warp < 12) {
    // ============================================================ epilogue A: ConvT accumulators -> HR tile (smem)
    const int half = (warp - 4) >> 2;         // parities {0,1} or {2,3}
    const int q = warp & 3;
    const int m = q * 32 + lane;              // TMEM lane = input pixel of the tile
    const int ty = m >> 3, tx = m & 7;
    const uint32_t swz = (uint32_t)(m & 7);
    int it = 0;
    long long ta_wait = 0, ta_busy = 0;
    TileWalk tw(blockIdx.x, p.tiles_x, p.tiles_y, gridDim.x);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it, tw.next()) {
      const int y0 = tw.y0(), x0 = tw.x0();
      // does the 32x16 HR tile reach outside the image?
      const bool border_tile = y0 < 0 || x0 < 0 || y0 + TH > p.h || x0 + TW > p.w;
      long long t0 = TT0();
      mbar_wait(bar_tfull + 8 * half, (uint32_t)it & 1u, 7);
      TTACC(ta_wait, t0); t0 = TT0();
      tc_fence_after();
      // both 32-column pieces of an accumulator are requested before the first is consumed
      uint32_t v[2][32];
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        const int acc = half * 2 + a2;
        const uint32_t tad = tmem_base + (uint32_t)acc * 64u + ((uint32_t)(q * 32) << 16);
        tmem_ld32(tad, v[0]);
        tmem_ld32(tad + 32, v[1]);
        tmem_ld_wait();
        if (a2 == 1) {                          // this warp has read everything of its half
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * half);
        }
        const int Y = 2 * (y0 + ty) + (acc >> 1), X = 2 * (x0 + tx) + (acc & 1);
        // pixels outside the image are conv_out's zero padding, not relu(bias): only tiles on the image border
        // have any (warp-uniform test), everywhere else the mask costs nothing
        const bool outside = border_tile && !(Y >= 0 && Y < H && X >= 0 && X < W);
        uint8_t* row = sm + kOffHr + (uint32_t)acc * kHrBlock + (uint32_t)m * 128u;
#pragma unroll
        for (int pc = 0; pc < 2; ++pc)
#pragma unroll
          for (int c = 0; c < 4; ++c) {         // four 16-byte chunks = 32 channels
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ch = pc * 32 + c * 8 + j * 2;
              const float a0 = fmaxf(__uint_as_float(v[pc][c * 8 + j * 2]) + bias_up_s[ch], 0.f);
              const float a1 = fmaxf(__uint_as_float(v[pc][c * 8 + j * 2 + 1]) + bias_up_s[ch + 1], 0.f);
              oh[j] = __floats2half2_rn(a0, a1);
            }
            if (outside) o = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(row + ((((uint32_t)(pc * 4 + c)) ^ swz) << 4)) = o;
          }
      }
      fence_proxy_async_smem();               // generic-proxy stores -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hrfull + 8 * half);
      TTACC(ta_busy, t0);
    }
    if (TIMING && warp == 4 && lane == 0) {
      p.dbg[blockIdx.x * TT_SLOTS + TT_EA_WAIT] = ta_wait; p.dbg[blockIdx.x * TT_SLOTS + TT_EA_BUSY] = ta_busy;
    }
  } else if (warp >= 12) {
    // ============================================================ epilogue B: shift-add, residual, stores
    // Two groups of 4 warps; group R produces output row R of every 2x2 quad (and the contributions that
    // leave the quad through its top (R = 0) / bottom (R = 1) edge).  Horizontal neighbours are lanes +-1 of
    // the same warp (shuffles); vertical / diagonal neighbours go through 8 float4 slots per quad in shared
    // memory, double-buffered by tile parity so one 256-thread barrier per tile suffices.
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    const bool up = ty > 0, dn = ty < TH - 1, lf = tx > 0, rt = tx < TW - 1;
    // the group's row is a compile-time constant of the body (register-resident tap tables)
    auto run_b = [&](auto Rtag) {
    constexpr int R = decltype(Rtag)::value;
    int it = 0;
    long long tb_wait = 0, tb_tmem = 0, tb_exch = 0, tb_resid = 0, tb_store = 0, tb_top = 0;
    const long long t_eb0 = TT0();
    long long t_top0 = t_eb0;
    TileWalk tw(blockIdx.x, p.tiles_x, p.tiles_y, gridDim.x);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it, tw.next()) {
      const int img = tw.img, y0 = tw.y0(), x0 = tw.x0();
      // ---- the two HR pixels this thread produces (row R of its quad)
      const int gy = y0 + ty, gx = x0 + tx;   // input pixel
      const int Yt = 2 * ty + R, Y = 2 * gy + R, X0 = 2 * gx;
      const bool row_ok = Yt >= 1 && Yt <= 2 * TH - 2 && Y >= 0 && Y < H;
      const bool ok0 = row_ok && tx >= 1 && X0 >= 0 && X0 < W;
      const bool ok1 = row_ok && tx <= TW - 2 && X0 + 1 >= 0 && X0 + 1 < W;
      const size_t HWs = (size_t)H * W;
      float* dst = p.y + (size_t)img * p.cout_real * HWs + (size_t)Y * W + X0;
      // accumulate mode: y already holds upsample_func(lr_curr); its read is issued before the wait for the
      // accumulators so the DRAM / L2 latency hides behind the MMAs of this tile
      float yo[2][3] = {};
      if (p.accumulate) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k < p.cout_real) {
            if (ok0 && ok1) {
              const float2 t2 = *reinterpret_cast<const float2*>(dst + k * HWs);
              yo[0][k] = t2.x; yo[1][k] = t2.y;
            } else if (ok0) yo[0][k] = dst[k * HWs];
            else if (ok1) yo[1][k] = dst[k * HWs + 1];
          }
        }
      }
      TTACC(tb_top, t_top0);
      long long t0 = TT0();
      mbar_wait(bar_d2full, (uint32_t)it & 1u, 8);
      TTACC(tb_wait, t0); t0 = TT0();
      tc_fence_after();
      // accumulators of this thread: own[c] = out(row R, col c); side[0/1] = row R, col -1 / col 2 (left / right
      // quad); edge[c] = row (R ? 2 : -1), col c (quad below / above); corner[0/1] = that row, col -1 / 2
      float own[2][3] = {}, side[2][3] = {}, edge[2][3] = {}, corner[2][3] = {};
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        const int spy = blk >> 1, spx = blk & 1;      // the source pixel's position inside the quad
        // out[P] = sum_t D2[P + (dy,dx)][t]  <=>  source S adds its tap-t product to pixel S - (dy,dx).
        // Group R owns target rows {-1, 0} (R = 0) or {1, 2} (R = 1): from a source in row spy it needs the
        // taps with spy - dy in that set -- 6 taps (24 columns) of the two near blocks, 3 taps of the far ones.
        uint32_t v[32];
        constexpr bool near_blk_c = false; (void)near_blk_c;
        const bool near_blk = spy == R;                                // compile time after unrolling
        const int tap0 = R == 0 ? (near_blk ? 3 : 6) : 0;             // first tap held in v
        const uint32_t tad = tmem_base + kD2Col + (uint32_t)blk * 64u + (uint32_t)tap0 * 4u + ((uint32_t)(q * 32) << 16);
        if (R == 0) {
          if (near_blk) tmem_ld32(tad, v); else tmem_ld16(tad, v);      // taps 3..8 (+2 unused) | taps 6..8 (+1 unused)
        } else {
          if (near_blk) tmem_ld32(tad, v); else tmem_ld16(tad, v);      // taps 0..7 (6 used) | taps 0..3 (3 used)
        }
        tmem_ld_wait();
        if (blk == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_d2empty);
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          const int rr = spy - dy, cc = spx - dx;      // target row / column relative to the quad, in [-1, 2]
          const bool mine = R == 0 ? (rr == -1 || rr == 0) : (rr == 1 || rr == 2);
          if (!mine) continue;
          const bool in_row = rr == R;                 // the quad's own row R (else the row beyond the edge)
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float val = __uint_as_float(v[(t - tap0) * 4 + k]);
            if (in_row) {
              if (cc >= 0 && cc <= 1) own[cc >= 0 && cc <= 1 ? cc : 0][k] += val;
              else side[cc < 0 ? 0 : 1][k] += val;
            } else {
              if (cc >= 0 && cc <= 1) edge[cc >= 0 && cc <= 1 ? cc : 0][k] += val;
              else corner[cc < 0 ? 0 : 1][k] += val;
            }
          }
        }
      }
      TTACC(tb_tmem, t0); t0 = TT0();
      // slots (slot-major: E[slot][quad], conflict-free 16-byte accesses):
      //   0,1 = U[c] (edge of group 0)  2,3 = D[c] (edge of group 1)  4 = UL  5 = UR  6 = DL  7 = DR
      float4* E = reinterpret_cast<float4*>(sm + kOffEx) + (size_t)(it & 1) * (8 * 128);
      E[(R * 2 + 0) * 128 + m] = make_float4(edge[0][0], edge[0][1], edge[0][2], 0.f);
      E[(R * 2 + 1) * 128 + m] = make_float4(edge[1][0], edge[1][1], edge[1][2], 0.f);
      E[(4 + R * 2 + 0) * 128 + m] = make_float4(corner[0][0], corner[0][1], corner[0][2], 0.f);
      E[(4 + R * 2 + 1) * 128 + m] = make_float4(corner[1][0], corner[1][1], corner[1][2], 0.f);
      // horizontal: my column 0 receives the left quad's side[1] (its col 2), column 1 the right quad's side[0]
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float from_left = __shfl_up_sync(0xFFFFFFFFu, side[1][k], 1);
        const float from_right = __shfl_down_sync(0xFFFFFFFFu, side[0][k], 1);
        if (lf) own[0][k] += from_left;
        if (rt) own[1][k] += from_right;
      }
      named_bar_sync(1, 256);
      {
        // row 0 receives from the quad ABOVE what its group 1 sent down (D[c], DL, DR); row 1 from the quad BELOW
        // what its group 0 sent up (U[c], UL, UR)
        const bool vert = R == 0 ? up : dn;
        const int nb = R == 0 ? m - 8 : m + 8;
        const int s_edge = R == 0 ? 2 : 0, s_corner = R == 0 ? 6 : 4;
        if (vert) {
          const float4 e0 = E[(s_edge + 0) * 128 + nb], e1 = E[(s_edge + 1) * 128 + nb];
          own[0][0] += e0.x; own[0][1] += e0.y; own[0][2] += e0.z;
          own[1][0] += e1.x; own[1][1] += e1.y; own[1][2] += e1.z;
          if (lf) {   // the quad diagonally left sent its corner[1] (col 2 of its edge row) to my column 0
            const float4 e = E[(s_corner + 1) * 128 + nb - 1];
            own[0][0] += e.x; own[0][1] += e.y; own[0][2] += e.z;
          }
          if (rt) {   // the quad diagonally right sent its corner[0] (col -1) to my column 1
            const float4 e = E[(s_corner + 0) * 128 + nb + 1];
            own[1][0] += e.x; own[1][1] += e.y; own[1][2] += e.z;
          }
        }
      }
      TTACC(tb_exch, t0); t0 = TT0();
      // ---- bias, residual, stores
      float o[2][3];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) o[j][k] = (own[j][k] + bias_out_s[k]) + yo[j][k];
      if (LRS != 0 && (ok0 || ok1) && !(p.flags & 1)) {
        // upsample_func(lr_curr) at (Y, X0), (Y, X0+1): both lie in one LR cell (2 | LRS); vertical pass first
        // (net_utils.py:144-151).  All 16 x C loads are issued before the first use: one L2 round trip per tile.
        const int ly = Y / LRS, lx = X0 / LRS, dyy = Y - ly * LRS, dxx = X0 - lx * LRS;
        int ro[4], co[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ro[i] = tg_clampi(ly - 1 + i, 0, p.lh - 1) * p.lw;
          co[i] = tg_clampi(lx - 1 + i, 0, p.lw - 1);
        }
        const int plane = p.lh * p.lw;
        const float* pl = p.lr + (size_t)img * p.cout_real * plane;
        float sv[3][4][4];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int kk = k < p.cout_real ? k : 0;             // out_nc < 3: harmless duplicate loads
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sv[k][i][j] = __ldg(pl + kk * plane + ro[i] + co[j]);
        }
        const float ky0 = p.taps[dyy][0], ky1 = p.taps[dyy][1], ky2 = p.taps[dyy][2], ky3 = p.taps[dyy][3];
        const float a0 = p.taps[dxx][0], a1 = p.taps[dxx][1], a2 = p.taps[dxx][2], a3 = p.taps[dxx][3];
        const float b0 = p.taps[dxx + 1][0], b1 = p.taps[dxx + 1][1], b2 = p.taps[dxx + 1][2], b3 = p.taps[dxx + 1][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float col[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) col[j] = ky0 * sv[k][0][j] + ky1 * sv[k][1][j] + ky2 * sv[k][2][j] + ky3 * sv[k][3][j];
          o[0][k] += a0 * col[0] + a1 * col[1] + a2 * col[2] + a3 * col[3];
          o[1][k] += b0 * col[0] + b1 * col[1] + b2 * col[2] + b3 * col[3];
        }
      }
      TTACC(tb_resid, t0); t0 = TT0();
      if (ok0 || ok1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k < p.cout_real) {
            if (ok0 && ok1) *reinterpret_cast<float2*>(dst + k * HWs) = make_float2(o[0][k], o[1][k]);   // X0 even: aligned
            else if (ok0) dst[k * HWs] = o[0][k];
            else dst[k * HWs + 1] = o[1][k];
          }
        }
        if (p.y_u8 != nullptr && !(p.flags & 2)) {
          uint8_t* d8 = p.y_u8 + ((size_t)img * HWs + (size_t)Y * W + X0) * p.cout_real;
          if (p.cout_real == 3 && ok0 && ok1) {         // 6 bytes at an even address: three 16-bit stores
            uint16_t* d16 = reinterpret_cast<uint16_t*>(d8);
            d16[0] = (uint16_t)(q8(o[0][0]) | (q8(o[0][1]) << 8));
            d16[1] = (uint16_t)(q8(o[0][2]) | (q8(o[1][0]) << 8));
            d16[2] = (uint16_t)(q8(o[1][1]) | (q8(o[1][2]) << 8));
          } else {
            for (int k = 0; k < p.cout_real; ++k) {
              if (ok0) d8[k] = (uint8_t)q8(o[0][k]);
              if (ok1) d8[p.cout_real + k] = (uint8_t)q8(o[1][k]);
            }
          }
        }
      }
      TTACC(tb_store, t0);
      t_top0 = TT0();
    }
    if (TIMING && warp == 12 && lane == 0) {
      unsigned long long* o = p.dbg + blockIdx.x * TT_SLOTS;
      o[TT_EB_WAIT] = tb_wait; o[TT_EB_TMEM] = tb_tmem; o[TT_EB_EXCH] = tb_exch; o[TT_EB_RESID] = tb_resid;
      o[TT_EB_STORE] = tb_store; o[TT_EB_TOTAL] = clock64() - t_eb0; o[TT_EB_TOP] = tb_top;
    }
    };
    if (warp < 16) run_b(std::integral_constant<int, 0>{}); else run_b(std::integral_constant<int, 1>{});
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  if (TIMING && threadIdx.x == 0) p.dbg[blockIdx.x * TT_SLOTS + TT_KERNEL] = clock64() - t_kernel0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tail_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    tried = true;
  }
  return fn;
}

}  // namespace

unsigned long long* tg_conv_timer_buffer();     // tg_conv_tcgen05.cu (tg_debug_set_conv_timers)

extern "C" {

int tg_convT_convout_tcgen05(const tg_tail_desc* d, void* stream) {
  TG_REQUIRE(d != nullptr, TG_E_INVALID, "convT_convout: null descriptor");
  TG_REQUIRE(d->x && d->w_up && d->b_up && d->w_out && d->b_out && d->y, TG_E_INVALID, "convT_convout: null pointer");
  TG_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, TG_E_INVALID, "convT_convout: bad size");
  TG_REQUIRE(d->cout_real >= 1 && d->cout_real <= 3, TG_E_UNSUPPORTED, "convT_convout: out_nc=%d (1..3)", d->cout_real);
  TG_REQUIRE(d->reserved == 0, TG_E_INVALID, "convT_convout: reserved must be 0");
  TG_REQUIRE(!(d->accumulate && d->lr), TG_E_INVALID, "convT_convout: accumulate and lr are exclusive");
  TG_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->w_up & 15) == 0 && ((uintptr_t)d->w_out & 15) == 0 &&
                 ((uintptr_t)d->y & 7) == 0, TG_E_INVALID, "convT_convout: pointer alignment");
  if (d->lr != nullptr) {
    TG_REQUIRE(d->lr_scale == 2 || d->lr_scale == 4, TG_E_UNSUPPORTED, "convT_convout: lr_scale %d (2 or 4)", d->lr_scale);
    TG_REQUIRE((2 * d->h) % d->lr_scale == 0 && (2 * d->w) % d->lr_scale == 0, TG_E_INVALID,
               "convT_convout: output size is not lr_scale x the LR size");
    TG_REQUIRE(d->up_mode == TG_UP_BICUBIC || d->up_mode == TG_UP_BILINEAR, TG_E_INVALID, "convT_convout: up_mode");
  }
  TailParams p;
  p.w_up = reinterpret_cast<const unsigned char*>(d->w_up);
  p.w_out = reinterpret_cast<const unsigned char*>(d->w_out);
  p.b_up = d->b_up; p.b_out = d->b_out; p.lr = d->lr; p.y = d->y; p.y_u8 = d->y_u8;
  p.n = d->n; p.h = d->h; p.w = d->w; p.cout_real = d->cout_real;
  p.lr_scale = d->lr ? d->lr_scale : 2; p.up_mode = d->up_mode;
  p.lh = 2 * d->h / p.lr_scale; p.lw = 2 * d->w / p.lr_scale;
  // HR rows -1 .. 2h-1 are covered in strips of 30 (the first strip starts at the even row -2)
  p.tiles_y = tg_ceil_div(2 * d->h + 1, 2 * kStepY);
  p.tiles_x = tg_ceil_div(2 * d->w + 1, 2 * kStepX);
  p.num_tiles = p.tiles_x * p.tiles_y * d->n;
  p.idesc_up = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc_out = (1u << 4) | ((uint32_t)(TG_TAPN_ROWS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  for (int dd = 0; dd < 5; ++dd) {        // host restatement of tg_up_taps (tg_common.cuh)
    float* k = p.taps[dd];
    k[0] = k[1] = k[2] = k[3] = 0.f;
    const int S = p.lr_scale;
    if (dd >= S) continue;
    if (p.up_mode == TG_UP_BICUBIC) {
      const float a = -0.75f, t = (float)dd / (float)S, t2 = t * t, t3 = t2 * t;
      k[0] = a * t - 2.f * a * t2 + a * t3;
      k[1] = 1.f - (a + 3.f) * t2 + (a + 2.f) * t3;
      k[2] = -a * t + (2.f * a + 3.f) * t2 - (a + 2.f) * t3;
      k[3] = a * t2 - a * t3;
    } else {
      const float src = ((float)dd + 0.5f) / (float)S - 0.5f;
      if (src < 0.f) { const float f = src + 1.f; k[0] = 1.f - f; k[1] = f; }
      else           { k[1] = 1.f - src; k[2] = src; }
    }
  }
  p.accumulate = d->accumulate;
  p.flags = 0;
  if (const char* e = getenv("TG_TAIL_FLAGS")) p.flags = atoi(e);
  p.dbg = tg_conv_timer_buffer();

  EncodeTiledFn fn = tail_encode_fn();
  TG_REQUIRE(fn != nullptr, TG_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {64, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
  cuuint64_t strides[3] = {128, (cuuint64_t)d->w * 128, (cuuint64_t)d->h * d->w * 128};
  cuuint32_t box[4] = {64, (cuuint32_t)kBoxW, (cuuint32_t)kBoxH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&p.map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->x), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TG_REQUIRE(r == CUDA_SUCCESS, TG_E_DRIVER, "convT_convout: cuTensorMapEncodeTiled failed (%d)", (int)r);

  static TgPerDeviceOnce attr_once;
  const cudaError_t attr_err = attr_once.run([] {
    cudaError_t e, err = cudaSuccess;
#define TG_TAIL_ATTR(S, M)                                                                                              \
    e = cudaFuncSetAttribute(tail_tcgen05_kernel<false, S, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes); \
    if (e != cudaSuccess) err = e;                                                                                        \
    e = cudaFuncSetAttribute(tail_tcgen05_kernel<true, S, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);  \
    if (e != cudaSuccess) err = e;
    TG_TAIL_ATTR(0, 0) TG_TAIL_ATTR(2, TG_UP_BICUBIC) TG_TAIL_ATTR(4, TG_UP_BICUBIC) TG_TAIL_ATTR(2, TG_UP_BILINEAR)
    TG_TAIL_ATTR(4, TG_UP_BILINEAR)
#undef TG_TAIL_ATTR
    return err;
  });
  TG_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "convT_convout: cudaFuncSetAttribute: %s",
             cudaGetErrorString(attr_err));
  int sms = 0;
  int rc = tg_device_sm_count(&sms);
  if (rc != TG_OK) return rc;
  int grid = d->max_ctas > 0 && d->max_ctas < sms ? d->max_ctas : sms;
  if (grid > p.num_tiles) grid = p.num_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t lerr;
#define TG_TAIL_LAUNCH(S, M)                                                                                      \
  lerr = p.dbg ? tg_launch(tail_tcgen05_kernel<true, S, M>, dim3(grid), dim3(kThreads), kSmemBytes, st, p)        \
               : tg_launch(tail_tcgen05_kernel<false, S, M>, dim3(grid), dim3(kThreads), kSmemBytes, st, p)
  if (d->lr == nullptr) TG_TAIL_LAUNCH(0, 0);
  else if (d->lr_scale == 4 && d->up_mode == TG_UP_BICUBIC) TG_TAIL_LAUNCH(4, TG_UP_BICUBIC);
  else if (d->lr_scale == 2 && d->up_mode == TG_UP_BICUBIC) TG_TAIL_LAUNCH(2, TG_UP_BICUBIC);
  else if (d->lr_scale == 2) TG_TAIL_LAUNCH(2, TG_UP_BILINEAR);
  else TG_TAIL_LAUNCH(4, TG_UP_BILINEAR);
#undef TG_TAIL_LAUNCH
  TG_REQUIRE(lerr == cudaSuccess, (int)lerr, "convT_convout: launch failed: %s", cudaGetErrorString(lerr));
  TG_CUDA_LAUNCH_CHECK("convT_convout");
  return TG_OK;
}

}  // extern "C"
