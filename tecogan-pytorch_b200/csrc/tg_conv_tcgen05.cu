// 3x3 convolution / stride-2 transposed convolution as a persistent, warp-specialised tcgen05
// implicit GEMM for sm_100a.
//
//   M tile  = 128 output pixels = a 16x8 patch of one image (TMEM lane m <-> pixel (m>>3, m&7))
//   N       = 64 output channels per CTA (layers with 128 / 256 are split over 2 / 4 CTAs); 48 for
//             the thin heads (MODE_TAPN: 9 taps x 4 couts in N, 3x3 shift-add in the epilogue)
//   K       = 64-channel chunks x 9 taps; UMMA K = 16 -> 4 MMAs per (tap, chunk)
//   A       : NHWC fp16 activations, fetched by TMA (4-D tiled map, 128B swizzle, OOB zero fill =
//             the conv's zero padding) either as ONE halo box (18x10 px) per (tile, chunk) whose
//             nine shifted views are addressed through the UMMA descriptor (start address +=
//             (dy*10+dx)*128 B, 8-row-group stride = 10*128 B), or as one 16x8 box per tap.
//   B       : weights pre-packed on the device in the exact swizzled smem image
//             (tg_pack_*_weights), either resident in smem for the whole kernel (SRNet, thin
//             FNet layers) or streamed per (tap, chunk) with cp.async.bulk (fat FNet layers).
//   D       : fp32 accumulators in TMEM, up to 8 tile buffers in flight (the epilogue of tile i
//             overlaps the MMAs of tiles i+1..).  The transposed conv keeps 4 parity accumulators
//             (1/2/2/4 taps); each epilogue thread stores its pixel's 2x2 outputs (pixel shuffle).
//   roles   : warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane; HALO convs issue two
//             tiles interleaved, thin heads four), warp 2 = TMEM allocator, warps 4.. = 2 epilogue
//             groups of 4 warps (4 groups for the transposed conv and the thin heads) taking tiles
//             round-robin: tcgen05.ld -> bias/act/residual -> fp16 -> the pixel's 128-byte NHWC
//             row straight to global with 256-bit stores; or the thin-head TAPN epilogues.
//
// Replaces the nn.Conv2d / nn.ConvTranspose2d library calls K1, K10, K11, K12 of SURVEY.md 2.1.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "tg_common.cuh"
#include "tg_epilogue.cuh"
#include "tg_tcgen05.cuh"

namespace {

constexpr int TH = 16, TW = 8;
// 4 control warps (TMA producer, MMA issuer, TMEM allocator, spare) + G epilogue groups of 4 warps.
// The transposed conv (4 accumulators per tile) and the thin heads (tiny MMA work per tile) are
// epilogue-bound: they run 4 groups (640 threads, <= 102 registers); the plain convs 2 groups.
__host__ __device__ constexpr int epi_groups(int kind, int mode) {
  return (mode == 2 /*MODE_TAPN*/ || kind == TG_CONVT_3X3_S2) ? 4 : 2;
}
// up to four tensor maps: [0] = the NHWC input; TG_CONV_3X3_S2 reads the input's four parity planes
struct TgMaps { CUtensorMap m[4]; };
__host__ __device__ constexpr int conv_threads(int kind, int mode) { return 128 + 128 * epi_groups(kind, mode); }
constexpr int kMaxStages = 8;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kHeaderBytes = 2048;   // barriers + tmem ptr (first 1 KB) + bias (second 1 KB)
constexpr uint32_t kTapABytes = TH * TW * 128;  // 16 KB
constexpr uint32_t kSmemLimit = 232448;   // 227 KB opt-in limit per CTA
// A-operand modes (template parameter MODE)
constexpr int MODE_TAP = 0;    // one 16x8 box per (tap, chunk)
constexpr int MODE_HALO = 1;   // one 18x10 halo box per (tile, chunk), taps = descriptor shifts
constexpr int MODE_TAPN = 2;   // thin heads: one 16x8 box per (tile, chunk), N = 9 taps x 4 couts,
                               // 3x3 shift-add in the epilogue; tiles overlap by one pixel ring
constexpr int kTapnStepY = TH - 2, kTapnStepX = TW - 2;   // 14 x 6 valid outputs per TAPN tile
constexpr uint32_t kTapnEBytes = 128 * 9 * 16;            // exchange buffer: [128 px][9 taps] float4

struct KParams {
  tg_conv_desc d;
  int tiles_x, tiles_y, num_tiles;
  int chunks, n_acc;
  int halo, b_resident;
  int box_w, box_h, org_x, org_y;
  int step_y, step_x;              // output pixels a tile advances by (16x8; 14x6 for MODE_TAPN)
  uint32_t acc_stride;             // TMEM columns per accumulator buffer
  int n_buf;                       // accumulator buffers in flight (512 / acc_stride, <= 8, even)
  int dbg_flags;                   // diagnostics (TG_DBG_FLAGS): 1 = TAPN skip global RMW, 2 = skip exchange
  int n_stages;
  int ksteps;                      // UMMA k-steps (16 channels) per 64-channel chunk that can hold non-zero input (1..4)
  int n_split, bn;                 // output channels are split over n_split CTAs of bn columns
  uint32_t stage_bytes, a_bytes, b_tile_bytes, b_stage_bytes;
  uint32_t off_b, off_stage, off_staging;
  uint32_t idesc;
  unsigned long long* dbg;         // optional per-CTA role timers (tg_debug_set_conv_timers), else null
};

// role-timer slots (cycles, per CTA): see tools/conv_timers.py
enum { T_PROD_WAIT_EMPTY = 0, T_MMA_WAIT_TEMPTY, T_MMA_WAIT_FULL, T_MMA_ISSUE, T_MMA_TOTAL,
       T_EPI_WAIT_STORE, T_EPI_WAIT_TFULL, T_EPI_COMPUTE, T_EPI_STORE, T_EPI_TOTAL, T_KERNEL, T_PROLOGUE,
       T_TILES, T_SLOTS = 16 };
#define TG_T0() (timing ? clock64() : 0)
#define TG_ACC(var, t0) do { if (timing) var += clock64() - (t0); } while (0)

// two MMA issuer warps (1 and 3) instead of one: halo convs whose tile is one smem stage, with even stage and
// accumulator counts (TG_DBG_FLAGS bit 16 = single issuer, for A/B measurements)
template <int MODE>
__device__ __forceinline__ bool conv_dual_issue(const KParams& p) {
  return MODE == MODE_HALO && p.chunks == 1 && p.n_stages >= 2 && (p.n_stages & 1) == 0 && p.n_buf >= 2 &&
         !(p.dbg_flags & 16);
}

struct TileCoord { int n, y0, x0, nb; };
__device__ __forceinline__ TileCoord tile_coord(const KParams& p, int tile) {
  TileCoord t;
  const int per_img = p.tiles_x * p.tiles_y;
  const int sp = tile / p.n_split;          // CTAs that share an A tile are adjacent (L2 reuse)
  t.nb = tile - sp * p.n_split;
  t.n = sp / per_img;
  const int r = sp - t.n * per_img;
  t.y0 = (r / p.tiles_x) * p.step_y;
  t.x0 = (r % p.tiles_x) * p.step_x;
  return t;
}

// ------------------------------------------------------------------ the kernel
// BWD = data-gradient instantiation: epilogue y = (acc + bias [+ residual]) * act'(mask)  (TG_ACT_DRELU /
// TG_ACT_DLRELU02; TG_ACT_NONE = no derivative).  Kept out of the forward instantiations so their
// register allocation and code are untouched.
// POOL = TG_EPI_NHWC_F16_POOL2 instantiation: 2x2 max over the tile's pixels by warp shuffles (lane ^ 1 = x
// neighbour, lane ^ 8 = y neighbour: a warp holds four 8-pixel rows of the 16x8 tile), one store per 2x2 block.
template <int KIND, int MODE, bool TIMING, bool BWD = false, bool POOL = false>
__global__ void __launch_bounds__(conv_threads(KIND, MODE), 1)
conv_tcgen05_kernel(const __grid_constant__ TgMaps maps, const KParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;   // 128B swizzle atoms need 1024B alignment
  uint8_t* sm = smem_raw + (base - raw);

  // shuffle broadcast: the compiler then knows the warp index is warp-uniform and keeps role-loop counters,
  // barrier addresses and UMMA descriptors in uniform registers (no R2UR in front of every MMA)
  const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const tg_conv_desc& d = p.d;
  constexpr bool timing = TIMING;   // role timers compiled out of the production instantiation
  const long long t_kernel0 = timing ? clock64() : 0;

  // header: barriers
  const uint32_t bar_full = base;                       // [kMaxStages]
  const uint32_t bar_empty = base + 8 * kMaxStages;     // [kMaxStages]
  const uint32_t bar_tfull = base + 16 * kMaxStages;    // [8]
  const uint32_t bar_tempty = bar_tfull + 64;           // [8]
  const uint32_t bar_b = bar_tempty + 64;               // [1]
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + 16 * kMaxStages + 136);
  float* bias_s = reinterpret_cast<float*>(sm + 1024);

  // warps that release an accumulator buffer: the 4 warps of the group that owns the tile (for the
  // transposed conv two groups share every tile, two parity accumulators each -> 8 warps)
  const int epi_warps_active = KIND == TG_CONVT_3X3_S2 ? 8 : 4;
  static_assert(!(KIND == TG_CONV_3X3_S2 && MODE != MODE_TAP), "the stride-2 conv runs in tap mode only");

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.m[0]);
    if (KIND == TG_CONV_3X3_S2) { tma_prefetch_desc(&maps.m[1]); tma_prefetch_desc(&maps.m[2]); tma_prefetch_desc(&maps.m[3]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < p.n_buf; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, epi_warps_active);
    }
    mbar_init(bar_b, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_s)), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (timing && threadIdx.x == 0) p.dbg[blockIdx.x * T_SLOTS + T_PROLOGUE] = clock64() - t_kernel0;

  const uint32_t smem_b = base + p.off_b;
  const uint32_t smem_stage0 = base + p.off_stage;
  const unsigned char* wglob = reinterpret_cast<const unsigned char*>(d.weights);
  const int n_tiles_w = (MODE == MODE_TAPN ? 1 : 9) * p.chunks;

  // Resident weights do not depend on the previous kernel (they are static during graph replay; on
  // the eager path tg_pack_* never triggers its dependents early and a bias copy separates it from
  // the conv): start their load, then join the programmatic-dependent-launch wait.  Everything up to here (barrier init, TMEM allocation, bias,
  // weights) overlaps the predecessor's tail; its OUTPUT is only read after tg_pdl_wait().
  if (warp == 0 && lane == 0 && p.b_resident) {
    // this CTA's 64-row slice of every weight tile (nb is fixed per CTA: gridDim.x % n_split == 0)
    const uint32_t nb = blockIdx.x % (uint32_t)p.n_split;
    mbar_expect_tx(bar_b, (uint32_t)n_tiles_w * p.b_stage_bytes);
    for (int t = 0; t < n_tiles_w; ++t)
      bulk_load(smem_b + t * p.b_stage_bytes, wglob + (size_t)t * p.b_tile_bytes + (size_t)nb * p.b_stage_bytes,
                p.b_stage_bytes, bar_b);
  }
  tg_pdl_wait();
  tg_pdl_trigger();
  // bias may have been written by the immediately preceding kernel (host-side refresh): read it
  // only after the wait.  The epilogue warps that consume bias_s pass the barrier below first.
  for (int i = threadIdx.x; i < d.cout; i += conv_threads(KIND, MODE)) bias_s[i] = d.bias[i];
  __syncthreads();

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long tw = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(p, tile);
        if (MODE != MODE_TAP) {
          for (int c = 0; c < p.chunks; ++c) {
            const long long t0 = TG_T0();
            mbar_wait(bar_empty + 8 * stage, phase ^ 1, 1);
            TG_ACC(tw, t0);
            mbar_expect_tx(bar_full + 8 * stage, p.a_bytes);
            tma_load_4d(smem_stage0 + stage * p.stage_bytes, &maps.m[0], bar_full + 8 * stage, c * 64,
                        tc.x0 + p.org_x, tc.y0 + p.org_y, tc.n);
            if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
          }
        } else {
#pragma unroll
          for (int g = 0; g < 9; ++g) {
            constexpr int kDummy = 0; (void)kDummy;
            const TgGroup gr = tg_group(KIND, g);
            for (int c = 0; c < p.chunks; ++c) {
              const long long t0 = TG_T0();
              mbar_wait(bar_empty + 8 * stage, phase ^ 1, 2);
              TG_ACC(tw, t0);
              const uint32_t sa = smem_stage0 + stage * p.stage_bytes;
              mbar_expect_tx(bar_full + 8 * stage, p.a_bytes + (p.b_resident ? 0u : p.b_stage_bytes));
              tma_load_4d(sa, &maps.m[KIND == TG_CONV_3X3_S2 ? tg_s2_plane(g) : 0], bar_full + 8 * stage, c * 64,
                          tc.x0 + gr.dx, tc.y0 + gr.dy, tc.n);
              if (!p.b_resident)
                bulk_load(sa + kTapABytes,
                          wglob + (size_t)(g * p.chunks + c) * p.b_tile_bytes + (size_t)tc.nb * p.b_stage_bytes,
                          p.b_stage_bytes, bar_full + 8 * stage);
              if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      if (timing) p.dbg[blockIdx.x * T_SLOTS + T_PROD_WAIT_EMPTY] = tw;
    }
  } else if (warp == 1 || (warp == 3 && conv_dual_issue<MODE>(p))) {
    // ============================================================ MMA issuer(s)
    // One warp consumes the smem stages strictly in order (a parity wait is only sound while the
    // waiter is at most one phase ahead of the barrier).  The fixed latencies between tiles
    // (mbarrier polls, commit -> epilogue hand-off, MMA pipeline fill/drain) are hidden by depth
    // instead: up to n_buf accumulator buffers are in flight in TMEM.  The warp walks the
    // (warp-uniform) pipeline; one elected lane issues.
    if (conv_dual_issue<MODE>(p)) {
      // TWO issuers (warps 1 and 3) take the tiles of this CTA alternately: an N=64 MMA holds the tensor pipe
      // for 48 cycles and the pipe hides only ~180 cycles without a new instruction, while one warp needs
      // ~80 cycles per MMA for issue + bookkeeping (tools/mma_probe.cu) -- with one issuer the pipe idles
      // 40 % of the time.  Stage it % n_stages, accumulator it % n_buf, both counts even: a stage / buffer
      // always belongs to the same issuer, who therefore sees every completion of the barriers it waits on
      // by parity.  All lanes walk the loop on warp-uniform values (uniform datapath), one elected lane issues.
      const int w = warp == 1 ? 0 : 1;
      if (p.b_resident) { mbar_wait(bar_b, 0, 3); }
      constexpr int kBoxW = (KIND == TG_CONV_3X3) ? TW + 2 : TW + 1;
      constexpr int kOrg = (KIND == TG_CONV_3X3) ? -1 : 0;
      const uint32_t a_hi32 = (uint32_t)(make_sdesc(0, (uint32_t)kBoxW * 128u) >> 32);
      const uint32_t b_hi32 = (uint32_t)(make_sdesc(0, 1024u) >> 32);
      const uint32_t a_lo0 = ((smem_stage0 & 0x3FFFFu) >> 4) + 0x10000u;   // + LBO field (1 << 16)
      const uint32_t b_lo0 = ((smem_b & 0x3FFFFu) >> 4) + 0x10000u;
      const uint32_t btb16 = p.b_stage_bytes >> 4;
      const uint32_t st16 = p.stage_bytes >> 4;
      int stage = w, buf = w, it = 0;
      uint32_t phase = 0, bphase = 0;
      long long tw_tempty = 0, tw_full = 0, t_issue = 0;
      const bool tm = timing && w == 0;
      const long long t_mma0 = tm ? clock64() : 0;
      for (int tile = blockIdx.x + w * (int)gridDim.x; tile < p.num_tiles; tile += 2 * (int)gridDim.x, ++it) {
        long long t0 = tm ? clock64() : 0;
        mbar_wait(bar_tempty + 8 * buf, bphase ^ 1, 4);
        if (tm) { tw_tempty += clock64() - t0; t0 = clock64(); }
        mbar_wait(bar_full + 8 * stage, phase, 5);
        if (tm) { tw_full += clock64() - t0; t0 = clock64(); }
        tc_fence_after();
        const uint32_t sa = a_lo0 + (uint32_t)stage * st16;
        const uint32_t d_base = tmem_base + (uint32_t)buf * p.acc_stride;
        if (elect_one_sync()) {
#pragma unroll
          for (int g = 0; g < 9; ++g) {
            const TgGroup gr = tg_group(KIND, g);
            const bool first_of_acc = (g == 0) || (tg_group(KIND, g > 0 ? g - 1 : 0).acc != gr.acc);
            const uint32_t off = (uint32_t)((gr.dy - kOrg) * kBoxW + (gr.dx - kOrg)) * 8u;
            const uint32_t dcol = d_base + (uint32_t)gr.acc * (uint32_t)p.bn;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < p.ksteps) umma_f16_words(dcol, sa + off + 2u * k, a_hi32, b_lo0 + (uint32_t)g * btb16 + 2u * k, b_hi32, p.idesc,
                             (first_of_acc && k == 0) ? 0u : 1u);
          }
          umma_commit(bar_empty + 8 * stage);
          umma_commit(bar_tfull + 8 * buf);
        }
        __syncwarp();
        if (tm) t_issue += clock64() - t0;
        stage += 2;
        if (stage >= p.n_stages) { stage -= p.n_stages; phase ^= 1u; }
        buf += 2;
        if (buf >= p.n_buf) { buf -= p.n_buf; bphase ^= 1u; }
      }
      if (tm && lane == 0) {
        unsigned long long* o = p.dbg + blockIdx.x * T_SLOTS;
        o[T_MMA_WAIT_TEMPTY] = tw_tempty; o[T_MMA_WAIT_FULL] = tw_full; o[T_MMA_ISSUE] = t_issue;
        o[T_MMA_TOTAL] = clock64() - t_mma0; o[T_TILES] = 2 * it;
      }
    } else {
      if (p.b_resident) { mbar_wait(bar_b, 0, 3); }
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t acc_stride = p.acc_stride;
      // descriptor templates: everything but the 14-bit start address is constant per kernel
      constexpr int kBoxW = (KIND == TG_CONV_3X3) ? TW + 2 : TW + 1;
      constexpr int kOrg = (KIND == TG_CONV_3X3) ? -1 : 0;
      const uint64_t a_hi = make_sdesc(0, (MODE == MODE_HALO && !(p.dbg_flags & 8)) ? (uint32_t)kBoxW * 128u : 1024u);
      const uint64_t b_hi = make_sdesc(0, 1024u);
      const uint32_t btb16 = p.b_stage_bytes >> 4;      // resident weight tiles are 64-row slices
      const uint32_t smem_b16 = (smem_b & 0x3FFFFu) >> 4;
      long long tw_tempty = 0, tw_full = 0, t_issue = 0;
      const long long t_mma0 = TG_T0();
      // Software-pipelined barrier polling: the barriers of the NEXT smem stage (and, at a tile
      // boundary, the next tile's TMEM buffer) are polled while MMAs of the current stage are still
      // queued in the tensor pipe, so the poll latencies (~100-200 cycles each) do not drain it.
      int buf = 0;
      uint32_t bphase = 0;
      const bool pair_mode = MODE == MODE_HALO && p.chunks == 1 && p.n_stages >= 2 && p.n_buf >= 2 &&
                             !(p.dbg_flags & 16);
      if (pair_mode) {
        // Two tiles are issued interleaved, MMA by MMA, into two TMEM accumulators.  Back-to-back
        // MMAs that accumulate into the SAME accumulator serialise on its read-modify-write
        // (measured ~77 cycles per 128x64x16 MMA); alternating between two independent
        // accumulators lets the tensor pipe overlap them (the transposed conv, whose taps already
        // spread over 4 accumulators, runs the same MMAs at ~48 cycles).
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
          const bool two = tile + (int)gridDim.x < p.num_tiles;
          const int bufA = buf;
          const uint32_t bphA = bphase;
          const int bufB = (bufA + 1 == p.n_buf) ? 0 : bufA + 1;
          const uint32_t bphB = (bufA + 1 == p.n_buf) ? (bphA ^ 1u) : bphA;
          const int stA = stage;
          const uint32_t phA = phase;
          const int stB = (stA + 1 == p.n_stages) ? 0 : stA + 1;
          const uint32_t phB = (stA + 1 == p.n_stages) ? (phA ^ 1u) : phA;
          long long t0 = TG_T0();
          mbar_wait(bar_tempty + 8 * bufA, bphA ^ 1, 4);
          if (two) mbar_wait(bar_tempty + 8 * bufB, bphB ^ 1, 4);
          TG_ACC(tw_tempty, t0);
          t0 = TG_T0();
          mbar_wait(bar_full + 8 * stA, phA, 5);
          if (two) mbar_wait(bar_full + 8 * stB, phB, 5);
          TG_ACC(tw_full, t0);
          tc_fence_after();
          const uint32_t saA = ((smem_stage0 + stA * p.stage_bytes) & 0x3FFFFu) >> 4;
          const uint32_t saB = ((smem_stage0 + stB * p.stage_bytes) & 0x3FFFFu) >> 4;
          const uint32_t dA = tmem_base + bufA * acc_stride, dB = tmem_base + bufB * acc_stride;
          const long long t_i0 = TG_T0();
          if (elect_one_sync()) {
#pragma unroll
            for (int g = 0; g < 9; ++g) {
              const TgGroup gr = tg_group(KIND, g);
              const bool first_of_acc = (g == 0) || (tg_group(KIND, g > 0 ? g - 1 : 0).acc != gr.acc);
              const uint32_t off = (uint32_t)((gr.dy - kOrg) * kBoxW + (gr.dx - kOrg)) * 8u;
              const uint32_t b16 = smem_b16 + (uint32_t)g * btb16;
              const uint32_t dcol = (uint32_t)gr.acc * (uint32_t)p.bn;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k >= p.ksteps) continue;
                const uint32_t accf = (first_of_acc && k == 0) ? 0u : 1u;
                umma_f16(dA + dcol, a_hi | (uint64_t)(saA + off + 2u * k), b_hi | (uint64_t)(b16 + 2u * k), p.idesc, accf);
                if (two)
                  umma_f16(dB + dcol, a_hi | (uint64_t)(saB + off + 2u * k), b_hi | (uint64_t)(b16 + 2u * k), p.idesc, accf);
              }
            }
            umma_commit(bar_empty + 8 * stA);
            if (two) umma_commit(bar_empty + 8 * stB);
            umma_commit(bar_tfull + 8 * bufA);
            if (two) umma_commit(bar_tfull + 8 * bufB);
          }
          __syncwarp();
          TG_ACC(t_issue, t_i0);
          if (two) {
            stage = (stB + 1 == p.n_stages) ? 0 : stB + 1;
            phase = (stB + 1 == p.n_stages) ? (phB ^ 1u) : phB;
            buf = (bufB + 1 == p.n_buf) ? 0 : bufB + 1;
            bphase = (bufB + 1 == p.n_buf) ? (bphB ^ 1u) : bphB;
          } else {
            it -= 1;      // only one tile issued in this round (it is advanced by 2 below)
          }
        }
      } else if (MODE == MODE_TAPN && p.chunks == 1 && p.n_stages >= 4 && p.n_buf >= 4 && !(p.dbg_flags & 16)) {
        // Thin heads: a tile is only 4 MMAs, so the fixed poll / commit latencies of the issue loop
        // dominate; issue up to four tiles per round (4 accumulators, 4 smem stages).
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += 4 * gridDim.x) {
          int cnt = 0, bufs[4], sts[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (tile + j * (int)gridDim.x < p.num_tiles) {
              bufs[j] = buf; sts[j] = stage;
              mbar_wait(bar_tempty + 8 * buf, bphase ^ 1, 4);
              mbar_wait(bar_full + 8 * stage, phase, 5);
              if (++buf == p.n_buf) { buf = 0; bphase ^= 1u; }
              if (++stage == p.n_stages) { stage = 0; phase ^= 1u; }
              cnt = j + 1;
            }
          }
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < cnt) {
                const uint32_t sa16 = ((smem_stage0 + sts[j] * p.stage_bytes) & 0x3FFFFu) >> 4;
                const uint32_t dcol = tmem_base + bufs[j] * acc_stride;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < p.ksteps) umma_f16(dcol, a_hi | (uint64_t)(sa16 + 2u * k), b_hi | (uint64_t)(smem_b16 + 2u * k), p.idesc,
                           k == 0 ? 0u : 1u);
                umma_commit(bar_empty + 8 * sts[j]);
                umma_commit(bar_tfull + 8 * bufs[j]);
              }
            }
          }
          __syncwarp();
          it += cnt;
        }
      } else {
      if (blockIdx.x < p.num_tiles) {
        mbar_wait(bar_tempty, 1, 4);              // fresh barrier: passes immediately
        mbar_wait(bar_full, 0, 5);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const bool has_next = tile + (int)gridDim.x < p.num_tiles;
        const int nbuf = (buf + 1 == p.n_buf) ? 0 : buf + 1;
        const uint32_t nbphase = (buf + 1 == p.n_buf) ? (bphase ^ 1u) : bphase;
        const uint32_t d_base = tmem_base + buf * acc_stride;
        if (MODE != MODE_TAP) {
          constexpr int NG = MODE == MODE_TAPN ? 1 : 9;   // TAPN: one group, all taps live in N
          constexpr int NG_HEAD = NG > 2 ? NG - 2 : 0;    // groups issued before the look-ahead poll
          for (int c = 0; c < p.chunks; ++c) {
            const bool last = c == p.chunks - 1;
            const int nstage = (stage + 1 == p.n_stages) ? 0 : stage + 1;
            const uint32_t nphase = (stage + 1 == p.n_stages) ? (phase ^ 1u) : phase;
            const uint32_t sa16 = ((smem_stage0 + stage * p.stage_bytes) & 0x3FFFFu) >> 4;
            const long long t_i0 = TG_T0();
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              if (elect_one_sync()) {
#pragma unroll
                for (int g = (part == 0 ? 0 : NG_HEAD); g < (part == 0 ? NG_HEAD : NG); ++g) {
                  const TgGroup gr = MODE == MODE_TAPN ? TgGroup{0, 0, 0, 0, 0} : tg_group(KIND, g);
                  // first MMA into an accumulator (per tile) overwrites, the rest accumulate
                  const bool first_of_acc = (g == 0) || (tg_group(KIND, g > 0 ? g - 1 : 0).acc != gr.acc);
                  const uint32_t a16 = (MODE == MODE_TAPN || (p.dbg_flags & 4))   // flag 4: timing experiment only
                                           ? sa16
                                           : sa16 + (uint32_t)((gr.dy - kOrg) * kBoxW + (gr.dx - kOrg)) * 8u;
                  const uint32_t b16 = smem_b16 + (uint32_t)(g * p.chunks + c) * btb16;
                  const uint32_t dcol = d_base + (uint32_t)gr.acc * (uint32_t)p.bn;
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (k < p.ksteps) umma_f16(dcol, a_hi | (uint64_t)(a16 + 2u * k), b_hi | (uint64_t)(b16 + 2u * k), p.idesc,
                             (first_of_acc && k == 0 && c == 0) ? 0u : 1u);
                }
                if (part == 1) {
                  umma_commit(bar_empty + 8 * stage);
                  if (last) umma_commit(bar_tfull + 8 * buf);
                }
              }
              __syncwarp();
              if (part == 0) {
                TG_ACC(t_issue, t_i0);
                if (!last || has_next) {
                  long long t0 = TG_T0();
                  if (last) { mbar_wait(bar_tempty + 8 * nbuf, nbphase ^ 1, 4); TG_ACC(tw_tempty, t0); t0 = TG_T0(); }
                  mbar_wait(bar_full + 8 * nstage, nphase, 5);
                  TG_ACC(tw_full, t0);
                  tc_fence_after();
                }
              }
            }
            stage = nstage; phase = nphase;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 9; ++g) {
            const TgGroup gr = tg_group(KIND, g);
            const bool first_of_acc = (g == 0) || (tg_group(KIND, g > 0 ? g - 1 : 0).acc != gr.acc);
            const uint32_t dcol = d_base + (uint32_t)gr.acc * (uint32_t)p.bn;
            for (int c = 0; c < p.chunks; ++c) {
              const bool last = g == 8 && c == p.chunks - 1;
              const int nstage = (stage + 1 == p.n_stages) ? 0 : stage + 1;
              const uint32_t nphase = (stage + 1 == p.n_stages) ? (phase ^ 1u) : phase;
              const uint32_t sa = smem_stage0 + stage * p.stage_bytes;
              const uint32_t a16 = (sa & 0x3FFFFu) >> 4;
              const uint32_t b16 = p.b_resident ? smem_b16 + (uint32_t)(g * p.chunks + c) * btb16
                                                : ((sa + kTapABytes) & 0x3FFFFu) >> 4;
              if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < p.ksteps) umma_f16(dcol, a_hi | (uint64_t)(a16 + 2u * k), b_hi | (uint64_t)(b16 + 2u * k), p.idesc,
                           (first_of_acc && k == 0 && c == 0) ? 0u : 1u);
                umma_commit(bar_empty + 8 * stage);
                if (last) umma_commit(bar_tfull + 8 * buf);
              }
              __syncwarp();
              // look ahead: poll the next stage while the 4 MMAs just issued execute
              if (!last || has_next) {
                long long t0 = TG_T0();
                if (last) { mbar_wait(bar_tempty + 8 * nbuf, nbphase ^ 1, 4); TG_ACC(tw_tempty, t0); t0 = TG_T0(); }
                mbar_wait(bar_full + 8 * nstage, nphase, 6);
                TG_ACC(tw_full, t0);
                tc_fence_after();
              }
              stage = nstage; phase = nphase;
            }
          }
        }
        buf = nbuf; bphase = nbphase;
      }
      }   // !pair_mode
      if (timing && lane == 0) {
        unsigned long long* o = p.dbg + blockIdx.x * T_SLOTS;
        o[T_MMA_WAIT_TEMPTY] = tw_tempty; o[T_MMA_WAIT_FULL] = tw_full; o[T_MMA_ISSUE] = t_issue;
        o[T_MMA_TOTAL] = clock64() - t_mma0; o[T_TILES] = it;
      }
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue
    // G groups of 4 warps take the tiles of this CTA round-robin (tile it -> group it % G, TMEM buffer
    // it % n_buf), so one group's TMEM -> registers -> global chain overlaps the other groups' and
    // all overlap the MMAs.  Warp q of a group reads TMEM lanes [32q, 32q+32) = tile rows, all bn
    // columns.
    const int group = (warp - 4) >> 2;
    const int gtid = threadIdx.x - 128 - group * 128;   // 0..127 inside the group
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;   // row of the tile = TMEM lane
    const int ty = r >> 3, tx = r & 7;
    const uint32_t acc_stride = p.acc_stride;
    long long te_store_wait = 0, te_tfull = 0, te_compute = 0, te_store = 0;
    const long long t_epi0 = TG_T0();
    int it = group;
    // A TMEM buffer must always be drained by the same group(s) (parity waits may not run two
    // phases ahead): TG tile-level groups with n_buf % TG == 0.  The transposed conv has only two
    // buffers, so its 4 groups work as 2 pairs -- both groups of a pair take every tile of the pair,
    // group 2j handles parity accumulators 0,1 and group 2j+1 accumulators 2,3.
    constexpr int G = epi_groups(KIND, MODE);
    constexpr bool kSplitAcc = KIND == TG_CONVT_3X3_S2;
    constexpr int TG = kSplitAcc ? G / 2 : G;
    const int tgroup = kSplitAcc ? (group >> 1) : group;
    const int acc_lo = kSplitAcc ? 2 * (group & 1) : 0;
    const int acc_hi = kSplitAcc ? acc_lo + 2 : p.n_acc;
    it = tgroup;
    for (int tile = blockIdx.x + tgroup * gridDim.x; tile < p.num_tiles; tile += TG * gridDim.x, it += TG) {
      const int buf = it % p.n_buf;
      const uint32_t bphase = (uint32_t)(it / p.n_buf) & 1u;
      const TileCoord tc = tile_coord(p, tile);
      // MODE_TAPN: thread = input position (halo ring included); interior positions are outputs
      const int py = tc.y0 + ty + (MODE == MODE_TAPN ? -1 : 0), px = tc.x0 + tx + (MODE == MODE_TAPN ? -1 : 0);
      const bool interior = MODE != MODE_TAPN || (ty >= 1 && ty <= TH - 2 && tx >= 1 && tx <= TW - 2);
      const bool inb = interior && py < d.h && px < d.w;
      long long t_s = TG_T0();
      // operands that come from global memory are fetched BEFORE waiting for the accumulator so
      // their latency hides behind the MMAs of this tile
      uint4 res[8];
      // the residual input exists only for the plain conv (validated on the host)
      constexpr bool kCanRes = KIND != TG_CONVT_3X3_S2 && MODE != MODE_TAPN;
      const bool has_res = kCanRes && (d.epilogue == TG_EPI_NHWC_F16) && (d.residual != nullptr) && inb;
      if (has_res) {
        const uint4* rp = reinterpret_cast<const uint4*>(
            reinterpret_cast<const __half*>(d.residual) +
            (((size_t)tc.n * d.h + py) * d.w + px) * d.cout + tc.nb * p.bn);
#pragma unroll
        for (int i = 0; i < 8; i += 2) ld_global_256(rp + i, res[i], res[i + 1]);
      }
      // derivative epilogue: the stored forward output whose sign gives act' -- fetched per 32-channel piece right
      // before that piece's TMEM load (keeping all 128 bytes live across the accumulator wait spilled registers)
      const bool has_mask = BWD && kCanRes && d.act >= TG_ACT_DRELU && inb;
      const uint4* mp = nullptr;
      if (BWD && has_mask)
        mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(d.mask) +
                                            (((size_t)tc.n * d.h + py) * d.w + px) * d.cout + tc.nb * p.bn);
      mbar_wait(bar_tfull + 8 * buf, bphase, 7);
      TG_ACC(te_tfull, t_s);
      t_s = TG_T0();
      tc_fence_after();
      if (d.epilogue == TG_EPI_NHWC_F16 || POOL) {
        // Each thread owns one output pixel = 64 channels = one contiguous 128-byte NHWC row: it
        // goes straight from registers to global memory (8 x 16-byte stores complete the line),
        // so the epilogue costs no shared-memory bandwidth -- the MMA operand reads need all of it.
        for (int acc = acc_lo; acc < acc_hi; ++acc) {
          int oy = py, ox = px, OW = d.w, OH = d.h;
          if (KIND == TG_CONVT_3X3_S2) { oy = 2 * py + (acc >> 1); ox = 2 * px + (acc & 1); OW = 2 * d.w; OH = 2 * d.h; }
          if (POOL) { OH = d.h >> 1; OW = d.w >> 1; oy = py >> 1; ox = px >> 1; }
          uint4* orow = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(d.y) +
                                                 (((size_t)tc.n * OH + oy) * OW + ox) * d.cout + tc.nb * p.bn);
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {                // bn == 64: two 32-column pieces
            uint32_t v[32];
            uint4 msk[BWD ? 4 : 1];
            if (BWD && has_mask) {
              ld_global_256(mp + pc * 4, msk[0], msk[BWD ? 1 : 0]);
              ld_global_256(mp + pc * 4 + 2, msk[BWD ? 2 : 0], msk[BWD ? 3 : 0]);
            }
            tmem_ld32(tmem_base + buf * acc_stride + acc * p.bn + pc * 32 + ((uint32_t)(q * 32) << 16), v);
            tmem_ld_wait();
            if (acc == acc_hi - 1 && pc == 1) {
              // all TMEM reads of this warp for this buffer are done -> hand it back to the MMA
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            }
            const float4* bias4 = reinterpret_cast<const float4*>(bias_s + tc.nb * p.bn + pc * 32);
            uint4 ov[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __half2* o = reinterpret_cast<__half2*>(&ov[i]);
              const __half2* rh = reinterpret_cast<const __half2*>(&res[pc * 4 + i]);
              const float4 b0 = bias4[i * 2], b1 = bias4[i * 2 + 1];
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int cidx = i * 8 + j * 2;
                float a0, a1;
                if (!BWD) {
                  a0 = tg_epi_val(__uint_as_float(v[cidx]), bb[j * 2], d.act);
                  a1 = tg_epi_val(__uint_as_float(v[cidx + 1]), bb[j * 2 + 1], d.act);
                  if (has_res) {
                    const float2 rf = __half22float2(rh[j]);
                    a0 += rf.x; a1 += rf.y;
                  }
                } else {
                  a0 = __uint_as_float(v[cidx]) + bb[j * 2];
                  a1 = __uint_as_float(v[cidx + 1]) + bb[j * 2 + 1];
                  if (has_res) {
                    const float2 rf = __half22float2(rh[j]);
                    a0 += rf.x; a1 += rf.y;
                  }
                  if (has_mask) {
                    const float2 mf = __half22float2(reinterpret_cast<const __half2*>(&msk[BWD ? i : 0])[j]);
                    a0 *= tg_dact(mf.x, d.act); a1 *= tg_dact(mf.y, d.act);
                  }
                }
                o[j] = __floats2half2_rn(a0, a1);
              }
            }
            if (POOL) {
              // max over the 2x2 block (all 32 lanes take part; floor pooling: blocks that reach outside the
              // image are not stored, so out-of-range pixels never contribute to a stored value)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint32_t* wv = reinterpret_cast<uint32_t*>(&ov[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, wv[j], 1);
                  __half2 mx = __hmax2(*reinterpret_cast<__half2*>(&wv[j]), *reinterpret_cast<__half2*>(&o1));
                  uint32_t m32 = *reinterpret_cast<uint32_t*>(&mx);
                  uint32_t o8 = __shfl_xor_sync(0xFFFFFFFFu, m32, 8);
                  mx = __hmax2(mx, *reinterpret_cast<__half2*>(&o8));
                  wv[j] = *reinterpret_cast<uint32_t*>(&mx);
                }
              }
              if (((tx | ty) & 1) == 0 && py + 1 < d.h && px + 1 < d.w) {
                st_global_256(orow + pc * 4, ov[0], ov[1]);
                st_global_256(orow + pc * 4 + 2, ov[2], ov[3]);
              }
            } else if (inb) {
              st_global_256(orow + pc * 4, ov[0], ov[1]);
              st_global_256(orow + pc * 4 + 2, ov[2], ov[3]);
            }
          }
        }
        TG_ACC(te_compute, t_s);
      } else {
        // MODE_TAPN heads: D[pos][tap*4+co] = x[pos] . W[tap][co]; out[p] = sum_taps D[p+off(tap)][tap].
        // Positions exchange their nine float4 tap products through shared memory.
        uint32_t v[48];
        const uint32_t tad = tmem_base + buf * acc_stride + ((uint32_t)(q * 32) << 16);
        tmem_ld32(tad, v);
        tmem_ld16(tad + 32, v + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        float4* E = reinterpret_cast<float4*>(sm + p.off_staging + (uint32_t)group * kTapnEBytes);
        if (!(p.dbg_flags & 2))
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
          E[r * 9 + tap] = make_float4(__uint_as_float(v[tap * 4]), __uint_as_float(v[tap * 4 + 1]),
                                       __uint_as_float(v[tap * 4 + 2]), __uint_as_float(v[tap * 4 + 3]));
        if (!(p.dbg_flags & 2)) named_bar_sync(1 + group, 128);
        if (inb && !(p.dbg_flags & 1)) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!(p.dbg_flags & 2))
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const float4 e = E[(r + (tap / 3 - 1) * TW + (tap % 3 - 1)) * 9 + tap];
            a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
          }
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            if (d.epilogue == TG_EPI_FLOW_NCHW_F32) tg_epi_flow(d, tc.n, py, px, d.h, d.w, ch, av[ch]);
            else tg_epi_out(d, tc.n, py, px, d.h, d.w, ch, av[ch]);
          }
        }
        // one exchange buffer per group (smem goes to TMA stages instead: the thin heads are bound
        // by bytes in flight): everyone must be done reading before the next tile overwrites it
        named_bar_sync(1 + group, 128);
        TG_ACC(te_compute, t_s);
      }
    }
    if (timing && gtid == 0 && group == 0) {
      unsigned long long* o = p.dbg + blockIdx.x * T_SLOTS;
      o[T_EPI_WAIT_STORE] = te_store_wait; o[T_EPI_WAIT_TFULL] = te_tfull; o[T_EPI_COMPUTE] = te_compute;
      o[T_EPI_STORE] = te_store; o[T_EPI_TOTAL] = clock64() - t_epi0;
    }
  }

  // ------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  if (timing && threadIdx.x == 0) p.dbg[blockIdx.x * T_SLOTS + T_KERNEL] = clock64() - t_kernel0;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// NHWC fp16 tensor [n][h][w][c] with explicit element strides for w/h/n (convT parity views)
int encode_nhwc(CUtensorMap* m, const void* ptr, int c, int w, int h, int n, size_t sw, size_t sh,
                size_t sn, int box_c, int box_w, int box_h) {
  EncodeTiledFn fn = get_encode_fn();
  TG_REQUIRE(fn != nullptr, TG_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TG_REQUIRE(r == CUDA_SUCCESS, TG_E_DRIVER,
             "cuTensorMapEncodeTiled failed (%d) c=%d w=%d h=%d n=%d box=%dx%dx%d", (int)r, c, w, h, n,
             box_c, box_w, box_h);
  return TG_OK;
}

}  // namespace

static unsigned long long* g_conv_timers = nullptr;
unsigned long long* tg_conv_timer_buffer() { return g_conv_timers; }   // shared with tg_chain_tcgen05.cu

extern "C" {

int tg_debug_set_conv_timers(void* device_buffer) {
  g_conv_timers = reinterpret_cast<unsigned long long*>(device_buffer);
  return TG_OK;
}

int tg_conv_validate(const tg_conv_desc* d, const char* who) {
  TG_REQUIRE(d != nullptr, TG_E_INVALID, "%s: null descriptor", who);
  TG_REQUIRE(d->x && d->weights && d->bias && d->y, TG_E_INVALID, "%s: null pointer", who);
  TG_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, TG_E_INVALID, "%s: bad size n=%d h=%d w=%d", who, d->n, d->h, d->w);
  TG_REQUIRE(d->kind == TG_CONV_3X3 || d->kind == TG_CONVT_3X3_S2 || d->kind == TG_CONV_3X3_S2, TG_E_INVALID,
             "%s: kind", who);
  TG_REQUIRE(d->act >= TG_ACT_NONE && d->act <= TG_ACT_DLRELU02, TG_E_INVALID, "%s: act", who);
  TG_REQUIRE((d->act >= TG_ACT_DRELU) == (d->mask != nullptr), TG_E_INVALID,
             "%s: mask must be given exactly for TG_ACT_DRELU / TG_ACT_DLRELU02", who);
  TG_REQUIRE(!(d->act >= TG_ACT_DRELU && (d->epilogue != TG_EPI_NHWC_F16 || d->kind == TG_CONVT_3X3_S2)),
             TG_E_UNSUPPORTED, "%s: derivative epilogues need NHWC output and a conv3x3 / conv3x3s2 layer", who);
  TG_REQUIRE(!(d->kind == TG_CONV_3X3_S2 && d->epilogue != TG_EPI_NHWC_F16), TG_E_UNSUPPORTED,
             "%s: conv3x3s2 needs the NHWC epilogue", who);
  TG_REQUIRE(d->cin == 64 || d->cin == 128 || d->cin == 256, TG_E_UNSUPPORTED,
             "%s: cin=%d (stored channels must be 64, 128 or 256)", who, d->cin);
  if (d->epilogue == TG_EPI_NHWC_F16_POOL2)
    TG_REQUIRE(d->kind == TG_CONV_3X3 && d->residual == nullptr && d->act <= TG_ACT_LRELU02 && d->h >= 2 && d->w >= 2,
               TG_E_UNSUPPORTED, "%s: the pooled epilogue needs a conv3x3 without residual / derivative epilogue", who);
  if (d->epilogue == TG_EPI_NHWC_F16 || d->epilogue == TG_EPI_NHWC_F16_POOL2) {
    TG_REQUIRE(d->cout == 64 || d->cout == 128 || d->cout == 256, TG_E_UNSUPPORTED,
               "%s: cout=%d (64, 128 or 256 for the NHWC epilogue)", who, d->cout);
    TG_REQUIRE(!(d->residual && d->kind == TG_CONVT_3X3_S2), TG_E_UNSUPPORTED, "%s: residual with convT", who);
    TG_REQUIRE(!(d->kind == TG_CONVT_3X3_S2 && d->cout != 64), TG_E_UNSUPPORTED,
               "%s: convT needs cout == 64 (4 parity accumulators in TMEM)", who);
  } else if (d->epilogue == TG_EPI_FLOW_NCHW_F32 || d->epilogue == TG_EPI_OUT_NCHW_F32) {
    TG_REQUIRE(d->kind == TG_CONV_3X3 && d->cout == TG_TAPN_ROWS && d->cout_real >= 1 && d->cout_real <= 4,
               TG_E_UNSUPPORTED, "%s: NCHW epilogues need conv3x3, cout=48 (tap-major N), cout_real<=4", who);
    TG_REQUIRE(d->residual == nullptr, TG_E_UNSUPPORTED, "%s: residual with NCHW epilogue", who);
  } else {
    TG_REQUIRE(false, TG_E_INVALID, "%s: epilogue %d", who, d->epilogue);
  }
  return TG_OK;
}

int tg_conv_tcgen05(const tg_conv_desc* d, void* stream) {
  int rc = tg_conv_validate(d, "conv_tcgen05");
  if (rc != TG_OK) return rc;
  TG_REQUIRE(d->a_mode >= TG_AMODE_AUTO && d->a_mode <= TG_AMODE_TAP, TG_E_INVALID, "conv_tcgen05: a_mode");
  TG_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->weights & 15) == 0 && ((uintptr_t)d->y & 15) == 0,
             TG_E_INVALID, "conv_tcgen05: pointers must be 16-byte aligned");

  KParams p;
  p.d = *d;
  p.dbg = g_conv_timers;
  const bool tapn = d->epilogue == TG_EPI_FLOW_NCHW_F32 || d->epilogue == TG_EPI_OUT_NCHW_F32;
  const bool pool = d->epilogue == TG_EPI_NHWC_F16_POOL2;
  p.step_y = tapn ? kTapnStepY : TH;
  p.step_x = tapn ? kTapnStepX : TW;
  p.tiles_x = tg_ceil_div(d->w, p.step_x);
  p.tiles_y = tg_ceil_div(d->h, p.step_y);
  p.num_tiles = p.tiles_x * p.tiles_y * d->n;
  p.chunks = d->cin / 64;
  TG_REQUIRE(d->cin_real >= 0 && d->cin_real <= d->cin, TG_E_INVALID, "conv_tcgen05: cin_real=%d outside [0, cin=%d]",
             d->cin_real, d->cin);
  p.ksteps = (p.chunks == 1 && d->cin_real > 0) ? (d->cin_real + 15) / 16 : 4;
  p.n_acc = d->kind == TG_CONVT_3X3_S2 ? 4 : 1;
  p.b_tile_bytes = (uint32_t)d->cout * 128u;

  // Output channels beyond 64 are split over CTAs (N = 64 per CTA): cout/64 x more CTAs on the
  // low-resolution FNet layers, and each CTA only needs its own 64-row slice of every weight tile.
  p.n_split = (!tapn && d->cout > 64) ? d->cout / 64 : 1;
  p.bn = d->cout / p.n_split;
  p.b_stage_bytes = (uint32_t)p.bn * 128u;
  const uint32_t b_total = (tapn ? 1u : 9u) * p.chunks * p.b_stage_bytes;   // resident slice per CTA
  // NHWC: 2 groups x 2-deep ring of 16 KB store staging; TAPN: 2 groups x 2 exchange buffers
  // TAPN: one exchange buffer per epilogue group
  uint32_t staging = tapn ? (uint32_t)epi_groups(TG_CONV_3X3, MODE_TAPN) * kTapnEBytes : 0u;
  const int hbox_w = d->kind != TG_CONVT_3X3_S2 ? TW + 2 : TW + 1;
  const int hbox_h = d->kind != TG_CONVT_3X3_S2 ? TH + 2 : TH + 1;
  const uint32_t halo_bytes = (uint32_t)hbox_w * hbox_h * 128u;
  const uint32_t halo_stage = (halo_bytes + 1023u) & ~1023u;
  uint32_t fixed = 1024u /*align slack*/ + kHeaderBytes + staging;

  const bool can_resident_halo = fixed + b_total + 2u * halo_stage <= kSmemLimit;
  const bool can_resident_tap = fixed + b_total + 2u * kTapABytes <= kSmemLimit;
  int mode = d->a_mode;
  if (tapn) mode = TG_AMODE_TAP;          // placeholder; thin heads always run MODE_TAPN below
  TG_REQUIRE(!(d->kind == TG_CONV_3X3_S2 && mode == TG_AMODE_HALO), TG_E_UNSUPPORTED,
             "conv_tcgen05: conv3x3s2 runs in tap mode (a stride-2 view is not a UMMA descriptor)");
  if (d->kind == TG_CONV_3X3_S2) mode = TG_AMODE_TAP;
  if (mode == TG_AMODE_AUTO) mode = can_resident_halo ? TG_AMODE_HALO : TG_AMODE_TAP;
  TG_REQUIRE(!(mode == TG_AMODE_HALO && !can_resident_halo), TG_E_UNSUPPORTED,
             "conv_tcgen05: halo mode needs the weights resident in smem (cin=%d cout=%d)", d->cin, d->cout);
  p.halo = mode == TG_AMODE_HALO;
  p.b_resident = p.halo ? 1 : (can_resident_tap ? 1 : 0);
  p.num_tiles *= p.n_split;
  p.idesc = (1u << 4) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  TG_REQUIRE(p.bn == 64 || (tapn && p.bn == TG_TAPN_ROWS), TG_E_UNSUPPORTED,
             "conv_tcgen05: per-CTA N must be 64 (48 for the thin heads)");
  TG_REQUIRE(!tapn || p.b_resident, TG_E_UNSUPPORTED, "conv_tcgen05: thin head weights must fit in smem");
  p.acc_stride = tapn ? 64u : (uint32_t)(p.n_acc * p.bn);
  p.n_buf = (int)(kTmemCols / p.acc_stride);
  if (p.n_buf > 8) p.n_buf = 8;
  p.n_buf &= ~1;                        // even: epilogue group g owns the buffers of parity g
  p.dbg_flags = 0;
  if (const char* e = getenv("TG_DBG_FLAGS")) p.dbg_flags = atoi(e);
  if (const char* e = getenv("TG_DBG_NBUF")) {   // diagnostics only (tools/conv_timers.py)
    const int v = atoi(e) & ~1;
    if (v >= 2 && v <= p.n_buf) p.n_buf = v;
  }
  if (tapn) {
    p.box_w = TW; p.box_h = TH; p.org_x = -1; p.org_y = -1;
    p.a_bytes = kTapABytes;
    p.stage_bytes = kTapABytes;
  } else if (p.halo) {
    p.box_w = hbox_w; p.box_h = hbox_h;
    p.org_x = d->kind == TG_CONV_3X3 ? -1 : 0;
    p.org_y = p.org_x;
    p.a_bytes = halo_bytes;
    p.stage_bytes = halo_stage;
  } else {
    p.box_w = TW; p.box_h = TH; p.org_x = 0; p.org_y = 0;
    p.a_bytes = kTapABytes;
    p.stage_bytes = kTapABytes + (p.b_resident ? 0u : p.b_stage_bytes);
  }
  const uint32_t avail = kSmemLimit - fixed - (p.b_resident ? b_total : 0u);
  int stages = (int)(avail / p.stage_bytes);
  const int want = p.halo ? 6 : kMaxStages;
  if (stages > want) stages = want;
  if (p.halo && p.chunks == 1 && stages >= 2) stages &= ~1;   // even: each of the two MMA issuers owns its stages
  TG_REQUIRE(stages >= 2, TG_E_UNSUPPORTED, "conv_tcgen05: shared memory budget (cin=%d cout=%d)", d->cin, d->cout);
  p.n_stages = stages;
  p.off_b = kHeaderBytes;
  p.off_stage = kHeaderBytes + (p.b_resident ? b_total : 0u);
  p.off_staging = p.off_stage + (uint32_t)stages * p.stage_bytes;
  const uint32_t smem_bytes = 1024u + p.off_staging + staging;
  TG_REQUIRE(smem_bytes <= kSmemLimit, TG_E_UNSUPPORTED, "conv_tcgen05: smem %u > limit", smem_bytes);

  // tensor maps
  TgMaps map_a;
  if (d->kind != TG_CONV_3X3_S2) {
    rc = encode_nhwc(&map_a.m[0], d->x, d->cin, d->w, d->h, d->n, (size_t)d->cin, (size_t)d->w * d->cin,
                     (size_t)d->h * d->w * d->cin, 64, p.box_w, p.box_h);
    if (rc != TG_OK) return rc;
    map_a.m[1] = map_a.m[2] = map_a.m[3] = map_a.m[0];
  } else {
    // x [n,2h,2w,cin]: parity plane (py,px) = pixels (2i+py, 2j+px), each an [n,h,w,cin] strided view
    const size_t W2 = (size_t)2 * d->w, C = (size_t)d->cin;
    for (int pl = 0; pl < 4; ++pl) {
      const __half* base = reinterpret_cast<const __half*>(d->x) + ((size_t)(pl >> 1) * W2 + (pl & 1)) * C;
      rc = encode_nhwc(&map_a.m[pl], base, d->cin, d->w, d->h, d->n, 2 * C, 2 * W2 * C,
                       (size_t)4 * d->h * d->w * C, 64, p.box_w, p.box_h);
      if (rc != TG_OK) return rc;
    }
  }

  static TgPerDeviceOnce attr_once;
  const cudaError_t attr_err = attr_once.run([] {
    cudaError_t e, err = cudaSuccess;
#define TG_SET_ATTR(K, H)                                                                                 \
    e = cudaFuncSetAttribute(conv_tcgen05_kernel<K, H, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)kSmemLimit);                                                            \
    if (e != cudaSuccess) err = e;                                                                        \
    e = cudaFuncSetAttribute(conv_tcgen05_kernel<K, H, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)kSmemLimit);                                                            \
    if (e != cudaSuccess) err = e;
    TG_SET_ATTR(TG_CONV_3X3, MODE_HALO) TG_SET_ATTR(TG_CONV_3X3, MODE_TAP) TG_SET_ATTR(TG_CONV_3X3, MODE_TAPN)
    TG_SET_ATTR(TG_CONVT_3X3_S2, MODE_HALO) TG_SET_ATTR(TG_CONVT_3X3_S2, MODE_TAP)
#undef TG_SET_ATTR
#define TG_SET_ATTR_POOL(H)                                                                                                  \
    e = cudaFuncSetAttribute(conv_tcgen05_kernel<TG_CONV_3X3, H, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)kSmemLimit);                                                                               \
    if (e != cudaSuccess) err = e;
    TG_SET_ATTR_POOL(MODE_HALO) TG_SET_ATTR_POOL(MODE_TAP)
#undef TG_SET_ATTR_POOL
#define TG_SET_ATTR_BWD(K, H)                                                                                  \
    e = cudaFuncSetAttribute(conv_tcgen05_kernel<K, H, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)kSmemLimit);                                                                 \
    if (e != cudaSuccess) err = e;
    TG_SET_ATTR_BWD(TG_CONV_3X3, MODE_HALO) TG_SET_ATTR_BWD(TG_CONV_3X3, MODE_TAP) TG_SET_ATTR_BWD(TG_CONV_3X3_S2, MODE_TAP)
#undef TG_SET_ATTR_BWD
    return err;
  });
  TG_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "conv_tcgen05: cudaFuncSetAttribute: %s",
             cudaGetErrorString(attr_err));

  int sms = 0;
  rc = tg_device_sm_count(&sms);
  if (rc != TG_OK) return rc;
  int grid = d->max_ctas > 0 ? d->max_ctas : sms;
  if (grid > p.num_tiles) grid = p.num_tiles;
  grid -= grid % p.n_split;             // every CTA keeps one fixed N slice (resident weights)
  if (grid < p.n_split) grid = p.n_split;
  // always request the full carve-out: exactly one CTA per SM, so the 512-column TMEM
  // allocation can never contend
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t lerr = cudaSuccess;
  const bool bwd = d->act >= TG_ACT_DRELU || d->kind == TG_CONV_3X3_S2;
  if (pool) {
    if (p.halo)
      lerr = tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_HALO, false, false, true>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_HALO)), kSmemLimit, st, map_a, p);
    else
      lerr = tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_TAP, false, false, true>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_TAP)), kSmemLimit, st, map_a, p);
  } else if (bwd) {
    if (d->kind == TG_CONV_3X3_S2)
      lerr = tg_launch(conv_tcgen05_kernel<TG_CONV_3X3_S2, MODE_TAP, false, true>, dim3(grid), dim3(conv_threads(TG_CONV_3X3_S2, MODE_TAP)), kSmemLimit, st, map_a, p);
    else if (p.halo)
      lerr = tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_HALO, false, true>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_HALO)), kSmemLimit, st, map_a, p);
    else
      lerr = tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_TAP, false, true>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_TAP)), kSmemLimit, st, map_a, p);
  } else if (tapn) {
    lerr = p.dbg ? tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_TAPN, true>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_TAPN)), kSmemLimit, st, map_a, p)
                 : tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_TAPN, false>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_TAPN)), kSmemLimit, st, map_a, p);
  } else if (d->kind == TG_CONV_3X3) {
    if (p.halo) lerr = p.dbg ? tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_HALO, true>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_HALO)), kSmemLimit, st, map_a, p)
                 : tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_HALO, false>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_HALO)), kSmemLimit, st, map_a, p);
    else        lerr = p.dbg ? tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_TAP, true>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_TAP)), kSmemLimit, st, map_a, p)
                 : tg_launch(conv_tcgen05_kernel<TG_CONV_3X3, MODE_TAP, false>, dim3(grid), dim3(conv_threads(TG_CONV_3X3, MODE_TAP)), kSmemLimit, st, map_a, p);
  } else {
    if (p.halo) lerr = p.dbg ? tg_launch(conv_tcgen05_kernel<TG_CONVT_3X3_S2, MODE_HALO, true>, dim3(grid), dim3(conv_threads(TG_CONVT_3X3_S2, MODE_HALO)), kSmemLimit, st, map_a, p)
                 : tg_launch(conv_tcgen05_kernel<TG_CONVT_3X3_S2, MODE_HALO, false>, dim3(grid), dim3(conv_threads(TG_CONVT_3X3_S2, MODE_HALO)), kSmemLimit, st, map_a, p);
    else        lerr = p.dbg ? tg_launch(conv_tcgen05_kernel<TG_CONVT_3X3_S2, MODE_TAP, true>, dim3(grid), dim3(conv_threads(TG_CONVT_3X3_S2, MODE_TAP)), kSmemLimit, st, map_a, p)
                 : tg_launch(conv_tcgen05_kernel<TG_CONVT_3X3_S2, MODE_TAP, false>, dim3(grid), dim3(conv_threads(TG_CONVT_3X3_S2, MODE_TAP)), kSmemLimit, st, map_a, p);
  }
  TG_REQUIRE(lerr == cudaSuccess, (int)lerr, "conv_tcgen05: launch failed: %s", cudaGetErrorString(lerr));
  TG_CUDA_LAUNCH_CHECK("conv_tcgen05");
  return TG_OK;
}

}  // extern "C"
