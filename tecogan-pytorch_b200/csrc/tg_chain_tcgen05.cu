// A chain of 64->64 3x3 convolutions (SRNet conv_in + residual blocks) as ONE persistent,
// warp-specialised tcgen05 launch for sm_100a -- the layer loop lives inside the kernel.
//
//   tiles   : 16x8 output pixels (M = 128), static assignment tile = cta + k*grid for every layer
//   A       : one 18x10 halo box per tile by TMA (zero fill = conv padding), 3 stages; the nine
//             taps are nine shifted UMMA-descriptor views of the box (see tg_conv_tcgen05.cu)
//   B       : packed weights double-buffered in smem (layer l in buffer l & 1, 72 KB each): layer l+1
//             streams in (one bulk copy) as soon as both MMA issuers have retired layer l-1 -- a whole
//             layer ahead of its first use, so there is no reload bubble at layer boundaries
//   D       : one 64-column fp32 accumulator per tile, 8 tiles in flight in TMEM (splitting the taps
//             over two partial accumulators was measured: no gain).
//   layers  : a tile of layer l needs the tiles of layer l-1 under its halo.  Epilogue groups
//             publish `flags[tile] = epoch*32 + l + 1` (release, gpu scope) after their stores; a
//             checker warp polls the <= 9 flags of each of the next three tiles (sliding window, one
//             load per lane and round) ahead of the TMA producer and hands it an in-order "verified"
//             counter.  No launch,
//             pipeline fill/drain, weight reload bubble or grid barrier between layers.
//   roles   : warp 0 TMA producer, warps 1 and 12 MMA issuers (even / odd tiles of the CTA's sequence),
//             warp 2 TMEM allocator + weight streamer, warp 3 dependency checker, warps 4..11 two
//             epilogue groups alternating tiles.
//             TWO issuers because an N=64 MMA occupies the tensor pipe for 48 cycles and the pipe hides
//             only ~180 cycles without a new instruction: one warp needs ~80 cycles per MMA for the issue
//             plus its bookkeeping (barrier polls, commits, descriptor arithmetic), so with one issuer
//             the pipe idles 40 % of the time; two warps' instruction streams interleave in the pipe and
//             it stays fed (tools/mma_probe.cu: 97 -> 52 cycles per MMA with 40 cycles of other work
//             per MMA and issuer).
//
// Replaces the 21 nn.Conv2d launches of SRNet.conv_in / ResidualBlock (tecogan_nets.py:92-100,
// 111-116, 139-141) per step.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "tg_common.cuh"
#include "tg_epilogue.cuh"
#include "tg_tcgen05.cuh"

namespace {

constexpr int TH = 16, TW = 8, BOXW = TW + 2, BOXH = TH + 2;
constexpr int kThreads = 416;                           // 13 warps, see `roles`
constexpr int kIssuerBWarp = 12;
constexpr int kMaxMaps = 4;
constexpr int kStages = 3;
constexpr int kSlots = 18;                              // two weight buffers of nine taps: layer l in buffer l & 1
constexpr int kBufs = 8;                               // tiles in flight in TMEM (8 x 64 columns)
constexpr uint32_t kAccStride = 64;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTapWBytes = 64 * 128;              // one tap: [64 cout rows][64 cin] fp16
constexpr uint32_t kHaloBytes = BOXW * BOXH * 128;     // 23,040
constexpr uint32_t kStageBytes = (kHaloBytes + 1023u) & ~1023u;
constexpr uint32_t kOffBias = 1024;                    // [layers][64] fp32
constexpr uint32_t kOffW = kOffBias + TG_CHAIN_MAX_LAYERS * 256;
constexpr uint32_t kOffStage = kOffW + kSlots * kTapWBytes;
constexpr uint32_t kSmemBytes = 1024 /*align slack*/ + kOffStage + kStages * kStageBytes;
static_assert(kOffW % 1024 == 0 && kOffStage % 1024 == 0, "swizzle atoms need 1024B alignment");
static_assert(kSmemBytes <= 232448, "227 KB opt-in shared memory limit");
constexpr int kSyncFlags = 32;                         // uint32 index of the first tile flag

struct ChainLayerDev {
  const unsigned char* w;
  const float* bias;
  const __half* res;
  __half* y;
  int map, act;
};
struct ChainParams {
  CUtensorMap maps[kMaxMaps];
  ChainLayerDev layers[TG_CHAIN_MAX_LAYERS];
  uint32_t* sync;                  // [0] epoch, [1] finished CTAs, [kSyncFlags + tile] progress
  int n_layers, n, h, w, tiles_x, tiles_y, num_tiles;
  uint32_t idesc;
  unsigned long long* dbg;         // optional per-CTA timers (same slots as tg_conv_tcgen05)
  uint32_t xflags;                 // TIMING builds only: ablation switches (TG_CHAIN_ABLATE), results are then WRONG
};

__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t saddr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
// non-blocking phase test (mbarrier.try_wait may suspend the thread for a while)
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// named barrier over `nthreads` threads + AND-reduction of a predicate
__device__ __forceinline__ uint32_t named_bar_red_and(int id, int nthreads, uint32_t pred) {
  uint32_t out;
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 q, %3, 0;\nbar.red.and.pred p, %1, %2, q;\nselp.b32 %0, 1, 0, p;\n}\n"
      : "=r"(out)
      : "r"(id), "r"(nthreads), "r"(pred)
      : "memory");
  return out;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// generic-proxy global writes <-> async-proxy (TMA) global reads
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
// 256-bit load served by L2 (strong, gpu scope): the data was written by another SM during this
// kernel, so neither the L1 nor the non-coherent path may be used
__device__ __forceinline__ void ld_global_256_l2(const void* ptr, uint4& a, uint4& b) {
  asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(ptr)
               : "memory");
}

enum { CT_PROD_FLAGS = 0, CT_PROD_EMPTY, CT_MMA_TOTAL, CT_MMA_WAIT, CT_EPI_TFULL, CT_EPI_TOTAL, CT_KERNEL, CT_TILES,
       CT_MMA_ISSUE0, CT_MMA_LOOK, CT_MMA_ISSUE1, CT_MMA_BOUNDARY, CT_CHK_ITERS, CT_CHK_FENCE, CT_CHK_HITS, CT_CHK_TOTAL,
       CT_SLOTS = 16 };

// event trace of CTA 0 (TIMING builds): trace[seq * 8 + event] = clock64, behind the 148 x 16 timer slots
#define CT_TRACE(seq, ev) do { if (timing && b == 0) cp.dbg[148 * CT_SLOTS + (size_t)(seq) * 8 + (ev)] = clock64(); } while (0)

template <bool TIMING>
__global__ void __launch_bounds__(kThreads, 1)
conv_chain_kernel(const __grid_constant__ ChainParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  constexpr bool timing = TIMING;
  const long long t_kernel0 = timing ? clock64() : 0;
  // ablation switches of the TIMING build (tools/conv_timers.py chain-ablate): 1 no dependency waits, 2 no epilogue
  // global loads/stores, 4 no flag publication, 8 no TMA loads, 16 one MMA per tile, 32 no bias reads
  const uint32_t xf = timing ? cp.xflags : 0u;

  // warp index through a shuffle broadcast: the compiler then KNOWS it is warp-uniform and keeps the role
  // loops' counters, barrier addresses and UMMA descriptors in uniform registers (a plain threadIdx.x >> 5
  // leaves them in vector registers and pays an R2UR per descriptor word in front of every MMA)
  const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t bar_full = base;                 // [2 issuers][kStages]: per ISSUER -- with an odd stage count an issuer meets
                                                  // a stage only every other time it is filled, and a parity wait must see
                                                  // every completion of its barrier (tests/test_chain_protocol_model.py, bug 4)
  const uint32_t bar_empty = base + 64;           // [kStages <= 8]
  const uint32_t bar_tfull = base + 128;          // [kBufs]
  const uint32_t bar_tempty = base + 192;         // [kBufs]
  const uint32_t bar_wfull = base + 256;          // [2] the weights of a layer have landed in buffer l & 1
  const uint32_t bar_wfree = base + 272;          // [2] both issuers' MMAs of the layer in buffer l & 1 have retired
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + 432);
  volatile uint32_t* epoch_s = reinterpret_cast<volatile uint32_t*>(sm + 436);
  const uint32_t deps_ok_addr = base + 440;       // tiles (in this CTA's sequence) whose dependencies are verified
  float* bias_s = reinterpret_cast<float*>(sm + kOffBias);
  const uint32_t smem_w0 = base + kOffW;
  const uint32_t smem_stage0 = base + kOffStage;

  const int G = (int)gridDim.x, b = (int)blockIdx.x;
  const int L = cp.n_layers;
  const int n_my = (cp.num_tiles - 1 - b) / G + 1;      // grid <= num_tiles (host)
  const int total = L * n_my;
  const int per_img = cp.tiles_x * cp.tiles_y;
  uint32_t* flags = cp.sync + kSyncFlags;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kMaxMaps; ++i) tma_prefetch_desc(&cp.maps[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2 * kStages; ++s) mbar_init(bar_full + 8 * s, 1);
    for (int i = 0; i < kBufs; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_wfull + 8 * i, 1); mbar_init(bar_wfree + 8 * i, 2); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_s)), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // Weights are static across the step (packed at module init / refresh, never by the predecessor
  // kernel): the first two layers' are requested before joining the PDL wait.
  if (warp == 2 && lane == 0) {
    for (int l = 0; l < 2 && l < L; ++l) {
      mbar_expect_tx(bar_wfull + 8 * l, 9 * kTapWBytes);
      bulk_load(smem_w0 + (uint32_t)l * 9u * kTapWBytes, cp.layers[l].w, 9 * kTapWBytes, bar_wfull + 8 * l);
    }
  }
  tg_pdl_wait();
  tg_pdl_trigger();
  for (int i = threadIdx.x; i < L * 64; i += kThreads) bias_s[i] = cp.layers[i >> 6].bias[i & 63];
  if (threadIdx.x == 0) {
    *epoch_s = ld_acquire_gpu(cp.sync);
    *reinterpret_cast<volatile uint32_t*>(sm + 440) = (uint32_t)n_my;   // layer 0 has no dependencies
  }
  __syncthreads();
  const uint32_t fbase = (*epoch_s) << 5;              // flags of this launch: fbase + layers done

  if (warp == 0) {
    // ============================================================ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    long long tw_flags = 0, tw_empty = 0;
    for (int l = 0; l < L; ++l) {
      const CUtensorMap* map = &cp.maps[cp.layers[l].map];
      for (int k = 0; k < n_my; ++k) {
        const int tile = b + k * G;
        const int n = tile / per_img;
        const int r = tile - n * per_img;
        const int ty = r / cp.tiles_x, tx = r - ty * cp.tiles_x;
        if (l > 0 && !(xf & 1u)) {
          // the dependency checker (warp 3) has verified the halo tiles of every tile up to *deps_ok
          const long long t0 = timing ? clock64() : 0;
          const uint32_t seq = (uint32_t)(l * n_my + k);
          if (lane == 0 && ld_acquire_cta_shared(deps_ok_addr) <= seq) {
            const long long s0 = clock64();
            while (ld_acquire_cta_shared(deps_ok_addr) <= seq) {
              __nanosleep(20);                          // do not hammer shared memory: the MMA operands need it
              if (clock64() - s0 > 3000000000LL) {
                if (b < 2) printf("tg_conv_chain: dependency timeout block=%d layer=%d tile=%d\n", b, l, tile);
                __trap();
              }
            }
          }
          __syncwarp();
          if (timing) tw_flags += clock64() - t0;
        }
        if (lane == 0) {
          if (l > 0) fence_proxy_async_global();
          const long long t0 = timing ? clock64() : 0;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1, 1);
          if (timing) tw_empty += clock64() - t0;
          const uint32_t fb = bar_full + 8 * (((l * n_my + k) & 1) * kStages + stage);   // the consuming issuer's barrier
          if (xf & 8u) {
            mbar_expect_tx(fb, 0);
          } else {
            mbar_expect_tx(fb, kHaloBytes);
            tma_load_4d(smem_stage0 + stage * kStageBytes, map, fb, 0, tx * TW - 1, ty * TH - 1, n);
          }
          CT_TRACE(l * n_my + k, 5);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
    if (timing && lane == 0) {
      cp.dbg[b * CT_SLOTS + CT_PROD_FLAGS] = tw_flags;
      cp.dbg[b * CT_SLOTS + CT_PROD_EMPTY] = tw_empty;
    }
  } else if (warp == 1 || warp == kIssuerBWarp) {
    // ============================================================ MMA issuers
    // Issuer w (0: warp 1, 1: warp 12) owns the tiles g = w, w + 2, ... of the CTA's sequence; A stage
    // g % kStages, accumulator g % kBufs.  All 32 lanes run the loop on warp-uniform values (uniform
    // datapath: one add per descriptor), one elected lane issues.  Per layer and issuer: ENTER = wait for
    // the layer's weights, EXIT = commit behind the issuer's last tile of the layer onto wfree[l & 1]
    // (count 2: the buffer is refilled when BOTH issuers' MMAs of the layer have retired).  An issuer
    // with no tile in a layer (one tile per CTA and layer) enters and exits it all the same, in order,
    // so every parity wait below stays at most one phase behind its barrier.
    const int w = warp == 1 ? 0 : 1;
    const uint32_t a_hi32 = (uint32_t)(make_sdesc(0, (uint32_t)BOXW * 128u) >> 32);   // 8-row groups = consecutive tile rows
    const uint32_t b_hi32 = (uint32_t)(make_sdesc(0, 1024u) >> 32);
    // descriptor low words carry the LBO field (1 << 16) beside the address: one add per operand
    const uint32_t a_lo0 = ((smem_stage0 & 0x3FFFFu) >> 4) + 0x10000u;
    const uint32_t b_lo0 = ((smem_w0 & 0x3FFFFu) >> 4) + 0x10000u;
    const bool tm = timing && w == 0;
    const long long t_mma0 = tm ? clock64() : 0;
    long long tw = 0, t_i0 = 0, t_lk = 0, t_i1 = 0, t_bd = 0;
    int g = w, l = 0, k = w, entered = -1;
    while (k >= n_my) { k -= n_my; ++l; }
    // A stage g % kStages; its 'landed' barrier full[w][stage] completes once per 2 * kStages tiles (kStages is odd)
    static_assert(kStages % 2 == 1, "phase of full[w][stage] below assumes an odd stage count");
    const uint32_t my_full = bar_full + 8 * (w * kStages);
    int stage = w % kStages, buf = w, fuse = 0;     // fuse: g / (2 kStages) = fills of (w, stage) so far, advanced with g
    int fcnt = w;                                   // g % (2 kStages)
    uint32_t bphase = 0;
    if (g < total) {
      mbar_wait(bar_tempty + 8 * buf, 1, 4);     // fresh barrier: passes immediately
      mbar_wait(my_full + 8 * stage, 0, 5);
      tc_fence_after();
    }
    for (; g < total; g += 2) {
      // position of this issuer's next tile
      int nk = k + 2, nl = l;
      while (nk >= n_my) { nk -= n_my; ++nl; }
      const bool has_next = g + 2 < total;
      const bool last_mine = nl != l;                  // my last tile of layer l
      int nstage = stage + 2;
      if (nstage >= kStages) nstage -= kStages;
      int nfcnt = fcnt + 2, nfuse = fuse;
      if (nfcnt >= 2 * kStages) { nfcnt -= 2 * kStages; ++nfuse; }
      const uint32_t nphase = (uint32_t)nfuse & 1u;
      int nbuf = buf + 2; uint32_t nbphase = bphase;
      if (nbuf >= kBufs) { nbuf -= kBufs; nbphase ^= 1u; }
      const uint32_t sa = a_lo0 + (uint32_t)stage * (kStageBytes >> 4);
      const uint32_t sb = b_lo0 + (uint32_t)(l & 1) * (9u * (kTapWBytes >> 4));
      const uint32_t d0 = tmem_base + (uint32_t)buf * kAccStride;
      bool next_ready = false;
      if (lane == 0) CT_TRACE(g, 6);
      if (entered < l) {
        const long long tb0 = tm ? clock64() : 0;
        while (entered < l) {
          ++entered;
          mbar_wait(bar_wfull + 8 * (entered & 1), (uint32_t)(entered >> 1) & 1u, 3);
          if (entered < l) {                           // a layer without a tile of mine: nothing to retire
            if (elect_one_sync()) mbar_arrive(bar_wfree + 8 * (entered & 1));
            __syncwarp();
          }
        }
        tc_fence_after();
        if (tm) t_bd += clock64() - tb0;
      }
      // The 36 MMAs are issued in two parts (taps 0-3 | taps 4-8) with the look-ahead poll between them.
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const long long tp0 = tm ? clock64() : 0;
        if (elect_one_sync()) {
#pragma unroll
          for (int i = (part == 0 ? 0 : 16); i < (part == 0 ? 16 : 36); ++i) {
            const int tap = i >> 2, kk = i & 3;
            if ((xf & 16u) && i > 0) continue;
            const uint32_t off = (uint32_t)((tap / 3) * BOXW + tap % 3) * 8u;   // (dy+1, dx+1) pixels, 128 B each
            umma_f16_words(d0, sa + off + 2u * kk, a_hi32, sb + (uint32_t)tap * (kTapWBytes >> 4) + 2u * kk, b_hi32,
                           cp.idesc, i >= 1 ? 1u : 0u);
          }
          if (part == 1) {
            umma_commit(bar_empty + 8 * stage);
            umma_commit(bar_tfull + 8 * buf);
            if (last_mine) umma_commit(bar_wfree + 8 * (l & 1));
          }
        }
        __syncwarp();
        if (tm) { if (part == 0) t_i0 += clock64() - tp0; else t_i1 += clock64() - tp0; }
        const long long tl0 = tm ? clock64() : 0;
        if (part == 0 && has_next) {
          // look ahead while MMAs are queued in the tensor pipe -- but never BLOCK before this tile
          // is committed: the next tile may (transitively) depend on this one through the flags
          // (test_wait: try_wait may suspend the warp for hundreds of cycles)
          uint32_t r = mbar_test_wait(bar_tempty + 8 * nbuf, nbphase ^ 1);
          r &= mbar_test_wait(my_full + 8 * nstage, nphase);
          next_ready = __all_sync(0xFFFFFFFFu, r != 0);
          if (next_ready) tc_fence_after();
          if (tm) t_lk += clock64() - tl0;
        }
      }
      if (has_next && !next_ready) {
        const long long t0 = tm ? clock64() : 0;
        mbar_wait(bar_tempty + 8 * nbuf, nbphase ^ 1, 4);
        mbar_wait(my_full + 8 * nstage, nphase, 5);
        if (tm) tw += clock64() - t0;
        tc_fence_after();
      }
      stage = nstage; fcnt = nfcnt; fuse = nfuse; buf = nbuf; bphase = nbphase; k = nk; l = nl;
    }
    if (tm && lane == 0) {
      cp.dbg[b * CT_SLOTS + CT_MMA_TOTAL] = clock64() - t_mma0;
      cp.dbg[b * CT_SLOTS + CT_MMA_WAIT] = tw;
      cp.dbg[b * CT_SLOTS + CT_TILES] = total;
      cp.dbg[b * CT_SLOTS + CT_MMA_ISSUE0] = t_i0;
      cp.dbg[b * CT_SLOTS + CT_MMA_LOOK] = t_lk;
      cp.dbg[b * CT_SLOTS + CT_MMA_ISSUE1] = t_i1;
      cp.dbg[b * CT_SLOTS + CT_MMA_BOUNDARY] = t_bd;
    }
  } else if (warp == 2) {
    // ============================================================ weight streamer
    // Layer l lives in buffer l & 1.  It may be (re)filled once both issuers have retired layer l - 2
    // (wfree[l & 1], one completion per use of the buffer): a whole layer before its first MMA.
    if (lane == 0) {
      for (int l = 2; l < L; ++l) {
        mbar_wait(bar_wfree + 8 * (l & 1), (uint32_t)((l - 2) >> 1) & 1u, 8);
        mbar_expect_tx(bar_wfull + 8 * (l & 1), 9 * kTapWBytes);
        bulk_load(smem_w0 + (uint32_t)(l & 1) * 9u * kTapWBytes, cp.layers[l].w, 9 * kTapWBytes, bar_wfull + 8 * (l & 1));
      }
    }
  } else if (warp == 3) {
    // ============================================================ dependency checker
    // Runs ahead of the TMA producer over a sliding window of the next THREE tiles of this CTA's sequence:
    // tile q is watched by the nine lanes 9 * ((q - n_my) % 3) .. + 8, one lane per tile under its halo, ONE
    // relaxed gpu-scope load per lane and round (strong loads of one thread do not overlap: with three per lane
    // a round took ~2000 cycles, all of it latency in the dependency loop between CTAs), one acquire fence per
    // publication.  The in-order prefix of verified tiles is published through a shared-memory counter and the
    // window slides on.
    constexpr int W = 3;
    const int sub = lane / 9, nbr = lane - sub * 9;
    int q = n_my + (sub < W ? sub : 0);
    const uint32_t* f = nullptr;
    uint32_t need = 0;
    bool ok = true;
    auto setup = [&]() {
      f = nullptr;
      need = 0;
      if (sub < W && q < total) {
        const int l = q / n_my, k = q - l * n_my;
        const int tile = b + k * G;
        const int n = tile / per_img;
        const int r = tile - n * per_img;
        const int ty = r / cp.tiles_x, tx = r - ty * cp.tiles_x;
        const int yy = ty + nbr / 3 - 1, xx = tx + nbr % 3 - 1;
        if (yy >= 0 && yy < cp.tiles_y && xx >= 0 && xx < cp.tiles_x) {
          f = flags + (size_t)n * per_img + yy * cp.tiles_x + xx;
          need = fbase + (uint32_t)l;                   // flag >= need  <=>  layer l-1 of that tile is published
        }
      }
      ok = f == nullptr;
    };
    setup();
    int head = n_my;                                    // first tile of the sequence not yet verified
    long long t_s = clock64();
    long long c_iters = 0, c_fence = 0, c_hits = 0;
    const long long c_t0 = timing ? clock64() : 0;
    while (head < total && !(xf & 1u)) {
      if (timing) ++c_iters;
      if (!ok) ok = (int)(ld_relaxed_gpu(f) - need) >= 0;
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
      int p = 0;
      while (p < W && head + p < total) {
        const int grp = (head + p - n_my) % W;
        if (((m >> (9 * grp)) & 0x1FFu) != 0x1FFu) break;
        ++p;
      }
      if (p > 0) {
        const long long tf0 = timing ? clock64() : 0;
        if (!(xf & 64u)) fence_acq_rel_gpu();           // relaxed polls + fence = acquire (every lane)
        if (timing) { c_fence += clock64() - tf0; c_hits += p; }
        __syncwarp();                                   // the other lanes' acquires happen-before the publication
        if (lane == 0) {
          st_release_cta_shared(deps_ok_addr, (uint32_t)(head + p));
          for (int i = 0; i < p; ++i) CT_TRACE(head + i, 4);
        }
        head += p;
        if (sub < W) {
          bool moved = false;
          while (q < head) { q += W; moved = true; }
          if (moved) setup();
        }
        t_s = clock64();
      } else if (clock64() - t_s > 3000000000LL) {
        if (b < 2 && lane == 0) printf("tg_conv_chain: flag timeout block=%d seq=%d of %d\n", b, head, total);
        __trap();
      }
    }
    if (timing && lane == 0) {
      cp.dbg[b * CT_SLOTS + CT_CHK_ITERS] = c_iters;
      cp.dbg[b * CT_SLOTS + CT_CHK_FENCE] = c_fence;
      cp.dbg[b * CT_SLOTS + CT_CHK_HITS] = c_hits;
      cp.dbg[b * CT_SLOTS + CT_CHK_TOTAL] = clock64() - c_t0;
    }
  } else if (warp >= 4 && warp < 12) {
    // ============================================================ epilogue (2 groups alternate tiles)
    const int grp = (warp - 4) >> 2;
    const int gtid = threadIdx.x - 128 - grp * 128;
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // tile row = TMEM lane
    const int ry = r >> 3, rx = r & 7;
    const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
    long long te_tfull = 0;
    const long long t_epi0 = timing ? clock64() : 0;
    // Publication of a finished tile (flag = layers done, release at gpu scope) is DEFERRED: the
    // gpu-scope fence has to wait until the tile's stores have reached L2, and done right after the
    // stores it stalls the group for that round trip on every tile (measured: the epilogue became the
    // bottleneck).  The group barrier at the top of the next tile orders the 128 threads' stores
    // before thread 0 (CTA scope; its fence + release store is cumulative over them -- the
    // grid-barrier pattern); thread 0 then publishes right before the NEXT tile's stores, when the
    // old ones have long landed.  If the next accumulator is not ready yet the group would only
    // wait, so it publishes at once -- a tile is never held back while its CTA is idle, which also
    // keeps the dependency graph free of cycles.  The consumers order the async proxy (TMA)
    // behind their acquire with fence.proxy.async.
    bool pending = false;
    uint32_t* pend_flag = nullptr;
    uint32_t pend_val = 0;
    int pend_seq = 0;
    for (int g = grp; g < total; g += 2) {
      const int l = g / n_my, k = g - l * n_my;
      const int buf = g & (kBufs - 1);
      const uint32_t bphase = (uint32_t)(g / kBufs) & 1u;
      const int tile = b + k * G;
      const int n = tile / per_img;
      const int rr = tile - n * per_img;
      const int ty = rr / cp.tiles_x, tx = rr - ty * cp.tiles_x;
      const int py = ty * TH + ry, px = tx * TW + rx;
      const bool inb = py < cp.h && px < cp.w;
      const ChainLayerDev& ly = cp.layers[l];
      const size_t pix = ((size_t)n * cp.h + py) * cp.w + px;
      const float slope = tg_act_slope(ly.act);
      uint4 res[8];
      const bool has_res = ly.res != nullptr && inb && !(xf & 2u);
      if (has_res) {
        const uint4* rp = reinterpret_cast<const uint4*>(ly.res + pix * 64);
#pragma unroll
        for (int i = 0; i < 8; i += 2) ld_global_256_l2(rp + i, res[i], res[i + 1]);
      }
      const long long t0 = timing ? clock64() : 0;
      const uint32_t acc_ready = named_bar_red_and(1 + grp, 128, mbar_test_wait(bar_tfull + 8 * buf, bphase));
      if (!acc_ready) {
        if (pending) {
          if (gtid == 0) { fence_acq_rel_gpu(); st_relaxed_gpu(pend_flag, pend_val); CT_TRACE(pend_seq, 3); }
          pending = false;
        }
        mbar_wait(bar_tfull + 8 * buf, bphase, 7);
      }
      if (timing) te_tfull += clock64() - t0;
      if (gtid == 0) CT_TRACE(g, 1);
      tc_fence_after();
      uint4* orow = reinterpret_cast<uint4*>(ly.y + pix * 64);
      const uint32_t tad = tmem_base + (uint32_t)buf * kAccStride + lane_bits;
#pragma unroll
      for (int pc = 0; pc < 2; ++pc) {
        uint32_t v[32];
        tmem_ld32(tad + pc * 32, v);
        tmem_ld_wait();
        if (pc == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        const float4* bias4 = reinterpret_cast<const float4*>(bias_s + l * 64 + pc * 32);
        uint4 ov[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __half2* o = reinterpret_cast<__half2*>(&ov[i]);
          const __half2* rh = reinterpret_cast<const __half2*>(&res[pc * 4 + i]);
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (!(xf & 32u)) { b0 = bias4[i * 2]; b1 = bias4[i * 2 + 1]; }
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = i * 8 + j * 2;
            float a0 = __uint_as_float(v[c]) + bb[j * 2];
            float a1 = __uint_as_float(v[c + 1]) + bb[j * 2 + 1];
            a0 = fmaxf(a0, a0 * slope);
            a1 = fmaxf(a1, a1 * slope);
            if (has_res) {
              const float2 rf = __half22float2(rh[j]);
              a0 += rf.x; a1 += rf.y;
            }
            o[j] = __floats2half2_rn(a0, a1);
          }
        }
        if (pc == 0 && pending) {
          if (gtid == 0) { fence_acq_rel_gpu(); st_relaxed_gpu(pend_flag, pend_val); CT_TRACE(pend_seq, 3); }
          pending = false;
        }
        if (inb && !(xf & 2u)) {
          st_global_256(orow + pc * 4, ov[0], ov[1]);
          st_global_256(orow + pc * 4 + 2, ov[2], ov[3]);
        }
      }
      if (gtid == 0) CT_TRACE(g, 2);
      if (l + 1 < L && (xf & 128u)) {
        // ablation: publish right behind the stores (group barrier + fence + flag), no deferral
        named_bar_sync(1 + grp, 128);
        if (gtid == 0) { fence_acq_rel_gpu(); st_relaxed_gpu(flags + tile, fbase + (uint32_t)l + 1u); CT_TRACE(g, 3); }
      } else if (l + 1 < L && !(xf & 4u)) {
        pending = true;
        pend_flag = flags + tile;
        pend_val = fbase + (uint32_t)l + 1u;
        pend_seq = g;
      }
    }
    if (pending) {
      named_bar_sync(1 + grp, 128);
      if (gtid == 0) { fence_acq_rel_gpu(); st_relaxed_gpu(pend_flag, pend_val); CT_TRACE(pend_seq, 3); }
    }
    if (timing && gtid == 0 && grp == 0) {
      cp.dbg[b * CT_SLOTS + CT_EPI_TFULL] = te_tfull;
      cp.dbg[b * CT_SLOTS + CT_EPI_TOTAL] = clock64() - t_epi0;
    }
  }

  // ------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  if (threadIdx.x == 0) {
    // last CTA out advances the epoch: the flags of this launch can never satisfy a later one
    __threadfence();
    const uint32_t done = atomicAdd(cp.sync + 1, 1u);
    if (done == (uint32_t)G - 1u) {
      atomicExch(cp.sync + 1, 0u);
      __threadfence();
      atomicExch(cp.sync, (*epoch_s) + 1u);
    }
    if (timing) cp.dbg[b * CT_SLOTS + CT_KERNEL] = clock64() - t_kernel0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn chain_encode_fn() {
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

}  // namespace

extern unsigned long long* tg_conv_timer_buffer();

extern "C" {

size_t tg_conv_chain_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  const size_t tiles = (size_t)tg_ceil_div(w, TW) * tg_ceil_div(h, TH) * n;
  return (kSyncFlags + tiles) * sizeof(uint32_t);
}

int tg_conv_chain_tcgen05(const tg_chain_layer* layers, int n_layers, int n, int h, int w, void* sync_ws,
                          int max_ctas, void* stream) {
  TG_REQUIRE(layers != nullptr && sync_ws != nullptr, TG_E_INVALID, "conv_chain: null pointer");
  TG_REQUIRE(n_layers >= 1 && n_layers <= TG_CHAIN_MAX_LAYERS, TG_E_UNSUPPORTED,
             "conv_chain: n_layers=%d (1..%d)", n_layers, TG_CHAIN_MAX_LAYERS);
  TG_REQUIRE(n > 0 && h > 0 && w > 0, TG_E_INVALID, "conv_chain: bad size n=%d h=%d w=%d", n, h, w);
  TG_REQUIRE(((uintptr_t)sync_ws & 15) == 0, TG_E_INVALID, "conv_chain: sync_ws must be 16-byte aligned");

  ChainParams p;
  p.n_layers = n_layers; p.n = n; p.h = h; p.w = w;
  p.tiles_x = tg_ceil_div(w, TW);
  p.tiles_y = tg_ceil_div(h, TH);
  p.num_tiles = p.tiles_x * p.tiles_y * n;
  p.sync = reinterpret_cast<uint32_t*>(sync_ws);
  p.idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.dbg = tg_conv_timer_buffer();
  p.xflags = 0;
  if (p.dbg != nullptr) {
    const char* e = getenv("TG_CHAIN_ABLATE");
    if (e != nullptr) p.xflags = (uint32_t)atoi(e);
  }

  const void* bufs[kMaxMaps];
  int n_maps = 0;
  for (int l = 0; l < n_layers; ++l) {
    const tg_chain_layer& s = layers[l];
    TG_REQUIRE(s.x && s.weights && s.bias && s.y, TG_E_INVALID, "conv_chain: layer %d: null pointer", l);
    TG_REQUIRE(s.act >= TG_ACT_NONE && s.act <= TG_ACT_LRELU02, TG_E_INVALID, "conv_chain: layer %d: act", l);
    TG_REQUIRE(s.reserved == 0, TG_E_INVALID, "conv_chain: layer %d: reserved must be 0", l);
    TG_REQUIRE(s.y != s.x, TG_E_INVALID, "conv_chain: layer %d: y aliases x (halo reads of other tiles)", l);
    TG_REQUIRE(((uintptr_t)s.x & 15) == 0 && ((uintptr_t)s.y & 31) == 0 && ((uintptr_t)s.weights & 15) == 0 &&
                   ((uintptr_t)s.residual & 31) == 0 && ((uintptr_t)s.bias & 3) == 0,
               TG_E_INVALID, "conv_chain: layer %d: pointer alignment", l);
    int m = -1;
    for (int i = 0; i < n_maps; ++i)
      if (bufs[i] == s.x) m = i;
    if (m < 0) {
      TG_REQUIRE(n_maps < kMaxMaps, TG_E_UNSUPPORTED, "conv_chain: more than %d distinct input buffers", kMaxMaps);
      EncodeTiledFn fn = chain_encode_fn();
      TG_REQUIRE(fn != nullptr, TG_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
      cuuint64_t dims[4] = {64, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
      cuuint64_t strides[3] = {128, (cuuint64_t)w * 128, (cuuint64_t)h * w * 128};
      cuuint32_t box[4] = {64, (cuuint32_t)BOXW, (cuuint32_t)BOXH, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = fn(&p.maps[n_maps], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s.x), dims, strides,
                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      TG_REQUIRE(r == CUDA_SUCCESS, TG_E_DRIVER, "conv_chain: cuTensorMapEncodeTiled failed (%d)", (int)r);
      bufs[n_maps] = s.x;
      m = n_maps++;
    }
    p.layers[l].w = reinterpret_cast<const unsigned char*>(s.weights);
    p.layers[l].bias = s.bias;
    p.layers[l].res = reinterpret_cast<const __half*>(s.residual);
    p.layers[l].y = reinterpret_cast<__half*>(s.y);
    p.layers[l].map = m;
    p.layers[l].act = s.act;
  }
  for (int i = n_maps; i < kMaxMaps; ++i) p.maps[i] = p.maps[0];
  for (int l = n_layers; l < TG_CHAIN_MAX_LAYERS; ++l) p.layers[l] = p.layers[0];

  static TgPerDeviceOnce attr_once;
  const cudaError_t attr_err = attr_once.run([] {
    cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(conv_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  });
  TG_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "conv_chain: cudaFuncSetAttribute: %s",
             cudaGetErrorString(attr_err));

  int sms = 0;
  int rc = tg_device_sm_count(&sms);
  if (rc != TG_OK) return rc;
  // Every CTA must be resident at once (tiles wait on tiles of other CTAs).  Two guards:
  //  (1) the grid never exceeds what the occupancy calculator says fits on this device / partition;
  //  (2) the launch is COOPERATIVE, so the driver only starts the grid when all of its CTAs can be
  //      co-scheduled -- a second chain on another stream, an MPS client or a green-context SM
  //      partition delays or fails the launch instead of dead-locking it into the watchdog trap.
  // TECOGAN_B200_CHAIN_COOP=0 restores the plain PDL launch (A/B measurements on a whole GPU only).
  int per_sm = 0;
  cudaError_t oerr = p.dbg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv_chain_kernel<true>, kThreads, kSmemBytes)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv_chain_kernel<false>, kThreads, kSmemBytes);
  TG_REQUIRE(oerr == cudaSuccess, (int)oerr, "conv_chain: occupancy query: %s", cudaGetErrorString(oerr));
  TG_REQUIRE(per_sm >= 1, TG_E_UNSUPPORTED, "conv_chain: kernel does not fit on an SM of this device");
  int grid = (max_ctas > 0 && max_ctas < sms) ? max_ctas : sms;     // 1 CTA per SM (512 TMEM columns each)
  if (grid > p.num_tiles) grid = p.num_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  static int coop = -1;
  if (coop < 0) {
    const char* e = getenv("TECOGAN_B200_CHAIN_COOP");
    coop = (e != nullptr && e[0] == '0') ? 0 : 1;
    int dev = 0, can = 0;
    if (coop && (cudaGetDevice(&dev) != cudaSuccess ||
                 cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !can))
      coop = 0;
  }
  cudaError_t lerr;
  if (coop)
    lerr = p.dbg ? tg_launch_cooperative(conv_chain_kernel<true>, dim3(grid), dim3(kThreads), kSmemBytes, st, p)
                 : tg_launch_cooperative(conv_chain_kernel<false>, dim3(grid), dim3(kThreads), kSmemBytes, st, p);
  else
    lerr = p.dbg ? tg_launch(conv_chain_kernel<true>, dim3(grid), dim3(kThreads), kSmemBytes, st, p)
                 : tg_launch(conv_chain_kernel<false>, dim3(grid), dim3(kThreads), kSmemBytes, st, p);
  TG_REQUIRE(lerr == cudaSuccess, (int)lerr, "conv_chain: launch failed: %s", cudaGetErrorString(lerr));
  TG_CUDA_LAUNCH_CHECK("conv_chain");
  return TG_OK;
}

}  // extern "C"
