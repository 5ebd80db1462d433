// tcgen05.mma issue / execution probe for sm_100a (diagnostic tool, not part of the library).
//
// One thread per CTA issues K MMAs (M = 128, K = 16, fp16 in / fp32 accumulate, both operands in
// shared memory) and timestamps every issue and the completion of a trailing tcgen05.commit.  It
// answers, per variant: how long the issuing thread is held per MMA, how long the tensor pipe needs
// per MMA, and how much unrelated work between MMAs the pipe's queue hides.
//
//   make -C tecogan-pytorch_b200/csrc probe      (nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17)
//   tools/_build/mma_probe            (prints one line per variant, cycles of CTA 0 / median over CTAs)
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <utility>

#include "../tecogan-pytorch_b200/csrc/tg_tcgen05.cuh"

namespace {

constexpr int K = 72;                 // MMAs per experiment (two conv tiles' worth)
constexpr int kSlotsOut = K + 8;      // [0..K) issue stamps, K: after commit, K+1: completion

enum Variant {
  V_N64 = 0, V_N64_SAME, V_N128, V_N256, V_N64_COLL_A, V_WS_N64, V_WS_N64_SHARE_B, V_WS_N128_SHARE_B,
  V_N64_GAP200, V_N64_GAP500, V_N64_GAP1000, V_N64_GAP2000,
  V_N64_ONE_D, V_N64_FRESH_D, V_N64_ALT_D1, V_N64_ALT_D4, V_N64_COMMIT36, V_COUNT
};
const char* kNames[V_COUNT] = {
  "N=64 distinct operands", "N=64 same operands", "N=128", "N=256", "N=64 pairs sharing A (collector::a fill/lastuse)",
  ".ws N=64 (collector::b0 discard)", ".ws N=64 pairs sharing B (b0 fill/lastuse)", ".ws N=128 pairs sharing B",
  "N=64, 200-cycle gap after every 36", "N=64, 500-cycle gap after every 36", "N=64, 1000-cycle gap after every 36",
  "N=64, 2000-cycle gap after every 36",
  "N=64, one accumulator for all 72", "N=64, new accumulator at 36 with accumulate=0", "N=64, two accumulators alternating every MMA",
  "N=64, two accumulators alternating every 4", "N=64, commit after 36 (other barrier)"};

__device__ __forceinline__ void umma_coll_a(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc, int use) {
  if (use)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// mode 0 discard, 1 fill, 2 lastuse
__device__ __forceinline__ void umma_ws(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc, int mode) {
  if (mode == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::discard [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else if (mode == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ unsigned int vclock() {
  unsigned int t;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(t)::"memory");
  return t;
}
__device__ __forceinline__ uint32_t idesc_for(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int v>
__global__ void __launch_bounds__(128, 1) probe_kernel(unsigned int* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar = base;                       // one mbarrier
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + 16);
  const uint32_t a0 = base + 1024;                 // A region: 18x10 halo box, 23 KB (+ slack)
  const uint32_t b0 = base + 1024 + 32 * 1024;     // B region: up to 256 rows x 128 B x several slices
  // operands: small finite fp16 values
  for (int i = threadIdx.x; i < (200 * 1024 - 2048) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sm + 1024)[i] = 0x2C002C00u;     // 0.0625 , 0.0625
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_s)), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s;
  unsigned int* o = out + (size_t)blockIdx.x * V_COUNT * kSlotsOut;

  if (threadIdx.x < 32 && elect_one_sync()) {   // elected lane of a converged warp: operands stay in uniform registers
    const uint64_t a_hi = make_sdesc(0, 10u * 128u);
    const uint64_t b_hi = make_sdesc(0, 1024u);
    const uint32_t sa16 = (a0 & 0x3FFFFu) >> 4, sb16 = (b0 & 0x3FFFFu) >> 4;
    uint32_t phase = 0;
    {
      constexpr int n = (v == V_N128 || v == V_WS_N128_SHARE_B) ? 128 : (v == V_N256 ? 256 : 64);
      const uint32_t idesc = idesc_for(n);
      const uint32_t bslice16 = (uint32_t)(n * 128) >> 4;        // one tap's weights: n rows x 128 B
      constexpr int gap = v == V_N64_GAP200 ? 200 : v == V_N64_GAP500 ? 500 : v == V_N64_GAP1000 ? 1000 : v == V_N64_GAP2000 ? 2000 : 0;
      unsigned int* ov = o + v * kSlotsOut;
      // warm the path once
      umma_f16(tmem, a_hi | sa16, b_hi | sb16, idesc, 0);
      umma_commit(bar);
      mbar_wait(bar, phase, 1); phase ^= 1;
      tc_fence_after();
      unsigned int st[K];
      const unsigned int t0 = clock();
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const int j = i % 36, tap = j >> 2, kk = j & 3;
        const uint32_t aoff = (uint32_t)((tap / 3) * 10 + tap % 3) * 8u + 2u * kk;
        uint64_t da = a_hi | (uint64_t)(sa16 + aoff);
        uint64_t db = b_hi | (uint64_t)(sb16 + (uint32_t)(tap % 5) * bslice16 + 2u * kk);
        uint32_t d = tmem + ((i / 36) & 1) * 256u;
        uint32_t acc = i >= 1;
        if (v == V_N64_ONE_D) d = tmem;
        if (v == V_N64_FRESH_D) acc = (i % 36) != 0;
        if (v == V_N64_ALT_D1) { d = tmem + (i & 1) * 256u; acc = i >= 2; }
        if (v == V_N64_ALT_D4) { d = tmem + ((i >> 2) & 1) * 256u; acc = i >= 8; }
        if (v == V_N64_SAME) { da = a_hi | sa16; db = b_hi | sb16; }
        if (v == V_N64_COLL_A) {
          da = a_hi | (uint64_t)(sa16 + (uint32_t)((i >> 1) % 18) * 8u);
          umma_coll_a(d, da, db, idesc, i >= 1, i & 1);
        } else if (v == V_WS_N64) {
          umma_ws(d, da, db, idesc, i >= 1, 0);
        } else if (v == V_WS_N64_SHARE_B || v == V_WS_N128_SHARE_B) {
          db = b_hi | (uint64_t)(sb16 + (uint32_t)((i >> 1) % 5) * bslice16 + 2u * ((i >> 1) & 3));
          umma_ws(d + (i & 1) * 128u, da, db, idesc, i >= 2, (i & 1) ? 2 : 1);
        } else {
          umma_f16(d, da, db, idesc, acc);
          if (v == V_N64_COMMIT36 && i == 35) umma_commit(bar + 8);
        }
        st[i] = clock();
        if (gap && (i % 36) == 35) {
          const unsigned int g0 = vclock();
          while (vclock() - g0 < (unsigned)gap) { }
        }
      }
      umma_commit(bar);
      const unsigned int t1 = clock();
      mbar_wait(bar, phase, 2); phase ^= 1;
      const unsigned int t2 = clock();
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < K; ++i) ov[i] = st[i] - t0;
      ov[K] = t1 - t0;
      ov[K + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// Two issuing warps (one elected lane each, own accumulator and barrier), G cycles of unrelated work
// after EVERY MMA: does a second issuer hide the first one's bookkeeping?
template <int G, int WARPS>
__global__ void __launch_bounds__(128, 1) probe2_kernel(unsigned int* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar = base;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + 32);
  const uint32_t a0 = base + 1024, b0 = base + 1024 + 64 * 1024;
  for (int i = threadIdx.x; i < (200 * 1024 - 2048) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sm + 1024)[i] = 0x2C002C00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_s)), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s;
  const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
  if (warp < WARPS && elect_one_sync()) {
    const uint64_t a_hi = make_sdesc(0, 10u * 128u), b_hi = make_sdesc(0, 1024u);
    const uint32_t sa16 = ((a0 + warp * 32768u) & 0x3FFFFu) >> 4, sb16 = (b0 & 0x3FFFFu) >> 4;
    const uint32_t idesc = idesc_for(64);
    const uint32_t d = tmem + warp * 256u;
    const unsigned int t0 = vclock();
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int j = i % 36, tap = j >> 2, kk = j & 3;
      const uint32_t aoff = (uint32_t)((tap / 3) * 10 + tap % 3) * 8u + 2u * kk;
      umma_f16(d, a_hi | (uint64_t)(sa16 + aoff), b_hi | (uint64_t)(sb16 + (uint32_t)tap * 512u + 2u * kk), idesc, i >= 1);
      if (G) {
        const unsigned int g0 = vclock();
        while (vclock() - g0 < (unsigned)G) { }
      }
    }
    umma_commit(bar + 8 * warp);
    mbar_wait(bar + 8 * warp, 0, 3);
    const unsigned int t2 = vclock();
    out[blockIdx.x * 2 + warp] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int G, int WARPS>
void run_probe2(int grid, int smem) {
  unsigned int* d = nullptr;
  cudaMalloc(&d, grid * 8);
  cudaMemset(d, 0, grid * 8);
  cudaFuncSetAttribute(probe2_kernel<G, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe2_kernel<G, WARPS><<<grid, 128, smem>>>(d);
  cudaDeviceSynchronize();
  std::vector<unsigned int> h(grid * 2);
  cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
  std::vector<double> t;
  for (int b = 0; b < grid; ++b) t.push_back(std::max(h[2 * b], h[2 * b + 1]));
  std::sort(t.begin(), t.end());
  printf("issuers=%d, %3d cycles of other work after every MMA: %6.0f cycles for %d MMAs per issuer = %.1f cycles per MMA (all issuers)\n",
         WARPS, G, t[t.size() / 2], K, t[t.size() / 2] / (K * WARPS));
  cudaFree(d);
}

}  // namespace

template <int... Vs>
void launch_all(int grid, int smem, unsigned int* d_out, std::integer_sequence<int, Vs...>) {
  (cudaFuncSetAttribute(probe_kernel<Vs>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), ...);
  for (int rep = 0; rep < 2; ++rep) ((probe_kernel<Vs><<<grid, 128, smem>>>(d_out)), ...);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  unsigned int* d_out = nullptr;
  const size_t n_out = (size_t)grid * V_COUNT * kSlotsOut;
  cudaMalloc(&d_out, n_out * 4);
  cudaMemset(d_out, 0, n_out * 4);
  const int smem = 201 * 1024;
  launch_all(grid, smem, d_out, std::make_integer_sequence<int, V_COUNT>{});
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<unsigned int> h(n_out);
  cudaMemcpy(h.data(), d_out, n_out * 4, cudaMemcpyDeviceToHost);
  printf("grid=%d CTAs, K=%d MMAs (M=128, K=16) per variant; cycles\n", grid, K);
  for (int v = 0; v < V_COUNT; ++v) {
    std::vector<double> done, issue_end, first8, last;
    for (int b = 0; b < grid; ++b) {
      const unsigned int* o = h.data() + ((size_t)b * V_COUNT + v) * kSlotsOut;
      done.push_back(o[K + 1]);
      issue_end.push_back(o[K - 1]);
      first8.push_back(o[7]);
      last.push_back((o[K - 1] - o[K - 9]) / 8.0);
    }
    auto med = [](std::vector<double> x) { std::sort(x.begin(), x.end()); return x[x.size() / 2]; };
    printf("%-52s done %7.0f (%.1f/MMA)  issue_end %7.0f  first8 issued by %5.0f  last8 spacing %5.1f\n", kNames[v], med(done),
           med(done) / K, med(issue_end), med(first8), med(last));
    if (v == 0 || v == 2 || v >= V_N64_GAP200) {
      const unsigned int* o = h.data() + (size_t)v * kSlotsOut;
      printf("    cta0 issue stamps:");
      for (int i = 0; i < K; ++i) printf(" %u", o[i]);
      printf("\n");
    }
  }
  cudaFree(d_out);
  run_probe2<0, 1>(grid, smem);  run_probe2<0, 2>(grid, smem);
  run_probe2<20, 1>(grid, smem); run_probe2<20, 2>(grid, smem);
  run_probe2<40, 1>(grid, smem); run_probe2<40, 2>(grid, smem);
  run_probe2<80, 1>(grid, smem); run_probe2<80, 2>(grid, smem);
  return 0;
}
