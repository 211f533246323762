// Persistent two-layer SeqLSTM for the FEW-ROW encoder LSTMs (question / history: R = B*10 = 320 rows, T = 20 / 40 steps,
// encoders/mn-att-ques-im-hist.lua:27-45, hrea-ques-im-hist.lua:36,76, lf-*.lua): ONE launch runs both stacked layers over
// all T time steps, forward (k_enc_pair_fwd) or BPTT (k_enc_pair_bwd), instead of 3 launches per time step.
//
// Why (profiles/r01_ncu_full_encoder_small_steps.md, r02_launches_f16_v1.md): the per-step kernels are latency-bound — 360
// launches, 8.3 ms of serialised device time per training step for < 1 % of the FLOPs, each launch re-streaming the 4 MB
// recurrent weight through L2.  Here the weight is STATIONARY: every CTA keeps a 128 KB fp16 slice of it in shared memory
// for the whole sequence, and per step only the (128-row x K) fp16 state panel moves (TMA, 4-stage ring):
//
//   layer-1 CTA (slice of 32 hidden units): gates = xproj1[t] (batched GEMM, up front) + h1_{t-1} Wh1^T        K = H
//   layer-2 CTA (slice of 16 hidden units): gates = [h1_t | h2_{t-1}] [Wx2 | Wh2]^T + b2                       K = 2H
//
// tcgen05 kind::f16, M = 128 rows, N = 4 gates x slice, fp32 accumulators in TMEM; the pointwise half runs in the epilogue
// warps (thread = row) with the cell state held in REGISTERS across the sequence.  A group of (H/32 + H/16) CTAs owns one
// 128-row block; groups loop over row blocks.  Time steps are chained through global-memory flags: a CTA publishes
// "step t of my slice is stored" (fence + atomicAdd), the TMA producer of a consumer CTA spins on the count of the row
// block (ld.acquire), so layer 2 trails layer 1 by one step (wavefront) with no cluster / grid barrier.  A stuck wait traps.
//
// Numerics: fp16 operands (h, weights) with fp32 accumulation = the VD_MATH_F16 class (10-bit mantissa like TF32); cell
// state, gate pre-activations, saved activations and all gradients fp32.
#include <cuda.h>
#include <cuda_fp16.h>
#include "../../include/visdial_b200.h"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vd {
namespace tc {

constexpr int EP_THREADS = 192;          // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (thread = row)
constexpr int EP_STAGES = 4;
constexpr int EP_STAGE_BYTES = 128 * 64 * 2;      // 128 rows x 64 halves
constexpr int EP_W_BYTES_MAX = 131072;            // weight slice: N x K x 2 B = 256*H bytes (H <= 512)
constexpr int EP_SMEM = EP_W_BYTES_MAX + EP_STAGES * EP_STAGE_BYTES + 1024 + 256;

__host__ __device__ constexpr uint32_t ep_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void ep_umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ uint32_t ep_pack2(float a, float b) {
  uint32_t r;
  asm("{\n .reg .f16 lo, hi;\n cvt.rn.satfinite.f16.f32 lo, %1;\n cvt.rn.satfinite.f16.f32 hi, %2;\n mov.b32 %0, {lo, hi};\n}"
      : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint4 ep_pack8(const float* v) {
  return make_uint4(ep_pack2(v[0], v[1]), ep_pack2(v[2], v[3]), ep_pack2(v[4], v[5]), ep_pack2(v[6], v[7]));
}
// spin until *flag >= want (published with fence + atomicAdd by the producers of that step); traps instead of hanging
__device__ __forceinline__ void wait_flag(const int* flag, int want) {
  uint32_t spins = 0;
  for (;;) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= want) break;
    if (++spins > 64) __nanosleep(40);
    if (spins > (1u << 27)) __trap();
  }
  asm volatile("fence.proxy.async;" ::: "memory");       // the TMA (async proxy) reads what generic-proxy stores published
}
__device__ __forceinline__ void ld8g(const float* p, float* d) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void st8g(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

struct EncFwdParams {
  int T, R, H, RB;                 // RB = number of 128-row blocks
  int nS1, nS2, groups;            // slices of layer 1 (H/32), layer 2 (H/16); CTA groups (each owns row blocks g, g+groups, ...)
  float* gates1; float* c1; float* h1; __half* h1_16;      // gates1: in = x-projection (+bias), out = activated gates
  float* gates2; float* c2; float* h2; __half* h2_16;
  const float* bias2;
  const int32_t* mask;             // (T,R) token ids for maskzero, or null
  int* flags;                      // [2][RB][T]: completed slices of (layer, row block, step)
};

__global__ void __launch_bounds__(EP_THREADS, 1)
k_enc_pair_fwd(const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmH2,
               const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const EncFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;                                        // resident weight slice: KBW tiles of [N rows][128 B]
  uint8_t* stages = smem + EP_W_BYTES_MAX;
  uint64_t* full = (uint64_t*)(stages + EP_STAGES * EP_STAGE_BYTES);
  uint64_t* empty = full + EP_STAGES;
  uint64_t* tfull = empty + EP_STAGES;
  uint64_t* tempty = tfull + 1;
  uint64_t* wbar = tempty + 1;
  uint32_t* tmem_slot = (uint32_t*)(wbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, T = p.T;
  const int per_group = p.nS1 + p.nS2;
  const int group = blockIdx.x / per_group, idx = blockIdx.x % per_group;
  const int layer = idx < p.nS1 ? 0 : 1;
  const int slice = layer == 0 ? idx : idx - p.nS1;
  const int HS = layer == 0 ? 32 : 16;             // hidden units of this slice
  const int N = 4 * HS;                            // accumulator columns [i | f | o | g]
  const int KB1 = H / 64;                          // k-blocks of one H-wide operand
  const int KBW = layer == 0 ? KB1 : 2 * KB1;      // k-blocks of the resident weight slice
  int* flag1 = p.flags;
  int* flag2 = p.flags + (size_t)p.RB * T;

  if (threadIdx.x == 0) {
    for (int s = 0; s < EP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 4); mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- weights, once: tile kb = 4 gate boxes of HS rows x 64 halves
      mbar_expect_tx(wbar, (uint32_t)(KBW * N * 128));
      for (int kb = 0; kb < KBW; ++kb)
        for (int g = 0; g < 4; ++g)
          tma_load_2d(wsm + kb * N * 128 + g * HS * 128, layer == 0 ? &tmW1 : &tmW2, wbar, kb * 64, g * H + slice * HS);
      // ---- per step: the state panel(s) of this row block
      int s = 0; uint32_t ph = 0;
      auto load_panel = [&](const CUtensorMap* tm, int row0) {
        for (int kb = 0; kb < KB1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], EP_STAGE_BYTES);
          tma_load_2d(stages + s * EP_STAGE_BYTES, tm, &full[s], kb * 64, row0);
          if (++s == EP_STAGES) { s = 0; ph ^= 1; }
        }
      };
      for (int rb = group; rb < p.RB; rb += p.groups) {
        for (int t = 0; t < T; ++t) {
          if (layer == 0) {
            if (t == 0) continue;
            wait_flag(flag1 + (size_t)rb * T + (t - 1), p.nS1);
            load_panel(&tmH1, (t - 1) * p.R + rb * 128);
          } else {
            if (t > 0) {                                    // recurrent half first: it is ready one step earlier
              wait_flag(flag2 + (size_t)rb * T + (t - 1), p.nS2);
              load_panel(&tmH2, (t - 1) * p.R + rb * 128);
            }
            wait_flag(flag1 + (size_t)rb * T + t, p.nS1);
            load_panel(&tmH1, t * p.R + rb * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: same (row block, step, panel, k-block) order as the producer
    const uint32_t idesc = ep_idesc_f16(128, N);
    mbar_wait(wbar, 0);
    tc_fence_after();
    int s = 0; uint32_t ph = 0; uint32_t nuse = 0;
    for (int rb = group; rb < p.RB; rb += p.groups) {
      for (int t = 0; t < T; ++t) {
        const int npan = layer == 0 ? (t > 0 ? 1 : 0) : (t > 0 ? 2 : 1);
        if (npan == 0) continue;
        mbar_wait(tempty, (nuse & 1) ^ 1);                   // the epilogue has drained the accumulator of the previous step
        tc_fence_after();
        ++nuse;
        uint32_t first = 1;
        for (int pan = 0; pan < npan; ++pan) {
          // weight k-blocks: layer 1 -> 0..KB1-1; layer 2 -> recurrent half KB1..2KB1-1 first (when present), then the x half
          const int kb0 = layer == 0 ? 0 : ((npan == 2 && pan == 0) ? KB1 : 0);
          for (int kb = 0; kb < KB1; ++kb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t sa = smem_u32(stages + s * EP_STAGE_BYTES);
              const uint32_t sb = smem_u32(wsm + (kb0 + kb) * N * 128);
              const uint64_t adesc = make_desc(sa, 16, 1024), bdesc = make_desc(sb, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ep_umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
                first = 0;
              }
              umma_commit(&empty[s]);
              if (pan == npan - 1 && kb == KB1 - 1) umma_commit(tfull);
            }
            __syncwarp();
            first = 0;
            if (++s == EP_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ---- epilogue: thread = row of the 128-row block; cell state in registers across the sequence
    const int q = warp & 3;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int j0 = slice * HS;
    float* gates = layer == 0 ? p.gates1 : p.gates2;
    float* cst = layer == 0 ? p.c1 : p.c2;
    float* hst = layer == 0 ? p.h1 : p.h2;
    __half* h16 = layer == 0 ? p.h1_16 : p.h2_16;
    int* flag = layer == 0 ? flag1 : flag2;
    uint32_t nuse = 0;
    for (int rb = group; rb < p.RB; rb += p.groups) {
      const int64_t row = (int64_t)rb * 128 + q * 32 + lane;
      const bool row_ok = row < p.R;
      float c[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) c[e] = 0.f;
      for (int t = 0; t < T; ++t) {
        const bool has_acc = layer == 1 || t > 0;
        const int64_t tr = (int64_t)t * p.R + row;
        const float keep = (row_ok && p.mask && p.mask[tr] == 0) ? 0.f : 1.f;
        const int nsub = HS / 8;
        // additive term of the pre-activation: layer 1 = this step's x-projection (+ bias) rows, layer 2 = the bias.  The
        // loads of sub-tile s+1 are issued before the math of sub-tile s, those of sub-tile 0 before the accumulator is awaited.
        float xn[4][8];
        auto load_x = [&](int sub) {
          const int j = j0 + sub * 8;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (!row_ok) {
#pragma unroll
              for (int e = 0; e < 8; ++e) xn[g][e] = 0.f;
            } else if (layer == 0) ld8g(gates + tr * 4 * H + g * H + j, xn[g]);
            else ld8g(p.bias2 + g * H + j, xn[g]);
          }
        };
        load_x(0);
        if (layer == 0 && row_ok && t + 1 < T) {             // next step's x-projection rows: pull them into L2 now
#pragma unroll
          for (int g = 0; g < 4; ++g)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gates + (tr + p.R) * 4 * H + g * H + j0) : "memory");
        }
        if (has_acc) { mbar_wait(tfull, nuse & 1); tc_fence_after(); ++nuse; }
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          if (sub < nsub) {
            float a[4][8], hn[8];
            if (has_acc) {
#pragma unroll
              for (int g = 0; g < 4; ++g) tmem_ld8(taddr + g * HS + sub * 8, a[g]);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int e = 0; e < 8; ++e) a[g][e] = 0.f;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int e = 0; e < 8; ++e) a[g][e] += xn[g][e];
            if (sub + 1 < nsub) load_x(sub + 1);
            const int j = j0 + sub * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float gi = fsigmoid(a[0][e]), gf = fsigmoid(a[1][e]), go = fsigmoid(a[2][e]), gg = ftanh(a[3][e]);
              const float c_ = (gf * c[sub * 8 + e] + gi * gg) * keep;       // maskzero: state reset on an all-zero input row
              a[0][e] = gi * keep; a[1][e] = gf * keep; a[2][e] = go * keep; a[3][e] = gg * keep;
              c[sub * 8 + e] = c_;
              hn[e] = go * ftanh(c_) * keep;
            }
            if (row_ok) {
              *reinterpret_cast<uint4*>(h16 + tr * H + j) = ep_pack8(hn);   // what the next step's TMA reads goes out first
#pragma unroll
              for (int g = 0; g < 4; ++g) st8g(gates + tr * 4 * H + g * H + j, a[g]);
              st8g(cst + tr * H + j, &c[sub * 8]);
              st8g(hst + tr * H + j, hn);
            }
          }
        }
        if (has_acc) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
        }
        // publish step t of this slice: CTA barrier, then ONE gpu-scope fence (cumulative over what the barrier made
        // visible to the signalling thread) and the count-in — the grid-sync idiom of cooperative groups
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          __threadfence();
          atomicAdd(flag + (size_t)rb * T + t, 1);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}


// ------------------------------------------------------------------------------------------------
// BPTT of the pair, same structure mirrored in time and in roles:
//   layer-2 CTA (slice of 32 hidden units): dh2_t = da2_{t+1} Wh2            (K = 4H)  [+ dL/dh2 at the last step]
//   layer-1 CTA (slice of 16 hidden units): dh1_t = da2_t Wx2 + da1_{t+1} Wh1 (K = 8H)  [+ dL/dh1 at the last step]
// then the SeqLSTM backward pointwise (saved gates, c_{t-1}, c_t; dc carried in registers) -> da_t as fp32 (what the weight /
// input gradients after the kernel read) and fp16 (the A operand of the steps that follow).  Layer 2 runs ahead of layer 1;
// layer 1's producer streams the da2_t half first (ready early) and the da1_{t+1} half when its own previous step is published.
struct EncBwdParams {
  int T, R, H, RB;
  int nS1, nS2, groups;            // slices of layer 1 (H/16), layer 2 (H/32)
  const float* gates1; const float* c1; float* da1; __half* da1_16;
  const float* gates2; const float* c2; float* da2; __half* da2_16;
  const float* dh_last1; const float* dc_last1; const float* dh_last2; const float* dc_last2;   // (R,H) each or null
  const int32_t* mask;
  int* flags;                      // [2][RB][T]
};

__global__ void __launch_bounds__(EP_THREADS, 1)
k_enc_pair_bwd(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const EncBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;
  uint8_t* stages = smem + EP_W_BYTES_MAX;
  uint64_t* full = (uint64_t*)(stages + EP_STAGES * EP_STAGE_BYTES);
  uint64_t* empty = full + EP_STAGES;
  uint64_t* tfull = empty + EP_STAGES;
  uint64_t* tempty = tfull + 1;
  uint64_t* wbar = tempty + 1;
  uint32_t* tmem_slot = (uint32_t*)(wbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, T = p.T;
  const int per_group = p.nS1 + p.nS2;
  const int group = blockIdx.x / per_group, idx = blockIdx.x % per_group;
  const int layer = idx < p.nS2 ? 1 : 0;                  // 1 = layer 2 (top), 0 = layer 1
  const int slice = layer == 1 ? idx : idx - p.nS2;
  const int HS = layer == 1 ? 32 : 16;                    // hidden units (= accumulator columns) of this slice
  const int KBP = 4 * H / 64;                             // k-blocks of one (rows, 4H) da panel
  const int KBW = layer == 1 ? KBP : 2 * KBP;             // k-blocks of the resident weight slice
  int* flag1 = p.flags;
  int* flag2 = p.flags + (size_t)p.RB * T;

  if (threadIdx.x == 0) {
    for (int s = 0; s < EP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 4); mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wbar, (uint32_t)(KBW * HS * 128));
      for (int kb = 0; kb < KBW; ++kb)
        tma_load_2d(wsm + kb * HS * 128, layer == 1 ? &tmW2 : &tmW1, wbar, kb * 64, slice * HS);
      int s = 0; uint32_t ph = 0;
      auto load_panel = [&](const CUtensorMap* tm, int row0) {
        for (int kb = 0; kb < KBP; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], EP_STAGE_BYTES);
          tma_load_2d(stages + s * EP_STAGE_BYTES, tm, &full[s], kb * 64, row0);
          if (++s == EP_STAGES) { s = 0; ph ^= 1; }
        }
      };
      for (int rb = group; rb < p.RB; rb += p.groups) {
        for (int t = T - 1; t >= 0; --t) {
          if (layer == 1) {
            if (t == T - 1) continue;
            wait_flag(flag2 + (size_t)rb * T + (t + 1), p.nS2);
            load_panel(&tmA2, (t + 1) * p.R + rb * 128);
          } else {
            wait_flag(flag2 + (size_t)rb * T + t, p.nS2);            // da2_t: the layer above is ahead
            load_panel(&tmA2, t * p.R + rb * 128);
            if (t < T - 1) {
              wait_flag(flag1 + (size_t)rb * T + (t + 1), p.nS1);
              load_panel(&tmA1, (t + 1) * p.R + rb * 128);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ep_idesc_f16(128, HS);
    mbar_wait(wbar, 0);
    tc_fence_after();
    int s = 0; uint32_t ph = 0; uint32_t nuse = 0;
    for (int rb = group; rb < p.RB; rb += p.groups) {
      for (int t = T - 1; t >= 0; --t) {
        const int npan = layer == 1 ? (t < T - 1 ? 1 : 0) : (t < T - 1 ? 2 : 1);
        if (npan == 0) continue;
        mbar_wait(tempty, (nuse & 1) ^ 1);
        tc_fence_after();
        ++nuse;
        uint32_t first = 1;
        for (int pan = 0; pan < npan; ++pan) {
          const int kb0 = pan * KBP;                       // layer 1: panel 0 = Wx2 half, panel 1 = Wh1 half; layer 2: Wh2
          for (int kb = 0; kb < KBP; ++kb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t sa = smem_u32(stages + s * EP_STAGE_BYTES);
              const uint32_t sb = smem_u32(wsm + (kb0 + kb) * HS * 128);
              const uint64_t adesc = make_desc(sa, 16, 1024), bdesc = make_desc(sb, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ep_umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
                first = 0;
              }
              umma_commit(&empty[s]);
              if (pan == npan - 1 && kb == KBP - 1) umma_commit(tfull);
            }
            __syncwarp();
            first = 0;
            if (++s == EP_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int j0 = slice * HS;
    const float* gates = layer == 1 ? p.gates2 : p.gates1;
    const float* cst = layer == 1 ? p.c2 : p.c1;
    float* da = layer == 1 ? p.da2 : p.da1;
    __half* da16 = layer == 1 ? p.da2_16 : p.da1_16;
    const float* dh_last = layer == 1 ? p.dh_last2 : p.dh_last1;
    const float* dc_last = layer == 1 ? p.dc_last2 : p.dc_last1;
    int* flag = layer == 1 ? flag2 : flag1;
    uint32_t nuse = 0;
    for (int rb = group; rb < p.RB; rb += p.groups) {
      const int64_t row = (int64_t)rb * 128 + q * 32 + lane;
      const bool row_ok = row < p.R;
      float dc[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) dc[e] = 0.f;
      for (int t = T - 1; t >= 0; --t) {
        const bool has_acc = layer == 0 || t < T - 1;
        const int64_t tr = (int64_t)t * p.R + row;
        const float keep = (row_ok && p.mask && p.mask[tr] == 0) ? 0.f : 1.f;
        const int nsub = HS / 8;
        if (has_acc) { mbar_wait(tfull, nuse & 1); tc_fence_after(); ++nuse; }
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          if (sub < nsub) {
            float dh[8], g[4][8], cp[8], cc[8], out[4][8];
            if (has_acc) { tmem_ld8(taddr + sub * 8, dh); tmem_ld_wait(); }
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) dh[e] = 0.f;
            }
            const int j = j0 + sub * 8;
            if (row_ok) {
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) ld8g(gates + tr * 4 * H + gg * H + j, g[gg]);
              ld8g(cst + tr * H + j, cc);
              if (t > 0) ld8g(cst + (tr - p.R) * H + j, cp);
              else {
#pragma unroll
                for (int e = 0; e < 8; ++e) cp[e] = 0.f;
              }
              if (t == T - 1) {
                if (dh_last) { float x[8]; ld8g(dh_last + row * H + j, x);
#pragma unroll
                  for (int e = 0; e < 8; ++e) dh[e] += x[e]; }
                if (dc_last) { float x[8]; ld8g(dc_last + row * H + j, x);
#pragma unroll
                  for (int e = 0; e < 8; ++e) dc[sub * 8 + e] = x[e]; }
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float gi = g[0][e], gf = g[1][e], go = g[2][e], gg_ = g[3][e];
                const float tcv = ftanh(cc[e]);
                const float dhe = dh[e] * keep;
                const float d = (dc[sub * 8 + e] + dhe * go * (1.f - tcv * tcv)) * keep;
                out[0][e] = d * gg_ * gi * (1.f - gi);
                out[1][e] = d * cp[e] * gf * (1.f - gf);
                out[2][e] = dhe * tcv * go * (1.f - go);
                out[3][e] = d * gi * (1.f - gg_ * gg_);
                dc[sub * 8 + e] = d * gf;
              }
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) {
                *reinterpret_cast<uint4*>(da16 + tr * 4 * H + gg * H + j) = ep_pack8(out[gg]);
                st8g(da + tr * 4 * H + gg * H + j, out[gg]);
              }
            }
          }
        }
        if (has_acc) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          __threadfence();
          atomicAdd(flag + (size_t)rb * T + t, 1);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}

// ------------------------------------------------------------------------------------------------ host
static CUtensorMap ep_tmap_h(const __half* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (enc_lstm) failed (%d): rows %lld cols %lld ld %lld box %d", (int)r, (long long)rows,
             (long long)cols, (long long)ld, box_rows);
    throw CudaError(-3, buf);
  }
  return tm;
}

}  // namespace tc

bool enc_pair_shape_ok(int64_t R, int H, int sm_count) {
  return H % 64 == 0 && H <= 512 && R >= 64 && (H / 32 + H / 16) <= sm_count;
}

// flags: int32 [2 * RB * T] (zeroed here).  gates1 holds the layer-1 x-projection (+ bias) on entry.
void enc_pair_forward(LaunchCtx& cx, int T, int64_t R, int H, const __half* W1h16, const __half* W2cat16, const float* bias2,
                      const int32_t* mask, float* gates1, float* c1, float* h1, __half* h1_16, float* gates2, float* c2, float* h2,
                      __half* h2_16, int* flags) {
  using namespace tc;
  VD_REQUIRE(enc_pair_shape_ok(R, H, cx.sm_count), VD_E_STATE, "enc_pair_forward: shape");
  EncFwdParams p = {};
  p.T = T; p.R = (int)R; p.H = H; p.RB = cdiv(R, 128);
  p.nS1 = H / 32; p.nS2 = H / 16;
  p.groups = std::max(1, std::min(p.RB, cx.sm_count / (p.nS1 + p.nS2)));
  p.gates1 = gates1; p.c1 = c1; p.h1 = h1; p.h1_16 = h1_16;
  p.gates2 = gates2; p.c2 = c2; p.h2 = h2; p.h2_16 = h2_16;
  p.bias2 = bias2; p.mask = mask; p.flags = flags;
  VD_CUDA_CHECK(cudaMemsetAsync(flags, 0, (size_t)2 * p.RB * T * sizeof(int), cx.stream));
  const int64_t TR = (int64_t)T * R;
  CUtensorMap tH1 = ep_tmap_h(h1_16, TR, H, H, 128), tH2 = ep_tmap_h(h2_16, TR, H, H, 128);
  CUtensorMap tW1 = ep_tmap_h(W1h16, 4 * (int64_t)H, H, H, 32), tW2 = ep_tmap_h(W2cat16, 4 * (int64_t)H, 2 * (int64_t)H, 2 * (int64_t)H, 16);
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA_CHECK(cudaFuncSetAttribute(k_enc_pair_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, EP_SMEM));
    attr_set = true;
  }
  k_enc_pair_fwd<<<p.groups * (p.nS1 + p.nS2), EP_THREADS, EP_SMEM, cx.stream>>>(tH1, tH2, tW1, tW2, p);
  check_launch(cx, "k_enc_pair_fwd");
}


// gates*/c* = the activations the forward saved; da*/da*_16 out (all T steps); flags int32 [2 * RB * T]
void enc_pair_backward(LaunchCtx& cx, int T, int64_t R, int H, const __half* B1cat16, const __half* Whb2_16, const int32_t* mask,
                       const float* gates1, const float* c1, const float* gates2, const float* c2, const float* dh_last1,
                       const float* dc_last1, const float* dh_last2, const float* dc_last2, float* da1, __half* da1_16, float* da2,
                       __half* da2_16, int* flags) {
  using namespace tc;
  VD_REQUIRE(enc_pair_shape_ok(R, H, cx.sm_count), VD_E_STATE, "enc_pair_backward: shape");
  EncBwdParams p = {};
  p.T = T; p.R = (int)R; p.H = H; p.RB = cdiv(R, 128);
  p.nS1 = H / 16; p.nS2 = H / 32;
  p.groups = std::max(1, std::min(p.RB, cx.sm_count / (p.nS1 + p.nS2)));
  p.gates1 = gates1; p.c1 = c1; p.da1 = da1; p.da1_16 = da1_16;
  p.gates2 = gates2; p.c2 = c2; p.da2 = da2; p.da2_16 = da2_16;
  p.dh_last1 = dh_last1; p.dc_last1 = dc_last1; p.dh_last2 = dh_last2; p.dc_last2 = dc_last2;
  p.mask = mask; p.flags = flags;
  VD_CUDA_CHECK(cudaMemsetAsync(flags, 0, (size_t)2 * p.RB * T * sizeof(int), cx.stream));
  const int64_t TR = (int64_t)T * R;
  const int64_t G = 4 * (int64_t)H;
  CUtensorMap tA1 = ep_tmap_h(da1_16, TR, G, G, 128), tA2 = ep_tmap_h(da2_16, TR, G, G, 128);
  CUtensorMap tW1 = ep_tmap_h(B1cat16, H, 2 * G, 2 * G, 16), tW2 = ep_tmap_h(Whb2_16, H, G, G, 32);
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA_CHECK(cudaFuncSetAttribute(k_enc_pair_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, EP_SMEM));
    attr_set = true;
  }
  k_enc_pair_bwd<<<p.groups * (p.nS1 + p.nS2), EP_THREADS, EP_SMEM, cx.stream>>>(tA1, tA2, tW1, tW2, p);
  check_launch(cx, "k_enc_pair_bwd");
}

}  // namespace vd
