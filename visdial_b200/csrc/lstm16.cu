// SeqLSTM over MANY rows (the disc decoder's option LSTM, decoders/disc.lua:4-20: R = N*100 = 32 000 rows at B = 32)
// with 16-bit operands and 16-bit saved state: VD_MATH_F16.
//
// Why: ncu of the TF32 step kernels (profiles/r01_*) showed them bound by BYTES, not by the tensor pipe — the L2->SM
// operand stream of fp32 tiles (1.0 GB per launch, ~80 % of the ~6300 B/clk LTS cap) and the HBM traffic of the fp32
// saved activations (0.5 GB forward, 1.05 GB backward per step).  A TF32 operand keeps 10 mantissa bits of the fp32
// word it reads; an fp16 word carries the same 10 bits in half the bytes (the exponent range is what is given up: h and
// the weights live well inside it, the gradients are scaled by a power of two chosen from max|dL/dh_T| so that they do
// too — exact, undone in the weight-gradient epilogue).  So: h_t, the x-projection table, the activated gates and da_t
// are stored as fp16, the contractions run as tcgen05 kind::f16 (2x the TF32 rate) with fp32 accumulation in TMEM,
// and c_t, dc, every accumulator, the weight gradients and the final h_T that meets the encoder stay fp32.
//
//   k_lstm16_fwd   : gates = h_{t-1} Wh^T (tcgen05, CTA pairs 256x256) + P16[token] + bias -> pointwise -> fp16 gates,
//                    fp32 c_t, fp16 h_t (+ fp32 h_T on the last step)
//   k_lstm16_bwd   : dh = da_{t+1} Wh (tcgen05) -> backward pointwise -> fp16 da_t, fp32 dc carry
//   k_atb16        : dWh += inv_scale * h^T da  (both operands MN-major fp16, split-K, red.global.add)
//   k_lstm16_first / k_lstm16_bwd_last / k_segsum16 / k_cvt16 / k_amax / k_pick_scale : streaming helpers
#include <cuda.h>
#include <cuda_fp16.h>
#include "../../include/visdial_b200.h"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vd {
namespace tc {

constexpr int BM16 = 128;        // rows per CTA (UMMA M = 256 per pair)
constexpr int BK16 = 64;         // halves per k-block = one 128-byte swizzle row
constexpr int UK16 = 16;         // kind::f16: 32 bytes per instruction
// epilogue warps: 16 = four per TMEM lane quarter, each with its own share of the tile's columns.  The pointwise halves are
// latency-bound per warp (gather / saved-activation loads, MUFU chains): four warps per scheduler hide what two could not
// (forward step 176 -> 109 us, with the pad-token gather skip).  The backward step stays at 8: its 16-warp variant needs 160 KB
// of staging, which leaves two 32 KB pipeline stages for a K = 2048 main loop — measured 209 us against 138 us with 8 warps / 4 stages
template <int MODE> struct EW16T { static constexpr int N = MODE == 0 ? 16 : 8; };
constexpr int STAGE16 = 32768;   // 16 KB of A (this CTA's 128 rows) + 16 KB of B (this CTA's half of the 256-column tile)

// instruction descriptor, kind::f16: D = f32 (c_format 1), A = B = f16 (format 0)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// sigmoid / tanh as {FMUL, MUFU.EX2, FADD, MUFU.RCP [, FFMA]}: abs error ~1e-7 like fsigmoid / ftanh of tc_ptx.cuh, without the
// range fix-ups of __fdividef / copysign (ex2 -> +inf gives rcp -> 0, ex2 -> 0 gives 1: both limits are exact)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// Activations of the fp16 option LSTM: ONE special-function op each (tanh.approx.f32, relative error 2^-11 — the rounding class of the fp16
// the gates and h are stored in right after; sigmoid(x) = 0.5 tanh(x/2) + 0.5) instead of ex2 + rcp (two): the pointwise half of the
// forward step is issue / MUFU-bound.  Measured at the benched size (tests/test_c4_b32_gpu.py): rank agreement with the fp64
// oracle 0.9781 vs 0.9777, top-1 0.99375 both, R@k deltas 0 both.  -DVD_F16_EX2_ACTIVATIONS restores the two-op forms.
#ifndef VD_F16_EX2_ACTIVATIONS
__device__ __forceinline__ float tanh16(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sig16(float x) { return fmaf(0.5f, tanh16(0.5f * x), 0.5f); }
#else
__device__ __forceinline__ float sig16(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh16(float x) { return fmaf(2.f, rcp_approx(1.f + ex2_approx(-2.8853900817779268f * x)), -1.f); }
#endif

// ---- fp16 <-> fp32 packing (round to nearest even, saturating: a scaled gradient that outgrows the range clamps to
// +-65504 instead of becoming inf)
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  asm("{\n .reg .f16 lo, hi;\n cvt.rn.satfinite.f16.f32 lo, %1;\n cvt.rn.satfinite.f16.f32 hi, %2;\n mov.b32 %0, {lo, hi};\n}"
      : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4 u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// ---- epilogue staging tiles (one set per epilogue warp, 32 rows each).  Two geometries, both laid out exactly like the
// TMA box that leaves (or could enter) them, so stores are single bulk-tensor instructions:
//   S32: [32 rows][16 halves] = 32-byte rows, 16-byte chunk c of row r at  r*32 + ((c ^ ((r>>2)&1)) << 4)   (SWIZZLE_32B)
//   S64: [32 rows][16 floats] = 64-byte rows, 16-byte chunk c of row r at  r*64 + ((c ^ ((r>>1)&3)) << 4)   (SWIZZLE_64B)
// Both are bank-conflict free for "thread = row" 16-byte accesses and for the coalesced global side.
constexpr int S32_BYTES = 1024, S64_BYTES = 2048;
__device__ __forceinline__ uint4* s32_at(uint8_t* base, int row, int c) {
  return reinterpret_cast<uint4*>(base + row * 32 + ((c ^ ((row >> 2) & 1)) << 4));
}
__device__ __forceinline__ float4* s64_at(uint8_t* base, int row, int c) {
  return reinterpret_cast<float4*>(base + row * 64 + ((c ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ const void* shfl_vptr(const void* p, int src_lane) {
  unsigned long long v = (unsigned long long)p;
  unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src_lane), hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src_lane);
  return (const void*)(((unsigned long long)hi << 32) | lo);
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// global -> S32 tile: `mine` = this lane's row base (16 halves) or nullptr (zeros); 2 lanes per row, 2 passes
__device__ __forceinline__ void s32_load(uint8_t* base, const void* mine, int lane) {
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int row = ps * 16 + (lane >> 1), c = lane & 1;
    const uint8_t* src = (const uint8_t*)shfl_vptr(mine, row);
    uint4* dst = s32_at(base, row, c);
    if (src) cp16(dst, src + c * 16); else *dst = make_uint4(0, 0, 0, 0);
  }
}
// global -> S64 tile: 4 lanes per row, 4 passes
__device__ __forceinline__ void s64_load(uint8_t* base, const void* mine, int lane) {
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int row = ps * 8 + (lane >> 2), c = lane & 3;
    const uint8_t* src = (const uint8_t*)shfl_vptr(mine, row);
    float4* dst = s64_at(base, row, c);
    if (src) cp16(dst, src + c * 16); else *dst = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void cp_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
}
__device__ __forceinline__ void s64_get8(uint8_t* base, int row, int sub, float* d) {
  const float4 a = *s64_at(base, row, sub * 2), b = *s64_at(base, row, sub * 2 + 1);
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void s64_put8(uint8_t* base, int row, int sub, const float* v) {
  *s64_at(base, row, sub * 2) = make_float4(v[0], v[1], v[2], v[3]);
  *s64_at(base, row, sub * 2 + 1) = make_float4(v[4], v[5], v[6], v[7]);
}

struct Lstm16Params {
  int R, H;
  // forward
  const __half* ptable; const int32_t* tok;   // x-projection table (V+1, 4H) WITHOUT bias, token id per row
  const float* bias;                           // (4H) fp32, added in the epilogue
  const float* c_prev;                         // (R,H) fp32 or null (zeros)
  const int32_t* mask_ids;                     // maskzero ids or null
  float* h32_out;                              // optional fp32 copy of h (the step whose h meets fp32 consumers)
  int save_gates;
  // backward
  const __half* gsave; const float* c_cur; float* dc_carry;
};
struct Lstm16Maps { CUtensorMap g16, c, h16; };   // [R,4H] fp16 gates / da ; [R,H] fp32 c / dc ; [R,H] fp16 h

// ------------------------------------------------------------------------------------------------
// MODE 0 = forward step, MODE 1 = backward step.  2-CTA clusters, persistent over the tile list.
template <int MODE>
struct Cfg16 {
  static constexpr int EW = EW16T<MODE>::N;
  static constexpr int THREADS = 64 + 32 * EW;
  // per-warp staging: forward {4 gate tiles S32, c tile S64, h tile S32} = 7 KB (inputs land here by cp.async, outputs
  // overwrite them in place and leave by TMA); backward {4 gate tiles S32, c_prev, c_t, dc S64} = 10 KB
  static constexpr int STG_PER_WARP = MODE == 0 ? (4 * S32_BYTES + S64_BYTES + S32_BYTES) : (4 * S32_BYTES + 3 * S64_BYTES);
  static constexpr int BIAS_BYTES = MODE == 0 ? 4 * 512 * 4 : 0;                     // fp32 bias of all 4H gate columns (H <= 512)
  static constexpr int STG_BYTES = EW * STG_PER_WARP + BIAS_BYTES;
  static constexpr int STAGES = (232448 - 1024 - 256 - STG_BYTES) / STAGE16;      // fwd 3, bwd 4
  static constexpr int TOTAL = STAGES * STAGE16 + STG_BYTES + 1024 + 256;
};

template <int MODE>
__global__ void __launch_bounds__(Cfg16<MODE>::THREADS, 1)
k_lstm16(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
         const __grid_constant__ Lstm16Maps em, const Lstm16Params p) {
  using C = Cfg16<MODE>;
  constexpr int STAGES = C::STAGES;
  constexpr int BN = 256;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cta = (int)(blockIdx.x >> 1), ncta = (int)(gridDim.x >> 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stg_all = smem + STAGES * STAGE16;
  uint64_t* full = (uint64_t*)(stg_all + C::STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H;
  const int num_m = (p.R + 2 * BM16 - 1) / (2 * BM16);
  const int num_n = MODE == 0 ? H / 64 : H / 256;
  const int num_tiles = num_m * num_n;
  const int num_kb = (MODE == 0 ? H : 4 * H) / BK16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], C::EW * 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_cg2(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: this CTA's 128 rows of A and its half of the B tile; all transactions of the pair
      // complete on the LEADER's full barrier
      int s = 0; uint32_t ph = 0;
      for (int tile = cta; tile < num_tiles; tile += ncta) {
        const int m0 = (tile / num_n) * 2 * BM16 + (int)rank * BM16;
        const int nt = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE16;
          uint8_t* sb = sa + 16384;
          const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
          if (leader) mbar_expect_tx(&full[s], 2 * STAGE16);
          tma_load_2d_cg2(sa, &tmA, bar, kb * BK16, m0);
          if (MODE == 0) {
            // tile columns = [i | f | o | g] of 64 hidden units: the leader stages gate blocks i,f, the peer o,g
#pragma unroll
            for (int g = 0; g < 2; ++g)
              tma_load_2d_cg2(sb + g * 8192, &tmB, bar, kb * BK16, ((int)rank * 2 + g) * H + nt * 64);
          } else {
            tma_load_2d_cg2(sb, &tmB, bar, kb * BK16, nt * BN + (int)rank * 128);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && !leader) {
    // peer CTA: its MMA warp only takes part in TMEM alloc / dealloc
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    constexpr uint32_t idesc = make_idesc_f16(2 * BM16, BN, 0, 0);
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int tile = cta; tile < num_tiles; tile += ncta, ++it) {
      const int buf = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      mbar_wait(&tempty[buf], bph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + s * STAGE16);
          const uint64_t adesc = make_desc(sa, 16, 1024);
          const uint64_t bdesc = make_desc(sa + 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK16 / UK16; ++k)
            umma_f16_cg2(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          umma_commit_cg2(&empty[s]);
          if (kb == num_kb - 1) umma_commit_cg2(&tfull[buf]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* stg = stg_all + (warp - 2) * C::STG_PER_WARP;
    if constexpr (MODE == 0) {
      // ---- forward.  64 hidden units per tile = four column groups of 16; the four warps of a TMEM lane quarter take one
      // group each.  Per tile and warp: inputs (x-projection rows gathered from the fp16 table, previous cell) -> staging by
      // cp.async (issued before the accumulator is awaited), TMEM + staging -> gates / c / h in place, out by TMA.
      const int grp = (warp - 2) >> 2;
      float* sBias = reinterpret_cast<float*>(stg_all + C::EW * C::STG_PER_WARP);
      for (int i = (int)threadIdx.x - 64; i < 4 * H; i += 32 * C::EW) sBias[i] = __ldg(p.bias + i);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * C::EW) : "memory");          // epilogue warps only
      uint8_t* sG = stg; uint8_t* sC = stg + 4 * S32_BYTES; uint8_t* sH = sC + S64_BYTES;
      int it = 0;
      for (int tile = cta; tile < num_tiles; tile += ncta, ++it) {
        const int buf = it & 1;
        const uint32_t bph = (it >> 1) & 1;
        const int m0 = (tile / num_n) * 2 * BM16 + (int)rank * BM16;
        const int nt = tile % num_n;
        const int64_t row = (int64_t)m0 + q * 32 + lane;
        const bool row_ok = row < p.R;
        const float keep = (row_ok && p.mask_ids && p.mask_ids[row] == 0) ? 0.f : 1.f;
        const int r0 = m0 + q * 32;
        const int j = nt * 64 + grp * 16;                      // first hidden unit of this warp's group
        const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16) + grp * 16;
        // A pad token's x-projection is exactly zero (LookupTableMaskZero: embedding row 0 is zero, and the table carries no
        // bias), so finished sequences skip the gather: late time steps, where most of the 100 x 20-token options have ended,
        // would otherwise send every row of the machine to the same 4 KB of L2 (measured: 143 us at t = 1 -> 219 us at t = 19)
        const int32_t tk = row_ok ? __ldg(p.tok + row) : 0;
        const __half* prow = tk != 0 ? p.ptable + (int64_t)tk * 4 * H + j : nullptr;
        const float* cprow = (row_ok && p.c_prev) ? p.c_prev + row * H + j : nullptr;
        if (lane == 0) bulk_wait_read0();                      // the previous tile's TMA stores have read the staging tiles
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 4; ++g) s32_load(sG + g * S32_BYTES, prow ? prow + g * H : nullptr, lane);
        s64_load(sC, cprow, lane);
        mbar_wait(&tfull[buf], bph);                           // the loads fly while the MMAs of this tile finish
        tc_fence_after();
        cp_wait_all();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          float a[4][8], cp[8], hn[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld8(taddr + g * 64 + sub * 8, a[g]);
          tmem_ld_wait();
          s64_get8(sC, lane, sub, cp);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float x[8];
            unpack8(*s32_at(sG + g * S32_BYTES, lane, sub), x);
            const float4 b0 = *reinterpret_cast<const float4*>(sBias + g * H + j + sub * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(sBias + g * H + j + sub * 8 + 4);
            a[g][0] += x[0] + b0.x; a[g][1] += x[1] + b0.y; a[g][2] += x[2] + b0.z; a[g][3] += x[3] + b0.w;
            a[g][4] += x[4] + b1.x; a[g][5] += x[5] + b1.y; a[g][6] += x[6] + b1.z; a[g][7] += x[7] + b1.w;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float gi = sig16(a[0][e]), gf = sig16(a[1][e]), go = sig16(a[2][e]), gg = tanh16(a[3][e]);
            const float c_ = (gf * cp[e] + gi * gg) * keep;
            a[0][e] = gi * keep; a[1][e] = gf * keep; a[2][e] = go * keep; a[3][e] = gg * keep;
            cp[e] = c_; hn[e] = go * tanh16(c_) * keep;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) *s32_at(sG + g * S32_BYTES, lane, sub) = pack8(a[g]);
          s64_put8(sC, lane, sub, cp);
          *s32_at(sH, lane, sub) = pack8(hn);
          if (p.h32_out && row_ok) {                  // last step only: the fp32 h that meets the encoder output
            float4* o = reinterpret_cast<float4*>(p.h32_out + row * H + j + sub * 8);
            o[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
            o[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
          }
        }
        tc_fence_before();                            // the accumulator has been read: hand it back before the stores leave
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (!leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
          else mbar_arrive(&tempty[buf]);
          if (p.save_gates) {
#pragma unroll
            for (int g = 0; g < 4; ++g) tma_store_2d(&em.g16, sG + g * S32_BYTES, g * H + j, r0);
          }
          tma_store_2d(&em.c, sC, j, r0);
          tma_store_2d(&em.h16, sH, j, r0);
          bulk_commit();
        }
        __syncwarp();
      }
    } else {
    int it = 0;
    for (int tile = cta; tile < num_tiles; tile += ncta, ++it) {
      const int buf = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      const int m0 = (tile / num_n) * 2 * BM16 + (int)rank * BM16;
      const int nt = tile % num_n;
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      const bool row_ok = row < p.R;
      const bool masked = row_ok && p.mask_ids && p.mask_ids[row] == 0;
      const float keep = masked ? 0.f : 1.f;
      const int r0 = m0 + q * 32;
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      {
        // ---- backward: 256 hidden units per tile, the warp's column slice = 256 / (EW/4) of them, in groups of 16
        constexpr int NSL = C::EW / 4, GPW = 256 / NSL / 16;      // column slices per tile, groups per warp
        const int j0 = nt * 256 + half * (256 / NSL);
        const __half* grow = row_ok ? p.gsave + row * 4 * H : nullptr;
        const float* cprow = (row_ok && p.c_prev) ? p.c_prev + row * H : nullptr;
        const float* ccrow = row_ok ? p.c_cur + row * H : nullptr;
        const float* dcrow = row_ok ? p.dc_carry + row * H : nullptr;
        uint8_t* sG = stg; uint8_t* sCP = stg + 4 * S32_BYTES; uint8_t* sCC = sCP + S64_BYTES; uint8_t* sDC = sCC + S64_BYTES;
        bool waited = false;
#pragma unroll 1
        for (int grp = 0; grp < GPW; ++grp) {
          const int j = j0 + grp * 16;
          const int tc0 = half * (256 / NSL) + grp * 16;
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g) s32_load(sG + g * S32_BYTES, grow ? grow + g * H + j : nullptr, lane);
          s64_load(sCP, cprow ? cprow + j : nullptr, lane);
          s64_load(sCC, ccrow ? ccrow + j : nullptr, lane);
          s64_load(sDC, dcrow ? dcrow + j : nullptr, lane);
          if (!waited) { mbar_wait(&tfull[buf], bph); tc_fence_after(); waited = true; }
          cp_wait_all();
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            float dh[8], g[4][8], cp[8], cc[8], dc[8], out[4][8], dcn[8];
            tmem_ld8(taddr + tc0 + sub * 8, dh);
            tmem_ld_wait();
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) unpack8(*s32_at(sG + gg * S32_BYTES, lane, sub), g[gg]);
            s64_get8(sCP, lane, sub, cp);
            s64_get8(sCC, lane, sub, cc);
            s64_get8(sDC, lane, sub, dc);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float gi = g[0][e], gf = g[1][e], go = g[2][e], gg_ = g[3][e];
              const float tcv = tanh16(cc[e]);
              const float d = (dc[e] + dh[e] * go * (1.f - tcv * tcv)) * keep;
              const float dhe = dh[e] * keep;
              out[0][e] = d * gg_ * gi * (1.f - gi);
              out[1][e] = d * cp[e] * gf * (1.f - gf);
              out[2][e] = dhe * tcv * go * (1.f - go);
              out[3][e] = d * gi * (1.f - gg_ * gg_);
              dcn[e] = d * gf;
            }
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) *s32_at(sG + gg * S32_BYTES, lane, sub) = pack8(out[gg]);
            s64_put8(sDC, lane, sub, dcn);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) tma_store_2d(&em.g16, sG + gg * S32_BYTES, gg * H + j, r0);
            tma_store_2d(&em.c, sDC, j, r0);
            bulk_commit();
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                // the accumulator buffer may be overwritten by the leader's MMAs
        if (!leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
        else mbar_arrive(&tempty[buf]);
      }
    }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all TMA stores performed
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_cg2(tmem_base, 2 * BN); }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of the recurrent block:  C[m,n] += inv_scale * sum_k A[k,m] B[k,n],  A = h (K rows x H) fp16,
// B = da (K rows x 4H) fp16 — both MN-major (the contraction index is the row index in HBM).  A TMA box of 64 columns
// x KB rows lands as KB rows of 128 bytes = the canonical MN-major SWIZZLE_128B layout (8 k-rows per 1024-byte atom:
// SBO = 1024; 64-column groups LBO bytes apart); one K = 16 instruction consumes two atoms.
constexpr int A16_KB = 64;                 // k-rows per stage (4 MMAs)
constexpr int A16_BN = 256;
constexpr int A16_THREADS = 192;
constexpr int A16_A_BYTES = 2 * A16_KB * 128;              // 128 columns of A = 2 boxes
constexpr int A16_B_BYTES = (A16_BN / 64) * A16_KB * 128;  // 4 boxes
constexpr int A16_STAGE = A16_A_BYTES + A16_B_BYTES;       // 48 KB
constexpr int A16_STAGES = 4;
constexpr int A16_TOTAL = A16_STAGES * A16_STAGE + 1024 + 256;
struct Atb16Params { int M, N; int64_t K, k_per_split; float* C; int64_t ldc; const float* inv_scale; };

__global__ void __launch_bounds__(A16_THREADS, 1)
k_atb16(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Atb16Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + A16_STAGES * A16_STAGE);
  uint64_t* empty = full + A16_STAGES;
  uint64_t* tfull = empty + A16_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = (p.N + A16_BN - 1) / A16_BN;
  const int m0 = (blockIdx.x / num_n) * BM16, n0 = (blockIdx.x % num_n) * A16_BN;
  const int64_t kbeg = (int64_t)blockIdx.y * p.k_per_split;
  const int64_t kend = min(p.K, kbeg + p.k_per_split);
  const int num_kb = (int)((kend - kbeg + A16_KB - 1) / A16_KB);

  if (threadIdx.x == 0) {
    for (int s = 0; s < A16_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, A16_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * A16_STAGE;
        uint8_t* sb = sa + A16_A_BYTES;
        mbar_expect_tx(&full[s], A16_STAGE);
        const int krow = (int)(kbeg + (int64_t)kb * A16_KB);     // k_per_split % A16_KB == 0: only the global K tail is partial
#pragma unroll
        for (int g = 0; g < 2; ++g) tma_load_2d(sa + g * A16_KB * 128, &tmA, &full[s], m0 + g * 64, krow);
#pragma unroll
        for (int g = 0; g < A16_BN / 64; ++g) tma_load_2d(sb + g * A16_KB * 128, &tmB, &full[s], n0 + g * 64, krow);
        if (++s == A16_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(BM16, A16_BN, 1, 1);
    int s = 0; uint32_t ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + s * A16_STAGE);
        const uint32_t sb = sa + A16_A_BYTES;
#pragma unroll
        for (int k = 0; k < A16_KB / UK16; ++k) {
          const uint64_t adesc = make_desc(sa + k * 2048, A16_KB * 128, 1024, 2);
          const uint64_t bdesc = make_desc(sb + k * 2048, A16_KB * 128, 1024, 2);
          umma_f16(tmem_base, adesc, bdesc, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        if (kb == num_kb - 1) umma_commit(tfull);
      }
      __syncwarp();
      if (++s == A16_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (num_kb > 0) {
    const int q = warp & 3;
    const float sc = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int m = m0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < A16_BN; c += 8) {
      if (n0 + c >= p.N) break;
      float v[8];
      tmem_ld8(taddr + c, v);
      tmem_ld_wait();
      if (m < p.M) {
        float* crow = p.C + (int64_t)m * p.ldc + n0 + c;
        if (n0 + c + 8 <= p.N && ((p.ldc & 3) == 0)) {
          red_add_v4(crow, v[0] * sc, v[1] * sc, v[2] * sc, v[3] * sc);
          red_add_v4(crow + 4, v[4] * sc, v[5] * sc, v[6] * sc, v[7] * sc);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + c + j < p.N) atomicAdd(crow + j, v[j] * sc);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, A16_BN); }
}

// ------------------------------------------------------------------------------------------------ streaming helpers
// t = 0 without initial state: pre-activation = P16[tok] + bias.  One thread per (row, 8 hidden units).
__global__ void __launch_bounds__(256)
k_lstm16_first(const __half* __restrict__ ptable, const int32_t* __restrict__ tok, const float* __restrict__ bias,
               const float* __restrict__ c_prev, const int32_t* __restrict__ mask_ids, __half* __restrict__ gates,
               float* __restrict__ c_out, __half* __restrict__ h16, float* __restrict__ h32, int64_t R, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int H8 = H >> 3;
  if (idx >= R * H8) return;
  const int64_t r = idx / H8;
  const int j = (int)(idx % H8) * 8;
  const float keep = (mask_ids && mask_ids[r] == 0) ? 0.f : 1.f;
  const int32_t tk = tok[r];
  const __half* src = ptable + (int64_t)tk * 4 * H;
  float a[4][8], cp[8], cn[8], hn[8];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    unpack8(tk != 0 ? __ldg(reinterpret_cast<const uint4*>(src + g * H + j)) : make_uint4(0, 0, 0, 0), a[g]);   // pad: exactly zero
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + g * H + j)), b1 = __ldg(reinterpret_cast<const float4*>(bias + g * H + j) + 1);
    a[g][0] += b0.x; a[g][1] += b0.y; a[g][2] += b0.z; a[g][3] += b0.w; a[g][4] += b1.x; a[g][5] += b1.y; a[g][6] += b1.z; a[g][7] += b1.w;
  }
  if (c_prev) {
    const float4 c0 = *reinterpret_cast<const float4*>(c_prev + r * H + j), c1 = *(reinterpret_cast<const float4*>(c_prev + r * H + j) + 1);
    cp[0] = c0.x; cp[1] = c0.y; cp[2] = c0.z; cp[3] = c0.w; cp[4] = c1.x; cp[5] = c1.y; cp[6] = c1.z; cp[7] = c1.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) cp[e] = 0.f;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float gi = fsigmoid(a[0][e]), gf = fsigmoid(a[1][e]), go = fsigmoid(a[2][e]), gg = ftanh(a[3][e]);
    const float c_ = gf * cp[e] + gi * gg;
    a[0][e] = gi * keep; a[1][e] = gf * keep; a[2][e] = go * keep; a[3][e] = gg * keep;
    cn[e] = c_ * keep; hn[e] = go * ftanh(c_) * keep;
  }
  if (gates) {
#pragma unroll
    for (int g = 0; g < 4; ++g) __stcs(reinterpret_cast<uint4*>(gates + r * 4 * H + g * H + j), pack8(a[g]));
  }
  float4* co = reinterpret_cast<float4*>(c_out + r * H + j);
  co[0] = make_float4(cn[0], cn[1], cn[2], cn[3]); co[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
  *reinterpret_cast<uint4*>(h16 + r * H + j) = pack8(hn);
  if (h32) {
    float4* ho = reinterpret_cast<float4*>(h32 + r * H + j);
    ho[0] = make_float4(hn[0], hn[1], hn[2], hn[3]); ho[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
  }
}

// t = T-1 of the BPTT: no recurrent gradient yet; dh = scale * dh_last (the power-of-two gradient scale enters here)
__global__ void __launch_bounds__(256)
k_lstm16_bwd_last(const __half* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_cur,
                  const float* __restrict__ dh_last, const float* __restrict__ scale, const int32_t* __restrict__ mask_ids,
                  float* __restrict__ dc_carry, __half* __restrict__ da, int64_t R, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int H8 = H >> 3;
  if (idx >= R * H8) return;
  const int64_t r = idx / H8;
  const int j = (int)(idx % H8) * 8;
  const float keep = (mask_ids && mask_ids[r] == 0) ? 0.f : 1.f;
  const float s = __ldg(scale);
  float g[4][8], cp[8], cc[8], dh[8], out[4][8], dcn[8];
#pragma unroll
  for (int gg = 0; gg < 4; ++gg) unpack8(__ldcs(reinterpret_cast<const uint4*>(gates + r * 4 * H + gg * H + j)), g[gg]);
  auto ld8f = [](const float* p, float* d) {
    const float4 x0 = *reinterpret_cast<const float4*>(p), x1 = *(reinterpret_cast<const float4*>(p) + 1);
    d[0] = x0.x; d[1] = x0.y; d[2] = x0.z; d[3] = x0.w; d[4] = x1.x; d[5] = x1.y; d[6] = x1.z; d[7] = x1.w;
  };
  if (c_prev) ld8f(c_prev + r * H + j, cp);
  else {
#pragma unroll
    for (int e = 0; e < 8; ++e) cp[e] = 0.f;
  }
  ld8f(c_cur + r * H + j, cc);
  ld8f(dh_last + r * H + j, dh);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float gi = g[0][e], gf = g[1][e], go = g[2][e], gg_ = g[3][e];
    const float dhe = dh[e] * s * keep;
    const float tcv = tanh16(cc[e]);
    const float d = dhe * go * (1.f - tcv * tcv);
    out[0][e] = d * gg_ * gi * (1.f - gi);
    out[1][e] = d * cp[e] * gf * (1.f - gf);
    out[2][e] = dhe * tcv * go * (1.f - go);
    out[3][e] = d * gi * (1.f - gg_ * gg_);
    dcn[e] = d * gf;
  }
#pragma unroll
  for (int gg = 0; gg < 4; ++gg) *reinterpret_cast<uint4*>(da + r * 4 * H + gg * H + j) = pack8(out[gg]);
  float4* o = reinterpret_cast<float4*>(dc_carry + r * H + j);
  o[0] = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]); o[1] = make_float4(dcn[4], dcn[5], dcn[6], dcn[7]);
}

// 2-D fp32 -> fp16 (weights, projection table)
__global__ void k_cvt16(__half* __restrict__ dst, int64_t ldd, const float* __restrict__ src, int64_t lds, int64_t rows, int cols) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = cols >> 2;
  if (idx >= rows * c4) return;
  const int64_t r = idx / c4;
  const int c = (int)(idx % c4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
  *reinterpret_cast<uint2*>(dst + r * ldd + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
}

// max |x| -> bits[0] (atomicMax on the IEEE bit pattern of a non-negative float)
__global__ void __launch_bounds__(256) k_amax(const float* __restrict__ x, int64_t n4, uint32_t* __restrict__ bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));
}
// scale = the power of two that brings max|dh| to [2^9, 2^10): 6 binades of head-room below the fp16 maximum for the
// growth of da through the recurrence, 24 binades (normal + subnormal) below for the small gradients
__global__ void k_pick_scale(const uint32_t* __restrict__ bits, float* __restrict__ out2) {
  const float amax = __uint_as_float(bits[0]);
  float s = 1.f;
  if (amax > 0.f && isfinite(amax)) {
    int e; frexpf(amax, &e);                         // amax = f * 2^e, f in [0.5, 1)
    int k = 10 - e;
    k = max(-60, min(60, k));
    s = ldexpf(1.f, k);
  }
  out2[0] = s; out2[1] = 1.f / s;
}

// segmented row sum over fp16 rows (see k_segsum_rows in pointwise.cu): out[tok,:] += inv_scale * sum of X[perm[p],:]
constexpr int SEG16_ROWS = 64;
__global__ void __launch_bounds__(256)
k_segsum16(const __half* __restrict__ X, int64_t ldx, const int32_t* __restrict__ perm, const int32_t* __restrict__ sorted_tok,
           int64_t n, float* __restrict__ out, int ncols, const float* __restrict__ inv_scale) {
  const int64_t p0 = (int64_t)blockIdx.x * SEG16_ROWS, p1 = min(n, p0 + SEG16_ROWS);
  const float sc = inv_scale ? __ldg(inv_scale) : 1.f;
  for (int c0 = threadIdx.x * 8; c0 < ncols; c0 += blockDim.x * 8) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    int cur = sorted_tok[p0];
    auto flush = [&](int tok) {
      float* o = out + (int64_t)tok * ncols + c0;
      red_add_v4(o, acc[0] * sc, acc[1] * sc, acc[2] * sc, acc[3] * sc);
      red_add_v4(o + 4, acc[4] * sc, acc[5] * sc, acc[6] * sc, acc[7] * sc);
    };
    for (int64_t p = p0; p < p1; ++p) {
      const int tok = sorted_tok[p];
      if (tok != cur) {
        flush(cur);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        cur = tok;
      }
      float v[8];
      unpack8(__ldcs(reinterpret_cast<const uint4*>(X + (int64_t)perm[p] * ldx + c0)), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
    flush(cur);
  }
}

// ------------------------------------------------------------------------------------------------ host
static CUtensorMap make_tmap16(const void* base, CUtensorMapDataType dt, int esize, int64_t rows, int64_t cols, int64_t ld,
                               int box_rows, int box_cols, CUtensorMapSwizzle swz) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&tm, dt, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (lstm16) failed (%d): base %p rows %lld cols %lld ld %lld box %dx%d", (int)r, base,
             (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
    throw CudaError(-3, buf);
  }
  return tm;
}
static CUtensorMap tmap_h(const __half* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
                          CUtensorMapSwizzle swz) {
  return make_tmap16(base, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, rows, cols, ld, box_rows, box_cols, swz);
}
static CUtensorMap tmap_f(const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
                          CUtensorMapSwizzle swz) {
  return make_tmap16(base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, cols, ld, box_rows, box_cols, swz);
}

template <int MODE>
static void launch16(LaunchCtx& cx, const CUtensorMap& tA, const CUtensorMap& tB, const Lstm16Maps& em, const Lstm16Params& p,
                     int num_tiles) {
  using C = Cfg16<MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA_CHECK(cudaFuncSetAttribute(k_lstm16<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    attr_set = true;
  }
  // persistent, balanced waves (see gemm_tc.cu::launch)
  const int pmax = cx.sms() / 2;
  int pairs = num_tiles;
  if (num_tiles > pmax) { const int rounds = cdiv(num_tiles, pmax); pairs = cdiv(num_tiles, rounds); }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::TOTAL;
  cfg.stream = cx.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  VD_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_lstm16<MODE>, tA, tB, em, p));
  check_launch(cx, "k_lstm16");
}

}  // namespace tc

bool lstm16_shape_ok(int64_t R, int H) { return H % 256 == 0 && H <= 512 && R >= 1024; }   // the forward kernel keeps the 4H bias in 8 KB of shared memory

void lstm16_step_fwd(LaunchCtx& cx, int64_t R, int H, const __half* h_prev16, const __half* Wh16, const __half* ptable16,
                     const int32_t* tok, const float* bias, const float* c_prev, const int32_t* mask_ids, __half* gates16,
                     float* c_out, __half* h16_out, float* h32_out) {
  using namespace tc;
  VD_REQUIRE(lstm16_shape_ok(R, H), VD_E_STATE, "lstm16_step_fwd: shape");
  Lstm16Params p = {};
  p.R = (int)R; p.H = H; p.ptable = ptable16; p.tok = tok; p.bias = bias; p.c_prev = c_prev; p.mask_ids = mask_ids;
  p.h32_out = h32_out; p.save_gates = gates16 != nullptr;
  CUtensorMap tA = tmap_h(h_prev16, R, H, H, BM16, BK16, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tB = tmap_h(Wh16, 4 * (int64_t)H, H, H, 64, BK16, CU_TENSOR_MAP_SWIZZLE_128B);
  Lstm16Maps em;
  em.g16 = gates16 ? tmap_h(gates16, R, 4 * (int64_t)H, 4 * (int64_t)H, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B)
                   : tmap_h(h16_out, R, H, H, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B);
  em.c = tmap_f(c_out, R, H, H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
  em.h16 = tmap_h(h16_out, R, H, H, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B);
  launch16<0>(cx, tA, tB, em, p, cdiv(R, 2 * BM16) * (H / 64));
}

void lstm16_step_bwd(LaunchCtx& cx, int64_t R, int H, const __half* da_next16, const __half* Whb16, const __half* gates16,
                     const float* c_prev, const float* c_cur, float* dc_carry, const int32_t* mask_ids, __half* da16) {
  using namespace tc;
  VD_REQUIRE(lstm16_shape_ok(R, H), VD_E_STATE, "lstm16_step_bwd: shape");
  Lstm16Params p = {};
  p.R = (int)R; p.H = H; p.gsave = gates16; p.c_prev = c_prev; p.c_cur = c_cur; p.dc_carry = dc_carry; p.mask_ids = mask_ids;
  CUtensorMap tA = tmap_h(da_next16, R, 4 * (int64_t)H, 4 * (int64_t)H, BM16, BK16, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tB = tmap_h(Whb16, H, 4 * (int64_t)H, 4 * (int64_t)H, 128, BK16, CU_TENSOR_MAP_SWIZZLE_128B);
  Lstm16Maps em;
  em.g16 = tmap_h(da16, R, 4 * (int64_t)H, 4 * (int64_t)H, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B);
  em.c = tmap_f(dc_carry, R, H, H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
  em.h16 = em.c;
  launch16<1>(cx, tA, tB, em, p, cdiv(R, 2 * BM16) * (H / 256));
}

void lstm16_first_step(LaunchCtx& cx, int64_t R, int H, const __half* ptable16, const int32_t* tok, const float* bias,
                       const float* c_prev, const int32_t* mask_ids, __half* gates16, float* c_out, __half* h16_out, float* h32_out) {
  const int64_t n = R * (H / 8);
  tc::k_lstm16_first<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(ptable16, tok, bias, c_prev, mask_ids, gates16, c_out,
                                                                          h16_out, h32_out, R, H);
  check_launch(cx, "k_lstm16_first");
}

void lstm16_bwd_last(LaunchCtx& cx, int64_t R, int H, const __half* gates16, const float* c_prev, const float* c_cur,
                     const float* dh_last, const float* scale, const int32_t* mask_ids, float* dc_carry, __half* da16) {
  const int64_t n = R * (H / 8);
  tc::k_lstm16_bwd_last<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(gates16, c_prev, c_cur, dh_last, scale, mask_ids,
                                                                             dc_carry, da16, R, H);
  check_launch(cx, "k_lstm16_bwd_last");
}

void cvt_f32_to_f16(LaunchCtx& cx, __half* dst, int64_t ldd, const float* src, int64_t lds, int64_t rows, int cols) {
  VD_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, VD_E_STATE, "cvt_f32_to_f16: alignment");
  const int64_t n = rows * (cols / 4);
  if (n == 0) return;
  tc::k_cvt16<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(dst, ldd, src, lds, rows, cols);
  check_launch(cx, "k_cvt16");
}

// scale2[0] = power-of-two scale for max|x| -> [2^9, 2^10), scale2[1] = its inverse; bits = 1 scratch word
void pick_grad_scale(LaunchCtx& cx, const float* x, int64_t n, uint32_t* bits, float* scale2) {
  VD_REQUIRE(n % 4 == 0, VD_E_STATE, "pick_grad_scale: n % 4");
  VD_CUDA_CHECK(cudaMemsetAsync(bits, 0, sizeof(uint32_t), cx.stream));
  const int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256, 4 * 148);
  tc::k_amax<<<blocks, 256, 0, cx.stream>>>(x, n / 4, bits);
  check_launch(cx, "k_amax");
  tc::k_pick_scale<<<1, 1, 0, cx.stream>>>(bits, scale2);
  check_launch(cx, "k_pick_scale");
}

void segsum_rows16(LaunchCtx& cx, const __half* X, int64_t ldx, const int32_t* perm, const int32_t* sorted_tok, int64_t n,
                   float* out, int ncols, const float* inv_scale) {
  VD_REQUIRE(ncols % 8 == 0 && ldx % 8 == 0, VD_E_STATE, "segsum_rows16: alignment");
  if (n == 0) return;
  tc::k_segsum16<<<(unsigned)((n + tc::SEG16_ROWS - 1) / tc::SEG16_ROWS), 256, 0, cx.stream>>>(X, ldx, perm, sorted_tok, n, out, ncols,
                                                                                               inv_scale);
  check_launch(cx, "k_segsum16");
}

// C[M,N] += inv_scale * A^T B, A (K x M) and B (K x N) fp16 row-major
void gemm_atb16(LaunchCtx& cx, int M, int N, int64_t K, const __half* A, int64_t lda, const __half* B, int64_t ldb, float* C,
                int64_t ldc, const float* inv_scale) {
  using namespace tc;
  VD_REQUIRE(M % 64 == 0 && N % 64 == 0 && K >= A16_KB && lda % 8 == 0 && ldb % 8 == 0, VD_E_STATE, "gemm_atb16: shape");
  const int tiles = cdiv(M, BM16) * cdiv(N, A16_BN);
  const int64_t cta_cap = cx.sm_budget > 0 ? cx.sm_budget : 2LL * cx.sm_count;
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(cta_cap / tiles, K / (A16_KB * 4)));
  int64_t kps = ((K + splits - 1) / splits + A16_KB - 1) / A16_KB * A16_KB;
  splits = (K + kps - 1) / kps;
  Atb16Params p = {M, N, K, kps, C, ldc, inv_scale};
  CUtensorMap tA = tmap_h(A, K, M, lda, A16_KB, 64, CU_TENSOR_MAP_SWIZZLE_128B),
              tB = tmap_h(B, K, N, ldb, A16_KB, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA_CHECK(cudaFuncSetAttribute(k_atb16, cudaFuncAttributeMaxDynamicSharedMemorySize, A16_TOTAL));
    attr_set = true;
  }
  dim3 grid(tiles, (unsigned)splits);
  k_atb16<<<grid, A16_THREADS, A16_TOTAL, cx.stream>>>(tA, tB, p);
  check_launch(cx, "k_atb16");
}

}  // namespace vd
