// Dense contractions on the 5th-gen tensor cores: tcgen05.mma kind::tf32 (fp32 operands in shared memory,
// TF32 multiply, fp32 accumulate in TMEM), operands staged by TMA (128B-swizzled tiles), one persistent CTA
// per SM, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2-9 =
// epilogue (TMEM -> registers -> fused pointwise -> global).  Double-buffered TMEM accumulators let the
// epilogue of tile i overlap the main loop of tile i+1.
//
//   MODE_GENERIC : C = act(beta*C + bias + A B^T)                          (nn.Linear and friends)
//   MODE_LSTM_FWD: gates = A_h Wh^T (+ x-projection) -> SeqLSTM pointwise   (one launch per time step)
//   MODE_LSTM_BWD: dh = da_{t+1} Wh -> SeqLSTM backward pointwise -> da_t   (one launch per time step)
//   k_tc_atb     : C += A^T B with both operands MN-major                   (weight gradients, split-K)
#include <cuda.h>
#include <mutex>
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vd {
namespace tc {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int BK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction
constexpr int EPI_WARPS = 8;     // 2 warps per TMEM lane quarter, each takes half of the tile's columns
// threads per CTA = 64 (TMA producer warp + MMA issuer warp) + 32 * epilogue warps (template parameter EW)

enum { MODE_GENERIC = 0, MODE_LSTM_FWD = 1, MODE_LSTM_BWD = 2, MODE_LSE = 3, MODE_DLOGIT = 4 };
// MODE_LSE    : vocabulary projection whose (rows, V) logits never leave the chip: per row and column slice only the running
//               max, the sum of exponentials and the target's logit are written (gen.lua:23-24 + the criterion of model.lua:33-36
//               / utils.computeLhood, utils.lua:86-102)
// MODE_DLOGIT : the same contraction recomputed in the backward pass with the epilogue  C = keep * (exp(x - lse[row]) - onehot)

struct Params {
  int M, N, K;                    // GEMM sizes (N = output columns; LSTM_FWD: N = 4H, LSTM_BWD: N = H)
  int H;
  // generic
  float* C; int64_t ldc; float beta; const float* bias; int act;
  // lstm fwd
  float* gates;                   // (R,4H): in = x-projection pre-activations (if has_xproj), out = activated gates
  int has_xproj;
  const float* ptable; const int32_t* tok;   // optional gathered x-projection: ptable[tok[r], 4H]
  const float* c_prev; float* c_out; float* h_out; const int32_t* mask_ids;
  // lstm bwd
  const float* gsave; const float* c_cur; const float* dh_ext; float* dc_carry; float* da;
  // fused vocabulary softmax (MODE_LSE / MODE_DLOGIT): 1-based target class per row (0 = none), maskzero ids per row
  const int32_t* tgt; const int32_t* row_ids; const float* lse;
  float* part_max; float* part_sum; float* tgt_logit; int nparts;      // (M, nparts) partials, nparts = column tiles * slices
};

__device__ __forceinline__ void ld8(const float* p, float* d) {
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(p)), x1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
  d[0] = x0.x; d[1] = x0.y; d[2] = x0.z; d[3] = x0.w; d[4] = x1.x; d[5] = x1.y; d[6] = x1.z; d[7] = x1.w;
}
__device__ __forceinline__ void add8(const float* p, float* d) {
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(p)), x1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
  d[0] += x0.x; d[1] += x0.y; d[2] += x0.z; d[3] += x0.w; d[4] += x1.x; d[5] += x1.y; d[6] += x1.z; d[7] += x1.w;
}
__device__ __forceinline__ void st8(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
// streaming store: saved activations are not read again before the backward pass — keep them out of L2's way
__device__ __forceinline__ void st8_cs(float* p, const float* v) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
}

// Epilogue staging (CTA-pair LSTM kernels): per epilogue warp, ARR arrays of [32 rows][16 floats].  The 16-byte
// chunks of a row are XOR-swizzled by (row>>1)&3 so that both access patterns are bank-conflict free: "thread =
// row" (the TMEM side) and "4 lanes = one row's 64 bytes" (the global side, coalesced 64-byte runs instead of one
// 16-byte piece of 32 different lines per instruction, which is what saturated L1TEX before).
constexpr int STG_ARR_BYTES = 32 * 16 * 4;
template <int BN, int MODE, int CG, int EW = EPI_WARPS> struct StageCfg {
  static constexpr bool ON = MODE == MODE_GENERIC || MODE == MODE_DLOGIT || (CG == 2 && (BN == 256 || (BN == 128 && MODE == MODE_LSTM_BWD)));
  static constexpr int ARR = !ON ? 0 : ((MODE == MODE_GENERIC || MODE == MODE_DLOGIT) ? 1 : MODE == MODE_LSTM_FWD ? 6 : 7);
  static constexpr int BYTES = EW * ARR * STG_ARR_BYTES;
};
template <int BN, int CG = 1, int STG_BYTES = 0, int MAXST = 16> struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 4;        // 16 KB (this CTA's 128 rows)
  static constexpr int B_BYTES = (BN / CG) * BK * 4; // a CTA pair splits the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int WANT = (B_BYTES == 32768) ? 4 : (B_BYTES == 16384) ? 6 : (B_BYTES == 8192) ? 8 : 10;   // small tiles are latency-bound: deeper
  static constexpr int FIT = (232448 - 1024 - 256 - STG_BYTES) / STAGE_BYTES;
  static constexpr int STAGES0 = WANT < FIT ? WANT : FIT;
  static constexpr int STAGES = STAGES0 < MAXST ? STAGES0 : MAXST;   // MAXST = 4: the half-size CTAs that share an SM
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
};

__device__ __forceinline__ float4* stg_at(float* stg, int arr, int row, int q) {
  return reinterpret_cast<float4*>(stg + (arr * 32 + row) * 16 + ((q ^ ((row >> 1) & 3)) << 2));
}
// global -> staging: `mine` = this lane's row base (16 floats) or nullptr (zeros); coalesced 4 lanes per row.
// cp.async (LDGSTS): no register staging, so every array of a group is in flight at once and the global latency is
// paid once per group instead of once per array; the caller ends the group with stg_load_wait().
__device__ __forceinline__ void stg_load(float* stg, int arr, const float* mine, int lane) {
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int row = ps * 8 + (lane >> 2), q = lane & 3;
    const float* src = shfl_ptr(mine, row);
    float4* dst = stg_at(stg, arr, row, q);
    if (src) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src + q * 4) : "memory");
    else *dst = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// staging -> global through the TMA (one bulk tensor store per array instead of 4 passes of shuffles + LDS + STG per
// lane): the staging arrays are laid out exactly like a SWIZZLE_64B box of 32 rows x 16 floats.
struct EpiMaps { CUtensorMap g4, c, h; };      // [R,4H] gates / da;  [R,H] c_out / dc_carry;  [R,H] h_out
__device__ __forceinline__ void stg_load_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
}
template <bool STREAM>
__device__ __forceinline__ void stg_store(float* stg, int arr, float* mine, int lane) {
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int row = ps * 8 + (lane >> 2), q = lane & 3;
    float* dst = const_cast<float*>(shfl_ptr(mine, row));
    if (dst) {
      const float4 v = *stg_at(stg, arr, row, q);
      if (STREAM) __stcs(reinterpret_cast<float4*>(dst) + q, v); else reinterpret_cast<float4*>(dst)[q] = v;
    }
  }
}
__device__ __forceinline__ void stg_get8(float* stg, int arr, int row, int sub, float* d) {
  const float4 a = *stg_at(stg, arr, row, sub * 2), b = *stg_at(stg, arr, row, sub * 2 + 1);
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void stg_put8(float* stg, int arr, int row, int sub, const float* v) {
  *stg_at(stg, arr, row, sub * 2) = make_float4(v[0], v[1], v[2], v[3]);
  *stg_at(stg, arr, row, sub * 2 + 1) = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------------
// EW = epilogue warps.  8 (320 threads, one CTA per SM) for the SM-filling kernels; 4 (192 threads, 4 pipeline stages,
// TWO CTAs per SM) for the few-row encoder kernels: those are latency-bound, so a CTA that owns a whole SM mostly
// waits — and, run beside the option stream, what they cost is SM time, not FLOPs.
template <int BN, int MODE, int CG, int EW = EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * EW, EW == 4 ? 2 : 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ EpiMaps em, const Params p) {
  using SC = StageCfg<BN, MODE, CG, EW>;
  using L = SmemLayout<BN, CG, SC::BYTES, (EW == 4 ? 4 : 16)>;
  constexpr int NH = EW / 4;                         // epilogue warps per TMEM lane quarter = column slices per tile
  constexpr int STAGES = L::STAGES;
  constexpr int TM = BM * CG;                        // rows per tile: 128, or 256 for a CTA pair
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int cta = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // tile-loop index of this CTA (pair)
  const int ncta = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* stg_all = (float*)(smem + STAGES * L::STAGE_BYTES);
  uint64_t* full = (uint64_t*)(smem + STAGES * L::STAGE_BYTES + SC::BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + TM - 1) / TM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], EW * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CG == 2) cluster_sync_all();                  // peer barriers exist before anyone signals them
  if (warp == 1) { if (CG == 2) tmem_alloc_cg2(tmem_slot, L::TMEM_COLS); else tmem_alloc(tmem_slot, L::TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (every CTA stages its own rows of A and its share of B) =====
      int s = 0; uint32_t ph = 0;
      for (int tile = cta; tile < num_tiles; tile += ncta) {
        const int m0 = (tile / num_n) * TM + (int)rank * BM;
        const int nt = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (CG == 1) {
            mbar_expect_tx(&full[s], L::STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full[s], kb * BK, m0);
            if (MODE == MODE_LSTM_FWD) {
              // interleave the 4 gate blocks of this hidden-unit slice: tile columns = [i | f | o | g]
              constexpr int HB = BN / 4;
#pragma unroll
              for (int g = 0; g < 4; ++g) tma_load_2d(sb + g * HB * BK * 4, &tmB, &full[s], kb * BK, g * p.H + nt * HB);
            } else {
              tma_load_2d(sb, &tmB, &full[s], kb * BK, nt * BN);
            }
          } else {
            // all transactions of the pair complete on the LEADER's full barrier
            const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
            if (leader) mbar_expect_tx(&full[s], 2 * L::STAGE_BYTES);
            tma_load_2d_cg2(sa, &tmA, bar, kb * BK, m0);
            if (MODE == MODE_LSTM_FWD) {
              constexpr int HB = BN / 4;              // leader stages gate blocks [i | f], peer [o | g]
#pragma unroll
              for (int g = 0; g < 2; ++g)
                tma_load_2d_cg2(sb + g * HB * BK * 4, &tmB, bar, kb * BK, ((int)rank * 2 + g) * p.H + nt * HB);
            } else {
              tma_load_2d_cg2(sb, &tmB, bar, kb * BK, nt * BN + (int)rank * (BN / 2));
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && !leader) {
    // peer CTA of a pair: its MMA warp only takes part in TMEM alloc / dealloc
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(TM, BN, 0, 0);
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
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
          const uint64_t adesc = make_desc(sa, 16, 1024);
          const uint64_t bdesc = make_desc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = adesc + (uint64_t)(k * UMMA_K * 4 >> 4), bd = bdesc + (uint64_t)(k * UMMA_K * 4 >> 4);
            if (CG == 2) umma_tf32_cg2(d_tmem, ad, bd, idesc, (kb | k) ? 1u : 0u);
            else umma_tf32(d_tmem, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
          if (CG == 2) { umma_commit_cg2(&empty[s]); if (kb == num_kb - 1) umma_commit_cg2(&tfull[buf]); }
          else { umma_commit(&empty[s]); if (kb == num_kb - 1) umma_commit(&tfull[buf]); }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (num_kb == 0 && lane == 0) {                 // K == 0: nothing to wait for, release the epilogue(s)
        mbar_arrive(&tfull[buf]);
        if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tfull[buf]), 1));
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int it = 0;
    for (int tile = cta; tile < num_tiles; tile += ncta, ++it) {
      const int buf = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      const int m0 = (tile / num_n) * TM + (int)rank * BM;
      const int nt = tile % num_n;
      mbar_wait(&tfull[buf], bph);
      tc_fence_after();
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);

      if (MODE == MODE_LSE) {
        // online softmax statistics of this warp's column slice of the tile: nothing but (max, sum exp, target logit) leaves
        const int n0 = nt * BN;
        const int tcol = row_ok ? p.tgt[row] - 1 : -1;
        float mrun = -INFINITY, srun = 0.f, tl = 0.f;
        bool has = false;
#pragma unroll 1
        for (int c = half * (BN / NH); c < (half + 1) * (BN / NH); c += 8) {
          if (n0 + c >= p.N) break;                 // warp-uniform
          float v[8];
          tmem_ld8(taddr + c, v);
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n = n0 + c + j;
            const bool ok = n < p.N;
            const float x = ok ? v[j] + (p.bias ? __ldg(p.bias + n) : 0.f) : -INFINITY;
            v[j] = x;
            if (ok && n == tcol) { tl = x; has = true; }
            mx = fmaxf(mx, x);
          }
          const float mnew = fmaxf(mrun, mx);
          float add = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) add += __expf(v[j] - mnew);      // exp(-inf) = 0 for the clipped columns
          srun = srun * __expf(mrun - mnew) + add;
          mrun = mnew;
        }
        if (row_ok) {
          const int64_t pi = row * p.nparts + nt * NH + half;
          p.part_max[pi] = mrun; p.part_sum[pi] = srun;
          if (has) p.tgt_logit[row] = tl;
        }
      } else if (MODE == MODE_GENERIC || MODE == MODE_DLOGIT) {
        // 16 output columns at a time through the warp's staging array; C leaves (and, for beta != 0, enters)
        // as 64-byte-swizzled boxes: TMA tensor store / cp.async load.  Rows >= M and columns >= N are clipped.
        const int n0 = nt * BN;
        float* stg = stg_all + (warp - 2) * (SC::ARR * 32 * 16);
#pragma unroll 1
        for (int c = half * (BN / NH); c < (half + 1) * (BN / NH); c += 16) {
          if (n0 + c >= p.N) break;                 // warp-uniform
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          if (p.beta != 0.f) {
            const bool full16 = n0 + c + 16 <= p.N;
            stg_load(stg, 0, (row_ok && full16) ? p.C + row * p.ldc + n0 + c : nullptr, lane);
            stg_load_wait();
            if (!full16 && row_ok) {               // ragged last group: element-wise
#pragma unroll
              for (int sub = 0; sub < 2; ++sub) {
                float t[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { const int n = n0 + c + sub * 8 + j; t[j] = n < p.N ? p.C[row * p.ldc + n] : 0.f; }
                stg_put8(stg, 0, lane, sub, t);
              }
            }
          }
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            float v[8], old[8];
            tmem_ld8(taddr + c + sub * 8, v);
            tmem_ld_wait();
            if (p.beta != 0.f) stg_get8(stg, 0, lane, sub, old);
            if (MODE == MODE_DLOGIT) {
              // d loss / d logit of the sum criterion: softmax - onehot on kept rows (input token != pad, target != pad)
              const int tg = row_ok ? p.tgt[row] : 0;
              const bool keepr = row_ok && tg > 0 && p.row_ids[row] != 0;
              const float l = keepr ? p.lse[row] : 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int n = n0 + c + sub * 8 + j;
                const float x = v[j] + ((p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f);
                v[j] = keepr ? __expf(x - l) - (n == tg - 1 ? 1.f : 0.f) : 0.f;
              }
            } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int n = n0 + c + sub * 8 + j;
              float x = v[j];
              if (p.bias && n < p.N) x += __ldg(p.bias + n);
              if (p.beta != 0.f) x += p.beta * old[j];
              v[j] = p.act == 1 ? ftanh(x) : x;
            }
            }
            stg_put8(stg, 0, lane, sub, v);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&em.g4, stg, n0 + c, m0 + q * 32); bulk_commit(); }
          __syncwarp();
        }
      } else if (MODE == MODE_LSTM_FWD) {
        constexpr int HB = BN / 4;                  // hidden units per tile
        const int H = p.H, j0 = nt * HB;
        const bool masked = row_ok && p.mask_ids && p.mask_ids[row] == 0;
        const float* prow = (row_ok && p.ptable) ? p.ptable + (int64_t)p.tok[row] * 4 * H : nullptr;
        float* grow = p.gates ? p.gates + row * 4 * H : nullptr;
        if constexpr (SC::ON) {
          // staged epilogue: 16 hidden units at a time through the warp's swizzled staging tile
          float* stg = stg_all + (warp - 2) * (SC::ARR * 32 * 16);
#pragma unroll 1
          for (int c = half * (HB / NH); c < (half + 1) * (HB / NH); c += 16) {
            const int j = j0 + c;
            if (lane == 0) bulk_wait_read0();    // the previous group's TMA stores have finished reading the staging tile
            __syncwarp();
            // phase 1: coalesced global -> staging (x-projection rows of the 4 gates, previous cell)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float* src = nullptr;
              if (row_ok) src = p.has_xproj ? grow + g * H + j : (prow ? prow + g * H + j : nullptr);
              stg_load(stg, g, src, lane);
            }
            stg_load(stg, 4, (row_ok && p.c_prev) ? p.c_prev + row * H + j : nullptr, lane);
            stg_load_wait();
            // phase 2: thread = row; TMEM accumulators + staged inputs -> gates, c, h back into the staging tile
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              float a[4][8], cn[8], hn[8], cp[8];
#pragma unroll
              for (int g = 0; g < 4; ++g) tmem_ld8(taddr + g * HB + c + sub * 8, a[g]);
              tmem_ld_wait();
              stg_get8(stg, 4, lane, sub, cp);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float x[8];
                stg_get8(stg, g, lane, sub, x);
                if (p.bias) add8(p.bias + g * H + j + sub * 8, x);     // null when folded into the x-projection
#pragma unroll
                for (int e = 0; e < 8; ++e) a[g][e] += x[e];
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float gi = fsigmoid(a[0][e]), gf = fsigmoid(a[1][e]), go = fsigmoid(a[2][e]), gg = ftanh(a[3][e]);
                const float c_ = gf * cp[e] + gi * gg;
                const float keep = masked ? 0.f : 1.f;
                a[0][e] = gi * keep; a[1][e] = gf * keep; a[2][e] = go * keep; a[3][e] = gg * keep;
                cn[e] = c_ * keep; hn[e] = go * ftanh(c_) * keep;
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) stg_put8(stg, g, lane, sub, a[g]);
              stg_put8(stg, 4, lane, sub, cn);
              stg_put8(stg, 5, lane, sub, hn);
            }
            // phase 3: staging -> global by TMA tensor stores (rows beyond R are clipped by the tensor map)
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int r0 = m0 + q * 32;
              if (p.gates) {
#pragma unroll
                for (int g = 0; g < 4; ++g) tma_store_2d(&em.g4, stg + g * 512, g * H + j, r0);
              }
              tma_store_2d(&em.c, stg + 4 * 512, j, r0);
              tma_store_2d(&em.h, stg + 5 * 512, j, r0);
              bulk_commit();
            }
            __syncwarp();
          }
        } else {
#pragma unroll 1
        for (int c = half * (HB / NH); c < (half + 1) * (HB / NH); c += 8) {
          const int j = j0 + c;
          float a[4][8], x[4][8], cp[8];
          // issue the global loads first so that they overlap the TMEM read
          if (row_ok && !masked) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (p.bias) ld8(p.bias + g * H + j, x[g]);
              else {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[g][e] = 0.f;
              }
              if (p.has_xproj) add8(grow + g * H + j, x[g]);
              if (prow) add8(prow + g * H + j, x[g]);
            }
            if (p.c_prev) ld8(p.c_prev + row * H + j, cp);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) cp[e] = 0.f;
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld8(taddr + g * HB + c, a[g]);
          tmem_ld_wait();
          if (row_ok) {
            float cn[8], hn[8];
            if (masked) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { a[0][e] = a[1][e] = a[2][e] = a[3][e] = 0.f; cn[e] = 0.f; hn[e] = 0.f; }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float zi = (p.K ? a[0][e] : 0.f) + x[0][e], zf = (p.K ? a[1][e] : 0.f) + x[1][e];
                const float zo = (p.K ? a[2][e] : 0.f) + x[2][e], zg = (p.K ? a[3][e] : 0.f) + x[3][e];
                const float gi = fsigmoid(zi), gf = fsigmoid(zf), go = fsigmoid(zo), gg = ftanh(zg);
                const float c_ = gf * cp[e] + gi * gg;
                a[0][e] = gi; a[1][e] = gf; a[2][e] = go; a[3][e] = gg;
                cn[e] = c_; hn[e] = go * ftanh(c_);
              }
            }
            if (grow) {
#pragma unroll
              for (int g = 0; g < 4; ++g) st8_cs(grow + g * H + j, a[g]);
            }
            st8(p.c_out + row * H + j, cn);
            st8(p.h_out + row * H + j, hn);
          }
          __syncwarp();
        }
        }   // !SC::ON
      } else {   // MODE_LSTM_BWD: accumulator = dh_rec for hidden units [nt*BN, nt*BN + BN)
        const int H = p.H, j0 = nt * BN;
        const bool masked = row_ok && p.mask_ids && p.mask_ids[row] == 0;
        if constexpr (SC::ON) {
          float* stg = stg_all + (warp - 2) * (SC::ARR * 32 * 16);
#pragma unroll 1
          for (int c = half * (BN / NH); c < (half + 1) * (BN / NH); c += 16) {
            const int j = j0 + c;
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            // phase 1: saved gates (4), c_prev, c_t, dc carry -> staging, coalesced
#pragma unroll
            for (int g = 0; g < 4; ++g) stg_load(stg, g, row_ok ? p.gsave + row * 4 * H + g * H + j : nullptr, lane);
            stg_load(stg, 4, (row_ok && p.c_prev) ? p.c_prev + row * H + j : nullptr, lane);
            stg_load(stg, 5, row_ok ? p.c_cur + row * H + j : nullptr, lane);
            stg_load(stg, 6, row_ok ? p.dc_carry + row * H + j : nullptr, lane);
            stg_load_wait();
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              float dh[8], g[4][8], cp[8], cc[8], dc[8], out[4][8], dcn[8];
              tmem_ld8(taddr + c + sub * 8, dh);
              tmem_ld_wait();
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) stg_get8(stg, gg, lane, sub, g[gg]);
              stg_get8(stg, 4, lane, sub, cp);
              stg_get8(stg, 5, lane, sub, cc);
              stg_get8(stg, 6, lane, sub, dc);
              if (p.dh_ext && row_ok) {          // only the last time step has an external gradient here
                float ex[8];
                ld8(p.dh_ext + row * H + j + sub * 8, ex);
#pragma unroll
                for (int e = 0; e < 8; ++e) dh[e] += ex[e];
              }
              const float keep = masked ? 0.f : 1.f;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float gi = g[0][e], gf = g[1][e], go = g[2][e], gg_ = g[3][e];
                const float tcv = ftanh(cc[e]);
                const float d = (dc[e] + dh[e] * go * (1.f - tcv * tcv)) * keep;
                const float dhe = dh[e] * keep;
                out[0][e] = d * gg_ * gi * (1.f - gi);
                out[1][e] = d * cp[e] * gf * (1.f - gf);
                out[2][e] = dhe * tcv * go * (1.f - go);
                out[3][e] = d * gi * (1.f - gg_ * gg_);
                dcn[e] = d * gf;
              }
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) stg_put8(stg, gg, lane, sub, out[gg]);
              stg_put8(stg, 6, lane, sub, dcn);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int r0 = m0 + q * 32;
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) tma_store_2d(&em.g4, stg + gg * 512, gg * H + j, r0);
              tma_store_2d(&em.c, stg + 6 * 512, j, r0);
              bulk_commit();
            }
            __syncwarp();
          }
        } else {
#pragma unroll 1
        for (int c = half * (BN / NH); c < (half + 1) * (BN / NH); c += 8) {
          const int j = j0 + c;
          float dh[8], g[4][8], cp[8], cc[8], dc[8], ex[8];
          if (row_ok && !masked) {
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) ld8(p.gsave + row * 4 * H + gg * H + j, g[gg]);
            if (p.c_prev) ld8(p.c_prev + row * H + j, cp);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) cp[e] = 0.f;
            }
            ld8(p.c_cur + row * H + j, cc);
            ld8(p.dc_carry + row * H + j, dc);
            if (p.dh_ext) ld8(p.dh_ext + row * H + j, ex);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) ex[e] = 0.f;
            }
          }
          tmem_ld8(taddr + c, dh);
          tmem_ld_wait();
          if (row_ok) {
            float out[4][8], dcn[8];
            if (masked) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { out[0][e] = out[1][e] = out[2][e] = out[3][e] = 0.f; dcn[e] = 0.f; }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float dhe = (p.K ? dh[e] : 0.f) + ex[e];
                const float gi = g[0][e], gf = g[1][e], go = g[2][e], gg_ = g[3][e];
                const float tcv = ftanh(cc[e]);
                const float d = dc[e] + dhe * go * (1.f - tcv * tcv);
                out[0][e] = d * gg_ * gi * (1.f - gi);
                out[1][e] = d * cp[e] * gf * (1.f - gf);
                out[2][e] = dhe * tcv * go * (1.f - go);
                out[3][e] = d * gi * (1.f - gg_ * gg_);
                dcn[e] = d * gf;
              }
            }
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) st8(p.da + row * 4 * H + gg * H + j, out[gg]);
            st8(p.dc_carry + row * H + j, dcn);
          }
          __syncwarp();
        }
        }   // !SC::ON
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                // the accumulator buffer may be overwritten by the (leader's) MMAs
        if (CG == 2 && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
        else mbar_arrive(&tempty[buf]);
      }
    }
    if (SC::ON && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all TMA stores performed
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody tears down while the peer still signals / reads
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, L::TMEM_COLS); else tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradients: C[M,N] += sum_k A[k,m] B[k,n].  Both operands are "MN-major" (the contraction index k is
// the row index of the activations in HBM): a TMA box of 32 columns x KB rows lands in shared memory as KB rows
// of 128 bytes; with the 128B_ATOM_32B swizzle this is the canonical MN-major SWIZZLE_128B_BASE32B UMMA layout, the
// only MN-major layout tf32 operands may use (4 k-rows per 512-byte atom, SBO = 512; one K=8 MMA consumes two
// atoms; 32-column groups are LBO bytes apart).  One (tile, K-split) per CTA; the fp32 partial sums are reduced
// into C with vector red.global.add.
constexpr int ATB_KB = 32;       // k-rows per pipeline stage (4 MMAs)
constexpr int ATB_BN = 256;
constexpr int ATB_THREADS = 192;
struct AtbParams { int M, N; int64_t K, k_per_split; float* C; int64_t ldc; };
struct AtbSmem {
  static constexpr int A_BYTES = BM * ATB_KB * 4;         // 4 boxes of [32 rows x 128 B]
  static constexpr int B_BYTES = ATB_BN * ATB_KB * 4;     // 8 boxes
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 4;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

__global__ void __launch_bounds__(ATB_THREADS, 1)
k_tc_atb(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const AtbParams p) {
  using L = AtbSmem;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = (p.N + ATB_BN - 1) / ATB_BN;
  const int m0 = (blockIdx.x / num_n) * BM, n0 = (blockIdx.x % num_n) * ATB_BN;
  const int64_t kbeg = (int64_t)blockIdx.y * p.k_per_split;
  const int64_t kend = min(p.K, kbeg + p.k_per_split);
  const int num_kb = (int)((kend - kbeg + ATB_KB - 1) / ATB_KB);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, ATB_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_expect_tx(&full[s], L::STAGE_BYTES);
        const int krow = (int)(kbeg + (int64_t)kb * ATB_KB);
        // k_per_split is a multiple of ATB_KB, so only the global K tail (zero-filled by TMA) can be partial
#pragma unroll
        for (int g = 0; g < BM / 32; ++g) tma_load_2d(sa + g * ATB_KB * 128, &tmA, &full[s], m0 + g * 32, krow);
#pragma unroll
        for (int g = 0; g < ATB_BN / 32; ++g) tma_load_2d(sb + g * ATB_KB * 128, &tmB, &full[s], n0 + g * 32, krow);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(BM, ATB_BN, 1, 1);
    int s = 0; uint32_t ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < ATB_KB / UMMA_K; ++k) {
          const uint64_t adesc = make_desc(sa + k * 1024, ATB_KB * 128, 512, 1);
          const uint64_t bdesc = make_desc(sb + k * 1024, ATB_KB * 128, 512, 1);
          umma_tf32(tmem_base, adesc, bdesc, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        if (kb == num_kb - 1) umma_commit(tfull);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (num_kb > 0) {
    const int q = warp & 3;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int m = m0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < ATB_BN; c += 8) {
      if (n0 + c >= p.N) break;
      float v[8];
      tmem_ld8(taddr + c, v);
      tmem_ld_wait();
      if (m < p.M) {
        float* crow = p.C + (int64_t)m * p.ldc + n0 + c;
        if (n0 + c + 8 <= p.N && ((p.ldc & 3) == 0)) {
          red_add_v4(crow, v[0], v[1], v[2], v[3]);
          red_add_v4(crow + 4, v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + c + j < p.N) atomicAdd(crow + j, v[j]);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, ATB_BN); }
}

// ------------------------------------------------------------------------------------------------ host
PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  VD_REQUIRE(fn != nullptr, -3, "cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

// 2-D fp32 tensor map: `rows` x `cols` (cols contiguous), row pitch ld floats, box = box_rows x box_cols floats
static CUtensorMap make_tmap(const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols = BK,
                             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): base %p rows %lld cols %lld ld %lld box %dx%d", (int)r, (void*)base,
             (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
    throw CudaError(-3, buf);
  }
  return tm;
}

static bool tma_ok(const float* p, int64_t ld) { return ((uintptr_t)p % 16 == 0) && (ld % 4 == 0); }

template <int BN, int MODE, int CG = 1, int EW = EPI_WARPS>
static void launch(LaunchCtx& cx, const CUtensorMap& tA, const CUtensorMap& tB, const Params& p, int num_tiles,
                   const EpiMaps* epi = nullptr) {
  using L = SmemLayout<BN, CG, StageCfg<BN, MODE, CG, EW>::BYTES, (EW == 4 ? 4 : 16)>;
  constexpr int THREADS = 64 + 32 * EW;
  static EpiMaps none = {};
  const EpiMaps& em = epi ? *epi : none;
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm<BN, MODE, CG, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    if (EW == 4)      // two of these CTAs per SM: ask for the full shared-memory carve-out
      VD_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm<BN, MODE, CG, EW>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  // Persistent grid, balanced waves: with pmax CTAs (pairs) available the tiles need ceil(tiles/pmax) rounds; the
  // smallest grid that still finishes in that many rounds is used, so no CTA idles through a ragged last round and the
  // SMs not needed stay free for concurrent streams (250 backward tiles: 63 pairs x 4 rounds instead of 74 x 3.4).
  auto balanced = [&](int pmax) {
    if (num_tiles <= pmax) return num_tiles;
    const int rounds = cdiv(num_tiles, pmax);
    return cdiv(num_tiles, rounds);
  };
  if (CG == 1) {
    int grid = balanced(cx.sms());
    k_tc_gemm<BN, MODE, 1, EW><<<grid, THREADS, L::TOTAL, cx.stream>>>(tA, tB, em, p);
  } else {
    // CTA pairs: a 2-CTA cluster per 256-row tile, one pair per TPC (num_tiles counts 256-row tiles here)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * balanced(cx.sms() / 2));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = cx.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    VD_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_tc_gemm<BN, MODE, 2, EW>, tA, tB, em, p));
  }
  check_launch(cx, "k_tc_gemm");
}

static bool small_ew4() {          // VD_SMALL_EW4=0: the few-row kernels as one 320-thread CTA per SM (A/B)
  static int v = -1;
  if (v < 0) { const char* e = getenv("VD_SMALL_EW4"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

static int small_narrow() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VD_SMALL_NARROW"); v = e ? atoi(e) : 0; }
  return v;
}

static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VD_CTA_PAIRS"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

}  // namespace tc

// ---- public entry points ------------------------------------------------------------------------
bool gemm_tn_tc(LaunchCtx& cx, int M, int N, int K, const float* A, int64_t lda, const int32_t* a_gather, const float* B,
                int64_t ldb, float* C, int64_t ldc, float beta, const float* bias, int act) {
  using namespace tc;
  if (a_gather) return false;                               // gathered rows go through the projection table instead
  if (M < 64 || N < 16 || K < 32) return false;             // tiny contractions stay on CUDA cores
  if (!tma_ok(A, lda) || !tma_ok(B, ldb) || !tma_ok(C, ldc)) return false;
  Params p = {};
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.beta = beta; p.bias = bias; p.act = act;
  EpiMaps em = {};
  em.g4 = make_tmap(C, M, N, ldc, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);      // C leaves through TMA tensor stores
  const int tiles256 = cdiv(M, BM) * cdiv(N, 256);
  if (N > 128 && tiles256 >= cx.sm_count / 2) {
    CUtensorMap tA = make_tmap(A, M, K, lda, BM), tB = make_tmap(B, N, K, ldb, 256);
    launch<256, MODE_GENERIC>(cx, tA, tB, p, tiles256, &em);
  } else if (N > 64 && cdiv(M, BM) * cdiv(N, 128) >= cx.sm_count / 2) {
    CUtensorMap tA = make_tmap(A, M, K, lda, BM), tB = make_tmap(B, N, K, ldb, 128);
    launch<128, MODE_GENERIC>(cx, tA, tB, p, cdiv(M, BM) * cdiv(N, 128), &em);
  } else {
    CUtensorMap tA = make_tmap(A, M, K, lda, BM), tB = make_tmap(B, N, K, ldb, 64);
    if (small_ew4()) launch<64, MODE_GENERIC, 1, 4>(cx, tA, tB, p, cdiv(M, BM) * cdiv(N, 64), &em);
    else launch<64, MODE_GENERIC>(cx, tA, tB, p, cdiv(M, BM) * cdiv(N, 64), &em);
  }
  return true;
}

bool gemm_atb_tc(LaunchCtx& cx, int M, int N, int64_t K, const float* A, int64_t lda, const int32_t* a_gather, const float* B,
                 int64_t ldb, float* C, int64_t ldc) {
  using namespace tc;
  if (a_gather) return false;
  if (M < 32 || N < 32 || K < 64) return false;
  if (!tma_ok(A, lda) || !tma_ok(B, ldb)) return false;
  const int tiles = cdiv(M, BM) * cdiv(N, ATB_BN);
  // whole waves: the largest split count with tiles * splits <= 2 * #SM (one CTA per SM at this smem footprint)
  // (under an SM budget — the option stream sharing the GPU with the encoder's chains — a single wave of budget CTAs, so
  // that the reserved SMs really stay free: a second wave would be placed on them)
  const int64_t cta_cap = cx.sm_budget > 0 ? cx.sm_budget : 2LL * cx.sm_count;
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(cta_cap / tiles, K / (ATB_KB * 4)));
  int64_t kps = ((K + splits - 1) / splits + ATB_KB - 1) / ATB_KB * ATB_KB;
  splits = (K + kps - 1) / kps;
  AtbParams p = {M, N, K, kps, C, ldc};
  CUtensorMap tA = make_tmap(A, K, M, lda, ATB_KB, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B),
              tB = make_tmap(B, K, N, ldb, ATB_KB, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  static bool attr_set = false;
  if (!attr_set) {
    VD_CUDA_CHECK(cudaFuncSetAttribute(k_tc_atb, cudaFuncAttributeMaxDynamicSharedMemorySize, AtbSmem::TOTAL));
    attr_set = true;
  }
  dim3 grid(tiles, (unsigned)splits);
  k_tc_atb<<<grid, ATB_THREADS, AtbSmem::TOTAL, cx.stream>>>(tA, tB, p);
  check_launch(cx, "k_tc_atb");
  return true;
}

// Vocabulary projection with the softmax statistics fused into the epilogue (no (rows, V) tensor in HBM):
//   part_max / part_sum (M, nparts): per column slice running max and sum of exp(x - max);  tgt_logit[m] = x[m, tgt[m]-1]
int vocab_lse_nparts(int N) { return cdiv(N, 256) * (tc::EPI_WARPS / 4); }
bool vocab_lse_tc(LaunchCtx& cx, int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias,
                  const int32_t* tgt, float* part_max, float* part_sum, float* tgt_logit) {
  using namespace tc;
  if (M < 64 || N < 256 || K < 32) return false;
  if (!tma_ok(A, lda) || !tma_ok(B, ldb)) return false;
  Params p = {};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.tgt = tgt; p.part_max = part_max; p.part_sum = part_sum; p.tgt_logit = tgt_logit;
  p.nparts = vocab_lse_nparts(N);
  CUtensorMap tA = make_tmap(A, M, K, lda, BM), tB = make_tmap(B, N, K, ldb, 256);
  launch<256, MODE_LSE>(cx, tA, tB, p, cdiv(M, BM) * cdiv(N, 256));
  return true;
}
// C[m,n] = keep[m] * (exp(A B^T + bias - lse[m]) - [n == tgt[m]-1])   (backward of log-softmax + ClassNLL, recomputed)
bool vocab_dlogits_tc(LaunchCtx& cx, int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias,
                      const int32_t* tgt, const int32_t* row_ids, const float* lse, float* C, int64_t ldc) {
  using namespace tc;
  if (M < 64 || N < 256 || K < 32) return false;
  if (!tma_ok(A, lda) || !tma_ok(B, ldb) || !tma_ok(C, ldc)) return false;
  Params p = {};
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias; p.tgt = tgt; p.row_ids = row_ids; p.lse = lse;
  EpiMaps em = {};
  em.g4 = make_tmap(C, M, N, ldc, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
  CUtensorMap tA = make_tmap(A, M, K, lda, BM), tB = make_tmap(B, N, K, ldb, 256);
  launch<256, MODE_DLOGIT>(cx, tA, tB, p, cdiv(M, BM) * cdiv(N, 256), &em);
  return true;
}

// One SeqLSTM forward step on the tensor cores: gates = h_prev Wh^T (+ xproj | + ptable[tok]) + bias, then the
// pointwise half, fused.  WtS_h = transposed shadow weight offset to the h columns: [4H, ld] with K = H.
bool lstm_step_fwd_tc(LaunchCtx& cx, int64_t R, int H, const float* h_prev, const float* WtS_h, int64_t ldw, const float* bias,
                      float* gates, int has_xproj, const float* ptable, const int32_t* tok, const float* c_prev, float* c_out,
                      float* h_out, const int32_t* mask_ids) {
  using namespace tc;
  if (H % 64 != 0 || R < 1) return false;
  if (!tma_ok(WtS_h, ldw) || (h_prev && !tma_ok(h_prev, H))) return false;
  Params p = {};
  p.M = (int)R; p.N = 4 * H; p.K = h_prev ? H : 0; p.H = H;      // K == 0: no recurrent term (t = 0 without h0)
  if (!h_prev) h_prev = WtS_h;                                      // any valid address for the (unused) tensor map
  p.bias = bias; p.gates = gates; p.has_xproj = has_xproj; p.ptable = ptable; p.tok = tok;
  p.c_prev = c_prev; p.c_out = c_out; p.h_out = h_out; p.mask_ids = mask_ids;
  CUtensorMap tA = make_tmap(h_prev, p.K ? R : 128, H, H, BM);
  const int tiles_big = cdiv(R, BM) * (H / 64);
  if (tiles_big >= cx.sm_count) {            // 64 hidden units (x 4 gates = 256 columns) per tile
    CUtensorMap tB = make_tmap(WtS_h, 4 * (int64_t)H, H, ldw, 64);
    if (use_cta_pairs() && p.K > 0) {
      EpiMaps em;                              // epilogue outputs leave through TMA tensor stores (64B-swizzled boxes)
      em.g4 = make_tmap(gates ? gates : c_out, R, gates ? 4 * (int64_t)H : H, gates ? 4 * (int64_t)H : H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
      em.c = make_tmap(c_out, R, H, H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
      em.h = make_tmap(h_out, R, H, H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
      launch<256, MODE_LSTM_FWD, 2>(cx, tA, tB, p, cdiv(R, 2 * BM) * (H / 64), &em);
    }
    else launch<256, MODE_LSTM_FWD>(cx, tA, tB, p, tiles_big);
  } else if (small_narrow() == 2) {          // experiment: few fat CTAs (64 hidden units per tile)
    CUtensorMap tB = make_tmap(WtS_h, 4 * (int64_t)H, H, ldw, 64);
    launch<256, MODE_LSTM_FWD>(cx, tA, tB, p, tiles_big);
  } else {                                   // few rows (encoder LSTMs): 16 hidden units per tile, 4x the CTAs
    CUtensorMap tB = make_tmap(WtS_h, 4 * (int64_t)H, H, ldw, 16);
    if (small_ew4()) launch<64, MODE_LSTM_FWD, 1, 4>(cx, tA, tB, p, cdiv(R, BM) * (H / 16));
    else launch<64, MODE_LSTM_FWD>(cx, tA, tB, p, cdiv(R, BM) * (H / 16));
  }
  return true;
}

// One SeqLSTM backward step: dh_rec = da_next Wh (Wh rows = reference layout rows D.., [H, 4H]) fused with the
// backward pointwise half producing da_t and the cell-gradient carry.
bool lstm_step_bwd_tc(LaunchCtx& cx, int64_t R, int H, const float* da_next, const float* Wh, const float* gsave,
                      const float* c_prev, const float* c_cur, const float* dh_ext, float* dc_carry, const int32_t* mask_ids,
                      float* da) {
  using namespace tc;
  if (H % 128 != 0 || R < 1) return false;
  if (!tma_ok(Wh, 4 * H) || (da_next && !tma_ok(da_next, 4 * H))) return false;
  Params p = {};
  p.M = (int)R; p.N = H; p.K = da_next ? 4 * H : 0; p.H = H;       // K == 0: last time step, no recurrent gradient
  if (!da_next) da_next = Wh;
  p.gsave = gsave; p.c_prev = c_prev; p.c_cur = c_cur; p.dh_ext = dh_ext;
  p.dc_carry = dc_carry; p.mask_ids = mask_ids; p.da = da;
  CUtensorMap tA = make_tmap(da_next, p.K ? R : 128, 4 * (int64_t)H, 4 * (int64_t)H, BM);
  const int tiles_big = cdiv(R, BM) * (H / 128);
  if (tiles_big >= cx.sm_count) {
    CUtensorMap tB = make_tmap(Wh, H, 4 * (int64_t)H, 4 * (int64_t)H, 128);
    static int bwd_bn = -1;
    if (bwd_bn < 0) { const char* e = getenv("VD_BWD_BN"); bwd_bn = (e && atoi(e) == 128) ? 128 : 256; }
    if (use_cta_pairs() && p.K > 0 && bwd_bn == 256 && H % 256 == 0) {
      EpiMaps em;
      em.g4 = make_tmap(da, R, 4 * (int64_t)H, 4 * (int64_t)H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
      em.c = make_tmap(dc_carry, R, H, H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
      em.h = em.c;
      launch<256, MODE_LSTM_BWD, 2>(cx, tA, tB, p, cdiv(R, 2 * BM) * (H / 256), &em);
    } else if (use_cta_pairs() && p.K > 0) {
      // 128 hidden units per pair-tile: twice the tiles of the 256-wide variant -> a fuller last wave (250 vs 500
      // tiles over 74 CTA pairs)
      CUtensorMap tB64 = make_tmap(Wh, H, 4 * (int64_t)H, 4 * (int64_t)H, 64);
      EpiMaps em;
      em.g4 = make_tmap(da, R, 4 * (int64_t)H, 4 * (int64_t)H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
      em.c = make_tmap(dc_carry, R, H, H, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B);
      em.h = em.c;
      launch<128, MODE_LSTM_BWD, 2>(cx, tA, tB64, p, cdiv(R, 2 * BM) * (H / 128), &em);
    } else launch<128, MODE_LSTM_BWD>(cx, tA, tB, p, tiles_big);
  } else if (small_narrow() >= 1) {          // experiment: 128 hidden units per tile
    CUtensorMap tB = make_tmap(Wh, H, 4 * (int64_t)H, 4 * (int64_t)H, 128);
    launch<128, MODE_LSTM_BWD>(cx, tA, tB, p, tiles_big);
  } else {
    CUtensorMap tB = make_tmap(Wh, H, 4 * (int64_t)H, 4 * (int64_t)H, 32);
    if (small_ew4()) launch<32, MODE_LSTM_BWD, 1, 4>(cx, tA, tB, p, cdiv(R, BM) * (H / 32));
    else launch<32, MODE_LSTM_BWD>(cx, tA, tB, p, cdiv(R, BM) * (H / 32));
  }
  return true;
}

}  // namespace vd
