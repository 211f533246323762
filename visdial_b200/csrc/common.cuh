// Shared definitions for the visdial_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <stdexcept>
#include <vector>
#include <map>

namespace vd {

struct CudaError : std::runtime_error {
  int code;
  CudaError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define VD_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      char _buf[512];                                                                    \
      snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
               cudaGetErrorString(_e));                                                  \
      throw vd::CudaError(_e == cudaErrorMemoryAllocation ? -5 : -3, _buf);              \
    }                                                                                    \
  } while (0)

#define VD_REQUIRE(cond, code, msg)                                                      \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      char _buf[512];                                                                    \
      snprintf(_buf, sizeof(_buf), "%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);    \
      throw vd::CudaError(code, _buf);                                                   \
    }                                                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Dropout RNG: Philox4x32-10, counter = (q_lo, q_hi, site, iteration), key = (seed_lo, seed_hi),
// q = element_index / 4, word = element_index % 4.  keep <=> word >= thresh (thresh = p * 2^32).
// The numpy twin used by the tests is oracle/philox.py.
// ---------------------------------------------------------------------------------------------
struct DropCfg {
  uint32_t seed_lo, seed_hi, iter, thresh;  // thresh == 0 -> dropout disabled (identity)
  float scale;                              // 1 / (1 - p)
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                       uint32_t c3, uint32_t k0, uint32_t k1,
                                                       uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// multiplicative factor (0 or scale) of element `idx` at dropout site `site`
__device__ __forceinline__ float drop_factor(const DropCfg& d, uint32_t site, uint64_t idx) {
  if (d.thresh == 0) return 1.f;
  uint64_t q = idx >> 2;
  uint32_t o[4];
  philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), site, d.iter, d.seed_lo, d.seed_hi, o);
  return o[idx & 3] >= d.thresh ? d.scale : 0.f;
}
// factors of the 4 elements idx4*4 .. idx4*4+3 with one Philox call
__device__ __forceinline__ void drop_factor4(const DropCfg& d, uint32_t site, uint64_t idx4, float f[4]) {
  if (d.thresh == 0) { f[0] = f[1] = f[2] = f[3] = 1.f; return; }
  uint32_t o[4];
  philox4x32_10((uint32_t)idx4, (uint32_t)(idx4 >> 32), site, d.iter, d.seed_lo, d.seed_hi, o);
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = o[i] >= d.thresh ? d.scale : 0.f;
}

// Dropout sites (mirrors oracle/visdial_oracle.py)
enum : uint32_t {
  SITE_QEMBED = 0, SITE_HEMBED = 1, SITE_HATT = 2, SITE_IMG_TR = 3, SITE_U_OUT = 5,
  SITE_FUSION = 6, SITE_IMG_FC7 = 7, SITE_HOP0 = 16
};

// ---------------------------------------------------------------------------------------------
// Launch context: stream + launch accounting + optional per-class event bracketing.
// ---------------------------------------------------------------------------------------------
struct KStat {
  int64_t launches = 0;
  double flops = 0, bytes = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  double ms = 0;
};

struct LaunchCtx {
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  int profiling = 0;          // 0 off, 1 every class, 2 only the roofline class (big lstm_step / lstm_step_bwd launches)
  int sm_count = 148;
  int sm_budget = 0;          // > 0: persistent / single-wave kernels size their grids to this many SMs (the rest stay free
                              // for a concurrent higher-priority stream); 0 = all of them
  int sms() const { return sm_budget > 0 ? sm_budget : sm_count; }
  std::map<std::string, KStat> stats;
  std::vector<cudaEvent_t> free_events;

  cudaEvent_t get_event() {
    if (!free_events.empty()) { cudaEvent_t e = free_events.back(); free_events.pop_back(); return e; }
    cudaEvent_t e; VD_CUDA_CHECK(cudaEventCreate(&e)); return e;
  }
  // bracket one launch (or a short group) of class `name`
  struct Scope {
    LaunchCtx* cx; KStat* st; cudaEvent_t a = nullptr, b = nullptr;
    Scope(LaunchCtx* c, const char* name, double flops, double bytes) : cx(c), st(nullptr) {
      if (!cx->profiling) return;
      if (cx->profiling == 2 && strcmp(name, "lstm_step") != 0 && strcmp(name, "lstm_step_bwd") != 0) return;
      st = &cx->stats[name];
      st->launches++; st->flops += flops; st->bytes += bytes;
      a = cx->get_event(); b = cx->get_event();
      cudaEventRecord(a, cx->stream);
    }
    ~Scope() {
      if (!st) return;
      cudaEventRecord(b, cx->stream);
      st->pending.emplace_back(a, b);
    }
  };
  void collect() {
    for (auto& kv : stats) {
      for (auto& p : kv.second.pending) {
        cudaEventSynchronize(p.second);
        float ms = 0; cudaEventElapsedTime(&ms, p.first, p.second);
        kv.second.ms += ms;
        free_events.push_back(p.first); free_events.push_back(p.second);
      }
      kv.second.pending.clear();
    }
  }
};

inline void check_launch(LaunchCtx& cx, const char* what) {
  cx.launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof(buf), "kernel launch failed: %s -> %s", what, cudaGetErrorString(e));
    throw CudaError(-3, buf);
  }
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace vd
