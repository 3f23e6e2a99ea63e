// Shared host/device helpers for libmegacrn_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/megacrn_b200.h"
#include "../../include/megacrn_b200_debug.h"

namespace mcrn {

// ---- error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define MCRN_CUDA_OK(expr)                                                                 \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::mcrn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MCRN_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define MCRN_TRY(expr)             \
  do {                             \
    int _s = (expr);               \
    if (_s != MCRN_OK) return _s;  \
  } while (0)

// Every kernel launch of the library goes through this so launches are counted and
// launch errors are surfaced as status codes (never exceptions).
#define MCRN_LAUNCH(kernel, grid, block, smem, stream, ...)                                \
  do {                                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                            \
    ::mcrn::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t _e = cudaPeekAtLastError();                                                \
    if (_e != cudaSuccess) {                                                               \
      ::mcrn::set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, #kernel,     \
                        cudaGetErrorString(_e));                                           \
      return MCRN_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

// Launch of a kernel of the recurrent chain (fused AGCN forward / backward, step glue) with programmatic dependent launch:
// the next kernel of the chain is scheduled as soon as every CTA of this one has passed `griddepcontrol.launch_dependents`,
// runs its prologue (barrier init, TMEM allocation, tensor-map prefetch) while this one drains, and blocks in
// `griddepcontrol.wait` until this grid has completed and flushed.  Works eagerly and under stream capture (the edge
// becomes a programmatic graph edge).  g_pdl_chain = 0 (MCRN_PDL_CHAIN=0 / mcrn_set_option("pdl", 0)): plain launches.
extern int g_pdl_chain;     // bit 0: fused forward, bit 1: fused backward, bit 2: step glue; bit 3: trigger the dependents late (before the epilogue)
template <class... KArgs, class... Args>
int launch_chain(int pdl_bit, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const char* name, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (g_pdl_chain & pdl_bit) ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (le != cudaSuccess) {
    set_error("launch of %s failed: %s", name, cudaGetErrorString(le));
    return MCRN_ERR_CUDA;
  }
  return MCRN_OK;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- model geometry ---------------------------------------------------------------
struct Geo {
  int B, N, T_in, T_out, Cin, Cout, Ycov, H, D, M, d, cheb_k;
  int KS;       // number of non-identity supports: 2*(cheb_k-1)
  int NB;       // 1 + KS blocks in every XP buffer
  int Cdec;     // decoder input channels: Cout + Ycov
  int ldS;      // leading dimension of a support matrix row (multiple of 4 floats)
  int64_t R;    // rows of the node-major state: N*B
  int L;        // num_layers (stacked cells per encoder / decoder, model/MegaCRN.py:62-63)
};

static inline int support_ld(int n) { return (n + 3) / 4 * 4; }

int make_geo(const mcrn_dims* dm, Geo* g);

}  // namespace mcrn
