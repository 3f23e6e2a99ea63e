// Host side of the tcgen05 engine: TMA tensor-map encoding (driver entry point resolved at run
// time, so the library has no link-time dependency on libcuda) and the eligibility rules.
#include "gemm_tc.cuh"

#include <cudaTypedefs.h>

#include <stdlib.h>

#include <mutex>

namespace mcrn {
namespace tc {

float* g_dbg = nullptr;
int g_pdl = getenv("MCRN_PDL") ? atoi(getenv("MCRN_PDL")) : 0;
// 3 = hi/lo as 2*NBX K-segments on 128x128 tiles, 2 CTAs/SM: measured fastest at C2 (7.68 ms/step vs 7.86 / 7.94 / 8.07
// for the A-sharing variants 1 / 2 / 0): these GEMMs are bound by per-CTA latency, not by staged bytes.
int g_hilo_cfg = getenv("MCRN_HILO_CFG") ? atoi(getenv("MCRN_HILO_CFG")) : 3;
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_once;

static void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  else cudaGetLastError();
}

int encode_tensor_map(CUtensorMap* out, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                      const uint32_t box[4], bool mn_major) {
  std::call_once(g_once, resolve_encode);
  if (!g_encode) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return MCRN_ERR_CUDA; }
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] box=[%u,%u,%u,%u] base=%p",
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
              (unsigned long long)strides_bytes[2], box[0], box[1], box[2], box[3], (const void*)base);
    return MCRN_ERR_CUDA;
  }
  return MCRN_OK;
}

}  // namespace tc
namespace fusedh {
// fp16 operands of the fused AGCN kernel: K-major, 128-byte swizzle.
int encode_tensor_map_h(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                        const uint32_t box[4]) {
  std::call_once(tc::g_once, tc::resolve_encode);
  if (!tc::g_encode) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return MCRN_ERR_CUDA; }
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = tc::g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, bx, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(fp16) failed (%d): dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] box=[%u,%u,%u,%u] base=%p",
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
              (unsigned long long)strides_bytes[2], box[0], box[1], box[2], box[3], base);
    return MCRN_ERR_CUDA;
  }
  return MCRN_OK;
}
}  // namespace fusedh
namespace tc {

static bool ok_stride(int64_t elems) { return elems > 0 && (elems % 4) == 0 && elems * 4 < ((int64_t)1 << 40); }

bool eligible(const GemmDesc& g) {
  if (g.prec_exact) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) return false;
  if (g.M < 1 || g.N < 1 || g.Kseg < 1 || g.nseg > 16) return false;
  if (g.b_sub == 2 && !(g.a_k == 1 && g.b_n == 1 && g.b_seg && !g.use_map && g.splits == 1)) return false;
  if (g.b_sub > 2) return false;
  // A: exactly one of (K contiguous, M contiguous); the other stride 16-byte aligned
  if (g.a_k == 1) { if (!ok_stride(g.a_row)) return false; }
  else if (g.a_row == 1) { if (!ok_stride(g.a_k)) return false; }
  else return false;
  if (g.b_k == 1) { if (!ok_stride(g.b_n)) return false; }
  else if (g.b_n == 1) { if (!ok_stride(g.b_k)) return false; }
  else return false;
  if (g.a_seg && g.nseg_a() > 1 && !ok_stride(g.a_seg)) return false;
  if (g.b_seg && g.nseg_b() > 1 && !ok_stride(g.b_seg)) return false;
  if (g.a_batch && g.nbatch > 1 && !ok_stride(g.a_batch)) return false;
  if (g.b_batch && g.nbatch > 1 && !ok_stride(g.b_batch)) return false;
  return true;
}

}  // namespace tc
}  // namespace mcrn
