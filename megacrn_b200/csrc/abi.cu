// extern "C" surface of libmegacrn_b200.so (see include/megacrn_b200.h).
#include <mutex>
#include <unordered_map>

#include "engine.cuh"
#include "agcn_fused.cuh"
#include "plan.cuh"

namespace mcrn {
int g_engine = 0;
extern int g_simt_mask;
extern int g_fused, g_fused_parts, g_bwd_fused;
const char* last_error();
int forward_impl(const Geo& g, const Plan& p, const mcrn_params* prm, const float* x, const float* y_cov,
                 const float* labels, const uint8_t* tf, float* output, float* h_att, float* query, float* pos,
                 float* neg, float* ws, cudaStream_t st, bool reuse_prologue);
int backward_impl(const Geo& g, const Plan& p, const mcrn_params* prm, const uint8_t* tf, const float* d_output,
                  const float* d_hatt, const float* d_query, const float* d_pos, const float* d_neg,
                  const mcrn_params* grads, float* ws, cudaStream_t st);
int forward_impl_layers(const Geo& g, const Plan& p, const mcrn_params* prm, const mcrn_layer_params* up, const float* x,
                        const float* y_cov, const float* labels, const uint8_t* tf, float* output, float* h_att, float* query,
                        float* pos, float* neg, float* ws, cudaStream_t st);
int backward_impl_layers(const Geo& g, const Plan& p, const mcrn_params* prm, const mcrn_layer_params* up, const uint8_t* tf,
                         const float* d_output, const float* d_hatt, const float* d_query, const float* d_pos,
                         const float* d_neg, const mcrn_params* grads, const mcrn_layer_params* ugrads, float* ws,
                         cudaStream_t st);
int supports_forward_entry(const Geo& g, const Plan& p, float* ws, const float* mem, const float* we1,
                           const float* we2, float* S, float* Sr, cudaStream_t st);
int trainer_loss_impl(const Geo& g, const float* output, const float* labels, const float* query, const float* pos,
                      const float* neg, float mean, float std, float lamb, float lamb1, const float* mask_count, float* loss_out,
                      float* d_output, float* d_query, float* scratch, cudaStream_t st);
int mask_count_impl(const float* labels, int64_t n, float mean, float std, float* count_out, cudaStream_t st);

int adam_step_impl(const Geo& g, const mcrn_params* prm, const mcrn_params* grads, const mcrn_params* m, const mcrn_params* v,
                   float* state, float beta1, float beta2, float eps, float max_norm, cudaStream_t st);

int adam_step_layers_impl(const Geo& g, const mcrn_params* const q[4], const mcrn_layer_params* const u[4], float* state,
                          float beta1, float beta2, float eps, float max_norm, cudaStream_t st);
bool set_option(const char* name, int value);
int probe_mn16_entry(const void* A, const void* B, float* C, unsigned lbo, unsigned sbo, unsigned layout, unsigned kstep,
                     unsigned b_major, cudaStream_t st);
static std::atomic<uint64_t> g_mode_epoch{0};
static std::mutex g_mu;
static std::unordered_map<const void*, uint64_t> g_saved;   // workspace -> dims hash of the last saving forward

static uint64_t dims_hash(const mcrn_dims* d) {
  const int32_t* v = reinterpret_cast<const int32_t*>(d);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof(mcrn_dims) / sizeof(int32_t); ++i) h = (h ^ (uint64_t)(uint32_t)v[i]) * 1099511628211ull;
  return h;
}

// One device per process (one process per GPU is the deployment model): the side streams, events and
// the per-kernel shared-memory attributes of the library are created for the device of the first call.  A later call with
// another current device fails loudly instead of launching on streams of the wrong device.
static std::atomic<int> g_bound_device{-1};
static int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(e)); cudaGetLastError(); return MCRN_ERR_NO_DEVICE; }
  int expected = -1;
  if (!g_bound_device.compare_exchange_strong(expected, dev) && expected != dev) {
    set_error("libmegacrn_b200 is bound to device %d by its first call; the current device is %d (one process per GPU)", expected, dev);
    return MCRN_ERR_STATE;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess || major != 10) {
    set_error("device %d has compute capability major %d; libmegacrn_b200 is built for sm_100a only", dev, major);
    cudaGetLastError();
    return MCRN_ERR_NO_DEVICE;
  }
  return MCRN_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int check_params(const mcrn_params* p, const char* what) {
  if (!p) { set_error("%s is null", what); return MCRN_ERR_BAD_POINTER; }
  const float* const* v = reinterpret_cast<const float* const*>(p);
  for (size_t i = 0; i < sizeof(mcrn_params) / sizeof(float*); ++i) {
    if (!v[i]) { set_error("%s: tensor #%zu is null", what, i); return MCRN_ERR_BAD_POINTER; }
    if (!aligned16(v[i]) && i != 13) { /* proj_b may be a 4-byte tensor; alignment still comes from the allocator */ }
  }
  return MCRN_OK;
}
static int check_layer_params(const mcrn_layer_params* up, int layers, const char* what) {
  if (layers <= 1) return MCRN_OK;
  if (!up) { set_error("%s is null but num_layers = %d", what, layers); return MCRN_ERR_BAD_POINTER; }
  for (int l = 0; l + 1 < layers; ++l) {
    const float* const* v = reinterpret_cast<const float* const*>(up + l);
    for (size_t i = 0; i < sizeof(mcrn_layer_params) / sizeof(float*); ++i)
      if (!v[i]) { set_error("%s[%d]: tensor #%zu is null", what, l, i); return MCRN_ERR_BAD_POINTER; }
  }
  return MCRN_OK;
}
static int single_layer_only(const Geo& g, const char* what) {
  if (g.L != 1) { set_error("%s: num_layers = %d needs the *_layers entry (mcrn_forward_layers / mcrn_backward_layers)", what, g.L); return MCRN_ERR_BAD_DIMS; }
  return MCRN_OK;
}
}  // namespace mcrn

using namespace mcrn;

extern "C" {

int mcrn_abi_version(void) { return MCRN_ABI_VERSION; }
const char* mcrn_last_error(void) { return mcrn::last_error(); }
int mcrn_device_ok(void) { return check_device(); }
uint64_t mcrn_launch_count(void) { return g_launches.load(); }
int mcrn_set_engine(int engine) {
  if (engine < 0 || engine > 2) { set_error("engine must be 0, 1 or 2"); return MCRN_ERR_BAD_DIMS; }
  g_engine = engine;
  g_mode_epoch.fetch_add(1);
  return MCRN_OK;
}
int mcrn_get_engine(void) { return g_engine; }
int mcrn_set_debug_mask(int mask) { g_simt_mask = mask; g_mode_epoch.fetch_add(1); return MCRN_OK; }
uint64_t mcrn_mode_epoch(void) { return g_mode_epoch.load(); }
int mcrn_debug_probe_mn16(const void* A, const void* B, float* C, unsigned lbo, unsigned sbo, unsigned layout, unsigned kstep,
                          unsigned b_major, void* stream) {
  if (!A || !B || !C) { set_error("mcrn_debug_probe_mn16: null pointer"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  return probe_mn16_entry(A, B, C, lbo, sbo, layout, kstep, b_major, static_cast<cudaStream_t>(stream));
}
int mcrn_set_option(const char* name, int value) {
  if (!set_option(name, value)) { set_error("mcrn_set_option: unknown option '%s'", name ? name : "(null)"); return MCRN_ERR_BAD_DIMS; }
  g_mode_epoch.fetch_add(1);
  return MCRN_OK;
}
int mcrn_debug_fused_timeline(long long* device_slots, int which) {
  fused::g_dbg_timeline = device_slots; fused::g_dbg_which = which; fused::g_dbg_count = 0;
  return MCRN_OK;
}
int mcrn_debug_launch_spans(unsigned long long* device_slots, int max_launches) {
  fused::g_dbg_span = device_slots; fused::g_dbg_span_n = 0; fused::g_dbg_span_cap = device_slots ? max_launches : 0;
  return MCRN_OK;
}
int mcrn_adam_step(const mcrn_dims* dims, const mcrn_params* params, const mcrn_params* grads, const mcrn_params* exp_avg,
                   const mcrn_params* exp_avg_sq, float* dev_state, float beta1, float beta2, float eps, float max_grad_norm,
                   void* stream) {
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  MCRN_TRY(single_layer_only(g, "mcrn_adam_step"));
  MCRN_TRY(check_params(params, "params"));
  MCRN_TRY(check_params(grads, "grads"));
  MCRN_TRY(check_params(exp_avg, "exp_avg"));
  MCRN_TRY(check_params(exp_avg_sq, "exp_avg_sq"));
  if (!dev_state) { set_error("mcrn_adam_step: dev_state is null"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  return adam_step_impl(g, params, grads, exp_avg, exp_avg_sq, dev_state, beta1, beta2, eps, max_grad_norm,
                        static_cast<cudaStream_t>(stream));
}
int mcrn_adam_step_layers(const mcrn_dims* dims, const mcrn_params* params, const mcrn_layer_params* upper,
                          const mcrn_params* grads, const mcrn_layer_params* upper_grads, const mcrn_params* exp_avg,
                          const mcrn_layer_params* upper_exp_avg, const mcrn_params* exp_avg_sq,
                          const mcrn_layer_params* upper_exp_avg_sq, float* dev_state, float beta1, float beta2, float eps,
                          float max_grad_norm, void* stream) {
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  if (g.L == 1) return mcrn_adam_step(dims, params, grads, exp_avg, exp_avg_sq, dev_state, beta1, beta2, eps, max_grad_norm, stream);
  const mcrn_params* const q[4] = {params, grads, exp_avg, exp_avg_sq};
  const mcrn_layer_params* const u[4] = {upper, upper_grads, upper_exp_avg, upper_exp_avg_sq};
  const char* names[4] = {"params", "grads", "exp_avg", "exp_avg_sq"};
  for (int a = 0; a < 4; ++a) {
    MCRN_TRY(check_params(q[a], names[a]));
    MCRN_TRY(check_layer_params(u[a], g.L, names[a]));
  }
  if (!dev_state) { set_error("mcrn_adam_step_layers: dev_state is null"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  return adam_step_layers_impl(g, q, u, dev_state, beta1, beta2, eps, max_grad_norm, static_cast<cudaStream_t>(stream));
}
int mcrn_kernel_timing(int enable) {
  fused::g_prof.enabled = enable ? 1 : 0;
  if (enable) fused::g_prof.count = 0;
  return MCRN_OK;
}
int mcrn_kernel_timing_read(int kernel_class, float* ms_total, int* launches) {
  if (!ms_total || !launches) { set_error("mcrn_kernel_timing_read: null output"); return MCRN_ERR_BAD_POINTER; }
  *ms_total = 0.f; *launches = 0;
  for (int i = 0; i < fused::g_prof.count; ++i) {
    if (fused::g_prof.cls[i] != kernel_class) continue;
    float ms = 0.f;
    if (cudaEventSynchronize(fused::g_prof.ev[i][1]) != cudaSuccess || cudaEventElapsedTime(&ms, fused::g_prof.ev[i][0], fused::g_prof.ev[i][1]) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    *ms_total += ms; *launches += 1;
  }
  return MCRN_OK;
}
int mcrn_set_bwd_fused(int fused) {
  if (fused < 0 || fused > 2) { set_error("mcrn_set_bwd_fused: 0, 1 or 2"); return MCRN_ERR_BAD_DIMS; }
  g_bwd_fused = fused;
  g_mode_epoch.fetch_add(1);
  return MCRN_OK;
}
int mcrn_set_fused(int fused, int weight_parts) {
  if (fused < 0 || fused > 2 || weight_parts < 1 || weight_parts > 2) { set_error("mcrn_set_fused: fused in {0,1,2}, weight_parts in {1,2}"); return MCRN_ERR_BAD_DIMS; }
  g_fused = fused; g_fused_parts = weight_parts;
  g_mode_epoch.fetch_add(1);
  return MCRN_OK;
}
int mcrn_support_ld(int n) { return support_ld(n); }

size_t mcrn_workspace_bytes(const mcrn_dims* dims, uint32_t flags) {
  Geo g;
  if (make_geo(dims, &g) != MCRN_OK) return 0;
  Plan p;
  make_plan(g, (flags & MCRN_FWD_SAVE_FOR_BACKWARD) != 0, &p);
  return p.bytes;
}

int mcrn_forward(const mcrn_dims* dims, const mcrn_params* params, const float* x, const float* y_cov,
                 const float* labels, const uint8_t* teacher_forcing, float* output, float* h_att, float* query,
                 float* pos, float* neg, void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  MCRN_TRY(single_layer_only(g, "mcrn_forward"));
  MCRN_TRY(check_params(params, "params"));
  if (!x || !y_cov || !output || !h_att || !query || !pos || !neg || !workspace) {
    set_error("mcrn_forward: null tensor pointer");
    return MCRN_ERR_BAD_POINTER;
  }
  bool any_tf = false;
  if (teacher_forcing)
    for (int t = 0; t < g.T_out; ++t) any_tf |= teacher_forcing[t] != 0;
  if (any_tf && !labels) { set_error("teacher forcing requested but labels is null"); return MCRN_ERR_BAD_POINTER; }
  if (!aligned16(workspace)) { set_error("workspace must be 16-byte aligned"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  Plan p;
  bool save = (flags & MCRN_FWD_SAVE_FOR_BACKWARD) != 0;
  make_plan(g, save, &p);
  if (workspace_bytes < p.bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, p.bytes);
    return MCRN_ERR_WORKSPACE;
  }
  int s = forward_impl(g, p, params, x, y_cov, labels, teacher_forcing, output, h_att, query, pos, neg,
                       static_cast<float*>(workspace), static_cast<cudaStream_t>(stream),
                       (flags & MCRN_FWD_REUSE_PROLOGUE) != 0);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (s == MCRN_OK && save) g_saved[workspace] = dims_hash(dims);
    else g_saved.erase(workspace);
  }
  return s;
}

int mcrn_backward(const mcrn_dims* dims, const mcrn_params* params, const float* x, const float* y_cov,
                  const float* labels, const uint8_t* teacher_forcing, const float* d_output, const float* d_h_att,
                  const float* d_query, const float* d_pos, const float* d_neg, const mcrn_params* grads,
                  void* workspace, size_t workspace_bytes, void* stream) {
  (void)x; (void)y_cov; (void)labels;
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  MCRN_TRY(single_layer_only(g, "mcrn_backward"));
  MCRN_TRY(check_params(params, "params"));
  MCRN_TRY(check_params(grads, "grads"));
  if (!workspace) { set_error("workspace is null"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  Plan p;
  make_plan(g, true, &p);
  if (workspace_bytes < p.bytes) { set_error("workspace too small: %zu < %zu", workspace_bytes, p.bytes); return MCRN_ERR_WORKSPACE; }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_saved.find(workspace);
    if (it == g_saved.end() || it->second != dims_hash(dims)) {
      set_error("mcrn_backward: no matching mcrn_forward(MCRN_FWD_SAVE_FOR_BACKWARD) on this workspace");
      return MCRN_ERR_STATE;
    }
  }
  const int s = backward_impl(g, p, params, teacher_forcing, d_output, d_h_att, d_query, d_pos, d_neg, grads,
                              static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
  {
    // the backward consumes the saved state (accumulators are not idempotent): a second mcrn_backward on the same
    // forward is an error, as it is for freed autograd buffers
    std::lock_guard<std::mutex> lk(g_mu);
    g_saved.erase(workspace);
  }
  return s;
}

int mcrn_forward_layers(const mcrn_dims* dims, const mcrn_params* params, const mcrn_layer_params* upper, const float* x,
                        const float* y_cov, const float* labels, const uint8_t* teacher_forcing, float* output, float* h_att,
                        float* query, float* pos, float* neg, void* workspace, size_t workspace_bytes, uint32_t flags,
                        void* stream) {
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  if (g.L == 1)
    return mcrn_forward(dims, params, x, y_cov, labels, teacher_forcing, output, h_att, query, pos, neg, workspace,
                        workspace_bytes, flags, stream);
  MCRN_TRY(check_params(params, "params"));
  MCRN_TRY(check_layer_params(upper, g.L, "upper"));
  if (!x || !y_cov || !output || !h_att || !query || !pos || !neg || !workspace) {
    set_error("mcrn_forward_layers: null tensor pointer");
    return MCRN_ERR_BAD_POINTER;
  }
  bool any_tf = false;
  if (teacher_forcing)
    for (int t = 0; t < g.T_out; ++t) any_tf |= teacher_forcing[t] != 0;
  if (any_tf && !labels) { set_error("teacher forcing requested but labels is null"); return MCRN_ERR_BAD_POINTER; }
  if (!aligned16(workspace)) { set_error("workspace must be 16-byte aligned"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  Plan p;
  const bool save = (flags & MCRN_FWD_SAVE_FOR_BACKWARD) != 0;
  make_plan(g, save, &p);
  if (workspace_bytes < p.bytes) { set_error("workspace too small: %zu < %zu", workspace_bytes, p.bytes); return MCRN_ERR_WORKSPACE; }
  const int s = forward_impl_layers(g, p, params, upper, x, y_cov, labels, teacher_forcing, output, h_att, query, pos, neg,
                                    static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (s == MCRN_OK && save) g_saved[workspace] = dims_hash(dims);
    else g_saved.erase(workspace);
  }
  return s;
}

int mcrn_backward_layers(const mcrn_dims* dims, const mcrn_params* params, const mcrn_layer_params* upper, const float* x,
                         const float* y_cov, const float* labels, const uint8_t* teacher_forcing, const float* d_output,
                         const float* d_h_att, const float* d_query, const float* d_pos, const float* d_neg,
                         const mcrn_params* grads, const mcrn_layer_params* upper_grads, void* workspace,
                         size_t workspace_bytes, void* stream) {
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  if (g.L == 1)
    return mcrn_backward(dims, params, x, y_cov, labels, teacher_forcing, d_output, d_h_att, d_query, d_pos, d_neg, grads,
                         workspace, workspace_bytes, stream);
  MCRN_TRY(check_params(params, "params"));
  MCRN_TRY(check_params(grads, "grads"));
  MCRN_TRY(check_layer_params(upper, g.L, "upper"));
  MCRN_TRY(check_layer_params(upper_grads, g.L, "upper_grads"));
  if (!workspace) { set_error("workspace is null"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  Plan p;
  make_plan(g, true, &p);
  if (workspace_bytes < p.bytes) { set_error("workspace too small: %zu < %zu", workspace_bytes, p.bytes); return MCRN_ERR_WORKSPACE; }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_saved.find(workspace);
    if (it == g_saved.end() || it->second != dims_hash(dims)) {
      set_error("mcrn_backward_layers: no matching mcrn_forward_layers(MCRN_FWD_SAVE_FOR_BACKWARD) on this workspace");
      return MCRN_ERR_STATE;
    }
  }
  const int s = backward_impl_layers(g, p, params, upper, teacher_forcing, d_output, d_h_att, d_query, d_pos, d_neg, grads,
                                     upper_grads, static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_saved.erase(workspace);
  }
  return s;
}

int mcrn_supports_fwd(const mcrn_dims* dims, const float* memory, const float* we1, const float* we2,
                      float* supports_out, void* workspace, size_t workspace_bytes, void* stream) {
  return mcrn_supports_fwd2(dims, memory, we1, we2, supports_out, nullptr, workspace, workspace_bytes, stream);
}

int mcrn_supports_fwd2(const mcrn_dims* dims, const float* memory, const float* we1, const float* we2,
                       float* supports_out, float* supports_tc_out, void* workspace, size_t workspace_bytes,
                       void* stream) {
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  if (!memory || !we1 || !we2 || !supports_out || !workspace) { set_error("null pointer"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  Plan p;
  make_plan(g, false, &p);
  if (workspace_bytes < p.bytes) { set_error("workspace too small"); return MCRN_ERR_WORKSPACE; }
  return supports_forward_entry(g, p, static_cast<float*>(workspace), memory, we1, we2, supports_out,
                                supports_tc_out, static_cast<cudaStream_t>(stream));
}

int mcrn_gemm(int M, int N, int K, const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b,
              float* C, int ldc, int engine, void* stream) {
  if (M < 1 || N < 1 || K < 1 || !A || !B || !C) { set_error("mcrn_gemm: bad arguments"); return MCRN_ERR_BAD_DIMS; }
  MCRN_TRY(check_device());
  GemmDesc q;
  q.A = A; q.M = M; q.N = N; q.Kseg = K;
  if (trans_a) { q.a_row = 1; q.a_k = lda; } else { q.a_row = lda; q.a_k = 1; }
  q.B = B;
  if (trans_b) { q.b_k = 1; q.b_n = ldb; } else { q.b_k = ldb; q.b_n = 1; }
  EpiStore e{C, ldc, 0, 1.0f, nullptr, nullptr};
  int saved = g_engine;
  if (engine) g_engine = engine;
  int s = gemm(q, e, static_cast<cudaStream_t>(stream));
  g_engine = saved;
  return s;
}

// Debug aid (tools/tc_diag.py): one tcgen05 GEMM with the first shared-memory stage dumped to `dbg`.
int mcrn_debug_tc_gemm(int M, int N, int K, const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b,
                       float* C, int ldc, float* dbg, void* stream) {
  tc::g_dbg = dbg;
  int s = mcrn_gemm(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, 2, stream);
  tc::g_dbg = nullptr;
  return s;
}

int mcrn_trainer_loss(const mcrn_dims* dims, const float* output, const float* labels, const float* query,
                      const float* pos, const float* neg, float scaler_mean, float scaler_std, float lamb,
                      float lamb1, float* loss_out, float* d_output, float* d_query, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return mcrn_trainer_loss_dp(dims, output, labels, query, pos, neg, scaler_mean, scaler_std, lamb, lamb1, nullptr, loss_out,
                              d_output, d_query, workspace, workspace_bytes, stream);
}

int mcrn_trainer_loss_dp(const mcrn_dims* dims, const float* output, const float* labels, const float* query,
                         const float* pos, const float* neg, float scaler_mean, float scaler_std, float lamb,
                         float lamb1, const float* mask_count, float* loss_out, float* d_output, float* d_query,
                         void* workspace, size_t workspace_bytes, void* stream) {
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  if (!output || !labels || !query || !pos || !neg || !loss_out || !workspace) { set_error("null pointer"); return MCRN_ERR_BAD_POINTER; }
  if (workspace_bytes < 256) { set_error("mcrn_trainer_loss needs 256 bytes of device scratch"); return MCRN_ERR_WORKSPACE; }
  MCRN_TRY(check_device());
  return trainer_loss_impl(g, output, labels, query, pos, neg, scaler_mean, scaler_std, lamb, lamb1, mask_count, loss_out,
                           d_output, d_query, static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
}

int mcrn_mask_count(const float* labels, int64_t n, float scaler_mean, float scaler_std, float* count_out, void* stream) {
  if (!labels || !count_out || n < 1) { set_error("mcrn_mask_count: bad arguments"); return MCRN_ERR_BAD_POINTER; }
  MCRN_TRY(check_device());
  return mask_count_impl(labels, n, scaler_mean, scaler_std, count_out, static_cast<cudaStream_t>(stream));
}

// ---- host-buffer entries ---------------------------------------------------------------
namespace {
struct HostStage {
  size_t psz[14];        // floats per parameter tensor, mcrn_params field order
  size_t n_x, n_ycov, n_lab, n_out, n_bnd;
  size_t total;          // floats
};
HostStage host_stage(const Geo& g) {
  HostStage h;
  const size_t k2 = 2 * (size_t)g.cheb_k;
  size_t v[14] = {(size_t)g.M * g.d, (size_t)g.H * g.d, (size_t)g.N * g.M, (size_t)g.N * g.M,
                  k2 * (g.Cin + g.H) * 2 * g.H, (size_t)2 * g.H, k2 * (g.Cin + g.H) * g.H, (size_t)g.H,
                  k2 * (g.Cdec + g.D) * 2 * g.D, (size_t)2 * g.D, k2 * (g.Cdec + g.D) * g.D, (size_t)g.D,
                  (size_t)g.Cout * g.D, (size_t)g.Cout};
  h.total = 0;
  for (int i = 0; i < 14; ++i) { h.psz[i] = v[i]; h.total += (v[i] + 63) / 64 * 64; }
  h.n_x = (size_t)g.B * g.T_in * g.N * g.Cin;
  h.n_ycov = (size_t)g.B * g.T_out * g.N * g.Ycov;
  h.n_lab = (size_t)g.B * g.T_out * g.N * g.Cout;
  h.n_out = h.n_lab;
  h.n_bnd = (size_t)g.B * g.N * g.d;
  for (size_t n : {h.n_x, h.n_ycov, h.n_lab, h.n_out, h.n_bnd, h.n_bnd, h.n_bnd, h.n_bnd}) h.total += (n + 63) / 64 * 64;
  return h;
}
}  // namespace

size_t mcrn_host_workspace_bytes(const mcrn_dims* dims, uint32_t flags) {
  Geo g;
  if (make_geo(dims, &g) != MCRN_OK) return 0;
  Plan p;
  make_plan(g, (flags & MCRN_FWD_SAVE_FOR_BACKWARD) != 0, &p);
  return p.bytes + host_stage(g).total * sizeof(float);
}

int mcrn_forward_host(const mcrn_dims* dims, const mcrn_params* host_params, const float* x, const float* y_cov,
                      const float* labels, const uint8_t* teacher_forcing, float* output, float* h_att, float* query,
                      float* pos, float* neg, void* device_workspace, size_t workspace_bytes, uint32_t flags,
                      void* stream) {
  Geo g;
  MCRN_TRY(make_geo(dims, &g));
  MCRN_TRY(single_layer_only(g, "mcrn_forward_host"));
  MCRN_TRY(check_params(host_params, "host_params"));
  if (!x || !y_cov || !output || !h_att || !query || !pos || !neg || !device_workspace) {
    set_error("mcrn_forward_host: null pointer");
    return MCRN_ERR_BAD_POINTER;
  }
  MCRN_TRY(check_device());
  Plan p;
  make_plan(g, (flags & MCRN_FWD_SAVE_FOR_BACKWARD) != 0, &p);
  HostStage h = host_stage(g);
  if (workspace_bytes < p.bytes + h.total * sizeof(float)) { set_error("host workspace too small"); return MCRN_ERR_WORKSPACE; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* base = reinterpret_cast<float*>(static_cast<char*>(device_workspace) + p.bytes);
  size_t off = 0;
  auto take = [&](size_t n) { float* q = base + off; off += (n + 63) / 64 * 64; return q; };
  mcrn_params dp;
  float** dpv = reinterpret_cast<float**>(&dp);
  float* const* hpv = reinterpret_cast<float* const*>(host_params);
  for (int i = 0; i < 14; ++i) {
    dpv[i] = take(h.psz[i]);
    MCRN_CUDA_OK(cudaMemcpyAsync(dpv[i], hpv[i], h.psz[i] * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  float *dx = take(h.n_x), *dy = take(h.n_ycov), *dl = take(h.n_lab), *dout = take(h.n_out);
  float *dha = take(h.n_bnd), *dqu = take(h.n_bnd), *dpo = take(h.n_bnd), *dne = take(h.n_bnd);
  MCRN_CUDA_OK(cudaMemcpyAsync(dx, x, h.n_x * sizeof(float), cudaMemcpyHostToDevice, st));
  if (h.n_ycov) MCRN_CUDA_OK(cudaMemcpyAsync(dy, y_cov, h.n_ycov * sizeof(float), cudaMemcpyHostToDevice, st));
  if (labels) MCRN_CUDA_OK(cudaMemcpyAsync(dl, labels, h.n_lab * sizeof(float), cudaMemcpyHostToDevice, st));
  MCRN_TRY(mcrn_forward(dims, &dp, dx, dy, labels ? dl : nullptr, teacher_forcing, dout, dha, dqu, dpo, dne,
                        device_workspace, p.bytes, flags, stream));
  MCRN_CUDA_OK(cudaMemcpyAsync(output, dout, h.n_out * sizeof(float), cudaMemcpyDeviceToHost, st));
  MCRN_CUDA_OK(cudaMemcpyAsync(h_att, dha, h.n_bnd * sizeof(float), cudaMemcpyDeviceToHost, st));
  MCRN_CUDA_OK(cudaMemcpyAsync(query, dqu, h.n_bnd * sizeof(float), cudaMemcpyDeviceToHost, st));
  MCRN_CUDA_OK(cudaMemcpyAsync(pos, dpo, h.n_bnd * sizeof(float), cudaMemcpyDeviceToHost, st));
  MCRN_CUDA_OK(cudaMemcpyAsync(neg, dne, h.n_bnd * sizeof(float), cudaMemcpyDeviceToHost, st));
  MCRN_CUDA_OK(cudaStreamSynchronize(st));
  return MCRN_OK;
}

}  // extern "C"
