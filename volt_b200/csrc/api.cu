// extern "C" entry points of libvolt_b200.so (declared in include/volt_b200.h).
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/volt_b200.h"
#include "params.cuh"

namespace volt {

int launch_rollout_pack(const float* scalars, const float* Vt, const float* x, int B, int n, float* series, cudaStream_t st);
int launch_gather_scalar(const float* scalars, int B, int idx, float* out, cudaStream_t st);

// ---- error state
static thread_local char g_err[512] = "";
static long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (strstr(what, "_kernel") && !strstr(what, "cudaFunc")) ++g_launches;
  if (e == cudaSuccess) return VOLT_OK;
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return VOLT_ERR_CUDA;
}

// ---- cached workspaces, one arena per (device, slot, stream)
// Work submitted to one stream is ordered, so kernels that share an arena never overlap; two streams (or the library's own
// copy / compute streams behind the host-buffer entry) get separate arenas and therefore cannot race on the scratch.
// Grown on demand (device-wide synchronisation before the old block is freed), released by volt_release_workspaces().
constexpr int kSlots = 16;
struct Ws {
  int slot = 0;
  cudaStream_t stream = nullptr;
  void* ptr = nullptr;
  size_t bytes = 0;
};
static std::vector<Ws> g_ws[16];
static std::mutex g_mu;

int get_workspace(size_t bytes, void** ptr, int slot, cudaStream_t stream, int* created) {
  if (created) *created = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16 || slot < 0 || slot >= kSlots) {
    set_error("get_workspace: bad device/slot");
    return VOLT_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  Ws* w = nullptr;
  for (Ws& e : g_ws[dev])
    if (e.slot == slot && e.stream == stream) { w = &e; break; }
  if (bytes == 0) bytes = 256;
  if (!w || w->bytes < bytes) {
    // Nothing may be allocated while `stream` is being captured into a CUDA graph (torch.cuda.graph captures on a side
    // stream of its own).  A graph keeps the pointers it was captured with wherever it is replayed, so the capture
    // borrows an arena that an eager warm-up call has already created for this slot: replays then share that arena with
    // eager calls on the warm-up stream, which is also where callers replay (stream order keeps them apart).
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
      for (Ws& e : g_ws[dev])
        if (e.slot == slot && e.bytes >= bytes) { *ptr = e.ptr; return VOLT_OK; }
      set_error("workspace (slot %d, %zu bytes) cannot be created during stream capture: run the call once eagerly first", slot, bytes);
      return VOLT_ERR_ALLOC;
    }
  }
  if (!w) {
    g_ws[dev].emplace_back();
    w = &g_ws[dev].back();
    w->slot = slot;
    w->stream = stream;
  }
  if (w->bytes < bytes) {
    if (w->ptr) {
      cudaDeviceSynchronize();
      cudaFree(w->ptr);
      w->ptr = nullptr;
      w->bytes = 0;
    }
    const size_t want = (bytes + (size_t)(1 << 20) - 1) / (size_t)(1 << 20) * (size_t)(1 << 20);
    cudaError_t e = cudaMalloc(&w->ptr, want);
    if (e != cudaSuccess) {
      set_error("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
      w->ptr = nullptr;
      return VOLT_ERR_ALLOC;
    }
    w->bytes = want;
    if (created) *created = 1;
  }
  *ptr = w->ptr;
  return VOLT_OK;
}

static int release_workspaces() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return VOLT_ERR_CUDA;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaDeviceSynchronize();
  for (Ws& e : g_ws[dev])
    if (e.ptr) cudaFree(e.ptr);
  g_ws[dev].clear();
  return VOLT_OK;
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 0;
  return dev;
}

int sm_count() {
  static int cached[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return 148;
  if (!cached[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

static int device_check() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return VOLT_ERR_ARCH;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("volt_b200 requires an sm_100 (B200) device; found compute capability %d.x", major);
    return VOLT_ERR_ARCH;
  }
  return VOLT_OK;
}

// small helper kernels that only glue outputs together
__global__ void rollout_pack_kernel(const float* __restrict__ scalars, const float* __restrict__ Vt, const float* __restrict__ x,
                                    int B, int n, float* __restrict__ series) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* o = series + (size_t)b * NSERIES;
  o[0] = scalars[(size_t)b * VOLT_NSCALARS + VOLT_S_Z2Z2];
  o[1] = scalars[(size_t)b * VOLT_NSCALARS + VOLT_S_Z1Z2];
  o[2] = Vt[(size_t)b * n + n - 1];
  o[3] = x[1] - x[0];
  o[4] = scalars[(size_t)b * VOLT_NSCALARS + VOLT_S_JITTER];
  o[5] = o[6] = o[7] = 0.f;
}
int launch_rollout_pack(const float* scalars, const float* Vt, const float* x, int B, int n, float* series, cudaStream_t st) {
  rollout_pack_kernel<<<(B + 127) / 128, 128, 0, st>>>(scalars, Vt, x, B, n, series);
  return check_cuda(cudaGetLastError(), "rollout_pack_kernel");
}
__global__ void gather_scalar_kernel(const float* __restrict__ scalars, int B, int idx, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = scalars[(size_t)b * VOLT_NSCALARS + idx];
}
int launch_gather_scalar(const float* scalars, int B, int idx, float* out, cudaStream_t st) {
  gather_scalar_kernel<<<(B + 127) / 128, 128, 0, st>>>(scalars, B, idx, out);
  return check_cuda(cudaGetLastError(), "gather_scalar_kernel");
}

// [GPyTorch] GaussianLikelihood noise transform and the step's scalar outputs for the kernels that do not fuse them
// (SIMT A/B kernel, multi-CTA long-series path): same arithmetic and summation order as the fused epilogue of chol_tc.cu.
__global__ void noise_from_raw_kernel(const float* __restrict__ raw, int stride, int B, float* __restrict__ noise) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) noise[b] = noise_from_raw_dev(raw[(size_t)b * stride]);
}
__global__ void __launch_bounds__(256) raw_finish_kernel(const float* __restrict__ raw, int stride, const float* __restrict__ noise,
                                                         int B, float* __restrict__ scalars, float* __restrict__ loss_out,
                                                         LossExchange ex) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    float* o = scalars + (size_t)b * VOLT_NSCALARS;
    o[VOLT_S_DRAW] = o[VOLT_S_DNOISE] * sigmoid_dev(raw[(size_t)b * stride]);
    o[VOLT_S_NOISE] = noise[b];
    acc += o[VOLT_S_MLL];
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0 && loss_out) loss_out[0] = -tot;
  if ((ex.peers || ex.totals) && threadIdx.x < 32) exchange_partial_warp(ex, -tot, threadIdx.x);
}

// rank-local sum of the partial losses the ranks pushed into this rank's slots for step `seq` (fixed rank order: every
// rank gets the same bits).  A slot that never arrives (a peer that failed) ends the wait after ~2 s with NaN.
__global__ void loss_gather_kernel(const unsigned long long* __restrict__ slots, int world, int ring, unsigned int seq,
                                   float* __restrict__ out) {
  const float tot = exchange_sum_warp(slots, world, ring, seq, threadIdx.x);
  if (threadIdx.x == 0) out[0] = tot;
}
__global__ void push_zero_kernel(LossExchange ex) { exchange_partial_warp(ex, 0.f, threadIdx.x); }
// the push alone, from a value in device memory (volt_loss_push): a one-warp kernel that fits next to the resident step kernels
__global__ void push_value_kernel(LossExchange ex, const float* value) { exchange_partial_warp(ex, *value, threadIdx.x); }

// implementation switch for the batched MLL kernel: 1 = tcgen05 (default), 0 = SIMT fp32 (kept for A/B measurement)
static int g_mll_impl = -1;
int launch_mll_batched(MllParams p, cudaStream_t st) {
  if (g_mll_impl < 0) {
    const char* e = getenv("VOLT_MLL_IMPL");
    g_mll_impl = (e && (e[0] == 's' || e[0] == '0')) ? 0 : 1;
  }
  // A few very long series: one CTA per series would leave the GPU idle -> multi-CTA right-looking path, series by series.
  if (g_mll_impl && p.resid && !p.resid2 && !p.L_out && !p.U_out && !p.z_out && p.T >= 1536 && p.B <= 16) {
    for (int b = 0; b < p.B; ++b) {
      int s = launch_mll_large(p, b, st);
      if (s) return s;
    }
    return VOLT_OK;
  }
  return g_mll_impl ? launch_mll_batched_tc(p, st) : launch_mll_batched_simt(p, st);
}

static MllParams base_params(int B, int T, const float* resid, const float* noise, int noise_stride, float jitter, int max_tries,
                             float* scalars, float* alpha, int* info) {
  MllParams p;
  memset(&p, 0, sizeof(p));
  p.B = B;
  p.T = T;
  p.resid = resid;
  p.diag_add = noise;
  p.diag_stride = noise_stride;
  p.jitter = jitter;
  p.max_tries = max_tries;
  p.scalars = scalars;
  p.alpha = alpha;
  p.info = info;
  p.do_inverse = resid ? 1 : 0;
  return p;
}

}  // namespace volt

using namespace volt;

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* volt_last_error(void) { return g_err; }
int volt_abi_version(void) { return VOLT_ABI_VERSION; }
int volt_device_check(void) { return device_check(); }
long long volt_launch_count(void) { return g_launches; }
int volt_release_workspaces(void) { return release_workspaces(); }
int volt_set_mll_impl(int impl) {
  const int prev = g_mll_impl;
  g_mll_impl = impl ? 1 : 0;
  return prev;
}

int volt_cumtrapz(const float* x, int x_batched, const float* y, int B, int T, int vol_mode, int half_last, float* V, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(x && y && V, "volt_cumtrapz: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 2, "volt_cumtrapz: need B >= 1 and T >= 2 (got B=%d, T=%d)", B, T);
  VOLT_REQUIRE(vol_mode >= 0 && vol_mode <= 2, "volt_cumtrapz: bad vol_mode %d", vol_mode);
  return launch_cumtrapz(x, x_batched, y, B, T, vol_mode, half_last, V, ST(stream));
}

int volt_vol_cov(const float* x, int x_batched, const float* vol, int vol_mode, int B, int T, const float* add_diag, int add_stride,
                 float* K, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(x && vol && K, "volt_vol_cov: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 2, "volt_vol_cov: need B >= 1 and T >= 2 (got B=%d, T=%d)", B, T);
  void* V = nullptr;
  int s = get_workspace((size_t)B * T * sizeof(float), &V, 1, ST(stream));
  if (s) return s;
  s = launch_cumtrapz(x, x_batched, vol, B, T, vol_mode, 1, (float*)V, ST(stream));
  if (s) return s;
  return launch_vol_cov((const float*)V, add_diag, add_stride, B, T, K, ST(stream));
}

int volt_bm_cov(const float* x1, int n1, const float* x2, int n2, const float* vol, float* K, void* stream) {
  VOLT_REQUIRE(x1 && x2 && vol && K, "volt_bm_cov: null pointer");
  VOLT_REQUIRE(n1 >= 1 && n2 >= 1, "volt_bm_cov: empty input");
  return launch_bm_cov(x1, n1, x2, n2, vol, K, ST(stream));
}

int volt_ewma(const float* y, int S, int T, int k, float* out, void* stream) {
  if (S == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(y && out, "volt_ewma: null pointer");
  VOLT_REQUIRE(S >= 1 && T >= 1 && k >= 1, "volt_ewma: need S,T,k >= 1 (got %d,%d,%d)", S, T, k);
  void* w = nullptr;
  int s = get_workspace((size_t)k * sizeof(float), &w, 2, ST(stream));
  if (s) return s;
  s = launch_ewma_weights(k, (float*)w, ST(stream));
  if (s) return s;
  return launch_ewma(y, S, T, k, (const float*)w, out, ST(stream));
}

int volt_ma_mean(const float* y, int S, int T, int k, int kind, float theta, const float* latent, float* out, float* e_out,
                 float* ee_out, float* resid_out, void* stream) {
  if (S == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(y && out, "volt_ma_mean: null pointer");
  VOLT_REQUIRE(S >= 1 && T >= 1 && k >= 1, "volt_ma_mean: need S,T,k >= 1 (got %d,%d,%d)", S, T, k);
  VOLT_REQUIRE(kind >= VOLT_MA_EWMA && kind <= VOLT_MA_MEANREVERT, "volt_ma_mean: bad kind %d", kind);
  VOLT_REQUIRE(kind != VOLT_MA_MEANREVERT || latent, "volt_ma_mean: meanrevert needs latent");
  void* w = nullptr;
  int s = get_workspace((size_t)k * sizeof(float), &w, 2, ST(stream));
  if (s) return s;
  s = launch_ewma_weights(k, (float*)w, ST(stream));
  if (s) return s;
  return launch_ma_paths(y, S, T, k, (const float*)w, kind, theta, latent, out, e_out, ee_out, resid_out, ST(stream));
}

int volt_mll_grad_vol(const float* x, int x_batched, const float* vol, int vol_mode, const float* resid, const float* noise,
                      int noise_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                      void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(x && vol && resid && scalars, "volt_mll_grad_vol: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 2, "volt_mll_grad_vol: need B >= 1 and T >= 2 (got B=%d, T=%d)", B, T);
  void* V = nullptr;
  int s = get_workspace((size_t)B * T * sizeof(float), &V, 1, ST(stream));
  if (s) return s;
  s = launch_cumtrapz(x, x_batched, vol, B, T, vol_mode, 1, (float*)V, ST(stream));
  if (s) return s;
  MllParams p = base_params(B, T, resid, noise, noise_stride, jitter, max_tries, scalars, alpha, info);
  p.kind = KIND_VOL;
  p.V = (const float*)V;
  return launch_mll_batched(p, ST(stream));
}

static int mll_step_impl(const float* x, int x_batched, const float* vol, int vol_mode, const float* resid, const float* raw_noise,
                         int raw_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                         float* loss_out, const LossExchange& ex, void* stream);

int volt_mll_grad_vol_raw(const float* x, int x_batched, const float* vol, int vol_mode, const float* resid, const float* raw_noise,
                          int raw_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                          float* loss_out, void* stream) {
  return mll_step_impl(x, x_batched, vol, vol_mode, resid, raw_noise, raw_stride, B, T, jitter, max_tries, scalars, alpha, info, loss_out,
                       LossExchange{}, stream);
}

int volt_mll_step_sharded(const float* x, int x_batched, const float* vol, int vol_mode, const float* resid, const float* raw_noise,
                          int raw_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                          float* loss_out, const void* peer_slot_ptrs, const void* local_slots, float* prev_totals, int lag, int world,
                          int rank, int ring, unsigned int seq, void* stream) {
  VOLT_REQUIRE(world >= 1 && rank >= 0 && rank < world && ring >= 2 && lag >= 1 && lag < ring,
               "volt_mll_step_sharded: bad exchange description (world=%d rank=%d ring=%d lag=%d)", world, rank, ring, lag);
  VOLT_REQUIRE(!prev_totals || seq > (unsigned)lag, "volt_mll_step_sharded: prev_totals needs seq > lag");
  VOLT_REQUIRE(!prev_totals || local_slots, "volt_mll_step_sharded: exchange with prev_totals needs local_slots");
  LossExchange ex;
  ex.peers = reinterpret_cast<const unsigned long long*>(peer_slot_ptrs);
  ex.mine = reinterpret_cast<const unsigned long long*>(local_slots);
  ex.totals = prev_totals;
  ex.world = world; ex.rank = rank; ex.ring = ring; ex.lag = lag; ex.seq = seq;
  return mll_step_impl(x, x_batched, vol, vol_mode, resid, raw_noise, raw_stride, B, T, jitter, max_tries, scalars, alpha, info, loss_out,
                       ex, stream);
}

int volt_loss_push(const float* value, const void* peer_slot_ptrs, int world, int rank, int ring, unsigned int seq, void* stream) {
  VOLT_REQUIRE(value && peer_slot_ptrs && world >= 1 && rank >= 0 && rank < world && ring >= 2, "volt_loss_push: bad arguments");
  LossExchange ex;
  memset(&ex, 0, sizeof(ex));
  ex.peers = reinterpret_cast<const unsigned long long*>(peer_slot_ptrs);
  ex.world = world; ex.rank = rank; ex.ring = ring; ex.lag = 1; ex.seq = seq;
  push_value_kernel<<<1, 32, 0, ST(stream)>>>(ex, value);
  return check_cuda(cudaGetLastError(), "push_value_kernel");
}

int volt_loss_gather(const void* local_slots, int world, int ring, unsigned int seq, float* out, void* stream) {
  VOLT_REQUIRE(local_slots && out && world >= 1 && ring >= 2, "volt_loss_gather: bad arguments");
  loss_gather_kernel<<<1, 32, 0, ST(stream)>>>(reinterpret_cast<const unsigned long long*>(local_slots), world, ring, seq, out);
  return check_cuda(cudaGetLastError(), "loss_gather_kernel");
}

static int mll_step_impl(const float* x, int x_batched, const float* vol, int vol_mode, const float* resid, const float* raw_noise,
                         int raw_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                         float* loss_out, const LossExchange& ex, void* stream) {
  if (B == 0) {   // empty shard: the step's partial loss is 0
    if (loss_out) VOLT_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), ST(stream)));
    if (ex.peers || ex.totals) {
      push_zero_kernel<<<1, 32, 0, ST(stream)>>>(ex);
      return check_cuda(cudaGetLastError(), "push_zero_kernel");
    }
    return VOLT_OK;
  }
  VOLT_REQUIRE(x && vol && resid && raw_noise && scalars, "volt_mll_grad_vol_raw: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 2, "volt_mll_grad_vol_raw: need B >= 1 and T >= 2 (got B=%d, T=%d)", B, T);
  VOLT_REQUIRE(!(ex.peers || ex.totals) || loss_out, "volt_mll_step_sharded: the exchange needs loss_out");
  cudaStream_t st = ST(stream);
  void* V = nullptr;
  int s = get_workspace((size_t)B * T * sizeof(float), &V, 1, st);
  if (s) return s;
  void* aux = nullptr;   // [0] completion counter of the fused loss reduction | [64...] noise (B) for the unfused kernels
  int created = 0;
  s = get_workspace(256 + (size_t)B * sizeof(float), &aux, 13, st, &created);
  if (s) return s;
  if (created) VOLT_CUDA(cudaMemsetAsync(aux, 0, 256, st));
  s = launch_cumtrapz(x, x_batched, vol, B, T, vol_mode, 1, (float*)V, st);
  if (s) return s;
  if (g_mll_impl < 0) {
    const char* e = getenv("VOLT_MLL_IMPL");
    g_mll_impl = (e && (e[0] == 's' || e[0] == '0')) ? 0 : 1;
  }
  const bool fused = g_mll_impl && !(T >= 1536 && B <= 16);
  if (fused) {
    MllParams p = base_params(B, T, resid, nullptr, 0, jitter, max_tries, scalars, alpha, info);
    p.kind = KIND_VOL;
    p.V = (const float*)V;
    p.raw_noise = raw_noise;
    p.raw_stride = raw_stride;
    p.loss_out = loss_out;
    p.done_counter = (unsigned int*)aux;
    p.ex = ex;
    return launch_mll_batched_tc(p, st);
  }
  float* noise = (float*)aux + 64;
  noise_from_raw_kernel<<<(B + 127) / 128, 128, 0, st>>>(raw_noise, raw_stride, B, noise);
  s = check_cuda(cudaGetLastError(), "noise_from_raw_kernel");
  if (s) return s;
  MllParams p = base_params(B, T, resid, noise, 1, jitter, max_tries, scalars, alpha, info);
  p.kind = KIND_VOL;
  p.V = (const float*)V;
  s = launch_mll_batched(p, st);
  if (s) return s;
  raw_finish_kernel<<<1, 256, 0, st>>>(raw_noise, raw_stride, noise, B, scalars, loss_out, ex);
  return check_cuda(cudaGetLastError(), "raw_finish_kernel");
}

int volt_mll_grad_bm(const float* x, const float* scale, int scale_stride, const float* resid, const float* noise, int noise_stride,
                     int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(x && scale && resid && scalars, "volt_mll_grad_bm: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 1, "volt_mll_grad_bm: need B,T >= 1");
  MllParams p = base_params(B, T, resid, noise, noise_stride, jitter, max_tries, scalars, alpha, info);
  p.kind = KIND_BM;
  p.x = x;
  p.scale = scale;
  p.scale_stride = scale_stride;
  return launch_mll_batched(p, ST(stream));
}

int volt_mll_grad_bm_inv(const float* x, const float* scale, int scale_stride, const float* resid, const float* noise, int noise_stride,
                         int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info, float* linv_t,
                         void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(x && scale && resid && scalars && linv_t, "volt_mll_grad_bm_inv: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 1, "volt_mll_grad_bm_inv: need B,T >= 1");
  VOLT_REQUIRE(g_mll_impl, "volt_mll_grad_bm_inv: needs the tensor-core implementation (VOLT_MLL_IMPL=tc)");
  MllParams p = base_params(B, T, resid, noise, noise_stride, jitter, max_tries, scalars, alpha, info);
  p.kind = KIND_BM;
  p.x = x;
  p.scale = scale;
  p.scale_stride = scale_stride;
  p.U_out = linv_t;
  return launch_mll_batched(p, ST(stream));
}

int volt_mll_grad_dense(const float* K, long long k_bstride, int ld, const float* resid, const float* noise, int noise_stride, int B,
                        int T, float jitter, int max_tries, float* scalars, float* alpha, int* info, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(K && resid && scalars, "volt_mll_grad_dense: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 1 && ld >= T, "volt_mll_grad_dense: bad shape");
  MllParams p = base_params(B, T, resid, noise, noise_stride, jitter, max_tries, scalars, alpha, info);
  p.kind = KIND_DENSE;
  p.dense = K;
  p.dense_bstride = k_bstride;
  p.ldd = ld;
  return launch_mll_batched(p, ST(stream));
}

int volt_mll_grad_vol_host(const float* x, const float* vol, const float* resid, const float* noise, int noise_stride, int B, int T,
                           float jitter, int max_tries, float* scalars, float* alpha, int* info) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(x && vol && resid && noise && scalars, "volt_mll_grad_vol_host: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 2, "volt_mll_grad_vol_host: need B >= 1 and T >= 2");
  const size_t bt = (size_t)B * T;
  const size_t n_noise = noise_stride ? (size_t)B : 1;
  // one staging buffer: x | vol | resid | noise | scalars | alpha | info
  const size_t floats = (size_t)T + bt + bt + n_noise + (size_t)B * VOLT_NSCALARS + bt + (size_t)B;
  // Pipeline: the inputs of the first `B0` series are copied and ONE kernel is launched for the whole batch; the copy of
  // the remaining series runs on the copy stream underneath it, followed by a 4-byte flag.  The persistent CTAs take
  // series b, b + grid, ... in order and wait for the flag before touching a series >= B0 (by the time a CTA gets
  // there the data has long arrived).  The flag is written by the copy engine, so it cannot depend on an SM the
  // waiting CTAs hold.  The prefix sums (CumTrapz) are built inside the kernel: no second launch, no second tail.
  struct HostCtx { cudaStream_t s_copy, s_comp; cudaEvent_t ev0; int* h_flags; };   // per device
  static HostCtx g_ctx[16] = {};
  HostCtx& hc = g_ctx[device_slot()];
  cudaStream_t& s_copy = hc.s_copy;
  cudaStream_t& s_comp = hc.s_comp;
  cudaEvent_t& ev0 = hc.ev0;
  int*& h_flags = hc.h_flags;      // pinned {0, 0 | 1 | timeout read-back}
  if (!s_copy) {
    VOLT_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    VOLT_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    VOLT_CUDA(cudaEventCreateWithFlags(&ev0, cudaEventDisableTiming));
    VOLT_CUDA(cudaHostAlloc(&h_flags, 4 * sizeof(int), cudaHostAllocDefault));
    h_flags[0] = 0;
    h_flags[1] = 0;
    h_flags[2] = 1;
  }
  if (g_mll_impl < 0) {
    const char* e = getenv("VOLT_MLL_IMPL");
    g_mll_impl = (e && (e[0] == 's' || e[0] == '0')) ? 0 : 1;
  }
  // Direct form: when every buffer of the call is page-locked host memory (cudaHostAlloc / cudaHostRegister, e.g. torch's
  // pin_memory()), it is already mapped into the device's address space: the kernel reads each series' 12 T input bytes
  // straight from it at the start of that series (one PCIe round trip per series, hidden under the 0.4 ms the series
  // takes) and writes the 64 output bytes (+ alpha) straight back.  No staging copies, no arrival flag, nothing before or
  // after the launch but the final synchronisation.  VOLT_E2E_DIRECT=0 forces the staged pipeline below (also taken for
  // pageable buffers, the SIMT implementation and the multi-CTA path of very long series).
  {
    static const int direct_env = [] { const char* e = getenv("VOLT_E2E_DIRECT"); return e ? atoi(e) : 1; }();
    auto mapped = [](const void* h, void** d) -> bool {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return false; }
      if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
      *d = a.devicePointer;
      return true;
    };
    void *mx, *mv, *mr, *mn, *ms, *ma = nullptr, *mi = nullptr;
    if (direct_env && g_mll_impl && !(T >= 1536 && B <= 16) && mapped(x, &mx) && mapped(vol, &mv) && mapped(resid, &mr) &&
        mapped(noise, &mn) && mapped(scalars, &ms) && (!alpha || mapped(alpha, &ma)) && (!info || mapped(info, &mi))) {
      void* dinfo = mi;
      if (!dinfo) {   // the kernel always records the status
        int s = get_workspace((size_t)B * sizeof(int), &dinfo, 12, s_comp);
        if (s) return s;
      }
      MllParams p = base_params(B, T, (const float*)mr, (const float*)mn, noise_stride, jitter, max_tries, (float*)ms, (float*)ma, (int*)dinfo);
      p.kind = KIND_VOL;
      p.vol_in = (const float*)mv;
      p.x_in = (const float*)mx;
      p.x_batched = 0;
      p.vol_mode = VOLT_VOL_SIGMA;
      p.stage_in = 1;
      int s = launch_mll_batched_tc(p, s_comp);
      if (s) return s;
      VOLT_CUDA(cudaStreamSynchronize(s_comp));
      return VOLT_OK;
    }
  }
  // every arena of this entry (staging, factor scratch, prefix sums) is keyed by the library's own compute stream, so the
  // call cannot race with work the caller has in flight on their streams through the device-pointer entry points
  void* ws = nullptr;
  int s = get_workspace(floats * sizeof(float), &ws, 3, s_comp);
  if (s) return s;
  float* d_x = (float*)ws;
  float* d_vol = d_x + T;
  float* d_res = d_vol + bt;
  float* d_noise = d_res + bt;
  float* d_scal = d_noise + n_noise;
  float* d_alpha = d_scal + (size_t)B * VOLT_NSCALARS;
  int* d_info = (int*)(d_alpha + bt);
  void* vflag = nullptr;
  s = get_workspace(128, &vflag, 12, s_comp);
  if (s) return s;
  int* d_flag = (int*)vflag;
  if (g_mll_impl < 0) {
    const char* e = getenv("VOLT_MLL_IMPL");
    g_mll_impl = (e && (e[0] == 's' || e[0] == '0')) ? 0 : 1;
  }
  // first chunk: VOLT_E2E_CHUNK0 series (developer knob), default one series per SM: the CTAs that own three series of a
  // c2-sized batch are the first ones to be scheduled, and they are the critical path.  Gating needs whole cache lines
  // per series (T % 32 == 0: a line shared by a copied and a not-yet-copied series could be cached stale) and the
  // batched tensor-core kernel (not the multi-CTA path for a few very long series, not the SIMT kernel).
  static const int chunk0_env = [] { const char* e = getenv("VOLT_E2E_CHUNK0"); return e ? atoi(e) : 0; }();
  const int slots = chunk0_env > 0 ? chunk0_env : sm_count();   // (two per SM for the control-warp instance was measured: 572 k vs 577 k evals/s)
  const bool batched_tc = g_mll_impl && !(T >= 1536 && B <= 16);
  const bool gated = batched_tc && (T % 32 == 0) && B >= 2 * slots;
  const int B0 = gated ? slots : B;
  const size_t n0 = (size_t)B0 * T;
  VOLT_CUDA(cudaMemcpyAsync(d_flag, h_flags, 2 * sizeof(int), cudaMemcpyHostToDevice, s_copy));   // arrival flag, timeout flag
  VOLT_CUDA(cudaMemcpyAsync(d_x, x, (size_t)T * 4, cudaMemcpyHostToDevice, s_copy));
  VOLT_CUDA(cudaMemcpyAsync(d_noise, noise, n_noise * 4, cudaMemcpyHostToDevice, s_copy));
  VOLT_CUDA(cudaMemcpyAsync(d_vol, vol, n0 * 4, cudaMemcpyHostToDevice, s_copy));
  VOLT_CUDA(cudaMemcpyAsync(d_res, resid, n0 * 4, cudaMemcpyHostToDevice, s_copy));
  VOLT_CUDA(cudaEventRecord(ev0, s_copy));
  VOLT_CUDA(cudaStreamWaitEvent(s_comp, ev0, 0));
  if (batched_tc) {
    MllParams p = base_params(B, T, d_res, d_noise, noise_stride, jitter, max_tries, d_scal, alpha ? d_alpha : nullptr, d_info);
    p.kind = KIND_VOL;
    p.vol_in = d_vol;
    p.x_in = d_x;
    p.x_batched = 0;
    p.vol_mode = VOLT_VOL_SIGMA;
    p.ready = gated ? d_flag : nullptr;
    p.ready_from = B0;
    p.ready_timeout = d_flag + 1;
    // ~0.6 us per poll (measured): 20 ms plus the time the remaining bytes would need at 1 GB/s
    p.ready_spins = 32768 + (long long)((bt - n0) * 8 / 600);
    s = launch_mll_batched_tc(p, s_comp);
  } else {
    s = volt_mll_grad_vol(d_x, 0, d_vol, VOLT_VOL_SIGMA, d_res, d_noise, noise_stride, B, T, jitter, max_tries, d_scal,
                          alpha ? d_alpha : nullptr, d_info, s_comp);
  }
  if (s) return s;
  if (gated) {   // submitted after the launch: the kernel does not wait for the host to queue these
    const cudaError_t e1 = cudaMemcpyAsync(d_vol + n0, vol + n0, (bt - n0) * 4, cudaMemcpyHostToDevice, s_copy);
    const cudaError_t e2 = cudaMemcpyAsync(d_res + n0, resid + n0, (bt - n0) * 4, cudaMemcpyHostToDevice, s_copy);
    // the flag must be raised whatever happened above: CTAs are waiting for it (they also give up after a while)
    cudaError_t e3 = cudaMemcpyAsync(d_flag, h_flags + 2, sizeof(int), cudaMemcpyHostToDevice, s_copy);
    if (e3 != cudaSuccess) e3 = cudaMemcpy(d_flag, h_flags + 2, sizeof(int), cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
      cudaStreamSynchronize(s_comp);
      return check_cuda(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3), "volt_mll_grad_vol_host: copy of the later series");
    }
  }
  cudaStream_t st = s_comp;
  h_flags[3] = 0;
  if (gated) VOLT_CUDA(cudaMemcpyAsync(h_flags + 3, d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  for (int pass = 0; pass < 2; ++pass) {
    VOLT_CUDA(cudaMemcpyAsync(scalars, d_scal, (size_t)B * VOLT_NSCALARS * 4, cudaMemcpyDeviceToHost, st));
    if (alpha) VOLT_CUDA(cudaMemcpyAsync(alpha, d_alpha, bt * 4, cudaMemcpyDeviceToHost, st));
    if (info) VOLT_CUDA(cudaMemcpyAsync(info, d_info, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    VOLT_CUDA(cudaStreamSynchronize(st));
    if (pass == 1 || !gated || h_flags[3] == 0) break;
    // A CTA gave up waiting for the arrival flag (streams serialised by a profiler or CUDA_LAUNCH_BLOCKING).  Every input
    // has landed once the copy stream is idle: run the batch again without the gate.
    VOLT_CUDA(cudaStreamSynchronize(s_copy));
    MllParams p = base_params(B, T, d_res, d_noise, noise_stride, jitter, max_tries, d_scal, alpha ? d_alpha : nullptr, d_info);
    p.kind = KIND_VOL;
    p.vol_in = d_vol;
    p.x_in = d_x;
    p.vol_mode = VOLT_VOL_SIGMA;
    s = launch_mll_batched_tc(p, st);
    if (s) return s;
  }
  return VOLT_OK;
}

int volt_potrf(const float* A, long long a_bstride, int lda, const float* add_diag, int add_stride, int B, int T, float jitter,
               int max_tries, float* L, long long l_bstride, int ldl, float* jitter_used, int* info, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(A && L, "volt_potrf: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 1 && lda >= T && ldl >= T, "volt_potrf: bad shape");
  void* sc = nullptr;
  int s = get_workspace((size_t)B * VOLT_NSCALARS * sizeof(float), &sc, 4, ST(stream));
  if (s) return s;
  MllParams p = base_params(B, T, nullptr, add_diag, add_stride, jitter, max_tries, (float*)sc, nullptr, info);
  p.kind = KIND_DENSE;
  p.dense = A;
  p.dense_bstride = a_bstride;
  p.ldd = lda;
  p.L_out = L;
  p.L_bstride = l_bstride;
  p.ldl = ldl;
  s = launch_mll_batched(p, ST(stream));
  if (s) return s;
  if (jitter_used) return launch_gather_scalar((const float*)sc, B, VOLT_S_JITTER, jitter_used, ST(stream));
  return VOLT_OK;
}

int volt_potrs(const float* L, long long l_bstride, int ldl, int B, int T, float* rhs, long long r_bstride, int nrhs,
               int forward_only, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(L && rhs, "volt_potrs: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 1 && nrhs >= 1 && ldl >= T, "volt_potrs: bad shape");
  return launch_chol_solve(L, l_bstride, ldl, B, T, rhs, r_bstride, nrhs, forward_only ? 1 : 0, ST(stream));
}

int volt_bmgp_posterior(const float* x, const float* y, int B, int T, const float* xs, int H, const float* vol, int vol_stride,
                        const float* noise, int noise_stride, float* mean, float* cov, int* info, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(x && y && xs && vol && noise && mean && cov, "volt_bmgp_posterior: null pointer");
  VOLT_REQUIRE(B >= 1 && T >= 1 && H >= 1, "volt_bmgp_posterior: bad shape");
  const size_t wfl = (size_t)B * T * (H + 1) + (size_t)B * H * H + (size_t)B * H + (size_t)B * VOLT_NSCALARS;
  void* ws = nullptr;
  int s = get_workspace(wfl * sizeof(float), &ws, 5, ST(stream));
  if (s) return s;
  float* W0 = (float*)ws;
  float* Kss = W0 + (size_t)B * T * (H + 1);
  float* mean_s = Kss + (size_t)B * H * H;
  float* scal = mean_s + (size_t)B * H;
  void* Lw = nullptr;
  s = get_workspace((size_t)B * T * T * sizeof(float), &Lw, 6, ST(stream));
  if (s) return s;
  s = launch_bm_posterior_pack(x, B, T, xs, H, y, vol, vol_stride, W0, Kss, mean_s, nullptr, ST(stream));
  if (s) return s;
  MllParams p = base_params(B, T, nullptr, noise, noise_stride, 1e-6f, 3, scal, nullptr, info);
  p.kind = KIND_BM;
  p.x = x;
  p.scale = vol;
  p.scale_stride = vol_stride;
  p.L_out = (float*)Lw;
  p.L_bstride = (long long)T * T;
  p.ldl = T;
  s = launch_mll_batched(p, ST(stream));
  if (s) return s;
  s = launch_chol_solve((const float*)Lw, (long long)T * T, T, B, T, W0, (long long)T * (H + 1), H + 1, 1, ST(stream));
  if (s) return s;
  return launch_posterior(W0, B, T, H, Kss, mean_s, mean, cov, ST(stream));
}

int volt_mvn_sample(const float* mean, const float* cov, const float* eps, int B, int H, int S, float jitter, int exp_out,
                    float* samples, int* info, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(mean && cov && eps && samples, "volt_mvn_sample: null pointer");
  VOLT_REQUIRE(B >= 1 && H >= 1 && S >= 1, "volt_mvn_sample: bad shape");
  void* Lc = nullptr;
  int s = get_workspace((size_t)B * H * H * sizeof(float), &Lc, 7, ST(stream));
  if (s) return s;
  s = volt_potrf(cov, (long long)H * H, H, nullptr, 0, B, H, jitter, 3, (float*)Lc, (long long)H * H, H, nullptr, info, stream);
  if (s) return s;
  return launch_mvn_sample(mean, (const float*)Lc, eps, B, H, S, exp_out, samples, ST(stream));
}

int volt_rollout(const float* x, const float* logy, const float* vol, int vol_mode, const float* pred_vol, const float* eps, int B,
                 int n, int S, int H, int mean_kind, int k, float mr_theta, const float* mr_latent, const float* resid_given,
                 const float* mean_test, int use_theta, float theta, const float* latent, int joint, float jitter,
                 unsigned long long seed, float* samples, int* draw_info, int* series_info, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(x && logy && vol && pred_vol && samples, "volt_rollout: null pointer");
  VOLT_REQUIRE(B >= 1 && n >= 2 && S >= 1 && H >= 1, "volt_rollout: bad shape (B=%d n=%d S=%d H=%d)", B, n, S, H);
  VOLT_REQUIRE(mean_kind >= VOLT_MA_EWMA && mean_kind <= VOLT_MA_GIVEN, "volt_rollout: bad mean_kind %d", mean_kind);
  const bool ma = mean_kind != VOLT_MA_GIVEN;
  VOLT_REQUIRE(!ma || k >= 1, "volt_rollout: moving-average means need k >= 1");
  VOLT_REQUIRE(ma || (resid_given && (mean_test || !joint)), "volt_rollout: VOLT_MA_GIVEN needs resid_given (and mean_test)");
  VOLT_REQUIRE(!(joint && ma && H > 1), "volt_rollout: the moving-average means support one test point per call (EWMA.py:48-54)");
  VOLT_REQUIRE(mean_kind != VOLT_MA_MEANREVERT || mr_latent, "volt_rollout: meanrevert needs mr_latent");
  VOLT_REQUIRE(!use_theta || latent, "volt_rollout: use_theta needs latent");
  cudaStream_t st = ST(stream);
  // workspace: Vt (B,n) | path,e,ee (B,n+1)x3 | resid (B,n) | scalars (B,16) | series (B,4) | w (k) | sinfo (B)
  const size_t bn = (size_t)B * n, bn1 = (size_t)B * (n + 1);
  const size_t fl = bn + 3 * bn1 + bn + (size_t)B * VOLT_NSCALARS + (size_t)B * NSERIES + (size_t)(k > 0 ? k : 1) + (size_t)B;
  void* ws = nullptr;
  int s = get_workspace(fl * sizeof(float), &ws, 8, st);
  if (s) return s;
  float* Vt = (float*)ws;
  float* path = Vt + bn;
  float* e_tr = path + bn1;
  float* ee_tr = e_tr + bn1;
  float* resid = ee_tr + bn1;
  float* scal = resid + bn;
  float* series = scal + (size_t)B * VOLT_NSCALARS;
  float* w = series + (size_t)B * NSERIES;
  int* sinfo = (int*)(w + (k > 0 ? k : 1));
  s = launch_cumtrapz(x, 0, vol, B, n, vol_mode, 0, Vt, st);
  if (s) return s;
  const float* r1 = resid_given;
  if (ma) {
    s = launch_ewma_weights(k, w, st);
    if (s) return s;
    s = launch_ma_paths(logy, B, n, k, w, mean_kind, mr_theta, mr_latent, path, e_tr, ee_tr, resid, st);
    if (s) return s;
    r1 = resid;
  }
  MllParams p = base_params(B, n, r1, nullptr, 0, jitter, 3, scal, nullptr, series_info ? series_info : sinfo);
  p.kind = KIND_VOL;
  p.V = Vt;
  p.resid2 = Vt;
  p.do_inverse = 0;
  if (g_mll_impl < 0) {
    const char* e = getenv("VOLT_MLL_IMPL");
    g_mll_impl = (e && (e[0] == 's' || e[0] == '0')) ? 0 : 1;
  }
  // Tensor-core batched kernel: it writes the rollout kernel's per-series inputs itself and raises a flag per series; the
  // rollout kernel is launched as a programmatic dependent and starts on the finished series while the prep kernel's last
  // wave is still running.  (Other kernels: pack kernel + plain stream order.)
  static const int overlap_env = [] { const char* e = getenv("VOLT_ROLLOUT_OVERLAP"); return e ? atoi(e) : 1; }();
  const bool overlap = overlap_env && g_mll_impl && !(n >= 1536 && B <= 16);
  int* flags = nullptr;
  if (overlap) {
    void* fl_ws = nullptr;
    s = get_workspace((size_t)B * sizeof(int), &fl_ws, 14, st);
    if (s) return s;
    flags = (int*)fl_ws;
    VOLT_CUDA(cudaMemsetAsync(flags, 0, (size_t)B * sizeof(int), st));
    p.pack_out = series;
    p.pack_x = x;
    p.series_flag = flags;
  }
  s = launch_mll_batched(p, st);
  if (s) return s;
  if (!overlap) {
    s = launch_rollout_pack(scal, Vt, x, B, n, series, st);
    if (s) return s;
  }
  RolloutParams q;
  memset(&q, 0, sizeof(q));
  q.B = B; q.n = n; q.S = S; q.H = H; q.k = ma ? k : 0; q.mean_kind = mean_kind; q.joint = joint;
  q.w = ma ? w : nullptr;
  q.ytrain = logy;
  q.e_train = ma ? e_tr : nullptr;
  q.ee_train = ma ? ee_tr : nullptr;
  q.mean_test = mean_test;
  q.series = series;
  q.series_info = series_info ? series_info : sinfo;
  q.pred_vol = pred_vol;
  q.eps = eps;
  q.latent = latent;
  q.theta = theta;
  q.use_theta = use_theta;
  q.mr_latent = mr_latent;
  q.mr_theta = mr_theta;
  q.jitter = jitter;
  q.seed = seed;
  q.samples = samples;
  q.info = draw_info;
  q.series_flag = flags;
  return launch_rollout(q, st);
}

int volt_rollout_normals(unsigned long long seed, int B, int S, int H, int joint, float* eps, void* stream) {
  VOLT_REQUIRE(B >= 0 && S >= 0 && H >= 0 && (eps || (long long)B * S * H == 0), "volt_rollout_normals: bad arguments");
  // series beyond 65535 are numbered on (b_offset + b) inside volt_rollout's chunks: one global numbering here
  return launch_rollout_normals(seed, 0, B, S, H, joint, eps, ST(stream));
}

int volt_rollout_stats(const float* samples, int B, int S, int H, const float* truth, const float* strike, int exp_flag,
                       float* ecdf, float* mean, float* sd, float* nll, float* payoff, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(samples, "volt_rollout_stats: null samples");
  VOLT_REQUIRE(B >= 1 && S >= 1 && H >= 1, "volt_rollout_stats: need B, S, H >= 1 (got %d, %d, %d)", B, S, H);
  VOLT_REQUIRE(!nll || truth, "volt_rollout_stats: nll needs truth");
  VOLT_REQUIRE(!payoff || strike, "volt_rollout_stats: payoff needs strike");
  VOLT_REQUIRE(!ecdf || truth, "volt_rollout_stats: ecdf needs truth");
  return launch_rollout_stats(samples, B, S, H, truth, strike, exp_flag, ecdf, mean, sd, nll, payoff, ST(stream));
}

int volt_gpcv_rows(const float* chol_var, const float* W, const float* var_mean, const float* y, const float* gh_t,
                   const float* gh_w, int nq, int B, int n, float inv_n, float* grad_chol, float* rows, void* stream) {
  if (B == 0) return VOLT_OK;   // empty batch (e.g. an empty shard): nothing to do
  VOLT_REQUIRE(chol_var && W && var_mean && y && gh_t && gh_w && grad_chol && rows, "volt_gpcv_rows: null pointer");
  VOLT_REQUIRE(B >= 1 && n >= 1 && nq >= 1 && nq <= 128, "volt_gpcv_rows: need B, n >= 1 and 1 <= nq <= 128 (got %d, %d, %d)", B, n, nq);
  return launch_gpcv_rows(chol_var, W, var_mean, y, gh_t, gh_w, nq, B, n, inv_n, grad_chol, rows, ST(stream));
}

int volt_gemm_nt(const float* A, long long lda, long long a_bstride, const float* B, long long ldb, long long b_bstride, float* C,
                 long long ldc, long long c_bstride, int M, int N, int K, int batch, int subtract, void* stream) {
  if (batch == 0 || M == 0 || N == 0) return VOLT_OK;
  VOLT_REQUIRE(A && B && C, "volt_gemm_nt: null pointer");
  VOLT_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "volt_gemm_nt: bad shape (M=%d N=%d K=%d batch=%d)", M, N, K, batch);
  VOLT_REQUIRE(lda >= K && ldb >= K && ldc >= N, "volt_gemm_nt: row strides smaller than the rows");
  VOLT_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
               "volt_gemm_nt: operands must be 16-byte aligned");
  return launch_gemm_nt(A, lda, a_bstride, B, ldb, b_bstride, C, ldc, c_bstride, M, N, K, batch, subtract ? 1 : 0, 0, 0, ST(stream));
}

int volt_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long count, float lr, float beta1,
                   float beta2, float eps, int step, const float* step_dev, void* stream) {
  VOLT_REQUIRE(param && grad && exp_avg && exp_avg_sq, "volt_adam_step: null pointer");
  VOLT_REQUIRE(count >= 1 && (step >= 1 || step_dev), "volt_adam_step: need count >= 1 and step >= 1 (or a device step counter)");
  return launch_adam(param, grad, exp_avg, exp_avg_sq, count, lr, beta1, beta2, eps, step < 1 ? 1 : step, step_dev, ST(stream));
}

}  // extern "C"
