// TORCH_LIBRARY(volt, ...) registration of the hot-path operators ("Extension ops"): a thin shim over the
// plain-C ABI of libvolt_b200.so -- every op below only checks / allocates tensors and forwards device pointers plus the
// current CUDA stream to the C entry point.  Built by __graft_entry__.build() into libvolt_torch.so (g++, no nvcc) and
// loaded with torch.ops.load_library (volt_b200/torch_ops.py); callable as torch.ops.volt.<name>.
//
//   vol_cov(x, vol, add_diag?)                 VolatilityKernel.forward              voltron/kernels/VolKernel.py:18-41
//   bm_cov(x1, x2, vol)                        BMKernel.forward                      voltron/kernels/BMKernel.py:38-51
//   ewma(y, k, mode)                           EWMA / EWMAMean family                voltron/means/EWMA.py:20-135
//   potrf_(A) -> info                          torch.linalg.cholesky_ex in place     voltron/rollout_utils.py:35,46
//   mll_fwd_bwd(x, vol, resid, noise)          mll(...); loss.backward()             voltron/train_utils.py:247-250
//   gp_predict(L, Kx, r)                       cholesky_solve posterior pieces       voltron/rollout_utils.py:36-44
//   rollout(...)                               Rollouts                              voltron/rollout_utils.py:57-93
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>

#include "../../include/volt_b200.h"

namespace {

void check(int status, const char* what) {
  TORCH_CHECK(status == VOLT_OK, what, " failed with status ", status, ": ", volt_last_error());
}

at::Tensor f32c(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (volt_b200 has no CPU path)");
  return t.to(at::kFloat).contiguous();
}

void* stream_of(const at::Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

at::Tensor vol_cov(const at::Tensor& x, const at::Tensor& vol, std::optional<double> add_diag) {
  c10::cuda::CUDAGuard guard(vol.device());
  const auto xd = f32c(x, "x");
  const int64_t T = vol.size(-1);
  const auto vd = f32c(vol, "vol").reshape({-1, T});
  const int64_t B = vd.size(0);
  TORCH_CHECK(xd.dim() == 1 && xd.size(0) == T, "vol_cov: x must be (T,)");
  at::Tensor ad;
  if (add_diag.has_value()) ad = at::full({1}, *add_diag, vd.options());
  auto K = at::empty({B, T, T}, vd.options());
  check(volt_vol_cov(xd.data_ptr<float>(), 0, vd.data_ptr<float>(), VOLT_VOL_SIGMA, (int)B, (int)T,
                     add_diag.has_value() ? ad.data_ptr<float>() : nullptr, 0, K.data_ptr<float>(), stream_of(vd)),
        "volt_vol_cov");
  auto shape = vol.sizes().vec();
  shape.push_back(T);
  return K.reshape(shape);
}

at::Tensor bm_cov(const at::Tensor& x1, const at::Tensor& x2, const at::Tensor& vol) {
  c10::cuda::CUDAGuard guard(x1.device());
  const auto a = f32c(x1, "x1").reshape({-1}), b = f32c(x2, "x2").reshape({-1}), v = f32c(vol, "vol").reshape({-1});
  auto K = at::empty({a.numel(), b.numel()}, a.options());
  check(volt_bm_cov(a.data_ptr<float>(), (int)a.numel(), b.data_ptr<float>(), (int)b.numel(), v.data_ptr<float>(),
                    K.data_ptr<float>(), stream_of(a)),
        "volt_bm_cov");
  return K;
}

// mode: VOLT_MA_EWMA / DEWMA / TEWMA (the mean-reverting variant needs its latent mean: use the Python wrapper)
at::Tensor ewma(const at::Tensor& y, int64_t k, int64_t mode) {
  c10::cuda::CUDAGuard guard(y.device());
  const int64_t T = y.size(-1);
  const auto yd = f32c(y, "y").reshape({-1, T});
  const int64_t S = yd.size(0);
  TORCH_CHECK(mode >= VOLT_MA_EWMA && mode <= VOLT_MA_TEWMA, "ewma: mode must be 0 (EWMA), 1 (DEWMA) or 2 (TEWMA)");
  auto out = at::empty({S, T + 1}, yd.options());
  check(volt_ma_mean(yd.data_ptr<float>(), (int)S, (int)T, (int)k, (int)mode, 0.f, nullptr, out.data_ptr<float>(), nullptr, nullptr,
                     nullptr, stream_of(yd)),
        "volt_ma_mean");
  auto shape = y.sizes().vec();
  shape.back() = T + 1;
  return out.reshape(shape);
}

// In place: the lower triangle of A (B,T,T) is replaced by its Cholesky factor (strict upper triangle zeroed); returns info (B)
// with torch.linalg.cholesky_ex semantics.  No jitter: the psd_safe_cholesky policy lives in the caller, as in the reference.
at::Tensor potrf_(at::Tensor A) {
  c10::cuda::CUDAGuard guard(A.device());
  TORCH_CHECK(A.is_cuda() && A.scalar_type() == at::kFloat && A.is_contiguous() && A.dim() >= 2, "potrf_: contiguous float32 CUDA tensor");
  const int64_t T = A.size(-1);
  TORCH_CHECK(A.size(-2) == T, "potrf_: square matrices");
  const int64_t B = A.numel() / (T * T);
  auto info = at::empty({B}, A.options().dtype(at::kInt));
  check(volt_potrf(A.data_ptr<float>(), T * T, (int)T, nullptr, 0, (int)B, (int)T, 0.f, 0, A.data_ptr<float>(), T * T, (int)T, nullptr,
                   info.data_ptr<int>(), stream_of(A)),
        "volt_potrf");
  return info;
}

// One exact MLL + gradient evaluation per series: (mll[B], dMLL/dnoise[B], alpha[B,T], logdet[B]).
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> mll_fwd_bwd(const at::Tensor& x, const at::Tensor& vol, const at::Tensor& resid,
                                                                       const at::Tensor& noise) {
  c10::cuda::CUDAGuard guard(resid.device());
  const int64_t T = resid.size(-1);
  const auto r = f32c(resid, "resid").reshape({-1, T});
  const int64_t B = r.size(0);
  auto v = f32c(vol, "vol").reshape({-1, T});
  if (v.size(0) != B) v = v.expand({B, T}).contiguous();
  const auto xd = f32c(x, "x");
  TORCH_CHECK(xd.dim() == 1 && xd.size(0) == T, "mll_fwd_bwd: x must be (T,)");
  const auto nz = f32c(noise, "noise").reshape({-1});
  TORCH_CHECK(nz.numel() == 1 || nz.numel() == B, "mll_fwd_bwd: noise must be a scalar or (B,)");
  auto scal = at::empty({B, VOLT_NSCALARS}, r.options());
  auto alpha = at::empty({B, T}, r.options());
  auto info = at::empty({B}, r.options().dtype(at::kInt));
  check(volt_mll_grad_vol(xd.data_ptr<float>(), 0, v.data_ptr<float>(), VOLT_VOL_SIGMA, r.data_ptr<float>(), nz.data_ptr<float>(),
                          nz.numel() == 1 ? 0 : 1, (int)B, (int)T, 1e-6f, 3, scal.data_ptr<float>(), alpha.data_ptr<float>(),
                          info.data_ptr<int>(), stream_of(r)),
        "volt_mll_grad_vol");
  return {scal.select(1, VOLT_S_MLL).contiguous(), scal.select(1, VOLT_S_DNOISE).contiguous(), alpha,
          scal.select(1, VOLT_S_LOGDET).contiguous()};
}

// Posterior pieces from a Cholesky factor L (B,T,T) of K_tr: mean = Kx^T K_tr^-1 r (B,H) and the covariance reduction
// Kx^T K_tr^-1 Kx (B,H,H), with Kx (B,T,H), r (B,T)  (rollout_utils.py:36,44: two cholesky_solve calls).
std::tuple<at::Tensor, at::Tensor> gp_predict(const at::Tensor& L, const at::Tensor& Kx, const at::Tensor& r) {
  c10::cuda::CUDAGuard guard(L.device());
  const int64_t T = L.size(-1);
  const auto Ld = f32c(L, "L").reshape({-1, T, T});
  const int64_t B = Ld.size(0), H = Kx.size(-1);
  auto W = at::cat({f32c(Kx, "Kx").reshape({B, T, H}), f32c(r, "r").reshape({B, T, 1})}, 2).contiguous();
  check(volt_potrs(Ld.data_ptr<float>(), T * T, (int)T, (int)B, (int)T, W.data_ptr<float>(), T * (H + 1), (int)(H + 1), 1, stream_of(Ld)),
        "volt_potrs");
  const auto Wk = W.narrow(2, 0, H), v = W.narrow(2, H, 1);
  return {at::matmul(Wk.transpose(1, 2), v).squeeze(-1), at::matmul(Wk.transpose(1, 2), Wk)};
}

// Rollouts for B series x S draws x H steps with an EWMA(k) mean (the shipped drivers' configuration).
// x_train (n), y_train (B,n) log prices, vol_train (B,n) sigma, test_x (H) (only its length enters: uniform grid),
// pred_vol (B,S,H), eps (B,S,H) or None (in-kernel Philox with `seed`); theta / latent_mean: rollout_utils.py:41-42.
at::Tensor rollout(const at::Tensor& x_train, const at::Tensor& y_train, const at::Tensor& vol_train, const at::Tensor& test_x,
                   const at::Tensor& pred_vol, const std::optional<at::Tensor>& eps, int64_t k, std::optional<double> theta,
                   const std::optional<at::Tensor>& latent_mean, int64_t seed) {
  c10::cuda::CUDAGuard guard(pred_vol.device());
  const auto xd = f32c(x_train, "x_train").reshape({-1});
  const int64_t n = xd.numel();
  const auto ly = f32c(y_train, "y_train").reshape({-1, n});
  const int64_t B = ly.size(0);
  auto vd = f32c(vol_train, "vol_train").reshape({-1, n});
  if (vd.size(0) != B) vd = vd.expand({B, n}).contiguous();
  const int64_t H = test_x.numel();
  const auto pv = f32c(pred_vol, "pred_vol").reshape({B, -1, H});
  const int64_t S = pv.size(1);
  at::Tensor ep, lat;
  if (eps.has_value()) ep = f32c(*eps, "eps").reshape({B, S, H});
  const bool use_theta = theta.has_value() && latent_mean.has_value();
  if (use_theta) {
    lat = f32c(*latent_mean, "latent_mean").reshape({-1});
    if (lat.numel() == 1) lat = lat.expand({B}).contiguous();
  }
  auto out = at::empty({B, S, H}, pv.options());
  auto dinfo = at::empty({B, S}, pv.options().dtype(at::kInt));
  auto sinfo = at::empty({B}, pv.options().dtype(at::kInt));
  check(volt_rollout(xd.data_ptr<float>(), ly.data_ptr<float>(), vd.data_ptr<float>(), VOLT_VOL_SIGMA, pv.data_ptr<float>(),
                     eps.has_value() ? ep.data_ptr<float>() : nullptr, (int)B, (int)n, (int)S, (int)H, VOLT_MA_EWMA, (int)k, 0.5f, nullptr,
                     nullptr, nullptr, use_theta ? 1 : 0, use_theta ? (float)*theta : 0.f, use_theta ? lat.data_ptr<float>() : nullptr, 0,
                     1e-4f, (unsigned long long)seed, out.data_ptr<float>(), dinfo.data_ptr<int>(), sinfo.data_ptr<int>(), stream_of(pv)),
        "volt_rollout");
  return out;
}

}  // namespace

TORCH_LIBRARY(volt, m) {
  m.def("vol_cov(Tensor x, Tensor vol, float? add_diag=None) -> Tensor");
  m.def("bm_cov(Tensor x1, Tensor x2, Tensor vol) -> Tensor");
  m.def("ewma(Tensor y, int k, int mode=0) -> Tensor");
  m.def("potrf_(Tensor(a!) A) -> Tensor");
  m.def("mll_fwd_bwd(Tensor x, Tensor vol, Tensor resid, Tensor noise) -> (Tensor, Tensor, Tensor, Tensor)");
  m.def("gp_predict(Tensor L, Tensor Kx, Tensor r) -> (Tensor, Tensor)");
  m.def("rollout(Tensor x_train, Tensor y_train, Tensor vol_train, Tensor test_x, Tensor pred_vol, Tensor? eps=None, int k=25, "
        "float? theta=None, Tensor? latent_mean=None, int seed=0) -> Tensor");
}

TORCH_LIBRARY_IMPL(volt, CUDA, m) {
  m.impl("vol_cov", &vol_cov);
  m.impl("bm_cov", &bm_cov);
  m.impl("ewma", &ewma);
  m.impl("potrf_", &potrf_);
  m.impl("mll_fwd_bwd", &mll_fwd_bwd);
  m.impl("gp_predict", &gp_predict);
  m.impl("rollout", &rollout);
}
