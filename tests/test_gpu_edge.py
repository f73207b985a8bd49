"""GPU edge cases of the hot path through the C ABI (`pytest -m gpu`): degenerate shapes (empty batch, T = 1, T = 2),
many short series (grid-stride over the persistent CTAs), batched / 2-D time grids, NaN inputs, the diag / batch-mode
kernel calls of the GPyTorch protocol.  Everything is compared with the CPU oracle on the same inputs."""
import warnings

import pytest
import torch

from oracle import volt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vb():
    import volt_b200

    volt_b200._lib.require_device()
    return volt_b200


def relerr(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64).cpu(), torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_empty_batch_is_a_no_op(vb):
    """B = 0: every entry point returns empty outputs without launching (ragged shards of a small job can be empty)."""
    T = 16
    x = torch.arange(T, dtype=torch.float32) / 252
    e = torch.empty(0, T)
    K = vb.ops.vol_cov(x.cuda(), e.cuda())
    assert tuple(K.shape) == (0, T, T)
    out = vb.batched.mll_and_grad(x.cuda(), e.cuda(), e.cuda(), torch.empty(0).cuda())
    assert out["mll"].numel() == 0 and out["alpha"].shape == (0, T) and float(out["loss"]) == 0.0
    V = vb.ops.cumtrapz(e.cuda(), x.cuda())
    assert tuple(V.shape) == (0, T)
    m = vb.ops.ewma(e.cuda(), 5)
    assert tuple(m.shape) == (0, T + 1)


@pytest.mark.parametrize("T", [2, 3])
def test_tiny_series(vb, T):
    """The shortest series the reference accepts (CumTrapz needs two grid points, VolKernel.py:5)."""
    B = 4
    x, vol, logy = O.synth_series(B, T, seed=5)
    resid = logy - logy.mean(-1, keepdim=True)
    raw = torch.tensor([1e-5, -2.0, 0.5, 2.0])
    K = vb.ops.vol_cov(x.cuda(), vol.cuda()).cpu()
    assert torch.equal(K, O.vol_kernel(x, vol))
    out = vb.batched.mll_and_grad(x.cuda(), vol.cuda(), resid.cuda(), raw.cuda(), check=True)
    for b in range(B):
        ref = O.volt_mll_and_grad(x.double(), vol[b].double(), resid[b].double(), raw[b].double())
        assert relerr(out["mll"][b], ref["mll"]) < 1e-4
        assert relerr(out["draw_noise"][b], ref["draw_noise"]) < 2e-3
        assert relerr(out["alpha"][b], ref["alpha"]) < 2e-3


def test_single_point_dense_mll(vb):
    """T = 1 through the dense entry point: MLL = -0.5 (r^2 / a + log a + log 2 pi), d/da closed form."""
    Kd = torch.tensor([[[2.0]], [[0.5]]])
    r = torch.tensor([[0.3], [-1.2]])
    noise = torch.tensor([0.1, 0.7])
    out = vb.ops.mll_grad("dense", None, Kd.cuda(), r.cuda(), noise.cuda())
    a = Kd.reshape(-1) + noise
    mll = -0.5 * (r.reshape(-1) ** 2 / a + a.log() + torch.log(torch.tensor(2 * torch.pi)))
    dmll = 0.5 * (r.reshape(-1) ** 2 / a ** 2 - 1 / a)
    assert relerr(out["scalars"][:, 0], mll) < 1e-5
    assert relerr(out["scalars"][:, 1], dmll) < 1e-5
    assert relerr(out["alpha"].reshape(-1), r.reshape(-1) / a) < 1e-5
    assert int(out["info"].abs().sum()) == 0


def test_many_short_series_grid_stride(vb):
    """5000 series of 40 points: ~11 series per persistent CTA; spot-check against the oracle and batch == loop."""
    B, T = 5000, 40
    x, vol, logy = O.synth_series(B, T, seed=11)
    resid = logy - logy.mean(-1, keepdim=True)
    raw = torch.linspace(-3.0, 2.0, B)
    out = vb.batched.mll_and_grad(x.cuda(), vol.cuda(), resid.cuda(), raw.cuda(), check=True)
    assert int(out["info"].abs().sum()) == 0
    for b in (0, 443, 444, 887, 888, 2500, 4999):
        ref = O.volt_mll_and_grad(x.double(), vol[b].double(), resid[b].double(), raw[b].double())
        assert relerr(out["mll"][b], ref["mll"]) < 1e-4
        assert relerr(out["draw_noise"][b], ref["draw_noise"]) < 2e-3
        one = vb.batched.mll_and_grad(x.cuda(), vol[b:b + 1].cuda(), resid[b:b + 1].cuda(), raw[b:b + 1].cuda())
        assert torch.equal(one["mll"][0], out["mll"][b]) and torch.equal(one["alpha"][0], out["alpha"][b])


def test_batched_time_grid(vb):
    """CumTrapz takes dx per row when x is 2-D (VolKernel.py:5-8): series with different sampling steps in one batch."""
    B, T = 3, 130
    _, vol, logy = O.synth_series(B, T, seed=3)
    dts = torch.tensor([1 / 252, 1 / 365, 1 / 52])
    x = torch.arange(T, dtype=torch.float32)[None, :] * dts[:, None]
    K = vb.ops.vol_cov(x.cuda(), vol.cuda()).cpu()
    assert torch.equal(K, O.vol_kernel(x, vol))
    resid = logy - logy.mean(-1, keepdim=True)
    noise = torch.tensor([0.3, 0.05, 1.0])
    out = vb.ops.mll_grad("vol", x.cuda(), vol.cuda(), resid.cuda(), noise.cuda())
    for b in range(B):
        ref = O.exact_mll(O.vol_kernel(x[b].double(), vol[b].double()), resid[b].double(), noise[b].double())
        assert relerr(out["scalars"][b, 0], ref) < 1e-4


def test_nan_input_is_flagged_not_hidden(vb):
    """A NaN in one series fails that series' factorisation (info > 0, like cholesky_ex) and leaves the others exact."""
    B, T = 4, 96
    x, vol, logy = O.synth_series(B, T, seed=9)
    vol[2, 17] = float("nan")
    resid = logy - logy.mean(-1, keepdim=True)
    raw = torch.zeros(B)
    out = vb.batched.mll_and_grad(x.cuda(), vol.cuda(), resid.cuda(), raw.cuda(), check=False)
    info = out["info"].cpu()
    assert int(info[2]) > 0 and int(info[[0, 1, 3]].abs().sum()) == 0
    for b in (0, 1, 3):
        ref = O.volt_mll_and_grad(x.double(), vol[b].double(), resid[b].double(), raw[b].double())
        assert relerr(out["mll"][b], ref["mll"]) < 1e-4
    with pytest.raises(Exception):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vb.batched.mll_and_grad(x.cuda(), vol.cuda(), resid.cuda(), raw.cuda(), check=True)


def test_kernel_protocol_diag_and_shapes(vb):
    """VolatilityKernel through the GPyTorch call protocol: (T,1) inputs, diag=True, batch inputs, .evaluate()."""
    T = 77
    x, vol, _ = O.synth_series(2, T, seed=21)
    k = vb.VolatilityKernel()
    K = k(x.cuda().unsqueeze(-1), vol[0].cuda().unsqueeze(-1)).evaluate().cpu()
    assert torch.equal(K, O.vol_kernel(x, vol[0]))
    d = k(x.cuda().unsqueeze(-1), vol[0].cuda().unsqueeze(-1), diag=True)
    d = d.evaluate() if hasattr(d, "evaluate") else d
    assert torch.equal(d.cpu().reshape(-1), O.vol_kernel(x, vol[0], diag=True))
    Kb = k(x.cuda().expand(2, T).unsqueeze(-1), vol.cuda().unsqueeze(-1)).evaluate().cpu()
    assert torch.equal(Kb, O.vol_kernel(x, vol))


@pytest.mark.parametrize("B,T,pinned", [(700, 64, True), (1024, 128, False), (400, 96, True), (300, 100, True), (5, 512, False),
                                         (1, 2048, True)])
def test_host_buffer_entry_equals_device_entry(vb, B, T, pinned):
    """volt_mll_grad_vol_host (host pointers, copies inside the call; for large batches one kernel launch whose later
    series wait for an arrival flag while their inputs are still being copied) == the device-pointer entry, bit for bit."""
    lib = vb._lib.load()
    x, vol, logy = vb.batched.synth_series(B, T)
    resid = (logy - logy.mean(-1, keepdim=True)).contiguous()
    noise = torch.linspace(0.05, 1.5, B)
    ref = vb.ops.mll_grad("vol", x.cuda(), vol.cuda(), resid.cuda(), noise.cuda(), check=False)
    torch.cuda.synchronize()
    pin = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    hx, hv, hr, hn = pin(x.contiguous()), pin(vol.contiguous()), pin(resid), pin(noise)
    hs, ha, hi = pin(torch.empty(B, 16)), pin(torch.empty(B, T)), pin(torch.empty(B, dtype=torch.int32))
    for _ in range(3):   # repeated calls reuse the staging buffers and the flag
        hs.zero_(); ha.zero_()
        vb._lib.check(lib.volt_mll_grad_vol_host(hx.data_ptr(), hv.data_ptr(), hr.data_ptr(), hn.data_ptr(), 1, B, T, 1e-6, 3,
                                                  hs.data_ptr(), ha.data_ptr(), hi.data_ptr()), "volt_mll_grad_vol_host")
        assert torch.equal(hs[:, :8], ref["scalars"][:, :8].cpu())
        assert torch.equal(ha, ref["alpha"].cpu())
        assert int(hi.abs().sum()) == 0


def test_one_process_two_devices(vb):
    """Function attributes (opt-in shared memory), helper streams and workspaces are per device: the same process can
    drive a second GPU (skipped on a one-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    B, T = 6, 320
    x, vol, logy = O.synth_series(B, T, seed=4)
    resid = logy - logy.mean(-1, keepdim=True)
    raw = torch.linspace(-2.0, 1.0, B)
    outs = []
    for d in (0, 1, 0):
        with torch.cuda.device(d):
            dev = torch.device("cuda", d)
            o = vb.batched.mll_and_grad(x.to(dev), vol.to(dev), resid.to(dev), raw.to(dev), check=True)
            xl, vl, ll = O.synth_series(1, 2048, seed=6)                      # multi-CTA path: its own streams / attributes
            ol = vb.batched.mll_and_grad(xl.to(dev), vl.to(dev), (ll - ll.mean()).to(dev), torch.zeros(1, device=dev))
            pv = vol[:, -1:, None].expand(B, 64, 5).contiguous().to(dev)
            r, _, _ = vb.ops.rollout(x.to(dev), logy.to(dev), vol.to(dev), pv, eps=torch.ones(B, 64, 5, device=dev), k=5)
            hs = torch.empty(B, 16).pin_memory()
            hv, hr, hn = vol.contiguous(), resid.contiguous(), vb.batched.noise_from_raw(raw).contiguous()   # kept alive over the call
            vb._lib.check(vb._lib.load().volt_mll_grad_vol_host(x.data_ptr(), hv.data_ptr(), hr.data_ptr(), hn.data_ptr(), 1, B, T, 1e-6, 3,
                                                                hs.data_ptr(), None, None), "volt_mll_grad_vol_host")
            outs.append((o["mll"].cpu(), ol["mll"].cpu(), r.cpu(), hs[:, 0].clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    for a, b in zip(outs[0], outs[2]):
        assert torch.equal(a, b)
    assert torch.equal(outs[0][3], outs[0][0])
