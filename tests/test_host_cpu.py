"""CPU-side tests (no GPU needed): the C-ABI library loads and exports every symbol include/volt_b200.h declares, the
host mirror keeps the reference's parameter order / grad flags, the product path fails loudly without a GPU and never
imports the oracle, and the multi-rank sharding logic works under a world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "volt_b200.h")
LIB = os.path.join(ROOT, "volt_b200", "csrc", "libvolt_b200.so")


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    return LIB


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(volt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_and_library_exports_same_symbols(built):
    syms = header_symbols()
    assert len(syms) >= 18
    lib = ctypes.CDLL(built)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"libvolt_b200.so does not export {missing}"
    from volt_b200 import _lib

    assert sorted(_lib.EXPORTED) == syms
    lib.volt_abi_version.restype = ctypes.c_int
    assert lib.volt_abi_version() == 1


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    for path in ("voltron/kernels/VolKernel.py", "voltron/kernels/BMKernel.py", "voltron/means/EWMA.py",
                 "voltron/train_utils.py", "voltron/rollout_utils.py", "voltron/models/BMGP.py"):
        assert path in src


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "volt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
    assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(ROOT, "voltron", "__init__.py")).read(), flags=re.M)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_ops_fail_loudly_without_gpu(built):
    import volt_b200
    from volt_b200._lib import VoltLibraryError

    x = torch.arange(8) / 252.0
    with pytest.raises(VoltLibraryError):
        volt_b200.ops.vol_cov(x, torch.ones(8))
    with pytest.raises(VoltLibraryError):
        volt_b200.ops.ewma(torch.ones(8), 3)
    with pytest.raises(VoltLibraryError):
        volt_b200.VolatilityKernel()(x, torch.ones(8)).evaluate()


def test_missing_library_is_an_error(tmp_path, built):
    from volt_b200 import _lib

    saved, saved_lib = _lib.LIB_PATH, _lib._lib
    try:
        _lib.LIB_PATH, _lib._lib = str(tmp_path / "nope.so"), None
        with pytest.raises(_lib.VoltLibraryError):
            _lib.load()
    finally:
        _lib.LIB_PATH, _lib._lib = saved, saved_lib


@pytest.mark.parametrize("mean_func,flags", [("ewma", [True, False, False]), ("dewma", [True, False, False]),
                                              ("tewma", [True, False, False]), ("meanrevert", [True, False, False]),
                                              ("constant", [True, True, False, False]),
                                              ("loglinear", [True, True, True, False, False]),
                                              ("linear", [True, True, True, False, False])])
def test_parameter_order_and_grad_flags(golden, mean_func, flags):
    """Registration order [likelihood, mean_module, covar_module, vol_lh, vol_model] and the positional grad_flags of
    train_utils.py:201-227 (no GPU work: train_iters=0 only builds the model)."""
    import volt_b200 as vb

    g = golden["train_volt_ewma"]
    vmod, vlh = vb.TrainVolModel(g["x"], g["vol"], train_iters=0)
    volt, lh = vb.TrainVoltMagpieModel(g["x"], g["px"][1:], vmod, vlh, g["vol"], train_iters=0, k=g["k"], mean_func=mean_func)
    assert [p.requires_grad for p in volt.parameters()] == flags
    names = [n for n, _ in volt.named_parameters()]
    assert names[0] == "likelihood.noise_covar.raw_noise" and names[-1] == "vol_model.covar_module.raw_vol"
    if mean_func == "ewma":
        assert names == g["param_names"]
    # raw_noise := 1e-5 (RAW value, train_utils.py:222) -> noise = softplus(1e-5) + 1e-4
    assert float(lh.raw_noise) == pytest.approx(1e-5)
    assert float(lh.noise) == pytest.approx(0.69325, abs=1e-4)
    # TrainVolModel: `vol_lh.noise.data = ...` is a no-op in the reference (:71): raw_noise stays 0, vol = 0.2
    assert float(vlh.raw_noise) == 0.0
    assert float(vmod.covar_module.vol) == pytest.approx(0.2, abs=1e-6)


def test_data_model_parameters():
    import volt_b200 as vb

    x = torch.arange(16) / 252.0
    px = torch.linspace(10, 11, 16)
    vol = torch.full((16,), 0.2)
    vmod, vlh = vb.TrainVolModel(x, vol, train_iters=0)
    m, lh = vb.TrainDataModel(x, px, vmod, vlh, vol, train_iters=0)
    assert [n for n, _ in m.named_parameters()] == ["likelihood.noise_covar.raw_noise", "mean_module.weights", "mean_module.bias",
                                                   "vol_lh.noise_covar.raw_noise", "vol_model.covar_module.raw_vol"]
    assert [p.requires_grad for p in m.parameters()] == [True, True, True, False, False]
    assert float(m.mean_module.bias) == pytest.approx(float(px.mean()), rel=1e-6)  # initialize_from_data


def test_voltron_alias_exposes_reference_names():
    import voltron
    from voltron.kernels import BMKernel, VolatilityKernel  # noqa: F401
    from voltron.means import DEWMAMean, EWMAMean, LogLinearMean, MeanRevertingEMAMean, TEWMAMean  # noqa: F401
    from voltron.models import BMGP, VoltMagpie, VoltronGP  # noqa: F401
    from voltron.rollout_utils import GeneratePrediction, Rollouts  # noqa: F401
    from voltron.train_utils import TrainDataModel, TrainVolModel, TrainVoltMagpieModel  # noqa: F401

    assert voltron.Rollouts is Rollouts


def test_shard_bounds_cover_everything():
    from volt_b200.batched import shard_bounds

    for n in (1, 7, 1024, 4096, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synth_series_matches_oracle_generator():
    from oracle import volt_oracle as O
    from volt_b200.batched import synth_series

    a, b = synth_series(3, 64, start=2), O.synth_series(5, 64)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1][2:]) and torch.equal(a[2], b[2][2:])


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from volt_b200 import batched
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = batched.dist_info()
B = 37
lo, hi = batched.shard_bounds(B, rank, world)
per_series = torch.arange(B, dtype=torch.float64) * 0.5 - 3.0          # stand-in for the per-series -MLL
loss = batched.all_reduce_sum(per_series[lo:hi].sum().reshape(1).clone())
assert abs(float(loss) - float(per_series.sum())) < 1e-12, (float(loss), float(per_series.sum()))
counts = batched.all_reduce_sum(torch.tensor([hi - lo], dtype=torch.float64))
assert int(counts) == B
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_series_sharding_under_gloo_world2(tmp_path):
    """N > 1 path on CPU: two ranks own contiguous blocks of series; the only collective is the scalar-loss all-reduce."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bench_reference_arm_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json

    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


# ---------------------------------------------------------------------------------------------------------------------------
# bookkeeping of the multi-GPU loss exchange (volt_b200.batched.LossExchange) with the device parts replaced: which step's
# total a loss object reads, when it is copied out of the ring, and that dropped losses cost nothing
# ---------------------------------------------------------------------------------------------------------------------------
class _FakeExchange:
    """LossExchange with CPU tensors: the 'kernel' of step s writes totals[(s - LAG) % RING]; the gather returns the truth."""

    def __new__(cls, truth):
        from volt_b200 import batched

        class Fake(batched.LossExchange):
            def __init__(self, truth):
                self.world, self.rank, self.device = 2, 0, torch.device("cpu")
                self.slots = torch.zeros(self.RING * self.world, dtype=torch.int64)
                self.peers, self.side, self.push_mode = 1, None, "kernel"
                self.totals = torch.zeros(self.RING + 1)
                self.seq, self._keep, self._pending, self._launched = 0, [], [], None
                self.truth, self.gathers, self.events = truth, [], 0

            def _record_event(self):
                self.events += 1
                return self.events

            def _wait_event(self, ev):
                assert ev <= self.events

            def _total(self, loss):
                self._pending = [r for r in self._pending if r() is not None and r() is not loss]
                self.gathers.append(loss._seq)
                return torch.tensor(self.truth[loss._seq])

            def step(self):
                desc, loss = self.next()
                peers, mine, totals, lag, world, rank, ring, seq = desc
                assert (totals is None) == (seq <= lag) and seq == self.seq
                if totals is not None:                       # what volt_mll_step_sharded does on the device
                    totals[(seq - lag) % ring] = self.truth[seq - lag]
                self.launched()
                return loss

        return Fake(truth)


def test_loss_exchange_bookkeeping_training_loop_pattern():
    truth = {s: float(100 + s) for s in range(1, 60)}
    ex = _FakeExchange(truth)
    held = []
    for s in range(1, 40):
        held.append((s, ex.step()))
        if len(held) > ex.LAG:                               # read the loss of LAG steps ago: never needs a gather
            t, l = held.pop(0)
            assert float(l.wait()) == truth[t]
    assert ex.gathers == []
    for t, l in held:                                        # the newest LAG steps do
        assert float(l.wait()) == truth[t]
    assert ex.gathers == list(range(40 - ex.LAG, 40))


def test_loss_exchange_bookkeeping_late_and_dropped_losses():
    truth = {s: float(7 * s) for s in range(1, 80)}
    ex = _FakeExchange(truth)
    kept = {}
    for s in range(1, 41):
        loss = ex.step()
        if s % 3 == 0:
            kept[s] = loss                                   # held across many reuses of its ring entry; the others are dropped
        if s == 20:
            assert float(loss.wait()) == truth[20] and ex.gathers == [20]   # newest step, read at once: one gather
    for s, l in kept.items():
        assert float(l.wait()) == truth[s], s                # copied out before its entry was overwritten
    # step 39 needs a gather only if it is one of the newest LAG steps (40 was the last one launched); nothing else does
    assert ex.gathers == ([20, 39] if ex.LAG >= 2 else [20])
    assert len(ex._pending) <= ex.RING                       # dropped losses do not accumulate
