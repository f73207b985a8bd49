"""voltron.option_utils mirror -- the sample-based evaluation helpers of the reference that consume a rollout tensor
(voltron/option_utils.py:28-52).  The pandas bookkeeping of `Pricer` (option chains, dates) is outside the hot path; its
two reductions are exposed directly: `ECDF` (same name and meaning as the reference) and `CallValuation`.
Both run the one-pass CUDA reduction `volt_rollout_stats`; many series at once go through `volt_b200.ops.rollout_stats`."""
import torch

from . import ops


def ECDF(sample_pxs, true_px):
    """Fraction of sampled prices whose log is below the log of the realised price -- option_utils.py:48-52.
    sample_pxs (S,) prices, true_px 0-dim tensor or float.  Returns a Python float like the reference."""
    smp = torch.as_tensor(sample_pxs, dtype=torch.float32).log().reshape(-1, 1)   # (S, H=1); the sort is irrelevant to a count
    log_px = torch.as_tensor(true_px, dtype=torch.float32).log().reshape(1, 1)
    return ops.rollout_stats(smp, truth=log_px)["ecdf"].item()


def CallValuation(mc_pxs, strike):
    """Monte-Carlo value of a call per expiry: mean(max(px - K, 0)) over the draws -- option_utils.py:37.
    mc_pxs (S,E) prices, strike float or (E,).  Returns an (E,) tensor on the CPU."""
    px = torch.as_tensor(mc_pxs, dtype=torch.float32)
    if px.dim() == 1:
        px = px.unsqueeze(-1)
    k = torch.as_tensor(strike, dtype=torch.float32).reshape(-1).expand(px.shape[-1])
    return ops.rollout_stats(px, strike=k)["payoff"][0].cpu()
