"""Golden vectors for the evaluation reductions, produced by running the reference's OWN
voltron/option_utils.py (unchanged; it needs only numpy / torch / pandas) in the build container:

    python tests/golden/make_golden_eval.py        ->  tests/golden/eval_golden.pt

`Pricer` is driven with a tiny synthetic option chain so that both of its reductions -- the Monte-Carlo call valuation
(option_utils.py:37) and the sample percentile ECDF (:39, :48-52) -- come straight out of the reference code.
"""
import importlib.util
import os

import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_option_utils", "/root/reference/voltron/option_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def main():
    g = torch.Generator().manual_seed(7)
    G = {}
    # ---- ECDF on its own: generic, all below, all above, exact ties, a single draw
    cases = []
    for S, shift in ((64, 0.0), (257, 0.3), (50, -5.0), (50, 5.0), (1, 0.0)):
        px = torch.exp(2.3 + 0.2 * torch.randn(S, generator=g))
        true_px = torch.exp(torch.tensor(2.3 + shift))
        cases.append(dict(sample_pxs=px, true_px=true_px, ecdf=ref.ECDF(px, true_px)))
    px = torch.tensor([9.0, 10.0, 10.0, 11.0, 12.0])
    cases.append(dict(sample_pxs=px, true_px=torch.tensor(10.0), ecdf=ref.ECDF(px, torch.tensor(10.0))))
    G["ecdf"] = cases
    # ---- Pricer: S draws x E expiries against a small chain
    S, E = 200, 3
    mc = torch.exp(2.3 + 0.15 * torch.randn(S, E, generator=g).cumsum(1))
    edays = [pd.Timestamp("2020-01-17"), pd.Timestamp("2020-02-21"), pd.Timestamp("2020-03-20")]
    rows = []
    for e in edays:
        for K in (8.0, 10.0, 11.5):
            rows.append(dict(expiration=e, strike=K, bid=0.1, ask=0.2))
    options = pd.DataFrame(rows)
    true_pxs = torch.tensor([10.2, 9.1, 12.4])
    df = ref.Pricer(mc, options, edays, true_pxs, 10.0)
    G["pricer"] = dict(mc_pxs=mc, true_pxs=true_pxs, strikes=torch.tensor(df["Strike"].to_numpy(), dtype=torch.float32),
                       expiry_idx=torch.tensor([edays.index(pd.Timestamp(e)) for e in df["Expiry"]]),
                       valuation=torch.tensor(df["Voltron"].to_numpy(), dtype=torch.float64),
                       percentile=torch.tensor(df["Sample_Percentile"].to_numpy(), dtype=torch.float64))
    torch.save(G, os.path.join(HERE, "eval_golden.pt"))
    print("wrote eval_golden.pt:", len(cases), "ECDF cases,", len(df), "priced options")


if __name__ == "__main__":
    main()
