"""Series-sharded loss exchange check (run under torchrun on >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/exchange_check.py
Every rank runs steps of batched.mll_and_grad on its own shard; the loss delivered by the in-kernel peer push
(volt_mll_step_sharded + volt_loss_gather) must equal, bit for bit on every rank, the rank-ordered sum of the partials
gathered with NCCL, with losses waited late / out of order / not at all, and with an empty shard on the last rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volt_b200 import batched, ops  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
    torch.cuda.synchronize()
    os.dup2(saved, 1)

    T = 400
    ok = True
    for B in (96, 5):
        Bl = 0 if (rank == world - 1 and B == 5) else B          # empty shard on the last rank in the second round
        x, vol, logy = batched.synth_series(max(Bl, 1), T, 1.0 / 252, start=rank * B)
        xd, vd, yd = x.to(dev), vol.to(dev)[:Bl], logy.to(dev)[:Bl]
        resid = ops.ma_mean("ewma", yd, 20, want_resid=True)[1] if Bl else yd
        held = []
        for s in range(27):
            raw = torch.full((Bl,), -3.0 + 0.05 * s, device=dev)
            out = batched.mll_and_grad(xd, vd, resid, raw)
            assert isinstance(out["loss"], batched._ExchangedLoss), "peer exchange not active"
            parts = [torch.zeros(1, device=dev) for _ in range(world)]
            dist.all_gather(parts, out["partial_loss"].reshape(1))
            want = torch.zeros((), device=dev)
            for p in parts:
                want = want + p.reshape(())
            held.append((out["loss"], want))
            if s >= 21:                                          # training-loop pattern: read the previous step's loss
                if len(held) > 1:
                    l, w = held.pop(0)
                    if not torch.equal(l.wait(), w):
                        ok = False
                        print(f"rank {rank} B {B} step {s} (loop): got {float(l.wait())!r} want {float(w)!r}", file=sys.stderr)
            elif 12 < s < 20:                                    # hold 8 un-waited losses across the reuse of their entries
                pass
            elif s % 3 == 0 or s == 20:                          # wait late and newest-first
                for l, w in reversed(held):
                    got = l.wait()
                    if not torch.equal(got, w):
                        ok = False
                        print(f"rank {rank} B {B} step {s}: got {float(got)!r} want {float(w)!r}", file=sys.stderr)
                held = []
        torch.cuda.synchronize()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("EXCHANGE_OK" if float(flag) == 1.0 else "EXCHANGE_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
