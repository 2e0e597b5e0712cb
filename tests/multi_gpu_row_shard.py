"""Row-sharded RMHMC/HMC across ranks (NCCL all-reduce of every build) == the unsharded run == the oracle.

Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
            --master-port 29533 tests/multi_gpu_row_shard.py
Every rank holds N/world rows of the same synthetic data set and runs ALL chains under the same host
tape; rank 0 also runs the unsharded problem on its own GPU.  FP64 sums are associated differently
(per-shard partials added by NCCL), so agreement is to ~1e-12, not bitwise.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import riemannhamiltonianmontecarlo_b200 as r  # noqa: E402
from oracle import blr_oracle as bo  # noqa: E402
from riemannhamiltonianmontecarlo_b200.engine import shard_rows  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    out = {}
    for dim, n_rows, tag in [(15, 1003, "d15"), (64, 1500, "d64")]:
        xx, t = r.datasets.synthetic_logistic(n_rows, dim, 5000 + dim)
        b, e = shard_rows(n_rows, rank, world)
        n_iter, burn, c = 6, 1, 5
        tapes = [bo.make_tape(n_iter, dim, 9300 + i) for i in range(c)]
        st = bo.stack_tapes(tapes)
        data = r.LogisticData(xx[b:e], t[b:e], device=f"cuda:{local}", row_shard=(rank, world))
        s = r.RMHMCSampler(data, c, 4, 0.4, 4)
        s.set_tape(st["z"], st["u_step"], st["z_dir"], st["u_acc"])
        s.set_samples(n_iter - burn, burn)
        s.run(n_iter)
        sharded = s.samples.cpu().numpy()
        acc = s.state()["accepted"]
        # HMC, sharded
        hs = r.HMCSampler(data, c, 15, 0.05)
        hs.set_tape(st["z"], st["u_step"], st["u_acc"])
        hs.set_samples(n_iter - burn, burn)
        hs.run(n_iter)
        hmc_sharded = hs.samples.cpu().numpy()
        data.close()
        # all ranks must hold identical results
        ref = torch.from_numpy(sharded).cuda()
        dist.broadcast(ref, src=0)
        same = bool((ref.cpu().numpy() == sharded).all())
        if rank == 0:
            # the oracle (CPU restatement of rmhmc.py / hmc.py) on the UNSHARDED data, same tape
            o_ref, o_infos = bo.rmhmc_chains(xx, t, tapes, n_iter=n_iter, burn_in=burn, n_leapfrog=4, step_size=0.4, n_fixed=4)
            h_ref, _ = bo.hmc_chains(xx, t, tapes, n_iter=n_iter, burn_in=burn, n_leapfrog=15, step_size=0.05)
            full, _, info = r.rmhmc_batched(xx, t, c, n_iter, burn, 4, 0.4, 4, draws=st, device=f"cuda:{local}")
            hfull, _, _ = r.hmc_batched(xx, t, c, n_iter, burn, 15, 0.05, draws=st, device=f"cuda:{local}")
            err = float(np.abs(sharded[:, 1:] - full[:, 1:]).max() / np.abs(full[:, 1:]).max())
            herr = float(np.abs(hmc_sharded[:, 1:] - hfull[:, 1:]).max() / np.abs(hfull[:, 1:]).max())
            oerr = float(np.abs(sharded[:, 1:] - o_ref[:, 1:]).max() / np.abs(o_ref[:, 1:]).max())
            oherr = float(np.abs(hmc_sharded[:, 1:] - h_ref[:, 1:]).max() / np.abs(h_ref[:, 1:]).max())
            out[tag] = {"rel_err_rmhmc": err, "rel_err_hmc": herr, "accept_equal": bool(np.array_equal(acc, info["accepted"])),
                        "rel_err_rmhmc_vs_oracle": oerr, "rel_err_hmc_vs_oracle": oherr,
                        "accept_equal_oracle": bool(np.array_equal(acc, [i["accepted"].sum() for i in o_infos]))}
        flags = torch.tensor([1.0 if same else 0.0], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            out[tag]["ranks_bit_identical"] = bool(flags.item() == 1.0)
    if rank == 0:
        ok = all(v["rel_err_rmhmc"] < 1e-10 and v["rel_err_hmc"] < 1e-10 and v["accept_equal"] and v["ranks_bit_identical"]
                 and v["rel_err_rmhmc_vs_oracle"] < 1e-9 and v["rel_err_hmc_vs_oracle"] < 1e-9 and v["accept_equal_oracle"]
                 for v in out.values())
        print(json.dumps({"world": world, "ok": ok, **out}))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
