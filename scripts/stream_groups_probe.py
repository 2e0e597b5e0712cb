"""Probe: do G independent chain groups on G CUDA streams (one handle each) fill the GPU better than one batch when the
per-GPU batch is small (strong-scaled configs[3]: 8192 chains per GPU)?  Prints chain-leapfrog-steps/s per variant."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import riemannhamiltonianmontecarlo_b200 as r

xx, t = r.datasets.shaped("german")
out = {}
for total in (8192, 16384, 65536):
    for groups in (1, 2, 4):
        c = total // groups
        streams = [torch.cuda.Stream() for _ in range(groups)]
        datas, samplers = [], []
        for g in range(groups):
            with torch.cuda.stream(streams[g]):
                d = r.LogisticData(xx, t)
                s = r.RMHMCSampler(d, c, 6, 0.5, 6)
                s.set_philox(7, g * c)
                datas.append(d); samplers.append(s)
        R = 40
        def run(n):
            for _ in range(n):
                for g in range(groups):
                    with torch.cuda.stream(streams[g]):
                        samplers[g].advance(10)
        run(2); torch.cuda.synchronize()
        t0 = time.perf_counter(); run(R // 10); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out[f"{total}x{groups}"] = {"ms_per_round": dt / R * 1e3, "chain_rounds_per_s": total * R / dt}
        print(total, groups, out[f"{total}x{groups}"], flush=True)
        for d in datas:
            d.close()
print(json.dumps(out))
