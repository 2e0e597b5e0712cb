"""CPU tests of the host side: the C-ABI library builds, loads and exports every symbol the header
declares; host helpers; loud failure without a GPU; the bench CPU arm; multi-rank plumbing (gloo)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rmhmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:rmhmc|hmc|mmala|blr)_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(built_library):
    from riemannhamiltonianmontecarlo_b200 import _capi
    lib = _capi.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/rmhmc_b200.h but not exported"
        assert name in _capi.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_capi.SIGNATURES) == set(syms)
    assert b"sm_100a" in lib.rmhmc_version()


def test_library_is_sm100a_only(built_library):
    out = subprocess.run(["cuobjdump", "-lelf", built_library], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import riemannhamiltonianmontecarlo_b200 as r
    from riemannhamiltonianmontecarlo_b200._capi import RmhmcError
    xx, t = r.datasets.shaped("australian")
    with pytest.raises(RmhmcError):
        r.RMHMC(xx, t, 10, 2)
    with pytest.raises(RmhmcError):
        r.HMC(xx, t, 10, 2)
    with pytest.raises(RmhmcError):
        r.CalculateESS(np.zeros((10, 2)), 9)
    with pytest.raises(RmhmcError):
        r.ac(np.zeros(10), 5)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "riemannhamiltonianmontecarlo_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn


def test_host_helpers_match_reference_semantics(golden):
    import riemannhamiltonianmontecarlo_b200 as r
    fx = golden("tools_ess")
    assert abs(r.LogNormPDF(np.zeros((1, 15)), fx["lnp_w"], 100) - float(fx["lnp"])) < 1e-12
    assert abs(r.LogNormPDF(np.zeros((15, 1)), fx["lnp_w"], 100) - float(fx["lnp"])) < 1e-12
    for i, v in fx["nextpow2"]:
        assert r.nextpow2(int(i)) == int(v)


def test_csv_preprocessing_matches_main_py(tmp_path):
    from riemannhamiltonianmontecarlo_b200 import datasets
    rng = np.random.default_rng(1)
    raw = np.hstack([rng.normal(3, 2, (40, 4)), rng.integers(1, 3, (40, 1))])
    p = tmp_path / "toy.csv"
    np.savetxt(p, raw, delimiter=",")
    xx, t = datasets.load_csv(str(p), relabel_12=True)
    x = raw[:, :-1]
    expect = np.hstack([np.ones((40, 1)), (x - x.mean(0)) / x.std(0)])      # main.py:34-41 (ddof=0)
    assert np.allclose(xx, expect) and set(np.unique(t)) <= {0.0, 1.0} and t.shape == (40, 1)


def test_packed_index_tables_roundtrip():
    # python mirror of csrc/common.cuh index math: packed triples enumerate i<=j<=k exactly once
    def tri(n): return n * (n + 1) * (n + 2) // 6
    def triple_index(i, j, k, d):
        n, jj, kk = d - i, j - i, k - i
        return tri(d) - tri(d - i) + jj * n - jj * (jj - 1) // 2 + (kk - jj)
    for d in (3, 15, 25, 32):
        seen = [triple_index(i, j, k, d) for i in range(d) for j in range(i, d) for k in range(j, d)]
        assert seen == list(range(tri(d)))


def test_bench_reference_arm_prints_contract_json():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--ref-iters-per-step", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "min_ess_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_row_shards_partition_the_rows():
    from riemannhamiltonianmontecarlo_b200.engine import shard_rows
    for n, world in [(1000, 1), (1003, 2), (690, 4), (10_000_000, 8), (5, 8)]:
        spans = [shard_rows(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    """One rank of a chain-sharded run on CPU: the repo's partition (engine.shard_chains), its Gelman-Rubin sufficient
    statistics (engine.chain_moments, the host mirror of what rmhmc_stats_gather exchanges) and the finishing formula
    (engine.rhat_from_moments = k_rhat_finish), with gloo standing in for the library's NCCL all-reduce."""
    import ctypes
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from riemannhamiltonianmontecarlo_b200 import _capi
    from riemannhamiltonianmontecarlo_b200.engine import chain_moments, rhat_from_moments, shard_chains
    from oracle import blr_oracle as bo
    # the communicators are bootstrapped with rank 0's NCCL unique id (128 bytes from the C ABI) sent over torch.distributed
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        assert _capi.load().rmhmc_comm_unique_id(buf) == 0
    box = [buf.raw]
    dist.broadcast_object_list(box, src=0)
    assert len(box[0]) == 128 and any(box[0])
    # every rank builds the same global (C, S, D) array and keeps only ITS chains
    c_total, n_s, d = 13, 200, 3
    rng = np.random.default_rng(77)
    full = rng.normal(0, 1, (c_total, n_s, d)).cumsum(axis=1) * 0.05 + rng.normal(0, 1, (c_total, 1, d)) * np.array([0.0, 0.3, 1.0])
    c0, c1 = shard_chains(c_total, rank, world)
    mine = full[c0:c1]
    ess_sum = torch.from_numpy(np.sum([bo.ess(mine[i], n_s - 1)[:, 0] for i in range(c1 - c0)], axis=0))
    buf_t = torch.from_numpy(np.concatenate([chain_moments(mine).ravel(), [float(c1 - c0)]]))
    dist.all_reduce(ess_sum, op=dist.ReduceOp.SUM)
    dist.all_reduce(buf_t, op=dist.ReduceOp.SUM)
    tmax = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    rhat = rhat_from_moments(buf_t[:-1].numpy().reshape(3, d), int(buf_t[-1].item()), n_s)
    q.put((rank, (c0, c1), ess_sum.tolist(), rhat.tolist(), float(tmax.item())))
    dist.destroy_process_group()


def test_two_rank_chain_sharded_statistics_over_gloo():
    """N > 1 host logic of BASELINE.json configs[3] on CPU (world_size 2, gloo): the partition covers every chain once
    and the combined statistics equal the single-process oracle on all chains."""
    import torch.multiprocessing as mp
    from oracle import blr_oracle as bo
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    c_total, n_s, d = 13, 200, 3
    rng = np.random.default_rng(77)
    full = rng.normal(0, 1, (c_total, n_s, d)).cumsum(axis=1) * 0.05 + rng.normal(0, 1, (c_total, 1, d)) * np.array([0.0, 0.3, 1.0])
    ess_ref = np.sum([bo.ess(full[i], n_s - 1)[:, 0] for i in range(c_total)], axis=0)
    rhat_ref = bo.rhat(full)
    assert [r[1] for r in res] == [(0, 7), (7, 13)]
    for rank, span, ess_sum, rhat, tmax in res:
        assert np.allclose(ess_sum, ess_ref, rtol=1e-12) and np.allclose(rhat, rhat_ref, rtol=1e-12) and tmax == 2.0


def test_dataset_dispatch_and_ripley_cubic_basis(tmp_path):
    """datasets.load_dataset: the MATLAB drivers' per-data-set preparation (BLR_RMHMC.m:9-178) -- relabelling for german /
    heart, the cubic basis [1, X, X^2, X^3] for ripley (D = 7), population-std standardisation as in main.py:34-37."""
    from riemannhamiltonianmontecarlo_b200 import datasets
    rng = np.random.default_rng(11)
    x = rng.normal(1.0, 2.0, (50, 2))
    lab01 = (rng.random(50) < 0.5).astype(float)
    np.savetxt(tmp_path / "ripley.csv", np.hstack([x, lab01[:, None]]), delimiter=",")
    np.savetxt(tmp_path / "german.csv", np.hstack([x, lab01[:, None] + 1.0]), delimiter=",")
    xx, t = datasets.load_dataset("Ripley", str(tmp_path))
    xs = (x - x.mean(0)) / x.std(0)
    assert xx.shape == (50, 7) and np.array_equal(t[:, 0], lab01)
    assert np.array_equal(xx[:, 0], np.ones(50))
    assert np.allclose(xx[:, 1:3], xs) and np.allclose(xx[:, 3:5], xs ** 2) and np.allclose(xx[:, 5:7], xs ** 3)
    xg, tg = datasets.load_dataset("german", str(tmp_path))
    assert xg.shape == (50, 3) and np.array_equal(tg[:, 0], lab01)          # {1, 2} -> {0, 1}
    assert np.array_equal(xg, datasets.load_csv(str(tmp_path / "german.csv"), relabel_12=True)[0])
    with pytest.raises(ValueError):
        datasets.load_dataset("mnist", str(tmp_path))
    from riemannhamiltonianmontecarlo_b200 import harness
    assert set(harness.HMC_STEP_SIZES) == set(datasets.DATASETS)


def test_thread_per_chain_linear_algebra_core(tmp_path):
    """csrc/tpc_core.h (packed in-place Cholesky / solve / inverse, one thread per chain) compiled for the host and
    checked against dense arithmetic; the two-column factorisation must be bit-identical to the one-column one."""
    exe = tmp_path / "tpc_host_test"
    src = os.path.join(ROOT, "tests", "native", "tpc_host_test.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", str(exe), src], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-2000:]
