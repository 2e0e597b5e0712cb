"""Host side of the batched samplers: device buffers (torch) + calls into the C ABI.

``LogisticData``   a data set bound to a device handle; also exposes the parity seams
                   (metric / partials / Cholesky) the reference inlines in rmhmc.py.
``RMHMCSampler``   many independent RMHMC chains (rmhmc.py:37-191 per chain).
``HMCSampler``     many independent Euclidean-HMC chains (hmc.py:38-89 per chain).

PyTorch is plumbing here (allocation, H2D/D2H, streams); every number is produced by the CUDA
kernels in ``csrc/``.
"""
from __future__ import annotations

import ctypes
from ctypes import c_int64, c_void_p

import numpy as np

from . import _capi

HUGE_ITERS = 1 << 60


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous row range [begin, end) of shard ``rank`` (row-sharded large-N mode)."""
    base, rem = divmod(n_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_chains(n_chains: int, rank: int, world: int):
    """Contiguous chain range [begin, end) of rank ``rank`` when ``n_chains`` chains are partitioned over ``world``
    GPUs (BASELINE.json configs[3]); ``begin`` is also the rank's Philox ``chain_offset``."""
    return shard_rows(n_chains, rank, world)


def chain_moments(samples):
    """Host mirror of the sufficient statistics rmhmc_stats_gather exchanges for Gelman-Rubin: for a (C, S, D) array the
    (3, D) sums over chains of the chain means, squared chain means and chain variances (ddof = 1)."""
    s = np.asarray(samples, dtype=np.float64)
    m, v = s.mean(axis=1), s.var(axis=1, ddof=1)
    return np.stack([m.sum(axis=0), (m * m).sum(axis=0), v.sum(axis=0)])


def rhat_from_moments(moments, c_total: int, n_samples: int):
    """Rhat per parameter from moments summed over ALL ranks (csrc/ess_kernel.cuh: k_rhat_finish)."""
    a, b, w = np.asarray(moments, dtype=np.float64)
    w = w / c_total
    b_over_s = (b - a * a / c_total) / (c_total - 1.0)
    return np.sqrt(((n_samples - 1.0) / n_samples * w + b_over_s) / w)


PROFILE_KINDS = ["metric_fp", "metric_closing", "partials", "chain_turn", "chain_solve", "quad_pass", "leverage_gemm",
                 "trace_pass", "i8_vslice", "i8_gemm", "allreduce", "i8_vslice_closing", "i8_qdigits"]


def device_peaks(device=0):
    """(FP64 DMMA TFLOP/s, INT8 tcgen05 TOP/s) issue peaks measured live on ``device`` (blr_device_peaks)."""
    torch = _capi.require_cuda()
    lib = _capi.load()
    dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
    a, b = ctypes.c_double(0), ctypes.c_double(0)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.blr_device_peaks(dev.index or 0, c_void_p(stream), ctypes.byref(a), ctypes.byref(b))
    if rc != 0:
        raise _capi.RmhmcError(f"blr_device_peaks failed (code {rc})")
    return a.value, b.value


class LogisticData:
    """``(XX, t)`` of the Bayesian logistic regression, resident on one GPU.

    ``row_shard=(rank, world)``: ``xx`` / ``t`` hold only this rank's rows (see :func:`shard_rows`);
    every metric / partials build is then summed over the ranks with NCCL inside the library
    (``torch.distributed`` must be initialised: it carries the NCCL unique id to the other ranks).
    """

    PARTIALS = {"tensor": 0, "matrix_free": 1}
    METRIC = {"dmma": 0, "i8": 1}
    REGIME = {"auto": 0, "small": 1, "large": 2}

    def __init__(self, xx, t, alpha: float = 100.0, device: str | int = "cuda:0", row_shard=None, partials=None,
                 metric=None, regime=None):
        torch = _capi.require_cuda()
        self._lib = _capi.load()
        self.torch = torch
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        xx = np.ascontiguousarray(xx, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1)
        if xx.ndim != 2 or xx.shape[0] != t.shape[0]:
            raise ValueError("XX must be (N, D) and t must have N entries")
        self.n_rows, self.dim = xx.shape
        self.alpha = float(alpha)
        self.h2d_bytes = xx.nbytes + t.nbytes
        xx_d = torch.from_numpy(xx).to(self.device)
        t_d = torch.from_numpy(t).to(self.device)
        handle = c_void_p()
        rc = self._lib.rmhmc_create(ctypes.byref(handle), self.device.index or 0, self.n_rows, self.dim,
                                    self.alpha, _ptr(xx_d), _ptr(t_d))
        _capi.check(rc, None, "rmhmc_create")
        self.handle = handle
        self.row_shard = row_shard
        if row_shard is not None:
            self._init_comm(*row_shard)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _capi.check(self._lib.rmhmc_set_stream(self.handle, c_void_p(stream)), self.handle, "rmhmc_set_stream")
        if partials is not None:
            self.set_partials_mode(partials)
        if metric is not None:
            self.set_metric_mode(metric)
        if regime is not None:
            self.set_launch_regime(regime)

    def set_metric_mode(self, mode: str):
        """``"dmma"``: fused FP64 kernel (legacy tensor path); ``"i8"``: INT8-slice build on the tcgen05 tensor cores
        (include/rmhmc_b200.h).  Drops the handle's chains: call before creating a sampler."""
        _capi.check(self._lib.rmhmc_set_metric_mode(self.handle, self.METRIC[mode]), self.handle, "rmhmc_set_metric_mode")

    @property
    def metric_mode(self) -> str:
        return "i8" if self._lib.rmhmc_get_metric_mode(self.handle) == 1 else "dmma"

    def set_launch_regime(self, regime: str):
        """Pin the chain-count dependent kernel variants: ``"small"`` / ``"large"`` / ``"auto"``."""
        _capi.check(self._lib.rmhmc_set_launch_regime(self.handle, self.REGIME[regime]), self.handle,
                    "rmhmc_set_launch_regime")

    def set_partials_mode(self, mode: str):
        """``"tensor"``: build the packed partials tensor per chain (the reference's formulation);
        ``"matrix_free"`` (default when it fits): traces and quadratic forms as passes over the data.
        Drops the handle's chains: call before creating a sampler."""
        _capi.check(self._lib.rmhmc_set_partials_mode(self.handle, self.PARTIALS[mode]), self.handle,
                    "rmhmc_set_partials_mode")

    @property
    def partials_mode(self) -> str:
        return "matrix_free" if self._lib.rmhmc_get_partials_mode(self.handle) == 1 else "tensor"

    def _init_comm(self, rank: int, world: int):
        import torch.distributed as dist
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            rc = self._lib.rmhmc_comm_unique_id(buf)
            if rc != 0:
                raise _capi.RmhmcError(f"rmhmc_comm_unique_id failed (code {rc})")
        box = [buf.raw]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        _capi.check(self._lib.rmhmc_comm_init(self.handle, int(world), int(rank), box[0]), self.handle, "rmhmc_comm_init")

    def init_stats_comm(self, rank: int, world: int):
        """NCCL communicator for the end-of-run statistics of a chain-sharded run (``stats_gather``); the unique id
        travels over ``torch.distributed`` (plumbing)."""
        import torch.distributed as dist
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            rc = self._lib.rmhmc_comm_unique_id(buf)
            if rc != 0:
                raise _capi.RmhmcError(f"rmhmc_comm_unique_id failed (code {rc})")
        box = [buf.raw]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        _capi.check(self._lib.rmhmc_stats_comm_init(self.handle, int(world), int(rank), box[0]), self.handle,
                    "rmhmc_stats_comm_init")

    def stats_gather(self, ess=None, samples=None, scalars=None):
        """Sum of the per-chain ESS (C, D) over all chains of all ranks, Gelman-Rubin Rhat of the sample window
        (C, S, D) over all chains of all ranks, and the element-wise sum of ``scalars`` -- one NCCL all-reduce inside the
        library (rmhmc_stats_gather).  Returns (ess_sum (D,), rhat (D,), scalars) as device tensors (None where no input)."""
        t = self.torch
        n_chains = ess.shape[0] if ess is not None else samples.shape[0]
        ess_sum = t.zeros(self.dim, dtype=t.float64, device=self.device) if ess is not None else None
        rhat = t.zeros(self.dim, dtype=t.float64, device=self.device) if samples is not None else None
        if ess is not None:
            ess = ess.contiguous()
        n_samples = cs = rs = 0
        if samples is not None:
            assert samples.dim() == 3 and samples.stride(2) == 1
            n_samples, cs, rs = samples.shape[1], samples.stride(0), samples.stride(1)
        if scalars is not None:
            scalars = scalars.to(t.float64).contiguous()
        _capi.check(self._lib.rmhmc_stats_gather(self.handle, _ptr(ess), int(n_chains), _ptr(samples), int(n_samples), int(cs),
                                                 int(rs), _ptr(ess_sum), _ptr(rhat), _ptr(scalars),
                                                 0 if scalars is None else scalars.numel()), self.handle, "rmhmc_stats_gather")
        return ess_sum, rhat, scalars

    def update(self, xx_dev, t_dev):
        """Re-upload the design matrix / labels from device tensors of the bound shape."""
        assert xx_dev.shape == (self.n_rows, self.dim) and t_dev.numel() == self.n_rows
        _capi.check(self._lib.rmhmc_update_data(self.handle, _ptr(xx_dev), _ptr(t_dev)), self.handle, "rmhmc_update_data")

    def close(self):
        if getattr(self, "handle", None):
            self._lib.rmhmc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, arr, dtype=np.float64):
        return self.torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(self.device)

    def _empty(self, *shape, dtype=None):
        return self.torch.empty(*shape, dtype=dtype or self.torch.float64, device=self.device)

    # ---------------------------------------------------------------- parity seams
    def metric(self, theta):
        """G (C,D,D), grad (C,D), logjoint (C,) at theta (C,D)   [rmhmc.py:51-57, :100, :31-34]."""
        th = self._dev(np.atleast_2d(theta))
        c = th.shape[0]
        g, grad, lj = self._empty(c, self.dim, self.dim), self._empty(c, self.dim), self._empty(c)
        _capi.check(self._lib.rmhmc_metric(self.handle, c, _ptr(th), _ptr(g), _ptr(grad), _ptr(lj)),
                    self.handle, "rmhmc_metric")
        return g.cpu().numpy(), grad.cpu().numpy(), lj.cpu().numpy()

    def metric_partials(self, theta):
        """dG (C,D,D,D) and tr(G^-1 dG_d) (C,D) at theta (C,D)   [rmhmc.py:64-77]."""
        th = self._dev(np.atleast_2d(theta))
        c = th.shape[0]
        dg, tr = self._empty(c, self.dim, self.dim, self.dim), self._empty(c, self.dim)
        _capi.check(self._lib.rmhmc_metric_partials(self.handle, c, _ptr(th), _ptr(dg), _ptr(tr)),
                    self.handle, "rmhmc_metric_partials")
        return dg.cpu().numpy(), tr.cpu().numpy()

    def chol_logdet(self, g):
        """Lower Cholesky factor, inverse and sum log diag of SPD matrices g (C,D,D)   [rmhmc.py:58-60,171]."""
        gd = self._dev(g)
        c = gd.shape[0]
        l, gi, ld = self._empty(c, self.dim, self.dim), self._empty(c, self.dim, self.dim), self._empty(c)
        _capi.check(self._lib.rmhmc_chol_logdet(self.handle, c, _ptr(gd), _ptr(l), _ptr(gi), _ptr(ld)),
                    self.handle, "rmhmc_chol_logdet")
        return l.cpu().numpy(), gi.cpu().numpy(), ld.cpu().numpy()


    def leapfrog(self, theta, mom, direction, n_steps, step_size: float, n_fixed: int):
        """Deterministic generalized leapfrog seam (rmhmc.py:96-163): ``n_steps[c]`` steps from
        ``(theta[c], mom[c])`` in direction ``direction[c]``; returns (theta, mom, H_start, H_end)."""
        t = self.torch
        th, mo = self._dev(np.atleast_2d(theta)), self._dev(np.atleast_2d(mom))
        c = th.shape[0]
        di = self._dev(np.asarray(direction).reshape(c), dtype=np.int32)
        ns = self._dev(np.asarray(n_steps).reshape(c), dtype=np.int32)
        o_th, o_mo, h0, h1 = self._empty(c, self.dim), self._empty(c, self.dim), self._empty(c), self._empty(c)
        _capi.check(self._lib.rmhmc_leapfrog(self.handle, c, _ptr(th), _ptr(mo), _ptr(di), _ptr(ns), float(step_size),
                                             int(n_fixed), _ptr(o_th), _ptr(o_mo), _ptr(h0), _ptr(h1)),
                    self.handle, "rmhmc_leapfrog")
        return o_th.cpu().numpy(), o_mo.cpu().numpy(), h0.cpu().numpy(), h1.cpu().numpy()


class _SamplerBase:
    _is_hmc = False

    def __init__(self, data: LogisticData, n_chains: int, theta0=None, _init=None):
        self.data = data
        self._lib = data._lib
        self.torch = data.torch
        self.h = data.handle
        self.n_chains = int(n_chains)
        self.dim = data.dim
        self._keep = {}          # device buffers the handle points into
        self.samples = None
        self.burn_in = 0
        self.trace = None
        th0 = None if theta0 is None else data._dev(np.asarray(theta0, dtype=np.float64).reshape(self.n_chains, self.dim))
        if _init is not None:
            _capi.check(_init(th0), self.h, "chains_init")
        else:
            fn = self._lib.hmc_chains_init if self._is_hmc else self._lib.rmhmc_chains_init
            _capi.check(fn(self.h, self.n_chains, _ptr(th0)), self.h, "chains_init")
        self._gen = self._lib.rmhmc_chain_generation(self.h)

    def _check_alive(self):
        """A handle owns one chain set: a later sampler / seam call on the same LogisticData replaced this one's chains."""
        if self._lib.rmhmc_chain_generation(self.h) != self._gen:
            raise _capi.RmhmcError("this sampler's chains were replaced by a later chains_init / seam call on the same "
                                   "LogisticData handle (one chain set per handle)")

    # ---------------------------------------------------------------- randomness
    def set_philox(self, seed: int, chain_offset: int = 0):
        _capi.check(self._lib.rmhmc_set_philox(self.h, int(seed) & (2**64 - 1), int(chain_offset)), self.h, "set_philox")

    # ---------------------------------------------------------------- outputs
    def set_samples(self, capacity: int, burn_in: int):
        """Allocate the (C, capacity, D) sample store; row it-burn_in receives the state after iteration it."""
        self.samples = self.torch.zeros(self.n_chains, int(capacity), self.dim, dtype=self.torch.float64,
                                        device=self.data.device)
        self.burn_in = int(burn_in)
        _capi.check(self._lib.rmhmc_set_samples(self.h, _ptr(self.samples), int(capacity), int(burn_in)),
                    self.h, "set_samples")
        return self.samples

    def set_trace(self, n_iters: int, n_leapfrog: int):
        t, dev, c, d = self.torch, self.data.device, self.n_chains, self.dim
        nan = float("nan")
        self.trace = {
            "theta_steps": t.full((c, n_iters, n_leapfrog, d), nan, dtype=t.float64, device=dev),
            "mom_end": t.zeros(c, n_iters, d, dtype=t.float64, device=dev),
            "theta_end": t.zeros(c, n_iters, d, dtype=t.float64, device=dev),
            "mom0": t.zeros(c, n_iters, d, dtype=t.float64, device=dev),
            "h_current": t.zeros(c, n_iters, dtype=t.float64, device=dev),
            "h_proposed": t.zeros(c, n_iters, dtype=t.float64, device=dev),
            "flags": t.zeros(c, n_iters, dtype=t.int32, device=dev),
        }
        tr = self.trace
        _capi.check(self._lib.rmhmc_set_trace(self.h, int(n_iters), _ptr(tr["theta_steps"]), _ptr(tr["mom_end"]),
                                              _ptr(tr["theta_end"]), _ptr(tr["mom0"]), _ptr(tr["h_current"]),
                                              _ptr(tr["h_proposed"]), _ptr(tr["flags"])), self.h, "set_trace")

    def trace_numpy(self):
        out = {k: v.cpu().numpy() for k, v in self.trace.items()}
        fl = out["flags"]
        out["accepted"] = (fl & 1).astype(bool)
        out["used_uniform"] = (fl & 2).astype(bool)
        out["direction"] = np.where(fl & 16, 1, -1)
        out["n_steps"] = fl >> 8
        return out

    def state(self):
        self._check_alive()
        t, dev, c = self.torch, self.data.device, self.n_chains
        theta = t.empty(c, self.dim, dtype=t.float64, device=dev)
        iters = t.empty(c, dtype=t.int64, device=dev)
        acc = t.empty(c, dtype=t.int64, device=dev)
        lf = t.empty(c, dtype=t.int64, device=dev)
        rm = t.empty(c, dtype=t.int32, device=dev)
        rp = t.empty(c, dtype=t.int32, device=dev)
        _capi.check(self._lib.rmhmc_read_state(self.h, _ptr(theta), _ptr(iters), _ptr(acc), _ptr(lf), _ptr(rm), _ptr(rp)),
                    self.h, "read_state")
        return {"theta": theta.cpu().numpy(), "iters": iters.cpu().numpy(), "accepted": acc.cpu().numpy(),
                "leapfrogs": lf.cpu().numpy(), "renorm_momentum": rm.cpu().numpy(), "renorm_position": rp.cpu().numpy()}

    def launch_count(self) -> int:
        return int(self._lib.rmhmc_launch_count(self.h))

    _advance_fn = "rmhmc_advance"

    def advance(self, n_rounds: int, it_stop: int = HUGE_ITERS):
        """Enqueue ``n_rounds`` rounds without synchronising (free-running chains)."""
        self._check_alive()
        _capi.check(getattr(self._lib, self._advance_fn)(self.h, int(n_rounds), int(it_stop)), self.h, self._advance_fn)

    def profile(self, enable: bool):
        _capi.check(self._lib.rmhmc_profile_enable(self.h, 1 if enable else 0), self.h, "profile_enable")

    def profile_read(self):
        """{kind: (milliseconds, launches)} per kernel class."""
        out = {}
        for k, nm in enumerate(PROFILE_KINDS):
            ms, n = ctypes.c_double(0), c_int64(0)
            _capi.check(self._lib.rmhmc_profile_read(self.h, k, ctypes.byref(ms), ctypes.byref(n)), self.h, "profile_read")
            out[nm] = (ms.value, n.value)
        return out


class RMHMCSampler(_SamplerBase):
    """C independent RMHMC chains; every ``round`` advances each chain by one generalized leapfrog step."""

    def __init__(self, data: LogisticData, n_chains: int, n_leapfrog: int = 6, step_size: float = 0.5,
                 n_fixed: int = 4, theta0=None, student_t: bool = False):
        super().__init__(data, n_chains, theta0)
        self.n_leapfrog, self.step_size, self.n_fixed = int(n_leapfrog), float(step_size), int(n_fixed)
        _capi.check(self._lib.rmhmc_configure(self.h, self.n_leapfrog, self.step_size, self.n_fixed), self.h, "configure")
        self.student_t = bool(student_t)      # BLR_RMHMC_StudentT.m: Student-t kinetic energy (include/rmhmc_b200.h)
        _capi.check(self._lib.rmhmc_set_momentum_family(self.h, 1 if student_t else 0), self.h, "rmhmc_set_momentum_family")

    def set_tape(self, z, u_step, z_dir, u_acc, it_base: int = 0, z_chi=None):
        """Host draws in the C-ABI layout: z (W,C,D), u_step/z_dir/u_acc (W,C); Student-t: z_chi (W,C)."""
        d = self.data
        self._keep["tape"] = [d._dev(z), d._dev(u_step), d._dev(z_dir), d._dev(u_acc)]
        w = self._keep["tape"][0].shape[0]
        zt, us, zd, ua = self._keep["tape"]
        assert zt.shape == (w, self.n_chains, self.dim) and us.shape == (w, self.n_chains)
        _capi.check(self._lib.rmhmc_set_tape(self.h, int(it_base), int(w), _ptr(zt), _ptr(us), _ptr(zd), _ptr(ua)),
                    self.h, "set_tape")
        if z_chi is not None:
            self._keep["tape_chi"] = d._dev(z_chi)
            assert self._keep["tape_chi"].shape == (w, self.n_chains)
            _capi.check(self._lib.rmhmc_set_tape_chi(self.h, _ptr(self._keep["tape_chi"])), self.h, "set_tape_chi")

    def set_trace(self, n_iters: int):          # noqa: D102
        super().set_trace(n_iters, self.n_leapfrog)

    def run(self, it_stop: int) -> int:
        """Run rounds until every chain has completed ``it_stop`` iterations; returns the round count."""
        self._check_alive()
        rounds = c_int64(0)
        _capi.check(self._lib.rmhmc_run(self.h, int(it_stop), ctypes.byref(rounds)), self.h, "rmhmc_run")
        return rounds.value


class HMCSampler(_SamplerBase):
    """C independent Euclidean-HMC chains (identity mass)."""

    _is_hmc = True
    _advance_fn = "hmc_advance"

    def __init__(self, data: LogisticData, n_chains: int, n_leapfrog: int = 100, step_size: float = 0.14, theta0=None,
                 fused=None, rounds_per_launch: int = 0):
        super().__init__(data, n_chains, theta0)
        self.n_leapfrog, self.step_size = int(n_leapfrog), float(step_size)
        _capi.check(self._lib.hmc_configure(self.h, self.n_leapfrog, self.step_size), self.h, "hmc_configure")
        if fused is not None or rounds_per_launch:
            self.set_fused(True if fused is None else fused, rounds_per_launch)

    def set_fused(self, fused: bool, rounds_per_launch: int = 0):
        """Many leapfrog rounds per launch with the chain state in registers (default) or three launches per round."""
        _capi.check(self._lib.hmc_set_fused(self.h, int(bool(fused)), int(rounds_per_launch)), self.h, "hmc_set_fused")

    def set_tape(self, z, u_step, u_acc, it_base: int = 0):
        d = self.data
        self._keep["tape"] = [d._dev(z), d._dev(u_step), d._dev(u_acc)]
        zt, us, ua = self._keep["tape"]
        _capi.check(self._lib.hmc_set_tape(self.h, int(it_base), int(zt.shape[0]), _ptr(zt), _ptr(us), _ptr(ua)),
                    self.h, "hmc_set_tape")

    def set_trace(self, n_iters: int):          # noqa: D102
        super().set_trace(n_iters, 1)

    def run(self, it_stop: int) -> int:
        self._check_alive()
        rounds = c_int64(0)
        _capi.check(self._lib.hmc_run(self.h, int(it_stop), ctypes.byref(rounds)), self.h, "hmc_run")
        return rounds.value


class MMALASampler(_SamplerBase):
    """C independent (simplified) manifold-MALA chains; one round = one MCMC iteration of every chain.

    MATLAB-only in the reference (BLR_mMALA.m / BLR_mMALA_Simp.m); see include/rmhmc_b200.h.
    """

    _advance_fn = "mmala_advance"

    def __init__(self, data: LogisticData, n_chains: int, step_size: float = 1.0, simplified: bool = False, theta0=None,
                 iwls: bool = False):
        lib, h, c = data._lib, data.handle, int(n_chains)
        variant = 2 if iwls else (1 if simplified else 0)        # 2: the IWLS proposal of code/iwls.py
        super().__init__(data, n_chains, theta0,
                         _init=lambda th0: lib.mmala_chains_init(h, c, _ptr(th0), variant, float(step_size)))
        self.step_size, self.simplified, self.iwls = float(step_size), bool(simplified), bool(iwls)

    def proposal_moments(self):
        """(mean (C, D), L (C, D, D)) of every chain's current proposal N(mean, L L^T) as numpy arrays; one chain: (D,), (D, D)."""
        t, dev, c, d = self.torch, self.data.device, self.n_chains, self.dim
        mean, chol = t.empty(c, d, dtype=t.float64, device=dev), t.empty(c, d, d, dtype=t.float64, device=dev)
        _capi.check(self._lib.mmala_read_proposal(self.h, _ptr(mean), _ptr(chol)), self.h, "mmala_read_proposal")
        mean, chol = mean.cpu().numpy(), chol.cpu().numpy()
        return (mean[0], chol[0]) if c == 1 else (mean, chol)

    def set_tape(self, z, u_acc, it_base: int = 0):
        d = self.data
        self._keep["tape"] = [d._dev(z), d._dev(u_acc)]
        zt, ua = self._keep["tape"]
        assert zt.shape == (zt.shape[0], self.n_chains, self.dim) and ua.shape == (zt.shape[0], self.n_chains)
        _capi.check(self._lib.mmala_set_tape(self.h, int(it_base), int(zt.shape[0]), _ptr(zt), _ptr(ua)), self.h,
                    "mmala_set_tape")

    def set_trace(self, n_iters: int):          # noqa: D102
        super().set_trace(n_iters, 1)

    def run(self, it_stop: int) -> int:
        self._check_alive()
        rounds = c_int64(0)
        _capi.check(self._lib.mmala_run(self.h, int(it_stop), ctypes.byref(rounds)), self.h, "mmala_run")
        return rounds.value


def ess_batched(samples, max_lag: int | None = None):
    """ESS of every (chain, parameter) series of a device or host array (C, S, D) -> numpy (C, D).

    Exactly ``tools.CalculateESS(samples[c], max_lag)`` (tools.py:32-74), computed on the GPU.
    """
    torch = _capi.require_cuda()
    lib = _capi.load()
    if isinstance(samples, np.ndarray):
        samples = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float64)).cuda()
    assert samples.dim() == 3 and samples.dtype == torch.float64 and samples.stride(2) == 1
    c, s, d = samples.shape
    if max_lag is None:
        max_lag = s - 1
    out = torch.empty(c, d, dtype=torch.float64, device=samples.device)
    stream = torch.cuda.current_stream(samples.device).cuda_stream
    rc = lib.blr_ess_batched(samples.device.index or 0, c_void_p(stream), _ptr(samples), c, s, d,
                             samples.stride(0), samples.stride(1), int(max_lag), _ptr(out))
    if rc != 0:
        raise _capi.RmhmcError(f"blr_ess_batched failed (code {rc})")
    return out


def rhat_batched(samples):
    """Gelman-Rubin Rhat per parameter of a device or host array (C, S, D) -> device tensor (D,)."""
    torch = _capi.require_cuda()
    lib = _capi.load()
    if isinstance(samples, np.ndarray):
        samples = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float64)).cuda()
    assert samples.dim() == 3 and samples.dtype == torch.float64 and samples.stride(2) == 1
    c, s, d = samples.shape
    out = torch.empty(d, dtype=torch.float64, device=samples.device)
    stream = torch.cuda.current_stream(samples.device).cuda_stream
    rc = lib.blr_rhat(samples.device.index or 0, c_void_p(stream), _ptr(samples), c, s, d, samples.stride(0), samples.stride(1),
                      _ptr(out))
    if rc != 0:
        raise _capi.RmhmcError(f"blr_rhat failed (code {rc})")
    return out


def ess_ragged(samples, starts, counts):
    """ESS per (chain, parameter) over per-chain row windows of a device array (C, S, D).

    ``starts`` / ``counts`` are int64 device tensors (C,): chain c uses rows
    [starts[c], starts[c]+counts[c]).  Returns a device tensor (C, D).
    """
    torch = _capi.require_cuda()
    lib = _capi.load()
    c, s, d = samples.shape
    assert samples.dtype == torch.float64 and samples.stride(2) == 1
    starts = starts.to(torch.int64).contiguous()
    counts = counts.to(torch.int64).contiguous()
    max_s = int(counts.max().item())
    out = torch.zeros(c, d, dtype=torch.float64, device=samples.device)
    if max_s < 2:
        return out
    stream = torch.cuda.current_stream(samples.device).cuda_stream
    rc = lib.blr_ess_ragged(samples.device.index or 0, c_void_p(stream), _ptr(samples), c, max_s, d,
                            samples.stride(0), samples.stride(1), _ptr(starts), _ptr(counts), _ptr(out))
    if rc != 0:
        raise _capi.RmhmcError(f"blr_ess_ragged failed (code {rc})")
    return out


def autocorr_batched(series, n_lag: int):
    """tools.ac for every row of ``series`` (n_series, S) -> device tensor (n_series, n_lag+1)."""
    torch = _capi.require_cuda()
    lib = _capi.load()
    if isinstance(series, np.ndarray):
        series = torch.from_numpy(np.ascontiguousarray(series, dtype=np.float64)).cuda()
    series = series.contiguous()
    n, s = series.shape
    out = torch.empty(n, int(n_lag) + 1, dtype=torch.float64, device=series.device)
    stream = torch.cuda.current_stream(series.device).cuda_stream
    rc = lib.blr_autocorr(series.device.index or 0, c_void_p(stream), _ptr(series), n, s, int(n_lag), _ptr(out))
    if rc != 0:
        raise _capi.RmhmcError(f"blr_autocorr failed (code {rc})")
    return out
