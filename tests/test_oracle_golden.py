"""CPU tests: the oracle restatement against the committed golden fixtures (generated from the
unmodified reference by tests/golden/make_golden.py) -- bit-exact, since the oracle keeps the
reference's NumPy operation order -- and, when /root/reference is mounted, against the live code.
"""
import numpy as np
import pytest

from oracle import blr_oracle as bo
from oracle import ref_live


def _tapes(fx, with_dir=True):
    c = fx["z"].shape[1]
    out = []
    for i in range(c):
        out.append(bo.DrawTape(z=fx["z"][:, i], u_step=fx["u_step"][:, i],
                               z_dir=fx["z_dir"][:, i] if with_dir else np.zeros(fx["z"].shape[0]),
                               u_acc=fx["u_acc"][:, i]))
    return out


@pytest.mark.parametrize("name", ["rmhmc_australian_shaped", "rmhmc_german_real", "rmhmc_pima_real"])
def test_rmhmc_oracle_reproduces_reference_fixture(golden, name):
    fx = golden(name)
    tapes = _tapes(fx)
    n_chain = min(len(tapes), 2)          # keep the CPU suite short
    for c in range(n_chain):
        s, info = bo.rmhmc_chain(fx["xx"], fx["t"], tapes[c], int(fx["n_iter"]), int(fx["burn_in"]),
                                 int(fx["n_leapfrog"]), float(fx["step_size"]), int(fx["n_fixed"]), record=True)
        assert np.array_equal(s[1:], fx["samples"][c, 1:])          # bit-exact
        assert np.all(np.isnan(s[0]))                                 # row 0 never written (rmhmc.py:28,190)
        assert np.array_equal(info["accepted"], fx["accepted"][c])
        for it, rec in enumerate(info["records"]):
            assert rec.n_steps == fx["n_steps"][c, it] and rec.direction == fx["direction"][c, it]
            assert rec.h_current == fx["h_current"][c, it] and rec.h_proposed == fx["h_proposed"][c, it]
            assert rec.used_uniform == fx["used_uniform"][c, it]
            for k in range(rec.n_steps):
                assert np.array_equal(rec.theta_steps[k], fx["theta_steps"][c, it, k])


def test_uniform_is_consumed_only_when_ratio_not_positive(golden):
    fx = golden("rmhmc_australian_shaped")
    assert np.array_equal(fx["used_uniform"], ~(fx["ratio"] > 0))
    assert fx["used_uniform"].any() and (~fx["used_uniform"]).any()


@pytest.mark.parametrize("name", ["hmc_australian_shaped", "hmc_pima_real", "hmc_german_shaped"])
def test_hmc_oracle_reproduces_reference_fixture(golden, name):
    fx = golden(name)
    tapes = _tapes(fx, with_dir=False)
    s, info = bo.hmc_chain(fx["xx"], fx["t"], tapes[0], int(fx["n_iter"]), int(fx["burn_in"]),
                           int(fx["n_leapfrog"]), float(fx["step_size"]), record=True)
    assert np.array_equal(s, fx["samples"][0])
    assert np.array_equal(info["accepted"], fx["accepted"][0])
    assert np.all(s[0] == 0.0)                                        # hmc.py:28,83


@pytest.mark.parametrize("name", ["iwls_australian_shaped", "iwls_pima_real"])
def test_iwls_oracle_reproduces_reference_fixture(golden, name):
    fx = golden(name)
    n = fx["z"].shape[0]
    tape = bo.DrawTape(z=fx["z"][:, 0], u_step=np.zeros(n), z_dir=np.zeros(n), u_acc=fx["u_acc"][:, 0])
    s, info = bo.iwls_chain(fx["xx"], fx["t"], tape, int(fx["n_iter"]), int(fx["burn_in"]), record=True)
    assert np.array_equal(s, fx["samples"][0])                        # bit-exact
    assert np.array_equal(info["accepted"], fx["accepted"][0])
    assert [r["used_uniform"] for r in info["records"]] == list(fx["used_uniform"][0])


@pytest.mark.reference
@pytest.mark.skipif(not ref_live.available(), reason="/root/reference not mounted")
def test_iwls_oracle_matches_live_reference_on_a_fresh_tape():
    from riemannhamiltonianmontecarlo_b200 import datasets
    xx, t = datasets.shaped("australian")
    tape = bo.make_tape(12, xx.shape[1], 4242)
    w_ref, info = ref_live.run_iwls(xx, t, tape, 12, 3)
    w_orc, oinfo = bo.iwls_chain(xx, t, tape, 12, 3, record=True)
    assert np.array_equal(w_ref, w_orc)
    assert np.array_equal(info["uniform_used"], [r["used_uniform"] for r in oinfo["records"]])


def test_tools_oracle_reproduces_reference_fixture(golden):
    fx = golden("tools_ess")
    x = fx["x"]
    assert np.array_equal(bo.ess(x, x.shape[0] - 1), fx["ess_full"])
    assert np.array_equal(bo.ess(x[:599], 598), fx["ess_599"])
    assert np.array_equal(bo.ess(x, 50), fx["ess_lag50"])
    for j in range(x.shape[1]):
        assert np.array_equal(bo.autocorr(x[:, j], 200), fx["ac_200"][:, j])
    assert bo.log_norm_pdf(np.zeros((1, 15)), fx["lnp_w"], 100) == fx["lnp"]
    for i, v in fx["nextpow2"]:
        assert bo.next_pow2(int(i)) == int(v)


def test_ac_uses_the_ports_nfft_quirk():
    # nFFT = nextpow2(n)+1 (tools.py:23): circular aliasing adds lin[nFFT-k] for k > nFFT-n
    rng = np.random.default_rng(3)
    x = np.cumsum(rng.standard_normal(600))          # strongly autocorrelated
    n, n_fft = 600, 1025
    y = x - x.mean()
    lin = np.correlate(y, y, mode="full")[n - 1:]
    k = 500                                           # n_fft - k = 525 < n: aliased
    expect = (lin[k] + lin[n_fft - k]) / lin[0]
    assert abs(bo.autocorr(x, 599)[k] - expect) < 1e-9


def test_rhat_of_identical_and_shifted_chains():
    rng = np.random.default_rng(0)
    base = rng.standard_normal((4, 500, 3))
    assert np.all(np.abs(bo.rhat(base) - 1) < 0.02)
    shifted = base + np.arange(4)[:, None, None] * 3.0
    assert np.all(bo.rhat(shifted) > 2)


def test_synthetic_generator_is_deterministic_and_standardised():
    from riemannhamiltonianmontecarlo_b200 import datasets
    xx, t = datasets.shaped("australian")
    xx2, t2 = datasets.shaped("australian")
    assert np.array_equal(xx, xx2) and np.array_equal(t, t2)
    assert xx.shape == (690, 15) and t.shape == (690, 1) and set(np.unique(t)) == {0.0, 1.0}
    assert np.all(xx[:, 0] == 1) and np.allclose(xx[:, 1:].mean(0), 0) and np.allclose(xx[:, 1:].std(0), 1)
    fx = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "rmhmc_australian_shaped.npz"))
    assert np.array_equal(fx["xx"], xx) and np.array_equal(fx["t"], t)   # the fixture's data is this generator's


@pytest.mark.reference
@pytest.mark.skipif(not ref_live.available(), reason="/root/reference not mounted")
def test_oracle_matches_live_reference_on_a_fresh_tape():
    from riemannhamiltonianmontecarlo_b200 import datasets
    xx, t = datasets.shaped("australian")
    tape = bo.make_tape(14, xx.shape[1], 31337)
    w_ref, info = ref_live.run_rmhmc(xx, t, tape, 14, 3, 6, 0.5, 4)
    w_orc, oinfo = bo.rmhmc_chain(xx, t, tape, 14, 3, 6, 0.5, 4, record=True)
    assert np.array_equal(w_ref[1:], w_orc[1:])
    assert np.array_equal(info["uniform_used"], [r.used_uniform for r in oinfo["records"]])
    tools = ref_live.load_reference()["tools"]
    x = np.random.default_rng(5).standard_normal((700, 2)).cumsum(axis=0) * 0.05 + np.random.default_rng(6).standard_normal((700, 2))
    assert np.array_equal(tools.CalculateESS(x, 699), bo.ess(x, 699))


def test_matrix_free_identities_equal_the_reference_partials():
    """The two contractions the reference forms InvGdG for (rmhmc.py:77, :105-107), evaluated without it.

    tr(G^-1 dG_d) = sum_n c_n x_nd (x_n^T G^-1 x_n) and p^T G^-1 dG_d G^-1 p = sum_n c_n x_nd (x_n . G^-1 p)^2 with
    c_n = v_n (1 - 2 p_n): what the CUDA engine's MATRIX_FREE partials mode computes (include/rmhmc_b200.h).
    """
    from riemannhamiltonianmontecarlo_b200 import datasets
    xx, t = datasets.shaped("australian")
    rng = np.random.default_rng(5)
    w = rng.normal(0, 0.4, (xx.shape[1], 1))
    mom = rng.normal(0, 3.0, (xx.shape[1], 1))
    _, p, v, g = bo.fisher_metric(xx, w)
    inv_g = np.linalg.inv(g)
    inv_g_dg, tr_ref = bo.metric_partials(xx, p, v, inv_g)                     # the reference's formulation
    u = inv_g.dot(mom)
    last_ref = np.array([bo._scalar(mom.T.dot(inv_g_dg[d]).dot(u)) for d in range(xx.shape[1])])
    c = v * (1 - 2 * p[:, 0])
    lev = np.einsum("na,ab,nb->n", xx, inv_g, xx)
    tr_mf = xx.T.dot(c * lev)
    quad_mf = xx.T.dot(c * xx.dot(u)[:, 0] ** 2)
    assert np.abs(tr_mf - tr_ref[:, 0]).max() < 1e-12 * np.abs(tr_ref).max()
    assert np.abs(quad_mf - last_ref).max() < 1e-12 * np.abs(last_ref).max()
