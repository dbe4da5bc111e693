"""Pin oracle/returns.py against the reference's own outputs (tests/golden/returns.npz,
made by tests/golden/make_golden.py from /root/reference) - bit-exact - and against the
hand-checkable vectors of SURVEY.md section 9.1."""
import numpy as np
import pytest

from oracle import returns as O

CASES = ["kat", "t1", "t2b1", "cfg1", "edges", "ragged", "sparse", "cfg2"]
GL = [(0.99, 1.0), (0.99, 0.98), (0.99, 0.95), (1.0, 0.9), (0.9, 0.8)]


def _inputs(g, name):
    return (g[f"{name}/reward"], g[f"{name}/value"], g[f"{name}/done"], g[f"{name}/bv"])


@pytest.mark.parametrize("name", CASES)
def test_gae_and_discount_bit_exact(golden, name):
    g = golden("returns")
    reward, value, done, bv = _inputs(g, name)
    n = 0
    for (gam, lam) in GL:
        key = f"{name}/g{gam}_l{lam}"
        if key + "/gae_adv" not in g:
            continue
        adv, ret = O.generalized_advantage_estimation(reward, value, done, bv, gam, lam)
        assert np.array_equal(adv, g[key + "/gae_adv"])
        assert np.array_equal(ret, g[key + "/gae_ret"])
        assert np.array_equal(O.discount_return(reward, done, bv, gam), g[key + "/disc_ret"])
        n += 1
    assert n >= 3


@pytest.mark.parametrize("name", CASES)
def test_valid_and_nstep_bit_exact(golden, name):
    g = golden("returns")
    reward, value, done, bv = _inputs(g, name)
    assert np.array_equal(O.valid_from_done(done), g[f"{name}/valid"])
    n = 0
    for key in g.files:
        if not key.startswith(name + "/n") or not key.endswith("/ret"):
            continue
        tag = key.split("/")[1]  # n3_t0_g0.99
        ns, tr, gam = tag.split("_")
        ret, done_n = O.discount_return_n_step(reward, done, int(ns[1:]), float(gam[1:]),
                                               do_truncated=bool(int(tr[1:])))
        assert ret.dtype == np.float32 and done_n.dtype == bool
        assert np.array_equal(ret, g[key])
        assert np.array_equal(done_n, g[key[:-4] + "/done_n"])
        n += 1
    assert n >= 1


@pytest.mark.parametrize("name", ["cfg1", "edges", "ragged", "cfg2"])
def test_process_returns(golden, name):
    g = golden("returns")
    reward, value, done, bv = _inputs(g, name)
    n = 0
    for lam in (1.0, 0.98):
        for mbr in (True, False):
            for norm in (False, True):
                key = f"{name}/pr_l{lam}_m{int(mbr)}_n{int(norm)}"
                if key + "/ret" not in g:
                    continue
                ret, adv, valid = O.process_returns(reward, done, value, bv, 0.99, lam,
                                                    use_valid=not mbr, normalize_advantage=norm)
                assert np.array_equal(ret, g[key + "/ret"])
                assert np.array_equal(adv, g[key + "/adv"])
                if mbr:
                    assert valid is None and key + "/valid" not in g
                else:
                    assert np.array_equal(valid, g[key + "/valid"])
                n += 1
    assert n >= 4


def test_known_answers_survey_9_1(golden):
    """Closed-form values from SURVEY.md 9.1 (T=4,B=2,gamma=.9,lambda=.8)."""
    g = golden("returns")
    reward, value, done, bv = _inputs(g, "kat")
    adv, ret = O.generalized_advantage_estimation(reward, value, done, bv, 0.9, 0.8)
    np.testing.assert_allclose(adv, [[0.536, 1.550557], [-0.2, 1.91744], [2.836, 0.552], [1.3, -1.4]],
                               rtol=1e-6)
    np.testing.assert_allclose(ret, [[1.036, 1.650557], [0.0, 2.21744], [2.836, 0.352], [2.3, -1.0]],
                               rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(O.discount_return(reward, done, bv, 0.9),
                               [[1.0, 1.881], [0.0, 2.09], [3.07, 0.1], [2.3, -1.0]], rtol=1e-6)
    assert np.array_equal(O.valid_from_done(done), [[1, 1], [1, 1], [0, 1], [0, 1]])
    r3, d3 = O.discount_return_n_step(reward, done, 3, 0.9)
    np.testing.assert_allclose(r3, [[1.0, 2.61], [0.0, 2.09]], rtol=1e-6)
    assert np.array_equal(d3, [[True, False], [True, True]])
    r3t, d3t = O.discount_return_n_step(reward, done, 3, 0.9, do_truncated=True)
    np.testing.assert_allclose(r3t, [[1, 2.61], [0, 2.09], [1.45, 0.1], [0.5, -1]], rtol=1e-6)
    assert np.array_equal(d3t, [[1, 0], [1, 1], [0, 1], [0, 1]])
