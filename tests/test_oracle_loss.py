"""Pin oracle/pg_loss.py against the reference's PPO.loss / A2C.loss outputs and gradients
(tests/golden/loss.npz) and the known-answer vector of SURVEY.md 9.2."""
import numpy as np
import pytest

from oracle import pg_loss as L

CASES = ["kat", "kat_valid", "n37_a18", "n37_a18_valid", "a2c_cfg", "a2c_cfg_valid", "ppo_cfg",
         "ppo_cfg_valid", "n1"]


def load_case(g, name):
    d = {k: g[f"{name}/{k}"] for k in ("p_new", "p_old", "value", "action", "adv", "ret")}
    d["valid"] = g[f"{name}/valid"] if f"{name}/valid" in g.files else None
    d["clip"], d["c_v"], d["c_ent"] = (float(x) for x in g[f"{name}/hyper"])
    return d


@pytest.mark.parametrize("name", CASES)
def test_ppo_loss_matches_reference(golden, name):
    g = golden("loss")
    c = load_case(g, name)
    o = L.ppo_loss(c["p_new"], c["value"], c["p_old"], c["action"], c["ret"], c["adv"], c["valid"],
                   c["clip"], c["c_v"], c["c_ent"])
    assert [o["loss"], o["entropy"], o["perplexity"]] == list(g[f"{name}/ppo/scalars"])
    assert np.array_equal(o["grad_prob"], g[f"{name}/ppo/grad_prob"])
    assert np.array_equal(o["grad_value"], g[f"{name}/ppo/grad_value"])


@pytest.mark.parametrize("name", CASES)
def test_a2c_loss_matches_reference(golden, name):
    g = golden("loss")
    c = load_case(g, name)
    o = L.a2c_loss(c["p_new"], c["value"], c["action"], c["ret"], c["adv"], c["valid"], c["c_v"], c["c_ent"])
    assert [o["loss"], o["entropy"], o["perplexity"]] == list(g[f"{name}/a2c/scalars"])
    assert np.array_equal(o["grad_prob"], g[f"{name}/a2c/grad_prob"])
    assert np.array_equal(o["grad_value"], g[f"{name}/a2c/grad_value"])


def test_known_answers_survey_9_2(golden):
    g = golden("loss")
    c = load_case(g, "kat")
    o = L.ppo_loss(c["p_new"], c["value"], c["p_old"], c["action"], c["ret"], c["adv"], None, 0.1, 1.0, 0.01)
    np.testing.assert_allclose([o["loss"], o["entropy"], o["perplexity"]],
                               [-0.67151588, 0.90158784, 2.49443078], rtol=1e-6)
    np.testing.assert_allclose(o["grad_value"], [-0.125, -0.05, 0.125, 0.1], rtol=1e-6)
    np.testing.assert_allclose(o["grad_prob"][1, 0], 0.251223, rtol=1e-5)
    np.testing.assert_allclose(o["grad_prob"][3, 1], -0.3009658, rtol=1e-5)
    assert np.all(np.abs(o["grad_prob"][[0, 2]]) < 4e-3)  # clipped rows: entropy gradient only
    o = L.ppo_loss(c["p_new"], c["value"], c["p_old"], c["action"], c["ret"], c["adv"],
                   np.array([1, 1, 0, 1], np.float32), 0.1, 1.0, 0.01)
    np.testing.assert_allclose([o["loss"], o["entropy"], o["perplexity"]],
                               [-0.20155776, 0.98910648, 2.69435906], rtol=1e-6)
    assert np.all(o["grad_prob"][2] == 0) and o["grad_value"][2] == 0
    np.testing.assert_allclose(o["grad_value"], [-0.1666667, -0.0666667, 0, 0.1333333], rtol=1e-6)


def test_sample_categorical_inverse_cdf():
    p = np.array([[0.2, 0.5, 0.3], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], np.float32)
    assert list(L.sample_categorical(p, [0.0, 0.5, 0.5])) == [0, 0, 2]
    assert list(L.sample_categorical(p, [0.2, 0.999, 0.0])) == [1, 0, 2]
    assert list(L.sample_categorical(p, [0.71, 0.0, 0.999])) == [2, 0, 2]


def test_philox_known_answers():
    """oracle/philox.py against the Random123 known-answer vectors of Philox4x32-10 (kat_vectors: counter/key all
    zero -> 6627e8d5...; all ones -> 408f276d...; pi digits -> d16cfe09...): first output word."""
    from oracle.philox import philox4x32_10_word0, uniforms
    import numpy as np
    assert int(philox4x32_10_word0([0], 0, 0)[0]) == 0x6627e8d5
    # counter ffffffff x4, key ffffffff x2
    assert int(philox4x32_10_word0([0xFFFFFFFFFFFFFFFF], 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)[0]) == 0x408f276d
    # counter 243f6a88 85a308d3 13198a2e 03707344, key a4093822 299f31d0
    assert int(philox4x32_10_word0([0x85a308d3243f6a88], 0x0370734413198a2e, 0x299f31d0a4093822)[0]) == 0xd16cfe09
    u = uniforms(1000, 5, 42)
    assert u.dtype == np.float32 and float(u.min()) >= 0.0 and float(u.max()) < 1.0
