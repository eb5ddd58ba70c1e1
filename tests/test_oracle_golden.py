"""Pins the oracle (oracle/tinyda_oracle.py) against the reference's own outputs: every
golden fixture was produced by the unmodified reference under injected streams."""
import numpy as np
import pytest

import golden_io
from oracle import tinyda_oracle as orc

RTOL = 1e-11   # float64 restatement vs reference: only summation-order level differences


@pytest.mark.parametrize("name", golden_io.names())
def test_oracle_matches_reference(name):
    g = golden_io.load(name)
    spec = g["spec"]
    out, chains = orc.run_chains(spec, g["theta0"], g["z"], g["u"], g["iterations"], g["archive0"])
    for l in range(spec["n_levels"]):
        ref = g["ref"][l]
        assert out[l]["acc"].shape == ref["acc"].shape, (name, l)
        assert np.array_equal(out[l]["acc"], ref["acc"]), "accept/reject decisions differ at level %d" % l
        np.testing.assert_allclose(out[l]["theta"], ref["theta"], rtol=RTOL, atol=1e-13)
        np.testing.assert_allclose(out[l]["prior"], ref["prior"], rtol=RTOL, atol=1e-11)
        np.testing.assert_allclose(out[l]["like"], ref["like"], rtol=1e-9, atol=1e-9)
        if "F" in ref:
            np.testing.assert_allclose(out[l]["F"], ref["F"], rtol=1e-10, atol=1e-13)
    # identical stream consumption (control-flow dependent)
    consumed = np.array([[ch.S.nz, ch.S.nu] for ch in chains])
    assert np.array_equal(consumed, g["consumed"])


def test_oracle_matches_reference_on_the_long_cfg2_fixture():
    """BASELINE cfg2 at its real shape, 8 chains x 200 fine iterations (tests/golden/long/, made by
    make_golden.py --long from the unmodified reference): decisions of both levels identical, fine-level
    states and densities at summation-order accuracy."""
    g = golden_io.load("da_pcn_cfg2", long=True)
    out, chains = orc.run_chains(g["spec"], g["theta0"], g["z"], g["u"], g["iterations"], None)
    c_ref, f_ref = g["ref"]
    assert np.array_equal(out[0]["acc"], c_ref["acc"]) and np.array_equal(out[1]["acc"], f_ref["acc"])
    assert f_ref["acc"].shape == (8, 201) and c_ref["acc"].shape == (8, 2000)
    np.testing.assert_allclose(out[1]["theta"], f_ref["theta"], rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(out[1]["prior"], f_ref["prior"], rtol=RTOL, atol=1e-11)
    np.testing.assert_allclose(out[1]["like"], f_ref["like"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out[0]["like"], c_ref["like"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(np.array([[ch.S.nz, ch.S.nu] for ch in chains]), g["consumed"])


def test_dream_ensemble_sync_every_one_is_the_lock_step_rule_and_larger_values_differ():
    """DreamEnsemble(sync_every=K): K = 1 reproduces the dream_shared fixture (checked above through run_chains);
    K = 4 gives the same first step (every chain sees the initial rows only) and different chains afterwards (at the
    second step the lock-step rule already counts one more row per chain)."""
    import copy
    from oracle import tinyda_oracle as orc
    g = golden_io.load("dream_shared")
    spec4 = copy.deepcopy(g["spec"])
    spec4["proposal"]["sync_every"] = 4
    a, _ = orc.run_chains(g["spec"], g["theta0"], g["z"], g["u"], g["iterations"], g["archive0"])
    b, _ = orc.run_chains(spec4, g["theta0"], g["z"], g["u"], g["iterations"], g["archive0"])
    assert np.array_equal(a[0]["theta"][:, :2], b[0]["theta"][:, :2])
    assert not np.array_equal(a[0]["theta"], b[0]["theta"])
