"""Pin the numpy oracle to outputs of the reference's own code (tools/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import lapf_oracle as orc


@pytest.fixture(scope="module", params=[2, 3])
def gold(request, golden_dir):
    z = np.load(os.path.join(golden_dir, "reference_exec_%dbody.npz" % request.param), allow_pickle=True)
    return request.param, z


HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}


def test_tables_match_reference(gold):
    nbody, z = gold
    lay = orc.layout_for(nbody)
    assert np.array_equal(np.asarray(lay.widths), z["widths"])
    assert sorted(lay.lognorm) == sorted(z["lognorm"].tolist())
    norm = [i for i in range(lay.nparam) if i not in lay.lognorm]
    assert norm == sorted(z["norm"].tolist())
    assert orc.SIGMA_GUESS == float(z["sigma"])


def test_frame_prep_matches_reference(gold):
    _, z = gold
    img = z["image"].astype(np.float64)
    mask, err = orc.frame_prep(img, HEADER)
    assert orc.saturation_level(HEADER) == float(z["satlevel"])
    assert orc.read_noise(HEADER) == float(z["readnoise"])
    assert np.array_equal(mask, z["mask"])
    assert mask.sum() >= 1  # the synthetic frame carries hot pixels
    assert np.array_equal(err, z["err"])


def test_initial_parameters_match_reference(gold):
    nbody, z = gold
    lay = orc.layout_for(nbody)
    p = orc.initial_parameters(z["image"].astype(np.float64), z["guess"], lay)
    assert np.array_equal(p, z["p_init"][:-1])


def test_model_and_chi2_match_reference(gold):
    nbody, z = gold
    lay = orc.layout_for(nbody)
    img = z["image"].astype(np.float64)
    n = int(z["size"])
    mask, err = orc.frame_prep(img, HEADER)
    w = orc.weight_map(img, HEADER)
    for q, m_ref, c_ref in zip(z["vec_params"], z["vec_models"], z["vec_chi2"]):
        m = orc.model_image(q, lay, n, n)
        np.testing.assert_allclose(m, m_ref, rtol=1e-13, atol=0)
        assert orc.chi_squared(img, m, err, mask) == pytest.approx(c_ref, rel=1e-12)
        assert orc.chi_squared_weighted(img, m, w) == pytest.approx(c_ref, rel=1e-12)
    c0 = orc.chi_squared_weighted(img, orc.model_image(z["p_init"][:-1], lay, n, n), w)
    assert c0 == pytest.approx(float(z["chi_init"]), rel=1e-12)


def test_proposals_match_reference(gold):
    _, z = gold
    st = orc.NumpyStream(20190531)
    with np.errstate(all="ignore"):
        got_n = [orc.propose(st, 0, a, b, False) for a, b in z["prop_in"]]
        got_l = [orc.propose(st, 0, a, b, True) for a, b in z["prop_in"]]
    np.testing.assert_array_equal(np.array(got_n), z["prop_normal"])
    np.testing.assert_array_equal(np.array(got_l), z["prop_log"])
    # log10 of zero gives 10**-inf = 0; log10 of a negative value gives nan (apf_step2.py:67)
    assert got_l[3] == 0.0 and np.isnan(got_l[4])
    # accept rule continues on the same stream
    for (a, b), (yes, pa, dice) in zip(z["accept_in"], z["accept_out"]):
        assert orc.accept_rule(st, 0, a, b) == bool(yes)
    assert z["accept_out"][4, 0] == 0.0      # nan proposal is rejected
    assert z["accept_out"][5, 0] == 1.0      # overflow to inf is accepted


def test_loop_replays_reference_update_by_update(gold):
    nbody, z = gold
    lay = orc.layout_for(nbody)
    img = z["image"].astype(np.float64)
    w = orc.weight_map(img, HEADER)
    trace = z["loop_trace"]
    st = orc.NumpyStream(int(z["loop_seed"]))
    res = orc.run_chain(img, w, lay, z["p_init"][:-1], st, n_updates=len(trace), burn_in=0)
    assert res.rows.shape == (len(trace) + 1, lay.nparam + 1)
    assert np.all(np.isnan(res.rows[0]))
    np.testing.assert_allclose(res.rows[1:, :-1], trace[:, :-1], rtol=0, atol=0)
    np.testing.assert_allclose(res.rows[1:, -1], trace[:, -1], rtol=1e-12)
    assert np.array_equal(res.tries, z["loop_tries"])
    assert np.array_equal(res.accepts, z["loop_accepts"])
