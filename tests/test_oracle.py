"""Self-tests of the numpy oracle (T0 of SURVEY.md section 4): the restated astropy formula
against its geometric meaning, model assembly, masking, and the random stream."""
import math

import numpy as np
import pytest

from oracle import lapf_oracle as orc


def test_abc_form_equals_rotated_gaussian():
    """astropy's a, b, c form == A exp(-((x'/sx)^2 + (y'/sy)^2)/2) with the counter-clockwise
    rotation x' = dx cos t + dy sin t, y' = -dx sin t + dy cos t."""
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[:40, :50].astype(np.float64)
    for _ in range(20):
        amp, x0, y0 = rng.uniform(1, 1e4), rng.uniform(5, 45), rng.uniform(5, 35)
        sx, sy, th = rng.uniform(1, 8), rng.uniform(1, 8), rng.uniform(-4, 4)
        xp = (xx - x0) * math.cos(th) + (yy - y0) * math.sin(th)
        yp = -(xx - x0) * math.sin(th) + (yy - y0) * math.cos(th)
        ref = amp * np.exp(-0.5 * ((xp / sx) ** 2 + (yp / sy) ** 2))
        np.testing.assert_allclose(orc.gaussian2d(xx, yy, amp, x0, y0, sx, sy, th), ref, rtol=1e-12, atol=1e-300)


def test_separable_and_integral_identities():
    yy, xx = np.mgrid[:201, :201].astype(np.float64)
    g = orc.gaussian2d(xx, yy, 3.0, 100.25, 99.5, 4.0, 4.0, 0.0)
    gx = np.exp(-0.5 * ((np.arange(201) - 100.25) / 4.0) ** 2)
    gy = np.exp(-0.5 * ((np.arange(201) - 99.5) / 4.0) ** 2)
    np.testing.assert_allclose(g, 3.0 * np.outer(gy, gx), rtol=1e-12, atol=1e-300)
    g2 = orc.gaussian2d(xx, yy, 7.0, 100.0, 100.0, 3.0, 5.0, 0.7)
    assert g2.sum() == pytest.approx(2 * math.pi * 7.0 * 3.0 * 5.0, rel=1e-9)
    # theta and theta + pi describe the same ellipse
    np.testing.assert_allclose(g2, orc.gaussian2d(xx, yy, 7.0, 100.0, 100.0, 3.0, 5.0, 0.7 + math.pi), rtol=1e-9)


@pytest.mark.parametrize("nbody", [2, 3])
def test_model_assembly_and_floor(nbody):
    lay = orc.layout_for(nbody)
    rng = np.random.default_rng(nbody)
    p = np.abs(rng.normal(size=lay.nparam)) + 1.0
    p[:2 * nbody] = rng.uniform(8, 24, 2 * nbody)
    for o in range(nbody):
        p[lay.i_amp(o)] = rng.uniform(100, 1000)
    m = orc.model_image(p, lay, 32, 32)
    # far from every source the model is the floor slot: p[12] in both layouts
    q = p.copy()
    for o in range(nbody):
        q[lay.i_amp(o)] = q[lay.bkgd_index]          # zero amplitudes: only the floor is left
    np.testing.assert_allclose(orc.model_image(q, lay, 32, 32), p[12], rtol=1e-14)
    assert lay.floor_index == 12
    assert (lay.names[12] == "sigmax2") if nbody == 2 else (lay.names[12] == "bkgd")
    # the opt-in floor uses bkgd
    m_fix = orc.model_image(p, lay, 32, 32, floor_index=lay.bkgd_index)
    np.testing.assert_allclose(m_fix - m, p[lay.bkgd_index] - p[12], rtol=1e-9, atol=1e-9)
    # amplitudes: narrow + wide at the centre of an isolated object add up to total - bkgd
    one = p.copy()
    for o in range(1, nbody):
        one[lay.i_amp(o)] = one[lay.bkgd_index]
    one[lay.i_dx] = one[lay.i_dy] = 0.0
    one[0], one[1] = 16.0, 16.0
    peak = orc.model_image(one, lay, 32, 32)[16, 16] - one[12]
    assert peak == pytest.approx(one[lay.i_amp(0)] - one[lay.bkgd_index], rel=1e-12)
    # an origin shift moves the grid, not the model values
    m_shift = orc.model_image(p + np.array([100 if i < 2 * nbody and i % 2 == 0 else (200 if i < 2 * nbody else 0)
                                            for i in range(lay.nparam)]), lay, 32, 32, origin=(100, 200))
    np.testing.assert_allclose(m_shift, m, rtol=1e-9)


def test_chi_square_masking_and_weight_form():
    rng = np.random.default_rng(5)
    data = rng.normal(size=(16, 16)) * 50 + 100
    data[3, 4] = 30000.0
    hdr = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
    mask, err = orc.frame_prep(data, hdr)
    assert mask.sum() == 1 and mask[3, 4]
    model = np.full_like(data, 100.0)
    c = orc.chi_squared(data, model, err, mask)
    # the numpy masked-array computation of the reference (apf_step2.py:135-136, :188)
    ma = np.ma.masked_greater(data, 0.8 * orc.saturation_level(hdr))
    assert c == pytest.approx(float(np.sum(((ma - model) / err) ** 2)), rel=1e-14)
    assert orc.chi_squared_weighted(data, model, orc.weight_map(data, hdr)) == pytest.approx(c, rel=1e-13)
    assert orc.chi_squared(model, model, err, mask) == 0.0
    hdr3 = {"itime": 2.0, "coadds": 4, "multisam": 8, "sampmode": 3}
    assert orc.saturation_level(hdr3) == 4 * 24000.0 * (1.0 - 0.1 * 7.0 / 2.0)
    assert orc.read_noise(hdr3) == (38.0 / math.sqrt(8.0)) * 2.0


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10."""
    assert orc.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert orc.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert orc.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_device_stream_moments():
    ks, zs, us = zip(*(orc.device_draws(7, 3, t, 16) for t in range(20000)))
    assert set(ks) == set(range(16))
    counts = np.bincount(ks, minlength=16)
    assert counts.min() > 1100 and counts.max() < 1400
    assert abs(np.mean(zs)) < 0.03 and abs(np.std(zs) - 1.0) < 0.03
    assert abs(np.mean(us) - 0.5) < 0.01 and min(us) >= 0.0 and max(us) < 1.0
    # streams of different walkers / seeds differ
    assert orc.device_draws(7, 3, 0, 16) != orc.device_draws(7, 4, 0, 16) != orc.device_draws(8, 3, 0, 16)


def test_run_chain_rules():
    """Row accounting (nan seed row, burn-in, thin), counters and the two stop rules."""
    lay = orc.layout_for(2)
    from olpefit_b200 import synth
    ox, oy = synth.stamp_origin(32)
    img, truth = synth.make_frame(0, 2, region=(oy, oy + 32, ox, ox + 32))
    img = img.astype(np.float64)
    hdr = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
    w = orc.weight_map(img, hdr)
    r = orc.run_chain(img, w, lay, truth, orc.PhiloxStream(1, 0, 16), origin=(ox, oy), n_updates=60, burn_in=10, thin=5)
    assert r.rows.shape == (1 + 11, 17) and np.all(np.isnan(r.rows[0])) and r.tries.sum() == 60
    assert np.all(r.accepts <= r.tries)
    r2 = orc.run_chain(img, w, lay, truth, orc.PhiloxStream(1, 0, 16), origin=(ox, oy), accept_min=3, burn_in=0)
    assert r2.tries.min() == 3 and r2.rows.shape[0] == r2.n_updates + 1
    # rejected proposals still append the (unchanged) state: rows change exactly `accepts` times
    prev = np.vstack([truth[None], r2.rows[1:-1, :-1]])
    changed = np.any(r2.rows[1:, :-1] != prev, axis=1)
    assert int(changed.sum()) == int(r2.accepts.sum()) < r2.n_updates


def test_block_factorisation_identity():
    """The algebra behind the device's factorised pixel loop, restated in
    float64: around the centre (xa, yb) of a 2 x 4 pixel block every component value is
    E_k * C_k,ij * T_c,ij with one exponential per block (E), lane constants (C) and block factors
    shared by the shape class (T).  Checked against the oracle's own component evaluation."""
    rng = np.random.default_rng(5)
    for _ in range(20):
        sx, sy, th = rng.uniform(1.5, 7.0), rng.uniform(1.5, 7.0), rng.uniform(-1.6, 1.6)
        a = 0.5 * (np.cos(th) ** 2 / sx ** 2 + np.sin(th) ** 2 / sy ** 2)
        b = 0.5 * (np.sin(2 * th) / sx ** 2 - np.sin(2 * th) / sy ** 2)
        c = 0.5 * (np.sin(th) ** 2 / sx ** 2 + np.cos(th) ** 2 / sy ** 2)
        sa, sb, sc = -a * np.log2(np.e), -b * np.log2(np.e), -c * np.log2(np.e)   # q = sa dx^2 + sb dx dy + sc dy^2
        amp = rng.uniform(10, 1e4)
        x0c, y0c = rng.uniform(20, 44, 2)            # first component of the class
        x0k, y0k = x0c + rng.normal(0, 6), y0c + rng.normal(0, 6)
        xa, yb = 4 * rng.integers(0, 16) + 1.5, 2 * rng.integers(0, 32) + 0.5
        dxa, dyb_c, dy0 = xa - x0k, yb - y0c, y0c - y0k
        E = 2.0 ** (sa * dxa ** 2 + sb * dxa * (yb - y0k) + sc * (yb - y0k) ** 2)
        for i in (-0.5, 0.5):
            for j in (-1.5, -0.5, 0.5, 1.5):
                C = amp * 2.0 ** (j * (sa * (2 * dxa + j) + sb * dy0) + i * (sb * dxa + 2 * sc * dy0))
                T = 2.0 ** (j * sb * (dyb_c + i) + 2 * i * sc * dyb_c + sc / 4)
                direct = amp * orc.gaussian2d(np.array([[xa + j]]), np.array([[yb + i]]), 1.0, x0k, y0k, sx, sy, th)[0, 0]
                assert E * C * T == pytest.approx(direct, rel=1e-11)
