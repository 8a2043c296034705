"""GPU parity tests: the CUDA path (through the C ABI) against the numpy oracle and the
reference-executed golden vectors.  Tolerance for the pointwise tests is the one BASELINE.json
states: 1e-5 relative (FP32 device arithmetic vs the reference's FP64)."""
import math
import os

import numpy as np
import pytest

from oracle import lapf_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5
HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}


@pytest.fixture(scope="module")
def gpu():
    import torch
    from olpefit_b200 import frame, model, sampler, synth
    assert torch.cuda.is_available()
    return {"torch": torch, "frame": frame, "model": model, "sampler": sampler, "synth": synth}


def _domain(gpu, image, origin, nbody):
    return gpu["frame"].prepare_domain(image, HEADER, origin=origin, nbody=nbody)


def _random_vectors(truth_local, nbody, n, rng, spread=1.0):
    out = []
    for _ in range(n):
        q = truth_local.copy()
        q[:2 * nbody] += rng.normal(0, 1.5 * spread, 2 * nbody)
        q[2 * nbody:2 * nbody + 2] += rng.normal(0, 0.3, 2)
        q[2 * nbody + 2:3 * nbody + 2] *= 10 ** rng.normal(0, 0.1, nbody)
        q[3 * nbody + 2] = rng.uniform(0.05, 0.6)
        q[3 * nbody + 3] *= 10 ** rng.normal(0, 0.2)
        q[3 * nbody + 4:3 * nbody + 8] *= 10 ** rng.normal(0, 0.08, 4)
        q[3 * nbody + 8:3 * nbody + 10] = rng.uniform(-1.6, 1.6, 2)
        # FP32-representable values, so device and oracle see identical inputs
        out.append(q.astype(np.float32).astype(np.float64))
    return np.array(out)


@pytest.mark.parametrize("nbody", [2, 3])
def test_k1_matches_reference_executed_vectors(gpu, golden_dir, nbody):
    z = np.load(os.path.join(golden_dir, "reference_exec_%dbody.npz" % nbody), allow_pickle=True)
    img = z["image"]
    dom = _domain(gpu, img, (0, 0), nbody)
    # frame preparation on the device equals the reference's mask + error map
    w_ref = np.where(z["mask"], 0.0, 1.0 / z["err"] ** 2)
    np.testing.assert_allclose(dom.weight[0].cpu().numpy(), w_ref, rtol=2e-7)
    assert np.array_equal(dom.weight[0].cpu().numpy() == 0, z["mask"])
    model, chi2 = dom.model_chi2(z["vec_params"], want_model=True)
    model, chi2 = model.cpu().numpy(), chi2.cpu().numpy()
    for i in range(len(z["vec_params"])):
        np.testing.assert_allclose(model[i], z["vec_models"][i], rtol=RTOL, atol=0)
        assert chi2[i] == pytest.approx(z["vec_chi2"][i], rel=RTOL)


@pytest.mark.parametrize("nbody", [2, 3])
@pytest.mark.parametrize("size", [32, 64, 128])
def test_k1_stamp_matches_oracle(gpu, nbody, size):
    synth = gpu["synth"]
    lay = orc.layout_for(nbody)
    ox, oy = synth.stamp_origin(size)
    img, truth = synth.make_frame(3, nbody, region=(oy, oy + size, ox, ox + size))
    dom = _domain(gpu, img, (ox, oy), nbody)
    w = orc.weight_map(img.astype(np.float64), HEADER)
    tl = truth.copy()
    tl[0:2 * nbody:2] -= ox
    tl[1:2 * nbody:2] -= oy
    vecs = _random_vectors(tl, nbody, 48, np.random.default_rng(7 * size + nbody))
    vecs[:, 0:2 * nbody:2] += ox     # frame coordinates (integers added to FP32 values stay exact)
    vecs[:, 1:2 * nbody:2] += oy
    model, chi2 = dom.model_chi2(vecs, want_model=True)
    model, chi2 = model.cpu().numpy(), chi2.cpu().numpy()
    worst = 0.0
    for i, q in enumerate(vecs):
        m = orc.model_image(q, lay, size, size, origin=(ox, oy))
        rel = np.max(np.abs(model[i] - m) / np.abs(m))
        worst = max(worst, rel)
        assert rel < RTOL
        c = orc.chi_squared_weighted(img.astype(np.float64), m, w)
        assert chi2[i] == pytest.approx(c, rel=RTOL)
    # chi2-only call (no model store) gives the identical number
    _, chi2b = dom.model_chi2(vecs, want_model=False)
    assert np.array_equal(chi2b.cpu().numpy(), chi2)
    print("worst per-pixel relative error S=%d nbody=%d: %.2e" % (size, nbody, worst))


def test_k1_multi_frame_and_fix_bkgd(gpu):
    synth = gpu["synth"]
    lay = orc.layout_for(2)
    size, nf = 64, 5
    stamps, origins = synth.make_stamps(nf, size)
    dom = gpu["frame"].prepare_domain(stamps, HEADER, origin=origins, nbody=2, floor_index=9)
    rng = np.random.default_rng(11)
    frame_of = rng.integers(0, nf, 40).astype(np.int32)
    vecs = np.array([synth.truth_parameters(2, int(f)) for f in frame_of]).astype(np.float32).astype(np.float64)
    _, chi2 = dom.model_chi2(vecs, frame_of=frame_of)
    chi2 = chi2.cpu().numpy()
    for i, f in enumerate(frame_of):
        img = stamps[f].astype(np.float64)
        m = orc.model_image(vecs[i], lay, size, size, origin=tuple(origins[f]), floor_index=9)
        assert chi2[i] == pytest.approx(orc.chi_squared_weighted(img, m, orc.weight_map(img, HEADER)), rel=RTOL)


@pytest.mark.parametrize("nbody", [2, 3])
def test_k1_full_frame_matches_oracle(gpu, nbody):
    """The reference's own pixel domain: the whole 1024 x 1024 frame (apf_step2.py:94,237)."""
    synth = gpu["synth"]
    lay = orc.layout_for(nbody)
    img, truth = synth.make_frame(0, nbody)
    dom = _domain(gpu, img, (0, 0), nbody)
    vecs = _random_vectors(truth, nbody, 3, np.random.default_rng(5), spread=0.3)
    vecs = np.vstack([truth.astype(np.float32).astype(np.float64)[None], vecs])
    model, chi2 = dom.model_chi2(vecs, want_model=True)
    model, chi2 = model.cpu().numpy(), chi2.cpu().numpy()
    img64 = img.astype(np.float64)
    w = orc.weight_map(img64, HEADER)
    for i, q in enumerate(vecs):
        m = orc.model_image(q, lay, 1024, 1024)
        assert np.max(np.abs(model[i] - m) / np.abs(m)) < RTOL
        assert chi2[i] == pytest.approx(orc.chi_squared_weighted(img64, m, w), rel=RTOL)


def test_k1_ragged_domain_and_empty_batch(gpu):
    synth = gpu["synth"]
    lay = orc.layout_for(2)
    ox, oy = synth.stamp_origin(64)
    img, truth = synth.make_frame(1, 2, region=(oy, oy + 50, ox, ox + 77))   # 50 x 77: generic path
    dom = _domain(gpu, img, (ox, oy), 2)
    q = truth.astype(np.float32).astype(np.float64)
    model, chi2 = dom.model_chi2(q[None], want_model=True)
    m = orc.model_image(q, lay, 50, 77, origin=(ox, oy))
    assert np.max(np.abs(model[0].cpu().numpy() - m) / np.abs(m)) < RTOL
    img64 = img.astype(np.float64)
    assert chi2.item() == pytest.approx(orc.chi_squared_weighted(img64, m, orc.weight_map(img64, HEADER)), rel=RTOL)
    _, none = dom.model_chi2(np.zeros((0, 16)))
    assert none.numel() == 0


def test_k1_nan_and_degenerate_parameters(gpu):
    """nan in, nan out (the accept rule then rejects, apf_step2.py:141-146); zero amplitude is fine."""
    synth = gpu["synth"]
    ox, oy = synth.stamp_origin(32)
    img, truth = synth.make_frame(0, 2, region=(oy, oy + 32, ox, ox + 32))
    dom = _domain(gpu, img, (ox, oy), 2)
    a = truth.copy(); a[7] = np.nan
    b = truth.copy(); b[10] = np.nan
    c = truth.copy(); c[7] = c[9]          # companion amplitude equals background: zero component
    _, chi2 = dom.model_chi2(np.array([a, b, c]))
    chi2 = chi2.cpu().numpy()
    assert np.isnan(chi2[0]) and np.isnan(chi2[1]) and np.isfinite(chi2[2])


# ---------------------------------------------------------------------------------------------
# sampler
# ---------------------------------------------------------------------------------------------
def _sampler_setup(gpu, nbody, size, n_frames=1):
    synth = gpu["synth"]
    stamps, origins = synth.make_stamps(n_frames, size, nbody)
    dom = gpu["frame"].prepare_domain(stamps, HEADER, origin=origins, nbody=nbody)
    guess = synth.step1_guess(stamps[0], nbody, origin=tuple(origins[0]))
    p0 = gpu["frame"].initial_parameters(stamps[0], guess, nbody, origin=tuple(origins[0]))
    return dom, stamps, origins, p0


@pytest.mark.parametrize("nbody,size,team", [(2, 32, 1), (2, 64, 1), (3, 32, 1), (2, 128, 1), (3, 64, 1),
                                             (2, 64, 4), (2, 64, 16), (2, 32, 4), (3, 128, 4), (2, 128, 16)])
def test_sampler_replays_oracle_stream(gpu, nbody, size, team):
    """Same Philox stream on both sides: the device chain must follow the float64 oracle chain
    update by update (same parameter picked, same decision, same values) until FP32 rounding of
    chi-square flips a borderline accept -- which must not happen early."""
    lay = orc.layout_for(nbody)
    dom, stamps, origins, p0 = _sampler_setup(gpu, nbody, size)
    n_upd, walkers, seed = 240, 3, 1234
    init = np.tile(p0, (walkers, 1))
    with gpu["sampler"].GibbsSampler(dom, init, seed=seed, burn_in=0, thin=1, id_base=5, id_stride=3,
                                     team_warps=team) as s:
        chain = s.run(n_upd).cpu().numpy()
        st, tries, accepts = (t.cpu().numpy() for t in s.state())
    img = stamps[0].astype(np.float64)
    w = orc.weight_map(img, HEADER)
    for wi in range(walkers):
        stream = orc.PhiloxStream(seed, 5 + 3 * wi, lay.nparam)
        res = orc.run_chain(img, w, lay, p0, stream, origin=tuple(origins[0]), n_updates=n_upd, burn_in=0)
        ref = res.rows[1:]
        dev = chain[:, wi, :]
        same = np.all(np.isclose(dev[:, :-1], ref[:, :-1], rtol=1e-9, atol=1e-12), axis=1)
        first_bad = int(np.argmin(same)) if not same.all() else n_upd
        assert first_bad >= 120, "device chain left the oracle chain at update %d" % first_bad
        np.testing.assert_allclose(dev[:first_bad, -1], ref[:first_bad, -1], rtol=RTOL)
        if first_bad == n_upd:
            assert np.array_equal(tries[wi], res.tries) and np.array_equal(accepts[wi], res.accepts)
            np.testing.assert_allclose(st[wi, :-1], res.params, rtol=1e-9)
        assert tries[wi].sum() == n_upd


def test_sampler_initial_chi2_and_counters(gpu):
    lay = orc.layout_for(2)
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 64)
    img = stamps[0].astype(np.float64)
    c0 = orc.chi_squared_weighted(img, orc.model_image(p0, lay, 64, 64, origin=tuple(origins[0])),
                                  orc.weight_map(img, HEADER))
    with gpu["sampler"].GibbsSampler(dom, np.tile(p0, (70, 1)), seed=9) as s:
        st, tries, accepts = s.state()
        assert st[:, -1].cpu().numpy() == pytest.approx(c0, rel=RTOL)     # apf_step2.py:283-289
        assert int(tries.sum()) == 0 and int(accepts.sum()) == 0           # apf_step2.py:276
        s.run(64, record=False)
        st, tries, accepts = s.state()
        assert np.all(tries.sum(dim=1).cpu().numpy() == 64)
        assert bool((accepts <= tries).all())
        stats = s.stats()
        assert np.array_equal(stats["tries"].cpu().numpy(), tries.sum(dim=0).cpu().numpy())
        assert np.array_equal(stats["accepts"].cpu().numpy(), accepts.sum(dim=0).cpu().numpy())
        assert int(stats["min_tries"]) == int(tries.min())
        assert s.count == 64
        # exponentials really evaluated: everything, minus what far-field culling may skip
        full = 64 * 70 * 64 * 64 * 4
        assert 0.5 * full < int(stats["exps"]) <= full
    off = gpu["model"].PixelDomain(dom.data, dom.weight, dom.origin, nbody=2, cull=False)
    with gpu["sampler"].GibbsSampler(off, np.tile(p0, (3, 1)), seed=9) as s:
        s.run(10, record=False)
        assert int(s.stats()["exps"]) == 10 * 3 * 64 * 64 * 4


def test_sampler_rows_burn_in_thin_and_split_runs(gpu):
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 32, n_frames=3)
    torch = gpu["torch"]
    walkers = 37
    init = np.tile(p0, (walkers, 1))
    frame_of = (np.arange(walkers) % 3).astype(np.int32)
    kw = dict(seed=77, burn_in=10, thin=4)
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, **kw) as a:
        assert a.rows_for(9) == 0 and a.rows_for(10) == 1 and a.rows_for(50) == 11
        whole = a.run(50)
        sa = a.state()
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, **kw) as b:
        parts = [b.run(7), b.run(13), b.run(30)]
        sb = b.state()
    assert whole.shape == (11, walkers, 17)
    assert torch.equal(whole, torch.cat(parts, dim=0))          # bitwise: launches are seamless
    for x, y in zip(sa, sb):
        assert torch.equal(x, y)
    # reference rule with thin = 1 and burn-in 0: one row per update, also for rejected proposals
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=77, burn_in=0, thin=1) as c:
        assert c.run(20).shape[0] == 20


def test_sampler_sharding_invariance(gpu):
    """Chains depend on (seed, global walker id) only: 1 shard vs 2 interleaved shards, bitwise."""
    torch = gpu["torch"]
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 32, n_frames=2)
    walkers = 24
    init = np.tile(p0, (walkers, 1))
    frame_of = (np.arange(walkers) % 2).astype(np.int32)
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=3) as s:
        whole = s.run(40)
    halves = []
    for r in range(2):
        with gpu["sampler"].GibbsSampler(dom, init[r::2], frame_of[r::2], seed=3, id_base=r, id_stride=2) as s:
            halves.append(s.run(40))
    assert torch.equal(whole[:, 0::2], halves[0]) and torch.equal(whole[:, 1::2], halves[1])


def test_sampler_negative_log_parameter_is_never_accepted(gpu):
    """log10 of a negative value is nan in the reference (apf_step2.py:67): such proposals are
    always rejected, everything else keeps moving."""
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 32)
    p = p0.copy()
    p[7] = -5.0                      # companion amplitude (log10-proposed) negative
    with gpu["sampler"].GibbsSampler(dom, p[None], seed=1) as s:
        s.run(400, record=False)
        st, tries, accepts = (t.cpu().numpy() for t in s.state())
    assert tries[0, 7] > 0 and accepts[0, 7] == 0 and st[0, 7] == -5.0
    assert accepts[0].sum() > 0 and np.isfinite(st[0, -1])


@pytest.mark.parametrize("size", [32, 64])
def test_nan_proposals_cost_no_pass_and_change_nothing(gpu, size):
    """The batched kernel skips the pass of a proposal that is nan by construction (DESIGN.md section
    4).  Walkers with a negative log-proposed parameter sit between walkers without (on 32-pixel
    stamps they share passes with them): the parameter never moves, every other one does, the state's
    chi-square is the stateless operator's bit for bit, and every chain is the chain the walker has
    when it runs alone."""
    torch = gpu["torch"]
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, size)
    walkers = 71
    init = np.tile(p0, (walkers, 1))
    neg = np.arange(walkers) % 2 == 0
    init[neg, 9] = -2.5                               # background: log10-proposed (apf_step2.py:217), negative start
    with gpu["sampler"].GibbsSampler(dom, init, seed=9, id_base=300, team_warps=1) as s:
        chain = s.run(200)
        st, tries, accepts = (t.cpu().numpy() for t in s.state())
    assert (tries[neg, 9] > 0).all() and (accepts[neg, 9] == 0).all() and (st[neg, 9] == -2.5).all()
    assert accepts[~neg, 9].sum() > 0
    assert (accepts.sum(axis=1) > 5).all() and np.isfinite(st[:, -1]).all()
    _, c = dom.model_chi2(np.ascontiguousarray(st[:, :-1]))
    assert np.array_equal(c.cpu().numpy(), st[:, -1])
    for w in (0, 1, 36, 69, 70):
        with gpu["sampler"].GibbsSampler(dom, init[w:w + 1], seed=9, id_base=300 + w, team_warps=1) as s:
            assert torch.equal(chain[:, w], s.run(200)[:, 0]), "walker %d" % w


def test_sampler_moments_match_chain(gpu):
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 32, n_frames=2)
    walkers = 20
    frame_of = (np.arange(walkers) % 2).astype(np.int32)
    with gpu["sampler"].GibbsSampler(dom, np.tile(p0, (walkers, 1)), frame_of, seed=5, burn_in=20, thin=2) as s:
        chain = s.run(200).cpu().numpy()
        st = s.stats()
    mom = st["moments"].cpu().numpy()
    assert int(st["rows"]) == chain.shape[0]
    assert st["walkers_per_frame"].cpu().numpy().tolist() == [10, 10]
    for f in range(2):
        sub = chain[:, frame_of == f, :]                       # [rows, walkers_f, P+1]
        means = sub.mean(axis=0) - mom[f, :, 0]                # centred on the per-frame reference
        np.testing.assert_allclose(mom[f, :-1, 0], p0)         # ... which is the starting point here
        np.testing.assert_allclose(mom[f, :-1, 1], means.sum(axis=0)[:-1], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(mom[f, :-1, 2], (means ** 2).sum(axis=0)[:-1], rtol=1e-7, atol=1e-14)
        np.testing.assert_allclose(mom[f, :, 3], (sub.std(axis=0) ** 2).sum(axis=0), rtol=1e-6, atol=1e-12)


def test_team_mode_agrees_with_single_warp_mode(gpu):
    """Several warps per walker change only the order of the FP64 partial sums: chi-square agrees
    to rounding, and the chains stay together until a borderline accept flips (not within the first
    updates).  Split runs and sharding stay bitwise reproducible for a fixed team size."""
    torch = gpu["torch"]
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 64, n_frames=2)
    walkers = 21
    init = np.tile(p0, (walkers, 1))
    frame_of = (np.arange(walkers) % 2).astype(np.int32)
    out = {}
    for team in (1, 4, 16):
        with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=8, team_warps=team) as s:
            out[team] = s.run(150)
            st, tries, acc = s.state()
            assert int(tries.sum()) == 150 * walkers
        with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=8, team_warps=team) as s:
            again = torch.cat([s.run(60), s.run(90)], dim=0)
        assert torch.equal(out[team], again)
    for team in (4, 16):
        a, b = out[1].cpu().numpy(), out[team].cpu().numpy()
        np.testing.assert_allclose(a[:40, :, :-1], b[:40, :, :-1], rtol=1e-9)   # same decisions, same values
        np.testing.assert_allclose(a[:40, :, -1], b[:40, :, -1], rtol=1e-6)     # chi-square: FP32 partials regrouped
        assert np.mean(np.isclose(a[..., :-1], b[..., :-1], rtol=1e-9)) > 0.9


def test_device_random_stream_equals_oracle_stream(gpu):
    """The Philox stream is defined once and restated by the oracle: same
    parameter index, same normal and uniform (to double rounding) for every update."""
    import math
    torch = gpu["torch"]
    from olpefit_b200 import _lib
    lib = _lib.load()
    for seed, walker, t0, npar in ((1234, 5, 0, 16), (2 ** 40 + 17, 70000, 2 ** 33, 19)):
        n = 300
        k = torch.empty(n, dtype=torch.int32, device="cuda")
        z = torch.empty(n, dtype=torch.float64, device="cuda")
        lnu = torch.empty(n, dtype=torch.float64, device="cuda")
        _lib.check(lib.lapf_philox_draws(seed, walker, t0, n, npar, k.data_ptr(), z.data_ptr(), lnu.data_ptr(), None))
        k, z, lnu = k.cpu().numpy(), z.cpu().numpy(), lnu.cpu().numpy()
        for i in range(n):
            ko, zo, uo = orc.device_draws(seed, walker, t0 + i, npar)
            assert k[i] == ko
            assert z[i] == pytest.approx(zo, rel=1e-12, abs=1e-14)
            assert lnu[i] == pytest.approx(math.log(uo) if uo > 0 else -math.inf, rel=1e-13)
        assert set(k.tolist()) == set(range(npar))


# ---------------------------------------------------------------------------------------------
# whole-frame chi-square at stamp cost, and full-size consistency
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbody,size,rtol", [(2, 128, 3e-7), (3, 128, 3e-7), (2, 64, 3e-6)])
def test_whole_frame_domain_equals_reference_domain(gpu, nbody, size, rtol):
    """The reference sums chi-square over the whole 1024 x 1024 frame (apf_step2.py:94,134-137).
    Cut-out + exact sums of the outside pixels against the constant floor gives the same number,
    up to the Gaussian wings the cut-out truncates: nothing at 128 pixels, a few 1e-7 of chi-square
    at 64 pixels (wide component, sigma ~ 6.4 px, cut at 5 sigma)."""
    synth = gpu["synth"]
    lay = orc.layout_for(nbody)
    img, truth = synth.make_frame(4, nbody)
    ox, oy = synth.stamp_origin(size)
    full = gpu["frame"].prepare_domain(img, HEADER, nbody=nbody)                       # generic kernel
    cut = gpu["frame"].prepare_domain(img, HEADER, size=size, cut=(ox, oy), nbody=nbody, whole_frame=True)
    assert cut.outside is not None and tuple(cut.origin[0].tolist()) == (ox, oy)
    vecs = np.vstack([truth.astype(np.float32).astype(np.float64)[None],
                      _random_vectors(truth, nbody, 5, np.random.default_rng(3), spread=0.3)])
    _, c_full = full.model_chi2(vecs)
    _, c_cut = cut.model_chi2(vecs)
    np.testing.assert_allclose(c_cut.cpu().numpy(), c_full.cpu().numpy(), rtol=rtol)
    img64 = img.astype(np.float64)
    w = orc.weight_map(img64, HEADER)
    for q, c in zip(vecs[:3], c_cut.cpu().numpy()):
        ref = orc.chi_squared_weighted(img64, orc.model_image(q, lay, 1024, 1024), w)
        assert c == pytest.approx(ref, rel=rtol)


def test_sampler_on_whole_frame_domain_replays_full_frame_oracle(gpu):
    """The sampler with the outside sums follows the oracle run on the FULL frame, i.e. the
    reference's own pixel domain, update by update."""
    synth = gpu["synth"]
    lay = orc.layout_for(2)
    img, truth = synth.make_frame(0, 2)
    ox, oy = synth.stamp_origin(64)
    ox, oy = synth.stamp_origin(128)
    dom = gpu["frame"].prepare_domain(img, HEADER, size=128, cut=(ox, oy), nbody=2, whole_frame=True)
    guess = synth.step1_guess(img, 2, sky_xy=(100, 120))
    p0 = gpu["frame"].initial_parameters(img, guess, 2)
    n_upd = 48
    with gpu["sampler"].GibbsSampler(dom, p0[None], seed=99) as s:
        chain = s.run(n_upd).cpu().numpy()[:, 0, :]
    img64 = img.astype(np.float64)
    res = orc.run_chain(img64, orc.weight_map(img64, HEADER), lay, p0, orc.PhiloxStream(99, 0, 16),
                        n_updates=n_upd, burn_in=0)
    np.testing.assert_allclose(chain[:, :-1], res.rows[1:, :-1], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(chain[:, -1], res.rows[1:, -1], rtol=1e-7)
    assert res.accepts.sum() > 5


# walkers per CTA item of the batched kernel (warps x walkers per warp, LAPF_DISPATCH_BATCH in lapf.cu)
BATCH_CHUNK = {(2, 32): 512, (2, 64): 512, (2, 128): 192, (3, 32): 512, (3, 64): 512, (3, 128): 144}


@pytest.mark.parametrize("nbody,size,walkers,frames,n_upd", [
    (2, 64, 65536, 100, 48),                       # BASELINE configs[2] at full size
    (2, 32, 148 * (2 * 512 + 37), 5, 32),          # every CTA: two full items (32 walkers per warp) + a ragged one
    (3, 32, 148 * (2 * 512 + 37), 5, 32),
    (3, 64, 148 * (512 + 37), 3, 32),              # config 4's kernel, full warps
    (2, 128, 148 * (192 + 29), 3, 16),             # TMEM holds the weight plane only (TM = 2), 16 walkers per warp
    (3, 128, 148 * (144 + 29), 3, 16),             # 12 walkers per warp
])
def test_full_size_batch_state_is_consistent_with_k1(gpu, nbody, size, walkers, frames, n_upd):
    """The batched kernel at the occupancy the benchmark times (many walkers per warp, coefficient
    images in shared memory, TMEM pixel store), every shape: after a run, the chi-square every
    walker carries equals the stateless operator K1 applied to its parameters -- bit for bit --
    (apf_step2.py:314-327: a walker's chi-square is that of its vector), every update was counted
    once, and the batch statistics add up."""
    torch = gpu["torch"]
    synth = gpu["synth"]
    W, F, S = walkers, frames, size
    P = 3 * nbody + 10
    assert W >= 148 * BATCH_CHUNK[(nbody, size)] or W == 65536
    stamps, origins = synth.make_stamps(F, S, nbody)
    dom = gpu["frame"].prepare_domain(stamps, HEADER, origin=origins, nbody=nbody)
    frame_of = (np.arange(W) % F).astype(np.int32)
    p_frame = np.array([synth.truth_parameters(nbody, f) for f in range(F)])
    thin = 16
    with gpu["sampler"].GibbsSampler(dom, p_frame[frame_of], frame_of, seed=31, burn_in=0, thin=thin) as s:
        chain = s.run(n_upd)
        st, tries, acc = s.state()
        stats = s.stats()
    assert chain.shape == (n_upd // thin, W, P + 1)
    assert torch.equal(chain[-1], st)                                     # last recorded row = final state
    _, chi = dom.model_chi2(st[:, :P], frame_of=frame_of)
    assert torch.equal(chi, st[:, P])                                     # K1 == K2, bitwise
    assert bool((tries.sum(dim=1) == n_upd).all()) and bool((acc <= tries).all())
    assert int(stats["tries"].sum()) == n_upd * W
    assert 0.05 < float(acc.sum()) / float(tries.sum()) < 0.9             # chains move (a wrong chi-square freezes them)
    assert stats["walkers_per_frame"].cpu().numpy().tolist() == np.bincount(frame_of, minlength=F).tolist()
    # chi-square per pixel of chains started at the in-model truth stays of order one
    assert 0.8 < float(st[:, P].median()) / (S * S) < 1.3


@pytest.mark.parametrize("nbody,size,walkers", [(2, 64, 600), (3, 64, 600), (2, 32, 600), (3, 32, 600),
                                                (2, 128, 230), (3, 128, 170)])
def test_dense_warps_replay_oracle_stream(gpu, nbody, size, walkers):
    """The oracle replay of test_sampler_replays_oracle_stream with MANY walkers per warp (more
    than one CTA item of one frame: full lanes, the lane == i pick-up of the passes, a ragged second
    item).  Walkers from different warps and lanes are replayed by the float64 oracle on the same
    Philox stream; the device chain must follow update by update."""
    lay = orc.layout_for(nbody)
    dom, stamps, origins, p0 = _sampler_setup(gpu, nbody, size)
    assert walkers > BATCH_CHUNK[(nbody, size)]
    n_upd, seed = 160, 4321
    init = np.tile(p0, (walkers, 1))
    with gpu["sampler"].GibbsSampler(dom, init, seed=seed, burn_in=0, thin=1, id_base=11, id_stride=2) as s:
        chain = s.run(n_upd).cpu().numpy()
        st, tries, accepts = (t.cpu().numpy() for t in s.state())
    assert np.all(tries.sum(axis=1) == n_upd)
    img = stamps[0].astype(np.float64)
    w = orc.weight_map(img, HEADER)
    chunk = BATCH_CHUNK[(nbody, size)]
    picks = sorted({0, 1, 15, 16, 17, chunk // 2 + 3, chunk - 1, chunk, chunk + 5, walkers - 1})
    early = 0
    for wi in picks:
        gid = 11 + 2 * wi
        stream = orc.PhiloxStream(seed, gid, lay.nparam)
        res = orc.run_chain(img, w, lay, p0, stream, origin=tuple(origins[0]), n_updates=n_upd, burn_in=0,
                            record_trace=True)
        ref = res.rows[1:]
        dev = chain[:, wi, :]
        same = np.all(np.isclose(dev[:, :-1], ref[:, :-1], rtol=1e-9, atol=1e-12), axis=1)
        first_bad = int(np.argmin(same)) if not same.all() else n_upd
        np.testing.assert_allclose(dev[:first_bad, -1], ref[:first_bad, -1], rtol=RTOL)
        if first_bad == n_upd:
            assert np.array_equal(tries[wi], res.tries) and np.array_equal(accepts[wi], res.accepts)
            continue
        # the chains part at update first_bad: legitimate only if FP32 rounding of chi-square (within
        # the stated 1e-5) can flip that decision, i.e. the float64 margin of the accept rule is tiny
        k, new, chi_t, ok = res.trace[first_bad]
        if first_bad > 0:
            chi_c = ref[first_bad - 1, -1]
        else:
            chi_c = orc.chi_squared_weighted(img, orc.model_image(p0, lay, size, size, origin=tuple(origins[0])), w)
        u = orc.device_draws(seed, gid, first_bad, lay.nparam)[2]
        margin = abs(math.log(u) + 0.5 * (chi_t - chi_c))
        assert margin <= 0.5 * RTOL * (abs(chi_t) + abs(chi_c)), (
            "walker %d left the oracle chain at update %d with a clear decision (margin %.3g)" % (wi, first_bad, margin))
        early += first_bad < 80
    assert early <= 2, "%d of %d replayed walkers met a borderline decision before update 80" % (early, len(picks))


@pytest.mark.parametrize("nbody,size", [(2, 64), (2, 128), (3, 128)])
def test_far_field_culling_changes_nothing_visible(gpu, nbody, size):
    """Components are skipped only where they are provably below 2^-24 of the floor: the model
    image and chi-square with and without the culling agree to FP32 rounding, and a nan still
    reaches chi-square."""
    synth, model = gpu["synth"], gpu["model"]
    ox, oy = synth.stamp_origin(size)
    img, truth = synth.make_frame(2, nbody, region=(oy, oy + size, ox, ox + size))
    on = _domain(gpu, img, (ox, oy), nbody)
    off = model.PixelDomain(on.data, on.weight, on.origin, nbody=nbody, cull=False)
    tl = truth.copy()
    tl[0:2 * nbody:2] -= ox
    tl[1:2 * nbody:2] -= oy
    vecs = _random_vectors(tl, nbody, 40, np.random.default_rng(size + nbody))
    vecs[:, 0:2 * nbody:2] += ox
    vecs[:, 1:2 * nbody:2] += oy
    vecs[5, 0] += 40.0                       # a source outside the stamp
    vecs[6, 2 * nbody + 2] = vecs[6, 3 * nbody + 3]      # zero-amplitude star
    m_on, c_on = on.model_chi2(vecs, want_model=True)
    m_off, c_off = off.model_chi2(vecs, want_model=True)
    m_on, m_off = m_on.cpu().numpy(), m_off.cpu().numpy()
    # skipped terms are below half an ulp of the pixel they would be added to (amplitudes are
    # non-negative here) and the partial sums keep their order: identical to the last bit
    assert np.array_equal(m_on, m_off)
    assert np.array_equal(c_on.cpu().numpy(), c_off.cpu().numpy())
    bad = vecs[:2].copy()
    bad[0, 3 * nbody + 4] = np.nan
    bad[1, 2 * nbody + 3] = np.nan
    _, c_bad = on.model_chi2(bad)
    assert bool(np.all(np.isnan(c_bad.cpu().numpy())))


def test_reference_host_loop_with_device_operator(gpu):
    """INTEGRATION.md level 2: the reference's own host loop (here: the oracle's restatement of
    apf_step2.py:300-351 on numpy's Mersenne-Twister stream) with build_analytical_model +
    chi_squared swapped for the device operator follows the all-CPU chain."""
    synth, model = gpu["synth"], gpu["model"]
    lay = orc.layout_for(2)
    ox, oy = synth.stamp_origin(64)
    img, truth = synth.make_frame(0, 2, region=(oy, oy + 64, ox, ox + 64))
    dom = _domain(gpu, img, (ox, oy), 2)
    img64 = img.astype(np.float64)
    w = orc.weight_map(img64, HEADER)
    guess = synth.step1_guess(img, 2, origin=(ox, oy))
    p0 = gpu["frame"].initial_parameters(img, guess, 2, origin=(ox, oy))
    m = model.build_analytical_model(p0, dom).cpu().numpy()
    ref = orc.model_image(p0, lay, 64, 64, origin=(ox, oy))
    assert np.max(np.abs(m - ref) / np.abs(ref)) < RTOL
    assert model.chi_squared(p0, dom) == pytest.approx(orc.chi_squared_weighted(img64, ref, w), rel=RTOL)
    cpu = orc.run_chain(img64, w, lay, p0, orc.NumpyStream(77), origin=(ox, oy), n_updates=150)
    dev = orc.run_chain(img64, w, lay, p0, orc.NumpyStream(77), origin=(ox, oy), n_updates=150,
                        chi2_fn=lambda q: model.chi_squared(q, dom))
    assert np.array_equal(dev.tries, cpu.tries) and np.array_equal(dev.accepts, cpu.accepts)
    np.testing.assert_array_equal(dev.rows[1:, :-1], cpu.rows[1:, :-1])      # same decisions, same values
    np.testing.assert_allclose(dev.rows[1:, -1], cpu.rows[1:, -1], rtol=RTOL)


def test_checkpoint_resume_is_exact_and_widths_can_be_retuned(gpu):
    torch = gpu["torch"]
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 32, n_frames=2)
    walkers = 9
    init = np.tile(p0, (walkers, 1))
    fo = (np.arange(walkers) % 2).astype(np.int32)
    kw = dict(seed=4, burn_in=5, thin=3)
    with gpu["sampler"].GibbsSampler(dom, init, fo, **kw) as a:
        whole = a.run(90)
        sa, sta = a.state(), a.stats()
    with gpu["sampler"].GibbsSampler(dom, init, fo, **kw) as b:
        first = b.run(40)
        blob = b.save().clone()
    with gpu["sampler"].GibbsSampler(dom, init + 1.0, fo, seed=999, burn_in=5, thin=3) as c:   # other start, other seed
        c.load(blob)
        assert c.count == 40
        rest = c.run(50)
        sc, stc = c.state(), c.stats()
    assert torch.equal(whole, torch.cat([first, rest], dim=0))
    for x, y in zip(sa, sc):
        assert torch.equal(x, y)
    assert torch.equal(sta["moments"], stc["moments"]) and int(sta["exps"]) == int(stc["exps"])
    # wider jumps, fewer acceptances: set_widths takes effect for the following runs only
    from olpefit_b200 import layout
    w0, _ = layout.default_widths(2)
    with gpu["sampler"].GibbsSampler(dom, init, fo, seed=4) as s:
        s.run(200, record=False)
        acc0 = float(s.stats()["accepts"].sum())
        s.set_widths(w0 * 8.0)
        s.run(200, record=False)
        acc1 = float(s.stats()["accepts"].sum()) - acc0
    assert acc1 < 0.8 * acc0
    # the checkpoint carries the jump widths: a chain tuned during burn-in continues with the tuned scale
    with gpu["sampler"].GibbsSampler(dom, init, fo, seed=4) as a:
        a.set_widths(w0 * 3.0)
        a.run(30, record=False)
        blob = a.save().clone()
        want = a.run(40)
    with gpu["sampler"].GibbsSampler(dom, init, fo, seed=4) as b:            # default widths until load()
        b.load(blob)
        assert torch.equal(b.run(40), want)
    with pytest.raises(Exception):
        s2 = gpu["sampler"].GibbsSampler(dom, init, fo, seed=4)
        try:
            s2.set_widths(-w0)
        finally:
            s2.close()


def test_error_paths_on_device(gpu):
    """Bad arguments are refused with a message, before anything is launched."""
    torch = gpu["torch"]
    from olpefit_b200 import _lib
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 32, n_frames=2)
    S = gpu["sampler"].GibbsSampler
    with pytest.raises(_lib.LapfError, match="frame_of"):
        S(dom, np.tile(p0, (3, 1)), np.array([0, 1, 2], dtype=np.int32))          # frame 2 does not exist
    with pytest.raises(_lib.LapfError, match="thin"):
        S(dom, p0[None], thin=0)
    with pytest.raises(_lib.LapfError, match="team_warps"):
        S(dom, p0[None], team_warps=3)
    with pytest.raises(_lib.LapfError, match="team_warps"):
        S(dom, p0[None], team_warps=16)                                            # 32-pixel stamps: 4 row steps
    img, _ = gpu["synth"].make_frame(0, 2, region=(480, 530, 470, 547))
    ragged = gpu["frame"].prepare_domain(img, HEADER, origin=(470, 480), nbody=2)
    with pytest.raises(_lib.LapfError, match="square stamps"):
        S(ragged, p0[None])                                                        # K1 handles it, the sampler does not
    with S(dom, np.tile(p0, (4, 1))) as s:
        small = torch.empty((2, 4, 17), dtype=torch.float64, device="cuda")
        with pytest.raises(ValueError):
            s.run(10, out=small)
        assert _lib.load().lapf_sampler_run(s._h, 10, small.data_ptr(), 2, None) == -1
        assert b"rows" in _lib.load().lapf_last_error()
        assert _lib.load().lapf_sampler_run(s._h, -1, None, 0, None) == -1
        assert s.count == 0                                                         # nothing ran
        with pytest.raises(_lib.LapfError, match="too small"):
            _lib.check(_lib.load().lapf_sampler_load(s._h, small.data_ptr(), 16, None))
    # out-of-range frame index in the stateless operator: nan for that vector only
    _, c = dom.model_chi2(np.stack([p0, p0, p0]), frame_of=np.array([0, 7, 1], dtype=np.int32))
    c = c.cpu().numpy()
    assert np.isfinite(c[0]) and np.isnan(c[1]) and np.isfinite(c[2])


# ---------------------------------------------------------------------------------------------
# factorised pixel loop, TMEM pixel store, batched sampler
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbody,size", [(2, 32), (2, 64), (3, 64), (2, 128), (3, 32)])
def test_factorised_loop_agrees_with_plain_loop(gpu, nbody, size):
    """The factorised pixel loop (one exponential per 4-pixel group and component, DESIGN.md
    section 4) against the plain loop (one per pixel and component) on the same vectors: both are
    within the stated 1e-5 of the float64 oracle, they agree with each other to FP32 rounding, and
    a sampler forced onto the plain loop follows the same chain."""
    synth, model = gpu["synth"], gpu["model"]
    lay = orc.layout_for(nbody)
    ox, oy = synth.stamp_origin(size)
    img, truth = synth.make_frame(4, nbody, region=(oy, oy + size, ox, ox + size))
    fast = _domain(gpu, img, (ox, oy), nbody)
    plain = model.PixelDomain(fast.data, fast.weight, fast.origin, nbody=nbody, plain_loop=True)
    tl = truth.copy()
    tl[0:2 * nbody:2] -= ox
    tl[1:2 * nbody:2] -= oy
    vecs = _random_vectors(tl, nbody, 32, np.random.default_rng(100 * size + nbody))
    vecs[:, 0:2 * nbody:2] += ox
    vecs[:, 1:2 * nbody:2] += oy
    m_f, c_f = (t.cpu().numpy() for t in fast.model_chi2(vecs, want_model=True))
    m_p, c_p = (t.cpu().numpy() for t in plain.model_chi2(vecs, want_model=True))
    assert np.max(np.abs(m_f - m_p) / np.abs(m_p)) < 5e-6
    np.testing.assert_allclose(c_f, c_p, rtol=2e-6)
    for i in (0, 7, 19):
        m = orc.model_image(vecs[i], lay, size, size, origin=(ox, oy))
        assert np.max(np.abs(m_f[i] - m) / np.abs(m)) < RTOL and np.max(np.abs(m_p[i] - m) / np.abs(m)) < RTOL
    # chains: same stream, same decisions until FP32 rounding flips a borderline accept
    init = np.tile(truth.astype(np.float32).astype(np.float64), (5, 1))
    out = []
    for dom in (fast, plain):
        with gpu["sampler"].GibbsSampler(dom, init, seed=21) as s:
            out.append(s.run(120).cpu().numpy())
    np.testing.assert_allclose(out[0][:40, :, :-1], out[1][:40, :, :-1], rtol=1e-9)
    np.testing.assert_allclose(out[0][:40, :, -1], out[1][:40, :, -1], rtol=2e-6)
    assert np.mean(np.isclose(out[0][..., :-1], out[1][..., :-1], rtol=1e-9)) > 0.9


def test_wild_vectors_leave_the_factorised_loop(gpu):
    """Outside its safe range (factors that would overflow or underflow on their own) the
    factorised loop hands the vector to the plain loop: needle-sharp cores, sources far outside
    the stamp, amplitudes enormous against the floor, elongated rotated profiles.  All of them still
    match the float64 oracle to the stated tolerance.  (Axis ratios beyond ~8 at 45 degrees do not, in
    either loop: the FP32 quadratic form cancels there, 2e-5 at 11:1 -- DESIGN.md section 5.)"""
    synth = gpu["synth"]
    lay = orc.layout_for(2)
    size = 64
    ox, oy = synth.stamp_origin(size)
    img, truth = synth.make_frame(6, 2, region=(oy, oy + size, ox, ox + size))
    dom = gpu["frame"].prepare_domain(img, HEADER, origin=(ox, oy), nbody=2, floor_index=9)
    t = truth.astype(np.float32).astype(np.float64)
    vecs = []
    v = t.copy(); v[10] = 0.35; v[11] = 0.4; vecs.append(v)                    # needle core (sigma 0.35 px)
    v = t.copy(); v[0] -= 300.0; vecs.append(v)                                 # star 300 px left of the stamp
    v = t.copy(); v[3] += 2000.0; vecs.append(v)                                # companion far above it
    v = t.copy(); v[9] = 1e-4; vecs.append(v)                                   # floor tiny against the amplitudes
    v = t.copy(); v[6] = 3e13; vecs.append(v)                                   # enormous amplitude
    v = t.copy(); v[10] = 0.6; v[11] = 1.8; v[14] = 0.78; vecs.append(v)        # sharp 3:1 core at 45 degrees
    v = t.copy(); v[12] = 9.0; v[13] = 2.2; v[15] = -0.7; vecs.append(v)        # elongated 4:1 wing
    v = t.copy(); v[10] = 0.05; vecs.append(v)                                  # sub-pixel delta
    vecs = np.array(vecs).astype(np.float32).astype(np.float64)
    model, chi2 = dom.model_chi2(vecs, want_model=True)
    model, chi2 = model.cpu().numpy(), chi2.cpu().numpy()
    img64 = img.astype(np.float64)
    w = orc.weight_map(img64, HEADER)
    for i, q in enumerate(vecs):
        m = orc.model_image(q, lay, size, size, origin=(ox, oy), floor_index=9)
        rel = np.max(np.abs(model[i] - m) / np.abs(m))
        assert rel < RTOL, "vector %d: %.2e" % (i, rel)
        assert chi2[i] == pytest.approx(orc.chi_squared_weighted(img64, m, w), rel=RTOL)
    # and a chain started from a wild point is the chain the plain loop gives, bit for bit
    plain = gpu["model"].PixelDomain(dom.data, dom.weight, dom.origin, nbody=2, floor_index=9, plain_loop=True)
    res = []
    for d in (dom, plain):
        with gpu["sampler"].GibbsSampler(d, vecs[:1], seed=2) as s:
            res.append(s.run(30))
    assert gpu["torch"].equal(res[0], res[1])


@pytest.mark.parametrize("nbody", [2, 3])
def test_two_vectors_per_pass_are_independent_of_their_partner(gpu, nbody):
    """32-pixel stamps are evaluated two vectors per pass, one per half warp.
    Chi-square and the model image must stay pure functions of the vector: alone (a batch of one),
    in an odd batch, paired with a tame or with a wild partner (which takes the plain loop, so both
    loops run in that pass), in either half of the warp -- always the same bits, and always the
    float64 oracle to the stated tolerance."""
    synth = gpu["synth"]
    lay = orc.layout_for(nbody)
    size = 32
    ox, oy = synth.stamp_origin(size)
    img, truth = synth.make_frame(8, nbody, region=(oy, oy + size, ox, ox + size))
    dom = _domain(gpu, img, (ox, oy), nbody)
    tl = truth.copy()
    tl[0:2 * nbody:2] -= ox
    tl[1:2 * nbody:2] -= oy
    tame = _random_vectors(tl, nbody, 9, np.random.default_rng(5 + nbody))
    tame[:, 0:2 * nbody:2] += ox
    tame[:, 1:2 * nbody:2] += oy
    wild = tame[:4].copy()
    sx, sy, sx2, _ = lay.i_sigma
    wild[0, sx] = 0.3; wild[0, sy] = 0.35                        # needle core: outside the safe range
    wild[1, 0] -= 400.0                                          # star far outside the stamp
    wild[2, lay.i_amp(0)] = 3e13                                 # enormous amplitude
    wild[3, sx2] = float("nan")                                  # nan shape
    vecs = np.concatenate([tame, wild]).astype(np.float32).astype(np.float64)   # 13 vectors: odd
    n = len(vecs)
    model, chi2 = (t.cpu().numpy() for t in dom.model_chi2(vecs, want_model=True))
    img64 = img.astype(np.float64)
    w = orc.weight_map(img64, HEADER)
    for i in range(n - 1):                                       # the last one is nan
        m = orc.model_image(vecs[i], lay, size, size, origin=(ox, oy))
        assert np.max(np.abs(model[i] - m) / np.abs(m)) < RTOL, i
        assert chi2[i] == pytest.approx(orc.chi_squared_weighted(img64, m, w), rel=RTOL)
    assert np.isnan(chi2[-1])
    # alone
    for i in range(n):
        m1, c1 = (t.cpu().numpy() for t in dom.model_chi2(vecs[i:i + 1], want_model=True))
        assert np.array_equal(c1, chi2[i:i + 1], equal_nan=True), i
        assert np.array_equal(m1[0], model[i], equal_nan=True), i
    # every ordered pair (a, b): a in the lower half warp, b in the upper
    ia, ib = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    pairs = np.stack([vecs[ia.ravel()], vecs[ib.ravel()]], axis=1).reshape(-1, vecs.shape[1])
    _, cp = dom.model_chi2(pairs)
    cp = cp.cpu().numpy().reshape(n, n, 2)
    assert np.array_equal(cp[..., 0], np.broadcast_to(chi2[:, None], (n, n)), equal_nan=True)
    assert np.array_equal(cp[..., 1], np.broadcast_to(chi2[None, :], (n, n)), equal_nan=True)
    # and the sampler (walkers in both halves, an odd number of them, tame and wild starting points):
    # the state's chi-square is the stateless operator's on the state, bit for bit
    init = vecs[:n - 1][[0, 9, 1, 10, 2, 11, 3]]                 # 7 walkers; 9-11 are wild
    with gpu["sampler"].GibbsSampler(dom, init, seed=17) as s:
        s.run(64)
        st = s.state()[0].cpu().numpy()
    _, c = dom.model_chi2(np.ascontiguousarray(st[:, :-1]))
    assert np.array_equal(c.cpu().numpy(), st[:, -1])


@pytest.mark.parametrize("size", [32, 64])
def test_batched_sampler_partitions(gpu, size):
    """The batched kernel deals walkers to CTAs, warps and lanes: ragged frames, frames without
    walkers, fewer walkers than SMs, more walkers of one frame than a CTA holds at once, items
    that run across a frame boundary with both stamps resident (two pixel-store slots; on 32-pixel
    stamps also two walkers per pass).  Whatever the shape, a walker's chain is the chain it has
    when it runs alone under the same id."""
    torch = gpu["torch"]
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, size, n_frames=7)
    rng = np.random.default_rng(8)

    def solo(frame, gid, start, n_upd, seed):
        with gpu["sampler"].GibbsSampler(dom, start[None], np.array([frame], dtype=np.int32), seed=seed,
                                         id_base=gid) as s:
            return s.run(n_upd)[:, 0]

    # ragged: 7 frames with very different walker counts (one of them empty), distinct starting points
    counts = [1, 40, 3, 0, 600, 2, 33]
    frame_of = np.repeat(np.arange(7), counts).astype(np.int32)
    rng.shuffle(frame_of)
    W = frame_of.size
    init = np.tile(p0, (W, 1))
    init[:, 0:4] += rng.normal(0, 0.05, (W, 4))
    init = init.astype(np.float32).astype(np.float64)
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=12, id_base=1000) as s:
        chain = s.run(24)
        st, tries, acc = s.state()
    assert bool((tries.sum(dim=1) == 24).all())
    for w in (0, 1, 17, 333, W - 1, int(np.flatnonzero(frame_of == 0)[0]), int(np.flatnonzero(frame_of == 5)[1])):
        assert torch.equal(chain[:, w], solo(int(frame_of[w]), 1000 + w, init[w], 24, 12)), "walker %d" % w

    # more walkers of one frame than a CTA holds at once (and a few of another frame)
    W = 90000
    frame_of = np.zeros(W, dtype=np.int32)
    frame_of[-37:] = 2
    init = np.tile(p0, (W, 1))
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=5) as s:
        chain = s.run(3)
        st, tries, acc = s.state()
    assert bool((tries.sum(dim=1) == 3).all())
    for w in (0, 511, 512, 607, 608, 44444, W - 38, W - 1):
        assert torch.equal(chain[:, w], solo(int(frame_of[w]), w, init[w], 3, 5)), "walker %d" % w


def test_stream_of_epochs_reuses_domain_and_sampler(gpu):
    """A service processing batch after batch (bench.py's e2e): new frames and origins go into the
    SAME device buffers (prepare_domain(into=...)), the sampler is reset, chain rows leave through
    the double-buffered ChainStreamer.  Every batch must equal a fresh domain + fresh sampler."""
    torch, synth, frame, sampler = gpu["torch"], gpu["synth"], gpu["frame"], gpu["sampler"]
    size, nf, walkers, n_upd = 64, 3, 40, 48
    frame_of = (np.arange(walkers) % nf).astype(np.int32)
    batches = []
    for b in range(3):
        stamps, origins = synth.make_stamps(nf, size, 2)
        stamps = stamps[::-1].copy() if b == 1 else stamps + np.float32(b)        # different pixels ...
        origins = origins[::-1].copy() if b == 1 else origins + b                # ... and different origins
        p = []
        for f in range(nf):
            g = synth.step1_guess(stamps[f], 2, origin=tuple(origins[f]))
            p.append(frame.initial_parameters(stamps[f], g, 2, origin=tuple(origins[f])))
        batches.append((stamps, origins, np.asarray(p)[frame_of]))

    fresh = []
    for b, (stamps, origins, init) in enumerate(batches):
        dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=2)
        with sampler.GibbsSampler(dom, init, frame_of, seed=50 + b, thin=4) as s:
            fresh.append((s.run(n_upd).cpu().numpy(), [t.cpu().numpy() for t in s.state()]))

    stamps, origins, init = batches[0]
    dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=2)
    got = []
    with sampler.GibbsSampler(dom, init, frame_of, seed=50, thin=4) as s:
        streamer = sampler.ChainStreamer(s, n_upd)
        for b, (stamps, origins, init) in enumerate(batches):
            if b > 0:
                frame.prepare_domain(torch.from_numpy(stamps).to(dom.device), HEADER, origin=origins, nbody=2, into=dom)
                s.reset(init, seed=50 + b)
            prev = streamer.run(n_upd)
            if prev is not None:
                got.append(prev.copy())
            if b == len(batches) - 1:
                state = [t.cpu().numpy() for t in s.state()]
        got.append(streamer.finish().copy())
    assert len(got) == 3
    for b in range(3):
        assert np.array_equal(got[b], fresh[b][0]), "batch %d" % b
    for x, y in zip(state, fresh[2][1]):
        assert np.array_equal(x, y)


# ---------------------------------------------------------------------------------------------
# round 2: statistics without the chain, compact chain rows, outside sums
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbody,team", [(2, 1), (3, 1), (2, 4)])
def test_device_sketches_match_the_recorded_chain(gpu, nbody, team):
    """lapf_sampler_sketch: separation / position-angle histograms and sums accumulated in the
    record step equal what numpy gets from the chain rows the same run recorded
    (apf_step3.py:255-256,283-291,436-437) -- counts exactly, quantiles to one bin width."""
    torch = gpu["torch"]
    from olpefit_b200 import stats
    dom, stamps, origins, p0 = _sampler_setup(gpu, nbody, 32, n_frames=3)
    walkers = 50
    frame_of = (np.arange(walkers) % 3).astype(np.int32)
    init = np.array([gpu["synth"].truth_parameters(nbody, f) for f in range(3)])[frame_of]   # chains stay near the truth
    n_bins, sep_bin, pa_bin = 4096, 1e-3, 5e-3
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=8, burn_in=20, thin=3, team_warps=team) as s:
        s.enable_sketch(n_bins=n_bins, sep_bin=sep_bin, pa_bin=pa_bin)
        chain = torch.cat([s.run(100), s.run(140)], dim=0).cpu().numpy()
        sk = s.sketch()
    hist, summ = sk["hist"].cpu().numpy(), sk["summary"].cpu().numpy()
    res = {k: v.cpu().numpy() for k, v in stats.sketch_summary(sk, pixscale=1.0).items()}
    rows = chain.shape[0]
    for f in range(3):
        sel = frame_of == f
        for o in range(1, nbody):
            dx = chain[:, sel, 2 * o] - chain[:, sel, 0]
            dy = chain[:, sel, 2 * o + 1] - chain[:, sel, 1]
            for q, vals in enumerate((np.sqrt(dx * dx + dy * dy), np.degrees(np.arctan2(-dx, dy)))):
                width = (sep_bin, pa_bin)[q]
                c = summ[f, o - 1, q, 0]
                assert summ[f, o - 1, q, 3] == rows * sel.sum() == hist[f, o - 1, q].sum()
                raw = (vals - c) / width
                pos = np.floor(raw) + n_bins // 2
                b = np.where(pos < 0, 0, np.where(pos >= n_bins, n_bins + 1, pos + 1)).astype(int)
                ref_hist = np.bincount(b.ravel(), minlength=n_bins + 2)
                # a value within rounding of a bin edge may fall on either side (walkers that have not moved
                # yet sit exactly on the edge at the centre: device and numpy sqrt / atan2 differ in the last bit)
                on_edge = int((np.abs(raw - np.rint(raw)) < 1e-6).sum())
                assert np.abs(hist[f, o - 1, q] - ref_hist).sum() <= 2 * on_edge
                assert summ[f, o - 1, q, 1] == pytest.approx((vals - c).sum(), rel=1e-9, abs=1e-9)
                assert summ[f, o - 1, q, 2] == pytest.approx(((vals - c) ** 2).sum(), rel=1e-9)
                # a histogram quantile is exact to its bin: the share of the values below (quantile - bin)
                # is at most the level, the share below (quantile + bin) at least the level
                qs = res["sep_q" if q == 0 else "pa_q"][f, o - 1]
                for level, v in zip((0.15865, 0.5, 0.84135), qs):
                    assert np.mean(vals < v - width) <= level <= np.mean(vals <= v + width), (f, o, q, level, v)
                assert res["sep_std" if q == 0 else "pa_std"][f, o - 1] == pytest.approx(vals.std(), rel=1e-6)
    # a quantile that falls outside the histogram comes back as nan, never as a made-up number
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=8, thin=5) as s:
        s.enable_sketch(n_bins=2, sep_bin=1e-6, pa_bin=1e-6)
        s.run(50, record=False)
        tiny = stats.sketch_summary(s.sketch())
        assert bool(torch.isnan(tiny["sep_q"]).any()) and float(tiny["outside"].max()) > 0.5
    # reset zeroes the sketches
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, seed=8, thin=5) as s:
        s.enable_sketch(n_bins=64, sep_bin=0.01, pa_bin=0.05)
        s.run(50, record=False)
        assert int(s.sketch()["hist"].sum()) == 10 * walkers * 2 * (nbody - 1)
        s.reset(init, seed=9)
        assert int(s.sketch()["hist"].sum()) == 0


def test_float32_difference_rows_equal_the_float64_rows(gpu):
    """LAPF_CHAIN_F32_DELTA: the same run, rows leaving as float32 differences from the starting
    point: exactly float32(row - start), through run() and through the ChainStreamer."""
    torch = gpu["torch"]
    dom, stamps, origins, p0 = _sampler_setup(gpu, 2, 32, n_frames=2)
    walkers = 600
    frame_of = (np.arange(walkers) % 2).astype(np.int32)
    init = np.tile(p0, (walkers, 1))
    init[:, 0:4] += np.random.default_rng(1).normal(0, 0.02, (walkers, 4))
    kw = dict(seed=17, burn_in=0, thin=2)
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, **kw) as a:
        full = a.run(40)
        start = a.start()
    assert torch.equal(start[:, :16], torch.as_tensor(init, device=start.device))
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, **kw) as b:
        b.set_chain_format("f32delta")
        small = b.run(40)
        assert small.dtype == torch.float32 and small.shape == full.shape
        assert torch.equal(small, (full - start[None]).float())
        streamer = gpu["sampler"].ChainStreamer(b, 20)
        assert streamer.run(20) is None
        seg = streamer.finish()
        assert seg.dtype == np.float32 and seg.shape == (10, walkers, 17)
    with gpu["sampler"].GibbsSampler(dom, init, frame_of, **kw) as c:
        more = c.run(60)[20:]
    assert np.array_equal(seg, (more - start[None]).float().cpu().numpy())
    # positions survive to ~1e-7 of the distance travelled although they leave in 32 bits (as plain
    # floats 512.3 would keep only 3e-5)
    back = start[None] + small.double()
    moved = (full[..., :4] - start[None, :, :4]).abs()
    assert float(((back[..., :4] - full[..., :4]).abs() - 6e-8 * moved).max()) <= 0.0


def test_outside_sums_kernel_matches_float64_numpy(gpu):
    """lapf_frame_outside (lapf_problem.outside): sum w, sum w d, sum w d^2 over the pixels outside
    the cut-out with the device weight map, against numpy in float64."""
    synth, frame = gpu["synth"], gpu["frame"]
    imgs = np.stack([synth.make_frame(f, 2)[0] for f in range(2)])
    imgs[1, 200:210, 300:310] = 30000.0                       # saturated pixels far from the objects: masked
    imgs[0, 5, 7] = np.nan                                    # a dead pixel: no weight
    cuts = np.array([[450, 448], [460, 470]], dtype=np.int32)
    dom = frame.prepare_domain(imgs, HEADER, size=128, cut=cuts, whole_frame=True)
    got = dom.outside.cpu().numpy()
    sat, rn = frame.saturation_level(HEADER), frame.read_noise(HEADER)
    for f in range(2):
        v = imgs[f].astype(np.float64)
        ok = np.isfinite(v) & ~(v > 0.8 * sat)
        w = np.where(ok, (1.0 / (rn * rn + np.abs(np.where(ok, v, 0.0)))).astype(np.float32).astype(np.float64), 0.0)
        d = np.where(ok, v, 0.0)
        inside = np.zeros_like(ok)
        inside[cuts[f, 1]:cuts[f, 1] + 128, cuts[f, 0]:cuts[f, 0] + 128] = True
        w = np.where(inside, 0.0, w)
        np.testing.assert_allclose(got[f], [w.sum(), (w * d).sum(), (w * d * d).sum()], rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# round 2: the pointwise parity test, wide (>= 1000 vectors per shape, masked cores,
# negative pixels, both floors), with the worst errors written down as an artefact
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbody", [2, 3])
@pytest.mark.parametrize("size", [32, 64, 128])
def test_k1_pointwise_parity_1000_vectors(gpu, nbody, size):
    """1,000 seeded parameter vectors per shape in four frame variants -- the synthetic epoch as it
    is, its star core above 0.8 x satlevel (masked pixels, apf_step2.py:188), a sky-subtracted copy
    with negative pixels (err^2 = readnoise^2 + |image|, :207-210), and the --fix-bkgd floor -- against
    the float64 oracle: per-pixel model within 1e-5 relative, chi-square within 1e-5 relative.  The
    worst errors go to gpurun_out/r02_parity_errors.json (copied to profiles/)."""
    import json
    synth, model = gpu["synth"], gpu["model"]
    lay = orc.layout_for(nbody)
    ox, oy = synth.stamp_origin(size)
    base, truth = synth.make_frame(5, nbody, region=(oy, oy + size, ox, ox + size))
    tl = truth.copy()
    tl[0:2 * nbody:2] -= ox
    tl[1:2 * nbody:2] -= oy
    rng = np.random.default_rng(1000 * size + nbody)
    vecs = _random_vectors(tl, nbody, 1000, rng)
    vecs[:, 0:2 * nbody:2] += ox
    vecs[:, 1:2 * nbody:2] += oy
    grid = orc.pixel_grid(size, size, (ox, oy))
    variants = {"plain": (base, None), "masked_core": (base * np.float32(1.6), None),
                "negative_pixels": (base - np.float32(120.0), None), "fix_bkgd": (base, 3 * nbody + 3)}
    report = {}
    for name, (img, floor_index) in variants.items():
        dom = gpu["frame"].prepare_domain(img, HEADER, origin=(ox, oy), nbody=nbody, floor_index=floor_index)
        img64 = img.astype(np.float64)
        w = orc.weight_map(img64, HEADER)
        if name == "masked_core":
            assert (w == 0).sum() > 4 and np.array_equal(dom.weight[0].cpu().numpy() == 0, w == 0)
        if name == "negative_pixels":
            assert (img64 < 0).sum() > size
        worst_px, worst_chi = 0.0, 0.0
        for lo in range(0, 1000, 250):                     # model images of 250 vectors at a time
            mod, chi = dom.model_chi2(vecs[lo:lo + 250], want_model=True)
            mod, chi = mod.cpu().numpy(), chi.cpu().numpy()
            for i in range(250):
                m = orc.model_image(vecs[lo + i], lay, size, size, grid=grid, floor_index=floor_index)
                worst_px = max(worst_px, float(np.max(np.abs(mod[i] - m) / np.abs(m))))
                c = orc.chi_squared_weighted(img64, m, w)
                worst_chi = max(worst_chi, abs(chi[i] - c) / abs(c))
        report[name] = {"worst_pixel_rel": worst_px, "worst_chi2_rel": worst_chi}
        assert worst_px < RTOL and worst_chi < RTOL, (name, worst_px, worst_chi)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_parity_errors.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    try:
        with open(path) as fh:
            allrep = json.load(fh)
    except Exception:
        allrep = {"tolerance": RTOL, "vectors_per_shape_and_variant": 1000, "oracle": "float64 numpy, oracle/lapf_oracle.py"}
    allrep["%d-body %dx%d" % (nbody, size, size)] = report
    with open(path, "w") as fh:
        json.dump(allrep, fh, indent=1, sort_keys=True)
    print("worst errors", nbody, size, report)


def test_selftest_compares_trial_chi2_with_k1_and_restores_the_sampler(gpu, monkeypatch):
    """lapf_sampler_selftest (run by lapf_sampler_create): three recorded probe updates whose TRIAL
    chi-squares must equal the stateless operator bit for bit -- rejected proposals included, which
    no state-based check sees -- and a sampler that continues exactly as if nothing had happened."""
    torch = gpu["torch"]
    from olpefit_b200 import _lib
    lib = _lib.load()
    for nbody, size, walkers in ((2, 32, 700), (3, 64, 600), (2, 128, 250)):
        dom, stamps, origins, p0 = _sampler_setup(gpu, nbody, size, n_frames=2)
        frame_of = (np.arange(walkers) % 2).astype(np.int32)
        init = np.tile(p0, (walkers, 1))
        kw = dict(seed=3, burn_in=4, thin=3)
        with gpu["sampler"].GibbsSampler(dom, init, frame_of, **kw) as a:          # self-test inside create()
            first = a.run(30)
            sa = a.state()
            _lib.check(lib.lapf_sampler_selftest(a._h, None))                          # and in the middle of a run
            assert a.count == 30
            for x, y in zip(sa, a.state()):
                assert torch.equal(x, y)
            rest = a.run(30)
            stats_a = a.stats()
        monkeypatch.setenv("LAPF_NO_SELFTEST", "1")
        with gpu["sampler"].GibbsSampler(dom, init, frame_of, **kw) as b:
            whole = b.run(60)
            stats_b = b.stats()
        monkeypatch.delenv("LAPF_NO_SELFTEST")
        assert torch.equal(torch.cat([first, rest], dim=0), whole)
        assert torch.equal(stats_a["moments"], stats_b["moments"]) and int(stats_a["exps"]) == int(stats_b["exps"])


def test_tma_cut_out_equals_the_plain_cut_out(gpu, monkeypatch):
    """lapf_frame_prep fetches the cut-outs through a 3-D TMA tensor map over the frames when their
    rows are 16-byte aligned; the plain kernel (LAPF_NO_TMA_PREP, or frames a tensor map cannot
    describe) does the same arithmetic: identical bits, also for cut-outs that stick out of the frame,
    masked and non-finite pixels, and sizes that are not a multiple of the box height."""
    torch, frame = gpu["torch"], gpu["frame"]
    rng = np.random.default_rng(4)
    for fy, fx, size, cuts in ((200, 256, (64, 64), [(10, 20), (-7, 150), (230, -3)]),
                               (96, 132, (40, 36), [(0, 0), (100, 60), (5, 7)]),
                               (64, 130, (32, 32), [(3, 4), (90, 30), (0, 0)])):          # fx % 4 != 0: plain kernel both times
        imgs = rng.normal(50.0, 40.0, (3, fy, fx)).astype(np.float32)
        imgs[0, 30:33, 40:44] = 30000.0
        imgs[1, 12, 13] = np.nan
        imgs[2, 50, 60] = np.inf
        cut = np.array(cuts, dtype=np.int32)
        a = frame.prepare_domain(imgs, HEADER, size=size, cut=cut)
        monkeypatch.setenv("LAPF_NO_TMA_PREP", "1")
        b = frame.prepare_domain(imgs, HEADER, size=size, cut=cut)
        monkeypatch.delenv("LAPF_NO_TMA_PREP")
        assert torch.equal(a.data, b.data) and torch.equal(a.weight, b.weight)
        # and against numpy for one frame
        sat, rn = frame.saturation_level(HEADER), frame.read_noise(HEADER)
        ny, nx = size
        for f in range(3):
            x0, y0 = cuts[f]
            ref_d = np.zeros((ny, nx), np.float32)
            ref_w = np.zeros((ny, nx), np.float32)
            for r in range(ny):
                for c in range(nx):
                    y, x = y0 + r, x0 + c
                    if 0 <= y < fy and 0 <= x < fx:
                        v = imgs[f, y, x]
                        if np.isfinite(v) and not (float(v) > 0.8 * sat):
                            ref_d[r, c] = v
                            ref_w[r, c] = np.float32(1.0 / (rn * rn + abs(float(v))))
            assert np.array_equal(a.data[f].cpu().numpy(), ref_d)
            np.testing.assert_allclose(a.weight[f].cpu().numpy(), ref_w, rtol=1e-7)
