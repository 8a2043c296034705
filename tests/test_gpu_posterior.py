"""Statistical parity (BASELINE.json: posterior medians and 68% intervals of separation and
position angle must match a reference CPU run within Monte-Carlo error on the same frame).

The CPU side is tests/golden/posterior_2body_s32.npz, produced by tools/make_posterior_fixture.py:
64 (2-body) / 48 (3-body) walkers of the float64 numpy oracle with numpy's Mersenne-Twister stream (the reference's
stream), 50,000 updates each, first 10,000 dropped, every 10th kept.  The GPU side runs the same
configuration with 2048 walkers on the Philox stream.  Bit-equal chains are impossible by design
(the reference is unseeded); the comparison is distributional, with standard errors taken from
the spread between the CPU walkers."""
import os

import numpy as np
import pytest

from oracle import lapf_oracle as orc

pytestmark = pytest.mark.gpu
HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
QS = [15.865, 50.0, 84.135]


def _device_chain(z, plain_loop=False, walkers=2048, seed=20260101):
    """The fixture's configuration on the device: same frame, same stamp, same starting point, same
    burn-in / thinning; 2048 walkers on the Philox stream.  Returns (rows [rows, walkers, P+1], acceptance)."""
    import torch
    from olpefit_b200 import frame, model, sampler, synth
    size, nbody, epoch = int(z["size"]), int(z["nbody"]), int(z["epoch"])
    ox, oy = (int(v) for v in z["origin"])
    domain = str(z["domain"]) if "domain" in z.files else "stamp"
    if domain == "frame":
        # the reference's pixel domain: the whole 1024 x 1024 frame (apf_step2.py:94,134-137)
        img, _ = synth.make_frame(epoch, nbody)
        dom = frame.prepare_domain(img, HEADER, size=size, cut=(ox, oy), nbody=nbody, whole_frame=True)
    else:
        img, _ = synth.make_frame(epoch, nbody, region=(oy, oy + size, ox, ox + size))
        dom = frame.prepare_domain(img, HEADER, origin=(ox, oy), nbody=nbody)
    if plain_loop:
        dom = model.PixelDomain(dom.data, dom.weight, dom.origin, nbody=nbody, outside=dom.outside, plain_loop=True)
    with sampler.GibbsSampler(dom, np.tile(z["p0"], (walkers, 1)), seed=seed,
                              burn_in=int(z["burn"]), thin=int(z["thin"])) as s:
        chain = s.run(int(z["updates"]))
        st = s.stats()
        acc = (st["accepts"].double() / st["tries"].double()).cpu().numpy()
    rows = chain.cpu().numpy()
    assert rows.shape[0] == (int(z["updates"]) - int(z["burn"])) // int(z["thin"]) + 1
    return rows, acc


def _compare_quantiles(name, vals, q_cpu, q_walker_cpu, n_cpu):
    """z-test of the pooled 16/50/84 % quantiles: standard errors from the spread between walkers."""
    n_gpu = vals.shape[1]
    q_gpu = np.percentile(vals, QS)
    se_cpu = q_walker_cpu.std(axis=0, ddof=1) / np.sqrt(n_cpu)
    se_gpu = np.percentile(vals, QS, axis=0).std(axis=1, ddof=1) / np.sqrt(n_gpu)
    zscore = (q_gpu - q_cpu) / np.sqrt(se_cpu ** 2 + se_gpu ** 2)
    print(name, "gpu", q_gpu, "cpu", q_cpu, "z", zscore)
    assert np.all(np.abs(zscore) < 5.0), (name, q_gpu, q_cpu, zscore)
    # and the 68% interval has the same width to a few per cent
    assert (q_gpu[2] - q_gpu[0]) == pytest.approx(q_cpu[2] - q_cpu[0], rel=0.05)


# the benchmark shapes (64: factorised loop + far-field culling + both planes in TMEM; 128 on the
# reference's whole-frame domain: weight plane in TMEM) next to the 32-pixel ones, and the plain loop
@pytest.mark.parametrize("fixture,plain_loop", [("posterior_2body_s32.npz", False), ("posterior_3body_s32.npz", False),
                                                ("posterior_2body_s64.npz", False), ("posterior_2body_s64.npz", True),
                                                ("posterior_2body_s128_frame.npz", False)])
def test_separation_and_position_angle_posterior_match_cpu_reference(golden_dir, fixture, plain_loop):
    from olpefit_b200 import chains
    z = np.load(os.path.join(golden_dir, fixture))
    rows, acc = _device_chain(z, plain_loop)
    n_cpu, walkers = int(z["walkers"]), rows.shape[1]

    sep, pa = chains.separation_pa(rows[..., 0], rows[..., 1], rows[..., 2], rows[..., 3])
    _compare_quantiles("sep", sep, z["sep_q"], z["sep_q_walker"], n_cpu)
    _compare_quantiles("pa", pa, z["pa_q"], z["pa_q_walker"], n_cpu)

    # every sampled parameter (not only the astrometry): pooled mean within Monte-Carlo error
    flat = rows.reshape(-1, rows.shape[-1])
    mean_gpu = flat.mean(axis=0)
    se_cpu = z["param_mean_walker"].std(axis=0, ddof=1) / np.sqrt(n_cpu)
    se_gpu = rows.mean(axis=0).std(axis=0, ddof=1) / np.sqrt(walkers)
    zscore = (mean_gpu - z["param_mean"]) / np.sqrt(se_cpu ** 2 + se_gpu ** 2)
    print("parameter z-scores", np.round(zscore, 2))
    assert np.all(np.abs(zscore) < 5.0), zscore
    # acceptance rates with the reference jump widths agree too
    np.testing.assert_allclose(acc, z["acceptance"], atol=0.02)
    # truth is inside the posterior bulk (sanity of the whole chain: synthetic truth is in-model)
    t_sep, t_pa = orc.separation_pa(*z["truth"][:4])
    assert np.percentile(sep, 0.5) < t_sep < np.percentile(sep, 99.5)
    assert np.percentile(pa, 0.5) < t_pa < np.percentile(pa, 99.5)


def test_posterior_from_the_step1_guess_matches_cpu_reference(golden_dir):
    """Started from the raw step-1 guess (apf_step2.py:258-273) instead of the truth, at the
    benchmark's 64-pixel stamp.  From there the reference's sampler itself is bimodal: a few per cent
    of its walkers collapse the companion onto the star.  Both sides must show the same behaviour:
    the same share of walkers on the star-companion solution, and there the same posterior."""
    from olpefit_b200 import chains
    z = np.load(os.path.join(golden_dir, "posterior_2body_s64_guess.npz"))
    rows, acc = _device_chain(z)
    n_cpu, walkers = int(z["walkers"]), rows.shape[1]
    sep, pa = chains.separation_pa(rows[..., 0], rows[..., 1], rows[..., 2], rows[..., 3])
    t_sep, t_pa = orc.separation_pa(*z["truth"][:4])
    main = (np.abs(np.median(sep, axis=0) - t_sep) < 3.0) & (np.abs(np.median(pa, axis=0) - t_pa) < 2.0)
    f_gpu, f_cpu = main.mean(), z["main_mode"].mean()
    se = np.sqrt(max(f_cpu * (1 - f_cpu), 1.0 / n_cpu) / n_cpu + f_gpu * (1 - f_gpu) / walkers)
    print("share of walkers on the main mode: gpu %.3f cpu %.3f (se %.3f)" % (f_gpu, f_cpu, se))
    assert abs(f_gpu - f_cpu) < 4.0 * se
    cm = z["main_mode"]
    _compare_quantiles("sep", sep[:, main], z["sep_q_main"], z["sep_q_walker"][cm], int(cm.sum()))
    _compare_quantiles("pa", pa[:, main], z["pa_q_main"], z["pa_q_walker"][cm], int(cm.sum()))


def test_device_side_statistics_match_numpy(golden_dir):
    """Step-3 statistics computed on the device (olpefit_b200/stats.py) equal the numpy / oracle
    versions, per frame, on a multi-epoch batch; GR from the K4 moments equals GR from the rows."""
    import torch
    from olpefit_b200 import chains, frame, sampler, stats, synth
    nf, walkers, size = 3, 96, 32
    stamps, origins = synth.make_stamps(nf, size)
    dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=2)
    frame_of = (np.arange(walkers) % nf).astype(np.int32)
    p0 = np.array([synth.truth_parameters(2, f) for f in range(nf)])[frame_of]
    with sampler.GibbsSampler(dom, p0, frame_of, seed=6, burn_in=2000, thin=5) as s:
        chain = s.run(6000)
        st = s.stats()
    summ = stats.per_frame_summary(chain, frame_of, nf)
    rows = chain.cpu().numpy()
    for f in range(nf):
        sub = rows[:, frame_of == f, :]
        sep, pa = orc.separation_pa(sub[..., 0], sub[..., 1], sub[..., 2], sub[..., 3])
        np.testing.assert_allclose(summ["sep_q"][f].cpu().numpy(), np.percentile(sep, QS), rtol=1e-10)
        np.testing.assert_allclose(summ["pa_q"][f].cpu().numpy(), np.percentile(pa, QS), rtol=1e-10)
        assert float(summ["sep_std"][f]) == pytest.approx(sep.std(), rel=1e-9)
        # the companion drifts by (0.01, -0.005) px per epoch in the synthetic frames: sep follows the truth
        t_sep, t_pa = orc.separation_pa(*synth.truth_parameters(2, f)[:4])
        assert abs(float(summ["sep_q"][f, 1]) - t_sep) < 4.0 and abs(float(summ["pa_q"][f, 1]) - t_pa) < 2.0
        psrf_t, rc_t = stats.gelman_rubin(chain[:, torch.as_tensor(frame_of == f, device=chain.device), :16])
        _, rc_m = chains.gelman_rubin_from_moments(st["moments"][f, :16].cpu().numpy(), rows.shape[0], sub.shape[1])
        for j in range(16):
            assert float(rc_t[j]) == pytest.approx(orc.gelman_rubin(sub[:, :, j])[1], rel=1e-8)
        np.testing.assert_allclose(rc_m, rc_t.cpu().numpy(), rtol=1e-6)
    assert summ["walkers"].tolist() == [32, 32, 32]
