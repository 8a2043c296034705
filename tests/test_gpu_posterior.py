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


@pytest.mark.parametrize("fixture", ["posterior_2body_s32.npz", "posterior_3body_s32.npz"])
def test_separation_and_position_angle_posterior_match_cpu_reference(golden_dir, fixture):
    import torch
    from olpefit_b200 import chains, frame, sampler, synth

    z = np.load(os.path.join(golden_dir, fixture))
    size, nbody, epoch = int(z["size"]), int(z["nbody"]), int(z["epoch"])
    ox, oy = (int(v) for v in z["origin"])
    img, _ = synth.make_frame(epoch, nbody, region=(oy, oy + size, ox, ox + size))
    dom = frame.prepare_domain(img, HEADER, origin=(ox, oy), nbody=nbody)
    walkers = 2048
    with sampler.GibbsSampler(dom, np.tile(z["p0"], (walkers, 1)), seed=20260101,
                              burn_in=int(z["burn"]), thin=int(z["thin"])) as s:
        chain = s.run(int(z["updates"]))
        st = s.stats()
        acc = (st["accepts"].double() / st["tries"].double()).cpu().numpy()
    rows = chain.cpu().numpy()                                    # [rows, walkers, P+1]
    n_cpu = int(z["walkers"])
    assert rows.shape[0] == (int(z["updates"]) - int(z["burn"])) // int(z["thin"]) + 1

    sep, pa = chains.separation_pa(rows[..., 0], rows[..., 1], rows[..., 2], rows[..., 3])
    for name, vals, q_cpu, q_walker in (("sep", sep, z["sep_q"], z["sep_q_walker"]),
                                        ("pa", pa, z["pa_q"], z["pa_q_walker"])):
        q_gpu = np.percentile(vals, QS)
        # Monte-Carlo standard error of each pooled CPU quantile from the spread between walkers;
        # the same for the GPU walkers (32x more of them)
        se_cpu = q_walker.std(axis=0, ddof=1) / np.sqrt(n_cpu)
        se_gpu = np.percentile(vals, QS, axis=0).std(axis=1, ddof=1) / np.sqrt(walkers)
        zscore = (q_gpu - q_cpu) / np.sqrt(se_cpu ** 2 + se_gpu ** 2)
        print(name, "gpu", q_gpu, "cpu", q_cpu, "z", zscore)
        assert np.all(np.abs(zscore) < 5.0), (name, q_gpu, q_cpu, zscore)
        # and the 68% interval has the same width to a few per cent
        assert (q_gpu[2] - q_gpu[0]) == pytest.approx(q_cpu[2] - q_cpu[0], rel=0.05)

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
