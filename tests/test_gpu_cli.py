"""End-to-end drop-in test of the three command lines on a synthetic NIRC2-like FITS frame:
same arguments, same input file, same output files as the reference scripts, and the output
loads through the step-3 ingest (apf_step3.py:169-214)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_case(tmp_path, nbody, tag=""):
    from olpefit_b200 import frame, synth
    d = tmp_path / ("2009_%dbody%s" % (nbody, tag))
    d.mkdir()
    img, truth = synth.make_frame(0, nbody)
    path = str(d / "N2.20090531.29966.LDIF.fits")
    frame.write_fits(path, img, synth.HEADER)
    guess = synth.step1_guess(img, nbody, sky_xy=(100, 120))
    with open(str(d / "29966_initialguess"), "w") as fh:          # apf_step1.py:172-175
        fh.write(" ".join(str(v) for v in guess) + "\n")
    return path, str(d / "29966_apf_results"), truth


def test_step2a_then_step2_with_2a_guess(tmp_path):
    from olpefit_b200 import chains, cli
    path, out, truth = _write_case(tmp_path, 2)
    assert cli.main_step2a([path, "--n-steps", "700", "--seed", "11", "--quiet"]) == 0
    a = np.genfromtxt(os.path.join(out, "step2a.csv"), delimiter=",")
    assert a.shape == (701, 17) and np.all(np.isnan(a[0])) and np.all(np.isfinite(a[1:]))
    assert a[-1, 16] < a[1, 16]                                  # chi-square went down
    assert os.path.exists(os.path.join(out, "step2a_acceptance_rate"))

    args = [path, "-i", "2a", "--walkers", "6", "--accept-min", "30", "--burn-in", "100",
            "--seed", "12", "--segment", "128", "--quiet"]
    assert cli.main_step2(args) == 0
    files = [os.path.join(out, "%d_finalarray_mpi.csv" % w) for w in range(6)]
    assert all(os.path.exists(f) for f in files)
    first = np.genfromtxt(files[0], delimiter=",")
    assert first.shape[1] == 17 and np.all(np.isnan(first[0]))
    cols, npos = chains.ingest(out, 6)                            # equal lengths or this raises
    assert npos == 4 and cols.shape[0] == 17 and cols.shape[2] == 6
    assert np.all(np.isfinite(cols))
    # all walkers start from the last row of step2a.csv (apf_step2.py:254-256)
    for w in range(6):
        rate = open(os.path.join(out, "%d_acceptance_rate.csv" % w)).read()
        vals = np.array(rate.replace("[", " ").replace("]", " ").split(), dtype=float)
        assert vals.shape == (16,) and np.all((vals >= 0) & (vals <= 1))
    # different walkers, different chains; same seed, same chains
    assert not np.array_equal(cols[:, :, 0], cols[:, :, 1])
    keep = first.copy()
    assert cli.main_step2(args) == 0
    assert np.array_equal(np.genfromtxt(files[0], delimiter=",")[1:], keep[1:])
    # separation / position angle of the fitted companion are in the right place
    sep, pa = chains.separation_pa(cols[0], cols[1], cols[2], cols[3])
    assert np.median(sep) == pytest.approx(119.7, abs=4.0)
    assert np.median(pa) == pytest.approx(-51.4, abs=2.0)


def test_step2_default_guess_packed_format_and_3body(tmp_path):
    from olpefit_b200 import chains, cli
    path, out, _ = _write_case(tmp_path, 3)
    script_args = [path, "--walkers", "5", "--accept-min", "12", "--seed", "3", "--stamp", "32",
                   "--thin", "4", "--format", "bin", "--quiet"]
    assert cli.main_step2_3body(script_args) == 0
    arr, meta = chains.read_packed(os.path.join(out, "chains_rank0"))
    assert arr.shape[1:] == (5, 20) and meta["burn_in"] == 0 and meta["thin"] == 4
    assert arr.shape[0] == meta["count"] // 4
    assert chains.unpack_to_csv(os.path.join(out, "chains_rank0"), os.path.join(out, "csv")) == 5
    back = np.genfromtxt(os.path.join(out, "csv", "4_finalarray_mpi.csv"), delimiter=",")
    assert back.shape == (arr.shape[0] + 1, 20) and np.array_equal(back[1:], arr[:, 4, :])

    path2, out2, _ = _write_case(tmp_path, 2, tag="_b")
    # step-1 guess, reference burn-in of 6000 overridden, --fix-bkgd floor
    assert cli.main_step2([path2, "--walkers", "3", "--accept-min", "10", "--burn-in", "0", "--seed", "5",
                           "--fix-bkgd", "--quiet"]) == 0
    c = np.genfromtxt(os.path.join(out2, "0_finalarray_mpi.csv"), delimiter=",")
    assert c.shape[1] == 17 and c.shape[0] > 10


def test_checkpoint_resume_and_adapt_flags(tmp_path):
    """--checkpoint / --resume continue the same chains (the files of an interrupted run followed
    by its resumption equal the files of the uninterrupted run); --adapt only acts during burn-in."""
    from olpefit_b200 import cli
    common = ["--walkers", "4", "--burn-in", "0", "--seed", "21", "--stamp", "32", "--segment", "64", "--quiet"]
    path_a, out_a, _ = _write_case(tmp_path, 2, tag="_whole")
    assert cli.main_step2([path_a, "--accept-min", "24"] + common) == 0
    whole = np.genfromtxt(os.path.join(out_a, "2_finalarray_mpi.csv"), delimiter=",")
    path_b, out_b, _ = _write_case(tmp_path, 2, tag="_parts")
    assert cli.main_step2([path_b, "--accept-min", "9", "--checkpoint"] + common) == 0
    part = np.genfromtxt(os.path.join(out_b, "2_finalarray_mpi.csv"), delimiter=",")
    assert os.path.exists(os.path.join(out_b, "checkpoint_rank0.pt")) and part.shape[0] < whole.shape[0]
    assert cli.main_step2([path_b, "--accept-min", "24", "--resume"] + common) == 0
    both = np.genfromtxt(os.path.join(out_b, "2_finalarray_mpi.csv"), delimiter=",")
    assert both.shape == whole.shape and np.array_equal(both[1:], whole[1:])
    # the device-side statistics continue too: the summary of the two parts is the summary of the whole run
    import json
    sa = json.load(open(os.path.join(out_a, "step2_summary.json")))
    sb = json.load(open(os.path.join(out_b, "step2_summary.json")))
    assert sa["sep_pa_companion"] == sb["sep_pa_companion"] and sa["gelman_rubin"] == sb["gelman_rubin"]
    path_c, out_c, _ = _write_case(tmp_path, 2, tag="_adapt")
    assert cli.main_step2([path_c, "--accept-min", "40", "--adapt", "--walkers", "8", "--burn-in", "600",
                           "--seed", "21", "--stamp", "32", "--quiet"]) == 0
    assert np.genfromtxt(os.path.join(out_c, "0_finalarray_mpi.csv"), delimiter=",").shape[1] == 17


def test_three_body_entry_point_with_all_defaults(tmp_path):
    """3body/apf_step2_3body.py IMAGE with nothing but the stop rule shortened: 24 walkers, 128-pixel
    stamp, whole-frame domain, automatic team size (16 warps per walker on a 3-body 128-pixel stamp)."""
    from olpefit_b200 import chains, cli
    path, out, _ = _write_case(tmp_path, 3)
    assert cli.main_step2_3body([path, "--accept-min", "6", "--seed", "4", "--quiet"]) == 0
    cols, npos = chains.ingest(out, 24)
    assert npos == 6 and cols.shape[0] == 20 and cols.shape[2] == 24 and np.all(np.isfinite(cols))
    assert os.path.exists(os.path.join(out, "step2_summary.json"))


def test_many_epochs_in_one_run_and_device_side_summary(tmp_path):
    """--frames: several epochs, each with its own step-1 file and results directory, advance in one
    batched sampler; step2_summary.json per epoch comes from statistics reduced on the device and
    agrees with the chain files; --no-chains writes the summary alone."""
    import json
    from olpefit_b200 import chains, cli, frame, synth
    paths, outs = [], []
    for f in range(3):
        d = tmp_path / ("epoch%d" % f)
        d.mkdir()
        img, truth = synth.make_frame(f, 2)
        p = str(d / ("N2.2009053%d.2996%d.LDIF.fits" % (f, f)))
        frame.write_fits(p, img, synth.HEADER)
        paths.append(p)
        outs.append(chains.results_dir(p))
        os.makedirs(outs[-1])
        # every epoch starts from its own step-2a result (apf_step2.py:248-256), here the in-model truth
        chains.write_walker_csv(outs[-1] + "step2a.csv", np.concatenate([truth, [0.0]])[None])
    lst = tmp_path / "frames.txt"
    lst.write_text("# epochs of one target\n" + "\n".join(paths[1:]) + "\n")
    args = [paths[0], "-i", "2a", "--frames", str(lst), "--walkers", "6", "--accept-min", "40", "--burn-in", "300", "--seed", "9",
            "--stamp", "32", "--domain", "stamp", "--thin", "2", "--quiet"]
    assert cli.main_step2(args) == 0
    for f in range(3):
        cols, npos = chains.ingest(outs[f], 6)
        assert npos == 4 and cols.shape[2] == 6 and np.all(np.isfinite(cols))
        summ = json.load(open(os.path.join(outs[f], "step2_summary.json")))
        sep, pa = chains.separation_pa(cols[0], cols[1], cols[2], cols[3])
        s = summ["sep_pa_companion"]
        assert s["rows"] == sep.size
        # histogram quantiles are exact to one bin (5e-4 pixel = 0.005 mas; 2e-3 degrees)
        for vals, key, width in ((sep, "sep_mas", 5e-4 * 9.952), (pa, "pa_deg", 2e-3)):
            for level, name in ((0.15865, "lo"), (0.5, "median"), (0.84135, "hi")):
                v = s[key][name]
                assert np.mean(vals < v - width) <= level <= np.mean(vals <= v + width), (key, name, v)
        assert s["sep_mas"]["std"] == pytest.approx(np.std(sep), rel=1e-6)
        assert s["pa_deg"]["mean"] == pytest.approx(np.mean(pa), abs=1e-9)
        assert len(summ["gelman_rubin"]) == 16
        gr = [chains.gelman_rubin(cols[j])[1] for j in range(4)]
        np.testing.assert_allclose(summ["gelman_rubin"][:4], gr, rtol=1e-6)
    # the epochs differ (the companion drifts), and so do their chains
    a = np.genfromtxt(os.path.join(outs[0], "0_finalarray_mpi.csv"), delimiter=",")
    b = np.genfromtxt(os.path.join(outs[2], "0_finalarray_mpi.csv"), delimiter=",")
    assert a.shape == b.shape and not np.array_equal(a[1:], b[1:])
    # statistics only: same summary, no chain files touched
    for o in outs:
        for w in range(6):
            os.unlink(os.path.join(o, "%d_finalarray_mpi.csv" % w))
    before = [json.load(open(os.path.join(o, "step2_summary.json"))) for o in outs]
    assert cli.main_step2(args + ["--no-chains"]) == 0
    for f in range(3):
        assert not os.path.exists(os.path.join(outs[f], "0_finalarray_mpi.csv"))
        assert json.load(open(os.path.join(outs[f], "step2_summary.json"))) == before[f]


def test_packed_float32_chain_and_its_resume(tmp_path):
    """--format bin --chain-dtype f32, interrupted and resumed: the packed file of the two parts
    equals the packed file of the uninterrupted run (--resume used to truncate it)."""
    from olpefit_b200 import chains, cli
    common = ["--walkers", "5", "--burn-in", "0", "--seed", "33", "--stamp", "32", "--segment", "64", "--format", "bin",
              "--chain-dtype", "f32", "--quiet"]
    path_a, out_a, _ = _write_case(tmp_path, 2, tag="_whole")
    assert cli.main_step2([path_a, "--accept-min", "20"] + common) == 0
    whole, meta = chains.read_packed(os.path.join(out_a, "chains_rank0"))
    assert meta["dtype"] == "float32" and whole.shape[1:] == (5, 17)
    path_b, out_b, _ = _write_case(tmp_path, 2, tag="_parts")
    assert cli.main_step2([path_b, "--accept-min", "8", "--checkpoint"] + common) == 0
    part, _ = chains.read_packed(os.path.join(out_b, "chains_rank0"))
    assert 0 < part.shape[0] < whole.shape[0]
    assert cli.main_step2([path_b, "--accept-min", "20", "--resume"] + common) == 0
    both, meta_b = chains.read_packed(os.path.join(out_b, "chains_rank0"))
    assert both.shape == whole.shape and np.array_equal(both, whole) and meta_b["count"] == meta["count"]
