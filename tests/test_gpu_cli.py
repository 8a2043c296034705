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
    path_c, out_c, _ = _write_case(tmp_path, 2, tag="_adapt")
    assert cli.main_step2([path_c, "--accept-min", "40", "--adapt", "--walkers", "8", "--burn-in", "600",
                           "--seed", "21", "--stamp", "32", "--quiet"]) == 0
    assert np.genfromtxt(os.path.join(out_c, "0_finalarray_mpi.csv"), delimiter=",").shape[1] == 17
