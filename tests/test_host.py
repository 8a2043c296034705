"""CPU tests of the host side: the C-ABI library loads and exports what include/lapf.h declares,
argument validation fails loudly without a GPU, chain files and FITS I/O round-trip, and the
step-3 statistics agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import lapf_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from olpefit_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from olpefit_b200 import _lib
    header = open(os.path.join(ROOT, "include", "lapf.h")).read()
    declared = set(re.findall(r"\b(lapf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.lapf_abi_version() == 2


def test_tables_served_by_library_match_reference_tables(lib):
    from olpefit_b200 import layout
    for nbody in (2, 3):
        lay = orc.layout_for(nbody)
        w, is_log = layout.default_widths(nbody)
        assert layout.nparam(nbody) == lay.nparam
        assert np.array_equal(w, np.asarray(lay.widths))
        assert np.array_equal(is_log, lay.is_log())
        assert len(layout.names(nbody)) == lay.nparam and layout.names(nbody) == lay.names
    assert lib.lapf_num_params(4) < 0 and b"nbody" in lib.lapf_last_error()


def test_compute_calls_fail_loudly_without_a_device(lib):
    """No CPU fallback: validation errors are reported, and with valid-looking arguments a box
    without an sm_100 device gets LAPF_ERR_NO_DEVICE, never a silent CPU result."""
    import torch
    from olpefit_b200 import _lib
    bad = _lib.Problem(5, 32, 32, 1, 12, 0, 256, 512, 768)
    assert lib.lapf_model_chi2(C.byref(bad), 1024, 1, None, None, 2048, None) == -1
    bad = _lib.Problem(2, 32, 32, 1, 1, 0, 256, 512, 768)          # a position slot cannot be the floor
    assert lib.lapf_model_chi2(C.byref(bad), 1024, 1, None, None, 2048, None) == -1
    bad = _lib.Problem(2, 32, 32, 1, 12, 0, 260, 512, 768)         # misaligned for TMA
    assert lib.lapf_model_chi2(C.byref(bad), 1024, 1, None, None, 2048, None) == -1
    if not torch.cuda.is_available():
        ok = _lib.Problem(2, 32, 32, 1, 12, 0, 256, 512, 768)
        assert lib.lapf_model_chi2(C.byref(ok), 1024, 1, None, None, 2048, None) == -3
        out = (C.c_double * 4)()
        assert lib.lapf_measure_peaks(out) == -3
        with pytest.raises(_lib.LapfError):
            from olpefit_b200 import model
            model.PixelDomain(np.zeros((32, 32), np.float32), np.ones((32, 32), np.float32))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "olpefit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in src and "import oracle" not in src, f
    for f in ("apf_step2.py", "apf_step2a.py", os.path.join("3body", "apf_step2_3body.py")):
        src = open(os.path.join(ROOT, f)).read()
        assert "oracle" not in src


# ---------------------------------------------------------------------------------------------
# chain files
# ---------------------------------------------------------------------------------------------
def test_walker_csv_matches_reference_layout(tmp_path):
    """apf_step2.py:357-360: csv.writer rows, first row all nan; readable by np.genfromtxt exactly
    as apf_step3.py:169,181 does; values survive bit-for-bit (shortest round-trip decimals)."""
    import csv
    from olpefit_b200 import chains
    rng = np.random.default_rng(0)
    rows = rng.normal(size=(50, 17)) * 10.0 ** rng.integers(-8, 8, size=(50, 17))
    rows[3, 4] = 0.0
    rows[7, 2] = 15000.0
    rows[9, 16] = 1e22
    path = str(tmp_path / "0_finalarray_mpi.csv")
    chains.write_walker_csv(path, rows)
    back = np.genfromtxt(path, delimiter=",")
    assert back.shape == (51, 17) and np.all(np.isnan(back[0]))
    assert np.array_equal(back[1:], rows)
    raw = open(path, "rb").read()
    assert raw.count(b"\r\n") == 51 and raw.startswith(b"nan,nan")
    # the reference's own writer parses back to the same numbers
    ref = str(tmp_path / "ref.csv")
    with open(ref, "w", newline="") as fh:
        csv.writer(fh).writerows(np.vstack([np.full((1, 17), np.nan), rows]))
    assert np.array_equal(np.genfromtxt(ref, delimiter=",")[1:], back[1:])


def test_segment_append_equals_single_write(tmp_path):
    from olpefit_b200 import chains
    rng = np.random.default_rng(1)
    seg = rng.normal(size=(12, 3, 20))                         # [rows, walkers, P+1], 3-body width
    whole = [str(tmp_path / ("w%d.csv" % w)) for w in range(3)]
    parts = [str(tmp_path / ("p%d.csv" % w)) for w in range(3)]
    chains.write_segment_csv(whole, seg, first=True)
    chains.write_segment_csv(parts, seg[:5], first=True)
    chains.write_segment_csv(parts, seg[5:], first=False)
    for a, b, w in zip(whole, parts, range(3)):
        assert open(a, "rb").read() == open(b, "rb").read()
        assert np.array_equal(np.genfromtxt(a, delimiter=",")[1:], seg[:, w, :])


def test_acceptance_file_and_paths(tmp_path):
    from olpefit_b200 import chains
    p = str(tmp_path / "0_acceptance_rate.csv")
    chains.write_acceptance(p, [1, 2, 3], [2, 4, 6])
    assert open(p).read() == str(np.array([0.5, 0.5, 0.5]))
    img = "../1RXSJ1609/2009/N2.20090531.29966.LDIF.fits"       # the reference's own example
    assert chains.results_dir(img) == "../1RXSJ1609/2009/29966_apf_results/"
    assert chains.initial_guess_path(img) == "../1RXSJ1609/2009/29966_initialguess"


def test_packed_chain_round_trip_and_unpack(tmp_path):
    from olpefit_b200 import chains
    rng = np.random.default_rng(2)
    base = str(tmp_path / "chains_rank0")
    w = chains.PackedChainWriter(base, 4, 17, {"seed": 7})
    a, b = rng.normal(size=(6, 4, 17)), rng.normal(size=(3, 4, 17))
    w.append(a)
    w.append(b)
    w.close(count=99)
    arr, meta = chains.read_packed(base)
    assert np.array_equal(arr, np.concatenate([a, b])) and meta["rows"] == 9 and meta["count"] == 99
    out = str(tmp_path / "csv")
    assert chains.unpack_to_csv(base, out) == 4
    assert np.array_equal(np.genfromtxt(os.path.join(out, "2_finalarray_mpi.csv"), delimiter=",")[1:], arr[:, 2, :])


def test_ingest_and_statistics_match_oracle(tmp_path):
    """The step-3 side (apf_step3.py:169-214, 260-291): ingest drops the nan row, adds 1 to the
    positions; Gelman-Rubin and separation / position angle equal the oracle's restatement."""
    from olpefit_b200 import chains
    rng = np.random.default_rng(3)
    ncor, n = 5, 40
    base = np.array([512.3, 511.7, 521.7, 519.2] + [1.0] * 13)
    data = base + rng.normal(size=(n, ncor, 17)) * 0.01
    for w in range(ncor):
        chains.write_walker_csv(str(tmp_path / ("%d_finalarray_mpi.csv" % w)), data[:, w, :])
    cols, npos = chains.ingest(str(tmp_path), ncor)
    assert cols.shape == (17, n, ncor) and npos == 4
    np.testing.assert_array_equal(cols[4:], np.moveaxis(data, 2, 0)[4:])
    np.testing.assert_allclose(cols[:4], np.moveaxis(data, 2, 0)[:4] + 1.0, rtol=0, atol=1e-12)
    for j in range(16):
        assert chains.gelman_rubin(cols[j]) == pytest.approx(orc.gelman_rubin(cols[j]))
    sep, pa = chains.separation_pa(cols[0], cols[1], cols[2], cols[3])
    sep_o, pa_o = orc.separation_pa(cols[0], cols[1], cols[2], cols[3])
    np.testing.assert_array_equal(sep, sep_o)
    np.testing.assert_array_equal(pa, pa_o)
    assert np.median(sep) == pytest.approx(12.03 * 9.952, rel=2e-3)
    assert np.median(pa) == pytest.approx(-51.4, abs=0.2)
    # unequal files are refused, like the fixed-length arrays of apf_step3.py:183
    chains.write_walker_csv(str(tmp_path / "1_finalarray_mpi.csv"), data[:-1, 1, :])
    with pytest.raises(ValueError):
        chains.ingest(str(tmp_path), ncor)


def test_gelman_rubin_from_device_moments_equals_direct():
    from olpefit_b200 import chains
    rng = np.random.default_rng(4)
    x = 3.0 + rng.normal(size=(200, 7)) * 0.1 + rng.normal(size=(1, 7)) * 0.05
    means = x.mean(axis=0)
    ref = x[0, 0]
    mom = np.array([ref, (means - ref).sum(), ((means - ref) ** 2).sum(), x.var(axis=0).sum()])
    psrf, rc = chains.gelman_rubin_from_moments(mom, 200, 7)
    psrf_d, rc_d = chains.gelman_rubin(x)
    assert psrf == pytest.approx(psrf_d, rel=1e-9) and rc == pytest.approx(rc_d, rel=1e-9)


# ---------------------------------------------------------------------------------------------
# FITS + header scalars
# ---------------------------------------------------------------------------------------------
def test_fits_round_trip_and_header_scalars(tmp_path):
    from olpefit_b200 import frame
    rng = np.random.default_rng(5)
    img = rng.normal(size=(40, 56)).astype(np.float32)
    path = str(tmp_path / "N2.20090531.29966.LDIF.fits")
    frame.write_fits(path, img, {"ITIME": 1.5, "COADDS": 10, "MULTISAM": 4, "SAMPMODE": 3, "OBJECT": "test"})
    back, hdr = frame.read_fits(path)
    assert back.dtype == np.float32 and np.array_equal(back, img)
    assert os.path.getsize(path) % 2880 == 0
    assert hdr["itime"] == 1.5 and hdr["COADDS"] == 10 and hdr["object"] == "test"
    assert frame.saturation_level(hdr) == orc.saturation_level(hdr) == 10 * 24000.0 * (1.0 - 0.1 * 3.0 / 1.5)
    assert frame.read_noise(hdr) == orc.read_noise(hdr) == (38.0 / 2.0) * np.sqrt(10.0)
    h2 = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
    assert frame.saturation_level(h2) == 22000.0 and frame.read_noise(h2) == 38.0
    # int16 with BSCALE/BZERO
    raw = (rng.integers(-100, 100, size=(8, 8))).astype(np.int16)
    p2 = str(tmp_path / "a.b.c.fits")
    frame.write_fits(p2, raw, {"BSCALE": 2.0, "BZERO": 10.0})
    b2, _ = frame.read_fits(p2)
    assert np.array_equal(b2, raw * 2.0 + 10.0)


def test_initial_parameters_match_oracle():
    from olpefit_b200 import frame, synth
    for nbody in (2, 3):
        lay = orc.layout_for(nbody)
        ox, oy = synth.stamp_origin(64, nbody)
        img, _ = synth.make_frame(2, nbody, region=(oy, oy + 64, ox, ox + 64))
        guess = synth.step1_guess(img, nbody, origin=(ox, oy))
        p = frame.initial_parameters(img, guess, nbody, origin=(ox, oy))
        g_local = guess - np.array([ox, oy] * nbody + [ox, oy], dtype=np.float64)
        ref = orc.initial_parameters(img, g_local, lay)      # same dtype: the median of float32 stays float32
        ref[0:2 * nbody:2] += ox
        ref[1:2 * nbody:2] += oy
        np.testing.assert_array_equal(p, ref)


def test_cli_parsers_keep_reference_flags():
    from olpefit_b200 import cli
    a = cli._parser("step2").parse_args(["x/N2.1.2.LDIF.fits", "-i", "2a"])
    assert a.image.endswith(".fits") and a.initial_guess_option == "2a" and a.accept_min == 100000 and a.walkers == 24
    a = cli._parser("step2").parse_args(["img.a.b.c.fits", "--initial_guess_option", "1"])
    assert a.initial_guess_option == "1"
    a = cli._parser("step2a").parse_args(["img.a.b.c.fits"])
    assert a.n_steps == 5000 and not hasattr(a, "initial_guess_option")
    a = cli._parser("step2_3body").parse_args(["img.a.b.c.fits"])
    assert not hasattr(a, "initial_guess_option")


def test_automatic_step1_writes_the_reference_file(tmp_path):
    """apf_step1.py:145-175 without the clicks: brightest pixel of the 21 x 21 box + 0.5, one line."""
    import apf_step1_auto
    from olpefit_b200 import frame, step1, synth
    img, truth = synth.make_frame(0, 2, hot_pixels=False)
    path = str(tmp_path / "N2.20090531.29966.LDIF.fits")
    frame.write_fits(path, img, synth.HEADER)
    assert apf_step1_auto.main([path, "--star", "509,514", "--companion", "521.7,519.2", "--sky", "100.9,120.2"]) == 0
    g = np.loadtxt(open(str(tmp_path / "29966_initialguess"), "rb"), delimiter=" ")     # apf_step2.py:262
    assert g.shape == (6,)
    iy, ix = np.unravel_index(np.argmax(img[503:524, 498:519]), (21, 21))
    assert (g[0], g[1]) == (498 + ix + 0.5, 503 + iy + 0.5)
    assert abs(g[0] - truth[0]) < 1.5 and abs(g[1] - truth[1]) < 1.5
    assert g[4] == 100 and g[5] == 120
    raw = step1.initial_guess(img, [(509, 514), (521.7, 519.2)], (100, 120), refine=False)
    assert raw[:4] == [509.0, 514.0, 521.7, 519.2]
    # the numbers feed step 2's starting point exactly like the reference's file does
    p = frame.initial_parameters(img, g, 2)
    assert p.shape == (16,) and p[6] == img[int(g[1] - 1), int(g[0] - 1)]


def test_packed_chain_float32_differences_and_resume(tmp_path):
    """--format bin --chain-dtype f32 stores float32 differences from the starting points; a resumed
    run appends to the same file instead of truncating it (chains.PackedChainWriter)."""
    from olpefit_b200 import chains
    rng = np.random.default_rng(3)
    start = np.array([[512.3, 511.7, 15000.0, 6.4, 1040.5]] * 4) + rng.normal(0, 1e-3, (4, 5))
    full = start[None] + rng.normal(0, 3e-3, (9, 4, 5)) * np.array([1, 1, 100.0, 0.01, 1.0])
    delta = (full - start[None]).astype(np.float32)
    base = str(tmp_path / "chains_rank0")
    w = chains.PackedChainWriter(base, 4, 5, {"seed": 1}, dtype="float32", start=start)
    w.append(delta[:5])
    w.close(count=5)
    with open(base + ".bin", "ab") as fh:                       # a torn segment of an interrupted run
        fh.write(b"\0" * 13)
    w = chains.PackedChainWriter(base, 4, 5, dtype="float32", resume=True)
    assert w.rows == 5
    w.append(delta[5:])
    w.close(count=9)
    arr, meta = chains.read_packed(base)
    assert arr.shape == (9, 4, 5) and meta["rows"] == 9 and meta["count"] == 9 and meta["seed"] == 1
    assert np.array_equal(arr, start[None] + delta.astype(np.float64))
    # positions keep 1e-9 absolute although they are stored in 32 bits
    assert np.max(np.abs(arr[..., 0] - full[..., 0])) < 2e-9
    with pytest.raises(ValueError):
        chains.PackedChainWriter(base, 3, 5, dtype="float32", resume=True)
    # float64 resume
    base2 = str(tmp_path / "c64")
    w = chains.PackedChainWriter(base2, 4, 5)
    w.append(full[:2]); w.close()
    w = chains.PackedChainWriter(base2, 4, 5, resume=True)
    w.append(full[2:]); w.close()
    assert np.array_equal(chains.read_packed(base2)[0], full)


def test_histogram_quantiles_match_numpy_percentiles():
    """stats.quantiles_from_hist (the host end of lapf_sampler_sketch) against np.percentile of the
    values the histogram was filled from: equal to within one bin width."""
    import torch
    from olpefit_b200 import stats
    rng = np.random.default_rng(11)
    n_bins, width, centre = 2048, 2e-3, 12.03
    x = np.concatenate([rng.normal(12.031, 0.04, 150000), rng.normal(12.2, 0.01, 3000)])
    pos = np.floor((x - centre) / width) + n_bins // 2
    b = np.where(pos < 0, 0, np.where(pos >= n_bins, n_bins + 1, pos + 1)).astype(int)
    hist = np.bincount(b, minlength=n_bins + 2)
    q = stats.quantiles_from_hist(torch.tensor(hist)[None], torch.tensor([centre]), torch.tensor([[width]]))[0].numpy()
    ref = np.percentile(x, [15.865, 50.0, 84.135])
    assert np.max(np.abs(q - ref)) < width
    # nothing recorded, or the quantile lies outside the range: nan, not a made-up number
    empty = stats.quantiles_from_hist(torch.zeros((1, n_bins + 2)), torch.tensor([centre]), torch.tensor([[width]]))
    assert bool(torch.isnan(empty).all())
    far = np.zeros(n_bins + 2); far[-1] = 10
    assert bool(torch.isnan(stats.quantiles_from_hist(torch.tensor(far)[None], torch.tensor([centre]), torch.tensor([[width]]))).all())


def test_frame_list_of_the_many_epoch_command_line(tmp_path):
    """--frames LIST: one FITS path per line, relative to the list file, comments and blank lines
    skipped, IMAGE first and never twice (cli._frame_list)."""
    import argparse
    from olpefit_b200 import cli
    d = tmp_path / "night"
    d.mkdir()
    lst = d / "frames.txt"
    lst.write_text("# epochs\n\nb.fits\n%s\n../night/a.fits\nc.fits\n" % (d / "c.fits"))
    got = cli._frame_list(argparse.Namespace(image=str(d / "a.fits"), frames=str(lst)))
    assert [os.path.basename(p) for p in got] == ["a.fits", "b.fits", "c.fits"]
    assert cli._frame_list(argparse.Namespace(image="x.fits", frames=None)) == ["x.fits"]
    # the new flags are additive: the reference's positional argument and -i still parse alone
    a = cli._parser("step2").parse_args(["img.fits", "-i", "2a", "--frames", "l.txt", "--no-chains", "--chain-dtype", "f32"])
    assert a.image == "img.fits" and a.initial_guess_option == "2a" and a.frames == "l.txt" and a.no_chains
