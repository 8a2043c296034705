"""Chain files (the step-2 -> step-3 hand-off) and the chain statistics step 3 derives from them.

File layout kept from the reference:
  <dir>/<N>_apf_results/<w>_finalarray_mpi.csv   apf_step2.py:357-360: no header, first row all
        nan (:278-279), then one row per recorded update: P parameters + chi-square
        (17 columns 2-body, 20 columns 3-body); comma separated, CRLF
  <dir>/<N>_apf_results/<w>_acceptance_rate.csv  apf_step2.py:362-365: str(accept/tries)
  step2a.csv / step2a_acceptance_rate            apf_step2a.py:324,329
plus a packed binary format for batches far beyond what per-walker text files can hold.

Statistics restated from apf_step3.py: ingest (:169-214), Gelman-Rubin (:260-278),
separation / position angle (:255-256,283-291), summary (:436-437).
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

from . import _lib

CSV_LEADING_NAN_ROW = 1
CSV_APPEND = 2

PIXSCALE_PRE2015 = 9.952    # mas / pixel, apf_step3.py:224
PIXSCALE_POST2015 = 9.971   # apf_step3.py:231


# ----------------------------------------------------------------------------------------------
# output directory and file names (apf_step2.py:164-173)
# ----------------------------------------------------------------------------------------------
def split_image_path(image_path: str):
    """(directory with trailing '/', image number) -- apf_step2.py:164-170."""
    d = image_path.split("/")
    directory = ""
    for part in d[:-1]:
        directory = directory + str(part) + "/"
    return directory, d[-1].split(".")[-3]


def results_dir(image_path: str) -> str:
    directory, number = split_image_path(image_path)
    return directory + number + "_apf_results/"


def initial_guess_path(image_path: str) -> str:
    directory, number = split_image_path(image_path)
    return directory + number + "_initialguess"


# ----------------------------------------------------------------------------------------------
# writers
# ----------------------------------------------------------------------------------------------
def write_walker_csv(path, rows, leading_nan=True, append=False):
    """Write (or append) one walker's chain rows [n, P+1] float64 through liblapf's native writer."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    if rows.ndim != 2:
        raise ValueError("rows must be [n, columns]")
    flags = (CSV_LEADING_NAN_ROW if leading_nan else 0) | (CSV_APPEND if append else 0)
    _lib.check(_lib.load().lapf_write_chain_csv(os.fsencode(path), rows.ctypes.data, rows.shape[0],
                                                rows.shape[1], rows.shape[1], flags))


_pool = None


def _writer_pool():
    global _pool
    if _pool is None:
        import concurrent.futures as cf
        _pool = cf.ThreadPoolExecutor(max_workers=max(1, min(32, (os.cpu_count() or 1))))
    return _pool


def write_segment_csv(paths, segment, first):
    """Append a time-major chain segment [rows, W, P+1] to the W per-walker files ``paths``.
    One file per walker is the reference's layout (apf_step2.py:357); the walkers' files are
    formatted and written concurrently (the native writer runs outside the GIL)."""
    seg = np.ascontiguousarray(segment, dtype=np.float64)
    nrow, nw, ncol = seg.shape
    lib = _lib.load()
    flags = CSV_LEADING_NAN_ROW if first else CSV_APPEND

    def one(w):
        base = seg.ctypes.data + w * ncol * 8
        return lib.lapf_write_chain_csv(os.fsencode(paths[w]), base, nrow, ncol, nw * ncol, flags)

    if nw == 1:
        _lib.check(one(0))
        return
    for rc in _writer_pool().map(one, range(nw)):
        _lib.check(rc)


def write_acceptance(path, accepts, tries):
    """apf_step2.py:362-365: the numpy repr of total_accept / total_tries."""
    with np.errstate(all="ignore"):
        rate = np.asarray(accepts, dtype=np.float64) / np.asarray(tries, dtype=np.float64)
    with open(path, "w") as fh:
        fh.write(str(rate))


class PackedChainWriter:
    """Packed binary chain: [row][walker][P+1] appended segment by segment to ``<base>.bin`` with a
    JSON sidecar ``<base>.json``; for batches of 10^4..10^6 walkers where one text file per walker
    is not workable.  dtype 'float64' holds the values; 'float32' holds differences from the
    walkers' starting points, which are stored once as float64 in ``<base>.start.bin``
    (LAPF_CHAIN_F32_DELTA).  ``resume=True`` appends to an existing file of the same shape (the
    sidecar tells the row count so far).  ``read_packed`` / ``unpack_to_csv`` convert back."""

    def __init__(self, base, n_walkers, n_cols, meta=None, dtype="float64", start=None, resume=False):
        self.base, self.n_walkers, self.n_cols = base, int(n_walkers), int(n_cols)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype("float64"), np.dtype("float32")):
            raise ValueError("packed chains are float64 or float32 (differences from the start)")
        self.rows = 0
        self.meta = dict(meta or {})
        if resume:
            with open(base + ".json") as fh:
                old = json.load(fh)
            if (old["n_walkers"], old["n_cols"], old["dtype"]) != (self.n_walkers, self.n_cols, self.dtype.name):
                raise ValueError("%s.bin holds %s walkers x %s columns of %s: cannot append %d x %d of %s"
                                 % (base, old["n_walkers"], old["n_cols"], old["dtype"], self.n_walkers,
                                    self.n_cols, self.dtype.name))
            have = os.path.getsize(base + ".bin")
            want = old["rows"] * self.n_walkers * self.n_cols * self.dtype.itemsize
            if have < want:
                raise ValueError("%s.bin is shorter than its sidecar says" % base)
            self.rows = int(old["rows"])
            self.meta = {**old, **self.meta}
            self.fh = open(base + ".bin", "r+b")
            self.fh.truncate(want)                      # drop a partial segment of an interrupted run
            self.fh.seek(want)
        else:
            self.fh = open(base + ".bin", "wb")
            if self.dtype == np.dtype("float32"):
                if start is None:
                    raise ValueError("float32 chains are differences: the starting points are needed")
                st = np.ascontiguousarray(start, dtype=np.float64)
                assert st.shape == (self.n_walkers, self.n_cols)
                st.tofile(base + ".start.bin")

    def append(self, segment):
        seg = np.ascontiguousarray(segment, dtype=self.dtype)
        assert seg.shape[1:] == (self.n_walkers, self.n_cols)
        self.fh.write(seg.tobytes())
        self.rows += seg.shape[0]

    def close(self, **extra):
        self.fh.close()
        self.meta.update(extra)
        self.meta.update({"rows": self.rows, "n_walkers": self.n_walkers, "n_cols": self.n_cols,
                          "dtype": self.dtype.name, "order": "[row][walker][column]",
                          "values": "differences from <base>.start.bin" if self.dtype == np.dtype("float32") else "absolute"})
        with open(self.base + ".json", "w") as fh:
            json.dump(self.meta, fh, indent=1)


def read_packed(base):
    """(chain [rows, walkers, columns] float64, meta) of a packed chain, whatever its storage type."""
    with open(base + ".json") as fh:
        meta = json.load(fh)
    shape = (meta["rows"], meta["n_walkers"], meta["n_cols"])
    dt = np.dtype(meta.get("dtype", "float64"))
    arr = np.fromfile(base + ".bin", dtype=dt, count=int(np.prod(shape))).reshape(shape)
    if dt == np.dtype("float32"):
        start = np.fromfile(base + ".start.bin", dtype=np.float64).reshape(shape[1:])
        arr = start[None] + arr.astype(np.float64)
    return arr, meta


def unpack_to_csv(base, out_dir, walkers=None):
    """Per-walker reference CSVs from a packed chain, so an unmodified apf_step3.py can read them."""
    arr, meta = read_packed(base)
    os.makedirs(out_dir, exist_ok=True)
    ids = range(arr.shape[1]) if walkers is None else walkers
    for w in ids:
        write_walker_csv(os.path.join(out_dir, "%d_finalarray_mpi.csv" % w), arr[:, w, :])
    return len(list(ids))


# ----------------------------------------------------------------------------------------------
# step-3 side: ingest and statistics
# ----------------------------------------------------------------------------------------------
def read_walker_csv(path):
    return np.genfromtxt(path, delimiter=",")


def ingest(input_directory, ncor, additional_burnin=1):
    """apf_step3.py:169-214: load ncor walker files into [length, ncor] arrays per column, drop
    the first ``additional_burnin`` rows (default 1 = the nan row, :81-84), add 1 to the positions
    (FITS pixels are 1-based, :211-214).  Every file must have the same number of rows (:183).
    Returns (columns [n_cols, length-burn, ncor], n_position_columns)."""
    if not additional_burnin:
        additional_burnin = 1
    first = read_walker_csv(os.path.join(input_directory, "0_finalarray_mpi.csv"))
    length, ncols = first.shape
    cols = np.zeros((ncols, length, ncor))
    for i in range(ncor):
        a = first if i == 0 else read_walker_csv(os.path.join(input_directory, "%d_finalarray_mpi.csv" % i))
        if a.shape != (length, ncols):
            raise ValueError("walker %d has %s rows/cols, walker 0 has %s" % (i, a.shape, (length, ncols)))
        cols[:, :, i] = a.T
    cols = cols[:, additional_burnin:length, :]
    npos = 4 if ncols == 17 else 6
    cols[:npos] += 1.0
    return cols, npos


def gelman_rubin(chains, python2_division=True):
    """apf_step3.py:262-276 for one parameter; ``chains`` is [n_rows, n_walkers].  Returns
    (PSRF, RC).  The reference computes (d+3)/(d+1) with d = 16 under Python 2: integer 1."""
    chains = np.asarray(chains, dtype=np.float64)
    n, m = float(chains.shape[0]), float(chains.shape[1])
    w = (1.0 / m) * np.sum(np.std(chains, axis=0) ** 2)
    b = (n / (m - 1.0)) * np.sum((np.mean(chains, axis=0) - np.mean(chains)) ** 2)
    psrf = (((n - 1.0) / n) * w + ((m + 1.0) / (m * n)) * b) / w
    factor = 1.0 if python2_division else 19.0 / 17.0
    return psrf, math.sqrt(factor * psrf)


def gelman_rubin_from_moments(moments, n_rows, n_walkers, python2_division=True):
    """The same statistic from the device-reduced sufficient statistics of ``lapf_sampler_stats``:
    moments[..., 0] = reference value r, [..., 1] = sum of (chain mean - r), [..., 2] = sum of
    (chain mean - r)^2, [..., 3] = sum of chain variances.  Equal-length chains make the overall
    mean the mean of the chain means."""
    mom = np.asarray(moments, dtype=np.float64)
    n, m = float(n_rows), np.asarray(n_walkers, dtype=np.float64)
    s_mean, s_mean2, s_var = mom[..., 1], mom[..., 2], mom[..., 3]
    w = s_var / m
    overall = s_mean / m
    b = (n / (m - 1.0)) * (s_mean2 - m * overall * overall)
    with np.errstate(divide="ignore", invalid="ignore"):   # a parameter that never moved has w = 0
        psrf = (((n - 1.0) / n) * w + ((m + 1.0) / (m * n)) * b) / w
    factor = 1.0 if python2_division else 19.0 / 17.0
    return psrf, np.sqrt(factor * psrf)


def separation_pa(xcs, ycs, xcc, ycc, pixscale=PIXSCALE_PRE2015):
    """apf_step3.py:255-256,283-291 (without the distortion lookup: its FITS tables are not part
    of the checkout): sep = sqrt(dx^2+dy^2)*pixscale [mas], pa = degrees(atan2(-dx, dy))."""
    dy = np.asarray(ycc) - np.asarray(ycs)
    dx = np.asarray(xcc) - np.asarray(xcs)
    return np.sqrt(dy * dy + dx * dx) * pixscale, np.degrees(np.arctan2(-dx, dy))


def summarize(values):
    """median, 16th/84th percentiles, mean, std (apf_step3.py:436-437 uses median and std)."""
    v = np.asarray(values, dtype=np.float64).ravel()
    lo, med, hi = np.percentile(v, [15.865, 50.0, 84.135])
    return {"median": float(med), "lo": float(lo), "hi": float(hi), "mean": float(v.mean()),
            "std": float(v.std())}
