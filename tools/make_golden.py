#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING the reference's own source lines.

Run in the authoring container only (needs /root/reference, read-only):

    python tools/make_golden.py

The reference scripts are Python-2 programs that run at import and need astropy, mpi4py
and a FITS file, so they cannot be imported.  Their function bodies and the sampler loop
body are, however, plain Python-3-clean numpy code.  This script slices those line ranges
out of the files where they lie, ``exec``s them unmodified in a namespace that provides

  * ``np`` (this image's numpy),
  * ``models.Gaussian2D`` -- a stand-in for the one missing third-party symbol, restating
    astropy's published ``Gaussian2D.evaluate`` formula (astropy/modeling/functional_models.py),
  * the module-level names the sliced code expects (``image``, ``imhdr``, ``xsize`` ...),

and stores inputs + outputs as small fixtures.  Nothing from the reference is copied into the
repository: only numbers it computed.
"""
import os
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from olpefit_b200 import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


class _Gaussian2D:
    """Stand-in for astropy.modeling.models.Gaussian2D (constructor keywords as used at
    apf_step2.py:98-101; ``__call__(x, y)`` evaluates on grids)."""

    def __init__(self, amplitude, x_mean, y_mean, x_stddev, y_stddev, theta):
        self.args = (amplitude, x_mean, y_mean, x_stddev, y_stddev, theta)

    def __call__(self, x, y):
        amplitude, x_mean, y_mean, x_stddev, y_stddev, theta = self.args
        cost2 = np.cos(theta) ** 2
        sint2 = np.sin(theta) ** 2
        sin2t = np.sin(2.0 * theta)
        xstd2 = x_stddev ** 2
        ystd2 = y_stddev ** 2
        xdiff = x - x_mean
        ydiff = y - y_mean
        a = 0.5 * ((cost2 / xstd2) + (sint2 / ystd2))
        b = 0.5 * ((sin2t / xstd2) - (sin2t / ystd2))
        c = 0.5 * ((sint2 / xstd2) + (cost2 / ystd2))
        return amplitude * np.exp(-((a * xdiff ** 2) + (b * xdiff * ydiff) + (c * ydiff ** 2)))


class _Models:
    Gaussian2D = _Gaussian2D


def ref_lines(relpath, first, last, dedent=False):
    with open(os.path.join(REF, relpath)) as fh:
        lines = fh.readlines()[first - 1:last]
    src = "".join(lines)
    return textwrap.dedent(src) if dedent else src


class _Rank0Comm:
    def barrier(self):
        pass


VARIANTS = {
    # file, functions, frame-prep ranges, tables, sigma, init, trackers, initial chi, loop body
    2: dict(path="apf_step2.py", funcs=[(63, 70), (78, 148)], prep=[(176, 188), (197, 210)],
            tables=[(215, 217), (234, 234)], sigma=(242, 245), init=(265, 273), track=(276, 276),
            chi0=[(283, 285), (289, 289)], body=(301, 333)),
    3: dict(path="3body/apf_step2_3body.py", funcs=[(63, 70), (78, 141)], prep=[(167, 179), (188, 201)],
            tables=[(220, 238), (292, 295)], sigma=(246, 249), init=(255, 265), track=(298, 298),
            chi0=[(307, 309), (313, 313)], body=(327, 368)),
}


def run_variant(nbody, size, n_trace, seed):
    v = VARIANTS[nbody]
    ox, oy = synth.stamp_origin(size, nbody)
    image32, truth = synth.make_frame(0, nbody, region=(oy, oy + size, ox, ox + size))
    # The reference indexes pixels from 0 at the array corner: feed it the cut-out as its whole
    # "image", with step-1 coordinates expressed relative to the cut-out.
    image = image32.astype(np.float64)
    guess = synth.step1_guess(image, nbody, origin=(ox, oy)) - np.array(
        [ox, oy] * nbody + [ox, oy], dtype=np.float64)
    ns = {"np": np, "models": _Models, "image": image, "rank": 1, "comm": _Rank0Comm(),
          "imhdr": {k.lower(): val for k, val in synth.HEADER.items()}, "guess": guess}
    for a, b in v["funcs"]:
        exec(ref_lines(v["path"], a, b), ns)
    for a, b in v["prep"]:
        exec(ref_lines(v["path"], a, b), ns)
    for a, b in v["tables"]:
        exec(ref_lines(v["path"], a, b), ns)
    exec(ref_lines(v["path"], *v["sigma"]), ns)
    ns["ysize"], ns["xsize"] = image.shape[1], image.shape[0]          # apf_step2.py:237
    exec(ref_lines(v["path"], *v["init"], dedent=True), ns)
    p_init = ns["parameters"].copy()
    exec(ref_lines(v["path"], *v["track"]), ns)
    for a, b in v["chi0"]:
        exec(ref_lines(v["path"], a, b), ns)
    chi_init = float(ns["parameters"][-1])

    out = {"size": size, "origin": np.array([ox, oy]), "image": image32, "guess": guess,
           "mask": np.ma.getmaskarray(ns["image_nanmask"]).copy(), "err": ns["err"].copy(),
           "satlevel": ns["satlevel"], "readnoise": ns["readnoise"], "sigma": ns["sigma"],
           "widths": np.asarray(ns["widths"], dtype=np.float64),
           "norm": np.array(ns["norm"]), "lognorm": np.array(ns["lognorm"]),
           "p_init": p_init, "chi_init": chi_init, "truth_local": None}

    # --- pointwise vectors: model images + chi-square for assorted parameter vectors ----------
    rng = np.random.default_rng(1000 + nbody)
    npar = len(p_init) - 1
    tl = truth.copy()
    for o in range(nbody):
        tl[2 * o] -= ox
        tl[2 * o + 1] -= oy
    out["truth_local"] = tl
    vecs = [tl, p_init[:npar].copy()]
    for _ in range(6):
        q = tl.copy()
        q[:2 * nbody] += rng.normal(0, 1.5, 2 * nbody)
        q[2 * nbody:2 * nbody + 2] += rng.normal(0, 0.3, 2)
        q[2 * nbody + 2:3 * nbody + 2] *= 10 ** rng.normal(0, 0.1, nbody)
        q[3 * nbody + 2] = rng.uniform(0.05, 0.6)
        q[3 * nbody + 3] *= 10 ** rng.normal(0, 0.2)
        q[3 * nbody + 4:3 * nbody + 8] *= 10 ** rng.normal(0, 0.08, 4)
        q[3 * nbody + 8:3 * nbody + 10] = rng.uniform(-1.6, 1.6, 2)
        vecs.append(q)
    vecs = np.array(vecs)
    models_, chis = [], []
    for q in vecs:
        m = ns["build_analytical_model"](np.concatenate([q, [0.0]]))
        models_.append(np.asarray(m))
        chis.append(float(ns["chi_squared"](ns["image_nanmask"], m, ns["err"])))
    out["vec_params"] = vecs
    out["vec_models"] = np.array(models_)
    out["vec_chi2"] = np.array(chis)

    # --- proposals and accept rule with the reference's global numpy stream -----------------
    np.random.seed(seed)
    out["prop_in"] = np.array([[12.5, 0.3], [0.2, 0.001], [3000.0, 0.02], [0.0, 0.02], [-4.0, 0.02]])
    with np.errstate(all="ignore"):
        out["prop_normal"] = np.array([ns["proposal"](a, b, 1)[0] for a, b in out["prop_in"]])
        out["prop_log"] = np.array([ns["logproposal"](a, b, 1)[0] for a, b in out["prop_in"]])
    acc_in = np.array([[100.0, 99.0], [100.0, 100.5], [100.0, 104.0], [100.0, 160.0],
                       [100.0, np.nan], [100.0, -1e6], [100.0, np.inf], [5000.0, 5000.0]])
    acc = []
    with np.errstate(all="ignore"):
        for a, b in acc_in:
            yes, pa, dice = ns["accept_reject"](a, b)
            acc.append([1.0 if yes == "yes" else 0.0, pa, dice])
    out["accept_in"] = acc_in
    out["accept_out"] = np.array(acc)

    # --- the sampler loop body, executed update by update -----------------------------------
    body = compile(ref_lines(v["path"], *v["body"], dedent=True), "<reference loop body>", "exec")
    np.random.seed(seed + 1)
    ns["count"] = 0
    trace = []
    with np.errstate(all="ignore"):
        for _ in range(n_trace):
            exec(body, ns)
            trace.append(ns["parameters"].copy())
    out["loop_seed"] = seed + 1
    out["loop_trace"] = np.array(trace)
    out["loop_tries"] = ns["total_tries"].copy()
    out["loop_accepts"] = ns["total_accept"].copy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for nbody in (2, 3):
        res = run_variant(nbody, size=32, n_trace=400, seed=20190531)
        path = os.path.join(OUT, "reference_exec_%dbody.npz" % nbody)
        np.savez_compressed(path, **res)
        print(path, os.path.getsize(path), "bytes; chi_init", res["chi_init"],
              "accepted", int(res["loop_accepts"].sum()), "of", int(res["loop_tries"].sum()))


if __name__ == "__main__":
    main()
