#!/usr/bin/env python
"""Experiment: model images and chi-squares of the stateless operator for fixed vectors, written
to an .npz so builds of the library (LAPF_LIB) can be compared offline and against the oracle.

    LAPF_LIB=build/exp/liblapf_split.so python tests/diag/exp_k1dump.py out.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from olpefit_b200 import frame, synth  # noqa: E402
from oracle import lapf_oracle as orc  # noqa: E402  (experiment script: the oracle is the checker here)

HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}

out = {}
for nbody, size in ((2, 32), (3, 32), (2, 64)):
    stamps, origins = synth.make_stamps(1, size, nbody)
    dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=nbody)
    rng = np.random.default_rng(5)
    p = np.tile(synth.truth_parameters(nbody, 0), (6, 1))
    p[1:, :2 * nbody] += rng.normal(0, 0.5, (5, 2 * nbody))
    p[1:, 3 * nbody + 8:] += rng.normal(0, 0.3, (5, 2))
    p = p.astype(np.float32).astype(np.float64)
    m, c = dom.model_chi2(p, want_model=True)
    lay = orc.layout_for(nbody)
    img = stamps[0].astype(np.float64)
    w = orc.weight_map(img, HEADER)
    ref = np.array([orc.model_image(q, lay, size, size, origin=tuple(origins[0])) for q in p])
    chi_ref = np.array([orc.chi_squared_weighted(img, r, w) for r in ref])
    mm = m.cpu().numpy()
    rel = np.abs(mm - ref) / np.abs(ref)
    print("%d-body %d px: worst per-pixel rel error vs oracle %.3e, chi2 rel %.3e" %
          (nbody, size, rel.max(), np.max(np.abs(c.cpu().numpy() - chi_ref) / chi_ref)), flush=True)
    if rel.max() > 1e-5:
        v, y, x = np.unravel_index(np.argmax(rel), rel.shape)
        bad = np.argwhere(rel[v] > 1e-5)
        print("   worst at vector %d pixel (row %d, col %d); %d bad pixels in that image, rows %s cols %s"
              % (v, y, x, len(bad), sorted(set(bad[:, 0]))[:20], sorted(set(bad[:, 1]))[:20]))
    out["m_%d_%d" % (nbody, size)] = mm
    out["c_%d_%d" % (nbody, size)] = c.cpu().numpy()
np.savez_compressed(sys.argv[1], **out)
