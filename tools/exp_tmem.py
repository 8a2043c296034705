#!/usr/bin/env python
"""Experiment harness for the TMEM / table-split anomaly: run the batched sampler of
the library named by LAPF_LIB (default: the shipped one) on every shape and compare the chi-square
each walker carries with the stateless operator K1 applied to its parameters (must be bitwise
equal), and report the acceptance rate (the failing build rejected every proposal).

    LAPF_LIB=build/exp/liblapf_split.so python tools/exp_tmem.py [nbody:size ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from olpefit_b200 import frame, sampler, synth  # noqa: E402

HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}


def one(nbody, size, walkers=1200, n_frames=3, n_upd=32):
    stamps, origins = synth.make_stamps(n_frames, size, nbody)
    dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=nbody)
    frame_of = (np.arange(walkers) % n_frames).astype(np.int32)
    p = np.array([synth.truth_parameters(nbody, f) for f in range(n_frames)])[frame_of]
    P = p.shape[1]
    with sampler.GibbsSampler(dom, p, frame_of, seed=31, burn_in=0, thin=8) as s:
        s.run(n_upd, record=False)
        st, tries, acc = s.state()
    _, chi = dom.model_chi2(st[:, :P], frame_of=frame_of)
    same = int((chi == st[:, P]).sum())
    rel = float(((chi - st[:, P]).abs() / chi.abs()).max())
    rate = float(acc.sum()) / float(tries.sum())
    ok = same == walkers
    print("%d-body %3d px: K1==K2 bitwise for %d / %d walkers (max rel diff %.2e), acceptance %.3f  %s"
          % (nbody, size, same, walkers, rel, rate, "OK" if ok else "MISMATCH"), flush=True)
    return ok


if __name__ == "__main__":
    shapes = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(2, 32), (3, 32), (2, 64), (3, 64), (2, 128), (3, 128)]
    print("library:", os.environ.get("LAPF_LIB", "(shipped)"))
    good = all([one(nb, sz) for nb, sz in shapes])
    sys.exit(0 if good else 1)
