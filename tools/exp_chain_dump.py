#!/usr/bin/env python
"""Experiment: dump the chain of a small batched run so two builds of the library (LAPF_LIB) can be
compared row by row offline.   usage: exp_chain_dump.py out.npz [nbody size walkers]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from olpefit_b200 import frame, sampler, synth  # noqa: E402

HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
nbody, size, walkers = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (2, 32, 1200)
n_frames, n_upd = 3, 32
stamps, origins = synth.make_stamps(n_frames, size, nbody)
dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=nbody)
frame_of = (np.arange(walkers) % n_frames).astype(np.int32)
p = np.array([synth.truth_parameters(nbody, f) for f in range(n_frames)])[frame_of]
with sampler.GibbsSampler(dom, p, frame_of, seed=31, burn_in=0, thin=1) as s:
    chain = s.run(n_upd).cpu().numpy()
    st, tries, acc = (t.cpu().numpy() for t in s.state())
print("acceptance", acc.sum() / tries.sum())
np.savez_compressed(sys.argv[1], chain=chain, tries=tries, acc=acc, frame_of=frame_of)
