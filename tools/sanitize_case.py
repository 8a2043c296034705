#!/usr/bin/env python
"""A tiny workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from olpefit_b200 import frame, sampler, synth
HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
for nbody, size, team in ((2, 32, 1), (2, 64, 1), (2, 32, 4), (3, 64, 16)):
    stamps, origins = synth.make_stamps(2, size, nbody)
    dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=nbody)
    p0 = frame.initial_parameters(stamps[0], synth.step1_guess(stamps[0], nbody, origin=tuple(origins[0])), nbody,
                                  origin=tuple(origins[0]))
    w = 5
    fo = (np.arange(w) % 2).astype(np.int32)
    m, c = dom.model_chi2(np.tile(p0, (w, 1)), frame_of=fo, want_model=True)
    with sampler.GibbsSampler(dom, np.tile(p0, (w, 1)), fo, seed=1, thin=3, team_warps=team) as s:
        ch = s.run(40)
        st = s.stats()
        print(nbody, size, team, float(c[0]), ch.shape, int(st["tries"].sum()), int(st["exps"]))
img, _ = synth.make_frame(0, 2, region=(480, 560, 470, 590))
dom = frame.prepare_domain(img, HEADER, origin=(470, 480), nbody=2)          # generic kernel
print(float(dom.model_chi2(synth.truth_parameters(2, 0)[None])[1][0]))
