"""Diagnosis: cost of an update by epoch of the synthetic benchmark (some epochs are ~7 % slower per walker)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from olpefit_b200 import frame, sampler, synth
HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
F, S, W, U = 100, 64, 65536, 128
stamps, origins = synth.make_stamps(F, S, 2)
dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=2)
p = []
for f in range(F):
    g = synth.step1_guess(stamps[f], 2, origin=tuple(origins[f]))
    p.append(frame.initial_parameters(stamps[f], g, 2, origin=tuple(origins[f])))
p = np.asarray(p)
print("epoch  ms/128 updates  comp-evals/update/pixel  acceptance   start: sx sy sx2 sy2 amp_s amp_c ratio")
from olpefit_b200 import model
plain = model.PixelDomain(dom.data, dom.weight, dom.origin, nbody=2, plain_loop=True)
nocull = model.PixelDomain(dom.data, dom.weight, dom.origin, nbody=2, cull=False) if "cull" in model.PixelDomain.__init__.__code__.co_varnames else None
variants = [("default", dom), ("plain loop", plain)] + ([("no culling", nocull)] if nocull is not None else [])
for name, d_ in variants:
  print(name)
  for f in (34, 50):
    fo = np.full(W, f, dtype=np.int32)
    with sampler.GibbsSampler(d_, p[fo], fo, seed=1) as s:
        s.run(U)
        e0 = int(s.stats(moments=False)["exps"].item())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s.run(U); b.record(); torch.cuda.synchronize()
        st = s.stats(moments=False)
        e1 = int(st["exps"].item())
        acc = float(st["accepts"].sum()) / float(st["tries"].sum())
    print("%3d   %8.3f   %8.3f   %.3f" % (f, a.elapsed_time(b), (e1 - e0) / (W * U) / (S * S), acc))
