#!/usr/bin/env python
"""Many epochs through the command line (BASELINE.json configs[2] as a user would run it): writes N
synthetic NIRC2-like FITS frames with their step-2a starting points, runs
`apf_step2.py EPOCH0.fits -i 2a --frames LIST --walkers W --no-chains` (statistics reduced on the
device, nothing but step2_summary.json leaves it) and compares each epoch's separation / position
angle with the truth the frame was drawn from.

    python tools/demo_epochs.py [--epochs 100] [--walkers 656] [--stamp 64] [--accept-min 1500] [--gpus 1]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from olpefit_b200 import chains, frame, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=100)
ap.add_argument("--walkers", type=int, default=128)
ap.add_argument("--stamp", type=int, default=32)
ap.add_argument("--accept-min", type=int, default=400)
ap.add_argument("--burn-in", type=int, default=3000)
ap.add_argument("--gpus", type=int, default=1)
a = ap.parse_args()

tmp = tempfile.mkdtemp(prefix="lapf_epochs_")
paths, truths = [], []
t0 = time.time()
for f in range(a.epochs):
    d = os.path.join(tmp, "epoch%04d" % f)
    os.makedirs(d)
    img, truth = synth.make_frame(f, 2)
    p = os.path.join(d, "N2.20090531.%05d.LDIF.fits" % (30000 + f))
    frame.write_fits(p, img, synth.HEADER)
    os.makedirs(chains.results_dir(p))
    start = truth.copy()
    start[:4] = np.round(start[:4] * 2) / 2 + 0.1          # half-pixel clicks, not the truth
    chains.write_walker_csv(chains.results_dir(p) + "step2a.csv", np.concatenate([start, [0.0]])[None])
    paths.append(p)
    truths.append(truth)
lst = os.path.join(tmp, "frames.txt")
open(lst, "w").write("\n".join(paths[1:]) + "\n")
print("wrote %d frames in %.1f s" % (a.epochs, time.time() - t0), flush=True)

args = [paths[0], "-i", "2a", "--frames", lst, "--walkers", str(a.walkers), "--accept-min", str(a.accept_min),
        "--burn-in", str(a.burn_in), "--thin", "16", "--stamp", str(a.stamp), "--seed", "7", "--no-chains"]
cmd = [sys.executable, os.path.join(ROOT, "apf_step2.py")] + args
if a.gpus > 1:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus), "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "apf_step2.py")] + args
t0 = time.time()
res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
wall = time.time() - t0
print(res.stdout[-3000:])
if res.returncode:
    print(res.stderr[-3000:])
    sys.exit(1)
print("command line took %.1f s for %d epochs x %d walkers" % (wall, a.epochs, a.walkers))
print("epoch   truth sep [mas]  median  (lo, hi)            truth pa [deg]  median   (lo, hi)         max GR(pos)  outside")
worst = 0.0
for f in list(range(min(a.epochs, 6))) + ([a.epochs - 1] if a.epochs > 6 else []):
    s = json.load(open(chains.results_dir(paths[f]) + "step2_summary.json"))
    c = s["sep_pa_companion"]
    t_sep, t_pa = chains.separation_pa(*truths[f][:4])
    print("%5d   %10.3f  %9.3f  (%.3f, %.3f)   %10.4f  %9.4f  (%.4f, %.4f)   %8.4f   %.3f" % (
        f, t_sep, c["sep_mas"]["median"], c["sep_mas"]["lo"], c["sep_mas"]["hi"], t_pa, c["pa_deg"]["median"],
        c["pa_deg"]["lo"], c["pa_deg"]["hi"], max(s["gelman_rubin"][:4]), c["outside_histogram"]))
pulls = []
for f in range(a.epochs):
    s = json.load(open(chains.results_dir(paths[f]) + "step2_summary.json"))["sep_pa_companion"]
    t_sep, t_pa = chains.separation_pa(*truths[f][:4])
    pulls.append(((s["sep_mas"]["median"] - t_sep) / max(s["sep_mas"]["std"], 1e-9), (s["pa_deg"]["median"] - t_pa) / max(s["pa_deg"]["std"], 1e-9)))
pulls = np.array(pulls)
print("pulls (median - truth) / std over %d epochs: sep mean %.2f rms %.2f, pa mean %.2f rms %.2f"
      % (a.epochs, pulls[:, 0].mean(), pulls[:, 0].std(), pulls[:, 1].mean(), pulls[:, 1].std()))
