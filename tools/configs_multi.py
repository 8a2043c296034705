#!/usr/bin/env python
"""BASELINE.json configs 3, 4 and 5 on N GPUs through torchrun, one JSON list (profiles/).
Config 5 (1,048,576 walkers over 1,000 epochs) is sharded the way the command line shards it: walker
w of epoch f is g = w F + f and rank r owns g = r (mod N), i.e. whole epochs f = r (mod N) -- 1000/N
epochs with all their 1,048 walkers per GPU; "--by-walker" runs the other cut (every GPU sees all
1,000 epochs with 1/N of their walkers) for comparison.
usage: tools/configs_multi.py N out.json [--by-walker]"""
import json, subprocess, sys

n = int(sys.argv[1])
fr = 1000 if "--by-walker" in sys.argv else 1000 // n
cut = "every GPU all 1,000 epochs" if "--by-walker" in sys.argv else "%d whole epochs per GPU" % fr
cases = [("config3: 2-body, 65,536 walkers/GPU x 100 epochs, 64x64 (the bench default, with e2e)", "--steps 20 --warmup 3 --no-cpu-baseline --no-latency"),
         ("config4: 3-body, 65,536 walkers/GPU x 100 epochs, 64x64", "--nbody 3 --steps 10 --warmup 3 --no-cpu-baseline --no-latency"),
         ("config5: 1,048,576 walkers x 1,000 frames, 32x32, " + cut, "--walkers %d --frames %d --stamp 32 --updates-per-step 64 --thin 32 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-latency" % (1048576 // n, fr)),
         ("config5: 1,048,576 walkers x 1,000 frames, 64x64, " + cut, "--walkers %d --frames %d --stamp 64 --updates-per-step 64 --thin 32 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-latency" % (1048576 // n, fr)),
         ("config5: 1,048,576 walkers x 1,000 frames, 128x128, " + cut, "--walkers %d --frames %d --stamp 128 --updates-per-step 64 --thin 32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-latency" % (1048576 // n, fr))]
if "--only5" in sys.argv:
    cases = cases[2:]
out = []
for port, (name, args) in enumerate(cases):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + port), "bench.py", "--gpus", str(n)] + args.split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1])
    except Exception:
        print("FAILED", name, r.stderr[-600:], flush=True)
        continue
    row = {"workload": name, "args": args, "n_gpus": d["n_gpus"], "pixel_evals_per_s": d["value"],
           "updates_per_s": d["gibbs_updates_per_sec"], "ms_per_step": d["ms_per_step"],
           "fp32_frac_executed": d["roofline"]["frac"], "fp32_frac_algorithmic": d["roofline"]["algorithmic"]["frac"],
           "e2e": d["e2e"], "clocks": d["clocks"]}
    out.append(row)
    print("%-70s px/s %.3e upd/s %.3e ms/step %.2f exec %.3f e2e %s" % (name, row["pixel_evals_per_s"], row["updates_per_s"],
          row["ms_per_step"], row["fp32_frac_executed"], ("%.3e" % d["e2e"]["value"]) if d["e2e"] else "-"), flush=True)
json.dump(out, open(sys.argv[2], "w"), indent=1)
