#!/usr/bin/env python
"""Run bench.py over stamp sizes / models / batch shapes and collect one JSON list (profiles/)."""
import json, subprocess, sys
out = []
cases = [("--stamp %d --nbody %d" % (s, nb)) for nb in (2, 3) for s in (32, 64, 128)]
cases += ["--walkers 64 --frames 1 --team 16 --updates-per-step 4096 --stamp 64",
          "--walkers 64 --frames 1 --team 16 --updates-per-step 2048 --stamp 128",
          "--walkers 1 --frames 1 --team 16 --updates-per-step 16000 --stamp 64",
          "--walkers 1048576 --frames 1000 --stamp 32 --updates-per-step 32 --thin 32"]
for c in cases:
    cmd = [sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--no-latency"] + c.split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print("FAILED", c, r.stderr[-400:]); continue
    row = {"args": c, "pixel_evals_per_s": d["value"], "updates_per_s": d["gibbs_updates_per_sec"],
           "ms_per_step": d["ms_per_step"], "fp32_frac_algorithmic": d["roofline"]["algorithmic"]["frac"],
           "fp32_frac_executed": d["roofline"]["frac"], "sfu_frac_algorithmic": d["roofline"]["sfu"]["frac"],
           "cycles_per_update_per_scheduler": d["roofline"]["executed"]["cycles_per_update_per_scheduler"],
           "component_evals_per_pixel_eval": d["roofline"]["executed"]["component_evals_per_pixel_eval"],
           "clocks": d["clocks"]}
    out.append(row)
    print("%-75s px/s %.3e upd/s %.3e fp32 alg %.3f exec %.3f sfu alg %.3f" % (c, row["pixel_evals_per_s"], row["updates_per_s"],
          row["fp32_frac_algorithmic"], row["fp32_frac_executed"], row["sfu_frac_algorithmic"]), flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.json", "w"), indent=1)
