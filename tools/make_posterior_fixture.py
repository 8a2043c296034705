#!/usr/bin/env python
"""Reference posterior for the statistical parity test (tests/test_gpu_posterior.py).

Runs the numpy oracle (the float64 restatement of apf_step2.py:300-351, numpy Mersenne-Twister
stream like the reference) as N independent CPU walkers on one synthetic 32x32 stamp and stores
what step 3 derives from such chains -- separation and position angle (apf_step3.py:255-256,
283-291) -- as pooled quantiles plus per-walker summaries (for Monte-Carlo standard errors).

    python tools/make_posterior_fixture.py [--walkers 64] [--updates 60000] [--burn 20000]
"""
import argparse
import concurrent.futures as cf
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
SIZE, EPOCH = 32, 0
NBODY = int(os.environ.get("LAPF_FIXTURE_NBODY", "2"))


def setup():
    from olpefit_b200 import synth
    from oracle import lapf_oracle as orc
    lay = orc.layout_for(NBODY)
    ox, oy = synth.stamp_origin(SIZE, NBODY)
    img32, truth = synth.make_frame(EPOCH, NBODY, region=(oy, oy + SIZE, ox, ox + SIZE))
    img = img32.astype(np.float64)
    w = orc.weight_map(img, HEADER)
    guess = synth.step1_guess(img32, NBODY, origin=(ox, oy))
    g_local = guess - np.array([ox, oy] * NBODY + [ox, oy], dtype=np.float64)
    # Start at the in-model truth.  From the raw step-1 guess the reference algorithm itself can
    # walk the companion component onto the (initially badly fitted) star; that is a property of
    # the sampler, not something a CPU/GPU comparison should depend on.
    p0 = truth.copy()
    return orc, lay, img, w, (ox, oy), p0, truth


def walker(job):
    seed, n_updates, burn, thin = job
    orc, lay, img, w, origin, p0, _ = setup()
    res = orc.run_chain(img, w, lay, p0, orc.NumpyStream(seed), origin=origin, n_updates=n_updates,
                        burn_in=burn, thin=thin)
    return res.rows[1:], res.tries, res.accepts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=64)
    ap.add_argument("--updates", type=int, default=50000)
    ap.add_argument("--burn", type=int, default=10000)
    ap.add_argument("--thin", type=int, default=10)
    a = ap.parse_args()
    orc, lay, img, w, origin, p0, truth = setup()
    t0 = time.time()
    jobs = [(5000 + i, a.updates, a.burn, a.thin) for i in range(a.walkers)]
    with cf.ProcessPoolExecutor(max_workers=os.cpu_count(), mp_context=mp.get_context("spawn")) as ex:
        out = list(ex.map(walker, jobs))
    rows = np.array([o[0] for o in out])                     # [walkers, rows, P+1]
    tries = np.array([o[1] for o in out])
    accepts = np.array([o[2] for o in out])
    sep, pa = orc.separation_pa(rows[..., 0], rows[..., 1], rows[..., 2], rows[..., 3])   # star -> first companion
    qs = [15.865, 50.0, 84.135]
    res = {
        "size": SIZE, "nbody": NBODY, "epoch": EPOCH, "p0": p0, "truth": truth, "origin": np.array(origin),
        "updates": a.updates, "burn": a.burn, "thin": a.thin, "walkers": a.walkers,
        "sep_q": np.percentile(sep, qs), "pa_q": np.percentile(pa, qs),
        "sep_q_walker": np.percentile(sep, qs, axis=1).T, "pa_q_walker": np.percentile(pa, qs, axis=1).T,
        "sep_mean_walker": sep.mean(axis=1), "pa_mean_walker": pa.mean(axis=1),
        "param_mean": rows.reshape(-1, rows.shape[-1]).mean(axis=0),
        "param_std": rows.reshape(-1, rows.shape[-1]).std(axis=0),
        "param_mean_walker": rows.mean(axis=1), "param_var_walker": rows.var(axis=1),
        "acceptance": accepts.sum(axis=0) / tries.sum(axis=0),
        "chi2_first_rows": rows[:, :5, -1], "chi2_last_rows": rows[:, -5:, -1],
    }
    path = os.path.join(ROOT, "tests", "golden", "posterior_%dbody_s32.npz" % NBODY)
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes in %.0f s" % (time.time() - t0))
    print("sep quantiles", res["sep_q"], " SE(median) ~", res["sep_q_walker"][:, 1].std() / np.sqrt(a.walkers))
    print("pa quantiles", res["pa_q"], " SE(median) ~", res["pa_q_walker"][:, 1].std() / np.sqrt(a.walkers))
    print("acceptance", np.round(res["acceptance"], 3))
    print("chi2 at first recorded rows", res["chi2_first_rows"][:3], "last", res["chi2_last_rows"][:3])
    t_sep, t_pa = orc.separation_pa(truth[0], truth[1], truth[2], truth[3])
    print("truth sep, pa:", t_sep, t_pa)


if __name__ == "__main__":
    main()
