"""Drop-in command lines for step 2 of LAPF, driving the CUDA sampler.

    apf_step2.py IMAGE [-i {1,2a}]        apf_step2.py:154-158   (mpiexec -n N  ->  --walkers N)
    apf_step2a.py IMAGE                   apf_step2a.py:146-148
    3body/apf_step2_3body.py IMAGE        3body/apf_step2_3body.py:147-149

Same positional argument and -i flag, same step-1 input (<dir>/<N>_initialguess), same outputs
(<dir>/<N>_apf_results/<w>_finalarray_mpi.csv, <w>_acceptance_rate.csv, step2a.csv,
step2a_acceptance_rate), same defaults (accept_min = 100000; burn_in = 6000 / 0 for 3-body and
2a; n_steps = 5000).  New flags are additive.  Under torchrun the walkers are sharded over the
GPUs; every rank writes the files of its own walkers, so nothing but the stop rule crosses ranks.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np


def _parser(kind):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("image", type=str)
    if kind == "step2":
        ap.add_argument("-i", "--initial_guess_option", type=str,
                        help="Choose which file to take in as initial guess.  Enter -i 1 for step 1, "
                             "-i 2a for step 2a.")
    if kind != "step2a":
        ap.add_argument("--walkers", type=int, default=24,
                        help="independent walkers (replaces the process count of mpiexec -n)")
        ap.add_argument("--accept-min", type=int, default=100000,
                        help="stop when every parameter of every walker has been tried this often")
    else:
        ap.add_argument("--n-steps", type=int, default=5000)
    ap.add_argument("--burn-in", type=int, default=None)
    ap.add_argument("--stamp", type=int, default=128, choices=[32, 64, 128],
                    help="side of the square cut-out, centred on the objects, in which the Gaussians are evaluated")
    ap.add_argument("--domain", choices=["frame", "stamp"], default="frame",
                    help="frame: chi-square over the whole image like the reference (pixels outside the stamp enter "
                         "through their exact sums against the constant floor); stamp: the cut-out only")
    ap.add_argument("--thin", type=int, default=1, help="record every thin-th update (1 = reference)")
    ap.add_argument("--seed", type=int, default=None, help="random seed (default: from the OS, like the reference's unseeded numpy)")
    ap.add_argument("--segment", type=int, default=4096, help="updates per kernel launch")
    ap.add_argument("--team", type=int, default=0, choices=[0, 1, 4, 16],
                    help="warps cooperating on one walker (0: pick from the total walker count)")
    ap.add_argument("--format", choices=["csv", "bin"], default="csv", help="per-walker reference CSV files, or one packed binary")
    ap.add_argument("--fix-bkgd", action="store_true",
                    help="2-body only: use bkgd (p[9]) as the constant floor instead of the reference's p[12]")
    ap.add_argument("--adapt", action="store_true",
                    help="tune the jump widths from the acceptance rates DURING BURN-IN ONLY (the reference's widths "
                         "are fixed; off by default so posteriors compare like for like)")
    ap.add_argument("--checkpoint", action="store_true",
                    help="write <results>/checkpoint_rank<r>.pt when done (exact state of every walker)")
    ap.add_argument("--resume", action="store_true",
                    help="continue from <results>/checkpoint_rank<r>.pt, appending to the existing chain files")
    ap.add_argument("--quiet", action="store_true")
    return ap


def _stamp_origin(params, nbody, size, shape):
    xs, ys = params[0:2 * nbody:2], params[1:2 * nbody:2]
    ox = int(round(float(np.mean(xs)))) - size // 2
    oy = int(round(float(np.mean(ys)))) - size // 2
    ox = min(max(ox, 0), max(shape[1] - size, 0))
    oy = min(max(oy, 0), max(shape[0] - size, 0))
    return ox, oy


def run(kind, nbody, argv=None):
    args = _parser(kind).parse_args(argv)
    import torch
    from . import chains, dist, frame, layout, sampler as smp

    rank, local_rank, world = dist.init()
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this implementation has no CPU path")
    torch.cuda.set_device(local_rank)
    say = (lambda *a: None) if (args.quiet or rank != 0) else (lambda *a: print(*a, flush=True))

    image, hdr = frame.read_fits(args.image)
    image = np.asarray(image, dtype=np.float64) if image.dtype.kind != "f" else image
    outdir = chains.results_dir(args.image)
    say(outdir)
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)                    # apf_step2.py:172-173
    dist.barrier()

    satlevel = frame.saturation_level(hdr)
    say("Max pixel value in image:", float(np.max(image)))
    say("Masking pixels greater than ", 0.8 * satlevel)

    P = layout.nparam(nbody)
    if kind == "step2" and getattr(args, "initial_guess_option", None) == "2a":   # apf_step2.py:248-256
        say("I am taking the initial guess from Step 2a output")
        a = np.genfromtxt(outdir + "step2a.csv", delimiter=",")
        parameters = np.array(a[-1][:P], dtype=np.float64)
    else:                                                                          # :258-273
        say("I am taking the initial guess from Step 1 output")
        guess = np.loadtxt(open(chains.initial_guess_path(args.image), "rb"), delimiter=" ")
        parameters = frame.initial_parameters(image, guess, nbody)

    if args.burn_in is not None:
        burn_in = args.burn_in
    else:
        burn_in = 6000 if (kind == "step2" and nbody == 2) else 0    # apf_step2.py:38; 3body :38; 2a: SURVEY C3
    floor_index = layout.bkgd_index(nbody) if (args.fix_bkgd and nbody == 2) else layout.REFERENCE_FLOOR_INDEX
    ox, oy = _stamp_origin(parameters, nbody, args.stamp, image.shape)
    dom = frame.prepare_domain(image.astype(np.float32), hdr, size=args.stamp, cut=(ox, oy), nbody=nbody,
                               floor_index=floor_index, device="cuda:%d" % local_rank,
                               whole_frame=(args.domain == "frame"))
    say("I have masked", int((dom.weight == 0).sum().item()), "pixels (inside the %dx%d domain at x0=%d y0=%d)"
        % (args.stamp, args.stamp, ox, oy))

    total_walkers = 1 if kind == "step2a" else args.walkers
    id_base, id_stride, n_local = dist.shard_ids(total_walkers, rank, world)
    seed = args.seed if args.seed is not None else int.from_bytes(os.urandom(8), "little")
    if world > 1:   # every rank must use rank 0's seed
        t = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda:%d" % local_rank)
        torch.distributed.broadcast(t, 0)
        seed = int(t.item())
    say("Random seed:", seed)
    if n_local == 0:
        say("rank has no walkers")
    # Few walkers: several warps per walker, so one update takes a fraction of the time; from about
    # 1,800 walkers per GPU the batched kernel (a warp owns 32 walkers) is faster (measured on B200:
    # 64 walkers 3.5e7 / 2.9e7 / 1.4e7 updates/s with 16 / 4 / 1 warps per walker, 592 walkers
    # 0.8e8 / 2.6e8 / 1.3e8, 2,048 walkers 0.8e8 / 2.2e8 / 2.4e8).  The choice depends on the TOTAL
    # walker count only, so every rank makes the same one.
    team = args.team
    if team == 0:
        team = 16 if (total_walkers <= 220 * world and args.stamp >= 64) else (4 if total_walkers <= 1800 * world else 1)
    sam = smp.GibbsSampler(dom, np.tile(parameters, (max(n_local, 1), 1)), seed=seed, burn_in=burn_in,
                           thin=args.thin, id_base=id_base, id_stride=id_stride, team_warps=team)
    st0, _, _ = sam.state()
    say("Found initial chi-squared:", float(st0[0, -1].item()))
    say("Initial guess:", np.concatenate([parameters, [float(st0[0, -1].item())]]))
    say()
    say("Beginning loop...")

    ids = [id_base + i * id_stride for i in range(n_local)]
    if kind == "step2a":
        paths = [outdir + "step2a.csv"]
        acc_paths = [outdir + "step2a_acceptance_rate"]
    else:
        paths = [outdir + "%d_finalarray_mpi.csv" % g for g in ids]
        acc_paths = [outdir + "%d_acceptance_rate.csv" % g for g in ids]
    packed = None
    if args.format == "bin":
        packed = chains.PackedChainWriter(outdir + "chains_rank%d" % rank, max(n_local, 1), P + 1,
                                          {"walker_ids": ids, "seed": seed, "burn_in": burn_in, "thin": args.thin,
                                           "nbody": nbody, "origin": [ox, oy], "stamp": args.stamp})

    streamer = smp.ChainStreamer(sam, args.segment)
    first = [True]
    ckpt_path = outdir + "checkpoint_rank%d.pt" % rank
    if args.resume:
        ck = torch.load(ckpt_path)
        if ck["n_walkers"] != sam.n_walkers or ck["nbody"] != nbody:
            raise SystemExit("checkpoint %s does not match this run" % ckpt_path)
        sam.load(ck["blob"])
        seed = ck["seed"]
        first[0] = False                      # append to the chain files that are already there
        say("Resumed at update", sam.count)
    widths, _is_log = layout.default_widths(nbody)

    def consume(seg):
        if seg is None or n_local == 0:
            return
        if packed is not None:
            packed.append(seg)
        else:
            chains.write_segment_csv(paths, seg, first[0])
        first[0] = False

    t_start = time.time()
    if kind == "step2a":
        left = args.n_steps                                   # apf_step2a.py:271
        while left > 0:
            n = min(args.segment, left)
            consume(streamer.run(n))
            left -= n
    else:
        while True:                                           # apf_step2.py:300
            stt = sam.stats(moments=False)
            tries, accepts, mn = dist.allreduce_stats(stt["tries"], stt["accepts"], stt["min_tries"])
            gap = args.accept_min - int(mn.item())
            if gap <= 0:
                break
            # min(tries) grows by at most one per update: `gap` updates can never overshoot
            n = min(args.segment, gap)
            if args.adapt and sam.count < burn_in:
                # burn-in only: nudge each width towards ~35% acceptance, then leave it alone
                n = min(n, max(burn_in - sam.count, 1), 512)
                if sam.count > 0:
                    rate = (accepts.double() / tries.double().clamp(min=1)).cpu().numpy()
                    widths = widths * np.clip(np.exp(rate - 0.35), 0.5, 2.0)
                    sam.set_widths(widths)
            consume(streamer.run(n))
            if sam.count % (10 * args.segment) < n:
                say("Loop count:", sam.count, " min tries:", int(mn.item()),
                    " acceptance rate:", (accepts.double() / tries.double().clamp(min=1)).cpu().numpy())
    consume(streamer.finish())
    if packed is None and first[0] and n_local:
        for pth in paths:                                     # nothing recorded: the nan row alone
            chains.write_walker_csv(pth, np.zeros((0, P + 1)))
    if packed is not None:
        packed.close(count=sam.count)
    if args.adapt:
        say("Jump widths after burn-in tuning:", widths)
    if args.checkpoint and n_local:
        torch.save({"blob": sam.save().cpu(), "n_walkers": sam.n_walkers, "nbody": nbody, "seed": seed,
                    "count": sam.count}, ckpt_path)

    # Gelman-Rubin statistic of the recorded rows (apf_step3.py:260-278) from device-side moments,
    # summed over ranks: a convergence read-out without re-reading any chain file
    if kind != "step2a" and total_walkers > 1:
        stt = sam.stats(moments=True)
        mom = dist.allreduce_sum(stt["moments"] if n_local else torch.zeros_like(stt["moments"]))
        n_rows = int(stt["rows"].item())
        if n_rows > 1:
            _, rc = chains.gelman_rubin_from_moments(mom[0, :P].cpu().numpy(), n_rows, total_walkers)
            say("Gelman-Rubin stat per parameter (recorded rows, all walkers):", np.round(rc, 4))
            say("GR for positions:", *np.round(rc[:2 * nbody], 4))
    _, tr, ac = sam.state()
    tr, ac = tr.cpu().numpy(), ac.cpu().numpy()
    for i in range(n_local):
        chains.write_acceptance(acc_paths[i], ac[i], tr[i])
    dist.barrier()
    say("Rank ", rank, "done with loop: %d updates per walker, %d walkers, %.1f s"
        % (sam.count, total_walkers, time.time() - t_start))
    sam.close()
    return 0


def main_step2(argv=None):
    return run("step2", 2, argv)


def main_step2a(argv=None):
    return run("step2a", 2, argv)


def main_step2_3body(argv=None):
    return run("step2_3body", 3, argv)


if __name__ == "__main__":
    sys.exit(main_step2())
