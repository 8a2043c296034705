"""Drop-in command lines for step 2 of LAPF, driving the CUDA sampler.

    apf_step2.py IMAGE [-i {1,2a}]        apf_step2.py:154-158   (mpiexec -n N  ->  --walkers N)
    apf_step2a.py IMAGE                   apf_step2a.py:146-148
    3body/apf_step2_3body.py IMAGE        3body/apf_step2_3body.py:147-149

Same positional argument and -i flag, same step-1 input (<dir>/<N>_initialguess), same outputs
(<dir>/<N>_apf_results/<w>_finalarray_mpi.csv, <w>_acceptance_rate.csv, step2a.csv,
step2a_acceptance_rate), same defaults (accept_min = 100000; burn_in = 6000 / 0 for 3-body and
2a; n_steps = 5000).  New flags are additive.  Under torchrun the walkers are sharded over the
GPUs; every rank writes the files of its own walkers, so nothing but the stop rule and the
summary statistics crosses ranks.

Many epochs in one run (additive): ``--frames LIST`` names a text file with one FITS path per line.
Every frame keeps the reference's layout -- its own <N>_initialguess, its own <N>_apf_results/
with ``--walkers`` chains -- but all of them advance in ONE batched sampler, and the per-epoch
separation / position angle (median, 68 % interval, mean, std; apf_step3.py:283-291,436-437) and
Gelman-Rubin statistics (:260-278) are reduced on the device and written to
<N>_apf_results/step2_summary.json, so ``--no-chains`` runs need no chain to leave the GPU.
"""
from __future__ import annotations

import argparse
import json
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
                        help="independent walkers per image (replaces the process count of mpiexec -n)")
        ap.add_argument("--accept-min", type=int, default=100000,
                        help="stop when every parameter of every walker has been tried this often")
    else:
        ap.add_argument("--n-steps", type=int, default=5000)
    ap.add_argument("--frames", type=str, default=None,
                    help="text file with one FITS path per line: more epochs for the same run (each with its own "
                         "<N>_initialguess and <N>_apf_results/); IMAGE is the first epoch")
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
    ap.add_argument("--chain-dtype", choices=["f64", "f32"], default="f64",
                    help="--format bin only: f32 stores float32 differences from each walker's starting point "
                         "(half the bytes off the device and on disk; chains.read_packed restores the values)")
    ap.add_argument("--no-chains", action="store_true",
                    help="record no chain rows at all: only the acceptance files and step2_summary.json (statistics "
                         "reduced on the device) are written")
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


def _frame_list(args):
    images = [args.image]
    if args.frames:
        base = os.path.dirname(os.path.abspath(args.frames))
        with open(args.frames) as fh:
            for ln in fh:
                ln = ln.strip()
                if not ln or ln.startswith("#"):
                    continue
                pth = ln if os.path.isabs(ln) else os.path.join(base, ln)
                if os.path.abspath(pth) not in [os.path.abspath(p) for p in images]:
                    images.append(pth)
    return images


def run(kind, nbody, argv=None):
    args = _parser(kind).parse_args(argv)
    import torch
    from . import chains, dist, frame, layout, sampler as smp, stats

    rank, local_rank, world = dist.init()
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this implementation has no CPU path")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    say = (lambda *a: None) if (args.quiet or rank != 0) else (lambda *a: print(*a, flush=True))

    images = _frame_list(args)
    n_frames = len(images)
    outdirs = [chains.results_dir(p) for p in images]
    say(outdirs[0] if n_frames == 1 else "%d epochs, results in %s ... %s" % (n_frames, outdirs[0], outdirs[-1]))
    if rank == 0:
        for d in outdirs:
            os.makedirs(d, exist_ok=True)                     # apf_step2.py:172-173
    dist.barrier()

    P = layout.nparam(nbody)
    from_2a = kind == "step2" and getattr(args, "initial_guess_option", None) == "2a"
    say("I am taking the initial guess from Step 2a output" if from_2a else "I am taking the initial guess from Step 1 output")

    def start_fn(image, hdr, path):
        if from_2a:                                                                # apf_step2.py:248-256
            a = np.genfromtxt(chains.results_dir(path) + "step2a.csv", delimiter=",")
            return np.array(a[-1][:P], dtype=np.float64)
        guess = np.loadtxt(open(chains.initial_guess_path(path), "rb"), delimiter=" ")   # :258-273
        return frame.initial_parameters(image, guess, nbody)

    if args.burn_in is not None:
        burn_in = args.burn_in
    else:
        burn_in = 6000 if (kind == "step2" and nbody == 2) else 0    # apf_step2.py:38; 3body :38; 2a: SURVEY C3
    floor_index = layout.bkgd_index(nbody) if (args.fix_bkgd and nbody == 2) else layout.REFERENCE_FLOOR_INDEX
    dom, params, cuts, headers = frame.load_epochs(images, start_fn, nbody=nbody, size=args.stamp,
                                                   floor_index=floor_index, device=device,
                                                   whole_frame=(args.domain == "frame"))
    if n_frames == 1:
        say("Masking pixels greater than ", 0.8 * frame.saturation_level(headers[0]))
    say("I have masked", int((dom.weight == 0).sum().item()), "pixels (inside the %dx%d domain%s)"
        % (args.stamp, args.stamp, " at x0=%d y0=%d" % tuple(cuts[0]) if n_frames == 1 else "s of all epochs"))

    per_frame = 1 if kind == "step2a" else args.walkers
    total_walkers = per_frame * n_frames
    id_base, id_stride, n_local = dist.shard_ids(total_walkers, rank, world)
    seed = args.seed if args.seed is not None else int.from_bytes(os.urandom(8), "little")
    if world > 1:   # every rank must use rank 0's seed
        t = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
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
    # Global walker g = w * n_frames + frame: walker w of epoch `frame` (the reference's rank w of that image;
    # one epoch: g = w).  Ranks own g = rank (mod world), so with the epochs a multiple of the ranks every rank
    # owns WHOLE epochs (frame = rank mod world): its warps fill with walkers of one stamp (1,048 per
    # epoch instead of 131 at a million walkers, 1,000 epochs, 8 GPUs: +26 % at 32 pixels), and the
    # random streams -- keyed by g -- do not depend on the number of ranks.
    ids = np.array([id_base + i * id_stride for i in range(n_local)], dtype=np.int64)
    ids_or_0 = ids if n_local else np.zeros(1, dtype=np.int64)
    frame_of = (ids_or_0 % n_frames).astype(np.int32)
    sam = smp.GibbsSampler(dom, params[frame_of], frame_of, seed=seed, burn_in=burn_in,
                           thin=args.thin, id_base=id_base, id_stride=id_stride, team_warps=team)
    st0, _, _ = sam.state()
    if n_local:
        say("Found initial chi-squared:", float(st0[0, -1].item()))
        say("Initial guess:", np.concatenate([params[frame_of[0]], [float(st0[0, -1].item())]]))
    say()
    say("Beginning loop...")

    if kind == "step2a":
        paths = [outdirs[f] + "step2a.csv" for f in frame_of[:n_local]]
        acc_paths = [outdirs[f] + "step2a_acceptance_rate" for f in frame_of[:n_local]]
    else:
        paths = [outdirs[g % n_frames] + "%d_finalarray_mpi.csv" % (g // n_frames) for g in ids]
        acc_paths = [outdirs[g % n_frames] + "%d_acceptance_rate.csv" % (g // n_frames) for g in ids]

    # separation / position angle of every recorded row, binned on the device (one centre per epoch,
    # the same on every rank: the starting point)
    want_sketch = kind != "step2a"
    if want_sketch:
        centers = np.zeros((n_frames, nbody - 1, 2))
        for f in range(n_frames):
            for o in range(1, nbody):
                dx, dy = params[f][2 * o] - params[f][0], params[f][2 * o + 1] - params[f][1]
                centers[f, o - 1] = (np.hypot(dx, dy), np.degrees(np.arctan2(-dx, dy)))
        sam.enable_sketch(centers=centers)

    record = not args.no_chains
    packed = None
    if args.format == "bin" and record:
        if args.chain_dtype == "f32":
            sam.set_chain_format("f32delta")
        packed = chains.PackedChainWriter(
            outdirs[0] + "chains_rank%d" % rank, max(n_local, 1), P + 1,
            {"walker_ids": ids.tolist(), "walkers_per_frame": per_frame, "frames": images, "seed": seed,
             "id_layout": "walker_id = walker * n_frames + frame",
             "burn_in": burn_in, "thin": args.thin, "nbody": nbody, "origin": [int(v) for v in cuts[0]],
             "cuts": cuts.tolist(), "stamp": args.stamp},
            dtype="float32" if args.chain_dtype == "f32" else "float64",
            start=sam.start().cpu().numpy() if args.chain_dtype == "f32" else None, resume=args.resume)

    streamer = smp.ChainStreamer(sam, args.segment) if record else None
    first = [True]
    ckpt_path = outdirs[0] + "checkpoint_rank%d.pt" % rank
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

    def advance(n):
        if record:
            consume(streamer.run(n))
        else:
            sam.run(n, record=False)

    t_start = time.time()
    if kind == "step2a":
        left = args.n_steps                                   # apf_step2a.py:271
        while left > 0:
            n = min(args.segment, left)
            advance(n)
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
            advance(n)
            if sam.count % (10 * args.segment) < n:
                say("Loop count:", sam.count, " min tries:", int(mn.item()),
                    " acceptance rate:", (accepts.double() / tries.double().clamp(min=1)).cpu().numpy())
    if record:
        consume(streamer.finish())
    if packed is None and record and first[0] and n_local:
        for pth in paths:                                     # nothing recorded: the nan row alone
            chains.write_walker_csv(pth, np.zeros((0, P + 1)))
    if packed is not None:
        packed.close(count=sam.count)
    if args.adapt:
        say("Jump widths after burn-in tuning:", widths)
    if args.checkpoint and n_local:
        torch.save({"blob": sam.save().cpu(), "n_walkers": sam.n_walkers, "nbody": nbody, "seed": seed,
                    "count": sam.count}, ckpt_path)

    # Summary per epoch from statistics reduced on the device and summed over ranks -- no chain file
    # is read back: Gelman-Rubin per parameter (apf_step3.py:260-278), separation and position angle
    # of each companion (median, 68 % interval, mean, std; :283-291,436-437; without the distortion,
    # rotation and refraction corrections, whose tables are not part of the checkout).
    summaries = [dict(image=images[f], walkers=per_frame, updates=sam.count) for f in range(n_frames)]
    stt = sam.stats(moments=True)
    n_rows = int(stt["rows"].item())
    if kind != "step2a" and per_frame > 1 and n_rows > 1:
        mom = dist.allreduce_sum(stt["moments"] if n_local else torch.zeros_like(stt["moments"]))
        if world > 1:                     # the reference value of a frame is rank-local; all starts of a frame are equal here
            mom[:, :, 0] = mom[:, :, 0] / world
        mom = mom.cpu().numpy()
        for f in range(n_frames):
            _, rc = chains.gelman_rubin_from_moments(mom[f, :P], n_rows, per_frame)
            summaries[f]["gelman_rubin"] = [float(v) for v in rc]
        say("Gelman-Rubin stat per parameter (recorded rows, all walkers%s):" % ("" if n_frames == 1 else ", first epoch"),
            np.round(summaries[0]["gelman_rubin"], 4))
        say("GR for positions:", *np.round(summaries[0]["gelman_rubin"][:2 * nbody], 4))
    if want_sketch and n_rows > 0:
        sk = sam.sketch()
        if not n_local:
            sk["hist"].zero_(); sk["summary"][..., 1:].zero_()
        sk["hist"] = dist.allreduce_sum(sk["hist"])
        sums = dist.allreduce_sum(sk["summary"][..., 1:].contiguous())
        sk["summary"] = torch.cat([sk["summary"][..., :1], sums], dim=-1)
        res = {k: v.cpu().numpy() for k, v in stats.sketch_summary(sk).items()}
        names = ["companion"] if nbody == 2 else ["b", "c"]
        for f in range(n_frames):
            for o, nm in enumerate(names):
                summaries[f]["sep_pa_%s" % nm] = {
                    "sep_mas": {"median": float(res["sep_q"][f, o, 1]), "lo": float(res["sep_q"][f, o, 0]),
                                "hi": float(res["sep_q"][f, o, 2]), "mean": float(res["sep_mean"][f, o]),
                                "std": float(res["sep_std"][f, o])},
                    "pa_deg": {"median": float(res["pa_q"][f, o, 1]), "lo": float(res["pa_q"][f, o, 0]),
                               "hi": float(res["pa_q"][f, o, 2]), "mean": float(res["pa_mean"][f, o]),
                               "std": float(res["pa_std"][f, o])},
                    "rows": int(res["count"][f, o]), "outside_histogram": float(res["outside"][f, o].max()),
                    "pixscale_mas": float(chains.PIXSCALE_PRE2015)}
        for f in range(min(n_frames, 8)):
            for nm in names:
                s_ = summaries[f]["sep_pa_%s" % nm]
                say("epoch %d %s: r = %.3f +%.3f -%.3f mas (std %.3f), pa = %.4f +%.4f -%.4f deg (std %.4f)  [%d rows]"
                    % (f, nm, s_["sep_mas"]["median"], s_["sep_mas"]["hi"] - s_["sep_mas"]["median"],
                       s_["sep_mas"]["median"] - s_["sep_mas"]["lo"], s_["sep_mas"]["std"], s_["pa_deg"]["median"],
                       s_["pa_deg"]["hi"] - s_["pa_deg"]["median"], s_["pa_deg"]["median"] - s_["pa_deg"]["lo"],
                       s_["pa_deg"]["std"], s_["rows"]))
    if rank == 0 and kind != "step2a":
        for f in range(n_frames):
            with open(outdirs[f] + "step2_summary.json", "w") as fh:
                json.dump(summaries[f], fh, indent=1)
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
