"""Two real GPUs under torchrun + NCCL (skipped on a one-GPU box; run with `gpurun --gpus 2`):
the command line sharded over two ranks writes exactly the files of the one-rank run, its
all-reduced summary equals the one-rank summary, and dist.gather_chains brings the shards'
chains to rank 0 in walker order (BASELINE north star: NCCL only for the cross-chain statistics
and the final chain gather)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def _torchrun(args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return res


def _write_epochs(tmp_path, tag, n=3):
    from olpefit_b200 import chains, frame, synth
    paths = []
    for f in range(n):
        d = tmp_path / ("%s_epoch%d" % (tag, f))
        d.mkdir()
        img, truth = synth.make_frame(f, 2)
        p = str(d / ("N2.2009053%d.2996%d.LDIF.fits" % (f, f)))
        frame.write_fits(p, img, synth.HEADER)
        os.makedirs(chains.results_dir(p))
        chains.write_walker_csv(chains.results_dir(p) + "step2a.csv", np.concatenate([truth, [0.0]])[None])
        paths.append(p)
    lst = tmp_path / (tag + "_frames.txt")
    lst.write_text("\n".join(paths[1:]) + "\n")
    return paths, str(lst)


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_command_line_on_two_gpus_equals_one_gpu(tmp_path):
    from olpefit_b200 import chains, cli
    common = ["--walkers", "7", "--accept-min", "30", "--burn-in", "50", "--seed", "77", "--stamp", "64", "--thin", "2",
              "--segment", "96", "--quiet"]
    p1, l1 = _write_epochs(tmp_path, "one")
    assert cli.main_step2([p1[0], "-i", "2a", "--frames", l1] + common) == 0
    p2, l2 = _write_epochs(tmp_path, "two")
    _torchrun([os.path.join(ROOT, "apf_step2.py"), p2[0], "-i", "2a", "--frames", l2] + common, 29621)
    for a, b in zip(p1, p2):
        da, db = chains.results_dir(a), chains.results_dir(b)
        for w in range(7):
            fa = open(da + "%d_finalarray_mpi.csv" % w, "rb").read()
            assert fa == open(db + "%d_finalarray_mpi.csv" % w, "rb").read(), (a, w)       # byte for byte
            assert open(da + "%d_acceptance_rate.csv" % w).read() == open(db + "%d_acceptance_rate.csv" % w).read()
        sa, sb = json.load(open(da + "step2_summary.json")), json.load(open(db + "step2_summary.json"))
        assert sa["sep_pa_companion"]["rows"] == sb["sep_pa_companion"]["rows"]
        for key in ("sep_mas", "pa_deg"):
            for k2 in ("median", "lo", "hi"):                                              # integer histograms: exact
                assert sa["sep_pa_companion"][key][k2] == sb["sep_pa_companion"][key][k2]
            for k2 in ("mean", "std"):                                                     # FP64 sums in another order
                assert sa["sep_pa_companion"][key][k2] == pytest.approx(sb["sep_pa_companion"][key][k2], rel=1e-9)
        np.testing.assert_allclose(sa["gelman_rubin"], sb["gelman_rubin"], rtol=1e-6)


_GATHER = r"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, %(root)r)
from olpefit_b200 import dist, frame, sampler, synth
HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
rank, local_rank, world = dist.init()
torch.cuda.set_device(local_rank)
dev = "cuda:%%d" %% local_rank
total, nf = 37, 3
stamps, origins = synth.make_stamps(nf, 32)
dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=2, device=dev)
p = np.array([synth.truth_parameters(2, f) for f in range(nf)])
base, stride, n_local = dist.shard_ids(total, rank, world)
ids = base + stride * np.arange(n_local)
with sampler.GibbsSampler(dom, p[ids %% nf], (ids %% nf).astype(np.int32), seed=5, thin=3, id_base=base, id_stride=stride) as s:
    local = s.run(60)
    st = s.stats(moments=False)
tries, accepts, mn = dist.allreduce_stats(st["tries"], st["accepts"], st["min_tries"])
parts = dist.gather_chains(local, rank, world)
if rank == 0:
    merged = dist.merge_interleaved(parts)
    allw = np.arange(total)
    with sampler.GibbsSampler(dom, p[allw %% nf], (allw %% nf).astype(np.int32), seed=5, thin=3) as s:
        whole = s.run(60)
        st1 = s.stats(moments=False)
    assert merged.shape == whole.shape and torch.equal(merged, whole), "gathered chains differ from the one-rank run"
    assert torch.equal(tries, st1["tries"]) and torch.equal(accepts, st1["accepts"]) and int(mn) == int(st1["min_tries"])
    print("GATHER_OK", tuple(merged.shape))
dist.barrier()
torch.distributed.destroy_process_group()
"""


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_nccl_gather_of_sharded_chains(tmp_path):
    script = tmp_path / "gather.py"
    script.write_text(_GATHER % {"root": ROOT})
    res = _torchrun([str(script)], 29622)
    assert "GATHER_OK (20, 37, 17)" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
