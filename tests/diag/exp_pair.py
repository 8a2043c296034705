"""Diagnosis for the half-warp (two vectors per pass) mode on 32-pixel stamps: dump the self-test of
a failing build and say which trial vectors differ (by how much, which path, which lane parity)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["LAPF_SELFTEST_DUMP"] = "/tmp/selftest.bin"
import torch
from olpefit_b200 import frame, sampler, _lib
from oracle import lapf_oracle as orc   # the checker (this script is test infrastructure: tests/diag)
from olpefit_b200 import synth
HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}

def REF(t, img, wmap, origin, nbody):
    lay = orc.layout_for(nbody)
    m = orc.model_image(t, lay, 32, 32, origin=(int(origin[0]), int(origin[1])))
    return orc.chi_squared_weighted(img, m, wmap)


def main():
    nbody = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    stamps, origins = synth.make_stamps(1, 32, nbody)
    dom = frame.prepare_domain(stamps, HEADER, origin=origins, nbody=nbody)
    guess = synth.step1_guess(stamps[0], nbody, origin=tuple(origins[0]))
    p0 = frame.initial_parameters(stamps[0], guess, nbody, origin=tuple(origins[0]))
    init = np.tile(p0, (W, 1))
    img = stamps[0].astype(np.float64)
    wmap = orc.weight_map(img, HEADER)
    try:
        s = sampler.GibbsSampler(dom, init, None, seed=3)
        print("self-test passed")
    except _lib.LapfError as e:
        print(str(e).splitlines()[0][:200])
    raw = open("/tmp/selftest.bin", "rb").read()
    U, Wn, P = np.frombuffer(raw[:24], dtype=np.int64)
    body = np.frombuffer(raw[24:], dtype=np.float64)
    n = U * Wn
    probe = body[:3 * n].reshape(U, Wn, 3)
    chi = body[3 * n:4 * n].reshape(U, Wn)
    trials = body[4 * n:].reshape(U, Wn, P)
    a, b = probe[..., 2], chi
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    print("mismatches", int(bad.sum()), "of", n)
    u, w = np.nonzero(bad)
    rel = np.abs(a[bad] - b[bad]) / np.abs(b[bad])
    print("relative differences: min %.3g median %.3g max %.3g" % (rel.min(), np.median(rel), rel.max()))
    print("by update:", np.bincount(u, minlength=U).tolist())
    print("by lane parity:", np.bincount(w % 2, minlength=2).tolist(), " by lane:", np.bincount(w % 32, minlength=32).tolist())
    print("by parameter:", np.bincount(probe[..., 0][bad].astype(int), minlength=P).tolist())
    # purity of the stateless operator: the same vectors alone, and with the partners swapped
    flat = trials.reshape(n, P)
    def k1(v):
        _, c = dom.model_chi2(torch.as_tensor(np.ascontiguousarray(v), device="cuda"))
        return c.cpu().numpy()
    c_pair = k1(flat)
    print("K1 again == K1 of the self-test:", bool(np.array_equal(c_pair, chi.reshape(-1), equal_nan=True)))
    swapped = flat.reshape(n // 2, 2, P)[:, ::-1].reshape(n, P)
    c_sw = k1(swapped).reshape(n // 2, 2)[:, ::-1].reshape(n)
    print("K1 with halves swapped differs in", int((c_sw != c_pair).sum()))
    dup = np.repeat(flat, 2, axis=0)                      # every vector paired with itself
    c_dup = k1(dup).reshape(n, 2)
    print("K1 self-paired: halves differ in", int((c_dup[:, 0] != c_dup[:, 1]).sum()),
          "; differs from pair-run in", int((c_dup[:, 0] != c_pair).sum()),
          "; differs from the sampler in", int((c_dup[:, 0] != a.reshape(-1)).sum()))
    # which of the two agrees with float64?
    for i in range(min(8, len(u))):
        t = trials[u[i], w[i]]
        ref = REF(t, img, wmap, origins[0], nbody)
        print("u %d w %d k %d sampler %.17g k1 %.17g oracle %.17g" % (u[i], w[i], int(probe[u[i], w[i], 0]), a[u[i], w[i]], b[u[i], w[i]], ref))

main()
