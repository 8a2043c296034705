import time, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
from olpefit_b200 import frame, sampler, synth
HEADER = {"itime": 1.0, "coadds": 1, "multisam": 1, "sampmode": 2}
dev = torch.device('cuda:0')
W, F, S, U, P = 65536, 100, 64, 128, 16
stamps, origins = synth.make_stamps(F, S, 2)
frame_of = (np.arange(W) % F).astype(np.int32)
p0 = frame.initial_parameters(stamps[0], synth.step1_guess(stamps[0], 2, origin=tuple(origins[0])), 2, origin=tuple(origins[0]))
init = np.tile(p0, (W, 1))
frames_h = torch.from_numpy(stamps).pin_memory(); init_h = torch.from_numpy(init).pin_memory()
dom = frame.prepare_domain(frames_h.to(dev), HEADER, origin=origins, nbody=2)
smp = sampler.GibbsSampler(dom, init_h.to(dev), torch.from_numpy(frame_of).to(dev), seed=1, thin=16)
rows = smp.rows_for(U)
chain = torch.empty((rows, W, P + 1), dtype=torch.float64, device=dev)
chain_h = torch.empty((rows, W, P + 1), dtype=torch.float64).pin_memory()
frames_d = torch.empty_like(frames_h, device=dev); init_d = torch.empty_like(init_h, device=dev)
def T(name, fn, n=5):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print('%-28s %8.2f ms (min %.2f)' % (name, 1e3 * np.mean(ts[1:]), 1e3 * min(ts)))
T('h2d frames+init', lambda: (frames_d.copy_(frames_h, non_blocking=True), init_d.copy_(init_h, non_blocking=True)))
T('prepare_domain into', lambda: frame.prepare_domain(frames_d, HEADER, origin=origins, nbody=2, into=dom))
T('reset', lambda: smp.reset(init_d, seed=3))
T('run', lambda: smp.run(U, out=chain))
T('d2h chain %d MB' % (chain_h.numel() * 8 >> 20), lambda: chain_h.copy_(chain, non_blocking=True))
T('stats', lambda: smp.stats(moments=False))
print('pinned?', chain_h.is_pinned(), frames_h.is_pinned())

tot_h = torch.empty((2 * P + 2,), dtype=torch.int64).pin_memory()
def whole():
    frames_d.copy_(frames_h, non_blocking=True); init_d.copy_(init_h, non_blocking=True)
    frame.prepare_domain(frames_d, HEADER, origin=origins, nbody=2, into=dom)
    smp.reset(init_d, seed=3)
    ch = smp.run(U, out=chain)
    chain_h.copy_(ch, non_blocking=True)
    stt = smp.stats(moments=False)
    tot_h.copy_(torch.cat([stt["tries"], stt["accepts"], stt["min_tries"].reshape(1), stt["exps"].reshape(1)]), non_blocking=True)
T('whole e2e step', whole, n=8)
def whole_nosync_prep():
    frames_d.copy_(frames_h, non_blocking=True); init_d.copy_(init_h, non_blocking=True)
    t0 = time.perf_counter(); frame.prepare_domain(frames_d, HEADER, origin=origins, nbody=2, into=dom); t1 = time.perf_counter()
    smp.reset(init_d, seed=3); t2 = time.perf_counter()
    ch = smp.run(U, out=chain); t3 = time.perf_counter()
    chain_h.copy_(ch, non_blocking=True); t4 = time.perf_counter()
    stt = smp.stats(moments=False); t5 = time.perf_counter()
    print('  host ms: prep %.2f reset %.2f run %.2f d2h-enqueue %.2f stats %.2f' % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)))
whole_nosync_prep(); torch.cuda.synchronize(); whole_nosync_prep(); torch.cuda.synchronize()
