"""Host-side driver of the batched Gibbs-within-Metropolis sampler (apf_step2.py:276-351).

The reference runs one MPI rank per walker (apf_step2.py:54-57); here one ``GibbsSampler`` owns
a batch of independent walkers on one GPU.  Per update and walker the device does exactly what
the reference loop does: pick a parameter uniformly (:302), propose (:306-309), evaluate
model + chi-square (:314-316), accept or reject (:318-327), count, and record a chain row once
count >= burn_in (:342-351).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import PixelDomain, _stream_ptr


class GibbsSampler:
    def __init__(self, domain: PixelDomain, init_params, frame_of=None, *, seed=0, burn_in=0, thin=1,
                 widths=None, id_base=0, id_stride=1, team_warps=0):
        self.lib = _lib.load()
        self.domain = domain
        dev = domain.device
        p = init_params if torch.is_tensor(init_params) else torch.as_tensor(np.asarray(init_params, dtype=np.float64))
        p = p.to(dev, torch.float64).reshape(-1, domain.nparam).contiguous()
        self.n_walkers = int(p.shape[0])
        self.nparam = domain.nparam
        fo = None
        if frame_of is not None:
            fo = frame_of if torch.is_tensor(frame_of) else torch.as_tensor(np.asarray(frame_of, dtype=np.int32))
            fo = fo.to(dev, torch.int32).contiguous()
            if fo.numel() != self.n_walkers:
                raise ValueError("frame_of must have one entry per walker")
        self._keep = (p, fo)   # the library copies init_params; frame_of is only read in create()
        w_arr = None
        if widths is not None:
            w_np = np.ascontiguousarray(widths, dtype=np.float64)
            if w_np.shape != (self.nparam,):
                raise ValueError("widths must have %d entries" % self.nparam)
            w_arr = (C.c_double * self.nparam)(*w_np.tolist())
        cfg = _lib.Config(domain.problem(), self.n_walkers, int(id_base), int(id_stride),
                          int(seed) & 0xFFFFFFFFFFFFFFFF, fo.data_ptr() if fo is not None else None,
                          p.data_ptr(), w_arr, int(burn_in), int(thin), int(team_warps))
        self.burn_in, self.thin = int(burn_in), int(thin)
        self.chain_dtype = torch.float64
        self._sketch = None
        handle = C.c_void_p()
        _lib.check(self.lib.lapf_sampler_create(C.byref(cfg), C.byref(handle), _stream_ptr(dev)))
        self._h = handle

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            torch.cuda.synchronize(self.domain.device)
            self.lib.lapf_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self, init_params, seed=0):
        """Start a new batch in this sampler: new starting points and seed; counters, moments
        and the update count return to zero and the initial chi-square is re-evaluated against
        the domain's current pixel buffers (see ``frame.prepare_domain(into=...)``)."""
        dev = self.domain.device
        p = init_params if torch.is_tensor(init_params) else torch.as_tensor(np.asarray(init_params, dtype=np.float64))
        p = p.to(dev, torch.float64).reshape(-1, self.nparam).contiguous()
        if p.shape[0] != self.n_walkers:
            raise ValueError("init_params must have one row per walker")
        self._keep = (p, self._keep[1])
        _lib.check(self.lib.lapf_sampler_reset(self._h, p.data_ptr(), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                               _stream_ptr(dev)))

    def save(self):
        """Checkpoint of the whole batch as a device uint8 tensor (see lapf_sampler_save)."""
        n = int(_lib.check(self.lib.lapf_sampler_checkpoint_bytes(self._h)))
        blob = torch.empty(n, dtype=torch.uint8, device=self.domain.device)
        _lib.check(self.lib.lapf_sampler_save(self._h, blob.data_ptr(), n, _stream_ptr(self.domain.device)))
        return blob

    def load(self, blob):
        """Restore a checkpoint made by ``save`` on a sampler of the same shape: the following runs
        continue the chains bit for bit."""
        blob = blob.to(self.domain.device).contiguous()
        _lib.check(self.lib.lapf_sampler_load(self._h, blob.data_ptr(), blob.numel(), _stream_ptr(self.domain.device)))

    def set_widths(self, widths):
        """New jump widths for the following runs (burn-in tuning; the reference's are fixed)."""
        w_np = np.ascontiguousarray(widths, dtype=np.float64)
        if w_np.shape != (self.nparam,):
            raise ValueError("widths must have %d entries" % self.nparam)
        _lib.check(self.lib.lapf_sampler_set_widths(self._h, (C.c_double * self.nparam)(*w_np.tolist())))

    def set_chain_format(self, fmt: str):
        """'f64': rows are the values themselves (default).  'f32delta': rows are float32 differences
        from the walker's starting point (``start()``) -- half the bytes to drain, and exact to ~1e-7
        of the distance travelled (see include/lapf.h)."""
        code = {"f64": 0, "f32delta": 1}[fmt]
        _lib.check(self.lib.lapf_sampler_set_chain_format(self._h, code))
        self.chain_dtype = torch.float32 if code else torch.float64

    def start(self):
        """[W, P+1] float64 device tensor: every walker's starting point and its chi-square."""
        out = torch.empty((self.n_walkers, self.nparam + 1), dtype=torch.float64, device=self.domain.device)
        _lib.check(self.lib.lapf_sampler_start(self._h, out.data_ptr(), _stream_ptr(self.domain.device)))
        return out

    # -- separation / position angle without the chain (apf_step3.py:255-256,283-291,436-437) ---------
    def enable_sketch(self, n_bins=16384, sep_bin=5e-4, pa_bin=2e-3, centers=None):
        """From now on every recorded row enters per-frame histograms of separation (pixels) and
        position angle (degrees) of each companion; call before the first ``run``.  ``centers``:
        [F, nbody-1, 2] centre values (give every rank the same ones), default: the starting point
        of each frame's first walker."""
        dev = self.domain.device
        c = None
        if centers is not None:
            c = centers if torch.is_tensor(centers) else torch.as_tensor(np.asarray(centers, dtype=np.float64))
            c = c.to(dev, torch.float64).reshape(self.domain.n_frames, self.domain.nbody - 1, 2).contiguous()
        _lib.check(self.lib.lapf_sampler_sketch_enable(self._h, int(n_bins), float(sep_bin), float(pa_bin),
                                                       c.data_ptr() if c is not None else None, _stream_ptr(dev)))
        self._sketch = (int(n_bins), float(sep_bin), float(pa_bin))

    def sketch(self):
        """dict(hist [F, nbody-1, 2, n_bins+2] int64, summary [F, nbody-1, 2, 4] float64 (centre, sum,
        sum of squares of value - centre, count), n_bins, sep_bin, pa_bin) as device tensors."""
        if self._sketch is None:
            raise _lib.LapfError("sketches are not enabled")
        dev = self.domain.device
        n_bins, sep_bin, pa_bin = self._sketch
        shape = (self.domain.n_frames, self.domain.nbody - 1, 2)
        hist = torch.empty(shape + (n_bins + 2,), dtype=torch.int32, device=dev)      # uint32 counts
        summ = torch.empty(shape + (4,), dtype=torch.float64, device=dev)
        _lib.check(self.lib.lapf_sampler_sketch(self._h, hist.data_ptr(), summ.data_ptr(), _stream_ptr(dev)))
        return {"hist": hist.long() & 0xFFFFFFFF, "summary": summ, "n_bins": n_bins, "sep_bin": sep_bin, "pa_bin": pa_bin}

    # -- the loop ------------------------------------------------------------------------
    @property
    def count(self) -> int:
        return int(self.lib.lapf_sampler_count(self._h))

    @property
    def launches(self) -> int:
        return int(self.lib.lapf_sampler_launches(self._h))

    def rows_for(self, n_updates: int) -> int:
        return int(_lib.check(self.lib.lapf_sampler_rows_for(self._h, int(n_updates))))

    def run(self, n_updates: int, record=True, out=None):
        """Advance all walkers by ``n_updates`` updates.  Returns the chain rows recorded by this
        call as a device tensor [rows, n_walkers, P+1] (float64, or float32 differences from
        ``start()`` after ``set_chain_format('f32delta')``; None when record is False)."""
        rows = self.rows_for(n_updates)
        chain = None
        if record:
            if out is not None:
                if out.dtype != self.chain_dtype or out.numel() < rows * self.n_walkers * (self.nparam + 1):
                    raise ValueError("out must be %s and hold %d rows" % (self.chain_dtype, rows))
                chain = out
            else:
                chain = torch.empty((rows, self.n_walkers, self.nparam + 1), dtype=self.chain_dtype,
                                    device=self.domain.device)
        _lib.check(self.lib.lapf_sampler_run(self._h, int(n_updates),
                                             chain.data_ptr() if chain is not None and rows > 0 else None,
                                             rows, _stream_ptr(self.domain.device)))
        if chain is not None and out is not None:
            return chain.reshape(-1)[: rows * self.n_walkers * (self.nparam + 1)].view(
                rows, self.n_walkers, self.nparam + 1)
        return chain

    def state(self):
        """(state [W, P+1], tries [W, P], accepts [W, P]) device tensors: parameters + chi-square,
        total_tries and total_accept of apf_step2.py:276."""
        dev = self.domain.device
        st = torch.empty((self.n_walkers, self.nparam + 1), dtype=torch.float64, device=dev)
        tr = torch.empty((self.n_walkers, self.nparam), dtype=torch.int32, device=dev)
        ac = torch.empty_like(tr)
        _lib.check(self.lib.lapf_sampler_state(self._h, st.data_ptr(), tr.data_ptr(), ac.data_ptr(),
                                               _stream_ptr(dev)))
        return st, tr, ac

    def stats(self, moments=True):
        """Batch statistics reduced on the device (K4).  Returns a dict of device tensors:
        tries[P], accepts[P], min_tries (scalar), exps (scalar: component evaluations, pixels x Gaussian components, really done),
        moments [F, P+1, 4] (reference, sum of centred chain means, of their squares, of chain
        variances), walkers_per_frame [F], rows (scalar)."""
        dev = self.domain.device
        pn, nf = self.nparam, self.domain.n_frames
        tot = torch.empty((2 * pn + 2,), dtype=torch.int64, device=dev)
        mom = torch.empty((nf, pn + 1, 4), dtype=torch.float64, device=dev) if moments else None
        cnt = torch.empty((nf + 1,), dtype=torch.int64, device=dev)
        _lib.check(self.lib.lapf_sampler_stats(self._h, tot.data_ptr(),
                                               mom.data_ptr() if mom is not None else None,
                                               cnt.data_ptr(), _stream_ptr(dev)))
        return {"tries": tot[:pn], "accepts": tot[pn:2 * pn], "min_tries": tot[2 * pn], "exps": tot[2 * pn + 1],
                "moments": mom, "walkers_per_frame": cnt[:nf], "rows": cnt[nf]}


class ChainStreamer:
    """Double-buffered chain output (K3): while the GPU computes segment i+1 into one device
    buffer, segment i travels to pinned host memory on a copy stream and is handed to the caller.
    Replaces the reference's in-memory hstack + whole-file rewrite (apf_step2.py:346-360).

        streamer = ChainStreamer(sampler, max_updates_per_segment)
        seg = streamer.run(n)      # launches n updates; returns the PREVIOUS segment (numpy) or None
        seg = streamer.finish()    # the last segment
    """

    def __init__(self, sampler: GibbsSampler, max_updates: int):
        self.s = sampler
        dev = sampler.domain.device
        cap_rows = max(1, -(-int(max_updates) // sampler.thin) + 1)
        shape = (cap_rows, sampler.n_walkers, sampler.nparam + 1)
        self.dev = [torch.empty(shape, dtype=sampler.chain_dtype, device=dev) for _ in range(2)]
        self.host = [torch.empty(shape, dtype=sampler.chain_dtype).pin_memory() for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.pending = None          # (buffer index, rows)
        self.i = 0
        self.max_updates = int(max_updates)

    def _collect(self):
        if self.pending is None:
            return None
        j, rows = self.pending
        self.pending = None
        self.done[j].synchronize()
        return self.host[j][:rows].numpy()

    def run(self, n_updates: int):
        if n_updates > self.max_updates:
            raise ValueError("segment longer than the streamer was sized for")
        s, i = self.s, self.i
        compute = torch.cuda.current_stream(s.domain.device)
        compute.wait_event(self.done[i])            # the copy out of dev[i] two segments ago is finished
        chain = s.run(n_updates, out=self.dev[i])
        rows = int(chain.shape[0])
        if rows:
            nbytes = rows * s.n_walkers * (s.nparam + 1) * self.dev[i].element_size()
            _lib.check(s.lib.lapf_chain_drain(self.dev[i].data_ptr(), self.host[i].data_ptr(), nbytes,
                                              compute.cuda_stream, self.copy_stream.cuda_stream))
        self.done[i].record(self.copy_stream)
        prev = self._collect()                      # overlaps with the segment just launched
        self.pending = (i, rows)
        self.i = 1 - i
        return prev

    def finish(self):
        return self._collect()
