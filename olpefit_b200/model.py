"""Host-side mirror of the reference's pixel operator, running on the GPU through liblapf.

    build_analytical_model(p)            apf_step2.py:106-124   (3-body: 3body/...:106-125)
    chi_squared(data, model, error)      apf_step2.py:134-137

The reference keeps image_nanmask / err as module globals (apf_step2.py:188,210); here they live
in a ``PixelDomain`` on the device.  torch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, layout


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class PixelDomain:
    """F frames of ny x nx pixels resident in HBM: data, weight (= 1/err^2, 0 where masked) and
    the frame coordinates of pixel [0][0] of each frame."""

    def __init__(self, data, weight, origin=None, nbody=2, floor_index=None, device="cuda", outside=None,
                 cull=True, plain_loop=False):
        _lib.load()
        if not torch.cuda.is_available():
            raise _lib.LapfError("no CUDA device: olpefit_b200 has no CPU path")
        self.device = torch.device(device)
        data = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32)) if not torch.is_tensor(data) else data
        weight = torch.as_tensor(np.ascontiguousarray(weight, dtype=np.float32)) if not torch.is_tensor(weight) else weight
        if data.dim() == 2:
            data, weight = data[None], weight[None]
        if data.shape != weight.shape or data.dim() != 3:
            raise ValueError("data and weight must both be [F, ny, nx]")
        self.n_frames, self.ny, self.nx = (int(v) for v in data.shape)
        self.data = data.to(self.device, torch.float32).contiguous()
        self.weight = weight.to(self.device, torch.float32).contiguous()
        if origin is None:
            origin = np.zeros((self.n_frames, 2), dtype=np.int32)
        origin = torch.as_tensor(np.ascontiguousarray(origin, dtype=np.int32)) if not torch.is_tensor(origin) else origin
        self.origin = origin.to(self.device, torch.int32).reshape(self.n_frames, 2).contiguous()
        self.nbody = int(nbody)
        self.nparam = layout.nparam(self.nbody)
        self.floor_index = layout.REFERENCE_FLOOR_INDEX if floor_index is None else int(floor_index)
        self.flags = (0 if cull else 1) | (2 if plain_loop else 0)   # LAPF_FLAG_NO_CULL, LAPF_FLAG_PLAIN_LOOP
        # optional [F, 3] float64: sum w, sum w d, sum w d^2 over the image pixels outside the cut-outs
        self.outside = None
        if outside is not None:
            o = outside if torch.is_tensor(outside) else torch.as_tensor(np.asarray(outside, dtype=np.float64))
            self.outside = o.to(self.device, torch.float64).reshape(self.n_frames, 3).contiguous()

    def problem(self) -> _lib.Problem:
        return _lib.Problem(self.nbody, self.ny, self.nx, self.n_frames, self.floor_index, self.flags,
                            self.data.data_ptr(), self.weight.data_ptr(), self.origin.data_ptr(),
                            self.outside.data_ptr() if self.outside is not None else None)

    # ------------------------------------------------------------------------------------
    def model_chi2(self, params, frame_of=None, want_model=False, want_chi2=True):
        """Evaluate B parameter vectors.  params: [B, P] (numpy or torch, float64).
        Returns (model [B, ny, nx] float32 or None, chi2 [B] float64 or None) as device tensors."""
        p = params if torch.is_tensor(params) else torch.as_tensor(np.asarray(params, dtype=np.float64))
        p = p.to(self.device, torch.float64).reshape(-1, self.nparam).contiguous()
        nb = p.shape[0]
        fo = None
        if frame_of is not None:
            fo = frame_of if torch.is_tensor(frame_of) else torch.as_tensor(np.asarray(frame_of, dtype=np.int32))
            fo = fo.to(self.device, torch.int32).contiguous()
            if fo.numel() != nb:
                raise ValueError("frame_of must have one entry per parameter vector")
        model = torch.empty((nb, self.ny, self.nx), dtype=torch.float32, device=self.device) if want_model else None
        chi2 = torch.empty((nb,), dtype=torch.float64, device=self.device) if want_chi2 else None
        prob = self.problem()
        _lib.check(_lib.load().lapf_model_chi2(
            C.byref(prob), p.data_ptr(), nb, fo.data_ptr() if fo is not None else None,
            model.data_ptr() if model is not None else None,
            chi2.data_ptr() if chi2 is not None else None, _stream_ptr(self.device)))
        return model, chi2


def build_analytical_model(p, domain: PixelDomain, frame=0):
    """Model image for one parameter vector (apf_step2.py:106-124) -> [ny, nx] float32 device tensor."""
    m, _ = domain.model_chi2(np.asarray(p, dtype=np.float64)[None, :domain.nparam],
                             frame_of=[frame], want_model=True, want_chi2=False)
    return m[0]


def chi_squared(p, domain: PixelDomain, frame=0) -> float:
    """chi-square of one parameter vector against frame ``frame`` (apf_step2.py:134-137 applied to
    the model of apf_step2.py:106-124; fused on the device, the model image is never stored)."""
    _, c = domain.model_chi2(np.asarray(p, dtype=np.float64)[None, :domain.nparam],
                             frame_of=[frame], want_model=False, want_chi2=True)
    return float(c.item())
