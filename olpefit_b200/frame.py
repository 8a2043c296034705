"""Frame input and preparation for step 2 (apf_step2.py:160-210).

* a minimal FITS primary-HDU reader/writer (the image ships no astropy): 2880-byte blocks,
  BITPIX 8/16/32/-32/-64, BSCALE/BZERO, and the header cards step 2 reads
  (ITIME, COADDS, MULTISAM, SAMPMODE; apf_step2.py:176-179);
* the header-derived scalars (saturation level, read noise; :181-204);
* the per-pixel part -- saturation mask, noise map and cut-out -- on the device
  (``lapf_frame_prep``), producing the (data, weight) pair the kernels consume.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from .model import PixelDomain, _stream_ptr

_BLOCK = 2880
_BITPIX_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


# ----------------------------------------------------------------------------------------------
# FITS
# ----------------------------------------------------------------------------------------------
def _parse_value(raw: str):
    raw = raw.strip()
    if raw.startswith("'"):
        end = raw.find("'", 1)
        while end != -1 and raw[end:end + 2] == "''":
            end = raw.find("'", end + 2)
        return raw[1:end].rstrip().replace("''", "'")
    raw = raw.split("/")[0].strip()
    if raw in ("T", "F"):
        return raw == "T"
    try:
        return int(raw)
    except ValueError:
        try:
            return float(raw.replace("D", "E"))
        except ValueError:
            return raw


class Header(dict):
    """Case-insensitive header mapping (astropy headers are; the reference uses lower case)."""

    def __getitem__(self, key):
        return dict.__getitem__(self, key.upper())

    def __setitem__(self, key, value):
        dict.__setitem__(self, key.upper(), value)

    def __contains__(self, key):
        return dict.__contains__(self, str(key).upper())

    def get(self, key, default=None):
        return dict.get(self, key.upper(), default)


def read_fits(path):
    """Primary HDU -> (image ndarray in native byte order, Header)."""
    with open(path, "rb") as fh:
        hdr = Header()
        done = False
        while not done:
            block = fh.read(_BLOCK)
            if len(block) < _BLOCK:
                raise ValueError("%s: truncated FITS header" % path)
            for i in range(0, _BLOCK, 80):
                card = block[i:i + 80].decode("ascii", "replace")
                key = card[:8].strip()
                if key == "END":
                    done = True
                    break
                if card[8:10] == "= " and key:
                    hdr[key] = _parse_value(card[10:])
        if not hdr.get("SIMPLE", False):
            raise ValueError("%s: not a simple FITS file" % path)
        naxis = int(hdr["NAXIS"])
        shape = [int(hdr["NAXIS%d" % (i + 1)]) for i in range(naxis)][::-1]
        bitpix = int(hdr["BITPIX"])
        count = int(np.prod(shape)) if shape else 0
        raw = np.frombuffer(fh.read(count * abs(bitpix) // 8), dtype=_BITPIX_DTYPE[bitpix])
    img = raw.reshape(shape)
    bscale, bzero = float(hdr.get("BSCALE", 1.0)), float(hdr.get("BZERO", 0.0))
    if bscale != 1.0 or bzero != 0.0:
        img = img.astype(np.float64) * bscale + bzero
    else:
        img = img.astype(img.dtype.newbyteorder("="))
    return img, hdr


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % value
    elif isinstance(value, (float, np.floating)):
        v = "%20s" % repr(float(value)).upper()
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
    s = "%-8s= %s" % (key.upper()[:8], v)
    if comment:
        s += " / " + comment
    return ("%-80s" % s)[:80]


def write_fits(path, image, header=None):
    """Write a primary HDU (float32/float64/int16/int32 image) with extra header cards."""
    image = np.asarray(image)
    bitpix = {np.dtype("float32"): -32, np.dtype("float64"): -64, np.dtype("int16"): 16,
              np.dtype("int32"): 32, np.dtype("uint8"): 8}[image.dtype]
    cards = [_card("SIMPLE", True), _card("BITPIX", bitpix), _card("NAXIS", image.ndim)]
    for i, n in enumerate(image.shape[::-1]):
        cards.append(_card("NAXIS%d" % (i + 1), int(n)))
    for k, v in (header or {}).items():
        if k.upper() in ("SIMPLE", "BITPIX", "NAXIS", "END") or k.upper().startswith("NAXIS"):
            continue
        cards.append(_card(k, v))
    cards.append("%-80s" % "END")
    hdr = "".join(cards).encode("ascii")
    hdr += b" " * (-len(hdr) % _BLOCK)
    data = image.astype(_BITPIX_DTYPE[bitpix]).tobytes()
    data += b"\0" * (-len(data) % _BLOCK)
    with open(path, "wb") as fh:
        fh.write(hdr)
        fh.write(data)


# ----------------------------------------------------------------------------------------------
# header scalars
# ----------------------------------------------------------------------------------------------
def saturation_level(header) -> float:
    """apf_step2.py:176-185."""
    itime = float(header["itime"]) * 1000.0
    coadds = float(header["coadds"])
    multisam = float(header["multisam"])
    if header["sampmode"] == 3:
        return coadds * 24000.0 * (1.0 - 0.1 * (multisam - 1.0) / (itime / 1000.0))
    return coadds * 22000.0


def read_noise(header) -> float:
    """apf_step2.py:197-204."""
    coadds = float(header["coadds"])
    multisam = float(header["multisam"])
    if header["sampmode"] == 3.0:
        return (38.0 / math.sqrt(multisam)) * math.sqrt(coadds)
    return 38.0 * math.sqrt(coadds)


# ----------------------------------------------------------------------------------------------
# device frame preparation
# ----------------------------------------------------------------------------------------------
def prepare_domain(frames, header, size=None, cut=None, origin=None, nbody=2, floor_index=None,
                   device="cuda", into: PixelDomain = None, whole_frame=False) -> PixelDomain:
    """Mask (image > 0.8*satlevel, apf_step2.py:188), noise map (err^2 = readnoise^2 + |image|,
    :197-210) and cut-outs, all on the device.

    frames  [F, fy, fx] (or one [fy, fx]) pixel arrays
    origin  frame coordinates (x0, y0) of frames[f][0][0]: one pair or [F, 2]; default (0, 0)
    cut     (x, y) of the cut-out inside each array: one pair or [F, 2]; default (0, 0)
    size    cut-out size, an int or (ny, nx); default: the whole array
    into    an existing PixelDomain of the same shape whose pixel buffers are overwritten in place
            (a new batch of epochs for a sampler that is then ``reset``)
    whole_frame
            also reduce the pixels of each array OUTSIDE its cut-out to (sum w, sum w d, sum w d^2).
            The reference evaluates chi-square over the whole frame (apf_step2.py:94,134-137); far
            from the objects its model is the constant floor, so those three sums reproduce the
            whole-frame chi-square exactly (as a quadratic in the floor) at cut-out cost.
    """
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.LapfError("no CUDA device: olpefit_b200 has no CPU path")
    dev = into.device if into is not None else torch.device(device)
    fr = frames if torch.is_tensor(frames) else torch.as_tensor(np.ascontiguousarray(frames, dtype=np.float32))
    if fr.dim() == 2:
        fr = fr[None]
    fr = fr.to(dev, torch.float32).contiguous()
    nf, fy, fx = (int(v) for v in fr.shape)
    if size is None:
        ny, nx = fy, fx
    else:
        ny, nx = (int(size), int(size)) if np.isscalar(size) else (int(size[0]), int(size[1]))

    def pairs(v):
        a = np.zeros((nf, 2), dtype=np.int64) if v is None else np.asarray(v, dtype=np.int64)
        return np.ascontiguousarray(np.broadcast_to(a.reshape(-1, 2), (nf, 2)))

    cut_np, org_np = pairs(cut), pairs(origin)
    if into is not None:
        if (into.n_frames, into.ny, into.nx) != (nf, ny, nx):
            raise ValueError("into has shape %s, new frames give %s" % ((into.n_frames, into.ny, into.nx), (nf, ny, nx)))
        data, weight = into.data, into.weight
        # cut-out positions and origins travel through a small pinned staging buffer kept on the
        # domain, and only when they changed: a pageable copy here would make the host wait for the
        # stream and serialise a stream of batches (bench.py e2e, ChainStreamer)
        new = np.stack([cut_np, org_np + cut_np]).astype(np.int32)
        st = getattr(into, "_stage", None)
        if st is None:
            st = {"host": torch.empty((2, nf, 2), dtype=torch.int32).pin_memory(),
                  "cut": torch.empty((nf, 2), dtype=torch.int32, device=dev),
                  "event": torch.cuda.Event(), "valid": False}
            into._stage = st
        if not (st["valid"] and np.array_equal(st["host"].numpy(), new)):
            st["event"].synchronize()            # the previous upload has left the staging buffer
            st["host"].numpy()[...] = new
            st["cut"].copy_(st["host"][0], non_blocking=True)
            into.origin.copy_(st["host"][1], non_blocking=True)
            st["event"].record(torch.cuda.current_stream(dev))
            st["valid"] = True
        cut_t = st["cut"]
    else:
        cut_t = torch.as_tensor(cut_np.astype(np.int32)).to(dev)
        data = torch.empty((nf, ny, nx), dtype=torch.float32, device=dev)
        weight = torch.empty_like(data)
    _lib.check(lib.lapf_frame_prep(fr.data_ptr(), nf, fy, fx, cut_t.data_ptr(), ny, nx,
                                   float(saturation_level(header)), float(read_noise(header)),
                                   data.data_ptr(), weight.data_ptr(), _stream_ptr(dev)))
    outside = None
    if whole_frame:
        # sum w, sum w d, sum w d^2 over the pixels outside each cut-out: one pass over the frames on the
        # device, FP64, fixed order (lapf_frame_outside)
        outside = into.outside if (into is not None and into.outside is not None) else \
            torch.empty((nf, 3), dtype=torch.float64, device=dev)
        _lib.check(lib.lapf_frame_outside(fr.data_ptr(), nf, fy, fx, cut_t.data_ptr(), ny, nx,
                                          float(saturation_level(header)), float(read_noise(header)),
                                          outside.data_ptr(), _stream_ptr(dev)))
    if into is not None:
        if outside is not None and into.outside is None:
            raise ValueError("into was created without whole_frame=True")
        return into
    return PixelDomain(data, weight, (org_np + cut_np).astype(np.int32), nbody=nbody,
                       floor_index=floor_index, device=device, outside=outside)


def stamp_cut(params, nbody, size, shape):
    """(x, y) of the size x size cut-out centred on the objects, kept inside a frame of ``shape``."""
    xs, ys = params[0:2 * nbody:2], params[1:2 * nbody:2]
    ox = int(round(float(np.mean(xs)))) - size // 2
    oy = int(round(float(np.mean(ys)))) - size // 2
    ox = min(max(ox, 0), max(shape[1] - size, 0))
    oy = min(max(oy, 0), max(shape[0] - size, 0))
    return ox, oy


def load_epochs(paths, start_fn, nbody=2, size=128, floor_index=None, device="cuda", whole_frame=True, chunk=32):
    """Many epochs into one PixelDomain (the reference runs one process group per image,
    apf_step2.py:154-161): every FITS frame is read, its starting point worked out by
    ``start_fn(image, header, path) -> parameters[P]`` (frame coordinates), and a size x size cut-out
    around the objects is masked, weighted and cut on the device (``lapf_frame_prep``), ``chunk``
    frames at a time through one pinned staging buffer, so the full frames never pile up anywhere.
    Returns (domain, parameters [F, P], cuts [F, 2], headers)."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.LapfError("no CUDA device: olpefit_b200 has no CPU path")
    dev = torch.device(device)
    nf = len(paths)
    data = torch.empty((nf, size, size), dtype=torch.float32, device=dev)
    weight = torch.empty_like(data)
    outside = torch.empty((nf, 3), dtype=torch.float64, device=dev) if whole_frame else None
    params, cuts, headers = [], [], []
    stage = None
    for f0 in range(0, nf, chunk):
        f1 = min(nf, f0 + chunk)
        scal = []
        for f in range(f0, f1):
            image, hdr = read_fits(paths[f])
            if image.ndim != 2:
                raise ValueError("%s: expected a 2-D image, got shape %s" % (paths[f], image.shape))
            if stage is None:
                fy, fx = image.shape
                if fy < size or fx < size:
                    raise ValueError("%s: frame %s is smaller than the %d-pixel stamp" % (paths[f], image.shape, size))
                stage = torch.empty((chunk, fy, fx), dtype=torch.float32).pin_memory()
                stage_d = torch.empty((chunk, fy, fx), dtype=torch.float32, device=dev)
                done = torch.cuda.Event()
            elif image.shape != (fy, fx):
                raise ValueError("%s: frame shape %s differs from %s" % (paths[f], image.shape, (fy, fx)))
            p = np.asarray(start_fn(image, hdr, paths[f]), dtype=np.float64)
            params.append(p)
            cuts.append(stamp_cut(p, nbody, size, image.shape))
            headers.append(hdr)
            scal.append((float(saturation_level(hdr)), float(read_noise(hdr))))
            if f == f0:
                done.synchronize()                      # the previous chunk has left the staging buffer
            stage[f - f0].copy_(torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)))
        n = f1 - f0
        stage_d[:n].copy_(stage[:n], non_blocking=True)
        done.record(torch.cuda.current_stream(dev))
        cut_t = torch.as_tensor(np.asarray(cuts[f0:f1], dtype=np.int32)).to(dev)
        g0 = 0
        while g0 < n:                                   # runs of frames with equal header scalars share a launch
            g1 = g0 + 1
            while g1 < n and scal[g1] == scal[g0]:
                g1 += 1
            sat, rn = scal[g0]
            _lib.check(lib.lapf_frame_prep(stage_d[g0:].data_ptr(), g1 - g0, fy, fx, cut_t[g0:].data_ptr(), size, size,
                                           sat, rn, data[f0 + g0:].data_ptr(), weight[f0 + g0:].data_ptr(), _stream_ptr(dev)))
            if whole_frame:
                _lib.check(lib.lapf_frame_outside(stage_d[g0:].data_ptr(), g1 - g0, fy, fx, cut_t[g0:].data_ptr(), size,
                                                  size, sat, rn, outside[f0 + g0:].data_ptr(), _stream_ptr(dev)))
            g0 = g1
    torch.cuda.synchronize(dev)
    cuts = np.asarray(cuts, dtype=np.int32).reshape(nf, 2)
    dom = PixelDomain(data, weight, cuts, nbody=nbody, floor_index=floor_index, device=device, outside=outside)
    return dom, np.asarray(params), cuts, headers


def initial_parameters(image, guess, nbody=2, origin=(0, 0)):
    """Starting point from the step-1 file (apf_step2.py:258-273; 3body/apf_step2_3body.py:252-265).
    ``guess`` holds the numbers of <N>_initialguess in frame coordinates; ``image`` is the frame
    as read from FITS, or a cut-out whose pixel [0][0] sits at frame coordinates ``origin``."""
    g = np.asarray(guess, dtype=np.float64)
    ox, oy = int(origin[0]), int(origin[1])
    sigma = (50.0 / 9.95) / 2.35                                   # apf_step2.py:242-245

    def pix(y, x):
        return image[int(y) - oy, int(x) - ox]

    def sky(bx, by):
        return np.median(image[int(by) - oy:int(by) - oy + 10, int(bx) - ox:int(bx) - ox + 10])

    if nbody == 2:
        xcs, ycs, xcc, ycc = g[0], g[1], g[2], g[3]
        amps, ampc = pix(ycs - 1, xcs - 1), pix(ycc - 1, xcc - 1)   # :267-268
        bkgd = sky(g[4], g[5])                                      # :270-271
        return np.array([xcs, ycs, xcc, ycc, 0., 0., amps, ampc, 0.2, bkgd, sigma, sigma, sigma * 3,
                         sigma * 3, 0., 0.], dtype=np.float64)
    xca, yca, xcb, ycb, xcc, ycc = g[:6]
    ampa = pix(yca - 0.5 + 1, xca - 0.5 + 1)                        # 3body :258-260
    ampb = pix(ycb - 0.5 + 1, xcb - 0.5 + 1)
    ampc = pix(ycc - 0.5 + 1, xcc - 0.5 + 1)
    bkgd = sky(g[6], g[7])
    return np.array([xca, yca, xcb, ycb, xcc, ycc, 0., 0., ampa, ampb, ampc, 0.2, bkgd, sigma, sigma,
                     sigma * 3, sigma * 3, 0., 0.], dtype=np.float64)
