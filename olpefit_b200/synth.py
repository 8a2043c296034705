"""Synthetic NIRC2-like frames for tests and benchmarks.

This is a DATA generator, not part of the fitting path: it renders a noiseless truth
image once per frame in float64 numpy and adds Poisson + read noise.  The sampler never
calls it.  The reference ships no sample data, so every test frame comes from here.

Truth convention: the 2-body reference fills the constant floor with p[12] (sigmax2,
apf_step2.py:119-120); frames are generated with the same convention so the truth is
inside the model family.
"""
from __future__ import annotations

import math

import numpy as np

SIGMA = (50.0 / 9.95) / 2.35  # apf_step2.py:242-245

HEADER = {"ITIME": 1.0, "COADDS": 1, "MULTISAM": 1, "SAMPMODE": 2}

# x, y of the star and companion(s), 0-based pixel coordinates in the 1024x1024 frame
STAR_XY = (512.3, 511.7)
COMP_XY = (521.7, 519.2)
COMP2_XY = (505.1, 522.4)
SKY_LEVEL = 50.0
READ_NOISE = 38.0


def truth_parameters(nbody=2, epoch=0):
    """Truth parameter vector in the reference layout (frame coordinates)."""
    cx = COMP_XY[0] + 0.01 * epoch
    cy = COMP_XY[1] - 0.005 * epoch
    shape = [SIGMA, 1.1 * SIGMA, 3.0 * SIGMA, 3.2 * SIGMA, 0.3, 0.5]
    if nbody == 2:
        return np.array([STAR_XY[0], STAR_XY[1], cx, cy, 0.3, -0.2, 15000.0, 300.0, 0.2, 6.4]
                        + shape, dtype=np.float64)
    if nbody == 3:
        return np.array([STAR_XY[0], STAR_XY[1], cx, cy, COMP2_XY[0], COMP2_XY[1], 0.3, -0.2,
                         15000.0, 300.0, 150.0, 0.2, 6.4] + shape, dtype=np.float64)
    raise ValueError("nbody must be 2 or 3")


def _ellipse(xx, yy, amp, x0, y0, sx, sy, th):
    ct, st = math.cos(th), math.sin(th)
    u = (xx - x0) * ct + (yy - y0) * st
    v = -(xx - x0) * st + (yy - y0) * ct
    return amp * np.exp(-0.5 * ((u / sx) ** 2 + (v / sy) ** 2))


def truth_image(p, nbody, region):
    """Noiseless truth over region = (y0, y1, x0, x1) in frame coordinates."""
    y0, y1, x0, x1 = region
    yy, xx = np.mgrid[y0:y1, x0:x1].astype(np.float64)
    n = nbody
    dx, dy = p[2 * n], p[2 * n + 1]
    ratio, bkgd = p[3 * n + 2], p[3 * n + 3]
    sx, sy, sx2, sy2, th, th2 = p[3 * n + 4: 3 * n + 10]
    floor = p[12]
    img = np.full(xx.shape, floor, dtype=np.float64)
    for o in range(n):
        a = p[2 * n + 2 + o] - bkgd
        img += _ellipse(xx, yy, a - a * ratio, p[2 * o], p[2 * o + 1], sx, sy, th)
        img += _ellipse(xx, yy, a * ratio, p[2 * o] + dx, p[2 * o + 1] + dy, sx2, sy2, th2)
    return img


def make_frame(epoch=0, nbody=2, region=(0, 1024, 0, 1024), hot_pixels=True, dtype=np.float32):
    """One noisy frame (or cut-out): Poisson(truth + sky) - sky + N(0, read noise).

    Returns (image, truth_parameters).  With ``hot_pixels`` a few pixels near the sources are
    driven above 0.8 x saturation so the mask path (apf_step2.py:188) is exercised.
    """
    p = truth_parameters(nbody, epoch)
    rng = np.random.default_rng(epoch)
    truth = truth_image(p, nbody, region)
    img = rng.poisson(truth + SKY_LEVEL).astype(np.float64) - SKY_LEVEL
    img += rng.normal(0.0, READ_NOISE, size=img.shape)
    if hot_pixels:
        y0, y1, x0, x1 = region
        for (hx, hy) in ((517, 505), (508, 516), (526, 524)):
            if x0 <= hx < x1 and y0 <= hy < y1:
                img[hy - y0, hx - x0] = 30000.0
    return img.astype(dtype), p


def stamp_origin(size, nbody=2):
    """Integer (x0, y0) of an S x S cut-out centred on the star-companion midpoint."""
    mx = 0.5 * (STAR_XY[0] + COMP_XY[0])
    my = 0.5 * (STAR_XY[1] + COMP_XY[1])
    return int(round(mx)) - size // 2, int(round(my)) - size // 2


def make_stamps(n_frames, size, nbody=2, dtype=np.float32):
    """``n_frames`` epochs cut to size x size.  Returns (stamps [F,S,S], origins [F,2] (x0,y0))."""
    ox, oy = stamp_origin(size, nbody)
    out = np.empty((n_frames, size, size), dtype=dtype)
    for f in range(n_frames):
        out[f], _ = make_frame(f, nbody, region=(oy, oy + size, ox, ox + size), dtype=dtype)
    origins = np.tile(np.array([[ox, oy]], dtype=np.int32), (n_frames, 1))
    return out, origins


def step1_guess(image, nbody=2, origin=(0, 0), sky_xy=None):
    """A step-1 style initial guess (apf_step1.py:145-163) in frame coordinates: the star at its
    brightest pixel + 0.5 (the reference searches a 21x21 box around the click); a companion at the
    pixel containing it + 0.5 (a faint companion sitting on the star's wing is not the brightest
    pixel of any box, the user's click decides); plus an empty-sky corner for the background box."""
    ox, oy = origin
    work = np.array(image, dtype=np.float64)
    work[work > 0.8 * 22000.0] = -np.inf  # do not lock on to hot pixels
    xm, ym = int(STAR_XY[0]) - ox, int(STAR_XY[1]) - oy
    ylo, xlo = max(ym - 10, 0), max(xm - 10, 0)
    box = work[ylo:ym + 11, xlo:xm + 11]
    iy, ix = np.unravel_index(np.argmax(box), box.shape)
    out = [xlo + ix + 0.5 + ox, ylo + iy + 0.5 + oy]
    for (sx, sy) in [COMP_XY] + ([COMP2_XY] if nbody == 3 else []):
        out += [math.floor(sx) + 0.5, math.floor(sy) + 0.5]
    if sky_xy is None:
        sky_xy = (ox + 1, oy + 1)
    out += [float(int(sky_xy[0])), float(int(sky_xy[1]))]
    return np.array(out, dtype=np.float64)
