"""Non-interactive step 1: the file step 2 starts from, without the matplotlib clicks.

The reference's step 1 (apf_step1.py) shows the frame, takes one click per object and one on empty
sky, refines each object click to the brightest pixel of the 21 x 21 box around it (+0.5: pixel
centre, apf_step1.py:145-163) and writes one line to <dir>/<N>_initialguess (:167-175).  Here the
clicks are given as numbers, so thousands of epochs can be prepared without a display; the
refinement and the file are the reference's.
"""
from __future__ import annotations

import numpy as np

from . import chains


def refine_click(image, x_click, y_click, half=11):
    """apf_step1.py:145-152: the brightest pixel of the box [click-11, click+10] + 0.5."""
    xm, ym = int(x_click), int(y_click)
    y_lo, x_lo = max(ym - half, 0), max(xm - half, 0)
    box = np.asarray(image)[y_lo:ym - half + 21, x_lo:xm - half + 21]
    iy, ix = np.unravel_index(np.argmax(box), box.shape)          # findmax, apf_step1.py:57-61
    return x_lo + ix + 0.5, y_lo + iy + 0.5


def initial_guess(image, objects, sky, refine=True):
    """``objects``: [(x, y), ...] approximate positions, star first (2 entries: 2-body, 3: 3-body);
    ``sky``: (x, y) of an empty region.  Returns the numbers of the step-1 file."""
    out = []
    for (x, y) in objects:
        out += list(refine_click(image, x, y)) if refine else [float(x), float(y)]
    out += [int(sky[0]), int(sky[1])]
    return out


def write_initial_guess(image_path, numbers):
    """apf_step1.py:167-175 / 3body/apf_step1_3body.py:196-199: one space-separated line."""
    path = chains.initial_guess_path(image_path)
    with open(path, "w") as fh:
        fh.write(" ".join(str(v) for v in numbers) + "\n")
    return path
