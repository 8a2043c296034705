"""Parameter layouts of the reference (apf_step2.py:108; 3body/apf_step2_3body.py:266-288),
served from liblapf so Python and CUDA cannot disagree."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

NAMES_2BODY = ("xcs", "ycs", "xcc", "ycc", "dx", "dy", "amps", "ampc", "ampratio", "bkgd",
               "sigmax", "sigmay", "sigmax2", "sigmay2", "theta", "theta2")
NAMES_3BODY = ("xca", "yca", "xcb", "ycb", "xcc", "ycc", "dx", "dy", "ampa", "ampb", "ampc",
               "ampratio", "bkgd", "sigmax", "sigmay", "sigmax2", "sigmay2", "theta", "theta2")

# slot the reference adds as the constant floor: p[12] in both scripts (apf_step2.py:120 --
# which is sigmax2 there -- and 3body/apf_step2_3body.py:121, bkgd)
REFERENCE_FLOOR_INDEX = 12


def nparam(nbody: int) -> int:
    return _lib.check(_lib.load().lapf_num_params(int(nbody)))


def names(nbody: int):
    return NAMES_2BODY if nbody == 2 else NAMES_3BODY


def bkgd_index(nbody: int) -> int:
    return 3 * nbody + 3


def default_widths(nbody: int):
    """(widths[P], is_log[P]) -- apf_step2.py:215-217,234 / 3body:220-238,292-295."""
    p = nparam(nbody)
    w = (C.c_double * p)()
    lg = (C.c_int32 * p)()
    _lib.check(_lib.load().lapf_default_widths(int(nbody), w, lg))
    return np.array(w[:], dtype=np.float64), np.array(lg[:], dtype=bool)
