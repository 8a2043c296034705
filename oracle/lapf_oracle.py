"""CPU oracle for the LAPF apf_step2 / apf_step2a / apf_step2_3body hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import it.
The product path (``olpefit_b200``) never imports anything from ``oracle/``
and fails loudly when the CUDA library is missing.

What it is: a float64 numpy restatement of the reference algorithm.  Every
function cites the reference lines (relative to /root/reference) it follows.

How it is pinned: the reference cannot run as a whole in this image (Python-2
syntax, astropy and mpi4py absent), and it ships no tests or golden vectors.
``tools/make_golden.py`` therefore executes the reference's OWN function
bodies (source lines apf_step2.py:63-70,78-148 and 3body/apf_step2_3body.py:
78-130, which are Python-3 clean) with a stand-in for the one missing
third-party symbol, ``astropy.modeling.models.Gaussian2D``, and commits the
outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this
oracle against those vectors.  The stand-in restates astropy's published
``Gaussian2D.evaluate`` formula (astropy/modeling/functional_models.py; the
reference pins no version -- Python-2 era, astropy <= 2.0.x -- and the formula
is unchanged from astropy 1.0 through 6.x).  So: model assembly, chi-square
and the accept rule are pinned to reference-executed outputs; the elliptical
Gaussian itself is pinned only to the restated formula plus the analytic
identities in tests/test_oracle.py ("parity unpinned" for that one symbol).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------
# Parameter layouts and proposal tables
# --------------------------------------------------------------------------


@dataclass(frozen=True)
class Layout:
    """Index map of the parameter vector.

    2-body: apf_step2.py:108,215-217,234,273.
    3-body: 3body/apf_step2_3body.py:108-109,220-238,265-295.
    """

    nbody: int
    names: tuple
    widths: tuple
    lognorm: tuple  # indices proposed in log10 space (apf_step2.py:217)
    floor_index: int  # slot used as the constant floor (apf_step2.py:119-120)
    bkgd_index: int

    @property
    def nparam(self) -> int:
        return len(self.names)

    @property
    def ncomp(self) -> int:
        return 2 * self.nbody

    def pos(self, obj: int):
        return 2 * obj, 2 * obj + 1

    @property
    def i_dx(self):
        return 2 * self.nbody

    @property
    def i_dy(self):
        return 2 * self.nbody + 1

    def i_amp(self, obj: int):
        return 2 * self.nbody + 2 + obj

    @property
    def i_ratio(self):
        return 3 * self.nbody + 2

    @property
    def i_sigma(self):  # sigmax, sigmay, sigmax2, sigmay2
        b = 3 * self.nbody + 4
        return b, b + 1, b + 2, b + 3

    @property
    def i_theta(self):
        b = 3 * self.nbody + 8
        return b, b + 1

    def is_log(self):
        out = np.zeros(self.nparam, dtype=bool)
        out[list(self.lognorm)] = True
        return out


LAYOUT_2BODY = Layout(
    nbody=2,
    names=("xcs", "ycs", "xcc", "ycc", "dx", "dy", "amps", "ampc", "ampratio", "bkgd",
           "sigmax", "sigmay", "sigmax2", "sigmay2", "theta", "theta2"),
    # apf_step2.py:234
    widths=(0.01, 0.01, 0.3, 0.3, 0.08, 0.09, 0.0025, 0.02, 0.001, 0.0008,
            0.002, 0.002, 0.001, 0.001, 0.008, 0.01),
    lognorm=(6, 7, 9, 10, 11, 12, 13),  # apf_step2.py:217
    floor_index=12,  # apf_step2.py:120 fills the floor with p[12] (sigmax2): reference behaviour
    bkgd_index=9,
)

LAYOUT_3BODY = Layout(
    nbody=3,
    names=("xca", "yca", "xcb", "ycb", "xcc", "ycc", "dx", "dy", "ampa", "ampb", "ampc",
           "ampratio", "bkgd", "sigmax", "sigmay", "sigmax2", "sigmay2", "theta", "theta2"),
    # 3body/apf_step2_3body.py:220-238
    widths=(0.01, 0.01, 0.3, 0.3, 0.3, 0.3, 0.08, 0.09, 0.0025, 0.02, 0.02, 0.001, 0.0008,
            0.002, 0.002, 0.001, 0.001, 0.008, 0.01),
    lognorm=(8, 9, 10, 12, 13, 14, 15, 16),  # 3body/apf_step2_3body.py:295
    floor_index=12,  # 3body/apf_step2_3body.py:121: p[12] is bkgd in this layout
    bkgd_index=12,
)


def layout_for(nbody: int) -> Layout:
    if nbody == 2:
        return LAYOUT_2BODY
    if nbody == 3:
        return LAYOUT_3BODY
    raise ValueError("nbody must be 2 or 3")


# NIRC2 PSF guess: apf_step2.py:242-245
FWHM_MAS = 50.0
PIXSCALE_STEP2 = 9.95
SIGMA_GUESS = (FWHM_MAS / PIXSCALE_STEP2) / 2.35

# --------------------------------------------------------------------------
# Pixel model
# --------------------------------------------------------------------------


def gaussian2d(x, y, amplitude, x_mean, y_mean, x_stddev, y_stddev, theta):
    """Rotated elliptical Gaussian: astropy ``Gaussian2D.evaluate`` restated.

    Third-party arithmetic invoked by the reference at apf_step2.py:98-102.
    ``x`` is the column index, ``y`` the row index.
    """
    ct = math.cos(theta)
    st = math.sin(theta)
    s2t = math.sin(2.0 * theta)
    vx = x_stddev * x_stddev
    vy = y_stddev * y_stddev
    a = 0.5 * (ct * ct / vx + st * st / vy)
    b = 0.5 * (s2t / vx - s2t / vy)
    c = 0.5 * (st * st / vx + ct * ct / vy)
    ddx = x - x_mean
    ddy = y - y_mean
    return amplitude * np.exp(-(a * ddx * ddx + b * ddx * ddy + c * ddy * ddy))


def psf_object(x, y, xc, yc, dx, dy, total, ratio, bkgd, sx, sy, sx2, sy2, th, th2):
    """Narrow core + wide wing of one object (apf_step2.py:78-103)."""
    amp = total - bkgd            # :95
    amp_wide = amp * ratio        # :96
    amp_narrow = amp - amp_wide   # :97
    core = gaussian2d(x, y, amp_narrow, xc, yc, sx, sy, th)                # :98-99
    wing = gaussian2d(x, y, amp_wide, xc + dx, yc + dy, sx2, sy2, th2)     # :100-101
    return wing + core            # :102


def pixel_grid(ny, nx, origin=(0, 0)):
    """Row/column coordinate grids (apf_step2.py:94); origin = (x0, y0) of the cut-out."""
    yy, xx = np.mgrid[:ny, :nx]
    return xx + origin[0], yy + origin[1]


def model_image(p, layout: Layout, ny, nx, origin=(0, 0), floor_index=None, grid=None):
    """Full model on an (ny, nx) grid.

    2-body: apf_step2.py:106-124 (floor = p[12], reference behaviour).
    3-body: 3body/apf_step2_3body.py:106-125.
    ``floor_index`` overrides the slot used for the floor (the ``--fix-bkgd`` opt-in).
    """
    p = np.asarray(p, dtype=np.float64)
    xx, yy = grid if grid is not None else pixel_grid(ny, nx, origin)
    fi = layout.floor_index if floor_index is None else floor_index
    sx, sy, sx2, sy2 = (p[i] for i in layout.i_sigma)
    th, th2 = (p[i] for i in layout.i_theta)
    total = np.zeros((ny, nx), dtype=np.float64)
    for obj in range(layout.nbody):
        ix, iy = layout.pos(obj)
        total = total + psf_object(xx, yy, p[ix], p[iy], p[layout.i_dx], p[layout.i_dy],
                                   p[layout.i_amp(obj)], p[layout.i_ratio], p[layout.bkgd_index],
                                   sx, sy, sx2, sy2, th, th2)
    return total + p[fi]


def chi_squared(data, model, err, mask=None):
    """Sum of squared, noise-weighted residuals over unmasked pixels.

    apf_step2.py:134-137 with the masked array of :188 (masked terms are skipped).
    """
    t = ((data - model) / err) ** 2
    if mask is not None:
        t = np.where(mask, 0.0, t)
    return float(np.sum(t))


def chi_squared_weighted(data, model, weight):
    """Same quantity written with w = 1/err^2 and w = 0 on masked pixels."""
    r = data - model
    return float(np.sum(weight * r * r))


# --------------------------------------------------------------------------
# Frame preparation
# --------------------------------------------------------------------------


def saturation_level(header):
    """apf_step2.py:176-185."""
    itime = float(header["itime"]) * 1000.0
    coadds = float(header["coadds"])
    multisam = float(header["multisam"])
    sampmode = header["sampmode"]
    if sampmode == 3:
        return coadds * 24000.0 * (1.0 - 0.1 * (multisam - 1.0) / (itime / 1000.0))
    return coadds * 22000.0


def read_noise(header):
    """apf_step2.py:197-204."""
    coadds = float(header["coadds"])
    multisam = float(header["multisam"])
    if header["sampmode"] == 3.0:
        return (38.0 / math.sqrt(multisam)) * math.sqrt(coadds)
    return 38.0 * math.sqrt(coadds)


def frame_prep(image, header):
    """Saturation mask and per-pixel error map (apf_step2.py:176-210).

    Returns (mask, err): mask True where image > 0.8*satlevel.
    """
    image = np.asarray(image, dtype=np.float64)
    mask = image > 0.8 * saturation_level(header)        # :188
    rn = read_noise(header)
    pois = np.sqrt(np.abs(image))                         # :207
    err = np.sqrt(rn * rn + pois * pois)                  # :210
    return mask, err


def weight_map(image, header):
    """w = 1/err^2, zero where masked: the device-side form of (mask, err)."""
    mask, err = frame_prep(image, header)
    w = 1.0 / (err * err)
    w[mask] = 0.0
    return w


# --------------------------------------------------------------------------
# Initial state
# --------------------------------------------------------------------------


def initial_parameters(image, guess, layout: Layout):
    """Initial parameter vector from the step-1 file.

    2-body: apf_step2.py:258-273.  3-body: 3body/apf_step2_3body.py:252-265.
    """
    g = np.asarray(guess, dtype=np.float64)
    s = SIGMA_GUESS
    if layout.nbody == 2:
        xcs, ycs, xcc, ycc = g[0], g[1], g[2], g[3]
        amps = image[int(ycs - 1), int(xcs - 1)]
        ampc = image[int(ycc - 1), int(xcc - 1)]
        box = image[int(g[5]):int(g[5]) + 10, int(g[4]):int(g[4]) + 10]
        bk = np.median(box)
        return np.array([xcs, ycs, xcc, ycc, 0.0, 0.0, amps, ampc, 0.2, bk, s, s, 3 * s, 3 * s,
                         0.0, 0.0], dtype=np.float64)
    xa, ya, xb, yb, xc, yc = g[:6]
    ampa = image[int(ya - 0.5 + 1), int(xa - 0.5 + 1)]
    ampb = image[int(yb - 0.5 + 1), int(xb - 0.5 + 1)]
    ampc = image[int(yc - 0.5 + 1), int(xc - 0.5 + 1)]
    box = image[int(g[7]):int(g[7]) + 10, int(g[6]):int(g[6]) + 10]
    bk = np.median(box)
    return np.array([xa, ya, xb, yb, xc, yc, 0.0, 0.0, ampa, ampb, ampc, 0.2, bk, s, s, 3 * s, 3 * s,
                     0.0, 0.0], dtype=np.float64)


# --------------------------------------------------------------------------
# Random streams
# --------------------------------------------------------------------------

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
PHILOX_TAG = 0x4C415046  # 'LAPF'


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11; Random123).  counter: 4 uint32, key: 2 uint32."""
    c0, c1, c2, c3 = (int(v) & 0xFFFFFFFF for v in counter)
    k0, k1 = (int(v) & 0xFFFFFFFF for v in key)
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & 0xFFFFFFFF
        hi1, lo1 = p1 >> 32, p1 & 0xFFFFFFFF
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & 0xFFFFFFFF, lo1, (hi0 ^ c3 ^ k1) & 0xFFFFFFFF, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def device_draws(seed, walker, t, nparam):
    """The (index, standard normal, uniform) triple the CUDA sampler uses for update ``t``.

    Stream definition (shared with make_draw in olpefit_b200/csrc/lapf_device.cuh): key = (seed low 32 bits,
    walker id); counter = (t low, t high, seed high 32 bits, 'LAPF').  Draw order follows the
    reference: index (apf_step2.py:302), normal (:64/:68), uniform (:143).
    """
    seed = int(seed)
    r0, r1, r2, r3 = philox4x32_10(
        (t & 0xFFFFFFFF, (t >> 32) & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, PHILOX_TAG),
        (seed & 0xFFFFFFFF, walker & 0xFFFFFFFF))
    k = (r0 * nparam) >> 32
    u1 = (r1 + 1) * 2.0 ** -32                 # (0, 1]: Box-Muller from two full words
    u2 = r2 * 2.0 ** -32                       # [0, 1)
    z = math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)
    # 53-bit uniform like numpy's rand() (apf_step2.py:143): the fourth word and the low 21 bits
    # of the first (its high bits picked the parameter)
    u = ((r3 << 21) | (r0 & 0x1FFFFF)) * 2.0 ** -53          # [0, 1)
    return k, z, u


class NumpyStream:
    """The reference's stream: numpy's global-style Mersenne Twister (apf_step2.py:64,68,143,302)."""

    def __init__(self, seed=None):
        self.rs = np.random.RandomState(seed)

    def draw_index(self, t, nparam):
        return int(self.rs.randint(0, nparam))

    def draw_normal(self, t, loc, scale):
        return float(self.rs.normal(loc, scale, 1)[0])

    def draw_uniform(self, t):
        return float(self.rs.rand())


class PhiloxStream:
    """Replays the device stream so CPU and GPU chains can be compared update by update."""

    def __init__(self, seed, walker, nparam):
        self.seed, self.walker, self.nparam = seed, walker, nparam
        self._t = None
        self._cur = None

    def _get(self, t):
        if self._t != t:
            self._cur = device_draws(self.seed, self.walker, t, self.nparam)
            self._t = t
        return self._cur

    def draw_index(self, t, nparam):
        return self._get(t)[0]

    def draw_normal(self, t, loc, scale):
        return loc + scale * self._get(t)[1]

    def draw_uniform(self, t):
        return self._get(t)[2]


# --------------------------------------------------------------------------
# Sampler
# --------------------------------------------------------------------------


def propose(stream, t, value, width, is_log):
    """apf_step2.py:63-70 -- normal, or normal in log10 space (non-positive value -> nan)."""
    if not is_log:
        return stream.draw_normal(t, value, width)
    with np.errstate(all="ignore"):
        lv = np.log10(value)
        # arrays of one element, like the reference: numpy's vector pow can differ from
        # Python's scalar pow by one ulp
        return float((10 ** np.array([stream.draw_normal(t, lv, width)]))[0])


def accept_rule(stream, t, chi_cur, chi_prop):
    """apf_step2.py:139-148: accept iff u < exp(-(chi_prop - chi_cur)/2); nan compares false."""
    with np.errstate(all="ignore"):
        p_accept = np.exp(-(chi_prop - chi_cur) / 2.0)
    u = stream.draw_uniform(t)
    return bool(u < p_accept)


@dataclass
class ChainResult:
    rows: np.ndarray          # [n_rows, P+1]; row 0 is all-nan like the reference file
    tries: np.ndarray
    accepts: np.ndarray
    params: np.ndarray        # final parameter vector
    chi2: float
    n_updates: int
    trace: list = field(default_factory=list)


def run_chain(data, weight, layout: Layout, p0, stream, *, origin=(0, 0), n_updates=None,
              accept_min=None, burn_in=0, thin=1, floor_index=None, widths=None,
              record_trace=False, chi0=None, chi2_fn=None):
    """One walker of the reference loop (apf_step2.py:276-351; step 2a: apf_step2a.py:271-337).

    ``data``/``weight`` are the pixel domain (full frame or a cut-out whose lower-left pixel
    is ``origin`` in frame coordinates; parameters stay in frame coordinates).  Stops after
    ``n_updates`` (step 2a rule) or when min(tries) >= accept_min (step 2 rule, :300).
    A chain row is appended after every ``thin``-th update once count >= burn_in (:342-351),
    whether or not the proposal was accepted.  ``chi2_fn(p) -> float`` replaces the
    build_analytical_model + chi_squared pair (:314-316), which is how the tests drive the device
    operator from the reference's host loop.
    """
    ny, nx = data.shape
    grid = pixel_grid(ny, nx, origin)
    p = np.array(p0, dtype=np.float64)
    npar = layout.nparam
    is_log = layout.is_log()
    w = np.asarray(layout.widths if widths is None else widths, dtype=np.float64)
    tries = np.zeros(npar)
    accepts = np.zeros(npar)
    if chi2_fn is None:
        def chi2_fn(q):
            return chi_squared_weighted(data, model_image(q, layout, ny, nx, grid=grid,
                                                          floor_index=floor_index), weight)
    chi = chi2_fn(p) if chi0 is None else chi0
    rows = [np.full(npar + 1, np.nan)]                       # :278-279
    trace = []
    count = 0
    while True:
        if n_updates is not None and count >= n_updates:
            break
        if n_updates is None and tries.min() >= accept_min:
            break
        k = stream.draw_index(count, npar)                   # :302
        tries[k] += 1                                        # :304
        new = propose(stream, count, p[k], w[k], is_log[k])  # :306-309
        trial = p.copy()                                     # :312-313
        trial[k] = new
        with np.errstate(all="ignore"):
            chi_t = chi2_fn(trial)                           # :314-316
        ok = accept_rule(stream, count, chi, chi_t)          # :318
        if ok:                                               # :321-327
            accepts[k] += 1
            p[k] = new
            chi = chi_t
        if record_trace:
            trace.append((k, new, chi_t, ok))
        count += 1                                           # :333
        if count >= burn_in and (count - burn_in) % thin == 0:   # :342-351
            rows.append(np.concatenate([p, [chi]]))
    return ChainResult(np.array(rows), tries, accepts, p, chi, count, trace)


# --------------------------------------------------------------------------
# Step-3 statistics used for posterior parity (apf_step3.py)
# --------------------------------------------------------------------------

PIXSCALE_PRE2015 = 9.952   # apf_step3.py:224
PIXSCALE_POST2015 = 9.971  # apf_step3.py:231


def separation_pa(xcs, ycs, xcc, ycc, pixscale=PIXSCALE_PRE2015):
    """apf_step3.py:255-256,283-291 without the distortion lookup (tables absent from the checkout)."""
    dy = np.asarray(ycc) - np.asarray(ycs)
    dx = np.asarray(xcc) - np.asarray(xcs)
    sep = np.sqrt(dy * dy + dx * dx) * pixscale
    pa = np.degrees(np.arctan2(-dx, dy))
    return sep, pa


def gelman_rubin(chains, python2_division=True):
    """apf_step3.py:262-276.  ``chains``: [n_rows, n_walkers] for one parameter.

    The reference evaluates (d+3)/(d+1) with d = 16 under Python 2, i.e. integer division = 1.
    """
    chains = np.asarray(chains, dtype=np.float64)
    n, m = float(chains.shape[0]), float(chains.shape[1])
    overall = np.mean(chains)
    wv = np.std(chains, axis=0) ** 2
    bv = (np.mean(chains, axis=0) - overall) ** 2
    w = (1.0 / m) * np.sum(wv)
    b = (n / (m - 1.0)) * np.sum(bv)
    pooled = ((n - 1.0) / n) * w + ((m + 1.0) / (m * n)) * b
    psrf = pooled / w
    d = 16
    factor = float((d + 3) // (d + 1)) if python2_division else (d + 3) / (d + 1)
    return psrf, math.sqrt(factor * psrf)
