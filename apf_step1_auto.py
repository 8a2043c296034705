#!/usr/bin/env python
"""LAPF step 1 without the clicks: writes <dir>/<N>_initialguess for apf_step2 / apf_step2a.

usage: python apf_step1_auto.py IMAGE --star X,Y --companion X,Y [--companion X,Y] --sky X,Y [--no-refine]
(the reference's apf_step1.py is interactive; positions are 0-based pixel coordinates as displayed there)
"""
import argparse
import sys

from olpefit_b200 import frame, step1


def _xy(text):
    x, y = text.split(",")
    return float(x), float(y)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("image")
    ap.add_argument("--star", type=_xy, required=True)
    ap.add_argument("--companion", type=_xy, action="append", required=True)
    ap.add_argument("--sky", type=_xy, required=True)
    ap.add_argument("--no-refine", action="store_true", help="keep the given positions instead of the brightest pixel nearby")
    a = ap.parse_args(argv)
    image, _ = frame.read_fits(a.image)
    numbers = step1.initial_guess(image, [a.star] + a.companion, a.sky, refine=not a.no_refine)
    print(step1.write_initial_guess(a.image, numbers), *numbers)
    return 0


if __name__ == "__main__":
    sys.exit(main())
