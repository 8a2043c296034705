#!/usr/bin/env python
"""LAPF step 2a on B200 -- drop-in for the reference's apf_step2a.py: one walker, n_steps updates,
writes step2a.csv whose last row seeds `apf_step2.py IMAGE -i 2a`."""
import sys

from olpefit_b200.cli import main_step2a

if __name__ == "__main__":
    sys.exit(main_step2a())
