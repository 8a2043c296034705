#!/usr/bin/env python
"""LAPF step 2, three-body model, on B200 -- drop-in for the reference's 3body/apf_step2_3body.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from olpefit_b200.cli import main_step2_3body  # noqa: E402

if __name__ == "__main__":
    sys.exit(main_step2_3body())
