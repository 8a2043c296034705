#!/usr/bin/env python
"""LAPF step 2 on B200 -- drop-in for the reference's apf_step2.py (same arguments, inputs, outputs).

usage: python apf_step2.py IMAGE [-i {1,2a}] [--walkers N] ...          (one GPU)
       python -m torch.distributed.run --nproc-per-node G apf_step2.py IMAGE ...   (G GPUs)
"""
import sys

from olpefit_b200.cli import main_step2

if __name__ == "__main__":
    sys.exit(main_step2())
