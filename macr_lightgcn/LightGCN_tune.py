#!/usr/bin/env python
"""Drop-in for the reference's `python macr_lightgcn/LightGCN_tune.py ...`: the training run of
LightGCN.py with a sweep of c over np.linspace(--start, --end, --step) at every evaluation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macr_b200.cli.lightgcn import main  # noqa: E402

if __name__ == "__main__":
    main(tune=True)
