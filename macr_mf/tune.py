#!/usr/bin/env python
"""Drop-in for the reference's `python ./macr_mf/tune.py ...` (README.md:101-105): the training
run of train.py with a sweep of c over np.linspace(--start, --end, --step) at every evaluation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macr_b200.cli.train_mf import main  # noqa: E402

if __name__ == "__main__":
    main(tune=True)
