#!/usr/bin/env python
"""Drop-in for the reference's `python ./macr_mf/train.py ...` (README.md:30-52): same flags,
executed by the B200-native path in macr_b200/."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macr_b200.cli.train_mf import main  # noqa: E402

if __name__ == "__main__":
    main()
