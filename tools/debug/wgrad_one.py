#!/usr/bin/env python
"""One failing wgrad configuration (for compute-sanitizer / HG_WU_DBG experiments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from wgrad_cfg import run
if __name__ == "__main__":
    for cfg in [(1, 48, 16, 32, 64, 1, 1), (1, 48, 32, 65, 64, 1, 1), (1, 32, 16, 65, 64, 1, 1), (1, 48, 16, 8, 64, 1, 1), (1, 48, 16, 5, 64, 1, 1)]:
        run(*cfg)
