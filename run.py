#!/usr/bin/env python
"""Drop-in for the reference's examples/cpu/inference/python/llm/run.py (same flags); see
isca-2025-lia_b200/run.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lia_b200  # noqa: E402,F401
from lia_b200.run import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
