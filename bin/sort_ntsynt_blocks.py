#!/usr/bin/env python3
"Drop-in for the reference's `visualization_scripts/sort_ntsynt_blocks.py` (see INTEGRATION.md)."
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), ".."))
from ntsynt_b200 import cli  # noqa: E402

if __name__ == "__main__":
    sys.exit(cli.main_sort_blocks())
