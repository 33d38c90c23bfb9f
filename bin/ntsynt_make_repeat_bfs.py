#!/usr/bin/env python3
"Drop-in for the reference's `ntsynt_make_repeat_bfs.py` (see INTEGRATION.md); runs the B200-native path."
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), ".."))
from ntsynt_b200 import cli  # noqa: E402

if __name__ == "__main__":
    sys.exit(cli.main_make_repeat_bfs())
