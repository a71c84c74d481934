#!/usr/bin/env python3
"""bench.py's e2e_frames record on its own (AdcDac frames from pinned host memory -> decode + loss + 4 cascades)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

for chunk in [int(v) for v in (sys.argv[1:] or ['32768'])]:
    print(json.dumps(bench.e2e_frames(0, chunk)), flush=True)
