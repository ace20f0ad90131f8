#!/usr/bin/env python
"""bench.py's ns_pressure workload at a small size (smoke check of the traced lhs= law on a native lattice)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
import gds_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
print(json.dumps(bench.pressure_workload(gds_b200.Context.default(), n, 3)))
