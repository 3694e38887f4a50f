#!/usr/bin/env python
"""Print the headline numbers of a bench.py log (last line = the JSON line): tools/show_bench.py LOG [label ...]"""
import json
import sys

try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(*sys.argv[2:], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "eik_ms", round(r.get("avg_launch_ms", 0), 3),
          "e2e", round((d.get("e2e") or {}).get("value", 0)), "launches", d.get("gpu_launches"))
except Exception as e:  # noqa: BLE001
    print(*sys.argv[2:], "FAILED", e, open(sys.argv[1]).read()[-400:])
