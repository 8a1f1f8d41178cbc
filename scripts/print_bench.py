#!/usr/bin/env python
"""One-line digest of bench.py JSON lines."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f" % d["ms_per_step"], "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"], 3),
              {k: round(v["ms_per_step"], 4) for k, v in d.get("stages", {}).items()})
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
