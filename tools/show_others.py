"""Print the `others` block of a bench.py JSON line read from stdin (scratch helper)."""
import json
import sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(tag, "value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4))
for k, v in d.get("others", {}).items():
    if isinstance(v, dict):
        print(tag, k[:14], {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("denominator",)})
