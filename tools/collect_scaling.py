"""Collect bench.py JSON lines of several N into one scaling summary (the driver computes its own efficiency from the
per-N values; this file is the builder's copy of the same arithmetic).
usage: python tools/collect_scaling.py out.json "how ..." N=file [N=file ...]"""
import json
import sys


def line_of(path):
    return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])


out, how, runs = sys.argv[1], sys.argv[2], {}
lines = {int(a.split("=")[0]): line_of(a.split("=")[1]) for a in sys.argv[3:]}
base = lines[1]["value"]
for n in sorted(lines):
    d = lines[n]
    runs[f"N={n}"] = {"value_GBs": d["value"], "ms_per_step": d["ms_per_step"], "efficiency_vs_N1": d["value"] / (base * n),
                      "e2e_GBs": d["e2e"]["value"], "e2e_warm_GBs": d["e2e_warm"]["value"], "parity_full_sum_rel_err": d.get("parity_full_sum_rel_err"),
                      "same_bits_on_every_rank": d.get("same_bits_on_every_rank"), "timed_region": d.get("timed_region"), "clocks": d.get("clocks"),
                      "roofline": {k: d["roofline"][k] for k in ("achieved", "frac", "kernel")}, "per_rank": d.get("per_rank"),
                      "others": {k: v for k, v in d.get("others", {}).items() if isinstance(v, dict) and ("us" in v or "ms" in v or "ms_per_step" in v)}}
json.dump({"how": how, "runs": runs}, open(out, "w"), indent=1)
for k, v in runs.items():
    print(k, round(v["value_GBs"]), "GB/s", round(v["ms_per_step"] * 1e3, 2), "us/step", "eff", round(v["efficiency_vs_N1"], 3), "e2e", round(v["e2e_GBs"], 1))
