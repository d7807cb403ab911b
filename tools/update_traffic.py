"""Refresh profiles/traffic.json from `ncu --set full` raw CSV pages of the two headline kernels.
usage: python tools/update_traffic.py <row_raw.csv> <all_raw.csv> <label>
The figure is dram__bytes_read.sum + dram__bytes_write.sum of ONE launch; the file is stamped with the hash of the
reduction kernels' sources (bench.source_stamp) so that bench.py can say whether the figure belongs to the current build."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def one(path):
    rows = list(csv.reader(open(path)))
    hdr, units, val = rows[0], rows[1], rows[2]
    g = lambda n: float(val[hdr.index(n)]) * UNIT.get(units[hdr.index(n)], 1.0)
    return {"kernel": val[hdr.index("Kernel Name")], "dram_read_bytes": g("dram__bytes_read.sum"), "dram_write_bytes": g("dram__bytes_write.sum"),
            "duration_us_under_ncu": float(val[hdr.index("gpu__time_duration.sum")]), "registers": int(float(val[hdr.index("launch__registers_per_thread")])),
            "grid": int(float(val[hdr.index("launch__grid_size")]))}


if __name__ == "__main__":
    import bench
    row, al = one(sys.argv[1]), one(sys.argv[2])
    out = {"reduce_row_block": row["dram_read_bytes"] + row["dram_write_bytes"], "reduce_all": al["dram_read_bytes"] + al["dram_write_bytes"],
           "source": f"{sys.argv[3]} (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, one launch each, inside `python bench.py --steps 2 --warmup 1`)",
           "source_stamp": bench.source_stamp(), "details": {"reduce_row_block": row, "reduce_all": al}}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
