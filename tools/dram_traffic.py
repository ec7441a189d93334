"""Sum dram__bytes_read/write over the launches of ONE forward step of an ncu csv
(ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none ... bench.py --steps 2 --no-extras)
and write profiles/r02_dram_traffic.json, the source of bench.py's roofline.traffic.
    python tools/dram_traffic.py gpurun_out/r02_dram.csv"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    d = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            d.setdefault(int(r[ii]), {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    ids = list(d)
    starts = [i for i in ids if "nchw_to_nhwc" in d[i]["k"]]
    step = [i for i in ids if starts[0] <= i < starts[1]] if len(starts) > 1 else ids
    rd = sum(d[i].get("dram__bytes_read.sum", 0) for i in step)
    wr = sum(d[i].get("dram__bytes_write.sum", 0) for i in step)
    out = {"bytes_per_step": rd + wr, "read": rd, "write": wr, "launches": len(step),
           "note": f"dram__bytes_read.sum + dram__bytes_write.sum over the {len(step)} launches of ONE cfg2 forward ({rd/1e9:.2f} GB read + "
                   f"{wr/1e9:.2f} GB written), ncu --cache-control none, {os.path.basename(path)}; recorded by tools/dram_traffic.py, not measured "
                   "in this run"}
    with open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1])
