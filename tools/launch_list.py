"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel-name totals of one forward step.
python tools/launch_list.py gpurun_out/x.csv"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    out = [(r[ki][:64], float(r[vi].replace(",", "")) / 1000) for r in rows[1:]]
    idx = [i for i, (k, _) in enumerate(out) if "nchw_to_nhwc" in k]
    step = out[idx[0]:idx[1]] if len(idx) > 1 else out
    agg, cnt = collections.Counter(), collections.Counter()
    for k, v in step:
        agg[k] += v
        cnt[k] += 1
    print(f"{len(step)} launches, {sum(v for _, v in step):.1f} us")
    for k, v in agg.most_common():
        print(f"{v:9.1f} us  {cnt[k]:4d} x {v / cnt[k]:7.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
