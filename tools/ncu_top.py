"""Summarise an .ncu-rep: per-kernel key metrics and the top stall sites of the source page (run here, no GPU)."""
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    keys = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
            "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]
    idx = [(k, hdr.index(k)) for k in keys if k in hdr]
    for r in rows[2:]:
        print({k: r[i][:60] for k, i in idx})
    # all tensor-ish metrics
    for j, h in enumerate(hdr):
        if "tensor" in h or "tmem" in h.lower():
            print(h, [r[j] for r in rows[2:]])


def source(rep, which, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kern = []
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    k = kern[which]
    h = k["hdr"]
    si, so = h.index("# Samples"), h.index("Source")
    stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[si]) for r in k["rows"])
    print(k["name"][:50], "samples", tot)
    agg = {}
    for r in k["rows"]:
        for i in stalls:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    order = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][si]))[:top]
    for i in order:
        r = k["rows"][i]
        st = sorted([(int(r[j]), h[j]) for j in stalls], reverse=True)[:2]
        prev = k["rows"][i - 1][so].strip()[:50] if i else ""
        print(f"{r[si]:>6} #{i:<5} {r[so].strip()[:60]:60} {st}   <- {prev}")


if __name__ == "__main__":
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        source(rep, int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 30)
    else:
        raw(rep)
