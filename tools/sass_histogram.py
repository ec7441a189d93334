"""SASS opcode histogram of the tensor-core / TMA / TMEM instructions per kernel of libclimsr_b200.so (cuobjdump -sass).
Evidence that the hot kernels are tcgen05 / TMEM / TMA code:  python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "climate-super-resolution_b200", "climsr_b200", "libclimsr_b200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "UTCATOMSWS", "REDG", "RED.", "ATOMG", "SHFL",
        "FADD2", "FFMA2", "FMUL2", "HMMA", "LDGSTS", "BAR.SYNC", "ACQBULK", "ELECT", "UCGABAR", "MEMBAR", "FENCE")


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    sizes = {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        sizes[cur] = int(m.group(1), 16) + 16
        ins = m.group(2)
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
        for k in KEYS:
            if ins.startswith(k):
                per[cur][k.rstrip(".")] += 1
        per[cur]["_total"] += 1
    print(f"# {os.path.basename(SO)}: instruction counts of the Blackwell-specific opcodes per kernel (cuobjdump -sass, sm_100a)")
    print("# UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load), UTMAPF = TMA prefetch,")
    print("# UBLKCP = cp.async.bulk, SYNCS = mbarrier, UTCBAR = tcgen05.commit, REDG = red.global, ELECT = elect.sync")
    for name, c in per.items():
        if not any(k in c for k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP")) and c["_total"] < 400:
            continue
        items = " ".join(f"{k}={v}" for k, v in sorted(c.items()) if k != "_total")
        print(f"{name[:110]:110s} {sizes.get(name, 0) / 1024:6.1f} KB  {items}")


if __name__ == "__main__":
    main()
