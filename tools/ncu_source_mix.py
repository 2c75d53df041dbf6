"""Aggregate an ncu source page (--page source --csv) by SASS opcode: executed share and stall-sample share.
usage: python tools/ncu_source_mix.py gpurun_out/prof.ncu-rep <kernel regex>"""
import collections
import csv
import subprocess
import sys


def analyze(rep, kre):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = None
    ex, st, conf = collections.Counter(), collections.Counter(), collections.Counter()
    tot = sttot = 0
    launches = 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            launches += 1
            continue
        if r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr) or launches > 1:
            continue
        src = r[hdr["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        n = int(r[hdr["Instructions Executed"]] or 0)
        s = int(r[hdr["Warp Stall Sampling (All Samples)"]] or 0)
        ex[op] += n
        st[op] += s
        tot += n
        sttot += s
        c = r[hdr["L1 Wavefronts Shared Excessive"]]
        if c and int(c) > 0:
            conf[src[:70]] += int(c)
    print(f"{kre}: warp instructions {tot}, stall samples {sttot}")
    for op, n in ex.most_common(16):
        print(f"  {op:10s} exec {n:12d} {100 * n / tot:5.1f}%   stall-samples {st[op]:7d} {100 * st[op] / max(sttot, 1):5.1f}%")
    print("  top excessive shared wavefronts:")
    for k, v in conf.most_common(6):
        print("    ", v, k)


if __name__ == "__main__":
    analyze(sys.argv[1], sys.argv[2])
