"""List the SASS instructions of a kernel that collect the most warp-stall samples, in address order, with the
dominant stall reasons per instruction -- enough to attribute stalls to the warp roles of a specialised kernel.
usage: python tools/ncu_hot_sass.py prof.ncu-rep <kernel regex> [min_pct]"""
import csv
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, launches, body = None, 0, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            launches += 1
            continue
        if r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            names = r
            continue
        if hdr is None or launches > 1 or len(r) < len(hdr):
            continue
        body.append(r)
    stall_cols = [i for i, h in enumerate(names) if h.startswith("stall_") or h.startswith("Stall")]
    if not stall_cols:  # ncu names the per-reason columns after the generic ones
        stall_cols = list(range(hdr["Warp Stall Sampling (Not-issued Samples)"] + 1, len(names)))
        stall_cols = [i for i in stall_cols if names[i].lower().startswith(("stall", "warp stall"))] or stall_cols
    tot = sum(int(r[hdr["Warp Stall Sampling (All Samples)"]] or 0) for r in body)
    print(f"{kre}: {len(body)} SASS instructions, {tot} stall samples; columns: {[names[i] for i in stall_cols][:40]}")
    for k, r in enumerate(body):
        s = int(r[hdr["Warp Stall Sampling (All Samples)"]] or 0)
        src = r[hdr["Source"]].strip()
        key = any(t in src for t in ("UTCHMMA", "UTMALDG", "LDTM", "SYNCS", "UTCBAR", "BAR.", "USETMAXREG", "FENCE", "UTMA"))
        if 100.0 * s / max(tot, 1) >= min_pct or (key and s > 0):
            reasons = []
            for i in stall_cols:
                try:
                    v = int(r[i] or 0)
                except ValueError:
                    continue
                if v > 0.25 * s and v > 0:
                    reasons.append(f"{names[i]}={v}")
            print(f"{k:5d} {100.0 * s / max(tot, 1):5.1f}% exec={r[hdr['Instructions Executed']]:>9s} {src[:80]:80s} {' '.join(reasons)[:120]}")


if __name__ == "__main__":
    main()
