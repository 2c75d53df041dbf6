"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name:
launches, total and mean duration, share of the summed GPU time.
usage: python tools/launch_summary.py profiles/r1f_launches_bench_fir.csv [out.md]"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = {h: i for i, h in enumerate(rows[0])}
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows[1:]:
        if r[hdr["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[hdr["Kernel Name"]])
        name = re.sub(r"^void ", "", name)[:90]
        ns = float(r[hdr["Metric Value"]])
        if r[hdr["Metric Unit"]] in ("us", "usecond"):
            ns *= 1e3
        elif r[hdr["Metric Unit"]] in ("ms", "msecond"):
            ns *= 1e6
        tot[name] += ns
        cnt[name] += 1
    allns = sum(tot.values())
    lines = ["| kernel | launches | total ms | mean ms | share of GPU time |", "|---|---|---|---|---|"]
    for name, ns in tot.most_common():
        lines.append(f"| `{name}` | {cnt[name]} | {ns / 1e6:.3f} | {ns / 1e6 / cnt[name]:.4f} | {100 * ns / allns:.1f} % |")
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(f"# Launch list summary of `{sys.argv[1]}`\n\n"
                    "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` on "
                    "`python bench.py --steps 2 --warmup 3 --no-cpu` (per-launch times under ncu are serialised and "
                    "cold-cache: compare shares, not absolutes).\n\n" + text + "\n")


if __name__ == "__main__":
    main()
