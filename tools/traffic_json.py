"""Build profiles/traffic.json (bench.py's roofline.traffic) from the per-workload ncu CSVs written by
tools/profile_round.sh: DRAM bytes read + written by ONE launch of the dominant kernel at full size.
usage: python tools/traffic_json.py gpurun_out/<tag>_traffic_ profiles/traffic.json"""
import csv
import json
import sys


def main():
    prefix, out = sys.argv[1], sys.argv[2]
    res = {}
    for w in ("fir", "decim", "interp", "iir_batch", "iir_scan", "autocorr"):
        rows = [r for r in csv.reader(open(f"{prefix}{w}.csv")) if len(r) > 10]
        hdr = {h: i for i, h in enumerate(rows[0])}
        per = {}
        for r in rows[1:]:
            per.setdefault(r[hdr["ID"]], {"kernel": r[hdr["Kernel Name"]]})[r[hdr["Metric Name"]]] = float(r[hdr["Metric Value"]])
        last = per[sorted(per, key=int)[-1]]  # the last captured launch (warm)
        res[w] = {"kernel": last["kernel"][:100], "dram_bytes_read": last["dram__bytes_read.sum"],
                  "dram_bytes_write": last["dram__bytes_write.sum"],
                  "dram_bytes": last["dram__bytes_read.sum"] + last["dram__bytes_write.sum"],
                  "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one launch at the bench's full size"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v["dram_bytes"] for k, v in res.items()}))


if __name__ == "__main__":
    main()
