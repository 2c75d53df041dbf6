"""Summarise an .ncu-rep (ncu --set full) into profiles/<name>.md + .csv: the counters the design
notes cite (duration, registers, occupancy, DRAM bytes, pipe utilisation, issue rate, stall mix).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tcgen05 (UTCMMA) pipe active %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tensor-core smem operand wavefronts % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA load bytes (L2 -> smem)"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second", "TMA load rate"),
    ("smsp__cycles_elapsed.avg.per_second", "SM clock during the capture"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [k for k, _ in KEYS if k in idx])
        w.writerow(["unit"] + [units[idx[k]] for k, _ in KEYS if k in idx])
        for r in data:
            w.writerow([r[idx["Kernel Name"]]] + [r[idx[k]] for k, _ in KEYS if k in idx])
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep.split('/')[-1]}` (ncu --set full --clock-control none)\n\n")
        for r in data:
            f.write(f"## {r[idx['Kernel Name']]}\n\n| counter | value |\n|---|---|\n")
            for k, nm in KEYS:
                if k in idx:
                    f.write(f"| {nm} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |\n")
            f.write("\n")
    print("wrote", out + ".md", out + ".csv")


if __name__ == "__main__":
    main()
