"""Summarise .ncu-rep files (read here, without a GPU) into markdown for profiles/:
    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/r01_x.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor pipe instructions"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks/SM"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard (cycles/issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex_throttle"),
    ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "selected"),
]


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"## {path}\n"]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        out.append(f"### `{name}`  (launch id {r[idx['ID']]})\n")
        out.append("| metric | value |\n|---|---|")
        for key, label in METRICS:
            if key in idx and r[idx[key]] not in ("", "n/a"):
                out.append(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    print("# ncu summary (captured with --set full --clock-control none under gpurun; read offline)\n")
    for p in sys.argv[1:]:
        print(summarise(p))
