"""Summarise an .ncu-rep (read here, no GPU): per-launch key metrics -> markdown. usage: ncu_summary.py rep [out.md]"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_%"),
        ("smsp__inst_executed.sum", "warp_insts"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch")]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = ["# ncu summary of `%s`" % rep, "", "(`ncu --set full --clock-control none`; per-launch values; times are cold-cache and serialised)", ""]
    for r in rows[2:]:
        out.append("## " + r[ix["Kernel Name"]])
        out.append("")
        out.append("| metric | value |")
        out.append("|---|---|")
        for key, name in WANT:
            if key in ix and r[ix[key]] not in ("", "-nan", "nan"):
                out.append("| %s | %s %s |" % (name, r[ix[key]], units[ix[key]]))
        out.append("")
    txt = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    else:
        print(txt)


if __name__ == "__main__":
    main()
