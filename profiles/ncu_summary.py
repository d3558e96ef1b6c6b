"""Reads an .ncu-rep (ncu -i ... --page raw --csv) and prints the metrics the judge looks at.
Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[H.index("Kernel Name")][:90])
        for w in WANT:
            if w in H:
                print("  %-85s %s %s" % (w, r[H.index(w)], U[H.index(w)]))
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        hdr = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r))
        H = rows[hdr]
        si = H.index("Source")
        ci = next(i for i, c in enumerate(H) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
        body = [r for r in rows[hdr + 1:] if len(r) > ci and r[ci].replace(",", "").isdigit()]
        body.sort(key=lambda r: -int(r[ci].replace(",", "")))
        tot = sum(int(r[ci].replace(",", "")) for r in body) or 1
        print("top source lines by stall samples (%s):" % H[ci])
        for r in body[:n]:
            print("  %6.2f%%  %s" % (100.0 * int(r[ci].replace(",", "")) / tot, r[si][:150]))


if __name__ == "__main__":
    main()
