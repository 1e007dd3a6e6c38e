#!/usr/bin/env python
"""Condense one-launch ``ncu --set full`` reports into the text summaries kept under
profiles/ (the .ncu-rep files themselves stay in the git-ignored gpurun_out/).

    python tools/ncu_summary.py gpurun_out/prof_pc_v14.ncu-rep "title" "command" > profiles/...txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
    "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = page(rep, "raw")
    hdr, units, val = rows[0], rows[1], rows[2]
    print(f"# ncu --set full --clock-control none, one launch: {title}")
    print(f"# command: {cmd} (B200, sm_100a)")
    print(f"# kernel: {val[hdr.index('Kernel Name')]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:<104s}{val[i]:>20s} {units[i]}")
    print("# warp stall reasons (average warps stalled per issue-active cycle)")
    st = [(float(val[i].replace(",", "") or 0), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and
          h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"{h:<104s}{v:>20.3f}")
    # hottest CUDA source lines by sampled stalls (needs -lineinfo + --import-source on)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                          "cuda,sass"], capture_output=True, text=True).stdout
    body, ci = [], None
    for r in csv.reader(io.StringIO(out)):
        if len(r) > 4 and r[0] == "Line No":
            ci = r.index("Warp Stall Sampling (All Samples)")
        elif ci is not None and len(r) > ci and r[0].isdigit() and r[ci].isdigit():
            body.append((int(r[ci]), int(r[0]), r[1].strip()))
    if body:
        tot = sum(b[0] for b in body)
        print(f"# hottest source lines by warp-stall samples (total {tot}; line numbers are "
              "per file, inlined helpers included)")
        for n, ln, txt in sorted(body, reverse=True)[:16]:
            print(f"{100.0 * n / tot:6.2f}%  L{ln:<5d} {txt[:100]}")


if __name__ == "__main__":
    main()
