#!/usr/bin/env python
"""Condense an .ncu-rep into the handful of numbers DESIGN.md / bench.py cite (run here, no GPU needed)."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "launch__shared_mem_per_block_dynamic",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# {rep}"]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"kernel: {d.get('Kernel Name', '?')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:85s} {d[k]:>16s} {units[hdr.index(k)]}")
        stalls = sorted(((float(d[h]), h) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")
                         and d[h] not in ("", "n/a")), reverse=True)[:8]
        lines.append("  warp stalls per issue-active cycle: " + ", ".join(
            f"{h.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in stalls))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 3:
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) == len(hdr)]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        lines.append(f"  sampled warp states: {tot}")
        st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = sorted(((sum(int(r[ix[h]]) for r in data), h) for h in st), reverse=True)[:8]
        lines.append("  " + ", ".join(f"{h}={100 * v / tot:.1f}%" for v, h in agg))
        lines.append("  hottest SASS instructions (share of samples):")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:10]:
            lines.append(f"    {100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:90]}")
        mn = {}
        for r in data:
            op = r[ix["Source"]].strip().split()
            if not op:
                continue
            name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            name = name.split(".")[0]
            mn[name] = mn.get(name, 0) + int(r[ix["Instructions Executed"]])
        tot_i = sum(mn.values()) or 1
        lines.append("  executed warp-instruction mix: " + ", ".join(
            f"{k}={100 * v / tot_i:.1f}%" for k, v in sorted(mn.items(), key=lambda x: -x[1])[:14]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
