#!/usr/bin/env python
"""Warp-state samples of one kernel by code region (run here on an .ncu-rep captured with --set full --import-source on).
Regions are cut at tcgen05.ld / tcgen05.st (LDTM / STTM), mbarrier waits (SYNCS), named barriers (BAR) and UTCHMMA."""
import csv
import subprocess
import sys


def main(rep, out, min_share=0.4):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    st_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1

    def opname(r):
        op = r[ix["Source"]].strip().split()
        if not op:
            return ""
        name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        return name.split(".")[0]

    cuts = [0]
    for i, r in enumerate(data):
        if opname(r) in ("LDTM", "STTM", "SYNCS", "BAR", "UTCHMMA", "UTCBAR", "EXIT"):
            cuts.append(i)
            cuts.append(i + 1)
    cuts.append(len(data))
    cuts = sorted(set(cuts))
    # merge adjacent tiny regions into runs: a region = [a, b) between consecutive cuts
    lines = [f"# {rep}: warp-state samples by code region; {tot} samples, {tot_i / 1e9:.1f} G warp-instructions executed",
             "# region = SASS index range between tcgen05.ld/st, mbarrier, named-barrier and UTCHMMA instructions"]
    regs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b <= a:
            continue
        smp = sum(int(r[ix["# Samples"]]) for r in data[a:b])
        ins = sum(int(r[ix["Instructions Executed"]]) for r in data[a:b])
        regs.append([a, b, smp, ins])
    # coalesce neighbouring regions that are both small
    merged = []
    for r in regs:
        if merged and (r[2] < 0.002 * tot and merged[-1][2] < 0.002 * tot):
            merged[-1][1] = r[1]
            merged[-1][2] += r[2]
            merged[-1][3] += r[3]
        else:
            merged.append(list(r))
    for a, b, smp, ins in merged:
        if smp < min_share / 100 * tot and ins < min_share / 100 * tot_i:
            continue
        sl = data[a:b]
        stalls = sorted(((sum(int(r[ix[h]]) for r in sl), h[6:]) for h in st_cols), reverse=True)[:4]
        mix = {}
        for r in sl:
            mix[opname(r)] = mix.get(opname(r), 0) + int(r[ix["Instructions Executed"]])
        mx = sorted(mix.items(), key=lambda x: -x[1])[:6]
        first = data[a][ix["Source"]].strip()[:40]
        lines.append(f"SASS [{a:5d},{b:5d}) {100 * smp / tot:5.1f}% samples {100 * ins / tot_i:5.1f}% instr | "
                     + " ".join(f"{h}={100 * v / max(smp, 1):.0f}" for v, h in stalls) + " | "
                     + " ".join(f"{k}={100 * v / max(ins, 1):.0f}" for k, v in mx) + f" | {first}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.4)
