"""Summarise an .ncu-rep (read with `ncu -i`) into a small markdown file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/fwd_r1.ncu-rep profiles/r1_fwd_v1.md "note" [kernel-name-regex]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


FILTER = []


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, *FILTER, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, *FILTER, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if r and r[0].startswith("0x") and len(r) > idx["# Samples"]]
    seen, first = set(), []
    for r in data:                      # keep the first captured launch only
        if r[0] in seen:
            break
        seen.add(r[0]); first.append(r)
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in first:
        for h in names:
            if idx[h] < len(r) and r[idx[h]]:
                tot[h] += int(r[idx[h]])
    top = sorted(first, key=lambda r: -int(r[idx["# Samples"]]))[:12]
    base = int(first[0][0], 16)
    return tot, [(hex(int(r[0], 16) - base), int(r[idx["# Samples"]]), r[1].strip()[:70]) for r in top], \
        sum(int(r[idx["# Samples"]]) for r in first)


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    if len(sys.argv) > 4:
        FILTER.extend(["-k", "regex:" + sys.argv[4]])
    hdr, units, rows = raw(rep)
    with open(dst, "w") as f:
        f.write(f"# ncu summary: {rep}\n\n{note}\n\n`ncu --set full --clock-control none --import-source on` "
                f"(per-launch numbers are cold-cache and serialised; a number printed under ncu is never a bench value)\n\n")
        for li, r in enumerate(rows[:2]):
            d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
            f.write(f"## launch {li}: {d.get('Kernel Name', ('', '?'))[1]}\n\n| metric | unit | value |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            f.write("\n")
        tot, top, n = stalls(rep)
        f.write(f"## warp-state samples, first launch (total {n})\n\n| stall reason | samples | share |\n|---|---|---|\n")
        for k, v in tot.most_common(10):
            f.write(f"| {k} | {v} | {100.0 * v / max(n, 1):.1f}% |\n")
        f.write("\n## hottest instructions\n\n| offset | samples | SASS |\n|---|---|---|\n")
        for a, s, t in top:
            f.write(f"| {a} | {s} | `{t}` |\n")


if __name__ == "__main__":
    main()
