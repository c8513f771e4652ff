"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep, read here without a GPU) into the few numbers the design
arguments use: duration, tensor-pipe / XU / DRAM utilisation, DRAM bytes, registers, and — from the source page — the
instructions that collected the most stall samples.

    python tools/ncu_summary.py gpurun_out/r2b_attn.ncu-rep [--top 12] [--kernel 0] > profiles/r2_attn_ncu_summary.txt
"""
import argparse
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("sm__cycles_elapsed.max.per_second", "SM clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_size", "cluster size"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
]


def ncu(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--kernel", type=int, default=-1, help="index of the launch to print the source page for (-1: none)")
    a = ap.parse_args()
    rows = ncu(a.report, "raw")
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {a.report}: {len(rows) - 2} launch(es); ncu --set full --clock-control none (cold caches, serialised)")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}: {r[ix['Kernel Name']][:110]}")
        for key, label in KEYS:
            if key in ix and r[ix[key]] not in ("", "n/a"):
                print(f"{label:28s} {r[ix[key]]:>16s} {units[ix[key]]}")
    if a.kernel >= 0:
        src = ncu(a.report, "source")
        blocks, cur = [], None
        for r in src:
            if r and r[0] == "Kernel Name":
                cur = []
                blocks.append(cur)
            elif cur is not None:
                cur.append(r)
        b = blocks[a.kernel]
        h, data = b[0], b[1:]
        jx = {c: i for i, c in enumerate(h)}
        stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[jx["# Samples"]]) for r in data)
        print(f"\n## source page of launch {a.kernel}: {len(data)} SASS instructions, {tot} stall samples; top {a.top}")
        for i in sorted(range(len(data)), key=lambda i: -int(data[i][jx["# Samples"]]))[:a.top]:
            r = data[i]
            n = int(r[jx["# Samples"]])
            st = sorted(((int(r[jx[s]]), s[6:]) for s in stalls), reverse=True)[:2]
            print(f"{100 * n / tot:5.1f}%  {r[jx['Source']].strip()[:64]:64s} exec={r[jx['Instructions Executed']]:>9s}  "
                  + " ".join(f"{s}={c}" for c, s in st if c))
    return 0


if __name__ == "__main__":
    sys.exit(main())
