#!/usr/bin/env python
"""Summarise an ncu report for profiles/: key raw metrics per captured launch plus the SASS-level instruction and
stall picture.  usage: python tools/ncu_summary.py rep.ncu-rep > profiles/x.txt   (needs `ncu` on PATH, no GPU)"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_active.avg", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"launch id {r[h.index('ID')]}: {r[h.index('Kernel Name')]}")
        for k in KEYS:
            if k in h:
                print(f"  {k:68s} {r[h.index(k)]:>16s} {units[h.index(k)]}")
        rd, wr = float(r[h.index('dram__bytes_read.sum')]), float(r[h.index('dram__bytes_write.sum')])
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rdb = rd * scale.get(units[h.index('dram__bytes_read.sum')], 1.0)
        wrb = wr * scale.get(units[h.index('dram__bytes_write.sum')], 1.0)
        print(f"  {'traffic = dram read + write (bytes per launch)':68s} {rdb + wrb:16.0f}")


def sass(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_summary_src.csv"
    open(tmp, "w").write(out)
    sys.stdout.flush()
    subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_sass.py"), tmp, str(top)])


if __name__ == "__main__":
    rep = sys.argv[1]
    print(f"# {rep}\n# ncu --set full --clock-control none --import-source on (one launch; cold cache, serialised: compare shares)")
    raw(rep)
    print("\n# SASS level (executed warp instructions by opcode and by innermost CUDA line; stall samples)")
    sass(rep, int(sys.argv[2]) if len(sys.argv) > 2 else 24)
