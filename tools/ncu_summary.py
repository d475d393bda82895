#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): headline metrics + stall samples per barrier-delimited code region.
Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--regions]"""
import csv, subprocess, sys, io
from collections import Counter

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__cycles_elapsed.avg', 'launch__shared_mem_per_block_dynamic',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'sm__cycles_active.avg',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for val in rows[2:]:
        print("==", val[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:85s} {val[i]:>16s} {units[i]}")
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    hdr, data = rows[1], rows[2:]
    iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[iN]) for r in data) or 1
    totE = sum(int(r[iE]) for r in data) or 1
    c = Counter()
    ops = Counter()
    for r in data:
        for i in stall:
            c[hdr[i][6:]] += int(r[i] or 0)
        t = r[iS].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] += int(r[iE])
    print(f"  samples {tot}  warp-instructions {totE}  sass lines {len(data)}")
    print("  stalls:", ", ".join(f"{k}={100 * v / tot:.1f}%" for k, v in c.most_common(9)))
    print("  opcodes:", ", ".join(f"{k}={100 * v / totE:.1f}%" for k, v in ops.most_common(12)))
    bars = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[iS]]
    prev = 0
    for b in bars + [len(data) - 1]:
        blk = data[prev:b + 1]
        sm = sum(int(r[iN]) for r in blk)
        e = sum(int(r[iE]) for r in blk)
        ff = sum(int(r[iE]) for r in blk if 'FFMA' in r[iS])
        cc = Counter()
        for r in blk:
            for i in stall:
                cc[hdr[i][6:]] += int(r[i] or 0)
        print(f"  region [{prev:5d},{b + 1:5d}) samples {100 * sm / tot:5.1f}%  instr {100 * e / totE:5.1f}% (ffma {100 * ff / max(e, 1):4.1f}%)  "
              + ", ".join(f"{k}={v}" for k, v in cc.most_common(5)))
        prev = b + 1


if __name__ == "__main__":
    main()
