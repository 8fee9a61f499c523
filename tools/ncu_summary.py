#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back into the text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv            # per-kernel share of the step
  python tools/ncu_summary.py kernel   gpurun_out/prof_score.ncu-rep      # key metrics of one --set full capture
  python tools/ncu_summary.py source   gpurun_out/prof_score.ncu-rep [N]  # top-N source lines by stall samples
  python tools/ncu_summary.py lanes    gpurun_out/prof_decoy.ncu-rep [N]  # top-N source lines by lost lane slots (divergence)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_shared_atom.sum",
    "sm__cycles_elapsed.max", "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def ncu_csv(rep, page):
    out = subprocess.check_output(["ncu", "-i", rep, "--page", page, "--csv"], text=True, stderr=subprocess.DEVNULL)
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = r[ki].replace("<unnamed>::", "")[:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# per-kernel device time over the whole command (ncu gpu__time_duration.sum, cold-cache, serialised: compare shares)")
    print("%-72s %6s %12s %7s" % ("kernel", "n", "ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %12.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("%-72s %6d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


def kernel(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("## %s" % name[:100])
        col = {h: i for i, h in enumerate(hdr)}
        for k in KEYS:
            if k in col:
                print("%-88s %-16s %s" % (k, units[col[k]], r[col[k]]))
        print("# warp stall reasons (warps per issue-active cycle)")
        st = [(float(r[i] or 0), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h, i in col.items()
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        for v, h in sorted(st, reverse=True)[:10]:
            print("  %-30s %.3f" % (h, v))


def source(rep, top=40):
    """CUDA source lines ranked by warp-stall samples (needs -lineinfo at compile time)."""
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= si or r[2] != "-":          # keep the per-source-line aggregate rows only
            continue
        try:
            v = float(r[si] or 0)
        except ValueError:
            continue
        st = sorted(((float(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        data.append((v, r[0], r[1], float(r[ii] or 0), st))
    tot = sum(d[0] for d in data) or 1
    print("# %s" % rows[1][1][:120])
    print("# top CUDA source lines by warp-stall samples; total samples %d" % tot)
    print("%7s %5s %12s  %-28s %s" % ("samples", "line", "warp-insts", "top stalls", "source"))
    for v, ln, text, ins, st in sorted(data, reverse=True)[:top]:
        sts = ",".join("%s:%d" % (h, x) for x, h in st if x > 0)
        print("%6.2f%% %5s %12d  %-28s %s" % (100 * v / tot, ln, ins, sts, text.strip()[:110]))


def lanes(rep, top=40):
    """CUDA source lines ranked by the lane slots they lose: warp instructions x (32 - active lanes)."""
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    ii, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= ti or r[2] != "-":
            continue
        try:
            ins, th = float(r[ii] or 0), float(r[ti] or 0)
        except ValueError:
            continue
        if ins > 0:
            data.append((ins, th, r[0], r[1]))
    tot, tt = sum(d[0] for d in data), sum(d[1] for d in data)
    lost = sum(d[0] * 32 - d[1] for d in data) or 1
    print("# %s" % rows[1][1][:120])
    print("# warp instructions %.4g, thread instructions %.4g: %.1f of 32 lanes active on average" % (tot, tt, tt / tot))
    print("%6s %7s %6s %7s  %s" % ("line", "%insts", "lanes", "%lost", "source"))
    for ins, th, ln, text in sorted(data, key=lambda d: -(d[0] * 32 - d[1]))[:top]:
        print("%6s %6.1f%% %6.1f %6.1f%%  %s" % (ln, 100 * ins / tot, th / ins, 100 * (ins * 32 - th) / lost, text.strip()[:110]))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "kernel":
        kernel(sys.argv[2])
    elif cmd == "lanes":
        lanes(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
    else:
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
