"""Key rows of an `ncu --set full` report for profiles/: duration, tensor-pipe / issue / memory throughput, DRAM bytes, top stall reasons.
usage: python tools/ncu_keyrows.py report.ncu-rep [kernel-substring] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ("gpu__time_duration.sum", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_uniform", "sm__throughput.avg.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed_op_shared", "sm__sass_inst_executed_op_shared",
        "tmem", "utc", "sm__ops_path_tensor")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


for vals in rows[2:]:
    if len(vals) != len(hdr):
        continue
    d = dict(zip(hdr, vals))
    if flt and flt not in d.get("Kernel Name", ""):
        continue
    print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:90], d.get("Grid Size"), d.get("Block Size")))
    for h, u in zip(hdr, units):
        if "sm__ops_path_tensor" in h:
            if h.endswith("utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"):
                print("  %-95s %-14s %s   <- tensor-pipe utilisation (bf16 MMA ops / peak)" % (h, u, d[h]))
            continue
        if any(k in h for k in KEYS) and "stalled" not in h and not any(x in h for x in (".max", ".min", ".sum.pct", "launch__cluster_", "imma", "dmma", "sm__mio", "utccp", "utcshift", "stsm", "_sp_sf", "shared_atom", "syslts")):
            print("  %-95s %-14s %s" % (h, u, d[h]))
    st = [(h, num(d[h])) for h in hdr if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    print("  -- top stall reasons (warps stalled per issue-active cycle)")
    for h, v in sorted(st, key=lambda t: -t[1])[:8]:
        print("  %-95s %s" % (h, v))
