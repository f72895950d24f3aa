"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python profiles/summarize_ncu.py launches <launches.csv>            -> per-kernel share of the step (gpu__time_duration)
  python profiles/summarize_ncu.py report <file.ncu-rep> [max_kernels] -> key metrics of every captured launch + stall mix"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__cluster_dim_x", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = r[ki].split("(zigp::")[0].split("(const")[0][:90]
        agg[n][0] += 1; agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-92s %6s %12s %6s" % ("kernel", "count", "total us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-92s %6d %12.1f %5.1f%%" % (n, c, t / 1e3, 100 * t / tot))


def report(path, maxk=8):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, units = r[0], r[1]
    for row in r[2:2 + maxk]:
        d = dict(zip(h, row))
        print("==", d.get("Kernel Name", "?")[:110], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print("   %-85s %16s %s" % (k, d[k], units[h.index(k)]))
        st = sorted(((k.replace("smsp__pcsamp_warps_issue_stalled_", ""), int(float(v))) for k, v in d.items()
                     if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v), key=lambda x: -x[1])
        tot = sum(v for _, v in st) or 1
        print("   stall samples:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in st[:8]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 8)
