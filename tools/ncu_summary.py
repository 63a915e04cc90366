"""Text summary of an .ncu-rep (ncu --set full): the metrics the roofline discussion uses + hottest source lines.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second"]


def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout


def main():
    rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== launch: {r[name_col][:100]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:100s} {r[i]:>16s} {units[i]}")
        stalls = sorted(((float(r[i] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", ""))
                         for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h),
                        reverse=True)
        print("   warp stall samples: " + ", ".join(f"{h}={int(a)}" for a, h in stalls[:6]))
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    h2, items = None, []
    for r in src:
        if "# Samples" in r:
            h2 = r
            continue
        if h2 and len(r) == len(h2):
            try:
                s = int(r[h2.index("# Samples")])
            except ValueError:
                continue
            if r[0] not in ("-", ""):
                items.append((s, r[0], r[1][:110]))
    seen, n = set(), 0
    print("== hottest source lines (samples, line, text) of the last profiled launch")
    for s, ln, txt in sorted(items, reverse=True):
        if (ln, txt) in seen:
            continue
        seen.add((ln, txt))
        print(f"   {s:7d}  {ln:>5s}  {txt}")
        n += 1
        if n >= 14:
            break


if __name__ == "__main__":
    main()
