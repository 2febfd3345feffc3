#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the metric,unit,value table kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_ncu_<kernel>_summary.csv"""
import csv
import io
import subprocess
import sys

KEEP = ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_requests_srcunit_tex.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w") as f:
        f.write("metric,unit,value\n")
        for n, u, v in sorted(zip(names, units, vals)):
            if n in KEEP or n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio"):
                f.write('%s,%s,"%s"\n' % (n, u, v) if "," in v else "%s,%s,%s\n" % (n, u, v))
    print("wrote", out)


if __name__ == "__main__":
    main()
