"""Text summary of an `ncu --set full --import-source on` report: selected raw metrics per kernel + the SASS mix / stall
table of tools/ncu_sass_summary.py.  Run where ncu is installed (no GPU needed):

    python tools/ncu_summary.py gpurun_out/fin/prof_encode.ncu-rep "header line" > profiles/<name>_ncu_summary.txt
"""
import csv
import io
import os
import subprocess
import sys
import tempfile

METRICS = [
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sectors_srcunit_tex_op_read.sum",
    "sm__cycles_elapsed.avg", "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    print("# " + (sys.argv[2] if len(sys.argv) > 2 else rep))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    here = os.path.dirname(os.path.abspath(__file__))
    for r in body:
        print("## selected raw metrics (ncu -i ... --page raw --csv)")
        print("kernel:", r[col["Kernel Name"]])
        for m in METRICS:
            if m in col and r[col[m]] not in ("", "no data"):
                print("%-90s %s %s" % (m, r[col[m]], units[col[m]]))
        with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", r[col["ID"]], "--launch-count", "1"],
                                 stdout=subprocess.PIPE, text=True).stdout
            f.write(src)
        print("## SASS summary (tools/ncu_sass_summary.py on --page source --csv)")
        out = subprocess.run([sys.executable, os.path.join(here, "ncu_sass_summary.py"), f.name], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        print(out)
        os.unlink(f.name)


if __name__ == "__main__":
    main()
