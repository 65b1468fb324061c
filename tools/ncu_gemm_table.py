"""Summarise an `ncu --set full` report of the tensor-core GEMM kernels as a markdown table (read on the CPU box):
python tools/ncu_gemm_table.py gpurun_out/<name>.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("kernel", "Kernel Name"), ("ms", "gpu__time_duration.sum"), ("SM MHz", "sm__cycles_elapsed.max.per_second"),
        ("DRAM read GB", "dram__bytes_read.sum"), ("DRAM write GB", "dram__bytes_write.sum"),
        ("tensor pipe active %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("shared-memory wavefronts % of peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warp instructions (M)", "smsp__inst_executed.sum"), ("registers", "launch__registers_per_thread"),
        ("smem KB / CTA", "launch__shared_mem_per_block_dynamic"), ("cluster", "launch__cluster_size")]
print("| " + " | ".join(n for n, _ in want) + " |")
print("|" + "---|" * len(want))
for r in data:
    out = []
    for n, k in want:
        v = r[col[k]] if k in col else "n/a"
        if n == "kernel":
            v = "`" + v.split("(")[0].replace("void ", "") + "`"
        elif n == "warp instructions (M)":
            v = "%.0f" % (float(v) / 1e6)
        elif n == "SM MHz":
            v = "%.0f" % (float(v) * 1000)
        else:
            try:
                v = "%.3f" % float(v) if float(v) < 100 else "%.0f" % float(v)
            except ValueError:
                pass
        out.append(v)
    print("| " + " | ".join(out) + " |")
