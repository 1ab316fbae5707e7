"""Turn an .ncu-rep (ncu --set full) into the compact per-kernel CSVs kept under profiles/:
one file per kernel, rows `metric,unit,value` for the metrics the roofline / DESIGN.md quote.
Usage: python profiles/ncu_extract.py gpurun_out/prof.ncu-rep r02   (runs `ncu -i` locally)."""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum|"
    r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__m_xbar2l1tex_read_bytes\.sum|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|l1tex__data_pipe_tc_wavefronts_mem_shared\.sum|"
    r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed\.sum|smsp__inst_executed\.sum|sm__inst_issued\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__cycles_elapsed\.max|smsp__cycles_active\.avg|launch__registers_per_thread|launch__grid_size|"
    r"launch__block_size|launch__shared_mem_per_block_dynamic|launch__occupancy_limit_.*|"
    r"smsp__inst_executed_op_tma_ld\.sum|sm__sass_inst_executed_op_shared_(ld|st)\.sum)$")


def main(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ki]).split("::")[-1].strip()
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:
            continue  # first captured launch of every kernel
        out = f"profiles/{tag}_ncu_{name.replace('_kernel', '')}.csv"
        with open(out, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            w.writerow(["kernel", "", r[ki]])
            for h, u, v in zip(hdr, units, r):
                if KEEP.match(h):
                    w.writerow([h, u, v])
        print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r02")
