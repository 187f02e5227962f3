"""Summarises .ncu-rep captures (ncu --set full) into the few numbers the roofline uses.
Usage: python profiles/ncu_summary.py gpurun_out/prof_r01_spmv.ncu-rep [...] > profiles/r01_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for rep in sys.argv[1:]:
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("==", rep)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        print("  kernel:", name)
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = (r[i], units[i])
                print("    %-78s %s %s" % (w, r[i], units[i]))

        def num(k):
            v, u = vals[k]
            v = float(v.replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
        try:
            tr = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            print("    traffic (dram read + write) per launch: %.4f GB; duration under ncu %.3f ms -> %.0f GB/s" %
                  (tr / 1e9, num("gpu__time_duration.sum") * 1e3, tr / num("gpu__time_duration.sum") / 1e9))
        except Exception as e:
            print("    (traffic n/a: %s)" % e)
