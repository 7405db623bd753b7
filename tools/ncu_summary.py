"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per captured launch."""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] + (" [%s]" % units[i] if units[i] else "") for i in idx] + ["dram GB/s (read+write)/time"])
        for r in rows[2:]:
            try:
                t = float(r[hdr.index("gpu__time_duration.sum")])
                tu = units[hdr.index("gpu__time_duration.sum")]
                secs = t * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9}.get(tu, 1e-6)
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
                wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
                bw = "%.1f" % ((rd + wr) / secs / 1e9)
            except Exception:
                bw = ""
            w.writerow([r[i] for i in idx] + [bw])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
