"""Summarises an `ncu --page raw --csv` export: one row per launch with the metrics the roofline discussion uses."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=float("nan")):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return d
print("| # | kernel | grid | us | DRAM rd GB | DRAM wr GB | DRAM GB/s | L2->L1 GB | issue act % | warps act % | tensor % | regs | top stalls |")
print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
for n, r in enumerate(rows[2:]):
    name = r[ix["Kernel Name"]][:60]
    t = g(r, "gpu__time_duration.sum")
    unit = rows[1][ix["gpu__time_duration.sum"]]
    us = t / 1000.0 if unit in ("ns", "nsecond") else t
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    def tobytes(v, k):
        u = rows[1][ix[k]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    rd, wr = tobytes(rd, "dram__bytes_read.sum"), tobytes(wr, "dram__bytes_write.sum")
    l2k = "lts__t_sectors_srcunit_tex_op_read.sum"
    l2 = g(r, l2k) * 32 if l2k in ix else float("nan")
    st = sorted(((g(r, k, 0.0), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stall_keys), reverse=True)[:3]
    print("| %d | `%s` | %s | %.0f | %.3f | %.3f | %.0f | %.2f | %.0f | %.0f | %.1f | %.0f | %s |" % (
        n, name, r[ix["Grid Size"]] if "Grid Size" in ix else "", us, rd / 1e9, wr / 1e9, (rd + wr) / us / 1e3, l2 / 1e9,
        g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", g(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")),
        g(r, "launch__registers_per_thread"), ", ".join("%s %.1f" % (k, v) for v, k in st)))
