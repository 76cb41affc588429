"""Per-launch table from an `ncu --page raw --csv` export of the conv launches of one eager forward.

    python tools/ncu_raw_table.py gpurun_out/r01_v5_convs_raw.csv [--names names.txt] [--json profiles/roofline_traffic.json]

Columns: duration, DRAM read / written, tensor pipe active (% of elapsed), L2 throughput %, L2->L1 bytes, grid.
--json writes the DRAM byte totals bench.py reports as roofline.traffic.
"""
import argparse
import csv
import json
import re


def col(hdr, name):
    for i, h in enumerate(hdr):
        if h == name:
            return i
    for i, h in enumerate(hdr):
        if h.endswith("." + name):
            return i
    raise KeyError(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--names", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [ln.strip() for ln in open(a.names)] if a.names else [""] * len(data)
    c_k = col(hdr, "Kernel Name")
    c_t = col(hdr, "gpu__time_duration.sum")
    c_r = col(hdr, "dram__bytes_read.sum")
    c_w = col(hdr, "dram__bytes_write.sum")
    c_tp = col(hdr, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    c_l2 = col(hdr, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    c_x = col(hdr, "l1tex__m_xbar2l1tex_read_bytes.sum")
    c_g = col(hdr, "Grid Size")

    def to_mb(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]

    def to_us(v, u):
        v = float(v.replace(",", ""))
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[u]

    print("| # | layer | kernel | us | DRAM rd MB | DRAM wr MB | tensor pipe % | L2 thr % | L2->L1 MB | grid |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    tot_t = tot_r = tot_w = 0.0
    for i, r in enumerate(data):
        k = re.sub(r"^.*conv_halo_kernel", "conv_halo_kernel", r[c_k])
        k = re.sub(r"\(.*$", "", k)
        t = to_us(r[c_t], units[c_t]); rd = to_mb(r[c_r], units[c_r]); wr = to_mb(r[c_w], units[c_w])
        tot_t += t; tot_r += rd; tot_w += wr
        grid = r[c_g].replace(",", "").strip("()").split()[0] if r[c_g] else ""
        print(f"| {i} | {names[i] if i < len(names) else ''} | {k} | {t:.1f} | {rd:.1f} | {wr:.1f} | {float(r[c_tp]):.1f} | "
              f"{float(r[c_l2]):.1f} | {to_mb(r[c_x], units[c_x]):.1f} | {grid} |")
    print(f"\nSum: {tot_t:.1f} us, DRAM read {tot_r:.1f} MB + write {tot_w:.1f} MB = {(tot_r + tot_w) / 1e3:.3f} GB")
    if a.json:
        json.dump({"conv_dram_bytes_per_step": (tot_r + tot_w) * 1e6, "dram_read_bytes": tot_r * 1e6,
                   "dram_write_bytes": tot_w * 1e6, "source": a.source}, open(a.json, "w"))


if __name__ == "__main__":
    main()
