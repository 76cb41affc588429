"""Summarise an ncu report of conv_halo_kernel / conv_tc_kernel per warp role (here, no GPU needed):

    python tools/ncu_roles.py gpurun_out/prof.ncu-rep

Prints the headline metrics and, from the source page, sample counts / instruction counts / top stall
reasons per role (roles are located by their marker instructions in the SASS)."""
import csv
import io
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout


def main():
    raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr, val = raw[0], raw[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
    for h, u, v in zip(hdr, raw[1], val):
        if h in want:
            print(f"{h:<70s} {v} {u}")
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
    h = src[1]
    data = src[2:]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    sidx = {c: h.index(c) for c in stalls}
    # role boundaries: split at big jumps of the per-instruction execution count pattern is fragile; instead
    # let the user eyeball: print cumulative samples in windows of 100 instructions with the dominant opcode
    tot = sum(int(r[isamp]) for r in data)
    print(f"total samples {tot}, instructions {len(data)}")
    win = 60
    for lo in range(0, len(data), win):
        rows = data[lo:lo + win]
        s = sum(int(r[isamp]) for r in rows)
        if s < tot * 0.01:
            continue
        e = sum(int(r[iex]) for r in rows)
        st = {c: sum(int(r[sidx[c]]) for r in rows) for c in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        tags = set()
        for r in rows:
            t = r[isrc]
            for k in ("UTCHMMA", "LDGSTS", "LDTM", "UTMALDG", "STG", "UTCBAR", "TRYWAIT", "BAR.SYNC", "LDG"):
                if k in t:
                    tags.add(k)
        print(f"[{lo:5d},{lo + win:5d}) samples {s:6d} ({100.0 * s / tot:4.1f}%) inst {e:10d}  {top}  {sorted(tags)}")
    # hottest individual instructions
    hot = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:25]
    print("hottest instructions:")
    for i in sorted(hot):
        r = data[i]
        st = sorted(((c, int(r[sidx[c]])) for c in stalls), key=lambda kv: -kv[1])[:2]
        print(f"  {i:5d} {r[isrc].strip()[:80]:<80s} exec {r[iex]:>9s} samples {r[isamp]:>6s} {st}")


if __name__ == "__main__":
    main()
