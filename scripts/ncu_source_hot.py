#!/usr/bin/env python
"""Per-instruction view of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`:
executed-instruction and stall-sample totals by opcode class and the hottest instructions, so that a change to a kernel
can be aimed at the code that actually costs time.  Usage: python scripts/ncu_source_hot.py <source_sass.csv> [top_n]"""
import collections
import csv
import io
import re
import sys


def load(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")  # the first line names the kernel
    hdr = rows[h0]
    return hdr, [r for r in rows[h0 + 1:] if len(r) == len(hdr)]


def col(hdr, *names):
    for n in names:
        for i, h in enumerate(hdr):
            if h.strip().lower() == n.lower():
                return i
    for n in names:
        for i, h in enumerate(hdr):
            if n.lower() in h.strip().lower():
                return i
    return None


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return 0.0


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr, rows = load(path)
    i_src = col(hdr, "Source")
    i_smp = col(hdr, "# Samples", "Sampling Data (All)", "Samples")
    i_exe = col(hdr, "Instructions Executed", "# Instructions Executed")
    i_addr = col(hdr, "Address")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.lower().startswith("stall_") and "not issued" not in h.lower()]
    print("# columns:", " | ".join(hdr[:40]))
    tot_s = sum(num(r[i_smp]) for r in rows) or 1.0
    tot_e = sum(num(r[i_exe]) for r in rows) if i_exe is not None else 0.0
    print("# %d SASS instructions, %.0f samples, %.0f warp-instructions executed" % (len(rows), tot_s, tot_e))
    by_op = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for r in rows:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r[i_src])
        op = m.group(1) if m else "?"
        by_op[op][0] += num(r[i_smp])
        by_op[op][1] += num(r[i_exe]) if i_exe is not None else 0.0
        by_op[op][2] += 1
    print("\n# by opcode: samples%, executed%, static count")
    for op, (s, e, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:30]:
        print("%-12s %6.2f%% %6.2f%% %5d" % (op, 100 * s / tot_s, 100 * e / max(tot_e, 1), n))
    if stall_cols:
        print("\n# stall columns total (samples)")
        for i, h in sorted(stall_cols, key=lambda ih: -sum(num(r[ih[0]]) for r in rows))[:12]:
            print("%-28s %8.0f" % (h, sum(num(r[i]) for r in rows)))
    print("\n# hottest instructions (samples%, executed, address, SASS, top stall)")
    order = sorted(range(len(rows)), key=lambda k: -num(rows[k][i_smp]))[:top_n]
    for k in sorted(order):
        r = rows[k]
        st = ""
        if stall_cols:
            j, h = max(stall_cols, key=lambda ih: num(r[ih[0]]))
            st = "%s=%s" % (h, r[j])
        print("%6.2f%% %10.0f %s  %-70s %s" % (100 * num(r[i_smp]) / tot_s, num(r[i_exe]) if i_exe is not None else 0,
                                          r[i_addr] if i_addr is not None else "", r[i_src][:70], st))
    # cumulative profile along the kernel: 40 equal slices of the instruction stream
    print("\n# samples / executed along the instruction stream (25 slices)")
    n = len(rows)
    for b in range(25):
        lo, hi = b * n // 25, (b + 1) * n // 25
        s = sum(num(r[i_smp]) for r in rows[lo:hi])
        e = sum(num(r[i_exe]) for r in rows[lo:hi]) if i_exe is not None else 0
        print("[%5d,%5d) %6.2f%% samples %6.2f%% executed   first: %s" % (lo, hi, 100 * s / tot_s, 100 * e / max(tot_e, 1), rows[lo][i_src][:50]))


if __name__ == "__main__":
    main()
