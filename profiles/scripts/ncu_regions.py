#!/usr/bin/env python
"""Attribute the warp-stall samples of one kernel in an ncu report to regions of its SASS, delimited by marker opcodes.

    python profiles/scripts/ncu_regions.py REPORT.ncu-rep KERNEL_REGEX [OPCODE]    (default marker opcode: DMMA)

Prints, for every maximal stretch between "runs" of the marker opcode (a run = consecutive markers less than 40
instructions apart), the share of samples and the top stall reasons.  Works on the box-less container (reads the report)."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    mark = sys.argv[3] if len(sys.argv) > 3 else "DMMA"
    # KERNEL_REGEX may carry an invocation index ("k_lq:2" = second captured launch matching k_lq)
    kern, _, inv = kern.partition(":")
    sel = ["--kernel-id", "::regex:%s:%s" % (kern, inv)] if inv else ["--kernel-name", "regex:" + kern]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    data = []
    for r in rows[2:]:
        if len(r) > 5 and r[0].startswith("0x"):
            data.append(r)
        elif r and r[0] == "Address":
            break                                            # second (not-issued) table
    isrc, ismp = hdr.index("Source"), hdr.index("# Samples")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ismp] or 0) for r in data)
    marks = [k for k, r in enumerate(data) if mark in r[isrc]]
    runs = []
    for k in marks:
        if runs and k - runs[-1][1] < 40:
            runs[-1][1] = k
        else:
            runs.append([k, k])
    bounds = [0]
    for a, b in runs:
        bounds += [a, b + 1]
    bounds.append(len(data))
    print(rows[0][1][:90] if rows and len(rows[0]) > 1 else "")
    print("kernel %s: %d instructions, %d samples, %d %s in %d runs" % (kern, len(data), tot, len(marks), mark, len(runs)))
    for n, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
        if hi <= lo:
            continue
        smp = sum(int(r[ismp] or 0) for r in data[lo:hi])
        st = {}
        for i, h in stall:
            v = sum(int(r[i] or 0) for r in data[lo:hi])
            if v:
                st[h[6:]] = v
        top = sorted(st.items(), key=lambda t: -t[1])[:4]
        nm = sum(1 for r in data[lo:hi] if mark in r[isrc])
        kind = "%s run (%d)" % (mark, nm) if n % 2 == 1 else "between"
        print("  [%5d,%5d) %-16s %6.2f%%  %s" % (lo, hi, kind, 100.0 * smp / max(tot, 1),
                                                 ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in top)))


if __name__ == "__main__":
    main()
