"""Warp-stall breakdown of one kernel from an `ncu --set full --import-source on` report, the way
profiles/r01_summary.md quotes it: totals per stall reason, a windowed walk over the SASS (which
region of the kernel holds the samples, with the memory / tensor / shuffle opcodes found there as
landmarks) and the hottest instructions.

    python tools/ncu_stalls.py profiles/r01l_pointresnet_tc_pool.ncu-rep [--window 100] [--top 25]

Needs `ncu` on PATH (it only reads the report; no GPU)."""
import argparse
import csv
import io
import re
import subprocess
from collections import Counter

LANDMARKS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "SHFL", "MATCH", "VOTE",
             "REDG", "ATOMG", "ATOMS", "LDG", "STG", "LDS", "STS", "LDGSTS", "BAR", "EXIT")


def opcode(sass):
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", sass)
    return m.group(1) if m else "?"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--window", type=int, default=100, help="SASS instructions per window")
    ap.add_argument("--top", type=int, default=25, help="hottest instructions to list")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    recs = []
    for r in data:
        try:
            n = int(r[ix["# Samples"]])
        except (ValueError, IndexError):
            continue
        recs.append((r[ix["Source"]], n, int(r[ix["Instructions Executed"]] or 0),
                     {h: int(r[ix[h]] or 0) for h in stalls}))
    total = sum(r[1] for r in recs) or 1
    print("%d SASS instructions, %d samples" % (len(recs), total))
    tot = Counter()
    for r in recs:
        tot.update(r[3])
    for h, v in tot.most_common(10):
        print("  %-26s %8d  %5.1f %%" % (h, v, 100.0 * v / total))
    print("\nwindow  samples  share  warp-instr  top stalls / landmark opcodes")
    for lo in range(0, len(recs), a.window):
        seg = recs[lo:lo + a.window]
        s = sum(r[1] for r in seg)
        if s == 0:
            continue
        st = Counter()
        for r in seg:
            st.update(r[3])
        marks = Counter(op for op in (opcode(r[0]) for r in seg) if op in LANDMARKS)
        print("%6d %8d %5.1f%% %10d  %s  %s" % (lo, s, 100.0 * s / total, sum(r[2] for r in seg),
                                               [(h[6:], v) for h, v in st.most_common(3)], dict(marks)))
    print("\nhottest instructions")
    for i, r in sorted(enumerate(recs), key=lambda x: -x[1][1])[:a.top]:
        top = sorted(r[3].items(), key=lambda x: -x[1])[:2]
        print("%6d %6d  %-70s %s" % (i, r[1], r[0].strip()[:70], [(h[6:], v) for h, v in top]))


if __name__ == "__main__":
    main()
