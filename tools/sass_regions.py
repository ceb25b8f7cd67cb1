"""Summarise an `ncu --page source --csv --print-source sass` export into straight-line regions:
share of executed warp instructions, execution count, average active threads, stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ithr, isamp = (hdr.index(c) for c in ("Address", "Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
ins = [(r[isrc].strip(), int(r[iex]), float(r[ithr]), int(r[isamp])) for r in rows[2:] if len(r) > isamp]
total = sum(i[1] for i in ins)
tsamp = sum(i[3] for i in ins)
print("total", total, "samples", tsamp)
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
start = 0
for i in range(1, len(ins) + 1):
    if i == len(ins) or ins[i][1] != ins[start][1]:
        seg = ins[start:i]
        share = 100.0 * sum(s[1] for s in seg) / total
        if share >= thresh:
            ops = " ".join(s[0].split()[0] if not s[0].startswith("@") else s[0].split()[1] for s in seg[:14])
            print(f"[{start:4d}-{i - 1:4d}] len={len(seg):3d} share={share:5.2f}% cnt={seg[0][1]:>11d} thr={seg[0][2]:4.1f} "
                  f"samp={100.0 * sum(s[3] for s in seg) / tsamp:5.2f}% {ops}")
        start = i
