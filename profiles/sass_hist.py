"""Opcode histogram (executed warp instructions) of one kernel from an ncu report.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv; python sass_hist.py sass.csv <units>
`units` = number of (warp, work item) pairs to normalise by (optional)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
secs = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        secs.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"]:
        cur["rows"].append(r)
for s in secs:
    h = s["hdr"]
    iS, iI, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    ops = collections.Counter()
    stall = collections.Counter()
    tot = 0
    for r in s["rows"]:
        try:
            n = int(r[iI])
        except ValueError:
            continue
        t = r[iS].split()
        op = t[1] if t[0].startswith("@") else t[0]
        if not op.startswith(("LDS", "ATOMS", "SHFL", "MUFU", "STG", "LDG", "RED", "ATOMG", "BAR")):
            op = op.split(".")[0]
        ops[op] += n
        stall[op] += int(r[iW] or 0)
        tot += n
    print("==", s["name"][:70], "total warp instr", tot, "" if not units else "per unit %.1f" % (tot / units))
    ts = sum(stall.values()) or 1
    for op, n in ops.most_common(36):
        print("  %-22s %6.2f%% %s  stall-samples %5.1f%%" % (op, 100 * n / tot, "" if not units else "%7.1f" % (n / units), 100 * stall[op] / ts))
