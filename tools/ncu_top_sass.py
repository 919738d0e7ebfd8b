"""From an `ncu --page source --csv --print-source sass` dump: per kernel, the SASS instructions with the most
executed warp instructions and the most stall samples (with the dominant stall reasons), plus totals per opcode class.
usage: python tools/ncu_top_sass.py dump.csv [N]"""
import csv
import sys
from collections import Counter


def main(path, n=25):
    rows = list(csv.reader(open(path)))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kern.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
            cur["data"].append(r)
    for k in kern:
        ix = {h: i for i, h in enumerate(k["hdr"])}
        stalls = [h for h in k["hdr"] if h.startswith("stall_")]
        d = k["data"]
        exe = [int(r[ix["Instructions Executed"]] or 0) for r in d]
        smp = [int(r[ix["Warp Stall Sampling (All Samples)"]] or 0) for r in d]
        te, ts = sum(exe) or 1, sum(smp) or 1
        print("==", k["name"][:100], "| SASS lines", len(d), "| warp instr", te, "| samples", ts)
        ops = Counter()
        for r, e in zip(d, exe):
            op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
            if op.startswith("@"):
                op = r[ix["Source"]].split()[1]
            ops[op.split(".")[0]] += e
        print("   opcode mix:", ", ".join("%s %.1f%%" % (o, 100.0 * c / te) for o, c in ops.most_common(18)))
        tot_reason = Counter()
        for r in d:
            for s in stalls:
                try:
                    tot_reason[s] += int(r[ix[s]] or 0)
                except ValueError:
                    pass
        tr = sum(tot_reason.values()) or 1
        print("   stall reasons:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * c / tr) for s, c in tot_reason.most_common(8)))
        print("   -- top by stall samples")
        for i in sorted(range(len(d)), key=lambda i: -smp[i])[:n]:
            r = d[i]
            rs = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
            print("   %5d %5.1f%% exe %9d  %-70s %s" % (i, 100.0 * smp[i] / ts, exe[i], r[ix["Source"]][:70], " ".join("%s:%d" % (b, a) for a, b in rs if a)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
