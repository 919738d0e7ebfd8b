"""Summarise an `ncu --page source --csv --print-source sass` dump: executed-instruction mix,
stall samples and the hottest SASS ranges.  usage: python tools/ncu_sass_summary.py dump.csv"""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    kernels = []
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
            cur["data"].append(r)
    seen = set()
    for k in kernels:
        hdr, data = k["hdr"], k["data"]
        key = (k["name"], len(data))
        if key in seen:
            continue          # ncu repeats the table once per profiled launch / section
        seen.add(key)
        ix = {h: i for i, h in enumerate(hdr)}
        tot = 0
        byop = collections.Counter()
        stall = collections.Counter()
        for r in data:
            n = int(r[ix["Instructions Executed"]])
            tot += n
            toks = r[ix["Source"]].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            byop[op.split(".")[0]] += n
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    stall[h] += int(r[ix[h]])
        print("==", k["name"])
        print("static SASS instructions:", len(data), " executed warp instructions:", tot)
        for op, n in byop.most_common(top):
            print("  %-10s %12d %5.1f%%" % (op, n, 100.0 * n / max(tot, 1)))
        s = sum(stall.values())
        print("stall samples:", ", ".join("%s %.1f%%" % (a[6:], 100.0 * b / max(s, 1)) for a, b in stall.most_common(8)))
        step = max(len(data) // 20, 1)
        print("executed share by SASS range:")
        for i in range(0, len(data), step):
            e = sum(int(r[ix["Instructions Executed"]]) for r in data[i:i + step])
            print("  [%5d,%5d) %5.1f%%" % (i, i + step, 100.0 * e / max(tot, 1)))


if __name__ == "__main__":
    main(sys.argv[1])
