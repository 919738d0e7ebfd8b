"""Join an `ncu --page source --csv --print-source sass` dump with `nvdisasm -g -c` output of the same cubin (both list a
kernel's SASS in address order) and print executed warp instructions + stall samples per CUDA source line.
usage: python tools/ncu_exec_by_line.py dump.csv all_sass.txt <ncu-kernel-substring> <mangled-substring> [N]"""
import csv
import re
import sys
from collections import Counter


def main(dump, sass, nkey, mkey, n=40):
    rows = list(csv.reader(open(dump)))
    cur, data, hdr = None, None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = r[1]
            if nkey in cur and data is None:
                data = []
                take = True
            else:
                take = False
        elif r and r[0] == "Address":
            hdr = r
        elif data is not None and take and hdr and len(r) >= len(hdr) - 2:
            data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    lines, inside, curline = [], False, None
    for line in open(sass):
        if line.startswith("//--------------------- .text."):
            inside = mkey in line
            curline = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            curline = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        s = line.strip()
        if s.startswith("/*") and ";" in s:
            lines.append(curline)
    print("ncu SASS rows", len(data), "nvdisasm instructions", len(lines))
    exe, smp = Counter(), Counter()
    for r, l in zip(data, lines):
        exe[l] += int(r[ix["Instructions Executed"]] or 0)
        smp[l] += int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    te, ts = sum(exe.values()) or 1, sum(smp.values()) or 1
    for l, c in sorted(exe.items(), key=lambda x: -x[1])[:n]:
        print("%-26s %5s  exe %5.1f%%  stall %5.1f%%" % (l[0] if l else "?", l[1] if l else "", 100.0 * c / te, 100.0 * smp[l] / ts))


if __name__ == "__main__":
    a = sys.argv
    main(a[1], a[2], a[3], a[4], int(a[5]) if len(a) > 5 else 40)
