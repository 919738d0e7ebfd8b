"""Static SASS instruction count per CUDA source line of one kernel, from `nvdisasm -g -c file.cubin` output
(compile with -lineinfo).  usage: python tools/sass_by_line.py all_sass.txt <mangled-name-substring> [N]"""
import re
import sys
from collections import Counter


def main(path, key, n=40):
    cnt, ops, cur, inside = Counter(), {}, None, False
    for line in open(path):
        if line.startswith("//--------------------- .text."):
            inside = key in line
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        s = line.strip()
        if s.startswith("/*") and ";" in s and cur:
            body = s.split("*/", 1)[1].strip()
            toks = body.split()
            op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
            cnt[cur] += 1
            ops.setdefault(cur, Counter())[op.split(".")[0]] += 1
    tot = sum(cnt.values())
    print("total SASS", tot)
    allops = Counter()
    for c in ops.values():
        allops.update(c)
    print("mix:", ", ".join("%s %d" % kv for kv in allops.most_common(24)))
    for (f, l), c in sorted(cnt.items(), key=lambda x: -x[1])[:n]:
        print("%-26s %4d  %4d  %s" % (f, l, c, dict(ops[(f, l)].most_common(5))))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
