"""From an `ncu --page source --csv --print-source sass` dump: per kernel, the stall-sample share and executed
instructions of consecutive SASS windows, split at barriers (BAR.SYNC) so that the windows line up with the
kernel's stages.  usage: python tools/ncu_hot_lines.py dump.csv"""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
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
        key = (k["name"], len(k["data"]))
        if key in seen:
            continue
        seen.add(key)
        ix = {h: i for i, h in enumerate(k["hdr"])}
        samp_col = [h for h in k["hdr"] if h.startswith("Warp Stall Sampling (All")] or [h for h in k["hdr"] if "Sampling" in h]
        sc = ix[samp_col[0]] if samp_col else None
        print("==", k["name"][:110])
        segs, cur_seg = [], {"n": 0, "exe": 0, "samp": 0, "lds": 0, "ldg": 0}
        for r in k["data"]:
            src = r[ix["Source"]]
            cur_seg["n"] += 1
            cur_seg["exe"] += int(r[ix["Instructions Executed"]])
            if sc is not None:
                try:
                    cur_seg["samp"] += int(r[sc])
                except ValueError:
                    pass
            if "LDS" in src: cur_seg["lds"] += int(r[ix["Instructions Executed"]])
            if "LDG" in src: cur_seg["ldg"] += int(r[ix["Instructions Executed"]])
            if "BAR.SYNC" in src or "EXIT" in src:
                segs.append(cur_seg)
                cur_seg = {"n": 0, "exe": 0, "samp": 0, "lds": 0, "ldg": 0}
        if cur_seg["n"]:
            segs.append(cur_seg)
        te = sum(s["exe"] for s in segs) or 1
        ts = sum(s["samp"] for s in segs) or 1
        for i, s in enumerate(segs):
            if s["exe"] == 0 and s["samp"] == 0:
                continue
            print("  stage %2d: %5d SASS  executed %5.1f%%  stall samples %5.1f%%  (LDS %d, LDG %d)"
                  % (i, s["n"], 100.0 * s["exe"] / te, 100.0 * s["samp"] / ts, s["lds"], s["ldg"]))


if __name__ == "__main__":
    main(sys.argv[1])
