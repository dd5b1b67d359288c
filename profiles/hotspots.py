#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from an ncu report (source page).
   python profiles/hotspots.py gpurun_out/prof.ncu-rep <kernel-name> [N]"""
import csv, io, subprocess, sys

def main(path, kernel, n=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"hdr": None, "data": []}; blocks.append(cur); continue
        if cur is None: continue
        if cur["hdr"] is None: cur["hdr"] = r; continue
        if len(r) == len(cur["hdr"]): cur["data"].append(r)
    import os
    b = blocks[int(os.environ.get("INST", "0"))]; hdr, data = b["hdr"], b["data"]
    col = lambda name: hdr.index(name)
    iS, iSrc, iEx = col("# Samples"), col("Source"), col("Instructions Executed")
    names = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_barrier", "stall_not_selected", "stall_mio", "stall_lg"]
    idx = [col(x) for x in names]
    tot = sum(int(r[iS]) for r in data)
    ex = sum(int(r[iEx]) for r in data)
    print("kernel %s: %d SASS instrs, %d samples, %d warp-instructions executed" % (kernel, len(data), tot, ex))
    print("stall totals: " + ", ".join("%s %.1f%%" % (nm[6:], 100.0 * sum(int(r[i]) for r in data) / tot) for nm, i in zip(names, idx)))
    for r in sorted(data, key=lambda r: -int(r[iS]))[:n]:
        st = " ".join("%s=%s" % (nm[6:9], r[i]) for nm, i in zip(names[:4], idx[:4]))
        print("%6d %5.1f%% ex=%9s %-34s | %s" % (int(r[iS]), 100.0 * int(r[iS]) / tot, r[iEx], st, r[iSrc].strip()[:80]))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
