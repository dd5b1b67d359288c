#!/usr/bin/env python
"""Print the headline fields of a bench.py JSON line:  python tools/show_bench.py gpurun_out/bench.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.4g %s  ms/step %.4f" % (d["value"], d.get("unit", ""), d["ms_per_step"]))
if "e2e" in d:
    print("e2e %.4g  ms/step %.4f" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
if "sustained" in d:
    print("sustained ms/step %.4f" % d["sustained"]["ms_per_step"])
if "roofline" in d:
    r = d["roofline"]
    print("roofline %s frac %.3f  step frac %.3f" % (r["kernel"], r["frac"], r["step"]["frac"]))
    print("phases us", {k: round(v * 1e3, 1) for k, v in r["phases_ms"].items()})
if "eval" in d:
    print("eval %.4g seqs/s  %.4f ms  frac %.3f" % (d["eval"]["value"], d["eval"]["ms"], d["eval"]["roofline"]["frac"]))
if "scoring_sweep" in d:
    print("sweep", [(p["Ls"], p["B_per_gpu"], round(p["ms"], 4), round(p["frac_of_hbm"], 3)) for p in d["scoring_sweep"]["points"]])
for k in ("dataset_resident", "movies", "sharded_10M"):
    if k in d:
        v = d[k].get("weak", d[k])
        print(k, "ms/step %.4f" % v["ms_per_step"])
