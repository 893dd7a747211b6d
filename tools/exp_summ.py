import json
rows=[json.loads(l) for l in open("gpurun_out/exp_mma.log") if l.startswith("{")]
par=[r for r in rows if r["kind"]=="parity"]
if par:
    w=max(par,key=lambda r:r["rel_err"]); print("parity cases", len(par), "worst", w)
for r in rows:
    if r["kind"]=="time": print(r["shape"], r["kernel"], r["cta"], "ms %.3f frac %.3f slow %d"%(r["ms"], r["frac"], r["slow_points"]))
