"""DRAM traffic of the local-correlation launches of one hot-path step from an ncu metrics CSV.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --profile-from-start off --csv --log-file gpurun_out/lc_traffic.csv python tools/profile_step.py
  python tools/lc_traffic.py gpurun_out/lc_traffic.csv profiles/r1_lc_dram_traffic.json
bench.py reads the JSON for roofline.traffic (bytes per step over the 14 local_correlation calls)."""
import collections, csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = {k: i for i, k in enumerate(rows[hdr])}
per = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) < len(h):
        continue
    name = r[h["Kernel Name"]]
    if "lc_" not in name:
        continue
    key = name.split("(")[0].replace("void ", "")
    v = float(r[h["Metric Value"]].replace(",", ""))
    unit = r[h["Metric Unit"]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
    d = per.setdefault(key, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "us": 0.0})
    m = r[h["Metric Name"]]
    if m == "dram__bytes_read.sum":
        d["dram_read_bytes"] += v * scale
        d["launches"] += 1
    elif m == "dram__bytes_write.sum":
        d["dram_write_bytes"] += v * scale
    elif m == "gpu__time_duration.sum":
        d["us"] += v * scale
tot = sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in per.values())
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (per-launch, serialised, caches flushed between launches) over tools/profile_step.py: 32 pairs, num_itr 2",
       "dram_bytes_per_step": tot, "per_kernel": per}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
