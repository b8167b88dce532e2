"""Build profiles/traffic.json (round 1: r1_traffic.json) from ncu CSVs (--metrics dram__bytes_read.sum,dram__bytes_write.sum) of
`python bench.py ... --steps 1 --warmup 1`: per workload and kernel, DRAM bytes per solve (= per bench step): the sum
over all launches of the capture divided by the number of solves in it (every solve launches k_plan_search once).
usage: python tools/traffic_json.py out.json workload=csv:"command" [workload=csv:"command" ...]"""
import csv, io, json, sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
KERNELS = ("k_tile_tensor", "k_tile_ffma", "k_direct", "k_small", "k_finalize", "k_plan_search")
out = {}
for spec in sys.argv[2:]:
    wl, rest = spec.split("=", 1)
    path, cmd = rest.split(":", 1)
    text = "".join(l for l in open(path) if l.startswith('"'))
    rows = list(csv.DictReader(io.StringIO(text)))
    per = defaultdict(lambda: defaultdict(dict))          # kernel -> launch id -> metric -> bytes
    for r in rows:
        name = r["Kernel Name"]
        k = next((k for k in KERNELS if k in name), None)
        if k is None or not r["Metric Name"].startswith("dram__bytes"):
            continue
        per[k][int(r["ID"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT[r["Metric Unit"]]
    res = {}
    n_solves = max(1, len(per.get("k_plan_search", {})))
    for k, launches in per.items():
        if k == "k_plan_search":
            continue
        rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in launches.values()) / n_solves
        wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in launches.values()) / n_solves
        res[k] = {"dram_bytes_per_step": rd + wr, "read": rd, "write": wr, "launches": len(launches) / n_solves,
                  "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, {cmd} "
                            f"(profiles/{path.split('/')[-1]}): all launches of the capture / its {n_solves} solves"}
    for k, v in res.items():                                # several captures may feed one workload (auto, then exact):
        out.setdefault(wl, {}).setdefault(k, v)             # a kernel keeps the entry of the first capture that ran it
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
