#!/usr/bin/env python
"""profiles/ncu_traffic.json <- dram__bytes_read.sum + dram__bytes_write.sum per launch of the `ncu --set full`
captures gpurun_out/<tag>_prof_<kernel>.ncu-rep (mean over the captured launches).
usage: scripts/ncu_traffic.py <tag> [workload]"""
import csv, glob, json, os, subprocess, sys

FAMILY = {"fused_layer_kernel": "edge_update_fused", "node_gemm_kernel": "node_gemm_pre", "knn_query_f32key_kernel": "knn_query",
          "knn_query_kernel": "knn_query", "tail_reduce_kernel": "edge_tail_reduce", "edge_aggregate_split_kernel": "edge_aggregate",
          "fill_slots_features_knn_kernel": "csc_build_edge_attr"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tag = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "headline_100k_k16_4x64"
path = os.path.join("profiles", "ncu_traffic.json")
doc = json.load(open(path)) if os.path.exists(path) else {}
entry = {}
for rep in sorted(glob.glob(f"gpurun_out/{tag}_prof_*.ncu-rep")):
    kernel = os.path.basename(rep)[len(tag) + 6:-8]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = []
    for r in rows[2:]:
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[ix[k]].replace(",", "")) * UNIT[units[ix[k]]]
        tot.append(b)
    entry[FAMILY.get(kernel, kernel)] = int(sum(tot) / len(tot))
doc[workload] = entry
doc["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (mean of the captured launches) of the ncu --set full "
                   "captures; cold L2: ncu flushes the caches before every launch")
doc["_source"] = f"profiles/{tag}_prof_*.txt (ncu --set full captures of this round, scripts/ncu_traffic.py {tag})"
json.dump(doc, open(path, "w"), indent=1)
print(json.dumps(entry))
