#!/usr/bin/env python
"""profiles/<tag>_kernel_traffic.json from an ncu kernel summary (tools/ncu_summary.py output): DRAM bytes per launch of every
kernel category of one evaluation and their sum (what bench.py's roofline.traffic reads).

    python tools/traffic_json.py profiles/r02_ncu_kernels_final.csv single_e 4096 fp16 1039552 > profiles/r02_kernel_traffic.json
"""
import csv
import json
import sys

src, workload, events, precision, cells = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
rows = list(csv.reader(open(src)))
h = rows[0]
name, rd, wr = 0, [i for i, c in enumerate(h) if c.startswith("dram__bytes_read.sum")][0], [i for i, c in enumerate(h) if c.startswith("dram__bytes_write.sum")][0]
scale = {"[Gbyte]": 1e9, "[Mbyte]": 1e6, "[Kbyte]": 1e3, "[byte]": 1.0}
sr = [v for k, v in scale.items() if k in h[rd]][0]
sw = [v for k, v in scale.items() if k in h[wr]][0]
# kernel -> (category, launches per evaluation); the capture window may hold a kernel of the neighbouring evaluation too: averages per kernel, then the counts
kern = {"embed_tc": ("embed", 1), "context_rows": ("embed", 1), "gemm_f32_big": ("adaln", 1), "modpq": ("adaln", 1), "attn3": ("attn", None), "head_fused": ("head", 1),
        "head_prep": ("head", 1), "head_chain": ("head", 1), "layer_chain_kernel<1, 1": ("feat0", 1), "layer_chain_kernel<0, 1": ("feat0", 1), "layer_chain_kernel": ("chain", None)}
seen = {}
for r in rows[1:]:
    if not r:
        continue
    k = next((k for k in kern if k in r[name]), None)
    if k is not None:
        seen.setdefault(k, []).append(float(r[rd]) * sr + float(r[wr]) * sw)
launch, total, counts = {}, 0.0, {}
for k, v in seen.items():
    c, n = kern[k]
    n = len(v) if n is None else n                 # attention / chain: every launch of the window belongs to one evaluation (6 layers)
    avg = sum(v) / len(v)
    launch[c] = launch.get(c, 0.0) + avg if kern[k][1] == 1 else avg
    counts[c] = counts.get(c, 0) + n
    total += avg * n
print(json.dumps({"source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch ({src})", "workload": workload, "events_per_gpu": events,
                  "precision": precision, "dram_bytes_per_launch": launch, "launches_per_evaluation": counts,
                  "dram_bytes_per_evaluation": total, "cells": cells, "dram_bytes_per_cell_per_evaluation": total / cells}, indent=1))
