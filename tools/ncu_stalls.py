#!/usr/bin/env python
"""Top stall sites of a kernel from `ncu -i rep --page source --csv` (SASS view)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
samp = ci['# Samples']
def I(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(I(r[samp]) for r in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(I(r[ci[s]]) for r in data) for s in stalls}
print(sorted(agg.items(), key=lambda x: -x[1])[:10])
for idx, r in sorted(enumerate(data), key=lambda x: -I(x[1][samp]))[:topn]:
    st = {s: I(r[ci[s]]) for s in stalls}
    best = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(idx, r[samp], r[1][:100], best)
