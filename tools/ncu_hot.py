"""Summarise an ncu --page source --csv dump: top instructions by stall samples with their stall mix.
usage: ncu -i X.ncu-rep --page source --csv > /tmp/x.csv; python tools/ncu_hot.py /tmp/x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        n = int(r[ci["# Samples"]])
    except (ValueError, IndexError):
        continue
    data.append((n, r))
tot = sum(n for n, _ in data) or 1
print("total samples", tot)
agg = {s: sum(int(r[ci[s]] or 0) for _, r in data) for s in stalls}
print("stall mix:", ", ".join(f"{k[6:]}={100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    mix = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{100 * n / tot:5.1f}%  {r[ci['Source']].strip()[:90]:90s} {mix}")
