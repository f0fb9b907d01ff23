"""Summarise an ncu launch list (csv of gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum]
per launch, `--kernel-name-base demangled`) into a markdown table: per kernel launches, total time, share of the
step, DRAM traffic per launch.  usage: python tools/summarize_launches.py launches.csv [out.md] [traffic.json]"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per_launch = collections.OrderedDict()   # launch ID -> dict(name, metrics)
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    e = per_launch.setdefault(d["ID"], {"name": d["Kernel Name"].split("(")[0].replace("void ", "").replace("yv6::", "")})
    val = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"]
    name = d["Metric Name"]
    if name == "gpu__time_duration.sum":
        e["us"] = val / 1e3 if unit == "ns" else (val if unit == "us" else val * 1e3)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        e[name] = val * scale
agg = collections.OrderedDict()
for e in per_launch.values():
    a = agg.setdefault(e["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1
    a["us"] += e.get("us", 0.0)
    a["rd"] += e.get("dram__bytes_read.sum", 0.0)
    a["wr"] += e.get("dram__bytes_write.sum", 0.0)
tot = sum(a["us"] for a in agg.values()) or 1.0
out = ["| kernel | launches | total us | share | avg us | DRAM read MB/launch | DRAM write MB/launch |", "|---|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    out.append(f"| {k} | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['us'] / a['n']:.1f} | "
               f"{a['rd'] / a['n'] / 1e6:.2f} | {a['wr'] / a['n'] / 1e6:.2f} |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        f.write(f"# ncu launch list summary ({len(per_launch)} launches; cold-cache, serialised per-launch times)\n\n" + text + "\n")
if len(sys.argv) > 3:
    conv = [a for k, a in agg.items() if k.startswith("conv_igemm_kernel")]
    n = sum(a["n"] for a in conv)
    if n:
        json.dump({"kernel": "yv6::conv_igemm_kernel", "launches": n,
                   "dram_bytes_per_launch": (sum(a["rd"] + a["wr"] for a in conv)) / n,
                   "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, averaged over the conv launches of the capture"},
                  open(sys.argv[3], "w"), indent=1)
