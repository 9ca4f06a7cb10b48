"""Summarise an ncu launch list (CSV from `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv --log-file ...`)
per kernel: launches, total time, share, DRAM bytes.   python tools/summarize_launches.py <csv> [out.md] [traffic.json]"""
import csv
import json
import re
import sys
from collections import defaultdict

src = sys.argv[1]
rows = {}
with open(src, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    k = rows.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        k["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    elif m.startswith("dram__bytes"):
        k[m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    return re.sub(r"^aab::", "", n)


agg = defaultdict(lambda: {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
for k in rows.values():
    fam = re.sub(r"<.*$", "", short(k["name"]))
    a = agg[fam]
    a["n"] += 1
    a["us"] += k.get("us", 0.0)
    a["rd"] += k.get("dram__bytes_read.sum", 0.0)
    a["wr"] += k.get("dram__bytes_write.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
out = [f"launches {len(rows)}, total {tot / 1e3:.2f} ms (ncu: serialised, cold-cache; compare shares)\n\n",
       "| kernel | launches | total ms | share | DRAM read GB | DRAM write GB | DRAM bytes / launch |\n|---|---|---|---|---|---|---|\n"]
for fam, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    out.append(f"| {fam} | {a['n']} | {a['us'] / 1e3:.3f} | {100 * a['us'] / tot:.1f} % | {a['rd'] / 1e9:.3f} | {a['wr'] / 1e9:.3f} | "
               f"{(a['rd'] + a['wr']) / max(a['n'], 1) / 1e6:.1f} MB |\n")
text = "".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "a").write(text)
if len(sys.argv) > 3 and "igemm_kernel" in agg:
    a = agg["igemm_kernel"]
    json.dump({"igemm_dram_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"], "igemm_launches": a["n"],
               "igemm_dram_read_bytes_total": a["rd"], "igemm_dram_write_bytes_total": a["wr"],
               "source": src, "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every igemm_kernel launch of "
               "one UNet forward (config 2), summed and divided by the launch count"}, open(sys.argv[3], "w"), indent=1)
