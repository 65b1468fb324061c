"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list as a markdown table:
python tools/launch_table.py gpurun_out/<name>.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = r[ik]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name).replace("gnn::", "")
    name = re.sub(r"\(int\)|\(bool\)", "", name)
    tot[name] += float(r[iv].replace(",", "")) / 1e6
    cnt[name] += 1
total = sum(tot.values())
print("| kernel | launches | total ms | ms/launch | share |")
print("|---|---:|---:|---:|---:|")
for k, v in tot.most_common():
    if v / total < 0.0005:
        continue
    print("| `%s` | %d | %.3f | %.3f | %.1f%% |" % (k, cnt[k], v, v / cnt[k], 100 * v / total))
print("| **total** | %d | %.3f | | 100%% |" % (sum(cnt.values()), total))
