"""Per-kernel-family time and DRAM traffic from an ncu CSV with gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum (one eager step, tools/ncu_step.py).  usage: summarize_launches.py in.csv out.txt label"""
import collections, csv, re, sys
src, dst, label = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(src).readlines()
hi = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
fam = collections.OrderedDict()
def val(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v * scale
ids = collections.defaultdict(dict)
for r in csv.DictReader(lines[hi:]):
    k = re.sub(r"^(void )?(b200sr::)?", "", r["Kernel Name"].split("(")[0])
    if k.startswith("at::") or "elementwise_kernel" in k:   # ATen kernels are not part of the step
        continue
    ids[(r["ID"], k)][r["Metric Name"]] = val(r)
for (i, k), m in ids.items():
    a = fam.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in fam.values())
out = [f"ncu gpu__time_duration + dram bytes per kernel family, one eager stage-2 step (tools/ncu_step.py), {label}"]
for k, a in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:34s} launches={a[0]:5d} time={a[1] / 1e3:8.3f} ms ({100 * a[1] / tot:4.1f}%)  "
               f"dram_traffic={a[2] / 1e6:9.1f} MB  per_launch={a[2] / 1e6 / a[0]:7.2f} MB")
out.append(f"total kernel time {tot / 1e3:.3f} ms, launches {sum(a[0] for a in fam.values())}")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
