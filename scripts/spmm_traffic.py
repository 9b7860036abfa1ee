"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of the aggregation launches of ONE eager
step of the benchmark mesh) -> profiles/spmm_dram_traffic_r2.json, stamped with the sha256 of the kernel sources so that
bench.py reports `roofline.traffic` only for the build the capture was taken on.
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmm \
      --launch-skip <one step of spmm launches> -c 48 --csv --log-file gpurun_out/spmm_dram_r2.csv \
      python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline
  python scripts/spmm_traffic.py gpurun_out/spmm_dram_r2.csv 224
"""
import csv, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    lid = r[ix["ID"]]
    d = per.setdefault(lid, {"kernel": r[ix["Kernel Name"]][:60]})
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    name = r[ix["Metric Name"]]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "nsecond": 1e-9,
            "usecond": 1e-6, "msecond": 1e-3, "second": 1}.get(unit, 1)
    d[name] = val * mult
launches = [d for d in per.values() if "dram__bytes_read.sum" in d]
n = len(launches)
tot = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in launches)
t = sum(d.get("gpu__time_duration.sum", 0.0) for d in launches)
src = [os.path.join(ROOT, "dual_dmp_b200", "csrc", f) for f in ("spmm.cu", "spmm_tile.cu")]
sha = hashlib.sha256(b"".join(open(p, "rb").read() for p in src)).hexdigest()
out = {"n": int(sys.argv[2]) if len(sys.argv) > 2 else 224, "launches": n, "dram_bytes_total": tot,
       "dram_bytes_per_launch_avg": tot / max(n, 1), "kernel_time_s_cold_cache": t,
       "dram_GBps_during_kernels": tot / t / 1e9 if t else None,
       "kernels": sorted({d["kernel"] for d in launches}), "source": os.path.basename(sys.argv[1]),
       "spmm_sources_sha256": sha}
json.dump(out, open(os.path.join(ROOT, "profiles", "spmm_dram_traffic_r2.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
