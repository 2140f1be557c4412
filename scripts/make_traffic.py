"""profiles/conv_tc_traffic.json from an ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of the
tensor-core convolution launches of the bench): DRAM bytes per frame of the dominant kernels, as bench.py's roofline.traffic."""
import csv, json, sys, collections
src, dst, frames = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ni, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(float)
launches = set()
for r in rows[hdr + 1:]:
    if len(r) < len(h) or "k_conv_" not in r[ni]:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui].lower()
    if "byte" in u:
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    tot[r[mi]] += v
    launches.add(r[0])
out = dict(source=src, frames=frames, conv_launches_per_frame=len(launches) / frames,
           dram_read_bytes_per_frame=tot["dram__bytes_read.sum"] / frames, dram_write_bytes_per_frame=tot["dram__bytes_write.sum"] / frames,
           dram_bytes_per_frame=(tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / frames)
json.dump(out, open(dst, "w"), indent=1)
print(out)
