"""Extract dram read+write bytes per launch of the FI kernels from ncu reports of bench.py and
write profiles/traffic.json (read by bench.py for roofline.traffic).
    python tools/ncu_traffic.py gpurun_out/bench_fi_bwd.ncu-rep gpurun_out/bench_fi_fwd.ncu-rep"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"source": "ncu --set full --clock-control none -k regex:fi_bwd_rows | fi_fwd_patch  -s 3 -c 1  python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-networks (tools/profile_round.sh)",
       "workload": "B=4 x 1920x1080, C=3, fs=4 (bench.py)"}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        def get(m):
            i = hdr.index(m)
            v = float(vals[i]); u = units[i]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tot = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
        key = "fi_bwd_bytes_per_launch" if "bwd" in name else "fi_fwd_bytes_per_launch"
        out[key] = tot
        out[key.replace("bytes_per_launch", "ncu_time_us")] = float(vals[hdr.index("gpu__time_duration.sum")])
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(out)
