#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of bench.py's dominant kernel: DRAM bytes of ONE
launch, stamped with the kernel name, the workload and a hash of the kernel sources so that bench.py only
quotes it for the code and workload it was measured on (roofline.traffic is null otherwise).
usage: python tools/ncu_traffic.py report.ncu-rep WORKLOAD [out.json]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402
rep, workload = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
def gb(k):
    u, v = d[k]
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
name = d["Kernel Name"][1]
kernel = "k_rx_ws" if "k_rx_ws" in name else "k_rx_fused" if "k_rx_fused" in name else "k_detect_lean" if "k_detect_lean" in name else name
t = {"kernel": kernel, "kernel_full": name, "workload": workload, "source_sha": kernel_source_sha(),
     "dram_bytes_read": gb("dram__bytes_read.sum"), "dram_bytes_write": gb("dram__bytes_write.sum"),
     "gpu_time_ms": float(d["gpu__time_duration.sum"][1]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(d["gpu__time_duration.sum"][0].replace("second", "s").replace("msecond", "ms"), 1.0),
     "source": os.path.basename(rep), "how": "ncu --set full --clock-control none, one launch"}
t["dram_bytes_per_launch"] = t["dram_bytes_read"] + t["dram_bytes_write"]
json.dump(t, open(out, "w"), indent=1)
print(json.dumps(t))
