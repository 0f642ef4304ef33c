"""Measured DRAM traffic of the step kernel from an ncu --set full capture -> traffic_bytes_per_env.json, stamped with the sha of the
kernel sources so that bench.py only reports it for the build it was measured on.
usage: python tools/update_traffic.py <rep.ncu-rep> <key e.g. flip | mix_dr> <envs> <capture tag> [out.json]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rep, key, n, tag = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
out_path = sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "gpurun_out", "traffic_bytes_per_env.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if "fpv_step_kernel" in r[4]]


def col(name):
    i = hdr.index(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return [float(r[i].replace(",", "")) * scale for r in data]


rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
per_env = sum(a + b for a, b in zip(rd, wr)) / len(rd) / n
from bench import kernel_source_sha  # noqa: E402
d = json.load(open(out_path)) if os.path.exists(out_path) else {}
if d.get("kernel_source_sha16") != kernel_source_sha():
    d = {}
d.update({key: round(per_env, 1), "kernel_source_sha16": kernel_source_sha(), "capture": tag,
          "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch / envs, ncu --set full --clock-control none, strict build"})
json.dump(d, open(out_path, "w"), indent=1)
print(key, per_env, "B/env-step over", len(rd), "launches")
