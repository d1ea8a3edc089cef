"""profiles/<round>_traffic.json from one `ncu --set full` report: DRAM bytes, time, instructions per launch of the
four kernels of a steady-state position (what bench.py's roofline.traffic reads).

    python tools/traffic_from_ncu.py gpurun_out/prof.ncu-rep profiles/r01_traffic.json [profiles/r01_ncu_full_raw.csv]
"""
import csv
import io
import json
import subprocess
import sys

rep, dst = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(txt)
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
LABELS = (("raster_spheres", ("membrane_from_field", "raster_gather", "raster_bin")),
          ("refract_membrane_hop", ("refract_lean_kernel<1, 0",)),
          ("refract_sample_ref_hop", ("refract_lean_kernel<2, 1",)),
          ("detect", ("detect_tile_kernel",)))


def val(r, key):
    v, u = float(r[col[key]].replace(",", "")), units[col[key]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(u, 1.0)
    return v * scale


out = {}
for label, needles in LABELS:
    for r in body:                      # the LAST matching launch: steady state, not position 0
        name = r[col["Kernel Name"]]
        if any(n in name for n in needles):
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            out[label] = {"kernel": name.split("(")[0][:60], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                          "gpu_time_us": val(r, "gpu__time_duration.sum"), "warp_instructions": val(r, "smsp__inst_executed.sum"),
                          "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          "registers": val(r, "launch__registers_per_thread"),
                          "source": "ncu --set full --clock-control none, one launch, cold caches (%s)" % rep.split("/")[-1]}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({k: (v["kernel"], round(v["dram_bytes"] / 1e6, 1), round(v["gpu_time_us"], 1)) for k, v in out.items()}))
