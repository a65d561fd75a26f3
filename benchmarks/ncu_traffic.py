"""profiles/ncu_traffic.json from an `ncu --set full` capture of one frame (benchmarks/one_frame.py):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, grouped by stage, plus each kernel's duration.
    python benchmarks/ncu_traffic.py gpurun_out/<capture>.ncu-rep profiles/ncu_traffic.json"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
label = sys.argv[3] if len(sys.argv) > 3 else rep  # what the JSON cites (the committed summary of the capture)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3,
                "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1)


kernels = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    kernels.append({"kernel": name.split("(")[0].replace("void ", ""), "us": round(val(r, "gpu__time_duration.sum"), 2),
                    "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum")})


n_frames = max(1, sum(1 for k in kernels if "raster_pair_kernel" in k["kernel"]))


def group(pred):
    """Kernels of ONE frame (the last captured one) and their DRAM bytes averaged over the captured frames."""
    sel = [k for k in kernels if pred(k["kernel"])]
    per_frame = sum(k["dram_read"] + k["dram_write"] for k in sel) / n_frames
    return sel[-(len(sel) // n_frames):], per_frame


frame = [k for k in kernels if "bsplat" in k["kernel"] or "proj_" in k["kernel"]]
# the last projection launch is the stand-alone stage call of one_frame.py
proj = [k for k in frame if "project_kernel" in k["kernel"]]
fused_proj, alone_proj = proj[0], proj[-1]
ras, ras_b = group(lambda n: "raster_" in n)
binn, bin_b = group(lambda n: any(t in n for t in ("onesweep", "bin_", "tile_finish")))
src = (f"{label} (ncu --set full --clock-control none, benchmarks/one_frame.py config3_1m_1080p; bytes per frame, "
       f"averaged over the {n_frames} captured frames)")
res = {
    "raster": {"kernel": " + ".join(sorted({k['kernel'] for k in ras})), "dram_bytes": ras_b,
               "per_kernel": ras, "source": src},
    "binning": {"kernels": " + ".join(k["kernel"].replace("bsplat::", "") for k in binn), "dram_bytes": bin_b,
                "per_kernel": binn, "source": src,
                "note": "writes of one kernel are often still in the 126 MB L2 when ncu ends the capture of that launch"},
    "projection": {"kernel": alone_proj["kernel"] + " (stand-alone stage call)",
                   "dram_bytes": alone_proj["dram_read"] + alone_proj["dram_write"],
                   "dram_bytes_read": alone_proj["dram_read"], "dram_bytes_write": alone_proj["dram_write"], "source": src},
    "projection_fused": {"kernel": fused_proj["kernel"] + " (inside the frame: epilogue products instead of the stage outputs)",
                         "dram_bytes": fused_proj["dram_read"] + fused_proj["dram_write"],
                         "dram_bytes_read": fused_proj["dram_read"], "dram_bytes_write": fused_proj["dram_write"],
                         "source": src},
}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v["dram_bytes"] for k, v in res.items()}))
