"""profiles/warp_kernel_traffic.json from an `ncu --set full` capture of the resampler bracket
(-k regex:"tps_warp_lattice|tps_nodes|tps_solve" on `bench.py --steps 2 --warmup 1`): DRAM bytes read + written by the
kernels of ONE bracket (solve + nodes + lattice kernel), which bench.py reports as roofline.traffic.  The file records the
SHA-1 of csrc/tps.cu at capture time; bench.py drops the number when the source has changed since.
Usage: python profiles/make_traffic_json.py gpurun_out/r2p_warp.ncu-rep H W frames Ho Wo"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, H, W, F, Ho, Wo = sys.argv[1], *[int(a) for a in sys.argv[2:7]]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__grid_size")}

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    per_kernel = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        grid = int(r[col["launch__grid_size"]].replace(",", ""))
        key = "tps_solve_kernel" if "tps_solve" in name else "tps_nodes_kernel" if "tps_nodes" in name else "tps_warp_lattice_kernel" if "tps_warp_lattice" in name else None
        if key is None:
            continue
        # the LAST launch of each kind in the capture that belongs to the 32-frame bracket (largest grid for the solve)
        if key not in per_kernel or grid >= per_kernel[key]["grid"]:
            per_kernel[key] = {"grid": grid, "dram_bytes_read": rd, "dram_bytes_write": wr, "name": name}
    out = {"kernel": "resampler bracket: " + " + ".join(sorted(per_kernel)),
           "source": "ncu --set full --clock-control none -k regex:tps_warp_lattice|tps_nodes|tps_solve on `bench.py --steps 2 --warmup 1` "
                     "(%s; summary in profiles/r02_warp_bracket_ncu_summary.txt)" % os.path.basename(rep),
           "height": H, "width": W, "frames_per_launch": F, "canvas": [Ho, Wo],
           "dram_bytes_read": sum(v["dram_bytes_read"] for v in per_kernel.values()),
           "dram_bytes_write": sum(v["dram_bytes_write"] for v in per_kernel.values()),
           "per_kernel": per_kernel,
           "tps_cu_sha1": hashlib.sha1(open(os.path.join(ROOT, "stabstitch2_b200", "csrc", "tps.cu"), "rb").read()).hexdigest()}
    json.dump(out, open(os.path.join(ROOT, "profiles", "warp_kernel_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
