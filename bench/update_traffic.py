"""Write profiles/k1_traffic.json from an `ncu --set full` capture of ONE vu_fused_pass launch of the bench workload
(developer tool):   python bench/update_traffic.py gpurun_out/<tag>_k1_tma_bench.ncu-rep [images_in_that_launch]

bench.py reports `roofline.traffic` from this file only while the digest of the kernel sources recorded here equals the
digest of the sources it runs (bench.k1_source_digest)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOAD, k1_source_digest  # noqa: E402


def main():
    rep = sys.argv[1]
    images = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, row = rows[0], rows[1], rows[2]

    def val(name):
        i = hdr.index(name)
        v = float(row[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    V = WORKLOAD["spatial"][0] * WORKLOAD["spatial"][1]
    data = {"source": f"ncu --set full --clock-control none, one launch of {row[hdr.index('Kernel Name')][:60]} ({os.path.basename(rep)})",
            "k1_source_digest": k1_source_digest(), "images_in_launch": images,
            "dram_bytes_read": rd, "dram_bytes_write": wr,
            "dram_bytes_per_voxel": (rd + wr) / (V * images),
            "dram_bytes_per_launch_at_images": {str(images): rd + wr}}
    with open(os.path.join(ROOT, "profiles", "k1_traffic.json"), "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(data))


if __name__ == "__main__":
    main()
