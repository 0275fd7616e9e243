"""Per-source-line instruction / stall-sample totals of one kernel in an .ncu-rep captured with --import-source on
(developer tool).   python bench/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, data = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if cur and hdr and len(r) == len(hdr):
            try:
                ie, sm = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            if r[2] == "-":  # a CUDA source line (aggregated), not one of its SASS instructions
                data.append((ie, sm, cur, r[0], r[1][:100]))
    ti, ts = sum(x[0] for x in data), sum(x[1] for x in data)
    print("total warp instructions", ti, "samples", ts)
    for ie, sm, f, ln, src in sorted(data, reverse=True)[:top]:
        print(f"{100 * ie / max(ti, 1):5.1f}% inst {100 * sm / max(ts, 1):5.1f}% smp  {f}:{ln}  {src.strip()}")


if __name__ == "__main__":
    main()
