"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on): share of the warp-stall samples and
of the executed instructions by CUDA source line.   usage: python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    sa, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
    cur_file = ""
    lines = []
    for r in rows:
        if r and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        if len(r) > ie and r[0].isdigit() and r[2] == "-" and r[sa].isdigit():
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[sa]), int(r[ie] or 0)))
    ts, ti = sum(x[3] for x in lines) or 1, sum(x[4] for x in lines) or 1
    print(f"total stall samples {ts}, warp instructions {ti}")
    for f, n, src, s, i in lines:
        if s * 100 / ts >= min_pct or i * 100 / ti >= min_pct:
            print(f"{f}:{n:<4d} samples {s * 100 / ts:5.1f}%  instr {i * 100 / ti:5.1f}%  {src[:100]}")


if __name__ == "__main__":
    main()
