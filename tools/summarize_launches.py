"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of ONE training step.
The step window is found from the launches themselves: from one level-0 ball query (the only launch of
ball_query_kernel with a (256, 16) grid at B=16 x 8192) to the next.
usage: python tools/summarize_launches.py gpurun_out/launches_r01.csv profiles/r01_launches.md
"""
import csv
import re
import sys
from collections import OrderedDict


OURS = re.compile(r"ball_query_kernel|ball_grid_build|ball_nearest|nearest_query_kernel|prefix_len_kernel|grid_subsample_kernel|group_points|"
                  r"transpose_kernel|count_kernel|scan_kernel|fill_kernel|sort_short_kernel|sort_long_kernel|aggregate_|"
                  r"nearest_gather|pseudogrid_|reduce_partials|inverse_|bn_stats|bn_apply|bn_bwd|bn_eval|chunk_count|chunk_scan|"
                  r"chunk_fill|chunk_sort|bbox_kernel|cell_count|cell_fill|nn_query|radius_patch|vote_kernel|voxel_id|barycentre|"
                  r"gemm_tf32|bn_finalize|pospool_tiles|pospool_scatter|pospool_fwd_pipelined|tile_plan|spatial_order|wgrad_")


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(?:<[0-9, ]+>)?)", name)
    return (m.group(1) if m else name)[:70]


def main(src, dst):
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((int(r["ID"]), r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", "")) / 1e3))
    marks = [i for i, r in enumerate(rows) if "ball_query_kernel" in r[1] and r[2].replace(" ", "").startswith("(256,16")]
    if len(marks) >= 2:
        window = rows[marks[0]:marks[1]]
        note = f"launches {rows[marks[0]][0]}..{rows[marks[1]][0] - 1}: one training step (level-0 ball query to the next one)"
    else:
        window, note = rows, "whole capture (step boundary not found)"
    total = sum(r[3] for r in window)
    agg = OrderedDict()
    for _, name, grid, us in window:
        k = short(name)
        a = agg.setdefault(k, [0.0, 0, bool(OURS.search(name))])
        a[0] += us
        a[1] += 1
    ours = sum(v[0] for v in agg.values() if v[2])
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n{note}\n\n")
        f.write(f"kernels in window: {len(window)}; summed device time {total / 1e3:.3f} ms "
                f"(ncu serialises launches and runs them cold: compare SHARES, not absolutes)\n\n")
        f.write(f"share of the step spent in this repo's own kernels (libd3d_b200.so): {ours / total:.1%}\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (us, n, mine) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
            f.write(f"| {'**' + k + '**' if mine else k} | {n} | {us:.1f} | {us / total:.1%} |\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
