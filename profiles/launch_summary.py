"""Turns an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv python bench.py ...`) into the
per-kernel table tracked under profiles/ (our kernels only; torch's input generators are dropped).
usage: python profiles/launch_summary.py gpurun_out/launches.csv profiles/rN_launches_bench_step.csv"""
import csv
import sys
from collections import OrderedDict

OURS = ("preprocess", "decode_kernel", "nms_", "finalize", "estimate", "warp_", "detect_fused", "select_kernel", "radix", "adjacency",
        "grid_", "cell_", "make_keys", "iota_keys", "gather_sorted", "map_keep", "low32", "scan_kernel", "crops_to_tensor", "invert_kernel", "resize_u8")


def main(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[kn].split("(")[0].replace("void ", "").replace("fd::", "")
        if not any(k in name for k in OURS):
            continue
        us = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "mean_us", "total_us", "share"])
        for k, (n, us) in agg.items():
            w.writerow([k, n, "%.2f" % (us / n), "%.2f" % us, "%.3f" % (us / tot)])
    print(open(dst).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
