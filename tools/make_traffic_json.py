"""`ncu --set full` raw-page CSV of ONE 1024^2 encode+tag pass (tools/ncu_full.sh) -> profiles/r02_dram_traffic.json:
DRAM bytes and time per tensor-kernel family, plus the sha256 of the kernel sources of the build that was profiled
(bench.kernel_source_sha) -- bench.py prints `roofline.traffic` only when that matches the build it is running.

    python tools/make_traffic_json.py gpurun_out/<tag>_raw.csv profiles/r02_dram_traffic.json
"""
import csv
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_sha  # noqa: E402

TENSOR = re.compile(r"conv3_fused_kernel|igemm_kernel|flash_d512_kernel|conv_in_kernel")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr = rows[0]
    col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    units = rows[1]
    scale = {}
    for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        u = units[col[n]].lower()
        scale[n] = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, None)
        assert scale[n] is not None, f"unexpected unit {u!r} for {n}"
    tu = units[col["gpu__time_duration.sum"]].lower()
    tscale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}[tu]
    by = {}
    for r in rows[2:]:
        if len(r) != len(hdr) or not TENSOR.search(r[col["Kernel Name"]]):
            continue
        name = re.sub(r"^.*?(conv3_fused_kernel<[^>]*>|igemm_kernel<[^>]*>|flash_d512_kernel|conv_in_kernel).*$", r"\1",
                      r[col["Kernel Name"]])
        d = by.setdefault(name, {"launches": 0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "time_us": 0.0})
        d["launches"] += 1
        d["dram_read_MB"] += float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale["dram__bytes_read.sum"]
        d["dram_write_MB"] += float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale["dram__bytes_write.sum"]
        d["time_us"] += float(r[col["gpu__time_duration.sum"]].replace(",", "")) * tscale
    tot = {k: sum(d[k] for d in by.values()) for k in ("launches", "dram_read_MB", "dram_write_MB", "time_us")}
    out = {"source": f"ncu --set full --clock-control none, tools/ncu_full.sh (ONE 1024x1024 image, one encode+tag pass; every "
                     f"tcgen05 kernel launch); raw page {os.path.basename(src)}",
           "kernel_source_sha256": kernel_source_sha(), "per_image_by_kernel": by, "per_image_all_tensor_kernels": tot}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(tot), "->", dst)


if __name__ == "__main__":
    main()
