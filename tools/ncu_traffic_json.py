"""profiles/ncu_den_traffic.json (what bench.py reports as roofline.traffic / l2_traffic) from the raw page of an
`ncu --set full` capture of one den_forward + den_backward launch pair:
    python tools/ncu_traffic_json.py profiles/r02_ncu_den_kernels_raw.csv > profiles/ncu_den_traffic.json"""
import csv
import json
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]

    def col(name):
        return hdr.index(name)

    def to_bytes(v, unit):
        return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

    def to_ms(v, unit):
        return float(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)

    kernels = []
    for r in rows[2:]:
        name = r[col("Kernel Name")]
        if "den_" not in name:
            continue
        kernels.append({
            "name": name,
            "dram_read_bytes": to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")]),
            "dram_write_bytes": to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")]),
            "duration_ms": to_ms(r[col("gpu__time_duration.sum")], units[col("gpu__time_duration.sum")]),
            "lts_sectors": float(r[col("lts__t_sectors.sum")]),
            "registers": int(float(r[col("launch__registers_per_thread")])),
            "inst_executed": float(r[col("smsp__inst_executed.sum")]),
            "issue_active_pct": float(r[col("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
        })
    out = {
        "dram_bytes_per_step": sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in kernels),
        "l2_bytes_per_step": 32.0 * sum(k["lts_sectors"] for k in kernels),
        "source": "ncu --set full --clock-control none, den_forward_kernel + den_backward_kernel of one bench.py step "
                  "(%s): dram__bytes_read.sum + dram__bytes_write.sum; l2 = lts__t_sectors.sum x 32 B" % path,
        "kernels": kernels,
    }
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
