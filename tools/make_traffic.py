#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --page raw --csv` export of bench.py's headline workload (1 GiB mixed per GPU):
DRAM bytes per launch of K1 and K2, tagged with the library build (b2lz4_version(), which carries a hash of the kernel
sources) so that bench.py can refuse the file once the kernels change.
usage: make_traffic.py raw.csv [out.json]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = {"_about": "dram__bytes_read.sum + dram__bytes_write.sum per launch from an `ncu --set full --clock-control none` capture "
                     "of bench.py's workload (1 GiB/GPU mixed, 64 KiB blocks); bench.py copies these into roofline.traffic when it "
                     "runs the same workload with the same build.  Written by tools/make_traffic.py.",
           "bytes_per_gpu": 1 << 30}
    import zig_lz4_b200 as z
    out["b2lz4_version"] = z.lib().b2lz4_version().decode()
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        return float(r[col[name]]) * UNIT.get(units[col[name]], 1)

    # the longest launch of each kernel: K1 also runs a short estimate pass (first KiB of every block) before the real one
    best = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = "k_compress_fast" if "k_compress_fast" in name else ("k_decompress" if "k_decompress" in name else None)
        if key and (key not in best or val(r, "gpu__time_duration.sum") > val(best[key], "gpu__time_duration.sum")):
            best[key] = r
    for key, r in best.items():
        name = r[col["Kernel Name"]]
        out[key] = {"traffic_bytes": int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")),
                    "dram_read_bytes": int(val(r, "dram__bytes_read.sum")), "dram_write_bytes": int(val(r, "dram__bytes_write.sum")),
                    "ncu_duration_ms": round(val(r, "gpu__time_duration.sum") * (1e-6 if units[col["gpu__time_duration.sum"]] == "ns" else 1e-3 if units[col["gpu__time_duration.sum"]] == "us" else 1), 3),
                    "inst_executed": int(float(r[col["smsp__inst_executed.sum"]])),
                    "issue_active_pct": round(float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]), 1),
                    "l2_hit_pct": round(float(r[col["lts__t_sector_hit_rate.pct"]]), 1),
                    "shared_bank_conflicts": int(float(r[col["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]])),
                    "registers": int(float(r[col["launch__registers_per_thread"]])),
                    "kernel": name.split("(")[0]}
    path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
