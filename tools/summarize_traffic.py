"""Per-kernel DRAM traffic and duration of ONE call of the hot path, from an ncu launch list taken with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file out.csv python tools/profile_call.py <pairs>
(the program makes a warm-up call and then the call that is summarised: the second half of the launch list).
Writes a JSON with, per kernel: launches, total ms, DRAM bytes read / written, achieved DRAM GB/s, and the path's
measured bytes per pixel per pair.
    python tools/summarize_traffic.py out.csv <pairs> <width> <height> > profiles/r02_traffic.json
"""
import collections
import csv
import json
import sys


def main(path, pairs, W, H):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0].replace("void ", "")})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[unit]
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            d["rd" if "read" in r["Metric Name"] else "wr"] = v * mult
    ls = [d for d in launches.values() if d["name"] != "k_synth_frames"]
    ls = ls[len(ls) // 2:]  # the second call
    agg = collections.OrderedDict()
    for d in ls:
        a = agg.setdefault(d["name"], {"launches": 0, "ms": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        a["launches"] += 1
        a["ms"] += d["us"] / 1e3
        a["dram_read"] += d.get("rd", 0.0)
        a["dram_write"] += d.get("wr", 0.0)
    tot_ms = sum(a["ms"] for a in agg.values())
    tot_b = sum(a["dram_read"] + a["dram_write"] for a in agg.values())
    out = {"what": f"one dofs3d_process_dev call of {pairs} pairs of {W}x{H}, ncu launch list (cold-cache, serialised)",
           "pairs": pairs, "launches": len(ls), "kernel_ms": round(tot_ms, 3), "dram_bytes": tot_b,
           "dram_bytes_per_pixel_per_pair": round(tot_b / (pairs * W * H), 1), "kernels": {}}
    for k, a in sorted(agg.items(), key=lambda kv: -(kv[1]["dram_read"] + kv[1]["dram_write"])):
        b = a["dram_read"] + a["dram_write"]
        out["kernels"][k] = {"launches": a["launches"], "ms": round(a["ms"], 3), "share_of_time": round(a["ms"] / tot_ms, 4),
                             "dram_read_MB": round(a["dram_read"] / 1e6, 1), "dram_write_MB": round(a["dram_write"] / 1e6, 1),
                             "bytes_per_pixel_per_pair": round(b / (pairs * W * H), 2),
                             "dram_GBps": round(b / (a["ms"] / 1e3) / 1e9, 1) if a["ms"] > 0 else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
