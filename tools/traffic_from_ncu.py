"""profiles/rNN_traffic.json from an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv) of bench.py at the bench configuration: DRAM bytes of the fused kernel and of every kernel of ONE step.

    python tools/traffic_from_ncu.py gpurun_out/x/launches_4096.csv 4096 > profiles/r02_traffic.json
"""
import collections
import csv
import json
import sys


def main(path, carriers, n_samples=1 << 20):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    launches = collections.OrderedDict()
    for r in rows:
        launches.setdefault(int(r[0]), {"name": r[4]})[r[12]] = float(r[14])
    seq = list(launches.values())
    # one step = from one fused-kernel launch up to (not including) the next fused-kernel launch that follows a finalize
    k1 = [i for i, l in enumerate(seq) if "k1_channelize" in l["name"]]
    fin = [i for i, l in enumerate(seq) if "k_finalize" in l["name"]]
    start = k1[0] if not any("k_edge" in l["name"] for l in seq[:k1[0]]) else min(i for i, l in enumerate(seq) if "k_edge" in l["name"])
    end = min(i for i in fin if i > k1[0]) + 1
    step = seq[start:end]
    kernels = {}
    for l in step:
        name = l["name"].replace("void ", "").replace("tetra::", "").split("(")[0]
        k = kernels.setdefault(name, {"read": 0.0, "write": 0.0, "ns": 0.0, "launches": 0})
        k["read"] += l.get("dram__bytes_read.sum", 0.0)
        k["write"] += l.get("dram__bytes_write.sum", 0.0)
        k["ns"] += l.get("gpu__time_duration.sum", 0.0)
        k["launches"] += 1
    k1_bytes = sum(v["read"] + v["write"] for n, v in kernels.items() if "k1_channelize" in n)
    tot = sum(v["read"] + v["write"] for v in kernels.values())
    samples = carriers * n_samples
    print(json.dumps({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum at the bench configuration "
                                "(%d carriers x 2^20), %s, the kernels of one step" % (carriers, path.split("/")[-1]),
                      "carriers": carriers, "n_samples": n_samples, "k1_dram_bytes_per_launch": k1_bytes, "step_dram_bytes": tot,
                      "k1_dram_bytes_per_sample": k1_bytes / samples, "step_dram_bytes_per_sample": tot / samples,
                      "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
