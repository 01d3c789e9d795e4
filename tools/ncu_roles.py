"""Per-segment stall-sample summary of an ncu source page (SASS), split at branch/barrier instructions."""
import collections
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    return h, rows[2:]


def main(rep, min_samples=300):
    h, data = load(rep)
    idx = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    total = sum(int(r[idx["# Samples"]] or 0) for r in data)
    print("total samples", total)
    tot = collections.Counter()
    for r in data:
        for s in stalls:
            tot[s] += int(r[idx[s]] or 0)
    print({k: "%.1f%%" % (100 * v / total) for k, v in tot.most_common(8)})
    seg, ops, start, ex_max = 0, collections.Counter(), 0, 0
    for n, r in enumerate(data):
        src = r[idx["Source"]].strip()
        smp = int(r[idx["# Samples"]] or 0)
        ex = int(r[idx["Instructions Executed"]] or 0)
        seg += smp
        ex_max = max(ex_max, ex)
        tok = src.split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        ops[op.split(".")[0]] += ex
        if any(k in src for k in ("BRA", "BAR.", "EXIT")):
            if seg >= min_samples:
                print(f"[{start:5d}-{n:5d}] samples={seg:6d} ({100*seg/total:4.1f}%) execs_max={ex_max:8d} end='{src[:40]}' "
                      f"mix={dict((k, v) for k, v in ops.most_common(6))}")
            seg, ops, start, ex_max = 0, collections.Counter(), n + 1, 0


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 300)
