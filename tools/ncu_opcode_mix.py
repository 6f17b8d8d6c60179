"""Per-opcode executed-instruction counts and stall-sample shares of ONE kernel from an ncu report taken with
`--set full --import-source on` (source page, SASS view):

    python tools/ncu_opcode_mix.py gpurun_out/x.ncu-rep > profiles/rNN_opcode_mix_<kernel>.txt

Needs ncu on PATH; no GPU.  The VABSDIFF4 row is the number of SAD instructions the kernel really issued (x 128 byte abs-diffs
per warp instruction = the `issued` abs-diffs of bench.py's roofline)."""
import collections
import csv
import re
import subprocess
import sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    name = rows[0][1]
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0
    tot_inst = sum(f(r, "Instructions Executed") for r in data)
    tot_samp = sum(f(r, "# Samples") for r in data)
    print("kernel:", name)
    print(f"warp instructions executed: {tot_inst:.0f}   stall samples: {tot_samp:.0f}   static SASS instructions: {len(data)}")
    op = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        if m:
            op[m.group(2).split(".")[0]] += f(r, "Instructions Executed")
    print("\nopcode          warp instructions   share")
    for k, v in op.most_common(16):
        print(f"{k:14s} {v:18.0f} {100 * v / tot_inst:6.1f}%")
    print("\nstall reason (all samples)        share")
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for k, v in sorted(((h, sum(f(r, h) for r in data)) for h in st), key=lambda kv: -kv[1])[:9]:
        print(f"{k:30s} {100 * v / tot_samp:6.1f}%")
    # the straight-line tile: the longest run of instructions executed more than half as often as the hottest one
    mx = max(f(r, "Instructions Executed") for r in data)
    best, cur = (0, 0), None
    for i, r in enumerate(data + [None]):
        hot = r is not None and f(r, "Instructions Executed") > 0.5 * mx
        if hot and cur is None:
            cur = i
        if not hot and cur is not None:
            if i - cur > best[1] - best[0]:
                best = (cur, i)
            cur = None
    a, b = best
    inst = sum(f(r, "Instructions Executed") for r in data[a:b])
    samp = sum(f(r, "# Samples") for r in data[a:b])
    print(f"\nhot tile: {b - a} SASS instructions in a row = {100 * inst / tot_inst:.1f}% of the executed instructions, {100 * samp / tot_samp:.1f}% of the stall samples")


if __name__ == "__main__":
    main(sys.argv[1])
