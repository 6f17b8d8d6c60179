"""Per-kernel register / stack / static-shared summary of the built library.

usage: python tools/resource_usage.py [video_codecs_b200/libhmb200.so] > profiles/rNN_resource_usage.txt
Runs on the CPU container (cuobjdump reads the embedded sm_100a cubin).
"""
import re
import subprocess
import sys


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else "video_codecs_b200/libhmb200.so"
    dump = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True,
                          check=True).stdout
    rows, name = [], None
    for l in dump.splitlines():
        l = l.strip()
        if l.startswith("Function "):
            name = l[len("Function "):].rstrip(":")
        elif l.startswith("REG:") and name:
            rows.append((name, l))
            name = None
    names = subprocess.run(["c++filt"], input="\n".join(n for n, _ in rows), capture_output=True,
                           text=True, check=True).stdout.splitlines()
    print(f"# cuobjdump --dump-resource-usage {so} (sm_100a), one line per kernel")
    print("# REG = registers per thread, STACK = bytes of per-thread stack frame, SHARED = static shared bytes")
    print("# (the TMA-staged search window is dynamic shared memory, set at launch, and not listed here)")
    for n, (_, r) in sorted(zip(names, rows)):
        n = re.sub(r"\(.*", "", n).replace("hmb200::", "").replace("void ", "")
        r = re.sub(r" (LOCAL|TEXTURE|SURFACE|SAMPLER):0", "", r)
        print(f"{n:58s} {r}")


if __name__ == "__main__":
    main()
