"""TEST INFRASTRUCTURE.  Builds the checker libraries:

  oracle/_build/libhmoracle.so   plain-C restatement (oracle/hm_oracle.c), always
  oracle/_ref/libhmref.so        the unmodified reference behind oracle/ref_harness.cpp, only where
                                 /root/reference exists (this container; the GPU box uses the prebuilt files)
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libhmoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhmref.so")
REF_ROOT = "/root/reference/hm-16.5rc1"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_oracle(force=False):
    src = os.path.join(HERE, "hm_oracle.c")
    if force or _stale(ORACLE_SO, [src]):
        os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-std=c99", "-Wall", "-o", ORACLE_SO, src])
    return ORACLE_SO


def build_reference(force=False):
    """Compiles the reference from /root/reference when present; returns the .so path or None."""
    if not os.path.isdir(REF_ROOT):
        return REF_SO if os.path.exists(REF_SO) else None
    if force or _stale(REF_SO, [os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "Makefile.ref")]):
        subprocess.check_call(["make", "-f", "Makefile.ref", "-j8", "all"], cwd=HERE,
                              stdout=subprocess.DEVNULL)
    return REF_SO


if __name__ == "__main__":
    print(build_oracle(force=True))
    print(build_reference())
