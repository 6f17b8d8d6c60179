"""Builds the GPU-routed HM-16.5 encoder used by the in-encoder parity test (BASELINE.json configs[0]).

    python integration/build_shim.py        (build container only: needs /root/reference and oracle/_ref/obj)

What it does, without copying any reference source into the repository:
  1. reads /root/reference/hm-16.5rc1/source/Lib/TLibEncoder/TEncSearch.cpp and injects three one-line forwarders
     (the same ones INTEGRATION.md shows a maintainer) into a scratch copy under integration/_build/ (git-ignored);
     likewise TLibCommon/TComRdCost.cpp with one call appended to TComRdCost::init() (the distortion-table hook);
  2. compiles that copy and integration/hm_shim.cpp (our binding code) against the reference headers;
  3. links them with the UNMODIFIED reference objects already built by oracle/Makefile.ref (oracle/_ref/obj, minus the
     stock TEncSearch.o) and video_codecs_b200/libhmb200.so into integration/_build/TAppEncoderB200.
The encoder settings are written next to the binary (bare key/value lines) for machines without /root/reference.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/hm-16.5rc1"
OUT = os.path.join(HERE, "_build")
OBJ = os.path.join(ROOT, "oracle", "_ref", "obj")
BIN = os.path.join(OUT, "TAppEncoderB200")

DECLS = '''
// ---- libhmb200 forwarders (integration/hm_shim.cpp) ----
class TComPic;
void hmb200_shim_ref_plane(TComPic* pic);
bool hmb200_shim_pattern_search(TEncSearch* self, TComPattern* key, Pel* piRefY, Int iRefStride, TComMv* lt, TComMv* rb, TComMv& rcMv, Distortion& ruiSAD);
class TComDataCU;
bool hmb200_shim_pattern_search_fast(TEncSearch* self, TComDataCU* pcCU, TComPattern* key, Pel* piRefY, Int iRefStride, TComMv* lt, TComMv* rb, TComMv& rcMv, Distortion& ruiSAD, const TComMv* pIntegerMv2Nx2NPred);
class TComYuv;
class TComMvField;
bool hmb200_shim_merge_estimation(TEncSearch* self, TComDataCU* pcCU, TComYuv* pcYuvOrg, Int iPUIdx, UInt uiAbsPartIdx, Int iWidth, Int iHeight, TComMvField* nb, UChar* dirs, Int numValid, UInt& uiInterDir, TComMvField* pacMvField, UInt& uiMergeIndex, Distortion& ruiCost);
bool hmb200_shim_pattern_search_frac(TEncSearch* self, Bool lossless, TComPattern* key, Pel* piRefY, Int iRefStride, TComMv* mvInt, TComMv& rcMvHalf, TComMv& rcMvQter, Distortion& ruiCost);
'''
FWD_SEARCH = "  if (hmb200_shim_pattern_search(this, pcPatternKey, piRefY, iRefStride, pcMvSrchRngLT, pcMvSrchRngRB, rcMv, ruiSAD)) return;\n"
FWD_FRAC = "  if (hmb200_shim_pattern_search_frac(this, bIsLosslessCoded, pcPatternKey, piRefY, iRefStride, pcMvInt, rcMvHalf, rcMvQter, ruiCost)) return;\n"
FWD_FAST = "  if (hmb200_shim_pattern_search_fast(this, pcCU, pcPatternKey, piRefY, iRefStride, pcMvSrchRngLT, pcMvSrchRngRB, rcMv, ruiSAD, pIntegerMv2Nx2NPred)) return;\n"
FWD_MERGE = "  if (hmb200_shim_merge_estimation(this, pcCU, pcYuvOrg, iPUIdx, uiAbsPartIdx, iWidth, iHeight, cMvFieldNeighbours, uhInterDirNeighbours, numValidMergeCand, uiInterDir, pacMvField, uiMergeIndex, ruiCost)) return;\n"
MERGE_ANCHOR = "  xRestrictBipredMergeCand( pcCU, iPUIdx, cMvFieldNeighbours, uhInterDirNeighbours, numValidMergeCand );\n"
FWD_PLANE = "  hmb200_shim_ref_plane(pcCU->getSlice()->getRefPic( eRefPicList, iRefIdxPred ));\n"


def inject_after_open_brace(src, anchor, line):
    """Inserts `line` right after the first '{' that follows the unique `anchor` (a function signature)."""
    assert src.count(anchor) == 1, f"anchor not unique: {anchor!r} x{src.count(anchor)}"
    i = src.index("{", src.index(anchor))
    j = src.index("\n", i) + 1
    return src[:j] + line + src[j:]


def patched_source():
    src = open(os.path.join(REF, "source/Lib/TLibEncoder/TEncSearch.cpp")).read()
    marker = '#include "TEncSearch.h"'
    assert src.count(marker) == 1
    src = src.replace(marker, marker + DECLS)
    src = inject_after_open_brace(src, "Void TEncSearch::xPatternSearch( TComPattern* pcPatternKey,", FWD_SEARCH)
    src = inject_after_open_brace(src, "Void TEncSearch::xPatternSearchFracDIF(", FWD_FRAC)
    src = inject_after_open_brace(src, "Void TEncSearch::xPatternSearchFast( TComDataCU*   pcCU,", FWD_FAST)
    src = inject_after_open_brace(src, "Void TEncSearch::xMotionEstimation( TComDataCU* pcCU,", FWD_PLANE)
    assert src.count(MERGE_ANCHOR) == 1
    src = src.replace(MERGE_ANCHOR, MERGE_ANCHOR + FWD_MERGE)          # xMergeEstimation: ahead of the candidate loop (:2868)
    return src


def patched_rdcost():
    """TComRdCost.cpp with one call appended to TComRdCost::init(): the distortion-table hook."""
    src = open(os.path.join(REF, "source/Lib/TLibCommon/TComRdCost.cpp")).read()
    marker = '#include "TComRdCost.h"'
    assert src.count(marker) == 1
    src = src.replace(marker, marker + "\nvoid hmb200_shim_dist_table(FpDistFunc* table, int n);\n")
    start = src.index("Void TComRdCost::init()")
    tail = "  m_iCostScale                 = 0;\n}"
    end = src.index(tail, start)
    return src[:end] + "  m_iCostScale                 = 0;\n  hmb200_shim_dist_table(m_afpDistortFunc, DF_TOTAL_FUNCTIONS);\n}" + src[end + len(tail):]


def build(force=False):
    if not os.path.isdir(REF):
        return BIN if os.path.exists(BIN) else None
    lib = os.path.join(ROOT, "video_codecs_b200", "libhmb200.so")
    deps = [os.path.join(HERE, "hm_shim.cpp"), os.path.abspath(__file__), os.path.join(ROOT, "include", "hmb200.h"), lib]
    if not force and os.path.exists(BIN) and all(os.path.getmtime(d) <= os.path.getmtime(BIN) for d in deps if os.path.exists(d)):
        return BIN
    if not os.path.isdir(OBJ):
        from oracle import build_oracle
        build_oracle.build_reference()
    os.makedirs(OUT, exist_ok=True)
    patched = os.path.join(OUT, "TEncSearch_shim.cpp")
    open(patched, "w").write(patched_source())
    inc = ["-I" + os.path.join(REF, "source/Lib"), "-I" + os.path.join(REF, "source/Lib/TLibEncoder"),
           "-I" + os.path.join(REF, "source/Lib/TLibCommon"), "-I" + os.path.join(ROOT, "include")]
    flags = ["-O3", "-fPIC", "-w", "-DMSYS_LINUX", "-D_LARGEFILE64_SOURCE", "-D_FILE_OFFSET_BITS=64", "-DMSYS_UNIX_LARGEFILE"]
    subprocess.check_call(["g++"] + flags + inc + ["-c", patched, "-o", os.path.join(OUT, "TEncSearch_shim.o")])
    os.remove(patched)                      # the scratch copy of the reference source does not stay in the tree
    patched_rd = os.path.join(OUT, "TComRdCost_shim.cpp")
    open(patched_rd, "w").write(patched_rdcost())
    subprocess.check_call(["g++"] + flags + inc + ["-c", patched_rd, "-o", os.path.join(OUT, "TComRdCost_shim.o")])
    os.remove(patched_rd)
    subprocess.check_call(["g++"] + flags + inc + ["-c", os.path.join(HERE, "hm_shim.cpp"), "-o", os.path.join(OUT, "hm_shim.o")])
    objs = []
    for sub in ("Lib/TLibCommon", "Lib/TLibEncoder", "Lib/TLibVideoIO", "Lib/TAppCommon", "Lib/libmd5", "App/TAppEncoder"):
        d = os.path.join(OBJ, sub)
        objs += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".o") and f not in ("TEncSearch.o", "TComRdCost.o")]
    subprocess.check_call(["g++", "-o", BIN, os.path.join(OUT, "TEncSearch_shim.o"), os.path.join(OUT, "TComRdCost_shim.o"),
                           os.path.join(OUT, "hm_shim.o")] + objs +
                          ["-L" + os.path.dirname(lib), "-lhmb200", "-Wl,-rpath,$ORIGIN/../../video_codecs_b200"])
    write_settings(os.path.join(OUT, "lowdelay_P_settings.cfg"))
    write_settings(os.path.join(OUT, "randomaccess_main10_settings.cfg"), "encoder_randomaccess_main10.cfg")
    return BIN


def write_settings(path, cfg="encoder_lowdelay_P_main.cfg"):
    """The encoder settings of BASELINE.json configs[0] (the values of the stock lowdelay-P main configuration) or
    configs[3] (random-access Main10), reduced to bare `Key : value` lines so that the test can run where
    /root/reference does not exist."""
    keep = []
    for line in open(os.path.join(REF, "cfg", cfg)):
        line = line.split("#", 1)[0].strip()
        if ":" in line:
            k, v = line.split(":", 1)
            keep.append(f"{k.strip()} : {' '.join(v.split())}")
    open(path, "w").write("\n".join(sorted(keep, key=lambda l: (not l.startswith("Frame"), l))) + "\n")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    print(build(force=True))
