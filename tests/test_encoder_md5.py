"""BASELINE.json configs[0]: HM-16.5 TAppEncoder, encoder_lowdelay_P_main.cfg, synthetic 416x240 8-bit,
FastSearch=0 SearchRange=64 — with every xPatternSearch / xPatternSearchFracDIF call routed through libhmb200
(integration/hm_shim.cpp), the bitstream md5 and every decoded-picture MD5 must equal the stock reference encoder's
(tests/golden/encoder_md5.json, produced by tests/golden/make_encoder_golden.py from the unmodified reference)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_encoder_golden as meg  # noqa: E402  (clip generator + command line only; no arithmetic of the path)

GOLD = os.path.join(ROOT, "tests", "golden", "encoder_md5.json")
BIN = os.path.join(ROOT, "integration", "_build", "TAppEncoderB200")
CFG = os.path.join(ROOT, "integration", "_build", "lowdelay_P_settings.cfg")


def _need_binary():
    if not (os.path.exists(BIN) and os.path.exists(CFG)):
        pytest.skip("integration/_build/TAppEncoderB200 not built (python integration/build_shim.py in the build container)")


def test_golden_file_is_consistent():
    gold = json.load(open(GOLD))
    for frames, g in gold.items():
        assert len(g["picture_md5"]) == int(frames) and len(g["bitstream_md5"]) == 32


@pytest.mark.gpu
@pytest.mark.parametrize("frames,mode,merge", [(3, "verify", "1"), (8, "gpu", "0"), (8, "gpu", "1")])
def test_encoder_bitstream_md5_matches_reference(tmp_path, frames, mode, merge):
    """merge = "1": xMergeEstimation's candidate loop goes through hmb200_merge_estimation_batch as well (verify mode compares
    every candidate's prediction error with xGetInterPredictionError)."""
    _need_binary()
    gold = json.load(open(GOLD))[str(frames)]
    yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
    meg.write_clip(yuv, frames)
    assert hashlib.md5(open(yuv, "rb").read()).hexdigest() == gold["yuv_md5"], "synthetic clip differs from the golden run's"
    env = dict(os.environ, HMB200_SHIM=mode, HMB200_SHIM_MERGE=merge)
    p = subprocess.run([BIN] + meg.encoder_args(CFG, yuv, frames, binf), capture_output=True, text=True, env=env, timeout=1500)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "integer searches" in p.stderr and " 0 integer searches" not in p.stderr, "the GPU path was not exercised"
    assert ("merge estimations" in p.stderr) == (merge == "1")
    assert [list(x) for x in meg.parse_md5_lines(p.stdout)] == [list(x) for x in gold["picture_md5"]]
    assert hashlib.md5(open(binf, "rb").read()).hexdigest() == gold["bitstream_md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("knob", ["HMB200_NO_SPECULATION", "HMB200_NO_ONE_FAST"])
def test_encoder_bitstream_md5_same_without_the_latency_paths(tmp_path, knob):
    """The in-encoder 1:1 path has three forms - whole-CU speculation (default), one launch set per PU (HMB200_NO_SPECULATION=1),
    round 1's kernels with a copy + stream synchronisation per call (HMB200_NO_ONE_FAST=1): the bitstream must not depend on it."""
    _need_binary()
    frames = 3
    gold = json.load(open(GOLD))[str(frames)]
    yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
    meg.write_clip(yuv, frames)
    env = dict(os.environ, HMB200_SHIM="gpu", HMB200_SHIM_MERGE="0")
    env[knob] = "1"
    p = subprocess.run([BIN] + meg.encoder_args(CFG, yuv, frames, binf), capture_output=True, text=True, env=env, timeout=1500)
    assert p.returncode == 0, p.stderr[-2000:]
    assert " 0 were answered from one" in p.stderr, "the knob did not switch the speculation off"
    assert hashlib.md5(open(binf, "rb").read()).hexdigest() == gold["bitstream_md5"]


@pytest.mark.gpu
def test_encoder_bitstream_md5_1080p_prefix(tmp_path):
    """Same check at BASELINE.json configs[2]'s geometry: 1920x1080 (ConformanceWindowMode=1 -> 1088 coded rows), I + P."""
    _need_binary()
    from video_codecs_b200 import synth
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "encoder_md5_1080p.json")))
    yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
    synth.write_yuv420(yuv, [synth.luma_frame(1920, 1080, t, seed=77) for t in range(2)], 8)
    assert hashlib.md5(open(yuv, "rb").read()).hexdigest() == gold["yuv_md5"]
    import time
    t0 = time.time()
    p = subprocess.run([BIN] + meg.args_1080p(CFG, yuv, 2, binf), capture_output=True, text=True,
                       env=dict(os.environ, HMB200_SHIM="gpu"), timeout=3000)
    assert p.returncode == 0, p.stderr[-2000:]
    print(f"\n1080p I+P through libhmb200: {time.time() - t0:.1f} s wall; " + "; ".join(l for l in p.stderr.splitlines() if l.startswith("hmb200 shim:")))
    assert [list(x) for x in meg.parse_md5_lines(p.stdout)] == [list(x) for x in gold["picture_md5"]]
    assert hashlib.md5(open(binf, "rb").read()).hexdigest() == gold["bitstream_md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("tag,mode", [("416x240", "verify"), ("1080p", "gpu")])
def test_encoder_bitstream_md5_tz_search(tmp_path, tag, mode):
    """The configuration file's own FastSearch = 1: every xPatternSearchFast (xTZSearch), bi-pred xPatternSearch and
    xPatternSearchFracDIF call goes through libhmb200; the bitstream must equal the stock encoder's."""
    _need_binary()
    from video_codecs_b200 import synth
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "encoder_md5_tz.json")))[tag]
    w, h = (416, 240) if tag == "416x240" else (1920, 1080)
    yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
    synth.write_yuv420(yuv, [synth.luma_frame(w, h, t, seed=77) for t in range(gold["frames"])], 8)
    assert hashlib.md5(open(yuv, "rb").read()).hexdigest() == gold["yuv_md5"]
    p = subprocess.run([BIN] + meg.args_tz(CFG, yuv, w, h, gold["frames"], binf) + gold["extra_args"], capture_output=True, text=True,
                       env=dict(os.environ, HMB200_SHIM=mode), timeout=3000)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "integer searches" in p.stderr and " 0 integer searches" not in p.stderr
    assert [list(x) for x in meg.parse_md5_lines(p.stdout)] == [list(x) for x in gold["picture_md5"]]
    assert hashlib.md5(open(binf, "rb").read()).hexdigest() == gold["bitstream_md5"]


CFG_RA10 = os.path.join(ROOT, "integration", "_build", "randomaccess_main10_settings.cfg")


@pytest.mark.gpu
def test_encoder_bitstream_md5_randomaccess_main10(tmp_path):
    """BASELINE.json configs[3]'s coding structure (encoder_randomaccess_main10.cfg: B slices, GOP 8, two lists, 10-bit,
    full search +-128) on a CPU-runnable picture: besides the uni-directional searches every bi-prediction refinement
    (xPatternSearch at +-BipredSearchRange on the signed `2*org - other prediction` pattern, TEncSearch.cpp:3686-3697, 3710-3728)
    and the 10-bit quarter-pel SATD refinement go through libhmb200; bitstream and picture MD5s must equal the stock
    encoder's (tests/golden/encoder_md5_ra10.json)."""
    _need_binary()
    if not os.path.exists(CFG_RA10):
        pytest.skip("integration/_build/randomaccess_main10_settings.cfg not written (python integration/build_shim.py)")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "encoder_md5_ra10.json")))
    yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
    meg.write_clip_ra10(yuv, gold["frames"])
    assert hashlib.md5(open(yuv, "rb").read()).hexdigest() == gold["yuv_md5"], "synthetic clip differs from the golden run's"
    p = subprocess.run([BIN] + meg.args_ra10(CFG_RA10, yuv, gold["frames"], binf), capture_output=True, text=True,
                       env=dict(os.environ, HMB200_SHIM="gpu", HMB200_SHIM_MERGE="1"), timeout=3000)     # bi-directional merge candidates on the GPU
    assert p.returncode == 0, p.stderr[-2000:]
    assert "integer searches" in p.stderr and " 0 integer searches" not in p.stderr, "the GPU path was not exercised"
    assert [list(x) for x in meg.parse_md5_lines(p.stdout)] == [list(x) for x in gold["picture_md5"]]
    assert hashlib.md5(open(binf, "rb").read()).hexdigest() == gold["bitstream_md5"]


@pytest.mark.gpu
def test_encoder_randomaccess_main10_verify_mode_prefix(tmp_path):
    """The same configuration in the shim's verify mode (every GPU result compared call by call with the reference
    body; the encoder aborts on the first mismatch) on the first two pictures, I + B: the host-side reference search at
    +-128 makes the full five-picture clip a three-minute test (it passed in that form, 183 s on the GPU box)."""
    _need_binary()
    if not os.path.exists(CFG_RA10):
        pytest.skip("integration/_build/randomaccess_main10_settings.cfg not written (python integration/build_shim.py)")
    yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
    meg.write_clip_ra10(yuv, 2)
    p = subprocess.run([BIN] + meg.args_ra10(CFG_RA10, yuv, 2, binf), capture_output=True, text=True,
                       env=dict(os.environ, HMB200_SHIM="verify", HMB200_SHIM_MERGE="1"), timeout=3000)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "integer searches" in p.stderr and " 0 integer searches" not in p.stderr, "the GPU path was not exercised"
    assert len(meg.parse_md5_lines(p.stdout)) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["verify", "gpu"])
def test_encoder_bitstream_md5_randomaccess_main10_tz(tmp_path, mode):
    """encoder_randomaccess_main10.cfg with its own FastSearch = 1 over one whole GOP (I + 8 B pictures, 10-bit): the TZ
    search on 16-bit planes, every bi-prediction refinement search and the 10-bit quarter-pel refinement run through
    libhmb200 (verify: each call also checked against the reference body); md5s from the stock encoder."""
    _need_binary()
    if not os.path.exists(CFG_RA10):
        pytest.skip("integration/_build/randomaccess_main10_settings.cfg not written (python integration/build_shim.py)")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "encoder_md5_ra10_tz.json")))
    yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
    meg.write_clip_ra10(yuv, gold["frames"])
    assert hashlib.md5(open(yuv, "rb").read()).hexdigest() == gold["yuv_md5"], "synthetic clip differs from the golden run's"
    p = subprocess.run([BIN] + meg.args_ra10(CFG_RA10, yuv, gold["frames"], binf, fast_search=1), capture_output=True, text=True,
                       env=dict(os.environ, HMB200_SHIM=mode), timeout=3000)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "integer searches" in p.stderr and " 0 integer searches" not in p.stderr, "the GPU path was not exercised"
    assert [list(x) for x in meg.parse_md5_lines(p.stdout)] == [list(x) for x in gold["picture_md5"]]
    assert hashlib.md5(open(binf, "rb").read()).hexdigest() == gold["bitstream_md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_kind", ["lowdelay8", "ra10"])
def test_distortion_table_hook_in_encoder(tmp_path, cfg_kind):
    """Row a2 of the scope table: forwarders installed in TComRdCost::m_afpDistortFunc at the end of TComRdCost::init()
    (TLibCommon/TComRdCost.cpp:224-276).  Every 499th call of each family - whatever the encoder asks the table for: RD-stage
    SSE of luma and chroma blocks, intra HADs, sub-pel SADs, bi-pred patterns - is also evaluated by hmb200_dist and must
    equal the reference member function (the shim aborts otherwise); the bitstream stays the stock encoder's."""
    _need_binary()
    if cfg_kind == "lowdelay8":
        gold = json.load(open(GOLD))["3"]
        frames, cfg = 3, CFG
        yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
        meg.write_clip(yuv, frames)
        args = meg.encoder_args(cfg, yuv, frames, binf)
    else:
        if not os.path.exists(CFG_RA10):
            pytest.skip("integration/_build/randomaccess_main10_settings.cfg not written")
        gold = None
        frames = 2
        yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
        meg.write_clip_ra10(yuv, frames)
        args = meg.args_ra10(CFG_RA10, yuv, frames, binf, fast_search=1)
    p = subprocess.run([BIN] + args, capture_output=True, text=True, env=dict(os.environ, HMB200_SHIM="gpu", HMB200_SHIM_TABLE="499"), timeout=3000)
    assert p.returncode == 0, p.stderr[-2000:]
    import re
    m = re.search(r"distortion-table hook: SAD (\d+)/(\d+) SADS (\d+)/(\d+) SSE (\d+)/(\d+) HADS (\d+)/(\d+)", p.stderr)
    assert m, p.stderr[-1500:]
    checked = [int(m.group(i)) for i in (1, 3, 5, 7)]
    assert checked[2] > 100 and checked[3] > 100 and sum(checked) > 1000, checked      # SSE and HADS are the table's big users
    if gold is not None:
        assert [list(x) for x in meg.parse_md5_lines(p.stdout)] == [list(x) for x in gold["picture_md5"]]
        assert hashlib.md5(open(binf, "rb").read()).hexdigest() == gold["bitstream_md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_kind", ["lowdelay8", "ra10_tz"])
def test_encoder_bitstream_md5_with_tile_columns(tmp_path, cfg_kind):
    """Two uniformly spaced tile columns (TileUniformSpacing / NumTileColumnsMinus1, App/TAppEncoder/TAppEncCfg.cpp:905-906 - the
    sharding unit of BASELINE.json configs[3]) in the routed encoder: predictors, merge candidates and therefore windows differ
    from the untiled runs along the column boundary; searches, refinements and merge estimation go through libhmb200 and the
    bitstream equals the stock encoder's with the same options (tests/golden/encoder_md5_tiles.json)."""
    _need_binary()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "encoder_md5_tiles.json")))[cfg_kind]
    yuv, binf = str(tmp_path / "clip.yuv"), str(tmp_path / "out.bin")
    if cfg_kind == "lowdelay8":
        meg.write_clip_tiles(yuv, gold["frames"])
        args = meg.encoder_args(CFG, yuv, gold["frames"], binf)
    else:
        if not os.path.exists(CFG_RA10):
            pytest.skip("integration/_build/randomaccess_main10_settings.cfg not written")
        meg.write_clip_tiles(yuv, gold["frames"], 10)
        args = meg.args_ra10(CFG_RA10, yuv, gold["frames"], binf, fast_search=1)
    assert hashlib.md5(open(yuv, "rb").read()).hexdigest() == gold["yuv_md5"]
    p = subprocess.run([BIN] + meg.tile_args(args), capture_output=True, text=True,
                       env=dict(os.environ, HMB200_SHIM="gpu", HMB200_SHIM_MERGE="1"), timeout=3000)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "integer searches" in p.stderr and " 0 integer searches" not in p.stderr
    assert [list(x) for x in meg.parse_md5_lines(p.stdout)] == [list(x) for x in gold["picture_md5"]]
    assert hashlib.md5(open(binf, "rb").read()).hexdigest() == gold["bitstream_md5"]
