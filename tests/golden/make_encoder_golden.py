"""Generates tests/golden/encoder_md5.json from the UNMODIFIED reference encoder (oracle/_ref/TAppEncoderStatic, built
from /root/reference/hm-16.5rc1 by oracle/Makefile.ref).  Run in the build container only:

    python tests/golden/make_encoder_golden.py

BASELINE.json configs[0]: encoder_lowdelay_P_main.cfg, synthetic 416x240 8-bit, FastSearch=0 SearchRange=64.  The
clip is video_codecs_b200.synth (seeded); the JSON records the bitstream md5 and the per-picture MD5 lines so that the
GPU-routed encoder (integration/) can be checked on machines without /root/reference."""
import hashlib
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from video_codecs_b200 import synth  # noqa: E402

W, H = 416, 240
CFG = "/root/reference/hm-16.5rc1/cfg/encoder_lowdelay_P_main.cfg"
ENC = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderStatic")


def encoder_args(cfg, yuv, frames, out_bin):
    return ["-c", cfg, "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(frames), "--FastSearch=0",
            "--SearchRange=64", "--SEIDecodedPictureHash=1", "-b", out_bin, "-o", ""]


def write_clip(path, frames):
    synth.write_yuv420(path, [synth.luma_frame(W, H, t, seed=77) for t in range(frames)], 8)


def parse_md5_lines(stdout):
    return re.findall(r"POC\s+(\d+).*?\[MD5:([0-9a-f,]+)\]", stdout)


def golden_1080p():
    """BASELINE.json configs[2] geometry, short prefix: 1920x1080 (coded 1088 rows, ConformanceWindowMode=1), 2 frames
    (one I, one P with a single reference): ~6 minutes of CPU for the stock encoder."""
    W2, H2, frames = 1920, 1080, 2
    yuv, binf = "/tmp/hmgold_1080.yuv", "/tmp/hmgold_1080.bin"
    synth.write_yuv420(yuv, [synth.luma_frame(W2, H2, t, seed=77) for t in range(frames)], 8)
    args = args_1080p(CFG, yuv, frames, binf)
    t0 = time.time()
    p = subprocess.run([ENC] + args, capture_output=True, text=True, check=True)
    out = {"bitstream_md5": hashlib.md5(open(binf, "rb").read()).hexdigest(), "bitstream_bytes": os.path.getsize(binf),
           "picture_md5": parse_md5_lines(p.stdout), "cpu_seconds": round(time.time() - t0, 1),
           "yuv_md5": hashlib.md5(open(yuv, "rb").read()).hexdigest()}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "encoder_md5_1080p.json"), "w"), indent=1)
    print(out)


def args_1080p(cfg, yuv, frames, out_bin):
    return ["-c", cfg, "-i", yuv, "-wdt", "1920", "-hgt", "1080", "-fr", "30", "-f", str(frames), "--FastSearch=0",
            "--SearchRange=64", "--SEIDecodedPictureHash=1", "--ConformanceWindowMode=1", "-b", out_bin, "-o", ""]


def golden_tz():
    """The cfg's own FastSearch = 1 (TZ search): 416x240 x 8 frames and 1920x1080 x 3 frames of the stock encoder."""
    out = {}
    for tag, (w, h, frames, extra) in {"416x240": (W, H, 8, []), "1080p": (1920, 1080, 3, ["--ConformanceWindowMode=1"])}.items():
        yuv, binf = f"/tmp/hmgold_tz_{tag}.yuv", f"/tmp/hmgold_tz_{tag}.bin"
        synth.write_yuv420(yuv, [synth.luma_frame(w, h, t, seed=77) for t in range(frames)], 8)
        args = args_tz(CFG, yuv, w, h, frames, binf) + extra
        t0 = time.time()
        p = subprocess.run([ENC] + args, capture_output=True, text=True, check=True)
        out[tag] = {"bitstream_md5": hashlib.md5(open(binf, "rb").read()).hexdigest(), "bitstream_bytes": os.path.getsize(binf),
                    "picture_md5": parse_md5_lines(p.stdout), "cpu_seconds": round(time.time() - t0, 1), "frames": frames,
                    "yuv_md5": hashlib.md5(open(yuv, "rb").read()).hexdigest(), "extra_args": extra}
        print(tag, out[tag]["bitstream_md5"], out[tag]["cpu_seconds"], flush=True)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "encoder_md5_tz.json"), "w"), indent=1)


def args_tz(cfg, yuv, w, h, frames, out_bin):
    return ["-c", cfg, "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "--FastSearch=1",
            "--SearchRange=64", "--SEIDecodedPictureHash=1", "-b", out_bin, "-o", ""]


CFG_RA10 = "/root/reference/hm-16.5rc1/cfg/encoder_randomaccess_main10.cfg"
RA10_FRAMES = 5


def args_ra10(cfg, yuv, frames, out_bin, fast_search=0):
    """BASELINE.json configs[3]'s coding structure on a CPU-runnable picture: encoder_randomaccess_main10.cfg (B slices,
    GOP 8, two lists, bi-prediction refinement at +-4 on a signed 16-bit pattern), 10-bit input and internal depth,
    full search +-128 (fast_search=1: the configuration's own TZ search instead)."""
    return ["-c", cfg, "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(frames), "--InputBitDepth=10",
            f"--FastSearch={fast_search}", "--SearchRange=128", "--SEIDecodedPictureHash=1", "-b", out_bin, "-o", ""]


def write_clip_ra10(path, frames):
    synth.write_yuv420(path, [synth.luma_frame(W, H, t, seed=77, bit_depth=10) for t in range(frames)], 10)


RA10_TZ_FRAMES = 9          # one whole GOP of 8 behind the I picture


def golden_ra10(fast_search=0):
    frames = RA10_TZ_FRAMES if fast_search else RA10_FRAMES
    yuv, binf = "/tmp/hmgold_ra10.yuv", "/tmp/hmgold_ra10.bin"
    write_clip_ra10(yuv, frames)
    t0 = time.time()
    p = subprocess.run([ENC] + args_ra10(CFG_RA10, yuv, frames, binf, fast_search), capture_output=True, text=True, check=True)
    out = {"bitstream_md5": hashlib.md5(open(binf, "rb").read()).hexdigest(), "bitstream_bytes": os.path.getsize(binf),
           "picture_md5": parse_md5_lines(p.stdout), "cpu_seconds": round(time.time() - t0, 1), "frames": frames,
           "yuv_md5": hashlib.md5(open(yuv, "rb").read()).hexdigest()}
    name = "encoder_md5_ra10_tz.json" if fast_search else "encoder_md5_ra10.json"
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", name), "w"), indent=1)
    print(out)


TILE_ARGS = ["--TileUniformSpacing=1", "--NumTileColumnsMinus1=1"]      # two uniformly spaced tile columns (TAppEncCfg.cpp:905-906)


TILE_W, TILE_H = 576, 240           # 9 CTUs wide: two uniform tile columns of 4 and 5 CTUs (a tile must be at least 256 luma samples wide)


def tile_args(args):
    """The untiled command line at 576x240 with two tile columns."""
    args = list(args)
    args[args.index("-wdt") + 1] = str(TILE_W)
    args[args.index("-hgt") + 1] = str(TILE_H)
    return args + TILE_ARGS


def write_clip_tiles(path, frames, bit_depth=8):
    synth.write_yuv420(path, [synth.luma_frame(TILE_W, TILE_H, t, seed=77, bit_depth=bit_depth) for t in range(frames)], bit_depth)


def golden_tiles():
    """Tiles enabled (BASELINE.json configs[3]'s sharding unit, in-encoder): two uniform tile columns.  Tiles restrict MV prediction
    and merge candidates across the column boundary, not the search window, so the routed calls see other predictors and windows
    than the untiled runs.  lowdelay-P 8-bit full search (3 frames) and random-access Main10 with the file's TZ search (one GOP)."""
    out = {}
    yuv, binf = "/tmp/hmgold_tiles.yuv", "/tmp/hmgold_tiles.bin"
    write_clip_tiles(yuv, 3)
    t0 = time.time()
    p = subprocess.run([ENC] + tile_args(encoder_args(CFG, yuv, 3, binf)), capture_output=True, text=True, check=True)
    out["lowdelay8"] = {"bitstream_md5": hashlib.md5(open(binf, "rb").read()).hexdigest(), "picture_md5": parse_md5_lines(p.stdout),
                        "cpu_seconds": round(time.time() - t0, 1), "frames": 3, "yuv_md5": hashlib.md5(open(yuv, "rb").read()).hexdigest()}
    print(out["lowdelay8"], flush=True)
    write_clip_tiles(yuv, RA10_TZ_FRAMES, 10)
    t0 = time.time()
    p = subprocess.run([ENC] + tile_args(args_ra10(CFG_RA10, yuv, RA10_TZ_FRAMES, binf, 1)), capture_output=True, text=True, check=True)
    out["ra10_tz"] = {"bitstream_md5": hashlib.md5(open(binf, "rb").read()).hexdigest(), "picture_md5": parse_md5_lines(p.stdout),
                      "cpu_seconds": round(time.time() - t0, 1), "frames": RA10_TZ_FRAMES, "yuv_md5": hashlib.md5(open(yuv, "rb").read()).hexdigest()}
    print(out["ra10_tz"], flush=True)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "encoder_md5_tiles.json"), "w"), indent=1)


def main():
    if "--tiles" in sys.argv:
        return golden_tiles()
    if "--1080p" in sys.argv:
        return golden_1080p()
    if "--ra10" in sys.argv:
        return golden_ra10()
    if "--ra10-tz" in sys.argv:
        return golden_ra10(fast_search=1)
    if "--tz" in sys.argv:
        return golden_tz()
    out = {}
    for frames in (3, 8):
        yuv, binf = f"/tmp/hmgold_{frames}.yuv", f"/tmp/hmgold_{frames}.bin"
        write_clip(yuv, frames)
        t0 = time.time()
        p = subprocess.run([ENC] + encoder_args(CFG, yuv, frames, binf), capture_output=True, text=True, check=True)
        dt = time.time() - t0
        out[str(frames)] = {"bitstream_md5": hashlib.md5(open(binf, "rb").read()).hexdigest(),
                            "bitstream_bytes": os.path.getsize(binf),
                            "picture_md5": parse_md5_lines(p.stdout), "cpu_seconds": round(dt, 1),
                            "yuv_md5": hashlib.md5(open(yuv, "rb").read()).hexdigest()}
        print(frames, out[str(frames)]["bitstream_md5"], dt, flush=True)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "encoder_md5.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
