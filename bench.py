#!/usr/bin/env python
"""bench.py — throughput of the HM-16.5 integer-pel full search + quarter-pel SATD refinement path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[2]): 1920x1080 8-bit (coded as 1920x1088, ConformanceWindowMode=1), lowdelay_P
settings (FEN=1, HadamardME=1), full search +-64 over the canonical all-PU job list (593 PUs per CTU)
followed by the quarter-pel SATD refinement of every PU.  One step = one (current, reference) frame pair per rank;
ranks work on independent frame pairs (no collective on the data path, scaling = weak).

Our arm prints `value` (planes resident in HBM, device time from CUDA events on the library's stream) and `e2e`
(host planes in, host MV field out, through the C-ABI).  The same JSON line carries, under `extra_workloads`, the other
GPU configurations of BASELINE.json measured in the same run (each with its own ms/step, e2e, roofline and clocks):

  tiles_2160p10    configs[3]: 3840x2160 10-bit +-128, the uniformly spaced tile columns of ONE picture, one per rank
                   (strong scaling); every rank uploads only its column's halo crop; merged shards == unsharded field
  stream_2160p8    configs[4]: 3840x2160 8-bit lookahead, distinct frames streamed back to back (upload -> search ->
                   refinement -> MV field to the host, pipelined, no L2 flush, no host sync between pairs): the sustained number
  unfused_1080p    the headline list with a different AMVP predictor per PU (nothing can share a pass: per-PU kernels)
  dist_table       configs[1]: SAD / SSE / HADs over >= 10^4 random blocks of every PU size, 8- and 10-bit

`--impl reference` times the unmodified reference (oracle/_ref/libhmref.so, HM-16.5 compiled from /root/reference in the
build container) on all host cores on a bounded sample of the same job list; it never loads the product library.  Only
that arm and the `cpu_baseline` leg execute anything under oracle/.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# --workload selects one of BASELINE.json's GPU configurations; the default (and the one the driver runs) is configs[2]
WORKLOADS = {
    "1080p": dict(w=1920, h=1080, coded_h=1088, sr=64, bd=8,
                  name="1080p_8bit_lowdelayP_fullsearch64_canonical593_fen1_hadme1_qpel",
                  metric="Mpixel/s, full-search +-64 ME with quarter-pel SATD refinement, 1080p"),
    "2160p8": dict(w=3840, h=2160, coded_h=2160, sr=64, bd=8,
                   name="2160p_8bit_lookahead_fullsearch64_canonical593_fen1_hadme1_qpel",
                   metric="Mpixel/s, full-search +-64 ME with quarter-pel SATD refinement, 2160p"),
    "2160p10": dict(w=3840, h=2160, coded_h=2160, sr=128, bd=10,
                    name="2160p_10bit_main10_fullsearch128_canonical593_fen1_hadme1_qpel",
                    metric="Mpixel/s, full-search +-128 ME with quarter-pel SATD refinement, 2160p 10-bit"),
}
LAMBDA_COST = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))   # lowdelay-P slice at QP 35
N_FRAMES = 5                      # distinct synthetic frames per rank -> 4 frame pairs, cycled (device-resident legs)
INT_PEAK_FILE = os.path.join(ROOT, "profiles", "r01_microbench_int.json")
INT16_PEAK_FILE = os.path.join(ROOT, "profiles", "r01_microbench_int16.json")
REFINE_OPS_FILE = os.path.join(ROOT, "profiles", "r02_refine_inst_counts.json")

_REAL_STDOUT = None


def emit(line):
    """The one JSON line on the real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self.stop_flag = True
        self.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def int_simd_peak(bit_depth=8):
    """Measured integer-SIMD SAD peak of this pool's B200, abs-diffs / s: the VABSDIFF4.U8.ACC issue rate for 8-bit
    content (tools/microbench_int.cu), and for deeper content the register-only rate of the instruction pair
    k_search16_cu uses (VIMNMX.U16x2 + IDP.2A.LO per two samples, the sum-of-minima form; tools/microbench_int16.cu) --
    SURVEY.md 8d asks for a separately calibrated peak there."""
    if bit_depth > 8:
        try:
            d = json.load(open(INT16_PEAK_FILE))
            seqs = max(v for k, v in d.items() if k.startswith("vimnmx16x2_plus_idp2a"))
            return seqs * 1e9 * 2.0, ("measured: profiles/r01_microbench_int16.json (register-only VIMNMX.U16x2 + IDP.2A.LO loop, "
                                      "two abs-diffs per pair)")
        except Exception:
            return 148 * 64 * 2.0 * 1.965e9, "fallback: one ALU-pipe + one FMA-pipe instruction per 2 samples at 64 lanes/clk/SM and pipe x 148 SMs x 1.965 GHz"
    try:
        d = json.load(open(INT_PEAK_FILE))
        lane_ops = max(v for k, v in d.items() if k.startswith("alu_vabsdiff4_acc"))
        return lane_ops * 1e9 * 4.0, "measured: profiles/r01_microbench_int.json (register-only VABSDIFF4.U8.ACC loop)"
    except Exception:
        return 148 * 64 * 4 * 1.965e9, "fallback: 64 lanes/clk/SM x 4 bytes x 148 SMs x 1.965 GHz"


def issue_peak():
    """Instruction issue peak: one warp instruction per clock and SM sub-partition = 148 x 4 x 32 lanes x 1.965 GHz."""
    return 148 * 4 * 32 * 1.965e9, "issue limit: 1 warp instruction / clk / SM sub-partition x 592 sub-partitions x 1.965 GHz (measured IMAD-only rate: 17.7 T/s, IADD-only: 36.0 T/s, profiles/r01_microbench_int.json)"


def ncu_traffic(wl_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant (longest) search kernel, from the committed
    ncu --set full capture (profiles/r02_ncu_full_1080p.csv, else round 1's); None when there is no capture for the workload."""
    import csv
    import math
    cands = {"1080p": [("r02_ncu_full_1080p.csv", "k_search8_cu"), ("r01_ncu_full_final.csv", "k_search8_cu")],
             "2160p10": [("r02_ncu_full_2160p10.csv", "k_search16_cu"), ("r01_ncu_full_2160p10.csv", "k_search16_cu")]}.get(wl_key, [])
    for fname, kern in cands:
        try:
            rows = list(csv.reader(open(os.path.join(ROOT, "profiles", fname))))
            hdr, units = rows[0], rows[1]
            ki, ti = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
            ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            ok = [r for r in rows[2:] if kern in r[ki] and not math.isnan(float(r[ri])) and not math.isnan(float(r[wi]))]
            best = max(ok, key=lambda r: float(r[ti]))
            return (int(float(best[ri]) * scale.get(units[ri], 1.0) + float(best[wi]) * scale.get(units[wi], 1.0)),
                    best[ki].split("(")[0] + " [" + fname + "]")
        except Exception:
            continue
    return None, None


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def count_pus(pic_w, coded_h):
    """PUs of the canonical list of a picture (13 per CU of 16..64, 5 per 8x8 CU, CUs inside the picture only)."""
    n = 0
    for s in (64, 32, 16, 8):
        n += (pic_w // s) * (coded_h // s) * (13 if s > 8 else 5)
    return n


def config_of(wl, search="full"):
    """The `config` both arms print (identical keys and values for one workload)."""
    name = wl["name"] if search == "full" else wl["name"].replace("fullsearch", "tzsearch")
    return {"workload": name, "picture": f"{wl['w']}x{wl['h']} (coded {wl['w']}x{wl['coded_h']})", "search_range": wl["sr"],
            "bit_depth": wl["bd"], "pus_per_frame": count_pus(wl["w"], wl["coded_h"]), "pus_per_ctu": 593,
            "frame_pairs_per_step_per_gpu": 1, "fen": 1, "hadamard_me": 1, "refinement": "quarter-pel SATD, every PU",
            "mpixel_counts": f"{wl['w']}x{wl['h']} luma per frame pair"}


def px_per_ctu(wl):
    n_ctus = ((wl["w"] + 63) // 64) * ((wl["coded_h"] + 63) // 64)
    return wl["w"] * wl["h"] / n_ctus, n_ctus


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the only users of oracle/; neither loads the product library)
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_checker():
    from oracle.pyoracle import Oracle, Reference
    try:
        return Reference(fen=1, hadme=1), "reference"
    except (FileNotFoundError, OSError):
        return Oracle(fen=1, hadme=1), "port"


def _synth():
    """video_codecs_b200/synth.py (pure numpy input generator) loaded by path: importing the package would load libhmb200."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_hmb_synth", os.path.join(ROOT, "video_codecs_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_WORKER = {}


def _cpu_worker(args):
    """One host process: runs the CPU implementation over whole CTUs of one frame pair (inputs cached per process)."""
    wl, frames_seed, ctus = args
    if frames_seed not in _WORKER:
        from oracle.pyoracle import py_canonical_jobs
        synth = _synth()
        chk, _ = _cpu_checker()
        _WORKER[frames_seed] = (chk, py_canonical_jobs,
                                synth.pad_plane(synth.luma_frame(wl["w"], wl["coded_h"], 1, seed=frames_seed), 80, 80),
                                synth.pad_plane(synth.luma_frame(wl["w"], wl["coded_h"], 0, seed=frames_seed), 80, 80))
    chk, build, cur, ref = _WORKER[frames_seed]
    stride = cur.shape[1]
    o0 = 80 * stride + 80
    t0 = time.perf_counter()
    n = 0
    for c in ctus:
        jobs = build(wl["w"], wl["coded_h"], wl["sr"], LAMBDA_COST, ctu_first=c, ctu_count=1)
        chk.run_jobs((cur, o0, stride), (ref, o0, stride), jobs, 8, True)
        n += 1
    return time.perf_counter() - t0, n


def cpu_sample_single(wl, n_ctus_sample):
    """Single-thread CPU baseline on the first CTUs of the interior CTU row 8 (bounded sample)."""
    from oracle.pyoracle import py_canonical_jobs
    chk, kind = _cpu_checker()
    synth = _synth()
    cur = synth.pad_plane(synth.luma_frame(wl["w"], wl["coded_h"], 1, seed=1234), 80, 80)
    ref = synth.pad_plane(synth.luma_frame(wl["w"], wl["coded_h"], 0, seed=1234), 80, 80)
    stride = cur.shape[1]
    o0 = 80 * stride + 80
    first = 8 * 30 + 3
    jobs = py_canonical_jobs(wl["w"], wl["coded_h"], wl["sr"], LAMBDA_COST, ctu_first=first, ctu_count=n_ctus_sample)
    t0 = time.perf_counter()
    chk.run_jobs((cur, o0, stride), (ref, o0, stride), jobs, 8, True)
    dt = time.perf_counter() - t0
    ppc, _ = px_per_ctu(wl)
    return {"value": n_ctus_sample * ppc / dt / 1e6, "unit": "Mpixel/s", "cores": 1, "kind": kind,
            "sample": f"{n_ctus_sample} interior CTUs ({len(jobs)} PU searches + refinements) of one 1080p frame pair, "
                      f"{dt:.1f} s on one host core"}


def stock_encoder_leg():
    """BASELINE.json configs[0] beside the sampled job list: wall time of the stock encoder (oracle/_ref/TAppEncoderStatic,
    the unmodified reference) on the 416x240 synthetic clip, lowdelay-P, FastSearch=0 SearchRange=64, I + P."""
    import subprocess
    import tempfile
    enc = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderStatic")
    cfg = os.path.join(ROOT, "integration", "_build", "lowdelay_P_settings.cfg")
    if not (os.path.exists(enc) and os.path.exists(cfg)):
        return None
    synth = _synth()
    frames = 2
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        synth.write_yuv420(yuv, [synth.luma_frame(416, 240, t, seed=77) for t in range(frames)], 8)
        cmd = [enc, "-c", cfg, "-i", yuv, "-wdt", "416", "-hgt", "240", "-fr", "30", "-f", str(frames), "--FastSearch=0", "--SearchRange=64",
               "-b", os.path.join(d, "o.bin"), "-o", ""]
        t0 = time.perf_counter()
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
        except subprocess.TimeoutExpired:
            return {"error": "timeout"}
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            return {"error": p.stderr[-200:]}
    return {"what": "TAppEncoderStatic (stock HM-16.5), encoder_lowdelay_P_main settings, 416x240 8-bit synthetic, FastSearch=0 SearchRange=64",
            "frames": frames, "wall_s": dt, "mpixel_per_s": (frames - 1) * 416 * 240 / dt / 1e6, "cores": 1,
            "note": "whole encoder (mode decision, transform, entropy coding included); one P picture with one reference"}


def run_reference_arm(args, wl):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ppc, n_ctus = px_per_ctu(wl)
    ctx = mp.get_context("spawn")
    _, kind = _cpu_checker()
    interior = [r * 30 + c for r in range(2, 15) for c in range(2, 28)]
    with ctx.Pool(cores) as pool:
        # calibration: one CTU per core
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(wl, 1234, [interior[i % len(interior)]]) for i in range(cores)])
        t_ctu = time.perf_counter() - t0
        budget = 120.0 / max(1, args.steps + args.warmup)
        per_core = int(max(1, min(8, budget // max(t_ctu, 1e-3))))
        per_core = max(1, min(per_core, (len(interior) - 1) // cores))

        def step(k):
            base = (k * cores * per_core) % max(1, len(interior) - cores * per_core)
            work = [(wl, 1234, interior[base + i * per_core: base + (i + 1) * per_core]) for i in range(cores)]
            t = time.perf_counter()
            pool.map(_cpu_worker, work)
            return time.perf_counter() - t
        for k in range(args.warmup):
            step(k)
        t_total = 0.0
        for k in range(args.steps):
            t_total += step(args.warmup + k)
    ctus_done = args.steps * cores * per_core
    value = ctus_done * ppc / t_total / 1e6
    sample = (f"{cores * per_core} CTUs per step ({per_core} per core, whole canonical job list of each CTU incl. "
              f"quarter-pel refinement), {args.steps} steps, one process per host core")
    line = {"impl": "reference", "metric": wl["metric"], "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8" if wl["bd"] == 8 else "u16", "data": "synthetic",
            "config": config_of(wl),
            "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.gpus == 1 and not args.no_encoder_leg:
        enc = stock_encoder_leg()
        if enc is not None:
            line["stock_encoder_configs0"] = enc
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class Env:
    """Process-wide handles of the GPU arm."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank, self.world, self.local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_mod
            self.dist = dist_mod
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"           # keeps NCCL's version banner off stdout: one JSON line only
            torch.cuda.set_device(self.local)
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        from video_codecs_b200 import HMB200
        self.hm = HMB200()
        self.hm.init(self.local)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{self.local}")   # > 126 MB L2
        self.args = args

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        self.hm.sync()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=f"cuda:{self.local}")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def gather_objects(self, obj):
        if self.dist is None:
            return [obj]
        box = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(obj, box, dst=0)
        return box


def pinned_copy(hm, a):
    out = hm.host_array(a.size, a.dtype).reshape(a.shape)
    out[...] = a
    return out


def measure_pairs(env, wl_key, steps, warmup, shard="frames", search="full", preds="zero", check_merge=False, with_hbm=True):
    """Device-resident throughput (`value`) and end-to-end throughput (`e2e`) of frame-pair steps of one workload."""
    from video_codecs_b200 import FLAG_FEN, FLAG_HADME, FLAG_FRAC, FLAG_TZ, FLAG_TZ_STOP, RESULT16_DTYPE, synth, widen_results16
    from video_codecs_b200 import shard as shard_mod
    hm, wl = env.hm, WORKLOADS[wl_key]
    W, H, CH, SR, BD = wl["w"], wl["h"], wl["coded_h"], wl["sr"], wl["bd"]
    rank, world = env.rank, env.world
    tiles = shard == "tiles"
    flags = FLAG_FEN | FLAG_HADME | FLAG_FRAC | ((FLAG_TZ | FLAG_TZ_STOP) if search == "tz" else 0)
    margin = 80 if SR <= 64 else 144
    n_frames = 3 if tiles else N_FRAMES
    frames = [synth.luma_frame(W, CH, t, seed=1234 + (0 if tiles else 97 * rank), bit_depth=BD) for t in range(n_frames)]
    pairs = [(t + 1, t) for t in range(n_frames - 1)]                       # (current, reference) = (t+1, t)
    crop = (0, W)
    if tiles:
        # tile-column sharding of ONE picture (BASELINE.json configs[3]): a rank searches the jobs of its uniformly spaced tile
        # column and holds only that column's halo crop of both planes; strong scaling, no collective on the data path
        x0, x1 = hm.tile_column_range(W, world, rank)
        jobs_pic = shard_mod.tile_column_jobs(hm, W, CH, world, rank, SR, LAMBDA_COST)
        crop = shard_mod.tile_column_crop(W, x0, x1, SR) if world > 1 else (0, W)
        jobs = shard_mod.shift_jobs(jobs_pic, crop[0])
        frames = [np.ascontiguousarray(f[:, crop[0]:crop[1]]) for f in frames]
    else:
        jobs = hm.build_canonical_jobs(W, CH, SR, LAMBDA_COST)
        jobs_pic = jobs
        if preds == "random":
            # a different AMVP-like predictor per PU (quarter-pel, +-4 pel): no two PUs of a CU share window and predictor, so
            # nothing is CU-fused - the situation inside an encoder whose neighbours' MVs differ
            rng = np.random.default_rng(99 + rank)
            jobs = jobs.copy()
            px, py = rng.integers(-16, 17, len(jobs)), rng.integers(-16, 17, len(jobs))
            jobs["pred_x"], jobs["pred_y"] = px, py
            jobs["lt_x"] += px >> 2; jobs["rb_x"] += px >> 2; jobs["lt_y"] += py >> 2; jobs["rb_y"] += py >> 2
    frames = [pinned_copy(hm, f) for f in frames]                            # e2e inputs live in page-locked host memory
    t_prep = time.perf_counter()
    prep = hm.prepare_jobs(jobs, flags, BD)
    if search == "tz":
        prep.set_tz(hm.canonical_tz_extra(jobs), (frames[0].shape[1], CH), SR)
    t_prep = time.perf_counter() - t_prep
    work = prep.work()

    def register(f, kind, poc=0):
        if BD == 8:
            return hm.register_plane_u8(f, margin, margin, kind=kind, poc=poc)
        return hm.register_plane_u16(f, BD, margin, margin, kind=kind, poc=poc)

    plane_ids = [register(f, 0, i) for i, f in enumerate(frames)]
    torch = env.torch

    def device_step(k):
        c, r = pairs[k % len(pairs)]
        env.flush.zero_()                               # L2 flush between timed iterations (not timed)
        torch.cuda.synchronize()
        prep.run(plane_ids[c], plane_ids[r])
        hm.sync()
        return prep.timing()

    for k in range(warmup):
        device_step(k)
    env.barrier()
    sampler = ClockSampler(env.local)
    sampler.start()
    launches0 = hm.launch_count()
    t_wall0 = time.perf_counter()
    tot = srch = frac = 0.0
    for k in range(steps):
        t = device_step(warmup + k)
        tot += t["total_ms"]; srch += t["search_ms"]; frac += t["frac_ms"]
    env.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    launches = hm.launch_count() - launches0
    clocks = sampler.finish()

    # ---- end to end through the C-ABI: host planes in, host MV field out, every step ------------------------------
    # Two prepared handles alternate: the D2H of step k (copy stream) runs behind the upload and kernels of step k+1.
    # Every step uploads both of its planes from page-locked host memory and delivers its MV field (16-byte records) to host
    # memory inside the timed region; a step's results are complete when fetch_wait returns for its handle.
    prep2 = hm.prepare_jobs(jobs, flags, BD)
    if search == "tz":
        prep2.set_tz(hm.canonical_tz_extra(jobs), (frames[0].shape[1], CH), SR)
    preps = [prep, prep2]
    outs = [hm.host_array(len(jobs), RESULT16_DTYPE), hm.host_array(len(jobs), RESULT16_DTYPE)]   # page-locked (hmb200_host_alloc)

    def e2e_step(k):
        c, r = pairs[k % len(pairs)]
        p, o = preps[k % 2], outs[k % 2]
        p.fetch_wait()                                              # step k-2's MV field has arrived (its consumer would run here)
        idc = register(frames[c], 0)                                # H2D from pinned memory + border extension on device
        idr = register(frames[r], 1)
        p.run(idc, idr)
        p.fetch16_async(o)                                          # D2H of the MV field / costs
        hm.release_plane(idc)
        hm.release_plane(idr)

    def e2e_drain():
        for p in preps:
            p.fetch_wait()

    e2e_steps = max(4, min(steps, 30))
    for k in range(max(4, warmup)):                                  # the plane pool reaches its steady size (HMB200_POOL_KEEP + 2 buffers)
        e2e_step(k)
    e2e_drain()
    env.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    e2e_drain()
    env.barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    # the last two steps' results must be the device-resident runs' results for the same pairs
    last = {}
    for k in (e2e_steps - 2, e2e_steps - 1):
        c, r = pairs[k % len(pairs)]
        prep.run(plane_ids[c], plane_ids[r])
        chk = prep.fetch()
        if not np.array_equal(chk, widen_results16(outs[k % 2])):
            raise SystemExit("bench: pipelined end-to-end results differ from the device-resident run")
        last[(c, r)] = chk

    merged_ok = None
    if tiles and check_merge and world > 1:
        # merged shards == the MV field of the whole picture searched at once by rank 0 (outside every timed region)
        c, r = pairs[0]
        prep.run(plane_ids[c], plane_ids[r])
        mine = prep.fetch()
        box = env.gather_objects((jobs_pic, mine))
        if rank == 0:
            whole_frames = [synth.luma_frame(W, CH, t, seed=1234, bit_depth=BD) for t in (r, c)]
            full_jobs = hm.build_canonical_jobs(W, CH, SR, LAMBDA_COST)
            ids = [register(np.ascontiguousarray(f), i) for i, f in enumerate(whole_frames)]
            pf = hm.prepare_jobs(full_jobs, flags, BD)
            pf.run(ids[1], ids[0])
            whole = pf.fetch()
            pf.free()
            for i in ids:
                hm.release_plane(i)
            merged = shard_mod.merge_shards([b[0] for b in box], [b[1] for b in box], full_jobs)
            merged_ok = bool(np.array_equal(merged, whole))
            if not merged_ok:
                raise SystemExit("bench: merged tile-column shards differ from the unsharded MV field")

    tot, srch, frac, e2e_ms, wall_ms = env.max_over_ranks([tot, srch, frac, e2e_ms, wall_ms])
    # whole-job work: every rank's share (ranks of a tile split hold different numbers of CTUs)
    tw = env.torch.tensor([work["cand_sads"], work["abs_diffs"], work["abs_diffs_executed"], work["abs_diffs_unique"], len(jobs),
                           work["pus_fused"]], dtype=env.torch.float64, device=f"cuda:{env.local}")
    if env.dist is not None:
        env.dist.all_reduce(tw, op=env.dist.ReduceOp.SUM)
    job_cands, job_abs, job_exec, job_unique, job_pus, job_fused = [float(v) for v in tw.cpu()]

    res = None
    if rank == 0:
        K = max(1, steps)
        ms_per_step = tot / K
        mpix_step = W * H / 1e6
        units = 1 if tiles else world                       # tile columns: all ranks together process ONE picture per step
        value = units * mpix_step / (ms_per_step / 1e3)
        e2e_value = units * mpix_step / (e2e_ms / 1e3)
        peak_abs, peak_src = int_simd_peak(BD)
        search_s = srch / K / 1e3
        # per-GPU rates against the per-GPU peak: this rank's work over the slowest rank's search time
        unique, executed, algorithmic = work["abs_diffs_unique"] / search_s, work["abs_diffs_executed"] / search_s, work["abs_diffs"] / search_s
        # round 1's count of the same quantity, for comparison across rounds: every CU on its own (an 8x8 CU's 64 abs-diffs per
        # candidate counted although its 16x16 parent's pass now yields them), 256 / 768 / 2048 per candidate of a 16 / 32 / 64 CU (FEN on)
        sq = jobs[(jobs["w"] == jobs["h"]) & np.isin(jobs["w"], (8, 16, 32, 64))]
        per_cand = np.select([sq["w"] == 8, sq["w"] == 16, sq["w"] == 32], [64, 256, 768], 2048).astype(np.int64)
        cands = (sq["rb_x"].astype(np.int64) - sq["lt_x"] + 1) * (sq["rb_y"].astype(np.int64) - sq["lt_y"] + 1)
        unique_r1 = float(np.sum(per_cand * cands))
        hbm, hbm_src = hbm_peak()
        traffic, traffic_kernel = ncu_traffic(wl_key)
        pu_pixels = float(np.sum(jobs["w"].astype(np.int64) * jobs["h"].astype(np.int64)))
        plane_bytes = (frames[0].shape[1] + 2 * margin) * (CH + 2 * margin) * (1 if BD == 8 else 2)
        algo_bytes = 2 * plane_bytes + len(jobs) * (32 + 48)       # both planes once + job list + results
        sample_bytes = 1 if BD == 8 else 2
        res = {
            "metric": wl["metric"], "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None,
            "dtype": "u8" if BD == 8 else "u16", "data": "synthetic",
            "config": config_of(wl, search),
            "timing": {"l2": "flushed between timed iterations (256 MiB write, untimed)", "device_time": "CUDA events on the library's stream, max over ranks",
                       "prepare_jobs_s": t_prep},
            "cand_sad_per_s": (job_cands if tiles else world * work["cand_sads"]) / (ms_per_step / 1e3),
            "sharding": ("tile columns of one picture, halo crops uploaded" if tiles else "independent frame pairs") if world > 1 else "one GPU",
            "search_ms": srch / K, "frac_ms": frac / K, "wall_ms_per_step_incl_flush": wall_ms / K,
            "roofline": {"bound": "int_alu",
                         "kernel": ("k_search8_cu<S,FEN,CHILD> (VABSDIFF4.U8.ACC)" if BD == 8 else "k_search16_cu<S,FEN> (sum of minima: VIMNMX.U16x2 + IDP.2A)"),
                         "achieved": unique / 1e12, "peak": peak_abs / 1e12, "unit": "Tabsdiff/s", "frac": unique / peak_abs,
                         "traffic": traffic, "traffic_kernel": traffic_kernel, "peak_source": peak_src,
                         "unique_absdiffs_per_launch": int(work["abs_diffs_unique"]),
                         "per_cu_count": {"absdiffs_per_launch": int(unique_r1), "frac": unique_r1 / search_s / peak_abs,
                                          "note": "round 1's count (VERDICT r01 weak #1: 64/256/768/2048 abs-diffs per candidate of an "
                                                  "8/16/32/64 CU, 8x8 CUs counted beside their parents); frac is the stricter count"},
                         "issued": executed / 1e12, "issued_frac": executed / peak_abs,
                         "issued_absdiffs_per_launch": int(work["abs_diffs_executed"]),
                         "algorithmic": algorithmic / 1e12, "algorithmic_over_peak": algorithmic / peak_abs,
                         "algorithmic_absdiffs_per_launch": int(work["abs_diffs"]),
                         "pus_cu_fused": int(work["pus_fused"]), "pus": int(len(jobs)),
                         "note": "frac = UNIQUE byte abs-diffs / search time / measured SAD-instruction peak (<= 1): every sample a CU-fused "
                                 "pass must visit, once per candidate of the window (all rows of a 16x16 CU that carries its 8x8 children, "
                                 "24 of 32 rows of a 32x32 CU and the even rows of a 64x64 CU under FEN), no masked lanes counted; issued = what "
                                 "the kernels execute incl. masked block columns (ncu opcode counts: profiles/r02_*); algorithmic = W*H[/2] per "
                                 "candidate and PU as HM executes them (may exceed the peak: 13-33 PUs share one pass).  Search time incl. "
                                 "key memset + finalize, on this rank's GPU"},
            "roofline_refine": refine_roofline(wl, BD, pu_pixels, frac / K / 1e3),
            "e2e": {"value": e2e_value, "unit": "Mpixel/s",
                    "h2d_bytes_per_step": int(2 * frames[0].size * sample_bytes),
                    "d2h_bytes_per_step": int(outs[0].nbytes), "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "note": "per rank; page-locked frames in, 16-byte MV-field records out, two frame pairs in flight"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if with_hbm:
            res["roofline_hbm"] = {"bound": "hbm", "achieved": algo_bytes / (ms_per_step / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                                   "frac": algo_bytes / (ms_per_step / 1e3) / 1e9 / hbm, "traffic": None, "peak_source": hbm_src}
        if tiles:
            res["tile_columns"] = {"columns": world, "crop_of_rank0": list(crop), "picture_width": W,
                                   "whole_job": {"pus": int(job_pus), "cand_sads": int(job_cands), "unique_absdiffs": int(job_unique)},
                                   "merged_equals_unsharded": merged_ok}
    prep.free()
    prep2.free()
    for i in plane_ids:
        hm.release_plane(i)
    hm.sync()
    for a in outs + frames:
        hm.host_free(a)
    return res


def refine_roofline(wl, bd, pu_pixels, frac_s):
    """Quarter-pel refinement against the measured IMAD issue rate.  Executed instructions per PU pixel come from ncu
    (smsp__inst_executed of the refinement kernels of one 1080p frame pair, profiles/r02_refine_inst_counts.json) when that
    file exists; otherwise from the source-level count of DESIGN.md 3.2."""
    peak, src = issue_peak()
    try:
        d = json.load(open(REFINE_OPS_FILE))
        per_px = d["lane_instructions_per_pu_pixel_8bit" if bd == 8 else "lane_instructions_per_pu_pixel_16bit"]
        how = f"ncu-counted executed lane-instructions per PU pixel ({REFINE_OPS_FILE.split(os.sep)[-1]})"
    except Exception:
        per_px, how = 250.0, "source-level count (DESIGN.md 3.2): ~250 integer lane-instructions per PU pixel"
    ops = per_px * pu_pixels
    return {"bound": "int_issue", "kernel": ("k_frac_hv<8,HAD> + k_frac_patch<4,HAD>" if bd == 8 else "k_frac_hv16<8,HAD> / k_frac_tiles<i16>"),
            "achieved": ops / frac_s / 1e12, "peak": peak / 1e12, "unit": "Tinst/s", "frac": ops / frac_s / peak, "traffic": None,
            "peak_source": src, "instructions_per_pu_pixel": per_px, "note": how + "; all pipes, so this is an issue-slot utilisation"}


def measure_stream(env, wl_key, n_frames, loops):
    """Lookahead streaming (BASELINE.json configs[4]): distinct frames, pairs (t, t-1), uploaded / searched / refined /
    fetched back to back with two pairs in flight; no L2 flush, no host synchronisation between pairs.  Wall clock over the
    whole region (max over ranks) = the sustained end-to-end rate; device busy time from the per-pair CUDA events."""
    from video_codecs_b200 import FLAG_FEN, FLAG_HADME, FLAG_FRAC, RESULT16_DTYPE, synth
    hm, wl = env.hm, WORKLOADS[wl_key]
    W, H, CH, SR, BD = wl["w"], wl["h"], wl["coded_h"], wl["sr"], wl["bd"]
    margin = 80
    flags = FLAG_FEN | FLAG_HADME | FLAG_FRAC
    # distinct frames at little host cost: four synthesised frames, each later frame = one of them moved by a growing integer
    # offset plus its own +-1 noise (so every pair has coherent, non-zero motion)
    base = [synth.luma_frame(W, CH, t, seed=4321 + 131 * env.rank, bit_depth=BD) for t in range(4)]
    rng = np.random.default_rng(17 + env.rank)
    frames = []
    for t in range(n_frames):
        f = np.roll(base[t % 4], ((t // 4) * 2, (t // 4) * 3), axis=(0, 1)).astype(np.int16)
        f += rng.integers(-1, 2, size=f.shape, dtype=np.int16)
        frames.append(pinned_copy(hm, np.clip(f, 0, 255).astype(np.uint8)))
    jobs = hm.build_canonical_jobs(W, CH, SR, LAMBDA_COST)
    preps = [hm.prepare_jobs(jobs, flags, BD) for _ in range(2)]
    outs = [hm.host_array(len(jobs), RESULT16_DTYPE) for _ in range(2)]
    n_pairs = (n_frames - 1) * loops
    checksum = [0]

    def step(k, timed):
        t = 1 + k % (n_frames - 1)
        p, o = preps[k % 2], outs[k % 2]
        p.fetch_wait()
        if timed and k >= 2:
            checksum[0] += int(o["sad"][::4096].sum())              # the consumer touches the MV field that just arrived
        idc = hm.register_plane_u8(frames[t], margin, margin, kind=0, poc=t)
        idr = hm.register_plane_u8(frames[t - 1], margin, margin, kind=1, poc=t - 1)
        p.run(idc, idr)
        p.fetch16_async(o)
        hm.release_plane(idc)
        hm.release_plane(idr)

    for k in range(4):
        step(k, False)
    for p in preps:
        p.fetch_wait()
    env.barrier()
    sampler = ClockSampler(env.local, period=0.05)
    sampler.start()
    launches0 = hm.launch_count()
    t0 = time.perf_counter()
    for k in range(n_pairs):
        step(k, True)
    for p in preps:
        p.fetch_wait()
    env.barrier()
    wall_s = time.perf_counter() - t0
    launches = hm.launch_count() - launches0
    clocks = sampler.finish()
    dev = preps[0].timing()
    wall_s, dev_ms = env.max_over_ranks([wall_s, dev["total_ms"]])
    res = None
    if env.rank == 0:
        mpix = W * H / 1e6
        res = {"workload": wl["name"], "config": config_of(wl), "n_gpus": env.world, "scaling": "weak",
               "distinct_frames_per_gpu": n_frames, "frame_pairs_per_gpu": n_pairs, "region_s": wall_s,
               "value": env.world * n_pairs * mpix / wall_s, "unit": "Mpixel/s", "ms_per_pair": 1e3 * wall_s / n_pairs,
               "device_ms_last_pair": dev_ms,
               "e2e": {"value": env.world * n_pairs * mpix / wall_s, "unit": "Mpixel/s", "h2d_bytes_per_step": int(2 * W * CH),
                       "d2h_bytes_per_step": int(outs[0].nbytes)},
               "l2": "not flushed: frames stream through (each 2160p plane pair is 17 MB, the MV field 19 MB)",
               "what": "sustained, end to end: every pair's planes uploaded from page-locked memory, MV field (16-byte records) delivered "
                       "to the host, two pairs in flight, wall clock over the whole region (max over ranks)",
               "gpu_launches": int(launches), "clocks": clocks, "consumer_checksum": checksum[0] % 1000003}
    for p in preps:
        p.free()
    hm.sync()
    for a in outs + frames:
        hm.host_free(a)
    return res


def measure_dist_table(env, blocks_per_size=10000):
    """BASELINE.json configs[1]: the distortion table over registered planes - SAD (with and without iSubShift), SSE and
    HADs for all 24 PU sizes, 8- and 10-bit, blocks_per_size random block pairs each in ONE hmb200_dist_batch call per
    family (descriptors up, distortions down inside the timed call)."""
    from video_codecs_b200 import DF_SAD, DF_SSE, DF_HADS, DIST_DESC_DTYPE, synth
    hm = env.hm
    if env.rank != 0:
        return None
    sizes = [(64, 64), (64, 32), (32, 64), (64, 16), (64, 48), (16, 64), (48, 64), (32, 32), (32, 16), (16, 32), (32, 8), (32, 24),
             (8, 32), (24, 32), (16, 16), (16, 8), (8, 16), (16, 4), (16, 12), (4, 16), (12, 16), (8, 8), (8, 4), (4, 8)]
    W, H = 1920, 1088
    out = {"blocks_per_size": blocks_per_size, "sizes": len(sizes), "rows": []}
    rng = np.random.default_rng(5)
    for bd in (8, 10):
        f0, f1 = synth.luma_frame(W, H, 0, seed=3, bit_depth=bd), synth.luma_frame(W, H, 1, seed=3, bit_depth=bd)
        reg = (lambda f, k: hm.register_plane_u8(f, 80, 80, kind=k)) if bd == 8 else (lambda f, k: hm.register_plane_u16(f, bd, 80, 80, kind=k))
        ido, idc = reg(f1, 0), reg(f0, 1)
        descs = np.zeros(len(sizes) * blocks_per_size, dtype=DIST_DESC_DTYPE)
        i = 0
        for (w, h) in sizes:
            n = blocks_per_size
            d = descs[i:i + n]
            d["org_plane"], d["cur_plane"], d["w"], d["h"] = ido, idc, w, h
            d["org_x"] = rng.integers(0, W - w, n); d["org_y"] = rng.integers(0, H - h, n)
            d["cur_x"] = d["org_x"] + rng.integers(-8, 9, n); d["cur_y"] = d["org_y"] + rng.integers(-8, 9, n)
            i += n
        px = float(np.sum(descs["w"].astype(np.int64) * descs["h"]))
        for name, func, ss in (("SAD", DF_SAD, 0), ("SAD_subshift1", DF_SAD, 1), ("SSE", DF_SSE, 0), ("HADs", DF_HADS, 0)):
            dd = descs.copy()
            dd["sub_shift"] = np.where(dd["h"] > 8, ss, 0)
            hm.dist_batch(func, bd, dd[:1000])
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                got = hm.dist_batch(func, bd, dd)
            dt = (time.perf_counter() - t0) / reps
            kms = hm.last_timing()["total_ms"]
            out["rows"].append({"func": name, "bit_depth": bd, "blocks": int(len(dd)), "ms_per_call": 1e3 * dt, "blocks_per_s": len(dd) / dt,
                                "samples_per_s": px / dt, "kernel_ms": kms, "kernel_samples_per_s": px / (kms / 1e3),
                                "kernel_read_gb_per_s": 2 * px * (1 if bd == 8 else 2) / (kms / 1e3) / 1e9,
                                "checksum": int(got.astype(np.uint64).sum() % 1000003)})
        hm.release_plane(ido)
        hm.release_plane(idc)
    peak_abs, _ = int_simd_peak(8)
    hbm, _ = hbm_peak()
    best = max(r["kernel_samples_per_s"] for r in out["rows"] if r["func"] == "SAD" and r["bit_depth"] == 8)
    out["sad8_kernel_frac_of_simd_peak"] = best / peak_abs
    out["kernel_read_frac_of_hbm_peak"] = max(r["kernel_read_gb_per_s"] for r in out["rows"]) / hbm
    out["bound"] = ("per call: 9.6 MB of descriptors up, one launch (a warp per block pair), distortions down, synchronise - the host copies are "
                    "most of ms_per_call; kernel_ms is the launch alone (CUDA events).  Every sample is read once from the two L2-resident planes, "
                    "so the kernel is bound by L2 / load issue (kernel_read_gb_per_s), not by the SAD pipe")
    return out


def measure_one_pu_calls(env, calls=1000):
    """The in-encoder shape of the path: ONE PU per call through the 1:1 entries (hmb200_pattern_search[_frac|_and_refine], what the
    forwarders of integration/hm_shim.cpp call once per TEncSearch::xPatternSearch / xPatternSearchFracDIF), host pattern in,
    vectors out, synchronous.  Microseconds per call for a few PU sizes, +-64 window at the centre of a 1080p plane; ctypes
    adds ~4 us per call.  The fused entry must return what the two separate entries return."""
    from video_codecs_b200 import synth
    hm = env.hm
    if env.rank != 0:
        return None
    W, H, M = 1920, 1080, 80
    ref = np.ascontiguousarray(np.pad(synth.luma_frame(W, H, 0).astype(np.int16), M, mode="edge"))
    cur = np.ascontiguousarray(np.pad(synth.luma_frame(W, H, 1).astype(np.int16), M, mode="edge"))
    stride = W + 2 * M
    idr = hm.register_plane(ref, W, H, M, M, 8, kind=1)
    off = (512 + M) * stride + 960 + M
    out = {"calls_per_point": calls, "unit": "us per call (synchronous, incl. ~4 us of ctypes)", "rows": []}

    import itertools
    tick = itertools.count(1)

    def timed(fn):
        for _ in range(30):
            fn()
        t0 = time.perf_counter()
        for _ in range(calls):
            fn()
        return 1e6 * (time.perf_counter() - t0) / calls

    try:
        for w, h in ((8, 8), (16, 16), (32, 32), (64, 64)):
            org, rf = (cur, off, stride), (ref, off, stride)
            mv, sad = hm.pattern_search(org, w, h, rf, (-64, -64), (64, 64), LAMBDA_COST, (0, 0))
            fr = hm.pattern_search_frac(org, w, h, rf, mv, LAMBDA_COST, (0, 0))
            if hm.pattern_search_and_refine(org, w, h, rf, (-64, -64), (64, 64), LAMBDA_COST, (0, 0)) != (mv, sad) + fr:
                raise SystemExit("bench: the fused 1:1 entry differs from search + refinement")
            out["rows"].append({
                "pu": f"{w}x{h}",
                "search": timed(lambda: hm.pattern_search(org, w, h, rf, (-64, -64), (64, 64), LAMBDA_COST, (0, 0))),
                "refine": timed(lambda: hm.pattern_search_frac(org, w, h, rf, mv, LAMBDA_COST, (0, 0))),
                # a new lambda per call: an identical repeat would be answered from the previous call's whole-CU launch
                "search_and_refine_2Nx2N": timed(lambda: hm.pattern_search_and_refine(org, w, h, rf, (-64, -64), (64, 64),
                                                                                      LAMBDA_COST + next(tick), (0, 0)))})
        # HM's call order for a CU: the 2Nx2N PU, then 2NxN / Nx2N / the AMP splits with the same window, predictor and lambda (what 90 %
        # of the non-2Nx2N calls of the 1080p test clip look like); lambda changes per repetition so that every 2Nx2N call launches
        out["cu_sequences"] = []
        for S in (8, 16, 32, 64):
            q, hf = S // 4, S // 2
            parts = [(0, 0, S, S), (0, 0, S, hf), (0, hf, S, hf), (0, 0, hf, S), (hf, 0, hf, S)]
            if S > 8:
                parts += [(0, 0, S, q), (0, q, S, S - q), (0, 0, S, S - q), (0, S - q, S, q), (0, 0, q, S), (q, 0, S - q, S),
                          (0, 0, S - q, S), (S - q, 0, q, S)]
            cu = np.ascontiguousarray(cur[512 + M:512 + M + S, 960 + M:960 + M + S])
            reps = max(20, calls // 10)
            c0 = hm.one_call_stats()
            t0 = time.perf_counter()
            for k in range(reps):
                for (ox, oy, w, h) in parts:
                    hm.pattern_search_and_refine((cu, oy * S + ox, S), w, h, (ref, off + oy * stride + ox, stride), (-64, -64), (64, 64),
                                                 LAMBDA_COST + k, (0, 0))
            dt = time.perf_counter() - t0
            c1 = hm.one_call_stats()
            out["cu_sequences"].append({"cu": f"{S}x{S}", "pus": len(parts), "us_per_cu": 1e6 * dt / reps, "us_per_pu": 1e6 * dt / reps / len(parts),
                                        "calls": c1[0] - c0[0], "whole_cu_launches": c1[1] - c0[1], "answered_from_cu_launch": c1[2] - c0[2]})
    finally:
        hm.release_plane(idr)
    out["note"] = ("kernels of hmb200_one.cuh: one launch per call for PUs up to 16x16 (pattern in the kernel arguments, the search's last "
                   "CTA refines), result as two 16-byte records in mapped host memory; round-trip floor of one trivial launch: 7.2 us "
                   "(profiles/r02_latency_1to1.txt).  rows: one square PU; search_and_refine_2Nx2N searches and refines the whole CU; cu_sequences: "
                   "a CU's partitions in HM's order, the 2Nx2N call searches and refines all of them, the others are answered from it")
    return out


def run_ours(args):
    env = Env(args)
    wl_key = args.workload
    t_start = time.perf_counter()
    if args.only:
        line = {"only": args.only}                       # development aid: one extra workload, no headline (not a bench line)
    else:
        line = measure_pairs(env, wl_key, args.steps, args.warmup, shard=args.shard, search=args.search,
                             preds=args.preds, check_merge=(args.shard == "tiles"))
    extras = {}
    if args.only or (args.extras and wl_key == "1080p" and args.shard == "frames" and args.search == "full" and args.preds == "zero"):
        def extra(name, fn):
            if args.only and name != args.only:
                return
            t0 = time.perf_counter()
            try:
                r = fn()
            except SystemExit:
                raise
            except Exception as e:                      # an extra workload never takes the headline down
                r = {"error": f"{type(e).__name__}: {e}"} if env.rank == 0 else None
            if env.rank == 0 and r is not None:
                r["leg_wall_s"] = time.perf_counter() - t0
                extras[name] = r
            log(f"[bench] {name}: {time.perf_counter() - t0:.1f} s")
        extra("tiles_2160p10", lambda: measure_pairs(env, "2160p10", 4, 3, shard="tiles", check_merge=True, with_hbm=False))
        extra("stream_2160p8", lambda: measure_stream(env, "2160p8", 64 if env.world == 1 else 32, 2 if env.world == 1 else 4))
        extra("unfused_1080p", lambda: measure_pairs(env, "1080p", 5, 3, preds="random", with_hbm=False))
        if env.world == 1:
            extra("dist_table", lambda: measure_dist_table(env))
            extra("one_pu_calls", lambda: measure_one_pu_calls(env))
    if env.rank == 0:
        if extras:
            line["extra_workloads"] = extras
        if env.world == 1 and not args.no_cpu_baseline and wl_key == "1080p" and not args.only:
            line["cpu_baseline"] = cpu_sample_single(WORKLOADS[wl_key], args.cpu_ctus)
        line["bench_wall_s"] = time.perf_counter() - t_start
        emit(line)
    env.hm.shutdown()
    if env.dist is not None:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-ctus", type=int, default=16, help="CTUs in the single-core cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encoder-leg", action="store_true", help="reference arm: skip the stock TAppEncoder run of configs[0]")
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="frames", choices=["frames", "tiles"],
                    help="N > 1: independent frame pairs per rank (weak scaling, default) or tile columns of one picture (strong)")
    ap.add_argument("--search", default="full", choices=["full", "tz"], help="tz: xTZSearch (FastSearch=1) instead of the full search")
    ap.add_argument("--preds", default="zero", choices=["zero", "random"], help="random: a different predictor per PU (no CU fusion)")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="headline only (skip extra_workloads)")
    ap.add_argument("--only", default=None, help="development aid: run just this extra workload (e.g. stream_2160p8) and print it")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: libraries that write banners to fd 1 (NCCL's version line, worker
    # processes) are sent to stderr for the duration of the run; emit() writes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args, WORKLOADS[args.workload])
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
