#!/usr/bin/env python3
"""bench.py -- ORB front-end throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of OrbExtractor::operator() over one batch of synthetic EuRoC-shaped frames
(blocks-v1, 752x480, 1000 features, 8 levels, 1.2, FAST 20/7).  `value` is frames/s with the
frames resident in HBM; `e2e` is the same batch through the C ABI with pinned HOST buffers
(H2D of the frames and D2H of keypoints + descriptors inside the timed region).  Also reported:
p50 latency of one blocking single-frame call, brute-force Hamming search (config 5:
1000 queries x 10M rows sharded over the ranks, NCCL all-gather of the per-rank top-2), the
roofline of the dominant kernel from live CUDA-event stage timings, and the reference's CPU
extractor timed on this host (oracle/_ref: the reference's orb_extractor.cc on the mini-cv shim).
With --impl reference only that CPU implementation is timed (all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT, NLEV, SCALE, INI_TH, MIN_TH = 752, 480, 1000, 8, 1.2, 20, 7
METRIC = "orb_frames_per_s_752x480_1000kp"
DB_ROWS, N_QUERIES = 10_000_000, 1000
KNN_XU_OPS, KNN_ALU_OPS = 6, 24   # per 256-bit pair in k_knn2's unrolled inner loop (profiles/r2_sass_excerpt.txt: 72 POPC, 235 LOP3, 36 VIMNMX, 17 IADD3 for 12 pairs;
                                  # the 57 IMAD that do the additions issue on the FMA pipe and are not the limiter)
WORKLOAD = "config 1 batched: blocks-v1 752x480, 1000 features, 8 levels, scale 1.2, FAST 20/7"


def level_sizes(w, h, nlev=NLEV):
    s, out = np.float32(1.0), []
    for l in range(nlev):
        if l:
            s = np.float32(np.float64(s) * np.float64(np.float32(SCALE)))
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def algorithmic_bytes(w, h, n_kp):
    """compulsory bytes per frame, per stage (every stage reads its input once and
    writes its output once, u8)."""
    px = [a * b for a, b in level_sizes(w, h)]
    return {
        "pyramid": sum(px[:-1]) + sum(px[1:]),
        "fast_blur": sum(px) + 2 * sum(px),   # FAST read + blur read/write (one fused kernel, reads the tile once)
        "describe": n_kp * 749 + n_kp * (512 + 32 + 28),
    }


def cv2_primitive_times(frames):
    """SURVEY.md 8(d) cross-check: the OpenCV primitives of the path (7 resizes, 8 clone + GaussianBlur, FAST with
    NMS at iniThFAST on the 8 levels) in real, SIMD-optimised OpenCV on one thread.  The shim under the compiled reference
    (oracle/minicv) is plain C, so this is the floor a reference built against real OpenCV cannot beat -- its
    quadtree, orientation and descriptor loops (the reference's own C++) come on top."""
    try:
        import cv2
    except ImportError:
        return None
    cv2.setNumThreads(1)
    fast = cv2.FastFeatureDetector_create(INI_TH, True)  # the reference detects at iniThFAST and retries at minThFAST only in empty cells
    sizes = level_sizes(W, H)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        for f in frames:
            lv = [f]
            for (w, h) in sizes[1:]:
                lv.append(cv2.resize(lv[-1], (w, h), interpolation=cv2.INTER_LINEAR))
            for im in lv:
                fast.detect(im)
                cv2.GaussianBlur(im.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        dt = (time.perf_counter() - t0) / len(frames)
        best = dt if best is None else min(best, dt)
    return {"ms_per_frame": 1e3 * best, "frames_per_s_bound": 1.0 / best, "cores": 1, "opencv": cv2.__version__,
            "what": "7 cv2.resize + 8 (copy + cv2.GaussianBlur 7x7) + 8 cv2.FAST(iniThFAST, nms) per 752x480 frame, "
                    "cv2.setNumThreads(1); excludes the reference's own quadtree / orientation / descriptor code"}


def synth_vocabulary(k, L, seed):
    """ORBvoc-shaped synthetic vocabulary tree in breadth-first node order (numpy, seeded): node arrays as
    orbv_create takes them.  A child's descriptor is its parent's with ~256 >> level bits flipped; leaf weights
    are idf-like positive numbers, 1 in 29 words is stopped (weight 0)."""
    rng = np.random.default_rng(seed)
    total = (k ** (L + 1) - 1) // (k - 1)
    parent = np.zeros(total, np.int32)
    leaf = np.zeros(total, np.uint8)
    desc = np.zeros((total, 32), np.uint8)
    weight = np.zeros(total, np.float64)
    first, prev_first, prev_n = 1, 0, 1
    for lev in range(1, L + 1):
        cnt = prev_n * k
        ids = np.arange(first, first + cnt)
        par = prev_first + (ids - first) // k
        parent[ids] = par
        if lev == 1:
            desc[ids] = rng.integers(0, 256, (cnt, 32), dtype=np.uint8)
        else:
            d = desc[par].copy()
            rows = np.arange(cnt)
            for _ in range(256 >> lev):
                pos = rng.integers(0, 256, cnt)
                d[rows, pos >> 3] ^= (1 << (pos & 7)).astype(np.uint8)
            desc[ids] = d
        if lev == L:
            leaf[ids] = 1
            w = (rng.integers(1, 998, cnt)).astype(np.float64) / 64.0
            w[28::29] = 0.0
            weight[ids] = w
        prev_first, prev_n, first = first, cnt, first + cnt
    return parent, leaf, desc, weight


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons.  The sampler is started well before the timed region
    (nvidia-smi takes ~0.1 s to produce its first row) and every row is stamped on arrival;
    summary(t0, t1) reports the rows that fell inside the timed region [t0, t1]."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu, power=False):
        self.rows, self.proc, self.power = [], None, power
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q + (",power.draw" if power else ""),
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        t_end = time.time() + timeout
        while self.proc and not self.rows and time.time() < t_end:
            time.sleep(0.01)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.02 and len(r) >= 6]
        where = "inside the timed region"
        if not rows:  # region shorter than the sampling period: the rows right around it
            near = sorted(self.rows, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))[:3]
            rows = [r for (_, r) in near if len(r) >= 6]
            where = "nearest to the timed region (region shorter than the sampling period)"
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k] == "Active" for r in rows)]
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(sm), "where": where}
        if self.power:
            pw = []
            for r in rows:
                try:
                    pw.append(float(r[6]))
                except (IndexError, ValueError):
                    pass
            out["power_w_median"] = float(np.median(pw)) if pw else None
            out["power_w_max"] = max(pw) if pw else None
        return out


# ---------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_rate(threads, frames_per_thread, first_frame=0):
    """Frames/s of the reference's own orb_extractor.cc (oracle/_ref) or, where that binary is
    missing, of the oracle port; one extractor object per thread (frame.cc:179-182 runs one
    extractor per image thread).  Returns (frames/s, kind)."""
    from oracle import oracle as O
    from oracle import ref as R
    kind = "reference" if R.available(try_build=False) else "port"
    mk = (lambda: R.Extractor(NFEAT, SCALE, NLEV, INI_TH, MIN_TH)) if kind == "reference" else \
         (lambda: O.Extractor(NFEAT, SCALE, NLEV, INI_TH, MIN_TH))
    imgs = [[O.blocks_v1(W, H, 1, first_frame + t * frames_per_thread + f) for f in range(frames_per_thread)]
            for t in range(threads)]
    exs = [mk() for _ in range(threads)]
    counts = [0] * threads

    def work(t):
        for im in imgs[t]:
            exs[t](im)
            counts[t] += 1

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return sum(counts) / dt, kind, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_thread = max(1, -(-args.batch // threads))  # the repo arm's frames per step, split over the host threads
    for _ in range(args.warmup):
        cpu_reference_rate(threads, 1)
    rates, total = [], 0.0
    for s in range(args.steps):
        r, kind, dt = cpu_reference_rate(threads, per_thread, first_frame=s * threads * per_thread)
        rates.append(r)
        total += dt
    value = threads * per_thread * args.steps / total
    sample = "%d steps x %d frames (%d threads x %d), blocks-v1 752x480" % (args.steps, threads * per_thread, threads, per_thread)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": args.batch},
        "workload_detail": {"frames_per_step_timed": threads * per_thread, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))



# ---------------------------------------------------------------------------- BASELINE configs 2, 3, 4
def algorithmic_bytes_total(w, h, n_kp, nlev=NLEV):
    """B = sum_{l<7} px_l + sum_{l>=1} px_l + sum px + 2 sum px + N * 749 + N * 572."""
    px = [a * b for a, b in level_sizes(w, h, nlev)]
    return sum(px[:-1]) + sum(px[1:]) + 3 * sum(px) + n_kp * (749 + 512 + 32 + 28)


def time_batches(torch, step, steps, warmup, barrier, max_over_ranks):
    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    return ms / steps


def run_other_configs(P, A, torch, world, rank, local, dev, barrier, max_over_ranks, peak, with_cpu):
    """BASELINE.json configs 2-4, each with its own roofline and, at N=1, a bounded CPU leg on
    the reference's own code (oracle/_ref).  Config 5 is the `matching` key."""
    out = {}
    stream = A.torch_stream(dev)

    def batch_case(w, h, nfeat, n_frames, chunk, first_frame, steps):
        ex = P.OrbExtractor(nfeat, SCALE, NLEV, INI_TH, MIN_TH, device=local, max_batch=chunk)
        frames = P.synth_frames("blocks", n_frames, w, h, seed=1, first_frame=first_frame, device=local)
        cap = ex.max_keypoints() + 8
        kps = torch.empty((n_frames, cap, 7), dtype=torch.float32, device=dev)
        desc = torch.empty((n_frames, cap, 32), dtype=torch.uint8, device=dev)
        n = torch.empty(n_frames, dtype=torch.int32, device=dev)
        nm = torch.empty(n_frames, dtype=torch.int32, device=dev)

        def step():
            ex.extract_batch_into(frames.data_ptr(), n_frames, w, h, frames.stride(1), frames.stride(0), A.MEM_DEVICE, (0, 0),
                                  kps.data_ptr(), desc.data_ptr(), cap, n.data_ptr(), nm.data_ptr(), stream)
        step()
        torch.cuda.synchronize()
        if int(n.min().item()) < 0:   # the geometry-free capacity was too small for this image size: size it from the first pass
            cap = int(-n.min().item()) + 8
            kps = torch.empty((n_frames, cap, 7), dtype=torch.float32, device=dev)
            desc = torch.empty((n_frames, cap, 32), dtype=torch.uint8, device=dev)
        ms = time_batches(torch, step, steps, 3, barrier, max_over_ranks)
        nh = n.cpu().numpy()
        assert (nh > 0).all(), "extraction failed"
        return ms, float(nh.mean()), ex, frames

    # ---- config 3: KITTI-shaped 1241x376, 2000 features, batch of 64 frames (every rank its own 64: weak)
    ms3, kp3, ex3, fr3 = batch_case(1241, 376, 2000, 64, 64, rank * 64, 20)
    b3 = algorithmic_bytes_total(1241, 376, kp3)
    v3 = world * 64 / (ms3 * 1e-3)
    out["config3"] = {"workload": "64 blocks-v1 frames 1241x376, 2000 features, 8 levels per GPU per step", "value": v3, "unit": "frames/s",
                      "ms_per_batch": ms3, "mean_keypoints_per_frame": kp3,
                      "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": b3, "achieved": b3 * v3 / world / 1e9, "peak": peak,
                                   "unit": "GB/s", "frac": b3 * v3 / world / 1e9 / peak},
                      "l2": "working set 0.7 GB per step (pyramid + blurred planes of 64 frames) exceeds the 126 MB L2"}
    if with_cpu:
        out["config3"]["cpu_baseline"] = cpu_reference_case(1241, 376, 2000, 12)
    del ex3, fr3

    # ---- config 4: 4096 frames 1280x720, 1000 features, sharded by frame over the ranks (strong scaling, no collective)
    total4 = 4096
    f0, f1 = total4 * rank // world, total4 * (rank + 1) // world
    ms4, kp4, ex4, fr4 = batch_case(1280, 720, 1000, f1 - f0, 256, f0, 3)
    b4 = algorithmic_bytes_total(1280, 720, kp4)
    v4 = total4 / (ms4 * 1e-3)
    out["config4"] = {"workload": "4096 blocks-v1 frames 1280x720, 1000 features, frames [r*4096/G, (r+1)*4096/G) on rank r, 256-frame chunks",
                      "value": v4, "unit": "frames/s", "scaling": "strong", "ms_per_pass": ms4, "frames_per_rank": f1 - f0,
                      "mean_keypoints_per_frame": kp4,
                      "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": b4, "achieved": b4 * v4 / world / 1e9, "peak": peak,
                                   "unit": "GB/s", "frac": b4 * v4 / world / 1e9 / peak}}
    if with_cpu:
        out["config4"]["cpu_baseline"] = cpu_reference_case(1280, 720, 1000, 8)
    del ex4, fr4
    torch.cuda.empty_cache()

    # ---- config 2: EuRoC-shaped stereo pair, 1200 features per side + ComputeStereoMatches, one blocking stereo frame at a
    # time through the C ABI with host buffers (frame.cc:139-235 = two extractions + :828-986), rank 0
    if rank == 0:
        import orb_slam_fusion_b200.orb_matcher  # noqa: F401
        m = P.ORBmatcher(device=local)
        exl = P.OrbExtractor(1200, SCALE, NLEV, INI_TH, MIN_TH, device=local, max_batch=1)
        exr = P.OrbExtractor(1200, SCALE, NLEV, INI_TH, MIN_TH, device=local, max_batch=1)
        pair = P.synth_frames("blocks", 1, W, H, seed=1, first_frame=3, device=local)
        right = P.synth_frames("blocks", 1, W, H, seed=1, first_frame=3, shift_x=12, noise_seed=2, device=local)
        hl, hr = pair[0].cpu().numpy(), right[0].cpu().numpy()
        bf, mb = np.float32(47.90639384423901), np.float32(0.11)   # settings/EuRoC.yaml: fx * baseline; minZ = mb (frame.cc:853-856)
        cap = 1400
        kl, dl = np.empty(cap, P.KP_DTYPE), np.empty((cap, 32), np.uint8)
        kr, dr = np.empty(cap, P.KP_DTYPE), np.empty((cap, 32), np.uint8)

        ur_buf, dp_buf = np.empty(cap, np.float32), np.empty(cap, np.float32)

        def stereo_frame():
            # frame.cc:139-235: both images in flight at once (the reference uses two threads, :179-182; here one host thread
            # and orbx_extract_begin / _end), then ComputeStereoMatches on the device-resident results in ONE call
            exl.extract_begin(hl)
            exr.extract_begin(hr)
            _, nl = exl.extract_end(kl, dl)
            _, nr = exr.extract_end(kr, dr)
            return m.stereo_matches_last(exl, exr, bf, mb, ur_buf, dp_buf), nl, nr
        for _ in range(20):
            (ur, dp), nl, nr = stereo_frame()
        lat = []
        for _ in range(300):
            t0 = time.perf_counter()
            stereo_frame()
            lat.append(time.perf_counter() - t0)
        p50 = 1e3 * float(np.median(lat))
        b2 = 2 * algorithmic_bytes_total(W, H, (nl + nr) / 2)
        out["config2"] = {"workload": "stereo pair 2 x 752x480, 1200 features per side: both extractions in flight together (orbx_extract_begin / _end), "
                                      "keypoints + descriptors back on the host, then ComputeStereoMatches (row band + SAD refinement + median cut) in "
                                      "one C-ABI call on the device-resident results",
                          "p50_ms_per_stereo_frame": p50, "stereo_frames_per_s": 1e3 / p50, "keypoints": [int(nl), int(nr)],
                          "stereo_matches": int((ur >= 0).sum()),
                          "roofline": {"bound": "latency", "algorithmic_bytes_per_stereo_frame": b2, "achieved": b2 / (p50 * 1e-3) / 1e9,
                                       "peak": peak, "unit": "GB/s", "frac": b2 / (p50 * 1e-3) / 1e9 / peak,
                                       "note": "one 1.1 MB pyramid per side is L2-resident: a blocking stereo frame is launch- and dependency-bound, "
                                               "SURVEY.md 8(d) asks for ms here"}}
        if with_cpu:
            out["config2"]["cpu_baseline"] = cpu_stereo_case(hl, hr, bf, mb, (kl[:nl].copy(), dl[:nl].copy(), kr[:nr].copy(), dr[:nr].copy(), ur.copy(), dp.copy()))
        del exl, exr, m
    return out


def cpu_reference_case(w, h, nfeat, n_frames):
    """The reference's extractor (oracle/_ref, else the oracle port) on one thread over n_frames frames of another geometry."""
    from oracle import oracle as O
    from oracle import ref as R
    kind = "reference" if R.available(try_build=False) else "port"
    ex = (R.Extractor if kind == "reference" else O.Extractor)(nfeat, SCALE, NLEV, INI_TH, MIN_TH)
    imgs = [O.blocks_v1(w, h, 1, f) for f in range(n_frames)]
    ex(imgs[0])
    t0 = time.perf_counter()
    for im in imgs:
        ex(im)
    dt = time.perf_counter() - t0
    return {"value": n_frames / dt, "unit": "frames/s", "cores": 1, "kind": kind,
            "sample": "%d blocks-v1 %dx%d frames, %d features, one extractor object on one thread" % (n_frames, w, h, nfeat)}


def cpu_stereo_case(hl, hr, bf, mb, gpu):
    """frame.cc:139-235 on the host: the reference's extractor on both images (sequentially, one thread) + the reference's own
    ComputeStereoMatches lines (oracle/_ref/libframe_ref.so); also checks the GPU's mvuRight / mvDepth against them."""
    from oracle import oracle as O
    from oracle import ref as R
    if not (R.available(try_build=False) and R.frame_available()):
        return None
    exl, exr = R.Extractor(1200, SCALE, NLEV, INI_TH, MIN_TH), R.Extractor(1200, SCALE, NLEV, INI_TH, MIN_TH)
    ol, orr = O.Extractor(1200, SCALE, NLEV, INI_TH, MIN_TH), O.Extractor(1200, SCALE, NLEV, INI_TH, MIN_TH)
    t = ol.tables()
    reps, t_ex, t_st = 5, 0.0, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        _, kl, dl = exl(hl)
        _, kr, dr = exr(hr)
        t1 = time.perf_counter()
        ol.compute_pyramid(hl)
        orr.compute_pyramid(hr)
        ll = [ol.level(l, with_border=True) for l in range(NLEV)]
        lr = [orr.level(l, with_border=True) for l in range(NLEV)]
        t2 = time.perf_counter()
        ur, dp = R.stereo_matches(ll, lr, kl, dl, kr, dr, t["scale"], t["inv_scale"], bf, mb)
        t3 = time.perf_counter()
        t_ex += t1 - t0
        t_st += t3 - t2
    # the reference's ComputeStereoMatches lines on the GPU's own keypoints / descriptors must give the GPU's mvuRight / mvDepth
    gkl, gdl, gkr, gdr, gpu_ur, gpu_dp = gpu
    wur, wdp = R.stereo_matches(ll, lr, gkl, gdl, gkr, gdr, t["scale"], t["inv_scale"], bf, mb)
    same = bool(wur.tobytes() == gpu_ur.tobytes() and wdp.tobytes() == gpu_dp.tobytes())
    assert same, "GPU ComputeStereoMatches != the reference's lines on the same stereo pair"
    return {"value": 1e3 * (t_ex + t_st) / reps, "unit": "ms per stereo frame", "cores": 1, "kind": "reference",
            "extraction_ms": 1e3 * t_ex / reps, "stereo_matches_ms": 1e3 * t_st / reps, "equal_to_gpu": same,
            "sample": "%d stereo frames: two extractions on one thread (the reference uses two, frame.cc:179-182) + the reference's "
                      "ComputeStereoMatches lines (frame.cc:828-986)" % reps}

# ---------------------------------------------------------------------------- B200 arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import orb_slam_fusion_b200 as P
    from orb_slam_fusion_b200 import _abi as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    debug = bool(os.environ.get("BENCH_DEBUG"))

    def log(msg):
        if debug:
            print("[bench rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout for the one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    log('process group ready')
    clocks = ClockSampler(local)
    B = args.batch
    frames = P.synth_frames("blocks", B, W, H, seed=1, first_frame=rank * B, device=local)
    ex = P.OrbExtractor(NFEAT, SCALE, NLEV, INI_TH, MIN_TH, device=local, max_batch=B)
    cap = ex.max_keypoints() + 8
    kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
    desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    n = torch.empty(B, dtype=torch.int32, device=dev)
    nm = torch.empty(B, dtype=torch.int32, device=dev)
    stream = A.torch_stream(dev)

    def step():
        ex.extract_batch_into(frames.data_ptr(), B, W, H, frames.stride(1), frames.stride(0), A.MEM_DEVICE, (0, 0),
                              kps.data_ptr(), desc.data_ptr(), cap, n.data_ptr(), nm.data_ptr(), stream)

    for _ in range(args.warmup):
        step()
    barrier()
    ex.set_profiling(True)
    ex.stage_times(reset=True)
    launches0 = ex.launch_count()
    clocks.wait_first()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t_region1 = time.time()
    ms_local = e0.elapsed_time(e1)
    ms = max_over_ranks(ms_local)
    ms_ranks = [ms_local]
    if world > 1:  # every rank's own device time: the reported value uses the slowest
        tl = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(tl, torch.tensor([ms_local], dtype=torch.float64, device=dev))
        ms_ranks = [float(t.item()) for t in tl]
    barrier()
    clocks.stop()
    clk = clocks.summary(t_region0, t_region1)
    stage_ms, chunks = ex.stage_times(reset=True)
    ex.set_profiling(False)
    launches = ex.launch_count() - launches0
    n_host = n.cpu().numpy()
    assert (n_host > 0).all(), "extraction failed"
    mean_kp = float(n_host.mean())
    value = world * B * args.steps / (ms * 1e-3)

    log('device-resident timing done: %.3f ms' % ms)
    # ---- the same step back to back for >= 2 s (the 20-step region above lasts ~50 ms): clocks and power under sustained load
    sus_steps = int(max(args.steps, np.ceil(2200.0 / max(ms / args.steps, 1e-3))))
    sclk = ClockSampler(local, power=True)
    sclk.wait_first()
    barrier()
    t_s0 = time.time()
    s0e, s1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0e.record()
    for _ in range(sus_steps):
        step()
    s1e.record()
    torch.cuda.synchronize()
    t_s1 = time.time()
    sus_ms = max_over_ranks(s0e.elapsed_time(s1e))
    barrier()
    sclk.stop()
    sustained = {"value": world * B * sus_steps / (sus_ms * 1e-3), "unit": "frames/s", "steps": sus_steps, "seconds": sus_ms * 1e-3,
                 "ms_per_step": sus_ms / sus_steps, "clocks": sclk.summary(t_s0, t_s1)}
    # ---- end to end through the C ABI with pinned host buffers
    eb = ex_e2e = None
    E2E_CHUNK = min(64, B)  # frames per pipelined chunk of the host-memory path (the handle's max_batch)
    eb = P.OrbExtractor(NFEAT, SCALE, NLEV, INI_TH, MIN_TH, device=local, max_batch=E2E_CHUNK)
    h_frames = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(frames)
    h_kps = torch.empty((B, cap, 7), dtype=torch.float32, pin_memory=True)
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_n = torch.empty(B, dtype=torch.int32, pin_memory=True)
    h_nm = torch.empty(B, dtype=torch.int32, pin_memory=True)

    def step_e2e():
        # pinned buffers, asynchronous host-memory mode: the H2D copy of every step's frames and the D2H
        # copy of its keypoints + descriptors are enqueued by the call; one orbx_sync ends the timed region
        eb.extract_batch_into(h_frames.data_ptr(), B, W, H, W, W * H, A.MEM_HOST_ASYNC, (0, 0), h_kps.data_ptr(),
                              h_desc.data_ptr(), cap, h_n.data_ptr(), h_nm.data_ptr(), None)

    for _ in range(max(1, args.warmup)):
        step_e2e()
    eb.sync()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    eb.sync()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert np.array_equal(h_n.numpy(), n_host), "host-memory path disagrees with device-memory path"
    # the host-memory path must return exactly what the device-memory path wrote: every keypoint record and descriptor row
    d_kps, d_desc, d_nm = kps.cpu().numpy(), desc.cpu().numpy(), nm.cpu().numpy()
    hk, hd = h_kps.numpy(), h_desc.numpy()
    assert np.array_equal(h_nm.numpy(), d_nm), "n_mono differs between the host-memory and the device-memory path"
    for f in range(B):
        c = int(n_host[f])
        assert hk[f, :c].tobytes() == d_kps[f, :c].tobytes() and np.array_equal(hd[f, :c], d_desc[f, :c]), \
            "frame %d: host-memory path returned other keypoints / descriptors than the device-memory path" % f
    del d_kps, d_desc, hk, hd
    e2e = {"value": world * B * args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": B * W * H,
           "d2h_bytes_per_step": B * (cap * (28 + 32) + 8), "ms_per_step": 1e3 * e2e_s / args.steps,
           "chunk_frames": E2E_CHUNK, "h2d_gbs": world * B * W * H * args.steps / e2e_s / 1e9,
           "h2d_gbs_per_rank": B * W * H * args.steps / e2e_s / 1e9,
           "checked": "n, n_mono, every keypoint record and descriptor row equal the device-memory path's"}
    del eb, h_kps, h_desc

    log('e2e done')
    # ---- p50 latency of one blocking single-frame call (config 1), rank 0
    p50 = None
    if rank == 0:
        ex1 = P.OrbExtractor(NFEAT, SCALE, NLEV, INI_TH, MIN_TH, device=local, max_batch=1)
        img = h_frames[0].numpy()
        cap1 = ex1.max_keypoints() if hasattr(ex1, "max_keypoints") else cap
        k1, d1 = np.empty(max(cap1, cap), P.KP_DTYPE), np.empty((max(cap1, cap), 32), np.uint8)
        for _ in range(20):
            ex1.extract_into(img, k1, d1)
        lat = []
        for _ in range(1000):  # the blocking C-ABI call orbx_extract: host frame in, keypoints + descriptors out
            t0 = time.perf_counter()
            ex1.extract_into(img, k1, d1)
            lat.append(time.perf_counter() - t0)
        p50 = 1e3 * float(np.median(lat))
        del ex1

    log('latency done')
    # ---- Hamming search: 1000 queries x 10M rows, rows sharded over the ranks (config 5)
    m = P.ORBmatcher(0.7, device=local)
    r0, r1 = DB_ROWS * rank // world, DB_ROWS * (rank + 1) // world
    db = P.synth_descriptors(r0, r1 - r0, seed=7, device=local)
    q = P.synth_descriptors(0, N_QUERIES, seed=8, device=local)
    # known answers in EVERY shard: query i's true match (8 flipped bits) sits at global row
    # i * 9973 + 12345 -- spread over all shards -- and every 10th query also has a decoy 2 bits further away in the shard
    # "opposite" to its match, so the ratio test must reject it and the merged second neighbour must come from another rank
    qi = torch.arange(N_QUERIES, device=dev, dtype=torch.int64)
    true_row = qi * 9973 + 12345
    decoy_row = (true_row + DB_ROWS // 2) % DB_ROWS
    planted = q.clone()
    planted[:, 0] ^= 0xFF
    decoy = planted.clone()
    decoy[:, 1] ^= 0x03
    mine = (true_row >= r0) & (true_row < r1)
    db[true_row[mine] - r0] = planted[mine]
    dmine = (decoy_row >= r0) & (decoy_row < r1) & (qi % 10 == 0)
    db[decoy_row[dmine] - r0] = decoy[dmine]
    from orb_slam_fusion_b200 import sharding

    def match_step():
        return sharding.sharded_knn2(m, q, db, r0, 0.7)   # orbm_knn2_sharded: ONE C-ABI call, one NCCL all-gather inside

    g_idx, g_dist, g_acc = match_step()
    torch.cuda.synchronize()
    has_decoy = (qi % 10 == 0)
    assert torch.equal(g_idx[:, 0], true_row) and bool((g_dist[:, 0] == 8).all()), "rank %d: merged nearest neighbour is not the planted row" % rank
    assert torch.equal(g_idx[has_decoy, 1], decoy_row[has_decoy]) and bool((g_dist[has_decoy, 1] == 10).all()), \
        "rank %d: merged second neighbour is not the planted decoy" % rank
    assert bool((g_dist[~has_decoy, 1] > 60).all()), "rank %d: an unexpected near row" % rank
    assert torch.equal(g_acc.bool(), ~has_decoy), "rank %d: ratio test result" % rank

    for _ in range(2):
        match_step()
    barrier()
    m0 = torch.cuda.Event(enable_timing=True)
    m1 = torch.cuda.Event(enable_timing=True)
    reps = 5
    m0.record()
    for _ in range(reps):
        match_step()
    m1.record()
    torch.cuda.synchronize()
    match_ms = max_over_ranks(m0.elapsed_time(m1)) / reps
    barrier()
    log('matching done')
    pairs_per_s = N_QUERIES * DB_ROWS / (match_ms * 1e-3)
    matching = {"value": pairs_per_s, "unit": "pair-distances/s", "ms_per_search": match_ms,
                "queries_per_s": N_QUERIES / (match_ms * 1e-3), "workload": "1000 queries x 10M rows, top-2 + ratio 0.7",
                "sharding": "database rows over %d rank(s); orbm_knn2_sharded: local top-2 as packed 64-bit keys, ONE ncclAllGather of "
                            "16 B per query and rank on the call's stream, merge + ratio in one kernel" % world,
                "verified": "every rank: planted true matches (all shards) are the merged nearest neighbours, planted decoys in the "
                            "opposite shard the merged second neighbours, ratio test rejects exactly those"}
    if rank == 0:
        popc = P.popc_peak(0, local)
        plain = P.popc_peak(1, local)
        # The bound that binds: a pair costs KNN_XU_OPS popc on the XU pipe (16 lanes / clk / SM, measured above as
        # popc_peak_per_s) and KNN_ALU_OPS LOP3 / IADD3 / VIMNMX on the ALU pipe (64 lanes / clk / SM = 4x the popc rate);
        # the kernel cannot beat the slower of the two pipes.  Counts from the SASS of k_knn2's inner loop (profiles/).
        alu_peak = 4.0 * popc
        two_pipe = min(popc / KNN_XU_OPS, alu_peak / KNN_ALU_OPS)
        matching.update({"popc_peak_per_s": popc, "plain_distance_peak_per_s": plain,
                         "roofline": {"bound": "xu+alu pipes", "xu_ops_per_pair": KNN_XU_OPS, "alu_ops_per_pair": KNN_ALU_OPS,
                                      "xu_bound_pairs_per_s": popc / KNN_XU_OPS, "alu_bound_pairs_per_s": alu_peak / KNN_ALU_OPS,
                                      "peak": two_pipe, "achieved": pairs_per_s / world, "unit": "pair-distances/s/GPU",
                                      "frac": pairs_per_s / world / two_pipe},
                         "frac_of_plain_distance_peak": pairs_per_s / (world * plain)})

    # ---- the tracker's projection matchers as whole calls (host buffers through the C ABI, like the reference's call sites):
    # SearchByProjection(Frame&, vector<MapPoint*>&) and SearchByProjection(CurrentFrame, LastFrame) on one frame of the batch,
    # one window per keypoint (a 3-px-off projection of the keypoint itself), rank 0
    if rank == 0:
        n0 = int(n_host[0])
        k0 = kps[0, :n0].cpu().numpy().view(P.KP_DTYPE).reshape(-1)
        d0 = desc[0, :n0].cpu().numpy()
        rng = np.random.default_rng(0)
        sfac = np.float32(SCALE) ** np.arange(NLEV, dtype=np.float32)
        qw = np.zeros(n0, P.WQ_DTYPE)
        qw["u"] = k0["x"] + rng.normal(0, 3, n0).astype(np.float32)
        qw["v"] = k0["y"] + rng.normal(0, 3, n0).astype(np.float32)
        qw["r"] = (np.float32(4.0 * 3.0) * sfac[np.clip(k0["octave"], 0, NLEV - 1)]).astype(np.float32)
        qw["min_level"], qw["max_level"] = k0["octave"] - 1, k0["octave"]
        geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)

        def p50_us(fn, reps=200):
            for _ in range(20):
                fn()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return 1e6 * float(np.median(ts))

        nm_a, _ = m.SearchByProjection(k0, d0, geom, qw, d0, None, None, None, None, 100, 0.8)
        nm_b, _ = m.SearchByProjectionLast(k0, d0, geom, qw, d0, k0["angle"], None, None, None, None, 100, True)
        matching["tracker"] = {
            "keypoints": n0, "windows": n0,
            "search_by_projection_us": p50_us(lambda: m.SearchByProjection(k0, d0, geom, qw, d0, None, None, None, None, 100, 0.8)),
            "search_by_projection_matches": int(nm_a),
            "search_by_projection_last_frame_us": p50_us(lambda: m.SearchByProjectionLast(k0, d0, geom, qw, d0, k0["angle"], None, None,
                                                                                          None, None, 100, True)),
            "search_by_projection_last_frame_matches": int(nm_b),
            "what": "p50 of one blocking whole-function call through the C ABI with host buffers (window search + greedy claim "
                    "[+ rotation histogram]), orb_matcher.cc:42-134 and :1518-1728 after the projection"}

    # ---- bag of words: Frame::ComputeBoW over this rank's batch of extracted descriptors,
    # ORBvoc-shaped synthetic vocabulary (k=10, L=6), device-resident, frames sharded like the extraction
    del db, q
    vk, vL = 10, 6
    vparent, vleaf, vdesc, vweight = synth_vocabulary(vk, vL, seed=7)
    voc = P.ORBVocabulary(vk, vL, vparent, vleaf, vdesc, vweight, device=local)
    for _ in range(2):
        bow = voc.transform_batch(desc, n, 4)
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(reps):
        bow = voc.transform_batch(desc, n, 4)
    b1.record()
    torch.cuda.synchronize()
    bow_ms = max_over_ranks(b0.elapsed_time(b1)) / reps
    barrier()
    bow_words = float(bow["bow_n"].float().mean().item())
    bow_out = {"value": world * B / (bow_ms * 1e-3), "unit": "frames/s", "ms_per_batch": bow_ms,
               "descriptors_per_s": world * float(n_host.sum()) / (bow_ms * 1e-3),
               "mean_words_per_frame": bow_words,
               "workload": "transform(features, BowVector, FeatureVector, 4) of %d frames x ~%d descriptors per rank, "
                           "synthetic vocabulary k=10 L=6 (%d nodes), TF-IDF / L1" % (B, int(mean_kp), len(vparent))}
    # ---- bag-of-words guided matching: ORBmatcher::SearchByBoW (orb_matcher.cc:215-389) of every frame of the batch
    # against its successor (key frame f -> frame f+1), on the FeatureVectors just computed, device-resident
    pairs = torch.stack([torch.arange(B, dtype=torch.int32), (torch.arange(B, dtype=torch.int32) + 1) % B], 1).to(dev)
    for _ in range(2):
        sb_n, sb_match = m.SearchByBoW(kps, desc, n, bow, pairs, None, 0.7, True)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps):
        sb_n, sb_match = m.SearchByBoW(kps, desc, n, bow, pairs, None, 0.7, True)
    s1.record()
    torch.cuda.synchronize()
    sb_ms = max_over_ranks(s0.elapsed_time(s1)) / reps
    barrier()
    bow_out["search_by_bow"] = {"value": world * B / (sb_ms * 1e-3), "unit": "frame pairs/s", "ms_per_batch": sb_ms,
                                "mean_matches_per_pair": float(sb_n.float().mean().item()),
                                "workload": "SearchByBoW(KF f, Frame f+1, nnratio 0.7, orientation check) over the %d frames "
                                            "of the batch, FeatureVectors at levelsup 4" % B}
    log('bag of words done')
    if rank == 0 and world == 1:  # the CPU restatement of DBoW2 on one thread, bounded sample (checker side: oracle/)
        from oracle import oracle as O
        vo = O.Vocabulary(vk, vL, vparent, vleaf, vdesc, vweight)
        d_host = desc[:32].cpu().numpy()
        t0 = time.perf_counter()
        for f in range(32):
            ids, vals, nodes, feats = vo.transform(d_host[f, :n_host[f]], 4)
        dt = time.perf_counter() - t0
        nb = int(bow["bow_n"][31].item())
        assert nb == len(ids) and np.array_equal(bow["bow_ids"][31, :nb].cpu().numpy().astype(np.uint32), ids)
        assert bow["bow_vals"][31, :nb].cpu().numpy().tobytes() == vals.tobytes(), "bag of words != CPU oracle"
        bow_out["cpu_baseline"] = {"value": 32 / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                                   "sample": "32 frames of this batch through oracle/bow_oracle.c (DBoW2 restatement)"}
        k_host = kps[:33].cpu().numpy().view(O.KP_DTYPE).reshape(33, -1)
        d33 = desc[:33].cpu().numpy()
        fvs = [O.pack_feature_vector(*vo.transform(d33[f, :n_host[f]], 4)[2:]) for f in range(33)]
        t0 = time.perf_counter()
        for f in range(32):
            wnm, want = O.search_by_bow(k_host[f, :n_host[f]], d33[f, :n_host[f]], None, fvs[f],
                                        k_host[f + 1, :n_host[f + 1]], d33[f + 1, :n_host[f + 1]], fvs[f + 1], 0.7, True)
        dt = time.perf_counter() - t0
        assert int(sb_n[31].item()) == wnm and np.array_equal(sb_match[31, :n_host[32]].cpu().numpy(), want), "SearchByBoW != CPU oracle"
        bow_out["search_by_bow"]["cpu_baseline"] = {"value": 32 / dt, "unit": "frame pairs/s", "cores": 1, "kind": "port",
                                                    "sample": "32 pairs of this batch through orc_search_by_bow (oracle/orb_oracle.c)"}
    del voc, bow

    # ---- BASELINE configs 2, 3, 4 (every rank takes part in 3 and 4)
    del kps, desc
    other_configs = run_other_configs(P, A, torch, world, rank, local, dev, barrier, max_over_ranks, measured_peaks()[0], world == 1)
    log('configs 2-4 done')

    sharding.close_comms()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event stage times over the timed region)
    peak, peak_src = measured_peaks()
    alg = algorithmic_bytes(W, H, mean_kp)
    per_launch_ms = {k: v / max(chunks, 1) for k, v in stage_ms.items()}
    dom = max(per_launch_ms, key=per_launch_ms.get)
    # algorithmic bytes of stages outside SURVEY 8(d)'s formula: the import copy is overhead (0), the
    # quadtree reads the candidate list once and writes the selection (5 B per entry)
    alg_stage = dict(alg)
    alg_stage["import"] = 0
    alg_stage["octree"] = 5 * mean_kp * 2
    dom_bytes = alg_stage[dom] * B
    achieved = dom_bytes / (per_launch_ms[dom] * 1e-3) / 1e9
    total_alg = sum(alg.values())
    traffic = None
    try:  # DRAM bytes of the dominant kernel from the committed ncu --set full capture, scaled per launch
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        if tr:
            traffic = tr["bytes_per_launch"] * B / tr["frames_per_launch"]
    except OSError:
        pass
    # The limiter of the dominant kernel is instruction issue, not HBM: executed warp-instructions per frame and the share of
    # them on the integer ALU pipe come from the committed ncu capture (profiles/inst_counts.json, like `traffic`);
    # issue bound = instructions / (4 SMSPs x 148 SMs x clock), ALU-pipe bound = 2 clocks per ALU warp-instruction per SMSP.
    issue = None
    try:
        ic = json.load(open(os.path.join(ROOT, "profiles", "inst_counts.json"))).get(dom)
        if ic:
            clk_hz = 1e6 * (clk.get("sm_mhz") or 1965.0)
            inst = ic["warp_inst_per_frame"] * B
            t_issue = inst / (592.0 * clk_hz) * 1e3
            t_alu = 2.0 * ic["alu_pipe_share"] * inst / (592.0 * clk_hz) * 1e3
            issue = {"warp_inst_per_launch": inst, "thread_inst_per_pixel": ic.get("thread_inst_per_pixel"),
                     "issue_bound_ms": t_issue, "frac_of_issue_bound": t_issue / per_launch_ms[dom],
                     "alu_pipe_share": ic["alu_pipe_share"], "alu_pipe_bound_ms": t_alu, "frac_of_alu_pipe_bound": t_alu / per_launch_ms[dom],
                     "source": ic.get("source")}
    except OSError:
        pass
    roofline = {"bound": "hbm", "binding_limit": "instruction issue / integer ALU pipe (see `issue`)", "issue": issue, "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": alg_stage[dom], "frames_per_launch": B,
                "launch_ms": per_launch_ms[dom],
                "pipeline": {"algorithmic_bytes_per_frame": total_alg,
                             "achieved": total_alg * (value / world) / 1e9, "frac": total_alg * (value / world) / 1e9 / peak},
                "stage_ms_per_launch": per_launch_ms,
                "stage_share": {k: v / sum(per_launch_ms.values()) for k, v in per_launch_ms.items()}}

    # ---- the reference's CPU extractor on this host, bounded sample, 1 thread
    cpu = None
    if world == 1:
        nfr = 0
        t0 = time.perf_counter()
        rate, kind, dt = cpu_reference_rate(1, 40)
        nfr += 40
        # extend the sample to ~10 s of CPU work
        extra = int(min(2000, max(0, (10.0 - dt) * rate)))
        if extra > 40:
            rate2, kind, dt2 = cpu_reference_rate(1, extra, first_frame=40)
            rate = (40 + extra) / (dt + dt2)
            nfr += extra
        cpu = {"value": rate, "unit": "frames/s", "cores": 1, "kind": kind,
               "sample": "%d blocks-v1 752x480 frames, one extractor object on one thread; host has %d logical cores"
                         % (nfr, os.cpu_count() or 0)}
        cpu["cv2_primitives"] = cv2_primitive_times(h_frames[:8].numpy())

    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_per_rank": [round(x / args.steps, 4) for x in ms_ranks],
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B},
        "workload_detail": {"mean_keypoints_per_frame": mean_kp,
                            "l2": "inputs larger than L2 (%.0f MB of frames + %.1f GB working set per step)"
                                  % (B * W * H / 1e6, B * 7.5e6 / 1e9)},
        "sustained": sustained, "configs": other_configs,
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "p50_ms_per_frame": p50, "matching": matching, "bow": bow_out,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="frames per step per GPU")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything any library prints to file descriptor 1 during the run
    # (NCCL's version banner, for one) is sent to stderr, and the line is written to the saved descriptor
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(out_fd, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
