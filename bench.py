#!/usr/bin/env python
"""bench.py — Mpix/s of baseline-JPEG encode, batched 1920x1280 4:2:0 frames (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one pass of the hot path (rgb_to_dct -> init_huffman -> write_jpg, reference
main/encoder.c:158,360,549) over one batch of 1024 synthetic frames PER GPU (weak scaling: every rank
encodes its own 1024 frames, no data-path collective — frames are independent).

  value : whole-job Mpix/s, inputs already resident in HBM, CUDA-event timed on the launching stream
  e2e   : same metric through the C-ABI call with HOST (pinned) buffers, H2D and D2H inside the timed region
  roofline : dominant kernel (k_pixels_to_tokens) against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline : the reference's own C (oracle/_ref/libref.so, built from /root/reference by oracle/Makefile)
                 timed on this box's host cores, one process per core, on a bounded sample of the workload

  parity_checked : SHA-256 of frames 0/1/121/1023 of the timed batch's output against tests/golden/golden.json (outside
                 every timed region): the bytes that were timed are the reference's bytes
  parity_live  : the streams the CPU arm produced in this run (the unmodified reference, 192 frames of the same batch by default)
                 compared byte for byte (size + SHA-256) with the device's output of the last timed step
  sub_records  : the other workloads of BASELINE.json (config 2 = the drop-in entry points one frame at a time, noise and ramp
                 classes, config 5 = 3840x2160, config 3 = the
                 comparator loop through jpegb200_compare_encode_batch beside the reference's loop on the host cores)

`--impl reference` times only that CPU arm (all host cores) and prints the same JSON shape.
"""
import argparse
import ctypes
import hashlib
import importlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

W, H, BATCH = 1920, 1280, 1024
KIND = "natural"
METRIC = "Mpix/s JPEG encode (1920x1280 4:2:0 batch)"
SLOT = 512 * 1024                    # bytes reserved per output stream (natural frames are ~247 KB, noise ~633 KB needs 1 MiB)
FRAME_MPIX = W * H / 1e6


def frames_mod():
    return importlib.import_module("jpeg-encoder-decoder_b200.frames")


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    first, count, reps = args
    import cpu_checkers
    fr = frames_mod()
    batch = np.stack([fr.GENERATORS[KIND](first + i, W, H) for i in range(count)])
    if cpu_checkers.Ref.available():
        chk, kind = cpu_checkers.Ref(), "reference"
    else:
        chk, kind = cpu_checkers.Oracle(), "port"
    chk.time_encode(batch[:1], 1)          # touch code and buffers
    sec, nbytes, streams = chk.time_encode(batch, reps, keep=True)
    # (frame index, bytes, SHA-256) of every stream this process produced: the GPU's bytes are compared with them further down
    return sec, nbytes, count * reps, kind, [(first + i, len(s), hashlib.sha256(s).hexdigest()) for i, s in enumerate(streams)]


def cpu_arm(frames_per_proc: int, nproc: int, reps: int = 1):
    """Reference C on `nproc` host cores (one PROCESS per core: encoder.c keeps global bit-buffer state,
    :383-384, so threads are unsafe).  Returns dict(value Mpix/s, cores, kind, sample, seconds, single)."""
    ctx = mp.get_context("fork")
    with ctx.Pool(nproc) as pool:
        res = pool.map(_cpu_worker, [(i * frames_per_proc, frames_per_proc, reps) for i in range(nproc)])
    wall = max(r[0] for r in res)
    nframes = sum(r[2] for r in res)
    per_core = [r[2] * FRAME_MPIX / r[0] for r in res]
    return dict(value=nframes * FRAME_MPIX / wall, unit="Mpix/s", cores=nproc, kind=res[0][3], streams=[t for r in res for t in r[4]],
                sample=f"{nframes} frames of the {KIND} {W}x{H} batch ({frames_per_proc} per process x {nproc} processes, "
                       f"{reps} pass), timed inside C around rgb_to_dct+init_huffman+write_jpg",
                seconds=wall, per_core_mpix_s=float(np.median(per_core)), jpeg_bytes_per_frame=sum(r[1] for r in res) / nframes)


def _cpu_dropin_worker(args):
    import cpu_checkers
    fr = frames_mod()
    chk = cpu_checkers.Ref() if cpu_checkers.Ref.available() else cpu_checkers.Oracle()
    out = {}
    for name, img in (("sample_640x640_bgr", fr.sample_bgr("640")), ("tile_1920x1280_bgr", fr.tile_bgr(1920, 1280))):
        batch = np.ascontiguousarray(img[None])
        chk.time_encode(batch, 1)
        sec, _ = chk.time_encode(batch, 3)
        out[name] = sec / 3
    return out, "reference" if cpu_checkers.Ref.available() else "port"


def cpu_dropin_arm():
    """One frame at a time through rgb_to_dct + init_huffman + write_jpg of the reference C, single thread (what app_main does)."""
    with mp.get_context("fork").Pool(1) as pool:
        return pool.map(_cpu_dropin_worker, [None])[0]


def _cpu_loop_worker(args):
    w, h, nframes, seed = args
    import cpu_checkers
    seq = frames_mod().moving_sequence(nframes, w, h, seed)
    chk = cpu_checkers.Ref() if cpu_checkers.Ref.available() else cpu_checkers.Oracle()
    chk.time_loop(seq[:2])
    sec, regions, nbytes = chk.time_loop(seq)
    return sec, regions, nbytes, nframes - 1, "reference" if cpu_checkers.Ref.available() else "port"


def cpu_loop_arm(w: int, h: int, nframes: int, nproc: int):
    """app_main's loop (subsample -> compare -> encode regions -> store, main.c:137-162) of the reference C on `nproc`
    host cores, every process on the same scene (frames_mod().moving_sequence)."""
    ctx = mp.get_context("fork")
    with ctx.Pool(nproc) as pool:
        res = pool.map(_cpu_loop_worker, [(w, h, nframes, 5)] * nproc)
    wall = max(r[0] for r in res)
    return dict(frames_per_s=sum(r[3] for r in res) / wall, cores=nproc, kind=res[0][4], regions_per_frame=res[0][1] / res[0][3],
                sample=f"{nframes - 1} frames of the {w}x{h} moving scene per process x {nproc} processes, timed inside C")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kind", default=KIND, choices=["natural", "noise", "ramp"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--shape", default=f"{W}x{H}", help="frame size WxH; 3840x2160 is SURVEY.md 8d config 5 (use --batch 256)")
    ap.add_argument("--frames-per-wave", type=int, default=64)
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--e2e-frames-per-wave", type=int, default=16, help="smaller waves keep the PCIe pipeline of the host path full")
    ap.add_argument("--e2e-lanes", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (noise / ramp / 3840x2160 / comparator loop)")
    ap.add_argument("--cpu-frames-per-proc", type=int, default=12)
    a = ap.parse_args()
    globals()["KIND"] = a.kind
    sw, sh = (int(v) for v in a.shape.lower().split("x"))
    globals().update(W=sw, H=sh, FRAME_MPIX=sw * sh / 1e6, METRIC=f"Mpix/s JPEG encode ({sw}x{sh} 4:2:0 batch)",
                     SLOT=SLOT if sw * sh <= 1920 * 1280 else 2 * 1024 * 1024)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    ncores = len(os.sched_getaffinity(0))
    config = {"workload": f"{a.batch} synthetic {W}x{H} BGR frames per GPU, class '{a.kind}' (tile of the two 640x640 samples, "
                          f"cyclically shifted per frame; SURVEY.md 8d config 4), 4:2:0, per-image optimal Huffman tables, 3 scans",
              "frames_per_gpu": a.batch, "global_frames": a.batch * world, "width": W, "height": H,
              "parallelism": f"frames sharded over {world} GPU(s), no collective",
              "l2": f"inputs ({a.batch * W * H * 3 / 1e9:.1f} GB per GPU) are far larger than L2; no explicit flush needed",
              "frames_per_wave": a.frames_per_wave, "lanes": a.lanes, "slot_bytes": SLOT}

    # ---------------- reference arm: CPU only, rank 0 only
    if a.impl == "reference":
        if rank != 0:
            return
        t0 = time.time()
        vals = []
        for _ in range(a.warmup + a.steps):
            r = cpu_arm(max(1, a.cpu_frames_per_proc // 4), ncores)
            vals.append(r)
        used = vals[a.warmup:] or vals
        v = float(np.mean([r["value"] for r in used]))
        r = used[-1]
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mpix/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1000 * float(np.mean([x["seconds"] for x in used])),
                "step_sample": f"one reference step = {r['sample']} (a bounded sample of the 1024-frame workload; ms_per_step is the time of that sample, "
                               f"value = its frames x Mpix / that time)",
                "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                 "per_core_mpix_s": r["per_core_mpix_s"]},
                "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": time.time() - t0}
        print(json.dumps(line))
        return

    # ---------------- CPU baseline first (before CUDA is initialised in this process: it forks)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu = cpu_arm(a.cpu_frames_per_proc, ncores)
        single = cpu_arm(min(a.cpu_frames_per_proc, 8), 1)
        cpu["single_thread_mpix_s"] = single["value"]
    cpu_loops = {}
    if rank == 0 and world == 1 and not a.no_cpu and not a.no_sub:
        for (lw, lh, lf) in ((640, 640, 33), (1920, 1280, 9)):
            cpu_loops[(lw, lh)] = (cpu_loop_arm(lw, lh, lf, ncores), cpu_loop_arm(lw, lh, lf, 1))

    cpu_dropin = cpu_dropin_arm() if rank == 0 and world == 1 and not a.no_cpu and not a.no_sub else None

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("jpeg-encoder-decoder_b200")
    fr = frames_mod()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    enc = pkg.Encoder(local, a.frames_per_wave, a.lanes)
    dev = torch.device("cuda", local)
    n = a.batch
    first = rank * n                                   # each rank owns a contiguous range of frame indices

    # ---- synthetic inputs, generated on the device (host generators in frames.py define them; parity tested)
    d_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
    if a.kind == "natural":
        tile = torch.from_numpy(fr.tile_bgr(W, H)).to(dev)
        for i in range(n):
            dx, dy = fr.natural_shift(first + i, W, H)
            d_in[i] = torch.roll(tile, shifts=(dy, dx), dims=(0, 1))
    else:
        for i in range(n):
            d_in[i] = torch.from_numpy(fr.GENERATORS[a.kind](first + i, W, H)).to(dev)
    slot = SLOT if a.kind != "noise" else max(1024 * 1024, W * H)
    d_out = torch.zeros((n, slot), dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        enc.encode_batch_ptr(d_in.data_ptr(), n, W, H, W * H * 3, d_out.data_ptr(), slot, d_sizes.data_ptr(), stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    sizes = d_sizes.cpu().numpy().astype(np.int64)
    assert (sizes > 0).all(), "an output did not fit its slot"
    jpeg_bytes = int(sizes.sum())
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["synthetic"]

    def parity(kind, w, h, first_frame, nfr, out, szs):
        """SHA-256 of the frames of this batch that tests/golden/golden.json pins (outputs of the reference itself)."""
        res = []
        for f in (0, 1, 121, 1023, 7):
            key = f"{kind}_{w}x{h}_f{f}"
            i = f - first_frame
            if key in golden and 0 <= i < nfr:
                got = hashlib.sha256(out[i, : int(szs[i])].cpu().numpy().tobytes()).hexdigest()
                res.append({"frame": f, "bytes": int(szs[i]), "sha256_matches_reference": got == golden[key]["sha256"]})
        return res
    # bookkeeping only (outside every timed region): the global (rank, offset, size) table a caller needs to find frame f's
    # stream in rank r's output; the data path itself has no collective
    shard = importlib.import_module("jpeg-encoder-decoder_b200.sharding")
    table = shard.gather_tables(sizes, n * world, rank, world, device=dev)
    assert table.shape == (n * world, 3) and int(table[first:first + n, 2].sum()) == jpeg_bytes

    # ---- device-resident timing: EXACTLY K steps between two events on the launching stream
    enc.lib.jpegb200_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    enc.lib.jpegb200_set_timing(enc.ctx, 1)
    l0 = enc.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record(stream)
        for _ in range(a.steps):
            step()
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    launches = enc.launches - l0
    k1_ms, k1_n = get_timing(enc)
    parity_checked = parity(a.kind, W, H, first, n, d_out, d_sizes.cpu().numpy())      # the output of the last timed step
    assert all(p["sha256_matches_reference"] for p in parity_checked), parity_checked
    # ... and against the streams the CPU arm produced in this very run (the unmodified reference on this box's cores)
    live_parity = None
    if cpu is not None:
        szs = d_sizes.cpu().numpy()
        same, checked = 0, 0
        for f, nb, digest in cpu["streams"]:
            i = f - first
            if 0 <= i < n:
                checked += 1
                same += int(int(szs[i]) == nb and hashlib.sha256(d_out[i, : int(szs[i])].cpu().numpy().tobytes()).hexdigest() == digest)
        live_parity = {"frames_compared": checked, "byte_identical": same, "against": f"the CPU arm's own output in this run (kind: {cpu['kind']})"}
        assert same == checked, live_parity
    # the same kernel timed alone (one lane: no other kernel shares the SMs), as a second reading for the roofline object:
    # in the timed region above its launches are time-sliced with the high-priority kernels of the other lanes
    enc.configure(a.frames_per_wave, 1)
    nsub = min(n, 4 * a.frames_per_wave)
    for _ in range(2):
        enc.encode_batch_ptr(d_in.data_ptr(), nsub, W, H, W * H * 3, d_out.data_ptr(), slot, d_sizes.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    get_timing(enc)
    for _ in range(3):
        enc.encode_batch_ptr(d_in.data_ptr(), nsub, W, H, W * H * 3, d_out.data_ptr(), slot, d_sizes.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    iso_ms, iso_n = get_timing(enc)
    iso_frames = 3 * nsub / iso_n if iso_n else 0
    enc.configure(a.frames_per_wave, a.lanes)
    enc.lib.jpegb200_set_timing(enc.ctx, 0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / a.steps
    value = world * n * FRAME_MPIX / (ms_per_step / 1e3)

    # ---- end to end through the host-buffer C ABI (pinned host memory in and out)
    e2e = None
    if not a.no_e2e:
        h_in = torch.empty((n, H, W, 3), dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in)
        h_out = torch.empty((n, slot), dtype=torch.uint8, pin_memory=True)
        h_sizes = torch.zeros(n, dtype=torch.int32, pin_memory=True)

        enc.configure(a.e2e_frames_per_wave, a.e2e_lanes)

        def estep():
            enc.encode_batch_host_ptr(h_in.data_ptr(), n, W, H, h_out.data_ptr(), slot, h_sizes.data_ptr())

        for _ in range(2):
            estep()
        assert np.array_equal(h_sizes.numpy().astype(np.int64), sizes), "host path and device path disagree"
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            estep()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * n * FRAME_MPIX / (dt / a.steps), "unit": "Mpix/s", "h2d_bytes_per_step": world * n * W * H * 3,
               "d2h_bytes_per_step": world * (jpeg_bytes + 4 * n), "ms_per_step": 1000 * dt / a.steps,
               "api": "jpegb200_encode_batch_host (pinned host buffers)", "frames_per_wave": a.e2e_frames_per_wave, "lanes": a.e2e_lanes}
        if rank == 0 and world == 1 and not a.no_sub:
            # the same frames as the camera's packed RGB565 (2 bytes per pixel over PCIe, unpacked on the device: SURVEY.md 8f rank 3)
            b_, g_, r_ = d_in[..., 0], d_in[..., 1], d_in[..., 2]
            h565 = torch.empty((n, H, W, 2), dtype=torch.uint8, pin_memory=True)
            h565.copy_(torch.stack([(r_ & 0xF8) | (g_ >> 5), ((g_ << 3) & 0xE0) | (b_ >> 3)], dim=-1))
            del b_, g_, r_

            def pstep():
                enc._check(enc.lib.jpegb200_encode_batch_host_fmt(enc.ctx, ctypes.c_void_p(h565.data_ptr()), 1, n, W, H, ctypes.c_void_p(h_out.data_ptr()), slot,
                                                                  ctypes.c_void_p(h_sizes.data_ptr())))

            for _ in range(2):
                pstep()
            assert (h_sizes.numpy() > 0).all()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                pstep()
            torch.cuda.synchronize()
            dtp = (time.perf_counter() - t0) / a.steps
            e2e["packed_rgb565"] = {"value": n * FRAME_MPIX / dtp, "unit": "Mpix/s", "h2d_bytes_per_step": n * W * H * 2,
                                    "d2h_bytes_per_step": int(h_sizes.numpy().astype(np.int64).sum()) + 4 * n, "ms_per_step": 1000 * dtp,
                                    "api": "jpegb200_encode_batch_host_fmt(JPEGB200_FMT_RGB565): the same frames as the camera's packed RGB565, unpacked on the device"}
            del h565

    # ---- sub-records (N = 1 only): the other workloads BASELINE.json names, each measured like `value` but with fewer steps
    sub_records = []
    if rank == 0 and world == 1 and not a.no_sub:
        peaks0 = {}
        try:
            peaks0 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak0 = float(peaks0.get("hbm_gbs", 6650.0))
        del d_in, d_out
        if not a.no_e2e:
            del h_in, h_out
        torch.cuda.empty_cache()
        enc.configure(a.frames_per_wave, a.lanes)

        def device_record(kind, sw_, sh_, nfr, label):
            x = torch.empty((nfr, sh_, sw_, 3), dtype=torch.uint8, device=dev)
            if kind == "natural":
                tile_ = torch.from_numpy(fr.tile_bgr(sw_, sh_)).to(dev)
                for i in range(nfr):
                    dx, dy = fr.natural_shift(i, sw_, sh_)
                    x[i] = torch.roll(tile_, shifts=(dy, dx), dims=(0, 1))
            else:
                for i in range(nfr):
                    x[i] = torch.from_numpy(fr.GENERATORS[kind](i, sw_, sh_)).to(dev)
            slot_ = max(1024 * 1024, sw_ * sh_) if kind == "noise" else (SLOT if sw_ * sh_ <= 1920 * 1280 else 2 * 1024 * 1024)
            o = torch.zeros((nfr, slot_), dtype=torch.uint8, device=dev)
            z = torch.zeros(nfr, dtype=torch.int32, device=dev)

            def st_():
                enc.encode_batch_ptr(x.data_ptr(), nfr, sw_, sh_, sw_ * sh_ * 3, o.data_ptr(), slot_, z.data_ptr(), stream.cuda_stream)

            for _ in range(3):
                st_()
            torch.cuda.synchronize()
            zs = z.cpu().numpy().astype(np.int64)
            assert (zs > 0).all(), (label, "an output did not fit its slot")
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            ea.record(stream)
            for _ in range(reps):
                st_()
            eb.record(stream)
            torch.cuda.synchronize()
            msr = ea.elapsed_time(eb) / reps
            alg = sw_ * sh_ * 3 + zs.sum() / nfr
            rec = {"workload": label, "metric": f"Mpix/s JPEG encode ({sw_}x{sh_} 4:2:0 batch)", "value": nfr * sw_ * sh_ / 1e6 / (msr / 1e3), "unit": "Mpix/s",
                   "frames": nfr, "steps": reps, "warmup": 3, "ms_per_step": msr, "jpeg_bytes_per_frame": float(zs.sum() / nfr),
                   "whole_step": {"achieved": alg * nfr / (msr / 1e3) / 1e9, "frac": alg * nfr / (msr / 1e3) / 1e9 / peak0, "unit": "GB/s", "peak": peak0},
                   "parity_checked": parity(kind, sw_, sh_, 0, nfr, o, zs)}
            assert all(p_["sha256_matches_reference"] for p_ in rec["parity_checked"]), rec
            del x, o, z
            torch.cuda.empty_cache()
            return rec

        # config 2: the drop-in boundary itself - one frame at a time through the reference's three entry points as this library
        # exports them (host buffers in and out on every call: planes back after rgb_to_dct, tables after init_huffman, the
        # file after write_jpg), beside the reference C doing the same on one core
        def dropin_record():
            api = pkg.RefAPI()
            enc_golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["encode"]
            rows = []
            for name, img in (("sample_640x640_bgr", fr.sample_bgr("640")), ("tile_1920x1280_bgr", fr.tile_bgr(1920, 1280))):
                h_, w_, _ = img.shape
                got = api.encode(img)
                ok_ = hashlib.sha256(got["jpg"].tobytes()).hexdigest() == enc_golden[name]["sha256"]
                assert ok_, ("drop-in entry points differ from the golden digest", name)
                # timed like a C caller: every buffer allocated once (app_main's are static), the three calls back to back
                C_ = ctypes
                area = pkg.Area(0, 0, w_, h_)
                n_ = w_ * h_
                Y_, Cb_, Cr_ = np.zeros(n_, np.int16), np.zeros(n_ // 4, np.int16), np.zeros(n_ // 4, np.int16)
                jpg_ = np.zeros(3 * n_, np.uint8)
                luma_, chroma_ = (pkg.HuffCode * 2)(), (pkg.HuffCode * 2)()
                src_ = np.ascontiguousarray(img)
                u8 = lambda x: x.ctypes.data_as(C_.POINTER(C_.c_uint8))
                i16 = lambda x: x.ctypes.data_as(C_.POINTER(C_.c_int16))
                f_ = api._libc.fopen(b"/dev/null", b"wb")
                api.set_dims(w_, h_)

                def one():
                    api.lib.rgb_to_dct(u8(src_), i16(Y_), i16(Cb_), i16(Cr_), area)
                    api.lib.init_huffman(i16(Y_), i16(Cb_), i16(Cr_), area, luma_, chroma_)
                    return api.lib.write_jpg(f_, u8(jpg_), i16(Y_), i16(Cb_), i16(Cr_), area, luma_, chroma_)

                nb_ = one()
                assert hashlib.sha256(jpg_[:nb_].tobytes()).hexdigest() == enc_golden[name]["sha256"], name
                reps = 10
                t0 = time.perf_counter()
                for _ in range(reps):
                    one()
                dt = (time.perf_counter() - t0) / reps
                # the same with the caller's buffers page-locked once (jpegb200_pin_host): no staging copy
                for arr_ in (src_, Y_, Cb_, Cr_, jpg_):
                    pkg.pin_host(arr_)
                one()
                t0 = time.perf_counter()
                for _ in range(reps):
                    one()
                dt_reg = (time.perf_counter() - t0) / reps
                for arr_ in (src_, Y_, Cb_, Cr_, jpg_):
                    pkg.unpin_host(arr_)
                api._libc.fclose(f_)
                row = {"image": name, "ms_per_frame": 1000 * dt, "mpix_s": w_ * h_ / 1e6 / dt, "ms_per_frame_buffers_page_locked": 1000 * dt_reg,
                       "bytes": int(got["jpg"].size), "sha256_matches_reference": ok_}
                if cpu_dropin:
                    row["reference_ms_per_frame_one_core"] = 1000 * cpu_dropin[0][name]
                rows.append(row)
            return {"workload": "config 2: rgb_to_dct + init_huffman + write_jpg, the reference's own entry points exported by libjpegb200.so, one frame per call "
                                "sequence, host buffers (through the ctypes binding; planes, tables and file returned to the host after every call)",
                    "metric": "ms per frame", "value": rows[-1]["ms_per_frame"], "unit": "ms", "higher_is_better": False, "frames": rows,
                    "cpu_baseline": None if not cpu_dropin else {"kind": cpu_dropin[1], "cores": 1, "unit": "ms", "value": rows[-1].get("reference_ms_per_frame_one_core"),
                                                                 "sample": "the same two images, 3 passes each, timed inside C"}}

        sub_records.append(dropin_record())
        sub_records.append(device_record("noise", 1920, 1280, 128, "config 4, class 'noise': 128 x 1920x1280 (splitmix64 bytes: every block busy, 632 KB per frame)"))
        sub_records.append(device_record("ramp", 1920, 1280, 128, "config 4, class 'ramp': 128 x 1920x1280 (R=G=B=(x+y+f)&255: every pixel on an exact-integer colour boundary)"))
        sub_records.append(device_record("natural", 3840, 2160, 256, "config 5: 256 x 3840x2160 natural"))

        # decoding side: 1024 device-resident streams of 1920x1280 natural -> B,G,R frames; checked against the
        # CPU statement of the same arithmetic (oracle/oracle_decode.c), which is also the CPU arm of this record
        def decode_record(nfr=1024, sw_=1920, sh_=1280):
            x = torch.empty((nfr, sh_, sw_, 3), dtype=torch.uint8, device=dev)
            tile_ = torch.from_numpy(fr.tile_bgr(sw_, sh_)).to(dev)
            for i in range(nfr):
                dx, dy = fr.natural_shift(i, sw_, sh_)
                x[i] = torch.roll(tile_, shifts=(dy, dx), dims=(0, 1))
            o = torch.zeros((nfr, SLOT), dtype=torch.uint8, device=dev)
            z = torch.zeros(nfr, dtype=torch.int32, device=dev)
            enc.encode_batch_ptr(x.data_ptr(), nfr, sw_, sh_, sw_ * sh_ * 3, o.data_ptr(), SLOT, z.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            back = torch.zeros_like(x)
            status = torch.full((nfr,), 7, dtype=torch.int32, device=dev)

            def st_():
                enc.decode_batch_ptr(o.data_ptr(), SLOT, z.data_ptr(), nfr, sw_, sh_, back.data_ptr(), sw_ * sh_ * 3, 0, status.data_ptr(), stream.cuda_stream)

            for _ in range(2):
                st_()
            torch.cuda.synchronize()
            assert not status.any().item(), "a stream did not decode"
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            ea.record(stream)
            for _ in range(reps):
                st_()
            eb.record(stream)
            torch.cuda.synchronize()
            msr = ea.elapsed_time(eb) / reps
            import cpu_checkers
            orc = cpu_checkers.Oracle()
            zs = z.cpu().numpy().astype(np.uint32)
            checked = []
            for i in (0, nfr - 1):                                   # bit-exact against the CPU statement
                want = orc.decode(o[i, :int(zs[i])].cpu().numpy().tobytes(), sw_, sh_)
                assert want["rc"] == 0 and np.array_equal(want["bgr"], back[i].cpu().numpy()), ("decode differs from the oracle decoder", i)
                checked.append(i)
            sample = o[:2].cpu().numpy()
            cpu_s = orc.time_decode(np.ascontiguousarray(sample), zs[:2], sw_, sh_)
            err = back[:8].float() - x[:8].float()
            rec = {"workload": f"decoding side: {nfr} device-resident streams of {sw_}x{sh_} natural -> B,G,R frames (jpegb200_decode_batch)",
                   "metric": f"Mpix/s JPEG decode ({sw_}x{sh_} 4:2:0 batch)", "value": nfr * sw_ * sh_ / 1e6 / (msr / 1e3), "unit": "Mpix/s", "frames": nfr, "steps": reps,
                   "warmup": 2, "ms_per_step": msr, "frames_equal_to_oracle_decoder": checked,
                   "psnr_db_vs_input": float(10 * torch.log10(255.0 ** 2 / (err ** 2).mean())),
                   "cpu_baseline": {"value": 2 * sw_ * sh_ / 1e6 / cpu_s, "unit": "Mpix/s", "cores": 1, "kind": "port",
                                    "sample": "2 streams, oracle/oracle_decode.c single thread (the reference has no working decoder: pixel arithmetic restated from its stubs)"}}
            del x, o, z, back
            torch.cuda.empty_cache()
            return rec

        sub_records.append(decode_record())

        # config 3: the comparator loop (main.c:137-162) through jpegb200_compare_encode_batch with HOST frames: H2D of the frames,
        # subsample + compare + device-built region jobs + encode, D2H of the streams, all inside the timed region
        for (lw, lh, nf) in ((640, 640, 65), (1920, 1280, 33)):
            seq = fr.moving_sequence(nf, lw, lh, 5)
            hs = torch.empty(seq.shape, dtype=torch.uint8, pin_memory=True)
            hs.copy_(torch.from_numpy(seq))
            host = hs.numpy()
            enc.compare_encode(host[0], seed=True)
            counts, boxes, jpgs = enc.compare_encode_batch(host[1:], max_regions=16)
            enc.compare_encode(host[0], seed=True)
            enc.compare_encode_batch(host[1:], max_regions=16)
            torch.cuda.synchronize()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                enc.compare_encode(host[0], seed=True)
                enc.compare_encode_batch(host[1:], max_regions=16)
            dt = (time.perf_counter() - t0) / reps
            nreg = sum(1 for row in jpgs for j in row if j is not None)
            rec = {"workload": f"config 3: comparator loop, {nf - 1} frames of a {lw}x{lh} moving scene per call (seed + jpegb200_compare_encode_batch, host frames in, "
                               f"streams out)", "metric": "frames/s subsample+compare+encode regions+store", "value": (nf - 1) / dt, "unit": "frames/s",
                   "regions_per_frame": nreg / (nf - 1), "jpeg_bytes": int(sum(len(j) for row in jpgs for j in row if j is not None)),
                   "ms_per_call": 1000 * dt, "steps": reps, "warmup": 2}
            if (lw, lh) in cpu_loops:
                multi, single1 = cpu_loops[(lw, lh)]
                rec["cpu_baseline"] = {"value": multi["frames_per_s"], "unit": "frames/s", "cores": multi["cores"], "kind": multi["kind"], "sample": multi["sample"],
                                       "single_thread_frames_per_s": single1["frames_per_s"], "regions_per_frame": multi["regions_per_frame"]}
            sub_records.append(rec)
            del hs

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_per_frame = W * H * 3 + jpeg_bytes / n                    # compulsory read + compulsory write
        roof = None
        if k1_n:
            k1_avg_ms = k1_ms / k1_n
            frames_per_launch = n * a.steps / k1_n
            ach = alg_per_frame * frames_per_launch / (k1_avg_ms / 1e3) / 1e9
            traffic = None                     # dram__bytes_read + dram__bytes_write of the kernel from the committed ncu --set full capture
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))["dram_bytes_per_frame"] * frames_per_launch
            except Exception:
                pass
            # `achieved` / `frac` come from the kernel's EXCLUSIVE duration (single lane: nothing else on the SMs; CUDA events on its
            # stream).  Inside the timed region its launches are time-sliced with the high-priority pass-2 kernels of the other
            # lanes, so an event pair around a launch there measures elapsed time, not kernel time: reported under `in_region`.
            iso_avg_ms = iso_ms / iso_n if iso_n else k1_avg_ms
            iso_fpl = iso_frames if iso_n else frames_per_launch
            ach_iso = alg_per_frame * iso_fpl / (iso_avg_ms / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": "k_pixels_to_tokens", "achieved": ach_iso, "peak": peak, "unit": "GB/s", "frac": ach_iso / peak,
                    "traffic": None if traffic is None else traffic * iso_fpl / frames_per_launch,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s",
                    "duration": "exclusive: the kernel alone on the GPU (single lane), CUDA events on its stream, live in this run",
                    "avg_launch_ms": iso_avg_ms, "launches_timed": iso_n, "frames_per_launch": iso_fpl,
                    "algorithmic_bytes_per_frame": alg_per_frame,
                    "in_region": {"what": "event pairs around the kernel's launches inside the timed region: elapsed time under time-slicing "
                                          "with the other lanes' high-priority kernels, not an exclusive duration",
                                  "avg_launch_ms": k1_avg_ms, "launches_timed": k1_n, "frames_per_launch": frames_per_launch,
                                  "achieved": ach, "frac": ach / peak},
                    "whole_step": {"achieved": alg_per_frame * n / (ms_per_step / 1e3) / 1e9, "frac": alg_per_frame * n / (ms_per_step / 1e3) / 1e9 / peak}}
        line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
                "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "per_core_mpix_s", "single_thread_mpix_s")},
                "clocks": clk.summary(), "jpeg_bytes_per_frame": jpeg_bytes / n, "parity_checked": parity_checked, "parity_live": live_parity,
                "sub_records": sub_records}
        print(json.dumps(line))
    enc.close()
    if world > 1:
        dist.destroy_process_group()


def get_timing(enc):
    ms, cnt = ctypes.c_double(0), ctypes.c_uint64(0)
    enc.lib.jpegb200_get_timing.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    enc.lib.jpegb200_get_timing(enc.ctx, ctypes.byref(ms), ctypes.byref(cnt))
    return ms.value, cnt.value


if __name__ == "__main__":
    main()
