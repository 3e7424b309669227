#!/usr/bin/env python
"""Development aid: time library variants (tools/debug/build_variant.sh) back to back on one GPU and check that they all
produce the same bytes.  For every library: device-resident batch in the bench configuration (3 lanes, 64 frames per wave)
and the dominant kernel alone (1 lane, CUDA events around its launches: jpegb200_set_timing level 1).
usage: time_variants.py [--batch 256] [--kind natural] lib1.so lib2.so ...   (re-executes itself once per library)"""
import argparse, ctypes, hashlib, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def one(a):
    import numpy as np
    import torch
    pkg = importlib.import_module("jpeg-encoder-decoder_b200")
    fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
    W, H, n = a.w, a.h, a.batch
    dev = torch.device("cuda", 0)
    d_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
    if a.kind == "natural":
        tile = torch.from_numpy(fr.tile_bgr(W, H)).to(dev)
        for i in range(n):
            dx, dy = fr.natural_shift(i, W, H)
            d_in[i] = torch.roll(tile, shifts=(dy, dx), dims=(0, 1))
    else:
        for i in range(n):
            d_in[i] = torch.from_numpy(fr.GENERATORS[a.kind](i, W, H)).to(dev)
    slot = 1024 * 1024
    d_out = torch.zeros((n, slot), dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    res = {"lib": os.path.basename(pkg.LIB_PATH)}
    enc = pkg.Encoder(0, a.frames_per_wave, a.lanes)
    L = enc.lib
    L.jpegb200_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.jpegb200_get_timing.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]

    def step(m=n):
        enc.encode_batch_ptr(d_in.data_ptr(), m, W, H, W * H * 3, d_out.data_ptr(), slot, d_sizes.data_ptr(), st.cuda_stream)

    def k1():
        ms, cnt = ctypes.c_double(0), ctypes.c_uint64(0)
        L.jpegb200_get_timing(enc.ctx, ctypes.byref(ms), ctypes.byref(cnt))
        return ms.value, cnt.value

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    sizes = d_sizes.cpu().numpy()
    out = d_out.cpu().numpy()
    h = hashlib.sha256()
    h.update(sizes.tobytes())
    for i in range(n):
        h.update(out[i, :sizes[i]].tobytes())
    res["digest"] = h.hexdigest()[:16]
    res["bytes_per_frame"] = float(sizes.mean())
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.steps):
            step()
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    res["ms_per_step"] = best
    res["gpix_s"] = n * W * H / best / 1e6
    # dominant kernel alone
    enc.configure(a.frames_per_wave, 1)
    m = min(n, 4 * a.frames_per_wave)
    for _ in range(2):
        step(m)
    torch.cuda.synchronize()
    L.jpegb200_set_timing(enc.ctx, 1)
    k1()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        step(m)
    e1.record(st)
    torch.cuda.synchronize()
    ms, cnt = k1()
    L.jpegb200_set_timing(enc.ctx, 0)
    res["k1_us_per_wave"] = 1000 * ms / cnt if cnt else None
    res["lane1_gpix_s"] = 5 * m * W * H / e0.elapsed_time(e1) / 1e6
    enc.close()
    print(json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--kind", default="natural")
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1280)
    ap.add_argument("--frames-per-wave", type=int, default=64)
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        return one(a)
    first = None
    for lib in a.libs or [os.path.join(ROOT, "jpeg-encoder-decoder_b200", "libjpegb200.so")]:
        env = dict(os.environ, JPEGB200_LIB=os.path.abspath(lib))
        args = [sys.executable, os.path.abspath(__file__), "--child"] + [x for x in sys.argv[1:] if not x.endswith(".so")]
        r = subprocess.run(args, env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(f"{os.path.basename(lib):28s} FAILED rc={r.returncode} {r.stderr[-400:]}")
            continue
        d = json.loads(line[-1])
        first = first or d["digest"]
        print(f"{d['lib']:28s} {d['gpix_s']:8.1f} Gpix/s ({d['ms_per_step']:.3f} ms/step)  k1 alone {d['k1_us_per_wave']:7.1f} us/wave  1-lane {d['lane1_gpix_s']:6.1f} Gpix/s"
              f"  {d['bytes_per_frame']:.1f} B/frame  digest {d['digest']} {'OK' if d['digest'] == first else 'DIFFERS'}", flush=True)


if __name__ == "__main__":
    main()
