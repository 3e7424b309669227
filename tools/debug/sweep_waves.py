#!/usr/bin/env python
"""Development aid: device-resident throughput of the headline workload over (frames_per_wave, lanes, tiles per warp of the
short-lived k_pixels_to_tokens CTAs).  usage: sweep_waves.py [--batch 512]   (re-executes itself per configuration, because
the tiles-per-warp override JPEGB200_TK_ITERS is read once per process)"""
import argparse, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def one(a):
    import torch
    pkg = importlib.import_module("jpeg-encoder-decoder_b200")
    fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
    W, H, n = 1920, 1280, a.batch
    dev = torch.device("cuda", 0)
    d_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
    tile = torch.from_numpy(fr.tile_bgr(W, H)).to(dev)
    for i in range(n):
        dx, dy = fr.natural_shift(i, W, H)
        d_in[i] = torch.roll(tile, shifts=(dy, dx), dims=(0, 1))
    slot = 512 * 1024
    d_out = torch.zeros((n, slot), dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    enc = pkg.Encoder(0, a.fpw, a.lanes)
    step = lambda: enc.encode_batch_ptr(d_in.data_ptr(), n, W, H, W * H * 3, d_out.data_ptr(), slot, d_sizes.data_ptr(), st.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(4):
            step()
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 4)
    print(json.dumps({"fpw": a.fpw, "lanes": a.lanes, "iters": os.environ.get("JPEGB200_TK_ITERS", "default"), "gpix_s": n * W * H / best / 1e6, "ms": best}))
    enc.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--fpw", type=int, default=64)
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        one(a)
    else:
        for fpw, lanes, iters in [(64, 3, None), (64, 3, 4), (64, 3, 16), (64, 3, 32), (64, 4, None), (64, 2, None), (128, 3, None), (32, 3, None), (32, 4, None), (32, 6, None),
                                  (128, 2, None), (16, 6, 4), (64, 4, 16), (128, 4, 16)]:
            env = dict(os.environ)
            if iters is not None:
                env["JPEGB200_TK_ITERS"] = str(iters)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--batch", str(a.batch), "--fpw", str(fpw), "--lanes", str(lanes)], env=env,
                               capture_output=True, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(line[-1] if line else f"FAILED {fpw} {lanes} {iters}: {r.stderr[-300:]}", flush=True)
