#!/usr/bin/env python
"""Per-stage CUDA-event timing of one device-resident batch (jpegb200_set_timing level 2).
Each launch is bracketed by its own event pair on its lane stream, so with several lanes the sums overlap
in wall-clock; run with --lanes 1 for a serial breakdown."""
import argparse, ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

STAGES = ["dct", "plane_masks", "symbol_stats", "build_huffman", "pack_tables", "block_bits", "scan", "pack", "count_ff", "layout", "stuff", "fix_blocks", "dc_fix", "run_bits"]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--frames-per-wave", type=int, default=8)
    ap.add_argument("--lanes", type=int, default=1)
    ap.add_argument("--kind", default="natural")
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1280)
    a = ap.parse_args()
    pkg = importlib.import_module("jpeg-encoder-decoder_b200")
    fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
    W, H, n = a.w, a.h, a.batch
    dev = torch.device("cuda", 0)
    enc = pkg.Encoder(0, a.frames_per_wave, a.lanes)
    d_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
    for i in range(n):
        d_in[i] = torch.from_numpy(fr.GENERATORS[a.kind](i, W, H)).to(dev)
    slot = 1024 * 1024
    d_out = torch.zeros((n, slot), dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    def step():
        enc.encode_batch_ptr(d_in.data_ptr(), n, W, H, W * H * 3, d_out.data_ptr(), slot, d_sizes.data_ptr(), st.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    L = enc.lib
    L.jpegb200_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.jpegb200_get_stage_timing.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    L.jpegb200_set_timing(enc.ctx, 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    reps = 3
    for _ in range(reps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * 16)(); cnt = (ctypes.c_uint64 * 16)()
    L.jpegb200_get_stage_timing(enc.ctx, ms, cnt)
    L.jpegb200_set_timing(enc.ctx, 0)
    total = e0.elapsed_time(e1)
    mpix = reps * n * W * H / 1e6
    print(f"{a.kind} {W}x{H} batch {n} G={a.frames_per_wave} lanes={a.lanes}: {total/reps:.3f} ms/step, {mpix/total/1e3*1e3:.0f} Mpix/s (with event overhead)")
    s = sum(ms)
    for i, name in enumerate(STAGES):
        if cnt[i]:
            print(f"  {name:14s} {ms[i]/reps:9.3f} ms/step  {100*ms[i]/s:5.1f}%  avg {1000*ms[i]/cnt[i]:8.1f} us x {cnt[i]//reps}  -> {mpix/ms[i]/1e3*1e3/1e3:8.1f} Gpix/s")
    print(f"  sum            {s/reps:9.3f} ms/step")
    print("  jpeg bytes/frame", float(d_sizes.float().mean()))
    enc.close()

if __name__ == "__main__":
    main()
