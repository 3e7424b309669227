#!/usr/bin/env python
"""Development aid: synchronisation statistics of the sub-sequence decoder on a batch of 1920x1280 frames of each content class."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
pkg = importlib.import_module("jpeg-encoder-decoder_b200"); fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
enc = pkg.Encoder(0)
dev = torch.device("cuda", 0)
W, H, n = 1920, 1280, 64
for kind in ("natural", "noise", "ramp"):
    x = torch.stack([torch.from_numpy(fr.GENERATORS[kind](i, W, H)) for i in range(n)]).to(dev)
    slot = 1024 * 1024
    o = torch.zeros((n, slot), dtype=torch.uint8, device=dev); z = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    enc.encode_batch_ptr(x.data_ptr(), n, W, H, W * H * 3, o.data_ptr(), slot, z.data_ptr(), st.cuda_stream)
    back = torch.zeros_like(x); status = torch.zeros(n, dtype=torch.int32, device=dev)
    f = lambda: enc.decode_batch_ptr(o.data_ptr(), slot, z.data_ptr(), n, W, H, back.data_ptr(), W * H * 3, 0, status.data_ptr(), st.cuda_stream)
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); f(); b.record(st); torch.cuda.synchronize()
    par = back.clone()
    stats = enc.decode_stats()
    enc.set_decode_sequential(True); f(); torch.cuda.synchronize(); enc.set_decode_sequential(False)
    print(kind, "ms", round(a.elapsed_time(b), 2), stats, "equal to the warp-per-scan decoder:", bool((par == back).all()), "status", int(status.abs().sum()))
