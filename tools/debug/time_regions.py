#!/usr/bin/env python
"""Development aid: CUDA-event time of jpegb200_encode_regions on the four golden comparator regions of the 640x640 pair
(all four have 3*x not a multiple of 16: the unaligned loader) and on the same regions of a
1920x1280 frame.  usage: time_regions.py lib1.so lib2.so ... (re-executes itself once per library)"""
import importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def one():
    import numpy as np, torch, hashlib
    pkg = importlib.import_module("jpeg-encoder-decoder_b200")
    fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
    enc = pkg.Encoder(0)
    res = {"lib": os.path.basename(pkg.LIB_PATH)}
    for name, img, areas in (("640", fr.sample_bgr("640_diffs"), [(2, 36, 112, 432), (358, 66, 256, 336), (406, 476, 192, 160), (146, 412, 176, 144)]),
                             ("1920", fr.natural_frame(7), [(2, 36, 112, 432), (358, 66, 256, 336), (1206, 476, 592, 560), (146, 812, 976, 444 // 16 * 16)])):
        H, W, _ = img.shape
        d = torch.from_numpy(img).cuda()
        slot = 1 << 20
        d_out = torch.zeros((len(areas), slot), dtype=torch.uint8, device="cuda")
        d_sizes = torch.zeros(len(areas), dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream()
        f = lambda: enc.encode_regions_ptr(d.data_ptr(), W, H, areas, d_out.data_ptr(), slot, d_sizes.data_ptr(), st.cuda_stream)
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        sizes = d_sizes.cpu().numpy(); out = d_out.cpu().numpy()
        h = hashlib.sha256()
        for i in range(len(areas)):
            h.update(out[i, :sizes[i]].tobytes())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 100
        e0.record(st)
        for _ in range(n):
            f()
        e1.record(st)
        torch.cuda.synchronize()
        res[name] = {"us_per_call": 1000 * e0.elapsed_time(e1) / n, "digest": h.hexdigest()[:12], "mpix": sum(a[2] * a[3] for a in areas) / 1e6}
    enc.close()
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        one()
    else:
        for lib in sys.argv[1:]:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, JPEGB200_LIB=os.path.abspath(lib)), capture_output=True, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(line[-1] if line else f"{lib} FAILED {r.stderr[-500:]}", flush=True)
