import importlib, time, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
pkg = importlib.import_module("jpeg-encoder-decoder_b200"); fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
W,H,n=1920,1280,256
tile=fr.tile_bgr(W,H)
pin=torch.empty((n,H,W,3),dtype=torch.uint8,pin_memory=True)
for i in range(n): pin[i]=torch.from_numpy(np.roll(tile,(16*(i%80),16*(i%120)),axis=(0,1)))
page=pin.numpy().copy()
slot=512*1024
enc=pkg.Encoder(0,16,3)
out_pin=torch.empty((n,slot),dtype=torch.uint8,pin_memory=True).numpy(); sizes=np.zeros(n,np.uint32)
out_page=np.empty((n,slot),np.uint8)
def run(src,dst):
    enc.encode_batch_host(src,slot,out=dst,sizes=sizes)
reg=page.copy(); out_reg=np.empty((n,slot),np.uint8); pkg.pin_host(reg); pkg.pin_host(out_reg)
for name,src,dst in (("pinned",pin.numpy(),out_pin),("pageable",page,out_page),("registered (jpegb200_pin_host)",reg,out_reg),("pageable",page,out_page)):
    run(src,dst)
    t0=time.perf_counter(); 
    for _ in range(3): run(src,dst)
    dt=(time.perf_counter()-t0)/3
    print(name, "%.1f ms  %.2f Gpix/s  %.1f GB/s in"%(1000*dt, n*W*H/dt/1e9, n*W*H*3/dt/1e9))

# the comparator loop (jpegb200_compare_encode_batch) with pinned and with pageable frames
seq = fr.moving_sequence(33, W, H, 5)
hs = torch.empty(seq.shape, dtype=torch.uint8, pin_memory=True); hs.copy_(torch.from_numpy(seq))
for name, host in (("pinned", hs.numpy()), ("pageable", seq), ("pinned", hs.numpy()), ("pageable", seq)):
    enc.compare_encode(host[0], seed=True); enc.compare_encode_batch(host[1:], max_regions=16)
    t0 = time.perf_counter()
    for _ in range(3):
        enc.compare_encode(host[0], seed=True); enc.compare_encode_batch(host[1:], max_regions=16)
    dt = (time.perf_counter() - t0) / 3
    print("comparator loop, 32 frames", name, "%.2f ms per call  %.0f frames/s" % (1000 * dt, 32 / dt))

# decoding side with host buffers: 64 streams -> 64 frames (472 MB) into pageable memory
jpgs = enc.encode_frames(page[:64])
import ctypes as C
slot_d = (max(len(j) for j in jpgs) + 15) & ~15
buf = np.zeros((64, slot_d), np.uint8); szs = np.zeros(64, np.uint32)
for i, j in enumerate(jpgs):
    buf[i, :len(j)] = np.frombuffer(bytes(j), np.uint8); szs[i] = len(j)
bgr = np.zeros((64, H, W, 3), np.uint8); status = np.zeros(64, np.int32)
def dec():
    enc._check(enc.lib.jpegb200_decode_batch_host(enc.ctx, buf.ctypes.data, slot_d, szs.ctypes.data, 64, W, H, bgr.ctypes.data, None, status.ctypes.data))
dec()
t0 = time.perf_counter(); dec(); dt = time.perf_counter() - t0
print("decode 64 streams to pageable frames: %.1f ms (%.2f Gpix/s), status ok %s" % (1000 * dt, 64 * W * H / dt / 1e9, bool((np.asarray(status) == 0).all())))
