import sys, importlib, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
pkg=importlib.import_module("jpeg-encoder-decoder_b200"); fr=importlib.import_module("jpeg-encoder-decoder_b200.frames")
enc=pkg.Encoder(0,64,3)
dev=torch.device("cuda",0)
nfr=64; W,H=1920,1280
x=torch.empty((nfr,H,W,3),dtype=torch.uint8,device=dev)
tile=torch.from_numpy(fr.tile_bgr(W,H)).to(dev)
for i in range(nfr):
    dx,dy=fr.natural_shift(i,W,H); x[i]=torch.roll(tile,shifts=(dy,dx),dims=(0,1))
slot=512*1024
o=torch.zeros((nfr,slot),dtype=torch.uint8,device=dev); z=torch.zeros(nfr,dtype=torch.int32,device=dev)
st=torch.cuda.current_stream()
enc.encode_batch_ptr(x.data_ptr(),nfr,W,H,W*H*3,o.data_ptr(),slot,z.data_ptr(),st.cuda_stream)
back=torch.zeros_like(x); status=torch.zeros(nfr,dtype=torch.int32,device=dev)
enc.decode_batch_ptr(o.data_ptr(),slot,z.data_ptr(),nfr,W,H,back.data_ptr(),W*H*3,0,status.data_ptr(),st.cuda_stream)
torch.cuda.synchronize()
