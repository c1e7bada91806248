import torch, time
n=1_000_000_000
h_in=torch.empty(n,dtype=torch.uint8).pin_memory(); h_out=torch.empty(n,dtype=torch.uint8).pin_memory()
d_in=torch.empty(n,dtype=torch.uint8,device='cuda'); d_out=torch.empty(n,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def run(h2d,d2h,chunks=1):
    torch.cuda.synchronize(); t=time.perf_counter()
    c=n//chunks
    for i in range(chunks):
        sl=slice(i*c,(i+1)*c)
        if h2d:
            with torch.cuda.stream(s1): d_in[sl].copy_(h_in[sl],non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[sl].copy_(d_out[sl],non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter()-t
for name,a,b,ch in (('h2d only',1,0,1),('d2h only',0,1,1),('both',1,1,1),('both 32 chunks',1,1,32)):
    run(a,b,ch); ts=[run(a,b,ch) for _ in range(3)]
    print(name,'%.1f ms  -> %.1f GB/s per direction'%(min(ts)*1e3, n/min(ts)/1e9))
