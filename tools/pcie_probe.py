import torch, time
n=18_464_768//4
h_in=torch.empty(n,dtype=torch.float32).pin_memory(); h_out=torch.empty(n,dtype=torch.float32).pin_memory()
d_in=torch.empty(n,dtype=torch.float32,device='cuda'); d_out=torch.ones(n,dtype=torch.float32,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(fn,reps=50):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps*1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in,non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out,non_blocking=True)
def both():
    h2d(); d2h()
for f in (h2d,d2h,both): t(f,5)
print("h2d ms",t(h2d),"GB/s",n*4/t(h2d)/1e6)
print("d2h ms",t(d2h),"GB/s",n*4/t(d2h)/1e6)
print("both ms",t(both))
# chunked both
def chunked(k):
    c=n//k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c],non_blocking=True)
            with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c],non_blocking=True)
    return f
for k in (2,4,8): print("both chunked",k,t(chunked(k)))
