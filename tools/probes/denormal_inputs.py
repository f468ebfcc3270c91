import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'cmsis-dsp_b200/python')
import cmsisdsp_b200 as cd
from oracle_lib import oracle
def relrms(a,b):
    a=a.astype(np.float64); b=b.astype(np.float64)
    d=np.sqrt(((a-b)**2).sum()); n=np.sqrt((b**2).sum()); return d/n if n>0 else d
rng=np.random.default_rng(1)
for scale in (1e30, 1e-30, 1e-36, 1e-38, 1e-41):
    for N in (64, 256, 1024, 4096):
        x=(rng.standard_normal((8,2*N))*scale).astype(np.float32)
        for ifft in (0,1):
            want=oracle().cfft("f32",N,x,ifft,1); got=cd.cfft_batch("f32",N,x,ifft,1)
            r=rng.standard_normal((8,N)).astype(np.float32)*np.float32(scale)
            wr=oracle().rfft(N,r,0); gr=cd.rfft_batch(N,r,0)
            print(f"scale {scale:g} N {N} ifft {ifft}: cfft relrms {relrms(got,want):.2e}  rfft fwd relrms {relrms(gr,wr):.2e}  zeros in got {int((got==0).sum())} want {int((want==0).sum())}")
