import os, sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import crop2seg_b200 as c2s
from crop2seg_b200 import _lib
from oracle import ltae4wtae_forward
from c2s_testlib import *
from golden_util import rel_err
for C,(b,t,h,w),lengths in [(128,(2,61,8,8),[61,27]),(128,(3,17,4,4),[17,0,1]),(64,(2,40,4,8),[40,33])]:
    kw=dict(in_channels=C,n_head=16,d_k=4,d_model=256)
    rng=np.random.RandomState(3)
    m=c2s.LTAE4WTAE(**kw); randomise(m,rng); m=m.cuda().eval()
    for zp in (False,True):
        m.assume_zero_padded=zp
        x,pos,pad=synth_inputs(rng,b,t,C,h,w,lengths)
        ref=ltae4wtae_forward(oracle_config('ltae4wtae',kw),oracle_params(m),bf16_round(x),pos,pad)
        os.environ.pop('C2S_LTAE_TC',None)
        with torch.no_grad(): a0=m(to_dev(x,dtype=torch.bfloat16),batch_positions=to_dev(pos),pad_mask=to_dev(pad))
        k0=_lib.last_kernel()
        os.environ['C2S_LTAE_TC']='1'
        with torch.no_grad(): a1=m(to_dev(x,dtype=torch.bfloat16),batch_positions=to_dev(pos),pad_mask=to_dev(pad))
        torch.cuda.synchronize()
        k1=_lib.last_kernel()
        print(C,(b,t,h,w),zp,k0,rel_err(a0.cpu().numpy(),ref),k1,rel_err(a1.cpu().numpy(),ref), float(a1.sum(2).sub(1).abs().max()))
