import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np, ria_b200
from ria_b200 import ofdm, sim, txsynth
dev=torch.device('cuda',0); ctx=ria_b200.Context(0)
cfg=ofdm.ModemConfig.high_throughput(ofdm.QAM64)
pool,_=txsynth.make_frame_pool(cfg,4,8,seed=3)
n=131072
rx=sim.awgn_batch(torch.from_numpy(pool).to(dev),n,25.0,seed=1,ctx=ctx)
dem=ofdm.OFDMDemodulator(cfg,ctx)
for label,c in (('rerun all (cfo +3.7 wrong)',3.7),('no rerun (cfo +0.2)',0.2)):
    cfo=torch.full((n,),c,device=dev); ph=torch.zeros(n,device=dev)
    for _ in range(2): dem.process_presynced_batch(rx,cfo,ph)
    torch.cuda.synchronize(); ctx.set_timing(True)
    for _ in range(3): dem.process_presynced_batch(rx,cfo,ph)
    torch.cuda.synchronize()
    print(label, {k: round(ctx.get_timing(k)[0]/3,3) for k in (11,12,13,1)})
    ctx.set_timing(False)
