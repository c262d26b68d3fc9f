import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import denseopticalflowsegmentation3d_b200 as d
W,H,B=1920,1080,8
c=d.Context(W,H,max_pairs=B)
fr=torch.empty((B+1,H,W,3),dtype=torch.uint8,device='cuda')
c.synth_frames_dev(1234,8,0,B+1,fr.data_ptr()); c.sync()
out=c.process(fr.cpu().numpy(), want_labels=False)
for s in out['stats'][:3]: print(dict(zip(s.dtype.names, s.tolist())))
