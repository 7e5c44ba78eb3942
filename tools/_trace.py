import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth
B=1024; n=synth.points_per_frame(2); p=synth.params(2)
host=torch.empty((B,n,4),dtype=torch.float32).pin_memory(); synth.frames(2,0,B,out=host.numpy()); dev=host.cuda()
counts=np.full(B,n,np.int32)
op=ObstacleProcessor(p,n,max_batch=B)
for i in range(4):
    if i==3: print("---- traced call", file=sys.stderr)
    op.process_batch_raw(dev.data_ptr(), n, counts)
