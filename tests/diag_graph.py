import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dca_b200.engine import DeviceEngine
from tests.util import synth_counts
from oracle import dca_oracle as O
dev = torch.device("cuda:0")
B, G = 512, 512
Y = synth_counts(B, G, 0); X, sf = O.normalize_inputs(Y)
eng = DeviceEngine(G, G, (64, 32, 64), "zinb-conddisp", True, max_batch=B, device=dev, seed=0)
Xd, Yd, sfd = (torch.from_numpy(a).to(dev) for a in (X, Y, sf))
st = torch.cuda.Stream(dev)
torch.cuda.synchronize()
with torch.cuda.stream(st):
    for i in range(4):
        eng.train_step(Xd, Yd, sfd); eng.apply_update(1e-3, 5.0)
        print(i, eng.read_loss(), eng.info()["step_graphs"])
