"""Does the streamed step slow down when the host runs far ahead of the device?  ms/step of the streaming loop for
several step counts, with and without a host-side wait every `sync_every` steps (run by hand on a GPU box)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dca_b200.engine import DeviceEngine
from dca_b200 import io as dio
from dca_b200.hostmem import pin_near_gpu
from bench import synth_on_device

dev = torch.device("cuda:0"); torch.cuda.set_device(0)
cells, genes, batch = 10000, 2000, 4096
X, Y, sf, zf, gmean, gstd = synth_on_device(cells, genes, dev, 1234)
eng = DeviceEngine(genes, genes, (64, 32, 64), "zinb-conddisp", True, max_batch=batch, device=dev, seed=0)
nb = 2
pc = dio.pack_counts(Y[: nb * batch].cpu().numpy(), 4, batch)
sf_h = pin_near_gpu(sf[: nb * batch].cpu(), 0)
ring = pin_near_gpu(torch.zeros(64, dtype=torch.float32), 0)
eng.set_input_transform(gmean, gstd, True, True)
st = torch.cuda.Stream(dev)


def run(k, sync_every=0):
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        eng.set_loss_ring(ring)
        eng.stream_begin(pc, sf_h, batch)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evs = []
        e0.record()
        for i in range(k):
            eng.stream_step(i % nb, (i + 1) % nb if i + 1 < k else -1)
            eng.apply_update(1e-3, 5.0, 1.0)
            if sync_every:
                ev = torch.cuda.Event(); ev.record(); evs.append(ev)
                if len(evs) > sync_every: evs.pop(0).synchronize()      # host stays at most `sync_every` steps ahead
        e1.record()
        eng.stream_end()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def resident(k):
    rows = torch.arange(batch, device=dev, dtype=torch.int32)
    with torch.cuda.stream(st):
        for _ in range(k):
            eng.train_step(X, Y, sf, rows=rows); eng.apply_update(1e-3, 5.0, 1.0)
    torch.cuda.synchronize()


run(8)
print("fresh engine:        30 steps %.3f ms/step, 60 steps %.3f" % (run(30), run(60)), flush=True)
resident(40)
print("after 40 resident:   30 steps %.3f ms/step, 60 steps %.3f" % (run(30), run(60)), flush=True)
eng.profile(True); resident(20); eng.profile_read(); eng.profile(False)
print("after profiled pass: 30 steps %.3f ms/step, 60 steps %.3f" % (run(30), run(60)), flush=True)
time.sleep(0.3)
print("after 0.3 s idle:    30 steps %.3f ms/step, 60 steps %.3f" % (run(30), run(60)), flush=True)
sys.exit(0)
for k in (6, 12, 30, 60, 120):
    print("steps %4d: free-running %.3f ms/step | host <= 2 steps ahead %.3f | <= 4 ahead %.3f" % (k, run(k), run(k, 2), run(k, 4)), flush=True)
