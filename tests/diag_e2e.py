import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dca_b200.engine import DeviceEngine
from bench import synth_on_device
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
cells, genes, batch = 10000, 2000, 4096
X, Y, sf, zf, gmean, gstd = synth_on_device(cells, genes, dev, 1234)
eng = DeviceEngine(genes, genes, (64, 32, 64), "zinb-conddisp", True, max_batch=batch, device=dev, seed=0)
nb = 2
cnt_h = torch.from_numpy(Y[: nb * batch].cpu().numpy().astype(np.uint16)).pin_memory(); sf_h = sf[: nb * batch].cpu().pin_memory()
eng.set_input_transform(gmean, gstd, True, True)
def run(k, update=True, side=False):
    st = torch.cuda.Stream(dev) if side else torch.cuda.current_stream(dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.stream_begin(cnt_h, sf_h, batch)
        e0.record()
        for i in range(k):
            eng.stream_step(i % nb, (i + 1) % nb if i + 1 < k else -1)
            if update: eng.apply_update(1e-3, 5.0, 1.0)
        e1.record()
        eng.stream_end()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
for side in (False, True):
    run(5, side=side)
    print("mode", os.environ.get("DCA_STREAM_DIAG", "0"), "side_stream", side, "ms/step %.3f" % run(40, side=side), eng.info())
# pure H2D of 16.4 MB chunks
d = torch.empty(batch * genes, dtype=torch.uint16, device=dev)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): d.copy_(cnt_h[(i % nb) * batch:(i % nb + 1) * batch].view(-1), non_blocking=True)
    e1.record(); torch.cuda.synchronize()
print("pure H2D 16.4MB: %.3f ms each -> %.1f GB/s" % (e0.elapsed_time(e1) / 20, batch * genes * 2 / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9))
