"""Where does the end-to-end (streaming) step spend its time?  Run by hand on a GPU box:
    PYTHONPATH=. python tests/diag_e2e.py            (DCA_STREAM_DIAG=1: copies + expansion only)
Prints ms/step of the streaming loop per host format, the resident step for comparison, the host-side
time of the Python loop (no GPU wait) and the raw H2D bandwidth at the per-step copy sizes."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dca_b200.engine import DeviceEngine
from dca_b200 import io as dio
from dca_b200.hostmem import pin_near_gpu, gpu_local_cpus
from bench import synth_on_device

dev = torch.device("cuda:0"); torch.cuda.set_device(0)
cells, genes, batch = int(os.environ.get("DIAG_CELLS", 10000)), int(os.environ.get("DIAG_GENES", 2000)), 4096
X, Y, sf, zf, gmean, gstd = synth_on_device(cells, genes, dev, 1234)
eng = DeviceEngine(genes, genes, (64, 32, 64), "zinb-conddisp", True, max_batch=batch, device=dev, seed=0)
nb = 2
counts = Y[: nb * batch].cpu().numpy()
print("gpu-local cpus:", sorted(gpu_local_cpus(0) or [])[:8], "... affinity now:", len(os.sched_getaffinity(0)))
sf_h = pin_near_gpu(sf[: nb * batch].cpu(), 0)
eng.set_input_transform(gmean, gstd, True, True)
formats = {"u16": pin_near_gpu(counts.astype(np.uint16), 0),
           "p8": dio.pack_counts(counts, 8, batch), "p4": dio.pack_counts(counts, 4, batch), "sp": dio.pack_counts(counts, "sparse", batch)}
loss_h = pin_near_gpu(torch.zeros(64, dtype=torch.float32), 0)
P = eng.n_params


def run(fmt, k, update=True, d2h=True):
    st = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.set_loss_ring(loss_h if d2h == 2 else None)
        eng.stream_begin(formats[fmt], sf_h, batch)
        e0.record()
        t0 = time.perf_counter()
        for i in range(k):
            eng.stream_step(i % nb, (i + 1) % nb if i + 1 < k else -1)
            if update: eng.apply_update(1e-3, 5.0, 1.0)
            if d2h == 1: loss_h[i % 64: i % 64 + 1].copy_(eng.grads[P:P + 1], non_blocking=True)
        host = time.perf_counter() - t0
        e1.record()
        eng.stream_end()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, host / k * 1e3


def resident(k):
    st = torch.cuda.Stream(dev)
    rows = torch.arange(batch, device=dev, dtype=torch.int32)
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(5):
            eng.train_step(X, Y, sf, rows=rows); eng.apply_update(1e-3, 5.0, 1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter()
        for _ in range(k):
            eng.train_step(X, Y, sf, rows=rows); eng.apply_update(1e-3, 5.0, 1.0)
        host = time.perf_counter() - t0
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, host / k * 1e3


print("mode DCA_STREAM_DIAG=%s" % os.environ.get("DCA_STREAM_DIAG", "0"))
if os.environ.get("DCA_STREAM_DIAG") == "2":        # device-side timeline of a few steps (printed by the library)
    for fmt in ("p4", "sp"):
        run(fmt, 5); print("timeline", fmt, flush=True); sys.stderr.flush()
        print("  -> %.3f ms/step" % run(fmt, 12, d2h=2)[0], flush=True)
    sys.exit(0)
print("resident step: gpu %.3f ms, host loop %.3f ms" % resident(60))
for fmt in formats:
    run(fmt, 5)
    for d2h in (1, 2, 0):        # 1: torch D2H copy per step, 2: loss ring (mapped host memory), 0: none
        g, h = run(fmt, 60, d2h=d2h)
        print("%-4s d2h=%d: gpu %.3f ms/step, host loop %.3f ms/step" % (fmt, d2h, g, h))
for mb, near in ((4, True), (16, True), (4, False), (16, False)):
    n = mb << 20
    hbuf = (pin_near_gpu(torch.empty(n, dtype=torch.uint8), 0) if near else torch.empty(n, dtype=torch.uint8).pin_memory()); dbuf = torch.empty(n, dtype=torch.uint8, device=dev)
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): dbuf.copy_(hbuf, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
    print("pure H2D %d MiB (near=%s): %.3f ms each -> %.1f GB/s" % (mb, near, e0.elapsed_time(e1) / 20, n / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9))
