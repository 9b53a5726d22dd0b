"""Launched under torchrun on >= 2 GPUs (not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_dist_smoke.py
Trains the accelerated path data-parallel through the public API and checks that (a) the replicas stay
bit-identical, (b) the loss history is identical on every rank and decreases."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import synth_counts  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from dca_b200.anndata_lite import AnnData
    from dca_b200.api import dca
    adata = AnnData(synth_counts(2000, 256, 0))
    out, net = dca(adata, ae_type="zinb-conddisp", epochs=4, batch_size=128, copy=True, return_model=True, return_info=True,
                   verbose=False)
    w = net.engine.params.clone()
    ref = w.clone(); dist.broadcast(ref, 0)
    same = bool(torch.equal(w, ref))
    h = out.uns["dca_loss_history"]
    hist = torch.tensor(h["loss"] + h["val_loss"], dtype=torch.float64, device="cuda")
    href = hist.clone(); dist.broadcast(href, 0)
    ok = same and bool(torch.equal(hist, href)) and h["loss"][-1] < h["loss"][0] and np.all(np.isfinite(out.X))
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist smoke: world=%d replicas_identical=%s loss %s -> %s : %s" % (world, same, h["loss"][0], h["loss"][-1],
                                                                              "OK" if flag.item() == 1.0 else "FAILED"))
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
