"""Fused head/loss/backward kernel vs the three-kernel path: per-tensor gradient differences (run by hand on a GPU)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import dca_oracle as O
from tests.util import synth_counts
from dca_b200.engine import DeviceEngine
from dca_b200 import _lib

DEV = "cuda:0"
def _t(a, dtype=torch.float32): return torch.as_tensor(np.ascontiguousarray(a)).to(DEV, dtype)
shapes = [(128, 64), (256, 264), (300, 2000)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for B, G in shapes:
    N = B + 37
    Y = synth_counts(N, G, 51); X, sf = O.normalize_inputs(Y)
    rows = torch.as_tensor(np.random.default_rng(3).permutation(N)[:B].astype(np.int32)).to(DEV)
    es = []
    for fused in (1, 0):
        _lib.set_tunable("fused_heads", fused)
        es.append(DeviceEngine(G, G, (64, 32, 64), "zinb-conddisp", max_batch=B, seed=9, gemm_path="tcgen05", ridge=0.01))
    e1, e2 = es
    Xd, Yd, sfd = _t(X), _t(Y), _t(sf)
    for step in range(3):
        for e in es: e.train_step(Xd, Yd, sfd, rows=rows)
        torch.cuda.synchronize()
        print("shape", B, G, "step", step, "loss fused %.7f unfused %.7f" % (e1.read_loss(), e2.read_loss()))
        G1, G2 = e1.grads.cpu().numpy(), e2.grads.cpu().numpy()
        for name, off, r, c in e2.param_info:
            g1, g2 = G1[off: off + r * c].reshape(r, c), G2[off: off + r * c].reshape(r, c)
            d = np.abs(g1 - g2)
            if step > 0 and d.max() <= 2e-4 * np.abs(g2).max() + 1e-9: continue
            line = "  %-18s %5dx%-5d max|g| %.3e max|diff| %.3e" % (name, r, c, np.abs(g2).max(), d.max())
            if d.max() > 2e-4 * np.abs(g2).max() + 1e-9 and c == G:
                blocks = [d[:, k:k + 64].max() / (np.abs(g2).max() + 1e-30) for k in range(0, G, 64)]
                line += "  rel diff per 64-gene block: " + " ".join("%.1e" % b for b in blocks[:32])
                rws = [d[k:k + 8, :].max() / (np.abs(g2).max() + 1e-30) for k in range(0, r, 8)]
                line += " | per 8-row block: " + " ".join("%.1e" % b for b in rws[:8])
            print(line[:900])
        for e in es: e.apply_update(1e-3, 5.0)
        torch.cuda.synchronize()
        e1.params.copy_(e2.params); e1.rms.copy_(e2.rms); e1.bn_state.copy_(e2.bn_state); e1.params_changed()
