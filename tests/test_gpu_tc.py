"""tcgen05 kernels against same-rounding references (bf16-rounded operands, fp32/fp64 accumulate)."""
import ctypes as C
import numpy as np
import pytest
import torch

from oracle import dca_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _L():
    from dca_b200 import _lib
    return _lib


def _bf(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV, torch.bfloat16).contiguous()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("N,K", [(64, 64), (256, 64), (64, 128), (128, 256)])
def test_tc_probe_operand_layouts(a_mn, b_mn, N, K):
    """D = A.B for every (K-major | MN-major) operand combination the Dense kernels rely on."""
    L = _L(); lib = L.load()
    M = 128
    rng = np.random.default_rng(N * 7 + K + a_mn * 2 + b_mn)
    A = rng.normal(0, 1, (M, K)).astype(np.float32)
    B = rng.normal(0, 1, (K, N)).astype(np.float32)
    Ab, Bb = _bf(A), _bf(B)
    ref = (Ab.double() @ Bb.double()).cpu().numpy()
    a_store = Ab.t().contiguous() if a_mn else Ab             # MN-major: stored [K x M]
    b_store = Bb.contiguous() if b_mn else Bb.t().contiguous()  # K-major B: stored [N x K]; MN-major: [K x N]
    D = torch.full((M, N), float("nan"), device=DEV)
    def run(al, asb, bl, bsb):
        D.fill_(float("nan"))
        st = lib.dca_tc_probe(a_store.data_ptr(), a_store.shape[0], a_store.shape[1], b_store.data_ptr(), b_store.shape[0],
                              b_store.shape[1], a_mn, b_mn, M, N, K, al, asb, bl, bsb, D.data_ptr(), None)
        L.check(st, "dca_tc_probe")
        torch.cuda.synchronize()
        got = D.cpu().numpy()
        return float(np.nanmax(np.abs(got - ref)) / np.max(np.abs(ref))) if np.isfinite(got).all() else float("inf")

    err = run(-1, -1, -1, -1)
    if not err < 1e-5:
        # diagnose: which (LBO, SBO) convention would have worked for the MN-major operands?
        notes = []
        for al, asb in ((-1, -1), (1024, K * 128), (K * 128, 1024), (0, 1024), (1024, 1024)):
            for bl, bsb in ((-1, -1), (1024, K * 128), (K * 128, 1024), (0, 1024), (1024, 1024)):
                e = run(al, asb, bl, bsb)
                if e < 1e-5:
                    notes.append("a(lbo,sbo)=(%d,%d) b(lbo,sbo)=(%d,%d)" % (al, asb, bl, bsb))
        pytest.fail("a_mn=%d b_mn=%d N=%d K=%d rel err %.3g; working overrides: %s" % (a_mn, b_mn, N, K, err, notes))


@pytest.mark.parametrize("B,G,nh", [(128, 256, 3), (4096, 2000, 3), (300, 1000, 2), (77, 200, 1)])
def test_tc_heads_fwd(B, G, nh):
    L = _L(); lib = L.load()
    rng = np.random.default_rng(B + G)
    H = np.maximum(rng.normal(0, 1, (B, 64)), 0).astype(np.float32)
    W = [rng.normal(0, 0.25, (64, G)).astype(np.float32) for _ in range(nh)]
    b = [rng.normal(0, 0.5, G).astype(np.float32) for _ in range(nh)]
    sf = np.exp(rng.normal(0, 0.3, B)).astype(np.float32)
    kinds = [2, 3, 4][:nh] if nh == 3 else ([2, 4] if nh == 2 else [2])
    Hb = _bf(H)
    WhT = torch.cat([_bf(w).t().contiguous() for w in W], 0).contiguous()       # [nh*G x 64]
    bias = torch.as_tensor(np.concatenate(b)).to(DEV)
    sfd = torch.as_tensor(sf).to(DEV)
    outs = [torch.full((B, G), float("nan"), device=DEV) for _ in range(3)]
    karr = (C.c_int32 * 3)(*(kinds + [0] * (3 - nh)))
    st = lib.dca_tc_heads_fwd(Hb.data_ptr(), B, WhT.data_ptr(), bias.data_ptr(), G, nh, C.byref(karr), sfd.data_ptr(),
                              outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), G, None)
    L.check(st, "dca_tc_heads_fwd")
    torch.cuda.synchronize()
    Hd = Hb.double().cpu().numpy()
    for i, kind in enumerate(kinds):
        Wd = WhT[i * G:(i + 1) * G].double().cpu().numpy().T
        z = Hd @ Wd + b[i]
        ref = {2: lambda z: O.mean_act(z) * sf[:, None], 3: O.disp_act, 4: O.sigmoid}[kind](z)
        got = outs[i].cpu().numpy()
        assert np.all(np.isfinite(got)), "head %d has non-finite / unwritten outputs" % i
        np.testing.assert_allclose(got, ref, rtol=3e-5, atol=1e-7, err_msg="head kind %d" % kind)
