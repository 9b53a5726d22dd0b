"""Host-side logic that needs no GPU: C-ABI export table, data preparation, CLI, callbacks."""
import ctypes as C
import os
import re
import numpy as np
import pandas as pd
import pytest

from oracle import dca_oracle as O
from tests.util import synth_counts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dca_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "dca_b200.h")).read()
    declared = set(re.findall(r"\b(dca_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dca_handle"}
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), "libdca_b200.so does not export %s" % name
    assert declared == set(_lib.PROTOTYPES), (declared ^ set(_lib.PROTOTYPES))
    assert lib.dca_version() == 100


def test_config_struct_matches_header_and_errors_are_reported():
    from dca_b200 import _lib
    lib = _lib.load()
    cfg = _lib.default_config()
    assert cfg.struct_bytes == C.sizeof(_lib.Config)
    assert list(cfg.hidden)[:3] == [64, 32, 64] and abs(cfg.bn_momentum - 0.99) < 1e-7
    n = C.c_size_t()
    cfg.n_in = 0
    assert lib.dca_arena_bytes(C.byref(cfg), C.byref(n)) == -1
    assert b"n_in" in lib.dca_last_error()
    cfg.n_in = cfg.n_out = 2000; cfg.max_batch = 4096
    assert lib.dca_arena_bytes(C.byref(cfg), C.byref(n)) == 0 and n.value > 3 * 4096 * 2000 * 4
    cfg.ae_type = 9                                    # nb-fork: (64, 32, 64) has the one decoder layer it needs
    assert lib.dca_arena_bytes(C.byref(cfg), C.byref(n)) == 0
    cfg.n_hidden = 5; cfg.hidden[3] = 16; cfg.hidden[4] = 8   # ... two decoder layers after 'center': unsupported, and said so
    assert lib.dca_arena_bytes(C.byref(cfg), C.byref(n)) == -3 and b"decoder layer" in lib.dca_last_error()
    cfg.n_hidden = 3; cfg.ae_type = 11
    assert lib.dca_arena_bytes(C.byref(cfg), C.byref(n)) == -3


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dca_b200 import _lib
    from dca_b200.engine import DeviceEngine
    with pytest.raises(_lib.DcaError):
        DeviceEngine(10, 10, (4, 2, 4), "zinb")


def test_normalize_matches_oracle_restatement():
    from dca_b200.anndata_lite import AnnData
    from dca_b200 import io
    Y = synth_counts(60, 25, 1)
    ad = AnnData(Y.copy())
    ad = io.read_dataset(ad, check_counts=True)
    ad = io.normalize(ad, filter_min_counts=False)
    X, sf = O.normalize_inputs(Y)
    np.testing.assert_allclose(ad.X, X, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ad.obs["size_factors"].values, sf, rtol=1e-6)
    np.testing.assert_array_equal(ad.raw.X, Y)
    assert abs(ad.X.mean(0)).max() < 1e-5 and abs(ad.X.std(0, ddof=1) - 1).max() < 1e-4
    assert set(ad.obs["dca_split"]) == {"train"}


def test_read_dataset_checks_counts_and_splits(tmp_path):
    from dca_b200.anndata_lite import AnnData
    from dca_b200 import io
    Y = synth_counts(40, 12, 2)
    bad = AnnData(Y + 0.5)
    with pytest.raises(AssertionError, match="unnormalized count data"):
        io.read_dataset(bad)
    ad = io.read_dataset(AnnData(Y), test_split=True)
    assert (ad.obs["dca_split"] == "test").sum() == 4
    # gene x cell TSV, transposed on read like the reference CLI does
    df = pd.DataFrame(Y.T.astype(int), index=["g%d" % i for i in range(12)], columns=["c%d" % i for i in range(40)])
    p = tmp_path / "counts.tsv"; df.to_csv(p, sep="\t")
    ad2 = io.read_dataset(str(p), transpose=True)
    assert ad2.shape == (40, 12) and list(ad2.var_names[:2]) == ["g0", "g1"]
    with pytest.raises(NotImplementedError):
        io.read_dataset(123)


def test_write_text_matrix_format(tmp_path):
    from dca_b200 import io
    m = np.array([[1.0, 2.5], [3.25, 4.125], [5, 6]], np.float32)
    f = tmp_path / "mean.tsv"
    io.write_text_matrix(m, str(f), rownames=["c0", "c1", "c2"], colnames=["g0", "g1"], transpose=True)
    lines = open(f).read().rstrip("\n").split("\n")
    assert lines[0] == "\tc0\tc1\tc2" and lines[1] == "g0\t1.000000\t3.250000\t5.000000"


def test_cli_flags_match_reference_defaults():
    from dca_b200.__main__ import parse_args
    a = parse_args(["in.tsv", "out"])
    assert (a.type, a.batchsize, a.hiddensize, a.epochs, a.earlystop, a.reducelr) == ("nb-conddisp", 32, "64,32,64", 300, 15, 10)
    assert a.sizefactors and a.norminput and a.loginput and a.batchnorm and a.checkcounts
    assert not (a.transpose or a.testsplit or a.saveweights or a.hyper or a.debug or a.tensorboard)
    assert a.gradclip == 5.0 and a.learningrate is None and a.optimizer == "RMSprop" and a.ridge == 0.0
    b = parse_args(["in.tsv", "out", "--nosizefactors", "--nobatchnorm", "-t", "--type", "zinb", "-r", "0.01", "-s", "16,2,16"])
    assert not b.sizefactors and not b.batchnorm and b.transpose and b.type == "zinb" and b.learningrate == 0.01


def test_ae_types_registry_keys():
    from dca_b200.network import AE_types
    assert set(AE_types) == {'normal', 'poisson', 'nb', 'nb-conddisp', 'nb-shared', 'nb-fork', 'zinb', 'zinb-conddisp',
                             'zinb-shared', 'zinb-fork', 'zinb-elempi'}
    from dca_b200 import _lib
    for key, cls in AE_types.items():                     # every registry key maps to an engine type of the same name
        assert cls.ae_type == key and key in _lib.AE_TYPE_IDS, key
    assert AE_types['zinb-elempi'](input_size=5, sharedpi=True).sharedpi is True      # dca/network.py:425-427
    import dca.api, dca.network            # alias package resolves
    assert dca.network.AE_types is AE_types


def test_plateau_and_early_stop_match_oracle_fit_logic():
    from dca_b200.train import PlateauAndStop
    vals = [5.0, 4.0, 4.0, 4.00005, 3.9, 3.95, 3.95, 3.95, 3.95]
    c = PlateauAndStop(1e-3, reduce_lr=2, early_stop=3)
    lrs, stopped = [], None
    for e, v in enumerate(vals):
        lrs.append(c.lr)
        if c.on_epoch_end(e, v):
            stopped = e; break
    assert stopped == 7
    assert lrs[:5] == [1e-3, 1e-3, 1e-3, 1e-3, 1e-4]


def test_shard_bounds():
    from dca_b200.dist import shard_bounds
    assert shard_bounds(10, 0, 1) == (0, 10)
    assert [shard_bounds(10, r, 4, equal=True) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    b = [shard_bounds(10, r, 4) for r in range(4)]
    assert b[0][0] == 0 and b[-1][1] == 10 and all(b[i][1] == b[i + 1][0] for i in range(3))


def test_pack_counts_round_trip_and_auto_width():
    """Packed host format of the streaming path (dca_stream_begin_packed): lossless, escapes listed row-sorted."""
    from dca_b200 import io
    rng = np.random.default_rng(0)
    C = rng.poisson(0.3, (130, 64)).astype(np.int64)
    C[3, 5] = 300; C[3, 6] = 15; C[10, 63] = 70000; C[129, 0] = 14; C[0, 1] = 255
    for bits in (4, 8, 16, "dense"):
        pc = io.pack_counts(C, bits, batch=32)
        assert np.array_equal(io.unpack_counts(pc), C.astype(np.float32))
        assert pc.indptr[0] == 0 and pc.indptr[-1] == len(pc.entries) and np.all(np.diff(pc.indptr) >= 0)
        esc = (1 << pc.bits) - 1
        assert len(pc.entries) == int((C >= esc).sum())
        assert pc.packed.shape == (130, 64 * pc.bits // 8 // pc.packed.itemsize)
    assert io.pack_counts(C, "dense").bits == 4                       # small counts: 4 bits win among the dense widths
    # sparse format (dca_stream_begin_sparse): non-zero bitmap + 4-bit codes in gene order, rows byte-aligned
    for bits in ("sparse", "auto"):
        pc = io.pack_counts(C, bits, batch=32)
        assert pc.bits == 1 and np.array_equal(io.unpack_counts(pc), C.astype(np.float32))
        assert pc.packed.shape == (130, 8) and pc.nib_indptr[0] == 0
        nnz = (C != 0).sum(1)
        assert np.array_equal(np.diff(pc.nib_indptr), (nnz + 1) // 2)
        assert len(pc.entries) == int((C >= 15).sum()) and pc.nbytes < io.pack_counts(C, 4).nbytes
    dense_rows = rng.poisson(3.0, (64, 64))                           # ~95 % non-zeros: auto falls back to a dense width
    assert io.pack_counts(dense_rows, "auto", batch=32).bits == 4
    with pytest.raises(ValueError, match="50 %"):
        io.pack_counts(dense_rows, "sparse", batch=32)
    assert io.pack_counts(np.full((8, 64), 100), "auto").bits == 8    # everything >= 15: 8 bits win
    assert io.pack_counts(np.full((8, 64), 1000), "auto").bits == 16
    with pytest.raises(ValueError):
        io.pack_counts(np.full((4, 12), 1.0))                         # genes not a multiple of 8
    with pytest.raises(ValueError):
        io.pack_counts(np.full((4, 16), 0.5))                         # not integer counts


def _pandas_tsv(m, path, rownames, colnames, transpose):
    if transpose:
        m = m.T; rownames, colnames = colnames, rownames
    pd.DataFrame(m, index=rownames, columns=colnames).to_csv(path, sep='\t', index=(rownames is not None),
                                                             header=(colnames is not None), float_format='%.6f')


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_native_text_writer_is_byte_identical_to_pandas(tmp_path, dtype):
    """dca_write_text_matrix (multi-threaded, fixed-point digits) == pandas to_csv(sep='\\t', float_format='%.6f'),
    the writer of dca/io.py:120-129: rounding ties, negative zero, NaN / inf, huge and tiny values, labels that
    need quoting, with and without labels, transposed and not."""
    from dca_b200 import io
    rng = np.random.default_rng(0)
    m = (rng.standard_normal((37, 23)) * np.exp(rng.normal(0, 4, (37, 23)))).astype(dtype)
    m[0, :8] = [0.0, -0.0, 1 / 128, 3 / 128, -5 / 128, 0.0000005, -0.0000005, 1234567.0000005]
    m[1, :6] = [np.nan, np.inf, -np.inf, 1e-12, -1e-12, 9.9999995]
    m[2, :4] = [8.1e9, -3.4e38 if dtype == np.float32 else -1.7e300, 0.1234565, 2.5e-7]
    m[3, :] = np.arange(23) / 128.0 + 0.5 / 128.0            # every value a rounding tie at the 7th decimal
    rows = ["cell%d" % i for i in range(37)]; rows[4] = 'we"ird\tname'
    cols = ["g%d" % j for j in range(23)]
    cases = [(rows, cols, False), (rows, cols, True), (None, cols, False), (rows, None, True), (None, None, False)]
    for k, (rn, cn, tr) in enumerate(cases):
        a, b = tmp_path / ("native%d.tsv" % k), tmp_path / ("pandas%d.tsv" % k)
        io.write_text_matrix(m, str(a), rownames=rn, colnames=cn, transpose=tr, threads=3)
        _pandas_tsv(m, str(b), rn, cn, tr)
        assert a.read_bytes() == b.read_bytes(), (dtype, k)
    big = rng.random((3000, 50)).astype(dtype)               # several chunks per thread
    a, b = tmp_path / "big_native.tsv", tmp_path / "big_pandas.tsv"
    io.write_text_matrix(big, str(a), rownames=None, colnames=None, transpose=True)
    _pandas_tsv(big, str(b), None, None, True)
    assert a.read_bytes() == b.read_bytes()


def test_hostmem_helpers_are_noops_without_a_gpu():
    """hostmem.py: without CUDA the NUMA helpers change nothing and return plain host tensors."""
    import os
    import torch
    from dca_b200 import hostmem
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    before = os.sched_getaffinity(0)
    assert hostmem.gpu_local_cpus(0) is None
    with hostmem.near_gpu(0):
        assert os.sched_getaffinity(0) == before
    t = hostmem.pin_near_gpu(np.arange(6, dtype=np.float32).reshape(2, 3))
    assert isinstance(t, torch.Tensor) and t.shape == (2, 3) and not t.is_cuda
    assert os.sched_getaffinity(0) == before


def test_packed_counts_batch_bytes():
    from dca_b200 import io
    C = np.zeros((10, 16), dtype=np.int64); C[2, 3] = 99; C[7, 0] = 20
    pc = io.pack_counts(C, 4, batch=4)
    assert pc.bytes_for_rows(0, 4) == 4 * 8 + 8 * 5 + 8 * 1          # tile + indptr segment + one overflow entry
    assert pc.bytes_for_rows(4, 8) == 4 * 8 + 8 * 5 + 8 * 1
    assert pc.bytes_for_rows(8, 10) == 2 * 8 + 8 * 3


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint16, np.int32, np.int64, np.int16])
def test_native_count_packer_equals_numpy_statement(dtype):
    """dca_count_escapes + dca_pack_counts (multi-threaded) == the NumPy statement of the packed format."""
    from dca_b200 import io
    rng = np.random.default_rng(3)
    C = rng.poisson(0.4, (300, 72)).astype(np.int64)
    C[3, 5] = 300; C[3, 6] = 15; C[10, 63] = 30000; C[299, 0] = 14; C[0, 1] = 255; C[0, 2] = 254; C[7, :] = 20
    C = C.astype(dtype)
    for bits in (4, 8, 16, "dense", "sparse", "auto"):
        a = io.pack_counts(C, bits, batch=64, native=True, threads=3)
        b = io.pack_counts(C, bits, batch=64, native=False)
        assert a.bits == b.bits and a.packed.dtype == b.packed.dtype
        assert np.array_equal(a.packed, b.packed) and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.entries, b.entries)
        if a.bits == 1:                                   # sparse: dca_sparse_counts + dca_pack_sparse
            n = int(a.nib_indptr[-1])
            assert np.array_equal(a.nib_indptr, b.nib_indptr) and np.array_equal(a.nibbles[:n], b.nibbles[:n])
        assert np.array_equal(io.unpack_counts(a), C.astype(np.float32))
    with pytest.raises(ValueError, match="non-negative integers"):
        io.pack_counts(np.where(C > 0, -1, 0).astype(np.int32) if np.issubdtype(dtype, np.integer) else C + 0.5)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the keys of
    the bench contract; it needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "c2"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "cells/sec" and line["value"] > 0
    # both arms print ONE shared metric string (the driver only divides values whose metric / unit / direction agree)
    import re
    src = open(os.path.join(root, "bench.py")).read()
    assert len(re.findall(r'"metric": METRIC\b', src)) == 2 and line["metric"] == re.search(r'^METRIC = "(.*)"$', src, re.M).group(1)
    assert line["warmup"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_normalize_mutates_a_real_anndata_in_place(monkeypatch):
    """dca/io.py:88-111 works on the caller's object (scanpy's pp functions are in-place): X, obs, raw and the
    filtered cell / gene sets must land on the SAME object, for a real anndata.AnnData as for the lite stand-in."""
    from tests.util import install_fake_anndata
    from dca_b200 import io
    from dca_b200.anndata_lite import AnnData, is_anndata
    Fake = install_fake_anndata(monkeypatch)
    Y = synth_counts(50, 20, 3)
    Y[7] = 0                                        # a zero-count cell: normalize_per_cell drops it
    Y[:, 5] = 0                                     # an all-zero gene: filter_genes drops it (filter_min_counts=True)
    for cls in (Fake, AnnData):
        ad = cls(Y.copy())
        assert is_anndata(ad)
        same = io.read_dataset(ad, check_counts=True)
        assert same is ad
        out = io.normalize(ad, filter_min_counts=True)
        assert out is ad                                             # mutated in place, like the reference
        assert ad.X.shape == (49, 19) and ad.raw.X.shape == (49, 19)
        keep_r = np.arange(50) != 7; keep_c = np.arange(20) != 5
        np.testing.assert_array_equal(ad.raw.X, Y[keep_r][:, keep_c])
        X, sf = O.normalize_inputs(Y[keep_r][:, keep_c])
        np.testing.assert_allclose(ad.X, X, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.asarray(ad.obs["size_factors"]), sf, rtol=1e-6)
        assert "n_counts" in ad.obs.columns and len(ad.obs) == 49 and len(ad.var) == 19
        # without filtering the zero-count cell is still dropped by normalize_per_cell, on the same object
        ad2 = cls(Y[:, keep_c].copy())
        assert io.normalize(ad2, filter_min_counts=False) is ad2 and ad2.X.shape == (49, 19) and ad2.raw.X.shape == (49, 19)


def test_optimizer_names_follow_keras_module_attributes():
    """`opt.__dict__[optimizer]` (dca/train.py:54-57) resolves the class names and the lower-case aliases keras/optimizers.py
    defines; anything else (TFOptimizer, a typo) must be refused before any device work starts."""
    from dca_b200 import _lib
    from dca_b200.train import train
    for name in ("RMSprop", "SGD", "Adagrad", "Adadelta", "Adam", "Adamax", "Nadam"):
        assert _lib.OPTIMIZERS[name] == _lib.OPTIMIZERS[name.lower()]
    assert _lib.OPTIMIZERS["RMSprop"] == (0, 1e-3) and _lib.OPTIMIZERS["Adadelta"][1] == 1.0 and _lib.OPTIMIZERS["Nadam"][1] == 2e-3
    with pytest.raises(NotImplementedError):
        train(None, None, optimizer="TFOptimizer")          # refused before adata / network are touched
    with pytest.raises(TypeError):
        train(None, None, callbacks=[])                     # model.fit keywords the fit loop does not implement


def test_activation_names_cover_the_reference_choices():
    """dca/hyper.py:32 samples ('relu', 'selu', 'elu', 'PReLU', 'linear', 'LeakyReLU'); network.py:41 lists the two advanced
    activations; every one must map to a dca_activation id (softmax is the one Keras name that is refused)."""
    from dca_b200 import _lib
    for name in ("relu", "selu", "elu", "PReLU", "linear", "LeakyReLU", "tanh", "sigmoid", "softplus"):
        assert name in _lib.ACTIVATION_IDS
    assert "softmax" not in _lib.ACTIVATION_IDS
    assert sorted(_lib.ACTIVATION_IDS.values()) == list(range(12))
