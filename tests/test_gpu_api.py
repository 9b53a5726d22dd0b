"""API-surface contract of the reference's only test (dca/test.py:6-59), on synthetic data
(paul15 needs a network download), plus the CLI round trip.  Needs a GPU."""
import os
import numpy as np
import pandas as pd
import pytest

from tests.util import synth_counts

pytestmark = pytest.mark.gpu


def _adata(n=300, g=120, seed=0):
    from dca_b200.anndata_lite import AnnData
    return AnnData(synth_counts(n, g, seed))


def test_api_denoise_and_info_keys():
    from dca_b200.api import dca
    adata = _adata()
    epochs = 1
    ret = dca(adata, mode='denoise', copy=True, epochs=epochs, verbose=True)
    assert not np.allclose(ret.X[:10], adata.X[:10])
    ret, model = dca(adata, mode='denoise', ae_type='nb-conddisp', copy=True, epochs=epochs,
                     return_model=True, return_info=True)
    assert not np.allclose(ret.X[:10], adata.X[:10])
    assert 'X_dca_dispersion' in ret.obsm_keys() and model is not None
    assert ret.obsm['X_dca_dispersion'].shape == adata.X.shape
    ret = dca(adata, mode='denoise', ae_type='nb', copy=True, epochs=epochs, return_model=False, return_info=True)
    assert not np.allclose(ret.X[:10], adata.X[:10])
    assert 'X_dca_dispersion' in ret.var_keys()
    ret = dca(adata, mode='denoise', ae_type='zinb', copy=True, epochs=epochs, return_model=False, return_info=True)
    assert 'X_dca_dropout' in ret.obsm_keys() and 'dca_loss_history' in ret.uns_keys()
    assert set(ret.uns['dca_loss_history']) == {'loss', 'val_loss', 'lr'}
    ret = dca(adata, mode='denoise', ae_type='zinb-conddisp', copy=True, epochs=2, return_info=True, batch_size=64)
    assert np.all(np.isfinite(ret.X)) and np.all(ret.X > 0)
    assert np.all((ret.obsm['X_dca_dropout'] >= 0) & (ret.obsm['X_dca_dropout'] <= 1))
    np.testing.assert_array_equal(ret.raw.X, adata.X)        # raw counts kept
    assert 'size_factors' in ret.obs.columns
    # the input object is untouched with copy=True
    assert 'dca_split' not in adata.obs.columns


def test_api_latent_modes():
    from dca_b200.api import dca
    adata = _adata(seed=1)
    hid_size = (10, 2, 10)
    for t in (None, 'nb-conddisp', 'nb', 'zinb'):
        kw = {} if t is None else {'ae_type': t}
        ret = dca(adata, mode='latent', hidden_size=hid_size, copy=True, epochs=1, **kw)
        assert 'X_dca' in ret.obsm_keys() and ret.obsm['X_dca'].shape[1] == hid_size[1]
        np.testing.assert_array_equal(ret.X, adata.X)         # latent mode restores raw counts (network.py:208-209)


def test_api_inplace_and_errors():
    from dca_b200.api import dca
    adata = _adata(seed=2)
    before = adata.X.copy()
    out = dca(adata, epochs=1)
    assert out is None and not np.allclose(adata.X[:10], before[:10]) and adata.raw is not None
    with pytest.raises(AssertionError, match='valid mode'):
        dca(_adata(), mode='full')
    with pytest.raises(AssertionError, match='AnnData'):
        dca(np.zeros((3, 3)))
    bad = _adata(); bad.X[:, 3] = 0
    with pytest.raises(AssertionError, match='all-zero genes'):
        dca(bad, epochs=1)
    with pytest.raises(KeyError):
        dca(_adata(), ae_type='gaussian', epochs=1)


def test_api_remaining_ae_types():
    """The rest of dca/test.py:31-41 ('zinb-elempi' with and without sharedpi) and every other registry key of
    dca/network.py:763-768 through the public API: outputs changed, info keys present with the reference's shapes."""
    from dca_b200.api import dca
    adata = _adata(seed=6)
    ret = dca(adata, mode='denoise', ae_type='zinb-elempi', copy=True, epochs=1, return_model=False, return_info=True)
    assert not np.allclose(ret.X[:10], adata.X[:10])
    assert 'X_dca_dropout' in ret.obsm_keys() and 'dca_loss_history' in ret.uns_keys()
    ret = dca(adata, mode='denoise', ae_type='zinb-elempi', copy=True, epochs=1, return_model=False, return_info=True,
              network_kwds={'sharedpi': True})
    assert not np.allclose(ret.X[:10], adata.X[:10])
    assert 'X_dca_dropout' in ret.obsm_keys() and ret.obsm['X_dca_dropout'].shape == adata.X.shape
    for t, disp_shape, has_pi in (('poisson', None, False), ('normal', None, False), ('nb-shared', (300, 1), False),
                                  ('zinb-shared', (300, 1), True), ('nb-fork', (300, 120), False), ('zinb-fork', (300, 120), True)):
        ret = dca(adata, mode='denoise', ae_type=t, copy=True, epochs=2, return_info=True, batch_size=64)
        assert np.all(np.isfinite(ret.X)) and not np.allclose(ret.X[:10], adata.X[:10]), t
        h = ret.uns['dca_loss_history']
        assert len(h['loss']) == 2 and np.all(np.isfinite(h['loss'])) and np.all(np.isfinite(h['val_loss'])), t
        if disp_shape is None:
            assert 'X_dca_dispersion' not in ret.obsm_keys()
        else:
            assert ret.obsm['X_dca_dispersion'].shape == disp_shape, t
        assert ('X_dca_dropout' in ret.obsm_keys()) == has_pi, t
        lat = dca(adata, mode='latent', ae_type=t, copy=True, epochs=1, hidden_size=(16, 4, 16))
        assert lat.obsm['X_dca'].shape == (300, 4), t


def test_api_mutates_a_real_anndata_in_place(monkeypatch):
    """dca(adata, copy=False) on a (duck-typed) real anndata.AnnData: the results land on the caller's object
    (dca/io.py:88-111, dca/api.py:166-211), and copy=True returns an object of the caller's class."""
    from tests.util import install_fake_anndata
    from dca_b200.api import dca
    Fake = install_fake_anndata(monkeypatch)
    Y = synth_counts(200, 64, 5)
    ad = Fake(Y.copy())
    out = dca(ad, ae_type='zinb-conddisp', epochs=1, return_info=True, mode='denoise')
    assert out is None
    assert not np.allclose(ad.X[:10], Y[:10]) and np.all(np.isfinite(ad.X))
    np.testing.assert_array_equal(ad.raw.X, Y)
    assert 'size_factors' in ad.obs.columns and 'dca_split' in ad.obs.columns
    assert set(ad.obsm) >= {'X_dca_dispersion', 'X_dca_dropout'} and 'dca_loss_history' in ad.uns
    ad2 = Fake(Y.copy())
    ret = dca(ad2, ae_type='nb-conddisp', epochs=1, copy=True)
    assert type(ret) is Fake and ret is not ad2 and 'dca_split' not in ad2.obs.columns
    np.testing.assert_array_equal(ad2.X, Y)


def test_training_reduces_loss_and_early_stop_history():
    from dca_b200.api import dca
    adata = _adata(400, 80, 3)
    ret = dca(adata, ae_type='zinb-conddisp', copy=True, epochs=12, return_info=True, batch_size=32, verbose=False)
    h = ret.uns['dca_loss_history']
    assert len(h['loss']) == len(h['val_loss']) == len(h['lr']) <= 12
    assert h['loss'][-1] < h['loss'][0]


def test_cli_round_trip(tmp_path):
    from dca_b200.__main__ import main
    Y = synth_counts(120, 60, 4).astype(int)
    df = pd.DataFrame(Y.T, index=["g%d" % i for i in range(60)], columns=["c%d" % i for i in range(120)])
    inp = tmp_path / "counts.tsv"; df.to_csv(inp, sep="\t")
    out = tmp_path / "out"
    main([str(inp), str(out), "--type", "zinb-conddisp", "-e", "2", "--saveweights"])
    for f in ("mean.tsv", "latent.tsv", "dispersion.tsv", "dropout.tsv", "model.pickle", "weights.npz"):
        assert (out / f).exists(), f
    mean = pd.read_csv(out / "mean.tsv", sep="\t", index_col=0)
    assert mean.shape == (60, 120) and list(mean.index[:2]) == ["g0", "g1"]      # gene x cell like the reference
    lat = pd.read_csv(out / "latent.tsv", sep="\t", index_col=0, header=None)
    assert lat.shape == (120, 32)
