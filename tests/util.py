import numpy as np


def synth_counts(n_cells, n_genes, seed=0, zero_frac=0.2):
    """Synthetic ZINB-like count matrix per SURVEY.md 8d (small host version)."""
    rng = np.random.default_rng(seed)
    gene_logmean = rng.normal(-1.0, 1.5, size=(1, n_genes))
    depth = np.exp(rng.normal(0, 0.35, size=(n_cells, 1)))
    lam = rng.gamma(2.0, depth * np.exp(gene_logmean) / 2.0)
    Y = rng.poisson(lam).astype(np.float32)
    Y[rng.random(Y.shape) < zero_frac] = 0
    dead = Y.sum(0) == 0
    Y[rng.integers(0, n_cells, dead.sum()), np.where(dead)[0]] = 1
    empty = Y.sum(1) == 0
    Y[np.where(empty)[0], rng.integers(0, n_genes, empty.sum())] = 1
    return Y


def rel_err(got, ref, floor_frac=1e-3):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    scale = np.maximum(np.abs(ref), floor_frac * np.max(np.abs(ref)) + 1e-30)
    return float(np.max(np.abs(got - ref) / scale))


class FakeRaw:
    def __init__(self, X, var):
        self.X, self.var = X, var

    @property
    def var_names(self):
        return self.var.index


class FakeAnnData:
    """Duck-typed stand-in for the REAL anndata.AnnData (the package is not in the image): deliberately NOT derived
    from dca_b200.anndata_lite.AnnData.  Tests install it as ``sys.modules['anndata'].AnnData`` so that
    ``is_anndata`` takes its real-anndata branch; it offers only anndata's own API (attribute assignment, copy(),
    boolean-Series row views, ``raw`` setter that freezes a copy, ``_inplace_subset_obs/_var``)."""

    def __init__(self, X, obs=None, var=None):
        import pandas as pd
        self.X = np.asarray(X, dtype=np.float32)
        n, g = self.X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=["c%d" % i for i in range(n)])
        self.var = var if var is not None else pd.DataFrame(index=["g%d" % i for i in range(g)])
        self.obsm, self.uns, self._raw = {}, {}, None

    n_obs = property(lambda s: s.X.shape[0])
    n_vars = property(lambda s: s.X.shape[1])
    obs_names = property(lambda s: s.obs.index)
    var_names = property(lambda s: s.var.index)
    raw = property(lambda s: s._raw)

    @raw.setter
    def raw(self, value):
        self._raw = None if value is None else FakeRaw(np.array(value.X, copy=True), value.var.copy())

    def obsm_keys(self): return list(self.obsm)
    def var_keys(self): return list(self.var.columns)

    def copy(self):
        c = FakeAnnData(self.X.copy(), self.obs.copy(), self.var.copy())
        c.obsm = {k: np.array(v, copy=True) for k, v in self.obsm.items()}; c.uns = dict(self.uns)
        c._raw = None if self._raw is None else FakeRaw(self._raw.X.copy(), self._raw.var.copy())
        return c

    def __getitem__(self, idx):
        idx = np.asarray(idx.values if hasattr(idx, "values") else idx)
        v = FakeAnnData(self.X[idx], self.obs[idx], self.var)
        v._raw = None if self._raw is None else FakeRaw(self._raw.X[idx], self._raw.var)
        return v

    def _inplace_subset_obs(self, mask):
        mask = np.asarray(mask)
        self.X = self.X[mask]; self.obs = self.obs[mask]
        if self._raw is not None:
            self._raw = FakeRaw(self._raw.X[mask], self._raw.var)

    def _inplace_subset_var(self, mask):
        mask = np.asarray(mask)
        self.X = self.X[:, mask]; self.var = self.var[mask]


def install_fake_anndata(monkeypatch):
    import sys, types
    mod = types.ModuleType("anndata")
    mod.AnnData = FakeAnnData
    monkeypatch.setitem(sys.modules, "anndata", mod)
    return FakeAnnData
