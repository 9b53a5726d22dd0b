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
