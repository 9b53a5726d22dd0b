"""Data preparation and text I/O with the surface of dca/io.py:53-131.

scanpy / anndata are not available in the image, so the four ``sc.pp`` calls the reference
makes (dca/io.py:91-109, dca/api.py:163) are restated in NumPy (semantics in SURVEY.md A.1 /
Appendix B); real AnnData objects are accepted when ``anndata`` is importable.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import pandas as pd

from .anndata_lite import AnnData, is_anndata


def _dense(X):
    return X.toarray() if hasattr(X, "toarray") else np.asarray(X)


def read_text_or_h5ad(path: str):
    """sc.read(path, first_column_names=True) -- dca/io.py:59."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".h5ad":
        try:
            import anndata  # type: ignore
        except ImportError as e:
            raise ImportError("reading .h5ad needs the anndata package, which is not installed") from e
        return anndata.read_h5ad(path)
    sep = "," if ext == ".csv" else "\t"
    tab = pd.read_csv(path, sep=sep, index_col=0)
    return AnnData(tab.values.astype(np.float32), obs=pd.DataFrame(index=tab.index.astype(str)),
                   var=pd.DataFrame(index=tab.columns.astype(str)))


def read_dataset(adata, transpose=False, test_split=False, copy=False, check_counts=True):
    """dca/io.py:53-85."""
    if is_anndata(adata):
        if copy:
            adata = adata.copy()
    elif isinstance(adata, str):
        adata = read_text_or_h5ad(adata)
    else:
        raise NotImplementedError

    if check_counts:
        # check if observations are unnormalized using first 10           (dca/io.py:63-70)
        X_subset = _dense(adata.X[:10])
        norm_error = 'Make sure that the dataset (adata.X) contains unnormalized count data.'
        assert np.all(X_subset.astype(int) == X_subset), norm_error

    if transpose:
        adata = adata.transpose()

    if test_split:
        from sklearn.model_selection import train_test_split
        train_idx, test_idx = train_test_split(np.arange(adata.n_obs), test_size=0.1, random_state=42)
        spl = pd.Series(['train'] * adata.n_obs)
        spl.iloc[test_idx] = 'test'
        adata.obs['dca_split'] = spl.values
    else:
        adata.obs['dca_split'] = 'train'

    adata.obs['dca_split'] = adata.obs['dca_split'].astype('category')
    print('dca: Successfully preprocessed {} genes and {} cells.'.format(adata.n_vars, adata.n_obs))
    return adata


def filter_genes_mask(X, min_counts=1):
    """sc.pp.filter_genes(X, min_counts=1) -> (mask, counts)   (dca/api.py:163)."""
    counts = np.asarray(_dense(X).sum(axis=0)).reshape(-1)
    return counts >= min_counts, counts


def _inplace_subset(adata, rows=None, cols=None):
    """Boolean-mask subsetting IN PLACE (what scanpy's filter_genes / filter_cells / normalize_per_cell do to the
    caller's object): anndata.AnnData and the lite stand-in both provide _inplace_subset_var / _inplace_subset_obs."""
    if cols is not None:
        adata._inplace_subset_var(np.asarray(cols))
    if rows is not None:
        adata._inplace_subset_obs(np.asarray(rows))


def normalize(adata, filter_min_counts=True, size_factors=True, normalize_input=True, logtrans_input=True):
    """dca/io.py:88-111 with scanpy's arithmetic restated:
    filter_genes/filter_cells(min_counts=1); raw copy; normalize_per_cell (each cell scaled to the
    median total count; zero-count cells dropped as scanpy does); size_factors = n_counts/median;
    log1p (natural); scale (zero mean, unit variance with ddof=1, no clipping).

    Like the reference, this MUTATES the object it is given -- X, obs['n_counts'], obs['size_factors'], raw and (when
    filtering) the set of cells / genes -- and returns the same object, for the lite stand-in and for a real
    anndata.AnnData alike (only attribute assignment and the two in-place subsetting methods are used)."""
    if filter_min_counts:
        gmask, _ = filter_genes_mask(adata.X, 1)                       # dca/io.py:90-92
        if not gmask.all():
            _inplace_subset(adata, cols=gmask)
        cmask = np.asarray(_dense(adata.X).sum(axis=1)).reshape(-1) >= 1
        if not cmask.all():
            _inplace_subset(adata, rows=cmask)

    if size_factors or normalize_input or logtrans_input:              # dca/io.py:94-97
        adata.raw = adata.copy()
    else:
        adata.raw = adata

    X = _dense(adata.X)
    if size_factors:                                                   # dca/io.py:99-101
        n_counts = np.asarray(X.sum(axis=1, dtype=np.float64)).reshape(-1)
        keep = n_counts >= 1                         # normalize_per_cell filters cells with < 1 count
        if not keep.all():
            _inplace_subset(adata, rows=keep)        # anndata subsets .raw along obs as well
            X = _dense(adata.X)
            n_counts = n_counts[keep]
        med = np.median(n_counts)
        adata.obs['n_counts'] = n_counts
        X = (X / (n_counts / med)[:, None]).astype(np.float32)
        adata.obs['size_factors'] = (n_counts / med).astype(np.float32)
    else:
        adata.obs['size_factors'] = np.float32(1.0)                    # dca/io.py:102-103

    if logtrans_input:                                                 # dca/io.py:105-106
        X = np.log1p(X)

    if normalize_input:                                                # dca/io.py:108-109
        mean = X.mean(axis=0, dtype=np.float64)
        var = X.var(axis=0, ddof=1, dtype=np.float64) if X.shape[0] > 1 else np.ones(X.shape[1])
        std = np.sqrt(var)
        std[std == 0] = 1.0
        X = ((X - mean) / std).astype(np.float32)

    adata.X = np.ascontiguousarray(X, dtype=np.float32)
    return adata


def read_genelist(filename):
    genelist = list(set(open(filename, 'rt').read().strip().split('\n')))
    assert len(genelist) > 0, 'No genes detected in genelist file'
    print('dca: Subset of {} genes will be denoised.'.format(len(genelist)))
    return genelist


def write_text_matrix(matrix, filename, rownames=None, colnames=None, transpose=False, threads=0):
    """TSV with '%.6f' values, the files of dca/io.py:120-129 byte for byte.  float32 / float64 matrices go through
    the multi-threaded native writer (dca_write_text_matrix); anything else through pandas like the reference."""
    import ctypes as C
    m = np.asarray(matrix)
    native = m.ndim == 2 and m.dtype in (np.float32, np.float64) and m.size > 0
    if native:
        from . import _lib
        lib = _lib.load()           # raises when libdca_b200.so has not been built
    if not native:                  # non-float / empty / 1-d input: the reference's own pandas call
        if transpose:
            m = m.T
            rownames, colnames = colnames, rownames
        pd.DataFrame(m, index=rownames, columns=colnames).to_csv(
            filename, sep='\t', index=(rownames is not None), header=(colnames is not None), float_format='%.6f')
        return
    m = np.ascontiguousarray(m)

    def labels(names, n):
        if names is None:
            return None, None
        vals = [str(v).encode() for v in list(names)]
        if len(vals) != n:
            raise ValueError("got %d labels for %d rows/columns" % (len(vals), n))
        arr = (C.c_char_p * n)(*vals)
        return arr, vals
    rn, _keep_r = labels(rownames, m.shape[0])
    cn, _keep_c = labels(colnames, m.shape[1])
    _lib.check(lib.dca_write_text_matrix(os.fsencode(filename), m.ctypes.data, int(m.dtype == np.float64), m.shape[0],
                                         m.shape[1], m.shape[1], rn, cn, int(bool(transpose)), int(threads)),
               "dca_write_text_matrix")


def read_pickle(inputfile):
    return pickle.load(open(inputfile, "rb"))


# ------------------------------------------------------------------------------------------------
# Packed host format of a raw count matrix for the streaming path (dca_stream_begin_packed, include/dca_b200.h).
# No counterpart in the reference (it feeds float matrices to Keras, dca/train.py:78-98); scRNA-seq counts
# are >80 % zeros and mostly < 15, so 4 bits per entry + a short overflow list carry the same information
# as the float32 matrix in 1/8 of the bytes that cross PCIe every step.
class PackedCounts:
    """bits-per-entry matrix + CSR overflow list; see pack_counts().  bits == 1 is the SPARSE format: ``packed`` is the
    non-zero bitmap (n_genes/8 bytes per row), ``nibbles`` the 4-bit codes of the non-zero counts in gene order (each row
    starts on a byte boundary at ``nib_indptr[row]``), codes of 15 escape into the overflow list."""

    def __init__(self, packed, bits, n_genes, indptr, entries, nib_indptr=None, nibbles=None):
        self.packed, self.bits, self.n_genes, self.indptr, self.entries = packed, bits, n_genes, indptr, entries
        self.nib_indptr, self.nibbles = nib_indptr, nibbles
        self._pinned = None            # pinned host copies, made by DeviceEngine.stream_begin on first use

    @property
    def n_rows(self):
        return self.packed.shape[0]

    @property
    def nbytes(self):
        extra = (self.nib_indptr.nbytes + self.nibbles.nbytes) if self.bits == 1 else 0
        return self.packed.nbytes + self.indptr.nbytes + self.entries.nbytes + extra

    def bytes_for_rows(self, r0, r1):
        """host->device bytes of one batch [r0, r1): tile + indptr segment + overflow entries (+ nibble stream)."""
        b = (r1 - r0) * self.packed.shape[1] * self.packed.itemsize + 8 * (r1 - r0 + 1) + \
            8 * int(self.indptr[r1] - self.indptr[r0])
        if self.bits == 1:
            b += 8 * (r1 - r0 + 1) + int(self.nib_indptr[r1] - self.nib_indptr[r0])
        return b


OVERFLOW_ENTRY = np.dtype([("gene", "<i4"), ("count", "<f4")])


_NATIVE_DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.uint16): 2, np.dtype(np.int32): 3,
                  np.dtype(np.int64): 4}


def _choose_bits(per_row, n, g, batch):
    """Smallest total bytes among 4 / 8 / 16 bits whose per-batch overflow fits the device staging capacity.
    per_row: int64 [3][n] escapes per row at each width."""
    best = None
    for w, b in enumerate((4, 8, 16)):
        pr = per_row[w]
        if batch:
            cap = max(4096, batch * g // 32)
            worst = max(int(pr[i:i + batch].sum()) for i in range(0, max(n, 1), batch)) if n else 0
            if worst > cap:
                continue
        total = n * g * b / 8.0 + 8.0 * float(pr.sum())
        if best is None or total < best[0]:
            best = (total, b)
    if best is None:
        raise ValueError("no packing width fits the overflow capacity")
    return best[1]


def _pack_sparse_native(C, batch, threads=0):
    """Multi-threaded packer of the library (dca_sparse_counts + dca_pack_sparse)."""
    from . import _lib
    lib = _lib.load()
    if C.dtype not in _NATIVE_DTYPES:
        C = C.astype(np.float64 if C.dtype.kind == "f" else np.int64)
    C = np.ascontiguousarray(C)
    n, g = C.shape
    if g > 65536:
        raise ValueError("the sparse format supports at most 65536 genes")
    dt = _NATIVE_DTYPES[C.dtype]
    nnz = np.zeros(n, dtype=np.int64); esc = np.zeros(n, dtype=np.int64)
    st = lib.dca_sparse_counts(C.ctypes.data, dt, n, g, g, nnz.ctypes.data, esc.ctypes.data, int(threads))
    if st != 0:
        msg = lib.dca_last_error().decode("utf-8", "replace")
        raise ValueError("counts must be non-negative integers" if "non-negative" in msg else msg)
    nib_indptr = np.zeros(n + 1, dtype=np.int64); np.cumsum((nnz + 1) // 2, out=nib_indptr[1:])
    indptr = np.zeros(n + 1, dtype=np.int64); np.cumsum(esc, out=indptr[1:])
    if batch:
        worst = max(int(nib_indptr[min(i + batch, n)] - nib_indptr[i]) for i in range(0, max(n, 1), batch)) if n else 0
        if worst > batch * g // 4 + 64:
            raise ValueError("more than 50 % non-zero entries in a batch: use a dense width (bits=4)")
    bitmap = np.empty((n, g // 8), dtype=np.uint8)
    nibbles = np.zeros(int(nib_indptr[-1]) + 16, dtype=np.uint8)
    entries = np.empty(int(indptr[-1]), dtype=OVERFLOW_ENTRY)
    _lib.check(lib.dca_pack_sparse(C.ctypes.data, dt, n, g, g, bitmap.ctypes.data, nib_indptr.ctypes.data, nibbles.ctypes.data,
                                   indptr.ctypes.data, entries.ctypes.data if len(entries) else None, int(threads)), "dca_pack_sparse")
    return PackedCounts(bitmap, 1, g, indptr, entries, nib_indptr, nibbles)


def _pack_sparse(C, batch):
    """NumPy statement of the sparse format (dca_stream_begin_sparse, include/dca_b200.h)."""
    n, g = C.shape
    if g > 65536:
        raise ValueError("the sparse format supports at most 65536 genes")
    nz = C != 0
    bitmap = np.packbits(nz, axis=1, bitorder="little")                    # [n, g/8]
    rows, cols = np.nonzero(nz)                                             # row-major: by row, then gene
    vals = C[rows, cols]
    codes = np.minimum(vals, 15).astype(np.uint8)
    cnt = np.bincount(rows, minlength=n).astype(np.int64)
    nib_indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum((cnt + 1) // 2, out=nib_indptr[1:])                          # every row starts on a byte boundary
    first = np.zeros(n + 1, dtype=np.int64); np.cumsum(cnt, out=first[1:])
    k = np.arange(rows.shape[0], dtype=np.int64) - first[rows]             # index of the code inside its row
    byte = nib_indptr[rows] + (k >> 1)
    nibbles = np.zeros(int(nib_indptr[-1]) + 16, dtype=np.uint8)           # + slack: the device reads whole bytes
    even = (k & 1) == 0
    nibbles[byte[even]] = codes[even]
    nibbles[byte[~even]] |= (codes[~even] << 4).astype(np.uint8)
    over = vals >= 15
    orow, ocol = rows[over], cols[over]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(orow, minlength=n), out=indptr[1:])
    entries = np.empty(orow.shape[0], dtype=OVERFLOW_ENTRY)
    entries["gene"] = ocol; entries["count"] = vals[over]
    if batch:
        worst = max(int(nib_indptr[min(i + batch, n)] - nib_indptr[i]) for i in range(0, max(n, 1), batch)) if n else 0
        if worst > batch * g // 4 + 64:
            raise ValueError("more than 50 % non-zero entries in a batch: use a dense width (bits=4)")
    return PackedCounts(np.ascontiguousarray(bitmap), 1, g, indptr, entries, nib_indptr, nibbles)


def pack_counts(counts, bits="auto", batch=None, native=True, threads=0):
    """Pack an integer-valued count matrix (cells x genes, any numeric dtype) into `bits` bits per entry.

    Counts >= 2**bits - 1 are stored as the escape value 2**bits - 1 and listed (row-sorted) in the overflow
    CSR: indptr int64[n_rows+1], entries {int32 gene, float32 count}.  bits='auto' picks the width in
    (4, 8, 16) with the fewest total bytes whose per-batch overflow (when `batch` is given) stays under
    batch*genes/32 entries (the device staging capacity).  native=True runs the multi-threaded packer of the
    library (dca_count_escapes / dca_pack_counts); native=False is the NumPy statement of the same format."""
    C = np.asarray(counts)
    if C.ndim != 2:
        raise ValueError("counts must be a 2-d matrix")
    n, g = C.shape
    if g % 8 != 0:
        raise ValueError("the number of genes must be a multiple of 8 for the packed format (got %d)" % g)
    if bits not in ("auto", "sparse", "dense") and bits not in (4, 8, 16):
        raise ValueError("bits must be 4, 8, 16, 'sparse', 'dense' (best dense width) or 'auto' (smallest of all)")
    if bits in ("sparse", "auto") and n > 0:
        if not native and C.size and (C.min() < 0 or np.any(C != np.floor(C))):
            raise ValueError("counts must be non-negative integers")
        nnz = int(np.count_nonzero(C))
        # sparse: 1 bit per entry + 4 bits per non-zero (+ 8 B per count >= 15); dense 4-bit: 4 bits per entry
        if bits == "sparse" or (nnz < 0.45 * C.size and C.size / 8.0 + nnz / 2.0 < 0.8 * C.size / 2.0):
            try:
                return _pack_sparse_native(C, batch, threads) if native else _pack_sparse(C, batch)
            except ValueError:
                if bits == "sparse":
                    raise
    if bits in ("sparse", "dense"):
        bits = "auto"
    if native and n > 0:
        return _pack_counts_native(C, bits, batch, threads)
    if C.size and (C.min() < 0 or np.any(C != np.floor(C))):
        raise ValueError("counts must be non-negative integers")
    if bits == "auto":
        per_row = np.stack([(C >= (1 << b) - 1).sum(1) for b in (4, 8, 16)]).astype(np.int64)
        bits = _choose_bits(per_row, n, g, batch)
    esc = (1 << bits) - 1
    over = C >= esc
    base = np.where(over, esc, C).astype(np.uint16 if bits == 16 else np.uint8)
    if bits == 4:
        packed = (base[:, 0::2] | (base[:, 1::2] << 4)).astype(np.uint8)
    else:
        packed = base
    rows, cols = np.nonzero(over)                     # row-major order: sorted by row, then gene
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    entries = np.empty(rows.shape[0], dtype=OVERFLOW_ENTRY)
    entries["gene"] = cols
    entries["count"] = C[rows, cols]
    return PackedCounts(np.ascontiguousarray(packed), bits, g, indptr, entries)


def _pack_counts_native(C, bits, batch, threads):
    from . import _lib
    lib = _lib.load()
    if C.dtype not in _NATIVE_DTYPES:
        C = C.astype(np.float64 if C.dtype.kind == "f" else np.int64)
    C = np.ascontiguousarray(C)
    n, g = C.shape
    dt = _NATIVE_DTYPES[C.dtype]
    per_row = np.zeros((3, n), dtype=np.int64)
    st = lib.dca_count_escapes(C.ctypes.data, dt, n, g, g, per_row.ctypes.data, int(threads))
    if st != 0:
        msg = lib.dca_last_error().decode("utf-8", "replace")
        raise ValueError("counts must be non-negative integers" if "non-negative" in msg else msg)
    if bits == "auto":
        bits = _choose_bits(per_row, n, g, batch)
    w = (4, 8, 16).index(bits)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(per_row[w], out=indptr[1:])
    packed = np.empty((n, g // 2) if bits == 4 else (n, g), dtype=np.uint16 if bits == 16 else np.uint8)
    entries = np.empty(int(indptr[-1]), dtype=OVERFLOW_ENTRY)
    _lib.check(lib.dca_pack_counts(C.ctypes.data, dt, n, g, g, bits, packed.ctypes.data, indptr.ctypes.data,
                                   entries.ctypes.data if len(entries) else None, int(threads)), "dca_pack_counts")
    return PackedCounts(packed, bits, g, indptr, entries)


def unpack_counts(pc: PackedCounts):
    """Inverse of pack_counts (float32 matrix) -- the host statement of what the device expansion produces."""
    if pc.bits == 1:
        nz = np.unpackbits(pc.packed, axis=1, bitorder="little")[:, :pc.n_genes].astype(bool)
        out = np.zeros((pc.n_rows, pc.n_genes), dtype=np.float32)
        rows, cols = np.nonzero(nz)
        cnt = np.bincount(rows, minlength=pc.n_rows).astype(np.int64)
        first = np.zeros(pc.n_rows + 1, dtype=np.int64); np.cumsum(cnt, out=first[1:])
        k = np.arange(rows.shape[0], dtype=np.int64) - first[rows]
        b = pc.nibbles[pc.nib_indptr[rows] + (k >> 1)]
        out[rows, cols] = np.where(k & 1, b >> 4, b & 0xF)
        orow = np.repeat(np.arange(pc.n_rows), np.diff(pc.indptr))
        out[orow, pc.entries["gene"]] = pc.entries["count"]
        return out
    if pc.bits == 4:
        out = np.empty((pc.n_rows, pc.n_genes), dtype=np.float32)
        out[:, 0::2] = pc.packed & 0xF
        out[:, 1::2] = pc.packed >> 4
    else:
        out = pc.packed.astype(np.float32)
    rows = np.repeat(np.arange(pc.n_rows), np.diff(pc.indptr))
    out[rows, pc.entries["gene"]] = pc.entries["count"]
    return out
