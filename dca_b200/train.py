"""Training driver with the surface of dca/train.py:35-191.

``train()`` replaces ``model.compile`` + ``model.fit`` (dca/train.py:54-98): the data stay
resident in HBM, every batch is one dca_train_step + gradient all-reduce (when launched under
torch.distributed) + dca_apply_update, validation is the tail ``validation_split`` of the rows
(taken before shuffling, as Keras does), and ReduceLROnPlateau / EarlyStopping are evaluated on
the host from one scalar per epoch.
"""
from __future__ import annotations

import os
import random
from typing import Optional

import numpy as np
import torch

from . import dist as D
from .engine import KERAS_DEFAULTS


class History:
    """Stand-in for keras.callbacks.History (only ``.history`` is used, dca/api.py:206)."""

    def __init__(self):
        self.history = {"loss": [], "val_loss": [], "lr": []}
        self.epoch = []


class PlateauAndStop:
    """ReduceLROnPlateau(monitor='val_loss', factor=0.1, min_delta=1e-4, cooldown=0, min_lr=0)
    followed by EarlyStopping(monitor='val_loss', min_delta=0) -- dca/train.py:70-75."""

    def __init__(self, lr, reduce_lr, early_stop, verbose=False):
        self.lr, self.reduce_lr, self.early_stop, self.verbose = lr, reduce_lr, early_stop, verbose
        self.best = np.inf; self.wait = 0
        self.es_best = np.inf; self.es_wait = 0
        self.stopped_epoch = None

    def on_epoch_end(self, epoch, monitor) -> bool:
        """Returns True when training should stop."""
        if monitor is None:
            return False
        if self.reduce_lr:
            if monitor < self.best - 1e-4:
                self.best = monitor; self.wait = 0
            else:
                self.wait += 1
                if self.wait >= self.reduce_lr:
                    new_lr = self.lr * 0.1
                    if self.verbose:
                        print("\nEpoch %05d: ReduceLROnPlateau reducing learning rate to %s." % (epoch + 1, new_lr))
                    self.lr = new_lr; self.wait = 0
        if self.early_stop:
            if monitor < self.es_best:
                self.es_best = monitor; self.es_wait = 0
            else:
                self.es_wait += 1
                if self.es_wait >= self.early_stop:
                    self.stopped_epoch = epoch
                    if self.verbose:
                        print("Epoch %05d: early stopping" % (epoch + 1))
                    return True
        return False


def _to_device(a, dtype, device):
    t = torch.as_tensor(np.ascontiguousarray(a))
    return t.to(device=device, dtype=dtype, non_blocking=False).contiguous()


def train(adata, network, output_dir=None, optimizer='RMSprop', learning_rate=None,
          epochs=300, reduce_lr=10, output_subset=None, use_raw_as_output=True,
          early_stop=15, batch_size=32, clip_grad=5., save_weights=False,
          validation_split=0.1, tensorboard=False, verbose=True, threads=None,
          **kwds):
    """Same signature as dca/train.py:35-39.  ``threads`` is accepted and ignored (GPU path).

    Extra keyword (``training_kwds`` of ``dca()``): ``stream`` -- False (default): the shard lives in HBM for the whole
    run, rows are reshuffled every epoch exactly like Keras; True / 'auto': train from HOST memory through
    dca_stream_* (raw counts bit-packed in pinned memory, every step copies its batch host->device while the previous
    one computes and normalises it on the device, dca/io.py:99-109 restated) -- for matrices that do not fit the GPU
    ('auto' switches when X + Y would exceed 60 % of the free device memory).  In streaming mode the training rows are
    shuffled ONCE and every epoch visits the batches in a new random order (documented deviation from Keras' per-epoch
    row shuffle).  Other keywords of the reference's model.fit (e.g. shuffle=False) are honoured or rejected loudly."""
    stream = kwds.pop('stream', False)
    shuffle = kwds.pop('shuffle', True)
    if kwds:
        raise TypeError("train() got keyword arguments the accelerated fit loop does not implement: %s" % sorted(kwds))
    from . import _lib as _L
    if optimizer not in _L.OPTIMIZERS:
        raise NotImplementedError("optimizer %r is not on the accelerated path (supported: %s)"
                                  % (optimizer, sorted(k for k in _L.OPTIMIZERS if not k.islower())))
    if tensorboard:
        raise NotImplementedError("tensorboard logging is not part of the accelerated path")
    if output_dir is not None:
        os.makedirs(output_dir, exist_ok=True)

    X = np.asarray(adata.X, dtype=np.float32)
    sf = np.asarray(adata.obs['size_factors'], dtype=np.float32).reshape(-1)
    if output_subset:
        raw_names = np.asarray(adata.raw.var_names)
        gene_idx = [np.where(raw_names == x)[0][0] for x in output_subset]
        Yh = adata.raw.X[:, gene_idx] if use_raw_as_output else adata.X[:, gene_idx]
    else:
        Yh = adata.raw.X if use_raw_as_output else adata.X
    Yh = np.asarray(Yh.toarray() if hasattr(Yh, "toarray") else Yh, dtype=np.float32)

    N = X.shape[0]
    split_at = int(N * (1. - validation_split)) if validation_split and 0. < validation_split < 1. else N
    rank, world = D.rank_world()

    # cells shard across ranks: contiguous row ranges, equal count per rank
    tr_lo, tr_hi = D.shard_bounds(split_at, rank, world, equal=True)
    va_lo, va_hi = D.shard_bounds(N - split_at, rank, world, equal=False)
    va_lo += split_at; va_hi += split_at

    if world > 1 and rank == 0 and (split_at % world) and verbose:
        print("dca: %d training cells do not divide over %d ranks; the last %d are left out of every epoch"
              % (split_at, world, split_at % world))
    eng = network.ensure_engine(max_batch=batch_size)
    dev = eng.device
    if stream == 'auto':
        free = torch.cuda.mem_get_info(dev)[0]
        stream = (tr_hi - tr_lo + va_hi - va_lo) * X.shape[1] * (4 + eng.params.element_size()) > 0.6 * free
    # opt.__dict__[optimizer](clipvalue=clip_grad[, lr=learning_rate])                   (dca/train.py:54-57)
    default_lr = eng.set_optimizer(optimizer)
    if learning_rate is None:
        learning_rate = default_lr
    if stream:
        return _fit_stream(eng, network, X, Yh, sf, (tr_lo, tr_hi), (va_lo, va_hi), batch_size, epochs, learning_rate, reduce_lr,
                           early_stop, clip_grad, world, rank, verbose, save_weights, output_dir, shuffle)
    Xd = _to_device(np.concatenate([X[tr_lo:tr_hi], X[va_lo:va_hi]]), eng.x_dtype, dev)
    Yd = _to_device(np.concatenate([Yh[tr_lo:tr_hi], Yh[va_lo:va_hi]]), torch.float32, dev)
    sfd = _to_device(np.concatenate([sf[tr_lo:tr_hi], sf[va_lo:va_hi]]), torch.float32, dev)
    n_tr = tr_hi - tr_lo
    n_va = va_hi - va_lo

    if world > 1:
        D.broadcast_(eng.params, src=0); D.broadcast_(eng.bn_state, src=0)
        eng.params_changed()
        if torch.distributed.get_backend() == "nccl":
            eng.comm_init()          # gradient exchange inside the library: one CUDA graph per step (dca_train_step_dp)
    eng.reset_optimizer()

    lr = float(learning_rate)
    ctl = PlateauAndStop(lr, reduce_lr, early_stop, verbose)
    hist = History()
    if verbose:
        print(network.summary())

    steps = (n_tr + batch_size - 1) // batch_size
    gscale = 1.0 / world
    # run on a non-default stream (CUDA-graph replay of the step needs a capturable stream)
    torch.cuda.synchronize(dev)
    prev_stream = torch.cuda.current_stream(dev)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    try:
        hist = _fit_loop(eng, network, Xd, Yd, sfd, n_tr, n_va, steps, batch_size, epochs, ctl, clip_grad, gscale, world, rank,
                         dev, hist, verbose, save_weights, output_dir, shuffle)
    finally:
        torch.cuda.synchronize(dev)
        torch.cuda.set_stream(prev_stream)
    if not hist.history["val_loss"]:
        del hist.history["val_loss"]
    return hist


def _epoch_end(eng, network, hist, ctl, epoch, epochs, n_va, world, rank, dev, verbose, save_weights, output_dir, best_val):
    """Epoch bookkeeping shared by the resident and the streaming loop: reduce the accumulators over ranks, history,
    ModelCheckpoint, ReduceLROnPlateau / EarlyStopping.  Returns (stop, best_val)."""
    acc = np.asarray(eng.read_epoch_acc(reset=True), dtype=np.float64)
    if world > 1:
        acc = D.all_reduce_sum_host(acc, dev)
        if eng.bn_state.numel():
            D.all_reduce_sum_(eng.bn_state); eng.bn_state.mul_(1.0 / world)
    loss = acc[0] / acc[1] if acc[1] > 0 else float("nan")
    if not np.isfinite(loss):
        loss = float("inf")                       # _nan2inf convention, dca/loss.py:148
    val = None
    if n_va > 0 or (world > 1 and acc[3] > 0):
        val = acc[2] / acc[3] + network.penalty_value()
        if not np.isfinite(val):
            val = float("inf")
    hist.epoch.append(epoch)
    hist.history["loss"].append(float(loss))
    hist.history["lr"].append(float(ctl.lr))
    if val is not None:
        hist.history["val_loss"].append(float(val))
    if verbose and rank == 0:
        print("Epoch %d/%d - loss: %.4f%s - lr: %g" % (epoch + 1, epochs, loss,
                                                    "" if val is None else " - val_loss: %.4f" % val, ctl.lr))
    if save_weights and output_dir is not None and rank == 0:
        mon = val if val is not None else loss
        if mon < best_val:                       # ModelCheckpoint(save_best_only=True), dca/train.py:64-69
            best_val = mon
            network.save_weights(os.path.join(output_dir, "weights.npz"))
    return ctl.on_epoch_end(epoch, val), best_val


def _fit_stream(eng, network, X, Yh, sf, tr, va, batch_size, epochs, learning_rate, reduce_lr, early_stop, clip_grad, world, rank,
                verbose, save_weights, output_dir, shuffle):
    """Training from host memory (dca_stream_*): see train().  X is only used for the (small, resident) validation rows;
    the training rows travel as bit-packed raw counts and are normalised on the device."""
    from . import io as dio
    from .hostmem import pin_near_gpu
    dev = eng.device
    (tr_lo, tr_hi), (va_lo, va_hi) = tr, va
    if eng.n_in != eng.n_out or Yh.shape[1] != X.shape[1]:
        raise NotImplementedError("stream=True needs the raw counts of the input genes as the target (no output_subset)")
    n_tr, n_va = tr_hi - tr_lo, va_hi - va_lo
    order0 = np.arange(tr_lo, tr_hi)
    if shuffle:
        np.random.shuffle(order0)                 # ONE row shuffle; the epochs permute whole batches
    Ytr = np.ascontiguousarray(Yh[order0]); sftr = np.ascontiguousarray(sf[order0])
    # the transform X = (log1p(y / sf) - mean_g) / std_g the host normalisation applied (dca/io.py:99-109), recovered from
    # (X, raw) of a few hundred rows: two unknowns per gene
    l = np.log1p(Yh / sf[:, None]) if n_tr + n_va <= 4096 else None
    if l is None:
        pick = np.linspace(0, Yh.shape[0] - 1, 4096).astype(np.int64)
        l = np.log1p(Yh[pick] / sf[pick, None]); xs = X[pick]
    else:
        xs = X
    lm, xm = l.mean(0, dtype=np.float64), xs.mean(0, dtype=np.float64)
    lv = ((l - lm) * (xs - xm)).sum(0, dtype=np.float64); xv = ((xs - xm) ** 2).sum(0, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        std = np.where(xv > 0, lv / xv, 1.0)       # l = mean + std * x  =>  std = cov(l, x) / var(x)
    std[~np.isfinite(std) | (std <= 0)] = 1.0
    mean = lm - std * xm
    if float(np.max(np.abs((l - mean) / std - xs))) > 1e-2:
        raise NotImplementedError("stream=True supports the default preprocessing only (size factors + log1p [+ scale], "
                                  "dca/io.py:99-109): adata.X is not (log1p(raw / size_factors) - mean_g) / std_g")
    eng.set_input_transform(mean, std, True, True)
    packed = dio.pack_counts(Ytr, "auto", batch=batch_size)
    sf_h = pin_near_gpu(torch.from_numpy(sftr.astype(np.float32)), dev.index or 0)
    Xv = _to_device(X[va_lo:va_hi], eng.x_dtype, dev) if n_va else None
    Yv = _to_device(Yh[va_lo:va_hi], torch.float32, dev) if n_va else None
    sfv = _to_device(sf[va_lo:va_hi], torch.float32, dev) if n_va else None
    if world > 1:
        D.broadcast_(eng.params, src=0); D.broadcast_(eng.bn_state, src=0)
        eng.params_changed()
        if torch.distributed.get_backend() == "nccl":
            eng.comm_init()
    eng.reset_optimizer()
    lr = float(learning_rate)
    ctl = PlateauAndStop(lr, reduce_lr, early_stop, verbose)
    hist = History()
    nb = (n_tr + batch_size - 1) // batch_size
    gscale = 1.0 / world
    torch.cuda.synchronize(dev)
    prev_stream = torch.cuda.current_stream(dev)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    best_val = np.inf
    try:
        for epoch in range(epochs):
            border = np.random.permutation(nb) if shuffle else np.arange(nb)
            eng.read_epoch_acc(reset=True)
            eng.stream_begin(packed, sf_h, batch_size)
            for k in range(nb):
                eng.stream_step(int(border[k]), int(border[k + 1]) if k + 1 < nb else -1)
                if world > 1:
                    eng.allreduce_grads() if getattr(eng, "_comm", False) else D.all_reduce_sum_(eng.grads)
                eng.apply_update(ctl.lr, clip_grad, gscale)
            eng.stream_end()
            for s0 in range(0, n_va, batch_size):
                e = min(s0 + batch_size, n_va)
                eng.eval_step(Xv[s0:e], Yv[s0:e], sfv[s0:e])
            stop, best_val = _epoch_end(eng, network, hist, ctl, epoch, epochs, n_va, world, rank, dev, verbose, save_weights,
                                        output_dir, best_val)
            if stop:
                break
    finally:
        torch.cuda.synchronize(dev)
        torch.cuda.set_stream(prev_stream)
    if not hist.history["val_loss"]:
        del hist.history["val_loss"]
    return hist


def _fit_loop(eng, network, Xd, Yd, sfd, n_tr, n_va, steps, batch_size, epochs, ctl, clip_grad, gscale, world, rank, dev, hist,
              verbose, save_weights, output_dir, shuffle=True):
    best_val = np.inf
    for epoch in range(epochs):
        # Keras: np.random.shuffle(index_array) with the global NumPy RNG (seeded in api.dca / CLI)
        order = np.arange(n_tr)
        if shuffle:
            np.random.shuffle(order)
        order_d = torch.from_numpy(order.astype(np.int32)).to(dev)
        eng.read_epoch_acc(reset=True)
        for s in range(steps):
            rows = order_d[s * batch_size: min((s + 1) * batch_size, n_tr)]
            if world > 1:
                eng.train_step_allreduce(Xd, Yd, sfd, rows=rows)     # NCCL all-reduce overlapped with the backward tail
            else:
                eng.train_step(Xd, Yd, sfd, rows=rows)
            eng.apply_update(ctl.lr, clip_grad, gscale)
        # validation pass: inference-mode BN over the held-out tail
        for s in range(n_tr, n_tr + n_va, batch_size):
            e = min(s + batch_size, n_tr + n_va)
            eng.eval_step(Xd[s:e], Yd[s:e], sfd[s:e])
        stop, best_val = _epoch_end(eng, network, hist, ctl, epoch, epochs, n_va, world, rank, dev, verbose, save_weights,
                                    output_dir, best_val)
        if stop:
            break
    return hist


def train_with_args(args):
    """CLI orchestration -- dca/train.py:103-191."""
    from . import io
    from .network import AE_types

    # set seed for reproducibility                                        (dca/train.py:114-117)
    random.seed(42)
    np.random.seed(42)
    torch.manual_seed(42)
    os.environ['PYTHONHASHSEED'] = '0'

    if args.hyper:
        raise NotImplementedError("--hyper (hyperopt search, dca/hyper.py) is outside the accelerated path")

    adata = io.read_dataset(args.input,
                            transpose=(not args.transpose),  # assume gene x cell by default
                            check_counts=args.checkcounts,
                            test_split=args.testsplit)

    adata = io.normalize(adata,
                         size_factors=args.sizefactors,
                         logtrans_input=args.loginput,
                         normalize_input=args.norminput)

    if args.denoisesubset:
        genelist = list(set(io.read_genelist(args.denoisesubset)))
        assert len(set(genelist) - set(adata.var_names.values)) == 0, \
            'Gene list is not overlapping with genes from the dataset'
        output_size = len(genelist)
    else:
        genelist = None
        output_size = adata.n_vars

    hidden_size = [int(x) for x in args.hiddensize.split(',')] if args.hiddensize.strip() else []
    hidden_dropout = [float(x) for x in args.dropoutrate.split(',')]
    if len(hidden_dropout) == 1:
        hidden_dropout = hidden_dropout[0]

    assert args.type in AE_types, 'loss type not supported'
    input_size = adata.n_vars

    net = AE_types[args.type](input_size=input_size,
                              output_size=output_size,
                              hidden_size=hidden_size,
                              l2_coef=args.l2,
                              l1_coef=args.l1,
                              l2_enc_coef=args.l2enc,
                              l1_enc_coef=args.l1enc,
                              ridge=args.ridge,
                              hidden_dropout=hidden_dropout,
                              input_dropout=args.inputdropout,
                              batchnorm=args.batchnorm,
                              activation=args.activation,
                              init=args.init,
                              debug=args.debug,
                              file_path=args.outputdir)
    net.save()
    net.build()

    losses = train(adata[adata.obs.dca_split == 'train'], net,
                   output_dir=args.outputdir,
                   learning_rate=args.learningrate,
                   epochs=args.epochs, batch_size=args.batchsize,
                   early_stop=args.earlystop,
                   reduce_lr=args.reducelr,
                   output_subset=genelist,
                   optimizer=args.optimizer,
                   clip_grad=args.gradclip,
                   save_weights=args.saveweights,
                   tensorboard=args.tensorboard,
                   verbose=True)

    if genelist:
        predict_columns = adata.var_names[[np.where(adata.var_names == x)[0][0] for x in genelist]]
    else:
        predict_columns = adata.var_names

    net.predict(adata, mode='full', return_info=True)
    net.write(adata, args.outputdir, mode='full', colnames=predict_columns)
    return losses
