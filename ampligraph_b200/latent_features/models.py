"""ScoringBasedEmbeddingModel facade: the reference's fit / predict / evaluate surface
(ampligraph/latent_features/models/ScoringBasedEmbeddingModel.py) re-hosted on
KGEEngine.  Only the loops and the host-side data plumbing live here; every number is
produced by libkge_b200.so.
"""
import pickle

import numpy as np
import torch

from .. import _lib
from ..datasets import DataIndexer, FilterIndex, as_triple_array
from ..engine import KGEEngine
from ..evaluation import hits_at_n_score, mr_score, mrr_score
from . import initializers, loss_functions, optimizers, regularizers
from .layers.scoring import SCORING_LAYER_REGISTRY


class History:
    """Stand-in for the Keras History callback object returned by fit()."""

    def __init__(self):
        self.history = {}
        self.epoch = []

    def _log(self, epoch, logs):
        self.epoch.append(epoch)
        for k, v in logs.items():
            self.history.setdefault(k, []).append(v)


class EarlyStopping:
    """tf.keras.callbacks.EarlyStopping subset used with fit(validation_data=...): stop when `monitor`
    (e.g. 'val_mrr') has not improved by `min_delta` for `patience` validations."""

    def __init__(self, monitor="val_mrr", min_delta=0.0, patience=0, mode="auto", restore_best_weights=False):
        self.monitor, self.min_delta, self.patience = monitor, abs(min_delta), patience
        if mode == "auto":
            mode = "min" if ("loss" in monitor or monitor.endswith("mr")) else "max"
        self.sign = 1.0 if mode == "max" else -1.0
        self.best, self.wait, self.stopped_epoch, self.model = None, 0, None, None

    def on_epoch_end(self, epoch, logs):
        if self.monitor not in logs:
            return
        v = self.sign * float(logs[self.monitor])
        if self.best is None or v > self.best + self.min_delta:
            self.best, self.wait = v, 0
        else:
            self.wait += 1
            if self.wait >= self.patience and self.model is not None:
                self.stopped_epoch = epoch
                self.model.stop_training = True


class ScoringBasedEmbeddingModel:
    """Same constructor as the reference (ScoringBasedEmbeddingModel.py:100-171)."""

    def __init__(self, eta, k, scoring_type="DistMult", seed=0, max_ent_size=None, max_rel_size=None):
        if scoring_type not in SCORING_LAYER_REGISTRY:
            raise KeyError(scoring_type)  # the reference indexes the registry directly (:146)
        self.eta, self.k, self.scoring_type, self.seed = int(eta), int(k), scoring_type, seed
        self.max_ent_size, self.max_rel_size = max_ent_size, max_rel_size
        self.scoring_layer = SCORING_LAYER_REGISTRY[scoring_type](k)
        if self.scoring_layer.kernel_id is None:
            raise NotImplementedError("scoring_type %r is the reference's random baseline: it is registered for parity of "
                                      "the registry only and has no CUDA kernel" % scoring_type)
        self.internal_k = self.scoring_layer.internal_k
        self.data_indexer = True
        self.is_fitted = False
        self.is_calibrated = False
        self.is_partitioned_training = False
        self._is_compiled = False
        self.engine = None
        self.device = 0
        self._loss_sum, self._loss_cnt = 0.0, 0  # never-reset running Mean metric (loss_functions.py:101,:224)
        self._initial_tables = None
        self._step = 0
        # set True (under torchrun, after torch.distributed.init_process_group) to train data-parallel over
        # all ranks -- every global batch is split across ranks -- and to rank against row shards
        self.distributed = False
        self._dp = None
        self.use_focusE = False

    # ------------------------------------------------------------------ compile
    def compile(self, optimizer="adam", loss=None, entity_relation_initializer="glorot_uniform",
                entity_relation_regularizer=None, **kwargs):
        """compile (:1145-1318).  entity_relation_initializer: anything tf.keras.initializers.get takes in the reference
        (a name such as 'glorot_uniform' / 'random_normal' / 'he_uniform', a {'class_name', 'config'} dict, an
        initializers.Initializer instance), an ndarray, a callable(shape)->ndarray, or a list of two of those
        [entities, relations] (EmbeddingLookupLayer.py:105-129).  entity_relation_regularizer: None | 'LP' | 'l3' |
        'l1' | 'l2' | 'l1_l2' | regularizers.LPRegularizer, or a list of two [entities, relations] (:131-155)."""
        self.optimizer = optimizers.get(optimizer)
        self.compiled_loss = loss_functions.get(loss)
        init = entity_relation_initializer
        if isinstance(init, (list, tuple)):
            assert len(init) == 2, "Incorrect length for initializer. Assumed 2 got {}".format(len(init))
            self._initializer = [initializers.get(init[0]), initializers.get(init[1])]
        else:
            self._initializer = [initializers.get(init), initializers.get(init)]
        if self.scoring_type == "RotatE":  # :1312-1315
            r = self._initializer[1]
            assert isinstance(r, initializers.Initializer) and r.name == "glorot_uniform", \
                "The relation initializer provided to a RotatE model must be glorot_uniform!"
        self._regularizer = regularizers.get_pair(entity_relation_regularizer)
        self._loss_sum, self._loss_cnt = 0.0, 0
        self._is_compiled = True
        self.engine = None  # tables are (re)built lazily on the first fit, like EmbeddingLookupLayer.build

    def _assert_compile_was_called(self):
        if not self._is_compiled:
            raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer, loss)`.")

    def is_fit(self):
        return self.is_fitted

    # ------------------------------------------------------------------ engine
    def _build_engine(self):
        name, lp = self.compiled_loss.kernel_params()
        reg = [r.kernel_params() if r is not None else None for r in self._regularizer]
        if self.scoring_type == "RotatE":
            self.scoring_layer.max_rel_size = self.max_rel_size  # :338
        def make_engine(alloc=None):
            return KGEEngine(self.scoring_type, self.k, self.eta, self.max_ent_size, self.max_rel_size,
                             loss=name or "pairwise", loss_params=lp, optimizer=self.optimizer.name,
                             optimizer_params=self.optimizer.hyperparams, regularizer=reg, device=self.device,
                             table_alloc=alloc)
        if self._world() > 1:
            from ..parallel import DataParallelTrainer
            self._dp = DataParallelTrainer(make_engine)
            self.engine = self._dp.eng
        else:
            self.engine = make_engine()
        dense = [None, None]
        for n, (init, rows, which) in enumerate(zip(self._initializer, (self.max_ent_size, self.max_rel_size), ("ent", "rel"))):
            if isinstance(init, initializers.Initializer):  # drawn on the device (Philox), no host table
                seed = init.seed if getattr(init, "seed", None) is not None else self.seed
                if init.name == "glorot_uniform":
                    self.engine.init_glorot_uniform(seed, only=which)
                else:
                    kind, a, b = init.spec(rows, self.internal_k)
                    self.engine.init_table(which, kind, a, b, seed)
            elif callable(init):
                dense[n] = np.asarray(init((rows, self.internal_k)), dtype=np.float32)
            else:
                dense[n] = np.asarray(init, dtype=np.float32)
        self.engine.set_embeddings(dense[0], dense[1])
        self._step = 0

    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size() if (self.distributed and dist.is_available() and dist.is_initialized()) else 1

    def _rank(self):
        import torch.distributed as dist
        return dist.get_rank() if self._world() > 1 else 0

    def _to_dev(self, a, dtype):
        """host array -> device tensor.  Pageable source, synchronous copy: these are the small per-chunk id / filter
        arrays of evaluate and the one-off upload of fit; pinning a fresh buffer per call (round 1) cost more than the
        copy.  The per-step training path stages batches through ONE persistent pinned ring (train_on_batches)."""
        return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).to(self.engine.device)

    # ------------------------------------------------------------------ indexing
    def _index(self, x, fit=False):
        x = as_triple_array(x)
        if self.data_indexer is False:  # use_indexer=False: data is already int ids
            return np.ascontiguousarray(x[:, :3], dtype=np.int32)
        if fit and not isinstance(self.data_indexer, DataIndexer):
            self.data_indexer = DataIndexer(x[:, :3])
        return self.data_indexer.get_indexes(x[:, :3], "t", "raw2ind")

    def get_indexes(self, X, type_of="t", order="raw2ind"):
        return self.data_indexer.get_indexes(X, type_of, order)

    def get_count(self, concept_type="e"):
        if concept_type == "e":
            return self.data_indexer.get_entities_count()
        if concept_type == "r":
            return self.data_indexer.get_relations_count()
        raise ValueError("Invalid Concept Type (expected 'e' or 'r')")

    # ------------------------------------------------------------------ fit
    def fit(self, x=None, batch_size=1000, epochs=100, verbose=True, callbacks=None, validation_split=0.0,
            validation_data=None, shuffle=True, initial_epoch=0, validation_batch_size=10,
            validation_corrupt_side="s,o", validation_freq=10, validation_burn_in=0, validation_filter=False,
            validation_entities_subset=None, partitioning_k=1, focusE=False, focusE_params={}):
        """fit (:544-883).  `shuffle` is accepted and ignored exactly like the reference
        (batches are the sequential slices of graph_data_loader.py:495-500)."""
        self._assert_compile_was_called()
        if partitioning_k != 1:
            raise NotImplementedError("bucket partitioning is replaced by HBM-resident tables (DESIGN.md)")
        x = as_triple_array(x)
        # FocusE (:342-368, :396-406, :468-543): numeric edge values in columns 3.. re-weight the scores
        self.use_focusE = bool(focusE) and x.shape[1] > 3
        if x.shape[1] > 3 and not focusE:
            print("Data shape is {}: not only triples were given, but focusE is not active!".format(x.shape[1]))
        if self.use_focusE:
            assert isinstance(focusE_params, dict), "focusE parameters need to be in a dict!"
            from .torch_losses import focuse_non_linearity
            self._focuse_nl = focuse_non_linearity(focusE_params.get("non_linearity", "linear"))
            self._focuse_stop = focusE_params.get("stop_epoch", 251)
            assert self._focuse_stop >= 0, "Invalid value for focusE stop_epoch: expected a value >=0 but got {}".format(self._focuse_stop)
            self._focuse_sw = focusE_params.get("structural_wt", 0.001)
            assert 0 <= self._focuse_sw <= 1, "Invalid focusE 'structural_wt' passed! It has to belong to [0,1]."
        triples = self._index(x, fit=True)
        if self.data_indexer is not False:
            self.max_ent_size = self.data_indexer.get_entities_count()
            self.max_rel_size = self.data_indexer.get_relations_count()
        elif self.max_ent_size is None or self.max_rel_size is None:
            self.max_ent_size = int(max(triples[:, 0].max(), triples[:, 2].max())) + 1
            self.max_rel_size = int(triples[:, 1].max()) + 1
        if self.engine is None:
            self._build_engine()
        eng = self.engine
        eng.set_hot_entities(triples=triples)  # the most frequent entities of the training set: their gradient rows are summed per warp
        data = self._to_dev(triples, np.int32)  # uploaded once; batches are device-side slices
        n = data.shape[0]
        weights = None
        if self.use_focusE:
            if len(triples) != len(x):
                raise ValueError("focusE needs every input triple to be indexable")
            weights = self._to_dev(np.asarray(x[:, 3:], dtype=np.float32), np.float32)
        batch_size = int(batch_size)
        history = History()
        self.stop_training = False
        user_loss = isinstance(self.compiled_loss, loss_functions.LossFunctionWrapper)
        world, rank = self._world(), self._rank()
        if world > 1 and (user_loss or self.use_focusE):
            raise NotImplementedError("user-callable losses / focusE are single-GPU only")
        for epoch in range(initial_epoch, epochs):
            if self.use_focusE and self._focuse_stop > 0:  # linear decay of the structural weight (:536-543)
                self._focuse_sw = max(1 - epoch / self._focuse_stop, 0.001)
            for start in range(0, n, batch_size):
                batch = data[start:start + batch_size]
                if self.use_focusE:
                    self._two_phase_step(batch, weights[start:start + batch_size])
                elif world > 1:  # this rank's slice of the global batch; the step is the global-batch step
                    from ..parallel import row_shard
                    lo, hi = row_shard(batch.shape[0], world, rank)
                    self._dp.train_step(batch[lo:hi], None, seed=self.seed + 7919 * rank, step=self._step)
                elif user_loss:
                    self._user_loss_step(batch)
                else:
                    eng.forward_backward(batch, None, seed=self.seed, step=self._step)
                    eng.apply_gradients()
                self._step += 1
                self._loss_cnt += 1
            # per-batch losses accumulate on the device; one read-back per epoch (the logged value is
            # the reference's never-reset running mean of per-batch SUM losses)
            if world > 1:
                self._dp.reduce_loss_()
            self._loss_sum += eng.read_loss()
            logs = {"loss": self._loss_sum / max(self._loss_cnt, 1)}
            validate = (epoch >= (validation_burn_in - 1) and validation_data is not None
                        and (epoch + 1) % validation_freq == 0)
            if validate:
                self.is_fitted = True
                ranks = self.evaluate(validation_data, batch_size=validation_batch_size or batch_size, verbose=False,
                                      use_filter=validation_filter, corrupt_side=validation_corrupt_side,
                                      entities_subset=validation_entities_subset, dataset_type="valid")
                logs.update({"val_mrr": mrr_score(ranks), "val_mr": mr_score(ranks),
                             "val_hits@1": hits_at_n_score(ranks, 1), "val_hits@10": hits_at_n_score(ranks, 10),
                             "val_hits@100": hits_at_n_score(ranks, 100)})
            history._log(epoch, logs)
            for cb in (callbacks or []):  # minimal Keras callback protocol: on_epoch_end + model.stop_training
                if getattr(cb, "model", None) is None:
                    try:
                        cb.model = self
                    except Exception:
                        pass
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            if verbose:
                print("Epoch %d/%d - " % (epoch + 1, epochs) + " - ".join("%s: %.4f" % kv for kv in logs.items()))
            if self.stop_training:
                break
        self.is_fitted = True
        self.history = history
        return history

    def train_on_batch(self, x):
        """Keras Model.train_on_batch (the reference model is a tf.keras.Model): one train_step on ONE
        host batch -- H2D copy of the batch, fused step, D2H read of the batch loss, which is returned.
        x: [B,3] raw triples (numpy) or an int32 torch tensor of already indexed ids (pinned for speed)."""
        self._assert_compile_was_called()
        if torch.is_tensor(x):
            host = x
        else:
            host = torch.as_tensor(self._index(x, fit=self.engine is None))
        if self.engine is None:
            if self.data_indexer is not False:
                self.max_ent_size = self.data_indexer.get_entities_count()
                self.max_rel_size = self.data_indexer.get_relations_count()
            self._build_engine()
        batch = host.to(self.engine.device, non_blocking=True)
        if self._world() > 1:  # this rank's share of the global batch; gradients are exchanged inside the step
            self._dp.train_step(batch, None, seed=self.seed + 7919 * self._rank(), step=self._step)
        elif isinstance(self.compiled_loss, loss_functions.LossFunctionWrapper):
            self._user_loss_step(batch)
        else:
            self.engine.forward_backward(batch, None, seed=self.seed, step=self._step)
            self.engine.apply_gradients()
        self._step += 1
        if self._world() > 1:
            self._dp.reduce_loss_()  # the loss of the GLOBAL batch, identical on every rank
        loss = self.engine.read_loss()
        self._loss_sum += loss
        self._loss_cnt += 1
        self.is_fitted = True
        return loss

    def train_on_batches(self, batches, prefetch=2):
        """Streamed training from HOST memory: one train_step per batch of `batches` (an iterable of [B,3] int32 id
        arrays, numpy or torch; the tf.data generator path of the reference's fit, graph_data_loader.py:472-523,:760-766
        with its prefetch(2)).  Every step copies ITS batch host->device and reads ITS loss device->host, but nothing
        blocks: batch i+1 travels on a copy stream through a persistent pinned ring while step i runs, and the 16-byte
        loss record of step i is copied asynchronously and read one step later.  Returns the list of per-batch losses."""
        self._assert_compile_was_called()
        it = iter(batches)
        first = next(it, None)
        if first is None:
            return []
        as_host = lambda x: x if torch.is_tensor(x) else torch.as_tensor(np.ascontiguousarray(x, dtype=np.int32))
        first = as_host(first)
        if self.engine is None:
            if self.max_ent_size is None or self.max_rel_size is None:
                raise ValueError("train_on_batches needs max_ent_size/max_rel_size (ids are already indexed)")
            self._build_engine()
        eng = self.engine
        if getattr(self, "hot_entities_from", None) is not None:  # streamed training sees no training set to count: an indexed
            eng.set_hot_entities(triples=self.hot_entities_from)   # [N,3] array may be attached for the hot-entity hint of fit()
            self.hot_entities_from = None
        dev = eng.device
        user_loss = isinstance(self.compiled_loss, loss_functions.LossFunctionWrapper)
        depth = max(2, int(prefetch))
        cap = first.shape[0]
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_pipe", None) is None or self._pipe["cap"] < cap or self._pipe["depth"] != depth:
            self._pipe = {"cap": cap, "depth": depth, "copy": torch.cuda.Stream(dev),
                          "pin": [torch.empty((cap, 3), dtype=torch.int32).pin_memory() for _ in range(depth)],
                          "dev": [torch.empty((cap, 3), dtype=torch.int32, device=dev) for _ in range(depth)],
                          "ready": [torch.cuda.Event() for _ in range(depth)], "free": [torch.cuda.Event() for _ in range(depth)],
                          "loss_pin": torch.zeros((depth, 2), dtype=torch.float64).pin_memory(),
                          "loss_ev": [torch.cuda.Event() for _ in range(depth)]}
            for e in self._pipe["free"]:
                e.record(main)
        P = self._pipe
        copy = P["copy"]

        def stage(i, hb):  # host batch -> (pinned ring slot ->) device ring slot, on the copy stream; never blocks on a STEP
            s = i % depth
            n = hb.shape[0]
            if n > P["cap"]:
                raise ValueError("batch of %d positives exceeds the first batch (%d) this pipeline was sized for" % (n, P["cap"]))
            if hb.is_pinned():
                src = hb
            else:
                P["ready"][s].synchronize()  # the previous H2D copy OUT of this pinned slot has finished (not the step)
                P["pin"][s][:n].copy_(hb)
                src = P["pin"][s][:n]
            copy.wait_event(P["free"][s])    # stream-ordered: the step that read this device slot last is done
            with torch.cuda.stream(copy):
                P["dev"][s][:n].copy_(src, non_blocking=True)
                P["ready"][s].record(copy)
            return s, n

        losses, pending = [], []  # pending: (slot, cumulative-loss record not yet read)
        eng.loss_acc.zero_()  # the accumulator is cumulative from here on; per-step losses are differences of records
        prev_total = 0.0
        queued = [stage(0, first)]
        i = 0
        while queued:
            nxt = next(it, None)
            if nxt is not None:
                queued.append(stage(i + 1, as_host(nxt)))  # overlaps step i
            s, n = queued.pop(0)
            main.wait_event(P["ready"][s])
            batch = P["dev"][s][:n]
            if self._world() > 1:
                self._dp.train_step(batch, None, seed=self.seed + 7919 * self._rank(), step=self._step)
            elif user_loss:
                self._user_loss_step(batch)
            else:
                eng.forward_backward(batch, None, seed=self.seed, step=self._step)
                eng.apply_gradients()
            P["free"][s].record(main)
            P["loss_pin"][s].copy_(eng.loss_acc, non_blocking=True)  # D2H of this step's (cumulative) loss record
            P["loss_ev"][s].record(main)
            pending.append(s)
            if len(pending) >= depth:  # read the record of step i-depth+1 (long complete) before its slot is reused
                ps = pending.pop(0)
                P["loss_ev"][ps].synchronize()
                tot = float(P["loss_pin"][ps].sum())
                losses.append(tot - prev_total)
                prev_total = tot
            self._step += 1
            self._loss_cnt += 1
            i += 1
        for ps in pending:
            P["loss_ev"][ps].synchronize()
            tot = float(P["loss_pin"][ps].sum())
            losses.append(tot - prev_total)
            prev_total = tot
        eng.loss_acc.zero_()
        self._loss_sum += sum(losses)
        self.is_fitted = True
        return losses

    def _user_loss_step(self, batch):
        self._two_phase_step(batch, None)

    def _two_phase_step(self, batch, weights):
        """LossFunctionWrapper / FocusE path: scores from the kernel, dL/dscore from torch autograd
        over a handful of elementwise ops, gradients from the kernel again."""
        eng, B = self.engine, batch.shape[0]
        neg = eng.generate_corruptions(batch, self.seed, self._step)
        tiled = batch.repeat(self.eta, 1)
        keep = (neg[:, 0] == tiled[:, 0]).to(torch.uint8).contiguous()
        repl = torch.where(keep.bool(), neg[:, 2], neg[:, 0]).contiguous()
        sp = torch.empty(B, dtype=torch.float32, device=eng.device)
        sn = torch.empty(B * self.eta, dtype=torch.float32, device=eng.device)
        eng.forward_backward(batch, (repl, keep), seed=self.seed, step=self._step, mode=_lib.STEP_FORWARD_ONLY,
                             scores_pos=sp, scores_neg=sn)
        spg, sng = sp.requires_grad_(True), sn.requires_grad_(True)
        fp, fn = spg, sng
        if weights is not None:  # compute_focusE_weights (:342-368) + score re-weighting (:396-406)
            w = weights.mean(1)
            sw = self._focuse_sw
            fp = self._focuse_nl(spg) * (sw + (1 - sw) * (1 - w))
            fn = self._focuse_nl(sng) * (sw + (1 - sw) * w.repeat(self.eta))
        if isinstance(self.compiled_loss, loss_functions.LossFunctionWrapper):
            loss = self.compiled_loss._user_losses(fp, fn.reshape(self.eta, -1)).sum()
        else:
            from .torch_losses import per_positive_loss
            name, lp = self.compiled_loss.kernel_params()
            loss = per_positive_loss(name, fp, fn.reshape(self.eta, -1), lp).sum()
        loss.backward()
        eng.forward_backward(batch, (repl, keep), seed=self.seed, step=self._step, mode=_lib.STEP_BACKWARD_EXT,
                             dpos=spg.grad.contiguous(), dneg=sng.grad.contiguous())  # step: the lazy rule's row stamp
        eng.apply_gradients()
        eng.loss_acc[0] += loss.detach().double()

    # ------------------------------------------------------------------ predict
    def predict(self, x, batch_size=32, verbose=0, callbacks=None):
        """predict (:1736-1823): fp32 scores; triples with unknown labels are dropped."""
        if not self.is_fitted:
            raise RuntimeError("Model has not been fitted.")
        t = self._index(x)
        if len(t) == 0:
            return np.zeros(0, np.float32)
        return self.engine.score(self._to_dev(t, np.int32)).cpu().numpy()

    # ------------------------------------------------------------------ evaluate
    def evaluate(self, x=None, batch_size=10, verbose=True, use_filter=False, corrupt_side="s,o",
                 entities_subset=None, ranking_strategy="worst", callbacks=None, dataset_type="test"):
        """evaluate (:1516-1692) -> int32 ranks [n, 2] for 's,o', [n, 1] otherwise.
        `batch_size` only lower-bounds the device chunk: results do not depend on it."""
        assert corrupt_side in ["s", "o", "s,o", "s+o"], "Invalid value for corrupt_side"
        assert ranking_strategy in ["best", "middle", "worst"], "Invalid value for ranking_strategy"
        if not self.is_fitted:
            raise RuntimeError("Model has not been fitted.")
        eng = self.engine
        t = self._index(x)
        n = len(t)
        sides = [s for s in ("s", "o") if s in corrupt_side]
        # candidates: all entities or a subset (:1634-1643)
        cand_ids, position_of = None, None
        if entities_subset is not None:
            sub = self.get_indexes(np.asarray(entities_subset), "e") if self.data_indexer is not False \
                else np.asarray(entities_subset, np.int32)
            cand_ids = self._to_dev(sub, np.int32)
            position_of = np.full(self.max_ent_size, -1, np.int64)
            position_of[sub] = np.arange(len(sub))
        # filters: dict of datasets (train/valid/test), or True = the evaluated data itself
        findex = None
        if isinstance(use_filter, dict) or use_filter is True:
            parts = [t] if use_filter is True else [self._index(v) for v in use_filter.values()]
            findex = FilterIndex(np.concatenate(parts), self.max_ent_size)
        out = np.zeros((n, len(sides)), np.int32)
        chunk = max(int(batch_size), 4096)
        for start in range(0, n, chunk):
            tb = t[start:start + chunk]
            td = self._to_dev(tb, np.int32)
            for j, side in enumerate(sides):
                off = idx = None
                if findex is not None:
                    o_np, i_np = findex.lookup(tb, side, position_of)
                    off, idx = self._to_dev(o_np, np.int64), self._to_dev(i_np, np.int32)
                if self._world() > 1 and cand_ids is None:
                    # row-sharded candidates: RAW counters summed over ranks, tie strategy applied once ('middle' is not
                    # additive over partitions, AbstractScoringLayer.py:232-244)
                    from ..parallel import allreduce_sum_, row_shard
                    lo, hi = row_shard(self.max_ent_size, self._world(), self._rank())
                    cnt = torch.zeros((len(tb), 3), dtype=torch.int32, device=eng.device)
                    eng.rank(td, side, ranking_strategy, off, idx, cand_begin=lo, n_cand=hi - lo, counts=cnt)
                    allreduce_sum_([cnt])
                    r = eng.finalize_ranks(cnt, ranking_strategy)
                else:
                    r = eng.rank(td, side, ranking_strategy, off, idx, cand_ids=cand_ids)
                out[start:start + len(tb), j] = r.cpu().numpy()
        if corrupt_side == "s+o":  # :1459-1463 sum BEFORE the +1
            out = out.sum(1, keepdims=True)
        return (out + 1).astype(np.int32)  # :1684

    # ------------------------------------------------------------------ calibration
    def calibrate(self, X_pos, X_neg=None, positive_base_rate=None, batch_size=32, epochs=50, verbose=0):
        """calibrate (:1922-2122) with CalibrationLayer (layers/calibration/calibrate.py:11-129): Platt scaling,
        two scalars (w, b) fitted with Adam on the kernel's scores of positives and negatives (given, or
        generated corruptions when only positive_base_rate is given)."""
        if not self.is_fitted:
            raise RuntimeError("Model has not been fitted.")
        self.is_calibrated = False
        pos = self._to_dev(self._index(X_pos), np.int32)
        pos_size = pos.shape[0]
        with_corruption = X_neg is None
        if with_corruption:
            assert positive_base_rate is not None, "Please provide the negatives or positive base rate!"
            neg, neg_size = None, pos_size
        else:
            neg = self._to_dev(self._index(X_neg), np.int32)
            neg_size = neg.shape[0]
            if positive_base_rate is None:
                positive_base_rate = pos_size / (pos_size + neg_size)
        if positive_base_rate is not None and (positive_base_rate <= 0 or positive_base_rate >= 1):
            raise ValueError("positive_base_rate must be a value between 0 and 1.")
        dev = self.engine.device
        w = torch.zeros((), device=dev, requires_grad=True)  # calibrate.py:60-63
        b = torch.tensor(float(np.log((neg_size + 1.0) / (pos_size + 1.0))), device=dev, requires_grad=True)
        opt = torch.optim.Adam([w, b], lr=1e-3, eps=1e-7)
        n_batches = int(np.ceil(pos_size / batch_size))
        nb_size = int(np.ceil(neg_size / n_batches)) if not with_corruption else None
        step = 0
        for _ in range(epochs):
            for i in range(n_batches):
                pb = pos[i * batch_size:(i + 1) * batch_size].contiguous()
                sp = self.engine.score(pb)
                if with_corruption:  # ONE corruption per positive (eta=1, :1886): the j=0 block of the tile-ordered tensor
                    sn = self.engine.score(self.engine.generate_corruptions(pb, self.seed, step)[:pb.shape[0]].contiguous())
                else:
                    sn = self.engine.score(neg[i * nb_size:(i + 1) * nb_size].contiguous())
                step += 1
                scores = torch.cat([sp, sn])
                logits = -(w * scores + b)
                labels = torch.cat([torch.full_like(sp, (pos_size + 1.0) / (pos_size + 2.0)),
                                    torch.full_like(sn, 1.0 / (neg_size + 2.0))])
                weights = torch.cat([torch.full_like(sp, sn.shape[0] / max(sp.shape[0], 1)),
                                     torch.full_like(sn, (1.0 - positive_base_rate) / positive_base_rate)])
                loss = (weights * torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="none")).mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
        self.calib_w, self.calib_b = float(w.detach()), float(b.detach())
        self.is_calibrated = True

    def predict_proba(self, x, batch_size=32, verbose=0, callbacks=None):
        """predict_proba (:2124-2212): sigmoid(-(w*score + b)) of the calibrated model."""
        if not self.is_calibrated:
            raise RuntimeError("Model has not been calibrated. Please call `model.calibrate(...)` before predicting probabilities.")
        sc = self.predict(x, batch_size=batch_size)
        return (1.0 / (1.0 + np.exp(self.calib_w * sc + self.calib_b))).astype(np.float32)

    # ------------------------------------------------------------------ embeddings / weights
    def get_embeddings(self, entities, embedding_type="e"):
        """get_embeddings (:2214): rows of the dense [rows, internal_k] tables."""
        if not self.is_fitted:
            raise RuntimeError("Model has not been fitted.")
        if embedding_type not in ("e", "r"):
            raise ValueError("Invalid entity type: %s" % embedding_type)
        ent, rel = self.engine.get_embeddings()
        idx = self.get_indexes(np.asarray(entities), embedding_type) if self.data_indexer is not False \
            else np.asarray(entities, np.int64)
        table = ent if embedding_type == "e" else rel
        return table[torch.as_tensor(idx, dtype=torch.long, device=table.device)].cpu().numpy()

    def save_weights(self, filepath):
        """Tables + optimizer slots + indexer in one pickle (the reference's Keras format is TF-specific)."""
        ent, rel = self.engine.get_embeddings()
        state = {"ent": ent.cpu().numpy(), "rel": rel.cpu().numpy(), "t": self.engine.t, "step": self._step,
                 "slots": {k: [None if s is None else s.cpu().numpy() for s in v] for k, v in self.engine.slots.items()},
                 "indexer": self.data_indexer, "max_ent_size": self.max_ent_size, "max_rel_size": self.max_rel_size}
        with open(filepath, "wb") as f:
            pickle.dump(state, f)

    def load_weights(self, filepath):
        self._assert_compile_was_called()
        with open(filepath, "rb") as f:
            state = pickle.load(f)
        self.data_indexer = state["indexer"]
        self.max_ent_size, self.max_rel_size = state["max_ent_size"], state["max_rel_size"]
        self._build_engine()
        self.engine.set_embeddings(state["ent"], state["rel"])
        self.engine.t, self._step = state["t"], state["step"]
        for k, v in state["slots"].items():
            for n, s in enumerate(v):
                if s is not None:
                    self.engine.slots[k][n].copy_(torch.as_tensor(s))
        self.is_fitted = True
