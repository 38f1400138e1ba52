"""Host-side data plumbing either side of the hot path (vectorised numpy/pandas).

* DataIndexer  -- label <-> int32 id mapping with the reference's first-seen order
  (ampligraph/datasets/data_indexer.py:373-428: scan rows, subject then object for
  entities; predicates separately), unknown labels dropped on lookup (:525-542).
* FilterIndex  -- known-true entities per (p,o) / (s,p) as CSR, built ONCE, replacing
  the per-batch pandas groupby/reindex + tf.ragged.constant of
  ampligraph/datasets/graph_data_loader.py:287-350, :382-439, :501-521.
"""
import numpy as np
import pandas as pd


def as_triple_array(x, sep="\t"):
    """Accept what the reference's DataSourceIdentifier accepts for an in-memory run
    (ampligraph/datasets/source_identifier.py:25-50, :134-136): a numpy array / nested list, a pandas
    DataFrame, or the path of a csv / txt / gz file with one `sep`-separated triple per line and no header."""
    if isinstance(x, str):
        ext = x.rsplit(".", 1)[-1].lower() if "." in x else ""
        if ext not in ("csv", "txt", "gz"):
            raise ValueError("Unsupported data source file type: %r (expected csv, txt or gz)" % x)
        return pd.read_csv(x, sep=sep, header=None).values
    if isinstance(x, pd.DataFrame):
        return x.values
    return np.asarray(x)


class DataIndexer:
    def __init__(self, X=None):
        self.ent_labels = np.empty(0, dtype=object)
        self.rel_labels = np.empty(0, dtype=object)
        self._ent_index = pd.Index([])
        self._rel_index = pd.Index([])
        if X is not None:
            self.update(X)

    # ids are assigned in first-seen order over s0,o0,s1,o1,... and p0,p1,...
    def update(self, X):
        X = np.asarray(X)
        ents = pd.unique(np.stack([X[:, 0], X[:, 2]], axis=1).ravel())
        rels = pd.unique(X[:, 1])
        new_e = ents[~pd.Index(ents).isin(self._ent_index)] if len(self._ent_index) else ents
        new_r = rels[~pd.Index(rels).isin(self._rel_index)] if len(self._rel_index) else rels
        self.ent_labels = np.concatenate([self.ent_labels, np.asarray(new_e, dtype=object)])
        self.rel_labels = np.concatenate([self.rel_labels, np.asarray(new_r, dtype=object)])
        self._ent_index = pd.Index(self.ent_labels)
        self._rel_index = pd.Index(self.rel_labels)

    def get_entities_count(self):
        return len(self.ent_labels)

    def get_relations_count(self):
        return len(self.rel_labels)

    def get_indexes(self, X, type_of="t", order="raw2ind"):
        """data_indexer.get_indexes: triples / entities / relations, raw->ind or ind->raw."""
        if type_of not in ("t", "e", "r"):
            raise ValueError("type_of must be 't', 'e' or 'r'")
        X = np.asarray(X)
        if order == "raw2ind":
            if type_of == "t":
                s = self._ent_index.get_indexer(X[:, 0])
                p = self._rel_index.get_indexer(X[:, 1])
                o = self._ent_index.get_indexer(X[:, 2])
                ok = (s >= 0) & (p >= 0) & (o >= 0)
                if not ok.all():
                    print("\n%d triples containing invalid keys skipped!\n" % int((~ok).sum()))
                return np.stack([s[ok], p[ok], o[ok]], axis=1).astype(np.int32)
            idx = (self._ent_index if type_of == "e" else self._rel_index).get_indexer(X.ravel())
            return idx[idx >= 0].astype(np.int32)
        if order == "ind2raw":
            if type_of == "t":
                X = X.astype(np.int64)
                return np.stack([self.ent_labels[X[:, 0]], self.rel_labels[X[:, 1]], self.ent_labels[X[:, 2]]], axis=1)
            lab = self.ent_labels if type_of == "e" else self.rel_labels
            return lab[X.astype(np.int64).ravel()]
        raise Exception("No such order available options: ind2raw, raw2ind, instead got {}.".format(order))


class FilterIndex:
    """CSR of known-true subjects per (p,o) and objects per (s,p) over a set of indexed triples."""

    def __init__(self, indexed_triples, n_ent):
        t = np.unique(np.asarray(indexed_triples, dtype=np.int64).reshape(-1, 3), axis=0)
        self.n_ent = int(n_ent)
        self._sub = self._group(t[:, 1] * self.n_ent + t[:, 2], t[:, 0])  # (p,o) -> subjects
        self._obj = self._group(t[:, 1] * self.n_ent + t[:, 0], t[:, 2])  # (s,p) -> objects

    @staticmethod
    def _group(keys, vals):
        order = np.lexsort((vals, keys))
        return keys[order], vals[order].astype(np.int32)

    def lookup(self, triples, side, position_of=None):
        """-> (offsets int64 [b+1], ids int32 [nnz]) for a batch of indexed test triples.
        position_of: optional int array mapping entity id -> candidate position (or -1) when an
        entities_subset is used (the mapping_dict of AbstractScoringLayer.py:266-275)."""
        t = np.asarray(triples, dtype=np.int64)
        if side == "s":
            keys, vals = self._sub
            q = t[:, 1] * self.n_ent + t[:, 2]
        else:
            keys, vals = self._obj
            q = t[:, 1] * self.n_ent + t[:, 0]
        lo = np.searchsorted(keys, q, side="left")
        hi = np.searchsorted(keys, q, side="right")
        lens = hi - lo
        off = np.zeros(len(t) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        total = int(off[-1])
        # gather the ranges [lo_i, hi_i) without a Python loop
        idx = np.repeat(lo - off[:-1], lens) + np.arange(total, dtype=np.int64)
        ids = vals[idx] if total else np.zeros(0, np.int32)
        if position_of is not None and total:
            pos = position_of[ids]
            keep = pos >= 0
            row = np.repeat(np.arange(len(t)), lens)[keep]
            ids = pos[keep].astype(np.int32)
            off = np.zeros(len(t) + 1, dtype=np.int64)
            np.cumsum(np.bincount(row, minlength=len(t)), out=off[1:])
        return off, np.ascontiguousarray(ids, dtype=np.int32)
