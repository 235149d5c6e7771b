"""Drop-in for the reference's `Newcode/NewLoadData.py` (LoadData, :6-62).

Same constructor, same attributes (`Total_data`, `Train_data`, `Test_data`, `n_user`, `n_item`, `features_M`,
`positive_feedback`, `train_set`) and the same random stream (one `np.random.shuffle` on the id table, :39),
but vectorised and pandas-3 / numpy-2 safe (`applymap` :34 and the per-row Python split loop :48-58 are gone).
Extra attributes used by the accelerated sampler/evaluator: `key_cols`, `in_positive_feedback(rows)`.
"""
from collections import defaultdict

import numpy as np
import pandas as pd


class LoadData(object):
    def __init__(self, path, dataset, ratio=0.9):
        self.path = path + dataset + "/"
        self.trainfile = self.path + dataset + ".libfm"
        total = pd.read_csv(self.trainfile, sep=' ', header=None)
        cols = ['label', 'user', 'item'] + ['feature' + str(i - 2) for i in range(3, total.shape[1])]
        total.columns = cols
        self.n_user = int(total['user'].nunique())
        self.n_item = int(total['item'].nunique())

        # one global id space, ids assigned first-seen in column-major order (NewLoadData.py:29-33);
        # pd.factorize numbers values by first appearance, which is exactly that order.
        tokens = total.values[:, 1:]
        codes, uniques = pd.factorize(tokens.T.reshape(-1), sort=False)
        self.features_M = int(len(uniques))
        ids = codes.reshape(tokens.shape[1], tokens.shape[0]).T.astype(np.int64)
        data = np.empty((len(total), total.shape[1]), dtype=np.int64)
        data[:, 0] = total['label'].to_numpy()
        data[:, 1:] = ids
        # users must be 0..n_user-1 and items n_user..n_user+n_item-1 (every topk relies on it, FM.py:175); this
        # silently breaks in the reference when a user token equals an item token (datasets `last`, `ml`).
        if ids[:, 0].max() >= self.n_user or ids[:, 1].min() < self.n_user or ids[:, 1].max() >= self.n_user + self.n_item:
            raise ValueError("%s: user/item tokens collide, the item id range is not contiguous" % self.trainfile)
        self.Total_data = pd.DataFrame(data.copy(), columns=cols)

        np.random.shuffle(data)                                  # NewLoadData.py:39
        test_size = int(len(data) * (1 - ratio))                 # :40
        self.key_cols = [c for c in range(1, data.shape[1]) if c != 2]
        keys = np.ascontiguousarray(data[:, self.key_cols])
        _, first_idx, inverse = np.unique(keys, axis=0, return_index=True, return_inverse=True)
        inverse = inverse.reshape(-1)
        is_first = np.zeros(len(data), dtype=bool)
        is_first[first_idx] = True
        # a row is a test row iff it is the first occurrence of its key and fewer than test_size test rows
        # precede it (:51-54); everything else trains (:55-58)
        cand = np.flatnonzero(is_first)
        test_mask = np.zeros(len(data), dtype=bool)
        test_mask[cand[:test_size]] = True
        train = data[~test_mask]
        test = data[test_mask]

        self.positive_feedback = defaultdict(set)
        self.train_set = defaultdict(set)
        tkeys = train[:, self.key_cols]
        for key, item, user in zip(map(tuple, tkeys.tolist()), train[:, 2].tolist(), train[:, 1].tolist()):
            self.positive_feedback[key].add(item)
            self.train_set[user].add(item)

        # compact membership structure for the vectorised sampler / evaluator: sorted (key_id, item) codes
        self._key_index = {}
        tinv = inverse[~test_mask]
        self._n_keys = int(inverse.max()) + 1 if len(inverse) else 0
        self._key_lookup_keys = keys[first_idx]                  # unique keys, lexicographically sorted
        span = self.n_user + self.n_item
        self._span = span
        self._pf_codes = np.unique(tinv.astype(np.int64) * span + train[:, 2])

        self.Train_data = pd.DataFrame(train, columns=cols)
        self.Test_data = pd.DataFrame(test, columns=cols)

    # -- helpers (not in the reference) ---------------------------------------------------------------
    def key_ids(self, rows):
        """Map rows [n, F] (ids without the label column) to the id of their key, -1 if the key never trained."""
        rows = np.asarray(rows, dtype=np.int64)
        kc = [c - 1 for c in self.key_cols]
        q = np.ascontiguousarray(rows[:, kc])
        uk = self._key_lookup_keys
        if len(uk) == 0:
            return np.full(len(rows), -1, dtype=np.int64)
        # 64-bit hash of the key columns -> binary search among the sorted hashes of the known keys -> verify the columns.
        # (A lexicographic searchsorted through a structured view gives the same ids but costs 50 ms per 10^5 rows: numpy
        # compares structured elements field by field in Python-object speed.)  Rows whose hash matches a different key --
        # a collision, or a run of equal hashes -- take the structured search.
        if getattr(self, "_key_hash", None) is None:
            h = self._hash_rows(np.ascontiguousarray(uk))
            self._key_hash_order = np.argsort(h, kind="stable")
            self._key_hash = h[self._key_hash_order]
        hq = self._hash_rows(q)
        p = np.clip(np.searchsorted(self._key_hash, hq), 0, len(uk) - 1)
        cand = self._key_hash_order[p]
        hit = (uk[cand] == q).all(axis=1)
        out = np.where(hit, cand, -1).astype(np.int64)
        redo = np.nonzero(~hit & (self._key_hash[p] == hq))[0]
        if len(redo):
            out[redo] = self._key_ids_lexicographic(q[redo])
        return out

    @staticmethod
    def _hash_rows(K):
        h = np.zeros(len(K), dtype=np.uint64)
        for j in range(K.shape[1]):
            h = (h ^ (K[:, j].astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))) * np.uint64(0xBF58476D1CE4E5B9)
            h ^= h >> np.uint64(29)
        return h

    def _key_ids_lexicographic(self, q):
        uk = self._key_lookup_keys
        dt = np.dtype([("f%d" % i, np.int64) for i in range(uk.shape[1])])
        ukv = np.ascontiguousarray(uk).view(dt).reshape(-1)
        qv = np.ascontiguousarray(q).view(dt).reshape(-1)
        pos = np.searchsorted(ukv, qv)
        pos = np.clip(pos, 0, len(ukv) - 1)
        hit = ukv[pos] == qv
        return np.where(hit, pos, -1)

    def in_positive_feedback(self, rows, items=None):
        """Vectorised `item in positive_feedback[key]` for rows [n, F]; items default to the rows' own item."""
        rows = np.asarray(rows, dtype=np.int64)
        items = rows[:, 1] if items is None else np.asarray(items, dtype=np.int64)
        kid = self.key_ids(rows)
        if items.ndim == 2:
            kid = kid[:, None]
        codes = kid * self._span + items
        pos = np.searchsorted(self._pf_codes, codes)
        pos = np.clip(pos, 0, max(len(self._pf_codes) - 1, 0))
        found = (self._pf_codes[pos] == codes) if len(self._pf_codes) else np.zeros(codes.shape, bool)
        return found & (kid >= 0)
