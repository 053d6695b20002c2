"""Test-time metrics of CIRS computed from the collectors' replay buffers (evaluation.py:10-77, 286-371):
coverage ``CV`` (distinct recommended items / catalogue size), ``CV_turn`` (distinct items / recommendations) and the
dominated-category rates ``ifeat_*`` per collector of a CollectorSet, with the reference's key prefixes.

The reference walks ``buffer.prev / next / last_index`` episode by episode on the host and looks every recommended
item up in a pandas frame.  Here the per-item part of the dominated-category arithmetic is folded ONCE into an int32
weight per catalogue item (``item_weights``), and the per-collect part -- the set of distinct items and the sum of the
weights over every stored transition -- is one device reduction over the buffer's ``d_act`` array
(``cirs_coverage_count``, csrc/util.cu) when the buffer was filled by the fused rollout; buffers filled through the
host interface (``buffer.add``) are reduced with numpy (host bookkeeping of host data, not a fallback of the kernels).
The buffer keeps the reference's index API too, so the reference's own callback also runs on it unchanged
(tests/test_host_cpu.py::test_coverage_callback_matches_reference).
"""
import numpy as np


def dominated_values(sorted_items, top_rate):
    """The most frequent category values whose cumulative share first exceeds ``top_rate`` (evaluation.py:19-33):
    ``sorted_items`` = [(value, count), ...] sorted by decreasing count."""
    counts = np.array([c for _, c in sorted_items], dtype=np.float64)
    cum = np.cumsum(counts / counts.sum())
    ind = int(np.searchsorted(cum, top_rate, side="right"))   # first index whose cumulative share is > top_rate
    return np.array([v for v, _ in sorted_items])[:max(ind, 1)]


def dominate_weights(cats, dom):
    """evaluation.py:38-47 ("feat" branch) per catalogue item: the per-value match COUNTS over the item's feature
    columns are combined with a bitwise OR into an integer -- for items whose feature columns are distinct (the real
    data) 1 iff the item carries a dominated category; the integer arithmetic is reproduced as written.
    ``cats``: int [n, n_feat].  Returns int64 [n]; the metric is ``weights[acts].sum() / len(acts)``."""
    acc = np.zeros(len(cats), dtype=np.int64)
    for v in dom:
        acc |= (cats == v).sum(axis=1)
    return acc


def dominate_rate(cats, dom):
    """The "feat" rate of a list of recommendations given their category rows (kept for callers of round 1)."""
    return float(dominate_weights(cats, dom).sum() / max(len(cats), 1))


class Callback_Coverage_Count:
    """Drop-in for evaluation.py:286-371.  ``df_item_val`` may be a pandas DataFrame with feature columns indexed by
    raw item id (as in the reference) or an int array [n_item, n_feat] indexed by encoded item id.
    ``item_feat_domination``: {"feat": [(value, count), ...]} (KuaiRec / KuaiRand: every ``feat*`` column, :19-47) or
    {feature name: [(value, count), ...], ...} (the per-feature branch, :49-77)."""

    def __init__(self, test_collector_set, df_item_val=None, need_transform=False, item_feat_domination=None,
                 lbe_photo=None, top_rate=0.6):
        self.collector_dict = test_collector_set.collector_dict
        mat = test_collector_set.env.mat
        self.num_items = (mat[0] if isinstance(mat, (list, tuple)) else mat).shape[-1]
        self.df_item_val, self.need_transform = df_item_val, need_transform
        self.item_feat_domination, self.lbe_photo, self.top_rate = item_feat_domination, lbe_photo, top_rate
        self._weights = None       # {metric key: int32 weight per ENCODED item}
        self._dev = {}

    def on_epoch_begin(self, epoch):
        pass

    def on_train_begin(self):
        pass

    def on_train_end(self):
        pass

    # ---- per catalogue item, once
    def _catalogue_rows(self):
        """Feature rows of every ENCODED item id 0 .. num_items-1 as {column name: int array}."""
        ids = np.arange(self.num_items)
        if self.need_transform and self.lbe_photo is not None:
            ids = self.lbe_photo.inverse_transform(ids)
        if hasattr(self.df_item_val, "loc"):
            frame = self.df_item_val.loc[ids]
            return {c: frame[c].to_numpy() for c in frame.columns}
        arr = np.asarray(self.df_item_val)[ids]
        return {f"feat{k}": arr[:, k] for k in range(arr.shape[1])}

    def item_weights(self):
        if self._weights is not None or self.item_feat_domination is None or self.df_item_val is None:
            return self._weights or {}
        cols = self._catalogue_rows()
        w = {}
        if "feat" in self.item_feat_domination:
            feat = np.stack([cols[c] for c in cols if str(c).startswith("feat")], axis=1).astype(int)
            dom = dominated_values(self.item_feat_domination["feat"], self.top_rate)
            w["ifeat_feat"] = dominate_weights(feat, dom)
        else:
            for name, sorted_items in self.item_feat_domination.items():
                dom = dominated_values(sorted_items, self.top_rate)
                w["ifeat_" + name] = np.isin(cols[name].astype(int), dom).astype(np.int64)   # :62-72: OR of booleans
        self._weights = w
        return w

    # ---- per collect
    def _reduce(self, buf, weights):
        """(distinct items, transitions, {key: weight sum}) over every stored transition of the last collect."""
        n = len(buf)
        if n == 0:
            return 0, 0, {k: 0 for k in weights}
        if getattr(buf, "_dev_valid", False) and getattr(buf, "_plan_ok", False) and not getattr(buf, "act_dim", 0):
            return self._reduce_device(buf, n, weights)
        acts = np.asarray(buf.act)[buf.sample_index(0)].astype(np.int64)
        return len(np.unique(acts)), n, {k: int(w[acts].sum()) for k, w in weights.items()}

    def _reduce_device(self, buf, n, weights):
        import torch
        from . import _lib
        dev = buf.device
        d = self._dev.setdefault(str(dev), {})
        if "bits" not in d:
            d["bits"] = torch.zeros((self.num_items + 31) // 32, dtype=torch.int32, device=dev)
            d["out"] = torch.zeros(3, dtype=torch.int64, device=dev)
            d["w"] = {k: torch.as_tensor(w.astype(np.int32), device=dev) for k, w in weights.items()}
        sums, hit = {}, 0
        for k in (list(weights) or [None]):
            _lib.call("cirs_coverage_count", n, _lib.ptr(buf.d_index), _lib.ptr(buf.d_act), self.num_items,
                      _lib.ptr(d["w"][k]) if k is not None else None, _lib.ptr(d["bits"]), _lib.ptr(d["out"]),
                      _lib.stream())
            o = d["out"].cpu().numpy()
            hit = int(o[0])
            if k is not None:
                sums[k] = int(o[2])
        return hit, n, sums

    def on_epoch_end(self, epoch, results=None, **kwargs):
        results = {} if results is None else results
        weights = self.item_weights()
        out = {}
        for name, collector in self.collector_dict.items():
            hit, n, sums = self._reduce(collector.buffer, weights)
            res = {"CV": hit / self.num_items, "CV_turn": hit / max(n, 1)}
            for k, s in sums.items():
                res[k] = s / max(n, 1)
            out.update(res if name == "FB" else {name + "_" + k: v for k, v in res.items()})
        results.update(out)
        return results
