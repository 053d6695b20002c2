"""Test-time metrics of CIRS computed from the collectors' replay buffers (evaluation.py:10-77, 286-371):
coverage ``CV`` (distinct recommended items / catalogue size), ``CV_turn`` (distinct items / recommendations) and the
dominated-category rate ``ifeat_feat`` per collector of a CollectorSet, with the reference's key prefixes.

The reference walks ``buffer.prev / next / last_index`` episode by episode on the host; here the same sets are read
in one shot from the env-major buffer (``sample_index(0)`` = every stored transition of the last collect).  The
buffer keeps the reference's index API too, so the reference's own callback also runs on it unchanged
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


def dominate_rate(cats, dom):
    """evaluation.py:38-47: per recommendation, the per-value match COUNTS are combined with a bitwise OR into an
    integer array that is then summed -- for items whose feature columns are distinct (the real data) this is the
    share of recommendations carrying a dominated category; the integer arithmetic is reproduced as written."""
    acc = np.zeros(len(cats), dtype=np.int64)
    for v in dom:
        acc |= (cats == v).sum(axis=1)
    return float(acc.sum() / max(len(cats), 1))


class Callback_Coverage_Count:
    """Drop-in for evaluation.py:286-371.  ``df_item_val`` may be a pandas DataFrame with ``feat*`` columns indexed by
    raw item id (as in the reference) or an int array [n_item, n_feat] indexed by encoded item id."""

    def __init__(self, test_collector_set, df_item_val=None, need_transform=False, item_feat_domination=None,
                 lbe_photo=None, top_rate=0.6):
        self.collector_dict = test_collector_set.collector_dict
        mat = test_collector_set.env.mat
        self.num_items = (mat[0] if isinstance(mat, (list, tuple)) else mat).shape[-1]
        self.df_item_val, self.need_transform = df_item_val, need_transform
        self.item_feat_domination, self.lbe_photo, self.top_rate = item_feat_domination, lbe_photo, top_rate

    def on_epoch_begin(self, epoch):
        pass

    def on_train_begin(self):
        pass

    def on_train_end(self):
        pass

    def _item_cats(self, acts):
        if self.need_transform and self.lbe_photo is not None:
            acts = self.lbe_photo.inverse_transform(acts)
        if hasattr(self.df_item_val, "loc"):
            return self.df_item_val.loc[acts].filter(regex="^feat", axis=1).to_numpy().astype(int)
        return np.asarray(self.df_item_val)[acts]

    def on_epoch_end(self, epoch, results=None, **kwargs):
        results = {} if results is None else results
        out = {}
        for name, collector in self.collector_dict.items():
            buf = collector.buffer
            acts = np.asarray(buf.act)[buf.sample_index(0)].astype(np.int64)
            hit = len(np.unique(acts))
            res = {"CV": hit / self.num_items, "CV_turn": hit / max(len(acts), 1)}
            if self.item_feat_domination is not None and "feat" in self.item_feat_domination and len(acts):
                cats = self._item_cats(acts)
                dom = dominated_values(self.item_feat_domination["feat"], self.top_rate)
                res["ifeat_feat"] = dominate_rate(cats, dom)
            out.update(res if name == "FB" else {name + "_" + k: v for k, v in res.items()})
        results.update(out)
        return results
