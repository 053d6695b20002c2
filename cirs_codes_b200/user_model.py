"""All-pairs user-model inference on the device: the producer of KuaishouEnv's ``normed_mat``.

Host mirror of ``KuaishouEnv.compute_normed_reward`` (environments/KuaishouRec/env/kuaishouEnv.py:113-145) over the
reference's DeepFM user model ``UserModel_Pairwise`` (core/user_model_pairwise.py:36-132, feature columns of
CIRS-UserModel-kuaishou.py:115-123).  The reference loops over the users in Python and runs one torch forward of
n_item rows per user; here the whole U x I table is one C-ABI call (csrc/user_model.cu, tcgen05 tensor cores) and it
can stay in HBM: ``compute_normed_reward(..., return_device=True)`` hands ``KuaishouVectorEnv(normed_mat=...)`` a CUDA
tensor without a host round trip.

No CPU fallback: without the built library / a CUDA device every function here raises ``CirsError``.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CirsError

_KEYS = {"emb_user": "embedding_dict.user_id.weight", "emb_item": "embedding_dict.photo_id.weight",
         "emb_feat": "embedding_dict.feat.weight", "lin_user": "linear.embedding_dict.user_id.weight",
         "lin_item": "linear.embedding_dict.photo_id.weight", "lin_feat": "linear.embedding_dict.feat.weight",
         "lin_dense": "linear.weight", "w1": "dnn.linears.0.weight", "b1": "dnn.linears.0.bias",
         "w2": "dnn.linears.1.weight", "b2": "dnn.linears.1.bias", "w_last": "last.weight"}


class UserModelWeights:
    """Device copy of a ``UserModel_Pairwise`` state_dict (torch tensors or numpy arrays, reference key names)."""

    def __init__(self, state_dict, device="cuda:0"):
        import torch
        _lib.require_cuda()
        _lib.load()
        sd = state_dict.state_dict() if hasattr(state_dict, "state_dict") else state_dict
        if "dnn.linears.2.weight" in sd or "dnn.linears.1.weight" not in sd:
            raise CirsError("user model: dnn_hidden_units must have exactly two layers (CIRS-UserModel-kuaishou.py:67)")
        self.device = torch.device(device)
        self.t = {}
        for k, name in _KEYS.items():
            if name not in sd:
                raise CirsError(f"user model state_dict lacks {name!r}")
            v = sd[name]
            v = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
            self.t[k] = torch.from_numpy(np.ascontiguousarray(v, np.float32)).to(self.device)
        ob = sd["out.bias"]
        self.out_bias = float(ob.detach().cpu().reshape(-1)[0] if torch.is_tensor(ob) else np.asarray(ob).reshape(-1)[0])
        self.emb_dim = int(self.t["emb_user"].shape[1])
        self.hidden = int(self.t["w1"].shape[0])
        self.n_dense = int(self.t["lin_dense"].shape[0])
        in_dim = int(self.t["w1"].shape[1])
        n_feat, rem = divmod(in_dim - self.n_dense - 2 * self.emb_dim, self.emb_dim)
        if rem or n_feat < 0 or self.t["emb_item"].shape[1] != self.emb_dim or self.t["emb_feat"].shape[1] != self.emb_dim:
            raise CirsError("user model: all sparse columns must share one embedding_dim "
                            "(args.entity_dim = args.feature_dim, CIRS-UserModel-kuaishou.py:153)")
        if tuple(self.t["w2"].shape) != (_lib.HIDDEN, _lib.HIDDEN) or self.hidden != _lib.HIDDEN:
            raise CirsError("user model: dnn_hidden_units must be (64, 64)")
        self.n_feat = int(n_feat)
        s = _lib.UserModelStruct()
        for k in _KEYS:
            setattr(s, k, _lib.ptr(self.t[k]))
        s.out_bias, s.emb_dim, s.n_feat, s.n_dense, s.hidden = self.out_bias, self.emb_dim, self.n_feat, self.n_dense, self.hidden
        self.struct = s


def _i32(x, device):
    import torch
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, np.int32)).to(device)


def predict_all(weights, users, items, item_feat, item_dense, normalise=True, out=None, return_minmax=False,
                workspace=None):
    """float32 CUDA tensor [n_user, n_item]: UserModel_Pairwise.forward on every (user, item) pair, min-max
    normalised over the table when ``normalise`` (kuaishouEnv.py:139-143).  ``users`` / ``items`` are RAW ids
    (``lbe.classes_``); ``item_feat`` int [n_item, n_feat]; ``item_dense`` float [n_item, n_dense].  Stream-ordered on
    torch's current stream, no synchronisation."""
    import torch
    dev = weights.device
    u, it = _i32(users, dev), _i32(items, dev)
    n_user, n_item = int(u.numel()), int(it.numel())
    if n_user == 0 or n_item == 0:
        raise CirsError("predict_all: empty user / item list")
    if int(u.max()) >= weights.t["emb_user"].shape[0] or int(it.max()) >= weights.t["emb_item"].shape[0] \
            or int(u.min()) < 0 or int(it.min()) < 0:
        raise CirsError("predict_all: id outside the embedding vocabulary")   # torch would raise IndexError
    feat = _i32(np.asarray(item_feat).reshape(n_item, -1) if not torch.is_tensor(item_feat) else item_feat, dev) \
        if weights.n_feat else None
    if feat is not None and (tuple(feat.shape) != (n_item, weights.n_feat) or int(feat.max()) >= weights.t["emb_feat"].shape[0]
                             or int(feat.min()) < 0):
        raise CirsError("predict_all: item_feat must be [n_item, n_feat] ids inside the feat vocabulary")
    dense = None
    if weights.n_dense:
        dense = item_dense if torch.is_tensor(item_dense) else torch.from_numpy(np.ascontiguousarray(item_dense, np.float32))
        dense = dense.to(device=dev, dtype=torch.float32).reshape(n_item, weights.n_dense).contiguous()
    if out is None:
        out = torch.empty((n_user, n_item), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (n_user, n_item) or out.dtype != torch.float32 or not out.is_contiguous():
        raise CirsError("predict_all: out must be a contiguous float32 [n_user, n_item] CUDA tensor")
    need = _lib.load().cirs_user_model_workspace_bytes(n_user, n_item, weights.emb_dim)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    mm = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("cirs_user_model_predict_all", C.byref(weights.struct), n_user, _lib.ptr(u), n_item, _lib.ptr(it),
                  _lib.ptr(feat), _lib.ptr(dense), 1 if normalise else 0, _lib.ptr(out), _lib.ptr(mm),
                  _lib.ptr(workspace), _lib.stream())
    return (out, mm) if return_minmax else out


def compute_normed_reward(user_model, lbe_user, lbe_photo, df_photo_env, device="cuda:0", return_device=False,
                          feat_columns=("feat0", "feat1", "feat2", "feat3"), dense_columns=("photo_duration",)):
    """Drop-in for ``KuaishouEnv.compute_normed_reward(user_model, lbe_user, lbe_photo, df_photo_env)``
    (kuaishouEnv.py:113-145): same arguments (a torch module / state_dict, two fitted label encoders, the item frame
    indexed by photo_id), same result -- float64 numpy [n_user, n_item] -- or, with ``return_device=True``, the float32
    CUDA tensor itself (what ``KuaishouVectorEnv(normed_mat=...)`` consumes)."""
    w = user_model if isinstance(user_model, UserModelWeights) else UserModelWeights(user_model, device)
    users, items = np.asarray(lbe_user.classes_), np.asarray(lbe_photo.classes_)
    info = df_photo_env.loc[items]
    item_feat = info[list(feat_columns)].to_numpy()
    item_dense = info[list(dense_columns)].to_numpy()
    out = predict_all(w, users, items, item_feat, item_dense, normalise=True)
    return out if return_device else out.cpu().numpy().astype(np.float64)
