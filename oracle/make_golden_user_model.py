"""Golden vectors for the all-pairs user-model inference, recorded from the REFERENCE itself.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    python -m oracle.make_golden_user_model        ->  tests/golden/user_model_deepfm.npz

Builds the reference's own ``UserModel_Pairwise`` (core/user_model_pairwise.py:36-94) with the feature columns of
``CIRS-UserModel-kuaishou.py:115-123`` (user_id, photo_id, four ``feat`` slots sharing one padded embedding table,
dense photo_duration; dnn = (64, 64)), gives it "trained-like" random weights, and calls the reference's own
``KuaishouEnv.compute_normed_reward`` (environments/KuaishouRec/env/kuaishouEnv.py:113-145) on stand-in label
encoders / item frame.  Recorded: every parameter (torch layout), the inputs, the raw predictions of three users and
the normalised table.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
warnings.filterwarnings("ignore")

import pandas as pd  # noqa: E402
import torch  # noqa: E402
from deepctr_torch.inputs import DenseFeat  # noqa: E402
from core.inputs import SparseFeatP  # noqa: E402
from core.user_model_pairwise import UserModel_Pairwise  # noqa: E402
from environments.KuaishouRec.env.kuaishouEnv import KuaishouEnv  # noqa: E402


class _Classes:
    """stand-in for the fitted sklearn LabelEncoder: compute_normed_reward only reads .classes_"""

    def __init__(self, classes):
        self.classes_ = np.asarray(classes)


def build_model(v_user, v_item, v_feat, dim, dnn, seed):
    x_columns = [SparseFeatP("user_id", v_user, embedding_dim=dim),
                 SparseFeatP("photo_id", v_item, embedding_dim=dim)] + \
                [SparseFeatP("feat{}".format(i), v_feat, embedding_dim=dim, embedding_name="feat", padding_idx=0)
                 for i in range(4)] + [DenseFeat("photo_duration", 1)]
    y_columns = [DenseFeat("y", 1)]
    model = UserModel_Pairwise(x_columns, y_columns, "regression", 1, dnn_hidden_units=dnn, seed=seed, device="cpu",
                               ab_columns=None)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            std = 0.3 if "embedding" in name else (0.25 if p.dim() > 1 else 0.1)
            p.copy_(torch.randn(p.shape, generator=g) * std)
        model.embedding_dict["feat"].weight[0].zero_()   # padding_idx = 0 row stays zero (user_model.py:568-579)
    return model.eval()


def case(path, n_user=37, n_item=300, v_user=50, v_item=400, v_feat=32, dim=16, dnn=(64, 64), seed=7):
    rng = np.random.Generator(np.random.PCG64(seed))
    model = build_model(v_user, v_item, v_feat, dim, dnn, seed)
    users = np.sort(rng.choice(v_user, n_user, replace=False))
    items = np.sort(rng.choice(v_item, n_item, replace=False))
    feats = np.zeros((v_item, 4), np.int64)
    for i in range(v_item):
        k = int(rng.integers(1, 5))
        feats[i, :k] = rng.integers(1, v_feat, k)
    dur = rng.uniform(3.0, 60.0, v_item).round(3)
    df_photo_env = pd.DataFrame({"feat0": feats[:, 0], "feat1": feats[:, 1], "feat2": feats[:, 2],
                                 "feat3": feats[:, 3], "photo_duration": dur}, index=np.arange(v_item))
    df_photo_env.index.name = "photo_id"
    normed = KuaishouEnv.compute_normed_reward(model, _Classes(users), _Classes(items), df_photo_env.copy())
    raw = {}
    item_np = np.concatenate([items[:, None], feats[items], dur[items, None]], axis=1)
    for k in (0, n_user // 2, n_user - 1):
        ui = torch.tensor(np.concatenate((np.ones((n_item, 1)) * users[k], item_np), axis=1), dtype=torch.float)
        raw[k] = model.forward(ui).detach().squeeze().numpy()
    out = {"users": users.astype(np.int32), "items": items.astype(np.int32),
           "item_feat": feats[items].astype(np.int32), "item_dense": dur[items].astype(np.float32)[:, None],
           "normed_mat": normed, "raw_rows": np.array(sorted(raw), np.int32),
           "raw_pred": np.stack([raw[k] for k in sorted(raw)]), "dim": np.int32(dim)}
    for name, p in model.state_dict().items():
        out["sd." + name] = p.detach().numpy()
    np.savez_compressed(path, **out)
    print(path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    case(os.path.join(ROOT, "tests", "golden", "user_model_deepfm.npz"))
