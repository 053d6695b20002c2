"""CPU-side tests (no GPU needed): the C-ABI library loads and exports every symbol include/cirs_b200.h declares,
the ctypes binding covers the header, and the host-side logic (parameter packing, replay-buffer index arithmetic,
minibatch splitting, feature columns) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "cirs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cirs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from cirs_codes_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cirs_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes and header disagree"
    lib.cirs_abi_version.restype = ctypes.c_int
    assert lib.cirs_abi_version() == _lib.ABI_VERSION


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import cirs_codes_b200 as cb
    from cirs_codes_b200._lib import CirsError
    with pytest.raises(CirsError):
        cb.KuaishouVectorEnv(2, np.zeros((3, 4)), [[1]] * 4, normed_mat=np.zeros((3, 4)))
    # the reward-table producer has no CPU path either
    from cirs_codes_b200 import user_model as um
    from oracle import user_model as oum

    class Lbe:
        classes_ = np.arange(3)
    with pytest.raises(CirsError):
        um.UserModelWeights(oum.synth_params(4, 5, 8))
    with pytest.raises(CirsError):
        um.compute_normed_reward(oum.synth_params(4, 5, 8), Lbe(), Lbe(), None)


def test_struct_sizes_match_c_layout():
    """sizeof() of the ctypes mirrors against the sizes the C compiler reports (compiled on the fly with gcc)."""
    import subprocess
    import tempfile
    from cirs_codes_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "cirs_b200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(cirs_user_model), sizeof(cirs_kuaishou_env), sizeof(cirs_encoder_layer),
  sizeof(cirs_tracker_weights), sizeof(cirs_policy_weights), sizeof(cirs_ppo_config)); return 0;}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                        os.path.join(d, "t")], check=True)
        got = [int(x) for x in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(s) for s in (_lib.UserModelStruct, _lib.KuaishouEnvStruct, _lib.EncoderLayerStruct, _lib.TrackerWeightsStruct,
                                       _lib.PolicyWeightsStruct, _lib.PPOConfigStruct)]
    assert got == want


def test_param_pack_roundtrip_policy_and_tracker():
    from cirs_codes_b200 import params, net
    torch.manual_seed(0)
    n = net.Net(20, hidden_sizes=[64, 64])
    a, c = net.Actor(n, 300), net.Critic(n)
    L = params.policy_layout(20, 300)
    flat = L.pack(params.policy_sd_from_reference(a.state_dict(), c.state_dict()), "cpu")
    assert flat.numel() == L.total and L.segs["actor.last.weight"].ld == 384 and L.n_trunk % 32 == 0
    a2, c2 = params.policy_sd_to_reference(L.unpack(flat))
    for k, v in a.state_dict().items():
        assert torch.equal(a2[k], v), k
    for k, v in c.state_dict().items():
        assert torch.equal(c2[k], v), k
    # k-major storage: Wt[in][ld]
    w = a.state_dict()["last.model.0.weight"]
    seg = L.segs["actor.last.weight"]
    view = flat[seg.offset:seg.offset + 64 * seg.ld].view(64, seg.ld)
    assert torch.equal(view[:, :300], w.t()) and torch.all(view[:, 300:] == 0)
    T = params.tracker_layout(32, 4, 128, 2, 20, 31, n_user=7, n_item=9)
    enc = torch.nn.TransformerEncoder(torch.nn.TransformerEncoderLayer(32, 4, 128, 0.0), 2, enable_nested_tensor=False)
    sd = {"transformer_encoder." + k: v for k, v in enc.state_dict().items()}
    sd.update({"embedding_dict.feat_user.weight": torch.randn(7, 32), "embedding_dict.feat_item.weight": torch.randn(9, 32),
               "ffn_user.weight": torch.randn(32, 32), "ffn_user.bias": torch.randn(32),
               "fnn_gate.weight": torch.randn(32, 33), "fnn_gate.bias": torch.randn(32),
               "decoder.weight": torch.randn(20, 32), "decoder.bias": torch.randn(20)})
    back = T.unpack(T.pack(sd, "cpu"))
    assert set(back) == set(sd)
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
    from oracle import nets
    assert torch.allclose(params.positional_encoding(31, 27), nets.positional_encoding(31, 27))


def test_replay_buffer_layout_and_index_arithmetic():
    """VectorReplayBuffer slot layout and prev / next / unfinished_index (tianshou/test/base/test_buffer.py's
    ReplayBufferManager cases restated for the env-major layout), on the host with device='cpu'."""
    from cirs_codes_b200.data import Batch, VectorReplayBuffer
    buf = VectorReplayBuffer(20, 4, device="cpu")           # 4 sub-buffers of 5 slots
    assert buf.maxsize == 20 and buf.sub_size == 5
    S = 3

    def add(ids, done):
        n = len(ids)
        b = Batch(obs=torch.ones(n, S) * len(buf), obs_next=torch.ones(n, S), act=np.array(ids) + 10,
                  rew=np.ones(n), done=np.array(done))
        return buf.add(b, buffer_ids=np.array(ids))

    ptr, ep_rew, ep_len, ep_idx = add([0, 1, 2, 3], [0, 0, 0, 0])
    assert list(ptr) == [0, 5, 10, 15] and list(ep_len) == [0, 0, 0, 0]
    ptr, ep_rew, ep_len, ep_idx = add([0, 1, 3], [0, 1, 0])
    assert list(ptr) == [1, 6, 16] and list(ep_len) == [0, 2, 0] and list(ep_rew) == [0, 2, 0] and ep_idx[1] == 5
    ptr, *_ = add([0, 3], [1, 0])
    assert list(ptr) == [2, 17]
    assert len(buf) == 9
    assert list(buf.sample_index(0)) == [0, 1, 2, 5, 6, 10, 15, 16, 17]
    assert list(buf.last_index) == [2, 6, 10, 17]
    assert list(buf.unfinished_index()) == [10, 17]
    assert list(buf.prev([0, 1, 2, 6, 10, 17])) == [0, 0, 1, 5, 10, 16]
    assert list(buf.next([0, 1, 2, 5, 6, 10, 16, 17])) == [1, 2, 2, 6, 6, 10, 17, 17]
    assert list(buf.act[[0, 5, 17]]) == [10, 11, 13] and buf.done[6] and not buf.done[5]
    b, idx = buf.sample(0)
    assert len(idx) == 9 and b.obs.shape == (9, S)
    buf.reset()
    assert len(buf) == 0


def test_split_indices_matches_oracle():
    from cirs_codes_b200.policy import split_indices
    from oracle import ppo
    for n, size in ((10, 3), (12, 4), (7, 16), (33, 8), (16, 16)):
        perm = np.random.default_rng(n).permutation(n)
        a, b = split_indices(n, size, perm), ppo.split_indices(n, size, perm)
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def test_feature_columns_surface():
    import cirs_codes_b200 as cb

    class E:
        mat = np.zeros((7, 9))

    u, a, f, hu, ha, hf = cb.get_dataset_columns(32, "KuaishouEnv-v0", E)
    assert (u[0].vocabulary_size, a[0].vocabulary_size, u[0].embedding_dim) == (7, 9, 32) and not hu and not ha and hf
    assert cb.compute_input_dim(a) == 32 and cb.build_input_features(u + f) == {"feat_user": (0, 1), "feat_feedback": (1, 2)}
    u, a, f, hu, ha, hf = cb.get_dataset_columns(27, "VirtualTB-v0")
    assert cb.compute_input_dim(u) == 88 and cb.compute_input_dim(a) == 27 and hu and ha


def test_trainer_loop_with_stub_collectors():
    """onpolicy_trainer control flow (core/trainer/onpolicy.py:156-240) on stub collectors / policy: collects until
    step_per_epoch, updates after every collect, evaluates once per epoch (+ once before training), hooks fire."""
    from cirs_codes_b200.trainer import MovAvg, onpolicy_trainer

    class Col:
        def __init__(self, policy, n_st):
            self.policy, self.n_st, self.buffer = policy, n_st, object()
            self.collect_step = self.collect_episode = 0
            self.collect_time = 1e-3
            self.calls = self.resets = 0

        def reset_stat(self):
            self.collect_step = self.collect_episode = 0

        def reset_env(self):
            self.resets += 1

        def reset_buffer(self, keep_statistics=False):
            pass

        def collect(self, n_step=None, n_episode=None):
            self.calls += 1
            self.collect_step += self.n_st
            self.collect_episode += n_episode
            return {"n/ep": n_episode, "n/st": self.n_st, "rew": float(self.calls), "rew_std": 0.0, "len": 3.0,
                    "rews": np.ones(n_episode), "lens": np.full(n_episode, 3)}

    class Pol:
        callbacks = []
        updates, mode = 0, None

        def train(self):
            self.mode = "train"

        def eval(self):
            self.mode = "eval"

        def update(self, sample_size, buffer, batch_size=None, repeat=1):
            assert self.mode == "train" and sample_size == 0
            self.updates += 1
            return {"loss": [1.0, 3.0], "loss/clip": [0.5, 0.5]}

    pol = Pol()
    tr, te = Col(pol, 40), Col(pol, 7)
    saved = []
    info = onpolicy_trainer(pol, tr, te, None, max_epoch=3, step_per_epoch=100, repeat_per_collect=2,
                            episode_per_test=5, batch_size=64, episode_per_collect=10, verbose=False,
                            save_model_fn=lambda epoch, policy: saved.append(epoch))
    assert tr.calls == 9 and pol.updates == 9            # ceil(100 / 40) = 3 collects per epoch
    assert te.calls == 4 and te.resets == 4              # one evaluation before training + one per epoch
    assert saved == [1, 2, 3]
    assert info["train_step"] == 360 and info["test_episode"] == 20 and info["best_reward"] == 4.0
    m = MovAvg(size=3)
    m.add([1.0, 2.0]); m.add(6.0); m.add(float("inf"))
    assert m.get() == 3.0 and m.add(9.0) == (2.0 + 6.0 + 9.0) / 3


def test_coverage_callback_matches_reference():
    """Callback_Coverage_Count (evaluation.py:286-371) on this package's buffer: our one-shot version against a brute
    force over the stored episodes, and -- in the build container, where /root/reference exists -- against the
    REFERENCE's own callback walking buffer.prev / next / last_index of the same buffer."""
    import importlib.util
    import pandas as pd
    from cirs_codes_b200.data import Batch, VectorReplayBuffer
    from cirs_codes_b200.evaluation import Callback_Coverage_Count
    rng = np.random.default_rng(4)
    n_item, B, L = 50, 6, 7
    cats = np.stack([rng.permutation(9)[:4] for _ in range(n_item)])   # distinct categories per item, like the data

    class Env:
        mat = [np.zeros((3, n_item))]

    class Col:
        pass

    class Set:
        env = Env()
        collector_dict = {}

    results, episodes = {}, {}
    for name in ("FB", "NX_0"):
        buf = VectorReplayBuffer(B * L, B, device="cpu")
        lens = rng.integers(1, L + 1, size=B)
        ready, t, eps = np.arange(B), 0, [[] for _ in range(B)]
        first = None
        while len(ready):
            acts = rng.integers(0, n_item, size=len(ready))
            done = lens[ready] == t + 1
            ptr, *_ = buf.add(Batch(obs=torch.zeros(len(ready), 2), obs_next=torch.zeros(len(ready), 2), act=acts,
                                    rew=np.ones(len(ready)), done=done), buffer_ids=ready)
            first = ptr if first is None else first
            for e, a in zip(ready, acts):
                eps[e].append(int(a))
            ready, t = ready[~done], t + 1
        c = Col()
        c.buffer = buf
        Set.collector_dict[name] = c
        episodes[name] = eps
        pre = "" if name == "FB" else name + "_"
        results[pre + "idxs"], results["n/ep"] = first, B
    dom = {"feat": [(3, 40), (5, 30), (1, 20), (7, 10)]}
    mine = Callback_Coverage_Count(Set, cats, need_transform=False, item_feat_domination=dom, top_rate=0.6)
    got = mine.on_epoch_end(1, dict(results))
    for name in ("FB", "NX_0"):
        pre = "" if name == "FB" else name + "_"
        flat = np.concatenate([np.array(e) for e in episodes[name]])
        assert got[pre + "CV"] == len(set(flat)) / n_item and got[pre + "CV_turn"] == len(set(flat)) / len(flat)
        assert got[pre + "ifeat_feat"] == np.mean([3 in cats[a] for a in flat])   # cumulative shares .4, .7: only value 3
    ref_path = "/root/reference/evaluation.py"
    if os.path.exists(ref_path):
        spec = importlib.util.spec_from_file_location("ref_evaluation", ref_path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        df = pd.DataFrame(cats, columns=[f"feat{i}" for i in range(4)])
        theirs = ref.Callback_Coverage_Count(Set, df, need_transform=False, item_feat_domination=dom, lbe_photo=None,
                                             top_rate=0.6).on_epoch_end(1, dict(results))
        for k in ("CV", "CV_turn", "ifeat_feat", "NX_0_CV", "NX_0_CV_turn", "NX_0_ifeat_feat"):
            assert abs(theirs[k] - got[k]) < 1e-12, (k, theirs[k], got[k])


def test_batch_item_assignment_and_split_match_tianshou():
    """D1: Batch.__setitem__(index) and Batch.split (tianshou/data/batch.py:244-267, 721-744) against tianshou's own
    Batch when the reference tree is present (build container); otherwise against their documented behaviour."""
    import torch
    from cirs_codes_b200.data import Batch
    b = Batch(obs=torch.arange(12.).reshape(6, 2), act=np.arange(6), info=Batch(), rew=np.zeros(6))
    b[np.array([1, 3])] = Batch(obs=torch.full((2, 2), -1.), act=np.array([10, 30]))
    assert b.act.tolist() == [0, 10, 2, 30, 4, 5] and b.obs[3].tolist() == [-1., -1.]
    assert b.rew.tolist() == [0.] * 6                       # a key missing on the right-hand side is zero-filled
    b[0] = {"act": 7}
    assert b.act[0] == 7 and b.obs[0].tolist() == [0., 0.]  # ... including tensors
    with pytest.raises(ValueError):
        b[0] = Batch(new_key=1)
    with pytest.raises(ValueError):
        b[0] = np.zeros(2)
    sizes = [len(x) for x in Batch(a=np.arange(10)).split(4, shuffle=False, merge_last=True)]
    assert sizes == [4, 6]
    sizes = [len(x) for x in Batch(a=np.arange(10)).split(4, shuffle=False)]
    assert sizes == [4, 4, 2]
    assert [len(x) for x in Batch(a=np.arange(3)).split(8)] == [3]
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):
        from oracle import ref_shims
        ref_shims.install()
        from tianshou.data import Batch as TB
        np.random.seed(3)
        mine = [x.a.tolist() for x in Batch(a=np.arange(23)).split(5, shuffle=True, merge_last=True)]
        np.random.seed(3)
        theirs = [x.a.tolist() for x in TB(a=np.arange(23)).split(5, shuffle=True, merge_last=True)]
        assert mine == theirs
        t = TB(obs=np.arange(12.).reshape(6, 2), act=np.arange(6), rew=np.zeros(6))
        m = Batch(obs=np.arange(12.).reshape(6, 2), act=np.arange(6), rew=np.ones(6))
        t.rew[:] = 1
        t[np.array([1, 3])] = TB(act=np.array([10, 30]))
        m[np.array([1, 3])] = Batch(act=np.array([10, 30]))
        assert np.array_equal(t.act, m.act) and np.array_equal(t.rew, m.rew) and np.array_equal(t.obs, m.obs)


def test_loggers_match_reference_format(tmp_path):
    """LoggerCallback_Policy's Info line and BasicLogger's scalar keys (util/utils.py:84-136,
    tianshou/utils/log_tools.py:84-200); compared with the reference's own classes when the reference tree is present."""
    import cirs_codes_b200 as cb
    results = {"n/ep": 4, "n/st": 20, "rew": 12.5, "CV": 0.123456, "CV_turn": 0.5, "ifeat_feat": 0.75,
               "NX_0_n/st": 16, "NX_0_rew": 10.0, "NX_0_CV": 0.2, "NX_0_CV_turn": 0.9, "NX_0_ifeat_feat": 0.5,
               "NX_10_n/st": 40, "NX_10_rew": 30.0, "NX_10_CV": 0.3, "NX_10_CV_turn": 1.0, "NX_10_ifeat_feat": 0.25}
    path = tmp_path / "log.txt"
    line = cb.LoggerCallback_Policy(str(path), 10).on_epoch_end(3, dict(results))
    assert line.startswith("Epoch: [3], Info: [{'num_test': 4, 'CV': '0.12346', 'CV_turn': '0.50000', 'ctr': '2.50000'")
    assert "'NX_10_ifeat_feat': 0.25" in line and path.read_text().strip() == line
    rec = cb.ScalarRecorder()
    lg = cb.BasicLogger(rec, train_interval=1, update_interval=1)
    r = {"n/ep": 2, "rews": np.array([1., 3.]), "lens": np.array([2, 4])}
    lg.log_train_data(r, 10)
    lg.log_test_data(dict(r), 10)
    lg.log_update_data({"loss": 0.5}, 7)
    lg.save_data(1, 10, 7, lambda *a: None)
    assert r["rew"] == 2.0 and set(rec.scalars) == {"train/n/ep", "train/rew", "train/len", "test/rew", "test/len",
                                                      "test/rew_std", "test/len_std", "loss", "save/epoch",
                                                      "save/env_step", "save/gradient_step"}
    assert lg.restore_data() == (1, 10, 7)
    if os.path.isdir("/root/reference"):
        from oracle import ref_shims
        ref_shims.install()
        import util.utils as ru
        got = []

        class _L:
            def info(self, msg):
                got.append(msg)

        ru.logger = _L()
        ru.LoggerCallback_Policy(str(tmp_path / "x.log"), 10).on_epoch_end(3, dict(results))
        assert got and got[0] == line
