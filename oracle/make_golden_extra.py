"""More golden vectors recorded by EXECUTING THE REFERENCE ITSELF (/root/reference); see oracle/make_golden.py.

TEST INFRASTRUCTURE ONLY -- run in the build container:

    python -m oracle.make_golden_extra collectors   # -> tests/golden/kuaishou_testcol.npz
    python -m oracle.make_golden_extra usergen      # -> tests/golden/taobao_usergen.npz

kuaishou_testcol: the three test-time collectors of core/collector_set.py:13-77 (FB, NX_0 = remove_recommended_ids,
  NX_x = remove_recommended_ids + force_length) on RAW KuaishouEnv test environments (kuaishouEnv.py:161-218) built with
  NON-IDENTITY label encoders (raw user ids 3u+5, raw item ids 2i+1), the reference's Callback_Coverage_Count
  (evaluation.py:286-371, both the "feat" and the per-feature branch of get_feat_dominate_dict :10-77) on their
  buffers, and teacher-forced steps of the SimulatedEnv training environment whose alpha_u / beta_i are indexed by RAW
  id through lbe_*.inverse_transform (simulated_env.py:157-161).
  Instrumentation (recording only): Categorical.sample made explicit (argmax p / q, q ~ Exp(1): bit-identical to
  torch.multinomial) so the race noise can be recorded -- scattered back to the ORIGINAL catalogue
  columns through the indices the reference's own masking returned (core/policy/utils.py:30-58).
taobao_usergen: VirtualTB's user generator (virtualTB/model/UserModel.py:40-60: z ~ U(0,1)^128 -> MLP -> 11 grouped
  softmax + multinomial -> one-hot) and click model (model/ActionModel.py:18-23) with the shipped weights
  (virtualTB/data/*.pt), their uniform seeds and race noise recorded the same way; and VirtualTB.step / reset
  (envs/virtualTB.py:74-113) driven with recorded actions.
"""
import os
import random
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

ref_shims.install()

import torch  # noqa: E402
import pandas as pd  # noqa: E402
import gym  # noqa: E402
from sklearn.preprocessing import LabelEncoder  # noqa: E402

from cirs_codes_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _sd(module, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def collectors_case(name="kuaishou_testcol", U=30, I=120, B=6, T=8, N=2, thr=1, d=16, nhead=4, force_length=5,
                    seed=23, tau=100.0, gamma_exposure=10.0):
    import core.policy.ppo as ref_ppo
    from core.collector_set import CollectorSet
    from core.policy.ppo import PPOPolicy
    from core.state_tracker import StateTrackerTransformer
    from core.inputs import get_dataset_columns
    from evaluation import Callback_Coverage_Count
    from tianshou.env import DummyVectorEnv
    from tianshou.utils.net.common import Net
    from tianshou.utils.net.discrete import Actor, Critic

    tb = synth.kuaishou_tables(U, I, seed=seed)
    raw_user = 3 * np.arange(U) + 5
    raw_item = 2 * np.arange(I) + 1
    lbe_user, lbe_photo = LabelEncoder().fit(raw_user), LabelEncoder().fit(raw_item)
    assert np.array_equal(lbe_user.classes_, raw_user) and np.array_equal(lbe_photo.classes_, raw_item)
    small = synth.cats_to_list_feat(tb["cats"])
    list_feat = [[] for _ in range(int(raw_item.max()) + 1)]          # indexed by RAW item id (kuaishouEnv.py:52)
    for j, r in enumerate(raw_item):
        list_feat[r] = small[j]
    dist = synth.jaccard_distance_matrix(tb["cats"])
    alpha_raw = np.full((int(raw_user.max()) + 1, 1), np.nan)
    beta_raw = np.full((int(raw_item.max()) + 1, 1), np.nan)
    alpha_raw[raw_user, 0], beta_raw[raw_item, 0] = tb["alpha_u"], tb["beta_i"]
    gym.register(id="KuaishouEnv-v0", entry_point="environments.KuaishouRec.env.kuaishouEnv:KuaishouEnv",
                 kwargs=dict(mat=tb["mat"].astype(np.float64), lbe_user=lbe_user, lbe_photo=lbe_photo,
                             num_leave_compute=N, leave_threshold=thr, max_turn=T, list_feat=list_feat,
                             df_photo_env=None, df_dist_small=pd.DataFrame(dist)))
    env = gym.make("KuaishouEnv-v0")
    gym.register(id="SimulatedEnv-v0", entry_point="core.env.simulatedEnv.simulated_env:SimulatedEnv",
                 kwargs=dict(user_model=torch.nn.Identity(), task_name="KuaishouEnv-v0", version="v1", tau=tau,
                             alpha_u=alpha_raw, beta_i=beta_raw, normed_mat=tb["normed_mat"].astype(np.float64),
                             gamma_exposure=gamma_exposure, r_decay=0.9))
    out = dict(cfg=np.array([U, I, B, T, N, thr, d, nhead, force_length, seed], dtype=np.int64),
               cfg_f=np.array([tau, gamma_exposure, 0.9]), raw_user=raw_user, raw_item=raw_item,
               alpha_raw=alpha_raw, beta_raw=beta_raw, **{k: v for k, v in tb.items()})

    # ---- (a) SimulatedEnv with raw-id-indexed alpha / beta: teacher-forced random actions
    random.seed(seed)
    sim = gym.make("SimulatedEnv-v0")
    rng = np.random.default_rng(seed)
    n_ep = 4
    for ep in range(n_ep):
        obs = sim.reset()
        out[f"sim/ep{ep}/user"] = np.asarray(obs).astype(np.int64).reshape(-1)
        acts, rews, dones = [], [], []
        for t in range(T):
            a = int(rng.integers(0, I)) if (t == 0 or rng.random() > 0.3) else int(acts[rng.integers(0, len(acts))])
            _, r, dn, _ = sim.step(a)
            acts.append(a); rews.append(float(r)); dones.append(bool(dn))
            if dn:
                break
        out[f"sim/ep{ep}/act"], out[f"sim/ep{ep}/rew"] = np.array(acts), np.array(rews)
        out[f"sim/ep{ep}/done"] = np.array(dones)
    out["sim/n_ep"] = np.array(n_ep)

    # ---- (b) the three test collectors
    envs_dict = {"FB": DummyVectorEnv([lambda: gym.make("KuaishouEnv-v0") for _ in range(B)]),
                 "NX_0": DummyVectorEnv([lambda: gym.make("KuaishouEnv-v0") for _ in range(B)]),
                 f"NX_{force_length}": DummyVectorEnv([lambda: gym.make("KuaishouEnv-v0") for _ in range(B)])}
    np.random.seed(seed)
    torch.manual_seed(seed)
    cols = get_dataset_columns(d, envname="KuaishouEnv-v0", env=env)
    tracker = StateTrackerTransformer(cols[0], cols[1], cols[2], dim_model=d, dim_state=20, dim_max_batch=B,
                                      dataset="KuaishouEnv-v0", has_user_embedding=cols[3],
                                      has_action_embedding=cols[4], has_feedback_embedding=cols[5], nhead=nhead,
                                      d_hid=128, nlayers=2, dropout=0.0, device="cpu", seed=seed, MAX_TURN=T)
    with torch.no_grad():
        for emb in tracker.embedding_dict.values():
            emb.weight.normal_(0, 0.1)
    net = Net(20, hidden_sizes=[64, 64], device="cpu")
    actor, critic = Actor(net, I, device="cpu"), Critic(net, device="cpu")
    for m in list(actor.modules()) + list(critic.modules()):
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.orthogonal_(m.weight)
            torch.nn.init.zeros_(m.bias)
    with torch.no_grad():
        actor.last.model[0].weight.mul_(6.0)      # a peaked distribution, so that masking matters for the samples
    optim = [torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=1e-3),
             torch.optim.Adam(tracker.parameters(), lr=1e-3)]
    policy = PPOPolicy(actor, critic, optim, torch.distributions.Categorical, discount_factor=0.95,
                       max_grad_norm=0.5, eps_clip=0.2, vf_coef=0.25, ent_coef=0.0, reward_normalization=1,
                       advantage_normalization=1, recompute_advantage=0, value_clip=1, gae_lambda=0.95,
                       action_space=env.action_space, action_bound_method="", action_scaling=False)
    out.update(_sd(tracker, "init/tracker/"))
    out.update(_sd(actor, "init/actor/"))
    out.update(_sd(critic, "init/critic/"))

    state = {"turns": None, "cur": None, "idx": None}
    orig_remove = ref_ppo.removed_recommended_id_from_embedding

    def remove_rec(logits, recommended_ids):
        lm, im = orig_remove(logits, recommended_ids)
        state["idx"] = im.numpy().copy()
        return lm, im

    ref_ppo.removed_recommended_id_from_embedding = remove_rec

    def sample_explicit(self, sample_shape=torch.Size()):
        p = self.probs
        q = torch.empty_like(p).exponential_(1)
        full = np.ones((p.shape[0], I), dtype=np.float32)
        cols_ = state["idx"] if state["idx"] is not None else np.tile(np.arange(I), (p.shape[0], 1))
        np.put_along_axis(full, cols_, q.numpy(), axis=1)
        state["cur"]["q"] = full
        state["cur"]["n_masked"] = np.array(I - p.shape[1])
        return torch.argmax(p / q, dim=-1)

    torch.distributions.Categorical.sample = sample_explicit
    orig_forward = policy.forward

    def fwd(batch, *a, **k):
        state["idx"] = None
        state["cur"] = {}
        state["turns"].append(state["cur"])
        state["cur"]["state"] = batch.obs.detach().numpy().copy()
        res = orig_forward(batch, *a, **k)
        state["cur"]["act"] = res.act.numpy().copy()
        return res

    policy.forward = fwd
    orig_build = tracker.build_state

    def build(**k):
        o = orig_build(**k)
        if k.get("obs") is not None:
            state["turns"] = []
            state["reset_obs"] = np.asarray(k["obs"]).copy()
            state["s0"] = o["obs"].detach().numpy().copy()
        elif k.get("obs_next") is not None:
            c = state["cur"]
            c["env_id"] = np.asarray(k["env_id"]).copy()
            c["obs_next_raw"] = np.asarray(k["obs_next"]).copy()
            c["rew"] = np.asarray(k["rew"], dtype=np.float64).copy()
            c["done"] = np.asarray(k["done"]).copy()
            c["state_next"] = o["obs_next"].detach().numpy().copy()
        return o

    cset = CollectorSet(policy, envs_dict, B * (T + 2), B, preprocess_fn=build, exploration_noise=False,
                        force_length=force_length)
    policy.eval()
    random.seed(seed + 1)
    results = {}
    for cname, collector in cset.collector_dict.items():       # one collector at a time, to record each one's turns
        res = collector.collect(n_episode=B)
        results.update(res if cname == "FB" else {cname + "_" + k: v for k, v in res.items()})
        P = f"{cname}/"
        out[P + "users"] = state["reset_obs"][:, 0].astype(np.int64)
        out[P + "s0"] = state["s0"]
        out[P + "n_turns"] = np.array(len(state["turns"]))
        for t, c in enumerate(state["turns"]):
            for k, v in c.items():
                out[P + f"turn{t}/{k}"] = v
        buf = collector.buffer
        idx = buf.sample_index(0)
        out[P + "buf/act"], out[P + "buf/rew"] = buf.act[idx].copy(), buf.rew[idx].copy()
        out[P + "buf/done"], out[P + "buf/lengths"] = buf.done[idx].copy(), buf._lengths.copy()
        for k in ("n/ep", "n/st", "rews", "lens", "idxs", "rew", "len", "rew_std", "len_std"):
            out[P + "res/" + k.replace("/", "_")] = np.asarray(res[k])

    # ---- (c) the reference's coverage / dominated-category callback on those buffers ("feat" branch and per-feature branch)
    feat_cols = {f"feat{k}": tb["cats"][:, k] for k in range(4)}
    df_item_val = pd.DataFrame(feat_cols, index=raw_item)
    flat = tb["cats"][tb["cats"] > 0]
    vals, cnts = np.unique(flat, return_counts=True)
    order = np.argsort(-cnts, kind="stable")
    dom_feat = {"feat": [(int(vals[i]), int(cnts[i])) for i in order]}
    cb = Callback_Coverage_Count(cset, df_item_val, True, dom_feat, lbe_photo, 0.6)
    r1 = cb.on_epoch_end(1, dict(results))
    dom_each = {}
    for k in range(2):
        v_, c_ = np.unique(tb["cats"][:, k], return_counts=True)
        o_ = np.argsort(-c_, kind="stable")
        dom_each[f"feat{k}"] = [(int(v_[i]), int(c_[i])) for i in o_]
    cb2 = Callback_Coverage_Count(cset, df_item_val, True, dom_each, lbe_photo, 0.5)
    r2 = cb2.on_epoch_end(1, dict(results))
    for tag, r in (("cov_feat", r1), ("cov_each", r2)):
        for k, v in r.items():
            if np.isscalar(v) and ("CV" in k or "ifeat" in k):
                out[f"{tag}/{k}"] = np.array(float(v))
    out["dom_feat"] = np.array(dom_feat["feat"], dtype=np.int64)
    for k, v in dom_each.items():
        out[f"dom_each/{k}"] = np.array(v, dtype=np.int64)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; lens:",
          {c: out[c + "/res/lens"].tolist() for c in cset.collector_dict},
          {k: float(v) for k, v in out.items() if k.startswith("cov_")})


def usergen_case(name="taobao_usergen", n_users=48, n_click=64, seed=31):
    """VirtualTB's generator and click model with the shipped weights; uniform seeds and race noise recorded."""
    from virtualTB.model.UserModel import UserModel
    from virtualTB.model.ActionModel import ActionModel
    from virtualTB.model.LeaveModel import LeaveModel
    torch.manual_seed(seed)
    um, am, lm = UserModel(), ActionModel(), LeaveModel()
    um.load(); am.load(); lm.load()
    out = {}
    out.update(_sd(um.generator_model, "generator/"))
    out.update(_sd(am.model, "action/"))
    noise = {}
    orig_multinomial = torch.multinomial

    def multinomial_explicit(p, num_samples, replacement=False, *, generator=None):
        assert num_samples == 1 and not replacement
        q = torch.empty_like(p).exponential_(1)
        noise.setdefault("q", []).append(q.numpy().copy())
        return torch.argmax(p / q, dim=-1, keepdim=True)

    # the explicit form draws the same sample as torch.multinomial from the same generator state
    st = torch.get_rng_state()
    probe = torch.softmax(torch.randn(5, 11), 1)
    torch.set_rng_state(st); a = orig_multinomial(probe, 1)
    torch.set_rng_state(st); b = multinomial_explicit(probe, 1)
    assert torch.equal(a, b), "explicit multinomial differs from torch's"
    noise.clear()
    torch.manual_seed(seed)
    torch.multinomial = multinomial_explicit
    zs, users, qs = [], [], []
    for _ in range(n_users):
        z = torch.rand((1, um.seed_dimesion))
        noise.clear()
        u = um.generate(z)
        zs.append(z.numpy()[0].copy()); users.append(u.detach().numpy()[0].copy())
        qs.append(np.concatenate([q[0] for q in noise["q"]]))        # 11 groups -> 88 race draws
    out["gen/z"], out["gen/user"], out["gen/q"] = np.array(zs), np.array(users), np.array(qs)
    # click model: predict(user, page, weight) -> (a, b); reward = a (virtualTB.py:84-86)
    g = torch.Generator().manual_seed(seed)
    xu = torch.tensor(np.array(users)[np.arange(n_click) % n_users])
    page = torch.randint(0, 10, (n_click, 1), generator=g).float()
    act = torch.rand(n_click, 27, generator=g) * 2 - 1
    res, qa = [], []
    for i in range(n_click):
        noise.clear()
        r = am.predict(xu[i:i + 1], page[i:i + 1], act[i:i + 1])
        res.append(r.numpy()[0].copy())
        qa.append(np.concatenate([q[0] for q in noise["q"]]))        # 11 + 10 race draws
    out["click/user"], out["click/page"], out["click/act"] = xu.numpy(), page.numpy(), act.numpy()
    out["click/result"], out["click/q"] = np.array(res), np.array(qa)
    torch.multinomial = orig_multinomial
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; clicks:", np.array(res)[:8, 0].tolist(),
          "user one-hot sums:", np.array(users).sum(1)[:4].tolist())


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1] if len(sys.argv) > 1 else "collectors"
    if what == "collectors":
        collectors_case()
    elif what == "usergen":
        usergen_case()
