"""CPU restatement of the reference's training / test environments (numpy, float64 like the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this file.
Parity pinned: against tests/golden/kuaishou_*.npz and tests/golden/taobao_*.npz, which were produced by executing
the reference itself (oracle/make_golden.py).

Each env object holds B independent environments (the reference holds one per Python object and loops over
them in DummyVectorEnv, tianshou/env/venvs.py:212-220); the arithmetic per environment follows the cited lines.
"""
import numpy as np


def _cats_to_mask(cats):
    m = np.zeros(cats.shape[0], dtype=np.uint32)
    for k in range(cats.shape[1]):
        c = cats[:, k].astype(np.uint32)
        m |= np.where(c > 0, np.uint32(1) << c, np.uint32(0)).astype(np.uint32)
    return m


class KuaishouSimOracle:
    """SimulatedEnv over KuaishouEnv (core/env/simulatedEnv/simulated_env.py:111-168,
    environments/KuaishouRec/env/kuaishouEnv.py:161-218).  ``simulated=False`` gives the raw KuaishouEnv used by
    the test collectors (reward = mat[u, a], no exposure)."""

    def __init__(self, mat, normed_mat, cats, alpha_u=None, beta_i=None, dist=None, *, max_turn=30,
                 num_leave_compute=1, leave_threshold=0, tau=100.0, gamma_exposure=10.0, r_decay=1.0,
                 version="v1", simulated=True, raw_user=None, raw_item=None):
        self.mat = np.asarray(mat, dtype=np.float64)
        self.normed = None if normed_mat is None else np.asarray(normed_mat, dtype=np.float64)
        self.cats = np.asarray(cats)
        self.mask = _cats_to_mask(self.cats)
        self.alpha = None if alpha_u is None else np.asarray(alpha_u, dtype=np.float64).reshape(-1)
        self.beta = None if beta_i is None else np.asarray(beta_i, dtype=np.float64).reshape(-1)
        self.dist = dist  # optional explicit I x I distance matrix (df_dist_small); None -> 1/Jaccard of cats
        self.T, self.N, self.thr = int(max_turn), int(num_leave_compute), leave_threshold
        self.tau, self.gamma_e, self.r_decay, self.version = float(tau), float(gamma_exposure), float(r_decay), version
        self.simulated = simulated
        # lbe_user.classes_ / lbe_photo.classes_: alpha_u / beta_i are then indexed by RAW id, through
        # lbe_*.inverse_transform(encoded) == classes_[encoded]  (simulated_env.py:157-161)
        self.raw_user = None if raw_user is None else np.asarray(raw_user)
        self.raw_item = None if raw_item is None else np.asarray(raw_item)

    def reset(self, users):
        """kuaishouEnv.py:182-190, simulated_env.py:59-72.  users are injected (the reference draws
        random.randint, kuaishouEnv.py:155-159)."""
        self.user = np.asarray(users, dtype=np.int64).copy()
        B = len(self.user)
        self.turn = np.zeros(B, dtype=np.int64)
        self.hist = np.zeros((B, self.T), dtype=np.int64)
        self.cum = np.zeros(B, dtype=np.float64)
        return self.user.reshape(B, 1).copy()

    def _distance(self, a, hist):
        if self.dist is not None:  # util.py:33-36  df_dist_small.iloc[action, hist]
            return np.asarray(self.dist)[a, hist].astype(np.float64)
        ma, mh = self.mask[a], self.mask[hist]
        inter = np.bitwise_count(ma & mh).astype(np.float64)
        union = np.bitwise_count(ma | mh).astype(np.float64)
        with np.errstate(divide="ignore"):
            return 1.0 / (inter / union)  # util.py:234-268: 1 / Jaccard, inf when disjoint

    def _leave(self, e, t, a):
        # kuaishouEnv.py:199-218
        if t == 0:
            return False
        seq = list(self.hist[e, :t])
        window = seq[t - self.N:t]  # python negative-slice quirk when t < N
        cnt = {}
        for it in window:
            for c in self.cats[it]:
                if c > 0:
                    cnt[int(c)] = cnt.get(int(c), 0) + 1
        for c in self.cats[a]:
            if c > 0 and cnt.get(int(c), 0) > self.thr:
                return True
        return False

    def step(self, act, env_ids=None):
        """One transition for the listed envs.  Returns (obs_next int64[n,1], rew f64[n], done bool[n])."""
        act = np.asarray(act).reshape(-1).astype(np.int64)
        ids = np.arange(len(self.user)) if env_ids is None else np.asarray(env_ids)
        rew = np.zeros(len(ids), dtype=np.float64)
        done = np.zeros(len(ids), dtype=bool)
        for k, e in enumerate(ids):
            a, t, u = int(act[k]), int(self.turn[e]), int(self.user[e])
            d = self._leave(e, t, a)
            if t >= self.T - 1:  # kuaishouEnv.py:167-168
                d = True
            if not self.simulated:
                r = self.mat[u, a]  # kuaishouEnv.py:171
            else:
                # exposure effect, simulated_env.py:147-168 + util.py:41-46
                if t == 0 or self.tau <= 0:
                    E = 0.0
                else:
                    hist = self.hist[e, :t]
                    dist = self._distance(a, hist)
                    E = float(np.sum(np.exp(-(t - np.arange(t)) * dist / self.tau)))
                    if self.alpha is not None:
                        u_id = u if self.raw_user is None else int(self.raw_user[u])
                        p_id = a if self.raw_item is None else int(self.raw_item[a])
                        E = E * self.alpha[u_id] * self.beta[p_id]
                    E = E * self.gamma_e
                r = self.normed[u, a]  # simulated_env.py:100
                # clip0 == np.amax(x, 0) is an identity on scalars (util.py:53-54)
                r = r / (1.0 + E) if self.version == "v1" else (r - E)
            if t < self.T:
                self.hist[e, t] = a  # simulated_env.py:123-124
            if self.simulated:
                n_prev = int(np.sum(self.hist[e, :t] == a))  # num_actions[a] - 1, simulated_env.py:129-132
                r = r * self.r_decay ** n_prev
            self.cum[e] += r
            self.turn[e] = t + 1
            rew[k], done[k] = r, d
        return act.reshape(-1, 1).copy(), rew, done


class TaobaoSimOracle:
    """SimulatedEnv over VirtualTB (simulated_env.py:77-168, virtualTB.py:74-133).  The real env's click /
    new-user draws only consume RNG and are discarded by SimulatedEnv (simulated_env.py:114,138), so only the
    exit test is restated.  ``reward_fn(x[n,118] f32) -> y[n]`` is the user model (UserModel_MMOE.forward)."""

    def __init__(self, reward_fn, *, max_turn=50, num_leave_compute=5, leave_threshold=3.0, tau=10.0,
                 gamma_exposure=10.0, version="v1"):
        self.reward_fn = reward_fn
        self.T, self.N, self.thr = int(max_turn), int(num_leave_compute), float(leave_threshold)
        self.tau, self.gamma_e, self.version = float(tau), float(gamma_exposure), version

    def reset(self, users):
        self.user = np.asarray(users, dtype=np.float64).copy()  # [B,88]
        B = len(self.user)
        self.turn = np.zeros(B, dtype=np.int64)
        self.hist = np.zeros((B, self.T, 27), dtype=np.float64)
        self.hist32 = np.zeros((B, self.T, 27), dtype=np.float32)  # VirtualTB.history_action keeps the f32 action
        self.prev_r = np.zeros(B, dtype=np.float64)
        self.cum = np.zeros(B, dtype=np.float64)
        return np.concatenate([self.user, np.zeros((B, 3))], axis=1)  # virtualTB.py:54-55

    def step(self, act, env_ids=None):
        act = np.asarray(act, dtype=np.float32).reshape(-1, 27)
        ids = np.arange(len(self.user)) if env_ids is None else np.asarray(env_ids)
        n = len(ids)
        rew, done = np.zeros(n), np.zeros(n, dtype=bool)
        obs = np.zeros((n, 30))
        for k, e in enumerate(ids):
            a, t = act[k], int(self.turn[e])
            d = False
            for tl in range(t - 1, max(-1, t - self.N), -1):  # virtualTB.py:126-133 (float32 norm)
                if np.linalg.norm(a - self.hist32[e, tl]) <= self.thr:
                    d = True
                    break
            if t >= self.T - 1:
                d = True
            if t == 0 or self.tau <= 0:
                E = 0.0
            else:
                dist = np.linalg.norm(a.astype(np.float64) - self.hist[e, :t], axis=1)  # util.py:24-28
                E = float(np.sum(np.exp(-(t - np.arange(t)) * dist / self.tau))) * self.gamma_e
            if t < self.T:
                self.hist[e, t] = a
                self.hist32[e, t] = a
            x = np.concatenate([self.user[e], [self.prev_r[e], 0.0, float(t)], a]).astype(np.float32)
            y = float(self.reward_fn(x[None, :])[0])  # simulated_env.py:79-86
            y = min(max(y, 0.0), 10.0)
            r = y / (1.0 + E) if self.version == "v1" else (y - E)
            self.prev_r[e] = r
            self.cum[e] += r
            self.turn[e] = t + 1
            rew[k], done[k] = r, d
            obs[k] = np.concatenate([a, [r, 0.0, float(t + 1)]])  # simulated_env.py:50
        return obs, rew, done


class VirtualTBOracle:
    """The raw VirtualTB environment's two networks (environments/VirtualTaobao/virtualTB/model/UserModel.py:13-60,
    model/ActionModel.py:6-23) in numpy float32, with the multinomial draws as explicit exponential races
    argmax_j p_j / q_j (what torch.multinomial(p, 1) computes) so that recorded noise can be replayed.
    ``gen`` / ``act``: state_dicts of generator_model / ActionModel.model ("0.weight", "0.bias", "2.weight", ...)."""

    GROUPS = (0, 8, 16, 27, 38, 49, 60, 62, 64, 67, 85, 88)  # UserModel.py:22-32

    def __init__(self, gen=None, act=None):
        f = lambda sd: None if sd is None else {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}  # noqa: E731
        self.gen, self.act = f(gen), f(act)

    @staticmethod
    def _softmax(x):
        e = np.exp(x - x.max(axis=1, keepdims=True))
        return e / e.sum(axis=1, keepdims=True)

    @staticmethod
    def _leaky(x):
        return np.where(x > 0, x, np.float32(0.01) * x)

    def generate(self, z, q):
        """UserModel.generate: z [n,128] uniform seeds, q [n,88] Exp(1) draws -> one-hot x 11 users [n,88]."""
        g = self.gen
        h = self._leaky(np.asarray(z, np.float32) @ g["0.weight"].T + g["0.bias"])
        x = h @ g["2.weight"].T + g["2.bias"]
        out = np.zeros_like(x)
        for lo, hi in zip(self.GROUPS[:-1], self.GROUPS[1:]):
            p = self._softmax(x[:, lo:hi])
            out[np.arange(len(x)), lo + np.argmax(p / np.asarray(q, np.float32)[:, lo:hi], axis=1)] = 1.0
        return out

    def click(self, user, page, action, q):
        """ActionModel.predict(user [n,88], page [n,1], action [n,27]) with q [n,21] -> int [n,2] = (a, b)."""
        a = self.act
        x = np.concatenate([user, page, action], axis=1).astype(np.float32)
        h = self._leaky(x @ a["0.weight"].T + a["0.bias"])
        h = self._leaky(h @ a["2.weight"].T + a["2.bias"])
        y = h @ a["4.weight"].T + a["4.bias"]
        q = np.asarray(q, np.float32)
        ca = np.argmax(self._softmax(y[:, :11]) / q[:, :11], axis=1)
        cb = np.argmax(self._softmax(y[:, 11:]) / q[:, 11:], axis=1)
        return np.stack([ca, cb], axis=1)
