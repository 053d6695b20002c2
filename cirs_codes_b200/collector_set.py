"""CollectorSet: the three test-time collectors of CIRS (core/collector_set.py:13-77) behind one ``collect``.

``envs_dict`` = {"FB": envs, "NX_0": envs, f"NX_{force_length}": envs} as built in CIRS-RL-kuaishou.py:213-221:
  FB      free browsing: the environment decides when the user leaves; nothing is masked
  NX_0    no repeated items (remove_recommended_ids), the environment decides when the user leaves
  NX_x    no repeated items and every episode is forced to last exactly ``force_length`` turns
Results of the non-FB collectors are prefixed with their name (``NX_0_rew``, ...), like the reference.
With this package's device-resident environments every collector runs the persistent rollout kernel; the
already-recommended set is a per-environment bitset maintained by the step kernel and masked inside the actor head.
"""
from .collector import Collector
from .data import VectorReplayBuffer


class CollectorSet:
    def __init__(self, policy, envs_dict, buffer_size, env_num, preprocess_fn=None, exploration_noise=False,
                 force_length=10):
        self.collector_dict = {}
        remove = {"FB": False, "NX_0": True, f"NX_{force_length}": True}
        force = {"FB": 0, "NX_0": 0, f"NX_{force_length}": force_length}
        for name, envs in envs_dict.items():
            self.collector_dict[name] = Collector(
                policy, envs, VectorReplayBuffer(buffer_size, env_num), preprocess_fn=preprocess_fn,
                exploration_noise=exploration_noise if name == "FB" else False,
                remove_recommended_ids=remove[name], force_length=force[name])
        self.env = envs_dict["FB"]
        self.policy, self.preprocess_fn = policy, preprocess_fn
        self.exploration_noise, self.env_num = exploration_noise, env_num
        self.collect_step = self.collect_episode = 0
        self.collect_time = 0.0

    def reset_stat(self):
        for c in self.collector_dict.values():
            c.reset_stat()

    def reset_buffer(self, keep_statistics=False):
        for c in self.collector_dict.values():
            c.reset_buffer(keep_statistics)

    def reset_env(self):
        for c in self.collector_dict.values():
            c.reset_env()

    def collect(self, n_step=None, n_episode=None, random=False, render=None, no_grad=True, users=None):
        all_res = {}
        for name, c in self.collector_dict.items():
            res = c.collect(n_step, n_episode, random, render, no_grad, users=users)
            all_res.update(res if name == "FB" else {name + "_" + k: v for k, v in res.items()})
        fb = self.collector_dict["FB"]
        self.collect_step, self.collect_episode, self.collect_time = fb.collect_step, fb.collect_episode, fb.collect_time
        return all_res
