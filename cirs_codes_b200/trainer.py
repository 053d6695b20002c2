"""The caller of the hot path: CIRS's on-policy training loop (core/trainer/onpolicy.py:30-252, a fork of tianshou
0.4.2's onpolicy_trainer) and its helpers (tianshou/trainer/utils.py:10-86, tianshou/utils/statistics.py:7-63), so
that CIRS-RL-kuaishou.py:320-334 runs on this package with only its imports changed.

Same keyword arguments, hooks and result dictionary.  What the loop does per epoch: until ``step_per_epoch``
transitions have been collected -- ``train_collector.collect(n_episode=episode_per_collect)`` then
``policy.update(0, buffer, batch_size, repeat)`` (the bench "step"); then one evaluation with ``test_collector``
(a Collector or a CollectorSet), the callbacks' ``on_epoch_end`` and ``save_model_fn(epoch=, policy=)``.
The progress bar of the reference (tqdm) is replaced by one optional line per epoch.

``save_checkpoint`` writes the reference's checkpoint layout {'policy', 'optim_RL', 'optim_state', 'state_tracker'}
(CIRS-RL-kuaishou.py:340-358); the two optimizer entries carry this package's flat Adam moments.
"""
import time
from collections import defaultdict

import numpy as np


class MovAvg:
    """Moving average over the last ``size`` scalars (tianshou/utils/statistics.py:7-63); lists are averaged in."""

    def __init__(self, size=100):
        self.size, self.cache = size, []

    def add(self, x):
        xs = np.asarray(x, dtype=np.float64).reshape(-1)
        self.cache.extend(v for v in xs if np.isfinite(v))
        if len(self.cache) > self.size:
            self.cache = self.cache[-self.size:]
        return self.get()

    def get(self):
        return float(np.mean(self.cache)) if self.cache else 0.0


class LazyLogger:
    """A logger that logs nothing (tianshou/utils/log_tools.py LazyLogger): the default of the trainer."""

    def log_train_data(self, result, step):
        pass

    def log_test_data(self, result, step):
        pass

    def log_update_data(self, losses, step):
        pass

    def save_data(self, epoch, env_step, gradient_step, save_checkpoint_fn=None):
        if save_checkpoint_fn:
            save_checkpoint_fn(epoch, env_step, gradient_step)

    def restore_data(self):
        return 0, 0, 0


def test_episode(policy, collector, test_fn, epoch, n_episode, logger=None, global_step=None, reward_metric=None):
    """tianshou/trainer/utils.py:10-31."""
    collector.reset_env()
    collector.reset_buffer()
    policy.eval()
    if test_fn:
        test_fn(epoch, global_step)
    result = collector.collect(n_episode=n_episode)
    if reward_metric:
        result["rews"] = reward_metric(result["rews"])
    if logger and global_step is not None:
        logger.log_test_data(result, global_step)
    return result


def gather_info(start_time, train_c, test_c, best_reward, best_reward_std):
    """tianshou/trainer/utils.py:34-86: the summary dictionary (``train_speed`` = env-steps per second, the unit of
    this repository's benchmark)."""
    duration = time.time() - start_time
    test_time = max(test_c.collect_time, 1e-9)
    model_time = duration - test_c.collect_time
    result = {
        "test_step": test_c.collect_step, "test_episode": test_c.collect_episode,
        "test_time": f"{test_c.collect_time:.2f}s", "test_speed": f"{test_c.collect_step / test_time:.2f} step/s",
        "best_reward": best_reward, "best_result": f"{best_reward:.2f} ± {best_reward_std:.2f}",
        "duration": f"{duration:.2f}s", "train_time/model": f"{model_time:.2f}s",
    }
    if train_c is not None:
        model_time -= train_c.collect_time
        train_speed = train_c.collect_step / max(duration - test_c.collect_time, 1e-9)
        result.update({
            "train_step": train_c.collect_step, "train_episode": train_c.collect_episode,
            "train_time/collector": f"{train_c.collect_time:.2f}s", "train_time/model": f"{model_time:.2f}s",
            "train_speed": f"{train_speed:.2f} step/s",
        })
    return result


def onpolicy_trainer(policy, train_collector, test_collector, state_tracker=None, max_epoch=1, step_per_epoch=1,
                     repeat_per_collect=1, episode_per_test=1, batch_size=64, step_per_collect=None,
                     episode_per_collect=None, train_fn=None, test_fn=None, stop_fn=None, save_fn=None,
                     save_checkpoint_fn=None, resume_from_log=False, reward_metric=None, logger=None, verbose=True,
                     test_in_train=True, save_model_fn=None):
    """core/trainer/onpolicy.py:30-252.  Only ``episode_per_collect`` collection is supported (CIRS always collects
    whole episodes, SURVEY §9 invariants)."""
    assert step_per_collect is None and episode_per_collect, "CIRS collects whole episodes (episode_per_collect)"
    logger = logger or LazyLogger()
    start_epoch, env_step, gradient_step = 0, 0, 0
    if resume_from_log:
        start_epoch, env_step, gradient_step = logger.restore_data()
    last_rew, last_len = 0.0, 0
    stat = defaultdict(MovAvg)
    start_time = time.time()
    train_collector.reset_stat()
    test_collector.reset_stat()
    test_in_train = test_in_train and train_collector.policy is policy
    test_result = test_episode(policy, test_collector, test_fn, start_epoch, episode_per_test, logger, None,
                               reward_metric)
    best_epoch, best_reward, best_reward_std = start_epoch, test_result["rew"], test_result["rew_std"]
    for cb in getattr(policy, "callbacks", []):
        cb.on_train_begin()

    for epoch in range(1 + start_epoch, 1 + max_epoch):
        policy.train()
        for cb in getattr(policy, "callbacks", []):
            cb.on_epoch_begin(epoch)
        collected, losses = 0, {}
        while collected < step_per_epoch:
            if train_fn:
                train_fn(epoch, env_step)
            result = train_collector.collect(n_step=step_per_collect, n_episode=episode_per_collect)
            if result["n/ep"] > 0 and reward_metric:
                result["rews"] = reward_metric(result["rews"])
            env_step += int(result["n/st"])
            collected += int(result["n/st"])
            logger.log_train_data(result, env_step)
            last_rew = result.get("rew", last_rew)
            last_len = result.get("len", last_len)
            if result["n/ep"] > 0 and test_in_train and stop_fn and stop_fn(result["rew"]):
                test_result = test_episode(policy, test_collector, test_fn, epoch, episode_per_test, logger, None)
                if stop_fn(test_result["rew"]):
                    if save_fn:
                        save_fn(policy)
                    logger.save_data(epoch, env_step, gradient_step, save_checkpoint_fn)
                    return gather_info(start_time, train_collector, test_collector, test_result["rew"],
                                       test_result["rew_std"])
                policy.train()
            losses = policy.update(0, train_collector.buffer, batch_size=batch_size, repeat=repeat_per_collect)
            gradient_step += max([1] + [len(v) for v in losses.values() if isinstance(v, list)])
            for k in losses:
                stat[k].add(losses[k])
                losses[k] = stat[k].get()
            logger.log_update_data(losses, gradient_step)
        test_result = test_episode(policy, test_collector, test_fn, epoch, episode_per_test, logger, None,
                                   reward_metric)
        rew, rew_std = test_result["rew"], test_result["rew_std"]
        if best_epoch < 0 or best_reward < rew:
            best_epoch, best_reward, best_reward_std = epoch, rew, rew_std
            if save_fn:
                save_fn(policy)
        logger.save_data(epoch, env_step, gradient_step, save_checkpoint_fn)
        for cb in getattr(policy, "callbacks", []):
            cb.on_epoch_end(epoch, test_result)
        if save_model_fn:
            save_model_fn(epoch=epoch, policy=policy)
        if verbose:
            print(f"Epoch #{epoch}: env_step {env_step} R_tra {last_rew:.3f} len_tra {last_len:.2f} "
                  f"loss {losses.get('loss', float('nan')):.4f} | test_reward: {rew:.6f} ± {rew_std:.6f}, "
                  f"best_reward: {best_reward:.6f} ± {best_reward_std:.6f} in #{best_epoch}", flush=True)
        if stop_fn and stop_fn(best_reward):
            break
    for cb in getattr(policy, "callbacks", []):
        cb.on_train_end()
    return gather_info(start_time, train_collector, test_collector, best_reward, best_reward_std)


def save_checkpoint(path, policy, state_tracker, optim=None):
    """The reference's checkpoint dictionary (CIRS-RL-kuaishou.py:340-358): 'policy' / 'state_tracker' are state dicts
    in the reference's parameter naming, 'optim_RL' / 'optim_state' are ``torch.optim.Adam.state_dict()``s in the
    reference's parameter order (the policy's and the tracker's torch optimizers are bridged to the device-side Adam
    state, PPOPolicy._bridge_optim) -- a checkpoint written here loads into the reference's optimizers and vice versa.
    The running return statistics (a plain attribute of tianshou's policy, not part of its state_dict) travel under the
    extra key 'ret_rms'."""
    import torch
    optim = optim or policy.optim
    torch.save({"policy": policy.state_dict(), "optim_RL": optim[0].state_dict(),
                "optim_state": optim[1].state_dict() if len(optim) > 1 else {},
                "state_tracker": state_tracker.state_dict(), "ret_rms": policy.ret_rms.t.detach().cpu()}, path)


def load_checkpoint(path, policy, state_tracker, optim=None):
    """Inverse of save_checkpoint; also reads a checkpoint written by the reference (no 'ret_rms' key).  Optimizer
    entries that cannot be restored raise instead of silently resetting Adam."""
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=False)
    optim = optim or policy.optim
    policy.load_state_dict(ck["policy"])
    state_tracker.load_state_dict(ck["state_tracker"])
    if "ret_rms" in ck:
        policy.ret_rms.t.copy_(torch.as_tensor(ck["ret_rms"], dtype=torch.float64))
    for opt, key in ((optim[0], "optim_RL"), (optim[1] if len(optim) > 1 else None, "optim_state")):
        if opt is None:
            continue
        st = ck.get(key)
        if not isinstance(st, dict) or "param_groups" not in st:
            raise ValueError(f"checkpoint entry '{key}' is not a torch.optim.Adam state_dict: the optimizer state "
                             "cannot be restored (re-save the checkpoint with this version, or drop the entry)")
        opt.load_state_dict(st)
    return ck
