"""Loggers around the hot path, with the reference's interfaces so that CIRS-RL-kuaishou.py:262-334 keeps its calls:

  BasicLogger            tianshou/utils/log_tools.py:84-200 -- what onpolicy_trainer's ``logger=`` argument receives
                         (``log_train_data / log_test_data / log_update_data / save_data / restore_data``); writes through
                         any object with ``add_scalar(key, value, global_step=)`` (a tensorboard ``SummaryWriter``).
  LoggerCallback_Policy  util/utils.py:84-136 -- the ``policy.callbacks`` entry that writes one ``Epoch: [e], Info: [{...}]``
                         line per epoch with ``num_test``, ``CV``, ``CV_turn``, ``ctr``, ``len_tra``, ``R_tra`` and the
                         ``ifeat_*`` rates for the three test collectors (the lines the paper's result tables are parsed
                         from, reproduce_results_of_our_paper/results_all_methods/*.log).  logzero is not a dependency:
                         the line goes to ``logging.getLogger("cirs")`` and, when given, is appended to ``logger_path``.
"""
import logging
import re

_log = logging.getLogger("cirs")


class BasicLogger:
    def __init__(self, writer, train_interval=1000, test_interval=1, update_interval=1000, save_interval=1):
        self.writer = writer
        self.train_interval, self.test_interval = train_interval, test_interval
        self.update_interval, self.save_interval = update_interval, save_interval
        self.last_log_train_step = self.last_log_test_step = -1
        self.last_log_update_step = self.last_save_step = -1

    def write(self, key, x, y, **kwargs):
        self.writer.add_scalar(key, y, global_step=x)

    def log_train_data(self, collect_result, step):
        if collect_result["n/ep"] > 0:
            collect_result["rew"] = collect_result["rews"].mean()
            collect_result["len"] = collect_result["lens"].mean()
            if step - self.last_log_train_step >= self.train_interval:
                self.write("train/n/ep", step, collect_result["n/ep"])
                self.write("train/rew", step, collect_result["rew"])
                self.write("train/len", step, collect_result["len"])
                self.last_log_train_step = step

    def log_test_data(self, collect_result, step):
        assert collect_result["n/ep"] > 0
        rews, lens = collect_result["rews"], collect_result["lens"]
        rew, rew_std, len_, len_std = rews.mean(), rews.std(), lens.mean(), lens.std()
        collect_result.update(rew=rew, rew_std=rew_std, len=len_, len_std=len_std)
        if step - self.last_log_test_step >= self.test_interval:
            self.write("test/rew", step, rew)
            self.write("test/len", step, len_)
            self.write("test/rew_std", step, rew_std)
            self.write("test/len_std", step, len_std)
            self.last_log_test_step = step

    def log_update_data(self, update_result, step):
        if step - self.last_log_update_step >= self.update_interval:
            for k, v in update_result.items():
                self.write(k, step, v)
            self.last_log_update_step = step

    def save_data(self, epoch, env_step, gradient_step, save_checkpoint_fn=None):
        if save_checkpoint_fn and epoch - self.last_save_step >= self.save_interval:
            self.last_save_step = epoch
            save_checkpoint_fn(epoch, env_step, gradient_step)
            self.write("save/epoch", epoch, epoch)
            self.write("save/env_step", env_step, env_step)
            self.write("save/gradient_step", gradient_step, gradient_step)

    def restore_data(self):
        """The reference re-reads its tensorboard event file; a writer that keeps ``scalars`` (e.g. ScalarRecorder)
        restores from memory, anything else starts from zero."""
        rec = getattr(self.writer, "scalars", None)
        if not rec or "save/epoch" not in rec:
            return 0, 0, 0
        epoch = rec["save/epoch"][-1][0]
        self.last_save_step = self.last_log_test_step = epoch
        gradient_step = rec["save/gradient_step"][-1][0]
        self.last_log_update_step = gradient_step
        env_step = rec.get("save/env_step", [(0, 0)])[-1][0]
        self.last_log_train_step = env_step
        return epoch, env_step, gradient_step


class ScalarRecorder:
    """A dependency-free writer for BasicLogger: keeps {key: [(step, value), ...]} and optionally appends
    ``step<TAB>key<TAB>value`` lines to a file."""

    def __init__(self, path=None):
        self.scalars, self.path = {}, path

    def add_scalar(self, key, value, global_step=None):
        self.scalars.setdefault(key, []).append((global_step, float(value)))
        if self.path:
            with open(self.path, "a") as f:
                f.write(f"{global_step}\t{key}\t{float(value)}\n")


class LoggerCallback_Policy:
    def __init__(self, logger_path=None, force_length=10):
        self.LOCAL_PATH, self.force_length = logger_path, force_length

    def on_epoch_begin(self, epoch, **kwargs):
        pass

    def on_train_begin(self, **kwargs):
        pass

    def on_train_end(self, **kwargs):
        pass

    def format(self, epoch, results):
        """util/utils.py:98-133: the Info dictionary of one epoch."""
        results_all = {}
        for prefix in ["", "NX_0_", f"NX_{self.force_length}_"]:
            num_test = results["n/ep"]
            len_tra = results[prefix + "n/st"] / num_test
            r_tra = results[prefix + "rew"]
            res = {"num_test": num_test, prefix + "CV": f"{results[prefix + 'CV']:.5f}",
                   prefix + "CV_turn": f"{results[prefix + 'CV_turn']:.5f}", prefix + "ctr": f"{r_tra / len_tra:.5f}",
                   prefix + "len_tra": len_tra, prefix + "R_tra": r_tra}
            pattern = re.compile(prefix + "ifeat_")
            results_all.update(res)
            results_all.update({k: v for k, v in results.items() if re.match(pattern, k)})
        return "Epoch: [{}], Info: [{}]".format(epoch, results_all)

    def on_epoch_end(self, epoch, results=None, **kwargs):
        line = self.format(epoch, results)
        _log.info(line)
        if self.LOCAL_PATH:
            with open(self.LOCAL_PATH, "a") as f:
                f.write(line + "\n")
        return line
