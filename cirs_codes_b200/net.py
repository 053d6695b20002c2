"""Parameter containers with tianshou's constructor surface and state_dict names.

Mirrors tianshou/utils/net/common.py:25-197 (MLP, Net) and tianshou/utils/net/discrete.py:11-114 (Actor, Critic)
as far as CIRS uses them (CIRS-RL-kuaishou.py:245-258): ``Net(state_dim, hidden_sizes=[64, 64])`` shared by
``Actor(net, n_items)`` and ``Critic(net)``.  They are plain torch modules living on the HOST so that the
reference's script code (orthogonal init over ``.modules()``, ``torch.optim.Adam(list(actor.parameters()) + ...)``,
``state_dict()`` / ``load_state_dict``) runs unchanged; they are never used for compute -- PPOPolicy packs them into
its flat device buffer (params.py) and every forward / backward runs in the CUDA kernels.
"""
import numpy as np
import torch
from torch import nn


class MLP(nn.Module):
    def __init__(self, input_dim, output_dim=0, hidden_sizes=(), device=None):
        super().__init__()
        sizes = [input_dim] + list(hidden_sizes)
        model = []
        for i, o in zip(sizes[:-1], sizes[1:]):
            model += [nn.Linear(i, o), nn.ReLU()]
        if output_dim > 0:
            model += [nn.Linear(sizes[-1], output_dim)]
        self.output_dim = output_dim or sizes[-1]
        self.model = nn.Sequential(*model)


class Net(nn.Module):
    def __init__(self, state_shape, action_shape=0, hidden_sizes=(), device="cpu", **kwargs):
        super().__init__()
        assert not kwargs.get("concat") and not kwargs.get("dueling_param"), "not on the CIRS hot path"
        self.device = device
        input_dim = int(np.prod(state_shape))
        action_dim = int(np.prod(action_shape))
        self.model = MLP(input_dim, action_dim, hidden_sizes)
        self.output_dim = self.model.output_dim
        self.hidden_sizes = list(hidden_sizes)
        self.input_dim = input_dim


class Actor(nn.Module):
    def __init__(self, preprocess_net, action_shape, hidden_sizes=(), softmax_output=True,
                 preprocess_net_output_dim=None, device="cpu"):
        super().__init__()
        assert not hidden_sizes and softmax_output, "CIRS uses Actor(net, n_items) with softmax output"
        self.device = device
        self.preprocess = preprocess_net
        self.output_dim = int(np.prod(action_shape))
        self.last = MLP(getattr(preprocess_net, "output_dim", preprocess_net_output_dim), self.output_dim)

    def to(self, *a, **k):  # parameters stay on the host; compute happens in the kernels
        return self


class ActorProb(nn.Module):
    """tianshou/utils/net/continuous.py:120-199 as CIRS-RL-taobao.py:206 uses it: ``ActorProb(net, action_shape,
    max_action=...)`` -- a ``mu`` layer on the shared trunk and a state-independent ``sigma_param``."""

    def __init__(self, preprocess_net, action_shape, hidden_sizes=(), max_action=1.0, device="cpu", unbounded=False,
                 conditioned_sigma=False, preprocess_net_output_dim=None):
        super().__init__()
        assert not hidden_sizes and not unbounded and not conditioned_sigma, "not on the CIRS hot path"
        self.device = device
        self.preprocess = preprocess_net
        self.output_dim = int(np.prod(action_shape))
        self.mu = MLP(getattr(preprocess_net, "output_dim", preprocess_net_output_dim), self.output_dim)
        self.sigma_param = nn.Parameter(torch.zeros(self.output_dim, 1))
        self._max = float(max_action)

    def to(self, *a, **k):
        return self


class Critic(nn.Module):
    def __init__(self, preprocess_net, hidden_sizes=(), last_size=1, preprocess_net_output_dim=None, device="cpu"):
        super().__init__()
        assert not hidden_sizes and last_size == 1
        self.device = device
        self.preprocess = preprocess_net
        self.output_dim = last_size
        self.last = MLP(getattr(preprocess_net, "output_dim", preprocess_net_output_dim), last_size)

    def to(self, *a, **k):
        return self


def orthogonal_init(*modules):
    """CIRS-RL-kuaishou.py:250-254."""
    for mod in modules:
        for m in mod.modules():
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight)
                nn.init.zeros_(m.bias)
