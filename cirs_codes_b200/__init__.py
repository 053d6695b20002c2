"""cirs_codes_b200 -- the CIRS rollout + PPO-update hot path as sm_100a CUDA kernels behind the reference's
Python interfaces (tianshou Collector / VectorEnv / policy.update, deepctr_torch feature columns).

Importing the package is cheap and CPU-safe; constructing any of the compute classes requires the built
libcirs_b200.so and a CUDA device (there is no CPU fallback on the product path).
"""
from .inputs import (DenseFeat, SparseFeat, SparseFeatP, VarLenSparseFeat, build_input_features,  # noqa: F401
                     compute_input_dim, get_dataset_columns, get_feature_names)
from .data import Batch, VectorReplayBuffer  # noqa: F401
from .net import Actor, ActorProb, Critic, Net, orthogonal_init  # noqa: F401
from .env import KuaishouVectorEnv, TaobaoVectorEnv, VirtualTBVectorEnv  # noqa: F401
from .state_tracker import StateTrackerTransformer  # noqa: F401
from .policy import PPOPolicy  # noqa: F401
from .collector import Collector  # noqa: F401
from .collector_set import CollectorSet  # noqa: F401
from .trainer import onpolicy_trainer, save_checkpoint, load_checkpoint  # noqa: F401
from .loggers import BasicLogger, LoggerCallback_Policy, ScalarRecorder  # noqa: F401
from .evaluation import Callback_Coverage_Count  # noqa: F401

__all__ = ["DenseFeat", "SparseFeat", "SparseFeatP", "VarLenSparseFeat", "build_input_features",
           "compute_input_dim", "get_dataset_columns", "get_feature_names", "Batch", "VectorReplayBuffer", "Actor",
           "ActorProb", "Critic", "Net", "orthogonal_init", "KuaishouVectorEnv", "TaobaoVectorEnv", "VirtualTBVectorEnv", "StateTrackerTransformer", "PPOPolicy",
           "Collector", "CollectorSet", "onpolicy_trainer", "save_checkpoint", "load_checkpoint", "BasicLogger",
           "LoggerCallback_Policy", "ScalarRecorder", "Callback_Coverage_Count"]
