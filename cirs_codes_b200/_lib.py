"""ctypes binding of libcirs_b200.so (C ABI declared in include/cirs_b200.h).

The product path has NO CPU fallback: importing this module without the built library, or calling an entry point
without a CUDA device, raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` (or
``make -C cirs_codes_b200/csrc``).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcirs_b200.so")
ABI_VERSION = 9
MAX_LAYERS = 4
HIDDEN = 64

fp = C.c_void_p  # device pointers travel as void*


class KuaishouEnvStruct(C.Structure):
    _fields_ = [("n_env", C.c_int32), ("max_turn", C.c_int32), ("num_leave_compute", C.c_int32),
                ("n_user", C.c_int32), ("n_item", C.c_int32), ("simulated", C.c_int32), ("version", C.c_int32),
                ("leave_threshold", C.c_float), ("tau", C.c_float), ("gamma_exposure", C.c_float),
                ("r_decay", C.c_float),
                ("normed_mat", fp), ("mat", fp), ("cat_mask", fp), ("alpha_u", fp), ("beta_i", fp), ("dist", fp),
                ("user", fp), ("turn", fp), ("hist", fp), ("cum_rew", fp), ("seen", fp)]


class MMOEStruct(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("h1", C.c_int32), ("h2", C.c_int32), ("n_expert", C.c_int32),
                ("expert_dim", C.c_int32),
                ("lin_w", fp), ("w1t", fp), ("b1", fp), ("w2t", fp), ("b2", fp), ("wet", fp), ("be", fp), ("wgt", fp),
                ("bg", fp), ("tower", fp), ("out_bias", C.c_float)]


class TaobaoEnvStruct(C.Structure):
    _fields_ = [("n_env", C.c_int32), ("max_turn", C.c_int32), ("num_leave_compute", C.c_int32),
                ("version", C.c_int32), ("map_action", C.c_int32), ("act_low", C.c_float), ("act_high", C.c_float),
                ("leave_threshold", C.c_double), ("tau", C.c_double), ("gamma_exposure", C.c_double),
                ("um", MMOEStruct),
                ("user", fp), ("turn", fp), ("hist", fp), ("prev_rew", fp), ("cum_rew", fp)]


class VirtualTBStruct(C.Structure):
    _fields_ = [(k, fp) for k in ("g1t", "g1b", "g2t", "g2b", "a1t", "a1b", "a2t", "a2b", "a3t", "a3b")]


class EncoderLayerStruct(C.Structure):
    _fields_ = [(k, fp) for k in ("in_wt", "in_b", "out_wt", "out_b", "l1_wt", "l1_b", "l2_wt", "l2_b",
                                  "n1_w", "n1_b", "n2_w", "n2_b")]


class TrackerWeightsStruct(C.Structure):
    _fields_ = [("d", C.c_int32), ("nhead", C.c_int32), ("d_hid", C.c_int32), ("nlayers", C.c_int32),
                ("dim_state", C.c_int32), ("max_len", C.c_int32), ("d_user_in", C.c_int32),
                ("d_item_in", C.c_int32), ("n_user", C.c_int32), ("n_item", C.c_int32),
                ("emb_user", fp), ("emb_item", fp), ("user_wt", fp), ("user_b", fp), ("gate_wt", fp),
                ("gate_b", fp), ("pe", fp), ("layer", EncoderLayerStruct * MAX_LAYERS), ("dec_wt", fp),
                ("dec_b", fp), ("flat", fp), ("n_flat", C.c_int64)]


class PolicyWeightsStruct(C.Structure):
    _fields_ = [("dim_state", C.c_int32), ("n_action", C.c_int32), ("ld_action", C.c_int32),
                ("w1t", fp), ("b1", fp), ("w2t", fp), ("b2", fp), ("w3t", fp), ("b3", fp), ("wv", fp), ("bv", fp),
                ("flat", fp), ("n_flat", C.c_int64), ("n_trunk", C.c_int64), ("sigma", fp),
                ("max_action", C.c_float)]


class PPOConfigStruct(C.Structure):
    _fields_ = [("eps_clip", C.c_float), ("vf_coef", C.c_float), ("ent_coef", C.c_float),
                ("max_grad_norm", C.c_float), ("value_clip", C.c_int32), ("norm_adv", C.c_int32),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float)]


class UserModelStruct(C.Structure):
    _fields_ = [(k, fp) for k in ("emb_user", "emb_item", "emb_feat", "lin_user", "lin_item", "lin_feat",
                                  "lin_dense", "w1", "b1", "w2", "b2", "w_last")] + \
               [("out_bias", C.c_float), ("emb_dim", C.c_int32), ("n_feat", C.c_int32), ("n_dense", C.c_int32),
                ("hidden", C.c_int32)]


i32, i64, u64, f64 = C.c_int32, C.c_int64, C.c_uint64, C.c_double
P = C.POINTER

# name -> (restype, argtypes); every name here must be declared in include/cirs_b200.h (tests check both ways)
PROTOTYPES = {
    "cirs_last_error": (C.c_char_p, []),
    "cirs_abi_version": (i32, []),
    "cirs_launch_count": (i64, []),
    "cirs_profile_enable": (None, [i32]),
    "cirs_profile_report": (i32, [C.c_char_p, i32]),
    "cirs_kuaishou_reset": (i32, [P(KuaishouEnvStruct), i32, fp, fp, fp, fp]),
    "cirs_kuaishou_step": (i32, [P(KuaishouEnvStruct), i32, fp, fp, fp, fp, fp, i32, fp, fp, fp, fp, i32, fp]),
    "cirs_taobao_reset": (i32, [P(TaobaoEnvStruct), i32, fp, fp, fp, fp]),
    "cirs_taobao_step": (i32, [P(TaobaoEnvStruct), i32, fp, fp, fp, fp, fp, fp, i32, fp, fp, fp, fp, fp, i32, fp]),
    "cirs_virtualtb_generate_users": (i32, [P(VirtualTBStruct), i32, fp, fp, u64, u64, fp, fp]),
    "cirs_virtualtb_step": (i32, [P(TaobaoEnvStruct), P(VirtualTBStruct), i32, fp, fp, fp, u64, u64, fp, fp, fp, i32,
                                  fp]),
    "cirs_tracker_step": (i32, [P(TrackerWeightsStruct), i32, i32, fp, fp, fp, i32, fp, fp, fp, fp, fp, fp, i64,
                                fp, i32, fp, fp, fp]),
    "cirs_tracker_train_workspace_bytes": (i64, [P(TrackerWeightsStruct), i32, i64]),
    "cirs_tracker_train": (i32, [P(TrackerWeightsStruct), P(TrackerWeightsStruct), i32, i32, fp, fp, fp, fp, fp,
                                 fp, i32, fp, fp, i32, fp, fp, fp, i64, i32, fp]),
    "cirs_tracker_train_fused_enable": (None, [i32]),
    "cirs_actor_workspace_bytes": (i64, [i32, i32]),
    "cirs_actor_sample": (i32, [P(PolicyWeightsStruct), i32, fp, fp, fp, i64, fp, u64, u64, fp, i32, fp, fp, fp, fp,
                                fp, fp]),
    "cirs_policy_eval": (i32, [P(PolicyWeightsStruct), i32, fp, fp, fp, fp, fp, fp, fp]),
    "cirs_policy_eval_dev": (i32, [P(PolicyWeightsStruct), i32, fp, fp, fp, fp, fp, fp, fp, fp]),
    "cirs_actorprob_sample": (i32, [P(PolicyWeightsStruct), i32, fp, fp, fp, i64, fp, u64, u64, fp, i32, fp, fp, fp,
                                    fp, fp]),
    "cirs_actorprob_eval": (i32, [P(PolicyWeightsStruct), i32, fp, fp, fp, fp, fp, fp]),
    "cirs_rollout_taobao": (i32, [P(TaobaoEnvStruct), P(TrackerWeightsStruct), P(PolicyWeightsStruct), fp, fp, fp,
                                  i32, fp, fp, fp, fp, fp, fp, fp, fp, fp, i32, u64, fp, i32, i32, i32, fp]),
    "cirs_rollout_workspace_bytes": (i64, [i32, i32]),
    "cirs_rollout_kuaishou": (i32, [P(KuaishouEnvStruct), P(TrackerWeightsStruct), P(PolicyWeightsStruct), fp, fp, fp,
                                    fp, fp, fp, fp, fp, i32, fp, fp, fp, fp, fp, fp, fp, fp, i32, u64, fp, i32, i32,
                                    i32, fp, fp]),
    "cirs_compute_returns": (i32, [i32, i32, fp, fp, fp, fp, fp, f64, f64, fp, fp, fp, fp, fp, fp]),
    "cirs_rms_update": (i32, [fp, fp, fp]),
    "cirs_adv_stats": (i32, [i32, fp, fp, fp, fp, fp]),
    "cirs_ppo_workspace_bytes": (i64, [i32, i32]),
    "cirs_ppo_minibatch": (i32, [P(PolicyWeightsStruct), P(PolicyWeightsStruct), P(PPOConfigStruct), i32, i32, fp,
                                 fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp]),
    "cirs_ppo_learn": (i32, [P(PolicyWeightsStruct), P(PolicyWeightsStruct), fp, fp, P(PPOConfigStruct), i32, i32,
                             fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, i64, fp, fp, fp, fp, fp, fp, i32, fp]),
    "cirs_update_plan": (i32, [i32, i32, fp, fp, fp, fp]),
    "cirs_gather_i32": (i32, [fp, fp, fp, i32, fp]),
    "cirs_zero": (i32, [fp, i64, fp]),
    "cirs_coverage_count": (i32, [i32, fp, fp, i32, fp, fp, fp, fp]),
    "cirs_comm_unique_id": (i32, [fp]),
    "cirs_comm_create": (i32, [fp, i32, i32, P(fp)]),
    "cirs_comm_destroy": (i32, [fp]),
    "cirs_comm_allreduce": (i32, [fp, fp, i64, i32, fp]),
    "cirs_comm_group_begin": (i32, [fp]),
    "cirs_comm_group_end": (i32, [fp]),
    "cirs_head_tc_enable": (None, [i32]),
    "cirs_head_tc_timeout": (i32, []),
    "cirs_head_tc_timeout_peek": (i32, [fp, fp]),
    "cirs_user_model_workspace_bytes": (i64, [i32, i32, i32]),
    "cirs_user_model_predict_all": (i32, [P(UserModelStruct), i32, fp, i32, fp, fp, fp, i32, fp, fp, fp, fp]),
    "cirs_user_model_tc_enable": (None, [i32]),
    "cirs_user_model_timeout": (i32, []),
    "cirs_user_model_debug_phases": (i32, [P(i64)]),
    "cirs_head_tc_debug_phases": (i32, [i32, P(i64), i32]),
    "cirs_clip_adam": (i32, [fp, fp, fp, fp, i64, i64, P(PPOConfigStruct), fp, fp, fp]),
}

_lib = None


class CirsError(RuntimeError):
    pass


def load():
    """dlopen the library (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CirsError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                        "(there is no CPU fallback for the product path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.cirs_abi_version() != ABI_VERSION:
        raise CirsError(f"ABI mismatch: library {lib.cirs_abi_version()} vs binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "C-ABI arguments must be contiguous CUDA tensors"
    return t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke an int-returning entry point and raise CirsError with cirs_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise CirsError(f"{name} failed ({rc}): {lib.cirs_last_error().decode()}")


def profile_report():
    """{kernel name: (launches, total ms)} since profiling was enabled; clears the records."""
    buf = C.create_string_buffer(1 << 16)
    load().cirs_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise CirsError("cirs_codes_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
