"""Checkpoint interchange with the reference.

The reference's snapshotter writes ``<exp_dir>/itrs/itr_%04d.pkl`` = ``pickle.dump(algo.policy)`` — the WHOLE policy object
(garage/experiment/snapshotter.py:100-104, mode 'gap_and_last') — and its evaluation scripts read it back with
``joblib.load`` (exp_runners/testing.py:68-77 ``load_policy``).  Such a pickle names the reference's classes
(``com_marl.torch.policies.*``, ``com_marl.torch.modules.*``, ``garage.torch.modules.*``, ``akro`` / ``gym`` spaces), none of
which exist here.  ``load_reference_checkpoint`` unpickles it through a module-alias shim — every class of those packages
becomes an empty ``nn.Module`` placeholder that just receives its pickled ``__dict__`` (torch's own ``Linear`` /
``Sequential`` / ``ModuleList`` / ``Parameter`` are the real ones) — reads the architecture attributes and the ``state_dict``
off the placeholder tree and builds the matching policy of this package (same parameter names, so ``load_state_dict`` is
exact).  The other direction needs no shim: ``save_state_dict`` writes a plain ``{name: ndarray}`` pickle that a reference
policy object loads with ``load_state_dict`` (identical names and shapes; tests/golden/make_golden_ckpt.py checks it).
"""
import io
import os
import pickle

import numpy as np
import torch
from torch import nn

from .policy import CentralizedCategoricalMLPPolicy, CommCategoricalMLPPolicy, DecCategoricalMLPPolicy
from .spaces import Box, Discrete, EnvSpec

_SHIMMED = ("com_marl", "garage", "akro", "gym", "custom_implement", "envs", "dowel")
_placeholders = {}


def _placeholder(module, name):
    key = (module, name)
    cls = _placeholders.get(key)
    if cls is None:
        # an nn.Module subclass: its __setstate__ takes the pickled __dict__ (with _parameters / _modules / _buffers for the
        # reference's modules; harmless extras for the plain objects such as spaces), and state_dict() walks the tree
        cls = _placeholders[key] = type(name, (nn.Module,), {"__module__": "com_marl_b200.checkpoint.shim." + module,
                                                              "_reference_class": f"{module}.{name}"})
    return cls


class _ShimUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in _SHIMMED:
            return _placeholder(module, name)
        return super().find_class(module, name)


def _attr(obj, *names, default=None):
    for k in names:
        if k in obj.__dict__:
            return obj.__dict__[k]
    return default


def load_reference_checkpoint(path, device="cuda", **policy_kwargs):
    """``itr_%04d.pkl`` of the reference -> CommCategoricalMLPPolicy / DecCategoricalMLPPolicy /
    CentralizedCategoricalMLPPolicy of this package with the checkpoint's weights."""
    with open(path, "rb") as f:
        data = f.read()
    ref = _ShimUnpickler(io.BytesIO(data)).load()
    if isinstance(ref, dict):                       # params.pkl / algo_backup.pkl hold {'algo': ...} or the algo itself
        ref = ref.get("algo", ref)
    if "policy" in getattr(ref, "__dict__", {}) or "policy" in getattr(ref, "_modules", {}):
        ref = ref.__dict__.get("policy") or ref._modules["policy"]
    kind = getattr(type(ref), "_reference_class", type(ref).__name__).rsplit(".", 1)[-1]
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    n = int(_attr(ref, "_n_agents"))
    if kind == "CommCategoricalMLPPolicy":
        D = int(_attr(ref, "_dec_obs_dim", default=sd["encoder._layers.0.linear.weight"].shape[1]))
        L = sum(1 for k in sd if k.startswith("gcn_layers.") and k.endswith(".weight"))
        cls, kw = CommCategoricalMLPPolicy, dict(n_gcn_layers=L, residual=bool(_attr(ref, "residual", default=True)),
                                                 gcn_bias=any(k.startswith("gcn_layers.") and k.endswith(".bias") for k in sd),
                                                 # layer widths / attention type as pickled (env_uitils.py:82-87)
                                                 encoder_hidden_sizes=(int(sd["encoder._layers.0.linear.weight"].shape[0]),),
                                                 embedding_dim=int(sd["encoder._output_layers.0.linear.weight"].shape[0]),
                                                 categorical_mlp_hidden_sizes=tuple(int(sd[f"categorical_output_layer._layers.{i}.linear.weight"].shape[0]) for i in range(3)),
                                                 attention_type="general" if "attention_layer.linear_in.weight" in sd else "dot")
    elif kind == "DecCategoricalMLPPolicy":
        D = int(sd["encoder._layers.0.linear.weight"].shape[1])
        cls, kw = DecCategoricalMLPPolicy, dict(hidden_sizes=(int(sd["encoder._layers.0.linear.weight"].shape[0]),
                                                              int(sd["encoder._output_layers.0.linear.weight"].shape[0]),
                                                              int(sd["_layers.0.linear.weight"].shape[0])))
    elif kind == "CentralizedCategoricalMLPPolicy":
        D = int(sd["_layers.0.linear.weight"].shape[1]) // n
        nl = _attr(ref, "_hidden_nonlinearity")
        cls, kw = CentralizedCategoricalMLPPolicy, dict(hidden_nonlinearity="relu" if getattr(nl, "__name__", "tanh") == "relu" else "tanh",
                                                        hidden_sizes=tuple(int(sd[f"_layers.{i}.linear.weight"].shape[0]) for i in range(3)))
    else:
        raise ValueError(f"{path}: unsupported policy class {kind!r} in the checkpoint")
    kw.update(policy_kwargs)
    env_spec = EnvSpec(Box(np.zeros(n * D, np.float32), np.ones(n * D, np.float32)), Discrete(5))
    pol = cls(env_spec, n, device=device, **kw)
    missing = set(pol.state_dict()) ^ set(sd)
    if missing:
        raise ValueError(f"{path}: parameter names differ from {cls.__name__}: {sorted(missing)}")
    pol.load_state_dict(sd)
    return pol


def load_policy(exp_dir, epoch, device="cuda", **policy_kwargs):
    """exp_runners/testing.py:68-77: ``<exp_dir>/itrs/itr_<epoch>.pkl``, then the zero-padded name"""
    for name in (f"itr_{epoch}.pkl", f"itr_{str(epoch).zfill(4)}.pkl"):
        p = os.path.join(exp_dir, "itrs", name)
        if os.path.isfile(p):
            return load_reference_checkpoint(p, device=device, **policy_kwargs)
    raise FileNotFoundError(f"{exp_dir}/itrs/itr_{str(epoch).zfill(4)}.pkl")


def save_state_dict(policy, path):
    """{parameter name: float32 ndarray} — what a reference policy object takes through
    ``load_state_dict({k: torch.as_tensor(v) ...})``; the names are the reference's own"""
    with open(path, "wb") as f:
        pickle.dump({k: v.detach().cpu().numpy() for k, v in policy.state_dict().items()}, f)
