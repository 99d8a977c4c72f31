"""com_marl_b200 — B200-native batched rollout engine for Com-MARL (PredatorPrey / Coverage step +
communication graph / packet loss + comm-GNN policy forward).  See DESIGN.md."""
from .scenario import ScenarioSpec  # noqa: F401

__all__ = ["ScenarioSpec", "BatchedEnv", "PredatorPreyWrapper", "CoverageWrapper", "CommCategoricalMLPPolicy",
           "DecCategoricalMLPPolicy", "RolloutEngine", "eval_model", "DeviceRolloutSampler"]


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import com_marl_b200` stays cheap
    if name in ("BatchedEnv", "PredatorPreyWrapper", "CoverageWrapper"):
        from . import envs
        return getattr(envs, name)
    if name in ("CommCategoricalMLPPolicy", "DecCategoricalMLPPolicy"):
        from . import policy
        return getattr(policy, name)
    if name == "RolloutEngine":
        from .rollout import RolloutEngine
        return RolloutEngine
    if name == "eval_model":
        from .evaluate import eval_model
        return eval_model
    if name == "DeviceRolloutSampler":
        from .sampler import DeviceRolloutSampler
        return DeviceRolloutSampler
    raise AttributeError(name)
