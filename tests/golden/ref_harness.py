"""Import harness for the UNMODIFIED reference (cnuns/Com-MARL) — golden-vector generation only.

The reference is pure Python and imports in this container once seven absent third-party
modules are stubbed in ``sys.modules``.  None of the stubbed modules does any
arithmetic on the hot path.  This file is used ONLY by ``tests/golden/make_golden.py`` (run by
hand in the build container, where ``/root/reference`` is mounted) to record input/output
vectors of the reference itself; nothing under ``tests/`` that runs on the GPU box imports it.

Nothing here is copied from the reference: the stubs are the minimal duck types its imports
touch (``gym.Env``, ``gym.Wrapper``, ``gym.spaces.Box/Discrete`` ...).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("COM_MARL_REFERENCE", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_gym():
    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = shape
            self.dtype = dtype

        def sample(self):
            raise NotImplementedError

        def contains(self, x):
            return True

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = n

        def sample(self):
            return int(np.random.randint(self.n))

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            low = np.asarray(low, dtype=dtype)
            high = np.asarray(high, dtype=dtype)
            super().__init__(low.shape, dtype)
            self.low, self.high = low, high

    class Env:
        metadata = {}
        reward_range = (-float("inf"), float("inf"))
        spec = None

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env
            self.action_space = env.action_space
            self.observation_space = env.observation_space
            self.metadata = getattr(env, "metadata", {})

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        def step(self, action):
            return self.env.step(action)

        def reset(self, **kw):
            return self.env.reset(**kw)

        @property
        def unwrapped(self):
            return self.env.unwrapped

    def np_random(seed=None):
        return np.random.RandomState(seed), (0 if seed is None else int(seed))

    def hash_seed(seed=None, max_bytes=8):
        return int(seed) * 2654435761 % (2 ** 32)

    space_mod = _mod("gym.spaces.space", Space=Space)
    utils_spaces = _mod("gym.spaces.utils", flatdim=lambda s: int(np.prod(s.shape)) if s.shape else s.n)
    spaces = _mod("gym.spaces", Space=Space, Discrete=Discrete, Box=Box, space=space_mod, utils=utils_spaces)
    seeding = _mod("gym.utils.seeding", np_random=np_random, hash_seed=hash_seed)
    gutils = _mod("gym.utils", seeding=seeding)

    class _Registry:
        def all(self):
            return []

    registration = _mod("gym.envs.registration", register=lambda *a, **k: None)
    genvs = _mod("gym.envs", registry=_Registry(), registration=registration)
    wrappers = _mod("gym.wrappers", Monitor=object)
    _mod("gym", Env=Env, Wrapper=Wrapper, spaces=spaces, utils=gutils, envs=genvs, wrappers=wrappers,
         Space=Space)


def _install_akro():
    class _A:
        pass

    class Discrete(_A):
        def __init__(self, n):
            self.n = n
            self.flat_dim = n
            self.shape = ()

    class Box(_A):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low)
            self.high = np.asarray(high)
            self.shape = self.low.shape
            self.flat_dim = int(np.prod(self.shape))

    def from_gym(space):
        if hasattr(space, "n"):
            return Discrete(space.n)
        return Box(space.low, space.high)

    for cls, qn in ((_A, "Space"), (Discrete, "Discrete"), (Box, "Box")):      # picklable like the real akro classes (checkpoints)
        cls.__module__, cls.__qualname__ = "akro", qn
    _mod("akro", Discrete=Discrete, Box=Box, from_gym=from_gym, Space=_A, Dict=_A, Tuple=_A, Image=_A)


def _install_dowel():
    class _Logger:
        def log(self, *a, **k):
            pass

        def add_output(self, *a, **k):
            pass

        def remove_all(self):
            pass

        def has_output_type(self, *a):
            return False

        def push_prefix(self, *a):
            pass

        def pop_prefix(self, *a):
            pass

        def dump_all(self, *a):
            pass

        def prefix(self, *a):
            import contextlib
            return contextlib.nullcontext()

    class _Tabular:
        def __init__(self):
            self.rows = {}

        def record(self, k, v):
            self.rows[k] = v

        def clear(self):
            self.rows.clear()

        def prefix(self, *a):
            import contextlib
            return contextlib.nullcontext()

    _mod("dowel", logger=_Logger(), tabular=_Tabular(), StdOutput=object, TextOutput=object,
         CsvOutput=object, TensorBoardOutput=object, LogOutput=object, TabularInput=_Tabular)


def _install_misc():
    class _Bar:
        def __init__(self, *a, **k):
            pass

        def update(self, *a, **k):
            pass

        def stop(self):
            pass

        active = False

    _mod("pyprind", ProgBar=_Bar)
    _mod("send2trash", send2trash=lambda *a, **k: None)

    def remote(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    _mod("ray", remote=remote, init=lambda *a, **k: None, is_initialized=lambda: False)


def install_stubs():
    if "gym" not in sys.modules:
        _install_gym()
    if "akro" not in sys.modules:
        _install_akro()
    if "dowel" not in sys.modules:
        _install_dowel()
    if "pyprind" not in sys.modules:
        _install_misc()
    # The reference imports a file it does not ship (com_marl/torch/modules/__init__.py:5).
    name = "com_marl.torch.modules.categorical_lstm_module"
    if name not in sys.modules:
        _mod(name, CategoricalLSTMModule=type("CategoricalLSTMModule", (), {}))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "envs", "ma_gym"))


def load_reference():
    """Returns a namespace with the reference classes/modules used for golden generation."""
    install_stubs()
    import envs as ref_envs  # noqa: E402
    import custom_implement.env_communication as ref_comm  # noqa: E402
    import custom_implement.gilbert_elliot_loss_model as ref_ge  # noqa: E402
    from com_marl.torch.policies.comm_categorical_mlp_policy import CommCategoricalMLPPolicy  # noqa: E402
    from garage.envs import GarageEnv  # noqa: E402
    import envs.ma_gym.envs.predator_prey.predator_prey as ref_pp_mod  # noqa: E402
    import envs.ma_gym.envs.coverage.coverage as ref_co_mod  # noqa: E402

    ns = types.SimpleNamespace(
        PredatorPreyWrapper=ref_envs.PredatorPreyWrapper,
        CoverageWrapper=ref_envs.CoverageWrapper,
        comm=ref_comm, ge=ref_ge, pp_mod=ref_pp_mod, co_mod=ref_co_mod,
        CommCategoricalMLPPolicy=CommCategoricalMLPPolicy, GarageEnv=GarageEnv)
    return ns


def scenario_params(scenario, map_size, sen, den, cap=2, loss=0.0, **over):
    """The ``params`` dict the reference runners hand to the env constructors.

    Follows exp_runners/env_uitils.py:171-217 (get_parser_to_args), predatorprey/utils_pp.py:69-89,
    coverage/utils_co.py:75-113 and the sizing rule n_agents = int(int(den*100)*(map/10)^2)
    (SURVEY.md §8 "Config -> concrete sizes").
    """
    base = int(den * 100)
    n_agents = int(base * (map_size / 10) ** 2)

    def _i(v):  # get_parser_to_args turns integral floats into ints
        return int(v) if float(v) == int(v) else v

    p = dict(grid_size=map_size, Rsen=sen, n_agents=n_agents, n_gcn_layers=2, loss_apply=1,
             curriculum_learning=0, calc_diameter=False, mode="train", trpl=_i(loss), tepl=_i(loss),
             channelType=None, Pgb=0.0196, Pbg=0.282, GE_INIT=1, trRcom=9, teRcom=9,
             n_eval_episodes=50, env_param_print=0, rm=0)
    if scenario == "pp":
        p.update(n_preys=n_agents, load=cap, max_env_steps=200, capture_reward=10, step_cost=0.1,
                 penalty=0, n_groups=1, n_nodes=None)
    else:
        p.update(n_groups=3, add_clock=0, obstComplex="Easy", load=2, rendering=0, capture_reward=2,
                 step_cost=0, penalty=1, revisit_penalty=0.5, lazy_penalty=1, max_env_steps=400)
    p.update(over)
    return p
