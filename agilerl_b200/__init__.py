"""agilerl_b200 — B200-native (sm_100a) implementation of AgileRL's population-parallel off-policy
``learn()`` hot path behind the reference's Python API.  CUDA only; no CPU fallback."""
__version__ = "0.2.0"

# reference LEAF module -> the module of this package that replaces it (BASELINE.json: "Subsystems replaced")
_LEAF = {
    "agilerl.components.replay_buffer": "agilerl_b200.components.replay_buffer",
    "agilerl.components.segment_tree": "agilerl_b200.components.segment_tree",
    "agilerl.components.sampler": "agilerl_b200.components.sampler",
    "agilerl.components.data": "agilerl_b200.components.data",
    "agilerl.algorithms.dqn": "agilerl_b200.algorithms.dqn",
    "agilerl.algorithms.dqn_rainbow": "agilerl_b200.algorithms.dqn_rainbow",
    "agilerl.algorithms.td3": "agilerl_b200.algorithms.td3",
    "agilerl.algorithms.ddpg": "agilerl_b200.algorithms.td3",
    "agilerl.algorithms.maddpg": "agilerl_b200.algorithms.maddpg",
    "agilerl.components.multi_agent_replay_buffer": "agilerl_b200.components.multi_agent_replay_buffer",
    "agilerl.hpo.tournament": "agilerl_b200.hpo.tournament",
    "agilerl.hpo.mutation": "agilerl_b200.hpo.mutation",
}
# reference PACKAGE -> ours: the alias package exports our names and serves everything else from the reference
_PACKAGES = {
    "agilerl.components": "agilerl_b200.components",
    "agilerl.algorithms": "agilerl_b200.algorithms",
    "agilerl.hpo": "agilerl_b200.hpo",
}


def _reference_exports(ref_name: str, ref_dir: str) -> dict:
    """name -> submodule, from the ``from .sub import A, B`` lines of the reference package's ``__init__``."""
    import os
    import re
    init = os.path.join(ref_dir, "__init__.py")
    table: dict = {}
    if not os.path.isfile(init):
        return table
    src = open(init).read()
    for m in re.finditer(r"from\s+(?:%s)?\.?([\w\.]+)\s+import\s+(\([^)]*\)|[^\n]+)" % re.escape(ref_name + "."), src):
        sub, names = m.group(1), m.group(2).strip("()")
        for n in re.split(r"[,\s]+", names):
            n = n.strip()
            if n and n.isidentifier() and n != "as":
                table.setdefault(n, sub)
    return table


def install(include_driver: bool | None = None) -> list[str]:
    """Make ``import agilerl.<replaced module>`` resolve to this package, so code written against the reference —
    first of all the UNCHANGED ``agilerl/training/train_off_policy.py`` with its ``from agilerl.algorithms import DDPG, DQN,
    TD3, RainbowDQN``, ``from agilerl.components import ...`` lines and the ``isinstance(memory, PrioritizedReplayBuffer)``
    dispatch of ``agilerl.components.sampler`` (sampler.py:71-72) — runs on the CUDA path.  Call it BEFORE importing
    ``agilerl.training``.

    Where the reference package is importable, only the replaced leaf modules (``agilerl.components.{replay_buffer,
    segment_tree, sampler, data, multi_agent_replay_buffer}``, ``agilerl.algorithms.{dqn, dqn_rainbow, td3, ddpg, maddpg}``, ``agilerl.hpo.{tournament,
    mutation}``) and the three packages that re-export them are aliased; every other name of those packages
    (``agilerl.algorithms.PPO``, ``agilerl.components.RolloutBuffer`` ...) still comes from the reference's own
    files, loaded lazily, and ``agilerl.utils``, ``agilerl.training``, ``agilerl.networks``, ``agilerl.modules`` stay
    the reference's.  Where it is not importable a bare ``agilerl`` namespace is created and
    ``agilerl.training.train_off_policy`` maps to this package's restatement of the driver (``include_driver`` forces
    either behaviour).  Returns the aliased module names."""
    import importlib
    import importlib.util
    import sys
    import types
    have_ref = "agilerl" in sys.modules or importlib.util.find_spec("agilerl") is not None
    if not have_ref:
        root = types.ModuleType("agilerl")
        root.__path__ = []
        sys.modules["agilerl"] = root
    done = []
    for ref_name, ours_name in _PACKAGES.items():
        ours = importlib.import_module(ours_name)
        ref_dirs = []
        if have_ref:
            try:
                spec = importlib.util.find_spec(ref_name)
            except (ImportError, ValueError, AttributeError):
                spec = None
            ref_dirs = list(getattr(spec, "submodule_search_locations", None) or [])
        pkg = types.ModuleType(ref_name)
        pkg.__path__ = ref_dirs                       # non-replaced submodules: the reference's own files
        pkg.__package__ = ref_name
        for k in getattr(ours, "__all__", [k for k in vars(ours) if not k.startswith("_")]):
            setattr(pkg, k, getattr(ours, k))
        table = {}
        for d in ref_dirs:
            table.update(_reference_exports(ref_name, d))

        def __getattr__(name, _table=table, _ref=ref_name, _pkg=pkg):
            sub = _table.get(name)
            if sub is None:
                raise AttributeError(f"module {_ref!r} has no attribute {name!r}")
            value = getattr(importlib.import_module(f"{_ref}.{sub}"), name)
            setattr(_pkg, name, value)
            return value
        pkg.__getattr__ = __getattr__
        sys.modules[ref_name] = pkg
        setattr(sys.modules[ref_name.rpartition(".")[0]], ref_name.rpartition(".")[2], pkg)
        done.append(ref_name)
    leaf = dict(_LEAF)
    if include_driver is True or (include_driver is None and not have_ref):
        for extra in ("training", "utils"):
            p = types.ModuleType(f"agilerl.{extra}")
            p.__path__ = []
            sys.modules[f"agilerl.{extra}"] = p
            setattr(sys.modules["agilerl"], extra, p)
        leaf["agilerl.training.train_off_policy"] = "agilerl_b200.training.train_off_policy"
        leaf["agilerl.utils.utils"] = "agilerl_b200.utils.utils"
    for ref_name, ours_name in leaf.items():
        mod = importlib.import_module(ours_name)
        sys.modules[ref_name] = mod
        parent, _, child = ref_name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], child, mod)
        done.append(ref_name)
    return done
