"""Stub the third-party modules the reference imports but this image lacks, so the
UNMODIFIED reference package under /root/reference can be imported in the build container."""
import sys
import types
from pathlib import Path

# The reference's own Python package: the source tree in the build container, or its `pip install --no-deps --target
# baseline/_ref` copy (git-ignored, but it travels to the GPU box with the snapshot), whichever exists.
_ROOT = Path(__file__).resolve().parents[1]
import os
_CANDIDATES = ([Path(os.environ["BLOKUS_REF_DIR"])] if os.environ.get("BLOKUS_REF_DIR") else []) + \
    [Path("/root/reference"), _ROOT / "baseline" / "_ref"]
REFERENCE = next((p for p in _CANDIDATES
                  if (p / "blokus_rl" / "alphazero" / "mcts.py").exists()), Path("/root/reference"))


def available() -> bool:
    return (REFERENCE / "blokus_rl" / "alphazero" / "mcts.py").exists()


def load_reference_mcts():
    """blokus_rl/alphazero/mcts.py imports only math + numpy: load it by file path, no stubs needed."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_mcts", REFERENCE / "blokus_rl" / "alphazero" / "mcts.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    import logging
    mod("coloredlogs", install=lambda *a, **k: None, ColoredFormatter=logging.Formatter)
    gym = mod("gymnasium", make=lambda *a, **k: None, Env=object)
    gym.wrappers = mod("gymnasium.wrappers", RecordEpisodeStatistics=lambda e, *a, **k: e, RecordVideo=lambda e, *a, **k: e)
    gym.vector = mod("gymnasium.vector", SyncVectorEnv=object)
    gym.spaces = mod("gymnasium.spaces", Discrete=object, Box=object)
    plt = mod("matplotlib.pyplot")
    mod("tensorboard")
    mod("torch.utils.tensorboard", SummaryWriter=lambda *a, **k: types.SimpleNamespace(add_scalar=lambda *a, **k: None, add_text=lambda *a, **k: None,
                                                                                         add_graph=lambda *a, **k: None, close=lambda: None))
    mod("matplotlib", pyplot=plt)
    mod("torchsummary", summary=lambda *a, **k: None)
    mod("pytablewriter", MarkdownTableWriter=object)
    mod("imageio", mimsave=lambda *a, **k: None)
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
