"""CPU tests of the C-ABI boundary: the library loads and exports exactly what include/blokus_b200.h declares.
No compute call is made (no GPU here)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from blokus_rl_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_functions():
    text = (ROOT / "include" / "blokus_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(blk_[a-z_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from blokus_rl_b200 import _lib
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in blokus_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_abi_version_and_struct_sizes(lib):
    from blokus_rl_b200 import _lib
    assert lib.blk_abi_version() == 3
    assert C.sizeof(_lib.BlkConfig) == 16 and C.sizeof(_lib.BlkInfo) == 44
    assert C.sizeof(_lib.BlkStepArgs) == 144 and C.sizeof(_lib.BlkRolloutArgs) == 112


def test_ctypes_structs_have_the_c_layout(tmp_path):
    """Every struct of include/blokus_b200.h: sizeof and the offset of its last field as gcc lays them out == ctypes'."""
    import subprocess
    from blokus_rl_b200 import _lib
    pairs = [("blk_config", _lib.BlkConfig, "device"), ("blk_info", _lib.BlkInfo, "sm_count"),
             ("blk_step_args", _lib.BlkStepArgs, "csr_offset"), ("blk_rollout_args", _lib.BlkRolloutArgs, "options"),
             ("blk_puct_forest", _lib.BlkPuctForest, "node_front"), ("blk_puct_expand_args", _lib.BlkPuctExpandArgs, "fuse_backup"),
             ("blk_puct_search_args", _lib.BlkPuctSearchArgs, "virtual_loss")]
    body = "".join(f'printf("%zu %zu\\n", sizeof({c}), offsetof({c}, {last}));' for c, _, last in pairs)
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "blokus_b200.h"\nint main(void){' + body + "return 0;}")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-I", str(ROOT / "include"), str(src), "-o", str(exe)])
    got = [tuple(int(x) for x in ln.split()) for ln in subprocess.check_output([str(exe)], text=True).splitlines()]
    want = [(C.sizeof(t), getattr(t, last).offset) for _, t, last in pairs]
    assert got == want


def test_no_cpu_fallback(lib):
    """Without a GPU the engine must refuse to start instead of silently computing on the host."""
    import torch
    from blokus_rl_b200 import BlokusEngine, EngineError, _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(EngineError):
        BlokusEngine(20, 4)
    h = C.c_void_p()
    cfg = _lib.BlkConfig(20, 4, 0, 0)
    rc = lib.blk_create(C.byref(cfg), C.byref(h))
    assert rc in (-2, -3) and not h.value
    assert b"fallback" in lib.blk_last_error() or b"cuda" in lib.blk_last_error().lower()


def test_argument_validation_without_gpu(lib):
    from blokus_rl_b200 import _lib
    h = C.c_void_p()
    for bad in (_lib.BlkConfig(21, 4, 0, 0), _lib.BlkConfig(20, 3, 0, 0), _lib.BlkConfig(20, 4, 7, 0)):
        assert lib.blk_create(C.byref(bad), C.byref(h)) == -1
    assert lib.blk_step(None, None, None) == -1
    assert lib.blk_reset(None, None, 0, None) == -1


def test_product_package_never_imports_oracle():
    for f in (ROOT / "blokus_rl_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_missing_library_fails_loudly(tmp_path):
    """No .so -> EngineError naming the build command; never a silent fallback."""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from blokus_rl_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.EngineError as e:\n    print('LOUD', 'no CPU fallback' in str(e) or 'missing' in str(e))\n" % str(ROOT))
    env = dict(os.environ, BLOKUS_B200_LIB=str(tmp_path / "nope.so"))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert "LOUD True" in out.stdout, out.stdout + out.stderr


def test_header_is_valid_c99_and_a_plain_c_program_links(lib, tmp_path):
    """The ABI is C, not C++: examples/c_abi_smoke.c compiles with gcc -std=c99 and links against the .so
    (it is run on the GPU box by tests/test_gpu_parity.py::test_plain_c_program_runs)."""
    import subprocess
    exe = tmp_path / "c_abi_smoke"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include",
           str(ROOT / "examples" / "c_abi_smoke.c"), "-o", str(exe), "-L", str(ROOT / "blokus_rl_b200"),
           "-lblokus_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{ROOT / 'blokus_rl_b200'}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert exe.exists()
