"""CPU check of the SB3 boundary: with `stable_baselines3.common.vec_env.VecEnv` importable,
`UAVVecEnv` IS a VecEnv subclass with no abstract method left -- so `DQN._wrap_env` / `VecFrameStack` take the batch as
it is instead of wrapping it in `DummyVecEnv([lambda: env])` (dqn.py:1276-1288).  SB3 is absent from the build image;
tests/_sb3_stub restates the abstract surface of SB3 2.7's VecEnv.  Runs in a subprocess so that the stub never leaks
into the other tests' imports.  (tests/test_gpu_boundary.py constructs and steps the subclass on a GPU.)"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

PROBE = r"""
import sys
sys.path.insert(0, %r)
sys.path.insert(0, %r)
from __graft_entry__ import import_package
pkg = import_package()
from stable_baselines3.common.vec_env import VecEnv
import uavsim_b200.vec_env as V
assert V.HAVE_SB3 and issubclass(pkg.UAVVecEnv, VecEnv), "not a VecEnv subclass"
assert not getattr(pkg.UAVVecEnv, "__abstractmethods__", None), pkg.UAVVecEnv.__abstractmethods__
for m in ("reset", "step_async", "step_wait", "close", "get_attr", "set_attr", "env_method", "env_is_wrapped", "seed", "step"):
    assert callable(getattr(pkg.UAVVecEnv, m)), m
print("ok")
"""


def test_uavvecenv_is_a_vecenv_subclass_when_sb3_is_importable():
    r = subprocess.run([sys.executable, "-c", PROBE % (str(ROOT / "tests" / "_sb3_stub"), str(ROOT))],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_duck_typed_surface_without_sb3():
    from __graft_entry__ import import_package
    pkg = import_package()
    import sys as _s
    V = _s.modules["uavsim_b200.vec_env"]
    assert V.HAVE_SB3 is False          # the build image has no SB3: same methods, object base
    for m in ("reset", "step_async", "step_wait", "close", "get_attr", "set_attr", "env_method", "env_is_wrapped",
              "seed", "step"):
        assert callable(getattr(pkg.UAVVecEnv, m)), m
