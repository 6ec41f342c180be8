"""Developer probe (GPU): float32 build against the float64 C oracle with the margin-aware harness; prints the
per-key errors, the flips with the oracle's margins, and the worst rewards.  tests/test_gpu_parity.py asserts the
same cases."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import import_package  # noqa: E402
from parity_util import run_parity_case  # noqa: E402

pkg = import_package()
E = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 48
cases = {
    "n50_replay": dict(N=50, grid=(1000, 1000), p_collect=0.3, start="random"),
    "n50_philox": dict(N=50, grid=(1000, 1000), p_collect=0.3, start="random", draws="philox"),
    "n20_replay": dict(N=20, grid=(500, 500), p_collect=0.3),
    "n20_philox": dict(N=20, grid=(500, 500), p_collect=0.3, draws="philox"),
    "n20_collect98": dict(N=20, grid=(200, 200), p_collect=0.98, start=(100.0, 100.0), draws="philox"),
    "dr_n20_philox": dict(N=20, P=20, grid=(300, 300), p_collect=0.4, start="random", domain_rand=True, obs_slots=50,
                          draws="philox"),
    "dr_n40_replay_restarts": dict(N=40, P=40, grid=(300, 300), p_collect=0.4, start="random", domain_rand=True,
                                   obs_slots=50, auto_reset=True, params_over=dict(max_battery=6.0),
                                   ref_kwargs=dict(max_battery=6.0)),
}
for name, kw in cases.items():
    res = run_parity_case(pkg, E=E, steps=T, seed=97, precision="f32", margin_aware=True, **kw)
    res.pop("first_bad", None) if not res["unmargined_flips"] else None
    print(name, json.dumps(res, default=str), flush=True)
