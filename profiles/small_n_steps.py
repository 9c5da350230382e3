"""Device-resident tick rate at the small, launch-bound configurations (cfg1 N = 64, cfg2 N = 4,096, N = 16,384)."""
import json
import os
import sys
import time
import tomllib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
from sfm_b200 import native, synth          # noqa: E402
from tests.gpu_util import make_context     # noqa: E402

cfg = tomllib.load(open(os.path.join(ROOT, 'carla-social-force-model_b200', 'config', 'sfm_config.toml'), 'rb'))
for k, n in ((1, None), (2, None), (3, 16384)):
    w = synth.make_config(k, n=n)
    ctx = make_context(w, cfg)
    ctx.step(50, True)
    ctx.synchronize()
    t0 = time.perf_counter()
    steps = 400
    ctx.step(steps, True)
    ctx.synchronize()
    sec = (time.perf_counter() - t0) / steps
    s = ctx.stats()
    print(json.dumps({'config': w.name, 'n_pedestrians': w.n, 'us_per_tick': sec * 1e6, 'agent_steps_per_s': w.n / sec,
                      'launches_per_tick': s['launches'] / (steps + 50)}), flush=True)
