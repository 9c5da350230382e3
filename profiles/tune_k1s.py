"""K1s variant sweep on the GPU box: every build/libsfm_*.so (profiles/build_k1s_variants.sh) timed on the pair force of
a flat N = 65,536 crowd (CUDA events inside the library), and compared with the first variant's output."""
import glob, os, subprocess, sys, json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, tomllib, json
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
import numpy as np
from sfm_b200 import native, synth
cfg = tomllib.load(open(os.path.join(ROOT, 'carla-social-force-model_b200/config/sfm_config.toml'), 'rb'))
n = int(os.environ.get('TUNE_N', '65536'))
w = synth.make_config(5, n=n)
ctx = native.Context(0)
ctx.set_params(native.params_from_config(cfg, 0.05))
ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
out = np.empty((n, 3))
for _ in range(3): ctx.force(native.PEDESTRIAN, out)
ctx.reset_stats(); ctx.set_profiling(True)
for _ in range(8): ctx.force(native.PEDESTRIAN, out)
s = ctx.stats()
np.save(os.environ['TUNE_OUT'], out)
print(json.dumps({'ms': s['ms_pairs'] / s['pair_launches'], 'fixup': s['fixup_rows']}))
'''.replace('ROOT', repr(ROOT))

import numpy as np
ref = None
libs = sorted(glob.glob(os.path.join(ROOT, 'build', 'libsfm_*.so')))
libs.sort(key=lambda p: (os.path.basename(p) != 'libsfm_base.so', p))
for lib in libs:
    name = os.path.basename(lib)[7:-3]
    out = f'/tmp/tune_{name}.npy'
    env = dict(os.environ, SFM_LIB=lib, TUNE_OUT=out)
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        print(f'{name}: FAILED {r.stderr[-300:]}')
        continue
    res = json.loads(r.stdout.strip().splitlines()[-1])
    f = np.load(out)
    if ref is None:
        ref = f
    d = np.abs(f - ref)
    rel = d / (1e-5 + 1e-4 * np.abs(ref))
    print(f"{name:>18}: {res['ms']:.3f} ms  fixup_rows={res['fixup']}  max|dF| vs base={d.max():.2e}  max err/tol={rel.max():.3f}", flush=True)
