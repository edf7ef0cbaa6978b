mkdir -p gpurun_out
python - <<'PY'
import os, sys, subprocess, time
sys.path.insert(0, '.')
from synth import files
d = '/tmp/tt2'
margs, sam = files.make_matealign_dataset(d + '/m', seed=4, n_pairs=300000)
for rep in range(3):
    t0 = time.time()
    p = subprocess.run(['defuse_b200/bin/matealign'] + margs, input=sam, env=dict(os.environ, DFB_TRACE='1'), capture_output=True)
    print('matealign wall', time.time() - t0, 'rc', p.returncode)
    print(p.stderr.decode()[-3500:])
PY
