mkdir -p gpurun_out
python - <<'PY'
import os, sys, subprocess, time
sys.path.insert(0, '.')
from synth import files
d = '/tmp/tt'
args = files.make_split_dataset(d, seed=1, n_clusters=200, pairs_per_cluster=100)
for rep in range(2):
    t0 = time.time()
    p = subprocess.run(['defuse_b200/bin/dosplitalign'] + args + ['-a', d + '/o.tmp'], env=dict(os.environ, DFB_TRACE='1'), capture_output=True)
    print('wall', time.time() - t0, 'rc', p.returncode)
    print(p.stderr.decode()[-3000:])
PY
