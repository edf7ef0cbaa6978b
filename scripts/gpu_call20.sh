TAG=${1:-r04i}
mkdir -p gpurun_out
for i in 1 2 3; do ./gpurun_variants/cudainit; done > gpurun_out/cudainit_$TAG.txt 2>&1
CUDA_VISIBLE_DEVICES=0 ./gpurun_variants/cudainit >> gpurun_out/cudainit_$TAG.txt 2>&1
CUDA_MODULE_LOADING=EAGER ./gpurun_variants/cudainit >> gpurun_out/cudainit_$TAG.txt 2>&1
nvidia-smi -q | grep -i "persistence" >> gpurun_out/cudainit_$TAG.txt 2>&1
cat gpurun_out/cudainit_$TAG.txt
# a tool run with the trace: where its start-up goes
python - <<'PY' 2>&1 | tail -30
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from synth import files
with tempfile.TemporaryDirectory() as d:
    a = files.make_split_dataset(os.path.join(d, "s"), seed=3, n_clusters=2000, pairs_per_cluster=100, n_chrom=8, genes_per_chrom=40)
    for rep in range(2):
        t0 = time.perf_counter()
        p = subprocess.run(["defuse_b200/bin/dosplitalign"] + a + ["-a", os.path.join(d, "o.tmp")], capture_output=True, env=dict(os.environ, DFB_TRACE="1"))
        print("run", rep, round(time.perf_counter() - t0, 3), "s rc", p.returncode)
        print(p.stderr.decode()[-2500:])
PY
