TAG=${1:-r04j}
mkdir -p gpurun_out
python - <<PY 2>&1 | tail -5
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from synth import files
with tempfile.TemporaryDirectory() as d:
    a = files.make_split_dataset(os.path.join(d, "s"), seed=3, n_clusters=2000, pairs_per_cluster=100, n_chrom=8, genes_per_chrom=40)
    for rep in range(2):
        t0 = time.perf_counter()
        p = subprocess.run(["defuse_b200/bin/dosplitalign"] + a + ["-a", os.path.join(d, "o.tmp")], capture_output=True, env=dict(os.environ, DFB_TRACE="1"))
        print("run", rep, round(time.perf_counter() - t0, 3), "s rc", p.returncode)
        open("gpurun_out/tool_trace_${TAG}_%d.txt" % rep, "wb").write(p.stderr)
PY
