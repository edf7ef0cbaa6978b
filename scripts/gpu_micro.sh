# usage: bash scripts/gpu_micro.sh <tag>   -- issue-rate microbenchmarks + the sweep-body variants (no profiler)
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/micro_${1:-m}.txt
import defuse_b200 as d
ctx = d.default_context(0)
info = ctx.device_info()
print(info)
names = {0:'VIADDMNMX.S16x2',1:'VIMNMX.U16x2',2:'VIMNMX3.S16x2',3:'LOP3',4:'IMAD',5:'IADD3',6:'PRMT',7:'dp_cell_body_s16x2',8:'SHFL.UP',9:'VIADDMNMX.S32',
         10:'HMNMX2',11:'HADD2',12:'HFMA2',13:'HSET2',14:'VIADDMNMX+HMNMX2 1:1',15:'VIADDMNMX+HFMA2 1:1',16:'VIADDMNMX+HADD2 1:1'}
for k in range(17):
    r, ms = ctx.microbench_issue_rate(k, 2000)
    print('%-24s %8.3f Gwarp-instr/s  %7.3f ms  -> %.2f warp-instr/clk/SM @%d MHz' % (names[k], r/1e9, ms, r/info['sm_count']/(info['clock_khz']*1e3), info['clock_khz']//1000))
# the sweep's paired-step body, S = 13 rows per lane, HR of them with the fp16 mismatch indicator (kind 100+HR with the
# row-maximum sink, 200+HR without): warp row-steps / s; one row-step = 64 cell updates per warp
for k in (100, 103, 105, 107, 109, 111, 113, 200, 207, 213):
    r, ms = ctx.microbench_issue_rate(k, 20000)
    print('sweep body S=13 HR=%-2d sink=%d  %8.3f Gwarp-rowsteps/s  %7.3f ms  -> %.2f clk/row-step/SMSP, %.2f TCUPS' % (
        k % 100, 1 if k < 200 else 0, r/1e9, ms, info['sm_count']*4*(info['clock_khz']*1e3)/r, r*64/1e12))
PY
cat gpurun_out/micro_${1:-m}.txt
