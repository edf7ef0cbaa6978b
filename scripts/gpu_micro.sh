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
PY
cat gpurun_out/micro_${1:-m}.txt
