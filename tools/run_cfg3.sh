N=$1
if [ "$N" = 1 ]; then
  python bench.py --workload cfg3 > gpurun_out/r02i_bench_cfg3_1gpu.json 2> gpurun_out/r02i_bench_cfg3_1gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload cfg3 > gpurun_out/r02i_bench_cfg3_${N}gpu.json 2> gpurun_out/r02i_bench_cfg3_${N}gpu.err
fi
python -c "
import json; d=json.loads(open('gpurun_out/r02i_bench_cfg3_${N}gpu.json').read().strip().splitlines()[-1]); e=d['e2e']; print($N, d['ms_per_step'], 'e2e', e['ms_per_step'], 'mono', e['monolithic']['ms_per_step'], d['checks'])"
