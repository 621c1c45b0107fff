#!/bin/bash
# Round-end measurement pass: tools/final_bench.sh <N gpus> <tag>  ->  gpurun_out/<tag>_*  (copied to profiles/ by hand)
N=$1; tag=$2
run() {  # run <name> <bench args...>
  name=$1; shift
  if [ "$N" = 1 ]; then
    python bench.py "$@" > gpurun_out/${tag}_${name}_${N}gpu.json 2> gpurun_out/${tag}_${name}_${N}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" \
      > gpurun_out/${tag}_${name}_${N}gpu.json 2> gpurun_out/${tag}_${name}_${N}gpu.err
  fi
  python - gpurun_out/${tag}_${name}_${N}gpu.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[1], "value %.4g %s" % (d.get("value", 0), d.get("unit")), "ms/step", d.get("ms_per_step"), "e2e ms", e.get("ms_per_step"),
          "dp_check", (d.get("dp_check") or {}).get("ok"), "roofline", (d.get("roofline") or {}).get("frac"))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
}
if [ "$N" = 1 ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_gputests_${N}gpu.log; cat gpurun_out/${tag}_gputests_${N}gpu.log
  run bench_cfg2 --steps 2000 --warmup 5
  run bench_cfg2_reference_arm --impl reference --steps 3 --warmup 1
  run bench_cfg3 --workload cfg3
  run bench_cfg4 --workload cfg4
  run bench_cfg5 --workload cfg5
  run bench_cfg5_reference_arm --workload cfg5 --impl reference --steps 3 --warmup 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_ncu_launches_cfg2.csv \
      python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-throughput > gpurun_out/${tag}_ncu_run.log 2>&1
else
  python -m pytest tests/test_gpu_engine.py -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_gputests_${N}gpu.log; cat gpurun_out/${tag}_gputests_${N}gpu.log
  run bench_cfg2 --steps 2000 --warmup 5
  run bench_cfg3 --workload cfg3
  run bench_cfg5 --workload cfg5
fi
