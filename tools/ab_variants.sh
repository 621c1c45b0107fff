#!/bin/bash
# A/B of library variants built with KVAE_NVCC_FLAGS=... KVAE_LIB_OUT=build_variants/lib_<name>.so (see tools/README.md):
#   tools/ab_variants.sh <tag> default tpb32 tpb64 ...   -> gpurun_out/<tag>_<name>.json + one summary line each
tag=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset KVAE_LIB; else export KVAE_LIB=$PWD/build_variants/lib_$v.so; fi
  python bench.py --steps 2000 --warmup 5 --no-e2e --no-cpu --no-throughput > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - "$v" gpurun_out/${tag}_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ks = {k: round(v * 1e6, 1) for k, v in d["roofline"]["kernel_seconds"].items()}
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 5), [round(b, 5) for b in d["ms_per_step_blocks"]], "kernels us", ks)
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
