#!/bin/bash
# quick A/B on the GPU box: kernel tests of the hop kernels, then the bench line in short form for
# the first tile kernel (PARESIS_TILE_CONFIG=1) and the lean one (0, production)
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_end_to_end.py -m gpu -x -q 2>&1 | tail -5
for cfg in ${CONFIGS:-1 0 1 0}; do
PARESIS_TILE_CONFIG=$cfg python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/quick_$cfg.log 2>&1
python - $cfg <<'EOF2'
import json, sys
cfg = sys.argv[1]
try:
    r=json.loads(open("gpurun_out/quick_%s.log" % cfg).read().strip().splitlines()[-1])
    print("cfg", cfg, "value", round(r["value"]), "ms/step", round(r["ms_per_step"],3), "e2e", round(r["e2e"]["value"]), {k:round(v["ms_per_launch"]*1e3,1) for k,v in r["kernel_shares"].items()})
except Exception as e:
    print("cfg", cfg, "failed", e); print(open("gpurun_out/quick_%s.log" % cfg).read()[-2000:])
EOF2
done
