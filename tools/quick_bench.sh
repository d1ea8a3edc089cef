#!/bin/bash
# quick A/B on the GPU box: kernel tests of the hop kernels, then the bench line in short form
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_end_to_end.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/quick.log 2>&1
python - <<'EOF2'
import json
r=json.loads(open("gpurun_out/quick.log").read().strip().splitlines()[-1])
print("value", round(r["value"]), "ms/step", round(r["ms_per_step"],3), "e2e", round(r["e2e"]["value"]), {k:round(v["ms_per_launch"]*1e3,1) for k,v in r["kernel_shares"].items()})
EOF2
