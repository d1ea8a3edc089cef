python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 0; do PARESIS_TILE_CONFIG=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench15_$c.log 2>&1; python - <<EOF2
import json
r=json.loads(open("gpurun_out/bench15_$c.log").read().strip().splitlines()[-1])
print("cfg $c value", round(r["value"]), "ms/step", round(r["ms_per_step"],3), "e2e", round(r["e2e"]["value"]), {k:round(v["ms_per_launch"]*1e3,1) for k,v in r["kernel_shares"].items()})
EOF2
done
