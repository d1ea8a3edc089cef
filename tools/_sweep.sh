python bench.py --steps 10 --warmup 3 > gpurun_out/bench18.log 2>&1; tail -1 gpurun_out/bench18.log
