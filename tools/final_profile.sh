#!/bin/bash
# Round-end evidence on the GPU box: full GPU test suite, default bench line, ncu launch list of the bench command,
# one full ncu capture of a steady-state position (each ncu pass only after its command ran clean without ncu).
set -o pipefail
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -4 gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err || { echo "bench failed"; tail -5 gpurun_out/bench_default.err; exit 1; }
tail -c 600 gpurun_out/bench_default.log; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --jobs-per-step 2 --no-cpu-baseline --skip-extras > gpurun_out/ncu_launches_final.log 2>&1
python tools/one_position.py 3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name 'regex:refract_|detect_tile|membrane_from_field' --launch-skip 4 -c 4 \
    -o gpurun_out/prof_final -f python tools/one_position.py 3 > gpurun_out/ncu_full_final.log 2>&1
tail -1 gpurun_out/ncu_full_final.log
# Fresnel model: launch list of two positions at 4096^2 and one full capture of its two kernels
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fresnel_launches_final.csv \
    python tools/fresnel_probe.py 4096 2 > gpurun_out/ncu_fresnel_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name 'regex:line_convolve|post_lines' --launch-skip 8 -c 2 \
    -o gpurun_out/prof_fresnel_final -f python tools/fresnel_probe.py 4096 2 > gpurun_out/ncu_fresnel_full.log 2>&1
tail -1 gpurun_out/ncu_fresnel_full.log
