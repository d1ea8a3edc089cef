#!/bin/bash
# Same-box A/B of library variants and knobs: label, env assignments, extra bench flags.  One line per run:
# whole-job image-sets/s and step time (3 positions in flight unless --slots), then the kernels timed one position at a time.
run() {
  label=$1; shift
  envs=$1; shift
  env $envs python bench.py --skip-extras --no-cpu-baseline "$@" 2>>gpurun_out/ab_err.log | LABEL=$label python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['kernel_shares']
print('%-16s %6d  %.2f ms  hop2 %.1f  hop1 %.1f  det %.1f  cut %.1f' % (os.environ['LABEL'], round(d['value']), d['ms_per_step'], s['refract_sample_ref_hop']['ms_per_launch']*1e3, s['refract_membrane_hop']['ms_per_launch']*1e3, s['detect']['ms_per_launch']*1e3, s['raster_spheres']['ms_per_launch']*1e3))"
}
# edit below: one `run LABEL "ENV=..." [bench flags]` per line, variants built with
#   python -m paresis_b200.build --variant NAME -DMACRO=VALUE   ->  PARESIS_B200_LIB=libparesis_b200_NAME.so
run base X=1
run rows14 PARESIS_LEAN_ROWS=14
run base X=1
