#!/bin/bash
# Same-box A/B of library variants and knobs: label, env assignments, extra bench flags.  One line per run:
# whole-job image-sets/s and step time (3 positions in flight unless --slots), then the kernels timed one position at a time.
run() {
  label=$1; shift
  envs=$1; shift
  env $envs python bench.py --skip-extras --no-cpu-baseline "$@" 2>/dev/null | LABEL=$label python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['kernel_shares']
print('%-16s %6d  %.2f ms  hop2 %.1f  hop1 %.1f  det %.1f  cut %.1f' % (os.environ['LABEL'], round(d['value']), d['ms_per_step'], s['refract_sample_ref_hop']['ms_per_launch']*1e3, s['refract_membrane_hop']['ms_per_launch']*1e3, s['detect']['ms_per_launch']*1e3, s['raster_spheres']['ms_per_launch']*1e3))"
}
R=PARESIS_LEAN_ROWS=32
run base X=1
run r16+mb3 "$R PARESIS_B200_LIB=libparesis_b200_mb3.so"
run base X=1
run r16+mb3 "$R PARESIS_B200_LIB=libparesis_b200_mb3.so"
run r16+mb2 "$R PARESIS_B200_LIB=libparesis_b200_mb2.so"
run r16+mb3s5 "$R PARESIS_B200_LIB=libparesis_b200_mb3s5.so"
run r16+mb3+s4 "$R PARESIS_B200_LIB=libparesis_b200_mb3.so" --slots 4
run r16+mb3+s5 "$R PARESIS_B200_LIB=libparesis_b200_mb3.so" --slots 5
run r16+mb3+s6 "$R PARESIS_B200_LIB=libparesis_b200_mb3.so" --slots 6
run r16+mb3 "$R PARESIS_B200_LIB=libparesis_b200_mb3.so"
run base X=1
