#!/bin/bash
# SURVEY 8d config 5 (tracer-count scaling at the tx0.1v3 column depth on the 1200x800 sample, passive tracers centred and
# on lw_lim) and the config 2 / config 3 shapes: one line each into gpurun_out/r2_config_sweeps.jsonl
mkdir -p gpurun_out
OUT=gpurun_out/r2_config_sweeps.jsonl
: > $OUT
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); p=d['phases_ms_per_step']
print(json.dumps({'workload':'$1','nt':$2,'passive_advect':'$3','ms_per_step':round(d['ms_per_step'],3),'solver_iterations':d.get('solver_iterations'),'phases_ms':{k:round(p[k],3) for k in ('TRACER_UPDATE','VMIX_TRACER_IMPLICIT','STATE','MOMENTUM_COLUMN','MOMENTUM_FINISH','SOLVER','HALO')}}))" >> $OUT; }
for adv in centered lw_lim; do
  for nt in 2 4 8 16 34; do
    if [ $adv = lw_lim ] && [ $nt = 2 ]; then continue; fi
    timeout 600 python bench.py --workload tx_sample --nt $nt --passive-advect $adv --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | line "tx_sample 1200x800x62" $nt $adv
  done
done
timeout 300 python bench.py --workload gx3v7 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | line "gx3v7 100x116x60 upwind3+del2+const vmix, ChronGear (config 2)" 2 upwind3
timeout 300 python bench.py --workload gx1v7 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | line "gx1v7 320x384x60 GM, P-CSI (config 3)" 2 centered
timeout 600 python bench.py --pbc --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | line "tx0.1v3 3600x2400x62 partial bottom cells (config 4, production variant)" 2 centered
cat $OUT
