#!/bin/bash
# round-2 GPU session ZP: PDL variants A/B (class mask SS2_PDL: 1 = SIMT kernels, 2 = tcgen05 kernels; trigger at the top
# of the tcgen05 kernels (libss2.so) or after the producer's last TMA load (libss2_late.so, -DSS2_PDL_LATE))
mkdir -p gpurun_out
BQ="--no-cpu-baseline --no-gpu-eager"
L=$PWD/stabstitch2_b200/libss2_late.so
run() { name=$1; shift; env "$@" timeout 600 python bench.py $BQ > gpurun_out/r2zp_$name.json 2> gpurun_out/r2zp_$name.err; }
for i in 1 2; do
run base_$i SS2_PDL=0
run small_$i SS2_PDL=1
run late3_$i SS2_PDL=3 SS2_LIB=$L
run late2_$i SS2_PDL=2 SS2_LIB=$L
done
run serial_base SS2_PDL=0 SS2_NET_OVERLAP=0
run serial_top3 SS2_PDL=3 SS2_NET_OVERLAP=0
run serial_late3 SS2_PDL=3 SS2_NET_OVERLAP=0 SS2_LIB=$L
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2zp_*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value %.1f ms %.3f e2e %.1f frac %.4f' % (d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac')), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
