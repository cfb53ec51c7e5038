for flags in "-DLAT_PREFETCH=0" "-DLAT_MINB=6" "-DLAT_NCELL=8" "-DLAT_RPI=2 -DLAT_MINB=6" "-DLAT_PREFETCH=4"; do
  SS2_NVCC_FLAGS="$flags" SS2_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  python bench.py --steps 6 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$flags', d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
done
