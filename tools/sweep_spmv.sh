run() { echo "== $*"; env "$@" timeout 60 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms_step %.2f spmv_ms %.4f GB/s %.0f frac %.3f asm %.2f solve %.2f' % (d['ms_per_step'], r['ms_per_launch'], r['achieved'], r['frac'], d['phases_ms_per_step']['assemble'], d['phases_ms_per_step']['solve']))
except Exception as e: print('FAILED', e)"; }
for cfg in "$@"; do run $cfg; done
