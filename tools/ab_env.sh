# A/B of one environment-controlled option on the same box: tools/ab_env.sh RDC_VEC_REVERSE "0 1" [bench args]
var=$1; vals=$2; shift 2
for rep in 1 2; do for v in $vals; do
  env $var=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$v', round(d['value'],2), 'steps/s', round(d['ms_per_step'],3), 'ms', {k: round(x,3) for k,x in d['phases_ms_per_step'].items()}, 'chk', d['solution_check']['l2'])"
done; done
