# the BASELINE.json configurations that fit one GPU
for spec in "adpm 28" "pihna 28" "pihna 119"; do set -- $spec
  python bench.py --model $1 --n $2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 n=$2', round(d['value'],2), 'steps/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'its', d.get('krylov_its_per_step'), 'spmv frac', round(d['roofline']['frac'],3), 'gmres30', (d.get('ksp_gmres30') or {}).get('ms_per_step'))"
done
