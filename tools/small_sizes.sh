# whole-step time at the per-rank sizes of 8-, 4- and 2-GPU runs of the headline mesh, on one GPU
for n in 60 75 95; do
  python bench.py --n $n --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('n=$n', round(d['value'],2), 'ms', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'its', d.get('krylov_its_per_step'), 'chk', d['solution_check']['l2'])"
done
