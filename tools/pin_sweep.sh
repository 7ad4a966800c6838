# whole-step time with and without the evict_first L2 priority on the SpMV's operator stream, one GPU, at the per-rank
# sizes of 8-, 4-, 2- and 1-GPU runs of the headline mesh.  (The evict_last / persisting-L2 / access-window variants
# recorded in profiles/r2_l2_policy_sweep.log were experiment code of that commit and are not in the tree.)
for n in 60 75 95 119; do for hint in 0 1; do
  RDC_L2_EVICT_FIRST=$hint python bench.py --n $n --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('n=$n evict_first=$hint', round(d['value'],2), 'steps/s', round(d['ms_per_step'],3), 'ms', d.get('phases_ms_per_step'), 'frac', round(d['roofline']['frac'],3))"
done; done
