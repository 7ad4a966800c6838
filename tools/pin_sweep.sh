for n in 60 75 95 119; do for mode in 0 1 2; do
  RDC_L2_MODE=$mode python bench.py --n $n --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/pin_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('n=$n mode=$mode', round(d['value'],2), 'ms', round(d['ms_per_step'],3), d.get('phases_ms_per_step'), 'frac', round(d['roofline']['frac'],3))"
done; done
