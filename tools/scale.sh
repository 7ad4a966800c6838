# usage: bash tools/scale.sh "1 2 4" [extra bench args]   -> one summary line per N (run under gpurun --gpus maxN)
NS="$1"; shift
for N in $NS; do
  if [ $N = 1 ]; then timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale_$N.json 2>gpurun_out/scale_err_$N.log
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale_$N.json 2>gpurun_out/scale_err_$N.log; fi
  tail -1 gpurun_out/scale_$N.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'steps/s %.1f ms %.2f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['krylov_its_per_step'], 'spmv_ms %.4f' % d['roofline']['ms_per_launch'])
except Exception as e: print('FAILED', e)"
  grep "rdc trace rank 0" gpurun_out/scale_err_$N.log | tail -1
done
