# per-rank problem size of an 8-GPU run (n=60: 1.3 M tets) on ONE GPU: kernel efficiency at small sizes
run() { echo "== $*"; env "$@" RDC_TRACE=1 timeout 60 python bench.py --n 60 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | grep -E "rdc trace|^\{" | tail -2 | cut -c1-330; }
for cfg in "$@"; do run $cfg; done
