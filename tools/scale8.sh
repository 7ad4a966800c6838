# N = 1 and N = 8 on the same box, driver settings (20 steps, 5 warm-up), no weak-scaling leg
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s8_1.json 2>gpurun_out/s8_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-weak > gpurun_out/s8_8.json 2>gpurun_out/s8_8.err
RDC_L2_EVICT_FIRST=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-weak > gpurun_out/s8_8_nohint.json 2>gpurun_out/s8_8_nohint.err
