mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm', torch.cuda.device_count())"
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --config 65b-tp --steps 20 --warmup 3 > gpurun_out/bench_tp2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_tp2.log)
tail -n 3 gpurun_out/bench_tp2.log | cut -c1-1500
