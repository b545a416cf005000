mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
(timeout 600 python tests/golden/make_fwd_golden.py gpurun_out/fwd_ref_triton.npz > gpurun_out/fwd_golden.log 2>&1; echo "rc=$?" >> gpurun_out/fwd_golden.log)
tail -n 8 gpurun_out/fwd_golden.log
(timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_all_gpu.log)
tail -n 6 gpurun_out/t_all_gpu.log
(timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_ours.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ours.log)
tail -n 2 gpurun_out/bench_ours.log
(timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log)
tail -n 2 gpurun_out/bench_ref.log
(timeout 200 python bench.py --config 13b-int3 --steps 30 --warmup 5 > gpurun_out/bench_13b.log 2>&1; echo "rc=$?" >> gpurun_out/bench_13b.log)
(timeout 200 python bench.py --config prefill --steps 5 --warmup 3 > gpurun_out/bench_prefill.log 2>&1; echo "rc=$?" >> gpurun_out/bench_prefill.log)
tail -n 2 gpurun_out/bench_13b.log gpurun_out/bench_prefill.log
(timeout 420 ncu --set full --clock-control none --import-source on -k regex:llama_decode_mega -s 2 -c 1 -f -o gpurun_out/r2_mega_final python tools/prof_mega.py 7b > gpurun_out/ncu_final.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_final.log)
tail -n 3 gpurun_out/ncu_final.log
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_launches.log)
tail -n 2 gpurun_out/ncu_launches.log
