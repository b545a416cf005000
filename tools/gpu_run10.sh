mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
(timeout 600 python tests/golden/make_fwd_golden.py gpurun_out/fwd_ref_triton.npz > gpurun_out/fwd_golden.log 2>&1; echo "rc=$?" >> gpurun_out/fwd_golden.log)
(timeout 900 python tools/refshim/ref_decode_bench.py 32 2047 12 > gpurun_out/ref_decode.log 2>&1; echo "rc=$?" >> gpurun_out/ref_decode.log)
(timeout 600 python tools/refshim/ref_decode_bench.py 32 0 12 > gpurun_out/ref_decode_ctx0.log 2>&1; echo "rc=$?" >> gpurun_out/ref_decode_ctx0.log)
tail -n 14 gpurun_out/fwd_golden.log; tail -n 4 gpurun_out/ref_decode.log gpurun_out/ref_decode_ctx0.log
