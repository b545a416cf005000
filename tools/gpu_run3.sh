mkdir -p gpurun_out
D=gptq-for-llama_b200/dev
for args in "1 0" "1 2047" "4 2047"; do
(timeout 200 python tools/dev_determinism.py $args >> gpurun_out/det.log 2>&1; echo "rc=$?" >> gpurun_out/det.log)
(GPTQ_B200_LIB=$D/libgptq_b200_fence.so timeout 200 python tools/dev_determinism.py $args >> gpurun_out/det.log 2>&1; echo "rc=$?" >> gpurun_out/det.log)
(GPTQ_B200_LIB=$D/libgptq_b200_exact.so timeout 200 python tools/dev_determinism.py $args >> gpurun_out/det.log 2>&1; echo "rc=$?" >> gpurun_out/det.log)
done
(GPTQ_B200_LIB=$D/libgptq_b200_trace.so timeout 200 python tools/trace_mega.py 7b > gpurun_out/trace_7b.log 2>&1; echo "rc=$?" >> gpurun_out/trace_7b.log)
cat gpurun_out/det.log; tail -12 gpurun_out/trace_7b.log
