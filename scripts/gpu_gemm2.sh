mkdir -p gpurun_out
GEMM_MAJORS=1 timeout 600 python scripts/bench_gemm.py > gpurun_out/bench_gemm_majors.log 2>&1; cat gpurun_out/bench_gemm_majors.log
