timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests.log 2>&1; echo "tests rc=$?"
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s4_attn_a.json 2> gpurun_out/s4_attn_a.err; echo "rc=$?"
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload multipart --events 1024 > gpurun_out/s4_attn_mp.json 2> gpurun_out/s4_attn_mp.err; echo "rc=$?"
