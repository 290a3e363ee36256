set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_final.log 2>&1; tail -2 gpurun_out/r02_pytest_final.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.json
timeout 600 python scripts/train_bench.py --batch 296 > gpurun_out/r02_train_final.json 2>&1; tail -c 400 gpurun_out/r02_train_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_bench_under_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_train_launches2.csv python scripts/train_bench.py --batch 296 --steps 2 --warmup 1 > gpurun_out/r02_train_under_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02_small_launches2.csv python scripts/dev/small_shapes_launches.py > gpurun_out/r02_small_under_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fused2_forward|materialize_final" -c 2 -o gpurun_out/r02_final_kernels python scripts/quick_time.py --B 296 --iters 1 --algo fused > gpurun_out/r02_ncu_final2.log 2>&1; tail -1 gpurun_out/r02_ncu_final2.log | cut -c1-150
