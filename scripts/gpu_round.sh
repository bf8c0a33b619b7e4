set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/b_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_|decode_|nms_" -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_r2c_list.log 2>&1
python scripts/prof_workload.py 4096 > gpurun_out/pw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"encode_|decode_|nms_" -s 8 -c 8 -o gpurun_out/prof_r2c -f python scripts/prof_workload.py 4096 > gpurun_out/ncu_r2c_full.log 2>&1
tail -5 gpurun_out/ncu_r2c_full.log
