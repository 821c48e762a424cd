cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches_fused_warm.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-large-batch > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/t_all.log
