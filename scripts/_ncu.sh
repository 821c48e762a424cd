cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sampler.py -m gpu -x -q > gpurun_out/t_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_q.log
ncu --set full --import-source on --clock-control none -k regex:sample_rows_small_kernel -s 11 -c 1 -o gpurun_out/sample_rows_src2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-batch --no-fuse > gpurun_out/ncu_src.log 2>&1
tail -2 gpurun_out/t_q.log
