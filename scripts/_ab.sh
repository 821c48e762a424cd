cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sampler.py tests/test_gpu_feature.py -m gpu -x -q > gpurun_out/t_emit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_emit.log
for i in 1 2; do
QV_SCAN_TICKETS=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-large-batch > gpurun_out/ab_tick_$i.json 2>> gpurun_out/ab.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-large-batch > gpurun_out/ab_direct_$i.json 2>> gpurun_out/ab.err
done
tail -2 gpurun_out/t_emit.log
