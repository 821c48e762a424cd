cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sampler.py tests/test_gpu_feature.py -m gpu -x -q > gpurun_out/t_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_h.log
for i in 1 2; do
QV_HEAVY_FIRST=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_noheavy_$i.json 2>> gpurun_out/ab.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_heavy_$i.json 2>> gpurun_out/ab.err
done
tail -2 gpurun_out/t_h.log
