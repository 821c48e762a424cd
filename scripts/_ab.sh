cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sampler.py tests/test_gpu_feature.py -m gpu -x -q > gpurun_out/t_emit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_emit.log
for i in 1 2; do
QV_EMIT_PER_HOP=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-large-batch > gpurun_out/ab_perhop_$i.json 2>> gpurun_out/ab.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-large-batch > gpurun_out/ab_all_$i.json 2>> gpurun_out/ab.err
done
tail -2 gpurun_out/t_emit.log
