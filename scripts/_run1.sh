cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_fuse.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_fuse.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fuse.json 2> gpurun_out/bench_fuse.err
python bench.py --steps 20 --warmup 5 --no-fuse --no-cpu-baseline --no-large-batch > gpurun_out/bench_nofuse.json 2>> gpurun_out/bench_fuse.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-large-batch > gpurun_out/bench_fuse2.json 2>> gpurun_out/bench_fuse.err
tail -3 gpurun_out/t_fuse.log; cat gpurun_out/smoke.log | tail -2
