cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/t_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_multi.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_final.json 2> gpurun_out/n2.err
$T bench.py --gpus 2 --steps 5 --warmup 2 --impl reference > gpurun_out/n2_ref.json 2>> gpurun_out/n2.err
tail -3 gpurun_out/t_multi.log
