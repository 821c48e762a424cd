cd /root/repo
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 4 --steps 20 --warmup 5 --no-large-batch > gpurun_out/n4_final.json 2> gpurun_out/n4.err
tail -c 300 gpurun_out/n4.err
