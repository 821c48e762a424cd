cd /root/repo
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:sample_rows_small_kernel -s 12 -c 1 -o gpurun_out/sample_rows_hop0 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-batch --no-fuse > gpurun_out/ncu_src.log 2>&1
