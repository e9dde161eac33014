N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29601 bench.py --gpus $N --steps 500 --warmup 20 > gpurun_out/r2_bench_A_n$N.json 2> gpurun_out/r2_bench_A_n$N.err; echo "A rc=$?"
timeout 600 $TR --master-port 29602 bench.py --gpus $N --workload C --scaling strong --steps 50 --warmup 5 --e2e-steps 0 > gpurun_out/r2_bench_C_strong_n$N.json 2> gpurun_out/r2_bench_C_strong_n$N.err; echo "C rc=$?"
timeout 600 $TR --master-port 29603 bench.py --gpus $N --workload D10 --scaling point --steps 20 --warmup 3 > gpurun_out/r2_bench_D10_point_n$N.json 2> gpurun_out/r2_bench_D10_point_n$N.err; echo "D10 rc=$?"
for f in A C_strong D10_point; do tail -2 gpurun_out/r2_bench_${f}_n$N.err | cut -c1-300; done
