set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m | head -12; nproc; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/shard_check.py > gpurun_out/r2f_shard8.log 2>&1; grep -E "world|Error|error" gpurun_out/r2f_shard8.log | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err; cut -c1-300 gpurun_out/r2f_bench_n8.json; tail -4 gpurun_out/r2f_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 5 --warmup 3 --no-piano > gpurun_out/r2f_bench_n4.json 2> gpurun_out/r2f_bench_n4.err; cut -c1-300 gpurun_out/r2f_bench_n4.json; tail -4 gpurun_out/r2f_bench_n4.err
SFB_SHARD_BLOCK=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 --no-piano --no-strong > gpurun_out/r2f_bench_n8_b1.json 2> gpurun_out/r2f_bench_n8_b1.err; cut -c1-200 gpurun_out/r2f_bench_n8_b1.json
