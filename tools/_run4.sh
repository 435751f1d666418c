set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m | head -8
python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2d_pytest.log; tail -6 gpurun_out/r2d_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_check.py > gpurun_out/r2d_shard.log 2>&1; grep -E "world|Error|error" gpurun_out/r2d_shard.log | tail -8
SFB_NO_PEER_FRAMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/shard_check.py > gpurun_out/r2d_shard_nopeer.log 2>&1; grep -E "world|Error|error" gpurun_out/r2d_shard_nopeer.log | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; cut -c1-400 gpurun_out/r2d_bench_n2.json; tail -5 gpurun_out/r2d_bench_n2.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; cut -c1-300 gpurun_out/r2d_bench_n1.json; tail -5 gpurun_out/r2d_bench_n1.err
