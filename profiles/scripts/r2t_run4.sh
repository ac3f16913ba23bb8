set -x
TAG=${1:-r2u}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 python -m pytest tests/test_gpu_sharded_index.py -x -q -m gpu 2>&1 | tail -3
SGX_SHARD_TRACE=1 timeout 400 $T tests/multigpu_index_check.py cfg4 1.0 3 2> gpurun_out/${TAG}_trace_peer.err | tail -1 | tee gpurun_out/${TAG}_trace_peer.json
grep "shard 0" gpurun_out/${TAG}_trace_peer.err | tail -3
