run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bench_allreduce.py 2>/dev/null | tail -1; }
run X=1
run NCCL_ALGO=Ring
run NCCL_ALGO=NVLS
run NCCL_ALGO=Tree
run NCCL_MIN_NCHANNELS=32
run NCCL_NVLS_CHUNKSIZE=65536
