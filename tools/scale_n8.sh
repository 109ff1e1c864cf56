mkdir -p gpurun_out/r3k; nvidia-smi -L | wc -l
run() { tag=$1; N=$2; shift 2; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus $N "$@" > gpurun_out/r3k/$tag.log 2>&1; echo $tag rc=$?; grep "^{" gpurun_out/r3k/$tag.log | cut -c1-220; }
run step_c2_n8 8 --steps 20 --warmup 5 --no-cpu-baseline
run job_c2_n8 8 --mode job --n-mc 100
run job_c5_n8 8 --workload C5 --mode job --n-mc 200
run step_c2_n2 2 --steps 20 --warmup 5 --no-cpu-baseline
run job_c2_n4 4 --mode job --n-mc 100
