# C5 (1024^3) Monte-Carlo job: one GPU with the per-rank share of the 8-GPU job (25 realisations), then N_mc = 200 on 8 GPUs
mkdir -p gpurun_out/r4b
python bench.py --workload C5 --mode job --n-mc 25 --no-cpu-baseline > gpurun_out/r4b/job_c5_n1_25.log 2>&1; echo n1 rc=$?; grep "^{" gpurun_out/r4b/job_c5_n1_25.log | cut -c1-200
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload C5 --mode job --n-mc 200 > gpurun_out/r4b/job_c5_n8.log 2>&1; echo n8 rc=$?; grep "^{" gpurun_out/r4b/job_c5_n8.log | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --mode job --n-mc 100 > gpurun_out/r4b/job_c2_n8.log 2>&1; echo c2n8 rc=$?; grep "^{" gpurun_out/r4b/job_c2_n8.log | cut -c1-200
