cd /root/repo
UB200_BRICK_PROFILE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29821 scripts/brick_lj.py --cells 63 --steps 200 2>/dev/null | grep '^{' | tee gpurun_out/r02h_brick8_prof.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29822 scripts/brick_lj.py --cells 63 --steps 200 --check 2>/dev/null | grep '^{' | tee gpurun_out/r02h_brick8.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29823 scripts/brick_lj.py --cells 63 --steps 200 2>/dev/null | grep '^{' | tee gpurun_out/r02h_brick4.json
