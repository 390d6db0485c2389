cd /root/repo
timeout 600 python -m pytest tests/test_brick_gpu.py -x -q -m gpu 2>&1 | tail -5
for prof in 1 0; do
UB200_BRICK_PROFILE=$prof timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2981$prof scripts/brick_lj.py --cells 63 --steps 200 2>/dev/null | grep '^{' | tee gpurun_out/r02g_brick2_prof$prof.json
done
