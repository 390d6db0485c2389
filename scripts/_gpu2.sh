cd /root/repo
timeout 900 python -m pytest tests/test_brick_gpu.py tests/test_rdf_gpu.py "tests/test_ref_parity_gpu.py::test_reference_parity[1000000-fcc]" tests/test_pse_gpu.py -q -m gpu 2>&1 | tail -6
for g in 1 0; do
UB200_BRICK_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2983$g scripts/brick_lj.py --cells 63 --steps 200 --check 2>/dev/null | grep '^{' | tee gpurun_out/r02l_brick2_graph$g.json
done
