cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rpyNearList|rpyNearTraversal|ibmGatherSorted|ibmSpreadRows|verletFill|pseNear" -s 12 -c 8 -o gpurun_out/r03j_pse -f python scripts/extra_step.py pse > gpurun_out/r03j_pse.log 2>&1; tail -1 gpurun_out/r03j_pse.log | cut -c1-150
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dpdTileTraversal|brickAdvancePush|brickUnpack|brickKick" -s 8 -c 6 -o gpurun_out/r03j_dpd -f python -c "
import json, sys; sys.path.insert(0, '.')
import torch, bench_extra
print(json.dumps(bench_extra.dpd(torch.device('cuda:0'), steps=3, warmup=2, equil=10)))" > gpurun_out/r03j_dpd.log 2>&1; tail -1 gpurun_out/r03j_dpd.log | cut -c1-150
