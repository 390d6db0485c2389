import json, sys; sys.path.insert(0, '.')
import torch
import bench_extra
print(json.dumps(bench_extra.dpd(torch.device('cuda:0'))))
