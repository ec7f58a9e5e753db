#!/usr/bin/env python
"""Where a training step of the IA-SSD SA stack spends its time: torch.profiler over 3 steps (after 3 warm-ups), top CUDA kernels
and the host/device balance.   SPSK_TRAIN_FUSED=1 python scripts/prof_train_backbone.py"""
import sys, time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from spsnet_b200 import backbone as bb, scenes  # noqa: E402
from spsnet_b200.configs import kitti_iassd_cfg  # noqa: E402

B = 8
torch.manual_seed(0)
net = bb.IASSD_Backbone(kitti_iassd_cfg(), num_class=3, input_channels=4).cuda().train()
pts = torch.from_numpy(np.ascontiguousarray(scenes.to_points(scenes.make_batch(7, B, 16384)))).cuda()


def step():
    out = net({"batch_size": B, "points": pts})
    loss = out["centers_features"].square().mean() + out["ctr_offsets"][:, 1:].square().mean()
    net.zero_grad(set_to_none=True)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"3 steps: host issue {t_host * 1e3 / 3:.2f} ms/step, wall {t_all * 1e3 / 3:.2f} ms/step")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
