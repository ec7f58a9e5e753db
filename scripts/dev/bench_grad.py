import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from spsnet_b200 import pointnet2_utils as pu
def dev_t(fn, it=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
rng = np.random.default_rng(0)
for B, C, N, M, S in [(8, 64, 4096, 1024, 32), (8, 256, 512, 256, 32), (8, 1, 16384, 4096, 32)]:
    idx = torch.from_numpy(rng.integers(0, N, (B, M, S)).astype(np.int32)).cuda()
    g = torch.randn(B, C, M, S, device="cuda")
    f = torch.randn(B, C, N, device="cuda", requires_grad=True)
    out = pu.grouping_operation(f, idx)
    def run():
        f.grad = None
        out.backward(g, retain_graph=True)
    os.environ.pop("SPSK_DETERMINISTIC_GRAD", None); a = dev_t(run)
    os.environ["SPSK_DETERMINISTIC_GRAD"] = "1"; d = dev_t(run)
    print(f"group_points_grad B={B} C={C} N={N} npoint={M} nsample={S}: atomics {a:.1f} us, sorted segment reduce {d:.1f} us")
