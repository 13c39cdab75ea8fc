"""A/B timing of C4-shaped inference (8-stack J=17, 256x192, batch 64 + mirrors)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from src.models import hg
torch.manual_seed(0)
m = hg(num_stacks=8, num_blocks=1, num_classes=17, mobile=False, skip_mode="sum").cuda().eval()
x = torch.randn(128, 3, 256, 192, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = m(x)[-1]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = m(x)[-1]
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"C4 inference 128 rows 256x192: {ms:.3f} ms/forward, {128 / ms * 1e3:.0f} rows/s, ragged_halo={'HG_NO_RAGGED_HALO' not in os.environ}, checksum {float(y.float().abs().mean()):.6f}")
