"""Which kernels cuBLAS picks for the dense equivalents of the bench GEMMs (run under ncu; tuning aid, not product)."""
import torch
dev = "cuda"
shapes = [(8192, 3072, 16384, "nt"), (8192, 8192, 3072, "nt"), (8192, 3072, 8192, "nn"), (16384, 8192, 3072, "tn")]
for m, k, n, lay in shapes:
    a = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
    b = torch.randn(n, k, device=dev, dtype=torch.bfloat16) if lay[1] == "t" else torch.randn(k, n, device=dev, dtype=torch.bfloat16)
    if lay[0] == "t":
        a = torch.randn(k, m, device=dev, dtype=torch.bfloat16).t()
    for _ in range(3):
        c = a @ (b.t() if lay[1] == "t" else b)
torch.cuda.synchronize()
