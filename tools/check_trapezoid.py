"""GPU check of bspb200_dev_potrf on a trapezoid (diagonal block + rows below) against torch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import baspacho_b200 as bsp
api = bsp.api()
torch.manual_seed(0)
for n, rb in ((390, 0), (1000, 700), (1700, 3000), (1488, 948), (1440, 1032), (2000, 1)):
    M = torch.randn(n, n, dtype=torch.float64, device="cuda")
    A11 = M @ M.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    A21 = torch.randn(rb, n, dtype=torch.float64, device="cuda")
    A = torch.cat([A11, A21]).contiguous()
    stream = torch.cuda.Stream() if len(sys.argv) > 1 else torch.cuda.current_stream()
    st = stream.cuda_stream
    worst = 0
    for rep in range(5):
        W = A.clone()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            junk = torch.randn(4096, 4096, device="cuda") @ torch.randn(4096, 4096, device="cuda")  # busy stream
            api.check(api.dev_potrf(0, n, rb, W.data_ptr(), n, st))
        torch.cuda.synchronize()
        L = torch.linalg.cholesky(A11)
        X = torch.linalg.solve_triangular(L, A21.T, upper=False).T if rb else A21
        e1 = (torch.tril(W[:n]) - L).abs().max().item() / L.abs().max().item()
        e2 = (W[n:] - X).abs().max().item() / max(1e-300, X.abs().max().item()) if rb else 0.0
        worst = max(worst, e1, e2)
    print(n, rb, "max rel err", worst)
