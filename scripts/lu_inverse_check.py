import torch, sys
sys.path.insert(0, "/root/repo")
from nf4ad_b200 import ops
torch.manual_seed(0)
for D in (16, 48, 64, 100, 784, 1000):
    L = torch.randn(D, D, device="cuda") * 0.05 + torch.eye(D, device="cuda")
    U = torch.randn(D, D, device="cuda") * 0.05 + torch.eye(D, device="cuda") * 1.5
    A = ops.lu_inverse(L, U)
    Lt = torch.tril(L, -1) + torch.eye(D, device="cuda"); Ut = torch.triu(U)
    ref = torch.linalg.inv((Lt.double() @ Ut.double()))
    old = ops.lu_solve(torch.eye(D, device="cuda"), L, U, None, transpose=True)
    print(D, "new err", float((A.double() - ref).abs().max() / ref.abs().max()), "old err", float((old.double() - ref).abs().max() / ref.abs().max()))
import time
D = 784
L = torch.randn(D, D, device="cuda") * 0.05; U = torch.randn(D, D, device="cuda") * 0.05 + torch.eye(D, device="cuda")
for fn, name in ((lambda: ops.lu_inverse(L, U), "lu_inverse"), (lambda: ops.lu_solve(torch.eye(D, device="cuda"), L, U, None, transpose=True), "lu_solve(eye)")):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 20 * 1e3, "us")
