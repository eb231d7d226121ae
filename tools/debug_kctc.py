"""Debug dump for the tcgen05 statistics kernel (not a test)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beer_b200 import ops
ops.require_cuda()
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
for variant in ('0',):
    os.environ['BEER_KCTC_VARIANT'] = variant
    for (M, D, N) in ((8, 40, 64), (100, 40, 640), (8, 20, 64)):
        g = torch.Generator().manual_seed(1)
        X = torch.randn(N, D, generator=g).cuda()
        post = torch.rand(N, M, generator=g).cuda()
        want = torch.cat([post.double().t() @ X.double(), -0.5 * post.double().t() @ (X.double() ** 2),
                          -0.5 * post.double().sum(0)[:, None], 0.5 * post.double().sum(0)[:, None]], 1)
        got = torch.zeros(M, 2 * D + 2, device='cuda', dtype=torch.float64)
        ops.accumulate_stats(X, got, pdf_post=post, tensor_cores=True)
        torch.cuda.synchronize()
        rel = ((got - want).abs() / want.abs().clamp(min=1e-6))
        ok = rel < 1e-4
        print(f'variant {variant} M={M} D={D} N={N}: nonzero {float((got != 0).double().mean()):.3f} '
              f'match {float(ok.double().mean()):.3f} maxrel {float(rel.max()):.3g}')
        print(' rows matching:', ok.all(dim=1).nonzero().flatten().tolist()[:40])
        print(' cols matching:', ok.all(dim=0).nonzero().flatten().tolist()[:90])
        print(' got[0,:6]', got[0, :6].tolist(), '\n want[0,:6]', want[0, :6].tolist())
        print(' got[1,:3]', got[1, :3].tolist(), ' want[1,:3]', want[1, :3].tolist())
        print(' counts got', got[:4, -1].tolist(), ' want', want[:4, -1].tolist())
