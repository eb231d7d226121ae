#!/usr/bin/env python
"""clock64 trace of one CTA of the mix16 kernels at the cfg3 shape (measure the critical path).

    python tools/trace_mix16.py > gpurun_out/trace_mix16.txt
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from beer_b200 import _lib, ops, synthetic
    dev = 'cuda'
    M, D, C, N = 8000, 40, 8, 400000
    Kp = M // C
    prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
    X = torch.randn(N, D, device=dev)
    logw = torch.full((M,), -2.0794, device=dev)
    W, bias, ref = ops.emission_prepare(*post, logw=logw)
    mx = ops.Mix16(M, D, C, dev)
    images = mx.build_images(X)
    mx.pack(W, bias, images['alpha'])
    llh2 = mx.emission(images)
    lp = torch.log2(torch.softmax(torch.randn(N, Kp, device=dev) * 4, dim=1)).contiguous()
    acc = torch.zeros(M, 2 * D + 2, device=dev, dtype=torch.float64)
    lrel = (lp - llh2).contiguous()      # the one-array form the forward-backward writes (BEER_FB_LPOST_RELATIVE)
    mx.accumulate(images, lrel, None, acc, relative=True)
    torch.cuda.synchronize()
    lib = _lib.load()
    for name, fn in (('emission (per chunk)', lambda: mx.emission(images, out=llh2)),
                     ('statistics (per tile)', lambda: mx.accumulate(images, lrel, None, acc, relative=True))):
        buf = torch.zeros(256 * 8, device=dev, dtype=torch.int64)
        lib.beer_mix16_set_trace(buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        lib.beer_mix16_set_trace(None)
        t = buf.cpu().numpy().reshape(256, 8)
        t0 = t[t > 0].min()
        rel = np.where(t > 0, t - t0, -1)
        print(f'== {name}: time stamps of CTA 0 in cycles since its first event; rows = items 100..131')
        print('item ' + ' '.join(f'{i:>9d}' for i in range(8)))
        for i in range(100, 132):
            print(f'{i:4d} ' + ' '.join(f'{v:9d}' for v in rel[i]))
        d = np.diff(rel[100:200], axis=0)
        print('mean period per item (cycles), per slot:', np.round(d.mean(axis=0), 0))


if __name__ == '__main__':
    main()
