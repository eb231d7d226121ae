#!/usr/bin/env python
"""Roofline denominators measured on the GPU this runs on:

    python tools/microbench.py > gpurun_out/microbench.json

* tcgen05 dispatch-limited peak of kind::tf32 and kind::f16 (beer_probe_mma: M = 128, N = 256 MMAs back to back,
  one CTA per SM) -> the 3-pass split of the statistics / emission kernels runs at a third of it;
* DRAM write-only ceilings (float4 stores, 32 KB bulk copies shared -> global) and the read-only rate over 4 GiB,
  next to the driver's copy figure in MEASURED_PEAKS.json.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from beer_b200 import ops
    ops.require_cuda()
    out = {'gpu': torch.cuda.get_device_name(0)}
    for kind in ('tf32', 'f16'):
        out[f'mma_{kind}_tflops'] = [round(ops.probe_mma_tflops(kind, n_mma=n), 1) for n in (2000, 20000, 200000)]
    for mode in ('fill_st', 'fill_bulk', 'read'):
        out[f'dram_{mode}_gbs'] = round(ops.probe_dram_gbs(mode), 1)
    if '--mma' in sys.argv:
        out['mma_cycles'] = {}
        for N in (64, 80, 160, 256):
            for n_acc in (1, 2):
                if n_acc * N > 480:
                    continue
                for a_tmem in (0, 1):
                    for elect in (0, 1):
                        out['mma_cycles'][f'N{N}_acc{n_acc}_{"ts" if a_tmem else "ss"}_{"elect" if elect else "lane0"}'] = round(
                            ops.probe_mma_shape_cycles(N, n_acc, 8, a_tmem, elect), 1)
    if '--tma' in sys.argv:
        out['tma_gbs'] = {}
        for mode in ('bulk', 'tensor'):
            for chunk, stages, issuers in ((4096, 1, 1), (4096, 8, 1), (4096, 8, 2), (4096, 8, 4), (4096, 16, 8), (16384, 1, 1),
                                           (16384, 8, 1), (16384, 8, 2), (16384, 8, 4), (32768, 4, 1), (32768, 4, 2),
                                           (32768, 4, 4), (65536, 3, 1), (65536, 3, 3)):
                copies = max(240, (64 << 20) // chunk)
                out['tma_gbs'][f'{mode}_{chunk}x{stages}i{issuers}'] = round(
                    ops.probe_tma_gbs(mode, chunk, stages, copies=copies, issuers=issuers), 0)
        out['tma_gbs_shared_walk'] = {
            f'bulk_{chunk}x{stages}i{issuers}': round(ops.probe_tma_gbs('bulk', chunk, stages, src_mib=mib, copies=copies,
                                                                        issuers=issuers, shared_walk=True), 0)
            for chunk, stages, issuers, mib, copies in ((16384, 8, 4, 64, 4096), (32768, 4, 2, 64, 2048), (49152, 4, 2, 3, 1400),
                                                        (49152, 2, 1, 3, 1400), (20480, 8, 4, 64, 3000))}
        out['tma_gbs_dram'] = {f'{mode}_32768x4': round(ops.probe_tma_gbs(mode, 32768, 4, src_mib=2048, copies=400), 0)
                               for mode in ('bulk', 'tensor')}
    a = torch.empty(1 << 30, device='cuda', dtype=torch.float32)
    b = torch.empty_like(a)
    t = ops._timed(lambda: b.copy_(a), 5)
    out['dram_copy_gbs'] = round(2 * a.numel() * 4 / t / 1e9, 1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
