"""tcgen05 probe (dmcf_umma_probe): numerics of the M = 64 / kind::f16 / streamed-A configuration and its time per conv-sized
tile with all SMs streaming the same filter.  python scripts/umma_probe.py"""
import sys
sys.path.insert(0, '.')
import ctypes as C
import numpy as np
import torch
from dmcf_b200 import _lib

lib = _lib.load()
dev = torch.device('cuda')


def pack_a(a):  # a [64, K] fp16 -> per k-step [chunk(2)][rowgroup(8)][8][8]
    k = a.shape[1]
    return a.reshape(8, 8, k // 16, 2, 8).permute(2, 3, 0, 1, 4).contiguous()


def pack_b(b, n):  # b [n, K] fp16 -> [K/8 chunks][(n/8) x 128 + 16 bytes]
    k = b.shape[1]
    ng = n // 8
    out = torch.zeros((k // 8, ng * 64 + 8), dtype=torch.float16, device=b.device)
    out[:, :ng * 64] = b.reshape(ng, 8, k // 8, 8).permute(2, 0, 1, 3).reshape(k // 8, ng * 64)
    return out.contiguous()


def run(ks, n, n_tiles, stages, passes, n_ctas, check=False):
    K = ks * 16
    g = torch.Generator(device='cpu').manual_seed(1)
    a = (torch.randn((64, K), generator=g) * 0.5).to(torch.float16).to(dev)
    bs = [(torch.randn((n, K), generator=g) * (0.5 if q == 0 else 0.001)).to(torch.float16).to(dev) for q in range(passes)]
    a_s = pack_a(a)
    b_s = torch.cat([pack_b(b, n) for b in bs], dim=0).contiguous()
    d = torch.zeros((128, n), dtype=torch.float32, device=dev)
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def call(tiles):
        rc = lib.dmcf_umma_probe(a_s.data_ptr(), b_s.data_ptr(), ks, n, tiles, stages, passes, n_ctas, d.data_ptr(), stats.data_ptr(), st)
        assert rc == 0, _lib.last_error() if hasattr(_lib, 'last_error') else rc
    call(1)
    torch.cuda.synchronize()
    if check:
        ref = a.float() @ sum(b.float() for b in bs).t()  # [64, n]
        got = d.cpu().numpy()
        ref = ref.cpu().numpy()
        # which accumulator lane holds row i?
        lanes = []
        for i in range(64):
            err = np.abs(got - ref[i][None, :]).max(axis=1)
            lanes.append(int(err.argmin()))
        worst = max(np.abs(got[lanes[i]] - ref[i]).max() for i in range(64))
        print('row -> lane', lanes)
        print('max abs err %.3e (max |ref| %.3e)' % (worst, np.abs(ref).max()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    call(n_tiles)
    torch.cuda.synchronize()
    e0.record(); call(n_tiles); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    us_tile = ms * 1e3 / n_tiles
    print(f'ks {ks} n {n} passes {passes} stages {stages} ctas {n_ctas}: {us_tile:.2f} us per tile, '
          f'{ks * 2048 * n_ctas / us_tile / 1e6:.2f} TB/s filter stream, {us_tile * 1965 / (ks * passes):.1f} clk per MMA; '
          f'clk per k-step: producer wait/copy, mma wait/issue/commit', [round(v / (n_tiles * ks), 1) for v in stats[:5].tolist()],
          flush=True)


if __name__ == '__main__':
    run(130, 16, 50, 24, 2, 1, check=True)
    for stages in (4, 8, 16, 24, 32):
        run(130, 16, 200, stages, 2, 148)
    for stages in (4, 6):
        run(130, 24, 200, stages, 2, 148)
    run(130, 16, 200, 24, 2, 1)
    run(130, 16, 200, 24, 1, 148)
    run(130, 8, 200, 24, 2, 148)
