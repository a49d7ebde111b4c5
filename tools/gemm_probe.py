"""Blind-debug helper: run the GEMM in every layout on structured inputs and dump diagnostics (gpurun_out/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_search_b200 import ops  # noqa: E402

out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
os.makedirs(out_dir, exist_ok=True)
log = open(os.path.join(out_dir, 'gemm_probe.txt'), 'w')


def P(*a):
    s = ' '.join(str(x) for x in a)
    print(s)
    log.write(s + '\n')
    log.flush()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def diag(name, out, ref):
    out, ref = out.double().cpu(), ref.double().cpu()
    e = rel(out, ref)
    P('%-40s rel %.3e  nan %d' % (name, e, int(torch.isnan(out).sum())))
    if e > 1e-3 or e != e:
        M, N = ref.shape
        bad = (out - ref).abs() > 1e-2 * ref.abs().max()
        P('   bad fraction %.3f; bad rows(first 16) %s; bad cols(first 16) %s' % (
            bad.double().mean().item(), bad.any(1).nonzero().view(-1)[:16].tolist(), bad.any(0).nonzero().view(-1)[:16].tolist()))
        for r0 in range(0, min(M, 128), 32):
            P('   rowblk %3d:' % r0, ' '.join('%.2f' % bad[r0:r0 + 32, c0:c0 + 32].double().mean().item() for c0 in range(0, min(N, 256), 32)))
        P('   out[0,:8]', out[0, :8].tolist())
        P('   ref[0,:8]', ref[0, :8].tolist())
        P('   out[1,:8]', out[1, :8].tolist())
        P('   ref[1,:8]', ref[1, :8].tolist())


def main():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 128, 16), (128, 128, 64), (128, 128, 128), (128, 128, 256), (256, 256, 512)]:
        A = torch.randn(M, K).to(torch.bfloat16)
        W = torch.randn(N, K).to(torch.bfloat16)
        Kp = max(K, 8)
        out = torch.full((M, N), float('nan'), device='cuda')
        try:
            ops.gemm(A.cuda(), W.cuda(), K, K, M, N, K, ops.EPI_STORE, out, N)
            torch.cuda.synchronize()
            diag('TN %dx%dx%d' % (M, N, K), out, A.double() @ W.double().t())
        except Exception as ex:  # noqa: BLE001
            P('TN %dx%dx%d EXC %r' % (M, N, K, ex))
            return
    for (M, N, K) in [(128, 128, 64), (128, 128, 128), (256, 256, 192)]:
        dY = torch.randn(M, K).to(torch.bfloat16)
        W = torch.randn(K, N).to(torch.bfloat16)
        out = torch.full((M, N), float('nan'), device='cuda')
        try:
            ops.gemm(dY.cuda(), W.cuda(), K, N, M, N, K, ops.EPI_STORE, out, N, b_layout=ops.MNMAJOR)
            torch.cuda.synchronize()
            diag('dgrad(B mn-major) %dx%dx%d' % (M, N, K), out, dY.double() @ W.double())
        except Exception as ex:  # noqa: BLE001
            P('dgrad EXC %r' % (ex,))
            return
    for (R, Nw, Kw, sp) in [(64, 128, 128, 1), (128, 128, 128, 1), (512, 256, 128, 2)]:
        dY = torch.randn(R, Nw).to(torch.bfloat16)
        X = torch.randn(R, Kw).to(torch.bfloat16)
        out = torch.zeros(Nw, Kw, device='cuda')
        try:
            ops.gemm(dY.cuda(), X.cuda(), Nw, Kw, Nw, Kw, R, ops.EPI_ATOMIC, out, Kw, a_layout=ops.MNMAJOR,
                     b_layout=ops.MNMAJOR, split_k=sp)
            torch.cuda.synchronize()
            diag('wgrad(A,B mn-major) R%d %dx%d' % (R, Nw, Kw), out, dY.double().t() @ X.double())
        except Exception as ex:  # noqa: BLE001
            P('wgrad EXC %r' % (ex,))
            return
    # quick timing of the big forward shapes
    for (M, N, K) in [(65792, 576, 256), (65792, 768, 256), (65792, 256, 768), (16640, 1536, 512), (4352, 3072, 1024), (8192, 8192, 8192)]:
        A = torch.randn(M, K, device='cuda').to(torch.bfloat16)
        W = torch.randn(N, K, device='cuda').to(torch.bfloat16)
        out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(A, W, K, K, M, N, K, ops.EPI_STORE, out, N)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(A, W, K, K, M, N, K, ops.EPI_STORE, out, N)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        ref = torch.matmul(A, W.t())
        P('time %6dx%5dx%5d  %.3f ms  %.1f TFLOP/s  %.1f GB/s   rel-vs-cublas %.2e' % (
            M, N, K, ms, 2.0 * M * N * K / ms / 1e9, (M * K + N * K + M * N) * 2 / ms / 1e6, rel(out, ref)))


if __name__ == '__main__':
    main()
