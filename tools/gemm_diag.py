"""Diagnostic for the tcgen05 GEMM on a real B200: prints the structure of any mismatch on small cases
(which rows / columns / k-blocks are wrong) so descriptor or swizzle mistakes can be read off one run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import sys

import torch

from saspa_aug_b200 import ops


def diag(M, N, K, pattern="rand"):
    torch.manual_seed(0)
    if pattern == "rand":
        a = torch.randn(M, K).to(torch.bfloat16).cuda()
        b = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16).cuda()
    elif pattern == "eye":  # A = row index in col 0.., B = identity-ish: out[m, n] = a[m, n] for n < K
        a = torch.arange(M * K).view(M, K).remainder(251).float().to(torch.bfloat16).cuda()
        b = torch.zeros(N, K)
        for i in range(min(N, K)):
            b[i, i] = 1
        b = b.to(torch.bfloat16).cuda()
    try:
        got = ops.gemm(a, b, out_fp32=True)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"[{M}x{N}x{K} {pattern}] EXCEPTION {e}")
        return False
    ref = a.float() @ b.float().t()
    err = (got - ref).abs()
    tol = 1e-2 * ref.abs().max().clamp_min(1e-3)
    bad = err > tol
    print(f"[{M}x{N}x{K} {pattern}] max err {err.max().item():.4g} (tol {tol.item():.3g}) bad {int(bad.sum())}/{bad.numel()}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten().tolist()
        cols = bad.any(0).nonzero().flatten().tolist()
        print("   bad rows:", rows[:24], "... n=", len(rows))
        print("   bad cols:", cols[:24], "... n=", len(cols))
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
        if pattern == "eye":
            print("   got[1,:16]", got[1, :16].tolist())
            print("   ref[1,:16]", ref[1, :16].tolist())
            print("   got[9,:16]", got[9, :16].tolist())
            print("   ref[9,:16]", ref[9, :16].tolist())
        return False
    return True


if __name__ == "__main__":
    ok = True
    for pat in ("eye", "rand"):
        for M, N, K in [(128, 32, 16), (128, 32, 64), (128, 64, 64), (128, 128, 64), (128, 128, 128), (128, 128, 512), (256, 256, 64),
                        (128, 160, 64), (384, 320, 320), (128, 128, 1024)]:
            ok &= diag(M, N, K, pat)
    print("GEMM DIAG", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
