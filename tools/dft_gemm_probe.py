"""Analysis only (not product code): how fast and how accurate is a 64-point DFT done as a GEMM on the tensor cores
(library GEMM through torch), against this repo's register FFT?  Answers the north-star clause "tensor cores only if a
64x64 DFT-as-GEMM variant measurably beats the FFT".  X[N,64] complex -> real GEMM [N,128] x [128,128]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
dev = o.device
N = 4_194_304
x = torch.randn((N, 64, 2), device=dev, dtype=torch.float32)
k = np.arange(64)
W = np.exp(-2j * np.pi * np.outer(k, k) / 64)                      # X = x @ W (symmetric)
B = np.zeros((128, 128), np.float64)                                 # [re, im] interleaved in, [re, im] interleaved out
B[0::2, 0::2] = W.real; B[1::2, 0::2] = -W.imag; B[0::2, 1::2] = W.imag; B[1::2, 1::2] = W.real
Bt = torch.tensor(B, device=dev, dtype=torch.float32)
A = x.reshape(N, 128)
ref = torch.fft.fft(torch.view_as_complex(x[:65536].contiguous().double()), dim=1)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def err(y):
    yc = torch.view_as_complex(y[:65536].reshape(-1, 64, 2).contiguous().double())
    return float((yc - ref).abs().max() / ref.abs().max())
out = {}
for name, tf32 in (("fp32 (no tensor cores)", False), ("tf32 tensor cores", True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    ms = t(lambda: torch.matmul(A, Bt)); y = torch.matmul(A, Bt)
    print("%-26s %.3f ms  %.2e transforms/s  max rel err %.1e" % (name, ms, N / ms * 1e3, err(y)))
# 3xTF32 split (hi/lo of both operands, three GEMMs): fp32-class accuracy
torch.backends.cuda.matmul.allow_tf32 = True
def split(v):
    hi = (v.view(torch.int32) & -8192).view(torch.float32)          # keep 10 mantissa bits
    return hi, v - hi
Bh, Bl = split(Bt)
def three():
    Ah, Al = split(A)
    return torch.matmul(Ah, Bh) + torch.matmul(Al, Bh) + torch.matmul(Ah, Bl)
ms = t(three); print("%-26s %.3f ms  %.2e transforms/s  max rel err %.1e" % ("3xTF32 (split, 3 GEMMs)", ms, N / ms * 1e3, err(three())))
# bf16 tensor cores (for completeness)
Ab, Bb = A.bfloat16(), Bt.bfloat16()
ms = t(lambda: torch.matmul(Ab, Bb)); print("%-26s %.3f ms  %.2e transforms/s  max rel err %.1e" % ("bf16 tensor cores", ms, N / ms * 1e3, err(torch.matmul(Ab, Bb).float())))
# this repo's FFT kernels on the same data (HBM in, HBM out)
for mode, name in ((pkg.MODE_FAST, "register FFT fast (this repo)"), (pkg.MODE_EXACT, "register FFT exact (this repo)")):
    ms = t(lambda: o.fft64(x, mode))
    y = o.fft64(x[:65536].contiguous(), mode)            # centred output: undo the shift for the error check
    yc = torch.view_as_complex(torch.roll(y, -32, dims=1).contiguous().double())
    print("%-30s %.3f ms  %.2e transforms/s  max rel err %.1e" % (name, ms, N / ms * 1e3, float((yc - ref).abs().max() / ref.abs().max())))
print("HBM floor for %d transforms in+out: %.3f ms at 6554 GB/s" % (N, N * 1024 / 6554e9 * 1e3))
