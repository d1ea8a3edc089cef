"""cuFFT timings that decide how the Fresnel transform is built (no product code: torch.fft = cuFFT).
2-D C2C at the reference's size N + 30 (Bluestein inside cuFFT) against 1-D batched transforms at sizes with
small prime factors, which is what a hand-written chirp-z around cuFFT would run."""
import sys
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = n + 30
dev = torch.device("cuda:0")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


x = torch.randn(p, p, dtype=torch.complex64, device=dev)
print("fft2 %d^2: %.3f ms" % (p, timed(lambda: torch.fft.fft2(x))))
print("fft rows %d x %d: %.3f ms" % (p, p, timed(lambda: torch.fft.fft(x, dim=1))))
print("fft cols %d x %d: %.3f ms" % (p, p, timed(lambda: torch.fft.fft(x, dim=0))))
y = torch.randn(n, n, dtype=torch.complex64, device=dev)
print("fft2 %d^2: %.3f ms" % (n, timed(lambda: torch.fft.fft2(y))))
good = []
for m in range(2 * p - 1, 4 * n + 1):
    r = m
    for q in (2, 3, 5, 7):
        while r % q == 0:
            r //= q
    if r == 1:
        good.append(m)
for m in good[:12] + [4 * n]:
    z = torch.randn(p, m, dtype=torch.complex64, device=dev)
    t = timed(lambda: torch.fft.fft(z, dim=1))
    gb = 2 * z.numel() * 8 / 1e9
    print("fft rows %d x %d: %.3f ms  (%.0f GB/s for one read + one write)" % (p, m, t, gb / t * 1e3))
    del z
# elementwise pass over the padded array, for scale
z = torch.randn(p, good[0], dtype=torch.complex64, device=dev)
w = torch.randn(good[0], dtype=torch.complex64, device=dev)
print("pointwise multiply %d x %d: %.3f ms" % (p, good[0], timed(lambda: z.mul_(w))))
