"""Diagnostic: can an H2D copy on one stream overlap a long kernel on another stream on this box (plain torch)?"""
import time, torch
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def work():
    with torch.cuda.stream(sa):
        for _ in range(12): y = x @ x
def copy():
    with torch.cuda.stream(sb):
        d.copy_(h, non_blocking=True)
for _ in range(2): work(); copy(); torch.cuda.synchronize()
t0 = time.perf_counter(); work(); torch.cuda.synchronize(); t1 = time.perf_counter()
copy(); torch.cuda.synchronize(); t2 = time.perf_counter()
work(); copy(); torch.cuda.synchronize(); t3 = time.perf_counter()
print(f"kernel {1e3*(t1-t0):.2f} ms, copy {1e3*(t2-t1):.2f} ms ({0.256/(t2-t1):.1f} GB/s), both {1e3*(t3-t2):.2f} ms")
