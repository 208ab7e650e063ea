"""Host -> device copy bandwidth: torch pinned memory against write-combined pinned memory (development aid)."""
import ctypes, time
import torch
n = 603982080
dev = torch.device('cuda:0')
dst = torch.empty(n, dtype=torch.uint8, device=dev)
def bw(src, label):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print('%-28s %.2f ms  %.1f GB/s  pinned=%s' % (label, ms, n / ms / 1e6, src.is_pinned()))
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
pin.random_(0, 255)
bw(pin, 'torch pinned')
try:
    from cuda import cudart
    err, ptr = cudart.cudaHostAlloc(n, cudart.cudaHostAllocWriteCombined)
    assert int(err) == 0, err
    buf = (ctypes.c_uint8 * n).from_address(int(ptr))
    wc = torch.frombuffer(buf, dtype=torch.uint8)
    t0 = time.time(); wc.copy_(pin); print('fill of the write-combined buffer %.1f ms' % ((time.time() - t0) * 1e3))
    bw(wc, 'write-combined pinned')
    err, ptr2 = cudart.cudaHostAlloc(n, cudart.cudaHostAllocDefault)
    buf2 = (ctypes.c_uint8 * n).from_address(int(ptr2))
    bw(torch.frombuffer(buf2, dtype=torch.uint8), 'cudaHostAlloc default')
except Exception as exc:
    print('cuda-python path failed:', repr(exc))
# two halves on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h = n // 2
def two():
    with torch.cuda.stream(s1): dst[:h].copy_(pin[:h], non_blocking=True)
    with torch.cuda.stream(s2): dst[h:].copy_(pin[h:], non_blocking=True)
two(); torch.cuda.synchronize()
t0 = time.time()
for _ in range(10): two()
torch.cuda.synchronize(); ms = (time.time() - t0) * 100
print('two halves on two streams    %.2f ms  %.1f GB/s' % (ms, n / ms / 1e6))
