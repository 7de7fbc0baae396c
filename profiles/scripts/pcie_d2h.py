import torch, time
dev = torch.device('cuda', 0)
n = 502560000 // 8
src = torch.randn(n, dtype=torch.float64, device=dev)
dst = torch.empty(n, dtype=torch.float64).pin_memory()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
one = t(lambda: dst.copy_(src, non_blocking=True))
print('one copy: %.2f ms  %.1f GB/s' % (one, n * 8 / one / 1e6))
for k in (2, 4, 8):
    streams = [torch.cuda.Stream(dev) for _ in range(k)]
    chunks_s = src.chunk(k); chunks_d = dst.chunk(k)
    def multi():
        cur = torch.cuda.current_stream()
        for st, a, b in zip(streams, chunks_s, chunks_d):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                b.copy_(a, non_blocking=True)
        for st in streams: cur.wait_stream(st)
    m = t(multi)
    print('%d concurrent chunks: %.2f ms  %.1f GB/s' % (k, m, n * 8 / m / 1e6))
# H2D concurrently with D2H
h = torch.randn(25128000 // 8, dtype=torch.float64).pin_memory()
