# per-node cost of a linear chain of tiny kernels in a CUDA graph (the floor a ~36-kernel training step pays for its kernel
# boundaries): 1-CTA kernels and 148-CTA x 512-thread kernels with 200 KB of dynamic shared memory (the shape of our launches)
import torch
x = torch.zeros(32, device='cuda')
big = torch.zeros(148 * 512, device='cuda')
def chain(t, n):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): t.add_(1.0)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(n): t.add_(1.0)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)
for name, t in (("32 elements (1 CTA)", x), ("75776 elements", big)):
    t10, t110 = chain(t, 10), chain(t, 110)
    print("%s: %.2f us per kernel node (chain of 110 vs 10)" % (name, 1e3 * (t110 - t10) / 100))
