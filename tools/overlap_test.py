"""Does a pinned H2D upload on a side stream overlap kernels (eager / CUDA graph) on this box?"""
import torch

dev = torch.device("cuda:0")
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
host = torch.empty(59_000_000 // 4, dtype=torch.float32).pin_memory()
dbuf = torch.empty_like(host, device=dev)
out_h = torch.empty(2_000_000, dtype=torch.bfloat16).pin_memory()
small = torch.zeros(2_000_000, dtype=torch.bfloat16, device=dev)
cs = torch.cuda.Stream()


def compute():
    for _ in range(8):
        torch.matmul(a, b)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def with_upload(run):
    def f():
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        cs.wait_event(ev)
        with torch.cuda.stream(cs):
            dbuf.copy_(host, non_blocking=True)
        run()
        main.wait_stream(cs)
    return f


print("compute eager            %.3f ms" % timed(compute))
print("upload alone             %.3f ms" % timed(lambda: dbuf.copy_(host, non_blocking=True)))
print("eager + upload           %.3f ms" % timed(with_upload(compute)))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    compute()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    compute()
print("graph                    %.3f ms" % timed(g.replay))
print("graph + upload           %.3f ms" % timed(with_upload(g.replay)))
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    compute()
    out_h.copy_(small, non_blocking=True)
print("graph(+D2H node)         %.3f ms" % timed(g2.replay))
print("graph(+D2H node) + upload %.3f ms" % timed(with_upload(g2.replay)))
