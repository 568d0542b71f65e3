import torch, time
x = torch.zeros(1024, device='cuda')
torch.cuda.synchronize()
for n in (1000, 10000):
    t = time.perf_counter()
    for _ in range(n):
        x.add_(1.0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{n} tiny launches: enqueue {(t1-t)/n*1e6:.2f} us/launch, total {(t2-t)/n*1e6:.2f} us/launch")
