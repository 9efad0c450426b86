import torch, time
dev = torch.device('cuda')
for shape in ((1, 1, 480, 640), (1, 1, 1024, 1224)):
    ts = [torch.rand(shape) for _ in range(30)]
    nbytes = ts[0].numel() * 4
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in ts: d = t.to(dev)
    torch.cuda.synchronize()
    base = (time.perf_counter() - t0) / len(ts) * 1e3
    stages = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(3)]
    evs = [None] * 3
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i, t in enumerate(ts):
        k = i % 3
        if evs[k] is not None: evs[k].synchronize()
        v = stages[k].view(torch.float32).view(shape)
        v.copy_(t)
        d = v.to(dev, non_blocking=True)
        evs[k] = torch.cuda.Event(); evs[k].record()
    torch.cuda.synchronize()
    staged = (time.perf_counter() - t0) / len(ts) * 1e3
    print(shape, f'pageable .to(): {base:.3f} ms   staged through pinned: {staged:.3f} ms   threads {torch.get_num_threads()}')
