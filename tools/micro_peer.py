"""Copy-engine peer bandwidth on one box (single process, two devices): cudaMemcpyPeerAsync d0 -> d1 for several sizes
and numbers of concurrent streams. Usage: python tools/micro_peer.py"""
import torch

assert torch.cuda.device_count() >= 2
d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
for mb in (0.25, 1, 4, 16, 47):
    n = int(mb * (1 << 20))
    for streams in (1, 2, 4):
        src = [torch.empty(n, dtype=torch.uint8, device=d0) for _ in range(streams)]
        dst = [torch.empty(n, dtype=torch.uint8, device=d1) for _ in range(streams)]
        sts = [torch.cuda.Stream(device=d0) for _ in range(streams)]
        torch.cuda.synchronize(d0)
        best = 1e9
        for rep in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.device(d0):
                a.record()
                for s in sts:
                    s.wait_event(a)
                for i, s in enumerate(sts):
                    with torch.cuda.stream(s):
                        for _ in range(4):
                            dst[i].copy_(src[i], non_blocking=True)
                for s in sts:
                    e = torch.cuda.Event()
                    e.record(s)
                    torch.cuda.current_stream().wait_event(e)
                b.record()
                torch.cuda.synchronize(d0)
            best = min(best, a.elapsed_time(b))
        tot = 4 * streams * n
        print(f"{mb:6.2f} MB x4 copies x {streams} streams: {best * 1000:8.1f} us  {tot / best / 1e6:7.1f} GB/s "
              f"({best * 1000 / 4:6.1f} us per copy slot)")
