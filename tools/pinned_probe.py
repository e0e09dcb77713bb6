"""Does freeing, from the main thread (current device 0), a pinned host tensor that received a non-blocking copy from
cuda:1 in another thread abort the process?  (diagnosis for LocalGPUWorker's D2H lane on multi-GPU boxes)"""
import sys
import threading

import torch

mode = sys.argv[1] if len(sys.argv) > 1 else "pinned"
box = {}


def work():
    x = torch.randint(0, 255, (4, 64, 64, 3), device="cuda:1", dtype=torch.uint8)
    if mode == "pinned":
        h = torch.empty(x.shape, dtype=torch.uint8, pin_memory=True)
        h.copy_(x, non_blocking=True)
        torch.cuda.current_stream(x.device).synchronize()
    else:
        h = x.cpu()
    box["h"] = h


t = threading.Thread(target=work)
t.start()
t.join()
s = int(box["h"].sum())
box["h"] = None
print(mode, "ok", s, flush=True)
