#!/bin/bash
nvidia-smi topo -m 2>&1 | head -20
nvidia-smi nvlink -s 2>&1 | head -12
python - <<'PY'
import torch, time
a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0"); b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:1")
print("can access peer", torch.cuda.can_device_access_peer(0, 1))
for _ in range(2):
    torch.cuda.synchronize(0); torch.cuda.synchronize(1); t = time.time(); b.copy_(a); torch.cuda.synchronize(0); torch.cuda.synchronize(1); dt = time.time() - t
print("peer copy 256 MiB: %.1f GB/s" % (0.268 / dt))
s = torch.empty(64, dtype=torch.uint8, device="cuda:0"); d = torch.empty(64, dtype=torch.uint8, device="cuda:1")
torch.cuda.synchronize(); t = time.time()
for _ in range(1000): d.copy_(s)
torch.cuda.synchronize(0); torch.cuda.synchronize(1); print("tiny peer copy: %.1f us each" % ((time.time() - t) * 1e3))
PY
