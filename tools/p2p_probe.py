import torch, time
print(torch.cuda.device_count(), [torch.cuda.can_device_access_peer(0, j) for j in range(1, torch.cuda.device_count())])
a = torch.empty(1 << 28, dtype=torch.float32, device="cuda:0")
b = torch.empty(1 << 28, dtype=torch.float32, device="cuda:1")
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t = time.time()
for _ in range(10):
    b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
print("peer copy GB/s:", 10 * a.numel() * 4 / (time.time() - t) / 1e9)
