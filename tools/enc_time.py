import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, labrador_ldpc_b200 as L
for code in range(9):
    c = L.LDPCCode(code); batch = max(4096, min(1 << 18, (1 << 28) // c.n()))
    data = torch.randint(0, 256, (batch, c.k() // 8), dtype=torch.uint8, device="cuda")
    cw = torch.empty((batch, c.n() // 8), dtype=torch.uint8, device="cuda")
    c.copy_encode_batch(data, cw); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): c.copy_encode_batch(data, cw)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5e3
    print("%s enc %.2f Mcw/s  %.1f Gbit/s info  %.1f GB/s (in+out)" % (c.name, batch / t / 1e6, batch * c.k() / t / 1e9, batch * (c.k() + c.n()) / 8 / t / 1e9))
