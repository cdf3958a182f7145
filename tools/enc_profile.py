"""One large copy_encode_batch per listed code (for ncu captures of the encoder kernels).
    python tools/enc_profile.py 8 6 [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, labrador_ldpc_b200 as L
codes = [int(a) for a in sys.argv[1:] if int(a) < 9] or [8]
frames = [int(a) for a in sys.argv[1:] if int(a) >= 9]
for code in codes:
    c = L.LDPCCode(code)
    batch = frames[0] if frames else max(4096, min(1 << 20, (1 << 31) // c.n()))
    data = torch.randint(0, 256, (batch, c.k() // 8), dtype=torch.uint8, device="cuda")
    cw = torch.empty((batch, c.n() // 8), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        c.copy_encode_batch(data, cw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): c.copy_encode_batch(data, cw)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5e3
    print("%s batch %d: %.2f Mcw/s, %.1f GB/s (in+out)" % (c.name, batch, batch / t / 1e6, batch * (c.k() + c.n()) / 8 / t / 1e9))
