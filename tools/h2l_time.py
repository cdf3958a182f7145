import os, sys
sys.path.insert(0, "/root/repo")
import torch, labrador_ldpc_b200 as L
c = L.LDPCCode(8); batch = 32768
hard = torch.randint(0, 256, (batch, c.n() // 8), dtype=torch.uint8, device="cuda")
for ty, dt in (("f32", torch.float32), ("i8", torch.int8), ("i16", torch.int16)):
    out = torch.empty((batch, c.n()), dtype=dt, device="cuda")
    for _ in range(3): c.hard_to_llrs_batch(hard, ty, llrs=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): c.hard_to_llrs_batch(hard, ty, llrs=out)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20e3
    print(ty, "h2l %.0f GB/s" % ((hard.numel() + out.numel() * out.element_size()) / t / 1e9))
