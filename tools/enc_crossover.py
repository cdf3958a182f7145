"""Encoders: time of one copy_encode_batch call vs batch size (run once per LABRADOR_LDPC_ENC_TM_FORM=1 / 2,
LABRADOR_LDPC_ENC_TC_COPIES=0 / 1, LABRADOR_LDPC_ENC_GENERATOR=1).   python tools/enc_crossover.py [codes...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, labrador_ldpc_b200 as L
for code in ([int(a) for a in sys.argv[1:]] or [3, 5, 6, 7, 8]):
    c = L.LDPCCode(code)
    row = []
    for batch in (1, 8, 64, 256, 1024, 2048, 4096, 8192, 16384, 65536, 262144):
        data = torch.randint(0, 256, (batch, c.k() // 8), dtype=torch.uint8, device="cuda")
        cw = torch.empty((batch, c.n() // 8), dtype=torch.uint8, device="cuda")
        for _ in range(3): c.copy_encode_batch(data, cw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): c.copy_encode_batch(data, cw)
        e1.record(); torch.cuda.synchronize()
        row.append("%d: %.1f us" % (batch, e0.elapsed_time(e1) / 20 * 1e3))
    print(c.name, " ".join("%s=%s" % (k[14:], v) for k, v in os.environ.items() if k.startswith("LABRADOR_LDPC_ENC")) or "default", " | ".join(row), flush=True)
