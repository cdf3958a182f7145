"""Quick device-side timing of decode_ms for one (code, type) -- development aid, not the bench."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import labrador_ldpc_b200 as L

def gen(code, batch, ebn0, ty="i8", seed=1):
    c = L.LDPCCode(code)
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    data = torch.randint(0, 256, (batch, c.k() // 8), dtype=torch.uint8, device="cuda", generator=g)
    cw = c.copy_encode_batch(data)
    bits = ((cw.unsqueeze(-1) >> torch.arange(7, -1, -1, device="cuda", dtype=torch.uint8)) & 1).reshape(batch, -1)
    sigma2 = 1.0 / (2.0 * (c.k() / c.n()) * 10.0 ** (ebn0 / 10.0))
    y = (1.0 - 2.0 * bits.float()) + (sigma2 ** 0.5) * torch.randn(bits.shape, device="cuda", generator=g)
    llr = 2.0 * y / sigma2
    if ty == "i8":
        return data, torch.clamp(torch.round(4.0 * llr), -31, 31).to(torch.int8).contiguous()
    if ty == "i16":
        return data, torch.clamp(torch.round(256.0 * llr), -8191, 8191).to(torch.int16).contiguous()
    if ty == "i32":
        return data, torch.round(65536.0 * llr).to(torch.int32).contiguous()
    if ty == "f32":
        return data, llr.contiguous()
    return data, llr.double().contiguous()

if __name__ == "__main__":
    code = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ty = sys.argv[2] if len(sys.argv) > 2 else "i8"
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
    ebn0 = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
    c = L.LDPCCode(code)
    data, llrs = gen(code, batch, ebn0, ty)
    out, ok, it = c.decode_ms_batch(llrs, 100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        c.decode_ms_batch(llrs, 100, output=out, success=ok, iters=it)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    itf = it.float()
    meta = os.environ.get("QUICK_TIME_META")
    if meta:      # facts about the (identical) launches of this run, for tools/ncu_export.py
        import json
        edges = {0: 512, 1: 1024, 2: 2048, 3: 4992, 4: 5888, 5: 7680, 6: 19968, 7: 23552, 8: 30720}[code]
        upd = torch.where(ok.bool(), it.double() + 1.0, it.double()).sum().item() * 2.0 * edges
        json.dump({"code": c.name, "llr_type": ty, "frames": batch, "ebn0_db": ebn0, "max_iters": 100,
                   "mean_iters": itf.mean().item(), "success": ok.float().mean().item(), "edge_updates": upd,
                   "variant": os.environ.get("LABRADOR_LDPC_TM_ARITH", "default"),
                   "command": "tools/quick_time.py " + " ".join(sys.argv[1:])}, open(meta, "w"))
    print("%s %s batch %d ebn0 %.1f kernel %s: %.3f ms  %.0f cw/s  %.2f Gbit/s info  succ %.4f iters mean %.2f max %d" % (
        c.name, ty, batch, ebn0, c.decode_ms_kernel_name(ty), ms, batch / ms * 1e3, batch * c.k() / ms / 1e6,
        ok.float().mean().item(), itf.mean().item(), int(it.max())))
