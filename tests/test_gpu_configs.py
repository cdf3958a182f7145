"""The BASELINE.json configurations at their full sizes.

C1 is small enough for the CPU oracle to decode every frame: full exact parity.  C2 and C4 are checked through
size-independent properties of the domain (decoded data == transmitted data on success, re-encoding the decoded
data gives the decoded codeword, iteration counts in range, the two decoders agree with each other on what a
codeword is) plus exact oracle parity on a prefix of the very same LLR bytes.  C3 is bench.py; C5 is
tests/test_gpu_parity.py::test_encode_kat_and_random + tools/sweep.py.
Frames come from the library's own counter-based generator (csrc/channel.cu, tests/test_gpu_channel.py).
"""
import numpy as np
import pytest

from test_gpu_parity import assert_exact, assert_float_parity

pytestmark = pytest.mark.gpu


def device_frames(torch, c, batch, ebn0, seed, ty, scale, limit):
    sigma2 = 1.0 / (2.0 * (c.k() / c.n()) * 10.0 ** (ebn0 / 10.0))
    data = c.random_data_batch(seed, 0, torch.empty((batch, c.k() // 8), dtype=torch.uint8, device="cuda"))
    cw = c.copy_encode_batch(data)
    llrs = c.awgn_batch(cw, sigma2 ** 0.5, scale * 2.0 / sigma2, seed, 0, ty, limit=limit)
    return data, cw, llrs


def check_properties(torch, c, data, cw, out, ok, iters, maxiters, min_success):
    okb = ok.bool()
    assert okb.float().mean().item() >= min_success
    kb, nb = c.k() // 8, c.n() // 8
    # a frame the decoder calls a success is a codeword: re-encoding its data part reproduces it
    reenc = c.copy_encode_batch(out[:, :kb].contiguous())
    assert torch.equal(reenc[okb], out[okb][:, :nb])
    # undetected errors (a different codeword) are possible but must be rare at these SNRs
    wrong = (out[okb][:, :kb] != data[okb]).any(dim=1).float().mean().item()
    assert wrong < 1e-3
    assert int(iters[okb].max()) < maxiters and bool((iters[~okb] == maxiters).all())


def test_c1_tc128_i8_100k_frames_full_parity(ldpc, oracle):
    """configs[0]: TC128 i8, 100 000 frames at each of Eb/N0 = 0..4 dB; the oracle decodes every frame."""
    import torch
    c = ldpc.LDPCCode.TC128
    fer = []
    for ebn0 in (0.0, 1.0, 2.0, 3.0, 4.0):
        _, _, llrs = device_frames(torch, c, 100_000, ebn0, 100 + int(ebn0), "i8", 4.0, 31)
        got = c.decode_ms_batch(llrs, 100)
        torch.cuda.synchronize()
        want = oracle.decode_ms_batch(0, llrs.cpu().numpy(), 100, nthreads=16)
        assert_exact([g.cpu().numpy() for g in got], want, "TC128 i8 %.0f dB" % ebn0)
        fer.append(1.0 - float(want[1].mean()))
    assert fer == sorted(fer, reverse=True) and fer[0] > 0.1 and fer[-1] < 0.02, fer


@pytest.mark.parametrize("ty,scale,limit", [("i16", 256.0, 8191), ("f32", 1.0, 0)])
def test_c2_tm2048_1m_frames_min_sum(ldpc, oracle, ty, scale, limit):
    """configs[1]: TM2048, 1 Mi frames (4 GiB of i16 / 8 GiB of f32 LLRs), punctured column exercised."""
    import torch
    c = ldpc.LDPCCode.TM2048
    batch = 1 << 20
    data, cw, llrs = device_frames(torch, c, batch, 2.0, 21, ty, scale, limit)
    out, ok, iters = c.decode_ms_batch(llrs, 100)
    torch.cuda.synchronize()
    check_properties(torch, c, data, cw, out, ok, iters, 100, 0.999)
    # the punctured 512 bits are recovered too: whole n+p output == systematic codeword + re-derived punctured parity
    ns = 16384                                   # prefix of the very same LLR bytes the GPU decoded
    sample = llrs[:ns].cpu().numpy()
    want = oracle.decode_ms_batch(5, sample, 100, nthreads=16)
    got = (out[:ns].cpu().numpy(), ok[:ns].cpu().numpy(), iters[:ns].cpu().numpy())
    if ty == "i16":
        assert_exact(got, want, "TM2048 i16 prefix")
    else:
        assert_float_parity(got, want, "TM2048 f32 prefix")
    assert np.array_equal(want[0][want[1].astype(bool)][:, : c.n() // 8], cw[:ns].cpu().numpy()[want[1].astype(bool)])
    if ty == "f32":
        # north_star: FER of the float decode within statistical noise of the oracle's on the same frames
        fer_g, fer_w = 1.0 - got[1].astype(bool).mean(), 1.0 - want[1].astype(bool).mean()
        assert abs(fer_g - fer_w) <= 3.0 * np.sqrt(max(fer_w, 1.0 / ns) / ns), (fer_g, fer_w)


def test_c2_tm2048_1m_frames_bit_flipping(ldpc, oracle):
    """configs[1], decode_bf leg: 1 Mi hard-decision frames with 0..6 bit errors each."""
    import torch
    c = ldpc.LDPCCode.TM2048
    batch = 1 << 20
    data = c.random_data_batch(31, 0, torch.empty((batch, c.k() // 8), dtype=torch.uint8, device="cuda"))
    cw = c.copy_encode_batch(data)
    rx = cw.clone()
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    rows = torch.arange(batch, device="cuda")
    nflip = torch.randint(0, 7, (batch,), device="cuda", generator=g)
    for j in range(6):
        pos = torch.randint(0, c.n(), (batch,), device="cuda", generator=g)
        mask = torch.where(nflip > j, (128 >> (pos % 8)), torch.zeros_like(pos)).to(torch.uint8)
        rx[rows, pos // 8] ^= mask
    out, ok, iters = c.decode_bf_batch(rx, 50)
    torch.cuda.synchronize()
    check_properties(torch, c, data, cw, out, ok, iters, 50, 0.95)
    clean = nflip == 0
    assert bool(ok[clean].all()) and torch.equal(out[clean][:, : c.n() // 8], cw[clean])
    want = oracle.decode_bf_batch(5, rx[:65536].cpu().numpy(), 50, nthreads=16)
    assert_exact((out[:65536].cpu().numpy(), ok[:65536].cpu().numpy(), iters[:65536].cpu().numpy()), want, "TM2048 bf prefix")
    # the min-sum decoder fed the same hard decisions (fused hard front end) recovers at least as many frames
    _, ok_ms, _ = c.decode_ms_hard_batch(rx[:65536].contiguous(), 50)
    assert int(ok_ms.sum()) >= int(ok[:65536].sum()) - 8


def test_c4_mixed_high_rate_batch(ldpc, oracle):
    """configs[3]: TM5120 (4 dB) + TM6144 (3 dB), i8, one mixed batch on concurrent streams."""
    import torch
    a, b = ldpc.LDPCCode.TM5120, ldpc.LDPCCode.TM6144
    half = 1 << 17
    da, cwa, la = device_frames(torch, a, half, 4.0, 41, "i8", 4.0, 31)
    db, cwb, lb = device_frames(torch, b, half, 3.0, 42, "i8", 4.0, 31)
    (oa, ka, ia), (ob, kb_, ib) = ldpc.decode_ms_mixed([(a, la), (b, lb)], 100)
    torch.cuda.synchronize()
    check_properties(torch, a, da, cwa, oa, ka, ia, 100, 0.999)
    check_properties(torch, b, db, cwb, ob, kb_, ib, 100, 0.999)
    for code, c, l, o, k, i in ((6, a, la, oa, ka, ia), (7, b, lb, ob, kb_, ib)):
        ns = 16384
        want = oracle.decode_ms_batch(code, l[:ns].cpu().numpy(), 100, nthreads=16)
        assert_exact((o[:ns].cpu().numpy(), k[:ns].cpu().numpy(), i[:ns].cpu().numpy()), want, "%s prefix" % c.name)
