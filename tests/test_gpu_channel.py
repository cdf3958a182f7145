"""GPU tests of the harness kernels (csrc/channel.cu): counter-based random data, AWGN channel, error counter.

These replace the per-trial set-up of the reference's Monte-Carlo driver (reference perftest/src/main.rs:9-28).
The Philox stream is checked bit for bit against a numpy restatement (itself pinned by the Random123 known
answers); the Gaussian samples, whose last bits depend on the device's logf / sincospif, are checked
against the same Box-Muller in numpy with a tolerance, and statistically.
"""
import numpy as np
import pytest

from frames import philox_words, quantise_soft

pytestmark = pytest.mark.gpu

SEED = 0x1234_5678_9ABC_DEF0


@pytest.mark.parametrize("code", [0, 3, 5, 8])
def test_random_data_is_the_philox_stream(ldpc, code):
    import torch
    c = ldpc.LDPCCode(code)
    kb = c.k() // 8
    first, batch = (1 << 32) - 3, 50                 # crosses the 32-bit boundary of the frame counter
    got = c.random_data_batch(SEED, first, np.zeros((batch, kb), np.uint8))
    words = philox_words(SEED, np.arange(first, first + batch), (kb + 15) // 16, 1)
    want = words.reshape(batch, -1).astype("<u4").view(np.uint8).reshape(batch, -1)[:, :kb]
    assert np.array_equal(got, want)
    # any split of the run, and device-resident output, give the same frames
    a = c.random_data_batch(SEED, first, np.zeros((17, kb), np.uint8))
    b = c.random_data_batch(SEED, first + 17, torch.zeros((batch - 17, kb), dtype=torch.uint8, device="cuda"))
    torch.cuda.synchronize()
    assert np.array_equal(np.concatenate([a, b.cpu().numpy()]), want)
    assert not np.array_equal(c.random_data_batch(SEED + 1, first, np.zeros((batch, kb), np.uint8)), want)


@pytest.mark.parametrize("code", [0, 5, 8])
def test_awgn_matches_box_muller_on_the_philox_stream(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    n, batch, first, sigma = c.n(), 64, 1000, 0.8
    data = c.random_data_batch(SEED, first, np.zeros((batch, c.k() // 8), np.uint8))
    cw = c.copy_encode_batch(data)
    assert np.array_equal(cw, oracle.copy_encode_batch(code, data))
    y = c.awgn_batch(cw, sigma, 1.0, SEED, first, "f32")
    w = philox_words(SEED, np.arange(first, first + batch), n // 4, 2).astype(np.float64)
    u = ((np.floor(w[..., [0, 2]] / 256) + 1.0) / 16777216.0)
    v = (np.floor(w[..., [1, 3]] / 256) / 16777216.0)
    r = np.sqrt(-2.0 * np.log(u))
    z = np.stack([r[..., 0] * np.cos(2 * np.pi * v[..., 0]), r[..., 0] * np.sin(2 * np.pi * v[..., 0]),
                  r[..., 1] * np.cos(2 * np.pi * v[..., 1]), r[..., 1] * np.sin(2 * np.pi * v[..., 1])], axis=-1)
    x = 1.0 - 2.0 * np.unpackbits(cw, axis=1).astype(np.float64)
    want = x + sigma * z.reshape(batch, n)
    assert np.allclose(y, want, rtol=0, atol=2e-5), np.abs(y - want).max()
    # quantised outputs are the quantisation of the f32 output (same product y * scale)
    soft = c.awgn_batch(cw, sigma, 12.5, SEED, first, "f32")
    assert np.array_equal(c.awgn_batch(cw, sigma, 12.5, SEED, first, "i8", limit=31), quantise_soft(soft, 1.0, 31, np.int8))
    assert np.array_equal(c.awgn_batch(cw, sigma, 800.0, SEED, first, "i16", limit=8191),
                          quantise_soft(c.awgn_batch(cw, sigma, 800.0, SEED, first, "f32"), 1.0, 8191, np.int16))
    # shard invariance
    part = c.awgn_batch(cw[20:], sigma, 1.0, SEED, first + 20, "f32")
    assert np.array_equal(part, y[20:])


def test_awgn_statistics(ldpc):
    import torch
    c = ldpc.LDPCCode(8)
    batch, sigma = 512, 1.25
    cw = torch.zeros((batch, c.n() // 8), dtype=torch.uint8, device="cuda")      # all-zero codeword: x = +1
    z = ((c.awgn_batch(cw, sigma, 1.0, 99, 0, "f32") - 1.0) / sigma).double().flatten()
    n = z.numel()                                                                 # 4.2 M samples
    assert abs(z.mean().item()) < 5.0 / n ** 0.5
    assert abs(z.var().item() - 1.0) < 5.0 * (2.0 / n) ** 0.5
    assert abs((z ** 3).mean().item()) < 5.0 * (15.0 / n) ** 0.5
    assert abs((z ** 4).mean().item() - 3.0) < 5.0 * (96.0 / n) ** 0.5
    assert abs((z.abs() > 3.0).double().mean().item() - 0.0026998) < 5.0 * (0.0027 / n) ** 0.5
    # neighbouring samples (the two outputs of one Box-Muller pair, and consecutive pairs) are uncorrelated
    assert abs((z[:-1] * z[1:]).mean().item()) < 5.0 / n ** 0.5
    assert abs((z[:-2] * z[2:]).mean().item()) < 5.0 / n ** 0.5


@pytest.mark.parametrize("code", [0, 4, 8])
def test_count_errors(ldpc, code):
    import torch
    c = ldpc.LDPCCode(code)
    rng = np.random.default_rng(code)
    batch = 300
    data = rng.integers(0, 256, (batch, c.k() // 8), dtype=np.uint8)
    decoded = rng.integers(0, 256, (batch, c.output_len()), dtype=np.uint8)
    decoded[::3, : c.k() // 8] = data[::3]
    decoded[5, 0] ^= 0x81
    want = np.unpackbits(decoded[:, : c.k() // 8] ^ data, axis=1).sum(axis=1).astype(np.uint32)
    assert np.array_equal(c.count_errors_batch(decoded, data), want)
    got = c.count_errors_batch(torch.from_numpy(decoded).cuda(), torch.from_numpy(data).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy().astype(np.uint32), want)


def test_monte_carlo_pipeline_end_to_end(ldpc, oracle):
    """generate -> encode -> channel (quantised) -> decode -> count, all on the device, against the oracle
    decoding the same LLR bytes."""
    import torch
    code = 5
    c = ldpc.LDPCCode(code)
    batch, ebn0 = 256, 1.8
    sigma = (1.0 / (2.0 * (c.k() / c.n()) * 10.0 ** (ebn0 / 10.0))) ** 0.5
    data = c.random_data_batch(7, 0, torch.zeros((batch, c.k() // 8), dtype=torch.uint8, device="cuda"))
    cw = c.copy_encode_batch(data)
    llrs = c.awgn_batch(cw, sigma, 4.0 * 2.0 / sigma ** 2, 7, 0, "i8", limit=31)
    out, ok, it = c.decode_ms_batch(llrs, 100)
    errs = c.count_errors_batch(out, data)
    torch.cuda.synchronize()
    want = oracle.decode_ms_batch(code, llrs.cpu().numpy(), 100, nthreads=8)
    assert np.array_equal(out.cpu().numpy(), want[0]) and np.array_equal(ok.cpu().numpy().astype(bool), want[1].astype(bool))
    werr = np.unpackbits(want[0][:, : c.k() // 8] ^ data.cpu().numpy(), axis=1).sum(axis=1)
    assert np.array_equal(errs.cpu().numpy().astype(np.int64), werr.astype(np.int64))
    assert 0 < ok.float().mean().item() <= 1.0
