#!/usr/bin/env python3
"""GPU version of the reference's BER Monte-Carlo driver (reference perftest/src/main.rs:9-70).

The reference runs, per SNR point, one endless `ms_trial` loop per CPU core: random data -> encode ->
hard_to_llrs (+-1) -> add Gaussian noise -> decode_ms::<f32>(..., 100) -> count data-bit errors, until
more than 50 M information bits or 5000 bit errors have been seen (main.rs:46-55), then prints
`code,snr,trials,bits,errors,BER` (main.rs:62).  Here every stage runs on the GPU in batches and
is a kernel of this library: counter-based random data and AWGN (csrc/channel.cu), encoder, decoder, error count.

    python perftest.py [--code TC512] [--snrs 0.8,0.9,...] [--llr f32] [--batch 65536] [--channel reference|ebn0]

`--channel reference` reproduces the reference's noise line exactly: std-dev = 1 / 10^(snr/10) added to the
+-1 LLRs (main.rs:15 -- note it has no code-rate term).  `--channel ebn0` uses the true Eb/N0 model of
SURVEY.md section 8d (sigma^2 = 1 / (2 (k/n) 10^(EbN0/10)), LLR = 2y/sigma^2).
The same stop rule and CSV line as the reference are used; `fer` and mean iterations are appended.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="TC512")
    ap.add_argument("--snrs", default="0.8,0.9,1.0,1.1,1.2,1.3,1.4,1.5,1.6,1.7,1.8,1.9,2.0,2.1,2.2")
    ap.add_argument("--llr", default="f32", choices=["i8", "i16", "f32", "f64"])
    ap.add_argument("--batch", type=int, default=1 << 16)
    ap.add_argument("--max-iters", type=int, default=100)
    ap.add_argument("--channel", default="reference", choices=["reference", "ebn0"])
    ap.add_argument("--max-bits", type=float, default=50e6)
    ap.add_argument("--max-errors", type=int, default=5000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()

    import torch
    import labrador_ldpc_b200 as L
    if not torch.cuda.is_available():
        raise SystemExit("perftest.py needs a CUDA device")
    c = L.LDPCCode[args.code]
    n, k = c.n(), c.k()
    kb = k // 8
    data = torch.empty((args.batch, kb), dtype=torch.uint8, device="cuda")
    cw = torch.empty((args.batch, n // 8), dtype=torch.uint8, device="cuda")
    t_start = time.time()
    for snr in [float(s) for s in args.snrs.split(",")]:
        trials = errors = frame_errors = 0
        iters_sum = 0
        if args.channel == "reference":                    # main.rs:15: +-1 plus N(0, (10^(-snr/10))^2), no rate term
            sigma, llr_scale = 1.0 / 10.0 ** (snr / 10.0), 1.0
            scale_i8, scale_i16 = 16.0, 2048.0
        else:                                              # true Eb/N0, LLR = 2 y / sigma^2
            sigma2 = 1.0 / (2.0 * (k / n) * 10.0 ** (snr / 10.0))
            sigma, llr_scale = sigma2 ** 0.5, 2.0 / sigma2
            scale_i8, scale_i16 = 4.0, 256.0
        while trials * k <= args.max_bits and errors <= args.max_errors:
            # every stage below is a kernel of this library; frames are numbered across the whole run
            c.random_data_batch(args.seed, trials, data)                                   # main.rs:10
            c.copy_encode_batch(data, cw)                                                  # main.rs:11-12
            if args.llr == "i8":                                                           # main.rs:13-18
                q = c.awgn_batch(cw, sigma, llr_scale * scale_i8, args.seed, trials, "i8", limit=31)
            elif args.llr == "i16":
                q = c.awgn_batch(cw, sigma, llr_scale * scale_i16, args.seed, trials, "i16", limit=8191)
            else:
                q = c.awgn_batch(cw, sigma, llr_scale, args.seed, trials, "f32")
                if args.llr == "f64":
                    q = q.double()
            out, ok, iters = c.decode_ms_batch(q, args.max_iters)                          # main.rs:19-22
            errs = c.count_errors_batch(out, data)                                         # main.rs:23-28
            errors += int(errs.sum().item())
            frame_errors += int((errs != 0).sum().item())
            iters_sum += int(iters.sum().item())
            trials += args.batch
        bits_total = trials * k
        n_err = max(1, errors)                                               # main.rs:60
        ber = n_err / bits_total
        print("%s,%.2f,%d,%d,%d,%.5e,fer=%.3e,mean_iters=%.2f" % (
            args.code, snr, trials, bits_total, n_err, ber, frame_errors / trials, iters_sum / trials), flush=True)
    print("# %s %s LLRs, channel=%s, kernel %s, %.1f s" % (
        args.code, args.llr, args.channel, c.decode_ms_kernel_name(args.llr), time.time() - t_start), file=sys.stderr)


if __name__ == "__main__":
    main()
