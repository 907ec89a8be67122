"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE in the build container.

    python tests/golden/make_golden.py

Every random draw the reference takes from torch's global generator is injected (oracle/ref_adapter.py)
and recorded in the fixture, so the fixtures pin the arithmetic, not torch's RNG streams.
The fixtures are small and committed; this script is the record of how they were made.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_adapter  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
AS = ref_adapter.load()
torch.set_num_threads(1)


def inputs():
    rng = np.random.default_rng(20251018)
    dy = np.array([0.5, -1.25, 2, 0, -0.75, 3, -0.25, 0.25], np.float32)       # SURVEY 8(c) dyadic vector
    cases = {
        "dyadic8": dy,
        "dyadic64": (rng.integers(-64, 65, 64) / 16.0).astype(np.float32),
        "gauss1000": rng.standard_normal(1000).astype(np.float32),
        "gauss4096": rng.standard_normal(4096).astype(np.float32),
        "gauss9001": rng.standard_normal(9001).astype(np.float32),
        "lognorm4096": (rng.lognormal(1.0, 2.0, 4096) * rng.choice([-1, 1], 4096)).astype(np.float32),
        "bern5000": (rng.random(5000) < 0.7).astype(np.float32),
        "sparse3000": (rng.standard_normal(3000) * (rng.random(3000) < 0.05)).astype(np.float32),
        "onehot17": np.eye(1, 17, 5, dtype=np.float32)[0] * -3.5,
        "zeros33": np.zeros(33, np.float32),
    }
    return cases


def gen_type():
    recs = {}
    i = 0
    for name, x in inputs().items():
        for R in (1, 2, 4, 6) if x.size <= 64 else (0.5, 1, 2, 3, 8):
            L1 = np.float32(torch.tensor(x).abs().sum().item())                 # the reference's own fp32 L1
            d = x.size
            for X in (0.0, 0.3, 0.5, 0.99, 0.123456):
                with ref_adapter.inject(rand=X):
                    q = AS.Type_unbiased_quantize(x, R).numpy()
                recs[f"u{i}_x"] = x; recs[f"u{i}_R"] = np.float64(R); recs[f"u{i}_X"] = np.float32(X)
                recs[f"u{i}_L1"] = L1; recs[f"u{i}_q"] = q; recs[f"u{i}_name"] = np.array(name)
                i += 1
    recs["n_unbiased"] = np.int64(i)
    j = 0
    for name, x in inputs().items():
        for R in (1, 2, 4, 6) if x.size <= 64 else (0.5, 1, 2, 3, 8):
            L1 = np.float32(torch.tensor(x).abs().sum().item())
            try:
                q = AS.Type_biased_quantize(x, R).numpy()
            except RuntimeError:        # torch.topk(k > d): the reference itself fails when |Delta| > d
                continue
            recs[f"b{j}_x"] = x; recs[f"b{j}_R"] = np.float64(R); recs[f"b{j}_L1"] = L1
            recs[f"b{j}_q"] = q; recs[f"b{j}_name"] = np.array(name)
            j += 1
    recs["n_biased"] = np.int64(j)
    np.savez_compressed(os.path.join(OUT, "type_quantizers.npz"), **recs)
    print("type_quantizers:", i, "unbiased,", j, "biased")


def gen_hadamard():
    rng = np.random.default_rng(7)
    H = AS.Hadamard(device="cpu")
    S = AS.HadamardSender(device="cpu")
    Rcv = AS.HadamardReceiver(device="cpu")
    recs = {}
    vecs = [np.arange(1, 9, dtype=np.float32)] + [rng.standard_normal(n).astype(np.float32) for n in (2, 64, 1024, 4096, 32768)]
    for i, v in enumerate(vecs):
        recs[f"h{i}_x"] = v
        recs[f"h{i}_y"] = H.hadamard(torch.tensor(v).clone()).numpy()
        recs[f"p{i}_y"] = AS.fast_walsh_hadamard_transform(torch.tensor(v).clone()).numpy()
    recs["n_h"] = np.int64(len(vecs))
    # randomized transform: non power-of-two lengths get padded (AS:131-139)
    k = 0
    for d, seed in ((1000, 123), (4096, 5), (5000, 77), (1, 3), (3, 9)):
        x = rng.standard_normal(d).astype(np.float32)
        dpad = 1 << int(np.ceil(np.log2(d))) if d > 1 else 1
        diag = H.random_diagonal(dpad, seed).numpy()
        y = S.randomized_hadamard_transform(torch.tensor(x), seed).numpy()
        z = Rcv.randomized_inverse_hadamard_transform(torch.tensor(y).clone(), seed).numpy()
        recs[f"r{k}_x"] = x; recs[f"r{k}_diag"] = diag; recs[f"r{k}_y"] = y; recs[f"r{k}_z"] = z
        k += 1
    recs["n_r"] = np.int64(k)
    np.savez_compressed(os.path.join(OUT, "hadamard.npz"), **recs)
    print("hadamard:", len(vecs), "plain,", k, "randomized")


def gen_drive():
    rng = np.random.default_rng(11)
    recs = {}
    k = 0
    for d in (2048, 5000, 100, 4097):
        x = rng.standard_normal(d).astype(np.float32)
        us, s0 = [], 0
        while s0 < d:
            ln = min(2048, d - s0)
            p2 = 1
            while p2 < ln:
                p2 <<= 1
            us.append(rng.random(p2).astype(np.float32))
            s0 += 2048
        with ref_adapter.inject(rand_like=list(us)):
            q = AS.DRIVE_quantize_Hadamard(x, 1).numpy()
        dsign = np.concatenate([(u > 0.5).astype(np.float32) * 2 - 1 for u in us])
        recs[f"d{k}_x"] = x; recs[f"d{k}_dsign"] = dsign; recs[f"d{k}_q"] = q
        k += 1
    recs["n"] = np.int64(k)
    np.savez_compressed(os.path.join(OUT, "drive.npz"), **recs)
    print("drive:", k)


def gen_eden():
    rng = np.random.default_rng(13)
    H = AS.Hadamard(device="cpu")
    recs = {}
    k = 0
    for d in (1000, 4096, 777):
        for nbits in (1, 2):
            for seed in (17, 64):
                x = rng.standard_normal(d).astype(np.float32)
                dpad = 1 << int(np.ceil(np.log2(d)))
                diag = H.random_diagonal(dpad, seed).numpy()
                snd = AS.EdenSender(device="cpu")
                comp = snd.compress({"vec": torch.tensor(x).clone(), "seed": seed, "nbits": nbits,
                                     "rotation_seed": 123, "nlevels": 2 ** nbits})
                with ref_adapter.inject(randint=seed):
                    q = AS.EDEN_quantize_Hadamard(x, nbits)
                rot = snd.randomized_hadamard_transform(torch.tensor(x).clone(), seed)
                recs[f"e{k}_x"] = x; recs[f"e{k}_diag"] = diag; recs[f"e{k}_nbits"] = np.int64(nbits)
                recs[f"e{k}_bins"] = comp["bins"].numpy().astype(np.int32)
                recs[f"e{k}_scale"] = np.float32(comp["scale"].item())
                recs[f"e{k}_norm"] = np.float32(torch.norm(rot, 2).item())
                recs[f"e{k}_q"] = np.asarray(q, np.float32)
                k += 1
    recs["n"] = np.int64(k)
    np.savez_compressed(os.path.join(OUT, "eden.npz"), **recs)
    print("eden:", k)


def gen_quicfl():
    """Receiver only (the sender tables are not shipped): the sender dict is synthesised."""
    rng = np.random.default_rng(17)
    H = AS.Hadamard(device="cpu")
    rcv = AS.QuicFLReceiver(device="cpu")
    recs = {}
    k = 0
    hl = {1: 64, 2: 32, 3: 16, 4: 16}
    for nbits in (1, 2, 3, 4):
        tab = rcv.recv_table[nbits].numpy()
        recs[f"table{nbits}"] = tab.astype(np.float32)
        assert tab.size == (2 ** nbits) * hl[nbits], tab.shape
        for d in (1000, 2048):
            dpad = 1 << int(np.ceil(np.log2(d)))
            X = rng.integers(0, 2 ** nbits, dpad)
            exact = rng.random(dpad) < 0.004
            ev = (rng.standard_normal(int(exact.sum())) * 3).astype(np.float32)
            prng_seed = int(rng.integers(0, 2 ** 16))
            gen = torch.Generator(device="cpu"); gen.manual_seed(prng_seed)
            h = torch.randint(0, hl[nbits], (dpad,), generator=gen).numpy()
            scale = torch.tensor(np.float32(np.sqrt(dpad) / (1.0 + rng.random())))
            diag = H.random_diagonal(dpad, 123).numpy()
            out = rcv.decompress({"X": torch.tensor(X), "exact_values": torch.tensor(ev), "exact_indeces": torch.tensor(exact),
                                  "prng_seed": prng_seed, "rotation_seed": 123, "dim": d, "scale": scale,
                                  "nbits": nbits, "h_len": hl[nbits]}).numpy()
            recs[f"q{k}_X"] = X.astype(np.int32); recs[f"q{k}_h"] = h.astype(np.int32); recs[f"q{k}_exact"] = exact.astype(np.uint8)
            recs[f"q{k}_ev"] = ev; recs[f"q{k}_scale"] = np.float32(scale.item()); recs[f"q{k}_diag"] = diag
            recs[f"q{k}_nbits"] = np.int64(nbits); recs[f"q{k}_d"] = np.int64(d); recs[f"q{k}_out"] = out
            k += 1
    recs["n"] = np.int64(k)
    np.savez_compressed(os.path.join(OUT, "quicfl_recv.npz"), **recs)
    print("quicfl:", k)


def gen_scalar_and_mean():
    rng = np.random.default_rng(19)
    recs = {}
    k = 0
    for d in (1000, 4096):
        for bits in (1, 2, 4):
            x = rng.standard_normal(d).astype(np.float32)
            u = rng.random(d).astype(np.float32)
            with ref_adapter.inject(rand_like=u):
                q = AS.Scalar_quantize(x, bits).numpy()
            recs[f"s{k}_x"] = x; recs[f"s{k}_u"] = u; recs[f"s{k}_bits"] = np.int64(bits); recs[f"s{k}_q"] = q
            k += 1
    recs["n_s"] = np.int64(k)
    # the reference's server loop ND:133-147 for the unbiased quantizer, n=10, d=1024 (BASELINE config 1)
    n, d, R = 10, 1024, 1
    Xm = rng.standard_normal((n, d)).astype(np.float32)
    Xs = rng.random(n).astype(np.float32)
    est = torch.zeros(d)
    L1s = []
    for c in range(n):
        L1s.append(np.float32(torch.tensor(Xm[c]).abs().sum().item()))
        with ref_adapter.inject(rand=float(Xs[c])):
            est += torch.as_tensor(AS.Type_unbiased_quantize(Xm[c], R)) / n
    recs["mean_X"] = Xm; recs["mean_Xs"] = Xs; recs["mean_L1"] = np.array(L1s, np.float32)
    recs["mean_R"] = np.float64(R); recs["mean_est"] = est.numpy()
    np.savez_compressed(os.path.join(OUT, "scalar_mean.npz"), **recs)
    print("scalar:", k, "+ mean")


def gen_eden_frac_and_kashin():
    """Fractional EDEN rates (AS:352-368, AS:401-421) and Kashin_quantize (AS:834-854) from the unmodified reference.
    Injected / recorded: the Bernoulli masks and uniforms, the dropped coordinates, EDEN's fp32 norm, Kashin's initial M."""
    rng = np.random.default_rng(29)
    H = AS.Hadamard(device="cpu")
    recs = {}
    k = 0
    for d, nbits, seed in ((1000, 1.5, 17), (4096, 1.25, 64), (777, 1.75, 3), (2048, 0.5, 9), (1500, 0.75, 41)):
        x = rng.standard_normal(d).astype(np.float32)
        dpad = 1 << int(np.ceil(np.log2(d)))
        diag = H.random_diagonal(dpad, seed).numpy()
        snd, rcv = AS.EdenSender(device="cpu"), AS.EdenReceiver(device="cpu")
        comp = snd.compress({"vec": torch.tensor(x).clone(), "seed": seed, "nbits": nbits, "rotation_seed": 123, "nlevels": 2 ** nbits})
        rot = snd.randomized_hadamard_transform(torch.tensor(x).clone() if d == dpad else torch.cat([torch.tensor(x), torch.zeros(dpad - d)]), seed)
        recs[f"f{k}_x"] = x; recs[f"f{k}_diag"] = diag; recs[f"f{k}_nbits"] = np.float64(nbits); recs[f"f{k}_seed"] = np.int64(seed)
        recs[f"f{k}_bins"] = comp["bins"].numpy().astype(np.int32)
        recs[f"f{k}_scale"] = np.float32(comp["scale"].item())
        recs[f"f{k}_norm"] = np.float32(torch.norm(rot, 2).item())
        if nbits > 1:
            g = torch.Generator(device="cpu"); g.manual_seed(seed * 7 + 13)                     # AS:389 / AS:402
            mask = torch.bernoulli(torch.ones(dpad) * (nbits - np.floor(nbits)), generator=g).bool().numpy()
            recs[f"f{k}_mask"] = mask.astype(np.uint8)
            recs[f"f{k}_drop"] = np.zeros(dpad, np.uint8)
            q = rcv.decompress(dict(comp))
        else:
            recs[f"f{k}_mask"] = np.zeros(dpad, np.uint8)
            perm = torch.randperm(dpad)
            saved = torch.randperm
            torch.randperm = lambda n_, device=None: perm                                      # AS:416: record the draw
            try:
                q = rcv.decompress(dict(comp))
            finally:
                torch.randperm = saved
            drop = np.zeros(dpad, np.uint8); drop[perm[:round(dpad * (1 - nbits))].numpy()] = 1
            recs[f"f{k}_drop"] = drop
        recs[f"f{k}_q"] = np.asarray(q, np.float32)
        k += 1
    recs["n_frac"] = np.int64(k)
    j = 0
    for d, bits, seed in ((1000, 2, 5), (1024, 1, 77), (900, 4, 12), (3000, 3, 30)):
        x = rng.standard_normal(d).astype(np.float32)
        snd = AS.KashinStochasticQuantizationSender(device="cpu")
        rcv = AS.KashinStochasticQuantizationReceiver(device="cpu")
        pdim = snd.kashin_padded_dim(d, 0.85)
        diag = H.random_diagonal(pdim, 123).numpy()
        U = torch.tensor(rng.random(pdim).astype(np.float32))
        # AS:81: Bernoulli(p) = [u < p] with recorded uniforms -- only inside StochasticQuantizationSender.compress (the rotation
        # diagonal of AS:117-120 is a Bernoulli draw too and must stay the generator's)
        saved, sq_compress, in_sq = torch.bernoulli, AS.StochasticQuantizationSender.compress, [False]

        def _sq(self, data, _orig=sq_compress):
            in_sq[0] = True
            try:
                return _orig(self, data)
            finally:
                in_sq[0] = False
        torch.bernoulli = lambda p, generator=None: (U < p).to(p.dtype) if in_sq[0] else saved(p, generator=generator)
        AS.StochasticQuantizationSender.compress = _sq
        try:
            data = snd.compress({"vec": torch.tensor(x).clone(), "seed": seed, "nbits": bits, "rotation_seed": 123,
                                 "nlevels": 2 ** bits, "niters": 3})
        finally:
            torch.bernoulli = saved
            AS.StochasticQuantizationSender.compress = sq_compress
        coeff, _ = snd.kashin_coefficients({"vec": torch.tensor(x).clone(), "rotation_seed": 123})
        q = rcv.decompress(data)
        recs[f"k{j}_x"] = x; recs[f"k{j}_diag"] = diag; recs[f"k{j}_bits"] = np.int64(bits); recs[f"k{j}_u"] = U.numpy()
        recs[f"k{j}_m0"] = np.float32((torch.norm(torch.tensor(x)) / np.sqrt(1.0 * pdim)).item())
        recs[f"k{j}_coeff"] = coeff.numpy(); recs[f"k{j}_bins"] = data["data"]["bins"].numpy()
        recs[f"k{j}_min"] = np.float32(data["data"]["min"].item()); recs[f"k{j}_step"] = np.float32(data["data"]["step"].item())
        recs[f"k{j}_q"] = np.asarray(q, np.float32)
        j += 1
    recs["n_kashin"] = np.int64(j)
    np.savez_compressed(os.path.join(OUT, "eden_frac_kashin.npz"), **recs)
    print("eden_frac:", k, "kashin:", j)


if __name__ == "__main__":
    if "--extra" not in sys.argv:                      # --extra: only the fixtures added in round 2
        gen_type(); gen_hadamard(); gen_drive(); gen_eden(); gen_quicfl(); gen_scalar_and_mean()
    gen_eden_frac_and_kashin()
