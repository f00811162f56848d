"""Seeded synthetic 16 kHz PCM (SURVEY.md §8d "value distributions").

There is no dataset in this environment (the reference's Hey-Snips wavs and
data/*.h5 are absent), so parity tests and the benchmark run on synthetic streams.
Each stream is a pure function of (seed, class, stream index); the host generator
below is numpy, `device_pcm` makes the same *kind* of signal on the GPU with torch
for benchmark-sized inputs (not bit-identical to the host generator — parity tests
always upload host-generated PCM).

Classes: 0 white noise s=0.1 FS | 1 white noise s=0.01 | 2 speech-like harmonic
stack with syllabic AM | 3 silence + clicks | 4 full-scale clipping sine |
5 pink-ish noise (octave-spaced random tones + white floor).
"""
from __future__ import annotations

import numpy as np

N_CLASSES = 6
SR = 16000


def _rng(seed: int, cls: int, stream: int) -> np.random.Generator:
    return np.random.default_rng([int(seed), int(cls), int(stream)])


def stream_float(n: int, cls: int, seed: int = 0, stream: int = 0) -> np.ndarray:
    """One float64 stream in roughly [-1.2, 1.2] (class 4 exceeds full scale on purpose)."""
    r = _rng(seed, cls, stream)
    t = np.arange(n) / SR
    if cls == 0:
        return 0.1 * r.standard_normal(n)
    if cls == 1:
        return 0.01 * r.standard_normal(n)
    if cls == 2:
        f0 = r.uniform(90, 250)
        vib = 1.0 + 0.03 * np.sin(2 * np.pi * r.uniform(4, 7) * t + r.uniform(0, 6.28))
        phase = 2 * np.pi * np.cumsum(f0 * vib) / SR
        formants = r.uniform([300, 900, 2200], [900, 2200, 3500])
        sweep = 1.0 + 0.3 * np.sin(2 * np.pi * r.uniform(0.5, 2.0) * t)[:, None]
        x = np.zeros(n)
        for h in range(1, 30):
            fh = f0 * h
            if fh > 7600:
                break
            env = sum(np.exp(-0.5 * ((fh - fm * sweep[:, 0]) / 180.0) ** 2) for fm in formants)
            x += (env + 0.02) / h ** 0.5 * np.sin(h * phase + r.uniform(0, 6.28))
        am = 0.55 + 0.45 * np.sin(2 * np.pi * r.uniform(2.5, 5.0) * t + r.uniform(0, 6.28))
        x = 0.25 * x * am / max(1e-9, np.abs(x).max()) * 2.0
        return x + 0.003 * r.standard_normal(n)
    if cls == 3:
        x = np.zeros(n)
        k = max(1, n // 4000)
        pos = r.integers(0, n, size=k)
        x[pos] = r.uniform(-0.9, 0.9, size=k)
        return x
    if cls == 4:
        return 1.2 * np.sin(2 * np.pi * r.uniform(200, 3000) * t + r.uniform(0, 6.28))
    if cls == 5:
        x = 0.002 * r.standard_normal(n)
        for o in range(8):
            lo = 40.0 * 2 ** o
            for _ in range(6):
                f = r.uniform(lo, min(2 * lo, 7900))
                x += 0.04 / 2 ** (o * 0.5) * np.sin(2 * np.pi * f * t + r.uniform(0, 6.28))
        return x
    raise ValueError("unknown synthetic class %d" % cls)


def stream_int16(n: int, cls: int, seed: int = 0, stream: int = 0) -> np.ndarray:
    x = stream_float(n, cls, seed, stream)
    return np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16)


def batch_int16(n_streams: int, n: int, seed: int = 0, first_stream: int = 0) -> np.ndarray:
    """[n_streams, n] int16; stream s is class (s % N_CLASSES).  `first_stream` lets a
    rank generate exactly its shard of a global batch."""
    out = np.empty((n_streams, n), np.int16)
    for i in range(n_streams):
        s = first_stream + i
        out[i] = stream_int16(n, s % N_CLASSES, seed, s)
    return out


def device_pcm(n_streams: int, n: int, seed: int, device, first_stream: int = 0):
    """Benchmark-sized int16 PCM generated on the GPU: per-stream gain-modulated
    gaussian noise plus a per-stream tone (values span the int16 range, a share of
    streams clips).  Rank-shardable through `first_stream`."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed) * 1000003 + int(first_stream))
    out = torch.empty((n_streams, n), dtype=torch.int16, device=device)
    step = max(1, min(n_streams, (1 << 26) // max(1, n)))
    t = torch.arange(n, device=device, dtype=torch.float32) / SR
    for s0 in range(0, n_streams, step):
        m = min(step, n_streams - s0)
        gain = torch.rand((m, 1), generator=g, device=device) * 0.3 + 0.005
        f = torch.rand((m, 1), generator=g, device=device) * 3000.0 + 100.0
        am = 0.6 + 0.4 * torch.sin(2 * torch.pi * 3.0 * t)[None, :]
        x = torch.randn((m, n), generator=g, device=device) * gain * am
        x += 0.2 * torch.sin(2 * torch.pi * f * t[None, :])
        out[s0:s0 + m] = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)
    return out
