"""Seeded synthetic IQ used by the golden generator, the parity tests and the
bench.  numpy PCG64 on the host; all outputs complex64."""
import numpy as np


def _cn(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.0)


def cfg1_burst(n_fft=1024, n_spectra=64, seed=1):
    """BASELINE.json configs[0]: one 64k-sample burst: strong off-bin tone,
    weak negative-frequency tone, noise, and a chirp sweeping bins 400->600."""
    rng = np.random.default_rng(seed)
    n = n_fft * n_spectra
    t = np.arange(n, dtype=np.float64)
    x = 0.5 * np.exp(2j * np.pi * 200.25 * t / n_fft)
    x += 0.05 * np.exp(2j * np.pi * (-317.0) * t / n_fft)
    x += 0.01 * _cn(rng, n)
    f0, f1 = 400.0 / n_fft, 600.0 / n_fft
    phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / n)
    x += 0.1 * np.exp(1j * phase)
    return x.astype(np.complex64)


def noise_tones(n_samples, n_fft=1024, seed=2, sigma=0.01, n_tones=8, amp=(0.02, 0.6)):
    """Continuous stream: white noise + n_tones fixed tones at seeded off-bin
    frequencies and amplitudes (BASELINE.json configs[1]/[4] style)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64)
    x = sigma * _cn(rng, n_samples)
    freqs = rng.uniform(-0.5, 0.5, n_tones)
    amps = np.exp(rng.uniform(np.log(amp[0]), np.log(amp[1]), n_tones))
    phs = rng.uniform(0, 2 * np.pi, n_tones)
    for f, a, p in zip(freqs, amps, phs):
        x += a * np.exp(1j * (2 * np.pi * f * t + p))
    return x.astype(np.complex64)


def burst_stress(n_fft=4096, n_spectra=256, seed=3, burst=64):
    """BASELINE.json configs[2]: alternating bursts of a strong tone and
    noise only, to exercise histogram rise and decay."""
    rng = np.random.default_rng(seed)
    n = n_fft * n_spectra
    t = np.arange(n, dtype=np.float64)
    x = 0.01 * _cn(rng, n)
    gate = ((np.arange(n) // (n_fft * burst)) % 2 == 0).astype(np.float64)
    x += 0.4 * gate * np.exp(2j * np.pi * (n_fft // 4 + 0.5) * t / n_fft)
    return x.astype(np.complex64)


def tone(n_fft, n_spectra, k, amp):
    t = np.arange(n_fft * n_spectra, dtype=np.float64)
    return (amp * np.exp(2j * np.pi * k * t / n_fft)).astype(np.complex64)


def impulse(n_fft, n_spectra):
    x = np.zeros((n_spectra, n_fft), np.complex64)
    x[:, 0] = 1.0
    return x.reshape(-1)


def overlap_windows(raw, n_fft, overlap, n_spectra):
    """What the reference's overlap_cc block emits (lib/overlap_cc_impl.cc:64-79):
    n_fft-sample windows hopping n_fft/overlap."""
    hop = n_fft // overlap
    idx = (np.arange(n_spectra)[:, None] * hop + np.arange(n_fft)[None, :])
    return np.ascontiguousarray(raw[idx].reshape(-1))
